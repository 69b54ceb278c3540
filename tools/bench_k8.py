#!/usr/bin/env python
"""K8 roofline: the fused gradient all-reduce + SGD kernel over NVLink peer memory, timed alone.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_k8.py [--batch 256] [--collective p2p|nccl]
Net = the C3 dual net (20x256, 19x19); --batch sets the training batch, i.e. the size of the batch-shaped
BN tensors that dominate |theta| (256 -> 2.05 G floats = 8.2 GB)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from agogo_b200 import _capi as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--collective", default="p2p")
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["AZ_TRAIN_COLLECTIVE"] = args.collective
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = K.load()
d = K.make_desc(K.GAME_WQ, 19, 19, 0, komi=7.5, sims=2, n_games=2, seed=1, device=lr, max_moves=4, flags=K.FLAG_FP32_TOWER,
                nn=dict(k=256, shared_layers=20, fc=512, batch_size=args.batch, features=18, action_space=362))
e = lib.create(d)
uid = K.comm_unique_id(lib)
t = torch.tensor(list(uid), dtype=torch.uint8).cuda()
dist.broadcast(t, 0)
e.comm_init(rank, world, bytes(t.cpu().tolist()))
dist.barrier()
ms, nbytes = e.comm_bench(1, args.iters)
out = torch.tensor([ms], dtype=torch.float64).cuda()
dist.all_reduce(out, op=dist.ReduceOp.MAX)
if rank == 0:
    gbs = nbytes / (out.item() / 1e3) / 1e9  # nbytes is already per direction (gradient reads in + parameter writes in)
    print(json.dumps({"kernel": "k_allreduce_sgd_p2p" if args.collective == "p2p" else "ncclAllReduce+k_sgd", "world": world,
                      "params": e.param_count()[1], "ms": out.item(), "nvlink_bytes_per_rank_per_direction": nbytes,
                      "gbs_per_direction": gbs, "peak_gbs_per_direction_measured": 770.0, "frac": gbs / 770.0}))
dist.barrier()
dist.destroy_process_group()
