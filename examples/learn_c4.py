#!/usr/bin/env python
"""BASELINE config C4: Connect-4 (c4 6x7, N=4), full AZ.Learn loop — self-play sharded by game over the GPUs,
dualnet training with the fused peer-memory gradient all-reduce, arena evaluation, promotion.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 examples/learn_c4.py \
        [--games 4096] [--sims 400] [--iters 2] [--nniters 2] [--arena 256]

Net = dual.DefaultConf(6, 7, 8) with Features=2 (K=16, 6 blocks, FC=32, batch 256), two-plane encoder,
mcts.Config{PUCT 1, M 6, N 7, DontPreferPass, DumbPass}.  Prints one JSON line per epoch from rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from agogo_b200 import _capi as K  # noqa: E402
from agogo_b200 import host  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=400)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--nniters", type=int, default=2)
ap.add_argument("--arena", type=int, default=256)
ap.add_argument("--max-examples", type=int, default=0)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lrank = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))

nn = host.DefaultConf(6, 7, 8)
nn.Features = 2
mc = host.MCTSConfig(PUCT=1.0, M=6, N=7, DumbPass=True, PassPreference=K.DONT_PREFER_PASS, Sims=args.sims)
conf = host.Config(Name="Connect 4", NNConf=nn, MCTSConf=mc, UpdateThreshold=0.52, Encoder=K.ENC_TWO_PLANE,
                   MaxExamples=args.max_examples)
per_gpu = (args.games + world - 1) // world
az = host.AZ(host.Game(K.GAME_C4, 6, 7, 4), conf, n_games=per_gpu, seed=2026, device=lrank, dist=dist)
t_last = [time.time()]


def on_epoch(ep, log):
    now = time.time()
    if rank == 0:
        print(json.dumps(dict(epoch=ep, seconds=round(now - t_last[0], 2), gpus=world, games=args.games, sims=args.sims, **log)),
              flush=True)
    t_last[0] = now


az.Learn(args.iters, args.games, args.nniters, args.arena, on_epoch=on_epoch)
if rank == 0:
    c = az.engine.counters()
    print(json.dumps(dict(done=True, engine_comm=az.engine_comm, sims=c["sims"], evals=c["evals"], launches=c["kernel_launches"])))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
