#!/usr/bin/env python
"""BASELINE config C5 (and C2/C3-shaped variants): 19x19 Go (wq) full AZ.Learn — self-play sharded by game over the GPUs,
dualnet training at DefaultConf's batch 256 with the fused peer-memory gradient all-reduce (K8), arena evaluation, promotion.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/learn_go.py \
        [--size 19] [--games 8192] [--sims 800] [--iters 2] [--nniters 1] [--arena 1024] [--max-moves 3]

Net = the C3/C5 dual net (20 blocks x 256, FC 512, batch 256, 18-plane WQEncoder), mcts.Config{PUCT 1, DontPreferPass,
DumbPass}.  A full C5 epoch is ~3.3e8 simulations per GPU (8192 games x ~400 plies x 800 sims / 8): hours.  --max-moves
caps every game (the engine's documented completion) so that one call exercises every phase of the
loop at the stated widths — 1024 concurrent games per GPU, 800 simulations per move, batch 256 — and reports seconds per
phase; the numbers are a REDUCED C5 and say so.  Epoch 0 plays with the dummy inferer (agogo.go:83-87), so --iters 2 is
the smallest run whose self-play goes through the tensor-core tower.  Prints one JSON line per epoch from rank 0."""
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
ap.add_argument("--size", type=int, default=19)
ap.add_argument("--games", type=int, default=8192)
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--nniters", type=int, default=1)
ap.add_argument("--arena", type=int, default=1024)
ap.add_argument("--max-moves", type=int, default=3)
ap.add_argument("--blocks", type=int, default=20)
ap.add_argument("--k", type=int, default=256)
ap.add_argument("--fc", type=int, default=512)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--max-examples", type=int, default=0)
ap.add_argument("--fast-tower", action="store_true")
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

s = args.size
nn = host.DualConfig(K=args.k, SharedLayers=args.blocks, FC=args.fc, BatchSize=args.batch, Width=s, Height=s, Features=18,
                     ActionSpace=s * s + 1)
mc = host.MCTSConfig(PUCT=1.0, M=s, N=s, DumbPass=True, PassPreference=K.DONT_PREFER_PASS, Sims=args.sims)
conf = host.Config(Name="wq %dx%d" % (s, s), NNConf=nn, MCTSConf=mc, UpdateThreshold=0.55, Encoder=K.ENC_WQ18,
                   MaxExamples=args.max_examples)
per_gpu = (args.games + world - 1) // world
t_start = time.time()
az = host.AZ(host.Game(K.GAME_WQ, s, s, 0, komi=7.5, max_moves=args.max_moves), conf, n_games=per_gpu, seed=2026, device=lrank,
             dist=dist, flags=K.FLAG_FAST_TOWER if args.fast_tower else 0)
t_last = [time.time()]
if rank == 0:
    print(json.dumps(dict(setup_seconds=round(t_last[0] - t_start, 2), gpus=world, games=args.games, games_per_gpu=per_gpu,
                          sims=args.sims, max_moves=args.max_moves, net="%d x %d, FC %d, batch %d" % (args.blocks, args.k, args.fc, args.batch),
                          params=int(az.engine.param_count()[1]), engine_comm=az.engine_comm,
                          reduced="games capped at %d plies (a full game is ~2 x %d plies)" % (args.max_moves, s * s))), flush=True)


def on_epoch(ep, log):
    now = time.time()
    if rank == 0:
        c = az.engine.counters()
        print(json.dumps(dict(epoch=ep, seconds=round(now - t_last[0], 2), rank0_sims=c["sims"], rank0_evals=c["evals"], **log)), flush=True)
    t_last[0] = now


az.Learn(args.iters, args.games, args.nniters, args.arena, on_epoch=on_epoch)
if rank == 0:
    c = az.engine.counters()
    print(json.dumps(dict(done=True, total_seconds=round(time.time() - t_start, 2), engine_comm=az.engine_comm, rank0_sims=c["sims"],
                          rank0_evals=c["evals"], launches=c["kernel_launches"])))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
