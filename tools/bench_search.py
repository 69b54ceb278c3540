#!/usr/bin/env python
"""Agent.Search on ONE position (GTP / analysis use): latency and simulations/s of az_search with
mcts.Config workers = 1 (canonical single worker: 800 batch-1 evaluations in sequence) against workers > 1 (rounds of
`workers` descents with virtual-loss flags, one evaluation batch per round).  One JSON line per setting."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from agogo_b200 import _capi as K  # noqa: E402

CONFIGS = {
    "C3": (19, 800, dict(k=256, shared_layers=20, fc=512, batch_size=256, features=18, action_space=362)),
    "C2": (9, 400, dict(k=64, shared_layers=6, fc=128, batch_size=256, features=18, action_space=82)),
}
lib = K.load()
which = sys.argv[1] if len(sys.argv) > 1 else "C3"
size, sims, nn = CONFIGS[which]
for workers in (1, 8, 32, 128):
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=sims, n_games=1, seed=1, nn=nn, max_moves=2 * size * size,
                    workers=workers)
    if len(sys.argv) > 2:
        d.act_scale_log2 = int(sys.argv[2])  # random-init nets have no trained BN statistics: headroom for the fp16 planes
    e = lib.create(d)
    e.net_init(0, 1334)
    e.set_inferer(0, K.INF_DUAL, 0)
    board = np.zeros(size * size, np.int32)
    board[[size * 3 + 3, size * 3 + size - 4]] = (K.BLACK, K.WHITE)
    e.search(0, board, K.BLACK, K.BLACK, move_number=2)  # warm-up: graph capture, weight staging
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        best, visits = e.search(0, board, K.BLACK, K.BLACK, move_number=2)
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"config": which, "workers": workers, "sims": sims, "ms_per_search": dt * 1e3, "sims_per_sec": sims / dt,
                      "best": int(best), "root_children_visited": int((visits > 1).sum()), "visits_sum": float(visits.sum())}), flush=True)
    e.close()
