#!/usr/bin/env python
"""Regenerates tests/golden/workers_search.npz: the tree after the first Search and the whole move record of small games
searched with mcts.Config workers > 1, produced by the CPU oracle's restatement of the fixed worker schedule
(oracle/mcts.hpp SearchRunWorkers).  The reference cannot produce this vector (its goroutine interleaving is unspecified
and no Go toolchain is available); the fixture pins the schedule itself: oracle and engine must both keep reproducing it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from agogo_b200 import _capi as K  # noqa: E402
from tests import helpers as H  # noqa: E402


def cases():
    rng = np.random.default_rng(123)
    t9 = rng.random((32, 10)).astype(np.float32); t9 /= t9.sum(axis=1, keepdims=True)
    v9 = rng.uniform(0.1, 0.9, 32).astype(np.float32)
    t8 = rng.random((32, 8)).astype(np.float32); t8 /= t8.sum(axis=1, keepdims=True)
    yield "ttt_w4", K.make_desc(K.GAME_MNK, 3, 3, 3, sims=30, nn=H.tiny_nn(3, 3, 10), n_games=2, seed=21, workers=4), t9, v9
    yield "c4_w6", K.make_desc(K.GAME_C4, 6, 7, 4, sims=40, nn=H.tiny_nn(6, 7, 8), n_games=2, seed=22, workers=6), t8, v9


def run(lib, desc, table, values):
    e = lib.create(desc)
    e.set_table(0, table, values); e.set_table(1, table[::-1].copy(), values[::-1].copy())
    e.arena_begin(2, False)
    e.search_begin(); e.search_run(desc.mcts.sims)
    trees = [e.tree_dump(g, t) for g in range(2) for t in (0, 1)]
    e.search_end()
    n = 2
    while n:
        n = e.arena_step()
    e.arena_finish()
    moves = [np.asarray(e.game_record(g)["moves"], np.int32) for g in range(2)]
    return trees, moves


if __name__ == "__main__":
    lib = K.load(os.path.join(ROOT, "oracle", "libazoracle.so"))
    out = {}
    for name, desc, table, values in cases():
        trees, moves = run(lib, desc, table, values)
        for i, t in enumerate(trees):
            out["%s_tree%d" % (name, i)] = t
        for g, m in enumerate(moves):
            out["%s_moves%d" % (name, g)] = m
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "workers_search.npz"), **out)
    print({k: v.shape for k, v in out.items()})
