"""The C++ oracle's search (oracle/mcts.hpp) against a second, independent Python restatement of the same Go sources
(tests/pyref_mcts.py): full tree after every Search — visits, W bits, P bits, child order — and the move sequence, for the
Example's scripted evaluator and for random evaluator tables, on 3x3 and 4x4/5x5 m,n,k games with tree reuse."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from tests import helpers as H
from tests import pyref_mcts as P
from tests.golden import rules_golden as G


def _run_pair(oracle, m, n, k, sims, table, values, seed=7):
    d = K.make_desc(K.GAME_MNK, m, n, k, sims=sims, nn=H.tiny_nn(m, n, m * n + 1), n_games=1, flags=K.FLAG_SHARED_TREE, seed=seed)
    e = oracle.create(d)
    e.set_inferer(0, K.INF_TABLE)
    e.set_table(0, table, values)
    e.arena_begin(1, False)
    g = P.MNK(m, n, k)
    t = P.MCTS(g, 1.0, sims, table, values, m, n)
    player, moves_py, ply, alive = P.BLACK, [], 0, 1
    while alive:
        ended, _ = g.ended()
        assert not ended
        best = t.search(player)
        alive = e.arena_step()
        got = e.tree_dump(0, 0).astype(np.int64) & 0xFFFFFFFF
        want = t.dump() & 0xFFFFFFFF
        assert got.shape == want.shape, (ply, got.shape, want.shape)
        assert (got == want).all(), (ply, np.argwhere(got != want)[:5], got[:4], want[:4])
        g.apply(player, best)
        moves_py.append(best)
        player = P.opponent(player)
        ply += 1
    rec = e.game_record(0)
    e.arena_finish()
    assert list(rec["moves"]) == moves_py
    ended, winner = g.ended()
    assert ended and rec["winner"] == winner
    assert e.counters()["evals"] == t.evals
    return moves_py, winner


@pytest.mark.parametrize("sims", [10, 50, 200])
def test_example_trees_vs_python(oracle, sims):
    rows = np.zeros((10, 10), np.float32)
    vals = np.zeros(10, np.float32)
    for mn, (hot, p, v) in enumerate(G.TTT_DUMMY_NN):
        rows[mn, hot] = p
        vals[mn] = v
    moves, winner = _run_pair(oracle, 3, 3, 3, sims, rows, vals)
    assert moves == G.TTT_EXPECTED_MOVES and winner == G.TTT_EXPECTED_WINNER


@pytest.mark.parametrize("m,n,k,sims,seed", [(3, 3, 3, 40, 1), (3, 3, 3, 150, 2), (4, 4, 3, 60, 3), (5, 5, 4, 30, 4), (4, 5, 4, 45, 5)])
def test_random_tables_vs_python(oracle, m, n, k, sims, seed):
    rng = np.random.default_rng(seed)
    A1 = m * n + 1
    table = rng.random((m * n + 2, A1)).astype(np.float32)
    table /= table.sum(axis=1, keepdims=True)
    values = rng.uniform(0.02, 0.98, m * n + 2).astype(np.float32)
    _run_pair(oracle, m, n, k, sims, table, values, seed)


@pytest.mark.parametrize("rows,cols,nn,sims,seed", [(6, 7, 4, 40, 11), (5, 5, 3, 60, 12), (4, 6, 4, 25, 13)])
def test_c4_search_vs_python(oracle, rows, cols, nn, sims, seed):
    """Connect-4 under the same flow: Pass is a legal child of every node, Apply never flips the side to move (so the
    whole descent evaluates for the searching colour), MoveNumber() is the constant 1, no tree reuse, DontPreferPass picks
    the first non-pass child when Pass sorts first."""
    rng = np.random.default_rng(seed)
    table = rng.random((3, cols + 1)).astype(np.float32)
    table[:, cols] *= 3  # make Pass competitive so that noPass matters
    table /= table.sum(axis=1, keepdims=True)
    values = rng.uniform(0.05, 0.95, 3).astype(np.float32)
    d = K.make_desc(K.GAME_C4, rows, cols, nn, sims=sims, nn=H.tiny_nn(rows, cols, cols + 1), n_games=1, flags=K.FLAG_SHARED_TREE,
                    seed=seed, mcts_m=rows, mcts_n=cols)
    e = oracle.create(d)
    e.set_inferer(0, K.INF_TABLE)
    e.set_table(0, table, values)
    e.arena_begin(1, False)
    g = P.C4(rows, cols, nn)
    t = P.MCTS(g, 1.0, sims, table, values, rows, cols)
    player, moves_py, alive, ply = P.BLACK, [], 1, 0
    while alive:
        assert not g.ended()[0]
        best = t.search(player)
        alive = e.arena_step()
        got = e.tree_dump(0, 0).astype(np.int64) & 0xFFFFFFFF
        want = t.dump() & 0xFFFFFFFF
        assert got.shape == want.shape and (got == want).all(), (ply, got[:3], want[:3])
        g.apply(player, best)
        moves_py.append(best)
        player = P.opponent(player)
        ply += 1
    rec = e.game_record(0)
    e.arena_finish()
    assert list(rec["moves"]) == moves_py and g.ended() == (True, rec["winner"])
