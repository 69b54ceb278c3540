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


@pytest.mark.parametrize("m,n,k,sims,dummy,seed", [(3, 3, 3, 30, (0, 0), 11), (3, 3, 3, 60, (1, 2), 12), (4, 4, 3, 25, (2, 1), 13)])
def test_arena_two_agents_vs_python(oracle, m, n, k, sims, dummy, seed):
    """Arena.Play with two agents, each searching its own tree (reuse across its own plies: two moves deep), the
    dummyInferer, the coin flips of the injected RNG stream, examples (two-plane encoder, Policies = normalised counts of
    the chosen move, colour -> outcome labels) and the win/loss/draw statistics: the oracle against the Python restatement."""
    G_ = 6
    d = K.make_desc(K.GAME_MNK, m, n, k, sims=sims, nn=H.tiny_nn(m, n, m * n + 1), n_games=G_, seed=seed)
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, dummy[0]); e.set_inferer(1, K.INF_DUMMY, dummy[1])
    run = H.play_and_collect(e, G_)
    state = P.derive_seed(seed, 0)
    ex_all, stats = [], [[0, 0, 0], [0, 0, 0]]
    for g in range(G_):
        state, r = P._splitmix(state)
        coin = r % 2
        def make(agent, game):
            return P.MCTS(game, 1.0, sims, None, None, m, n, evaluator=P.dummy_evaluator(m * n, dummy[agent]))
        moves, winner, a_player, examples, dumps = P.arena_play(lambda: P.MNK(m, n, k), make, coin)
        rec = run["records"][g]
        assert list(rec["moves"]) == moves and rec["winner"] == winner and rec["a_player"] == a_player and rec["n_examples"] == len(examples)
        for ply, (ta, tb) in enumerate(dumps):
            for t, want in enumerate((ta, tb)):
                got = run["dumps"][ply][g][t].astype(np.int64) & 0xFFFFFFFF
                assert got.shape == want.shape and (got == (want & 0xFFFFFFFF)).all(), (g, ply, t)
        ex_all += examples
        b_player = P.opponent(a_player)
        if winner == 0:
            stats[0][2] += 1; stats[1][2] += 1
        elif winner == a_player:
            stats[0][0] += 1; stats[1][1] += 1
        elif winner == b_player:
            stats[1][0] += 1; stats[0][1] += 1
    boards, pols, vals = run["examples"]
    assert len(ex_all) == len(vals)
    for i, (b, p, v) in enumerate(ex_all):
        assert (boards[i] == b).all() and (pols[i] == p).all() and vals[i] == v, i
    for a in (0, 1):
        assert tuple(run["stats"][a]) == tuple(float(x) for x in stats[a])


@pytest.mark.parametrize("size,sims,cap,dummy,seed", [(5, 20, 30, (1, 2), 21), (5, 40, 24, (0, 0), 22), (7, 12, 20, (2, 1), 23)])
def test_arena_go_vs_python(oracle, size, sims, cap, dummy, seed):
    """Go (wq) through the same Arena: Pass is a child of every node, two passes end a descent with combinedScore
    (blackScore - whiteScore - komi), Apply returns new states, no tree reuse, 18-plane WQEncoder with its history
    planes, the move cap; Board.check / Apply / Score from tests/pyref_rules.py."""
    G_ = 4
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=sims, n_games=G_, seed=seed, max_moves=cap,
                    nn=H.tiny_nn(size, size, size * size + 1, features=18))
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, dummy[0]); e.set_inferer(1, K.INF_DUMMY, dummy[1])
    run = H.play_and_collect(e, G_)
    state = P.derive_seed(seed, 0)
    ex_all = []
    for g in range(G_):
        state, r = P._splitmix(state)
        def make(agent, game):
            return P.MCTS(game, 1.0, sims, None, None, size, size, evaluator=P.dummy_evaluator(size * size, dummy[agent]))
        moves, winner, a_player, examples, dumps = P.arena_play(lambda: P.WQ(size, 7.5), make, r % 2, encoder=P.encode_wq18, max_moves=cap)
        rec = run["records"][g]
        assert list(rec["moves"]) == moves and rec["winner"] == winner and rec["a_player"] == a_player and rec["n_examples"] == len(examples), (g, rec, moves)
        for ply, pair in enumerate(dumps):
            for t, want in enumerate(pair):
                got = run["dumps"][ply][g][t].astype(np.int64) & 0xFFFFFFFF
                assert got.shape == want.shape and (got == (want & 0xFFFFFFFF)).all(), (g, ply, t)
        ex_all += examples
    boards, pols, vals = run["examples"]
    assert len(ex_all) == len(vals)
    for i, (b, p, v) in enumerate(ex_all):
        assert (boards[i] == b).all() and (pols[i] == p).all() and vals[i] == v, i


@pytest.mark.parametrize("game,workers,sims,seed", [("ttt", 2, 30, 31), ("ttt", 5, 48, 32), ("ttt", 64, 40, 33), ("go5", 4, 22, 34), ("go5", 7, 30, 35)])
def test_worker_schedule_vs_python(oracle, game, workers, sims, seed):
    """mcts.Config workers > 1: the oracle's SearchRunWorkers against the Python statement of the same fixed schedule
    (rounds of descents with virtual-loss flags — White-only penalty, node.go:150-152 — colliding leaves, short last round)."""
    G_ = 3
    if game == "ttt":
        m = n = 3
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=sims, nn=H.tiny_nn(3, 3, 10), n_games=G_, seed=seed, workers=workers)
        new_game, enc, cap, cells = (lambda: P.MNK(3, 3, 3)), None, 0, 9
    else:
        m = n = 5
        d = K.make_desc(K.GAME_WQ, 5, 5, 0, komi=7.5, sims=sims, n_games=G_, seed=seed, max_moves=16, workers=workers,
                        nn=H.tiny_nn(5, 5, 26, features=18))
        new_game, enc, cap, cells = (lambda: P.WQ(5, 7.5)), P.encode_wq18, 16, 25
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    run = H.play_and_collect(e, G_)
    state = P.derive_seed(seed, 0)
    for g in range(G_):
        state, r = P._splitmix(state)
        def make(agent, gm):
            return P.MCTS(gm, 1.0, sims, None, None, m, n, evaluator=P.dummy_evaluator(cells, 1 + agent), workers=workers)
        moves, winner, a_player, examples, dumps = P.arena_play(new_game, make, r % 2, encoder=enc, max_moves=cap)
        rec = run["records"][g]
        assert list(rec["moves"]) == moves and rec["winner"] == winner, (g, rec, moves)
        for ply, pair in enumerate(dumps):
            for t, want in enumerate(pair):
                got = run["dumps"][ply][g][t].astype(np.int64) & 0xFFFFFFFF
                assert got.shape == want.shape and (got == (want & 0xFFFFFFFF)).all(), (g, ply, t)


@pytest.mark.parametrize("name,conf", [
    ("sample-T1", dict(random_count=5, random_min_visits=1, random_temperature=1.0)),
    ("sample-T0.5", dict(random_count=9, random_min_visits=2, random_temperature=0.5)),
    ("preferpass-smart", dict(pass_preference=1, dumb_pass=False)),
    ("dontresign-smart", dict(pass_preference=2, dumb_pass=False, resign_percentage=0.3)),
    ("dontpreferpass-smart", dict(pass_preference=0, dumb_pass=False)),
])
def test_config_options_vs_python(oracle, name, conf):
    """The remaining mcts.Config options inside bestMove (search.go:341-390): temperature sampling of the root's children
    (randomizeChildren with the injected tree RNG, math32.Pow), PassPreference x DumbPass (the `LastMove().IsPass()` branch
    fires on the empty history too), shouldResign's early exits — oracle against the Python restatement, on tic-tac-toe
    for the sampling cases and on 5x5 Go for the pass logic."""
    G_, seed, sims = 4, 41, 24
    go = "smart" in name
    if go:
        d = K.make_desc(K.GAME_WQ, 5, 5, 0, komi=7.5, sims=sims, n_games=G_, seed=seed, max_moves=14,
                        nn=H.tiny_nn(5, 5, 26, features=18), pass_preference=conf["pass_preference"], dumb_pass=0)
        new_game, enc, cap, cells, m = (lambda: P.WQ(5, 7.5)), P.encode_wq18, 14, 25, 5
    else:
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=sims, nn=H.tiny_nn(3, 3, 10), n_games=G_, seed=seed)
        new_game, enc, cap, cells, m = (lambda: P.MNK(3, 3, 3)), None, 0, 9, 3
    d.mcts.random_count = conf.get("random_count", 0)
    d.mcts.random_min_visits = conf.get("random_min_visits", 0)
    d.mcts.random_temperature = conf.get("random_temperature", 0.0)
    d.mcts.resign_percentage = conf.get("resign_percentage", 0.0)
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    run = H.play_and_collect(e, G_)
    state = P.derive_seed(seed, 0)
    tree_seed = P.derive_seed(seed, 1)
    for g in range(G_):
        state, r = P._splitmix(state)
        def make(agent, gm):
            return P.MCTS(gm, 1.0, sims, None, None, m, m, evaluator=P.dummy_evaluator(cells, 1 + agent), conf=conf, tree_seed=P.derive_seed(tree_seed, 2 * g + agent))  # tree t of game g: stream 2g + t
        moves, winner, a_player, examples, dumps = P.arena_play(new_game, make, r % 2, encoder=enc, max_moves=cap)
        rec = run["records"][g]
        assert list(rec["moves"]) == moves and rec["winner"] == winner and rec["n_examples"] == len(examples), (name, g, rec, moves)
        for ply, pair in enumerate(dumps):
            for t, want in enumerate(pair):
                got = run["dumps"][ply][g][t].astype(np.int64) & 0xFFFFFFFF
                assert got.shape == want.shape and (got == (want & 0xFFFFFFFF)).all(), (name, g, ply, t)


@pytest.mark.parametrize("size,sims,cap,seed", [(3, 200, 40, 51), (4, 250, 60, 52), (5, 24, 60, 53)])
def test_arena_go_complete_rules_vs_python(oracle, size, sims, cap, seed):
    """AZ_FLAG_WQ_COMPLETE through the whole Arena, sampled play: ko, positional superko INSIDE the trees (a descent may not
    recreate a position of the game or of its own path — the tiny boards with many simulations are where it fires), own-eye
    and suicide bars, area-scored two-pass endings.  The oracle (group BFS, board list walked per candidate) against
    pyref_mcts.WQComplete (trial move + set of position tuples): every tree after every ply, moves, winners, examples."""
    G_ = 4
    cells = size * size
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=0.5, sims=sims, n_games=G_, seed=seed, max_moves=cap, flags=K.FLAG_WQ_COMPLETE,
                    nn=H.tiny_nn(size, size, cells + 1, features=18))
    conf = dict(random_count=cap, random_min_visits=0, random_temperature=1.0)
    d.mcts.random_count, d.mcts.random_temperature = cap, 1.0
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    run = H.play_and_collect(e, G_)
    state = P.derive_seed(seed, 0)
    tree_seed = P.derive_seed(seed, 1)
    rejected = [0]
    from tests import pyref_rules as R
    orig = R.wq_complete_check

    def counting(board, sz, player, move, ko=-1, positions=None):  # how often superko alone decides
        res = orig(board, sz, player, move, ko, positions)
        if not res[0] and positions and orig(board, sz, player, move, ko)[0]:
            rejected[0] += 1
        return res

    R.wq_complete_check = counting
    try:
        for g in range(G_):
            state, r = P._splitmix(state)
            def make(agent, gm):
                return P.MCTS(gm, 1.0, sims, None, None, size, size, evaluator=P.dummy_evaluator(cells, 1 + agent), conf=conf,
                              tree_seed=P.derive_seed(tree_seed, 2 * g + agent))
            moves, winner, a_player, examples, dumps = P.arena_play(lambda: P.WQComplete(size, 0.5), make, r % 2, encoder=P.encode_wq18, max_moves=cap)
            rec = run["records"][g]
            assert list(rec["moves"]) == moves and rec["winner"] == winner and rec["n_examples"] == len(examples), (g, rec, moves, winner)
            for ply, pair in enumerate(dumps):
                for t, want in enumerate(pair):
                    got = run["dumps"][ply][g][t].astype(np.int64) & 0xFFFFFFFF
                    assert got.shape == want.shape and (got == (want & 0xFFFFFFFF)).all(), (g, ply, t)
    finally:
        R.wq_complete_check = orig
    print("superko-only rejections inside the searches:", rejected[0])
    assert rejected[0] >= (10 if size < 5 else 1)
