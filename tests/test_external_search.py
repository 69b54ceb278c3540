"""Agent.Search on external positions (az_search) with tree reuse across calls (updateRoot / newRootState,
mcts/search.go:424-500): a game driven move by move from outside — each agent searching the positions it meets with the
state's move history attached — must reproduce the Arena's own game (whose per-ply re-rooting is pinned against the
Python restatement in test_oracle_mcts_pyref.py): same moves, same root visit counts."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from tests import helpers as H


def external_game(lib, m, n, k, sims, seed, with_moves=True, table=None, max_plies=64):
    """Two agents (own trees) alternate on one mnk game through az_search; returns moves, per-ply visit vectors."""
    d = K.make_desc(K.GAME_MNK, m, n, k, sims=sims, nn=H.tiny_nn(m, n, m * n + 1), n_games=2, seed=seed)
    e = lib.create(d)
    if table is None:
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    else:
        e.set_table(0, *table); e.set_table(1, table[0][::-1].copy(), table[1][::-1].copy())
    board = np.zeros(m * n, np.int32)
    moves, visits, hist = [], [], []
    player = K.BLACK
    for ply in range(min(m * n, max_plies)):
        agent = ply % 2  # A = Black moves first
        best, v = e.search(agent, board, player, player, move_number=ply, last_move=hist[-1][1] if hist else K.PASS,
                           moves=np.array(hist, np.int32).reshape(-1, 2) if (with_moves and hist) else None)
        moves.append(best); visits.append(v.copy())
        if best < 0 or board[best] != 0:
            break
        board[best] = player
        hist.append((player, best))
        ended, winner, _, _ = e.rules_status(board)
        if ended[0]:
            break
        player = K.WHITE if player == K.BLACK else K.BLACK
    e.close()
    return moves, visits


def arena_game(lib, m, n, k, sims, seed, table=None):
    d = K.make_desc(K.GAME_MNK, m, n, k, sims=sims, nn=H.tiny_nn(m, n, m * n + 1), n_games=2, seed=seed)
    e = lib.create(d)
    if table is None:
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    else:
        e.set_table(0, *table); e.set_table(1, table[0][::-1].copy(), table[1][::-1].copy())
    # the coin decides A's colour; retry seeds until A is Black so that the external driver's convention matches
    e.arena_begin(1, False)
    if e.game_record(0)["a_player"] != K.BLACK:
        e.arena_finish(); e.close()
        return None
    visits, n_act = [], 1
    while n_act:
        n_act = e.arena_step()
        rec = e.game_record(0)
        t = e.tree_dump(0, (len(rec["moves"]) - 1) % 2)
        visits.append(t)
    rec = e.game_record(0)
    e.arena_finish(); e.close()
    return list(rec["moves"]), visits


def _root_visits_from_dump(dump, A):
    v = np.zeros(A + 1, np.float32)
    for row in dump[1:]:
        if row[0] == 1:
            v[A if row[1] == K.PASS else row[1]] = row[2]
    return v


def check_external_equals_arena(lib, m, n, k, sims, table=None):
    for seed in range(1, 40):
        got = arena_game(lib, m, n, k, sims, seed, table)
        if got is not None:
            break
    a_moves, a_dumps = got
    x_moves, x_visits = external_game(lib, m, n, k, sims, seed, True, table)
    assert x_moves[:len(a_moves)] == a_moves, (a_moves, x_moves)
    for ply, dump in enumerate(a_dumps[:len(x_visits)]):
        assert (_root_visits_from_dump(dump, m * n) == x_visits[ply]).all(), ply
    # reuse really happened: from each agent's second search on, the root's children hold more visits than one search gives
    assert x_visits[2].sum() >= sims + (x_visits[2] > 0).sum()
    if sims > m * n:
        assert x_visits[2].sum() > sims + (x_visits[2] > 0).sum()
    # without the move history every call starts a fresh tree: each root child is born with one visit + `sims` descents
    f_moves, f_visits = external_game(lib, m, n, k, sims, seed, False, table)
    for ply, v in enumerate(f_visits):
        assert v.sum() == sims + (v > 0).sum(), ply
    return a_moves


def _table(A1, seed):
    rng = np.random.default_rng(seed)
    t = rng.random((40, A1)).astype(np.float32)
    t /= t.sum(axis=1, keepdims=True)
    return t, rng.uniform(0.05, 0.95, 40).astype(np.float32)


@pytest.mark.parametrize("m,n,k,sims,tab", [(3, 3, 3, 30, False), (3, 3, 3, 50, True), (4, 4, 3, 40, True), (5, 5, 4, 24, False)])
def test_oracle_external_reuse_equals_arena(oracle, m, n, k, sims, tab):
    check_external_equals_arena(oracle, m, n, k, sims, _table(m * n + 1, 3) if tab else None)


def check_reuse_edge_cases(lib_a, lib_b):
    """Same call sequences on two libraries (oracle / engine): re-search of the same position (depth 0), a jump to an
    unrelated line (fresh root), reset_tree, a continuation whose move is not a child (findChild fails inside Search)."""
    outs = []
    for lib in (lib_a, lib_b):
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=20, nn=H.tiny_nn(3, 3, 10), n_games=1, seed=5)
        e = lib.create(d)
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        out = []
        b0 = np.zeros(9, np.int32)
        out.append(e.search(0, b0, K.BLACK, K.BLACK, move_number=0))
        out.append(e.search(0, b0, K.BLACK, K.BLACK, move_number=0))            # depth 0: same root searched again
        b2 = b0.copy(); b2[4] = K.BLACK; b2[0] = K.WHITE
        mv = np.array([[K.BLACK, 4], [K.WHITE, 0]], np.int32)
        out.append(e.search(0, b2, K.BLACK, K.BLACK, move_number=2, last_move=0, moves=mv))   # continuation
        b2x = b0.copy(); b2x[8] = K.BLACK; b2x[7] = K.WHITE
        out.append(e.search(0, b2x, K.BLACK, K.BLACK, move_number=2, last_move=7, moves=np.array([[K.BLACK, 8], [K.WHITE, 7]], np.int32)))  # other line
        e.reset_tree(0)
        out.append(e.search(0, b2x, K.BLACK, K.BLACK, move_number=2, last_move=7))
        out.append(e.search(1, b2x, K.BLACK, K.BLACK, move_number=2, last_move=7))            # the other agent's tree
        outs.append(out)
        e.close()
    for (ba, va), (bb, vb) in zip(*outs):
        assert ba == bb and (va == vb).all(), (ba, bb, va, vb)
    # depth-0 re-search accumulates on the same root
    assert outs[0][1][1].sum() > outs[0][0][1].sum()


def test_oracle_reuse_edge_cases_self_consistent(oracle):
    check_reuse_edge_cases(oracle, oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k,sims,tab", [(3, 3, 3, 30, False), (3, 3, 3, 50, True), (4, 4, 3, 40, True), (5, 5, 4, 24, False)])
def test_engine_external_reuse_equals_arena_and_oracle(oracle, engine_lib, m, n, k, sims, tab):
    table = _table(m * n + 1, 3) if tab else None
    moves_o = check_external_equals_arena(oracle, m, n, k, sims, table)
    moves_e = check_external_equals_arena(engine_lib, m, n, k, sims, table)
    assert moves_o == moves_e


@pytest.mark.gpu
def test_engine_reuse_edge_cases_vs_oracle(oracle, engine_lib):
    check_reuse_edge_cases(oracle, engine_lib)
