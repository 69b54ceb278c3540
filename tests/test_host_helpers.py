"""Host-side mirrors of the reference's helpers around the path (agogo_b200/host.py): RotateBoard against the reference's
own test (encoding_helper_test.go:10-58), the two GameEncoders against the planes the engine's device-side twins put into
the Arena's examples (oracle library on CPU; bit for bit, -0.0 included), Agent.Search / Agent.Infer on a caller-owned
game.State."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from agogo_b200 import host as Hh
from tests import helpers as H
from tests import pyref_rules as R


def test_rotate_board_reference_case():
    W, N, B = 2.0, 0.0, 1.0   # any three values: the function moves float32s around
    board = [W, N, N, N, B,
             N, W, N, B, N,    # this line is to break rotational symmetry
             N, N, N, N, N,
             N, N, N, N, N,
             B, N, N, N, W]
    r = board
    seen = []
    for _ in range(4):
        r = Hh.RotateBoard(r, 5, 5)
        seen.append(r.tolist())
    assert seen[3] == board, "After 4 rotations the board should be the same"
    assert seen[0] != board and seen[1] != board and seen[0] != seen[2]
    # a quarter turn counter-clockwise, element by element as the reference moves them: new[i][j] = old[j][m-1-i]
    assert all(seen[0][i * 5 + j] == board[j * 5 + (4 - i)] for i in range(5) for j in range(5))
    with pytest.raises(ValueError):
        Hh.RotateBoard([0] * 6, 2, 3)
    r4 = Hh.RotateBoard(list(range(16)), 4, 4).tolist()
    assert all(r4[i * 4 + j] == j * 4 + (3 - i) for i in range(4) for j in range(4))


def _replay_wq(moves, size):
    """boards before every move of a wq game under the reference's rules as written (pyref_rules.wq_apply)"""
    b, player, before = [0] * (size * size), 1, []
    for mv in moves:
        before.append(list(b))
        if mv != K.PASS:
            _, _, b, _ = R.wq_apply(b, size, player, int(mv))
        player = 3 - player
    return before


def test_wq_encoder_equals_the_arena_examples(oracle):
    size = 5
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=0.5, sims=6, n_games=1, seed=3, max_moves=24,
                    nn=H.tiny_nn(size, size, size * size + 1, features=18))
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    e.arena_play(1, True)
    moves = e.game_record(0)["moves"]
    boards, _, _ = e.examples(clear=True)
    assert len(boards) == len(moves) >= 20
    before = _replay_wq(moves, size)
    for ply in range(len(moves)):
        st = Hh.State(board=before[ply], to_move=1 + ply % 2, move_number=ply, hist=before[max(0, ply - 8):ply])
        mine = Hh.WQEncoder(st)
        assert (mine.view(np.uint32) == boards[ply].view(np.uint32)).all(), ply
    assert np.signbit(boards[12][boards[12] == 0]).any()      # the negated group's empties are -0.0 and that is compared
    with pytest.raises(IndexError):
        Hh.State(board=before[9], move_number=9, hist=before[7:9]).Historical(3)
    e.close()


def test_two_plane_encoder_equals_the_arena_examples(oracle):
    d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=8, n_games=1, seed=5, nn=H.tiny_nn(3, 3, 9, features=2), encoder=K.ENC_TWO_PLANE)
    e = oracle.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    e.arena_play(1, True)
    moves = e.game_record(0)["moves"]
    boards, _, _ = e.examples(clear=True)
    b = [0] * 9
    for ply, mv in enumerate(moves):
        st = Hh.State(board=list(b), to_move=1 + ply % 2, move_number=ply)
        assert (Hh.EncodeBoard(st).view(np.uint32) == boards[ply].view(np.uint32)).all(), ply
        b[int(mv)] = 1 + ply % 2
    e.close()


def test_agent_search_and_infer_on_a_state(oracle):
    g = Hh.Game(K.GAME_MNK, 3, 3, 3)
    conf = Hh.Config(NNConf=Hh.DualConfig(K=3, SharedLayers=1, FC=8, BatchSize=2, Width=3, Height=3, Features=2, ActionSpace=9),
                     MCTSConf=Hh.MCTSConfig(PUCT=1.0, M=3, N=3, Sims=60, Budget=1000, RandomTemperature=1.0), Encoder=K.ENC_TWO_PLANE)
    az = Hh.AZ(g, conf, lib=oracle, n_games=1, seed=2)
    az.A.useDummy()
    # X to move with two in a row: the search takes the win
    st = Hh.State(board=[1, 1, 0, 2, 2, 0, 0, 0, 0], to_move=K.BLACK, move_number=4, last_move=4)
    assert az.A.Search(st) == 2
    az.A.Close()
    az.A.SwitchToInference()
    pol, val = az.A.Infer(st)
    p2, v2 = az.A.NNOutput(Hh.EncodeBoard(st)[None])
    # (the reference's own init overflows on a net this small — compare bits, NaNs included)
    assert pol.shape == (9,) and (pol.view(np.uint32) == p2[0].view(np.uint32)).all()
    assert np.float32(val).view(np.uint32) == v2[:1].view(np.uint32)[0]


def test_shuffle_batch_in_place_like_the_reference(oracle):
    """dualnet's TestShuffleBatch (dual_test.go:160-183): after a pass the three tensors are no longer in their original
    order.  Here additionally: az_train shuffles the caller's buffers in place (as the reference shuffles its tensors,
    meta.go:47), all three with the SAME row permutation, and that permutation is host.shuffle_rows' for the same seed
    (Fisher-Yates with j = r.Intn(i+1), meta.go:57-102)."""
    nn = dict(k=3, shared_layers=1, fc=8, batch_size=4, features=2, action_space=10)
    e = oracle.create(K.make_desc(K.GAME_MNK, 3, 3, 3, sims=2, n_games=1, seed=1, nn=nn))
    rng = np.random.default_rng(7)
    n = 12
    X = rng.uniform(150, 152, (n, 18)).astype(np.float32)
    Pi = rng.uniform(0, 1, (n, 10)).astype(np.float32)
    V = rng.uniform(0, 1, n).astype(np.float32)
    X0, P0, V0 = X.copy(), Pi.copy(), V.copy()
    e.train(1, X, Pi, V, 3, 1, lr=0.0, shuffle_seed=1234)
    assert not (X == X0).all() and not (Pi == P0).all() and not (V == V0).all()
    Xh, Ph, Vh = X0.copy(), P0.copy(), V0.copy()
    Hh.shuffle_rows(Xh, Ph, Vh, Hh.Rng(1234))
    assert (X == Xh).all() and (Pi == Ph).all() and (V == Vh).all()
    perm = [int(np.flatnonzero((X0 == row).all(axis=1))[0]) for row in X]
    assert sorted(perm) == list(range(n)) and (Pi == P0[perm]).all() and (V == V0[perm]).all()
    e.close()


def test_state_clone_eq_like_the_reference():
    """game/wq's TestGameBasics (game_test.go:9-20): a clone is equal; changing the side to move of the parent makes them
    unequal."""
    g = Hh.State(board=np.zeros(19 * 19, np.int32))
    g2 = g.Clone()
    assert g.Eq(g2), "Expected clones to be equal"
    g.SetToMove(K.WHITE)
    assert not g.Eq(g2), "Expected clones to be unequal after the parent object has changed"
    g2.SetToMove(K.WHITE)
    g2.board[3] = K.BLACK
    assert not g.Eq(g2) and g.Board()[3] == 0     # the clone owns its board
