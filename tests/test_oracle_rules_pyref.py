"""The C++ oracle's rules (oracle/mnk.hpp, c4.hpp, wq.hpp through az_rules_apply / az_rules_status) against a second,
independent plain-Python restatement of the same Go sources (tests/pyref_rules.py) on random positions — the reference's
own tests hold 8 + 6 + 7 boards (tests/golden); this widens the pin on the oracle itself."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from tests import helpers as H
from tests import pyref_rules as R


def _boards(rng, n, cells, p_empty):
    p = [p_empty, (1 - p_empty) / 2, (1 - p_empty) / 2]
    return rng.choice([0, 1, 2], size=(n, cells), p=p).astype(np.int32)


@pytest.mark.parametrize("m,n,k", [(3, 3, 3), (5, 5, 4), (4, 6, 3), (6, 4, 4)])
def test_mnk_rules_vs_python(oracle, m, n, k):
    rng = np.random.default_rng(m * 100 + n * 10 + k)
    e = H.rules_engine(oracle, K.GAME_MNK, m, n, k)
    boards = np.concatenate([_boards(rng, 150, m * n, 0.5), _boards(rng, 150, m * n, 0.15)])
    players = rng.integers(1, 3, len(boards)).astype(np.int32)
    moves = rng.integers(-1, m * n, len(boards)).astype(np.int32)
    chk, app, out, taken = e.rules_apply(boards, players, moves)
    ended, winner, sb, sw = e.rules_status(boards)
    for i, b in enumerate(boards.tolist()):
        c, a, o, t = R.mnk_apply(b, int(players[i]), int(moves[i]))
        assert (bool(chk[i]), bool(app[i]), out[i].tolist(), int(taken[i])) == (c, a, o, t), (i, b, moves[i])
        assert (bool(ended[i]), int(winner[i]), float(sb[i]), float(sw[i])) == R.mnk_status(b, m, n, k), (i, b)


@pytest.mark.parametrize("rows,cols,nn", [(6, 7, 4), (5, 5, 3), (4, 8, 4)])
def test_c4_rules_vs_python(oracle, rows, cols, nn):
    rng = np.random.default_rng(rows * 100 + cols * 10 + nn)
    e = H.rules_engine(oracle, K.GAME_C4, rows, cols, nn)
    boards = np.concatenate([_boards(rng, 150, rows * cols, 0.5), _boards(rng, 150, rows * cols, 0.1)])
    players = rng.integers(1, 3, len(boards)).astype(np.int32)
    moves = rng.integers(-1, cols, len(boards)).astype(np.int32)
    passes = rng.integers(0, 5, len(boards)).astype(np.int32)
    chk, app, out, taken = e.rules_apply(boards, players, moves)
    ended, winner, sb, sw = e.rules_status(boards, passes)
    for i, b in enumerate(boards.tolist()):
        c, a, o, t = R.c4_apply(b, rows, cols, int(players[i]), int(moves[i]))
        assert (bool(chk[i]), bool(app[i]), out[i].tolist(), int(taken[i])) == (c, a, o, t), (i, b, moves[i])
        assert (bool(ended[i]), int(winner[i]), float(sb[i]), float(sw[i])) == R.c4_status(b, rows, cols, nn, int(passes[i])), (i, b)


@pytest.mark.parametrize("size", [5, 7, 9])
def test_wq_rules_vs_python(oracle, size):
    rng = np.random.default_rng(size)
    e = H.rules_engine(oracle, K.GAME_WQ, size, size)
    cells = size * size
    base = np.concatenate([_boards(rng, 12, cells, 0.45), _boards(rng, 12, cells, 0.2)])
    boards, players, moves = [], [], []
    for j, b in enumerate(base):
        for mv in range(cells):  # every point, occupied ones included (Game.Check does not reject them)
            boards.append(b); players.append(1 + (mv + j) % 2); moves.append(mv)
    boards = np.array(boards, np.int32)
    chk, app, out, taken = e.rules_apply(boards, players, moves)
    for i, b in enumerate(boards.tolist()):
        c, a, o, t = R.wq_apply(b, size, players[i], moves[i])
        assert (bool(chk[i]), bool(app[i]), out[i].tolist(), int(taken[i])) == (c, a, o, t), (i, moves[i])
    _, _, sb, sw = e.rules_status(base)
    for i, b in enumerate(base.tolist()):
        assert (float(sb[i]), float(sw[i])) == (R.wq_score(b, size, 1), R.wq_score(b, size, 2)), i


def test_wq_complete_rules_oracle_vs_pyref(oracle):
    """AZ_FLAG_WQ_COMPLETE (our completion of the reference's unfinished Go rules): legality of every point (occupied,
    suicide, own-eye fill), captures and area scores of random 5x5 / 7x7 positions — oracle (group/liberty BFS) against
    the naive trial-move restatement in pyref_rules.py."""
    rng = np.random.default_rng(11)
    for size in (5, 7):
        d = K.make_desc(K.GAME_WQ, size, size, 0, komi=5.5, sims=1, n_games=1, flags=K.FLAG_WQ_COMPLETE,
                        nn=H.tiny_nn(size, size, size * size + 1, features=2), encoder=K.ENC_TWO_PLANE)
        e = oracle.create(d)
        cells = size * size
        for _ in range(40):
            b = rng.choice([0, 1, 2], size=cells, p=[0.35, 0.33, 0.32]).astype(np.int32)
            # drop dead groups so that the position is a possible Go position
            for p in range(cells):
                if b[p] and not R._wq_group(list(b), size, p)[1]:
                    for q in R._wq_group(list(b), size, p)[0]:
                        b[q] = 0
            boards = np.repeat(b[None], 2 * cells, axis=0)
            players = [1] * cells + [2] * cells
            moves = list(range(cells)) * 2
            check, applied, out, taken = e.rules_apply(boards, players, moves)
            for i, (pl, mv) in enumerate(zip(players, moves)):
                legal, captured, _ = R.wq_complete_check(list(b), size, pl, mv)
                assert bool(check[i]) == legal and bool(applied[i]) == legal, (size, b.tolist(), pl, mv, check[i], legal)
                if legal:
                    want = list(b)
                    want[mv] = pl
                    for q in captured:
                        want[q] = 0
                    assert out[i].tolist() == want and taken[i] == len(captured)
                else:
                    assert out[i].tolist() == b.tolist()
            _, _, sb, sw = e.rules_status(b[None], passes=[2])
            assert sb[0] == R.wq_area_score(list(b), size, 1) and sw[0] == R.wq_area_score(list(b), size, 2)
        e.close()


def _complete_game_vs_pyref(lib, size, sims, seed, max_moves):
    """One Arena game under AZ_FLAG_WQ_COMPLETE: after every search the root's children must be exactly the legal moves
    (+ Pass) of the Python restatement with its own ko and position-set tracking; returns how many kos, captures and
    superko-only rejections (a point legal under simple ko but recreating an earlier position) occurred."""
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=5.5, sims=sims, n_games=1, seed=seed, max_moves=max_moves,
                    flags=K.FLAG_WQ_COMPLETE, nn=H.tiny_nn(size, size, size * size + 1, features=18))
    d.mcts.random_count, d.mcts.random_temperature = max_moves, 1.0   # sampled play: varied games with fights
    e = lib.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    e.arena_begin(1, False)
    board, ko, kos, caps, sk = [0] * (size * size), -1, 0, 0, 0
    positions = set()
    player = 1
    n_act, ply = 1, 0
    while n_act:
        n_act = e.arena_step()
        rec = e.game_record(0)
        a_black = rec["a_player"] == 1
        agent = 0 if (ply % 2 == 0) == a_black else 1
        dump = e.tree_dump(0, agent)
        kids = {int(r[1]) for r in dump if r[0] == 1}
        legal = {p for p in range(size * size) if R.wq_complete_check(board, size, player, p, ko, positions)[0]} | {K.PASS}
        assert kids == legal, (ply, sorted(kids ^ legal), ko)
        sk += sum(1 for p in range(size * size) if p not in legal and R.wq_complete_check(board, size, player, p, ko)[0])
        mv = int(rec["moves"][ply])
        positions.add(tuple(board))
        if mv == K.PASS:
            ko = -1
        else:
            ok, captured, ko = R.wq_complete_check(board, size, player, mv, ko, positions)
            assert ok, (ply, mv)
            board[mv] = player
            for q in captured:
                board[q] = 0
            caps += len(captured)
            kos += ko >= 0
        assert e.game_state(0)["board"].tolist() == board, ply
        player = 3 - player
        ply += 1
    st = e.game_state(0)
    rec = e.game_record(0)
    e.arena_finish()
    if rec["moves"][-2:].tolist() == [K.PASS, K.PASS]:  # ended by two passes: area score + komi decides
        sb, sw = R.wq_area_score(board, size, 1), R.wq_area_score(board, size, 2) + 5.5
        assert e.game_record(0)["winner"] == (0 if sb == sw else (1 if sb > sw else 2))
    e.close()
    return kos, caps, sk, ply


def test_wq_complete_games_oracle_vs_pyref(oracle):
    kos = caps = sks = 0
    for seed in range(1, 9):
        k, c, sk, plies = _complete_game_vs_pyref(oracle, 5, 20, seed, 90)
        kos += k; caps += c; sks += sk
    print("kos", kos, "captures", caps, "superko-only rejections", sks)
    assert caps > 20 and kos > 0, (kos, caps)   # the games really fought (captures, at least one ko shape)


def _superko_external_vs_pyref(lib, size, seed, rounds):
    """Positional superko on external states: random positions handed to Agent.Search together with <= 14 earlier boards,
    some of which are exactly what a legal move (capturing or not) would recreate.  The root's children must be the
    restatement's legal set (+ Pass).  Returns how many points superko alone rejected."""
    cells = size * size
    d = K.make_desc(K.GAME_WQ, size, size, 0, komi=0.5, sims=2, n_games=1, seed=seed, flags=K.FLAG_WQ_COMPLETE,
                    nn=H.tiny_nn(size, size, cells + 1, features=2), encoder=K.ENC_TWO_PLANE)
    e = lib.create(d)
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    rng = np.random.default_rng(seed)
    sk = 0
    for _ in range(rounds):
        b = rng.choice([0, 1, 2], size=cells, p=[0.3, 0.35, 0.35]).tolist()
        for p in range(cells):  # drop dead groups: a possible Go position
            if b[p] and not R._wq_group(b, size, p)[1]:
                for q in R._wq_group(b, size, p)[0]:
                    b[q] = 0
        player = int(rng.integers(1, 3))
        results = []
        for p in range(cells):
            ok, captured, _ = R.wq_complete_check(b, size, player, p)
            if ok:
                after = list(b); after[p] = player
                for q in captured:
                    after[q] = 0
                results.append(after)
        n_hist = int(rng.integers(0, 15))   # more than the encoder's 8: superko reads every board the caller hands over
        window = []
        for _ in range(n_hist):
            if results and rng.random() < 0.6:
                window.append(results[int(rng.integers(len(results)))])
            else:
                window.append(rng.choice([0, 1, 2], size=cells).tolist())
        positions = {tuple(x) for x in window}
        legal = {p for p in range(cells) if R.wq_complete_check(b, size, player, p, -1, positions)[0]}
        sk += sum(1 for p in range(cells) if p not in legal and R.wq_complete_check(b, size, player, p)[0])
        e.reset_tree(0)
        _, visits = e.search(0, b, player, player, move_number=20 + n_hist, passes=0,
                             hist=np.array(window, np.int32) if window else None, last_move=0, ko=-1)
        kids = {(K.PASS if i == cells else i) for i in np.nonzero(visits)[0].tolist()}  # children are born with one visit
        assert kids == legal | {K.PASS}, (size, seed, b, player, window, sorted(kids ^ (legal | {K.PASS})))
    e.close()
    return sk


def test_wq_superko_external_oracle_vs_pyref(oracle):
    sk = sum(_superko_external_vs_pyref(oracle, size, seed, 40) for size, seed in ((3, 1), (5, 2), (7, 3), (9, 4)))
    print("superko-only rejections", sk)
    assert sk >= 100
