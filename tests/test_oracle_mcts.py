"""The oracle's search against the reference's only search KAT: mcts/example_test.go `Example`
(scripted dummyNN, documented move sequence 4,0,2,6,3,5,1,7,8, pinned output `WINNER None`)."""
import pytest

from tests import helpers as H


@pytest.mark.parametrize("sims", [10, 50, 200])
def test_oracle_ttt_example_kat(oracle, sims):
    H.check_ttt_kat(oracle, sims)


def test_oracle_search_external_position(oracle):
    """Agent.Search on a caller-supplied position: a tic-tac-toe board where X (to move) wins at 2."""
    import numpy as np
    from agogo_b200 import _capi as K
    from tests import helpers as H
    e = oracle.create(K.make_desc(K.GAME_MNK, 3, 3, 3, sims=60, nn=H.tiny_nn(3, 3, 10), n_games=1, seed=1))
    e.set_inferer(0, K.INF_DUMMY, 0)
    board = np.array([1, 1, 0, 2, 2, 0, 0, 0, 0], np.int32)
    best, visits = e.search(0, board, K.BLACK, K.BLACK, move_number=4)
    assert best in (2, 5, 6, 7, 8) and visits[[0, 1, 3, 4]].sum() == 0 and visits.sum() >= 60


@pytest.mark.parametrize("workers", [2, 5, 16, 100])
def test_oracle_concurrent_workers_invariants(oracle, workers):
    """mcts.Config workers > 1 — the fixed schedule of the reference's concurrent searchStates (search.go:112-130):
    every Search still runs exactly `sims` pipeline calls; every non-null call updates the root once; virtual-loss
    flags never outlive a round (the next Search's tree equals a single-worker continuation in structure); with the
    Example's scripted evaluator the documented draw is still found."""
    import numpy as np
    from agogo_b200 import _capi as K
    sims = 48
    e = oracle.create(K.make_desc(K.GAME_MNK, 3, 3, 3, sims=sims, nn=H.tiny_nn(3, 3, 10), n_games=3, seed=5, workers=workers))
    e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    e.arena_begin(3, True)
    e.search_begin(); e.search_run(sims)
    c = e.counters()
    assert c["sims"] == 3 * sims and c["searches"] == 3
    for g in range(3):
        for t in (0, 1):
            d = e.tree_dump(g, t)
            if d.shape[0] == 0:
                continue
            root = d[0]
            # root: born with 1 visit, +1 for its own expansion, +1 per non-null pipeline call from it
            kids = d[d[:, 0] == 1]
            assert root[2] >= 2 and kids[:, 2].sum() - len(kids) <= root[2]
    e.search_end()
    n = 3
    while n:
        n = e.arena_step()
    e.arena_finish()
    c = e.counters()
    assert c["sims"] == c["searches"] * sims
    assert c["evals"] + c["null_results"] <= c["sims"] + c["searches"]   # one evaluation per non-null call + the root's
    for g in range(3):
        assert 5 <= len(e.game_record(g)["moves"]) <= 9


def test_oracle_single_worker_is_default(oracle):
    """workers = 0 and workers = 1 are the canonical single-worker search."""
    from agogo_b200 import _capi as K
    runs = []
    for w in (0, 1):
        e = oracle.create(K.make_desc(K.GAME_MNK, 3, 3, 3, sims=30, nn=H.tiny_nn(3, 3, 10), n_games=2, seed=4, workers=w))
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        runs.append(H.play_and_collect(e, 2))
    H.assert_same_run(runs[0], runs[1], "workers 0 vs 1")


def test_oracle_workers_golden(oracle):
    H.check_workers_golden(oracle)
