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
