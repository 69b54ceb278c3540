"""The oracle's search against the reference's only search KAT: mcts/example_test.go `Example`
(scripted dummyNN, documented move sequence 4,0,2,6,3,5,1,7,8, pinned output `WINNER None`)."""
import pytest

from tests import helpers as H


@pytest.mark.parametrize("sims", [10, 50, 200])
def test_oracle_ttt_example_kat(oracle, sims):
    H.check_ttt_kat(oracle, sims)
