"""The oracle against every golden vector the reference's own rule tests hold."""
from tests import helpers as H


def test_oracle_mnk_golden(oracle):
    H.check_mnk_golden(oracle)


def test_oracle_c4_golden(oracle):
    H.check_c4_golden(oracle)


def test_oracle_wq_golden(oracle):
    H.check_wq_golden(oracle)
