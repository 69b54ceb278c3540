"""dual.Train restatement checks that need no GPU: the oracle's hand-derived backward pass against
finite differences, and the host-side training loop (gradient averaging, row shuffling) against the
library's own az_train."""
import numpy as np

from agogo_b200 import _capi as K
from agogo_b200 import host
from tests import helpers as H


def _data(rng, B, plane, A1):
    X = rng.normal(size=(B, plane)).astype(np.float32)
    Pi = np.zeros((B, A1), np.float32)
    Pi[np.arange(B), rng.integers(0, A1, B)] = 1
    V = rng.choice([-1.0, 0.0, 1.0], B).astype(np.float32)
    return X, Pi, V


def test_oracle_backward_matches_finite_differences(oracle):
    d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=2, n_games=1, seed=1,
                    nn=dict(k=3, shared_layers=2, fc=5, batch_size=6, features=2, action_space=10))
    e = oracle.create(d)
    p0 = H.tame_gammas([e], 1, 5, target=0.05)
    rng = np.random.default_rng(0)
    X, Pi, V = _data(rng, 6, 18, 10)
    g, c0 = e.train_grads(1, X, Pi, V)
    assert np.isfinite(g).all() and np.isfinite(c0)

    def cost(p):
        e.net_set(1, p)
        return e.train_grads(1, X, Pi, V)[1]

    nt, _ = e.param_count()
    checked = ok = 0
    for i in range(nt):
        name, shape, off, size = e.param_desc(i)
        for j in rng.integers(0, size, 3):
            k = off + int(j)
            eps = 2e-2 * max(1.0, abs(p0[k]))
            pp, pm = p0.copy(), p0.copy()
            pp[k] += eps
            pm[k] -= eps
            num = (cost(pp) - cost(pm)) / (2 * eps)
            checked += 1
            ok += abs(num - g[k]) <= 2e-3 + 2e-2 * abs(num)
    assert ok >= 0.9 * checked, (ok, checked)  # the misses are ReLU kinks inside +-eps


def test_host_training_loop_equals_az_train(oracle):
    """host.AZ._train with host-side gradient exchange (world=1) == the library's az_train, bit for bit."""
    nn = host.DualConfig(K=3, SharedLayers=2, FC=6, BatchSize=8, Width=3, Height=3, Features=2, ActionSpace=10)
    conf = host.Config(NNConf=nn, MCTSConf=host.MCTSConfig(M=3, N=3, Sims=4))
    a = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=2, seed=9)
    b = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=2, seed=9, host_allreduce=True)
    rng = np.random.default_rng(3)
    X, Pi, V = _data(rng, 24, 18, 10)
    ca = a._train(X.copy(), Pi.copy(), V.copy(), 3, 4, 1234)
    cb = b._train(X.copy(), Pi.copy(), V.copy(), 3, 4, 1234)
    assert (np.asarray(ca).view(np.uint32) == np.asarray(cb).view(np.uint32)).all()
    assert (a.engine.net_get(1).view(np.uint32) == b.engine.net_get(1).view(np.uint32)).all()
