"""N>1 host logic on CPU: world_size-2 `gloo` run of the sharded AZ.Learn (games sharded, examples
all-gathered, gradients averaged every step, win counts summed) over the oracle library."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from agogo_b200 import _capi as K
    from agogo_b200 import host
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = K.load(os.path.join(ROOT, "oracle", "libazoracle.so"))
    nn = host.DualConfig(K=3, SharedLayers=2, FC=8, BatchSize=10, Width=3, Height=3, Features=2, ActionSpace=10)
    mc = host.MCTSConfig(PUCT=1.0, M=3, N=3, Sims=12)
    conf = host.Config(NNConf=nn, MCTSConf=mc, UpdateThreshold=0.52)
    az = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=lib, n_games=8, seed=5, dist=dist)
    assert not az.engine_comm  # the oracle has no NCCL: the host group carries the gradients
    az.Learn(2, 10, 3, 6)
    np.save(os.path.join(out_dir, "params_%d.npy" % rank), np.concatenate([az.engine.net_get(0), az.engine.net_get(1)]))
    np.save(os.path.join(out_dir, "log_%d.npy" % rank),
            np.array([[*l["a"], *l["b"], l["n_examples"], l["batches"], l["promoted"], l["first_cost"], l["last_cost"]]
                      for l in az.log], np.float64))
    dist.destroy_process_group()


def test_two_rank_learn_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0, p1 = np.load(tmp_path / "params_0.npy"), np.load(tmp_path / "params_1.npy")
    l0, l1 = np.load(tmp_path / "log_0.npy"), np.load(tmp_path / "log_1.npy")
    assert np.isfinite(p0).all()
    assert (p0.view(np.uint32) == p1.view(np.uint32)).all(), "replicas diverged"
    assert (l0[:, :9] == l1[:, :9]).all()  # global statistics, example counts and decisions agree
    assert (l0[:, 0:3].sum(axis=1) == 6).all()  # 6 arena games in total across both ranks
    assert (l0[:, 6] > 0).all()
