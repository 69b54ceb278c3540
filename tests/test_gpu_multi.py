"""Multi-GPU (needs >= 2 devices; skipped otherwise): sharded AZ.Learn with the engine's own NCCL
gradient all-reduce — replicas must end bit-identical."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, collective):
    sys.path.insert(0, ROOT)
    os.environ["AZ_TRAIN_COLLECTIVE"] = collective
    import torch
    import torch.distributed as dist
    from agogo_b200 import _capi as K
    from agogo_b200 import host
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nn = host.DualConfig(K=16, SharedLayers=2, FC=32, BatchSize=16, Width=7, Height=6, Features=2, ActionSpace=8)
    mc = host.MCTSConfig(PUCT=1.0, M=6, N=7, Sims=16)
    conf = host.Config(NNConf=nn, MCTSConf=mc, UpdateThreshold=0.52)
    az = host.AZ(host.Game(K.GAME_C4, 6, 7, 4), conf, n_games=8, seed=5, dist=dist, device=rank, flags=K.FLAG_FP32_TOWER)
    assert az.engine_comm
    az.Learn(2, 4 * world, 2, 6)
    np.save(os.path.join(out_dir, "params_%s_%d.npy" % (collective, rank)),
            np.concatenate([az.engine.net_get(0), az.engine.net_get(1)]))
    dist.destroy_process_group()


def _ngpu():
    import subprocess
    try:
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout.strip().splitlines())
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_learn_nccl(tmp_path, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)  # (decided without importing torch: a cold import costs minutes on a fresh box)
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 2000) + 10 * world
    res = {}
    for i, collective in enumerate(("p2p", "nccl")):  # the fused peer-memory kernel, and plain ncclAllReduce + SGD
        mp.spawn(_worker, args=(world, port + i, str(tmp_path), collective), nprocs=world, join=True)
        p0 = np.load(tmp_path / ("params_%s_0.npy" % collective))
        assert np.isfinite(p0).all()
        for r in range(1, world):
            pr = np.load(tmp_path / ("params_%s_%d.npy" % (collective, r)))
            assert (p0.view(np.uint32) == pr.view(np.uint32)).all(), "replicas diverged (%s, rank %d)" % (collective, r)
        res[collective] = p0
    # same mathematics, different summation order inside the collective
    assert np.abs(res["p2p"] - res["nccl"]).max() <= 1e-4 * max(1.0, np.abs(res["nccl"]).max())
