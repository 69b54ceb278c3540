"""AZ.Learn composition (agogo.go:100-172): the host-side mirror (agogo_b200/host.py) driving the
C ABI must reproduce the oracle's native restatement of the same loop, epoch by epoch — statistics,
example counts, promotion decisions, training costs and the final weights (BASELINE config C1:
tic-tac-toe, cmd/tictactoe shapes; "plumbing, no GPU")."""
import ctypes as C

import numpy as np
import pytest

from agogo_b200 import _capi as K
from agogo_b200 import host


def _c1_conf(batch, sims):
    nn = host.DefaultConf(3, 3, 10)  # cmd/tictactoe/main.go:63-70
    nn.BatchSize, nn.Features, nn.K, nn.SharedLayers = batch, 2, 3, 3
    mc = host.MCTSConfig(PUCT=1.0, M=3, N=3, Timeout=100_000_000, PassPreference=K.DONT_PREFER_PASS, Budget=1000,
                         DumbPass=True, RandomCount=0, Sims=sims)
    return host.Config(Name="Tic Tac Toe", NNConf=nn, MCTSConf=mc, UpdateThreshold=0.52, Encoder=K.ENC_TWO_PLANE)


def _native_learn(oracle, az, iters, episodes, nniters, arena_games, threshold):
    """oracle::AZ::Learn through the oracle-only entry point azo_learn."""
    d = K.EngineDesc.from_buffer_copy(bytes(az.engine.desc))
    d.seed = az.seed
    f = oracle.dll.azo_learn
    f.restype = C.c_int
    log = np.zeros((iters, 11), np.float32)
    nfl = az.engine.param_count()[1]
    final = np.zeros(nfl, np.float32)
    rc = f(C.byref(d), C.c_double(threshold), C.c_int32(0), C.c_int32(iters), C.c_int32(episodes), C.c_int32(nniters),
           C.c_int32(arena_games), log.ctypes.data_as(C.POINTER(C.c_float)), final.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0, oracle.dll.az_last_error(None)
    return log, final


def test_c1_learn_host_vs_native_oracle(oracle):
    conf = _c1_conf(batch=20, sims=20)
    az = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=8, seed=42)
    assert az.engine.param_desc(0)[0] == "FilterInit"
    iters, episodes, nniters, arena_games = 3, 12, 4, 10
    az.Learn(iters, episodes, nniters, arena_games)
    log, final = _native_learn(oracle, az, iters, episodes, nniters, arena_games, conf.UpdateThreshold)
    for ep in range(iters):
        h, n = az.log[ep], log[ep]
        assert h["a"] == tuple(n[0:3]) and h["b"] == tuple(n[3:6]), (ep, h, n)
        assert h["n_examples"] == int(n[6]) and h["batches"] == int(n[7]) and int(h["promoted"]) == int(n[8])
        assert np.float32(h["first_cost"]) == n[9] and np.float32(h["last_cost"]) == n[10], (ep, h, n)
    assert (az.engine.net_get(0).view(np.uint32) == final.view(np.uint32)).all()


def test_invalid_configs_panic(oracle):
    conf = _c1_conf(20, 10)
    conf.MCTSConf.PUCT = 1.5  # mcts.Config.IsValid (tree.go:43-45)
    with pytest.raises(RuntimeError):
        host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle)
    conf = _c1_conf(20, 10)
    conf.NNConf.ActionSpace = 2  # dual.Config.IsValid (config.go:33-42)
    with pytest.raises(RuntimeError):
        host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle)


def test_round_kats():
    from tests.golden import rules_golden as G
    for a, want in G.ROUND:
        assert host.dual_round(a) == want


def test_cpp_host_learn_equals_python_host(oracle, tmp_path):
    """host/agogo.hpp (C++ host layer) linked against the oracle library reproduces the Python host's
    AZ.Learn exactly: same per-epoch lines and the same final weights."""
    import os
    import subprocess
    import zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "tictactoe_oracle")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", os.path.join(root, "host", "tictactoe.cpp"), "-o", exe,
                           "-L" + os.path.join(root, "oracle"), "-lazoracle", "-Wl,-rpath," + os.path.join(root, "oracle"),
                           "-fopenmp"])
    out = subprocess.check_output([exe, "2", "10", "3", "8", "16", "20", "77"], cwd=str(tmp_path)).decode().splitlines()
    conf = _c1_conf(batch=20, sims=16)
    az = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=64, seed=77)
    az.Learn(2, 10, 3, 8)
    for ep, l in enumerate(az.log):
        want = "epoch %d A %g %g %g B %g %g %g examples %d batches %d promoted %d cost %.9g %.9g" % (
            ep, *l["a"], *l["b"], l["n_examples"], l["batches"], int(l["promoted"]), np.float32(l["first_cost"]),
            np.float32(l["last_cost"]))
        assert out[ep] == want, (out[ep], want)
    p = az.engine.net_get(0)
    h = 2166136261
    for byte in p.tobytes():
        h = ((h ^ byte) * 16777619) & 0xFFFFFFFF
    assert out[-1].split()[1] == "%08x" % h


def test_arena_play_hooks(oracle):
    """(*Arena).Play(record, enc, aug): the OutputEncoder sees the MetaState after every move (arena.go:131-133) and the
    Augmenter multiplies every kept example (arena.go:115-121); Config.Augmenter is applied by Learn's batched self-play."""
    conf = _c1_conf(batch=20, sims=12)
    az = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=4, seed=8)
    az.setupSelfPlay(0)

    class Enc:
        def __init__(self):
            self.calls = []

        def Encode(self, ms):
            st = ms.State()
            self.calls.append((ms.Name(), ms.Epoch(), ms.GameNumber(), st["move_number"], int((st["board"] != 0).sum())))
            self.scores = (ms.Score(K.BLACK), ms.Score(K.WHITE), st["winner"])

        def Flush(self):
            return None

    def mirror(ex):  # a board symmetry: the kind of Augmenter the reference's README shows
        b = ex.Board.reshape(2, 3, 3)[:, :, ::-1].reshape(-1).copy()
        p = np.concatenate([ex.Policy[:9].reshape(3, 3)[:, ::-1].reshape(-1), ex.Policy[9:]])
        return [ex, host.Example(b, p, ex.Value)]

    enc = Enc()
    winner, plain = az.Play(True, None, None)
    az2 = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=4, seed=8)
    az2.setupSelfPlay(0)
    winner2, aug = az2.Play(True, enc, mirror, game_number=3)
    assert winner == K.NONE and winner2 == K.NONE                      # arena.go:178
    n_moves = len(az2.engine.game_record(0)["moves"])
    assert len(enc.calls) == n_moves and [c[3] for c in enc.calls] == list(range(1, n_moves + 1))
    assert all(c[0] == "Tic Tac Toe" and c[2] == 3 for c in enc.calls) and [c[4] for c in enc.calls] == list(range(1, n_moves + 1))
    assert len(aug) == 2 * len(plain) and len(plain) > 0
    # MetaState.Score / Arena.Score = the game's Score(p) (arena.go:191): mnk's 1 / -2 / 0 (mnk.go:147-155)
    want = {K.NONE: (0.0, 0.0), K.BLACK: (1.0, -2.0), K.WHITE: (-2.0, 1.0)}[enc.scores[2]]
    assert enc.scores[:2] == want and (az2.Score(K.BLACK), az2.Score(K.WHITE)) == want
    assert isinstance(az2, host.Arena) and (az2.Name(), az2.Epoch(), az2.GameNumber()) == ("Tic Tac Toe", 0, 3)
    assert az2.State()["move_number"] == n_moves
    import io
    buf = io.StringIO()
    az2.Log(buf)
    assert "A:" in buf.getvalue() and "B:" in buf.getvalue()
    for i, x in enumerate(plain):
        assert (aug[2 * i].Board == x.Board).all() and aug[2 * i].Value == x.Value
        assert (aug[2 * i + 1].Board.reshape(2, 3, 3) == x.Board.reshape(2, 3, 3)[:, :, ::-1]).all()
    # Config.Augmenter inside Learn's batched self-play
    conf.Augmenter = mirror
    az3 = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=4, seed=8)
    az3.setupSelfPlay(0)
    assert len(az3._play(1, True)) == len(aug)
