"""GPU parity tests: the CUDA engine through the C ABI against the reference's golden vectors and,
call for call, against the CPU oracle on identical seeds.  Bit-exact for trees, moves, examples,
labels, statistics and counters (integer / fp32-bit comparisons)."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from tests import helpers as H
from tests.golden import rules_golden as G

pytestmark = pytest.mark.gpu


def test_engine_mnk_golden(engine_lib):
    H.check_mnk_golden(engine_lib)


def test_engine_c4_golden(engine_lib):
    H.check_c4_golden(engine_lib)


def test_engine_wq_golden(engine_lib):
    H.check_wq_golden(engine_lib)


@pytest.mark.parametrize("sims", [10, 50, 200])
def test_engine_ttt_example_kat(engine_lib, sims):
    H.check_ttt_kat(engine_lib, sims)


def _pair(oracle, engine_lib, desc_fn):
    return oracle.create(desc_fn()), engine_lib.create(desc_fn())


def _setup_dummy(e):
    # agogo.go:83-87: epoch-0 self-play, dummyInferer captured with Player == None -> value 0
    e.set_inferer(0, K.INF_DUMMY, 0)
    e.set_inferer(1, K.INF_DUMMY, 0)


def test_parity_ttt_example_trees(oracle, engine_lib):
    """Full tree state after every Search of the Example game, 4 lockstep copies."""
    eo, eg = H.ttt_example_engine(oracle, 50, 4), H.ttt_example_engine(engine_lib, 50, 4)
    a, b = H.play_and_collect(eo, 4, record=True), H.play_and_collect(eg, 4, record=True)
    H.assert_same_run(a, b, "ttt-example")


@pytest.mark.parametrize("sims", [5, 40])
def test_parity_ttt_arena_dummy(oracle, engine_lib, sims):
    """C1's epoch-0 self-play: tic-tac-toe, two agents, two trees with reuse, dummy inferer."""
    def desc():
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=sims, nn=H.tiny_nn(3, 3, 10), n_games=16, seed=11)
    eo, eg = _pair(oracle, engine_lib, desc)
    for e in (eo, eg):
        _setup_dummy(e)
    H.assert_same_run(H.play_and_collect(eo, 16), H.play_and_collect(eg, 16), "ttt-arena")
    # a second batch continues the coin stream and the statistics
    H.assert_same_run(H.play_and_collect(eo, 16), H.play_and_collect(eg, 16), "ttt-arena-2")


def test_parity_gomoku_arena_dummy(oracle, engine_lib):
    def desc():
        return K.make_desc(K.GAME_MNK, 5, 5, 4, sims=12, nn=H.tiny_nn(5, 5, 26), n_games=6, seed=3)
    eo, eg = _pair(oracle, engine_lib, desc)
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUMMY, 1)
        e.set_inferer(1, K.INF_DUMMY, 2)
    H.assert_same_run(H.play_and_collect(eo, 6), H.play_and_collect(eg, 6), "gomoku")


def test_parity_c4_arena_dummy(oracle, engine_lib):
    def desc():
        return K.make_desc(K.GAME_C4, 6, 7, 4, sims=30, nn=H.tiny_nn(6, 7, 8), n_games=8, seed=5)
    eo, eg = _pair(oracle, engine_lib, desc)
    for e in (eo, eg):
        _setup_dummy(e)
    H.assert_same_run(H.play_and_collect(eo, 8), H.play_and_collect(eg, 8), "c4")


@pytest.mark.parametrize("size,sims,plies", [(5, 24, 40), (9, 16, 30)])
def test_parity_wq_arena_dummy(oracle, engine_lib, size, sims, plies):
    def desc():
        return K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=sims, n_games=4, seed=9, max_moves=plies,
                           nn=H.tiny_nn(size, size, size * size + 1, features=18))
    eo, eg = _pair(oracle, engine_lib, desc)
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUMMY, 1)
        e.set_inferer(1, K.INF_DUMMY, 2)
    H.assert_same_run(H.play_and_collect(eo, 4), H.play_and_collect(eg, 4), "wq")


@pytest.mark.parametrize("game,workers,sims", [
    ("ttt", 2, 40), ("ttt", 8, 50), ("ttt", 64, 50), ("gomoku", 3, 25), ("c4", 4, 30), ("c4shared", 4, 30), ("wq5", 4, 26), ("wq9", 16, 40),
])
def test_parity_concurrent_workers(oracle, engine_lib, game, workers, sims):
    """mcts.Config workers > 1 (the reference's concurrent searchStates, search.go:112-130, under the fixed schedule
    of include/agogo_b200.h): rounds of `workers` descents with virtual-loss flags, one evaluation batch per round.
    Trees after every ply, moves, examples and counters bit-identical to the oracle's restatement of the same
    schedule; sims not a multiple of workers leaves a short last round; workers > sims runs a single round."""
    def desc():
        if game == "ttt":
            return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=sims, nn=H.tiny_nn(3, 3, 10), n_games=8, seed=11, workers=workers)
        if game == "gomoku":
            return K.make_desc(K.GAME_MNK, 5, 5, 4, sims=sims, nn=H.tiny_nn(5, 5, 26), n_games=6, seed=3, workers=workers)
        if game in ("c4", "c4shared"):
            return K.make_desc(K.GAME_C4, 6, 7, 4, sims=sims, nn=H.tiny_nn(6, 7, 8), n_games=8, seed=5, workers=workers,
                               flags=K.FLAG_SHARED_TREE if game == "c4shared" else 0)
        size = 5 if game == "wq5" else 9
        return K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=sims, n_games=4, seed=9, max_moves=30, workers=workers,
                           nn=H.tiny_nn(size, size, size * size + 1, features=18))
    eo, eg = _pair(oracle, engine_lib, desc)
    n = eo.desc.n_games
    rng = np.random.default_rng(7)
    A1 = eo.desc.nn.action_space
    table = rng.random((64, A1)).astype(np.float32)
    table /= table.sum(axis=1, keepdims=True)
    values = rng.uniform(0.05, 0.95, 64).astype(np.float32)
    for e in (eo, eg):
        if game in ("ttt", "wq9"):
            e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        else:  # scripted evaluator: policy and value depend on the leaf's move number -> distinct priors per depth
            e.set_table(0, table, values); e.set_table(1, table[::-1].copy(), values[::-1].copy())
    a, b = H.play_and_collect(eo, n), H.play_and_collect(eg, n)
    H.assert_same_run(a, b, "workers-%s-%d" % (game, workers))
    ca, cb = dict(a["counters"]), dict(b["counters"])
    ca.pop("kernel_launches", None); cb.pop("kernel_launches", None)
    assert ca == cb
    assert ca["sims"] == ca["searches"] * sims


def test_engine_workers_golden(engine_lib):
    """The committed worker-schedule fixture (generated from the oracle) reproduced by the device engine, bit for bit."""
    H.check_workers_golden(engine_lib)


def test_concurrent_workers_dual_net(oracle, engine_lib):
    """The same schedule with the dual network in the loop (fp32 tower): batch = games x workers per round."""
    def desc():
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=24, nn=H.tiny_nn(3, 3, 10, k=4, layers=2, fc=8), n_games=4, seed=2,
                           workers=6, flags=K.FLAG_FP32_TOWER)
    eo, eg = _dual_pair(oracle, engine_lib, desc)
    H.assert_same_run(H.play_and_collect(eo, 4), H.play_and_collect(eg, 4), "workers-dual", float_ulps=64)


def test_engines_of_different_sizes_coexist(oracle, engine_lib):
    """Kernel attributes (dynamic shared memory opt-in) are per function and per device, not per engine: a small-board
    engine created after a large-board one must not shrink what the large one needs."""
    def big():
        return K.make_desc(K.GAME_WQ, 9, 9, 0, komi=7.5, sims=8, n_games=2, seed=9, max_moves=12, nn=H.tiny_nn(9, 9, 82, features=18))
    def small():
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=8, nn=H.tiny_nn(3, 3, 10), n_games=2, seed=1)
    eo, eg = _pair(oracle, engine_lib, big)
    so, sg = _pair(oracle, engine_lib, small)          # created while the 9x9 engines are alive
    for e in (eo, eg, so, sg):
        _setup_dummy(e)
    H.assert_same_run(H.play_and_collect(so, 2), H.play_and_collect(sg, 2), "small")
    H.assert_same_run(H.play_and_collect(eo, 2), H.play_and_collect(eg, 2), "big-after-small")


def test_wq_random_rules_vs_oracle(oracle, engine_lib):
    """Board.check / Board.Apply on random (also inconsistent) 7x7 positions, every point, both colours."""
    rng = np.random.default_rng(1)
    eo, eg = H.rules_engine(oracle, K.GAME_WQ, 7, 7), H.rules_engine(engine_lib, K.GAME_WQ, 7, 7)
    boards, players, moves = [], [], []
    for _ in range(60):
        b = rng.choice([0, 1, 2], size=49, p=[0.4, 0.3, 0.3]).astype(np.int32)
        for mv in range(49):
            boards.append(b); players.append(1 + (mv + len(boards)) % 2); moves.append(mv)
    boards = np.array(boards, np.int32)
    ra, rb = eo.rules_apply(boards, players, moves), eg.rules_apply(boards, players, moves)
    for x, y, name in zip(ra, rb, ("check", "applied", "boards", "taken")):
        assert (x == y).all(), name
    sa, sb = eo.rules_status(boards[::49]), eg.rules_status(boards[::49])
    for x, y in zip(sa, sb):
        assert (x == y).all()


def _dual_pair(oracle, engine_lib, desc_fn, seeds=(21, 22)):
    eo, eg = _pair(oracle, engine_lib, desc_fn)
    for e in (eo, eg):
        e.net_init(0, seeds[0]); e.net_init(1, seeds[1])
    assert (eo.net_get(0).view(np.uint32) == eg.net_get(0).view(np.uint32)).all(), "net init differs"
    H.tame_gammas([eo, eg], 0, seeds[0]); H.tame_gammas([eo, eg], 1, seeds[1])
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUAL); e.set_inferer(1, K.INF_DUAL)
    return eo, eg


def test_fp32_forward_vs_oracle(oracle, engine_lib):
    """dualnet forward, fp32 CUDA-core path, tolerance 1e-4 (north star); typically ~1e-6."""
    def desc():
        return K.make_desc(K.GAME_WQ, 9, 9, 0, komi=7.5, sims=4, n_games=8, seed=2, max_moves=8, flags=K.FLAG_FP32_TOWER,
                           nn=dict(k=8, shared_layers=2, fc=16, batch_size=4, features=18, action_space=82))
    eo, eg = _dual_pair(oracle, engine_lib, desc)
    rng = np.random.default_rng(0)
    planes = rng.choice([0.0, 1.0, -1.0, 0.001], size=(8, 18 * 81)).astype(np.float32)
    po, vo = eo.infer(0, planes)
    pg, vg = eg.infer(0, planes)
    assert np.isfinite(po).all() and np.isfinite(vo).all() and np.abs(vo).max() < 0.999
    assert np.abs(po - pg).max() < 1e-4 and np.abs(vo - vg).max() < 1e-4
    assert np.abs(pg.sum(axis=1) - 1).max() < 1e-5


def test_parity_ttt_arena_dual_fp32(oracle, engine_lib):
    """Two random dual nets (cmd/tictactoe shapes: K=3, 3 blocks, FC=8) driving the search."""
    def desc():
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=25, n_games=8, seed=4, flags=K.FLAG_FP32_TOWER,
                           nn=dict(k=3, shared_layers=3, fc=8, batch_size=100, features=2, action_space=10))
    eo, eg = _dual_pair(oracle, engine_lib, desc)
    H.assert_same_run(H.play_and_collect(eo, 8), H.play_and_collect(eg, 8), "ttt-dual", float_ulps=64)


def test_parity_wq_arena_dual_fp32(oracle, engine_lib):
    """9x9 Go with two small random dual nets: varied priors, so captures / occupied-point no-ops /
    the WQEncoder history planes all enter the trees; fp32 tower (op order identical to the oracle)."""
    def desc():
        return K.make_desc(K.GAME_WQ, 7, 7, 0, komi=7.5, sims=20, n_games=4, seed=13, max_moves=36, flags=K.FLAG_FP32_TOWER,
                           nn=dict(k=4, shared_layers=1, fc=8, batch_size=4, features=18, action_space=50))
    eo, eg = _dual_pair(oracle, engine_lib, desc)
    H.assert_same_run(H.play_and_collect(eo, 4), H.play_and_collect(eg, 4), "wq-dual", float_ulps=64)


def _wq_planes(rng, n, size):
    """Plausible WQEncoder planes: stones +-1 / 0 in history planes, one to-move plane set."""
    x = np.zeros((n, 18, size * size), np.float32)
    for b in range(n):
        for q in range(0, 7):
            board = rng.choice([0.0, 1.0, -1.0], size=size * size, p=[0.6, 0.2, 0.2]).astype(np.float32)
            x[b, q] = board
            x[b, 8 + q] = -board
        x[b, 16 if b % 2 == 0 else 17] = 1.0 if b % 2 == 0 else -1.0
    return x.reshape(n, -1)


@pytest.mark.parametrize("k,layers,size,fc,n", [(64, 6, 9, 128, 6), (128, 2, 9, 64, 3), (256, 3, 19, 512, 3)])
def test_tc_tower_vs_fp32_and_oracle(oracle, engine_lib, k, layers, size, fc, n):
    """tcgen05 tower (fp16 hi/lo split, 3 passes) against the engine's fp32 CUDA-core tower and the
    CPU oracle on the same weights: |dp|, |dv| < 1e-4 (north star tolerance)."""
    A1 = size * size + 1
    def desc(flags):
        return K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=2, n_games=8, seed=2, max_moves=4, flags=flags,
                           nn=dict(k=k, shared_layers=layers, fc=fc, batch_size=2, features=18, action_space=A1))
    eo = oracle.create(desc(0))
    e32 = engine_lib.create(desc(K.FLAG_FP32_TOWER))
    etc = engine_lib.create(desc(0))
    H.tame_gammas([eo, e32, etc], 0, 77)
    for e in (eo, e32, etc):
        e.set_inferer(0, K.INF_DUAL)
    planes = _wq_planes(np.random.default_rng(5), n, size)
    po, vo = eo.infer(0, planes)
    p32, v32 = e32.infer(0, planes)
    ptc, vtc = etc.infer(0, planes)
    assert np.isfinite(po).all() and np.abs(vo).max() < 0.9999
    assert np.abs(p32 - po).max() < 1e-4 and np.abs(v32 - vo).max() < 1e-4
    err_p, err_v = np.abs(ptc - po).max(), np.abs(vtc - vo).max()
    print("tc vs oracle: dp=%.3g dv=%.3g ; tc vs fp32: dp=%.3g dv=%.3g ; pmax=%.3g" %
          (err_p, err_v, np.abs(ptc - p32).max(), np.abs(vtc - v32).max(), po.max()))
    assert err_p < 1e-4 and err_v < 1e-4


def test_tc_selfplay_runs(oracle, engine_lib):
    """9x9 self-play with the tensor-core tower in the loop: moves legal-by-construction, finite
    outputs, same move sequences as the oracle for the first plies (priors agree to ~1e-6)."""
    def desc():
        return K.make_desc(K.GAME_WQ, 9, 9, 0, komi=7.5, sims=16, n_games=4, seed=21, max_moves=6,
                           nn=dict(k=64, shared_layers=2, fc=32, batch_size=2, features=18, action_space=82))
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    H.tame_gammas([eo, eg], 0, 31); H.tame_gammas([eo, eg], 1, 32)
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUAL); e.set_inferer(1, K.INF_DUAL)
    a, b = H.play_and_collect(eo, 4, dump_trees=False), H.play_and_collect(eg, 4, dump_trees=False)
    for ra, rb in zip(a["records"], b["records"]):
        assert list(ra["moves"]) == list(rb["moves"])
    assert a["counters"]["evals"] == b["counters"]["evals"]


def _train_data(rng, B, plane, A1):
    X = rng.choice([0.001, 1.0, -1.0], size=(B, plane)).astype(np.float32)
    Pi = np.zeros((B, A1), np.float32)
    Pi[np.arange(B), rng.integers(0, A1, B)] = 1
    V = rng.choice([-1.0, 0.0, 1.0], B).astype(np.float32)
    return X, Pi, V


@pytest.mark.parametrize("kind,m,n,k,nn", [
    (K.GAME_MNK, 3, 3, 3, dict(k=3, shared_layers=3, fc=8, batch_size=20, features=2, action_space=10)),   # cmd/tictactoe
    (K.GAME_C4, 6, 7, 4, dict(k=16, shared_layers=2, fc=32, batch_size=16, features=2, action_space=8)),   # C4 shapes
    # K = 64: the 3x3 convs (forward and backward-data) run on the tcgen05 kernel (hi/lo split, raw epilogue)
    (K.GAME_MNK, 9, 9, 5, dict(k=64, shared_layers=2, fc=32, batch_size=4, features=2, action_space=82)),
])
def test_train_grads_vs_oracle(oracle, engine_lib, kind, m, n, k, nn):
    """One Train step (BN train mode, xent-on-logits + MSE, reverse mode): CUDA gradients of every
    Model() tensor against the oracle's (itself pinned by finite differences)."""
    def desc():
        return K.make_desc(kind, m, n, k, sims=2, n_games=2, seed=1, nn=nn, flags=K.FLAG_FP32_TOWER)
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    H.tame_gammas([eo, eg], 1, 17, target=0.05)
    rng = np.random.default_rng(2)
    X, Pi, V = _train_data(rng, nn["batch_size"], 2 * m * n, nn["action_space"])
    go, co = eo.train_grads(1, X, Pi, V)
    gg, cg = eg.train_grads(1, X, Pi, V)
    assert abs(co - cg) <= 1e-5 * max(1, abs(co))
    scale = np.abs(go).max()
    assert np.abs(go - gg).max() <= 2e-4 * scale, (np.abs(go - gg).max(), scale)
    # and the solver step
    eo.train_apply(1, go, 0.1); eg.train_apply(1, gg, 0.1)
    assert np.abs(eo.net_get(1) - eg.net_get(1)).max() <= 1e-4 * max(1.0, scale)


def test_train_tensor_core_vs_fp32_full_width(engine_lib, monkeypatch):
    """K7 at the C3 width (19x19, K=256): the tcgen05 training convs (forward, backward-data, backward-filter with
    split-K) against the engine's own fp32 CUDA-core kernels on the same batch — size-independent A/B, the fp32 path
    being the one pinned against the oracle above."""
    nn = dict(k=256, shared_layers=2, fc=64, batch_size=8, features=18, action_space=362)
    def desc():
        return K.make_desc(K.GAME_WQ, 19, 19, 0, komi=7.5, sims=2, n_games=2, seed=1, nn=nn, max_moves=4, flags=K.FLAG_FP32_TOWER)
    rng = np.random.default_rng(5)
    X, Pi, V = _train_data(rng, 8, 18 * 361, 362)
    out = []
    for tc in ("0", "1"):
        monkeypatch.setenv("AZ_TRAIN_TC", tc)
        e = engine_lib.create(desc())
        H.tame_gammas([e], 1, 23, target=0.05)
        out.append(e.train_grads(1, X, Pi, V))
        e.close()
    (g0, c0), (g1, c1) = out
    assert np.isfinite(g1).all()
    assert abs(c0 - c1) <= 1e-5 * max(1, abs(c0))
    scale = np.abs(g0).max()
    assert np.abs(g0 - g1).max() <= 1e-4 * scale, (np.abs(g0 - g1).max(), scale)


def test_train_loop_vs_oracle(oracle, engine_lib):
    """dual.Train (meta.go:16-54): 3 batches x 5 passes with the per-pass row shuffle; costs and weights."""
    nn = dict(k=3, shared_layers=3, fc=8, batch_size=20, features=2, action_space=10)
    def desc():
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=2, n_games=2, seed=1, nn=nn, flags=K.FLAG_FP32_TOWER)
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    H.tame_gammas([eo, eg], 1, 17, target=0.05)
    X, Pi, V = _train_data(np.random.default_rng(4), 60, 18, 10)
    co = eo.train(1, X.copy(), Pi.copy(), V.copy(), 3, 5, lr=0.1, shuffle_seed=99)
    cg = eg.train(1, X.copy(), Pi.copy(), V.copy(), 3, 5, lr=0.1, shuffle_seed=99)
    assert np.isfinite(cg).all()
    assert np.abs(co - cg).max() <= 2e-3 * np.abs(co).max(), (co, cg)
    assert co[-1] < co[0]
    po, pg = eo.net_get(1), eg.net_get(1)
    assert np.abs(po - pg).max() <= 5e-3 * np.abs(po).max()


def test_c1_learn_on_device(oracle, engine_lib):
    """BASELINE config C1 shape on the device through the host mirror: tic-tac-toe AZ.Learn — epoch 0
    self-play (dummy inferer) is bit-exact against the oracle run; the whole loop completes with
    finite costs and a decision per epoch."""
    from agogo_b200 import host
    from tests.test_host_learn import _c1_conf
    conf = _c1_conf(batch=20, sims=20)
    ag = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=engine_lib, n_games=16, seed=42)
    ao = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=16, seed=42)
    for az in (ag, ao):
        az.setupSelfPlay(0)
    xg, xo = ag._play(12, True), ao._play(12, True)
    assert len(xg) == len(xo)
    for a, b in zip(xg, xo):
        assert (a.Board == b.Board).all() and (a.Policy == b.Policy).all() and a.Value == b.Value
    ag2 = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=engine_lib, n_games=16, seed=43)
    ag2.Learn(2, 12, 4, 10)
    assert len(ag2.log) == 2
    for l in ag2.log:
        assert np.isfinite(l["first_cost"]) and np.isfinite(l["last_cost"]) and l["n_examples"] > 0
        assert sum(l["a"]) == 10 and sum(l["b"]) == 10


def test_learn_loop_go_on_tensor_cores(engine_lib):
    """AZ.Learn end to end on a Go net wide enough for every tensor-core path (9x9, K = 64: tcgen05 tower in self-play
    and arena, tcgen05 forward / backward-data / backward-filter in Train) from the reference's own random init: two
    epochs complete, costs finite, every arena game counted, examples produced."""
    from agogo_b200 import host
    nn = host.DefaultConf(9, 9, 82)
    nn.K, nn.SharedLayers, nn.FC, nn.BatchSize, nn.Features = 64, 2, 32, 16, 18
    mc = host.MCTSConfig(PUCT=1.0, M=9, N=9, Timeout=100_000_000, PassPreference=K.DONT_PREFER_PASS, Budget=1000,
                         DumbPass=True, RandomCount=0, Sims=12)
    conf = host.Config(Name="go9", NNConf=nn, MCTSConf=mc, UpdateThreshold=0.55, Encoder=K.ENC_WQ18)
    az = host.AZ(host.Game(K.GAME_WQ, 9, 9, 0, komi=7.5, max_moves=14), conf, lib=engine_lib, n_games=8, seed=7)
    az.Learn(2, 8, 3, 6)
    assert len(az.log) == 2
    for l in az.log:
        assert np.isfinite(l["first_cost"]) and np.isfinite(l["last_cost"]) and l["n_examples"] >= 16
        assert sum(l["a"]) == 6 and sum(l["b"]) == 6


_C3_ORACLE = {}


def _c3_oracle_outputs(oracle, init, seed):
    """Oracle forward of the headline net (20 blocks x 256, 19x19, FC 512) on three positions, cached per weight set."""
    key = (init, seed)
    if key not in _C3_ORACLE:
        d = K.make_desc(K.GAME_WQ, 19, 19, 0, komi=7.5, sims=2, n_games=4, seed=2, max_moves=4,
                        nn=dict(k=256, shared_layers=20, fc=512, batch_size=2, features=18, action_space=362))
        eo = oracle.create(d)
        # "reference": BN scales with the spread the reference's own init has at DefaultConf's batch 256 (GlorotN over
        # [B,C,H,W] = sigma sqrt(2/((256+256)*361)), x 316 in test mode = 1.0396 per layer): the distribution the bench
        # nets have, without materialising 7.9 GB of batch-shaped parameters in the CPU oracle
        params = H.tame_gammas([eo], 0, seed, target=0.9 if init == "tamed" else 1.0396)
        eo.set_inferer(0, K.INF_DUAL)
        planes = _wq_planes(np.random.default_rng(8), 3, 19)
        po, vo = eo.infer(0, planes)
        _C3_ORACLE[key] = (params, planes, po, vo)
        eo.close()
    return _C3_ORACLE[key]


@pytest.mark.parametrize("init,seed", [("tamed", 99), ("tamed", 7), ("reference", 5), ("reference", 11)])
@pytest.mark.parametrize("mode", ["fp16x3-halo", "fp16x3-tap", "f8-halo", "f8-tap"])
def test_tc_tower_full_depth_c3(oracle, engine_lib, monkeypatch, mode, init, seed):
    """The headline net (20 blocks x 256, 19x19, FC 512) end to end against the oracle through all 41 conv layers, on four
    weight sets (two with the BN-scale spread of the reference's own init at batch 256: per-layer gain > 1, the nets
    bench.py runs).  The default precision (three fp16 passes, halo kernel — and the per-tap kernel it replaced) holds
    the north star's 1e-4 on all of them.  AZ_FLAG_FAST_TOWER (FP8 correction passes, ~14.5-bit operands) holds it with
    >= 5x margin on the well-conditioned nets but not on the reference-init ones (measured up to 1.4e-4 on the value
    head): that is why it is opt-in; its bound here is what was measured, stated, not the north star's."""
    env = {"fp16x3-halo": {}, "fp16x3-tap": {"AZ_TC_HALO": "0"}, "f8-halo": {}, "f8-tap": {"AZ_TC_FP8": "1"}}[mode]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fast = mode.startswith("f8")
    params, planes, po, vo = _c3_oracle_outputs(oracle, init, seed)
    d = K.make_desc(K.GAME_WQ, 19, 19, 0, komi=7.5, sims=2, n_games=4, seed=2, max_moves=4, flags=K.FLAG_FAST_TOWER if fast else 0,
                    nn=dict(k=256, shared_layers=20, fc=512, batch_size=2, features=18, action_space=362))
    etc = engine_lib.create(d)
    etc.net_set(0, params)
    etc.set_inferer(0, K.INF_DUAL)
    ptc, vtc = etc.infer(0, planes)
    etc.close()
    dp, dv = np.abs(ptc - po).max(), np.abs(vtc - vo).max()
    print("C3 depth %s %s/%d: dp=%.3g dv=%.3g pmax=%.3g |v|max=%.3g" % (mode, init, seed, dp, dv, po.max(), np.abs(vo).max()))
    assert np.isfinite(po).all() and np.isfinite(vo).all()
    tol = 1e-4 if not fast else (2e-5 if init == "tamed" else 3e-4)
    assert dp < tol and dv < tol, (mode, init, seed, dp, dv, tol)


@pytest.mark.parametrize("mode", ["fp16x3-halo", "fp16x3-tap", "f8-halo"])
def test_tc_tower_modes_small_nets(oracle, engine_lib, monkeypatch, mode):
    """9x9, K = 128 (2C = 256: the CTA-pair kernels in the flat 2-D layout, halo of 150 -> 160 rows): the three-pass halo
    kernel (default), the per-tap kernel and the FP8-correction halo kernel against the oracle."""
    if mode == "fp16x3-tap":
        monkeypatch.setenv("AZ_TC_HALO", "0")
    def desc():
        return K.make_desc(K.GAME_WQ, 9, 9, 0, komi=7.5, sims=2, n_games=8, seed=2, max_moves=4,
                           flags=K.FLAG_FAST_TOWER if mode == "f8-halo" else 0,
                           nn=dict(k=128, shared_layers=4, fc=64, batch_size=2, features=18, action_space=82))
    eo, etc = oracle.create(desc()), engine_lib.create(desc())
    H.tame_gammas([eo, etc], 0, 13)
    for e in (eo, etc):
        e.set_inferer(0, K.INF_DUAL)
    planes = _wq_planes(np.random.default_rng(6), 11, 9)
    (po, vo), (ptc, vtc) = eo.infer(0, planes), etc.infer(0, planes)
    print("9x9 K=128 mode %s: dp=%.3g dv=%.3g" % (mode, np.abs(ptc - po).max(), np.abs(vtc - vo).max()))
    assert np.abs(ptc - po).max() < 1e-4 and np.abs(vtc - vo).max() < 1e-4


def test_c2_shapes_selfplay(oracle, engine_lib):
    """BASELINE config C2 shapes (9x9 wq, 6-block x 64 net, FC 128, WQEncoder) at reduced game/sim counts:
    tensor-core tower in the loop, move sequences and evaluation counts equal to the oracle's."""
    def desc():
        return K.make_desc(K.GAME_WQ, 9, 9, 0, komi=7.5, sims=32, n_games=8, seed=77, max_moves=5,
                           nn=dict(k=64, shared_layers=6, fc=128, batch_size=2, features=18, action_space=82))
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    H.tame_gammas([eo, eg], 0, 41); H.tame_gammas([eo, eg], 1, 42)
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUAL); e.set_inferer(1, K.INF_DUAL)
    a, b = H.play_and_collect(eo, 8, dump_trees=False), H.play_and_collect(eg, 8, dump_trees=False)
    for ra, rb in zip(a["records"], b["records"]):
        assert list(ra["moves"]) == list(rb["moves"])
    for k in ("sims", "evals", "null_results", "created"):
        assert a["counters"][k] == b["counters"][k], k
    xa, xb = a["examples"], b["examples"]
    assert (xa[0].view(np.uint32) == xb[0].view(np.uint32)).all()  # WQEncoder planes incl. -0.0


def test_search_external_positions(oracle, engine_lib):
    """az_search (Agent.Search on a caller-supplied game.State): best move and root visit counts equal
    the oracle's for mnk, c4 and wq positions (history planes included for wq)."""
    rng = np.random.default_rng(11)
    cases = [
        (lambda: K.make_desc(K.GAME_MNK, 3, 3, 3, sims=60, nn=H.tiny_nn(3, 3, 10), n_games=2, seed=1),
         dict(board=[1, 1, 0, 2, 2, 0, 0, 0, 0], to_move=K.BLACK, player=K.BLACK, move_number=4)),
        (lambda: K.make_desc(K.GAME_C4, 6, 7, 4, sims=50, nn=H.tiny_nn(6, 7, 8), n_games=2, seed=1),
         dict(board=[0] * 35 + [1, 2, 1, 2, 0, 0, 0], to_move=K.WHITE, player=K.WHITE, move_number=1)),
    ]
    for mk, st in cases:
        eo, eg = oracle.create(mk()), engine_lib.create(mk())
        for e in (eo, eg):
            e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        for agent in (0, 1):
            bo, vo = eo.search(agent, st["board"], st["to_move"], st["player"], st["move_number"])
            bg, vg = eg.search(agent, st["board"], st["to_move"], st["player"], st["move_number"])
            assert bo == bg and (vo == vg).all(), (st, agent, bo, bg)
    # wq with a dual net (fp32 tower) and real history: positions produced by a short oracle game
    def wdesc():
        return K.make_desc(K.GAME_WQ, 7, 7, 0, komi=7.5, sims=30, n_games=2, seed=3, max_moves=12, flags=K.FLAG_FP32_TOWER,
                           nn=dict(k=4, shared_layers=1, fc=8, batch_size=4, features=18, action_space=50))
    eo, eg = oracle.create(wdesc()), engine_lib.create(wdesc())
    H.tame_gammas([eo, eg], 0, 5); H.tame_gammas([eo, eg], 1, 6)
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUAL); e.set_inferer(1, K.INF_DUAL)
    hist, board = [], np.zeros(49, np.int32)
    for mv, col in zip(rng.permutation(49)[:9], [1, 2] * 5):
        hist.append(board.copy())
        board = board.copy(); board[mv] = col
    bo, vo = eo.search(0, board, K.WHITE, K.WHITE, move_number=9, hist=np.array(hist[-8:]))
    bg, vg = eg.search(0, board, K.WHITE, K.WHITE, move_number=9, hist=np.array(hist[-8:]))
    assert bo == bg and np.abs(vo - vg).max() <= 1, (bo, bg, vo, vg)


def test_engine_error_paths(engine_lib):
    """Loud failures instead of silent fallbacks: invalid configs (the reference panics, agogo.go:42-47),
    unsupported search options, a tree pool that is too small, calls out of sequence."""
    with pytest.raises(K.AZError):
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=5, nn=H.tiny_nn(3, 3, 10)); d.mcts.puct = 0.0
        engine_lib.create(d)
    with pytest.raises(K.AZError):
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=5, nn=H.tiny_nn(3, 3, 10)); d.nn.action_space = 2
        engine_lib.create(d)
    with pytest.raises(K.AZError):
        d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=5, nn=H.tiny_nn(3, 3, 10)); d.mcts.random_count = 3
        engine_lib.create(d)
    d = K.make_desc(K.GAME_MNK, 3, 3, 3, sims=30, nn=H.tiny_nn(3, 3, 10), n_games=2)
    d.max_nodes_per_tree = 12
    e = engine_lib.create(d)
    with pytest.raises(K.AZError):
        e.arena_begin(2, False)          # no inferer yet
    e.set_inferer(0, K.INF_DUMMY, 0); e.set_inferer(1, K.INF_DUMMY, 0)
    with pytest.raises(K.AZError):
        e.search_run(1)                  # no arena running
    e.arena_begin(2, False)
    with pytest.raises(K.AZError) as ei:
        for _ in range(9):
            e.arena_step()
    assert "pool" in str(ei.value)
    # empty inference batch
    e2 = engine_lib.create(K.make_desc(K.GAME_MNK, 3, 3, 3, sims=5, nn=H.tiny_nn(3, 3, 10), n_games=2))
    e2.net_init(0, 1); e2.set_inferer(0, K.INF_DUAL)
    p, v = e2.infer(0, np.zeros((0, 18), np.float32))
    assert p.shape == (0, 10) and v.shape == (0,)


def test_graph_replay_equals_plain_launches(engine_lib):
    """The CUDA-graph replay of a wave and the plain launch sequence produce identical trees."""
    import os
    def run(no_graph):
        os.environ["AZ_NO_GRAPH"] = "1" if no_graph else "0"
        e = engine_lib.create(K.make_desc(K.GAME_C4, 6, 7, 4, sims=40, nn=H.tiny_nn(6, 7, 8), n_games=8, seed=5))
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        return H.play_and_collect(e, 8)
    try:
        H.assert_same_run(run(True), run(False), "graph")
    finally:
        os.environ.pop("AZ_NO_GRAPH", None)


def _check_tree_invariants(rows, sims_done_total):
    """Size-independent properties of a search tree dump (rows: depth, move, N, Wbits, Pbits, expanded, nchild)."""
    depth, N, nchild, expanded = rows[:, 0], rows[:, 2].astype(np.int64), rows[:, 6], rows[:, 5]
    P = rows[:, 4].copy().view(np.float32)
    W = rows[:, 3].copy().view(np.float32)
    assert np.isfinite(W).all() and np.isfinite(P).all()
    # walk the preorder dump: children of row i are the next rows at depth+1 until depth drops back
    stack = []
    kids = {}
    for i, d in enumerate(depth):
        while stack and depth[stack[-1]] >= d:
            stack.pop()
        if stack:
            kids.setdefault(stack[-1], []).append(i)
        stack.append(i)
    for i in range(len(rows)):
        ch = kids.get(i, [])
        assert len(ch) == nchild[i]
        if not expanded[i]:
            assert nchild[i] == 0
            continue
        s = float(P[ch].sum())  # priors of an expansion are normalised over the legal moves
        assert abs(s - 1.0) < 1e-3, s
        # visits: born with 1, +1 for its own expansion, +1 per simulation that went on into a child
        below = int(sum(N[c] - 1 for c in ch))
        assert N[i] - 2 == below, (i, N[i], below)
    assert N[0] - 2 <= sims_done_total


def test_full_size_properties_c3(engine_lib):
    """BASELINE config C3 at full width (19x19, 1024 concurrent games, 20x256 tensor-core net), reduced
    sims: properties that do not need the oracle — visit conservation in every sampled tree, priors
    normalised, evaluation count = searches + non-null simulations, bit-identical repeat run."""
    sims = 24
    def run():
        d = K.make_desc(K.GAME_WQ, 19, 19, 0, komi=7.5, sims=sims, n_games=1024, seed=5, max_moves=4,
                        nn=dict(k=256, shared_layers=20, fc=512, batch_size=2, features=18, action_space=362))
        e = engine_lib.create(d)
        H.tame_gammas([e], 0, 3); H.tame_gammas([e], 1, 4)
        e.set_inferer(0, K.INF_DUAL); e.set_inferer(1, K.INF_DUAL)
        e.arena_begin(1024, True)
        e.arena_step()
        trees = {g: [e.tree_dump(g, t) for t in (0, 1)] for g in (0, 1, 511, 1023)}
        recs = [e.game_record(g) for g in range(0, 1024, 97)]
        c = e.counters()
        e.arena_step()
        c2 = e.counters()
        e.arena_finish()
        ex = e.examples(clear=True)
        return trees, recs, c, c2, ex
    trees, recs, c, c2, ex = run()
    assert c["searches"] == 1024 and c["sims"] == 1024 * sims
    assert c["evals"] == c["searches"] + c["sims"] - c["null_results"]
    assert c2["searches"] == 2048
    for g, (ta, tb) in trees.items():
        searched = ta if len(ta) else tb
        assert (len(ta) == 0) != (len(tb) == 0)  # exactly one agent has moved in this game
        _check_tree_invariants(searched, sims)
    boards, pols, vals = ex
    assert boards.shape == (2048, 18 * 361) and np.isin(boards, [0.0, 1.0, -1.0]).all()
    assert (np.abs(pols.sum(axis=1) - 1) < 1e-6).all() and set(np.unique(vals)) <= {-1.0, 0.0, 1.0}
    trees2, recs2, c_b, _, ex2 = run()
    assert c_b == c
    for r1, r2 in zip(recs, recs2):
        assert list(r1["moves"]) == list(r2["moves"])
    assert (ex[0].view(np.uint32) == ex2[0].view(np.uint32)).all()
    for g in trees:
        for t in (0, 1):
            assert (trees[g][t] == trees2[g][t]).all()


@pytest.mark.parametrize("kind,m,n,k,dumb,pref", [
    (K.GAME_C4, 6, 7, 4, 0, K.DONT_PREFER_PASS), (K.GAME_C4, 6, 7, 4, 0, K.PREFER_PASS), (K.GAME_C4, 6, 7, 4, 1, K.DONT_RESIGN),
    (K.GAME_WQ, 5, 5, 0, 0, K.PREFER_PASS), (K.GAME_WQ, 5, 5, 0, 0, K.DONT_PREFER_PASS), (K.GAME_MNK, 3, 3, 3, 0, K.PREFER_PASS),
])
def test_parity_pass_preferences(oracle, engine_lib, kind, m, n, k, dumb, pref):
    """bestMove's pass heuristics (search.go:366-389): DumbPass=false and every PassPreference."""
    A = n if kind == K.GAME_C4 else m * n
    def desc():
        return K.make_desc(kind, m, n, k, komi=7.5, sims=20, n_games=6, seed=15, max_moves=30, dumb_pass=dumb,
                           pass_preference=pref, nn=H.tiny_nn(m, n, A + 1, features=18 if kind == K.GAME_WQ else 2))
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    H.assert_same_run(H.play_and_collect(eo, 6), H.play_and_collect(eg, 6), "pass-pref")


@pytest.mark.parametrize("kind,m,n,k,temp,minv", [(K.GAME_MNK, 3, 3, 3, 1.0, 0), (K.GAME_C4, 6, 7, 4, 0.5, 1), (K.GAME_MNK, 5, 5, 4, 2.0, 1)])
def test_parity_temperature_sampling(oracle, engine_lib, kind, m, n, k, temp, minv):
    """RandomCount > 0: randomizeChildren (tree.go:212-247) — (visits/maxVisits)^(1/T) sampling and the
    reference's swap loop, with the injected tree RNG."""
    A = n if kind == K.GAME_C4 else m * n
    def desc():
        d = K.make_desc(kind, m, n, k, sims=24, n_games=8, seed=23, max_moves=30, nn=H.tiny_nn(m, n, A + 1))
        d.mcts.random_count, d.mcts.random_temperature, d.mcts.random_min_visits = 6, temp, minv
        return d
    eo, eg = oracle.create(desc()), engine_lib.create(desc())
    for e in (eo, eg):
        e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
    a, b = H.play_and_collect(eo, 8), H.play_and_collect(eg, 8)
    H.assert_same_run(a, b, "temperature")
    assert len({tuple(r["moves"][:4]) for r in a["records"]}) >= 1


def test_arena_play_chunking_and_single_slot(oracle, engine_lib):
    """az_arena_play with more games than device slots runs them in chunks that continue the same coin
    stream; one-slot engines work: records, examples (game order) and statistics equal an all-at-once run."""
    def desc(g):
        return K.make_desc(K.GAME_MNK, 3, 3, 3, sims=15, nn=H.tiny_nn(3, 3, 10), n_games=g, seed=19)
    runs = []
    for lib, slots in ((oracle, 16), (engine_lib, 4), (engine_lib, 1)):
        e = lib.create(desc(slots))
        e.set_inferer(0, K.INF_DUMMY, 0); e.set_inferer(1, K.INF_DUMMY, 0)
        e.arena_play(10, True)
        recs = [e.game_record(g) for g in range(10)]
        runs.append((recs, e.examples(clear=True), e.stats(0), e.stats(1)))
    for other in runs[1:]:
        for ra, rb in zip(runs[0][0], other[0]):
            assert list(ra["moves"]) == list(rb["moves"]) and ra["winner"] == rb["winner"] and ra["a_player"] == rb["a_player"]
        for xa, xb in zip(runs[0][1], other[1]):
            assert xa.shape == xb.shape and (xa.view(np.uint32) == xb.view(np.uint32)).all()
        assert runs[0][2:] == other[2:]


@pytest.mark.parametrize("size,inferer,workers,sims,plies,temp", [
    (19, "dummy", 1, 96, 6, 0.0),     # all 362 priors equal: every Select is decided by the first-strict-max tie-break
    (19, "table", 1, 160, 8, 0.0),    # distinct priors per depth: rank sort over 362 entries, 3 x 128-wide Select strides
    (19, "ties", 1, 96, 6, 0.0),      # priors quantised to 8 levels: the stable rank sort must keep index order in ties
    (19, "table", 16, 160, 6, 0.0),   # the fixed worker schedule on 362-child nodes (virtual-loss flags, short last round)
    (19, "table", 1, 64, 10, 1.0),    # temperature sampling from the per-tree RNG spreads the 8 games over different lines
    (13, "table", 1, 128, 8, 0.0),    # 170 children: two strides, the second partial
    (13, "dummy", 16, 64, 6, 0.0),
])
def test_parity_wq_full_board_search(oracle, engine_lib, size, inferer, workers, sims, plies, temp):
    """The headline board size against the oracle, bit for bit: nodes with more than 128
    children take the multi-stride path of Node.Select (node.go:170-237) and the >128-entry rank sort of
    expandAndSimulate (search.go:314-330) on every simulation of BASELINE config C3; trees after every ply, moves,
    examples, statistics and counters equal the oracle's."""
    A1 = size * size + 1
    def desc():
        d = K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=sims, n_games=8, seed=31, max_moves=plies, workers=workers,
                        nn=H.tiny_nn(size, size, A1, features=18))
        if temp:
            d.mcts.random_count, d.mcts.random_temperature, d.mcts.random_min_visits = plies, temp, 0
        return d
    eo, eg = _pair(oracle, engine_lib, desc)
    rng = np.random.default_rng(size * 100 + sims)
    table = rng.random((96, A1)).astype(np.float32)
    if inferer == "ties":
        table = np.ceil(table * 8).astype(np.float32)
    table /= table.sum(axis=1, keepdims=True)
    values = rng.uniform(0.05, 0.95, 96).astype(np.float32)
    for e in (eo, eg):
        if inferer == "dummy":
            e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        else:
            e.set_table(0, table, values); e.set_table(1, table[::-1].copy(), values[::-1].copy())
    a, b = H.play_and_collect(eo, 8), H.play_and_collect(eg, 8)
    H.assert_same_run(a, b, "wq%d-%s-w%d" % (size, inferer, workers))
    c = a["counters"]
    assert c["sims"] == c["searches"] * sims
    assert c["select_children"] > 128 * c["select_levels"]  # the strided path really ran
    if temp:
        assert len({tuple(r["moves"]) for r in a["records"]}) >= 4


def test_wq_complete_rules_vs_oracle(oracle, engine_lib):
    """AZ_FLAG_WQ_COMPLETE (our completion of the Go rules): legality / captures of every point and area scores of random
    7x7 positions, engine against oracle (the oracle is pinned against the naive Python restatement on CPU)."""
    rng = np.random.default_rng(4)
    def eng(lib):
        d = K.make_desc(K.GAME_WQ, 7, 7, 0, komi=5.5, sims=1, n_games=1, flags=K.FLAG_WQ_COMPLETE,
                        nn=H.tiny_nn(7, 7, 50, features=2), encoder=K.ENC_TWO_PLANE)
        return lib.create(d)
    eo, eg = eng(oracle), eng(engine_lib)
    boards, players, moves = [], [], []
    for _ in range(50):
        b = rng.choice([0, 1, 2], size=49, p=[0.35, 0.33, 0.32]).astype(np.int32)
        for mv in range(49):
            boards.append(b); players.append(1 + (mv + len(boards)) % 2); moves.append(mv)
    boards = np.array(boards, np.int32)
    for x, y, name in zip(eo.rules_apply(boards, players, moves), eg.rules_apply(boards, players, moves), ("check", "applied", "boards", "taken")):
        assert (x == y).all(), name
    for x, y in zip(eo.rules_status(boards[::49], passes=[2] * 50), eg.rules_status(boards[::49], passes=[2] * 50)):
        assert (x == y).all()


@pytest.mark.parametrize("size,sims,plies,inferer,workers", [(5, 20, 90, "dummy", 1), (5, 24, 90, "table", 1), (9, 16, 60, "table", 1), (5, 24, 60, "table", 4),
                                                            (4, 200, 60, "table", 1), (3, 150, 40, "table", 1), (4, 120, 60, "table", 4)])
def test_parity_wq_complete_arena(oracle, engine_lib, size, sims, plies, inferer, workers):
    """Whole games under AZ_FLAG_WQ_COMPLETE with sampled play (captures, kos, eye-only endgames, two-pass endings scored
    by area + komi): trees after every ply, moves, examples, labels and statistics bit-identical to the oracle.  The tiny
    boards with many simulations are the positional-superko cases: the oracle rejects > 100 moves there that simple ko
    allows, most of them deep in the tree (positions of the descent itself), where the device compares 64-bit position
    hashes and the oracle whole boards."""
    A1 = size * size + 1
    def desc():
        d = K.make_desc(K.GAME_WQ, size, size, 0, komi=5.5, sims=sims, n_games=8, seed=17, max_moves=plies, workers=workers,
                        flags=K.FLAG_WQ_COMPLETE, nn=H.tiny_nn(size, size, A1, features=18))
        d.mcts.random_count, d.mcts.random_temperature = plies, 1.0
        return d
    eo, eg = _pair(oracle, engine_lib, desc)
    rng = np.random.default_rng(size)
    table = rng.random((128, A1)).astype(np.float32)
    table /= table.sum(axis=1, keepdims=True)
    values = rng.uniform(0.05, 0.95, 128).astype(np.float32)
    for e in (eo, eg):
        if inferer == "dummy":
            e.set_inferer(0, K.INF_DUMMY, 1); e.set_inferer(1, K.INF_DUMMY, 2)
        else:
            e.set_table(0, table, values); e.set_table(1, table[::-1].copy(), values[::-1].copy())
    a, b = H.play_and_collect(eo, 8), H.play_and_collect(eg, 8)
    H.assert_same_run(a, b, "wq-complete-%d-%s" % (size, inferer))
    assert len({tuple(r["moves"]) for r in a["records"]}) >= 4
    assert any(r["winner"] != 0 for r in a["records"]) or size == 9


def test_wq_complete_game_engine_vs_pyref(engine_lib):
    from tests.test_oracle_rules_pyref import _complete_game_vs_pyref
    kos = caps = 0
    for seed in range(1, 6):
        k, c, _, _ = _complete_game_vs_pyref(engine_lib, 5, 20, seed, 90)
        kos += k; caps += c
    assert caps > 10


def test_wq_superko_external_engine_vs_pyref(engine_lib):
    """Positional superko on caller-owned positions (az_state's earlier boards, up to 14 here): the device's legal set against the
    Python restatement, 160 positions with planted repetitions (capturing and non-capturing)."""
    from tests.test_oracle_rules_pyref import _superko_external_vs_pyref
    sk = sum(_superko_external_vs_pyref(engine_lib, size, seed, 40) for size, seed in ((3, 1), (5, 2), (7, 3), (9, 4)))
    assert sk >= 100


@pytest.mark.parametrize("size,layers,n", [(9, 6, 37), (9, 1, 300), (5, 3, 5), (10, 2, 9)])
def test_small_net_kernel_equals_layered_path(oracle, engine_lib, monkeypatch, size, layers, n):
    """k_net_small (K = 64: the whole network of a leaf in one CTA, activations resident in shared memory) against the
    per-layer kernels it replaces and against the oracle.  Same MMA sequence, same epilogue arithmetic, same head
    summation order: the two device paths agree to the last bits; n > 2 x SMs exercises the persistent sample loop."""
    A1 = size * size + 1
    def desc():
        return K.make_desc(K.GAME_WQ, size, size, 0, komi=7.5, sims=2, n_games=8, seed=2, max_moves=4,
                           nn=dict(k=64, shared_layers=layers, fc=128, batch_size=2, features=18, action_space=A1))
    planes = _wq_planes(np.random.default_rng(3), n, size)
    eo = oracle.create(desc())
    params = H.tame_gammas([eo], 0, 21)
    eo.set_inferer(0, K.INF_DUAL)
    po, vo = eo.infer(0, planes[:16])
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("AZ_TC_SMALLNET", mode)
        e = engine_lib.create(desc())
        e.net_set(0, params)
        e.set_inferer(0, K.INF_DUAL)
        outs[mode] = e.infer(0, planes)
        e.close()
    (p1, v1), (p0, v0) = outs["1"], outs["0"]
    print("small vs layered: dp=%.3g dv=%.3g ; small vs oracle: dp=%.3g dv=%.3g" %
          (np.abs(p1 - p0).max(), np.abs(v1 - v0).max(), np.abs(p1[:16] - po).max(), np.abs(v1[:16] - vo).max()))
    assert np.abs(p1 - p0).max() <= 1e-7 and np.abs(v1 - v0).max() <= 1e-6
    assert np.abs(p1[:16] - po).max() < 1e-4 and np.abs(v1[:16] - vo).max() < 1e-4
