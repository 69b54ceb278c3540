"""Host-side mirror of gorgonia/agogo's API layer for the self-play path, over the C ABI.

The reference's host is Go (agogo.go, arena.go, agent.go); this image has no Go toolchain, so the same
composition — same names, argument meaning and error behaviour — is provided here in Python over
`ctypes` (the cgo equivalent is go/agogo_b200.go + INTEGRATION.md).  Everything numeric happens inside
the engine library; this file only sequences calls the way `AZ.Learn` (agogo.go:100-172) does and
replaces the reference's time-seeded RNGs by the injected splitmix64 streams of DESIGN.md §2.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _capi as K

_M64 = (1 << 64) - 1


def _splitmix(state):
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def derive_seed(seed, stream):
    s = (seed ^ ((0xD1B54A32D192ED03 * (stream + 1)) & _M64)) & _M64
    return _splitmix(s)[1]


class Rng:
    def __init__(self, seed):
        self.s = seed & _M64

    def intn(self, n):
        self.s, v = _splitmix(self.s)
        return v % n


@dataclass
class DualConfig:
    """dual.Config (dualnet/config.go:4-16)."""
    K: int = 0
    SharedLayers: int = 0
    FC: int = 0
    L2: float = 0.0
    BatchSize: int = 0
    Width: int = 0
    Height: int = 0
    Features: int = 0
    ActionSpace: int = 0
    FwdOnly: bool = False

    def IsValid(self):  # config.go:33-42
        return (self.K >= 1 and self.ActionSpace >= 3 and self.SharedLayers >= 0 and self.FC > 1 and
                self.BatchSize >= 1 and self.Features > 0)


def dual_round(a):  # config.go:44-59
    n = a - 1
    for sh in (1, 2, 4, 8, 16):
        n |= n >> sh
    n += 1
    lt = n // 2
    return lt if (a - lt) < (n - a) else n


def DefaultConf(m, n, actionSpace):  # config.go:18-31
    k = dual_round((m * n) // 3)
    return DualConfig(K=k, SharedLayers=m, FC=2 * k, BatchSize=256, Width=n, Height=m, Features=18,
                      ActionSpace=actionSpace)


@dataclass
class MCTSConfig:
    """mcts.Config (mcts/tree.go:15-29) + Sims (the fixed-iteration mode the reference lacks)."""
    PUCT: float = 1.0
    Timeout: int = 0
    M: int = 0
    N: int = 0
    RandomCount: int = 0
    Budget: int = 0
    RandomMinVisits: int = 0
    RandomTemperature: float = 0.0
    DumbPass: bool = True
    ResignPercentage: float = 0.0
    PassPreference: int = K.DONT_PREFER_PASS
    Sims: int = 100
    Workers: int = 0  # concurrent pipeline calls per tree (the reference: runtime.NumCPU()); 0/1 = canonical single worker

    def IsValid(self):  # tree.go:43-45
        return 0 < self.PUCT <= 1


def DefaultConfig(boardSize):  # mcts/tree.go:31-41
    return MCTSConfig(PUCT=1.0, Timeout=100_000_000, M=boardSize, N=boardSize, DumbPass=True,
                      PassPreference=K.DONT_PREFER_PASS, Budget=10000)


@dataclass
class Game:
    """Which game.State the engine instantiates (mnk.New / c4.New / wq.New)."""
    kind: int
    m: int
    n: int
    k: int = 0
    komi: float = 0.0
    max_moves: int = 0
    zobrist_seed: int = 12345


@dataclass
class Config:
    """agogo.Config (datatypes.go:14-25).  Encoder is an az_encoder_kind; Augmenter stays a Python callable."""
    Name: str = ""
    NNConf: DualConfig = field(default_factory=DualConfig)
    MCTSConf: MCTSConfig = field(default_factory=MCTSConfig)
    UpdateThreshold: float = 0.0
    MaxExamples: int = 0
    Encoder: int = K.ENC_TWO_PLANE
    Augmenter: object = None       # Augmenter func(Example) []Example (datatypes.go:44-45), applied to every kept example
    OutputEncoder: object = None   # OutputEncoder{Encode(MetaState) error; Flush() error} (datatypes.go:36-42)


@dataclass
class Example:  # datatypes.go:38-42
    Board: np.ndarray
    Policy: np.ndarray
    Value: float


@dataclass
class State:
    """The game.State getters (game/state.go:128-169) the path reads from a caller-owned position — what `Agent.Search`,
    `Agent.Infer` and the encoders take.  `hist` holds the boards before the current one, oldest first (<= 8; any number under AZ_FLAG_WQ_COMPLETE)
    (`Historical(MoveNumber()-len(hist)) ... Historical(MoveNumber()-1)`); `moves` the tail of the history as
    (player, move) pairs, oldest first (lets the agent's tree be re-rooted across calls, search.go:424-500)."""
    board: np.ndarray
    to_move: int = K.BLACK
    move_number: int = 0
    passes: int = 0
    last_move: int = K.PASS
    hist: object = None
    moves: object = None
    ko: int = -1    # AZ_FLAG_WQ_COMPLETE only

    def Board(self):
        return np.asarray(self.board, np.int32).reshape(-1)

    def ToMove(self):
        return self.to_move

    def MoveNumber(self):
        return self.move_number

    def SetToMove(self, p):
        self.to_move = p

    def Clone(self):
        import copy
        return copy.deepcopy(self)

    def Eq(self, other):  # wq/game.go:123-160 / mnk.go:191-206: side to move, counters, board, then the known history
        def arr(x, w):
            return np.zeros((0, w), np.int32) if x is None else np.asarray(x, np.int32).reshape(-1, w)
        return (isinstance(other, State) and self.to_move == other.to_move and self.move_number == other.move_number
                and self.passes == other.passes and np.array_equal(self.Board(), other.Board())
                and np.array_equal(arr(self.moves, 2), arr(other.moves, 2)))

    def Historical(self, i):  # wq: the board before move i; the reference panics on a bad index too
        h = [] if self.hist is None else np.asarray(self.hist, np.int32).reshape(-1, self.Board().size)
        j = i - (self.move_number - len(h))
        if not 0 <= j < len(h):
            raise IndexError("index out of range")
        return h[j]


def EncodeTwoPlayerBoard(a, prealloc=None):
    """encoding_helper.go:10-26: black 1, white -1, everything else 0."""
    a = np.asarray(a)
    out = prealloc if prealloc is not None and len(prealloc) == len(a) else np.zeros(len(a), np.float32)
    out[:] = np.where(a == K.BLACK, np.float32(1), np.where(a == K.WHITE, np.float32(-1), np.float32(0)))
    return out


def EncodeBoard(state):
    """cmd/tictactoe/main.go:26-47 (the two-plane GameEncoder of the mnk / c4 commands): stones +-1 with empties at 0.001,
    then a plane of +-1 for the side to move (zeros when nobody is)."""
    board = EncodeTwoPlayerBoard(state.Board())
    board[board == 0] = np.float32(0.001)
    layer = np.zeros(len(board), np.float32)
    if state.ToMove() == K.BLACK:
        layer[:] = 1
    elif state.ToMove() == K.WHITE:
        layer[:] = -1
    return np.concatenate([board, layer])


def WQEncoder(state):
    """encoding_helper.go:29-68, quirks included: of the 2 x 8 history planes only i = 1..7 are ever written (plane 7 of
    each group and the CURRENT board never are), from Historical((MoveNumber()-1)-i) when that index is > 0; each plane
    carries BOTH colours (+1 black / -1 white), the "white" group being the negation — so its empty points are -0.0;
    the group of the side to move comes first and its to-move plane is filled with +-1."""
    lookback = 8
    board = state.Board()
    size = len(board)
    out = np.zeros(size * (2 * lookback + 2), np.float32)
    if state.ToMove() == K.BLACK:
        bs, ws, ns, player = 0, lookback * size, 2 * lookback * size, np.float32(1)
    else:
        bs, ws, ns, player = lookback * size, 0, (2 * lookback + 1) * size, np.float32(-1)
    current = state.MoveNumber() - 1
    for i in range(1, lookback):
        h = current - i
        if 0 < h < current:
            past = state.Historical(h)
            EncodeTwoPlayerBoard(past, out[bs:bs + size])                              # encodeBlack
            out[ws:ws + size] = EncodeTwoPlayerBoard(past) * np.float32(-1)            # encodeWhite: vecf32.Scale(-1)
        bs += size
        ws += size
    out[ns:ns + size] = player
    return out


def RotateBoard(board, m, n):
    """encoding_helper.go:80-108 (the building block of the commands' Augmenters): a quarter turn, counter-clockwise, of a
    square board given as a flat row-major slice — new[i][j] = old[j][m-1-i]; four of them are the identity
    (encoding_helper_test.go:10-58).  Returns a new array; non-square boards are an error, as in the reference."""
    if m != n:
        raise ValueError("Cannot handle m %d, n %d. This function only takes square boards" % (m, n))
    return np.rot90(np.asarray(board, np.float32).reshape(m, n), 1).reshape(-1).copy()


class Agent:
    """agent.go:14-121 — a view on one of the engine's two agents."""

    def __init__(self, az, idx):
        self._az, self.idx, self.Player = az, idx, K.NONE

    @property
    def Wins(self):
        return self._az.engine.stats(self.idx)[0]

    @property
    def Loss(self):
        return self._az.engine.stats(self.idx)[1]

    @property
    def Draw(self):
        return self._az.engine.stats(self.idx)[2]

    def SwitchToInference(self):  # agent.go:42-57
        self._az.engine.set_inferer(self.idx, K.INF_DUAL)

    def useDummy(self):  # agent.go:105-113: captures the agent's current colour
        self._az.engine.set_inferer(self.idx, K.INF_DUMMY, self.Player)

    def NNOutput(self, planes):  # agent.go:83-89
        return self._az.engine.infer(self.idx, planes)

    def Infer(self, g):  # agent.go:60-74: Enc(g) -> inferer.Infer; g is a host.State
        enc = WQEncoder if self._az.conf.Encoder == K.ENC_WQ18 else EncodeBoard
        policy, value = self._az.engine.infer(self.idx, enc(g)[None])
        return policy[0], float(value[0])

    def Search(self, g):  # agent.go:77-80: MCTS.Search(g.ToMove()) on a caller-owned position -> game.Single
        best, _ = self._az.engine.search(self.idx, g.Board(), g.to_move, g.to_move, move_number=g.move_number, passes=g.passes,
                                         hist=g.hist, last_move=g.last_move, moves=g.moves, ko=g.ko)
        return best

    def Close(self):  # agent.go:91-103: drops the inferer and the tree; the engine owns both until AZ.Close
        self._az.engine.reset_tree(self.idx)

    def resetStats(self):
        self._az.engine.reset_stats(self.idx)


class MetaState:
    """game.MetaState (game/state.go:171-177) as the OutputEncoder sees it after every move of Arena.Play."""

    def __init__(self, az, state, game_number):
        self._az, self._state, self._game_number = az, state, game_number

    def Name(self):
        return self._az.Name()

    def Epoch(self):
        return self._az.epoch

    def GameNumber(self):
        return self._game_number

    def Score(self, player):  # arena.go:191: float64(a.game.Score(p)) of the position the encoder is shown
        return self._az._game_score(self._state, player)

    def State(self):  # game.State getters of the running game: board, to_move, move_number, passes, ended, winner
        return self._state


def shuffle_rows(Xs, Pi, V, rng):
    """shuffleBatch (meta.go:57-102): Fisher-Yates over rows with j = r.Intn(i+1) (same stream as az_train)."""
    for i in range(len(V)):
        j = rng.intn(i + 1)
        if i != j:
            Xs[[i, j]] = Xs[[j, i]]
            Pi[[i, j]] = Pi[[j, i]]
            V[[i, j]] = V[[j, i]]


class Arena:
    """agogo.Arena (arena.go:18-233): the two agents, the game being played and the training bookkeeping; it fulfils
    game.MetaState (Name / Epoch / GameNumber / Score / State), which is what an OutputEncoder is handed after every move.
    Embedded in AZ as in the reference (agogo.go:22); MakeArena's two Dualers are the engine's nets 0 and 1."""
    engine = None
    epoch = 0        # training epoch (arena.go:33)
    gameNumber = 0   # which game of the evaluation arena this is (arena.go:34, agogo.go:144)
    name = "UNKNOWN GAME"  # arena.go:56-58
    _last_state = None
    oldCount = 0

    def Epoch(self):  # arena.go:182
        return self.epoch

    def GameNumber(self):  # arena.go:185
        return self.gameNumber

    def Name(self):  # arena.go:188
        return self.name

    def _game_score(self, st, player):
        """game.State.Score(p) of a position: mnk 1 / -2 / 0 (mnk.go:147-155), c4 1 / -1 / 0 (c4/game.go:75-84), wq as
        implemented by Board.Score (the Game method panics in the reference) or the area score under the complete rules"""
        _, _, sb, sw = self.engine.rules_status(np.asarray(st["board"], np.int32)[None], passes=[max(int(st.get("passes", 0)), 0)])
        return float(sb[0]) if player == K.BLACK else (float(sw[0]) if player == K.WHITE else 0.0)

    def Score(self, p):  # arena.go:191: the running game's Score(p)
        if self._last_state is None:
            raise RuntimeError("no game has been played")
        return self._game_score(self._last_state, p)

    def State(self):  # arena.go:194
        return self._last_state

    def Log(self, w):  # arena.go:197-203: the arena's log lines, then both agents' trees (MCTS.Log)
        for line in self.log:
            w.write("%s\n" % (line,))
        for nm, t in (("A", 0), ("B", 1)):
            w.write("\n%s:\n\n" % nm)
            try:
                rows = self.engine.tree_dump(0, t)
                w.write("%d nodes; root children (move, visits): %s\n" % (len(rows), [(int(r[1]), int(r[2])) for r in rows if r[0] == 1][:16]))
            except K.AZError:
                w.write("(no tree)\n")

    def newB(self, seed, killedA=False):  # arena.go:205-224: B gets a freshly initialised network
        if killedA:
            self.oldCount = 0
        self.engine.net_init(1, seed)
        self.oldCount += 1


class AZ(Arena):
    """agogo.AZ (agogo.go:21-39): New / SelfPlay / Learn / Save / Load over one engine handle.

    Multi-GPU: one AZ per rank (`dist` = an initialised torch.distributed module or None).
    Self-play and arena games are sharded by game with no communication; examples are all-gathered so every
    rank prepares the same batches and trains on its share of them; gradients are averaged across ranks every
    step — inside the engine over NCCL when `az_comm_init` succeeded, otherwise by the host through `dist`
    (the CPU/gloo tests) — so all replicas hold identical weights; win counts are summed."""

    def __init__(self, game, conf, lib=None, n_games=None, seed=1, device=0, flags=0, dist=None,
                 host_allreduce=False):  # agogo.New, agogo.go:41-73
        if not conf.NNConf.IsValid():
            raise RuntimeError("NNConf is not valid. Unable to proceed")  # the reference panics
        if not conf.MCTSConf.IsValid():
            raise RuntimeError("MCTSConf is not valid. Unable to proceed")
        self.lib = lib if lib is not None else K.load()
        self.conf, self.game, self.seed = conf, game, seed
        self.n_games = n_games or 64
        d = K.EngineDesc()
        d.game = K.GameDesc(game.kind, game.m, game.n, game.k, game.komi, game.max_moves, game.zobrist_seed)
        m, n = conf.MCTSConf, conf.NNConf
        d.mcts = K.MCTSConfig(m.PUCT, m.Timeout, m.M, m.N, m.RandomCount, m.Budget, m.RandomMinVisits,
                              m.RandomTemperature, int(m.DumbPass), m.ResignPercentage, m.PassPreference, m.Sims, m.Workers)
        d.nn = K.DualConfig(n.K, n.SharedLayers, n.FC, n.L2, n.BatchSize, n.Width, n.Height, n.Features, n.ActionSpace,
                            int(n.FwdOnly))
        d.encoder, d.n_games, d.device, d.flags = conf.Encoder, self.n_games, device, flags
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        d.seed = derive_seed(seed, 102 if self.rank == 0 else 102 + 7919 * self.rank)  # per-rank coin stream
        self.engine = self.lib.create(d)
        self.engine_comm = False
        self.host_allreduce = host_allreduce
        if self.world > 1 and not host_allreduce:
            self._init_engine_comm()
        self.engine.net_init(0, derive_seed(seed, 100))  # a.Init(), b.Init() (agogo.go:52-57)
        self.engine.net_init(1, derive_seed(seed, 101))
        self.A, self.B = Agent(self, 0), Agent(self, 1)
        self.useDummy = True
        self.epoch = 0
        self.gameNumber = 0
        self.name = conf.Name or "UNKNOWN GAME"
        self.log = []

    def _init_engine_comm(self):
        """NCCL communicator inside the engine: rank 0 creates the id, the host group broadcasts it."""
        import torch
        try:
            uid = K.comm_unique_id(self.lib)  # every rank probes; rank 0's id is the one used
        except K.AZError:
            return  # library without NCCL (the oracle): gradients go through the host group instead
        t = torch.tensor(list(uid), dtype=torch.uint8)
        if self.dist.get_backend() == "nccl":
            t = t.cuda()
        self.dist.broadcast(t, 0)
        self.engine.comm_init(self.rank, self.world, bytes(t.cpu().tolist()))
        self.engine_comm = True

    def _share(self, n):
        """games of this rank when n games are sharded round-robin"""
        return len(range(self.rank, n, self.world))

    def _gather_examples(self, ex):
        """Every rank ends up with all ranks' examples, rank-major (tensor all-gather of padded arrays: boards of a 19x19
        epoch are gigabytes, too much for pickled objects)."""
        if self.world == 1:
            return ex
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        n = len(ex)
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(cnt) for _ in range(self.world)]
        self.dist.all_gather(counts, cnt)
        counts = [int(c.item()) for c in counts]
        mx = max(counts)
        if mx == 0:
            return []
        plane = self.engine.plane
        A1 = self.engine.action_space + 1
        B = np.zeros((mx, plane), np.float32)
        P = np.zeros((mx, A1), np.float32)
        V = np.zeros(mx, np.float32)
        for i, x in enumerate(ex):
            B[i], P[i], V[i] = x.Board, x.Policy, x.Value
        out = []
        for arr in (B, P, V):
            t = torch.from_numpy(arr).to(dev)
            parts = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(parts, t)
            out.append([p.cpu().numpy() for p in parts])
        return [Example(out[0][r][i], out[1][r][i], float(out[2][r][i])) for r in range(self.world) for i in range(counts[r])]

    def _train(self, Xs, Pi, V, batches, nniters, seed):
        """dual.Train (meta.go:16-54) on this rank's share of the batches, gradients averaged over ranks."""
        e, bs = self.engine, self.conf.NNConf.BatchSize
        if self.world == 1 and not self.host_allreduce:
            return e.train(1, Xs, Pi, V, batches, nniters, lr=0.1, shuffle_seed=seed)
        usable = (batches // self.world) * self.world
        if usable == 0:
            raise RuntimeError("batches is nil, probably too few examples regarding the batchsize")
        mine = [b for b in range(usable) if b % self.world == self.rank]
        rows = np.concatenate([np.arange(b * bs, (b + 1) * bs) for b in mine])
        Xl, Pl, Vl = Xs[rows].copy(), Pi[rows].copy(), V[rows].copy()
        lseed = (seed + self.rank) & _M64
        if self.engine_comm:
            return e.train(1, Xl, Pl, Vl, len(mine), nniters, lr=0.1, shuffle_seed=lseed)
        import torch
        costs, rng = [], Rng(lseed)
        for _ in range(nniters):
            for bat in range(len(mine)):
                sl = slice(bat * bs, (bat + 1) * bs)
                g, c = e.train_grads(1, Xl[sl], Pl[sl], Vl[sl])
                if self.world > 1:
                    t = torch.from_numpy(g)
                    self.dist.all_reduce(t)
                    g = (t / self.world).numpy()
                e.train_apply(1, g, 0.1)
                costs.append(c)
            shuffle_rows(Xl, Pl, Vl, rng)
        return np.array(costs, np.float32)

    def _global_stats(self):
        aw, al, ad = self.engine.stats(0)
        bw, bl, bd = self.engine.stats(1)
        v = np.array([aw, al, ad, bw, bl, bd], np.float64)
        if self.world > 1:
            import torch
            t = torch.from_numpy(v)
            if self.dist.get_backend() == "nccl":
                t = t.cuda()
            self.dist.all_reduce(t)
            v = t.cpu().numpy()
        return tuple(np.float32(x) for x in v)

    # ---- self-play -----------------------------------------------------------------------------
    def setupSelfPlay(self, it):  # agogo.go:75-90
        self.A.SwitchToInference()
        self.B.SwitchToInference()
        if it == 0 and self.useDummy:
            self.A.useDummy()
            self.B.useDummy()

    def _play(self, n, record):
        """n x (Arena.Play(record, nil, aug); game.Reset()) (agogo.go:93-97,144-148), concurrently."""
        e = self.engine
        e.examples(clear=True)
        if n <= 0:
            return []
        e.arena_play(n, record)
        last = e.game_record(n - 1)
        self.A.Player = last["a_player"]
        self.B.Player = K.WHITE if last["a_player"] == K.BLACK else K.BLACK
        boards, pols, vals = e.examples(clear=True)
        ex = [Example(boards[i], pols[i], float(vals[i])) for i in range(len(vals))]
        if record and self.conf.Augmenter is not None:
            ex = [y for x in ex for y in self.conf.Augmenter(x)]
        return ex

    def SelfPlay(self, episodes=1):
        return self._play(episodes, True)

    def Play(self, record, enc=None, aug=None, game_number=0):
        """(*Arena).Play(record, enc, aug) (arena.go:80-179) with the reference's signature: ONE game, stepped ply by ply
        so that the OutputEncoder is called with the MetaState after every move (arena.go:131-133); the Augmenter is
        applied to every kept example (arena.go:115-121).  Returns (None, examples) like the reference (arena.go:178)."""
        e = self.engine
        e.examples(clear=True)
        e.arena_begin(1, record)
        rec = e.game_record(0)
        self.A.Player = rec["a_player"]
        self.B.Player = K.WHITE if rec["a_player"] == K.BLACK else K.BLACK
        active = 1
        self.gameNumber = game_number
        while active:
            active = e.arena_step()
            self._last_state = e.game_state(0)
            if enc is not None:
                enc.Encode(MetaState(self, self._last_state, game_number))
        e.arena_finish()
        boards, pols, vals = e.examples(clear=True)
        ex = [Example(boards[i], pols[i], float(vals[i])) for i in range(len(vals))]
        if aug is not None:
            ex = [y for x in ex for y in aug(x)]
        return K.NONE, ex

    @staticmethod
    def shuffleExamples(ex, seed):  # agogo.go:251-257 with an injected seed
        r = Rng(seed)
        for i in range(len(ex)):
            j = r.intn(i + 1)
            ex[i], ex[j] = ex[j], ex[i]

    def prepareExamples(self, ex, seed):  # agogo.go:211-249
        self.shuffleExamples(ex, seed)
        bs = self.conf.NNConf.BatchSize
        batches = len(ex) // bs
        total = batches * bs
        if batches == 0:
            return None, None, None, 0
        Xs = np.stack([x.Board for x in ex[:total]]).astype(np.float32)
        Pi = np.stack([x.Policy for x in ex[:total]]).astype(np.float32)
        V = np.array([x.Value for x in ex[:total]], np.float32)
        return Xs, Pi, V, batches

    # ---- AZ.Learn (agogo.go:100-172) ----------------------------------------------------------
    def Learn(self, iters, episodes, nniters, arenaGames, on_epoch=None):
        import time
        e = self.engine
        for self.epoch in range(iters):
            ep = self.epoch
            t0 = time.perf_counter()
            self.setupSelfPlay(ep)
            mine = self._play(self._share(episodes), True)
            t1 = time.perf_counter()
            ex = self._gather_examples(mine)
            t2 = time.perf_counter()
            if self.conf.MaxExamples > 0 and len(ex) > self.conf.MaxExamples:
                self.shuffleExamples(ex, derive_seed(self.seed, 1000 + 10 * ep))
                ex = ex[:self.conf.MaxExamples]
            Xs, Pi, V, batches = self.prepareExamples(ex, derive_seed(self.seed, 1001 + 10 * ep))
            if batches == 0:
                raise RuntimeError("batches is nil, probably too few examples regarding the batchsize")
            t3 = time.perf_counter()
            costs = self._train(Xs, Pi, V, batches, nniters, derive_seed(self.seed, 1002 + 10 * ep))
            t4 = time.perf_counter()
            self.B.SwitchToInference()
            self.A.resetStats()
            self.B.resetStats()
            self._play(self._share(arenaGames), False)
            aw, al, ad, bw, bl, bd = self._global_stats()
            t5 = time.perf_counter()
            with np.errstate(invalid="ignore", divide="ignore"):
                ratio = np.float32(bw) / (np.float32(bw) + np.float32(aw))
            promoted = bool(ratio > np.float32(self.conf.UpdateThreshold))  # NaN (0/0) never promotes
            if promoted:
                e.net_copy(0, 1)  # A.NN = B.NN (agogo.go:161)
            self.gameNumber = arenaGames                           # agogo.go:144: the loop variable after the arena games
            self.newB(derive_seed(self.seed, 200 + ep), promoted)  # arena.go:205-224
            self.log.append(dict(a=(float(aw), float(al), float(ad)), b=(float(bw), float(bl), float(bd)), n_examples=len(ex), batches=batches, promoted=promoted,
                                 first_cost=float(costs[0]), last_cost=float(costs[-1]),
                                 phase_seconds=dict(selfplay=round(t1 - t0, 3), gather=round(t2 - t1, 3), prepare=round(t3 - t2, 3),
                                                    train=round(t4 - t3, 3), arena=round(t5 - t4, 3),
                                                    train_steps=int(len(costs)))))
            if on_epoch is not None:
                on_epoch(ep, self.log[-1])
        return None

    # ---- checkpoint (agogo.go:175-209): the reference's gob container of the Model()-ordered tensors (gobfmt.py; the
    # tensor.Dense layout inside it is restated from memory, unverified), or a flat .npz when the name says so

    def Save(self, filename):
        nt, _ = self.engine.param_count()
        descs = [self.engine.param_desc(i) for i in range(nt)]
        params = self.engine.net_get(0)
        if str(filename).endswith(".npz"):
            np.savez(filename, params=params, names=np.array([d[0] for d in descs]),
                     shapes=np.array([list(d[1]) + [1] * (4 - len(d[1])) for d in descs]))
            return
        from . import gobfmt
        with open(filename, "wb") as f:  # os.O_CREATE|os.O_TRUNC|os.O_WRONLY
            f.write(gobfmt.save_stream([params[off:off + size].reshape(shape) for (_, shape, off, size) in descs]))

    def Load(self, filename):  # agogo.go:187-209: both A and B get the stored net, useDummy is cleared
        if str(filename).endswith(".npz"):
            flat = np.load(filename)["params"]
        else:
            from . import gobfmt
            nt, nf = self.engine.param_count()
            descs = [self.engine.param_desc(i) for i in range(nt)]
            with open(filename, "rb") as f:
                tensors = gobfmt.load_stream(f.read())
            if len(tensors) != nt:
                raise RuntimeError("checkpoint holds %d tensors, the net's Model() has %d" % (len(tensors), nt))
            flat = np.empty(nf, np.float32)
            for t, (name, shape, off, size) in zip(tensors, descs):
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError("checkpoint tensor %s has shape %r, expected %r" % (name, t.shape, shape))
                flat[off:off + size] = t.reshape(-1)
        self.engine.net_set(0, flat)
        self.engine.net_set(1, flat)
        self.useDummy = False
