"""A second, independent restatement of the reference's mcts package (tree.go, node.go, search.go, utils.go) in plain
Python with numpy float32 scalars — written from the Go sources, not from oracle/mcts.hpp — used only to cross-check the
C++ oracle's search bit for bit (tests/test_oracle_mcts_pyref.py).  Same canonical choices as the oracle: one worker,
exactly `sims` pipeline calls per Search, Go's unstable sort.Sort pinned to a stable sort.  Game: m,n,k (mnk.go) driven
the way mcts/example_test.go drives it — ONE tree searched by both colours, the caller applies the move to the shared
state.  Evaluator: a table keyed by state.MoveNumber() (the Example's dummyNN shape)."""
import numpy as np

f32 = np.float32
PASS, NIL = -1, -1
BLACK, WHITE = 1, 2
MAXTREESIZE = 25000000
INVALID, ACTIVE = 0, 1


def opponent(p):
    return WHITE if p == BLACK else BLACK


class MNK:  # game/mnk/mnk.go
    supports_undo = True

    def __init__(self, m, n, k):
        self.m, self.n, self.k = m, n, k
        self.board = [0] * (m * n)
        self.history, self.hist_ptr, self.next = [], 0, 0

    def clone(self):  # mnk.go:208-219
        c = MNK(self.m, self.n, self.k)
        c.board, c.history, c.hist_ptr, c.next = list(self.board), list(self.history), self.hist_ptr, self.next
        return c

    def action_space(self):
        return self.m * self.n

    def move_number(self):
        return len(self.history)

    def passes(self):
        return -1

    def last_move(self):
        return self.history[self.hist_ptr - 1] if self.history else (0, PASS)

    def check(self, player, move):
        if move == PASS:
            return False
        if move >= len(self.board):
            return False
        return self.board[move] == 0

    def apply(self, player, move):  # in place
        if not self.check(player, move):
            return self
        self.board[move] = player
        self.hist_ptr += 1
        if len(self.history) < self.hist_ptr:
            self.history.append((player, move))
        else:
            self.history[self.hist_ptr - 1] = (player, move)
        self.next = opponent(player)
        return self

    def undo_last_move(self):
        if self.history:
            self.board[self.history[self.hist_ptr - 1][1]] = 0
            self.hist_ptr -= 1

    def fwd(self):
        if self.history:
            self.hist_ptr += 1

    def eq(self, other):
        return self.board == other.board

    def is_winner(self, colour):
        from tests.pyref_rules import mnk_is_winner
        return mnk_is_winner(self.board, self.m, self.n, self.k, colour)

    def ended(self):
        if self.is_winner(BLACK):
            return True, BLACK
        if self.is_winner(WHITE):
            return True, WHITE
        return all(c != 0 for c in self.board), 0


class C4:  # game/c4/c4.go + game/c4/game.go, quirks kept: Apply never flips the player, MoveNumber() is always 1,
    # Passes() is always 0, Clone pads history / historical with two zero entries, Eq compares other's board with itself
    supports_undo = True

    def __init__(self, rows, cols, nn):
        self.rows, self.cols, self.nn = rows, cols, nn
        self.board = [0] * (rows * cols)
        self.history, self.historical = [], []
        self.next, self.hist_ptr, self.move_count, self.pass_count = 0, 0, 0, 0

    def clone(self):  # game.go:131-150
        c = C4(self.rows, self.cols, self.nn)
        c.board = list(self.board)
        c.history = list(self.history) + [(0, 0), (0, 0)]
        c.historical = [list(h) for h in self.historical] + [None, None]
        c.next, c.hist_ptr, c.move_count, c.pass_count = self.next, self.hist_ptr, self.move_count, self.pass_count
        return c

    def action_space(self):
        return self.cols

    def move_number(self):
        return self.move_count + 1

    def passes(self):
        return 0

    def last_move(self):
        return self.history[self.hist_ptr - 1] if self.history else (0, PASS)

    def _drop_row(self, col):
        for row in range(self.rows - 1, -1, -1):
            if self.board[row * self.cols + col] == 0:
                return row
        return None

    def check(self, player, move):
        return move == PASS or self._drop_row(move) is not None

    def apply(self, player, move):  # game.go:55-72, in place
        hb = list(self.board)
        ok = True
        if move != PASS:
            row = self._drop_row(move)
            if row is None:
                ok = False
            else:
                self.board[row * self.cols + move] = player
        if ok:
            self.history.append((player, move))
            self.historical.append(hb)
            self.hist_ptr += 1
        self.pass_count = self.pass_count + 1 if move == PASS else 0
        return self

    def undo_last_move(self):
        raise AssertionError("MoveNumber() is constant: newRootState never undoes a c4 move")

    def fwd(self):
        if self.history:
            self.hist_ptr += 1

    def eq(self, ot):  # game.go:98-129
        if self.hist_ptr != ot.hist_ptr or self.move_count != ot.move_count:
            return False
        if len(self.history) != len(ot.history) or len(self.historical) != len(ot.historical):
            return False
        if any(a != b for a, b in zip(self.history, ot.history)):
            return False
        for a, b in zip(self.historical, ot.historical):
            if a is not None and any(x != y for x, y in zip(a, b)):
                return False
        return True

    def ended(self):
        from tests.pyref_rules import c4_status
        e, w, _, _ = c4_status(self.board, self.rows, self.cols, self.nn, self.pass_count)
        return e, w


class WQ:  # game/wq/wq.go + game/wq/game.go.  What the reference leaves unfinished (it panics) follows the repository's
    # stated completion (DESIGN.md §2 "wq gap"): pass = board no-op + passes++ (a stone resets it), historical = board
    # before each move, Score = Board.Score as implemented, no undo (so no tree reuse), Apply returns a new state.
    supports_undo = False

    def __init__(self, size, komi):
        self.size, self.komi = size, komi
        self.board = [0] * (size * size)
        self.history, self.historical = [], []
        self.next, self.n_passes = BLACK, 0

    def clone(self):
        c = WQ(self.size, self.komi)
        c.board, c.history, c.historical = list(self.board), list(self.history), list(self.historical)
        c.next, c.n_passes = self.next, self.n_passes
        return c

    def action_space(self):
        return self.size * self.size

    def move_number(self):
        return len(self.history)

    def passes(self):
        return self.n_passes

    def last_move(self):  # game.go:54-59: (None, Pass) for an empty history
        return self.history[-1] if self.history else (0, PASS)

    def check(self, player, move):  # game.go:65-79: occupied points are not rejected
        from tests.pyref_rules import wq_board_check
        if move == PASS:
            return True
        if move >= len(self.board):
            return False
        return wq_board_check(self.board, self.size, player, move)[1]

    def apply(self, player, move):  # game.go:81-92: a new state; Board.Apply's error is ignored
        from tests.pyref_rules import wq_apply
        ns = self.clone()
        ns.historical.append(list(self.board))
        if move == PASS:
            ns.n_passes = self.n_passes + 1
        else:
            ns.board = wq_apply(self.board, self.size, player, move)[2]
            ns.n_passes = 0
        ns.next = opponent(player)
        ns.history.append((player, move))
        return ns

    def score(self, p):
        from tests.pyref_rules import wq_score
        return f32(wq_score(self.board, self.size, p))

    def ended(self):  # game.go:94-115
        if self.n_passes < 2:
            return False, 0
        ws, bs = self.score(WHITE), self.score(BLACK)
        return True, (0 if ws == bs else (WHITE if ws > bs else BLACK))


class WQComplete(WQ):  # AZ_FLAG_WQ_COMPLETE (OUR completion, include/agogo_b200.h): occupied points, suicide, simple ko, own
    # single-point eyes and POSITIONAL SUPERKO are illegal; area scoring, komi added to White in Ended.  The rules themselves
    # are the naive statement in tests/pyref_rules.py; this class only threads the ko point and the list of earlier
    # positions (`historical`: every board a move was played from, inside a search as well) through the states.
    complete = True

    def __init__(self, size, komi):
        super().__init__(size, komi)
        self.ko = -1

    def clone(self):
        c = WQComplete(self.size, self.komi)
        c.board, c.history, c.historical = list(self.board), list(self.history), list(self.historical)
        c.next, c.n_passes, c.ko = self.next, self.n_passes, self.ko
        return c

    def _check(self, player, move):
        from tests.pyref_rules import wq_complete_check
        return wq_complete_check(self.board, self.size, player, move, self.ko, {tuple(b) for b in self.historical})

    def check(self, player, move):
        if move == PASS:
            return True
        if move >= len(self.board):
            return False
        return self._check(player, move)[0]

    def apply(self, player, move):
        ns = self.clone()
        ns.historical.append(list(self.board))
        if move == PASS:
            ns.n_passes, ns.ko = self.n_passes + 1, -1
        else:
            ok, captured, ko = self._check(player, move)
            if ok:
                ns.board[move] = player
                for q in captured:
                    ns.board[q] = 0
            ns.n_passes, ns.ko = 0, (ko if ok else -1)
        ns.next = opponent(player)
        ns.history.append((player, move))
        return ns

    def score(self, p):
        from tests.pyref_rules import wq_area_score
        return f32(wq_area_score(self.board, self.size, p))

    def ended(self):
        if self.n_passes < 2:
            return False, 0
        ws, bs = f32(self.score(WHITE) + f32(self.komi)), self.score(BLACK)
        return True, (0 if ws == bs else (WHITE if ws > bs else BLACK))


class Node:
    __slots__ = ("move", "visits", "status", "black", "min_psa", "score", "vloss")

    def __init__(self):
        self.move, self.visits, self.status = 0, 0, 0
        self.black, self.min_psa, self.score, self.vloss = f32(0), f32(2.0), f32(0), f32(0)

    def has_children(self):
        return self.min_psa <= f32(1)

    def is_expandable(self, r):
        return f32(r) < self.min_psa

    def evaluate(self, player):  # node.go:147-159: the virtual loss is added for White only
        bs = self.black
        if player == WHITE:
            bs = f32(bs + self.vloss)
        s = f32(bs / f32(self.visits))
        if player == WHITE:
            s = f32(f32(1) - s)
        return s


class MCTS:
    def __init__(self, game, puct, sims, table, values, M, N, evaluator=None, workers=1, conf=None, tree_seed=0):
        self.g, self.puct, self.sims, self.table, self.values = game, f32(puct), sims, table, values
        # the rest of mcts.Config (tree.go:15-29); defaults = the Example's: DontPreferPass, DumbPass, no sampling, no resign
        self.conf = dict(pass_preference=0, dumb_pass=True, resign_percentage=0.0, random_count=0, random_min_visits=0,
                         random_temperature=0.0, M=M, N=N)
        self.conf.update(conf or {})
        self.rng_state = tree_seed  # MCTS.rand (tree.go:84), injected: splitmix64 stream
        self.workers = workers      # > 1: the fixed schedule of concurrent pipeline calls (include/agogo_b200.h)
        self.evaluator = evaluator  # callable(state) -> (policy, value); None = table keyed by MoveNumber()
        self.cached_policies = {}   # tree.go:75: (board, move) -> count
        self.max_depth = M * N
        self.nodes, self.children, self.freelist, self.freeables = [], [], [], []
        self.root, self.prev = NIL, None
        self.evals = 0

    # ---- tree.go
    def alloc(self):
        if not self.freelist:
            self.nodes.append(Node())
            self.children.append([])
            return len(self.nodes) - 1
        return self.freelist.pop()

    def new(self, move, score):
        n = self.alloc()
        nd = self.nodes[n]
        nd.move, nd.visits, nd.status, nd.score = move, 1, ACTIVE, f32(score)
        return n

    def free(self, n):
        self.children[n] = []
        self.freelist.append(n)
        nd = self.nodes[n]
        nd.move, nd.visits, nd.status, nd.black, nd.min_psa, nd.score = -1, 0, 0, f32(0), f32(2.0), f32(0)

    def clean_children(self, root):
        for kid in self.children[root]:
            self.nodes[kid].status = INVALID
            self.freeables.append(kid)
            self.clean_children(kid)
        self.children[root] = []

    def cleanup(self, old, new):
        for kid in self.children[old]:
            if kid != new:
                self.nodes[kid].status = INVALID
                self.freeables.append(kid)
                self.clean_children(kid)
        self.children[old] = [new]

    def find_child(self, n, move):
        for kid in self.children[n]:
            if self.nodes[kid].move == move:
                return kid
        return NIL

    def min_psa_ratio(self):
        ratio = f32(len(self.nodes)) / f32(MAXTREESIZE)  # node count: far below either threshold in these tests
        return f32(0.01) if ratio > 0.95 else (f32(0.001) if ratio > 0.5 else f32(0))

    # ---- the evaluator
    def infer(self, state):
        self.evals += 1
        if self.evaluator is not None:
            return self.evaluator(state)
        mn = state.move_number()
        if 0 <= mn < len(self.table):
            return self.table[mn], f32(self.values[mn])
        return np.zeros(self.table.shape[1], np.float32), f32(0)

    # ---- search.go
    def expand_and_simulate(self, parent, state, min_psa_ratio):
        n = self.nodes[parent]
        if not n.is_expandable(min_psa_ratio):
            return f32(0), False
        if state.passes() >= 2:
            return f32(0), False
        return self.expand_after_checks(parent, state, min_psa_ratio)

    def expand_after_checks(self, parent, state, min_psa_ratio):  # search.go:274-338
        n = self.nodes[parent]
        policy, value = self.infer(state)
        pass_prob = policy[len(policy) - 1]
        player = state.next
        if player == WHITE:
            value = f32(f32(1) - value)
        nodelist, legal_sum = [], f32(0)
        for i in range(self.g.action_space()):
            if state.check(player, i):
                nodelist.append([i, f32(policy[i])])
                legal_sum = f32(legal_sum + f32(policy[i]))
        if state.check(player, PASS):
            nodelist.append([PASS, f32(pass_prob)])
            legal_sum = f32(legal_sum + f32(pass_prob))
        if legal_sum > np.finfo(np.float32).smallest_subnormal:
            for p in nodelist:
                p[1] = f32(p[1] / legal_sum)
        else:
            with np.errstate(divide="ignore"):  # an empty list gives +Inf in Go too (float division), unused
                prob = f32(f32(1) / f32(len(nodelist)))
            for p in nodelist:
                p[1] = prob
        if not nodelist:
            return value, True
        nodelist.sort(key=lambda p: -float(p[1]))  # stable, best score first
        max_psa = nodelist[0][1]
        old_min = f32(max_psa * n.min_psa)
        new_min = f32(max_psa * min_psa_ratio)
        skipped = False
        for move, score in nodelist:
            if score < new_min:
                skipped = True
            elif score < old_min:
                if self.find_child(parent, move) == NIL:
                    self.children[parent].append(self.new(move, score))
        n.min_psa = f32(min_psa_ratio) if skipped else f32(0)
        return value, True

    def select(self, nid, player):  # node.go:170-237
        kids = self.children[nid]
        parent_visits = 0
        for kid in kids:
            c = self.nodes[kid]
            if c.status != INVALID:
                parent_visits += c.visits
        best, best_value = NIL, f32(-np.inf)
        numerator = np.sqrt(f32(parent_visits))
        for kid in kids:
            c = self.nodes[kid]
            if c.status != ACTIVE:
                continue
            assert c.visits > 0  # nodes are born with one visit (tree.go:110): the fpu branch is dead
            qsa = c.evaluate(player)
            denominator = f32(f32(1.0) + f32(c.visits))
            last = f32(numerator / denominator)
            puct = f32(f32(self.puct * c.score) * last)
            usa = f32(qsa + puct)
            if usa > best_value:
                best_value, best = usa, kid
        assert best != NIL
        return best

    def update(self, nid, score):
        nd = self.nodes[nid]
        nd.visits += 1
        nd.black = f32(nd.black + f32(score))

    def pipeline(self, cur, start, depth):
        """Returns the result or None for the null result."""
        depth += 1
        if depth > self.max_depth:
            return None
        player = cur.next
        n = self.nodes[start]
        ret = None
        if n.is_expandable(0) and cur.passes() >= 2:  # search.go:226-228, utils.go:62-67
            ret = f32(f32(cur.score(BLACK) - cur.score(WHITE)) - f32(cur.komi))
        elif n.is_expandable(0) and len(self.nodes) < MAXTREESIZE:
            had = n.has_children()
            value, ok = self.expand_and_simulate(start, cur, self.min_psa_ratio())
            if not had and ok:
                ret = value
        if n.has_children() and ret is None:
            nxt = self.select(start, player)
            move = self.nodes[nxt].move
            if cur.check(player, move):
                cur = cur.apply(player, move)
                ret = self.pipeline(cur, nxt, depth)
        if ret is not None:
            self.update(start, ret)
        return ret

    def run_workers(self, iterations):
        """search.go:112-130 starts runtime.NumCPU() pipeline calls at once; their interleaving is fixed here as the header
        states it: rounds of `workers` descents, each marking its path with the virtual-loss flag (search.go:222, a store
        of 3.0) and stopping where it needs an evaluation; null results and terminals complete (and clear their flags) at
        once; then the pending calls finish in start order — expansion (candidates an earlier call of the round created
        are found by findChild), Update along the path, undoVirtualLoss (a store of 0)."""
        left = iterations
        while left > 0:
            v = min(self.workers, left)
            left -= v
            pending = []
            for _ in range(v):
                cur, node, depth, path, ret, wait = self.g.clone(), self.root, 0, [], None, False
                while True:
                    depth += 1
                    if depth > self.max_depth:
                        break
                    player = cur.next
                    n = self.nodes[node]
                    n.vloss = f32(3.0)
                    path.append(node)
                    if n.is_expandable(0) and cur.passes() >= 2:
                        ret = f32(f32(cur.score(BLACK) - cur.score(WHITE)) - f32(cur.komi))
                        break
                    if n.is_expandable(0) and len(self.nodes) < MAXTREESIZE and n.is_expandable(self.min_psa_ratio()):
                        assert not n.has_children()
                        pending.append((path, cur))
                        wait = True
                        break
                    if not n.has_children():
                        break
                    nxt = self.select(node, player)
                    move = self.nodes[nxt].move
                    if not cur.check(player, move):
                        break
                    cur = cur.apply(player, move)
                    node = nxt
                if not wait:
                    self.finish_path(path, ret)
            for path, st in pending:  # a leaf an earlier call of the round expanded: min_psa is 0 now, nothing is created
                value, ok = self.expand_after_checks(path[-1], st, self.min_psa_ratio())
                self.finish_path(path, value if ok else None)

    def finish_path(self, path, ret):
        for nid in reversed(path):
            if ret is not None:
                self.update(nid, ret)
            self.nodes[nid].vloss = f32(0)

    def new_root_state(self):
        if self.root == NIL or self.prev is None:
            return False
        depth = self.g.move_number() - self.prev.move_number()
        if depth < 0:
            return False
        if depth > 0 and not self.g.supports_undo:  # wq: UndoLastMove panics in the reference -> a fresh root
            return False
        tmp = self.g.clone()
        for _ in range(depth):
            tmp.undo_last_move()
        if not tmp.eq(self.prev):
            return False
        for _ in range(depth):
            tmp.fwd()
            player, move = tmp.last_move()
            old = self.root
            new = self.find_child(old, move)
            if new == NIL:
                return False
            self.root = new
            self.cleanup(old, new)
            self.prev = self.prev.apply(player, move)
        if self.g.move_number() != self.prev.move_number():
            return False
        return self.g.eq(self.prev)

    def update_root(self):
        self.freeables = []
        player = self.g.next
        if not self.new_root_state() or self.root == NIL:
            if self.g.check(player, PASS):
                self.root = self.new(PASS, 0)
            else:
                for i in range(self.g.action_space()):
                    if self.g.check(player, i):
                        self.root = self.new(i, 0)
                        break
        self.prev = None
        if not self.children[self.root]:
            self.nodes[self.root].min_psa = f32(2.0)

    def prepare_root(self, player):
        root = self.nodes[self.root]
        had = len(self.children[self.root]) > 0
        value = f32(0)
        if root.is_expandable(0):
            value, _ = self.expand_and_simulate(self.root, self.g, self.min_psa_ratio())
        if not had:
            self.update(self.root, value)

    def fancy_less(self, player, a, b):  # utils.go:17-47
        li, lj = self.nodes[a], self.nodes[b]
        if li.visits != lj.visits:
            return li.visits > lj.visits
        if li.visits == 0:
            return li.score > lj.score
        return li.evaluate(player) > lj.evaluate(player)

    def randomize_children(self, of):  # tree.go:212-247
        accum, norm, vec = f32(0), f32(0), []
        kids = self.children[of]
        for kid in kids:
            visits = self.nodes[kid].visits
            if norm == 0:
                norm = f32(visits)
                if visits <= self.conf["random_min_visits"]:
                    return
            if visits > self.conf["random_min_visits"]:
                base = float(f32(f32(visits) / norm))
                expo = float(f32(f32(1) / f32(self.conf["random_temperature"])))
                accum = f32(accum + f32(base ** expo))  # math32.Pow = float32(math.Pow(float64, float64))
                vec.append(accum)
        self.rng_state, r = _splitmix(self.rng_state)
        rnd = f32(f32(r >> 40) * f32(1.0 / 16777216.0)) * accum  # rand.Float32() stand-in: 24 high bits
        rnd = f32(rnd)
        index = 0
        for i, a in enumerate(vec):
            if rnd < a:
                index = i
                break
        if index == 0:
            return
        for i in range(len(kids) - index):
            kids[i], kids[i + index] = kids[i + index], kids[i]

    def no_pass_best_move(self, best, best_score, player):  # search.go:531-563
        for kid in self.children[self.root]:
            mv = self.nodes[kid].move
            ok = self.g.check(player, mv)
            if mv != PASS and ok:
                nd = self.nodes[kid]
                return mv, (nd.evaluate(player) if nd.visits != 0 else f32(1))
        return best, best_score

    def should_resign(self, best_score, player):  # search.go:502-529
        c = self.conf
        if c["pass_preference"] == 2 or c["resign_percentage"] == 0:
            return False
        if self.g.move_number() <= (c["M"] * c["N"]) // 4:
            return False
        threshold = f32(0.1) if c["resign_percentage"] < 0 else f32(c["resign_percentage"])
        return not (best_score > threshold)

    def best_move(self):  # search.go:341-390
        c = self.conf
        player = self.g.next
        kids = self.children[self.root]
        out = []  # stable insertion sort by fancySort.Less
        for k in kids:
            i = len(out)
            while i > 0 and self.fancy_less(player, k, out[i - 1]):
                i -= 1
            out.insert(i, k)
        kids[:] = out
        if self.g.move_number() < c["random_count"]:
            self.randomize_children(self.root)
        if not kids:
            return PASS
        first = self.nodes[kids[0]]
        best, best_score = first.move, first.evaluate(player)
        root_score = self.nodes[self.root].score  # Node.Score(): the prior
        losing = (root_score > 0 and player == WHITE) or (root_score < 0 and player == BLACK)
        if c["pass_preference"] == 0 and best == PASS:
            best, best_score = self.no_pass_best_move(best, best_score, player)
        elif not c["dumb_pass"] and best == PASS:
            if losing:
                best, best_score = self.no_pass_best_move(best, best_score, player)
        elif not c["dumb_pass"] and self.g.last_move()[1] == PASS:
            if not losing:
                best = PASS
        if best == PASS and self.should_resign(best_score, player):
            best = -2
        return best

    def search(self, player):
        self.update_root()
        self.g.next = player
        for f in self.freeables:
            self.free(f)
        self.prepare_root(player)
        if self.workers > 1:
            self.run_workers(self.sims)
        else:
            for _ in range(self.sims):
                self.pipeline(self.g.clone(), self.root, 0)
        assert self.nodes[self.root].has_children()
        board_key = tuple(self.g.board)  # stands for current.Hash() taken before the move (search.go:96)
        best = self.best_move()
        self.prev = self.g.clone()
        self.cached_policies[(board_key, best)] = self.cached_policies.get((board_key, best), 0.0) + 1.0
        return best

    def policies(self, state):  # tree.go:128-142: counts of the moves Search chose at this board, normalised
        key = tuple(state.board)
        a1 = state.action_space() + 1
        ret = np.array([self.cached_policies.get((key, i), 0.0) for i in range(a1)], np.float32)
        total = f32(0)
        for v in ret:
            total = f32(total + v)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (ret / total).astype(np.float32)

    def dump(self):
        rows = []

        def rec(nid, d):
            n = self.nodes[nid]
            rows.append((d, n.move, n.visits, int(np.float32(n.black).view(np.uint32)), int(np.float32(n.score).view(np.uint32)),
                         1 if n.has_children() else 0, len(self.children[nid])))
            for kid in self.children[nid]:
                rec(kid, d + 1)
        rec(self.root, 0)
        return np.array(rows, np.int64)


# ---------------------------------------------------------------- arena.go / agent.go / dummy.go / cmd/tictactoe encoder
_M64 = (1 << 64) - 1


def _splitmix(s):  # the injected RNG specification (include/agogo_b200.h): splitmix64
    s = (s + 0x9E3779B97F4A7C15) & _M64
    z = s
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return s, z ^ (z >> 31)


def derive_seed(seed, stream):
    return _splitmix((seed ^ ((0xD1B54A32D192ED03 * (stream + 1)) & _M64)) & _M64)[1]


def encode_two_plane(state):  # cmd/tictactoe/main.go:26-47
    board = np.array([1.0 if c == BLACK else (-1.0 if c == WHITE else 0.001) for c in state.board], np.float32)
    nxt = state.next
    plane = np.full(len(state.board), 1.0 if nxt == BLACK else (-1.0 if nxt == WHITE else 0.0), np.float32)
    return np.concatenate([board, plane])


def dummy_evaluator(action_space, captured_player):  # dummy.go: uniform 1/outputSize over outputSize entries
    value = f32(1 if captured_player == BLACK else (-1 if captured_player == WHITE else 0))
    policy = np.full(action_space, f32(1) / f32(action_space), np.float32)
    return lambda state: (policy, value)


def encode_wq18(state):  # encoding_helper.go:29-68
    size = len(state.board)
    out = np.zeros(18 * size, np.float32)
    if state.next == BLACK:
        bstart, wstart, nstart, enc = 0, 8 * size, 16 * size, 1.0
    else:
        bstart, wstart, nstart, enc = 8 * size, 0, 17 * size, -1.0
    current = state.move_number() - 1
    for i in range(1, 8):
        h = current - i
        if 0 < h < current:
            past = state.historical[h]
            two = np.array([1.0 if c == BLACK else (-1.0 if c == WHITE else 0.0) for c in past], np.float32)
            out[bstart:bstart + size] = two
            out[wstart:wstart + size] = -two
        bstart += size
        wstart += size
    out[nstart:nstart + size] = enc
    return out


def arena_play(new_game, make_mcts, coin, record=True, encoder=None, max_moves=0):
    """Arena.Play (arena.go:80-179) for one game.  make_mcts(agent_index, game) builds the agent's fresh MCTS.  Returns
    (moves, winner, a_player, examples [(board, policy, value)], per-ply dumps [(tree A, tree B)])."""
    g = new_game()
    trees = [make_mcts(0, g), make_mcts(1, g)]
    players = [BLACK, WHITE] if coin == 0 else [WHITE, BLACK]
    cur = 0 if coin == 0 else 1
    g.next = players[cur]
    moves, examples, dumps, pass_count = [], [], [], 0
    while True:
        ended, winner = g.ended()
        if ended:
            break
        t = trees[cur]
        t.g = g                        # Agent.Search: MCTS.SetGame(g) ...
        best = t.search(players[cur])  # ... + Search(a.Player)
        pass_count = pass_count + 1 if best == PASS else 0
        if record:
            pol = t.policies(g)
            if np.isfinite(pol).all():
                examples.append([(encoder or encode_two_plane)(g), pol, float(players[cur])])
        dumps.append([tr.dump() if tr.root != NIL else np.zeros((0, 7), np.int64) for tr in trees])
        g = g.apply(players[cur], best)
        moves.append(best)
        cur ^= 1
        if pass_count >= 2:  # arena.go:134-136: `winner` keeps the value of the last loop condition (None)
            if getattr(g, "complete", False):  # OUR complete-rules mode: two passes end the game and it is scored
                _, winner = g.ended()
            break
        if max_moves and len(moves) >= max_moves:  # the repository's cap for games that never end
            break
    for ex in examples:
        ex[2] = 0.0 if winner == 0 else (1.0 if ex[2] == float(winner) else -1.0)
    return moves, winner, players[0], examples, dumps
