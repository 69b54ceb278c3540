"""Go Text Protocol front-end over the engine (the reference's internal/gtp + game/wq/gtp.go command set: boardsize,
clear_board, genmove, known_command, komi, list_commands, name, play, protocol_version, quit, showboard, undo, version,
plus final_score).  The board lives here (so `undo` works, which the reference's wq.Game cannot: UndoLastMove panics,
wq/game.go:119); legality and captures come from the engine's rules kernels (`az_rules_apply`), moves from
`Agent.Search` on the current position (`az_search`), under the complete-rules mode (AZ_FLAG_WQ_COMPLETE: occupied points,
suicide, simple ko, positional superko and own-eye fills are illegal; area scoring) — a Go program that cannot refuse a suicide or a ko
recapture is not playable against another one.

    python -m agogo_b200.gtp --size 9 --sims 200 [--k 64 --blocks 6 --fc 128] [--checkpoint net.model]
"""
import sys

import numpy as np

from . import _capi as K

COLS = "ABCDEFGHJKLMNOPQRSTUVWXYZ"  # no I


class GTPEngine:
    name, version = "agogo_b200", "0.2"
    known = ["boardsize", "clear_board", "final_score", "genmove", "known_command", "komi", "list_commands", "name", "play",
             "protocol_version", "quit", "showboard", "undo", "version"]

    def __init__(self, lib=None, size=9, komi=7.5, sims=100, nn=None, inferer=K.INF_DUMMY, params=None, seed=1,
                 name=None, version=None):  # gtp.New(g, name, version, known) (internal/gtp/gtp.go:47-57)
        if name is not None:
            self.name = name
        if version is not None:
            self.version = version
        self.lib = lib if lib is not None else K.load()
        self.sims, self.nn, self.inferer, self.params, self.seed = sims, nn, inferer, params, seed
        self.komi = komi
        self.engine = None
        self._new_engine(size)

    # ---- engine / position -------------------------------------------------------------------------------------------
    def _new_engine(self, size):
        if self.engine is not None:
            self.engine.close()
        self.size = size
        nn = dict(self.nn) if self.nn else dict(k=4, shared_layers=1, fc=8, batch_size=2)
        nn.update(features=18, action_space=size * size + 1)
        d = K.make_desc(K.GAME_WQ, size, size, 0, komi=self.komi, sims=self.sims, n_games=1, seed=self.seed,
                        flags=K.FLAG_WQ_COMPLETE, nn=nn)
        self.engine = self.lib.create(d)
        if self.inferer == K.INF_DUAL:
            if self.params is not None:
                self.engine.net_set(0, self.params)
            else:
                self.engine.net_init(0, self.seed)
            self.engine.set_inferer(0, K.INF_DUAL)
        else:
            self.engine.set_inferer(0, K.INF_DUMMY, 0)
        self.clear()

    def clear(self):
        self.board = np.zeros(self.size * self.size, np.int32)
        self.boards = []   # board before every move (Historical)
        self.moves = []    # (player, move)
        self.kos = [-1]    # ko point for the side to move, per position
        self.passes = 0

    def _vertex(self, mv):
        if mv == K.PASS:
            return "pass"
        r, c = divmod(mv, self.size)
        return "%s%d" % (COLS[c], self.size - r)

    def _parse_vertex(self, s):
        s = s.strip().upper()
        if s == "PASS":
            return K.PASS
        c, r = COLS.index(s[0]), self.size - int(s[1:])
        if not (0 <= c < self.size and 0 <= r < self.size):
            raise ValueError("vertex off board")
        return r * self.size + c

    @staticmethod
    def _colour(s):
        s = s.strip().lower()
        if s in ("b", "black"):
            return K.BLACK
        if s in ("w", "white"):
            return K.WHITE
        raise ValueError("invalid color")

    def _apply(self, player, mv):
        """play a move on the front-end's board through the engine's rules; returns False when illegal"""
        if mv != K.PASS:
            if mv == self.kos[-1]:
                return False
            check, applied, out, taken = self.engine.rules_apply(self.board[None], [player], [mv])
            if not (check[0] and applied[0]):
                return False
            new = out[0].copy()
            if any((new == b).all() for b in self.boards):  # positional superko over the whole game, as in the engine's
                return False                                 # search (genmove hands it every earlier board)
            ko = -1
            if taken[0] == 1:  # a lone stone that captured one stone and has no other liberty: simple ko
                gone = int(np.flatnonzero((self.board != 0) & (new == 0))[0])
                r, c = divmod(mv, self.size)
                nb = [mv + 1 if c + 1 < self.size else -1, mv + self.size if r + 1 < self.size else -1,
                      mv - 1 if c > 0 else -1, mv - self.size if r > 0 else -1]
                if all(a < 0 or (new[a] != player and (new[a] != 0 or a == gone)) for a in nb):
                    ko = gone
            self.boards.append(self.board)
            self.board = new
            self.kos.append(ko)
            self.passes = 0
        else:
            self.boards.append(self.board)
            self.kos.append(-1)
            self.passes += 1
        self.moves.append((player, mv))
        return True

    # ---- commands ----------------------------------------------------------------------------------------------------
    def cmd_protocol_version(self, args):
        return "2"

    def cmd_name(self, args):
        return self.name

    def cmd_version(self, args):
        return self.version

    def cmd_known_command(self, args):
        return "true" if args and args[0] in self.known else "false"

    def cmd_list_commands(self, args):
        return "\n".join(self.known)

    def cmd_quit(self, args):
        return ""

    def cmd_boardsize(self, args):
        n = int(args[0])
        if not 2 <= n <= 25:
            raise ValueError("unacceptable size")
        self._new_engine(n)
        return ""

    def cmd_clear_board(self, args):
        self.clear()
        self.engine.reset_tree(0)
        return ""

    def cmd_komi(self, args):
        self.komi = float(args[0])
        self._new_engine_keep_position()
        return ""

    def _new_engine_keep_position(self):
        saved = (self.board, self.boards, self.moves, self.kos, self.passes)
        self._new_engine(self.size)
        self.board, self.boards, self.moves, self.kos, self.passes = saved

    def cmd_play(self, args):
        player, mv = self._colour(args[0]), self._parse_vertex(args[1])
        if not self._apply(player, mv):
            raise ValueError("illegal move")
        return ""

    def cmd_genmove(self, args):
        player = self._colour(args[0])
        n = len(self.moves)
        hist = np.array(self.boards, np.int32) if self.boards else None   # the whole game: superko reads all of it
        best, _ = self.engine.search(0, self.board, player, player, move_number=n, passes=min(self.passes, 1),
                                     hist=hist, last_move=self.moves[-1][1] if self.moves else K.PASS, ko=self.kos[-1])
        if best == K.RESIGN:
            return "resign"
        if not self._apply(player, best):  # cannot happen: the search only proposes legal moves
            raise RuntimeError("engine proposed an illegal move")
        return self._vertex(best)

    def cmd_undo(self, args):
        if not self.moves:
            raise ValueError("cannot undo")
        self.moves.pop()
        self.kos.pop()
        self.board = self.boards.pop()
        self.passes = 0
        for _, mv in reversed(self.moves):
            if mv != K.PASS:
                break
            self.passes += 1
        self.engine.reset_tree(0)
        return ""

    def cmd_showboard(self, args):
        rows = []
        for r in range(self.size):
            rows.append("%2d %s" % (self.size - r, " ".join(".XO"[v] for v in self.board[r * self.size:(r + 1) * self.size])))
        return "\n" + "\n".join(rows) + "\n   " + " ".join(COLS[:self.size])

    def cmd_final_score(self, args):
        _, _, sb, sw = self.engine.rules_status(self.board[None], passes=[2])
        d = float(sb[0]) - float(sw[0]) - self.komi
        return "0" if d == 0 else ("B+%g" % d if d > 0 else "W+%g" % -d)

    def handle(self, line):
        """One GTP line -> response text ('' for comments / blank lines), quit flag."""
        line = line.split("#")[0].strip()
        if not line:
            return "", False
        parts = line.split()
        ident = ""
        if parts[0].isdigit():
            ident, parts = parts[0], parts[1:]
        cmd, args = parts[0].lower(), parts[1:]   # the reference lower-cases the line (gtp.go:111-113)
        fn = getattr(self, "cmd_" + cmd, None)
        if fn is None or cmd not in self.known:
            return "?%s Unknown command \"%s\"\n\n" % (ident, cmd), False   # gtp.go:103
        try:
            return "=%s %s\n\n" % (ident, fn(args)), cmd == "quit"
        except (ValueError, IndexError) as ex:
            return "?%s %s\n\n" % (ident, ex), False


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=9)
    ap.add_argument("--komi", type=float, default=7.5)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--fc", type=int, default=128)
    ap.add_argument("--checkpoint", default=None)
    a = ap.parse_args(argv)
    nn = dict(k=a.k, shared_layers=a.blocks, fc=a.fc, batch_size=2) if a.k else None
    eng = GTPEngine(size=a.size, komi=a.komi, sims=a.sims, nn=nn, inferer=K.INF_DUAL if a.k else K.INF_DUMMY)
    for line in sys.stdin:
        out, quit_ = eng.handle(line)
        if out:
            sys.stdout.write(out)
            sys.stdout.flush()
        if quit_:
            break


if __name__ == "__main__":
    main()
