"""GTP front-end (agogo_b200/gtp.py) over the oracle library on CPU: the reference's command set (internal/gtp,
game/wq/gtp.go), a scripted capture + ko sequence, undo, final_score, and a generated game that only ever proposes moves the
independent Python rules accept."""
from agogo_b200 import _capi as K
from agogo_b200.gtp import GTPEngine
from tests import pyref_rules as R


def _ok(eng, line):
    out, _ = eng.handle(line)
    assert out.startswith("="), (line, out)
    return out[1:].strip()


def test_gtp_basics_capture_ko_undo(oracle):
    g = GTPEngine(lib=oracle, size=5, komi=5.5, sims=10)
    assert _ok(g, "protocol_version") == "2" and _ok(g, "known_command genmove") == "true"
    assert set(_ok(g, "list_commands").split()) >= {"boardsize", "clear_board", "genmove", "komi", "play", "quit", "showboard", "undo"}
    assert g.handle("frobnicate")[0].startswith("?")
    out, _ = g.handle("7 name")
    assert out == "=7 agogo_b200\n\n"
    # a ko: black B4 A3 B2, white C4 C2 D3; black C3 (the ko stone); white captures it at B3 ...
    for c, v in (("b", "B4"), ("w", "C4"), ("b", "A3"), ("w", "D3"), ("b", "B2"), ("w", "C2"), ("b", "C3")):
        _ok(g, "play %s %s" % (c, v))
    _ok(g, "play w B3")                                         # captures C3: ko
    assert g.board[2 * 5 + 2] == 0 and g.kos[-1] == 2 * 5 + 2
    assert g.handle("play b C3")[0].startswith("? illegal")      # immediate recapture is barred
    assert g.handle("play b B3")[0].startswith("? illegal")      # occupied
    _ok(g, "play b E5"); _ok(g, "play w E1")                     # ko threat and answer ...
    _ok(g, "play b C3")                                          # ... now the recapture is legal
    assert g.board[2 * 5 + 1] == 0
    _ok(g, "undo")
    assert g.board[2 * 5 + 1] == K.WHITE and g.board[2 * 5 + 2] == 0
    # positional superko: after two passes the simple-ko bar is gone, but retaking would recreate the position before
    # white's capture — still illegal; so is it for the engine's own search (the ko point never becomes a root child)
    _ok(g, "undo"); _ok(g, "undo")                               # back to: white has just captured at B3
    assert g.kos[-1] == 2 * 5 + 2
    _ok(g, "play b pass"); _ok(g, "play w pass")
    assert g.kos[-1] == -1 and g.handle("play b C3")[0].startswith("? illegal")
    _, visits = g.engine.search(0, g.board, K.BLACK, K.BLACK, move_number=len(g.moves), passes=0, hist=g.boards,
                                last_move=K.PASS, ko=-1)
    assert visits[2 * 5 + 2] == 0 and visits[4 * 5 + 4] > 0
    _ok(g, "play b E5")
    assert "X" in _ok(g, "showboard")
    assert _ok(g, "final_score")[0] in "BW0"
    _ok(g, "boardsize 7")
    assert g.board.size == 49 and not g.moves
    assert g.handle("quit")[1]


def test_gtp_generated_game_is_legal(oracle):
    g = GTPEngine(lib=oracle, size=5, komi=5.5, sims=12)
    board, ko, player = [0] * 25, -1, 1
    for ply in range(70):
        v = _ok(g, "genmove %s" % ("b" if player == 1 else "w"))
        if v == "pass":
            ko = -1
            if ply and g.moves[-2][1] == K.PASS:
                break
        else:
            mv = g._parse_vertex(v)
            ok, captured, ko = R.wq_complete_check(board, 5, player, mv, ko)
            assert ok, (ply, v)
            board[mv] = player
            for q in captured:
                board[q] = 0
        assert g.board.tolist() == board and g.kos[-1] == ko, ply
        player = 3 - player


def test_gtp_general_like_the_reference(oracle):
    """internal/gtp's Test_General (gtp_test.go:9-31), same engine arguments (name "xx", version "1"), same four exchanges,
    same response text."""
    e = GTPEngine(lib=oracle, size=5, sims=4, name="xx", version="1")
    assert e.handle("version")[0] == "= 1\n\n"
    assert e.handle("known_command hello")[0] == "= false\n\n"
    assert e.handle("known_command name")[0] == "= true\n\n"
    assert e.handle("completelyUnheardOfCommand xxx")[0] == "? Unknown command \"completelyunheardofcommand\"\n\n"
    assert e.handle("name")[0] == "= xx\n\n"
