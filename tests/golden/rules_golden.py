"""Golden vectors transcribed from the reference's own rule tests (literal boards and expected
results; no reference code).  Sources (paths relative to the reference tree):
  game/mnk/mnk_test.go:9-137    TestTicTacToe, TestGomoku, TestTicTacToeEnded      (8 boards)
  game/c4/c4_test.go:9-101      TestGame_Ended                                     (6 boards)
  game/wq/wq_test.go:33-196     applyTests / TestBoard_Apply                       (7 cases)
  game/komi/komi_test.go:10-187 applyTests (same 7 boards; pins `taken` a second time)
  dualnet/config_test.go:5-17   correctRounds                                      (11 KATs)
Colours: Z = None(0), X = Black(1), O = White(2).
"""
Z, X, O = 0, 1, 2

# (m, n, k, board, expect) ; expect keys: winner_is (isWinner(p) true for p), ended, winner
MNK = [
    dict(m=3, n=3, k=3, board=[X, O, X, O, X, O, O, O, X], is_winner=X, ended=1),           # mnk_test.go:13-23
    dict(m=3, n=3, k=3, board=[X, O, O, X, O, X, O, X, X], is_winner=O),                      # mnk_test.go:26-33
    dict(m=7, n=7, k=5, board=[Z, X, Z, Z, Z, Z, Z,
                               Z, Z, X, Z, Z, Z, Z,
                               Z, Z, Z, X, Z, Z, Z,
                               Z, Z, Z, Z, X, Z, Z,
                               Z, Z, Z, Z, Z, X, Z,
                               Z, Z, Z, Z, Z, X, Z,
                               Z, Z, Z, Z, Z, X, Z], is_winner=X, ended=1),                   # mnk_test.go:41-55
    dict(m=7, n=7, k=5, board=[Z, Z, Z, Z, Z, Z, Z,
                               Z, Z, Z, Z, Z, O, Z,
                               Z, Z, Z, Z, O, Z, Z,
                               Z, Z, Z, O, Z, Z, Z,
                               Z, Z, O, Z, Z, Z, Z,
                               Z, O, Z, Z, Z, Z, Z,
                               Z, Z, Z, Z, Z, Z, Z], is_winner=O, ended=1),                   # mnk_test.go:57-72
    dict(m=3, n=3, k=3, board=[O, Z, X, Z, Z, X, Z, O, X], ended=1, winner=X),                # mnk_test.go:80-91
    dict(m=3, n=3, k=3, board=[O, O, O, Z, Z, X, X, O, X], ended=1, winner=O),                # mnk_test.go:93-104
    dict(m=3, n=3, k=3, board=[Z, Z, X, X, O, X, O, O, O], ended=1, winner=O),                # mnk_test.go:106-117
    dict(m=3, n=3, k=3, board=[O, Z, X, X, O, X, O, Z, O], ended=1, winner=O),                # mnk_test.go:119-130
]

# c4 6x7, N=4: (board, ended, winner)
C4 = [
    dict(board=[X, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, Z, Z, Z,
                X, Z, Z, Z, Z, Z, Z,
                O, O, Z, Z, X, Z, X,
                X, O, Z, O, X, Z, X], ended=0, winner=Z),                                    # c4_test.go:17-29
    dict(board=[X, O, X, O, X, O, X,
                O, O, X, O, X, O, O,
                X, X, X, O, X, O, X,
                O, X, O, X, O, X, O,
                X, O, O, X, O, O, X,
                X, X, O, X, O, X, X], ended=1, winner=Z),                                    # c4_test.go:32-44
    dict(board=[X, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, Z, Z, Z,
                O, Z, Z, X, Z, Z, Z,
                X, Z, X, Z, Z, Z, Z,
                O, X, Z, Z, X, Z, X,
                X, O, Z, O, X, Z, X], ended=1, winner=X),                                    # c4_test.go:47-58
    dict(board=[X, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, Z, Z, Z,
                O, X, Z, O, Z, Z, Z,
                X, Z, X, Z, Z, Z, Z,
                O, X, Z, X, X, Z, X,
                X, O, Z, O, X, Z, X], ended=1, winner=X),                                    # c4_test.go:61-72
    dict(board=[X, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, X, Z, Z,
                O, Z, Z, X, X, Z, Z,
                X, Z, Z, Z, X, Z, Z,
                O, X, Z, Z, X, Z, X,
                X, O, Z, O, O, Z, X], ended=1, winner=X),                                    # c4_test.go:75-86
    dict(board=[X, Z, Z, Z, Z, Z, Z,
                O, Z, Z, Z, Z, Z, Z,
                O, Z, Z, X, Z, Z, Z,
                X, Z, Z, Z, X, Z, Z,
                O, X, Z, Z, X, Z, X,
                O, X, X, X, X, Z, X], ended=1, winner=X),                                    # c4_test.go:89-100
]

# wq Board.Apply: size, board, (player, move), board2 (None if error), taken, whiteScore, blackScore
WQ = [
    dict(size=3, board=[Z] * 9, player=X, move=4, board2=[Z, Z, Z, Z, X, Z, Z, Z, Z], taken=0, white=0, black=3,
         err=False),                                                                          # wq_test.go:43-58
    dict(size=3, board=[Z, O, Z, O, X, O, Z, Z, Z], player=O, move=7, board2=[Z, O, Z, O, Z, O, Z, O, Z], taken=1,
         white=6, black=0, err=False),                                                        # wq_test.go:70-85
    dict(size=4, board=[Z, O, Z, Z, O, X, O, Z, O, X, O, Z, Z, Z, Z, Z], player=O, move=13,
         board2=[Z, O, Z, Z, O, Z, O, Z, O, Z, O, Z, Z, O, Z, Z], taken=2, white=9, black=0, err=False),  # :100-117
    dict(size=4, board=[Z, Z, Z, Z, Z, Z, Z, Z, Z, X, X, Z, X, O, O, Z], player=X, move=15,
         board2=[Z, Z, Z, Z, Z, Z, Z, Z, Z, X, X, Z, X, Z, Z, X], taken=2, white=0, black=4, err=False),  # :132-149
    dict(size=3, board=[Z, O, Z, O, Z, O, Z, O, Z], player=X, move=4, board2=None, taken=0, err=True),    # :160-171 suicide
    dict(size=3, board=[Z] * 9, player=X, move=15, board2=None, taken=0, err=True),           # :174-184 off-board
    dict(size=3, board=[Z] * 9, player=Z, move=15, board2=None, taken=0, err=True),           # :187-197 bad colour
]

# dual.round KATs (dualnet/config_test.go:5-17)
ROUND = [(0, 0), (1, 1), (2, 2), (3, 4), (5, 4), (8, 8), (10, 8), (31, 32), (33, 32), (80, 64), (100, 128)]

# mcts Example (mcts/example_test.go:38-72): dummyNN keyed by MoveNumber; NB `8 / 9` is Go untyped
# integer division = 0.  Rows: (hot index, prob, value)
TTT_DUMMY_NN = [(4, 0.9, 0.5), (0, 0.1, 0.5), (2, 0.9, 0.0), (6, 0.1, 0.0), (3, 0.9, 0.0), (5, 0.1, 0.5),
                (1, 0.9, 0.0), (7, 0.1, 0.0), (8, 0.9, 0.0)]
# the move sequence the reference documents (example_test.go:107-152) and its pinned output
# "WINNER None" (example_test.go:154-155)
TTT_EXPECTED_MOVES = [4, 0, 2, 6, 3, 5, 1, 7, 8]
TTT_EXPECTED_WINNER = Z
