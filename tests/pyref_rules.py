"""A second, deliberately naive restatement of the reference's board rules in plain Python lists — written from the Go
sources independently of oracle/*.hpp, quirks kept — used only to cross-check the C++ oracle on random positions
(tests/test_oracle_rules_pyref.py).  Colours: 0 None, 1 Black/Cross, 2 White/Nought.  Moves: -1 pass."""


def opponent(p):
    return 2 if p == 1 else 1


# ---------------------------------------------------------------- game/mnk/mnk.go
def mnk_is_winner(board, m, n, k, colour):  # mnk.go:221-290
    for i in range(m):  # rows: ++ / -- without reset
        row_count = 0
        for j in range(n):
            row_count += 1 if board[i * n + j] == colour else -1
        if row_count >= k:
            return True
    for j in range(n):  # columns: only the run that reaches the bottom counts
        count = 0
        i = 0
        while i * n + j < len(board):
            count = count + 1 if board[i * n + j] == colour else 0
            i += 1
        if count >= k:
            return True
    for i in range(m):  # down-right diagonals, no wrap guard
        j = 0
        while n - j > n - k and j < n:
            idx, diag = i * n + j, 0
            while board[idx] == colour:
                diag += 1
                if diag >= k:
                    return True
                idx += n + 1
                if idx >= m * n:
                    break
            j += 1
    for i in range(m):  # down-left diagonals
        j = n - 1
        while j >= k - 1:
            idx, diag = i * n + j, 0
            while board[idx] == colour:
                diag += 1
                if diag >= k:
                    return True
                idx += n - 1
                if idx >= m * n:
                    break
            j -= 1
    return False


def mnk_check(board, move):  # mnk.go:96-115 (resign not exercised)
    if move == -1:
        return False
    if move >= len(board):
        return False
    return board[move] == 0


def mnk_apply(board, player, move):  # mnk.go:117-137
    ok = mnk_check(board, move)
    out = list(board)
    if ok:
        out[move] = player
    return ok, ok, out, 0


def mnk_status(board, m, n, k):  # mnk.go:142-169
    def score(p):
        if mnk_is_winner(board, m, n, k, p):
            return 1.0
        if mnk_is_winner(board, m, n, k, opponent(p)):
            return -2.0
        return 0.0
    if mnk_is_winner(board, m, n, k, 1):
        ended, winner = True, 1
    elif mnk_is_winner(board, m, n, k, 2):
        ended, winner = True, 2
    else:
        ended, winner = all(c != 0 for c in board), 0
    return ended, winner, score(1), score(2)


# ---------------------------------------------------------------- game/c4/c4.go, game/c4/game.go
def c4_drop_row(board, rows, cols, col):  # c4.go:59-70
    for row in range(rows - 1, -1, -1):
        if board[row * cols + col] == 0:
            return row
    return None


def c4_apply(board, rows, cols, player, move):  # c4.go:47-57, game.go:53-72
    out = list(board)
    if move == -1:
        return True, True, out, 0
    row = c4_drop_row(board, rows, cols, move)
    if row is None:
        return False, False, out, 0
    out[row * cols + move] = player
    return True, True, out, 0


def c4_check_win(board, rows, cols, nn):  # c4.go:82-192
    it = [[board[y * cols + x] for x in range(cols)] for y in range(rows)]

    def scan(dx, dy):
        for x in range(cols):
            for y in range(rows):
                c = it[y][x]
                if c == 0:
                    continue
                winning = True
                for i in range(nn):
                    xx, yy = x + dx * i, y + dy * i
                    if 0 <= xx < cols and yy < rows:
                        if it[yy][xx] != c:
                            winning = False
                    else:
                        winning = False
                if winning:
                    return c
        return 0
    for dx, dy in ((0, 1), (1, 0), (-1, 1), (1, 1)):  # vertical, horizontal, TLBR (x-i, y+i), TRBL (x+i, y+i)
        w = scan(dx, dy)
        if w:
            return w
    return 0


def c4_status(board, rows, cols, nn, pass_count):  # game.go:75-84, 161-179
    w = c4_check_win(board, rows, cols, nn)

    def score(p):
        return 1.0 if w == p else (0.0 if w == 0 else -1.0)
    if w:
        ended, winner = True, w
    elif pass_count > 2:
        ended, winner = True, 0
    else:
        ended, winner = all(c != 0 for c in board), 0
    return ended, winner, score(1), score(2)


# ---------------------------------------------------------------- game/wq/wq.go, game/wq/game.go
ADJ = ((0, 1), (1, 0), (0, -1), (-1, 0))  # wq.go:316-321, (X = row, Y = col)


def wq_nolib(it, size, c, potential):  # wq.go:237-290
    ret = []
    found, founds = True, [c]
    while found:
        found = False
        group = []
        for f in founds:
            for d in ADJ:
                a = (f[0] + d[0], f[1] + d[1])
                if not (0 <= a[0] < size and 0 <= a[1] < size):
                    continue
                if it[a[0]][a[1]] == 0 and a != potential:
                    return []
                if it[f[0]][f[1]] != it[a[0]][a[1]]:
                    continue
                if a in group or a in ret:
                    continue
                group.append(a)
                found = True
        ret.extend(founds)
        founds = group
    return ret


def wq_board_check(board, size, player, move):  # wq.go:205-234 -> (captures with duplicates, ok)
    it = [[board[x * size + y] for y in range(size)] for x in range(size)]
    c = (move // size, move % size)
    captures = []
    for d in ADJ:
        a = (c[0] + d[0], c[1] + d[1])
        if not (0 <= a[0] < size and 0 <= a[1] < size):
            continue
        if it[a[0]][a[1]] == opponent(player):
            for nl in wq_nolib(it, size, a, c):
                captures.append(nl[0] * size + nl[1])
    if captures:
        return captures, True
    if wq_nolib(it, size, c, (-5, -5)):
        return [], False
    return [], True


def wq_apply(board, size, player, move):  # game.go:65-92 + wq.go:141-171 (move in [0, size*size))
    out = list(board)
    caps, ok = wq_board_check(board, size, player, move)
    check = ok  # Game.Check: no occupancy test
    if board[move] != 0 or not ok:
        return check, False, out, 0
    out[move] = player
    for p in caps:
        out[p] = 0
    return check, True, out, len(caps) & 0xFF


def wq_score(board, size, player):  # wq.go:173-202, as implemented
    bd = [False] * len(board)
    q = []
    reachable = 0.0
    for i, c in enumerate(board):
        if c == player:
            reachable += 1
            bd[i] = True
            q.append(i)
    while q:
        i = q.pop(0)
        for adj in (-size, 1, size, 1):
            a = i + adj
            if a >= size or a < 0:
                continue
            if not bd[a] and board[a] == 0:
                reachable += 1
                bd[a] = True
                q.append(a)
    return reachable


# ---- wq under AZ_FLAG_WQ_COMPLETE (OUR completion: include/agogo_b200.h) — a naive second statement of the rules ----
def _adj4(size, p):
    r, c = divmod(p, size)
    return [p + 1 if c + 1 < size else None, p + size if r + 1 < size else None, p - 1 if c > 0 else None,
            p - size if r > 0 else None]


def _wq_group(board, size, p):
    colour = board[p]
    seen, stack, libs = {p}, [p], set()
    while stack:
        q = stack.pop()
        for a in _adj4(size, q):
            if a is None:
                continue
            if board[a] == 0:
                libs.add(a)
            elif board[a] == colour and a not in seen:
                seen.add(a); stack.append(a)
    return seen, libs


def wq_complete_check(board, size, player, move, ko=-1, positions=None):
    """-> (legal, captured points (set), ko point the move creates or -1).  `positions`: the earlier positions of the game
    (tuples of the board) — positional superko: none of them may be recreated."""
    if move < 0 or move >= size * size or board[move] != 0 or move == ko:
        return False, set(), -1
    opp = 3 - player
    nbrs = [a for a in _adj4(size, move) if a is not None]
    captured = set()
    for a in nbrs:
        if board[a] == opp:
            g, libs = _wq_group(board, size, a)
            if libs == {move}:
                captured |= g
    trial = list(board)
    trial[move] = player
    for q in captured:
        trial[q] = 0
    _, libs = _wq_group(trial, size, move)
    if not libs:
        return False, set(), -1  # suicide
    if all(board[a] == player for a in nbrs):
        return False, set(), -1  # own single-point eye: never filled
    if positions is not None and tuple(trial) in positions:
        return False, set(), -1  # positional superko
    mine, _ = _wq_group(trial, size, move)
    new_ko = next(iter(captured)) if (len(captured) == 1 and len(mine) == 1 and len(libs) == 1) else -1
    return True, captured, new_ko


def wq_area_score(board, size, player):
    total, seen = 0, set()
    for i in range(size * size):
        if board[i] == player:
            total += 1
        elif board[i] == 0 and i not in seen:
            region, stack, colours = {i}, [i], set()
            while stack:
                q = stack.pop()
                for a in _adj4(size, q):
                    if a is None:
                        continue
                    if board[a] == 0:
                        if a not in region:
                            region.add(a); stack.append(a)
                    else:
                        colours.add(board[a])
            seen |= region
            if colours == {player}:
                total += len(region)
    return float(total)
