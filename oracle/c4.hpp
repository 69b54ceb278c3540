// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU restatement of game/c4/c4.go + game/c4/game.go (Connect-N), quirks included:
// Apply never flips nextToMove, MoveNumber is the constant 1, Passes is the constant 0,
// Pass is a legal move, Clone pads history/historical by 2 entries.
#pragma once
#include "game.hpp"

namespace oracle {

struct C4 : State {
  int rows, cols, nwin;
  std::vector<int32_t> data;  // rows x cols, row-major (c4.go:19-33)
  std::vector<PlayerMove> history;
  std::vector<std::vector<int32_t>> historical;
  Player nextToMove = None;
  int histPtr = 0, moveCount = 0, passCount = 0;

  C4(int r, int c, int n) : rows(r), cols(c), nwin(n), data(r * c, None) {}  // c4/game.go:24-33

  int32_t at(int y, int x) const { return data[y * cols + x]; }

  // c4.go:59-70 — returns false on "column full"; row/col outputs as in the reference
  bool boardCheck(PlayerMove mv, int* row, int* col) const {
    if (mv.single == PassMove) { *row = -1; *col = -1; return true; }
    *col = (int)mv.single;
    if (*col < 0 || *col >= cols) throw std::runtime_error("c4: index out of range");  // Go panic
    for (*row = rows - 1; *row >= 0; (*row)--)
      if (at(*row, *col) == None) return true;
    *row = -1; *col = -1;
    return false;
  }
  bool boardApply(PlayerMove mv) {  // c4.go:47-57
    if (mv.single == PassMove) return true;
    int row, col;
    if (!boardCheck(mv, &row, &col)) return false;
    data[row * cols + col] = mv.player;
    return true;
  }

  // c4.go:72-192
  int32_t checkDir(int dx, int dy) const {
    for (int x = 0; x < cols; x++)
      for (int y = 0; y < rows; y++) {
        int32_t c = at(y, x);
        bool winning = true;
        if (c != None) {
          for (int i = 0; i < nwin; i++) {
            int xx = x + dx * i, yy = y + dy * i;
            if (xx >= 0 && xx < cols && yy < rows) {
              if (at(yy, xx) != c) winning = false;
            } else winning = false;
          }
          if (winning) return c;
        }
      }
    return None;
  }
  int32_t checkWin() const {
    int32_t w;
    if ((w = checkDir(0, 1)) != None) return w;   // vertical   c4.go:86-108
    if ((w = checkDir(1, 0)) != None) return w;   // horizontal c4.go:111-133
    if ((w = checkDir(-1, 1)) != None) return w;  // TLBR (x-i, y+i) c4.go:135-157
    return checkDir(1, 1);                        // TRBL (x+i, y+i) c4.go:159-181
  }

  void BoardSize(int* a, int* b) const override { *a = rows; *b = cols; }
  void SetToMove(Player p) override { nextToMove = p; }
  Player ToMove() const override { return nextToMove; }
  PlayerMove LastMove() const override {  // c4/game.go:41-46
    if (!history.empty()) return history.at(histPtr - 1);
    return PlayerMove{None, -1};
  }
  int Passes() const override { return 0; }                  // c4/game.go:49
  int MoveNumber() const override { return moveCount + 1; }  // c4/game.go:51
  bool Check(PlayerMove mv) const override {                 // c4/game.go:53
    int r, c;
    return boardCheck(mv, &r, &c);
  }
  State* Apply(PlayerMove mv) override {  // c4/game.go:55-72 — in place, no player flip
    std::vector<int32_t> hb = data;
    if (boardApply(mv)) {
      history.push_back(mv);
      historical.push_back(hb);
      histPtr++;
    }
    if (mv.single == PassMove) passCount++;
    else passCount = 0;
    return this;
  }
  float Score(Player p) const override {  // c4/game.go:74-83
    int32_t w = checkWin();
    if (w == p) return 1;
    if (w == None) return 0;
    return -1;
  }
  void UndoLastMove() override {  // c4/game.go:85-97 (buggy in the reference; kept)
    histPtr--;
    PlayerMove last = history.at(histPtr - 1);
    int col = (int)last.single;
    int row;
    for (row = rows - 1; row >= 0; row--)
      if (at(row, col) == None) { row--; break; }
    if (row < 0 || col < 0 || col >= cols) throw std::runtime_error("c4: index out of range");
    data[row * cols + col] = None;
  }
  void Fwd() override { if (!history.empty()) histPtr++; }
  bool Eq(const State* other) const override {  // c4/game.go:105-137
    const C4* ot = dynamic_cast<const C4*>(other);
    if (!ot) return false;
    // `!ot.b.data.Eq(ot.b.data)` compares other with itself: always equal
    if (histPtr != ot->histPtr) return false;
    if (moveCount != ot->moveCount) return false;
    if (history.size() != ot->history.size()) return false;
    if (historical.size() != ot->historical.size()) return false;
    for (size_t i = 0; i < history.size(); i++)
      if (ot->history[i].player != history[i].player || ot->history[i].single != history[i].single) return false;
    for (size_t i = 0; i < historical.size(); i++)
      for (size_t j = 0; j < historical[i].size(); j++) {
        if (j >= ot->historical[i].size()) throw std::runtime_error("c4: index out of range");
        if (ot->historical[i][j] != historical[i][j]) return false;
      }
    return true;
  }
  State* Clone() const override {  // c4/game.go:139-157 — history/historical padded by +2
    C4* r = new C4(rows, cols, nwin);
    r->data = data;
    r->history = history;
    r->history.resize(history.size() + 2, PlayerMove{0, 0});
    r->historical = historical;
    r->historical.resize(historical.size() + 2);
    r->nextToMove = nextToMove;
    r->histPtr = histPtr;
    r->moveCount = moveCount;
    r->passCount = passCount;
    return r;
  }
  float AdditionalScore() const override { return 0; }
  bool Ended(Player* winner) const override {  // c4/game.go:161-179
    int32_t w = checkWin();
    if (w != None) { *winner = w; return true; }
    if (passCount > 2) { *winner = None; return true; }
    for (int32_t c : data)
      if (c == None) { *winner = None; return false; }
    *winner = None;
    return true;
  }
  void Reset() override {  // c4/game.go:183-193
    for (auto& c : data) c = None;
    historical.clear();
    history.clear();
    histPtr = 0; moveCount = 0; passCount = 0; nextToMove = 0;
  }
  int ActionSpace() const override { return cols; }  // c4/game.go:195
  const std::vector<int32_t>& Board() const override { return data; }
  uint32_t Hash() const override { return fnv_board_hash(data); }
  const std::vector<int32_t>& Historical(int i) const override { return historical.at(i); }
};

}  // namespace oracle
