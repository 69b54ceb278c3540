// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU restatement of game/mnk/mnk.go (m,n,k games / tic-tac-toe), quirks included.
#pragma once
#include "game.hpp"

namespace oracle {

struct MNK : State {
  std::vector<int32_t> board;
  int m, n, k;
  Player nextToMove = None;
  std::vector<PlayerMove> history;
  std::vector<std::vector<int32_t>> historical;
  int histPtr = 0;

  MNK(int m_, int n_, int k_) : board(m_ * n_, None), m(m_), n(n_), k(k_) {}  // mnk.go:35-44

  void BoardSize(int* a, int* b) const override { *a = m; *b = n; }
  const std::vector<int32_t>& Board() const override { return board; }
  const std::vector<int32_t>& Historical(int i) const override { return historical.at(i); }
  uint32_t Hash() const override { return fnv_board_hash(board); }  // mnk.go:70-76
  int ActionSpace() const override { return m * n; }                // mnk.go:78
  void SetToMove(Player p) override { nextToMove = p; }
  Player ToMove() const override { return nextToMove; }
  PlayerMove LastMove() const override {  // mnk.go:84-89
    if (!history.empty()) return history.at(histPtr - 1);
    return PlayerMove{None, PassMove};
  }
  int Passes() const override { return -1; }                        // mnk.go:92
  int MoveNumber() const override { return (int)history.size(); }   // mnk.go:94

  bool Check(PlayerMove mv) const override {  // mnk.go:96-115
    if (mv.single == ResignMove) return true;
    if (mv.single == PassMove) return false;
    if ((int)mv.single >= (int)board.size()) return false;
    if (board[mv.single] != None) return false;  // (negative index other than -1/-2 would panic in Go)
    return true;
  }

  State* Apply(PlayerMove mv) override {  // mnk.go:117-137 — mutates in place, returns self
    if (!Check(mv)) return this;
    if (mv.single < 0) throw std::runtime_error("index out of range");  // Resign passes Check, then board[-2] panics
    std::vector<int32_t> hb = board;
    board[mv.single] = mv.player;
    histPtr++;
    if ((int)history.size() < histPtr) history.push_back(mv);
    else history[histPtr - 1] = mv;
    historical.push_back(hb);
    nextToMove = opponent(mv.player);
    return this;
  }

  float Score(Player p) const override {  // mnk.go:142-150
    if (isWinner(p)) return 1;
    if (isWinner(opponent(p))) return -2;
    return 0;
  }
  float AdditionalScore() const override { return 0; }

  bool Ended(Player* winner) const override {  // mnk.go:156-169
    if (isWinner(Black)) { *winner = Black; return true; }
    if (isWinner(White)) { *winner = White; return true; }
    for (int32_t c : board)
      if (c == None) { *winner = None; return false; }
    *winner = None;
    return true;
  }

  void Reset() override {  // mnk.go:171-177 (historical and nextToMove are NOT reset)
    for (auto& c : board) c = None;
    history.clear();
    histPtr = 0;
  }

  void UndoLastMove() override {  // mnk.go:179-184
    if (!history.empty()) {
      board[history.at(histPtr - 1).single] = None;
      histPtr--;
    }
  }
  void Fwd() override {  // mnk.go:186-190
    if (!history.empty()) histPtr++;
  }

  bool Eq(const State* other) const override {  // mnk.go:192-206 — boards only
    const MNK* ot = dynamic_cast<const MNK*>(other);
    if (!ot) return false;
    if (board.size() != ot->board.size()) return false;
    for (size_t i = 0; i < board.size(); i++)
      if (board[i] != ot->board[i]) return false;
    return true;
  }

  State* Clone() const override {  // mnk.go:208-219 (historical copy copies 0 elements)
    MNK* r = new MNK(m, n, k);
    r->board = board;
    r->history = history;
    r->nextToMove = nextToMove;
    r->histPtr = histPtr;
    return r;
  }

  // mnk.go:221-290 — kept verbatim in behaviour: the row test never resets its counter, the
  // diagonal walks have no column-wrap guard.
  bool isWinner(Player p) const {
    int32_t colour = p;
    for (int i = 0; i < m; i++) {
      int rowCount = 0;
      for (int j = 0; j < n; j++) {
        if (board[i * n + j] == colour) rowCount++;
        else rowCount--;
      }
      if (rowCount >= k) return true;
    }
    for (int j = 0; j < n; j++) {
      int count = 0;
      for (int i = 0; i * n + j < (int)board.size(); i++) {
        if (board[i * n + j] == colour) count++;
        else count = 0;
      }
      if (count >= k) return true;
    }
    for (int i = 0; i < m; i++) {
      for (int j = 0; n - j > n - k && j < n; j++) {
        int idx = i * n + j;
        int diagCount = 0;
        while (board[idx] == colour) {
          diagCount++;
          if (diagCount >= k) return true;
          idx = idx + n + 1;
          if (idx >= m * n) break;
        }
      }
    }
    for (int i = 0; i < m; i++) {
      for (int j = n - 1; j >= k - 1; j--) {
        int idx = i * n + j;
        int diagCount = 0;
        while (board[idx] == colour) {
          diagCount++;
          if (diagCount >= k) return true;
          idx = idx + n - 1;
          if (idx >= m * n) break;
        }
      }
    }
    return false;
  }
};

}  // namespace oracle
