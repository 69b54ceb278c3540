// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of gorgonia/agogo's game contract (game/state.go).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// The product path (agogo_b200/csrc) never links or calls anything in oracle/.
//
// Follows: game/state.go:7-156 (Colour, Player, PlayerMove, Single, State).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// game/state.go:9-13
enum Colour : int32_t { None = 0, Black = 1, White = 2 };
typedef int32_t Player;  // game/state.go:41 (Player is a Colour)
typedef int32_t Single;  // game/state.go:110-123
static const Single PassMove = -1;    // mcts/mcts.go:21
static const Single ResignMove = -2;  // mcts/mcts.go:22

struct PlayerMove {  // game/state.go:69-72
  Player player;
  Single single;
};

inline Player opponent(Player p) {  // mcts/search.go:26-34 (panics otherwise)
  if (p == Black) return White;
  if (p == White) return Black;
  throw std::runtime_error("Unreachable");
}

// game/state.go:125-156.  The reference panics where marked; we throw std::runtime_error.
struct State {
  virtual ~State() {}
  virtual void BoardSize(int* m, int* n) const = 0;
  virtual const std::vector<int32_t>& Board() const = 0;
  virtual int ActionSpace() const = 0;
  virtual uint32_t Hash() const = 0;
  virtual Player ToMove() const = 0;
  virtual int Passes() const = 0;
  virtual int MoveNumber() const = 0;
  virtual PlayerMove LastMove() const = 0;
  virtual float Score(Player p) const = 0;
  virtual float AdditionalScore() const = 0;
  virtual bool Ended(Player* winner) const = 0;
  virtual void SetToMove(Player p) = 0;
  virtual bool Check(PlayerMove m) const = 0;
  // Apply returns the resulting state: either `this` (mnk, c4 mutate in place) or a fresh
  // object (wq clones).  Callers own the pointer semantics through StatePtr below.
  virtual State* Apply(PlayerMove m) = 0;
  virtual void Reset() = 0;
  virtual const std::vector<int32_t>& Historical(int i) const = 0;
  virtual void UndoLastMove() = 0;
  virtual void Fwd() = 0;
  virtual bool Eq(const State* other) const = 0;
  virtual State* Clone() const = 0;
  // not part of the Go interface: true iff UndoLastMove/Fwd are usable (wq's panic).
  virtual bool SupportsUndo() const { return true; }
  virtual bool CompleteRules() const { return false; }  // wq under AZ_FLAG_WQ_COMPLETE (OUR mode)
};

// Helper mirroring Go's `x = x.Apply(m).(game.State)`: returns the new pointer and frees the
// old object iff Apply produced a different one and `owned` is true.
inline State* apply_replace(State* s, PlayerMove m, bool owned) {
  State* n = s->Apply(m);
  if (n != s && owned) delete s;
  return n;
}

// FNV-1a 32 over the decimal-free "%v" rendering of each colour (mnk.go:70-76, c4/game.go:203-210):
// fmt.Fprintf(h, "%v", colour) prints "None" / "Black" / "White".
inline uint32_t fnv_board_hash(const std::vector<int32_t>& b) {
  uint32_t h = 2166136261u;
  for (int32_t v : b) {
    const char* s = v == Black ? "Black" : (v == White ? "White" : "None");
    for (const char* p = s; *p; ++p) {
      h ^= (uint8_t)*p;
      h *= 16777619u;
    }
  }
  return h;
}

}  // namespace oracle
