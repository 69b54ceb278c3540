// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU restatement of game/wq/wq.go (Board: Apply/check/nolib/Score, quirks included),
// game/wq/zobrist.go and game/wq/game.go.
//
// The reference's wq.Game cannot complete a game: Score/Reset/UndoLastMove/Fwd panic
// (game.go:117-121,178), applying a Pass indexes data[-1] (wq.go:146-151), `passes` and
// `historical` are never written (game.go:81-92).  Everything marked COMPLETION below is OUR
// minimal completion (documented in DESIGN.md §wq-gap), switchable only there; everything else
// follows the reference line by line.
#pragma once
#include <algorithm>

#include "game.hpp"
#include "rng.hpp"

namespace oracle {

struct Coord { int16_t X, Y; };  // game/state.go:92-94 ; X is the ROW here (wq.go:206-208)

struct WQBoard {
  int32_t size;
  std::vector<int32_t> data;
  std::shared_ptr<const std::vector<int32_t>> table;  // zobrist table [size*size][2] (zobrist.go:24-41)
  int32_t hash = 0;

  // zobrist.go:31-41: table[i] = r.Int31() from a time-seeded source.  Injected seed instead.
  WQBoard(int sz, uint64_t zseed) : size(sz), data(sz * sz, None) {
    auto t = std::make_shared<std::vector<int32_t>>(sz * sz * 2);
    Rng r(zseed);
    for (auto& v : *t) v = (int32_t)(r.next() >> 33);  // 31 bits, like Int31
    table = t;
  }
  WQBoard(const WQBoard&) = default;  // wq.go:74-86 Clone: copies data, table values and hash

  int32_t it(Coord c) const { return data[(int)c.X * size + c.Y]; }
  bool isCoordValid(Coord c) const {  // wq.go:304-315
    int32_t x = c.X, y = c.Y;
    if (x >= size || x < 0) return false;
    if (y >= size || y < 0) return false;
    return true;
  }
  static Coord adj(Coord c, int i) {  // wq.go:296-301, 317-322
    static const int16_t dx[4] = {0, 1, 0, -1}, dy[4] = {1, 0, -1, 0};
    return Coord{(int16_t)(c.X + dx[i]), (int16_t)(c.Y + dy[i])};
  }
  static bool ceq(Coord a, Coord b) { return a.X == b.X && a.Y == b.Y; }
  Single ltoi(Coord c) const { return (Single)((int32_t)c.X * size + (int32_t)c.Y); }

  void zupdate(PlayerMove m) {  // zobrist.go:44-56
    if (m.player == Black) hash ^= (*table)[m.single * 2 + 0];
    else if (m.player == White) hash ^= (*table)[m.single * 2 + 1];
  }

  // wq.go:237-290
  std::vector<Coord> nolib(Coord c, Coord potential) const {
    std::vector<Coord> retVal;
    bool found = true;
    std::vector<Coord> founds{c};
    while (found) {
      found = false;
      std::vector<Coord> group;
      for (Coord f : founds) {
        for (int i = 0; i < 4; i++) {
          Coord a = adj(f, i);
          if (!isCoordValid(a)) continue;
          if (it(a) == None && !ceq(a, potential)) return {};
          if (it(f) != it(a)) continue;
          bool potentialGroup = true;
          for (Coord g : group)
            if (ceq(g, a)) { potentialGroup = false; break; }
          if (potentialGroup)
            for (Coord l : retVal)
              if (ceq(l, a)) { potentialGroup = false; break; }
          if (potentialGroup) { group.push_back(a); found = true; }
        }
      }
      retVal.insert(retVal.end(), founds.begin(), founds.end());
      founds = group;
    }
    return retVal;
  }

  // wq.go:205-234 ; returns false on "Suicide is not a valid option"
  bool check(PlayerMove m, std::vector<Single>* captures) const {
    captures->clear();
    // Go integer division truncates toward zero, % keeps the dividend's sign
    Coord c{(int16_t)((int32_t)m.single / size), (int16_t)((int32_t)m.single % size)};
    for (int i = 0; i < 4; i++) {
      Coord a = adj(c, i);
      if (!isCoordValid(a)) continue;
      if (it(a) == opponent(m.player)) {
        for (Coord nl : nolib(a, c)) captures->push_back(ltoi(nl));
      }
    }
    if (!captures->empty()) return true;
    if (!isCoordValid(c)) throw std::runtime_error("wq: index out of range");  // b.it[f.X][f.Y] would panic
    std::vector<Coord> suicides = nolib(c, Coord{-5, -5});
    if (!suicides.empty()) return false;
    return true;
  }

  // wq.go:141-171 ; returns false on error, *taken = byte(len(captures))
  bool Apply(PlayerMove m, uint8_t* taken) {
    *taken = 0;
    if (!(m.player == Black || m.player == White)) return false;  // "Impossible player"
    if ((int32_t)m.single >= size * size) return false;           // "Impossible move"
    if (m.single < 0) throw std::runtime_error("wq: index out of range");  // b.data[-1] panics
    if (data[m.single] != None) return false;                     // "board location not empty"
    std::vector<Single> captures;
    if (!check(m, &captures)) return false;
    data[m.single] = m.player;
    zupdate(m);
    for (Single prisoner : captures) {
      data[prisoner] = None;
      zupdate(PlayerMove{opponent(m.player), prisoner});
    }
    *taken = (uint8_t)captures.size();
    return true;
  }

  // wq.go:173-202 — the reference's flood fill only ever expands inside row 0
  // (`a >= b.size` rejects everything else; adjacents are {-size, 1, size, 1}).  Kept.
  float Score(Player player) const {
    int32_t colour = player;
    int n = (int)data.size();
    std::vector<char> bd(n, 0);
    std::vector<int32_t> q;
    size_t qh = 0;
    const int32_t adjacents[4] = {-size, 1, size, 1};
    float reachable = 0;
    for (int32_t i = 0; i < n; i++)
      if (data[i] == colour) { reachable++; bd[i] = 1; q.push_back(i); }
    while (qh < q.size()) {
      int32_t i = q[qh++];
      for (int32_t ad : adjacents) {
        int32_t a = i + ad;
        if (a >= size || a < 0) continue;
        if (!bd[a] && data[a] == None) { reachable++; bd[a] = 1; q.push_back(a); }
      }
    }
    return reachable;
  }
};

struct WQ : State {
  WQBoard board;
  std::vector<PlayerMove> history;
  // COMPLETION: historical boards, filled by Apply (the reference never appends to it).  Kept as a
  // persistent list so that Clone stays O(1) like the reference's (which clones an empty slice).
  struct HistNode { std::vector<int32_t> board; std::shared_ptr<const HistNode> prev; int idx; };
  std::shared_ptr<const HistNode> hist;
  Player nextToMove = Black;                     // game.go:31
  float komi;
  int moveCount = 0, passes = 0, histPtr = 0, handicap = 0;
  uint8_t captures[2] = {0, 0};
  bool ends = false;
  // AZ_FLAG_WQ_COMPLETE (OURS, not the reference's: SURVEY §8f row 4): real Go rules instead of the reference's unfinished
  // ones — occupied points and true suicide are illegal, simple ko (the point of a single stone just captured by a lone
  // stone left with that single liberty may not be retaken immediately), own single-point eyes are never filled (the
  // "eye-ish situations" noPass expects Check to reject, search.go:543), POSITIONAL SUPERKO (game.go:77's TODO: a move may
  // not recreate the stones of any earlier position of the game — the `hist` list, compared board by board), area scoring
  // with komi decides the winner.
  bool complete = false;
  int32_t ko = -1;

  WQ(int boardSize, int handicap_, double komi_, uint64_t zseed)
      : board(boardSize, zseed), komi((float)komi_), handicap(handicap_) {}

  void BoardSize(int* a, int* b) const override { *a = board.size; *b = board.size; }
  const std::vector<int32_t>& Board() const override { return board.data; }
  const std::vector<int32_t>& Historical(int i) const override {
    const HistNode* n = hist.get();
    while (n && n->idx > i) n = n->prev.get();
    if (!n || n->idx != i) throw std::runtime_error("wq: index out of range");
    return n->board;
  }
  uint32_t Hash() const override { return (uint32_t)board.hash; }           // game.go:46
  int ActionSpace() const override { return (int)board.data.size(); }        // game.go:48
  void SetToMove(Player p) override { nextToMove = p; }
  Player ToMove() const override { return nextToMove; }
  PlayerMove LastMove() const override {  // game.go:54-59
    if (!history.empty()) return history.at(histPtr - 1);
    return PlayerMove{None, -1};
  }
  int Passes() const override { return passes; }
  int MoveNumber() const override { return (int)history.size(); }  // game.go:63

  bool Check(PlayerMove m) const override {  // game.go:65-79 — occupied points are NOT rejected
    if (m.single == ResignMove) return true;
    if (m.single == PassMove) return true;
    if ((int)m.single >= (int)board.data.size()) return false;
    if (complete) return completeCheck(m, nullptr, nullptr);
    std::vector<Single> caps;
    return board.check(m, &caps);
  }

  // ---- complete rules (OURS) ----
  // the group of the stone at p and the number of its distinct liberties
  void groupOf(int p, std::vector<int>* stones, int* libs) const {
    const int size = board.size, colour = board.data[p];
    std::vector<char> seen(board.data.size(), 0), libseen(board.data.size(), 0);
    stones->assign(1, p); seen[p] = 1; *libs = 0;
    for (size_t h = 0; h < stones->size(); h++) {
      const int q = (*stones)[h], r = q / size, c = q % size;
      const int nb[4] = {c + 1 < size ? q + 1 : -1, r + 1 < size ? q + size : -1, c > 0 ? q - 1 : -1, r > 0 ? q - size : -1};
      for (int a : nb) {
        if (a < 0) continue;
        if (board.data[a] == None) { if (!libseen[a]) { libseen[a] = 1; (*libs)++; } }
        else if (board.data[a] == colour && !seen[a]) { seen[a] = 1; stones->push_back(a); }
      }
    }
  }
  // legality of an on-board point; optionally the opponent stones it would capture (each once) and the ko point it creates
  bool completeCheck(PlayerMove m, std::vector<int>* captured, int32_t* new_ko) const {
    const int p = m.single, size = board.size;
    if (captured) captured->clear();
    if (new_ko) *new_ko = -1;
    if (p < 0 || board.data[p] != None || p == ko) return false;
    const int r = p / size, c = p % size;
    const int nb[4] = {c + 1 < size ? p + 1 : -1, r + 1 < size ? p + size : -1, c > 0 ? p - 1 : -1, r > 0 ? p - size : -1};
    const int opp = opponent(m.player);
    bool cap = false, empty_nbr = false, friend_safe = false, has_opp = false, has_friend = false;
    std::vector<int> caps;
    for (int a : nb) {
      if (a < 0) continue;
      if (board.data[a] == None) { empty_nbr = true; continue; }
      std::vector<int> g; int libs;
      groupOf(a, &g, &libs);
      if (board.data[a] == opp) {
        has_opp = true;
        if (libs == 1) {  // its only liberty is p
          cap = true;
          for (int st : g) if (std::find(caps.begin(), caps.end(), st) == caps.end()) caps.push_back(st);
        }
      } else {
        has_friend = true;
        if (libs >= 2) friend_safe = true;
      }
    }
    if (!empty_nbr && !has_opp) return false;                 // own single-point eye (or a 1x1 board): never filled
    if (!(cap || empty_nbr || friend_safe)) return false;     // suicide
    if (hist) {                                               // positional superko
      std::vector<int32_t> after = board.data;
      after[p] = m.player;
      for (int st : caps) after[st] = None;
      for (const HistNode* n = hist.get(); n; n = n->prev.get())
        if (n->board == after) return false;
    }
    if (captured) *captured = caps;
    if (new_ko && caps.size() == 1 && !has_friend && !empty_nbr) *new_ko = caps[0];
    return true;
  }
  float areaScore(Player player) const {  // Tromp-Taylor: stones + empty regions that touch only this colour
    const int size = board.size, n = (int)board.data.size();
    std::vector<char> seen(n, 0);
    float total = 0;
    for (int i = 0; i < n; i++) {
      if (board.data[i] == player) { total++; continue; }
      if (board.data[i] != None || seen[i]) continue;
      std::vector<int> region{i}; seen[i] = 1;
      int mask = 0;
      for (size_t h = 0; h < region.size(); h++) {
        const int q = region[h], r = q / size, c = q % size;
        const int nb[4] = {c + 1 < size ? q + 1 : -1, r + 1 < size ? q + size : -1, c > 0 ? q - 1 : -1, r > 0 ? q - size : -1};
        for (int a : nb) {
          if (a < 0) continue;
          if (board.data[a] == None) { if (!seen[a]) { seen[a] = 1; region.push_back(a); } }
          else mask |= board.data[a] == Black ? 1 : 2;
        }
      }
      if (mask == (player == Black ? 1 : 2)) total += (float)region.size();
    }
    return total;
  }

  State* Apply(PlayerMove m) override {  // game.go:81-92 — clones; Board.Apply's error is ignored
    WQ* ns = static_cast<WQ*>(Clone());
    {  // COMPLETION: record the board before the move (mnk.go:122-132 convention)
      auto hn = std::make_shared<HistNode>();
      hn->board = board.data; hn->prev = hist; hn->idx = hist ? hist->idx + 1 : 0;
      ns->hist = hn;
    }
    uint8_t caps = 0;
    if (m.single == PassMove) {
      ns->passes = passes + 1;  // COMPLETION: reference would panic on data[-1]
      ns->ko = -1;
    } else if (complete) {
      std::vector<int> captured; int32_t nk = -1;
      if ((int)m.single < (int)board.data.size() && completeCheck(m, &captured, &nk)) {
        ns->board.data[m.single] = m.player;
        ns->board.zupdate(m);
        for (int st : captured) { ns->board.data[st] = None; ns->board.zupdate(PlayerMove{opponent(m.player), (Single)st}); }
        caps = (uint8_t)captured.size();
      }
      ns->ko = nk;
      ns->passes = 0;
    } else {
      ns->board.Apply(m, &caps);
      ns->passes = 0;           // COMPLETION: passes counts consecutive passes
    }
    if (m.player != Black && m.player != White) throw std::runtime_error("wq: index out of range");
    ns->captures[m.player - 1] += caps;
    ns->nextToMove = opponent(m.player);
    ns->history.push_back(m);
    ns->histPtr++;
    ns->moveCount++;
    return ns;
  }

  bool Ended(Player* winner) const override {  // game.go:94-115
    bool ended = false;
    if (passes >= 2) ended = true;
    if (ends) ended = true;
    if (!ended) { *winner = None; return false; }
    float whiteScore = Score(White), blackScore = Score(Black);
    if (complete) whiteScore += komi;  // OURS: the reference compares the raw scores and leaves komi to combinedScore
    if (whiteScore == blackScore) *winner = None;
    else if (whiteScore > blackScore) *winner = White;
    else *winner = Black;
    return true;
  }

  void Reset() override {  // COMPLETION (reference panics): back to New()'s state
    for (auto& c : board.data) c = None;
    board.hash = 0;  // wq.go:132-137
    history.clear(); hist.reset(); ko = -1;
    nextToMove = Black; moveCount = 0; passes = 0; histPtr = 0;
    captures[0] = captures[1] = 0; ends = false;
  }
  bool SupportsUndo() const override { return false; }
  bool CompleteRules() const override { return complete; }
  void UndoLastMove() override { throw std::runtime_error("not implemented"); }  // game.go:119
  void Fwd() override { throw std::runtime_error("not implemented"); }           // game.go:121

  bool Eq(const State* other) const override {  // game.go:123-160 (only used by tree reuse, which wq never reaches)
    const WQ* ot = dynamic_cast<const WQ*>(other);
    if (!ot) return false;
    if (nextToMove != ot->nextToMove || komi != ot->komi || moveCount != ot->moveCount ||
        passes != ot->passes || handicap != ot->handicap)
      return false;
    if (captures[0] != ot->captures[0] || captures[1] != ot->captures[1]) return false;
    if (board.size != ot->board.size || board.hash != ot->board.hash || board.data != ot->board.data) return false;
    for (int i = 0, j = 0; i < histPtr && j < ot->histPtr; i++, j++)
      if (history[i].player != ot->history[j].player || history[i].single != ot->history[j].single) return false;
    return true;
  }

  State* Clone() const override {  // game.go:162-175 (+ COMPLETION: historical travels with the clone)
    WQ* ns = new WQ(*this);
    ns->ends = false;  // Clone does not copy `ends`
    return ns;
  }

  float Score(Player p) const override { return complete ? areaScore(p) : board.Score(p); }  // COMPLETION: Board.Score as implemented
  float AdditionalScore() const override { return komi; }
};

}  // namespace oracle
