// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU restatement of the reference's API layer on the self-play path: arena.go (Arena.Play
// 80-179, newB 205-224), agent.go (Agent 14-121), agogo.go (AZ.New 41-73, setupSelfPlay 75-90,
// SelfPlay 93-97, Learn 100-172, prepareExamples 211-249, shuffleExamples 251-257), dummy.go,
// encoding_helper.go (WQEncoder 29-68) and cmd/tictactoe/main.go:26-47 (two-plane encoder).
// Time-seeded RNGs (arena.go:61, agogo.go:252) are replaced by injected splitmix64 streams.
#pragma once
#include <functional>

#include "c4.hpp"
#include "dual.hpp"
#include "mcts.hpp"
#include "mnk.hpp"
#include "wq.hpp"

namespace oracle {

// ---- encoders ----
inline void EncodeTwoPlayerBoard(const std::vector<int32_t>& a, float* out) {  // encoding_helper.go:10-26
  for (size_t i = 0; i < a.size(); i++) out[i] = a[i] == Black ? 1.0f : (a[i] == White ? -1.0f : 0.0f);
}
inline std::vector<float> encodeBoard2(const State& a) {  // cmd/tictactoe/main.go:26-47
  const std::vector<int32_t>& b = a.Board();
  size_t n = b.size();
  std::vector<float> r(2 * n, 0.0f);
  EncodeTwoPlayerBoard(b, r.data());
  for (size_t i = 0; i < n; i++)
    if (r[i] == 0) r[i] = 0.001f;
  Player next = a.ToMove();
  if (next == Black) for (size_t i = 0; i < n; i++) r[n + i] = 1;
  else if (next == White) for (size_t i = 0; i < n; i++) r[n + i] = -1;
  return r;
}
inline std::vector<float> WQEncoder(const State& a) {  // encoding_helper.go:29-68
  const int lookback = 8, features = 2 * lookback + 2;
  const std::vector<int32_t>& board = a.Board();
  int size = (int)board.size();
  std::vector<float> r((size_t)size * features, 0.0f);
  Player next = a.ToMove();
  float encodedPlayer = 1;
  int blackStart, whiteStart, nextStart;
  if (next == Black) { blackStart = 0; whiteStart = lookback * size; nextStart = 2 * lookback * size; }
  else { blackStart = lookback * size; whiteStart = 0; nextStart = (2 * lookback + 1) * size; encodedPlayer = -1; }
  int current = a.MoveNumber() - 1;
  for (int i = 1; i < lookback; i++) {
    int h = current - i;
    if (h > 0 && h < current) {
      const std::vector<int32_t>& past = a.Historical(h);
      EncodeTwoPlayerBoard(past, r.data() + blackStart);                    // encodeBlack
      EncodeTwoPlayerBoard(past, r.data() + whiteStart);                    // encodeWhite:
      for (int j = 0; j < size; j++) r[whiteStart + j] = r[whiteStart + j] * -1.0f;  // vecf32.Scale(-1): 0 -> -0
    }
    blackStart += size; whiteStart += size;
  }
  for (int i = nextStart; i < nextStart + size; i++) r[i] = encodedPlayer;
  return r;
}
enum EncoderKind { ENC_TWO_PLANE = 0, ENC_WQ18 = 1 };
inline std::vector<float> encode(int kind, const State& s) { return kind == ENC_WQ18 ? WQEncoder(s) : encodeBoard2(s); }

// ---- inferers (datatypes.go:51-59) ----
struct Inferer {
  virtual ~Inferer() {}
  virtual void Infer(const std::vector<float>& a, const State& st, std::vector<float>* policy, float* value) = 0;
};
struct DummyInferer : Inferer {  // dummy.go
  int outputSize; Player currentPlayer;
  DummyInferer(int o, Player p) : outputSize(o), currentPlayer(p) {}
  void Infer(const std::vector<float>&, const State&, std::vector<float>* policy, float* value) override {
    *value = 0;
    if (currentPlayer == 1) *value = 1;
    else if (currentPlayer == 2) *value = -1;
    policy->assign(outputSize, 1 / (float)outputSize);
  }
};
// scripted table keyed by state.MoveNumber() (mcts/example_test.go:38-72 dummyNN shape)
struct TableInferer : Inferer {
  std::vector<std::vector<float>> rows; std::vector<float> values;
  void Infer(const std::vector<float>&, const State& st, std::vector<float>* policy, float* value) override {
    int mn = st.MoveNumber();
    if (mn < 0 || mn >= (int)rows.size()) { policy->assign(rows.empty() ? 1 : rows[0].size(), 0.0f); *value = 0; return; }
    *policy = rows[mn]; *value = values[mn];
  }
};
// dual.Infer / Inferencer.Infer (meta.go:125-190): snapshot of the weights at creation time
struct DualInferer : Inferer {
  Dual net;
  explicit DualInferer(const Dual& d) : net(d) {}
  void Infer(const std::vector<float>& a, const State&, std::vector<float>* policy, float* value) override {
    policy->resize(net.conf.ActionSpace);
    dual_infer(net, a.data(), 1, policy->data(), value);
  }
};

// ---- Agent (agent.go:14-121) ----
struct Agent : Inferencer {
  std::shared_ptr<Dual> NN;
  std::shared_ptr<MCTS> mcts;  // shared_ptr only so that the single-tree Example mode can alias it
  Player player = None;
  int enc = ENC_TWO_PLANE;
  float Wins = 0, Loss = 0, Draw = 0;
  std::shared_ptr<Inferer> inferer;
  void SwitchToInference() { inferer = std::make_shared<DualInferer>(*NN); }  // agent.go:42-57
  void useDummy(const State& g) { inferer = std::make_shared<DummyInferer>(g.ActionSpace(), player); }  // agent.go:105-113
  void Infer(const State& g, std::vector<float>* policy, float* value) override {  // agent.go:60-74
    std::vector<float> input = encode(enc, g);
    inferer->Infer(input, g, policy, value);
  }
  Single Search(State* g) { mcts->SetGame(g); return mcts->Search(player); }  // agent.go:77-80
  void resetStats() { Wins = Loss = Draw = 0; }
};

struct Example { std::vector<float> Board, Policy; float Value; };  // datatypes.go:38-42

inline bool validPolicies(const std::vector<float>& p) {  // arena.go:241-251
  for (float v : p) if (std::isinf(v) || std::isnan(v)) return false;
  return true;
}

struct GameRecord { std::vector<int32_t> moves; int32_t winner = 0; int32_t a_player = 0; int32_t n_examples = 0; };

// ---- Arena (arena.go:20-179) ----
struct Arena {
  Rng r;
  State* game;  // owned
  Agent A, B;
  Agent* currentPlayer = nullptr;
  MCTSConfig conf;
  uint64_t tree_seed;
  uint64_t games = 0;  // games this Arena has started: every tree gets its own MCTS.rand stream (tree.go:84 seeds each from the clock)
  int max_moves = 0;  // COMPLETION: 0 = unlimited (reference has no cap; wq games need one to end)
  bool shared_tree = false;  // mcts/example_test.go:74-156 usage: ONE MCTS searched by both colours
  std::vector<GameRecord> records;

  Arena(State* g, std::shared_ptr<Dual> a, std::shared_ptr<Dual> b, const MCTSConfig& c, int enc, uint64_t seed)
      : r(derive_seed(seed, 0)), game(g), conf(c), tree_seed(derive_seed(seed, 1)) {  // arena.go:42-71
    A.NN = a; A.enc = enc; B.NN = b; B.enc = enc;
    newTrees();
  }
  ~Arena() { delete game; }
  void newTrees() {  // tree t of game g: stream derive_seed(tree_seed, 2 g + t)
    A.mcts.reset(new MCTS(game, conf, &A, derive_seed(tree_seed, 2 * games)));
    if (shared_tree) B.mcts = A.mcts;
    else B.mcts.reset(new MCTS(game, conf, &B, derive_seed(tree_seed, 2 * games + 1)));
  }

  void switchPlayer() { currentPlayer = currentPlayer == &A ? &B : &A; }

  // arena.go:80-179, split at the loop boundaries so the C API can run many games in lockstep;
  // Play() is exactly playBegin + playStep* + playFinish.  (The reference always returns None
  // as the winner, arena.go:178; the real winner is kept in the GameRecord.)
  bool active = false, record = false, lastEnded = false;  // lastEnded: the loop condition's last Ended()
  Player winner = None;
  int passCount = 0;
  GameRecord rec;
  std::vector<Example> examples;

  void playBegin(int coin, bool record_) {  // arena.go:81-96
    rec = GameRecord(); examples.clear(); winner = None; passCount = 0; record = record_;
    if (coin == 0) { A.player = Black; B.player = White; currentPlayer = &A; }
    else { A.player = White; B.player = Black; currentPlayer = &B; }
    rec.a_player = A.player;
    game->SetToMove(currentPlayer->player);
    lastEnded = game->Ended(&winner);
    active = !lastEnded;
  }
  void searchBegin() { currentPlayer->mcts->SetGame(game); currentPlayer->mcts->SearchBegin(currentPlayer->player); }
  void searchRun(int n) { currentPlayer->mcts->SearchRun(n); }
  void searchEnd() {  // arena.go:98-137
    Single best = currentPlayer->mcts->SearchEnd();
    if (best == PassMove) passCount++; else passCount = 0;
    if (record) {
      Example ex;
      ex.Board = encode(currentPlayer->enc, *game);
      ex.Policy = currentPlayer->mcts->Policies(*game);
      ex.Value = (float)currentPlayer->player;
      if (validPolicies(ex.Policy)) examples.push_back(ex);
    }
    game = apply_replace(game, PlayerMove{currentPlayer->player, best}, true);
    rec.moves.push_back(best);
    switchPlayer();
    if (passCount >= 2) {  // arena.go:135-137 breaks before the loop condition re-evaluates Ended(): the winner stays None
      active = false; lastEnded = false;
      if (game->CompleteRules()) lastEnded = game->Ended(&winner);  // OUR complete-rules mode: two passes end the game, score it
      return;
    }
    if (max_moves > 0 && (int)rec.moves.size() >= max_moves) { active = false; lastEnded = false; return; }  // COMPLETION
    lastEnded = game->Ended(&winner);
    active = !lastEnded;
  }
  bool playStep() { searchBegin(); searchRun(conf.Sims); searchEnd(); return active; }
  std::vector<Example> playFinish() {  // arena.go:139-178
    A.mcts->Reset(); B.mcts->Reset();
    for (auto& ex : examples) {  // arena.go:146-155
      if (winner == None) ex.Value = 0;
      else if (ex.Value == (float)winner) ex.Value = 1;
      else ex.Value = -1;
    }
    if (winner == None) { A.Draw++; B.Draw++; }
    else if (winner == A.player) { A.Wins++; B.Loss++; }
    else if (winner == B.player) { B.Wins++; A.Loss++; }
    rec.winner = winner; rec.n_examples = (int)examples.size();
    records.push_back(rec);
    games++;
    newTrees();
    return examples;
  }
  std::vector<Example> Play(bool record_) {
    playBegin(r.intn(2), record_);
    while (active) playStep();
    return playFinish();
  }
};

// ---- AZ (agogo.go) ----
struct AZConfig {
  DualConfig nn; MCTSConfig mcts;
  double UpdateThreshold = 0; int MaxExamples = 0; int enc = ENC_TWO_PLANE;
};
struct EpochLog { float a_wins, a_loss, a_draw, b_wins, b_loss, b_draw; int n_examples, batches, promoted; float first_cost, last_cost; };

struct AZ {
  AZConfig conf;
  std::unique_ptr<Arena> arena;
  bool useDummyFlag = true;
  uint64_t seed;
  int epoch = 0;
  std::vector<EpochLog> log;
  std::vector<Example> last_examples;

  AZ(State* g, const AZConfig& c, uint64_t seed_) : conf(c), seed(seed_) {  // agogo.go:41-73
    if (!c.nn.IsValid()) throw std::runtime_error("NNConf is not valid. Unable to proceed");
    if (!c.mcts.IsValid()) throw std::runtime_error("MCTSConf is not valid. Unable to proceed");
    auto a = std::make_shared<Dual>(c.nn); a->Init(derive_seed(seed, 100));
    auto b = std::make_shared<Dual>(c.nn); b->Init(derive_seed(seed, 101));
    arena.reset(new Arena(g, a, b, c.mcts, c.enc, derive_seed(seed, 102)));
  }
  void setupSelfPlay(int iter) {  // agogo.go:75-90
    arena->A.SwitchToInference(); arena->B.SwitchToInference();
    if (iter == 0 && useDummyFlag) { arena->A.useDummy(*arena->game); arena->B.useDummy(*arena->game); }
  }
  std::vector<Example> SelfPlay() {  // agogo.go:93-97
    std::vector<Example> ex = arena->Play(true);
    arena->game->Reset();
    return ex;
  }
  static void shuffleExamples(std::vector<Example>& ex, uint64_t s) {  // agogo.go:251-257
    Rng r(s);
    for (size_t i = 0; i < ex.size(); i++) { int j = r.intn((int)i + 1); std::swap(ex[i], ex[j]); }
  }
  // agogo.go:100-172
  void Learn(int iters, int episodes, int nniters, int arenaGames) {
    for (epoch = 0; epoch < iters; epoch++) {
      EpochLog el{};
      std::vector<Example> ex;
      setupSelfPlay(epoch);
      for (int e = 0; e < episodes; e++) { auto x = SelfPlay(); ex.insert(ex.end(), x.begin(), x.end()); }
      if (conf.MaxExamples > 0 && (int)ex.size() > conf.MaxExamples) {
        shuffleExamples(ex, derive_seed(seed, 1000 + 10 * epoch));
        ex.resize(conf.MaxExamples);
      }
      // prepareExamples (agogo.go:211-249)
      shuffleExamples(ex, derive_seed(seed, 1001 + 10 * epoch));
      last_examples = ex;
      int batches = (int)ex.size() / conf.nn.BatchSize;
      int total = batches * conf.nn.BatchSize;
      el.n_examples = (int)ex.size(); el.batches = batches;
      if (batches == 0) throw std::runtime_error("batches is nil, probably too few examples regarding the batchsize");
      std::vector<float> Xs, Pi, V;
      for (int i = 0; i < total; i++) {
        Xs.insert(Xs.end(), ex[i].Board.begin(), ex[i].Board.end());
        Pi.insert(Pi.end(), ex[i].Policy.begin(), ex[i].Policy.end());
        V.push_back(ex[i].Value);
      }
      Rng tr(derive_seed(seed, 1002 + 10 * epoch));
      std::vector<float> costs;
      dual_train(*arena->B.NN, Xs, Pi, V, batches, nniters, 0.1f, &tr, &costs);
      if (!costs.empty()) { el.first_cost = costs.front(); el.last_cost = costs.back(); }
      arena->B.SwitchToInference();
      arena->A.resetStats(); arena->B.resetStats();
      for (int g = 0; g < arenaGames; g++) { arena->Play(false); arena->game->Reset(); }
      bool killedA = false;
      Agent &A = arena->A, &B = arena->B;
      el.a_wins = A.Wins; el.a_loss = A.Loss; el.a_draw = A.Draw; el.b_wins = B.Wins; el.b_loss = B.Loss; el.b_draw = B.Draw;
      if (B.Wins / (B.Wins + A.Wins) > (float)conf.UpdateThreshold) {  // NaN when 0/0: no promotion
        A.NN = B.NN;
        killedA = true;
      }
      el.promoted = killedA;
      // newB (arena.go:205-224): fresh random net every epoch
      B.NN = std::make_shared<Dual>(conf.nn);
      B.NN->Init(derive_seed(seed, 200 + epoch));
      log.push_back(el);
    }
  }
};

}  // namespace oracle
