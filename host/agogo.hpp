// agogo.hpp — C++ host layer over the C ABI (include/agogo_b200.h), mirroring the reference's Go
// API layer for the self-play path: agogo.Config / agogo.New / AZ.SelfPlay / AZ.Learn / AZ.Save /
// AZ.Load (agogo.go), Agent.SwitchToInference / useDummy (agent.go), with the reference's names,
// argument meaning and error behaviour (panics become std::runtime_error).  The reference's host is
// Go; there is no Go toolchain in this image, so the host above the C ABI is provided in C++ (this
// file, single process) and Python (agogo_b200/host.py, also multi-GPU).  Header-only; link against
// libagogo_b200.so.  Time-seeded RNGs of the reference are replaced by the injected splitmix64
// streams documented in DESIGN.md §2 (identical to host.py and to the oracle).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/agogo_b200.h"

namespace agogo {

inline uint64_t splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t derive_seed(uint64_t seed, uint64_t stream) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (stream + 1));
  return splitmix64(&s);
}

// dual.DefaultConf (dualnet/config.go:18-31) and its `round`
inline int dual_round(int a) {
  int n = a - 1;
  n |= n >> 1; n |= n >> 2; n |= n >> 4; n |= n >> 8; n |= n >> 16;
  n++;
  int lt = n / 2;
  return (a - lt) < (n - a) ? lt : n;
}
inline az_dual_config DefaultConf(int m, int n, int actionSpace) {
  az_dual_config c{};
  int k = dual_round((m * n) / 3);
  c.k = k; c.shared_layers = m; c.fc = 2 * k; c.batch_size = 256; c.width = n; c.height = m; c.features = 18;
  c.action_space = actionSpace;
  return c;
}
// mcts.DefaultConfig (mcts/tree.go:31-41)
inline az_mcts_config DefaultConfig(int boardSize) {
  az_mcts_config c{};
  c.puct = 1.0f; c.timeout_ns = 100000000; c.m = boardSize; c.n = boardSize; c.dumb_pass = 1; c.pass_preference = 0;
  c.budget = 10000; c.sims = 100;
  return c;
}

struct Example { std::vector<float> Board, Policy; float Value; };  // datatypes.go:38-42

// agogo.Config (datatypes.go:14-25)
struct Config {
  std::string Name;
  az_dual_config NNConf{};
  az_mcts_config MCTSConf{};
  double UpdateThreshold = 0;
  int MaxExamples = 0;
  int Encoder = AZ_ENC_TWO_PLANE;
};

struct EpochLog { float a[3], b[3]; int n_examples, batches; bool promoted; float first_cost, last_cost; };

class AZ {
 public:
  // agogo.New (agogo.go:41-73): invalid configs are a panic in the reference
  AZ(const az_game_desc& game, const Config& conf, int n_games, uint64_t seed, int device = 0, uint32_t flags = 0)
      : conf_(conf), seed_(seed) {
    az_engine_desc d{};
    d.game = game; d.mcts = conf.MCTSConf; d.nn = conf.NNConf; d.encoder = conf.Encoder; d.n_games = n_games;
    d.device = device; d.flags = flags; d.seed = derive_seed(seed, 102);
    if (az_engine_create(&d, &e_) != AZ_OK) throw std::runtime_error(az_last_error(nullptr));
    cells_ = game.m * game.n;
    A_ = game.kind == AZ_GAME_C4 ? game.n : cells_;
    plane_ = conf.NNConf.features * cells_;
    ck(az_net_init(e_, 0, derive_seed(seed, 100)));
    ck(az_net_init(e_, 1, derive_seed(seed, 101)));
  }
  ~AZ() { az_engine_destroy(e_); }
  AZ(const AZ&) = delete;

  // setupSelfPlay (agogo.go:75-90)
  void setupSelfPlay(int iter) {
    ck(az_agent_set_inferer(e_, 0, AZ_INF_DUAL, 0));
    ck(az_agent_set_inferer(e_, 1, AZ_INF_DUAL, 0));
    if (iter == 0 && useDummy_) {
      ck(az_agent_set_inferer(e_, 0, AZ_INF_DUMMY, a_player_));
      ck(az_agent_set_inferer(e_, 1, AZ_INF_DUMMY, b_player_));
    }
  }
  // n x (Arena.Play(record) ; game.Reset()) (agogo.go:93-97, 144-148), concurrently on the device
  std::vector<Example> Play(int n, bool record) {
    ck(az_examples_clear(e_));
    ck(az_arena_play(e_, n, record ? 1 : 0));
    int32_t ap = 0;
    ck(az_game_record(e_, n - 1, nullptr, 0, nullptr, nullptr, &ap, nullptr));
    a_player_ = ap; b_player_ = ap == AZ_BLACK ? AZ_WHITE : AZ_BLACK;
    int64_t cnt = 0;
    ck(az_examples_count(e_, &cnt));
    std::vector<float> b((size_t)cnt * plane_), p((size_t)cnt * (A_ + 1)), v(cnt);
    if (cnt) ck(az_examples_read(e_, 0, cnt, b.data(), p.data(), v.data()));
    ck(az_examples_clear(e_));
    std::vector<Example> ex(cnt);
    for (int64_t i = 0; i < cnt; i++) {
      ex[i].Board.assign(b.begin() + i * plane_, b.begin() + (i + 1) * plane_);
      ex[i].Policy.assign(p.begin() + i * (A_ + 1), p.begin() + (i + 1) * (A_ + 1));
      ex[i].Value = v[i];
    }
    return ex;
  }
  std::vector<Example> SelfPlay() { return Play(1, true); }

  static void shuffleExamples(std::vector<Example>& ex, uint64_t seed) {  // agogo.go:251-257
    uint64_t s = seed;
    for (size_t i = 0; i < ex.size(); i++) { size_t j = splitmix64(&s) % (i + 1); std::swap(ex[i], ex[j]); }
  }

  // AZ.Learn (agogo.go:100-172)
  void Learn(int iters, int episodes, int nniters, int arenaGames) {
    const int bs = conf_.NNConf.batch_size;
    for (epoch_ = 0; epoch_ < iters; epoch_++) {
      EpochLog el{};
      setupSelfPlay(epoch_);
      std::vector<Example> ex = Play(episodes, true);
      if (conf_.MaxExamples > 0 && (int)ex.size() > conf_.MaxExamples) {
        shuffleExamples(ex, derive_seed(seed_, 1000 + 10 * epoch_));
        ex.resize(conf_.MaxExamples);
      }
      shuffleExamples(ex, derive_seed(seed_, 1001 + 10 * epoch_));  // prepareExamples (agogo.go:211-249)
      int batches = (int)ex.size() / bs, total = batches * bs;
      if (batches == 0) throw std::runtime_error("batches is nil, probably too few examples regarding the batchsize");
      std::vector<float> Xs, Pi, V;
      for (int i = 0; i < total; i++) {
        Xs.insert(Xs.end(), ex[i].Board.begin(), ex[i].Board.end());
        Pi.insert(Pi.end(), ex[i].Policy.begin(), ex[i].Policy.end());
        V.push_back(ex[i].Value);
      }
      std::vector<float> costs((size_t)batches * nniters);
      ck(az_train(e_, 1, Xs.data(), Pi.data(), V.data(), batches, nniters, 0.1f, derive_seed(seed_, 1002 + 10 * epoch_), costs.data()));
      ck(az_agent_set_inferer(e_, 1, AZ_INF_DUAL, 0));  // a.B.SwitchToInference
      ck(az_agent_reset_stats(e_, 0)); ck(az_agent_reset_stats(e_, 1));
      Play(arenaGames, false);
      ck(az_agent_stats(e_, 0, &el.a[0], &el.a[1], &el.a[2]));
      ck(az_agent_stats(e_, 1, &el.b[0], &el.b[1], &el.b[2]));
      el.promoted = el.b[0] / (el.b[0] + el.a[0]) > (float)conf_.UpdateThreshold;  // NaN (0/0) never promotes
      if (el.promoted) ck(az_net_copy(e_, 0, 1));                                  // A.NN = B.NN
      ck(az_net_init(e_, 1, derive_seed(seed_, 200 + epoch_)));                    // newB (arena.go:205-224)
      el.n_examples = (int)ex.size(); el.batches = batches;
      el.first_cost = costs.front(); el.last_cost = costs.back();
      log.push_back(el);
    }
  }

  // AZ.Save / AZ.Load (agogo.go:175-209): the flat Model()-ordered payload; the gob container is Go-side
  void Save(const std::string& filename) {
    uint64_t n = 0;
    ck(az_net_param_count(e_, nullptr, &n));
    std::vector<float> p(n);
    ck(az_net_get_params(e_, 0, p.data(), n));
    FILE* f = fopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + filename);
    fwrite(&n, 8, 1, f); fwrite(p.data(), 4, n, f); fclose(f);
  }
  void Load(const std::string& filename) {
    FILE* f = fopen(filename.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + filename);
    uint64_t n = 0, want = 0;
    ck(az_net_param_count(e_, nullptr, &want));
    if (fread(&n, 8, 1, f) != 1 || n != want) { fclose(f); throw std::runtime_error("checkpoint does not match the net"); }
    std::vector<float> p(n);
    if (fread(p.data(), 4, n, f) != n) { fclose(f); throw std::runtime_error("short checkpoint"); }
    fclose(f);
    ck(az_net_set_params(e_, 0, p.data(), n)); ck(az_net_set_params(e_, 1, p.data(), n));
    useDummy_ = false;
  }
  std::vector<float> Params(int net) {
    uint64_t n = 0;
    ck(az_net_param_count(e_, nullptr, &n));
    std::vector<float> p(n);
    ck(az_net_get_params(e_, net, p.data(), n));
    return p;
  }
  az_engine* handle() { return e_; }
  std::vector<EpochLog> log;

 private:
  void ck(int rc) { if (rc != AZ_OK) throw std::runtime_error(az_last_error(e_)); }
  Config conf_;
  uint64_t seed_;
  az_engine* e_ = nullptr;
  int cells_ = 0, A_ = 0, plane_ = 0, epoch_ = 0;
  int32_t a_player_ = AZ_NONE, b_player_ = AZ_NONE;
  bool useDummy_ = true;
};

}  // namespace agogo
