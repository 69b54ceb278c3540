// host/tictactoe.cpp — the reference's one runnable program (cmd/tictactoe/main.go) on the engine:
// same Config (DefaultConf(3,3,10), BatchSize 100, Features 2, K 3, SharedLayers 3, PUCT 1,
// DontPreferPass, DumbPass), Learn(iters, episodes, nniters, arenaGames) then Save.
//   usage: tictactoe [iters episodes nniters arenaGames [sims batch seed]]   (default: 5 30 200 30)
// Prints one line per epoch: stats of A and B, examples, batches, promotion, first/last cost;
// then an FNV-1a hash of A's final weights (used by the tests to compare host layers).
#include <cstdlib>
#include <cstring>

#include "agogo.hpp"

int main(int argc, char** argv) {
  int iters = 5, episodes = 30, nniters = 200, arenaGames = 30, sims = 100, batch = 100;
  uint64_t seed = 1;
  if (argc >= 5) { iters = atoi(argv[1]); episodes = atoi(argv[2]); nniters = atoi(argv[3]); arenaGames = atoi(argv[4]); }
  if (argc >= 8) { sims = atoi(argv[5]); batch = atoi(argv[6]); seed = strtoull(argv[7], nullptr, 10); }
  agogo::Config conf;
  conf.Name = "Tic Tac Toe";
  conf.NNConf = agogo::DefaultConf(3, 3, 10);
  conf.NNConf.batch_size = batch; conf.NNConf.features = 2; conf.NNConf.k = 3; conf.NNConf.shared_layers = 3;
  conf.MCTSConf = agogo::DefaultConfig(3);
  conf.MCTSConf.budget = 1000; conf.MCTSConf.sims = sims;
  conf.UpdateThreshold = 0.52;
  az_game_desc g{};
  g.kind = AZ_GAME_MNK; g.m = 3; g.n = 3; g.k = 3; g.zobrist_seed = 12345;
  try {
    agogo::AZ az(g, conf, 64, seed);
    az.Learn(iters, episodes, nniters, arenaGames);
    for (size_t i = 0; i < az.log.size(); i++) {
      const agogo::EpochLog& l = az.log[i];
      printf("epoch %zu A %g %g %g B %g %g %g examples %d batches %d promoted %d cost %.9g %.9g\n", i, l.a[0], l.a[1], l.a[2],
             l.b[0], l.b[1], l.b[2], l.n_examples, l.batches, (int)l.promoted, l.first_cost, l.last_cost);
    }
    std::vector<float> p = az.Params(0);
    uint32_t h = 2166136261u;
    for (float v : p) { uint32_t b; memcpy(&b, &v, 4); for (int k = 0; k < 4; k++) { h ^= (b >> (8 * k)) & 0xff; h *= 16777619u; } }
    printf("weights_fnv %08x n %zu\n", h, p.size());
    az.Save("tictactoe.model");
  } catch (const std::exception& ex) {
    fprintf(stderr, "panic: %s\n", ex.what());
    return 2;
  }
  return 0;
}
