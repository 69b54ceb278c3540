// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// The reference seeds every RNG from time.Now() (mcts/tree.go:84, arena.go:61, agogo.go:252,
// dualnet/meta.go:58, game/wq/zobrist.go:32), so none of its random draws are reproducible.
// Oracle and engine both replace them with this injected, fully specified generator
// (splitmix64).  The specification — not the code — is shared with agogo_b200/csrc/rng.cuh.
#pragma once
#include <cmath>
#include <cstdint>

namespace oracle {

inline uint64_t splitmix64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { return splitmix64(&s); }
  // stand-in for rand.Intn(n): next() % n (documented; modulo bias is part of the spec)
  int intn(int n) { return (int)(next() % (uint64_t)n); }
  // uniform in [0,1): 24 high bits -> float, exact
  float uniform() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
  // standard normal via Box-Muller on two uniforms in (0,1]; computed in double, rounded to float
  float normal() {
    double u1 = ((double)(next() >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    double u2 = ((double)(next() >> 11)) * (1.0 / 9007199254740992.0);
    return (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2));
  }
};

// stream derivation: independent generator for (seed, stream index)
inline uint64_t derive_seed(uint64_t seed, uint64_t stream) {
  uint64_t s = seed ^ (0xD1B54A32D192ED03ull * (stream + 1));
  return splitmix64(&s);
}

}  // namespace oracle
