// agogo_b200 — game rules as warp-cooperative __device__ functions (K3 in SURVEY.md §2).
// Re-expression of game/mnk/mnk.go, game/c4/c4.go + game.go and game/wq/wq.go + game.go for one
// warp per game with the board (1 byte per point) in shared memory.  Not a translation: Go's
// per-move flood fills (wq.go:237-290) become one connected-component labelling + liberty count
// per position, after which every point's legality is an O(1) lookup.  Behaviour, quirks
// included, is pinned bit-exactly against the CPU oracle and the reference's golden boards.
//
// Conventions: every function is called by all 32 lanes with warp-uniform arguments unless it
// says "per lane"; `lane` = threadIdx.x & 31; shared arrays are private to the warp.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// mnk (game/mnk/mnk.go)
// isWinner, mnk.go:221-290, verbatim behaviour: the row counter is never reset, the diagonal
// walks have no column-wrap guard.  Run by one lane (boards are tiny).
__device__ inline bool mnk_is_winner_seq(const GameP& P, const uint8_t* b, int colour) {
  const int m = P.m, n = P.n, k = P.k;
  for (int i = 0; i < m; i++) {
    int rowCount = 0;
    for (int j = 0; j < n; j++) rowCount += (b[i * n + j] == colour) ? 1 : -1;
    if (rowCount >= k) return true;
  }
  for (int j = 0; j < n; j++) {
    int count = 0;
    for (int i = 0; i * n + j < P.cells; i++) count = (b[i * n + j] == colour) ? count + 1 : 0;
    if (count >= k) return true;
  }
  for (int i = 0; i < m; i++)
    for (int j = 0; n - j > n - k && j < n; j++) {
      int idx = i * n + j, diag = 0;
      while (b[idx] == colour) {
        if (++diag >= k) return true;
        idx += n + 1;
        if (idx >= m * n) break;
      }
    }
  for (int i = 0; i < m; i++)
    for (int j = n - 1; j >= k - 1; j--) {
      int idx = i * n + j, diag = 0;
      while (b[idx] == colour) {
        if (++diag >= k) return true;
        idx += n - 1;
        if (idx >= m * n) break;
      }
    }
  return false;
}

// ------------------------------------------------------------------------------------------------
// c4 (game/c4/c4.go:72-192): scan order vertical, horizontal, TLBR (x-i,y+i), TRBL (x+i,y+i);
// x-major then y, first hit wins.  One lane.
__device__ inline int c4_check_dir_seq(const GameP& P, const uint8_t* b, int dx, int dy) {
  const int rows = P.m, cols = P.n, nw = P.k;
  for (int x = 0; x < cols; x++)
    for (int y = 0; y < rows; y++) {
      int c = b[y * cols + x];
      if (c == C_NONE) continue;
      bool winning = true;
      for (int i = 0; i < nw; i++) {
        int xx = x + dx * i, yy = y + dy * i;
        if (xx >= 0 && xx < cols && yy < rows) {
          if (b[yy * cols + xx] != c) winning = false;
        } else winning = false;
      }
      if (winning) return c;
    }
  return C_NONE;
}
__device__ inline int c4_check_win_seq(const GameP& P, const uint8_t* b) {
  int w;
  if ((w = c4_check_dir_seq(P, b, 0, 1))) return w;
  if ((w = c4_check_dir_seq(P, b, 1, 0))) return w;
  if ((w = c4_check_dir_seq(P, b, -1, 1))) return w;
  return c4_check_dir_seq(P, b, 1, 1);
}
// lowest empty row of a column, -1 if full (c4.go:59-70)
__device__ inline int c4_drop_row(const GameP& P, const uint8_t* b, int col) {
  for (int row = P.m - 1; row >= 0; row--)
    if (b[row * P.n + col] == C_NONE) return row;
  return -1;
}

// ------------------------------------------------------------------------------------------------
// wq (game/wq/wq.go)
struct WqScratch {
  int* label;   // [cells] group id = smallest point index of the group, -1 for empty points
  int* libcnt;  // [cells] indexed by group id: number of DISTINCT empty neighbours of the group
  int* gsize;   // [cells] indexed by group id: stones in the group
  // positional superko (AZ_FLAG_WQ_COMPLETE): when both are set, wq_analyze also leaves in ghash[group id] the XOR of the
  // group's stones' 64-bit position keys, so that a candidate move's resulting position hash costs O(1) per lane
  unsigned long long* ghash = nullptr;      // [cells]
  const unsigned long long* zt64 = nullptr;  // [cells][2] position keys (black, white)
};

// neighbour order of wq.go:317-322: {0,+1},{+1,0},{0,-1},{-1,0} in (row, col)
__device__ inline int wq_nbr(int size, int p, int dir) {
  int r = p / size, c = p - r * size;
  switch (dir) {
    case 0: return c + 1 < size ? p + 1 : -1;
    case 1: return r + 1 < size ? p + size : -1;
    case 2: return c > 0 ? p - 1 : -1;
    default: return r > 0 ? p - size : -1;
  }
}

// Connected components (min-label propagation with pointer jumping) + liberty counts.
__device__ inline void wq_analyze(const GameP& P, const uint8_t* b, WqScratch s, int lane) {
  const int cells = P.cells, size = P.m;
  for (int i = lane; i < cells; i += 32) {
    s.label[i] = b[i] ? i : -1;
    s.libcnt[i] = 0;
    s.gsize[i] = 0;
    if (s.ghash) s.ghash[i] = 0ull;
  }
  __syncwarp();
  bool changed;
  do {
    changed = false;
    for (int i = lane; i < cells; i += 32) {
      int c = b[i];
      if (!c) continue;
      int l = s.label[i];
#pragma unroll
      for (int d = 0; d < 4; d++) {
        int j = wq_nbr(size, i, d);
        if (j >= 0 && b[j] == c) l = min(l, ((volatile int*)s.label)[j]);
      }
      l = min(l, ((volatile int*)s.label)[l]);
      if (l < s.label[i]) { ((volatile int*)s.label)[i] = l; changed = true; }
    }
    changed = __any_sync(FULL, changed);
    __syncwarp();
  } while (changed);
  for (int i = lane; i < cells; i += 32) {
    if (b[i]) {
      atomicAdd(&s.gsize[s.label[i]], 1);
      if (s.ghash && s.zt64) atomicXor(&s.ghash[s.label[i]], s.zt64[i * 2 + (b[i] == C_BLACK ? 0 : 1)]);
    } else {
      int seen[4];
      int ns = 0;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        int j = wq_nbr(size, i, d);
        if (j < 0 || !b[j]) continue;
        int g = s.label[j];
        bool dup = false;
        for (int t = 0; t < ns; t++) dup |= (seen[t] == g);
        if (!dup) { seen[ns++] = g; atomicAdd(&s.libcnt[g], 1); }
      }
    }
  }
  __syncwarp();
}

// Board.check (wq.go:205-234) for point p and mover `player`, from the analysis of the CURRENT
// board.  Per lane (p may differ per lane).  Works for occupied points exactly like the
// reference (Game.Check does not reject them, wq/game.go:65-79).  *captures_any tells whether
// any neighbouring opponent group has no liberty other than p.
__device__ inline bool wq_check_pt(const GameP& P, const uint8_t* b, WqScratch s, int p, int player,
                                   bool* captures_any, int ko = -1) {
  const int size = P.m, o = opp(player);
  if (P.wq_complete) {
    // OUR completion (include/agogo_b200.h, AZ_FLAG_WQ_COMPLETE): occupied points, the ko point, suicide and the
    // mover's own single-point eyes are illegal
    *captures_any = false;
    if (b[p] != C_NONE || p == ko) return false;
    bool cap = false, empty_nbr = false, friend_safe = false, has_opp = false;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      int a = wq_nbr(size, p, d);
      if (a < 0) continue;
      if (b[a] == C_NONE) { empty_nbr = true; continue; }
      const int libs = s.libcnt[s.label[a]];
      if (b[a] == o) { has_opp = true; if (libs == 1) cap = true; }
      else if (libs >= 2) friend_safe = true;
    }
    *captures_any = cap;
    if (!empty_nbr && !has_opp) return false;  // own eye
    return cap || empty_nbr || friend_safe;
  }
  const int need = b[p] == C_NONE ? 1 : 0;  // p itself is a liberty of its neighbours iff it is empty
  bool cap = false, empty_nbr = false;
#pragma unroll
  for (int d = 0; d < 4; d++) {
    int a = wq_nbr(size, p, d);
    if (a < 0) continue;
    if (b[a] == o && s.libcnt[s.label[a]] == need) cap = true;
    if (b[a] == C_NONE) empty_nbr = true;
  }
  *captures_any = cap;
  if (cap) return true;
  // suicide test nolib(c, {-5,-5}) (wq.go:229-233): for an empty point the "group" is {p} and any
  // empty neighbour is a liberty; for an occupied point it is the stone's own group.
  if (b[p] == C_NONE) return empty_nbr;
  return s.libcnt[s.label[p]] > 0;
}

// Positional superko (AZ_FLAG_WQ_COMPLETE; game.go:77's TODO): 64-bit hash of the position.  Whole warp.
__device__ inline unsigned long long wq_pos_hash(const GameP& P, const uint8_t* b, const unsigned long long* __restrict__ zt64, int lane) {
  unsigned long long h = 0ull;
  for (int i = lane; i < P.cells; i += 32)
    if (b[i]) h ^= zt64[i * 2 + (b[i] == C_BLACK ? 0 : 1)];
#pragma unroll
  for (int off = 16; off; off >>= 1) h ^= __shfl_xor_sync(FULL, h, off);
  return h;
}
// Hash of the position a LEGAL move of `player` at the empty point p leads to, from the analysis of the current board
// (with group hashes) and the current position's hash: the stone, minus every neighbouring opponent group whose only
// liberty is p (each once).  Per lane.
__device__ inline unsigned long long wq_hash_after(const GameP& P, const uint8_t* b, WqScratch s, int p, int player,
                                                   unsigned long long cur) {
  const int size = P.m, o = opp(player);
  unsigned long long h = cur ^ s.zt64[p * 2 + (player == C_BLACK ? 0 : 1)];
  int seen[4];
#pragma unroll
  for (int d = 0; d < 4; d++) {
    int a = wq_nbr(size, p, d);
    seen[d] = (a >= 0 && b[a] == o && s.libcnt[s.label[a]] == 1) ? s.label[a] : -1;
#pragma unroll
    for (int q = 0; q < d; q++) if (seen[q] == seen[d]) seen[d] = -1;
    if (seen[d] >= 0) h ^= s.ghash[seen[d]];
  }
  return h;
}

// Board.Apply (wq.go:141-171) on the board in shared memory; needs wq_analyze(b) beforehand.
// Returns whether the board-level apply succeeded; *taken = len(captures) with the reference's
// duplicates (a group adjacent to p through two neighbours is listed twice); *zhash is updated
// with one XOR per listing when ztable != nullptr (zobrist.go:44-56).
__device__ inline bool wq_board_apply(const GameP& P, uint8_t* b, WqScratch s, int p, int player, int lane,
                                      int* taken, int* zhash, const int* __restrict__ ztable, int ko = -1,
                                      int* ko_out = nullptr) {
  *taken = 0;
  if (ko_out) *ko_out = -1;
  if (!(player == C_BLACK || player == C_WHITE)) return false;
  if (p >= P.cells || p < 0) return false;
  if (b[p] != C_NONE) return false;
  bool cap;
  if (!wq_check_pt(P, b, s, p, player, &cap, ko)) return false;
  const int size = P.m, o = opp(player);
  int h = 0, tk = 0;
  if (ztable) h ^= ztable[p * 2 + (player == C_BLACK ? 0 : 1)];
  int groups[4];
  bool lone = true;  // complete rules: no friendly and no empty neighbour -> after capturing one stone this is a ko
#pragma unroll
  for (int d = 0; d < 4; d++) {
    int a = wq_nbr(size, p, d);
    groups[d] = (a >= 0 && b[a] == o && s.libcnt[s.label[a]] == 1) ? s.label[a] : -1;
    if (a >= 0 && b[a] != o) lone = false;
    if (P.wq_complete)  // each captured group once (the reference lists a group per neighbour through which it touches p)
      for (int q = 0; q < d; q++) if (groups[q] == groups[d]) groups[d] = -1;
  }
  __syncwarp();
  int lh = 0;
#pragma unroll
  for (int d = 0; d < 4; d++) {
    int g = groups[d];
    if (g < 0) continue;
    tk += s.gsize[g];
    for (int i = lane; i < P.cells; i += 32)
      if (s.label[i] == g) {
        b[i] = C_NONE;
        if (ztable) lh ^= ztable[i * 2 + (o == C_BLACK ? 0 : 1)];
      }
  }
  if (lane == 0) b[p] = (uint8_t)player;
#pragma unroll
  for (int off = 16; off; off >>= 1) lh ^= __shfl_xor_sync(FULL, lh, off);
  __syncwarp();
  if (zhash) *zhash ^= h ^ lh;
  *taken = tk;
  if (P.wq_complete && ko_out && tk == 1 && lone) {
#pragma unroll
    for (int d = 0; d < 4; d++) if (groups[d] >= 0) *ko_out = groups[d];  // a one-stone group's label is its point
  }
  return true;
}

// Area score (AZ_FLAG_WQ_COMPLETE): stones of the colour + empty regions whose border touches that colour only.  Whole
// warp; clobbers the analysis scratch (label = empty-region ids, libcnt = border colour mask, gsize = region sizes).
__device__ inline void wq_area_scores(const GameP& P, const uint8_t* b, WqScratch s, int lane, float* black, float* white) {
  const int cells = P.cells, size = P.m;
  for (int i = lane; i < cells; i += 32) { s.label[i] = b[i] ? -1 : i; s.libcnt[i] = 0; s.gsize[i] = 0; }
  __syncwarp();
  bool changed;
  do {
    changed = false;
    for (int i = lane; i < cells; i += 32) {
      if (b[i]) continue;
      int l = s.label[i];
#pragma unroll
      for (int d = 0; d < 4; d++) {
        int j = wq_nbr(size, i, d);
        if (j >= 0 && !b[j]) l = min(l, ((volatile int*)s.label)[j]);
      }
      l = min(l, ((volatile int*)s.label)[l]);
      if (l < s.label[i]) { ((volatile int*)s.label)[i] = l; changed = true; }
    }
    changed = __any_sync(FULL, changed);
    __syncwarp();
  } while (changed);
  int nb = 0, nw = 0;
  for (int i = lane; i < cells; i += 32) {
    if (b[i] == C_BLACK) nb++;
    else if (b[i] == C_WHITE) nw++;
    else {
      int mask = 0;
#pragma unroll
      for (int d = 0; d < 4; d++) {
        int j = wq_nbr(size, i, d);
        if (j >= 0 && b[j]) mask |= b[j] == C_BLACK ? 1 : 2;
      }
      atomicAdd(&s.gsize[s.label[i]], 1);
      if (mask) atomicOr(&s.libcnt[s.label[i]], mask);
    }
  }
  __syncwarp();
  for (int i = lane; i < cells; i += 32)
    if (!b[i] && s.label[i] == i) {
      if (s.libcnt[i] == 1) nb += s.gsize[i];
      else if (s.libcnt[i] == 2) nw += s.gsize[i];
    }
#pragma unroll
  for (int off = 16; off; off >>= 1) { nb += __shfl_xor_sync(FULL, nb, off); nw += __shfl_xor_sync(FULL, nw, off); }
  *black = (float)nb; *white = (float)nw;
  __syncwarp();
}

// Board.Score (wq.go:173-202): stones of the colour plus the empties its buggy flood fill reaches
// — only row 0 is ever expanded (`a >= b.size` rejects the rest; adjacents are {-size,1,size,1}).
// One lane.
__device__ inline float wq_score_seq(const GameP& P, const uint8_t* b, int colour) {
  const int size = P.m;
  int cnt = 0;
  for (int i = 0; i < P.cells; i++) cnt += (b[i] == colour);
  bool reach_next = false;
  for (int x = 0; x < size; x++) {
    if (b[x] == C_NONE) {
      bool reached = reach_next || (x + size < P.cells && b[x + size] == colour);
      cnt += reached;
      reach_next = reached;
    } else {
      reach_next = (b[x] == colour);
    }
  }
  return (float)cnt;
}

// ------------------------------------------------------------------------------------------------
// State.Ended (mnk.go:156-169, c4/game.go:161-179, wq/game.go:94-115).  `passes` is c4's
// passCount or wq's passes.  Result broadcast to the warp.
__device__ inline bool game_ended(const GameP& P, const uint8_t* b, int passes, int lane, int* winner,
                                  const WqScratch* sc = nullptr) {
  int e = 0, w = C_NONE;
  if (P.kind == KIND_WQ && P.wq_complete) {  // area scores, komi to White (warp-wide; needs the scratch)
    if (passes < 2) { *winner = C_NONE; return false; }
    float bs, ws;
    wq_area_scores(P, b, *sc, lane, &bs, &ws);
    ws = __fadd_rn(ws, P.komi);
    *winner = (ws == bs) ? C_NONE : (ws > bs ? C_WHITE : C_BLACK);
    return true;
  }
  if (lane == 0) {
    if (P.kind == KIND_MNK) {
      if (mnk_is_winner_seq(P, b, C_BLACK)) { e = 1; w = C_BLACK; }
      else if (mnk_is_winner_seq(P, b, C_WHITE)) { e = 1; w = C_WHITE; }
      else {
        e = 1;
        for (int i = 0; i < P.cells; i++) if (b[i] == C_NONE) { e = 0; break; }
      }
    } else if (P.kind == KIND_C4) {
      int cw = c4_check_win_seq(P, b);
      if (cw != C_NONE) { e = 1; w = cw; }
      else if (passes > 2) { e = 1; }
      else {
        e = 1;
        for (int i = 0; i < P.cells; i++) if (b[i] == C_NONE) { e = 0; break; }
      }
    } else {
      if (passes >= 2) {
        e = 1;
        float ws = wq_score_seq(P, b, C_WHITE), bs = wq_score_seq(P, b, C_BLACK);
        w = (ws == bs) ? C_NONE : (ws > bs ? C_WHITE : C_BLACK);
      }
    }
  }
  e = __shfl_sync(FULL, e, 0);
  *winner = __shfl_sync(FULL, w, 0);
  return e != 0;
}

// State.Score (mnk.go:142-150, c4/game.go:74-83, wq Board.Score); one lane's result broadcast.
__device__ inline float game_score(const GameP& P, const uint8_t* b, int player, int lane, const WqScratch* scr = nullptr) {
  float sc = 0;
  if (P.kind == KIND_WQ && P.wq_complete) {
    float bs, ws;
    wq_area_scores(P, b, *scr, lane, &bs, &ws);
    return player == C_BLACK ? bs : ws;
  }
  if (lane == 0) {
    if (P.kind == KIND_MNK) {
      if (mnk_is_winner_seq(P, b, player)) sc = 1;
      else if (mnk_is_winner_seq(P, b, opp(player))) sc = -2;
    } else if (P.kind == KIND_C4) {
      int w = c4_check_win_seq(P, b);
      sc = (w == player) ? 1.f : (w == C_NONE ? 0.f : -1.f);
    } else {
      sc = wq_score_seq(P, b, player);
    }
  }
  return __shfl_sync(FULL, sc, 0);
}

// FNV-1a 32 of fmt "%v" of every colour (mnk.go:70-76, c4/game.go:203-210); one lane, broadcast.
__device__ inline uint32_t fnv_board_hash(const GameP& P, const uint8_t* b, int lane) {
  uint32_t h = 2166136261u;
  if (lane == 0) {
    for (int i = 0; i < P.cells; i++) {
      const char* s = b[i] == C_BLACK ? "Black" : (b[i] == C_WHITE ? "White" : "None");
      for (; *s; ++s) { h ^= (uint8_t)*s; h *= 16777619u; }
    }
  }
  return __shfl_sync(FULL, h, 0);
}

// State.Check for a non-wq game, per lane (mnk.go:96-115 ; c4.go:59-70 via c4/game.go:53).
__device__ inline bool simple_check(const GameP& P, const uint8_t* b, int move) {
  if (move == MV_RESIGN) return P.kind == KIND_MNK;  // mnk accepts resign; c4 would index out of range
  if (P.kind == KIND_MNK) {
    if (move == MV_PASS) return false;
    if (move >= P.cells || move < 0) return false;
    return b[move] == C_NONE;
  }
  // c4
  if (move == MV_PASS) return true;
  if (move < 0 || move >= P.n) return false;
  return c4_drop_row(P, b, move) >= 0;
}
