// agogo_b200 — MCTS kernels: one warp per game, trees as flat SoA pools in HBM.
//   k_search_begin   updateRoot + prepareRoot            (mcts/search.go:473-500, 392-408)
//   k_select         pipeline descent: PUCT select, Check/Apply, leaf emit   (search.go:209-248, node.go:170-237)
//   k_expand_backup  expandAndSimulate + Update along the path               (search.go:259-339, node.go:70-76)
//   k_search_end     bestMove, cachedPolicies, Example, Apply, Ended         (search.go:341-390,152-161; arena.go:98-137)
// Compiled with -fmad=false: the PUCT arithmetic must match Go/amd64 bit for bit (no FMA), and
// every rounding-sensitive operation is additionally spelled with an explicit _rn intrinsic.
#include <stdexcept>

#include "mcts_dev.cuh"
#include "rules.cuh"

// ---------------------------------------------------------------------------------------------
// per-warp shared-memory workspace
struct WS {
  uint8_t* board;
  uint8_t* hist;
  WqScratch wq;
  float *fa, *fb;
  int *ia, *ib, *ic;
  uint32_t* st[5];
};
__host__ __device__ inline size_t ws_bytes(const GameP& P, int cellsP) {
  size_t n = cellsP + (size_t)(P.hist_len ? 8 * cellsP : 0);
  n += 3 * (size_t)cellsP * 4;                 // label, libcnt, gsize
  if (P.wq_complete) n += (size_t)cellsP * 8;  // ghash (positional superko)
  n += 10 * (size_t)((P.A + 2 + 3) & ~3) * 4;  // fa fb ia ib ic st[5]
  return (n + 15) & ~(size_t)15;
}
__device__ inline WS make_ws(const GameP& P, int cellsP, uint8_t* base) {
  WS w;
  w.board = base; base += cellsP;
  w.hist = base; if (P.hist_len) base += 8 * cellsP;
  w.wq.label = (int*)base; base += cellsP * 4;
  w.wq.libcnt = (int*)base; base += cellsP * 4;
  w.wq.gsize = (int*)base; base += cellsP * 4;
  if (P.wq_complete) { w.wq.ghash = (unsigned long long*)base; base += cellsP * 8; }  // cellsP % 16 == 0: 8-byte aligned
  int AP = (P.A + 2 + 3) & ~3;
  w.fa = (float*)base; base += AP * 4;
  w.fb = (float*)base; base += AP * 4;
  w.ia = (int*)base; base += AP * 4;
  w.ib = (int*)base; base += AP * 4;
  w.ic = (int*)base; base += AP * 4;
  for (int i = 0; i < 5; i++) { w.st[i] = (uint32_t*)base; base += AP * 4; }
  return w;
}
size_t mcts_ws_bytes(const GameP& P, int cellsP) { return ws_bytes(P, cellsP); }

__device__ inline void count(const EngineDev& E, int which, unsigned long long v, int lane) {
  if (lane == 0 && v) atomicAdd(&E.counters[which], v);
}
__device__ inline void raise(const EngineDev& E, int bit, int lane) {
  if (lane == 0) atomicOr(E.err, bit);
}

// ---------------------------------------------------------------------------------------------
// encoders (K4).  two-plane: cmd/tictactoe/main.go:26-47.  wq18: encoding_helper.go:29-68 —
// planes i=1..7 of each colour group hold Historical((MoveNumber-1)-i) when that index is > 0,
// both colours in every plane (+1 black / -1 white; the "white" group is the negation, so empty
// points become -0.0 there), plane 7 of each group and the current board are never written.
__device__ inline void encode_planes(const GameP& P, const uint8_t* board, const uint8_t* ring, int cellsP,
                                     int to_move, int move_number, float* __restrict__ out, int lane) {
  const int cells = P.cells;
  if (P.encoder == 0) {
    float pv = to_move == C_BLACK ? 1.0f : (to_move == C_WHITE ? -1.0f : 0.0f);
    for (int i = lane; i < cells; i += 32) {
      int c = board[i];
      out[i] = c == C_BLACK ? 1.0f : (c == C_WHITE ? -1.0f : 0.001f);
      out[cells + i] = pv;
    }
    return;
  }
  const bool blk = to_move == C_BLACK;
  const int bs = blk ? 0 : 8, ns = blk ? 16 : 17;
  const float ep = blk ? 1.0f : -1.0f;
  const int current = move_number - 1;
  for (int q = 0; q < 18; q++) {
    float* o = out + (size_t)q * cells;
    if (q >= 16) {
      float v = q == ns ? ep : 0.0f;
      for (int i = lane; i < cells; i += 32) o[i] = v;
      continue;
    }
    int idx = q & 7;
    bool black_group = (q & 8) == bs;
    int h = current - (idx + 1);
    if (idx < 7 && h > 0 && h < current) {
      const uint8_t* past = ring + (size_t)(h & 7) * cellsP;
      for (int i = lane; i < cells; i += 32) {
        int c = past[i];
        float v = c == C_BLACK ? 1.0f : (c == C_WHITE ? -1.0f : 0.0f);
        o[i] = black_group ? v : __fmul_rn(v, -1.0f);
      }
    } else {
      for (int i = lane; i < cells; i += 32) o[i] = 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// State.Check / State.Apply on the warp's working state (board + ring in shared memory).
struct St { int to_move, move_number, passes; int ko = -1; };  // ko: AZ_FLAG_WQ_COMPLETE only (-1 = none)

// warp-uniform move.  For wq the board analysis must be current when `analyzed` is true.
__device__ inline bool state_check(const GameP& P, WS& w, int player, int move, int lane, bool* analyzed, int ko = -1) {
  if (P.kind != KIND_WQ) return simple_check(P, w.board, move);
  if (move == MV_RESIGN || move == MV_PASS) return true;  // wq/game.go:66-71
  if (move >= P.cells || move < 0) return false;
  if (!*analyzed) { wq_analyze(P, w.board, w.wq, lane); *analyzed = true; }
  bool cap;
  return wq_check_pt(P, w.board, w.wq, move, player, &cap, ko);
}

// In-tree / root Apply of a move already known to pass Check (mnk.go:117-137, c4/game.go:55-72,
// wq/game.go:81-92 + COMPLETION for pass/passes/historical).  zhash may be null (in-tree).
__device__ inline void state_apply(const GameP& P, WS& w, int cellsP, St& s, int player, int move, int lane,
                                   bool* analyzed, int* zhash, const int* ztable) {
  __syncwarp();  // every lane's State.Check read of the board is done before one lane writes the move (racecheck: WAR)
  if (P.kind == KIND_MNK) {
    if (lane == 0) w.board[move] = (uint8_t)player;
    s.to_move = opp(player);
    s.move_number++;
    __syncwarp();
  } else if (P.kind == KIND_C4) {
    if (move != MV_PASS) {
      if (lane == 0) {
        int row = c4_drop_row(P, w.board, move);
        if (row >= 0) w.board[row * P.n + move] = (uint8_t)player;
      }
      __syncwarp();
    }
  } else {
    if (P.hist_len) {  // COMPLETION: historical = board before the move
      uint8_t* dst = w.hist + (size_t)(s.move_number & 7) * cellsP;
      for (int i = lane; i < P.cells; i += 32) dst[i] = w.board[i];
    }
    if (move == MV_PASS) {
      s.passes++;
      s.ko = -1;
    } else {
      if (!*analyzed) { wq_analyze(P, w.board, w.wq, lane); *analyzed = true; }
      int taken, nko = -1;
      wq_board_apply(P, w.board, w.wq, move, player, lane, &taken, zhash, ztable, s.ko, &nko);  // error ignored (game.go:84)
      s.ko = nko;
      s.passes = 0;
    }
    s.to_move = opp(player);
    s.move_number++;
    *analyzed = false;
    __syncwarp();
  }
}

// Node.Evaluate (node.go:147-159); virtualLoss is 0 for every candidate under 1-worker semantics.
__device__ inline float evaluate(float W, uint32_t N, int player) {
  float score = __fdiv_rn(W, __uint2float_rn(N));
  if (player == C_WHITE) score = __fsub_rn(1.0f, score);
  return score;
}

// Update (node.go:70-76, 263-270) for every node on the path
__device__ inline void backup(const EngineDev& E, size_t tb, const int* path, int path_len, float v, int lane) {
  for (int i = lane; i < path_len; i += 32) {
    size_t idx = tb + path[i];
    E.N[idx] = E.N[idx] + 1u;
    E.W[idx] = __fadd_rn(E.W[idx], v);
  }
  count(E, CNT_BACKUP, path_len, lane);
}

__device__ inline int alloc_nodes(const EngineDev& E, const GameP& P, int* ti, int n, int lane) {
  int a = (ti[TI_ALLOC] + 3) & ~3;  // 4-node (16 B per array) alignment of every child block
  if (a + n + 4 > P.max_nodes) { raise(E, ERR_POOL_EXHAUSTED, lane); return -1; }  // +4: vector loads may touch the pad
  __syncwarp();
  if (lane == 0) ti[TI_ALLOC] = a + n;
  __syncwarp();
  return a;
}

// ---------------------------------------------------------------------------------------------
__device__ inline unsigned long long dev_splitmix64(unsigned long long* s) {
  unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ inline unsigned long long dev_derive_seed(unsigned long long seed, unsigned long long stream) {
  unsigned long long s = seed ^ (0xD1B54A32D192ED03ull * (stream + 1));
  return dev_splitmix64(&s);
}
// game_base: games this engine started before this call — tree t of the c-th game draws from stream 2c + t of the
// engine's tree seed (the reference seeds every MCTS.rand from the clock, tree.go:84: all trees differ)
__global__ void k_arena_begin(GameP P, EngineDev E, int n_games, const int* __restrict__ coins, unsigned long long game_base) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= E.G) return;
  int* gi = E.gi + (size_t)g * GI_COUNT;
  uint8_t* b = E.board + (size_t)g * E.cellsP;
  for (int i = lane; i < E.cellsP; i += 32) b[i] = 0;
  if (P.hist_len) for (int i = lane; i < 8 * E.cellsP; i += 32) E.hist[(size_t)g * 8 * E.cellsP + i] = 0;
  for (int t = 0; t < E.T; t++) {
    int* ti = E.ti + ((size_t)g * E.T + t) * TI_COUNT;
    const unsigned long long ts = dev_derive_seed(E.tree_seed, 2ull * (game_base + (unsigned long long)g) + (unsigned long long)t);
    if (lane < TI_COUNT)
      ti[lane] = lane == TI_ROOT ? -1 : (lane == TI_RNG_LO ? (int)(unsigned)(ts & 0xffffffffu) : (lane == TI_RNG_HI ? (int)(unsigned)(ts >> 32) : 0));
  }
  for (int i = lane; i < E.V * WV_COUNT; i += 32) WV_OF(E, g, 0)[i] = 0;
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < GI_COUNT; i++) gi[i] = 0;
    if (g < n_games) {
      // arena.go:81-89: coin 0 -> A is Black and moves first, else B is Black and moves first
      int coin = coins[g];
      gi[GI_A_PLAYER] = coin == 0 ? C_BLACK : C_WHITE;
      gi[GI_CUR_AGENT] = coin == 0 ? 0 : 1;
      gi[GI_TO_MOVE] = C_BLACK;  // SetToMove(currentPlayer.Player), arena.go:91
      gi[GI_PASSES] = P.kind == KIND_MNK ? -1 : 0;
      gi[GI_LAST_MOVE] = MV_PASS;  // State.LastMove() of an empty history (mnk.go:84-89)
      gi[GI_KO] = -1;
      if (E.poshash) E.poshash[(size_t)g * (P.max_plies + 2)] = 0ull;  // the empty board; GI_N_POS = 0 positions before it
    }
  }
  __syncwarp();
  if (g < n_games) {
    int winner;
    bool ended = game_ended(P, b, 0, lane, &winner);
    if (lane == 0) { gi[GI_ACTIVE] = ended ? 0 : 1; gi[GI_WINNER] = winner; }
  }
}

// slot of every active game inside its agent's evaluation batch (one warp does the scan)
__global__ void k_assign_slots(EngineDev E, int n_games, int shared_tree) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  int cnt[2] = {0, 0};
  for (int base = 0; base < n_games; base += 32) {
    int g = base + lane;
    int active = 0, agent = 0;
    if (g < n_games) {
      active = E.gi[(size_t)g * GI_COUNT + GI_ACTIVE];
      agent = shared_tree ? 0 : E.gi[(size_t)g * GI_COUNT + GI_CUR_AGENT];
    }
    for (int a = 0; a < 2; a++) {
      unsigned m = __ballot_sync(FULL, active && agent == a);
      if (active && agent == a) {
        WV_OF(E, g, 0)[WV_SLOT] = cnt[a] + __popc(m & ((1u << lane) - 1));
        WV_OF(E, g, 0)[WV_AGENT] = a;
      }
      cnt[a] += __popc(m);
    }
  }
  if (lane == 0) {
    E.batch_count[0] = cnt[0]; E.batch_count[1] = cnt[1]; E.batch_count[2] = cnt[0]; E.batch_count[3] = cnt[1];
    *E.n_active = cnt[0] + cnt[1];
  }
}

// findChild (node.go:288-298): first child with the wanted move, or -1
__device__ inline int find_child(const EngineDev& E, size_t tb, int node, int move, int lane) {
  uint32_t meta = E.meta[tb + node];
  int nc = META_NCHILD(meta), first = E.first[tb + node];
  for (int base = 0; base < nc; base += 32) {
    int j = base + lane;
    bool hit = j < nc && META_MOVE(E.meta[tb + first + j]) == move;
    unsigned m = __ballot_sync(FULL, hit);
    if (m) return first + base + __ffs(m) - 1;
  }
  return -1;
}

// updateRoot + prepareRoot.  Leaves WV_STATUS = ST_LEAF (root evaluation wanted, path = [root])
// or ST_DONE.
__global__ void k_search_begin(GameP P, EngineDev E, int n_games) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  int* gi = E.gi + (size_t)g * GI_COUNT;
  int* wv = WV_OF(E, g, 0);
  for (int l = 1 + lane; l < E.V; l += 32) WV_OF(E, g, l)[WV_STATUS] = ST_IDLE;  // the root preparation is worker 0's
  if (!gi[GI_ACTIVE]) { if (lane == 0) wv[WV_STATUS] = ST_IDLE; return; }
  WS w = make_ws(P, E.cellsP, smem + (size_t)wib * ws_bytes(P, E.cellsP));
  const uint8_t* gb = E.board + (size_t)g * E.cellsP;
  for (int i = lane; i < P.cells; i += 32) w.board[i] = gb[i];
  __syncwarp();
  const int agent = gi[GI_CUR_AGENT];
  const int t = P.shared_tree ? 0 : agent;
  const int player = agent == 0 ? gi[GI_A_PLAYER] : opp(gi[GI_A_PLAYER]);
  int* ti = E.ti + ((size_t)g * E.T + t) * TI_COUNT;
  const size_t tb = ((size_t)g * E.T + t) * (size_t)P.max_nodes;
  const int mn = gi[GI_MOVE_NUMBER];
  count(E, CNT_SEARCHES, 1, lane);

  // ---- updateRoot (search.go:473-500) / newRootState (424-469)
  int root = ti[TI_ROOT];
  bool reused = false;
  if (root >= 0 && ti[TI_PREV_VALID] && P.kind == KIND_MNK) {
    int d = mn - ti[TI_PREV_MN];
    if (d >= 0) {
      // tmp.UndoLastMove x d == prev holds by construction: both are this game's own history.
      reused = true;
      const int16_t* hm = E.hmoves + (size_t)g * P.max_plies;
      for (int i = 0; i < d; i++) {
        int nr = find_child(E, tb, root, hm[ti[TI_PREV_MN] + i], lane);
        if (nr < 0) { reused = false; break; }
        root = nr;  // cleanup(): siblings become unreachable
      }
    }
  }
  if (!reused) {
    if (P.kind != KIND_MNK) { if (lane == 0) ti[TI_ALLOC] = 0; __syncwarp(); }  // no reuse possible: recycle the pool
    // New(Pass,0,0) if Check(Pass) else first legal move (search.go:478-487)
    int rmove = MV_PASS;
    bool have = true;
    if (P.kind == KIND_MNK) {
      have = false;
      for (int base = 0; base < P.A && !have; base += 32) {
        int i = base + lane;
        unsigned m = __ballot_sync(FULL, i < P.A && w.board[i] == C_NONE);
        if (m) { rmove = base + __ffs(m) - 1; have = true; }
      }
    }
    if (have) {
      int a = alloc_nodes(E, P, ti, 1, lane);
      if (a < 0) { if (lane == 0) wv[WV_STATUS] = ST_IDLE; return; }
      if (lane == 0) {
        E.N[tb + a] = 1; E.W[tb + a] = 0.0f; E.Pr[tb + a] = 0.0f;
        E.meta[tb + a] = META_MAKE(rmove, 0, 0); E.first[tb + a] = -1;
      }
      root = a;
    }
    // (no legal move and no pass: the reference keeps the stale root; cannot happen for an un-ended game)
  }
  __syncwarp();
  if (root < 0) { raise(E, ERR_ROOT_NO_CHILDREN, lane); if (lane == 0) wv[WV_STATUS] = ST_IDLE; return; }
  uint32_t rmeta = E.meta[tb + root];
  if (META_NCHILD(rmeta) == 0 && META_EXPANDED(rmeta)) {  // search.go:496-499
    rmeta = META_MAKE(META_MOVE(rmeta), 0, 0);
    if (lane == 0) E.meta[tb + root] = rmeta;
  }
  // ---- Search prologue (search.go:95-96): SetToMove, board hash
  uint32_t hash = P.kind == KIND_WQ ? (uint32_t)gi[GI_ZHASH] : fnv_board_hash(P, w.board, lane);
  if (lane == 0) {
    ti[TI_ROOT] = root;
    ti[TI_PREV_VALID] = 0;  // t.prev = nil (search.go:490)
    gi[GI_TO_MOVE] = player;
    wv[WV_TREE] = t; wv[WV_PLAYER] = player; wv[WV_HASH] = (int)hash;
    wv[WV_FLAGS] = 1;  // root preparation wave
    wv[WV_PATHLEN] = 1;
    E.path[(size_t)g * E.V * (P.maxDepth + 1)] = root;
  }
  // ---- prepareRoot (search.go:392-408)
  const bool hadChildren = META_NCHILD(rmeta) > 0;
  const bool expandable = !META_EXPANDED(rmeta);
  const int passes = P.kind == KIND_WQ ? gi[GI_PASSES] : (P.kind == KIND_MNK ? -1 : 0);
  if (expandable && passes < 2) {
    // leaf request on the root state
    uint8_t* lb = E.leaf_board + (size_t)g * E.V * E.cellsP;
    for (int i = lane; i < P.cells; i += 32) lb[i] = w.board[i];
    if (lane == 0) {
      wv[WV_STATUS] = ST_LEAF; wv[WV_TO_MOVE] = player; wv[WV_MOVE_NUMBER] = mn; wv[WV_PASSES] = passes;
      wv[WV_KO] = gi[GI_KO];
      wv[WV_NPATH] = 0;
      if (E.pathhash) E.pathhash[(size_t)g * E.V * (P.maxDepth + 2)] = E.poshash[(size_t)g * (P.max_plies + 2) + gi[GI_N_POS]];
    }
    // planes are written after slots are known (k_encode_roots)
  } else {
    if (!hadChildren && lane == 0) {  // root.Update(0)
      E.N[tb + root] += 1u;
      E.W[tb + root] = __fadd_rn(E.W[tb + root], 0.0f);
    }
    if (!hadChildren) count(E, CNT_BACKUP, 1, lane);
    if (lane == 0) wv[WV_STATUS] = ST_DONE;
  }
}

// planes of the root states that asked for an evaluation (needs slots)
__global__ void k_encode_roots(GameP P, EngineDev E, int n_games) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  const int* wv = WV_OF(E, g, 0);
  if (wv[WV_STATUS] != ST_LEAF) return;
  const int agent = wv[WV_AGENT];
  if (E.inf[agent].kind != INF_DUAL) return;
  float* out = E.nn_in + ((size_t)agent * E.GS + wv[WV_SLOT]) * P.plane;
  encode_planes(P, E.board + (size_t)g * E.cellsP, E.hist + (size_t)g * 8 * E.cellsP, E.cellsP, wv[WV_TO_MOVE],
                wv[WV_MOVE_NUMBER], out, lane);
}

// ---------------------------------------------------------------------------------------------
// K1: the descent part of pipeline() (search.go:209-248), one warp per game.  With mcts.Config workers = V > 1 the
// warp starts V pipeline calls one after the other (the fixed interleaving of the reference's concurrent
// searchStates, include/agogo_b200.h): each sets the virtual-loss flag on its path (search.go:222, node.go:248-253)
// and stops at the leaf it wants evaluated; null results and two-pass terminals complete at once and clear their flags.
__global__ void k_select(GameP P, EngineDev E, int n_games) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  const int workers = E.V > 1 ? min(*E.round_workers, E.V) : 1;
  if (E.V > 1 && blockIdx.x == 0 && threadIdx.x == 0) {
    E.batch_count[0] = E.batch_count[2] * workers;
    E.batch_count[1] = E.batch_count[3] * workers;
  }
  if (g >= n_games) return;
  int* gi = E.gi + (size_t)g * GI_COUNT;
  int* wv0 = WV_OF(E, g, 0);
  if (!gi[GI_ACTIVE]) { for (int l = lane; l < E.V; l += 32) WV_OF(E, g, l)[WV_STATUS] = ST_IDLE; return; }
  for (int l = workers + lane; l < E.V; l += 32) WV_OF(E, g, l)[WV_STATUS] = ST_IDLE;
  WS w = make_ws(P, E.cellsP, smem + (size_t)wib * ws_bytes(P, E.cellsP));
  const uint8_t* gb = E.board + (size_t)g * E.cellsP;
  const int t = wv0[WV_TREE];
  const int agent = wv0[WV_AGENT];
  const int slot0 = wv0[WV_SLOT], slot_stride = E.batch_count[2 + agent];
  const size_t tb = ((size_t)g * E.T + t) * (size_t)P.max_nodes;
  const int* ti = E.ti + ((size_t)g * E.T + t) * TI_COUNT;
  uint8_t* vl = E.vl ? E.vl + tb : nullptr;
  if (lane == 0) wv0[WV_FLAGS] = 0;

  for (int wk = 0; wk < workers; wk++) {
    int* wv = WV_OF(E, g, wk);
    __syncwarp();
    for (int i = lane; i < P.cells; i += 32) w.board[i] = gb[i];
    if (P.hist_len) {
      const uint8_t* gh = E.hist + (size_t)g * 8 * E.cellsP;
      for (int i = lane; i < 8 * E.cellsP; i += 32) w.hist[i] = gh[i];
    }
    __syncwarp();
    int* path = E.path + ((size_t)g * E.V + wk) * (P.maxDepth + 1);
    St s;
    s.to_move = gi[GI_TO_MOVE];
    s.move_number = gi[GI_MOVE_NUMBER];
    s.passes = P.kind == KIND_WQ ? gi[GI_PASSES] : (P.kind == KIND_MNK ? -1 : 0);
    s.ko = gi[GI_KO];
    // positional superko: the hashes of the positions this descent walks through (entry 0 = the root's)
    unsigned long long* ph = E.pathhash ? E.pathhash + ((size_t)g * E.V + wk) * (P.maxDepth + 2) : nullptr;
    int npath = 0;
    if (ph && lane == 0) ph[0] = E.poshash[(size_t)g * (P.max_plies + 2) + gi[GI_N_POS]];
    int node = ti[TI_ROOT];
    int depth = 0, path_len = 0;
    int status = ST_DONE;
    bool is_null = true;
    bool analyzed = false;
    unsigned long long sel_children = 0, sel_levels = 0;

    while (true) {
      depth++;
      if (depth > P.maxDepth) break;  // search.go:211-215: null result, nothing on the path is updated
      const int player = s.to_move;
      if (lane == 0) { path[path_len] = node; if (vl) vl[node] = 1; }
      path_len++;
      const uint32_t meta = E.meta[tb + node];
      if (!META_EXPANDED(meta)) {
        if (s.passes >= 2) {  // search.go:226-228: terminal by passes -> combinedScore (utils.go:62-67)
          float ws_ = game_score(P, w.board, C_WHITE, lane, &w.wq);
          float bs_ = game_score(P, w.board, C_BLACK, lane, &w.wq);
          analyzed = false;
          float v = __fsub_rn(__fsub_rn(bs_, ws_), P.komi);
          __syncwarp();
          backup(E, tb, path, path_len, v, lane);
          is_null = false;
        } else {
          // leaf: hand the state to the evaluator
          uint8_t* lb = E.leaf_board + ((size_t)g * E.V + wk) * E.cellsP;
          for (int i = lane; i < P.cells; i += 32) lb[i] = w.board[i];
          if (E.inf[agent].kind == INF_DUAL) {
            float* out = E.nn_in + ((size_t)agent * E.GS + slot0 + (size_t)wk * slot_stride) * P.plane;
            encode_planes(P, w.board, w.hist, E.cellsP, s.to_move, s.move_number, out, lane);
          }
          if (lane == 0) { wv[WV_TO_MOVE] = s.to_move; wv[WV_MOVE_NUMBER] = s.move_number; wv[WV_PASSES] = s.passes; wv[WV_KO] = s.ko; wv[WV_NPATH] = npath; }
          status = ST_LEAF;
          is_null = false;
        }
        break;
      }
      // ---- Node.Select (node.go:170-237).  Child blocks start on a 4-node boundary, so each lane pulls
      // four consecutive children's N / W / P with one 128-bit load per array (coalesced 512 B per warp).
      const int nc = META_NCHILD(meta), first = E.first[tb + node];
      const uint32_t* Nb = E.N + tb + first;
      const float* Wb = E.W + tb + first;
      const float* Pb = E.Pr + tb + first;
      uint32_t pv = 0;
      for (int j0 = lane * 4; j0 < nc; j0 += 128) {
        const uint4 n4 = *reinterpret_cast<const uint4*>(Nb + j0);
        pv += n4.x;
        if (j0 + 1 < nc) pv += n4.y;
        if (j0 + 2 < nc) pv += n4.z;
        if (j0 + 3 < nc) pv += n4.w;
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) pv += __shfl_xor_sync(FULL, pv, off);
      const float numerator = __fsqrt_rn(__uint2float_rn(pv));
      float bestv = -INFINITY;
      int besti = 0x7fffffff;
      for (int j0 = lane * 4; j0 < nc; j0 += 128) {
        const uint4 n4 = *reinterpret_cast<const uint4*>(Nb + j0);
        const float4 w4 = *reinterpret_cast<const float4*>(Wb + j0);
        const float4 p4 = *reinterpret_cast<const float4*>(Pb + j0);
        const uint32_t nn[4] = {n4.x, n4.y, n4.z, n4.w};
        float ww[4] = {w4.x, w4.y, w4.z, w4.w};
        const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
        if (vl && player == C_WHITE) {  // Node.Evaluate: blackScores += VirtualLoss() for White only (node.go:150-152)
          const uint32_t f4 = *reinterpret_cast<const uint32_t*>(vl + first + j0);
#pragma unroll
          for (int i = 0; i < 4; i++)
            if ((f4 >> (8 * i)) & 0xffu) ww[i] = __fadd_rn(ww[i], 3.0f);  // virtualLoss1, mcts.go:27
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          if (j0 + i >= nc) break;
          const uint32_t visits = nn[i];
          const float qsa = evaluate(ww[i], visits, player);  // visits >= 1 always (tree.go:110): fpu is dead
          const float denominator = __fadd_rn(1.0f, __uint2float_rn(visits));
          const float lastTerm = __fdiv_rn(numerator, denominator);
          const float puct = __fmul_rn(__fmul_rn(P.puct, pp[i]), lastTerm);
          const float usa = __fadd_rn(qsa, puct);
          if (usa > bestv) { bestv = usa; besti = j0 + i; }  // strict >: earliest child wins ties
        }
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        float ov = __shfl_xor_sync(FULL, bestv, off);
        int oi = __shfl_xor_sync(FULL, besti, off);
        if (oi != 0x7fffffff && (besti == 0x7fffffff || ov > bestv || (ov == bestv && oi < besti))) { bestv = ov; besti = oi; }
      }
      sel_children += nc;
      sel_levels++;
      if (besti == 0x7fffffff) { raise(E, ERR_NO_ACTIVE_CHILD, lane); break; }
      const int next = first + besti;
      const int move = META_MOVE(E.meta[tb + next]);
      // (positional superko is not re-tested here: the child exists because expansion found the move legal against
      // the same history — game positions + this very path — so the reference's second Check cannot disagree)
      if (!state_check(P, w, player, move, lane, &analyzed, s.ko)) break;  // illegal: null result, no retry
      state_apply(P, w, E.cellsP, s, player, move, lane, &analyzed, nullptr, nullptr);
      if (ph) {
        const unsigned long long h = wq_pos_hash(P, w.board, E.zt64, lane);
        npath++;
        if (lane == 0) ph[npath] = h;
      }
      node = next;
    }
    if (vl && status != ST_LEAF) {  // the call returned: undoVirtualLoss on every node it entered (search.go:254)
      __syncwarp();
      for (int i = lane; i < path_len; i += 32) vl[path[i]] = 0;
    }
    count(E, CNT_SIMS, 1, lane);
    count(E, CNT_NULL, is_null ? 1 : 0, lane);
    count(E, CNT_SEL_CHILDREN, sel_children, lane);
    count(E, CNT_SEL_LEVELS, sel_levels, lane);
    if (lane == 0) { wv[WV_STATUS] = status; wv[WV_PATHLEN] = path_len; }
  }
}

// dummy.go / scripted-table evaluators, run on device for every pending leaf
__global__ void k_infer_simple(GameP P, EngineDev E, int n_games) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  const int* wv0 = WV_OF(E, g, 0);
  const int agent = wv0[WV_AGENT];
  const InfererDev inf = E.inf[agent];
  if (inf.kind == INF_DUAL) return;
  for (int wk = 0; wk < E.V; wk++) {
    const int* wv = WV_OF(E, g, wk);
    if (wv[WV_STATUS] != ST_LEAF) continue;
    const size_t slot = (size_t)agent * E.GS + wv0[WV_SLOT] + (size_t)wk * E.batch_count[2 + agent];
    float* pol = E.policy + slot * E.Lmax;
    float* val = E.value + slot;
    if (inf.kind == INF_DUMMY) {
      float p = __fdiv_rn(1.0f, (float)inf.L);
      for (int i = lane; i < inf.L; i += 32) pol[i] = p;
      if (lane == 0) *val = inf.dummy_value;
    } else {
      int mn = P.kind == KIND_C4 ? 1 : wv[WV_MOVE_NUMBER];  // c4/game.go:51: MoveNumber() is moveCount + 1, never advanced
      bool ok = mn >= 0 && mn < inf.table_rows;
      for (int i = lane; i < inf.L; i += 32) pol[i] = ok ? inf.table[(size_t)mn * inf.L + i] : 0.0f;
      if (lane == 0) *val = ok ? inf.table_values[mn] : 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2: expandAndSimulate (search.go:259-339) on the evaluated leaf, then Update along the path.
__global__ void k_expand_backup(GameP P, EngineDev E, int n_games) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  const int* wv0 = WV_OF(E, g, 0);
  WS w = make_ws(P, E.cellsP, smem + (size_t)wib * ws_bytes(P, E.cellsP));
  const int agent = wv0[WV_AGENT], t = wv0[WV_TREE];
  const int L = E.inf[agent].L;
  const size_t tb = ((size_t)g * E.T + t) * (size_t)P.max_nodes;
  int* ti = E.ti + ((size_t)g * E.T + t) * TI_COUNT;
  // pending workers finish in start order (with V == 1: the one descent of this wave)
  for (int wk = 0; wk < E.V; wk++) {
  int* wv = WV_OF(E, g, wk);
  __syncwarp();
  if (wv[WV_STATUS] != ST_LEAF) continue;
  const uint8_t* lb = E.leaf_board + ((size_t)g * E.V + wk) * E.cellsP;
  for (int i = lane; i < P.cells; i += 32) w.board[i] = lb[i];
  __syncwarp();
  const size_t slot = (size_t)agent * E.GS + wv0[WV_SLOT] + (size_t)wk * E.batch_count[2 + agent];
  const float* pol = E.policy + slot * E.Lmax;
  float value = E.value[slot];
  const int player = wv[WV_TO_MOVE];
  if (player == C_WHITE) value = __fsub_rn(1.0f, value);  // search.go:278-280
  const int* path = E.path + ((size_t)g * E.V + wk) * (P.maxDepth + 1);
  const int path_len = wv[WV_PATHLEN];
  const int leaf = path[path_len - 1];
  count(E, CNT_EVALS, 1, lane);
  // a worker that reached a leaf an earlier worker of this round has expanded meanwhile: every candidate is found
  // by findChild / oldMinPsa is 0 (search.go:316-325) — nothing is created, the value is still backed up
  const bool already = META_EXPANDED(E.meta[tb + leaf]) != 0;

  // legal list in index order, then Pass (search.go:285-296)
  w.wq.zt64 = E.zt64;  // complete rules: the analysis also leaves the group hashes
  if (P.kind == KIND_WQ && !already) wq_analyze(P, w.board, w.wq, lane);
  const unsigned long long* gh_ = E.poshash ? E.poshash + (size_t)g * (P.max_plies + 2) : nullptr;
  const unsigned long long* ph_ = E.pathhash ? E.pathhash + ((size_t)g * E.V + wk) * (P.maxDepth + 2) : nullptr;
  const int n_pos = gh_ ? E.gi[(size_t)g * GI_COUNT + GI_N_POS] : 0, npath = gh_ ? wv[WV_NPATH] : 0;
  int nleg = 0;
  for (int base = 0; base < (already ? 0 : P.A); base += 32) {
    int i = base + lane;
    bool legal = false;
    if (i < P.A) {
      if (P.kind == KIND_WQ) { bool cap; legal = wq_check_pt(P, w.board, w.wq, i, player, &cap, wv[WV_KO]); }
      else legal = simple_check(P, w.board, i);
      if (legal && gh_) {  // positional superko: the move may not recreate a position of the game or of this path
        const unsigned long long h = wq_hash_after(P, w.board, w.wq, i, player, ph_[npath]);
        for (int j = 0; j < n_pos && legal; j++) legal = gh_[j] != h;
        for (int j = 0; j < npath && legal; j++) legal = ph_[j] != h;
      }
    }
    unsigned m = __ballot_sync(FULL, legal);
    if (legal) {
      int pos = nleg + __popc(m & ((1u << lane) - 1));
      w.fa[pos] = pol[i];
      w.ia[pos] = i;
    }
    nleg += __popc(m);
  }
  if (P.kind != KIND_MNK && !already) {  // Check(Pass): c4/game.go:53 and wq/game.go:69 accept, mnk.go:102 rejects
    if (lane == 0) { w.fa[nleg] = pol[L - 1]; w.ia[nleg] = MV_PASS; }
    nleg++;
  }
  __syncwarp();
  // legalSum in list order, fp32, sequential (search.go:288,295)
  float legalSum = 0.0f;
  if (lane == 0) for (int j = 0; j < nleg; j++) legalSum = __fadd_rn(legalSum, w.fa[j]);
  legalSum = __shfl_sync(FULL, legalSum, 0);
  if (legalSum > __int_as_float(1) /* math32.SmallestNonzeroFloat32 */) {
    for (int j = lane; j < nleg; j += 32) w.fa[j] = __fdiv_rn(w.fa[j], legalSum);
  } else {
    float prob = __fdiv_rn(1.0f, (float)nleg);
    for (int j = lane; j < nleg; j += 32) w.fa[j] = prob;
  }
  __syncwarp();
  if (nleg > 0) {
    // sort.Sort(byScore) pinned to a stable descending sort: rank by counting
    bool nan = false;
    for (int i = lane; i < nleg; i += 32) {
      float si = w.fa[i];
      nan |= (si != si);
      int rank = 0;
      for (int j = 0; j < nleg; j++) {
        float sj = w.fa[j];
        rank += (sj > si) || (sj == si && j < i);
      }
      w.fb[rank] = si;
      w.ib[rank] = w.ia[i];
    }
    if (__any_sync(FULL, nan)) raise(E, ERR_NAN_PRIOR, lane);
    __syncwarp();
    const float maxPsa = w.fb[0];
    const float oldMinPsa = __fmul_rn(maxPsa, 2.0f);  // n.MinPsaRatio() of an unexpanded node
    const float newMinPsa = __fmul_rn(maxPsa, 0.0f);  // minPsaRatio() == 0 (tree far below 50% of 25M nodes)
    // children = entries with !(score < newMinPsa) && score < oldMinPsa, in sorted order
    int nkeep = 0;
    for (int base = 0; base < nleg; base += 32) {
      int j = base + lane;
      bool keep = j < nleg && !(w.fb[j] < newMinPsa) && (w.fb[j] < oldMinPsa);
      unsigned m = __ballot_sync(FULL, keep);
      if (keep) w.st[0][j] = nkeep + __popc(m & ((1u << lane) - 1));
      else if (j < nleg) w.st[0][j] = 0xffffffffu;
      nkeep += __popc(m);
    }
    __syncwarp();
    int a = nkeep > 0 ? alloc_nodes(E, P, ti, nkeep, lane) : 0;
    if (a >= 0) {
      for (int j = lane; j < nleg; j += 32) {
        uint32_t pos = w.st[0][j];
        if (pos == 0xffffffffu) continue;
        size_t ci = tb + a + pos;
        E.N[ci] = 1u; E.W[ci] = 0.0f; E.Pr[ci] = w.fb[j];
        E.meta[ci] = META_MAKE(w.ib[j], 0, 0); E.first[ci] = -1;
      }
      if (lane == 0) {
        E.meta[tb + leaf] = META_MAKE(META_MOVE(E.meta[tb + leaf]), nkeep, 1);
        E.first[tb + leaf] = nkeep > 0 ? a : -1;
      }
      count(E, CNT_CREATED, nkeep, lane);
    }
  }
  __syncwarp();
  backup(E, tb, path, path_len, value, lane);
  if (E.vl) for (int i = lane; i < path_len; i += 32) E.vl[tb + path[i]] = 0;  // undoVirtualLoss (search.go:254)
  if (lane == 0) wv[WV_STATUS] = ST_DONE;
  }
}

// ---------------------------------------------------------------------------------------------
// K9: MCTS.Search epilogue + the body of Arena.Play's loop.
__global__ void k_search_end(GameP P, EngineDev E, int n_games, int record) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + wib;
  if (g >= n_games) return;
  int* gi = E.gi + (size_t)g * GI_COUNT;
  int* wv = WV_OF(E, g, 0);
  if (lane == 0) E.ex_valid[g] = 0;
  if (!gi[GI_ACTIVE]) return;
  WS w = make_ws(P, E.cellsP, smem + (size_t)wib * ws_bytes(P, E.cellsP));
  uint8_t* gb = E.board + (size_t)g * E.cellsP;
  for (int i = lane; i < P.cells; i += 32) w.board[i] = gb[i];
  if (P.hist_len) {
    const uint8_t* gh = E.hist + (size_t)g * 8 * E.cellsP;
    for (int i = lane; i < 8 * E.cellsP; i += 32) w.hist[i] = gh[i];
  }
  __syncwarp();
  const int t = wv[WV_TREE], player = wv[WV_PLAYER];
  const size_t tb = ((size_t)g * E.T + t) * (size_t)P.max_nodes;
  int* ti = E.ti + ((size_t)g * E.T + t) * TI_COUNT;
  const int root = ti[TI_ROOT];
  const uint32_t rmeta = E.meta[tb + root];
  if (!META_EXPANDED(rmeta)) {  // search.go:141-149 (argmax-of-policy fallback): unreachable for an un-ended game
    raise(E, ERR_ROOT_NO_CHILDREN, lane);
    if (lane == 0) gi[GI_ACTIVE] = 0;
    return;
  }
  // ---- bestMove (search.go:341-390): stable sort of the root's children by fancySort (utils.go:10-47)
  const int nc = META_NCHILD(rmeta), first = E.first[tb + root];
  int best = MV_PASS;
  bool analyzed = false;
  if (nc > 0) {
    for (int j = lane; j < nc; j += 32) {
      size_t ci = tb + first + j;
      w.st[0][j] = E.N[ci]; w.st[1][j] = __float_as_uint(E.W[ci]); w.st[2][j] = __float_as_uint(E.Pr[ci]);
      w.st[3][j] = E.meta[ci]; w.st[4][j] = (uint32_t)E.first[ci];
      w.fa[j] = evaluate(E.W[ci], E.N[ci], player);
    }
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {
      uint32_t ni = w.st[0][i];
      float pi = __uint_as_float(w.st[2][i]), ei = w.fa[i];
      int rank = 0;
      for (int j = 0; j < nc; j++) {
        uint32_t nj = w.st[0][j];
        bool less_ji, less_ij;  // Less(j,i), Less(i,j)
        if (nj != ni) { less_ji = nj > ni; less_ij = ni > nj; }
        else if (ni == 0) { float pj = __uint_as_float(w.st[2][j]); less_ji = pj > pi; less_ij = pi > pj; }
        else { float ej = w.fa[j]; less_ji = ej > ei; less_ij = ei > ej; }
        rank += less_ji || (!less_ij && !less_ji && j < i);
      }
      w.ia[i] = rank;
    }
    __syncwarp();
    if (P.random_count > 0 && (P.kind == KIND_C4 ? 1 : gi[GI_MOVE_NUMBER]) < P.random_count) {
      // randomizeChildren (tree.go:212-247) on the sorted list: temperature sample, then the reference's
      // swap loop.  Sequential by construction (cumulative fp32 sums, one RNG draw): one lane.
      for (int i = lane; i < nc; i += 32) w.ib[w.ia[i]] = i;  // sorted position -> original child
      __syncwarp();
      if (lane == 0) {
        for (int p = 0; p < nc; p++) w.ic[p] = p;  // content (sorted index) at each list position
        float accum = 0.0f, norm = 0.0f;
        int nacc = 0;
        bool abort = false;
        for (int p = 0; p < nc; p++) {
          const uint32_t visits = w.st[0][w.ib[p]];
          if (norm == 0.0f) {
            norm = __uint2float_rn(visits);
            if (visits <= P.random_min_visits) { abort = true; break; }
          }
          if (visits > P.random_min_visits) {
            // math32.Pow = float32(math.Pow(float64(x), float64(y)))
            const float x = __fdiv_rn(__uint2float_rn(visits), norm), y = __fdiv_rn(1.0f, P.random_temperature);
            accum = __fadd_rn(accum, (float)pow((double)x, (double)y));
            w.fb[nacc++] = accum;
          }
        }
        if (!abort) {
          unsigned long long rs = ((unsigned long long)(unsigned)ti[TI_RNG_HI] << 32) | (unsigned)ti[TI_RNG_LO];
          unsigned long long z = (rs += 0x9E3779B97F4A7C15ull);
          z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
          z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
          z = z ^ (z >> 31);
          ti[TI_RNG_LO] = (int)(unsigned)(rs & 0xffffffffu); ti[TI_RNG_HI] = (int)(unsigned)(rs >> 32);
          const float rnd = __fmul_rn((float)(z >> 40) * (1.0f / 16777216.0f), accum);
          int index = 0;
          for (int i = 0; i < nacc; i++) if (rnd < w.fb[i]) { index = i; break; }
          if (index != 0)
            for (int i = 0; i < nc - index; i++) { int t0 = w.ic[i]; w.ic[i] = w.ic[i + index]; w.ic[i + index] = t0; }
        }
        for (int p = 0; p < nc; p++) w.ib[w.ic[p]] = p;  // sorted index -> final position (reuses ib)
      }
      __syncwarp();
      for (int i = lane; i < nc; i += 32) w.ic[i] = w.ib[w.ia[i]];
      __syncwarp();
      for (int i = lane; i < nc; i += 32) w.ia[i] = w.ic[i];
      __syncwarp();
    }
    for (int i = lane; i < nc; i += 32) {  // the sort is in place in the reference (children slice aliases)
      size_t ci = tb + first + w.ia[i];
      E.N[ci] = w.st[0][i]; E.W[ci] = __uint_as_float(w.st[1][i]); E.Pr[ci] = __uint_as_float(w.st[2][i]);
      E.meta[ci] = w.st[3][i]; E.first[ci] = (int)w.st[4][i];
      w.ib[w.ia[i]] = META_MOVE(w.st[3][i]);  // moves in sorted order
      w.fb[w.ia[i]] = w.fa[i];                 // Evaluate(player) in sorted order
    }
    __syncwarp();
    best = w.ib[0];
    float bestScore = w.fb[0];
    // search.go:366-389: pass preferences, then resignation
    const float rootScore = E.Pr[tb + root];
    const bool passing_loses = (rootScore > 0.0f && player == C_WHITE) || (rootScore < 0.0f && player == C_BLACK);
    bool want_nopass = false;
    if (P.dont_prefer_pass && best == MV_PASS) want_nopass = true;
    else if (!P.dumb_pass && best == MV_PASS) want_nopass = passing_loses;
    else if (!P.dumb_pass && gi[GI_LAST_MOVE] == MV_PASS) { if (!passing_loses) best = MV_PASS; }
    if (want_nopass) {  // noPassBestMove (search.go:538-563): first child in sorted order that is not Pass and is legal
      if (P.kind == KIND_WQ) { wq_analyze(P, w.board, w.wq, lane); analyzed = true; }
      int found = -1;
      for (int base = 0; base < nc && found < 0; base += 32) {
        int j = base + lane;
        bool ok = false;
        if (j < nc) {
          int mv = w.ib[j];
          if (mv != MV_PASS) {
            if (P.kind == KIND_WQ) { bool cap; ok = (mv == MV_RESIGN) || (mv < P.cells && wq_check_pt(P, w.board, w.wq, mv, player, &cap, gi[GI_KO])); }
            else ok = simple_check(P, w.board, mv);
          }
        }
        unsigned m = __ballot_sync(FULL, ok);
        if (m) found = base + __ffs(m) - 1;
      }
      if (found >= 0) { best = w.ib[found]; bestScore = w.fb[found]; }  // visits >= 1 always: never the "not visited" 1.0
    }
    // shouldResign (search.go:502-535)
    if (best == MV_PASS && !P.dont_resign && P.resign_pct != 0.0f) {
      const int move_number = P.kind == KIND_C4 ? 1 : gi[GI_MOVE_NUMBER];
      const float thr = P.resign_pct < 0.0f ? 0.1f : P.resign_pct;
      if (move_number > P.maxDepth / 4 && !(bestScore > thr)) best = MV_RESIGN;
    }
  }
  if (best == MV_RESIGN) {  // arena.go:127: Apply(Resign) indexes board[-2] in mnk, c4 and wq alike
    raise(E, ERR_RESIGN_APPLIED, lane);
    if (lane == 0) gi[GI_ACTIVE] = 0;
    return;
  }
  // ---- t.prev = clone(current); cachedPolicies[{hash, best}]++ (search.go:152,161)
  const uint32_t hash = (uint32_t)wv[WV_HASH];
  int npol = ti[TI_NPOL];
  uint32_t* ph = E.pol_hash + ((size_t)g * E.T + t) * P.max_plies;
  int16_t* pm = E.pol_move + ((size_t)g * E.T + t) * P.max_plies;
  if (npol >= P.max_plies) { raise(E, ERR_PATH_OVERFLOW, lane); if (lane == 0) gi[GI_ACTIVE] = 0; return; }
  if (lane == 0) {
    ph[npol] = hash; pm[npol] = (int16_t)best;
    ti[TI_NPOL] = npol + 1;
    ti[TI_PREV_VALID] = 1; ti[TI_PREV_MN] = gi[GI_MOVE_NUMBER];
  }
  npol++;
  __syncwarp();
  // ---- Arena.Play body (arena.go:99-137)
  int arena_pass = gi[GI_ARENA_PASS];
  arena_pass = best == MV_PASS ? arena_pass + 1 : 0;
  if (record) {
    // Example{Board: Enc(game), Policy: MCTS.Policies(game), Value: colour} (arena.go:105-121)
    encode_planes(P, w.board, w.hist, E.cellsP, gi[GI_TO_MOVE], gi[GI_MOVE_NUMBER], E.ex_board + (size_t)g * P.plane, lane);
    float* ep = E.ex_policy + (size_t)g * (P.A + 1);
    for (int i = lane; i <= P.A; i += 32) w.fa[i] = 0.0f;
    __syncwarp();
    int total = 0;
    for (int e = lane; e < npol; e += 32) {
      int mv = pm[e];
      if (ph[e] == hash && mv >= 0 && mv <= P.A) { atomicAdd(&w.fa[mv], 1.0f); total++; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(FULL, total, off);
    __syncwarp();
    const float sum = (float)total;  // tree.go:134-138: sum of small integer counts, exact in any order
    bool bad = false;
    for (int i = lane; i <= P.A; i += 32) {
      float v = __fdiv_rn(w.fa[i], sum);
      ep[i] = v;
      bad |= !(fabsf(v) <= 3.4028234663852886e38f);  // NaN or Inf (arena.go:241-251)
    }
    bad = __any_sync(FULL, bad);
    if (lane == 0) { E.ex_valid[g] = bad ? 0 : 1; E.ex_value[g] = (float)player; gi[GI_N_EX] += bad ? 0 : 1; }
  }
  // game.Apply(PlayerMove{player, best}) (arena.go:127-130)
  St s;
  s.to_move = gi[GI_TO_MOVE]; s.move_number = gi[GI_MOVE_NUMBER];
  s.passes = P.kind == KIND_WQ ? gi[GI_PASSES] : (P.kind == KIND_MNK ? -1 : 0);
  s.ko = gi[GI_KO];
  int zhash = gi[GI_ZHASH];
  int c4pass = gi[GI_C4_PASS];
  int last_move = gi[GI_LAST_MOVE];  // State.LastMove(): the last entry State.Apply appended to history
  if (P.kind == KIND_MNK) {
    if (simple_check(P, w.board, best)) {
      if (lane == 0) E.hmoves[(size_t)g * P.max_plies + s.move_number] = (int16_t)best;
      state_apply(P, w, E.cellsP, s, player, best, lane, &analyzed, nullptr, nullptr);
      last_move = best;
    }
  } else if (P.kind == KIND_C4) {
    if (best == MV_PASS || (best >= 0 && best < P.n && c4_drop_row(P, w.board, best) >= 0)) {
      state_apply(P, w, E.cellsP, s, player, best, lane, &analyzed, nullptr, nullptr);
      last_move = best;
    }
    c4pass = best == MV_PASS ? c4pass + 1 : 0;  // c4/game.go:66-70
  } else {
    state_apply(P, w, E.cellsP, s, player, best, lane, &analyzed, &zhash, E.ztable);
    last_move = best;
    if (E.poshash) {  // positional superko: the position just left joins the game's list, the new one follows it
      const unsigned long long h = wq_pos_hash(P, w.board, E.zt64, lane);
      if (lane == 0) {
        const int np = gi[GI_N_POS] + 1;
        if (np <= P.max_plies) { E.poshash[(size_t)g * (P.max_plies + 2) + np] = h; gi[GI_N_POS] = np; }
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < P.cells; i += 32) gb[i] = w.board[i];
  if (P.hist_len) {
    uint8_t* gh = E.hist + (size_t)g * 8 * E.cellsP;
    for (int i = lane; i < 8 * E.cellsP; i += 32) gh[i] = w.hist[i];
  }
  const int n_moves = gi[GI_N_MOVES];
  int winner = gi[GI_WINNER];
  bool active = true;
  if (arena_pass >= 2) {                           // arena.go:135-137: break before Ended() is re-evaluated (winner stays None)
    active = false;
    if (P.wq_complete) { int wn; game_ended(P, w.board, 2, lane, &wn, &w.wq); winner = wn; }  // OUR complete rules: score it
  }
  else if (P.max_moves > 0 && n_moves + 1 >= P.max_moves) active = false;  // COMPLETION: move cap
  else {
    int wn;
    bool ended = game_ended(P, w.board, P.kind == KIND_C4 ? c4pass : s.passes, lane, &wn, &w.wq);
    winner = wn;
    active = !ended;
  }
  // the per-game move list is full but the reference would keep playing (no move cap set): an engine limit, reported —
  // silently ending the game here would count it as a draw and label its examples 0
  if (active && n_moves + 1 >= P.max_plies) { raise(E, ERR_PATH_OVERFLOW, lane); active = false; }
  if (lane == 0) {
    E.moves[(size_t)g * P.max_plies + n_moves] = (int16_t)best;
    gi[GI_N_MOVES] = n_moves + 1;
    gi[GI_TO_MOVE] = s.to_move; gi[GI_MOVE_NUMBER] = s.move_number;
    if (P.kind == KIND_WQ) { gi[GI_PASSES] = s.passes; gi[GI_KO] = s.ko; }
    gi[GI_ZHASH] = zhash; gi[GI_C4_PASS] = c4pass; gi[GI_LAST_MOVE] = last_move;
    gi[GI_ARENA_PASS] = arena_pass;
    gi[GI_CUR_AGENT] ^= 1;  // switchPlayer
    gi[GI_ACTIVE] = active ? 1 : 0;
    gi[GI_WINNER] = winner;
    wv[WV_STATUS] = ST_IDLE;
  }
}

// ---------------------------------------------------------------------------------------------
// stateless rules evaluation for az_rules_apply / az_rules_status
__global__ void k_rules_apply(GameP P, int cellsP, int n, const int* __restrict__ boards, const int* __restrict__ players,
                              const int* __restrict__ moves, int* check, int* applied, int* out_boards, int* taken,
                              const int* __restrict__ ztable) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + wib;
  if (q >= n) return;
  WS w = make_ws(P, cellsP, smem + (size_t)wib * ws_bytes(P, cellsP));
  for (int i = lane; i < P.cells; i += 32) w.board[i] = (uint8_t)boards[(size_t)q * P.cells + i];
  __syncwarp();
  const int player = players[q], move = moves[q];
  int ck = 0, ap = 0, tk = 0;
  if (P.kind == KIND_MNK) {
    ck = simple_check(P, w.board, move);
    ap = ck && move >= 0;
    if (ap && lane == 0) w.board[move] = (uint8_t)player;
  } else if (P.kind == KIND_C4) {
    ck = simple_check(P, w.board, move);
    ap = ck;
    if (ck && move != MV_PASS && lane == 0) w.board[c4_drop_row(P, w.board, move) * P.n + move] = (uint8_t)player;
  } else {
    if (move == MV_PASS) { ck = 1; ap = 1; }
    else {
      wq_analyze(P, w.board, w.wq, lane);
      bool cap;
      if ((player == C_BLACK || player == C_WHITE) && move >= 0 && move < P.cells)
        ck = wq_check_pt(P, w.board, w.wq, move, player, &cap);
      ap = wq_board_apply(P, w.board, w.wq, move, player, lane, &tk, nullptr, nullptr);
      tk &= 0xff;  // byte(len(captures))
    }
  }
  __syncwarp();
  for (int i = lane; i < P.cells; i += 32) out_boards[(size_t)q * P.cells + i] = w.board[i];
  if (lane == 0) { check[q] = ck; applied[q] = ap; taken[q] = tk; }
}

__global__ void k_rules_status(GameP P, int cellsP, int n, const int* __restrict__ boards, const int* __restrict__ passes,
                               int* ended, int* winner, float* sb, float* sw) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + wib;
  if (q >= n) return;
  WS w = make_ws(P, cellsP, smem + (size_t)wib * ws_bytes(P, cellsP));
  for (int i = lane; i < P.cells; i += 32) w.board[i] = (uint8_t)boards[(size_t)q * P.cells + i];
  __syncwarp();
  int wn;
  bool e = game_ended(P, w.board, passes[q], lane, &wn, &w.wq);
  float b = game_score(P, w.board, C_BLACK, lane, &w.wq), wh = game_score(P, w.board, C_WHITE, lane, &w.wq);
  if (lane == 0) { ended[q] = e; winner[q] = wn; sb[q] = b; sw[q] = wh; }
}

// ---------------------------------------------------------------------------------------------
// host-callable launchers
static inline dim3 grid_for(int n, int wpb) { return dim3((n + wpb - 1) / wpb); }
static const int WPB = 4;

// Raised to the device's opt-in maximum (not to this engine's need): the attribute is per function and per device,
// and engines of different board sizes share both.
void mcts_set_smem_limits(const GameP& P, int cellsP) {
  int dev = 0, optin = 0;
  CUDA_CHECK(cudaGetDevice(&dev));
  CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)(ws_bytes(P, cellsP) * WPB) > optin) throw std::runtime_error("board too large for the per-warp shared-memory workspace");
  CUDA_CHECK(cudaFuncSetAttribute(k_arena_begin, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_search_begin, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_expand_backup, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_search_end, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_rules_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CUDA_CHECK(cudaFuncSetAttribute(k_rules_status, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
}
#define SMEM(P, E) (ws_bytes(P, (E).cellsP) * WPB)
void launch_arena_begin(const GameP& P, const EngineDev& E, int n_games, const int* coins, unsigned long long game_base, cudaStream_t s) {
  k_arena_begin<<<grid_for(E.G, WPB), WPB * 32, SMEM(P, E), s>>>(P, E, n_games, coins, game_base); LAUNCH_CHECK();
}
void launch_assign_slots(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_assign_slots<<<1, 32, 0, s>>>(E, n_games, P.shared_tree); LAUNCH_CHECK();
}
void launch_search_begin(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_search_begin<<<grid_for(n_games, WPB), WPB * 32, SMEM(P, E), s>>>(P, E, n_games); LAUNCH_CHECK();
}
void launch_encode_roots(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_encode_roots<<<grid_for(n_games, WPB), WPB * 32, 0, s>>>(P, E, n_games); LAUNCH_CHECK();
}
void launch_select(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_select<<<grid_for(n_games, WPB), WPB * 32, SMEM(P, E), s>>>(P, E, n_games); LAUNCH_CHECK();
}
void launch_infer_simple(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_infer_simple<<<grid_for(n_games, WPB), WPB * 32, 0, s>>>(P, E, n_games); LAUNCH_CHECK();
}
void launch_expand_backup(const GameP& P, const EngineDev& E, int n_games, cudaStream_t s) {
  k_expand_backup<<<grid_for(n_games, WPB), WPB * 32, SMEM(P, E), s>>>(P, E, n_games); LAUNCH_CHECK();
}
void launch_search_end(const GameP& P, const EngineDev& E, int n_games, int record, cudaStream_t s) {
  k_search_end<<<grid_for(n_games, WPB), WPB * 32, SMEM(P, E), s>>>(P, E, n_games, record); LAUNCH_CHECK();
}
void launch_rules_apply(const GameP& P, int cellsP, int n, const int* boards, const int* players, const int* moves,
                        int* check, int* applied, int* out_boards, int* taken, cudaStream_t s) {
  k_rules_apply<<<grid_for(n, WPB), WPB * 32, ws_bytes(P, cellsP) * WPB, s>>>(P, cellsP, n, boards, players, moves, check,
                                                                            applied, out_boards, taken, nullptr); LAUNCH_CHECK();
}
void launch_rules_status(const GameP& P, int cellsP, int n, const int* boards, const int* passes, int* ended, int* winner,
                         float* sb, float* sw, cudaStream_t s) {
  k_rules_status<<<grid_for(n, WPB), WPB * 32, ws_bytes(P, cellsP) * WPB, s>>>(P, cellsP, n, boards, passes, ended, winner,
                                                                             sb, sw); LAUNCH_CHECK();
}
