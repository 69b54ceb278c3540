// agogo_b200 — device-side data model of the search engine (shared by mcts.cu and engine.cu).
//
// One warp per game.  Every tree is a flat SoA node pool in HBM (K1/K2/K9 of SURVEY.md §2):
//   N[] visits (u32) · W[] sum of backed-up values, black's perspective (f32) · P[] prior (f32)
//   meta[] = move+2 (16 bit) | n_children (15 bit) | expanded (1 bit) · first[] index of the
//   contiguous child block.  Children of a node are contiguous and ordered by prior (desc,
//   stable) — the order is semantically significant (first-max tie-break, node.go:226).
// The reference's Node (node.go:32-49) maps as: visits->N, blackScores->W, score->P,
// minPSARatioChildren in {2.0, 0}->expanded bit, status->implicit (detached subtrees are
// unreachable), virtualLoss->dead under the canonical 1-worker semantics, value->unused (fpu is
// unreachable because nodes are born with visits=1, tree.go:110).  Q is recomputed per Select
// because Node.Evaluate is colour-dependent (node.go:147-159); parent links are replaced by the
// per-descent path stack.
#pragma once
#include "common.cuh"

enum { GI_TO_MOVE, GI_MOVE_NUMBER, GI_PASSES, GI_ACTIVE, GI_WINNER, GI_ARENA_PASS, GI_A_PLAYER, GI_CUR_AGENT,
       GI_N_MOVES, GI_ZHASH, GI_N_EX, GI_C4_PASS, GI_N_HMOVES, GI_LAST_MOVE, GI_KO, GI_N_POS, GI_COUNT = 16 };
enum { TI_ROOT, TI_ALLOC, TI_PREV_VALID, TI_PREV_MN, TI_NPOL, TI_RNG_LO, TI_RNG_HI, TI_COUNT = 8 };
enum { WV_STATUS, WV_PATHLEN, WV_TO_MOVE, WV_MOVE_NUMBER, WV_PASSES, WV_SLOT, WV_AGENT, WV_FLAGS, WV_TREE,
       WV_PLAYER, WV_HASH, WV_KO, WV_NPATH, WV_COUNT = 16 };
enum { ST_IDLE = 0, ST_LEAF = 1, ST_DONE = 2 };
enum { INF_DUAL = 0, INF_DUMMY = 1, INF_TABLE = 2 };
enum { CNT_SEARCHES, CNT_SIMS, CNT_NULL, CNT_EVALS, CNT_SEL_CHILDREN, CNT_SEL_LEVELS, CNT_CREATED, CNT_BACKUP,
       CNT_COUNT = 16 };

#define WV_OF(E, g, l) ((E).wv + ((size_t)(g) * (E).V + (l)) * WV_COUNT)
#define META_MOVE(m) ((int)((m) & 0xFFFFu) - 2)
#define META_NCHILD(m) ((int)(((m) >> 16) & 0x7FFFu))
#define META_EXPANDED(m) (((m) >> 31) & 1u)
#define META_MAKE(move, nchild, expanded) \
  ((uint32_t)(((move) + 2) & 0xFFFF) | ((uint32_t)(nchild) << 16) | ((uint32_t)(expanded) << 31))

struct InfererDev {
  int kind;            // INF_*
  int L;               // policy row length: dual action_space, dummy A, table row_len
  float dummy_value;   // dummy.go: by captured colour
  const float* table;  // [n_rows][L]
  const float* table_values;
  int table_rows;
};

struct EngineDev {
  // per game
  uint8_t* board;   // [G][cellsP]
  uint8_t* hist;    // [G][8][cellsP]  boards before the last 8 moves (wq18 encoder)
  int32_t* gi;      // [G][GI_COUNT]
  int16_t* moves;   // [G][max_plies]  Arena move record
  int16_t* hmoves;  // [G][max_plies]  State history (successfully applied moves; mnk tree reuse)
  // per tree (T trees per game)
  int32_t* ti;         // [G*T][TI_COUNT]
  uint32_t* pol_hash;  // [G*T][max_plies]   cachedPolicies keys (tree.go:75, search.go:161)
  int16_t* pol_move;   // [G*T][max_plies]
  uint32_t* N; float* W; float* Pr; uint32_t* meta; int32_t* first;  // [G*T][max_nodes]
  // per wave
  int32_t* wv;          // [G][WV_COUNT]
  int32_t* path;        // [G][maxDepth+1]
  uint8_t* leaf_board;  // [G][cellsP]
  // evaluation batches, one per agent
  int32_t* batch_count;  // [4]: [0..1] evaluation batch sizes, [2..3] active games per agent
  float* nn_in;          // [2][G][plane]   fp32 planes, NCHW
  float* policy;         // [2][G][Lmax]
  float* value;          // [2][G]
  // example staging for one ply
  float* ex_board;   // [G][plane]
  float* ex_policy;  // [G][A+1]
  float* ex_value;   // [G]
  int32_t* ex_valid; // [G]
  // misc
  const int32_t* ztable;  // [cells][2] wq zobrist
  // positional superko (AZ_FLAG_WQ_COMPLETE only, else null): 64-bit position keys; per game the hashes of the positions
  // before the root (gi[GI_N_POS] of them) followed by the root position's own; per descent the positions along the path
  // (entry d = after d in-tree moves, entry 0 = the root position, entry wv[WV_NPATH] = the leaf)
  const unsigned long long* zt64;  // [cells][2]
  unsigned long long* poshash;     // [G][max_plies + 2]
  unsigned long long* pathhash;    // [G*V][maxDepth + 2]
  int32_t* err;           // error bits
  unsigned long long* counters;  // [CNT_COUNT]
  int32_t* n_active;      // [1]
  int cellsP, T, Lmax, G;
  // concurrent pipeline calls per tree (mcts.Config workers, search.go:112-130): every per-wave record exists V
  // times per game, evaluation batches hold GS = G*V slots per agent (slot = rank + worker * active games of the agent)
  int V, GS;
  uint8_t* vl;                 // [G*T][max_nodes] virtual-loss flags (node.go:41), nullptr when V == 1
  const int32_t* round_workers;  // [1] workers started in the current round (the last round of a search may be short)
  unsigned long long tree_seed;  // MCTS.rand seed (tree.go:84), injected
  InfererDev inf[2];
};
