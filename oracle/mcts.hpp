// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU restatement of the reference's mcts package: mcts/tree.go, mcts/node.go, mcts/search.go,
// mcts/utils.go, mcts/mcts.go.
//
// Canonical semantics: ONE worker, exactly Config.Sims iterations of
// `pipeline` per Search, in place of the reference's runtime.NumCPU() goroutines racing a
// wall-clock Timeout (search.go:112-133).  Go's unstable sort.Sort (search.go:314,353) is pinned
// to a stable sort.  Every other line follows the reference, dead branches included.
// Compile with -ffp-contract=off: Go/amd64 never fuses multiply-add.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <utility>

#include "game.hpp"
#include "rng.hpp"

namespace oracle {

// mcts/mcts.go:15-18
struct Inferencer {
  virtual ~Inferencer() {}
  virtual void Infer(const State& state, std::vector<float>* policy, float* value) = 0;
};

enum PassPreference { DontPreferPass = 0, PreferPass = 1, DontResign = 2 };  // mcts/mcts.go:31-38

struct MCTSConfig {  // mcts/tree.go:15-29 (+ Sims: the fixed-iteration mode the reference lacks)
  float PUCT = 1.0f;
  int64_t Timeout = 0;
  int M = 0, N = 0;
  int RandomCount = 0;
  int32_t Budget = 0;
  uint32_t RandomMinVisits = 0;
  float RandomTemperature = 0;
  bool DumbPass = true;
  float ResignPercentage = 0;
  int PassPref = DontPreferPass;
  int Sims = 0;
  int Workers = 1;  // concurrent pipeline calls per tree (the reference starts runtime.NumCPU(), search.go:112-130)
  bool IsValid() const { return PUCT > 0 && PUCT <= 1; }  // tree.go:43-45
};

static const int MAXTREESIZE = 25000000;  // search.go:22-24
static const int nilNode = -1;            // naughty.go:8-10

enum Status : uint32_t { Invalid = 0, Active = 1, Pruned = 2 };  // node.go:14-18

struct Node {  // node.go:32-49
  int32_t move = 0;
  uint32_t visits = 0;
  uint32_t status = 0;
  float blackScores = 0, virtualLoss = 0, minPSARatioChildren = 2.0f, score = 0, value = 0;
  int id = 0;
};

struct Counters {
  uint64_t sims = 0, null_results = 0, evals = 0, select_children = 0, select_levels = 0,
           created = 0, backup_nodes = 0, searches = 0;
};

struct MCTS {
  MCTSConfig conf;
  Inferencer* nn;
  Rng rnd;
  std::vector<Node> nodes;
  std::vector<std::vector<int>> children;
  std::vector<int> freelist, freeables;
  // searchState (search.go:53-63)
  State* current;               // NOT owned (the Arena's game)
  std::unique_ptr<State> prev;  // owned clone
  int root = nilNode;
  int depth = 0, maxDepth = 0;
  int32_t playouts = 0, nc = 0;
  std::map<std::pair<uint32_t, int32_t>, float> cachedPolicies;  // tree.go:47-51,75
  Counters cnt;
  bool no_active_child_panic = false;

  MCTS(State* game, const MCTSConfig& c, Inferencer* inf, uint64_t seed)  // tree.go:80-104
      : conf(c), nn(inf), rnd(seed), current(game) {
    maxDepth = c.M * c.N;
  }

  // ---- tree.go ----
  int alloc() {  // tree.go:145-168
    if (freelist.empty()) {
      Node n;
      n.id = (int)nodes.size();
      n.minPSARatioChildren = 2.0f;
      nodes.push_back(n);
      children.emplace_back();
      return (int)nodes.size() - 1;
    }
    int i = freelist.back();
    freelist.pop_back();
    return i;
  }
  int New(Single move, float score, float value) {  // tree.go:106-117
    int n = alloc();
    Node& N = nodes[n];
    N.move = move; N.visits = 1; N.status = Active; N.score = score; N.value = value;
    return n;
  }
  void freeNode(int n) {  // tree.go:170-180 + node.go:300-309
    children[n].clear();
    freelist.push_back(n);
    Node& N = nodes[n];
    N.move = -1; N.visits = 0; N.status = 0; N.blackScores = 0; N.minPSARatioChildren = 2.0f;
    N.score = 0; N.value = 0; N.virtualLoss = 0;
  }
  void cleanChildren(int rootn) {  // tree.go:198-209
    std::vector<int> kids = children[rootn];
    for (int kid : kids) {
      nodes[kid].status = Invalid;
      freeables.push_back(kid);
      cleanChildren(kid);
    }
    children[rootn].clear();
  }
  void cleanup(int oldRoot, int newRoot) {  // tree.go:183-196
    std::vector<int> kids = children[oldRoot];
    for (int kid : kids) {
      if (kid != newRoot) {
        nodes[kid].status = Invalid;
        freeables.push_back(kid);
        cleanChildren(kid);
      }
    }
    children[oldRoot].resize(1);
    children[oldRoot][0] = newRoot;
  }
  void SetGame(State* g) { current = g; }  // tree.go:120-124
  int Nodes() const { return (int)nodes.size(); }

  std::vector<float> Policies(const State& g) {  // tree.go:128-142
    uint32_t hash = g.Hash();
    float sum = 0;
    int asp = g.ActionSpace() + 1;
    std::vector<float> r(asp);
    for (int i = 0; i < asp; i++) {
      auto it = cachedPolicies.find({hash, (int32_t)i});
      float prob = it == cachedPolicies.end() ? 0.0f : it->second;
      r[i] = prob;
      sum += prob;
    }
    for (int i = 0; i < asp; i++) r[i] /= sum;
    return r;
  }

  void randomizeChildren(int of) {  // tree.go:212-247
    float accum = 0, norm = 0;
    std::vector<float> accumVector;
    std::vector<int>& kids = children[of];
    for (int kid : kids) {
      uint32_t visits = nodes[kid].visits;
      if (norm == 0) {
        norm = (float)visits;
        if (visits <= conf.RandomMinVisits) return;
      }
      if (visits > conf.RandomMinVisits) {
        // math32.Pow(x, y) = float32(math.Pow(float64(x), float64(y)))
        accum += (float)std::pow((double)((float)visits / norm), (double)(1 / conf.RandomTemperature));
        accumVector.push_back(accum);
      }
    }
    float r = rnd.uniform() * accum;
    int index = 0;
    for (size_t i = 0; i < accumVector.size(); i++)
      if (r < accumVector[i]) { index = (int)i; break; }
    if (index == 0) return;
    for (int i = 0; i < (int)kids.size() - index; i++) std::swap(kids[i], kids[i + index]);
  }

  void Reset() {  // tree.go:249-276 — the reference's Reset leaves an unusable tree that
                  // Arena.Play immediately replaces (arena.go:140-141,175-176); we just clear.
    freelist.clear(); freeables.clear(); nodes.clear(); children.clear();
    playouts = 0; cachedPolicies.clear(); root = nilNode; prev.reset();
  }

  // ---- node.go ----
  static bool HasChildren(const Node& n) { return n.minPSARatioChildren <= 1; }           // node.go:129
  static bool IsExpandable(const Node& n, float r) { return r < n.minPSARatioChildren; }  // node.go:132
  static float Evaluate(const Node& n, Player player) {  // node.go:147-159
    float bs = n.blackScores;
    if (player == White) bs += n.virtualLoss;
    float score = bs / (float)n.visits;
    if (player == White) score = 1 - score;
    return score;
  }
  static float NNEvaluate(const Node& n, Player player) {  // node.go:162-167
    if (player == White) return 1.0f - n.value;
    return n.value;
  }
  int Select(int nid, Player of) {  // node.go:170-237
    float sumScore = 0;
    uint32_t parentVisits = 0;
    const std::vector<int>& kids = children[nid];
    for (int kid : kids) {
      const Node& child = nodes[kid];
      if (child.status != Invalid) {
        uint32_t visits = child.visits;
        parentVisits += visits;
        if (visits > 0) sumScore += child.score;
      }
    }
    (void)sumScore;
    int best = nilNode;
    float bestValue = -std::numeric_limits<float>::infinity();
    float fpu = NNEvaluate(nodes[nid], of);
    float numerator = sqrtf((float)parentVisits);
    for (int kid : kids) {
      const Node& child = nodes[kid];
      if (child.status != Active) continue;
      float qsa = fpu;
      uint32_t visits = child.visits;
      if (visits > 0) qsa = Evaluate(child, of);
      float psa = child.score;
      float denominator = 1.0f + (float)visits;
      float lastTerm = numerator / denominator;
      float puct = conf.PUCT * psa * lastTerm;
      float usa = qsa + puct;
      if (usa > bestValue) { bestValue = usa; best = kid; }
    }
    cnt.select_children += kids.size();
    cnt.select_levels++;
    if (best == nilNode) { no_active_child_panic = true; throw std::runtime_error("Cannot return nil"); }
    return best;
  }
  void Update(int nid, float score) {  // node.go:70-76, 263-270
    nodes[nid].visits += 1;
    nodes[nid].blackScores = nodes[nid].blackScores + score;
    cnt.backup_nodes++;
  }
  int countChildren(int nid) {  // node.go:273-285
    int r = 0;
    for (int kid : children[nid]) {
      if (nodes[kid].status == Active) r += countChildren(kid);
      r++;
    }
    return r;
  }
  int findChild(int nid, Single move) {  // node.go:288-298
    for (int kid : children[nid])
      if (nodes[kid].move == move) return kid;
    return nilNode;
  }

  // ---- search.go ----
  static const uint32_t noResultBits = 0x7FE00000u;  // search.go:39-51
  static float noResult() { float f; uint32_t b = noResultBits; memcpy(&f, &b, 4); return f; }
  static bool isNullResult(float r) { uint32_t b; memcpy(&b, &r, 4); return b == noResultBits; }

  float minPsaRatio() const {  // search.go:81-90
    float ratio = (float)nc / (float)MAXTREESIZE;
    if (ratio > 0.95f) return 0.01f;
    if (ratio > 0.5f) return 0.001f;
    return 0;
  }
  static float combinedScore(const State& s) {  // utils.go:62-67
    float whiteScore = s.Score(White);
    float blackScore = s.Score(Black);
    float komi = s.AdditionalScore();
    return blackScore - whiteScore - komi;
  }

  // search.go:92-164, split into its three phases so that tests can stop between them;
  // Search() is exactly Begin + Run(Sims) + End.
  uint32_t searchBoardHash = 0;
  void SearchBegin(Player player) {  // search.go:93-109
    cnt.searches++;
    updateRoot();
    current->SetToMove(player);
    searchBoardHash = current->Hash();
    for (int f : freeables) freeNode(f);
    prepareRoot(player, *current);
    depth = 0;
  }
  // Workers > 1: the reference's concurrent searchStates under ONE fixed interleaving (its goroutine schedule is
  // otherwise unspecified).  A round starts `Workers` pipeline calls one after the other; each descends, setting the
  // virtual-loss flag on its path (search.go:222), until it needs an inference — the slow step during which the
  // next worker runs.  Null results and two-pass terminals complete (and clear their flags) on the spot.  Then the
  // pending workers finish in start order: expansion (children already present are found by findChild /
  // oldMinPsa = 0, search.go:318-325), Update along the path, undoVirtualLoss (a store of 0, node.go:255-260).
  struct Pending { std::vector<int> path; std::unique_ptr<State> st; bool hadChildren; };
  void finishPath(const std::vector<int>& path, float ret) {
    for (size_t i = path.size(); i-- > 0;) {
      if (!isNullResult(ret)) Update(path[i], ret);
      nodes[path[i]].virtualLoss = 0;
    }
  }
  void SearchRunWorkers(int iterations, int workers) {
    int left = iterations;
    while (left > 0) {
      const int v = std::min(workers, left);
      left -= v;
      std::vector<Pending> pend;
      for (int l = 0; l < v; l++) {
        std::unique_ptr<State> cur(current->Clone());
        std::vector<int> path;
        int node = root, d = 0;
        float ret = noResult();
        bool pending = false;
        while (true) {
          d++;
          if (d > maxDepth) break;  // search.go:211-215
          Player player = cur->ToMove();
          nodes[node].virtualLoss = 3.0f;
          path.push_back(node);
          bool isExpandable = IsExpandable(nodes[node], 0);
          if (isExpandable && cur->Passes() >= 2) { ret = combinedScore(*cur); break; }
          if (isExpandable && nc < MAXTREESIZE && IsExpandable(nodes[node], minPsaRatio())) {  // search.go:229, 264, 269
            Pending p; p.path = path; p.hadChildren = HasChildren(nodes[node]); p.st = std::move(cur);
            if (p.hadChildren) throw std::runtime_error("partially expanded node under concurrent search");
            pend.push_back(std::move(p));
            pending = true;
            break;
          }
          if (!HasChildren(nodes[node])) break;
          int next = Select(node, player);
          PlayerMove pm{player, nodes[next].move};
          if (!cur->Check(pm)) break;
          State* n = cur->Apply(pm);
          if (n != cur.get()) cur.reset(n);
          node = next;
        }
        cnt.sims++;
        if (pending) { playouts++; continue; }
        if (!isNullResult(ret)) playouts++; else cnt.null_results++;
        finishPath(path, ret);
      }
      for (Pending& p : pend) {
        float value; bool ok;
        expandBody(p.path.back(), *p.st, minPsaRatio(), &value, &ok);
        finishPath(p.path, ok ? value : noResult());
      }
    }
  }
  void SearchRun(int iterations) {  // canonical doSearch (search.go:166-202): 1 worker, fixed count
    if (conf.Workers > 1) { SearchRunWorkers(iterations, conf.Workers); return; }
    for (int it = 0; it < iterations; it++) {
      std::unique_ptr<State> cl(current->Clone());
      float res = pipeline(cl, root);
      cnt.sims++;
      if (!isNullResult(res)) playouts++;
      else cnt.null_results++;
    }
  }
  Single SearchEnd() {  // search.go:140-163
    if (!HasChildren(nodes[root])) {  // search.go:141-149
      std::vector<float> policy; float v;
      nn->Infer(*current, &policy, &v);
      cnt.evals++;
      int moveID = 0;
      float mx = -std::numeric_limits<float>::infinity();
      for (size_t i = 0; i < policy.size(); i++)
        if (policy[i] > mx) { mx = policy[i]; moveID = (int)i; }
      if (moveID > current->ActionSpace()) return PassMove;
      return (Single)moveID;
    }
    Single retVal = bestMove();
    prev.reset(current->Clone());
    cachedPolicies[{searchBoardHash, (int32_t)retVal}] += 1.0f;
    return retVal;
  }
  Single Search(Player player) {
    SearchBegin(player);
    SearchRun(conf.Sims);
    return SearchEnd();
  }

  // search.go:209-257
  float pipeline(std::unique_ptr<State>& cur, int start) {
    float retVal = noResult();
    depth++;
    if (depth > maxDepth) { depth--; return retVal; }
    Player player = cur->ToMove();
    int32_t nodeCount = nc;
    nodes[start].virtualLoss = 3.0f;  // addVirtualLoss, node.go:248-253 (a store, not an add)

    bool isExpandable = IsExpandable(nodes[start], 0);
    if (isExpandable && cur->Passes() >= 2) {
      retVal = combinedScore(*cur);
    } else if (isExpandable && nodeCount < MAXTREESIZE) {
      bool hadChildren = HasChildren(nodes[start]);
      float value; bool ok;
      expandAndSimulate(start, *cur, minPsaRatio(), &value, &ok);
      if (!hadChildren && ok) retVal = value;
    }
    if (HasChildren(nodes[start]) && isNullResult(retVal)) {
      int next = Select(start, player);
      Single move = nodes[next].move;
      PlayerMove pm{player, move};
      if (cur->Check(pm)) {
        State* n = cur->Apply(pm);
        if (n != cur.get()) cur.reset(n);
        retVal = pipeline(cur, next);
      }
    }
    if (!isNullResult(retVal)) Update(start, retVal);
    nodes[start].virtualLoss = 0;  // undoVirtualLoss
    depth--;
    return retVal;
  }

  struct Pair { Single Coord; float Score; };  // utils.go:49-53

  // search.go:259-339
  void expandAndSimulate(int parent, const State& state, float minPsaRatio_, float* value, bool* ok) {
    *value = 0; *ok = false;
    if (!IsExpandable(nodes[parent], minPsaRatio_)) return;
    if (state.Passes() >= 2) return;
    expandBody(parent, state, minPsaRatio_, value, ok);
  }
  // search.go:274-338: everything after the two early returns (a concurrent worker passed them at descent time)
  void expandBody(int parent, const State& state, float minPsaRatio_, float* value, bool* ok) {
    std::vector<float> policy;
    nn->Infer(state, &policy, value);
    cnt.evals++;
    float passProb = policy.at(policy.size() - 1);
    Player player = state.ToMove();
    if (player == White) *value = 1 - *value;

    std::vector<Pair> nodelist;
    float legalSum = 0;
    int asp = current->ActionSpace();
    for (int i = 0; i < asp; i++) {
      if (state.Check(PlayerMove{player, (Single)i})) {
        nodelist.push_back(Pair{(Single)i, policy.at(i)});
        legalSum += policy[i];
      }
    }
    if (state.Check(PlayerMove{player, PassMove})) {
      nodelist.push_back(Pair{PassMove, passProb});
      legalSum += passProb;
    }
    if (legalSum > std::numeric_limits<float>::denorm_min() /* math32.SmallestNonzeroFloat32 */) {
      for (auto& p : nodelist) p.Score /= legalSum;
    } else {
      float prob = 1 / (float)nodelist.size();
      for (auto& p : nodelist) p.Score = prob;
    }
    if (nodelist.empty()) { *ok = true; return; }
    // sort.Sort(byScore) — pinned to a stable sort
    std::stable_sort(nodelist.begin(), nodelist.end(),
                     [](const Pair& a, const Pair& b) { return a.Score > b.Score; });
    float maxPsa = nodelist[0].Score;
    float oldMinPsa = maxPsa * nodes[parent].minPSARatioChildren;
    float newMinPsa = maxPsa * minPsaRatio_;
    bool skippedChildren = false;
    for (const Pair& p : nodelist) {
      if (p.Score < newMinPsa) {
        skippedChildren = true;
      } else if (p.Score < oldMinPsa) {
        if (findChild(parent, p.Coord) == nilNode) {
          int nn_ = New(p.Coord, p.Score, *value);
          children[parent].push_back(nn_);
          cnt.created++;
        }
      }
    }
    nodes[parent].minPSARatioChildren = skippedChildren ? minPsaRatio_ : 0.0f;
    *ok = true;
  }

  // utils.go:10-47 — fancySort.Less
  bool fancyLess(Player underEval, int a, int b) const {
    const Node& li = nodes[a];
    const Node& lj = nodes[b];
    if (li.visits != lj.visits) return li.visits > lj.visits;
    if (li.visits == 0) return li.score > lj.score;
    return Evaluate(li, underEval) > Evaluate(lj, underEval);
  }

  Single bestMove() {  // search.go:341-390
    Player player = current->ToMove();
    int moveNum = current->MoveNumber();
    std::vector<int>& kids = children[root];
    std::stable_sort(kids.begin(), kids.end(), [&](int a, int b) { return fancyLess(player, a, b); });
    if (moveNum < conf.RandomCount) randomizeChildren(root);
    if (kids.empty()) return PassMove;
    const Node& firstChild = nodes[kids[0]];
    Single best = firstChild.move;
    float bestScore = Evaluate(firstChild, player);
    const Node& rootN = nodes[root];
    if (conf.PassPref == DontPreferPass && best == PassMove) {
      noPassBestMove(&best, &bestScore, root, *current, player);
    } else if (!conf.DumbPass && best == PassMove) {
      float score = rootN.score;
      if ((score > 0 && player == White) || (score < 0 && player == Black))
        noPassBestMove(&best, &bestScore, root, *current, player);
    } else if (!conf.DumbPass && current->LastMove().single == PassMove) {
      float score = rootN.score;
      if ((score > 0 && player == White) || (score < 0 && player == Black)) {
      } else best = PassMove;
    }
    if (best == PassMove && shouldResign(bestScore, player)) best = ResignMove;
    return best;
  }

  void prepareRoot(Player player, const State& state) {  // search.go:392-408
    bool hadChildren = !children[root].empty();
    bool expandable = IsExpandable(nodes[root], 0);
    float value = 0; bool ok;
    if (expandable) expandAndSimulate(root, state, minPsaRatio(), &value, &ok);
    if (hadChildren) value = Evaluate(nodes[root], player);
    else Update(root, value);
  }

  bool newRootState() {  // search.go:424-469
    if (root == nilNode || !prev) return false;
    int d = current->MoveNumber() - prev->MoveNumber();
    if (d < 0) return false;
    if (d > 0 && !current->SupportsUndo()) return false;  // COMPLETION for wq (UndoLastMove panics, wq/game.go:119)
    std::unique_ptr<State> tmp(current->Clone());
    for (int i = 0; i < d; i++) tmp->UndoLastMove();
    if (!tmp->Eq(prev.get())) return false;
    for (int i = 0; i < d; i++) {
      tmp->Fwd();
      PlayerMove move = tmp->LastMove();
      int oldRoot = root;
      int newRoot = findChild(oldRoot, move.single);
      if (newRoot == nilNode) return false;
      root = newRoot;
      cleanup(oldRoot, newRoot);
      State* n = prev->Apply(move);
      if (n != prev.get()) prev.reset(n);
    }
    if (current->MoveNumber() != prev->MoveNumber()) return false;
    if (!current->Eq(prev.get())) return false;
    return true;
  }

  void updateRoot() {  // search.go:473-500
    freeables.clear();
    Player player = current->ToMove();
    if (!newRootState() || root == nilNode) {
      if (current->Check(PlayerMove{player, PassMove})) {
        root = New(PassMove, 0, 0);
      } else {
        int asp = current->ActionSpace();
        for (int i = 0; i < asp; i++)
          if (current->Check(PlayerMove{player, (Single)i})) { root = New((Single)i, 0, 0); break; }
      }
    }
    prev.reset();
    nc = countChildren(root);
    if (children[root].empty()) nodes[root].minPSARatioChildren = 2.0f;
  }

  bool shouldResign(float bestScore, Player) const {  // search.go:502-535
    if (conf.PassPref == DontResign) return false;
    if (conf.ResignPercentage == 0) return false;
    int squares = conf.M * conf.N;
    int threshold = squares / 4;
    if (current->MoveNumber() <= threshold) return false;
    float resignThreshold = conf.ResignPercentage < 0 ? 0.1f : conf.ResignPercentage;
    if (bestScore > resignThreshold) return false;
    return true;
  }
  int noPass(int of, const State& state, Player player) {  // search.go:538-551
    for (int kid : children[of]) {
      Single move = nodes[kid].move;
      bool ok = state.Check(PlayerMove{player, move});
      if (move != PassMove && ok) return kid;
    }
    return nilNode;
  }
  void noPassBestMove(Single* best, float* bestScore, int of, const State& state, Player player) {  // search.go:553-563
    int np = noPass(of, state, player);
    if (np >= 0) {
      *best = nodes[np].move;
      *bestScore = 1;
      if (nodes[np].visits != 0) *bestScore = Evaluate(nodes[np], player);
    }
  }

  // ---- test support: canonical tree dump (DFS preorder, children in list order) ----
  struct DumpRow { int32_t depth, move; uint32_t visits, wbits, pbits; int32_t expanded, nchildren; };
  void dump(int nid, int d, std::vector<DumpRow>* out) const {
    const Node& n = nodes[nid];
    DumpRow r;
    r.depth = d; r.move = n.move; r.visits = n.visits;
    memcpy(&r.wbits, &n.blackScores, 4); memcpy(&r.pbits, &n.score, 4);
    r.expanded = HasChildren(n) ? 1 : 0;
    r.nchildren = (int)children[nid].size();
    out->push_back(r);
    for (int kid : children[nid]) dump(kid, d + 1, out);
  }
};

}  // namespace oracle
