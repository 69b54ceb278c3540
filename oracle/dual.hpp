// ORACLE — TEST INFRASTRUCTURE ONLY (see game.hpp).
// CPU fp32 restatement of the reference's dualnet package: dualnet/config.go, dualnet/dual.go
// (fwd 50-103, bwd 105-132, Model 134-142), dualnet/ermahagerdmonards.go (conv/batchnorm/res/
// share/linear/xent), dualnet/meta.go (Train 16-54, shuffleBatch 57-102, Infer 125-190).
//
// PARITY UNPINNED: the arithmetic of these ops lives in gorgonia.org/gorgonia
// v0.9.17-0.20210124090702-531c6df2c434 and gorgonia.org/tensor v0.9.18 (go.mod:5-16), whose
// sources are not in /root/reference, and no reference test asserts a numeric output
// (dualnet/dual_test.go only logs).  This file restates the published behaviour of those ops
// as recalled in SURVEY.md §8a "Canonical-dualnet uncertainties"; the choices are named:
//   (i)  BatchNorm(x, nil, nil, ...) creates learnable scale/bias with the FULL shape of x
//        ([B,C,H,W], GlorotN(1.0)); output = scale ⊙ xhat + bias.
//   (ii) train mode: per-channel batch mean / biased variance over (B,H,W), eps inside sqrt;
//        test mode after BatchNormOp.Reset() (meta.go:170-172): stored mean/var/ma are zero, so
//        xhat = x / sqrt(eps).
//   (iii) SoftMax = exp(x)/sum(exp(x)) without max subtraction.
//   (iv) GEMM/conv summation order: plain ascending (ci, ky, kx) fp32 accumulation.
// Inference reads batch row 0 of every batch-shaped parameter (meta.go:141-146 prefix copy).
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "rng.hpp"

namespace oracle {

struct DualConfig {  // dualnet/config.go:4-16
  int K = 0, SharedLayers = 0, FC = 0;
  double L2 = 0;
  int BatchSize = 0, Width = 0, Height = 0, Features = 0, ActionSpace = 0;
  bool FwdOnly = false;
  bool IsValid() const {  // config.go:33-42
    return K >= 1 && ActionSpace >= 3 && SharedLayers >= 0 && FC > 1 && BatchSize >= 1 && Features > 0;
  }
};

inline int dual_round(int a) {  // config.go:44-59
  int n = a - 1;
  n |= n >> 1; n |= n >> 2; n |= n >> 4; n |= n >> 8; n |= n >> 16;
  n++;
  int lt = n / 2;
  if ((a - lt) < (n - a)) return lt;
  return n;
}

inline DualConfig DefaultConf(int m, int n, int actionSpace) {  // config.go:18-31
  DualConfig c;
  int k = dual_round((m * n) / 3);
  c.K = k; c.SharedLayers = m; c.FC = 2 * k;
  c.BatchSize = 256; c.Width = n; c.Height = m; c.Features = 18; c.ActionSpace = actionSpace;
  return c;
}

struct ParamDesc {
  std::string name;
  int rank;
  int shape[4];
  size_t offset, size;
  int init;  // 0 zeros, 1 GlorotU(1.0), 2 GlorotN(1.0)
};

// One conv+BN(+ReLU) unit: filter [Co,Ci,k,k], gamma/beta [B,Co,H,W]
struct ConvBN { int filter, gamma, beta, Ci, Co, k; };

struct Dual {
  DualConfig conf;
  std::vector<ParamDesc> desc;  // Model() order (dual.go:134-142 = graph creation order)
  std::vector<float> params;
  std::vector<ConvBN> units;    // Init, (Layer1,Layer2)*SharedLayers, PolicyHead, ValueHead
  int pW, pB, vW, vB, voW, voB;

  explicit Dual(const DualConfig& c) : conf(c) { build(); }

  int add(const std::string& name, std::vector<int> shape, int init) {
    ParamDesc d;
    d.name = name; d.rank = (int)shape.size(); d.init = init;
    size_t sz = 1;
    for (int i = 0; i < 4; i++) { d.shape[i] = i < d.rank ? shape[i] : 1; if (i < d.rank) sz *= shape[i]; }
    d.offset = params.size(); d.size = sz;
    params.resize(params.size() + sz, 0.0f);
    desc.push_back(d);
    return (int)desc.size() - 1;
  }
  int addUnit(const std::string& name, int Ci, int Co, int k) {  // ermahagerdmonards.go:33-73
    ConvBN u;
    u.Ci = Ci; u.Co = Co; u.k = k;
    u.filter = add("Filter" + name, {Co, Ci, k, k}, 1);
    u.gamma = add("Filter" + name + "_conv_γ", {conf.BatchSize, Co, conf.Height, conf.Width}, 2);
    u.beta = add("Filter" + name + "_conv_β", {conf.BatchSize, Co, conf.Height, conf.Width}, 2);
    units.push_back(u);
    return (int)units.size() - 1;
  }
  void build() {  // dual.go:50-103
    int B = conf.BatchSize, HW = conf.Width * conf.Height, K = conf.K;
    addUnit("Init", conf.Features, K, 3);
    for (int i = 0; i < conf.SharedLayers; i++) {
      char buf[64];
      snprintf(buf, sizeof buf, "Layer1 of Shared Layer %d", i); addUnit(buf, K, K, 3);
      snprintf(buf, sizeof buf, "Layer2 of Shared Layer %d", i); addUnit(buf, K, K, 3);
    }
    addUnit("PolicyHead", K, 2, 1);
    pW = add("Policy_w", {2 * HW, conf.ActionSpace}, 2);
    pB = add("Policy_b", {B, conf.ActionSpace}, 0);
    addUnit("ValueHead", K, 1, 1);
    vW = add("Value_w", {HW, conf.FC}, 2);
    vB = add("Value_b", {B, conf.FC}, 0);
    voW = add("ValueOutput_w", {conf.FC, 1}, 2);
    voB = add("ValueOutput_b", {B, 1}, 0);
  }
  float* P(int i) { return params.data() + desc[i].offset; }
  const float* P(int i) const { return params.data() + desc[i].offset; }

  // Init: GlorotU(1.0) filters, GlorotN(1.0) BN affine + linear weights, zero biases
  // (ermahagerdmonards.go:39,80,82).  fan = (s0+s1)*prod(s[2:]); drawn from our injected RNG,
  // one independent stream per tensor (reference: Go's global math/rand, unreproducible).
  void Init(uint64_t seed) {
    for (size_t t = 0; t < desc.size(); t++) {
      ParamDesc& d = desc[t];
      float* p = params.data() + d.offset;
      if (d.init == 0) { for (size_t i = 0; i < d.size; i++) p[i] = 0; continue; }
      double field = 1;
      for (int i = 2; i < d.rank; i++) field *= d.shape[i];
      double fan = (double)(d.shape[0] + d.shape[1]) * field;
      double stdev = 1.0 * std::sqrt(2.0 / fan);
      Rng r(derive_seed(seed, t));
      if (d.init == 1) {
        float lim = (float)(stdev * std::sqrt(3.0));
        for (size_t i = 0; i < d.size; i++) p[i] = (2.0f * r.uniform() - 1.0f) * lim;
      } else {
        float sd = (float)stdev;
        for (size_t i = 0; i < d.size; i++) p[i] = r.normal() * sd;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// fp32 kernels (NCHW).  conv: cross-correlation, "same" padding (k-1)/2, stride 1, no bias.
// Per output point the taps are accumulated in ascending (ci, ky, kx) order with an unfused multiply
// and add.  The input is zero-padded once so that every tap is one long contiguous
// `acc[i] += w * src[i]` sweep (vectorisable); a tap that falls outside the board then adds w*0 = +-0,
// which leaves the running fp32 sum bit-identical to skipping it.
inline void conv_fwd(const float* x, const float* w, float* y, int B, int Ci, int Co, int H, int W, int k) {
  const int pad = (k - 1) / 2, HW = H * W, Hp = H + 2 * pad, Wp = W + 2 * pad, PP = Hp * Wp;
  const int span = (H - 1) * Wp + W;  // flat sweep over rows of pitch Wp (pad columns hold garbage, never stored)
  const bool par = (double)B * Co * Ci * HW * k * k > 5e7;  // threads only pay off on big layers (128-way fork/join costs ~ms)
  std::vector<float> xp((size_t)B * Ci * PP, 0.0f);
  for (int b = 0; b < B; b++)
    for (int ci = 0; ci < Ci; ci++) {
      const float* xi = x + ((size_t)b * Ci + ci) * HW;
      float* xo = xp.data() + ((size_t)b * Ci + ci) * PP + pad * Wp + pad;
      for (int yy = 0; yy < H; yy++)
        for (int xx = 0; xx < W; xx++) xo[yy * Wp + xx] = xi[yy * W + xx];
    }
#pragma omp parallel for collapse(2) schedule(static) if (par)
  for (int b = 0; b < B; b++)
    for (int co = 0; co < Co; co++) {
      std::vector<float> accv((size_t)H * Wp, 0.0f);
      float* acc = accv.data();
      for (int ci = 0; ci < Ci; ci++) {
        const float* xi = xp.data() + ((size_t)b * Ci + ci) * PP;
        const float* wk = w + ((size_t)co * Ci + ci) * k * k;
        for (int ky = 0; ky < k; ky++)
          for (int kx = 0; kx < k; kx++) {
            const float wv = wk[ky * k + kx];
            const float* src = xi + ky * Wp + kx;
            for (int i = 0; i < span; i++) acc[i] = acc[i] + wv * src[i];
          }
      }
      float* yo = y + ((size_t)b * Co + co) * HW;
      for (int yy = 0; yy < H; yy++)
        for (int xx = 0; xx < W; xx++) yo[yy * W + xx] = acc[yy * Wp + xx];
    }
}
// dX += conv_transpose(dY, w) ; dW += corr(x, dY)
inline void conv_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int Ci, int Co,
                     int H, int W, int k) {
  int pad = (k - 1) / 2, HW = H * W;
  const bool par = (double)B * Co * Ci * HW * k * k > 5e7;
  if (dx) {
#pragma omp parallel for collapse(2) schedule(static) if (par)
    for (int b = 0; b < B; b++)
      for (int ci = 0; ci < Ci; ci++) {
        float* dxi = dx + ((size_t)b * Ci + ci) * HW;
        for (int co = 0; co < Co; co++) {
          const float* dyo = dy + ((size_t)b * Co + co) * HW;
          const float* wk = w + ((size_t)co * Ci + ci) * k * k;
          for (int ky = 0; ky < k; ky++)
            for (int kx = 0; kx < k; kx++) {
              float wv = wk[ky * k + kx];
              int ddy = ky - pad, ddx = kx - pad;
              int y0 = ddy < 0 ? -ddy : 0, y1 = ddy > 0 ? H - ddy : H;
              int x0 = ddx < 0 ? -ddx : 0, x1 = ddx > 0 ? W - ddx : W;
              for (int yy = y0; yy < y1; yy++)
                for (int xx = x0; xx < x1; xx++) dxi[(yy + ddy) * W + xx + ddx] += wv * dyo[yy * W + xx];
            }
        }
      }
  }
#pragma omp parallel for collapse(2) schedule(static) if (par)
  for (int co = 0; co < Co; co++)
    for (int ci = 0; ci < Ci; ci++) {
      float* dwk = dw + ((size_t)co * Ci + ci) * k * k;
      for (int ky = 0; ky < k; ky++)
        for (int kx = 0; kx < k; kx++) {
          int ddy = ky - pad, ddx = kx - pad;
          int y0 = ddy < 0 ? -ddy : 0, y1 = ddy > 0 ? H - ddy : H;
          int x0 = ddx < 0 ? -ddx : 0, x1 = ddx > 0 ? W - ddx : W;
          float acc = 0;
          for (int b = 0; b < B; b++) {
            const float* xi = x + ((size_t)b * Ci + ci) * HW;
            const float* dyo = dy + ((size_t)b * Co + co) * HW;
            for (int yy = y0; yy < y1; yy++)
              for (int xx = x0; xx < x1; xx++) acc += xi[(yy + ddy) * W + xx + ddx] * dyo[yy * W + xx];
          }
          dwk[ky * k + kx] += acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Inference forward (meta.go:168-190 semantics, useful-work form: n independent samples, each
// computed exactly as row 0 of the reference's ActionSpace-sized batch).
//   planes [n, F, H, W] -> policy [n, A'] (softmax), value [n] (tanh)
struct InferScratch { std::vector<float> a, b, c, d; };

inline void unit_infer(const Dual& net, const ConvBN& u, const float* x, float* z, float* out, int n, bool relu) {
  int H = net.conf.Height, W = net.conf.Width, HW = H * W;
  conv_fwd(x, net.P(u.filter), z, n, u.Ci, u.Co, H, W, u.k);
  const float s = sqrtf(0.0f + (float)1e-5);  // sqrt(var + eps) with var = 0 (BN test mode after Reset)
  const float* g = net.P(u.gamma);  // row 0 of [B,Co,H,W]
  const float* be = net.P(u.beta);
  size_t chw = (size_t)u.Co * HW;
#pragma omp parallel for schedule(static) if ((double)n * chw > 2e7)
  for (int b = 0; b < n; b++)
    for (size_t i = 0; i < chw; i++) {
      float t = z[b * chw + i] / s;
      float y = g[i] * t;
      y = y + be[i];
      out[b * chw + i] = relu ? (y > 0 ? y : 0.0f) : y;
    }
}

inline void dual_infer(const Dual& net, const float* planes, int n, float* policy, float* value,
                       std::vector<float>* tower_out = nullptr) {
  const DualConfig& c = net.conf;
  int H = c.Height, W = c.Width, HW = H * W, K = c.K, A = c.ActionSpace;
  std::vector<float> z((size_t)n * K * HW), cur((size_t)n * K * HW), l1((size_t)n * K * HW), l2((size_t)n * K * HW);
  unit_infer(net, net.units[0], planes, z.data(), cur.data(), n, true);  // dual.go:59
  for (int i = 0; i < c.SharedLayers; i++) {                             // dual.go:62-65, share: 67-73
    unit_infer(net, net.units[1 + 2 * i], cur.data(), z.data(), l1.data(), n, true);
    unit_infer(net, net.units[2 + 2 * i], cur.data(), z.data(), l2.data(), n, true);
    for (size_t j = 0; j < cur.size(); j++) { float s = l1[j] + l2[j]; cur[j] = s > 0 ? s : 0.0f; }
  }
  if (tower_out) *tower_out = cur;
  const ConvBN& pu = net.units[1 + 2 * c.SharedLayers];
  const ConvBN& vu = net.units[2 + 2 * c.SharedLayers];
  std::vector<float> pz((size_t)n * 2 * HW), ph((size_t)n * 2 * HW), vz((size_t)n * HW), vh((size_t)n * HW);
  unit_infer(net, pu, cur.data(), pz.data(), ph.data(), n, true);  // dual.go:70-71
  unit_infer(net, vu, cur.data(), vz.data(), vh.data(), n, true);  // dual.go:85-86
  const float* Wp = net.P(net.pW); const float* bp = net.P(net.pB);
  const float* Wv = net.P(net.vW); const float* bv = net.P(net.vB);
  const float* Wo = net.P(net.voW); const float* bo = net.P(net.voB);
#pragma omp parallel for schedule(static) if ((double)n * HW * A > 2e7)
  for (int b = 0; b < n; b++) {
    std::vector<float> logits(A);
    for (int a = 0; a < A; a++) {  // linear: xw + b (ermahagerdmonards.go:75-84), bias row 0
      float acc = 0;
      for (int j = 0; j < 2 * HW; j++) acc += ph[(size_t)b * 2 * HW + j] * Wp[(size_t)j * A + a];
      logits[a] = acc + bp[a];
    }
    float sum = 0;
    for (int a = 0; a < A; a++) { logits[a] = expf(logits[a]); sum += logits[a]; }  // SoftMax, dual.go:81
    for (int a = 0; a < A; a++) policy[(size_t)b * A + a] = logits[a] / sum;
    std::vector<float> h(c.FC);
    for (int f = 0; f < c.FC; f++) {
      float acc = 0;
      for (int j = 0; j < HW; j++) acc += vh[(size_t)b * HW + j] * Wv[(size_t)j * c.FC + f];
      acc = acc + bv[f];
      h[f] = acc > 0 ? acc : 0.0f;  // dual.go:89-90
    }
    float acc = 0;
    for (int f = 0; f < c.FC; f++) acc += h[f] * Wo[f];
    acc = acc + bo[0];
    value[b] = tanhf(acc);  // dual.go:96
  }
}

// ---------------------------------------------------------------------------------------------
// Training step (meta.go:22-46): forward in BN train mode on one batch of exactly BatchSize
// samples, loss (dual.go:105-126), reverse-mode gradients for every Model() tensor, vanilla SGD
// w -= lr*g (meta.go:20; no momentum, no batch-size scaling).  Returns the cost.
struct TrainCache { std::vector<float> x, z, xn, y; std::vector<float> mean, var; };

inline void unit_train_fwd(const Dual& net, const ConvBN& u, const std::vector<float>& x, TrainCache* tc) {
  const DualConfig& c = net.conf;
  int B = c.BatchSize, H = c.Height, W = c.Width, HW = H * W, Co = u.Co;
  tc->x = x;
  tc->z.assign((size_t)B * Co * HW, 0);
  conv_fwd(x.data(), net.P(u.filter), tc->z.data(), B, u.Ci, Co, H, W, u.k);
  tc->mean.assign(Co, 0); tc->var.assign(Co, 0);
  tc->xn.resize(tc->z.size()); tc->y.resize(tc->z.size());
  float m = (float)((size_t)B * HW);
  const float* g = net.P(u.gamma); const float* be = net.P(u.beta);
  for (int co = 0; co < Co; co++) {
    float s = 0;
    for (int b = 0; b < B; b++) for (int i = 0; i < HW; i++) s += tc->z[((size_t)b * Co + co) * HW + i];
    float mean = s / m;
    float v = 0;
    for (int b = 0; b < B; b++) for (int i = 0; i < HW; i++) { float d = tc->z[((size_t)b * Co + co) * HW + i] - mean; v += d * d; }
    float var = v / m;
    tc->mean[co] = mean; tc->var[co] = var;
    float sd = sqrtf(var + (float)1e-5);
    for (int b = 0; b < B; b++) for (int i = 0; i < HW; i++) {
      size_t idx = ((size_t)b * Co + co) * HW + i;
      float xn = (tc->z[idx] - mean) / sd;
      tc->xn[idx] = xn;
      float y = g[idx] * xn + be[idx];
      tc->y[idx] = y > 0 ? y : 0.0f;
    }
  }
}
// dy: grad wrt the unit's (post-ReLU) output; accumulates dfilter/dgamma/dbeta into grads,
// adds the grad wrt the unit input into dx (if non-null).
inline void unit_train_bwd(const Dual& net, const ConvBN& u, const TrainCache& tc, const std::vector<float>& dy,
                           std::vector<float>* grads, std::vector<float>* dx) {
  const DualConfig& c = net.conf;
  int B = c.BatchSize, H = c.Height, W = c.Width, HW = H * W, Co = u.Co;
  float m = (float)((size_t)B * HW);
  const float* g = net.P(u.gamma);
  float* dg = grads->data() + net.desc[u.gamma].offset;
  float* db = grads->data() + net.desc[u.beta].offset;
  std::vector<float> dxn(tc.z.size()), dz(tc.z.size());
  for (size_t i = 0; i < dy.size(); i++) {
    float d = tc.y[i] > 0 ? dy[i] : 0.0f;  // ReLU
    dg[i] += d * tc.xn[i];
    db[i] += d;
    dxn[i] = d * g[i];
  }
  for (int co = 0; co < Co; co++) {
    float s1 = 0, s2 = 0;
    for (int b = 0; b < B; b++) for (int i = 0; i < HW; i++) {
      size_t idx = ((size_t)b * Co + co) * HW + i;
      s1 += dxn[idx]; s2 += dxn[idx] * tc.xn[idx];
    }
    float m1 = s1 / m, m2 = s2 / m;
    float sd = sqrtf(tc.var[co] + (float)1e-5);
    for (int b = 0; b < B; b++) for (int i = 0; i < HW; i++) {
      size_t idx = ((size_t)b * Co + co) * HW + i;
      dz[idx] = (dxn[idx] - m1 - tc.xn[idx] * m2) / sd;
    }
  }
  conv_bwd(tc.x.data(), net.P(u.filter), dz.data(), dx ? dx->data() : nullptr,
           grads->data() + net.desc[u.filter].offset, B, u.Ci, Co, H, W, u.k);
}

inline float dual_train_step(Dual& net, const float* X, const float* Pi, const float* V, float lr,
                             std::vector<float>* grads_out = nullptr) {
  const DualConfig& c = net.conf;
  int B = c.BatchSize, H = c.Height, W = c.Width, HW = H * W, K = c.K, A = c.ActionSpace, FC = c.FC;
  int nu = (int)net.units.size();
  std::vector<TrainCache> tc(nu);
  std::vector<std::vector<float>> blockIn(c.SharedLayers + 1);
  std::vector<float> x0(X, X + (size_t)B * c.Features * HW);
  unit_train_fwd(net, net.units[0], x0, &tc[0]);
  std::vector<float> cur = tc[0].y;
  for (int i = 0; i < c.SharedLayers; i++) {
    unit_train_fwd(net, net.units[1 + 2 * i], cur, &tc[1 + 2 * i]);
    unit_train_fwd(net, net.units[2 + 2 * i], cur, &tc[2 + 2 * i]);
    std::vector<float> nx(cur.size());
    for (size_t j = 0; j < nx.size(); j++) { float s = tc[1 + 2 * i].y[j] + tc[2 + 2 * i].y[j]; nx[j] = s > 0 ? s : 0.0f; }
    cur.swap(nx);
    blockIn[i + 1] = cur;
  }
  int pu = 1 + 2 * c.SharedLayers, vu = pu + 1;
  unit_train_fwd(net, net.units[pu], cur, &tc[pu]);
  unit_train_fwd(net, net.units[vu], cur, &tc[vu]);
  const float* Wp = net.P(net.pW); const float* bp = net.P(net.pB);
  const float* Wv = net.P(net.vW); const float* bv = net.P(net.vB);
  const float* Wo = net.P(net.voW); const float* bo = net.P(net.voB);
  std::vector<float> logits((size_t)B * A), h1((size_t)B * FC), vraw(B);
  const std::vector<float>& ph = tc[pu].y;  // [B, 2*HW]
  const std::vector<float>& vh = tc[vu].y;  // [B, HW]
  for (int b = 0; b < B; b++) {
    for (int a = 0; a < A; a++) {
      float acc = 0;
      for (int j = 0; j < 2 * HW; j++) acc += ph[(size_t)b * 2 * HW + j] * Wp[(size_t)j * A + a];
      logits[(size_t)b * A + a] = acc + bp[(size_t)b * A + a];
    }
    for (int f = 0; f < FC; f++) {
      float acc = 0;
      for (int j = 0; j < HW; j++) acc += vh[(size_t)b * HW + j] * Wv[(size_t)j * FC + f];
      acc = acc + bv[(size_t)b * FC + f];
      h1[(size_t)b * FC + f] = acc > 0 ? acc : 0.0f;
    }
    float acc = 0;
    for (int f = 0; f < FC; f++) acc += h1[(size_t)b * FC + f] * Wo[f];
    vraw[b] = acc + bo[b];
  }
  // cost (dual.go:105-126; xent on raw logits, ermahagerdmonards.go:106-147)
  float psum = 0, vsum = 0;
  for (size_t i = 0; i < logits.size(); i++) psum += -(Pi[i] * logits[i] + (1 - Pi[i]) * (1 - logits[i]));
  float pcost = psum / (float)logits.size();
  for (int b = 0; b < B; b++) { float d = vraw[b] - V[b]; vsum += d * d; }
  float vcost = vsum / (float)B;
  float cost = pcost + vcost;

  std::vector<float> grads(net.params.size(), 0.0f);
  auto G = [&](int i) { return grads.data() + net.desc[i].offset; };
  // d cost / d logits = (1 - 2*Pi) / (B*A') ; d cost / d vraw = 2 (vraw - V) / B
  std::vector<float> dlog(logits.size()), dph(ph.size(), 0.0f), dvh(vh.size(), 0.0f);
  float invBA = 1.0f / (float)logits.size();
  for (size_t i = 0; i < dlog.size(); i++) dlog[i] = (1 - 2 * Pi[i]) * invBA;
  for (int b = 0; b < B; b++)
    for (int a = 0; a < A; a++) {
      float d = dlog[(size_t)b * A + a];
      G(net.pB)[(size_t)b * A + a] += d;
      for (int j = 0; j < 2 * HW; j++) {
        G(net.pW)[(size_t)j * A + a] += ph[(size_t)b * 2 * HW + j] * d;
        dph[(size_t)b * 2 * HW + j] += Wp[(size_t)j * A + a] * d;
      }
    }
  for (int b = 0; b < B; b++) {
    float dv = 2 * (vraw[b] - V[b]) / (float)B;
    G(net.voB)[b] += dv;
    for (int f = 0; f < FC; f++) {
      float hv = h1[(size_t)b * FC + f];
      G(net.voW)[f] += hv * dv;
      float dh = hv > 0 ? Wo[f] * dv : 0.0f;
      G(net.vB)[(size_t)b * FC + f] += dh;
      for (int j = 0; j < HW; j++) {
        G(net.vW)[(size_t)j * FC + f] += vh[(size_t)b * HW + j] * dh;
        dvh[(size_t)b * HW + j] += Wv[(size_t)j * FC + f] * dh;
      }
    }
  }
  std::vector<float> dcur((size_t)B * K * HW, 0.0f);
  unit_train_bwd(net, net.units[pu], tc[pu], dph, &grads, &dcur);
  unit_train_bwd(net, net.units[vu], tc[vu], dvh, &grads, &dcur);
  for (int i = c.SharedLayers - 1; i >= 0; i--) {
    // out = relu(l1 + l2): l1,l2 >= 0 so the outer ReLU gates on (l1+l2) > 0
    std::vector<float> dl(dcur.size());
    const std::vector<float>& y1 = tc[1 + 2 * i].y; const std::vector<float>& y2 = tc[2 + 2 * i].y;
    for (size_t j = 0; j < dl.size(); j++) dl[j] = (y1[j] + y2[j]) > 0 ? dcur[j] : 0.0f;
    std::vector<float> dprev(dcur.size(), 0.0f);
    unit_train_bwd(net, net.units[1 + 2 * i], tc[1 + 2 * i], dl, &grads, &dprev);
    unit_train_bwd(net, net.units[2 + 2 * i], tc[2 + 2 * i], dl, &grads, &dprev);
    dcur.swap(dprev);
  }
  unit_train_bwd(net, net.units[0], tc[0], dcur, &grads, nullptr);
  if (grads_out) *grads_out = grads;
  if (lr != 0)
    for (size_t i = 0; i < net.params.size(); i++) net.params[i] = net.params[i] - lr * grads[i];  // VanillaSolver
  return cost;
}

// meta.go:57-102 shuffleBatch: Fisher-Yates over rows, j = r.Intn(i+1), injected RNG
inline void shuffle_batch(std::vector<float>& Xs, std::vector<float>& Pi, std::vector<float>& V, int rows, Rng* r) {
  size_t xr = Xs.size() / rows, pr = Pi.size() / rows;
  std::vector<float> tmp(xr > pr ? xr : pr);
  for (int i = 0; i < rows; i++) {
    int j = r->intn(i + 1);
    for (size_t t = 0; t < xr; t++) std::swap(Xs[i * xr + t], Xs[j * xr + t]);
    for (size_t t = 0; t < pr; t++) std::swap(Pi[i * pr + t], Pi[j * pr + t]);
    std::swap(V[i], V[j]);
  }
}

// meta.go:16-54
inline void dual_train(Dual& net, std::vector<float>& Xs, std::vector<float>& Pi, std::vector<float>& V, int batches,
                       int iterations, float lr, Rng* r, std::vector<float>* costs = nullptr) {
  const DualConfig& c = net.conf;
  size_t xr = (size_t)c.Features * c.Height * c.Width, pr = c.ActionSpace;
  int rows = batches * c.BatchSize;
  for (int it = 0; it < iterations; it++) {
    for (int bat = 0; bat < batches; bat++) {
      size_t s = (size_t)bat * c.BatchSize;
      float cost = dual_train_step(net, Xs.data() + s * xr, Pi.data() + s * pr, V.data() + s, lr);
      if (costs) costs->push_back(cost);
    }
    shuffle_batch(Xs, Pi, V, rows, r);
  }
}

}  // namespace oracle
