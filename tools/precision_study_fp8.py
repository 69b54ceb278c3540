#!/usr/bin/env python
"""Feasibility arithmetic for DESIGN.md §10: can the two correction passes of the fp16 hi/lo split (hi*lo + lo*hi) run on
the FP8 tensor path (kind::f8f6f4, twice the fp16 MMA rate) without leaving fp32-level accuracy?  One C = 256 layer, 19x19,
numpy emulation: products of exactly representable operands are exact in fp32, accumulation in fp32 (what the tensor core
does), E4M3 / E5M2 rounding emulated.  Compared against an fp64 convolution."""
import numpy as np

rng = np.random.default_rng(0)
C, H, W, NB = 256, 19, 19, 2


def to_fp8(x, mant, emin, emax):
    """Round-to-nearest-even to a (1, e, mant) float with exponent range [emin, emax] and subnormals; saturating."""
    x = x.astype(np.float64)
    sign, a = np.sign(x), np.abs(x)
    e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.clip(e, emin, emax)
    q = 2.0 ** (e - mant)
    r = np.round(a / q) * q
    r = np.minimum(r, (2 - 2.0 ** -mant) * 2.0 ** emax)
    return (sign * r).astype(np.float32)


E4M3 = dict(mant=3, emin=-6, emax=8)
E5M2 = dict(mant=2, emin=-14, emax=15)


def split16(x, target=13):
    m = np.abs(x).max()
    e = target - int(np.floor(np.log2(m)))
    xs = (x * np.float32(2.0 ** e)).astype(np.float32)
    hi = xs.astype(np.float16).astype(np.float32)
    lo = (xs - hi).astype(np.float32)
    return hi, lo, e


def scale8(x, fmt):
    """per-tensor power-of-two scale that puts absmax just under the format's maximum, then round"""
    m = np.abs(x).max()
    if m == 0:
        return x, 0
    e = (fmt["emax"] - 1) - int(np.floor(np.log2(m)))
    return to_fp8(x * np.float32(2.0 ** e), **fmt), e


x = np.maximum(rng.normal(0, 1, (NB, C, H, W)), 0).astype(np.float32)
lim = np.sqrt(6.0 / ((C + C) * 9))
w = rng.uniform(-lim, lim, (C, C, 3, 3)).astype(np.float32)
xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
ref = np.zeros((NB, C, H, W))
for ky in range(3):
    for kx in range(3):
        ref += np.einsum("oc,nchw->nohw", w[:, :, ky, kx].astype(np.float64), xp[:, :, ky:ky + H, kx:kx + W])
xp32 = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
cols = np.stack([xp32[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], axis=1)
A = cols.transpose(0, 3, 4, 1, 2).reshape(NB * H * W, 9 * C)
B = w.transpose(2, 3, 1, 0).reshape(9 * C, C)
ref2 = ref.transpose(0, 2, 3, 1).reshape(NB * H * W, C)
rms_out = np.sqrt((ref2 ** 2).mean())

ah, al, ea = split16(A)
bh, bl, eb = split16(B)
s16 = np.float32(2.0 ** -(ea + eb))


def report(name, y, cost):
    err = y.astype(np.float64) - ref2
    print("%-64s rms err %.2e  max %.2e  (rel. to rms(out) %.1e)   tensor cost %.2f" %
          (name, np.sqrt((err ** 2).mean()), np.abs(err).max(), np.sqrt((err ** 2).mean()) / rms_out, cost))


report("fp16 hi*hi + hi*lo + lo*hi  (today)", (ah @ bh + ah @ bl + al @ bh) * s16, 3.0)
report("fp16 hi*hi + hi*lo          (2 passes)", (ah @ bh + ah @ bl) * s16, 2.0)
report("fp16 hi*hi only", (ah @ bh) * s16, 1.0)
for fname, fmt in (("E4M3", E4M3), ("E5M2", E5M2)):
    # correction terms on the FP8 path: (hi8 of A) * (lo8 of B) + (lo8 of A) * (hi8 of B), each operand with its own scale
    ah8, eah = scale8(ah, fmt); al8, eal = scale8(al, fmt)
    bh8, ebh = scale8(bh, fmt); bl8, ebl = scale8(bl, fmt)
    corr = (ah8 @ bl8) * np.float32(2.0 ** -(eah + ebl)) + (al8 @ bh8) * np.float32(2.0 ** -(eal + ebh))
    report("fp16 hi*hi + %s (hi*lo + lo*hi)" % fname, (ah @ bh + corr) * s16, 2.0)
    # three-way split hi16 + mid8 + lo8: corrections from mid8 only (1 fp16 + 2 fp8 passes) and with lo8 too (4 fp8 passes)
    am8, eam = scale8(al, fmt); ar = al - am8 * np.float32(2.0 ** -eam); ar8, ear = scale8(ar, fmt)
    bm8, ebm = scale8(bl, fmt); br = bl - bm8 * np.float32(2.0 ** -ebm); br8, ebr = scale8(br, fmt)
    c2 = (ah8 @ bm8) * np.float32(2.0 ** -(eah + ebm)) + (am8 @ bh8) * np.float32(2.0 ** -(eam + ebh))
    # error of using ah8 instead of ah in the correction: add (ah - ah8) * bm8?  -> needs another pass; report as is
    report("fp16 hi*hi + %s (hi8*mid8 + mid8*hi8)" % fname, (ah @ bh + c2) * s16, 2.0)
report("fp32 GEMM (numpy)", A @ B, 0)


def mxfp4(x, axis):
    """E2M1 with one power-of-two (UE8M0) scale per 32 consecutive elements along `axis` (kind::mxf4, 4x the fp16 rate)."""
    x = np.moveaxis(x.astype(np.float64), axis, -1)
    shp = x.shape
    xb = x.reshape(shp[:-1] + (shp[-1] // 32, 32))
    m = np.abs(xb).max(axis=-1, keepdims=True)
    e = np.where(m > 0, np.floor(np.log2(np.where(m > 0, m, 1.0))) - 2, 0)  # block max lands in [4, 8): E2M1 max is 6
    y = xb / 2.0 ** e
    grid = np.array([0, 0.5, 1, 1.5, 2, 3, 4, 6])
    idx = np.abs(np.abs(y)[..., None] - grid).argmin(axis=-1)
    q = np.sign(y) * grid[idx] * 2.0 ** e
    return np.moveaxis(q.reshape(shp), -1, axis).astype(np.float32)


corr4 = mxfp4(ah, 1) @ mxfp4(bl, 0) + mxfp4(al, 1) @ mxfp4(bh, 0)
report("fp16 hi*hi + MXFP4 (hi*lo + lo*hi), block scale per 32", (ah @ bh + corr4) * s16, 1.5)

# ---- robustness: activations spanning many binades (ReLU of a heavy-tailed pre-activation) with FIXED fp8 scales chosen
# from the fp16 scale plus headroom (the engine cannot take an absmax of a layer's output before writing it)
print()
for tail, head in ((1.0, 3), (2.0, 3), (2.0, 6)):
    xh = (np.maximum(rng.normal(0, 1, (NB, C, H, W)), 0) * np.exp(tail * rng.normal(0, 1, (NB, C, H, W)))).astype(np.float32)
    xph = np.pad(xh, ((0, 0), (0, 0), (1, 1), (1, 1)))
    colsh = np.stack([xph[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], axis=1)
    Ah = colsh.transpose(0, 3, 4, 1, 2).reshape(NB * H * W, 9 * C)
    refh = Ah.astype(np.float64) @ B.astype(np.float64)
    rms_h = np.sqrt((refh ** 2).mean())
    # fp16 split at a fixed activation scale with `head` binades of headroom over this batch's absmax
    ah, al, ea = split16(Ah, target=13 - head)
    s16h = np.float32(2.0 ** -(ea + eb))
    p_hi = (E4M3["emax"] - 1 - head) - (13 - head)          # hi8 = E4M3(hi * 2^p_hi): same headroom in the fp8 format
    ah8 = to_fp8(ah * np.float32(2.0 ** p_hi), **E4M3)
    bl8 = to_fp8(bl * np.float32(2.0 ** -p_hi), **E4M3)      # the product keeps the accumulator's scale
    q = 5 + head                                            # |lo| <= 2^-11 * 2^(13-head): its bound maps to 2^7
    al8 = to_fp8(al * np.float32(2.0 ** q), **E4M3)
    bh8 = to_fp8(bh * np.float32(2.0 ** -q), **E4M3)        # weights (max in [2^6, 2^7)) pay for it: 2^(2-head) at most
    y3 = (ah @ bh + ah @ bl + al @ bh) * s16h
    y8 = (ah @ bh + ah8 @ bl8 + al8 @ bh8) * s16h
    for nm, y in (("3 fp16 passes", y3), ("fp16 + 2 x E4M3, fixed scales", y8)):
        err = y.astype(np.float64) - refh
        print("lognormal tail %.1f, headroom 2^%d: %-32s rel. rms err %.2e  (activation max/median %.0f)" %
              (tail, head, nm, np.sqrt((err ** 2).mean()) / rms_h, Ah.max() / np.median(Ah[Ah > 0])))
