// Batched-affine "pair rounds" of the bucket accumulation (host + device code; kernels in msm.cu, the CPU simulation the
// no-GPU tests run in standalone.cu).
//
// After the bucket sort every bucket owns a contiguous list of points.  One round replaces each list by the sums of its
// adjacent pairs (an odd last point is copied): len -> ceil(len / 2).  The additions are AFFINE,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// and the K additions of one thread share ONE field inversion (Montgomery's trick): a forward pass multiplies the
// denominators into a running product (1 product per addition; the running products go to a scratch array), the
// product is inverted once by division steps (fp_inv_fast, ~50 products), and a backward pass peels the individual
// inverses off (2 products) and finishes the additions (3 products): 6 + 50 / K field products per addition instead of
// the 10 of an XYZZ mixed addition (SURVEY.md 8d canonical count).  Group elements are exact, so the MSM result is
// bit-identical whichever way its terms were associated.
//
// Special pairs never enter the shared product (their denominator is replaced by "skip"):
//   P + infinity, infinity + P -> the other point;  P + (-P) -> infinity;  P + P -> tangent: the denominator is 2 y
//   (y != 0: these curves have no 2-torsion; a y = 0 input is treated as its own inverse, which is what it is).
#pragma once
#include "ec.cuh"

namespace nzcp {

enum PairKind { kPairNormal = 0, kPairDouble = 1, kPairFirst = 2, kPairSecond = 3, kPairInfinity = 4 };

// Where a round reads its points: round 1 gathers from the window table through the sorted entry list (index | sign
// << 31), later rounds read the previous round's array directly.
template <class F, bool FROM_TABLE>
struct PairSource {
  const Affine<F>* pts;
  const uint32_t* entries;
  HD F x(uint32_t idx) const {
    if (FROM_TABLE) idx = entries[idx] & 0x7fffffffu;
    return pts[idx].x;
  }
  HD Affine<F> point(uint32_t idx) const {
    if (FROM_TABLE) {
      const uint32_t e = entries[idx];
      Affine<F> p = pts[e & 0x7fffffffu];
      if (e >> 31) p.y = f_neg(p.y);
      return p;
    }
    return pts[idx];
  }
};

// Kind of the pair and, for the two kinds that need an inversion, its denominator.
template <class F>
HD int pair_classify(const Affine<F>& p1, const Affine<F>& p2, F& den) {
  if (p1.is_inf()) return kPairSecond;
  if (p2.is_inf()) return kPairFirst;
  den = f_sub(p2.x, p1.x);
  if (den.is_zero()) {
    if (p1.y == p2.y && !p1.y.is_zero()) {
      den = f_dbl(p1.y);
      return kPairDouble;
    }
    return kPairInfinity;
  }
  return kPairNormal;
}

// One thread of a pair round: outputs [t * K, t * K + K) of the round.  off_in / off_out: bucket offsets (n_buckets + 1
// entries) of the input and output point arrays; scratch: K x stride running products, element (i, t) at i * stride + t.
template <class F, bool FROM_TABLE, int K>
HD void msm_pair_round_body(uint32_t t, uint32_t stride, const PairSource<F, FROM_TABLE>& src, const uint32_t* off_in,
                            const uint32_t* off_out, uint32_t n_buckets, Affine<F>* dst, F* scratch) {
  const uint32_t n_out = off_out[n_buckets];
  if ((uint64_t)t * K >= n_out) return;
  const uint32_t o0 = t * K;
  const uint32_t cnt = n_out - o0 < (uint32_t)K ? n_out - o0 : (uint32_t)K;
  uint32_t s;
  {
    uint32_t lo = 0, hi = n_buckets;   // off_out[lo] <= o0 < off_out[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (off_out[mid] <= o0) lo = mid; else hi = mid;
    }
    s = lo;
  }
  uint32_t seg_out0 = off_out[s], seg_out1 = off_out[s + 1], seg_in0 = off_in[s], seg_len = off_in[s + 1] - seg_in0;
  F pre = F::one();
  bool have = false;
  for (uint32_t i = 0; i < cnt; i++) {
    const uint32_t o = o0 + i;
    while (o >= seg_out1) {   // next non-empty bucket
      s++;
      seg_out0 = seg_out1;
      seg_out1 = off_out[s + 1];
      seg_in0 = off_in[s];
      seg_len = off_in[s + 1] - seg_in0;
    }
    const uint32_t j = o - seg_out0, i0 = seg_in0 + 2 * j;
    if (2 * j + 1 < seg_len) {
      const F x1 = src.x(i0), x2 = src.x(i0 + 1);
      F den = f_sub(x2, x1);
      bool use = true;
      if (den.is_zero() || x1.is_zero() || x2.is_zero()) {   // rare: look at the whole points
        const int kind = pair_classify(src.point(i0), src.point(i0 + 1), den);
        use = kind == kPairNormal || kind == kPairDouble;
      }
      if (use) {
        pre = have ? f_mul(pre, den) : den;
        have = true;
      }
    }
    scratch[(size_t)i * stride + t] = pre;
  }
  F inv = have ? f_inv_fast(pre) : F::one();
  for (uint32_t i = cnt; i-- > 0;) {
    const uint32_t o = o0 + i;
    while (o < seg_out0) {    // previous non-empty bucket
      s--;
      seg_out1 = seg_out0;
      seg_out0 = off_out[s];
      seg_in0 = off_in[s];
      seg_len = off_in[s + 1] - seg_in0;
    }
    const uint32_t j = o - seg_out0, i0 = seg_in0 + 2 * j;
    if (2 * j + 1 >= seg_len) {     // odd point out: carried to the next round as it is
      dst[o] = src.point(i0);
      continue;
    }
    const Affine<F> p1 = src.point(i0), p2 = src.point(i0 + 1);
    F den;
    const int kind = pair_classify(p1, p2, den);
    if (kind >= kPairFirst) {
      dst[o] = kind == kPairFirst ? p1 : kind == kPairSecond ? p2 : Affine<F>::inf();
      continue;
    }
    const F before = i ? scratch[(size_t)(i - 1) * stride + t] : F::one();   // product of the denominators before this one
    const F dinv = f_mul(inv, before);
    inv = f_mul(inv, den);
    F num;
    if (kind == kPairDouble) {
      const F xx = f_sqr(p1.x);
      num = f_add(f_dbl(xx), xx);
    } else {
      num = f_sub(p2.y, p1.y);
    }
    const F lam = f_mul(num, dinv);
    const F x3 = f_sub(f_sub(f_sqr(lam), p1.x), p2.x);
    const F y3 = f_sub(f_mul(lam, f_sub(p1.x, x3)), p1.y);
    dst[o] = Affine<F>{x3, y3};
  }
}

}  // namespace nzcp
