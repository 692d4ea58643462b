// Batched-affine "pair rounds" of the bucket accumulation (host + device code; kernels in msm.cu, the CPU simulation the
// no-GPU tests run in standalone.cu).
//
// After the bucket sort every bucket owns a contiguous list of points.  One round replaces each list by the sums of its
// adjacent pairs (an odd last point is copied): len -> ceil(len / 2).  The additions are AFFINE,
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1,
// and the additions share their field inversions (Montgomery's trick, two levels).  A round is three kernels:
//   forward   thread t multiplies the denominators of its K pairs into a running product (1 product per addition; the
//             running products go to a scratch array) and publishes the total P_t;
//   invert    the P_t are inverted 32 at a time: prefix products, ONE inversion by division steps (fp_inv_fast, ~50
//             products' worth of issue slots, constant time), back-substitution: 3 + 50/32 products per P_t;
//   backward  thread t peels the individual inverses off 1 / P_t (2 products per addition) and finishes the additions
//             (3 products).
// 6 + ~5 / K field products per addition instead of the 10 of an XYZZ mixed addition (SURVEY.md 8d canonical count),
// and the two hot kernels are short threads of plain Montgomery products -- the long serial inversion sits in a small
// kernel of its own.  Group elements are exact, so the MSM result is bit-identical whichever way its terms were
// associated.
//
// Special pairs never enter the shared product (their denominator is replaced by "skip"):
//   P + infinity, infinity + P -> the other point;  P + (-P) -> infinity;  P + P -> tangent: the denominator is 2 y
//   (y != 0: these curves have no 2-torsion; a y = 0 input is treated as its own inverse, which is what it is).
#pragma once
#include "ec.cuh"


namespace nzcp {

enum PairKind { kPairNormal = 0, kPairDouble = 1, kPairFirst = 2, kPairSecond = 3, kPairInfinity = 4 };

#if defined(__CUDA_ARCH__)
// Read-only 16-byte loads with the .L2::64B fetch-size qualifier (T: a multiple of 16 bytes, 16-byte aligned).
template <class T>
__device__ __forceinline__ T gather_hinted(const T* p) {
  static_assert(sizeof(T) % 16 == 0, "16-byte granules");
  T r;
  uint4* d = reinterpret_cast<uint4*>(&r);
  const char* a = reinterpret_cast<const char*>(p);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++)
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w) : "l"(a + 16 * i));
  return r;
}
#endif

// Where a round reads its points: round 1 gathers from the window table through the sorted entry list (index | sign
// << 31), later rounds read the previous round's array directly.
template <class F, bool FROM_TABLE>
struct PairSource {
  const Affine<F>* pts;
  const uint32_t* entries;
  int pf = 0;   // bits 0-1: how a round asks for its next operands ahead of use: 0 = not at all (default), 1 = prefetch to
                // L1, 2 = to L2; bit 2: fetch-size hint on the table gathers (experiment)
  HD uint32_t slot(uint32_t idx) const { return FROM_TABLE ? entries[idx] : idx; }   // table index | sign, or the index
  // Gathers from the window table.  pf & 4 (experiment knob "gather_hint"): the loads carry the PTX .L2::64B fetch-size
  // qualifier, to find out whether the 128 bytes of DRAM traffic per gathered 64-byte point are a fetch-size default.
  HD F x_at(uint32_t slot) const {
#if defined(__CUDA_ARCH__)
    if (FROM_TABLE && (pf & 4)) return gather_hinted<F>(&pts[slot & 0x7fffffffu].x);
#endif
    return pts[FROM_TABLE ? (slot & 0x7fffffffu) : slot].x;
  }
  HD Affine<F> point_at(uint32_t slot) const {
    if (FROM_TABLE) {
      Affine<F> p;
#if defined(__CUDA_ARCH__)
      if (pf & 4) p = gather_hinted<Affine<F>>(&pts[slot & 0x7fffffffu]);
      else
#endif
        p = pts[slot & 0x7fffffffu];
      if (slot >> 31) p.y = f_neg(p.y);
      return p;
    }
    return pts[slot];
  }
  HD Affine<F> point(uint32_t idx) const { return point_at(slot(idx)); }
  // Optionally ask for the point's cache line ahead of its use (device only).  OFF by default: measured on the B200 with
  // four proofs in flight, the L1 prefetch of the next pair's operands COSTS 5-6 % proofs/s (124 -> 131 without it,
  // profiles/r02_window_rounds_sweep.md): 768 resident threads x 4 lines overrun the L1 and evict lines before their use,
  // and the extra requests queue in front of the demand loads of the kernels running beside this one.
  HD void prefetch(uint32_t slot, bool whole_point) const {
#if defined(__CUDA_ARCH__)
    if ((pf & 3) == 0) return;
    // 64-byte granules (the points are 64-byte aligned): G1 x or whole point = 1, G2 x = 1, G2 whole point = 2
    const char* a = reinterpret_cast<const char*>(pts + (FROM_TABLE ? (slot & 0x7fffffffu) : slot));
    if ((pf & 3) == 1) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
      if (sizeof(F) > 32 && whole_point) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + 64));
    } else {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      if (sizeof(F) > 32 && whole_point) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 64));
    }
#else
    (void)slot;
    (void)whole_point;
#endif
  }
};

// Position of an output slot inside the bucket structure: walks forward / backward over the (possibly empty) buckets.
struct PairWalk {
  const uint32_t* off_in;
  const uint32_t* off_out;
  uint32_t s, out0, out1, in0, len;
  HD void load() {
    out0 = off_out[s];
    out1 = off_out[s + 1];
    in0 = off_in[s];
    len = off_in[s + 1] - in0;
  }
  HD void forward_to(uint32_t o) {   // o >= out0
    if (o < out1) return;
    do s++; while (o >= off_out[s + 1]);
    load();
  }
  HD void backward_to(uint32_t o) {  // o < out1
    if (o >= out0) return;
    do s--; while (o < off_out[s]);
    load();
  }
  HD uint32_t first(uint32_t o) const { return in0 + 2 * (o - out0); }       // input slot of the pair's first point
  HD bool paired(uint32_t o) const { return 2 * (o - out0) + 1 < len; }      // false: odd point out, copied
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

// The pairs of thread t.  The K outputs of a thread are INTERLEAVED across its warp: warp w owns the outputs
// [w * 32 K, (w + 1) * 32 K) and lane l takes w * 32 K + i * 32 + l, i < K -- at every step the 32 lanes touch 32
// consecutive outputs, so the entry list, the previous round's point array and the round's own output are read and
// written as contiguous 2-4 KB runs per warp instead of 32 scattered lines (a thread's K pairs need not be adjacent:
// Montgomery's trick only wants them to be the same pairs in the forward and the backward pass).
static constexpr uint32_t kPairLanes = 32;
template <int K>
HD bool pair_thread_range(uint32_t t, const uint32_t* off_out, uint32_t n_buckets, uint32_t& o0, uint32_t& cnt) {
  const uint32_t n_out = off_out[n_buckets];
  const uint64_t first = (uint64_t)(t / kPairLanes) * (kPairLanes * K) + (t % kPairLanes);
  if (first >= n_out) return false;
  o0 = (uint32_t)first;
  const uint32_t left = (n_out - o0 + kPairLanes - 1) / kPairLanes;
  cnt = left < (uint32_t)K ? left : (uint32_t)K;
  return true;
}
// Threads 0 .. pair_n_prod - 1 are exactly the ones with work (and a denominator product to invert).
template <int K>
HD uint32_t pair_n_prod(uint32_t n_out) {
  const uint32_t full = n_out / (kPairLanes * K), rem = n_out - full * (kPairLanes * K);
  return full * kPairLanes + (rem < kPairLanes ? rem : kPairLanes);
}
HD void pair_walk_seek(PairWalk& wk, uint32_t o, uint32_t n_buckets) {
  uint32_t lo = 0, hi = n_buckets;   // off_out[lo] <= o < off_out[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (wk.off_out[mid] <= o) lo = mid; else hi = mid;
  }
  wk.s = lo;
  wk.load();
}

// Forward kernel body.  off_in / off_out: bucket offsets (n_buckets + 1 entries) of the input and output point arrays;
// scratch: K x stride running products, element (i, t) at i * stride + t; prod[t] = product of the thread's denominators
// (one() when it has none).  Runs one step ahead of itself: while pair i is multiplied the lines of pair i + 1 are on
// their way to L1.
//
// ops (optional, round 1): the thread also writes the two whole operands of every pair to ops[2 * (i * stride + t) + {0, 1}]
// (the second is infinity for an odd point out).  The gather from the window table has pulled their cache lines anyway;
// the backward pass then STREAMS the operands (coalesced: adjacent threads, adjacent 128 / 256 bytes) instead of gathering
// them a second time at random-access DRAM efficiency.
template <class F, bool FROM_TABLE, int K>
HD void msm_pair_forward_body(uint32_t t, uint32_t stride, const PairSource<F, FROM_TABLE>& src, const uint32_t* off_in,
                              const uint32_t* off_out, uint32_t n_buckets, F* scratch, F* prod, Affine<F>* ops = nullptr) {
  uint32_t o0, cnt;
  if (!pair_thread_range<K>(t, off_out, n_buckets, o0, cnt)) return;
  PairWalk wk{off_in, off_out, 0, 0, 0, 0, 0};
  pair_walk_seek(wk, o0, n_buckets);
  F pre = F::one();
  bool have = false;
  bool pair_cur = wk.paired(o0);
  uint32_t a_cur = src.slot(wk.first(o0)), b_cur = pair_cur ? src.slot(wk.first(o0) + 1) : 0;
  for (uint32_t i = 0; i < cnt; i++) {
    bool pair_nxt = false;
    uint32_t a_nxt = 0, b_nxt = 0;
    if (i + 1 < cnt) {
      const uint32_t on = o0 + (i + 1) * kPairLanes;
      wk.forward_to(on);
      pair_nxt = wk.paired(on);
      if (pair_nxt || ops) a_nxt = src.slot(wk.first(on));   // an odd point out is only needed when it is staged
      if (pair_nxt) {
        b_nxt = src.slot(wk.first(on) + 1);
        src.prefetch(a_nxt, false);
        src.prefetch(b_nxt, false);
      }
    }
    if (ops) {
      const Affine<F> p1 = src.point_at(a_cur);
      const Affine<F> p2 = pair_cur ? src.point_at(b_cur) : Affine<F>::inf();
      ops[2 * ((size_t)i * stride + t)] = p1;
      ops[2 * ((size_t)i * stride + t) + 1] = p2;
      if (pair_cur) {
        F den;
        const int kind = pair_classify(p1, p2, den);
        if (kind == kPairNormal || kind == kPairDouble) {
          pre = have ? f_mul(pre, den) : den;
          have = true;
        }
      }
    } else if (pair_cur) {
      const F x1 = src.x_at(a_cur), x2 = src.x_at(b_cur);
      F den = f_sub(x2, x1);
      bool use = true;
      if (den.is_zero() || x1.is_zero() || x2.is_zero()) {   // rare: look at the whole points
        const int kind = pair_classify(src.point_at(a_cur), src.point_at(b_cur), den);
        use = kind == kPairNormal || kind == kPairDouble;
      }
      if (use) {
        pre = have ? f_mul(pre, den) : den;
        have = true;
      }
    }
    scratch[(size_t)i * stride + t] = pre;
    pair_cur = pair_nxt;
    a_cur = a_nxt;
    b_cur = b_nxt;
  }
  prod[t] = pre;
}

// Invert kernel body: v[g * KI .. g * KI + KI) <- their inverses, one inversion for the group.  No element is zero
// (denominators of special pairs never enter a product).
template <class F, int KI>
HD void msm_pair_invert_body(uint32_t g, F* v, uint32_t n) {
  const uint32_t b = g * KI;
  if (b >= n) return;
  const uint32_t cnt = n - b < (uint32_t)KI ? n - b : (uint32_t)KI;
  F pre[KI];
  F run = v[b];
  pre[0] = run;
  for (uint32_t i = 1; i < cnt; i++) {
    run = f_mul(run, v[b + i]);
    pre[i] = run;
  }
  F inv = f_inv_fast(run);
  for (uint32_t i = cnt; i-- > 1;) {
    const F x = v[b + i];
    v[b + i] = f_mul(inv, pre[i - 1]);
    inv = f_mul(inv, x);
  }
  v[b] = inv;
}

// Backward kernel body: inv_prod[t] = 1 / prod[t].
// One affine addition (or its special cases) with the shared inverse peeled off `inv`; `before` = product of the thread's
// denominators before this pair's.
template <class F>
HD Affine<F> pair_finish(const Affine<F>& p1, const Affine<F>& p2, F& inv, const F& before) {
  F den;
  const int kind = pair_classify(p1, p2, den);
  if (kind >= kPairFirst) return kind == kPairFirst ? p1 : kind == kPairSecond ? p2 : Affine<F>::inf();
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
  return Affine<F>{x3, y3};
}

template <class F, bool FROM_TABLE, int K>
HD void msm_pair_backward_body(uint32_t t, uint32_t stride, const PairSource<F, FROM_TABLE>& src, const uint32_t* off_in,
                               const uint32_t* off_out, uint32_t n_buckets, Affine<F>* dst, const F* scratch,
                               const F* inv_prod, const Affine<F>* ops = nullptr) {
  uint32_t o0, cnt;
  if (!pair_thread_range<K>(t, off_out, n_buckets, o0, cnt)) return;
  if (ops) {   // operands staged by the forward pass: no bucket walk, no gather -- an odd point out is (P, infinity) -> P
    F inv = inv_prod[t];
    for (uint32_t i = cnt; i-- > 0;) {
      const Affine<F> p1 = ops[2 * ((size_t)i * stride + t)], p2 = ops[2 * ((size_t)i * stride + t) + 1];
      const F before = i ? scratch[(size_t)(i - 1) * stride + t] : F::one();
      dst[o0 + i * kPairLanes] = pair_finish(p1, p2, inv, before);
    }
    return;
  }
  PairWalk wk{off_in, off_out, 0, 0, 0, 0, 0};
  pair_walk_seek(wk, o0 + (cnt - 1) * kPairLanes, n_buckets);
  F inv = inv_prod[t];
  bool pair_cur;
  uint32_t a_cur, b_cur;
  {
    const uint32_t o = o0 + (cnt - 1) * kPairLanes;
    pair_cur = wk.paired(o);
    a_cur = src.slot(wk.first(o));
    b_cur = pair_cur ? src.slot(wk.first(o) + 1) : 0;
  }
  for (uint32_t i = cnt; i-- > 0;) {
    const uint32_t o = o0 + i * kPairLanes;
    bool pair_nxt = false;
    uint32_t a_nxt = 0, b_nxt = 0;
    if (i > 0) {
      wk.backward_to(o - kPairLanes);
      pair_nxt = wk.paired(o - kPairLanes);
      a_nxt = src.slot(wk.first(o - kPairLanes));
      src.prefetch(a_nxt, true);
      if (pair_nxt) {
        b_nxt = src.slot(wk.first(o - kPairLanes) + 1);
        src.prefetch(b_nxt, true);
      }
    }
    if (!pair_cur) {     // odd point out: carried to the next round as it is
      dst[o] = src.point_at(a_cur);
    } else {
      const F before = i ? scratch[(size_t)(i - 1) * stride + t] : F::one();   // product of the denominators before this one
      dst[o] = pair_finish(src.point_at(a_cur), src.point_at(b_cur), inv, before);
    }
    pair_cur = pair_nxt;
    a_cur = a_nxt;
    b_cur = b_nxt;
  }
}

}  // namespace nzcp
