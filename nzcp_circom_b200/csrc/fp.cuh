// BN254 prime fields Fr / Fq: 254-bit Montgomery arithmetic on 8 x 32-bit limbs, R = 2^256.
//
// Replaces (upstream, not vendored in /root/reference -- see package.json:12, yarn.lock:408-416,1132-1135)
// wasmcurves 0.1.0 build_f1m.js (f1m_mul / f1m_add / f1m_sub / f1m_toMontgomery / f1m_fromMontgomery), which
// snarkjs reaches through ffjavascript's Fr / F1 objects.  Same representation contract as the WASM code:
// values are canonical (< p), Montgomery form is x*R mod p, buffers are little-endian.
//
// Every function is __host__ __device__: the host build (g++ via nvcc) is what the CPU-side unit tests and
// the O(1) host glue (twiddle tables, proof finalisation) run; the device build swaps in IMAD carry-chain
// PTX for mul/add/sub (sm_100a: mad.lo.cc/madc.hi.cc pairs fuse to IMAD.WIDE.U32 + carry predicates).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

#ifndef NZCP_MUL_PTX
#define NZCP_MUL_PTX 1
#endif

namespace nzcp {

struct FrParams {
  HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqParams {
  HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
  }
  HD static constexpr uint32_t one(int i) {  // R mod q
    constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
  }
  HD static constexpr uint32_t r2(int i) {  // R^2 mod q
    constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return m[i];
  }
  static constexpr uint32_t INV = 0xe4866389u;  // -q^-1 mod 2^32
};

template <class P>
struct alignas(16) Fp {
  uint32_t v[8];

  HD static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  HD static Fp one() {  // Montgomery form of 1
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
    return r;
  }
  HD static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
    return r;
  }
  HD static Fp modulus() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::mod(i);
    return r;
  }
  HD bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  HD bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// ------------------------------------------------------------------------------------------ portable core
template <class P>
HD bool fp_geq_mod(const uint32_t* t) {  // t >= p ?
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    if (t[i] > P::mod(i)) return true;
    if (t[i] < P::mod(i)) return false;
  }
  return true;
}

template <class P>
HD Fp<P> fp_add_portable(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[8], u[8];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.v[i] + b.v[i];
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  // p < 2^254 so a+b < 2^255: no carry out of limb 7.
  int64_t bw = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    bw += (int64_t)t[i] - (int64_t)P::mod(i);
    u[i] = (uint32_t)bw;
    bw >>= 32;
  }
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = bw ? t[i] : u[i];
  return r;
}

template <class P>
HD Fp<P> fp_sub_portable(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[8];
  int64_t bw = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    bw += (int64_t)a.v[i] - (int64_t)b.v[i];
    t[i] = (uint32_t)bw;
    bw >>= 32;
  }
  uint32_t mask = bw ? 0xffffffffu : 0u;
  uint64_t c = 0;
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)t[i] + (P::mod(i) & mask);
    r.v[i] = (uint32_t)c;
    c >>= 32;
  }
  return r;
}

// CIOS Montgomery product, 32-bit digits, 64-bit accumulators.  a, b < p  ->  a*b/R mod p, canonical.
template <class P>
HD Fp<P> fp_mul_portable(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[9];
#pragma unroll
  for (int i = 0; i < 9; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)a.v[j] * b.v[i] + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[8] = (uint32_t)c;  // p < 2^254: never needs a 10th limb
    uint32_t m = t[0] * P::INV;
    c = ((uint64_t)m * P::mod(0) + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      c += (uint64_t)m * P::mod(j) + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[7] = (uint32_t)c;
    t[8] = (uint32_t)(c >> 32);
  }
  Fp<P> r;
  if (fp_geq_mod<P>(t)) {
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      bw += (int64_t)t[i] - (int64_t)P::mod(i);
      r.v[i] = (uint32_t)bw;
      bw >>= 32;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
  }
  return r;
}

// ------------------------------------------------------------------------------------------ device PTX core
#if defined(__CUDA_ARCH__)
// r = (t >= p) ? t - p : t, branch-free.
template <class P>
__device__ __forceinline__ void fp_final_sub(uint32_t* t) {
  uint32_t u[8], bw;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(bw)
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(P::mod(0)), "r"(P::mod(1)), "r"(P::mod(2)), "r"(P::mod(3)), "r"(P::mod(4)), "r"(P::mod(5)),
        "r"(P::mod(6)), "r"(P::mod(7)));
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = bw ? t[i] : u[i];
}

template <class P>
__device__ __forceinline__ Fp<P> fp_add_ptx(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[8];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  fp_final_sub<P>(t);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = t[i];
  return r;
}

template <class P>
__device__ __forceinline__ Fp<P> fp_sub_ptx(const Fp<P>& a, const Fp<P>& b) {
  uint32_t t[8], bw;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(bw)
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
        "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  Fp<P> r;
  // bw = 0xffffffff when a < b: add p back.
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
        "=r"(r.v[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(P::mod(0) & bw), "r"(P::mod(1) & bw), "r"(P::mod(2) & bw), "r"(P::mod(3) & bw), "r"(P::mod(4) & bw),
        "r"(P::mod(5) & bw), "r"(P::mod(6) & bw), "r"(P::mod(7) & bw));
  return r;
}

// One row of the product: (ev, od) hold T = ev + od * 2^32 (od is offset by one limb).
//   first row : od = (a1,a3,a5,a7)*bi ; ev = (a0,a2,a4,a6)*bi
//   other rows: T was divided by 2^32 by the previous reduction, so the arrays swap roles; the old even array
//               (now `od`) is consumed two limbs ahead (od[j+2]) and its stray limb od[1] lands in ev[0].
template <bool FIRST>
__device__ __forceinline__ void fp_mad_row(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
  if (FIRST) {
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(od[0]), "=r"(od[1]), "=r"(od[2]), "=r"(od[3]), "=r"(od[4]), "=r"(od[5]), "=r"(od[6]), "=r"(od[7])
        : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(bi));
    asm("mul.lo.u32 %0, %8, %12;\n\t"
        "mul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\t"
        "mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=r"(ev[0]), "=r"(ev[1]), "=r"(ev[2]), "=r"(ev[3]), "=r"(ev[4]), "=r"(ev[5]), "=r"(ev[6]), "=r"(ev[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(bi));
  } else {
    asm("add.cc.u32 %0, %0, %9;\n\t"          // ev0 += od1 (stray limb), carry continues into od chain
        "madc.lo.cc.u32 %8, %16, %24, %10;\n\t"   // od0 = lo(a1*bi) + od2
        "madc.hi.cc.u32 %9, %16, %24, %11;\n\t"   // od1 = hi(a1*bi) + od3
        "madc.lo.cc.u32 %10, %17, %24, %12;\n\t"  // od2 = lo(a3*bi) + od4
        "madc.hi.cc.u32 %11, %17, %24, %13;\n\t"  // od3 = hi(a3*bi) + od5
        "madc.lo.cc.u32 %12, %18, %24, %14;\n\t"  // od4 = lo(a5*bi) + od6
        "madc.hi.cc.u32 %13, %18, %24, %15;\n\t"  // od5 = hi(a5*bi) + od7
        "madc.lo.cc.u32 %14, %19, %24, 0;\n\t"    // od6 = lo(a7*bi)
        "madc.hi.u32 %15, %19, %24, 0;\n\t"       // od7 = hi(a7*bi) + carry
        "mad.lo.cc.u32 %0, %20, %24, %0;\n\t"     // ev += (a0,a2,a4,a6)*bi
        "madc.hi.cc.u32 %1, %20, %24, %1;\n\t"
        "madc.lo.cc.u32 %2, %21, %24, %2;\n\t"
        "madc.hi.cc.u32 %3, %21, %24, %3;\n\t"
        "madc.lo.cc.u32 %4, %22, %24, %4;\n\t"
        "madc.hi.cc.u32 %5, %22, %24, %5;\n\t"
        "madc.lo.cc.u32 %6, %23, %24, %6;\n\t"
        "madc.hi.cc.u32 %7, %23, %24, %7;\n\t"
        "addc.u32 %15, %15, 0;"                   // carry out of ev (limb 8) -> od7
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
          "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(bi));
  }
}

// Montgomery reduction row: m = ev0 * INV; T += m * p  (makes ev0 == 0).
template <class P>
__device__ __forceinline__ void fp_redc_row(uint32_t* ev, uint32_t* od) {
  uint32_t m = ev[0] * P::INV;
  asm("mad.lo.cc.u32 %8, %16, %24, %8;\n\t"  // od += (p1,p3,p5,p7)*m
      "madc.hi.cc.u32 %9, %16, %24, %9;\n\t"
      "madc.lo.cc.u32 %10, %17, %24, %10;\n\t"
      "madc.hi.cc.u32 %11, %17, %24, %11;\n\t"
      "madc.lo.cc.u32 %12, %18, %24, %12;\n\t"
      "madc.hi.cc.u32 %13, %18, %24, %13;\n\t"
      "madc.lo.cc.u32 %14, %19, %24, %14;\n\t"
      "madc.hi.u32 %15, %19, %24, %15;\n\t"   // top limbs of a and p are < 2^30: cannot carry out
      "mad.lo.cc.u32 %0, %20, %24, %0;\n\t"   // ev += (p0,p2,p4,p6)*m
      "madc.hi.cc.u32 %1, %20, %24, %1;\n\t"
      "madc.lo.cc.u32 %2, %21, %24, %2;\n\t"
      "madc.hi.cc.u32 %3, %21, %24, %3;\n\t"
      "madc.lo.cc.u32 %4, %22, %24, %4;\n\t"
      "madc.hi.cc.u32 %5, %22, %24, %5;\n\t"
      "madc.lo.cc.u32 %6, %23, %24, %6;\n\t"
      "madc.hi.cc.u32 %7, %23, %24, %7;\n\t"
      "addc.u32 %15, %15, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]),
        "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
      : "r"(P::mod(1)), "r"(P::mod(3)), "r"(P::mod(5)), "r"(P::mod(7)), "r"(P::mod(0)), "r"(P::mod(2)),
        "r"(P::mod(4)), "r"(P::mod(6)), "r"(m));
}

template <class P>
__device__ __forceinline__ Fp<P> fp_mul_ptx(const Fp<P>& a, const Fp<P>& b) {
  uint32_t ev[8], od[8];
  fp_mad_row<true>(ev, od, a.v, b.v[0]);
  fp_redc_row<P>(ev, od);
  fp_mad_row<false>(od, ev, a.v, b.v[1]);
  fp_redc_row<P>(od, ev);
  fp_mad_row<false>(ev, od, a.v, b.v[2]);
  fp_redc_row<P>(ev, od);
  fp_mad_row<false>(od, ev, a.v, b.v[3]);
  fp_redc_row<P>(od, ev);
  fp_mad_row<false>(ev, od, a.v, b.v[4]);
  fp_redc_row<P>(ev, od);
  fp_mad_row<false>(od, ev, a.v, b.v[5]);
  fp_redc_row<P>(od, ev);
  fp_mad_row<false>(ev, od, a.v, b.v[6]);
  fp_redc_row<P>(ev, od);
  fp_mad_row<false>(od, ev, a.v, b.v[7]);
  fp_redc_row<P>(od, ev);
  // last row used (od, ev) as (even, odd): result = (od >> 32) + ev
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
      : "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
  fp_final_sub<P>(ev);
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = ev[i];
  return r;
}

// ---- wide product + separate reduction (used by the lazy-reduction Fq2 product in ec.cuh) -------------------------
// T[0..15] = a * b as a full 512-bit integer: the same even/odd rows as fp_mul_ptx with the reduction rows left out;
// after row i the low limb of the running value is product limb i.  a, b < 2^255.
__device__ __forceinline__ void fp_mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t ev[8], od[8];
  fp_mad_row<true>(ev, od, a, b[0]);
  T[0] = ev[0];
  fp_mad_row<false>(od, ev, a, b[1]);
  T[1] = od[0];
  fp_mad_row<false>(ev, od, a, b[2]);
  T[2] = ev[0];
  fp_mad_row<false>(od, ev, a, b[3]);
  T[3] = od[0];
  fp_mad_row<false>(ev, od, a, b[4]);
  T[4] = ev[0];
  fp_mad_row<false>(od, ev, a, b[5]);
  T[5] = od[0];
  fp_mad_row<false>(ev, od, a, b[6]);
  T[6] = ev[0];
  fp_mad_row<false>(od, ev, a, b[7]);
  T[7] = od[0];
  // high half = (even >> 32) + odd, even = od, odd = ev
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7])
      : "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
#pragma unroll
  for (int i = 0; i < 8; i++) T[8 + i] = ev[i];
}

// After a reduction row zeroed even[0]: divide the running value by 2^32.  `ev` = the old odd array (becomes the even
// one), `od` = the old even array (rewritten as the new odd one): ev0 += od1, od[k] = od[k+2] + carry, top limbs 0.
__device__ __forceinline__ void fp_shift_row(uint32_t* ev, uint32_t* od) {
  asm("add.cc.u32 %0, %0, %2;\n\t"
      "addc.cc.u32 %1, %3, 0;\n\t"
      "addc.cc.u32 %2, %4, 0;\n\t"
      "addc.cc.u32 %3, %5, 0;\n\t"
      "addc.cc.u32 %4, %6, 0;\n\t"
      "addc.cc.u32 %5, %7, 0;\n\t"
      "addc.cc.u32 %6, %8, 0;\n\t"
      "addc.u32 %7, 0, 0;\n\t"
      "mov.u32 %8, 0;"
      : "+r"(ev[0]), "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]));
}

// 8-limb add / subtract with carry (borrow) in and out as 0/1 values -- halves of the 16-limb operations below.
__device__ __forceinline__ uint32_t fp_add8c(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t cin) {
  uint32_t cout;
  asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"   // carry flag := cin
      "addc.cc.u32 %0, %9, %17;\n\t"
      "addc.cc.u32 %1, %10, %18;\n\t"
      "addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t"
      "addc.cc.u32 %4, %13, %21;\n\t"
      "addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t"
      "addc.cc.u32 %7, %16, %24;\n\t"
      "addc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(cout)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
  return cout;
}
__device__ __forceinline__ uint32_t fp_sub8b(uint32_t* r, const uint32_t* a, const uint32_t* b, uint32_t bin) {
  uint32_t bout;
  asm("sub.cc.u32 %8, 0, %25;\n\t"            // borrow flag := bin
      "subc.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bout)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(bin));
  return bout & 1u;
}
__device__ __forceinline__ void fp_add16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t c = fp_add8c(r, a, b, 0u);
  fp_add8c(r + 8, a + 8, b + 8, c);
}
__device__ __forceinline__ void fp_sub16(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t bw = fp_sub8b(r, a, b, 0u);
  fp_sub8b(r + 8, a + 8, b + 8, bw);
}

// r = T / 2^256 mod p (canonical) for a 16-limb T < p * 2^256: Montgomery-reduce the LOW half (eight reduction rows, the
// running value never exceeds p + 2^224, so the row's no-carry-out assumption holds), then add the HIGH half:
// (T_lo + M p) / R <= p and T_hi < p, so one conditional subtraction finishes.
template <class P>
__device__ __forceinline__ void fp_redc_wide(uint32_t* r, const uint32_t* T) {
  uint32_t ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ev[i] = T[i];
    od[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    fp_redc_row<P>(ev, od);
    fp_shift_row(od, ev);   // even = od, odd = ev
    fp_redc_row<P>(od, ev);
    fp_shift_row(ev, od);   // even = ev, odd = od
  }
  // value = ev (positions 0..7) + od (positions 1..8; od[7] is zero), plus the high half of T
  uint32_t u[8], o[8];
  o[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) o[i] = od[i - 1];
  fp_add8c(u, ev, o, 0u);
  fp_add8c(u, u, T + 8, 0u);
  fp_final_sub<P>(u);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = u[i];
}
#endif  // __CUDA_ARCH__

// ------------------------------------------------------------------------------------------ host 64-bit core
// The O(1) host glue (proof finalisation: two 256-bit scalar multiplications and three affine conversions per proof)
// sits on the per-proof latency path; on the CPU the same CIOS product runs ~3x faster on 4 x 64-bit digits.
#if !defined(__CUDA_ARCH__)
template <class P>
inline Fp<P> fp_mul_host64(const Fp<P>& a, const Fp<P>& b) {
  typedef unsigned __int128 u128;
  uint64_t A[4], B[4], M[4];
  for (int i = 0; i < 4; i++) {
    A[i] = a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    B[i] = b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
    M[i] = P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
  }
  // -p^-1 mod 2^64 from the 32-bit constant by one Newton step: x' = x * (2 + p * x)  (x = -p^-1)
  uint64_t inv = (uint64_t)P::INV;
  inv = inv * (2 + M[0] * inv);
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)A[j] * B[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * inv;
    c = ((u128)m * M[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * M[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  bool ge = t[4] != 0;
  if (!ge) {
    ge = true;
    for (int i = 3; i >= 0; i--) {
      if (t[i] > M[i]) break;
      if (t[i] < M[i]) { ge = false; break; }
    }
  }
  if (ge) {
    uint64_t bw = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - M[i] - bw;
      t[i] = (uint64_t)d;
      bw = (uint64_t)(d >> 64) & 1;
    }
  }
  Fp<P> r;
  for (int i = 0; i < 4; i++) {
    r.v[2 * i] = (uint32_t)t[i];
    r.v[2 * i + 1] = (uint32_t)(t[i] >> 32);
  }
  return r;
}
#endif

// ------------------------------------------------------------------------------------------ public ops
template <class P>
HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && NZCP_MUL_PTX
  return fp_add_ptx(a, b);
#else
  return fp_add_portable(a, b);
#endif
}
template <class P>
HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && NZCP_MUL_PTX
  return fp_sub_ptx(a, b);
#else
  return fp_sub_portable(a, b);
#endif
}
template <class P>
HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(__CUDA_ARCH__) && NZCP_MUL_PTX
  return fp_mul_ptx(a, b);
#elif defined(__CUDA_ARCH__)
  return fp_mul_portable(a, b);
#else
  return fp_mul_host64(a, b);
#endif
}
template <class P>
HD Fp<P> fp_sqr(const Fp<P>& a) {
  return fp_mul(a, a);
}
template <class P>
HD Fp<P> fp_neg(const Fp<P>& a) {
  return a.is_zero() ? a : fp_sub(Fp<P>::modulus(), a);
}
template <class P>
HD Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add(a, a);
}
template <class P>
HD Fp<P> fp_to_mont(const Fp<P>& a) {
  return fp_mul(a, Fp<P>::r2());
}
template <class P>
HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = Fp<P>::zero();
  o.v[0] = 1;
  return fp_mul(a, o);
}
// a^e, e given as 8 little-endian limbs (plain integer).  a in Montgomery form.
template <class P>
HDN inline Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e) {
  Fp<P> r = Fp<P>::one();
  bool started = false;
  for (int i = 255; i >= 0; i--) {
    if (started) r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) {
      r = started ? fp_mul(r, a) : a;
      started = true;
    }
  }
  return r;
}
// Fermat inverse (a != 0), Montgomery in and out.
template <class P>
HDN inline Fp<P> fp_inv(const Fp<P>& a) {
  uint32_t e[8];
  for (int i = 0; i < 8; i++) e[i] = P::mod(i);
  e[0] -= 2;  // low limb of both moduli is > 2
  return fp_pow(a, e);
}

// ------------------------------------------------------------------------------------------ fast inversion
// Constant-time modular inversion by Bernstein-Yang division steps ("safegcd", half-delta variant: zeta = -(delta + 1/2)),
// 20 rounds of 30 division steps on the low words followed by one 2x2 matrix update of the full-width values (9 signed
// 30-bit limbs, 64-bit accumulators).  590 steps suffice for any odd modulus below 2^256 (Bernstein-Yang 2019 with
// Wuille's half-delta bound), so 600 do.  About 50 Montgomery products' worth of issue slots against ~340 for the Fermat
// ladder, with no data-dependent branch -- every lane of a warp runs the same instruction stream.  It is what makes the
// batched-affine bucket additions of msm.cu pay: one inversion per thread per 32..64 additions.
//
// Montgomery in, Montgomery out: the cofactor track starts at e = R^2 instead of 1, so for the input aR the loop ends with
// R^2 / (aR) = a^-1 R.  fp_inv_fast(0) = 0.
struct FpS30 {
  int32_t v[9];
};

template <class P>
HD void fp_s30_from_u32(FpS30& o, const uint32_t* x) {  // 8 x 32 unsigned -> 9 x 30
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int bit = 30 * k, w = bit >> 5, sh = bit & 31;
    uint32_t lo = x[w] >> sh;
    if (sh > 2 && w + 1 < 8) lo |= x[w + 1] << (32 - sh);
    o.v[k] = (int32_t)(lo & 0x3fffffffu);
  }
}
template <class P>
HD constexpr int32_t fp_mod_s30(int k) {
  const int bit = 30 * k, w = bit >> 5, sh = bit & 31;
  uint32_t lo = P::mod(w) >> sh;
  if (sh > 2 && w + 1 < 8) lo |= P::mod(w + 1) << (32 - sh);
  return (int32_t)(lo & 0x3fffffffu);
}
template <class P>
HD constexpr int32_t fp_r2_s30(int k) {
  const int bit = 30 * k, w = bit >> 5, sh = bit & 31;
  uint32_t lo = P::r2(w) >> sh;
  if (sh > 2 && w + 1 < 8) lo |= P::r2(w + 1) << (32 - sh);
  return (int32_t)(lo & 0x3fffffffu);
}

// 30 division steps on the low 32 bits of f and g: 2^30 (f', g') = [[u, v], [q, r]] (f, g).
HD void fp_divsteps30(int32_t& zeta_io, uint32_t f, uint32_t g, int32_t& uo, int32_t& vo, int32_t& qo, int32_t& ro) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  uint32_t zeta = (uint32_t)zeta_io;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)((int32_t)zeta >> 31);   // all ones when zeta < 0 (delta > 0)
    const uint32_t c2 = 0u - (g & 1u);               // all ones when g is odd
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;   // (f, u, v) negated when zeta < 0
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;                                         // swap: zeta < 0 and g odd
    zeta = (zeta ^ c1) - 1u;
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  zeta_io = (int32_t)zeta;
  uo = (int32_t)u;
  vo = (int32_t)v;
  qo = (int32_t)q;
  ro = (int32_t)r;
}

template <class P>
HDN inline Fp<P> fp_inv_fast(const Fp<P>& a) {
  const int32_t M30 = 0x3fffffff;
  // p^-1 mod 2^30 from INV = -p^-1 mod 2^32
  const uint32_t minv30 = (0u - P::INV) & 0x3fffffffu;
  FpS30 f, g, d, e;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    f.v[k] = fp_mod_s30<P>(k);
    d.v[k] = 0;
    e.v[k] = fp_r2_s30<P>(k);
  }
  fp_s30_from_u32<P>(g, a.v);
  int32_t zeta = -1;
#pragma unroll 1
  for (int round = 0; round < 20; round++) {
    int32_t u, v, q, r;
    fp_divsteps30(zeta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), u, v, q, r);
    // (d, e) <- [[u, v], [q, r]] (d, e) / 2^30 mod p, kept in (-2p, p)
    {
      const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
      int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
      int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
      int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
      md -= (int32_t)((minv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      me -= (int32_t)((minv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
      cd += (int64_t)fp_mod_s30<P>(0) * md;
      ce += (int64_t)fp_mod_s30<P>(0) * me;
      cd >>= 30;
      ce >>= 30;
#pragma unroll
      for (int i = 1; i < 9; i++) {
        cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i] + (int64_t)fp_mod_s30<P>(i) * md;
        ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i] + (int64_t)fp_mod_s30<P>(i) * me;
        d.v[i - 1] = (int32_t)cd & M30;
        e.v[i - 1] = (int32_t)ce & M30;
        cd >>= 30;
        ce >>= 30;
      }
      d.v[8] = (int32_t)cd;
      e.v[8] = (int32_t)ce;
    }
    // (f, g) <- [[u, v], [q, r]] (f, g) / 2^30 (exact)
    {
      int64_t cf = (int64_t)u * f.v[0] + (int64_t)v * g.v[0];
      int64_t cg = (int64_t)q * f.v[0] + (int64_t)r * g.v[0];
      cf >>= 30;
      cg >>= 30;
#pragma unroll
      for (int i = 1; i < 9; i++) {
        cf += (int64_t)u * f.v[i] + (int64_t)v * g.v[i];
        cg += (int64_t)q * f.v[i] + (int64_t)r * g.v[i];
        f.v[i - 1] = (int32_t)cf & M30;
        g.v[i - 1] = (int32_t)cg & M30;
        cf >>= 30;
        cg >>= 30;
      }
      f.v[8] = (int32_t)cf;
      g.v[8] = (int32_t)cg;
    }
  }
  // g = 0, f = +-1, d = +-(result) in (-2p, p): bring into [0, p) with the sign of f
  {
    const int32_t sign = f.v[8] >> 31;
    int32_t add = d.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      int32_t t = d.v[i] + (fp_mod_s30<P>(i) & add);
      d.v[i] = (t ^ sign) - sign;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      d.v[i + 1] += d.v[i] >> 30;
      d.v[i] &= M30;
    }
    add = d.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) d.v[i] += fp_mod_s30<P>(i) & add;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      d.v[i + 1] += d.v[i] >> 30;
      d.v[i] &= M30;
    }
  }
  Fp<P> o;
#pragma unroll
  for (int w = 0; w < 8; w++) {   // 9 x 30 -> 8 x 32
    const int bit = 32 * w, k = bit / 30, sh = bit % 30;
    uint32_t x = (uint32_t)d.v[k] >> sh;
    x |= (uint32_t)d.v[k + 1] << (30 - sh);
    o.v[w] = x;
  }
  return o;
}

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

}  // namespace nzcp
