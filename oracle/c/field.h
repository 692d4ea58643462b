/* ORACLE (test infrastructure, NOT product code) -- BN254 Fr / Fq on 4 x 64-bit limbs, Montgomery R = 2^256.
 *
 * Restates wasmcurves 0.1.0 build_f1m.js (f1m_mul / f1m_add / f1m_sub / f1m_toMontgomery / f1m_fromMontgomery):
 * the WASM code is a 64-bit-accumulator CIOS Montgomery product over little-endian limbs; this is the same algorithm
 * with native 64-bit digits.  Upstream is not vendored in /root/reference (package.json:12 -> snarkjs ^0.4.12;
 * yarn.lock:408-416 ffjavascript 0.2.48; yarn.lock:1132-1135 wasmcurves 0.1.0).  "parity unpinned" -- see
 * oracle/bn254.py for what pins it.
 */
#ifndef NZO_FIELD_H
#define NZO_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;

#define NZO_DEFINE_FIELD(N, P0, P1, P2, P3, INV, O0, O1, O2, O3, S0, S1, S2, S3)                              \
  static const uint64_t N##_P[4] = {P0, P1, P2, P3};                                                          \
  static const fe N##_ONE = {{O0, O1, O2, O3}}; /* R mod p   */                                                \
  static const fe N##_R2 = {{S0, S1, S2, S3}};  /* R^2 mod p */                                                \
  static inline int N##_geq_p(const uint64_t* t) {                                                            \
    for (int i = 3; i >= 0; i--) {                                                                            \
      if (t[i] > N##_P[i]) return 1;                                                                          \
      if (t[i] < N##_P[i]) return 0;                                                                          \
    }                                                                                                         \
    return 1;                                                                                                 \
  }                                                                                                           \
  static inline void N##_subp(uint64_t* t) {                                                                  \
    u128 b = 0;                                                                                               \
    for (int i = 0; i < 4; i++) {                                                                             \
      u128 d = (u128)t[i] - N##_P[i] - (uint64_t)b;                                                           \
      t[i] = (uint64_t)d;                                                                                     \
      b = (d >> 64) & 1;                                                                                      \
    }                                                                                                         \
  }                                                                                                           \
  static inline void N##_add(fe* r, const fe* a, const fe* b) {                                               \
    u128 c = 0;                                                                                               \
    uint64_t t[4];                                                                                            \
    for (int i = 0; i < 4; i++) {                                                                             \
      c += (u128)a->v[i] + b->v[i];                                                                           \
      t[i] = (uint64_t)c;                                                                                     \
      c >>= 64;                                                                                               \
    }                                                                                                         \
    if (N##_geq_p(t)) N##_subp(t);                                                                            \
    memcpy(r->v, t, 32);                                                                                      \
  }                                                                                                           \
  static inline void N##_sub(fe* r, const fe* a, const fe* b) {                                               \
    uint64_t t[4];                                                                                            \
    uint64_t bw = 0;                                                                                          \
    for (int i = 0; i < 4; i++) {                                                                             \
      u128 d = (u128)a->v[i] - b->v[i] - bw;                                                                  \
      t[i] = (uint64_t)d;                                                                                     \
      bw = (uint64_t)(d >> 64) & 1;                                                                           \
    }                                                                                                         \
    if (bw) {                                                                                                 \
      u128 c = 0;                                                                                             \
      for (int i = 0; i < 4; i++) {                                                                           \
        c += (u128)t[i] + N##_P[i];                                                                           \
        t[i] = (uint64_t)c;                                                                                   \
        c >>= 64;                                                                                             \
      }                                                                                                       \
    }                                                                                                         \
    memcpy(r->v, t, 32);                                                                                      \
  }                                                                                                           \
  static inline void N##_mul(fe* r, const fe* a, const fe* b) {                                               \
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};                                                                       \
    for (int i = 0; i < 4; i++) {                                                                             \
      u128 c = 0;                                                                                             \
      for (int j = 0; j < 4; j++) {                                                                           \
        c += (u128)a->v[j] * b->v[i] + t[j];                                                                  \
        t[j] = (uint64_t)c;                                                                                   \
        c >>= 64;                                                                                             \
      }                                                                                                       \
      c += t[4];                                                                                              \
      t[4] = (uint64_t)c;                                                                                     \
      t[5] = (uint64_t)(c >> 64);                                                                             \
      uint64_t m = t[0] * INV;                                                                                \
      c = ((u128)m * N##_P[0] + t[0]) >> 64;                                                                  \
      for (int j = 1; j < 4; j++) {                                                                           \
        c += (u128)m * N##_P[j] + t[j];                                                                       \
        t[j - 1] = (uint64_t)c;                                                                               \
        c >>= 64;                                                                                             \
      }                                                                                                       \
      c += t[4];                                                                                              \
      t[3] = (uint64_t)c;                                                                                     \
      t[4] = t[5] + (uint64_t)(c >> 64);                                                                      \
    }                                                                                                         \
    if (t[4] || N##_geq_p(t)) N##_subp(t);                                                                    \
    memcpy(r->v, t, 32);                                                                                      \
  }                                                                                                           \
  static inline void N##_sqr(fe* r, const fe* a) { N##_mul(r, a, a); }                                        \
  static inline int N##_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }         \
  static inline int N##_eq(const fe* a, const fe* b) { return memcmp(a->v, b->v, 32) == 0; }                  \
  static inline void N##_neg(fe* r, const fe* a) {                                                            \
    if (N##_is_zero(a)) { *r = *a; return; }                                                                  \
    fe z = {{0, 0, 0, 0}};                                                                                    \
    N##_sub(r, &z, a);                                                                                        \
  }                                                                                                           \
  static inline void N##_to_mont(fe* r, const fe* a) { N##_mul(r, a, &N##_R2); }                              \
  static inline void N##_from_mont(fe* r, const fe* a) {                                                      \
    fe o = {{1, 0, 0, 0}};                                                                                    \
    N##_mul(r, a, &o);                                                                                        \
  }                                                                                                           \
  static void N##_pow(fe* r, const fe* a, const uint64_t* e) {                                                \
    fe acc = N##_ONE, b = *a;                                                                                 \
    for (int i = 0; i < 256; i++) {                                                                           \
      if ((e[i >> 6] >> (i & 63)) & 1) N##_mul(&acc, &acc, &b);                                               \
      N##_sqr(&b, &b);                                                                                        \
    }                                                                                                         \
    *r = acc;                                                                                                 \
  }                                                                                                           \
  static void N##_inv(fe* r, const fe* a) {                                                                   \
    uint64_t e[4] = {N##_P[0] - 2, N##_P[1], N##_P[2], N##_P[3]};                                             \
    N##_pow(r, a, e);                                                                                         \
  }

NZO_DEFINE_FIELD(fr, 0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull,
                 0xc2e1f593efffffffull, 0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull,
                 0x0e0a77c19a07df2full, 0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                 0x0216d0b17f4e44a5ull)
NZO_DEFINE_FIELD(fq, 0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull,
                 0x87d20782e4866389ull, 0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull,
                 0x0e0a77c19a07df2full, 0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull,
                 0x06d89f71cab8351full)

/* Fq2 = Fq[u]/(u^2 + 1)  (ffjavascript src/f2field.js) */
typedef struct { fe c0, c1; } fe2;
static const fe2 fq2_ONE = {{{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
                            {{0, 0, 0, 0}}};
static inline void fq2_add(fe2* r, const fe2* a, const fe2* b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fe2* r, const fe2* a, const fe2* b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_neg(fe2* r, const fe2* a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }
static inline void fq2_mul(fe2* r, const fe2* a, const fe2* b) {
  fe t0, t1, t2, s0, s1;
  fq_mul(&t0, &a->c0, &b->c0);
  fq_mul(&t1, &a->c1, &b->c1);
  fq_add(&s0, &a->c0, &a->c1);
  fq_add(&s1, &b->c0, &b->c1);
  fq_mul(&t2, &s0, &s1);
  fq_sub(&r->c0, &t0, &t1);
  fq_sub(&t2, &t2, &t0);
  fq_sub(&r->c1, &t2, &t1);
}
static inline void fq2_sqr(fe2* r, const fe2* a) {
  fe s, d, t0, t1;
  fq_add(&s, &a->c0, &a->c1);
  fq_sub(&d, &a->c0, &a->c1);
  fq_mul(&t0, &s, &d);
  fq_mul(&t1, &a->c0, &a->c1);
  r->c0 = t0;
  fq_add(&r->c1, &t1, &t1);
}
static inline int fq2_is_zero(const fe2* a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fe2* a, const fe2* b) { return fq_eq(&a->c0, &b->c0) && fq_eq(&a->c1, &b->c1); }
static void fq2_inv(fe2* r, const fe2* a) {
  fe t0, t1, d;
  fq_sqr(&t0, &a->c0);
  fq_sqr(&t1, &a->c1);
  fq_add(&t0, &t0, &t1);
  fq_inv(&d, &t0);
  fq_mul(&r->c0, &a->c0, &d);
  fq_mul(&t1, &a->c1, &d);
  fq_neg(&r->c1, &t1);
}
static inline void fq2_from_mont(fe2* r, const fe2* a) { fq_from_mont(&r->c0, &a->c0); fq_from_mont(&r->c1, &a->c1); }
#endif
