/* ORACLE (test infrastructure, NOT product code) -- snarkjs groth16.prove restated in plain C for the host CPU.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * It is the checker and the reported CPU baseline, never the product path.
 *
 * PARITY STATUS: "parity unpinned".  The algorithm lives in third-party npm packages that are NOT vendored in
 * /root/reference (package.json:12 -> snarkjs ^0.4.12; yarn.lock:987-999 snarkjs 0.4.12, yarn.lock:408-416
 * ffjavascript 0.2.48, yarn.lock:1132-1135 wasmcurves 0.1.0) and the reference has no call site, test or golden
 * vector for it.  What pins this file: it must agree bit-for-bit with oracle/prover.py (pure
 * Python restatement) on the committed fixtures, and every proof must equal the toxic-waste closed form and pass the
 * pairing check (tests/test_oracle_c.py).
 *
 * Restated, by upstream function:
 *   snarkjs src/groth16_prove.js   groth16Prove, buildABC1, joinABC
 *   ffjavascript src/engine_fft.js  fft / ifft (natural order in and out; threads over butterfly blocks)
 *   ffjavascript src/engine_applykey.js batchApplyKey(buf, 1, inc)
 *   ffjavascript src/engine_multiexp.js _multiExp: c = pTSizes[log2 n], one task per (point slice, window) on the
 *       worker pool, partial sums added per window, Horner by c doublings
 *   wasmcurves build_multiexp.js / build_curve_jacobian_a0.js / build_f1m.js / build_fft.js / build_qap.js
 * Threading mirrors ffjavascript's worker pool with OpenMP: the same task decomposition, `threads` workers.
 */
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "field.h"

#define FT fe
#define FN(x) fq_##x
#define PT g1_jac
#define AT g1_aff
#define CN(x) g1_##x
#include "curve_tmpl.h"
#undef FT
#undef FN
#undef PT
#undef AT
#undef CN

#define FT fe2
#define FN(x) fq2_##x
#define PT g2_jac
#define AT g2_aff
#define CN(x) g2_##x
#include "curve_tmpl.h"
#undef FT
#undef FN
#undef PT
#undef AT
#undef CN

static double now_sec(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------------------------------------------------------ roots of unity */
/* ffjavascript f1field.js: nqr = 5, w[s] = nqr^t, w[k-1] = w[k]^2  (s = 28) */
static void fr_root(fe* w, int k) {
  uint64_t e[4] = {fr_P[0] - 1, fr_P[1], fr_P[2], fr_P[3]};
  uint64_t t[4];
  for (int i = 0; i < 4; i++) t[i] = (e[i] >> 28) | (i + 1 < 4 ? e[i + 1] << 36 : 0);
  fe five = {{5, 0, 0, 0}}, f5;
  fr_to_mont(&f5, &five);
  fr_pow(w, &f5, t);
  for (int i = 28; i > k; i--) fr_sqr(w, w);
}

/* ------------------------------------------------------------------------------------------------ FFT */
static size_t bitrev(size_t i, int bits) {
  size_t r = 0;
  for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
  return r;
}

/* natural-order radix-2 transform of Montgomery-form values (Fr.fft / Fr.ifft) */
static void fr_fft(fe* a, int log_n, int inverse, int threads) {
  size_t n = (size_t)1 << log_n;
  if (log_n == 0) return;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (j > i) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
  }
  fe w;
  fr_root(&w, log_n);
  if (inverse) fr_inv(&w, &w);
  fe* tw = (fe*)malloc((n / 2 ? n / 2 : 1) * sizeof(fe));
  tw[0] = fr_ONE;
  for (size_t j = 1; j < n / 2; j++) fr_mul(&tw[j], &tw[j - 1], &w);
  for (int s = 1; s <= log_n; s++) {
    size_t m = (size_t)1 << s, half = m >> 1, step = n >> s;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t b = 0; b < n / 2; b++) {
      size_t k = (b / half) * m, j = b % half;
      fe t, u = a[k + j];
      fr_mul(&t, &tw[j * step], &a[k + j + half]);
      fr_add(&a[k + j], &u, &t);
      fr_sub(&a[k + j + half], &u, &t);
    }
  }
  if (inverse) {
    fe nn = {{n, 0, 0, 0}}, ninv;
    fr_to_mont(&nn, &nn);
    fr_inv(&ninv, &nn);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) fr_mul(&a[i], &a[i], &ninv);
  }
  free(tw);
}

/* Fr.batchApplyKey(buf, first = 1, inc): element i *= inc^i */
static void fr_apply_key(fe* a, size_t n, const fe* inc, int threads) {
  int nt = threads < 1 ? 1 : threads;
  size_t chunk = (n + nt - 1) / nt;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int t = 0; t < nt; t++) {
    size_t b = (size_t)t * chunk, e = b + chunk < n ? b + chunk : n;
    if (b >= e) continue;
    uint64_t ex[4] = {b, 0, 0, 0};
    fe k;
    fr_pow(&k, inc, ex);
    for (size_t i = b; i < e; i++) {
      fr_mul(&a[i], &a[i], &k);
      fr_mul(&k, &k, inc);
    }
  }
}

/* ------------------------------------------------------------------------------------------------ multiexp */
static const int PT_SIZES[32] = {1, 1, 1, 1, 2, 3, 4, 5, 6, 7, 7, 8, 9, 10, 11, 12, 13, 13, 14, 15, 16, 16, 17, 17, 17, 17, 17, 17, 17, 17, 17, 17};

static int ilog2(size_t v) {
  int l = 0;
  while (v >>= 1) l++;
  return l;
}

#define DEFINE_MULTIEXP(G)                                                                                     \
  static void G##_multiexp(G##_jac* out, const G##_aff* bases, const uint8_t* scalars, size_t n, int threads) { \
    G##_set_inf(out);                                                                                          \
    if (n == 0) return;                                                                                        \
    int c = PT_SIZES[ilog2(n)];                                                                                \
    int n_win = (32 * 8 - 1) / c + 1;                                                                          \
    /* engine_multiexp.js: point slices of n / concurrency, clamped to [2^10, 2^22] */                         \
    size_t slice = n / (threads > 0 ? threads : 1);                                                            \
    if (slice > ((size_t)1 << 22)) slice = (size_t)1 << 22;                                                    \
    if (slice < ((size_t)1 << 10)) slice = (size_t)1 << 10;                                                    \
    size_t n_slices = (n + slice - 1) / slice;                                                                 \
    size_t n_tasks = n_slices * n_win;                                                                         \
    G##_jac* partial = (G##_jac*)malloc(n_tasks * sizeof(G##_jac));                                            \
    _Pragma("omp parallel num_threads(threads)") {                                                             \
      G##_jac* buckets = (G##_jac*)malloc(((size_t)1 << c) * sizeof(G##_jac));                                 \
      _Pragma("omp for schedule(dynamic, 1)") for (size_t t = 0; t < n_tasks; t++) {                           \
        size_t sl = t / n_win;                                                                                 \
        int w = (int)(t % n_win);                                                                              \
        size_t b = sl * slice, e = b + slice < n ? b + slice : n;                                              \
        G##_multiexp_window(&partial[t], bases + b, scalars + 32 * b, e - b, w * c, c, buckets);               \
      }                                                                                                        \
      free(buckets);                                                                                           \
    }                                                                                                          \
    G##_jac res;                                                                                               \
    G##_set_inf(&res);                                                                                         \
    for (int w = n_win - 1; w >= 0; w--) {                                                                     \
      if (!G##_is_inf(&res))                                                                                   \
        for (int k = 0; k < c; k++) G##_dbl(&res, &res);                                                       \
      for (size_t sl = 0; sl < n_slices; sl++) G##_add(&res, &res, &partial[sl * n_win + w]);                  \
    }                                                                                                          \
    free(partial);                                                                                             \
    *out = res;                                                                                                \
  }
DEFINE_MULTIEXP(g1)
DEFINE_MULTIEXP(g2)

/* ------------------------------------------------------------------------------------------------ containers */
typedef struct { const uint8_t* p; uint64_t len; } section_t;

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

static int parse_container(const uint8_t* b, size_t len, const char* magic, section_t* secs, int max_id) {
  if (len < 12 || memcmp(b, magic, 4) != 0) return -2;
  uint32_t nsec = rd32(b + 8);
  size_t pos = 12;
  for (uint32_t i = 0; i < nsec; i++) {
    if (pos + 12 > len) return -2;
    uint32_t id = rd32(b + pos);
    uint64_t sl = rd64(b + pos + 4);
    pos += 12;
    if (sl > len - pos) return -2;
    if ((int)id <= max_id && !secs[id].p) { secs[id].p = b + pos; secs[id].len = sl; }
    pos += sl;
  }
  return 0;
}

static void g1_out_plain(uint8_t* out, const g1_jac* p) {
  g1_aff a;
  g1_to_aff(&a, p);
  fe t;
  fq_from_mont(&t, &a.x); memcpy(out, t.v, 32);
  fq_from_mont(&t, &a.y); memcpy(out + 32, t.v, 32);
}
static void g2_out_plain(uint8_t* out, const g2_jac* p) {
  g2_aff a;
  g2_to_aff(&a, p);
  fe2 t;
  fq2_from_mont(&t, &a.x); memcpy(out, t.c0.v, 32); memcpy(out + 32, t.c1.v, 32);
  fq2_from_mont(&t, &a.y); memcpy(out + 64, t.c0.v, 32); memcpy(out + 96, t.c1.v, 32);
}

/* ------------------------------------------------------------------------------------------------ public C ABI */
int nzo_max_threads(void) { return omp_get_max_threads(); }

/* op: 0 mul 1 add 2 sub; field 0 Fr 1 Fq (Montgomery in/out for mul) */
int nzo_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    fe x, y, r;
    memcpy(x.v, a + 32 * i, 32);
    memcpy(y.v, b + 32 * i, 32);
    if (field == 0) { if (op == 0) fr_mul(&r, &x, &y); else if (op == 1) fr_add(&r, &x, &y); else fr_sub(&r, &x, &y); }
    else { if (op == 0) fq_mul(&r, &x, &y); else if (op == 1) fq_add(&r, &x, &y); else fq_sub(&r, &x, &y); }
    memcpy(out + 32 * i, r.v, 32);
  }
  return 0;
}

/* in-place natural-order (i)NTT of 2^log_n Montgomery-form Fr values */
int nzo_ntt(uint8_t* data, int log_n, int inverse, int threads) {
  if (log_n < 0 || log_n > 28) return -1;
  fr_fft((fe*)data, log_n, inverse, threads > 0 ? threads : 1);
  return 0;
}

/* the per-polynomial H pipeline: ifft -> batchApplyKey(1, inc) -> fft */
int nzo_ntt_coset(uint8_t* data, int log_n, int threads) {
  if (log_n < 0 || log_n > 28) return -1;
  if (threads < 1) threads = 1;
  fe inc;
  if (log_n == 28) { fe s = {{25, 0, 0, 0}}; fr_to_mont(&inc, &s); } else fr_root(&inc, log_n + 1);
  fr_fft((fe*)data, log_n, 1, threads);
  fr_apply_key((fe*)data, (size_t)1 << log_n, &inc, threads);
  fr_fft((fe*)data, log_n, 0, threads);
  return 0;
}

/* bases: Montgomery affine (64 / 128 B); scalars: plain 32-byte LE; out: plain affine */
int nzo_msm(const uint8_t* bases, const uint8_t* scalars, size_t n, int g2, uint8_t* out, int threads) {
  if (threads < 1) threads = 1;
  if (g2) {
    g2_jac r;
    g2_multiexp(&r, (const g2_aff*)bases, scalars, n, threads);
    g2_out_plain(out, &r);
  } else {
    g1_jac r;
    g1_multiexp(&r, (const g1_aff*)bases, scalars, n, threads);
    g1_out_plain(out, &r);
  }
  return 0;
}

/* groth16Prove.  proof: 8 x 32 B plain (A.x A.y B.x0 B.x1 B.y0 B.y1 C.x C.y).  parts (optional): msm A | B1 | B2 | C | H
 * = 64+64+128+64+64 B plain affine.  h_out (optional): domainSize x 32 B joinABC output.
 * stage_sec (optional, 8): 0 parse, 1 buildABC1, 2 ffts+join, 3 msm A, 4 msm B1, 5 msm B2, 6 msm C, 7 msm H.
 * returns 0, or -2 format, -3 not groth16, -4 curve, -5 witness length */
int nzo_prove(const uint8_t* zkey, size_t zlen, const uint8_t* wtns, size_t wlen, const uint8_t* r_le, const uint8_t* s_le,
              uint8_t* proof, uint8_t* parts, uint8_t* h_out, int threads, double* stage_sec) {
  if (threads < 1) threads = 1;
  double t0 = now_sec(), t1;
  section_t zs[11], ws[3];
  memset(zs, 0, sizeof zs);
  memset(ws, 0, sizeof ws);
  if (parse_container(zkey, zlen, "zkey", zs, 10)) return -2;
  if (parse_container(wtns, wlen, "wtns", ws, 2)) return -2;
  for (int i = 1; i <= 9; i++) if (!zs[i].p) return -2;
  if (!ws[1].p || !ws[2].p) return -2;
  if (rd32(zs[1].p) != 1) return -3;
  const uint8_t* h = zs[2].p;
  if (rd32(h) != 32 || rd32(h + 36) != 32) return -4;
  if (rd32(ws[1].p) != 32 || memcmp(ws[1].p + 4, h + 40, 32) != 0) return -4; /* wtns.q must equal zkey.r */
  uint32_t m = rd32(h + 72), npub = rd32(h + 76), n = rd32(h + 80);
  if (rd32(ws[1].p + 36) != m) return -5;
  int power = ilog2(n);
  g1_aff alpha1, beta1, delta1;
  g2_aff beta2, delta2;
  memcpy(&alpha1, h + 84, 64);
  memcpy(&beta1, h + 148, 64);
  memcpy(&beta2, h + 212, 128);
  memcpy(&delta1, h + 468, 64);
  memcpy(&delta2, h + 532, 128);
  const uint8_t* wit = ws[2].p;
  t1 = now_sec(); if (stage_sec) stage_sec[0] = t1 - t0; t0 = t1;

  /* buildABC1: single loop over the coefficient records, exactly as snarkjs runs it on its main thread */
  fe* A = (fe*)calloc((size_t)n * 3, sizeof(fe));
  fe *B = A + n, *C = B + n;
  uint32_t ncoef = rd32(zs[4].p);
  const uint8_t* cp = zs[4].p + 4;
  for (uint32_t i = 0; i < ncoef; i++) {
    const uint8_t* rec = cp + (size_t)i * 44;
    uint32_t mt = rd32(rec), c = rd32(rec + 4), sg = rd32(rec + 8);
    fe coef, w, t;
    memcpy(coef.v, rec + 12, 32);
    memcpy(w.v, wit + (size_t)sg * 32, 32);
    fr_mul(&t, &coef, &w);
    fe* dst = (mt == 0 ? A : B) + c;
    fr_add(dst, dst, &t);
  }
  for (uint32_t i = 0; i < n; i++) fr_mul(&C[i], &A[i], &B[i]);
  t1 = now_sec(); if (stage_sec) stage_sec[1] = t1 - t0; t0 = t1;

  fe inc;
  if (power == 28) { fe s25 = {{25, 0, 0, 0}}; fr_to_mont(&inc, &s25); } else fr_root(&inc, power + 1);
  for (int k = 0; k < 3; k++) {
    fe* X = A + (size_t)k * n;
    fr_fft(X, power, 1, threads);
    fr_apply_key(X, n, &inc, threads);
    fr_fft(X, power, 0, threads);
  }
  /* joinABC + batchFromMontgomery */
  fe* H = (fe*)malloc((size_t)n * sizeof(fe));
#pragma omp parallel for num_threads(threads) schedule(static)
  for (uint32_t i = 0; i < n; i++) {
    fe t;
    fr_mul(&t, &A[i], &B[i]);
    fr_sub(&t, &t, &C[i]);
    fr_from_mont(&H[i], &t);
  }
  if (h_out) memcpy(h_out, H, (size_t)n * 32);
  t1 = now_sec(); if (stage_sec) stage_sec[2] = t1 - t0; t0 = t1;

  g1_jac mA, mB1, mC, mH;
  g2_jac mB2;
  g1_multiexp(&mA, (const g1_aff*)zs[5].p, wit, m, threads);
  t1 = now_sec(); if (stage_sec) stage_sec[3] = t1 - t0; t0 = t1;
  g1_multiexp(&mB1, (const g1_aff*)zs[6].p, wit, m, threads);
  t1 = now_sec(); if (stage_sec) stage_sec[4] = t1 - t0; t0 = t1;
  g2_multiexp(&mB2, (const g2_aff*)zs[7].p, wit, m, threads);
  t1 = now_sec(); if (stage_sec) stage_sec[5] = t1 - t0; t0 = t1;
  g1_multiexp(&mC, (const g1_aff*)zs[8].p, wit + (size_t)(npub + 1) * 32, m - npub - 1, threads);
  t1 = now_sec(); if (stage_sec) stage_sec[6] = t1 - t0; t0 = t1;
  g1_multiexp(&mH, (const g1_aff*)zs[9].p, (const uint8_t*)H, n, threads);
  t1 = now_sec(); if (stage_sec) stage_sec[7] = t1 - t0; t0 = t1;
  free(H);
  free(A);
  if (parts) {
    g1_out_plain(parts, &mA);
    g1_out_plain(parts + 64, &mB1);
    g2_out_plain(parts + 128, &mB2);
    g1_out_plain(parts + 256, &mC);
    g1_out_plain(parts + 320, &mH);
  }
  /* blinding and final combination (tail of groth16Prove) */
  fe r, s, rm, sm, rs;
  memcpy(r.v, r_le, 32);
  memcpy(s.v, s_le, 32);
  fr_to_mont(&rm, &r);
  fr_to_mont(&sm, &s);
  fr_mul(&rs, &rm, &sm);
  fr_neg(&rs, &rs);
  fr_from_mont(&rs, &rs);
  g1_jac d1, pa, pb1, pc, t;
  g2_jac d2, pb, t2;
  g1_from_aff(&d1, &delta1);
  g2_from_aff(&d2, &delta2);
  g1_add_mixed(&pa, &mA, &alpha1);
  g1_mul(&t, &d1, r.v);
  g1_add(&pa, &pa, &t);
  g2_add_mixed(&pb, &mB2, &beta2);
  g2_mul(&t2, &d2, s.v);
  g2_add(&pb, &pb, &t2);
  g1_add_mixed(&pb1, &mB1, &beta1);
  g1_mul(&t, &d1, s.v);
  g1_add(&pb1, &pb1, &t);
  g1_add(&pc, &mC, &mH);
  g1_mul(&t, &pa, s.v);
  g1_add(&pc, &pc, &t);
  g1_mul(&t, &pb1, r.v);
  g1_add(&pc, &pc, &t);
  g1_mul(&t, &d1, rs.v);
  g1_add(&pc, &pc, &t);
  g1_out_plain(proof, &pa);
  g2_out_plain(proof + 64, &pb);
  g1_out_plain(proof + 192, &pc);
  return 0;
}
