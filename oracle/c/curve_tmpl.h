/* ORACLE (test infrastructure) -- Jacobian short-Weierstrass a = 0 group law, instantiated for G1 (F = Fq) and G2
 * (F = Fq2).  Restates wasmcurves 0.1.0 build_curve_jacobian_a0.js (g1m_/g2m_ add, addMixed, double, toAffine) and
 * ffjavascript 0.2.48 src/engine_multiexp.js (_multiExp / _multiExpChunk) + wasmcurves build_multiexp.js
 * (multiexpAffine_chunk: per-window bucket accumulation with mixed adds, running-sum reduction).
 * Include with:  #define FT, FN(x), PT, AT, CN(x)
 */
typedef struct { FT x, y; } AT;        /* affine, Montgomery, (0,0) = infinity (zkey convention) */
typedef struct { FT x, y, z; } PT;     /* Jacobian, z = 0 <=> infinity */

static inline int CN(aff_is_inf)(const AT* p) { return FN(is_zero)(&p->x) && FN(is_zero)(&p->y); }
static inline int CN(is_inf)(const PT* p) { return FN(is_zero)(&p->z); }
static inline void CN(set_inf)(PT* p) { p->x = FN(ONE); p->y = FN(ONE); memset(&p->z, 0, sizeof(FT)); }
static inline void CN(from_aff)(PT* r, const AT* a) {
  if (CN(aff_is_inf)(a)) { CN(set_inf)(r); return; }
  r->x = a->x; r->y = a->y; r->z = FN(ONE);
}

static void CN(dbl)(PT* r, const PT* p) { /* dbl-2009-l */
  if (CN(is_inf)(p)) { *r = *p; return; }
  FT A, B, C, D, E, F, t;
  FN(sqr)(&A, &p->x);
  FN(sqr)(&B, &p->y);
  FN(sqr)(&C, &B);
  FN(add)(&t, &p->x, &B);
  FN(sqr)(&t, &t);
  FN(sub)(&t, &t, &A);
  FN(sub)(&t, &t, &C);
  FN(add)(&D, &t, &t);
  FN(add)(&E, &A, &A);
  FN(add)(&E, &E, &A);
  FN(sqr)(&F, &E);
  FT z3;
  FN(mul)(&z3, &p->y, &p->z);
  FN(add)(&z3, &z3, &z3);
  FN(sub)(&t, &F, &D);
  FN(sub)(&r->x, &t, &D);
  FN(sub)(&t, &D, &r->x);
  FN(mul)(&t, &E, &t);
  FN(add)(&C, &C, &C);
  FN(add)(&C, &C, &C);
  FN(add)(&C, &C, &C);
  FN(sub)(&r->y, &t, &C);
  r->z = z3;
}

static void CN(add)(PT* r, const PT* p, const PT* q) { /* add-2007-bl without the 2x scaling */
  if (CN(is_inf)(p)) { *r = *q; return; }
  if (CN(is_inf)(q)) { *r = *p; return; }
  FT z1z1, z2z2, u1, u2, s1, s2, h, rr, hh, hhh, v, t;
  FN(sqr)(&z1z1, &p->z);
  FN(sqr)(&z2z2, &q->z);
  FN(mul)(&u1, &p->x, &z2z2);
  FN(mul)(&u2, &q->x, &z1z1);
  FN(mul)(&t, &q->z, &z2z2);
  FN(mul)(&s1, &p->y, &t);
  FN(mul)(&t, &p->z, &z1z1);
  FN(mul)(&s2, &q->y, &t);
  if (FN(eq)(&u1, &u2)) {
    if (FN(eq)(&s1, &s2)) { CN(dbl)(r, p); return; }
    CN(set_inf)(r);
    return;
  }
  FN(sub)(&h, &u2, &u1);
  FN(sub)(&rr, &s2, &s1);
  FN(sqr)(&hh, &h);
  FN(mul)(&hhh, &h, &hh);
  FN(mul)(&v, &u1, &hh);
  FT x3, y3, z3;
  FN(sqr)(&x3, &rr);
  FN(sub)(&x3, &x3, &hhh);
  FN(sub)(&x3, &x3, &v);
  FN(sub)(&x3, &x3, &v);
  FN(sub)(&t, &v, &x3);
  FN(mul)(&y3, &rr, &t);
  FN(mul)(&t, &s1, &hhh);
  FN(sub)(&y3, &y3, &t);
  FN(mul)(&z3, &p->z, &q->z);
  FN(mul)(&z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}

static void CN(add_mixed)(PT* r, const PT* p, const AT* q) { /* q affine (z = 1) */
  if (CN(aff_is_inf)(q)) { *r = *p; return; }
  if (CN(is_inf)(p)) { CN(from_aff)(r, q); return; }
  FT z1z1, u2, s2, h, rr, hh, hhh, v, t;
  FN(sqr)(&z1z1, &p->z);
  FN(mul)(&u2, &q->x, &z1z1);
  FN(mul)(&t, &p->z, &z1z1);
  FN(mul)(&s2, &q->y, &t);
  if (FN(eq)(&p->x, &u2)) {
    if (FN(eq)(&p->y, &s2)) { CN(dbl)(r, p); return; }
    CN(set_inf)(r);
    return;
  }
  FN(sub)(&h, &u2, &p->x);
  FN(sub)(&rr, &s2, &p->y);
  FN(sqr)(&hh, &h);
  FN(mul)(&hhh, &h, &hh);
  FN(mul)(&v, &p->x, &hh);
  FT x3, y3, z3;
  FN(sqr)(&x3, &rr);
  FN(sub)(&x3, &x3, &hhh);
  FN(sub)(&x3, &x3, &v);
  FN(sub)(&x3, &x3, &v);
  FN(sub)(&t, &v, &x3);
  FN(mul)(&y3, &rr, &t);
  FN(mul)(&t, &p->y, &hhh);
  FN(sub)(&y3, &y3, &t);
  FN(mul)(&z3, &p->z, &h);
  r->x = x3; r->y = y3; r->z = z3;
}

static void CN(neg)(PT* r, const PT* p) { r->x = p->x; FN(neg)(&r->y, &p->y); r->z = p->z; }

/* k * p, k = 256-bit plain little-endian (ffjavascript ec.js timesScalar: left-to-right double-and-add) */
static void CN(mul)(PT* r, const PT* p, const uint64_t* k) {
  PT acc;
  CN(set_inf)(&acc);
  for (int i = 255; i >= 0; i--) {
    CN(dbl)(&acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) CN(add)(&acc, &acc, p);
  }
  *r = acc;
}

static void CN(to_aff)(AT* r, const PT* p) {
  if (CN(is_inf)(p)) { memset(r, 0, sizeof *r); return; }
  FT zi, zi2, zi3;
  FN(inv)(&zi, &p->z);
  FN(sqr)(&zi2, &zi);
  FN(mul)(&zi3, &zi2, &zi);
  FN(mul)(&r->x, &p->x, &zi2);
  FN(mul)(&r->y, &p->y, &zi3);
}

/* ffjavascript engine_multiexp.js _multiExpChunk task: one window (bit offset `shift`, width c) over one slice of
 * points -> sum_d d * bucket[d].  wasmcurves build_multiexp.js multiexpAffine_chunk + reduceTable restated as the
 * running-sum reduction. `buckets` is caller scratch of 2^c entries. */
static void CN(multiexp_window)(PT* out, const AT* bases, const uint8_t* scalars, size_t n, int shift, int c, PT* buckets) {
  size_t nb = (size_t)1 << c;
  for (size_t d = 0; d < nb; d++) CN(set_inf)(&buckets[d]);
  for (size_t i = 0; i < n; i++) {
    const uint8_t* s = scalars + 32 * i;
    /* c-bit digit at bit offset shift of the 256-bit LE scalar */
    uint32_t d = 0;
    int byte = shift >> 3, bit = shift & 7;
    uint64_t acc = 0;
    for (int k = 0; k < 5 && byte + k < 32; k++) acc |= (uint64_t)s[byte + k] << (8 * k);
    d = (uint32_t)(acc >> bit) & (uint32_t)(nb - 1);
    if (d) CN(add_mixed)(&buckets[d], &buckets[d], &bases[i]);
  }
  PT run, acc2;
  CN(set_inf)(&run);
  CN(set_inf)(&acc2);
  for (size_t d = nb - 1; d >= 1; d--) {
    CN(add)(&run, &run, &buckets[d]);
    CN(add)(&acc2, &acc2, &run);
  }
  *out = acc2;
}
