// Sparse R1CS evaluation: A_T = A*w, B_T = B*w, C_T = A_T o B_T over the evaluation domain.
//
// Replaces snarkjs 0.4.12 src/groth16_prove.js buildABC1 (upstream, not vendored: package.json:12,
// yarn.lock:987-999): a single-threaded JS loop over the zkey section-4 records
//     out[m][c] += coef (x) witness[s]      (Montgomery product; coef is stored as c*R^2, witness is plain,
//                                             so the product is the Montgomery form of c*w -- SURVEY.md F7)
// followed by C_T[i] = A_T[i] (x) B_T[i].  Section 4 holds matrices A and B only.
//
// B200 design.  The COO records are converted once, at zkey load, to CSR (rows 0..n-1 = A, n..2n-1 = B), so a
// proof needs no atomics and the result does not depend on record order.  One thread owns constraint i: it walks
// row i of A and row i of B (<= 8 terms each for the NZCP shape), then writes A_T[i], B_T[i] and their product.
// Threads take the constraints of their 128-row block longest first (order[], built at zkey load), so a warp's lanes
// walk rows of similar length instead of idling behind the longest one; the stores stay inside the block's 12 KB.
// HBM-bound: 36 B per coefficient (value + column) + a 32 B gathered witness word, 96 B written per constraint.
#include "common.cuh"

namespace nzcp {

__device__ __forceinline__ Fr ld_fr(const Fr* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__device__ __forceinline__ Fr row_dot(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                      const Fr* __restrict__ val, const Fr* __restrict__ w, uint32_t row) {
  uint32_t b = row_ptr[row], e = row_ptr[row + 1];
  Fr acc = Fr::zero();
  for (uint32_t k = b; k < e; k++) acc = fp_add(acc, fp_mul(ld_fr(val + k), ld_fr(w + col[k])));
  return acc;
}

__global__ void __launch_bounds__(128)
r1cs_eval_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col, const Fr* __restrict__ val,
                 const uint32_t* __restrict__ order, const Fr* __restrict__ w, Fr* __restrict__ abc, uint32_t n) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t i = order[t];
  Fr a = row_dot(row_ptr, col, val, w, i);
  Fr b = row_dot(row_ptr, col, val, w, n + i);
  st_fr(abc + i, a);
  st_fr(abc + (size_t)n + i, b);
  st_fr(abc + 2 * (size_t)n + i, fp_mul(a, b));
}

void r1cs_eval(const R1csDevice& m, const Fr* witness, Fr* abc, cudaStream_t st) {
  r1cs_eval_kernel<<<div_up(m.n, 128), 128, 0, st>>>(m.row_ptr, m.col, m.val, m.order, witness, abc, m.n);
  NZCP_LAUNCH_CHECK();
}

}  // namespace nzcp
