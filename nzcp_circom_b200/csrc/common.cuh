// Shared host-side declarations for libnzcp_prover.so (internal; the public C ABI is include/nzcp_prover.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

#include "ec.cuh"

namespace nzcp {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define NZCP_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      char b__[512];                                                                                 \
      snprintf(b__, sizeof b__, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__,    \
               __LINE__, #expr);                                                                     \
      throw nzcp::CudaError(b__);                                                                    \
    }                                                                                                \
  } while (0)

extern std::atomic<uint64_t> g_launch_count;  // every kernel launch of this library (bench.py "gpu_launches")
#define NZCP_LAUNCH_CHECK()                \
  do {                                     \
    nzcp::g_launch_count.fetch_add(1);     \
    NZCP_CUDA(cudaGetLastError());         \
  } while (0)

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------- NTT (ntt.cu)
// Tables for one domain size n = 2^log_n (Montgomery-form Fr, device memory).
struct NttDomain {
  int log_n = 0;
  Fr* tw_fwd = nullptr;   // w^j, j < n/2, w = Fr.w[log_n]
  Fr* tw_inv = nullptr;   // w^-j
  Fr* coset_scale = nullptr;  // n^-1 * inc^bitrev(p), p < n   (iNTT scale fused with batchApplyKey)
  Fr* ninv_scale = nullptr;   // unused slot kept for standalone iNTT: single element n^-1 (device)
};
void ntt_domain_create(NttDomain* d, int log_n, cudaStream_t st);
void ntt_domain_destroy(NttDomain* d);
// H pipeline on `batch` polynomials laid out back to back (batch * n elements), in place:
//   evaluations on the 2^log_n subgroup (natural order)  ->  evaluations on the coset inc*subgroup (natural order)
void ntt_coset_pipeline(const NttDomain& d, Fr* data, int batch, cudaStream_t st);
// Standalone natural-order transforms (snarkjs Fr.fft / Fr.ifft semantics); `tmp` is an n-element scratch.
void ntt_forward(const NttDomain& d, Fr* data, Fr* tmp, cudaStream_t st);
void ntt_inverse(const NttDomain& d, Fr* data, Fr* tmp, cudaStream_t st);
// h[i] = fromMontgomery(a[i]*b[i] - c[i])
void ntt_join_abc(const Fr* a, const Fr* b, const Fr* c, Fr* h, size_t n, cudaStream_t st);

// ---------------------------------------------------------------------------------------------- R1CS (r1cs.cu)
struct R1csDevice {
  uint32_t n = 0;            // domain size (rows per matrix)
  uint64_t nnz = 0;
  uint32_t* row_ptr = nullptr;  // 2n+1 entries: rows 0..n-1 = A, n..2n-1 = B
  uint32_t* col = nullptr;      // signal index per coefficient
  Fr* val = nullptr;            // coef * R^2 mod r exactly as stored in zkey section 4
};
// abc = [A_T | B_T | C_T], each n Montgomery-form Fr.  witness: plain Fr, n_vars entries.
void r1cs_eval(const R1csDevice& m, const Fr* witness, Fr* abc, cudaStream_t st);

// ---------------------------------------------------------------------------------------------- MSM (msm.cu)
struct MsmPlan {
  size_t n_points = 0;
  int c = 0;          // window bits (signed digits)
  int n_windows = 0;
  size_t n_buckets = 0;  // per window = 2^(c-1)
  bool g2 = false;
  // device scratch
  uint32_t* counts = nullptr;    // n_windows * n_buckets (+1)
  uint32_t* offsets = nullptr;
  uint32_t* cursors = nullptr;
  uint32_t* entries = nullptr;   // n_windows * n_points
  uint32_t* task_cnt = nullptr;  // per bucket
  uint32_t* task_off = nullptr;  // per bucket (+1)
  uint2* tasks = nullptr;        // (start, len)
  void* partial = nullptr;       // XYZZ per task
  void* buckets = nullptr;       // XYZZ per bucket
  void* lvl_a[2] = {nullptr, nullptr};  // ping-pong level arrays (A sums)
  void* lvl_r[2] = {nullptr, nullptr};  // ping-pong level arrays (R weighted sums)
  uint32_t* heavy_list = nullptr;
  uint32_t* flags = nullptr;     // [0] heavy count, [1] error flag (scalar >= r), [2] total tasks
  void* window_out = nullptr;    // device: n_windows XYZZ
  void* window_host = nullptr;   // pinned host mirror
  size_t max_tasks = 0;
  size_t scratch_bytes = 0;
};
int msm_pick_window(size_t n_points);
void msm_plan_create(MsmPlan* p, size_t n_points, bool g2, int c_override);
void msm_plan_destroy(MsmPlan* p);
// Launch all kernels of one MSM on `st`; leaves per-window sums in p->window_host after the stream drains.
void msm_launch(MsmPlan* p, const void* bases, const Fr* scalars, size_t n_points, cudaStream_t st);
// Host: Horner over the window sums (after stream sync).  Throws if the device flagged a scalar >= r.
G1XYZZ msm_finish_g1(const MsmPlan* p);
G2XYZZ msm_finish_g2(const MsmPlan* p);

}  // namespace nzcp
