// Multi-scalar multiplication sum_i s_i * P_i over BN254 G1 and G2 -- signed-digit Pippenger, bucket method.
//
// Replaces (upstream, not vendored: yarn.lock:408-416, 1132-1135) ffjavascript src/engine_multiexp.js
// (G1.multiExpAffine / G2.multiExpAffine: window size pTSizes[log2 N], one _multiExpChunk task per window and point
// chunk on the worker pool, then Horner by doubling) and wasmcurves build_multiexp.js g1m_multiexpAffine_chunk.
// Inputs follow the same contract: bases are affine, Montgomery-form, (0,0) = infinity (zkey sections 5-9);
// scalars are 32-byte little-endian *plain* integers < r (the witness, or the from-Montgomery h vector).
// The sum is an exact group element, so its affine form is bit-identical to the reference's.
//
// B200 design (one MSM = one stream-ordered chain of launches, no host round trip until the end):
//   1 digits/count   signed c-bit digits (c = 16 at 2^20: 16 windows x 2^15 buckets); histogram with REDG atomics
//   2 scan           exclusive prefix sum of the 2^19 bucket counts (single block)
//   3 scatter        point index | sign<<31 written into its bucket's slot range (sorted-by-bucket entry list)
//   4 tasks          every bucket is cut into tasks of <= kTaskLen entries so that the 0/1-heavy witness
//                    distribution (27 % of all points land in one bucket) still load-balances
//   5 accumulate     one thread per task: XYZZ accumulator in registers, 8M+2S mixed adds, bases gathered by index
//                    (64 B / 128 B per point, L2-resident at these sizes)
//   6 combine        sum the tasks of a bucket (a block-wide tree for the few heavy buckets)
//   7 reduce         sum_v v * B_v per window by a radix-8 tree of running sums (no scalar multiplications)
//   8 host           Horner over the <= 64 window sums (O(1) group operations, as snarkjs does on its main thread)
#include "common.cuh"

namespace nzcp {

static constexpr int kTaskLen = 64;       // max entries per accumulate task
static constexpr int kLightTasks = 2;     // buckets with more tasks than this go to the block-wide combine
static constexpr int kHeavyThreads = 128;
static constexpr int kGroupLog = 3;       // radix of the bucket-reduction tree

struct DigitParams {
  uint32_t n_points;
  int c;
  int n_windows;
  uint32_t n_buckets;  // per window
};

__device__ __forceinline__ uint32_t scalar_bits(const uint32_t* s, int pos, int c) {
  if (pos >= 256) return 0;
  int idx = pos >> 5, sh = pos & 31;
  uint64_t lo = s[idx];
  uint64_t hi = (idx + 1 < 8) ? s[idx + 1] : 0;
  uint64_t v = (lo | (hi << 32)) >> sh;
  return (uint32_t)v & ((1u << c) - 1);
}

// COUNT pass: histogram.  SCATTER pass: cursors start at the bucket offsets; write the entry list.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
msm_digits_kernel(const Fr* __restrict__ scalars, DigitParams p, uint32_t* __restrict__ counts_or_cursors,
                  uint32_t* __restrict__ entries, uint32_t* __restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_points) return;
  uint32_t s[8];
  {
    const uint4* q = reinterpret_cast<const uint4*>(scalars + i);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
  }
  if (!SCATTER) {
    // scalars must be canonical (< r); anything else is a malformed witness
    bool ge = true;
#pragma unroll
    for (int k = 7; k >= 0; k--) {
      uint32_t m = FrParams::mod(k);
      if (s[k] != m) {
        ge = s[k] > m;
        break;
      }
    }
    if (ge) {
      atomicOr(&flags[1], 1u);
      return;
    }
  }
  if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;
  uint32_t carry = 0;
  for (int w = 0; w < p.n_windows; w++) {
    uint32_t d = scalar_bits(s, w * p.c, p.c) + carry;
    uint32_t neg = 0;
    if (d > p.n_buckets) {
      d = (1u << p.c) - d;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    if (d) {
      uint32_t b = (uint32_t)w * p.n_buckets + (d - 1);
      if (SCATTER) {
        uint32_t pos = atomicAdd(&counts_or_cursors[b], 1u);
        entries[pos] = i | (neg << 31);
      } else {
        atomicAdd(&counts_or_cursors[b], 1u);
      }
    }
  }
  if (!SCATTER && carry) atomicOr(&flags[1], 2u);
}

// Single-block exclusive scan: out[i] = sum_{j<i} f(in[j]), out[n] = total.  f = identity or ceil(x / kTaskLen).
template <bool TASKS>
__global__ void __launch_bounds__(1024)
msm_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ total) {
  __shared__ uint32_t sums[1024];
  uint32_t per = (n + 1023) / 1024;
  uint32_t b = threadIdx.x * per;
  uint32_t e = b + per < n ? b + per : n;
  uint32_t acc = 0;
  for (uint32_t i = b; i < e; i++) {
    uint32_t v = in[i];
    acc += TASKS ? (v + kTaskLen - 1) / kTaskLen : v;
  }
  sums[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = threadIdx.x >= (uint32_t)off ? sums[threadIdx.x - off] : 0;
    __syncthreads();
    sums[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? sums[threadIdx.x - 1] : 0;
  for (uint32_t i = b; i < e; i++) {
    out[i] = run;
    uint32_t v = in[i];
    run += TASKS ? (v + kTaskLen - 1) / kTaskLen : v;
  }
  if (threadIdx.x == 1023) {
    out[n] = sums[1023];
    if (total) *total = sums[1023];
  }
}

__global__ void __launch_bounds__(256)
msm_task_fill_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                     const uint32_t* __restrict__ task_off, uint2* __restrict__ tasks, uint32_t total_buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= total_buckets) return;
  uint32_t cnt = counts[b], off = offsets[b], t = task_off[b];
  for (uint32_t done = 0; done < cnt; done += kTaskLen, t++) {
    uint32_t len = cnt - done < (uint32_t)kTaskLen ? cnt - done : (uint32_t)kTaskLen;
    tasks[t] = make_uint2(off + done, len);
  }
}

template <class F>
__global__ void __launch_bounds__(128)
msm_accumulate_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ entries,
                      const uint2* __restrict__ tasks, const uint32_t* __restrict__ flags, XYZZ<F>* __restrict__ partial) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= flags[2]) return;
  uint2 tk = tasks[t];
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t k = 0; k < tk.y; k++) {
    uint32_t e = entries[tk.x + k];
    Affine<F> q = bases[e & 0x7fffffffu];
    if (q.is_inf()) continue;
    xyzz_madd(acc, q, (e >> 31) != 0);
  }
  partial[t] = acc;
}

template <class F>
__global__ void __launch_bounds__(128)
msm_combine_kernel(const uint32_t* __restrict__ task_off, const XYZZ<F>* __restrict__ partial,
                   XYZZ<F>* __restrict__ buckets, uint32_t total_buckets, uint32_t* __restrict__ heavy_list,
                   uint32_t* __restrict__ flags) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= total_buckets) return;
  uint32_t t0 = task_off[b], nt = task_off[b + 1] - t0;
  if (nt > (uint32_t)kLightTasks) {
    heavy_list[atomicAdd(&flags[0], 1u)] = b;
    return;
  }
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t j = 0; j < nt; j++) xyzz_add(acc, partial[t0 + j]);
  buckets[b] = acc;
}

template <class F>
__global__ void __launch_bounds__(kHeavyThreads)
msm_combine_heavy_kernel(const uint32_t* __restrict__ task_off, const XYZZ<F>* __restrict__ partial,
                         XYZZ<F>* __restrict__ buckets, const uint32_t* __restrict__ heavy_list,
                         const uint32_t* __restrict__ flags) {
  extern __shared__ unsigned char heavy_sm_raw[];
  XYZZ<F>* sm = reinterpret_cast<XYZZ<F>*>(heavy_sm_raw);
  uint32_t n_heavy = flags[0];
  for (uint32_t h = blockIdx.x; h < n_heavy; h += gridDim.x) {
    uint32_t b = heavy_list[h];
    uint32_t t0 = task_off[b], nt = task_off[b + 1] - t0;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t j = threadIdx.x; j < nt; j += blockDim.x) xyzz_add(acc, partial[t0 + j]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t stride = kHeavyThreads / 2; stride >= 1; stride >>= 1) {
      if (threadIdx.x < stride) {
        XYZZ<F> a = sm[threadIdx.x];
        xyzz_add(a, sm[threadIdx.x + stride]);
        sm[threadIdx.x] = a;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) buckets[b] = sm[0];
    __syncthreads();
  }
}

// One level of the bucket-reduction tree.  Input "buckets" P[0..G) of a group carry weights 0..G-1:
//   A = sum_j P[j],   S = sum_j j * P[j]   (running sums from the top),
//   R = sum_j Rin[j] + 2^shift * S         (Rin = weighted sums of the children's own subtrees; absent at level 1)
// so that at the top  sum_v v * B_v = R_top + A_top  (bucket id v-1 carries weight v-1, plus one A_top).
template <class F>
__global__ void __launch_bounds__(128)
msm_reduce_level_kernel(const XYZZ<F>* __restrict__ in_a, const XYZZ<F>* __restrict__ in_r, XYZZ<F>* __restrict__ out_a,
                        XYZZ<F>* __restrict__ out_r, uint32_t n_groups, int g_log, int shift) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const uint32_t G = 1u << g_log;
  const size_t base = (size_t)g << g_log;
  XYZZ<F> run = XYZZ<F>::inf(), S = XYZZ<F>::inf();
  for (uint32_t j = G - 1; j >= 1; j--) {
    xyzz_add(run, in_a[base + j]);
    xyzz_add(S, run);
  }
  xyzz_add(run, in_a[base]);
  for (int k = 0; k < shift; k++) S = xyzz_dbl(S);
  if (in_r) {
    for (uint32_t j = 0; j < G; j++) xyzz_add(S, in_r[base + j]);
  }
  out_a[g] = run;
  out_r[g] = S;
}

// ---------------------------------------------------------------------------------------------- host side
int msm_pick_window(size_t n) {
  int lg = 0;
  while (((size_t)2 << lg) <= n) lg++;
  if (lg <= 6) return 4;
  if (lg <= 9) return 6;
  if (lg <= 12) return 8;
  if (lg <= 14) return 10;
  if (lg <= 16) return 12;
  if (lg <= 18) return 14;
  return 16;
}

template <class T>
static T* dev_alloc(size_t count, size_t* total) {
  T* p = nullptr;
  size_t bytes = (count ? count : 1) * sizeof(T);
  NZCP_CUDA(cudaMalloc(&p, bytes));
  *total += bytes;
  return p;
}

void msm_plan_create(MsmPlan* p, size_t n_points, bool g2, int c_override) {
  *p = MsmPlan();
  p->n_points = n_points;
  p->g2 = g2;
  p->c = c_override > 0 ? c_override : msm_pick_window(n_points ? n_points : 1);
  if (p->c < 2 || p->c > 20) throw std::runtime_error("msm: window size out of range");
  p->n_windows = (254 + p->c - 1) / p->c;
  if (p->n_windows * p->c < 255) p->n_windows++;
  p->n_buckets = (size_t)1 << (p->c - 1);
  size_t tb = p->n_buckets * p->n_windows;
  size_t max_entries = (size_t)p->n_windows * n_points;
  if (max_entries >= ((size_t)1 << 32) || n_points >= ((size_t)1 << 31)) throw std::runtime_error("msm: too many points");
  p->max_tasks = tb + max_entries / kTaskLen + 1;
  size_t psz = g2 ? sizeof(G2XYZZ) : sizeof(G1XYZZ);
  size_t tot = 0;
  p->counts = dev_alloc<uint32_t>(tb + 1, &tot);
  p->offsets = dev_alloc<uint32_t>(tb + 1, &tot);
  p->cursors = dev_alloc<uint32_t>(tb + 1, &tot);
  p->entries = dev_alloc<uint32_t>(max_entries, &tot);
  p->task_off = dev_alloc<uint32_t>(tb + 1, &tot);
  p->tasks = dev_alloc<uint2>(p->max_tasks, &tot);
  p->partial = dev_alloc<unsigned char>(p->max_tasks * psz, &tot);
  p->buckets = dev_alloc<unsigned char>(tb * psz, &tot);
  size_t lvl = (tb >> kGroupLog) + p->n_windows;
  for (int i = 0; i < 2; i++) {
    p->lvl_a[i] = dev_alloc<unsigned char>(lvl * psz, &tot);
    p->lvl_r[i] = dev_alloc<unsigned char>(lvl * psz, &tot);
  }
  p->heavy_list = dev_alloc<uint32_t>(tb, &tot);
  p->flags = dev_alloc<uint32_t>(8, &tot);
  p->window_out = dev_alloc<unsigned char>(2 * p->n_windows * psz, &tot);
  NZCP_CUDA(cudaMallocHost(&p->window_host, 2 * p->n_windows * psz + 64));
  p->scratch_bytes = tot;
}

void msm_plan_destroy(MsmPlan* p) {
  cudaFree(p->counts);
  cudaFree(p->offsets);
  cudaFree(p->cursors);
  cudaFree(p->entries);
  cudaFree(p->task_off);
  cudaFree(p->tasks);
  cudaFree(p->partial);
  cudaFree(p->buckets);
  for (int i = 0; i < 2; i++) {
    cudaFree(p->lvl_a[i]);
    cudaFree(p->lvl_r[i]);
  }
  cudaFree(p->heavy_list);
  cudaFree(p->flags);
  cudaFree(p->window_out);
  if (p->window_host) cudaFreeHost(p->window_host);
  *p = MsmPlan();
}

template <class F>
static void msm_launch_t(MsmPlan* p, const Affine<F>* bases, const Fr* scalars, size_t n_points, cudaStream_t st) {
  if (n_points > p->n_points) throw std::runtime_error("msm: plan too small");
  const size_t tb = p->n_buckets * p->n_windows;
  const size_t psz = sizeof(XYZZ<F>);
  uint32_t* host_flags = reinterpret_cast<uint32_t*>((unsigned char*)p->window_host + 2 * p->n_windows * psz);
  NZCP_CUDA(cudaMemsetAsync(p->counts, 0, (tb + 1) * sizeof(uint32_t), st));
  NZCP_CUDA(cudaMemsetAsync(p->flags, 0, 8 * sizeof(uint32_t), st));
  XYZZ<F>* wout = reinterpret_cast<XYZZ<F>*>(p->window_out);
  if (n_points == 0) {
    NZCP_CUDA(cudaMemsetAsync(p->window_out, 0, 2 * p->n_windows * psz, st));
  } else {
    DigitParams dp{(uint32_t)n_points, p->c, p->n_windows, (uint32_t)p->n_buckets};
    unsigned gp = div_up(n_points, 256);
    msm_digits_kernel<false><<<gp, 256, 0, st>>>(scalars, dp, p->counts, nullptr, p->flags);
    NZCP_LAUNCH_CHECK();
    msm_scan_kernel<false><<<1, 1024, 0, st>>>(p->counts, p->offsets, (uint32_t)tb, nullptr);
    NZCP_LAUNCH_CHECK();
    NZCP_CUDA(cudaMemcpyAsync(p->cursors, p->offsets, (tb + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    msm_digits_kernel<true><<<gp, 256, 0, st>>>(scalars, dp, p->cursors, p->entries, p->flags);
    NZCP_LAUNCH_CHECK();
    msm_scan_kernel<true><<<1, 1024, 0, st>>>(p->counts, p->task_off, (uint32_t)tb, p->flags + 2);
    NZCP_LAUNCH_CHECK();
    msm_task_fill_kernel<<<div_up(tb, 256), 256, 0, st>>>(p->counts, p->offsets, p->task_off, p->tasks, (uint32_t)tb);
    NZCP_LAUNCH_CHECK();
    XYZZ<F>* partial = reinterpret_cast<XYZZ<F>*>(p->partial);
    XYZZ<F>* buckets = reinterpret_cast<XYZZ<F>*>(p->buckets);
    msm_accumulate_kernel<F><<<div_up(p->max_tasks, 128), 128, 0, st>>>(bases, p->entries, p->tasks, p->flags, partial);
    NZCP_LAUNCH_CHECK();
    msm_combine_kernel<F><<<div_up(tb, 128), 128, 0, st>>>(p->task_off, partial, buckets, (uint32_t)tb, p->heavy_list,
                                                         p->flags);
    NZCP_LAUNCH_CHECK();
    msm_combine_heavy_kernel<F><<<296, kHeavyThreads, kHeavyThreads * psz, st>>>(p->task_off, partial, buckets,
                                                                               p->heavy_list, p->flags);
    NZCP_LAUNCH_CHECK();
    // reduction tree over the c-1 bucket-index bits
    int bits_left = p->c - 1, shift = 0, level = 0;
    const XYZZ<F>* in_a = buckets;
    const XYZZ<F>* in_r = nullptr;
    size_t groups_in = tb;
    while (bits_left > 0) {
      int g_log = bits_left < kGroupLog ? bits_left : kGroupLog;
      size_t n_groups = groups_in >> g_log;
      bool last = (bits_left == g_log);
      XYZZ<F>* out_a = last ? wout : reinterpret_cast<XYZZ<F>*>(p->lvl_a[level & 1]);
      XYZZ<F>* out_r = last ? wout + p->n_windows : reinterpret_cast<XYZZ<F>*>(p->lvl_r[level & 1]);
      msm_reduce_level_kernel<F><<<div_up(n_groups, 128), 128, 0, st>>>(in_a, in_r, out_a, out_r, (uint32_t)n_groups,
                                                                       g_log, shift);
      NZCP_LAUNCH_CHECK();
      in_a = out_a;
      in_r = out_r;
      groups_in = n_groups;
      shift += g_log;
      bits_left -= g_log;
      level++;
    }
  }
  NZCP_CUDA(cudaMemcpyAsync(p->window_host, p->window_out, 2 * p->n_windows * psz, cudaMemcpyDeviceToHost, st));
  NZCP_CUDA(cudaMemcpyAsync(host_flags, p->flags, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
}

void msm_launch(MsmPlan* p, const void* bases, const Fr* scalars, size_t n_points, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    NZCP_CUDA(cudaFuncSetAttribute(msm_combine_heavy_kernel<Fq2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)(kHeavyThreads * sizeof(G2XYZZ))));
    attr_done = true;
  }
  if (p->g2)
    msm_launch_t<Fq2>(p, reinterpret_cast<const G2Affine*>(bases), scalars, n_points, st);
  else
    msm_launch_t<Fq>(p, reinterpret_cast<const G1Affine*>(bases), scalars, n_points, st);
}

template <class F>
static XYZZ<F> msm_finish_t(const MsmPlan* p) {
  const size_t psz = sizeof(XYZZ<F>);
  const XYZZ<F>* w = reinterpret_cast<const XYZZ<F>*>(p->window_host);
  const uint32_t* host_flags = reinterpret_cast<const uint32_t*>((const unsigned char*)p->window_host + 2 * p->n_windows * psz);
  if (host_flags[1]) throw std::runtime_error("witness/scalar value is not a canonical field element (>= r)");
  XYZZ<F> acc = XYZZ<F>::inf();
  for (int i = p->n_windows - 1; i >= 0; i--) {
    if (!acc.is_inf())
      for (int k = 0; k < p->c; k++) acc = xyzz_dbl(acc);
    XYZZ<F> t = w[i];
    xyzz_add(t, w[p->n_windows + i]);
    xyzz_add(acc, t);
  }
  return acc;
}

G1XYZZ msm_finish_g1(const MsmPlan* p) { return msm_finish_t<Fq>(p); }
G2XYZZ msm_finish_g2(const MsmPlan* p) { return msm_finish_t<Fq2>(p); }

}  // namespace nzcp
