// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/nzcp_prover.h"
#include "common.cuh"

namespace nzcp {

struct ApiError : std::runtime_error {
  int code;
  ApiError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

void set_last_error(const std::string& s);

// BN254 moduli as 32-byte little-endian strings (zkey / wtns / r1cs headers)
static const uint8_t kRBytes[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
static const uint8_t kQBytes[32] = {0x47, 0xfd, 0x7c, 0xd8, 0x16, 0x8c, 0x20, 0x3c, 0x8d, 0xca, 0x71, 0x68, 0x91, 0x6a, 0x81, 0x97,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};


template <class Fn>
static int api_guard(Fn&& fn) {
  try {
    fn();
    return NZCP_OK;
  } catch (const ApiError& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const CudaError& e) {
    set_last_error(e.what());
    return NZCP_E_CUDA;
  } catch (const std::bad_alloc&) {
    set_last_error("out of host memory");
    return NZCP_E_INTERNAL;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return NZCP_E_INTERNAL;
  }
}

static inline void use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    throw ApiError(NZCP_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                    "); libnzcp_prover has no CPU fallback");
  }
  if (device < 0 || device >= n) throw ApiError(NZCP_E_ARG, "device index out of range");
  NZCP_CUDA(cudaSetDevice(device));
}

// plain LE bytes <-> field elements (host)
template <class P>
static inline Fp<P> fp_from_bytes_plain(const uint8_t* b) {
  Fp<P> x;
  memcpy(x.v, b, 32);
  return x;
}
template <class P>
static inline void fp_to_bytes(const Fp<P>& x, uint8_t* b) {
  memcpy(b, x.v, 32);
}
static inline void g1_to_plain_bytes(const G1XYZZ& p, uint8_t out[64]) {
  G1Affine a = xyzz_to_affine(p);
  fp_to_bytes(fp_from_mont(a.x), out);
  fp_to_bytes(fp_from_mont(a.y), out + 32);
}
static inline void g2_to_plain_bytes(const G2XYZZ& p, uint8_t out[128]) {
  G2Affine a = xyzz_to_affine(p);
  fp_to_bytes(fp_from_mont(a.x.c0), out);
  fp_to_bytes(fp_from_mont(a.x.c1), out + 32);
  fp_to_bytes(fp_from_mont(a.y.c0), out + 64);
  fp_to_bytes(fp_from_mont(a.y.c1), out + 96);
}
static inline bool fr_bytes_canonical(const uint8_t* b) {
  Fr x = fp_from_bytes_plain<FrParams>(b);
  return !fp_geq_mod<FrParams>(x.v);
}

}  // namespace nzcp
