// Shared device/host helpers for the marl_sap_b200 kernels (sm_100a).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "marl_sap_b200.h"

// ----------------------------------------------------------------------------- errors
void sap_set_error(const char* fmt, ...);

#define SAP_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      sap_set_error(__VA_ARGS__);     \
      return (code);                  \
    }                                 \
  } while (0)

#define SAP_CUDA_LAUNCH_CHECK(name)                                          \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      sap_set_error("%s: %s", (name), cudaGetErrorString(e__));             \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

static inline bool sap_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define SAP_FULL_MASK 0xffffffffu
#define SAP_NUM_SMS 148

// ----------------------------------------------------------------------------- typed stores
// Values are produced in float64 (like the reference's numpy) and rounded ONCE to the scheme dtype,
// exactly like th.tensor(np.array(v), dtype=...) at episode_buffer.py:107-108.
__device__ __forceinline__ void sap_store_real(void* base, int64_t idx, int dtype, double v) {
  if (dtype == SAP_F16) {
    reinterpret_cast<__half*>(base)[idx] = __double2half(v);
  } else {
    reinterpret_cast<float*>(base)[idx] = (float)v;  // cvt.rn.f32.f64
  }
}
// the value a buffer field of `dtype` would hold, widened back to f32 (basic_controller.py:82 `.float()`)
__device__ __forceinline__ float sap_round_real(int dtype, double v) {
  return dtype == SAP_F16 ? __half2float(__double2half(v)) : (float)v;
}
__device__ __forceinline__ void sap_store_int(void* base, int64_t idx, int dtype, int64_t v) {
  switch (dtype) {
    case SAP_I64: reinterpret_cast<int64_t*>(base)[idx] = v; break;
    case SAP_I32: reinterpret_cast<int32_t*>(base)[idx] = (int32_t)v; break;
    case SAP_I16: reinterpret_cast<int16_t*>(base)[idx] = (int16_t)v; break;
    case SAP_U8: reinterpret_cast<uint8_t*>(base)[idx] = (uint8_t)v; break;
    case SAP_F32: reinterpret_cast<float*>(base)[idx] = (float)v; break;
    case SAP_F16: reinterpret_cast<__half*>(base)[idx] = __float2half((float)v); break;
    default: break;
  }
}
__host__ __device__ __forceinline__ int sap_dtype_size(int dtype) {
  switch (dtype) {
    case SAP_F32: case SAP_I32: return 4;
    case SAP_F16: case SAP_I16: return 2;
    case SAP_I64: return 8;
    case SAP_U8: return 1;
    default: return 0;
  }
}
__device__ __forceinline__ int64_t sap_field_off(const SapField& f, int b, int t) {
  return (int64_t)b * f.env_stride + (int64_t)t * f.t_stride;
}

// ----------------------------------------------------------------------------- ordering
// Total order used by every top-k on the path (stable numpy argsort):
//   DESC_IDX_ASC : (value desc, index asc)  == np.argsort(-x, kind="stable")[:k]
//   DESC_IDX_DESC: (value desc, index desc) == reversed np.argsort(x, kind="stable")[-k:]
// Values are float64 sums of fp32 inputs -> identical to the reference's float64 numbers.
__device__ __forceinline__ bool sap_better(double va, int ia, double vb, int ib, bool idx_desc) {
  // "a ranks before b".  ia/ib < 0 marks an empty candidate.
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  if (va > vb) return true;
  if (va < vb) return false;
  return idx_desc ? (ia > ib) : (ia < ib);
}

// Warp arg-best reduction over (value, index) candidates; every lane gets the winner.
__device__ __forceinline__ void sap_warp_argbest(double& v, int& i, bool idx_desc) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double ov = __shfl_xor_sync(SAP_FULL_MASK, v, off);
    int oi = __shfl_xor_sync(SAP_FULL_MASK, i, off);
    if (sap_better(ov, oi, v, i, idx_desc)) {
      v = ov;
      i = oi;
    }
  }
}

// ----------------------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al. 2011), written out here; one call = 4 x 32 random bits.
struct SapPhilox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ SapPhilox4 sap_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return SapPhilox4{c0, c1, c2, c3};
}
// uniform in [0,1) with 24 random bits, exactly representable in fp32
__device__ __forceinline__ float sap_u01(uint32_t bits) { return (float)(bits >> 8) * (1.0f / 16777216.0f); }
