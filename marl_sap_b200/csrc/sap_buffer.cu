// Episode-buffer data movement and benefit-tensor re-layout.
//
// Replaces /root/reference/src/components/episode_buffer.py:244-271 (ReplayBuffer ring insert and the
// fancy-index copy of sample), components/transforms.py:16-19 (OneHot) and the per-step window slicing of
// the envs (real_constellation_env.py:127,167-170).  All kernels are pure HBM streaming.
#include <stdarg.h>

#include "sap_common.cuh"

// ----------------------------------------------------------------------------- last error (thread-local)
static thread_local char g_sap_error[512] = "";
void sap_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_sap_error, sizeof(g_sap_error), fmt, ap);
  va_end(ap);
}
extern "C" const char* sap_last_error(void) { return g_sap_error; }
extern "C" int sap_abi_version(void) { return SAP_ABI_VERSION; }

namespace {

constexpr int kThreads = 256;

// [B,n,m,T] -> [B,T,n,m]: per env a (n*m) x T matrix transpose through a padded shared tile.
__global__ void __launch_bounds__(kThreads) sap_ingest_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int nm, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + (size_t)b * nm * T;
  float* d = dst + (size_t)b * nm * T;
  const int x0 = blockIdx.x * 32;  // along T (src inner)
  const int y0 = blockIdx.y * 32;  // along nm
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int y = y0 + r, x = x0 + tx;
    if (y < nm && x < T) tile[r][tx] = s[(size_t)y * T + x];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int x = x0 + r, y = y0 + tx;  // dst[x][y]
    if (x < T && y < nm) d[(size_t)x * nm + y] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(kThreads) sap_rows_copy16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                                                                   int64_t row16, int64_t ring_rows, int64_t dst_row0,
                                                                   int64_t src_row0, int64_t total16) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total16; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / row16, c = e - r * row16;
    const int64_t dr = (dst_row0 + r) % ring_rows;
    dst[dr * row16 + c] = src[(src_row0 + r) * row16 + c];
  }
}
__global__ void __launch_bounds__(kThreads) sap_rows_copy1_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                                                  int64_t row_bytes, int64_t ring_rows, int64_t dst_row0,
                                                                  int64_t src_row0, int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / row_bytes, c = e - r * row_bytes;
    const int64_t dr = (dst_row0 + r) % ring_rows;
    dst[dr * row_bytes + c] = src[(src_row0 + r) * row_bytes + c];
  }
}
__global__ void __launch_bounds__(kThreads) sap_rows_gather16_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                                                                     const int64_t* __restrict__ ids, int64_t row16,
                                                                     int64_t total16) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total16; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / row16, c = e - r * row16;
    dst[e] = src[ids[r] * row16 + c];
  }
}
__global__ void __launch_bounds__(kThreads) sap_rows_gather1_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                                                    const int64_t* __restrict__ ids, int64_t row_bytes,
                                                                    int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / row_bytes, c = e - r * row_bytes;
    dst[e] = src[ids[r] * row_bytes + c];
  }
}

__device__ __forceinline__ int64_t load_int(const void* base, int64_t idx, int dtype) {
  switch (dtype) {
    case SAP_I64: return reinterpret_cast<const int64_t*>(base)[idx];
    case SAP_I32: return reinterpret_cast<const int32_t*>(base)[idx];
    case SAP_I16: return reinterpret_cast<const int16_t*>(base)[idx];
    case SAP_U8: return reinterpret_cast<const uint8_t*>(base)[idx];
    default: return 0;
  }
}

__global__ void __launch_bounds__(kThreads) sap_onehot_kernel(const void* actions, int adt, void* onehot, int odt,
                                                              int64_t rows, int m) {
  const int64_t total = rows * m;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / m;
    const int j = (int)(e - r * m);
    sap_store_int(onehot, e, odt, load_int(actions, r, adt) == j ? 1 : 0);
  }
}

// "Constellation-like" synthetic benefits, the law of generate_benefits_over_time (mock_constellation_env.py:276-299),
// written straight into the planes layout [B, T, n, m]: per (env, task) a scale drawn from {1, 1, 1, 10}; per
// (env, agent, task) active with probability 1/4, then a Gaussian bump in time with centre U(0, T) and width
// U(width_min, width_max).  Draws: Philox4x32-10 keyed by (seed; element, episode).
__global__ void __launch_bounds__(kThreads) sap_benefit_generate_kernel(float* __restrict__ planes, int B, int n, int m, int T,
                                                                        float wmin, float wmax, uint64_t seed,
                                                                        uint64_t episode) {
  const int64_t total = (int64_t)B * n * m;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const uint32_t ep_lo = (uint32_t)episode, ep_hi = (uint32_t)(episode >> 32);
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(e % m);
    const int64_t bi = e / m;
    const int i = (int)(bi % n), b = (int)(bi / n);
    const SapPhilox4 rs = sap_philox4x32_10((uint32_t)(b * m + j), ep_lo, ep_hi, 0x5CA1Eu, k0, k1);   // :281
    const float scale = (rs.x >> 30) == 3u ? 10.f : 1.f;
    const SapPhilox4 r = sap_philox4x32_10((uint32_t)e, (uint32_t)(e >> 32) ^ ep_lo, ep_hi, 0xBEEFu, k0, k1);
    const bool active = sap_u01(r.x) > 0.75f;                                                          // :284
    const float center = sap_u01(r.y) * (float)T;                                                      // :288
    const float spread = wmin + sap_u01(r.z) * (wmax - wmin);                                          // :291
    const float sigma_2 = sqrtf(spread * spread / (-8.f * logf(0.05f)));                               // :292
    const float inv = 1.f / (2.f * sigma_2);
    float* dst = planes + (((int64_t)b * T) * n + i) * m + j;
    for (int t = 0; t < T; ++t) {
      const float d = (float)t - center;
      dst[(int64_t)t * n * m] = active ? scale * expf(-d * d * inv) : 0.f;                             // :297
    }
  }
}

// x[r, c] = act(x[r, c] + bias[c]) in place, 128-bit accesses (cols % 4 == 0, x 16-byte aligned) or scalar
template <bool kVec>
__global__ void __launch_bounds__(kThreads) sap_bias_act_kernel(float* __restrict__ x, const float* __restrict__ bias,
                                                                int64_t total, int cols, int relu) {
  if (kVec) {
    const int64_t total4 = total >> 2;
    const int cols4 = cols >> 2;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
      const int c4 = (int)(e % cols4);
      float4 v = reinterpret_cast<float4*>(x)[e];
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
      v.x = __fadd_rn(v.x, b.x);
      v.y = __fadd_rn(v.y, b.y);
      v.z = __fadd_rn(v.z, b.z);
      v.w = __fadd_rn(v.w, b.w);
      if (relu) {  // torch.relu: max(x, 0) with NaN propagated
        v.x = v.x < 0.f ? 0.f : v.x;
        v.y = v.y < 0.f ? 0.f : v.y;
        v.z = v.z < 0.f ? 0.f : v.z;
        v.w = v.w < 0.f ? 0.f : v.w;
      }
      reinterpret_cast<float4*>(x)[e] = v;
    }
  } else {
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
      float v = __fadd_rn(x[e], bias[e % cols]);
      if (relu) v = v < 0.f ? 0.f : v;
      x[e] = v;
    }
  }
}

// beta field [B, t_count, n, m, L] (time steps t0 .. t0 + t_count - 1) rebuilt from planes [*, T, n, m]; batch row b reads
// plane row rows[b] (rows == null: b itself, or 0 when the planes are shared)
__global__ void __launch_bounds__(kThreads) sap_beta_window_kernel(SapEnvDims d, const float* __restrict__ planes,
                                                                   const float* __restrict__ prios,
                                                                   const int64_t* __restrict__ rows, int t0, int t_count,
                                                                   void* beta, int dtype) {
  const int64_t nm = (int64_t)d.n * d.m;
  const int64_t total = (int64_t)d.B * t_count * nm;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int64_t bt = e / nm, x = e - bt * nm;
    const int b = (int)(bt / t_count), t = t0 + (int)(bt - (int64_t)b * t_count);
    const int j = (int)(x % d.m);
    const double pr = prios ? (double)prios[j] : 1.0;
    const int64_t prow = rows ? rows[b] : (d.shared_planes ? 0 : b);
    for (int l = 0; l < d.L; ++l) {
      const double v = (t + l < d.T) ? (double)planes[(prow * d.T + t + l) * nm + x] * pr : 0.0;
      sap_store_real(beta, e * d.L + l, dtype, v);
    }
  }
}

// per-plane {min, max}: one CTA per (env, time) plane
__global__ void __launch_bounds__(kThreads) sap_plane_stats_kernel(const float* __restrict__ planes, float* __restrict__ stats,
                                                                   int nm) {
  __shared__ float smin[kThreads / 32], smax[kThreads / 32];
  const float* p = planes + (size_t)blockIdx.x * nm;
  float lo = INFINITY, hi = -INFINITY;
  for (int e = threadIdx.x; e < nm; e += kThreads) {
    const float v = p[e];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(SAP_FULL_MASK, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(SAP_FULL_MASK, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = lo;
    smax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      lo = fminf(lo, smin[w]);
      hi = fmaxf(hi, smax[w]);
    }
    stats[2 * (size_t)blockIdx.x] = lo;
    stats[2 * (size_t)blockIdx.x + 1] = hi;
  }
}

inline unsigned grid_for(int64_t work) {
  int64_t blocks = (work + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)SAP_NUM_SMS * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace

extern "C" int sap_benefit_ingest(const float* src_nmT, float* dst_Tnm, int32_t B, int32_t n, int32_t m, int32_t T,
                                  void* stream) {
  SAP_REQUIRE(src_nmT && dst_Tnm, SAP_E_NULL, "sap_benefit_ingest: src/dst is null");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && T > 0 && B <= 65535, SAP_E_DIMS, "sap_benefit_ingest: bad dims B=%d n=%d m=%d T=%d",
              B, n, m, T);
  const int nm = n * m;
  dim3 grid((T + 31) / 32, (nm + 31) / 32, B);
  SAP_REQUIRE(grid.y <= 65535, SAP_E_DIMS, "sap_benefit_ingest: n*m too large");
  sap_ingest_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(src_nmT, dst_Tnm, nm, T);
  SAP_CUDA_LAUNCH_CHECK("sap_ingest_kernel");
  return SAP_OK;
}

extern "C" int sap_benefit_stats(const float* planes_Tnm, float* stats, int32_t B, int32_t n, int32_t m, int32_t T,
                                 void* stream) {
  SAP_REQUIRE(planes_Tnm && stats, SAP_E_NULL, "sap_benefit_stats: planes/stats is null");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && T > 0 && (int64_t)B * T < 2147483647LL, SAP_E_DIMS, "sap_benefit_stats: bad dims");
  sap_plane_stats_kernel<<<(unsigned)((int64_t)B * T), kThreads, 0, (cudaStream_t)stream>>>(planes_Tnm, stats, n * m);
  SAP_CUDA_LAUNCH_CHECK("sap_plane_stats_kernel");
  return SAP_OK;
}

extern "C" int sap_benefit_upload_host(const float* src_nmT_host, float* staging_dev, float* dst_Tnm, int32_t B,
                                       int32_t n, int32_t m, int32_t T, void* stream) {
  SAP_REQUIRE(src_nmT_host && staging_dev && dst_Tnm, SAP_E_NULL, "sap_benefit_upload_host: null pointer");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && T > 0, SAP_E_DIMS, "sap_benefit_upload_host: bad dims");
  cudaError_t e = cudaMemcpyAsync(staging_dev, src_nmT_host, sizeof(float) * (size_t)B * n * m * T,
                                  cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    sap_set_error("sap_benefit_upload_host: cudaMemcpyAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return sap_benefit_ingest(staging_dev, dst_Tnm, B, n, m, T, stream);
}

extern "C" int sap_buffer_insert(void* dst, const void* src, int64_t row_bytes, int64_t ring_rows, int64_t dst_row0,
                                 int64_t src_row0, int64_t count, void* stream) {
  SAP_REQUIRE(dst && src, SAP_E_NULL, "sap_buffer_insert: dst/src is null");
  SAP_REQUIRE(row_bytes > 0 && ring_rows > 0 && count >= 0 && dst_row0 >= 0 && dst_row0 < ring_rows && src_row0 >= 0 &&
                  count <= ring_rows,
              SAP_E_DIMS, "sap_buffer_insert: bad arguments");
  if (count == 0) return SAP_OK;
  if (row_bytes % 16 == 0 && sap_aligned16(dst) && sap_aligned16(src)) {
    const int64_t row16 = row_bytes / 16, total = row16 * count;
    sap_rows_copy16_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>((uint4*)dst, (const uint4*)src, row16,
                                                                                  ring_rows, dst_row0, src_row0, total);
  } else {
    const int64_t total = row_bytes * count;
    sap_rows_copy1_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>((uint8_t*)dst, (const uint8_t*)src,
                                                                                 row_bytes, ring_rows, dst_row0, src_row0,
                                                                                 total);
  }
  SAP_CUDA_LAUNCH_CHECK("sap_rows_copy_kernel");
  return SAP_OK;
}

extern "C" int sap_buffer_gather(void* dst, const void* src, const int64_t* ids, int64_t row_bytes, int64_t count,
                                 void* stream) {
  SAP_REQUIRE(dst && src && ids, SAP_E_NULL, "sap_buffer_gather: dst/src/ids is null");
  SAP_REQUIRE(row_bytes > 0 && count >= 0, SAP_E_DIMS, "sap_buffer_gather: bad arguments");
  if (count == 0) return SAP_OK;
  if (row_bytes % 16 == 0 && sap_aligned16(dst) && sap_aligned16(src)) {
    const int64_t row16 = row_bytes / 16, total = row16 * count;
    sap_rows_gather16_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>((uint4*)dst, (const uint4*)src, ids,
                                                                                    row16, total);
  } else {
    const int64_t total = row_bytes * count;
    sap_rows_gather1_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>((uint8_t*)dst, (const uint8_t*)src, ids,
                                                                                   row_bytes, total);
  }
  SAP_CUDA_LAUNCH_CHECK("sap_rows_gather_kernel");
  return SAP_OK;
}

extern "C" int sap_onehot(const void* actions, int32_t actions_dtype, void* onehot, int32_t onehot_dtype, int64_t rows,
                          int32_t m, void* stream) {
  SAP_REQUIRE(actions && onehot, SAP_E_NULL, "sap_onehot: actions/onehot is null");
  SAP_REQUIRE(rows >= 0 && m > 0, SAP_E_DIMS, "sap_onehot: bad dims");
  SAP_REQUIRE(actions_dtype == SAP_I64 || actions_dtype == SAP_I32 || actions_dtype == SAP_I16, SAP_E_DTYPE,
              "sap_onehot: actions must be i64|i32|i16");
  if (rows == 0) return SAP_OK;
  sap_onehot_kernel<<<grid_for(rows * m), kThreads, 0, (cudaStream_t)stream>>>(actions, actions_dtype, onehot,
                                                                              onehot_dtype, rows, m);
  SAP_CUDA_LAUNCH_CHECK("sap_onehot_kernel");
  return SAP_OK;
}

extern "C" int sap_benefit_generate(float* planes_Tnm, int32_t B, int32_t n, int32_t m, int32_t T, float width_min,
                                    float width_max, uint64_t seed, uint64_t episode, void* stream) {
  SAP_REQUIRE(planes_Tnm, SAP_E_NULL, "sap_benefit_generate: planes is null");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && T > 0, SAP_E_DIMS, "sap_benefit_generate: bad dims B=%d n=%d m=%d T=%d", B, n, m, T);
  SAP_REQUIRE(width_min > 0.f && width_max >= width_min, SAP_E_CONSTRAINT, "sap_benefit_generate: bad widths");
  sap_benefit_generate_kernel<<<grid_for((int64_t)B * n * m), kThreads, 0, (cudaStream_t)stream>>>(planes_Tnm, B, n, m, T,
                                                                                                width_min, width_max, seed,
                                                                                                episode);
  SAP_CUDA_LAUNCH_CHECK("sap_benefit_generate_kernel");
  return SAP_OK;
}

// out[r, c] = act(sum_k scale^k * y[r, k * cols + c] + bias[c]): the epilogue of a split-precision layer whose weight
// was split into `terms` fp16 pieces (W = W_0 + scale W_1 + scale^2 W_2) multiplied side by side by one tensor-core GEMM
__global__ void __launch_bounds__(kThreads) sap_split_bias_act_kernel(const float* __restrict__ y, int terms, float scale,
                                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                                      int64_t rows, int cols, int relu) {
  const int cols4 = cols >> 2;
  const int64_t total4 = rows * cols4;
  const float s2 = scale * scale;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total4; e += (int64_t)gridDim.x * kThreads) {
    const int64_t r = e / cols4;
    const int c4 = (int)(e - r * cols4);
    const float4* row = reinterpret_cast<const float4*>(y + r * (int64_t)terms * cols);
    float4 v = __ldcs(row + c4);  // streamed: every element is read once
    const float4 v1 = __ldcs(row + cols4 + c4);
    v.x = __fmaf_rn(scale, v1.x, v.x);
    v.y = __fmaf_rn(scale, v1.y, v.y);
    v.z = __fmaf_rn(scale, v1.z, v.z);
    v.w = __fmaf_rn(scale, v1.w, v.w);
    if (terms == 3) {
      const float4 v2 = __ldcs(row + 2 * cols4 + c4);
      v.x = __fmaf_rn(s2, v2.x, v.x);
      v.y = __fmaf_rn(s2, v2.y, v.y);
      v.z = __fmaf_rn(s2, v2.z, v.z);
      v.w = __fmaf_rn(s2, v2.w, v.w);
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    v.x = __fadd_rn(v.x, b.x);
    v.y = __fadd_rn(v.y, b.y);
    v.z = __fadd_rn(v.z, b.z);
    v.w = __fadd_rn(v.w, b.w);
    if (relu) {
      v.x = v.x < 0.f ? 0.f : v.x;
      v.y = v.y < 0.f ? 0.f : v.y;
      v.z = v.z < 0.f ? 0.f : v.z;
      v.w = v.w < 0.f ? 0.f : v.w;
    }
    reinterpret_cast<float4*>(out)[e] = v;
  }
}

extern "C" int sap_split_bias_act(const float* y_cat, int32_t terms, float scale, const float* bias, float* out, int64_t rows,
                                  int32_t cols, int32_t relu, void* stream) {
  SAP_REQUIRE(y_cat && bias && out, SAP_E_NULL, "sap_split_bias_act: y/bias/out is null");
  SAP_REQUIRE(rows >= 0 && cols > 0 && (terms == 2 || terms == 3), SAP_E_DIMS, "sap_split_bias_act: bad dims / terms");
  SAP_REQUIRE(cols % 4 == 0 && sap_aligned16(y_cat) && sap_aligned16(bias) && sap_aligned16(out), SAP_E_CONSTRAINT,
              "sap_split_bias_act: cols must be a multiple of 4 and the pointers 16-byte aligned");
  if (rows == 0) return SAP_OK;
  sap_split_bias_act_kernel<<<grid_for(rows * (cols >> 2)), kThreads, 0, (cudaStream_t)stream>>>(y_cat, terms, scale, bias, out,
                                                                                               rows, cols, relu);
  SAP_CUDA_LAUNCH_CHECK("sap_split_bias_act_kernel");
  return SAP_OK;
}

extern "C" int sap_bias_act(float* x, const float* bias, int64_t rows, int32_t cols, int32_t relu, void* stream) {
  SAP_REQUIRE(x && bias, SAP_E_NULL, "sap_bias_act: x/bias is null");
  SAP_REQUIRE(rows >= 0 && cols > 0, SAP_E_DIMS, "sap_bias_act: bad dims");
  if (rows == 0) return SAP_OK;
  const int64_t total = rows * cols;
  const bool vec = (cols % 4 == 0) && sap_aligned16(x) && sap_aligned16(bias);
  if (vec) sap_bias_act_kernel<true><<<grid_for(total >> 2), kThreads, 0, (cudaStream_t)stream>>>(x, bias, total, cols, relu);
  else sap_bias_act_kernel<false><<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(x, bias, total, cols, relu);
  SAP_CUDA_LAUNCH_CHECK("sap_bias_act_kernel");
  return SAP_OK;
}

extern "C" int sap_real_beta_rows(const SapEnvDims* dims, const float* planes, const float* task_prios,
                                  const int64_t* plane_rows, int32_t t0, int32_t t_count, void* beta, int32_t dtype,
                                  void* stream) {
  SAP_REQUIRE(dims && planes && beta, SAP_E_NULL, "sap_real_beta_rows: null pointer");
  SAP_REQUIRE(dtype == SAP_F32 || dtype == SAP_F16, SAP_E_DTYPE, "sap_real_beta_rows: beta must be f32|f16");
  SAP_REQUIRE(dims->B > 0 && dims->n > 0 && dims->m > 0 && dims->T > 0 && dims->L > 0, SAP_E_DIMS, "sap_real_beta_rows: bad dims");
  SAP_REQUIRE(t0 >= 0 && t_count > 0 && t0 + t_count <= dims->T + 1, SAP_E_DIMS,
              "sap_real_beta_rows: time steps [%d, %d) outside [0, T + 1 = %d]", t0, t0 + t_count, dims->T + 1);
  const int64_t total = (int64_t)dims->B * t_count * dims->n * dims->m;
  sap_beta_window_kernel<<<grid_for(total), kThreads, 0, (cudaStream_t)stream>>>(*dims, planes, task_prios, plane_rows, t0,
                                                                                 t_count, beta, dtype);
  SAP_CUDA_LAUNCH_CHECK("sap_beta_window_kernel");
  return SAP_OK;
}

extern "C" int sap_real_beta_window(const SapEnvDims* dims, const float* planes, const float* task_prios, void* beta,
                                    int32_t dtype, void* stream) {
  SAP_REQUIRE(dims, SAP_E_NULL, "sap_real_beta_window: null pointer");
  return sap_real_beta_rows(dims, planes, task_prios, nullptr, 0, dims->T + 1, beta, dtype, stream);
}
