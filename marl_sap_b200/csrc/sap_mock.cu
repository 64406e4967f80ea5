// MockConstellationEnv step + observation build: a pure streaming kernel, one CTA per environment.
//
// Replaces /root/reference/src/envs/mock_constellation_env.py:
//   reset :94-114, step :116-162, beta_hat :228-274 (evaluated at the chosen entries),
//   get_pretransition_data :164-175, and the buffer writes of runners/episode_runner.py:71-100.
// obs_i = [onehot(a_i) (m) | S[i,:,k] | ... | S[i,:,k+L-1]], zero rows past T.  With the device
// layout planes[B,T,n,m] every S[i,:,k+l] is a contiguous row, so the obs build is row copies:
// 128-bit loads/stores when m % 4 == 0 and the obs field is fp32.
#include "sap_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct MockParams {
  SapEnvDims d;
  const float* planes;
  const float* ttrans;
  double lambda_;
  const int64_t* actions;  // step: [B,n]
  const int64_t* prev0;    // reset: [B,n]
  int32_t* k;
  int32_t* prev;
  double* ep_return;
  int32_t* counts_out;
  SapBatchView view;
  int is_reset;
  int vec4;  // m % 4 == 0, fp32 obs, 16B-aligned bases
  int ain_vec4;
};

__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// kEnvWarps warps per environment: 8 (one CTA per env) or, for tiny envs such as the reference's own 10 x 10
// configuration, 1 (eight envs per CTA, warp-level synchronisation only): at 65 536 envs of 10 x 10 the launch is
// otherwise bound by the rate at which CTAs can be issued, not by memory.
template <int kEnvWarps>
__global__ void __launch_bounds__(kThreads) sap_mock_kernel(MockParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kT = kEnvWarps * 32;         // threads per env
  constexpr int kEnvsPerCta = kThreads / kT;
  const int sub = threadIdx.x / kT;          // which env of this CTA
  const SapEnvDims d = p.d;
  const int b = blockIdx.x * kEnvsPerCta + sub, n = d.n, m = d.m, T = d.T, L = d.L;
  if (b >= d.B) return;  // whole warps only (kT is a multiple of 32)
  int32_t* cnt = reinterpret_cast<int32_t*>(smem_raw) + (size_t)sub * (m + n);  // [m]
  int32_t* act = cnt + m;                                                       // [n] clamped actions (or -1 on reset)
  __shared__ double red_all[kWarps];
  double* red = red_all + sub * kEnvWarps;
  const int tid = threadIdx.x % kT, lane = tid & 31, warp = tid >> 5;
  auto env_sync = [&]() {
    if (kEnvWarps == 1) __syncwarp();
    else __syncthreads();
  };
  const float* env_planes = p.planes + (d.shared_planes ? (size_t)0 : (size_t)b * T * n * m);
  const SapBatchView& vw = p.view;
  const int obs_size = (L + 1) * m;

  int k_new = 0;
  if (!p.is_reset) {
    const int k_old = p.k[b];
    if (k_old >= T) return;
    for (int j = tid; j < m; j += kT) cnt[j] = 0;
    env_sync();
    for (int i = tid; i < n; i += kT) {
      int a = (int)p.actions[(size_t)b * n + i];
      a = min(max(a, 0), m - 1);
      act[i] = a;
      atomicAdd(&cnt[a], 1);  // :128-130
    }
    env_sync();
    double local_ret = 0.0;
    const float* cur = env_planes + (size_t)k_old * n * m;  // beta = S[:,:,k]  (:104,:157)
    for (int i = tid; i < n; i += kT) {
      const int a = act[i];
      const int pv = p.prev[(size_t)b * n + i];
      const double bv = (double)cur[(size_t)i * m + a];
      const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);  // :250-260
      const double meaningful = bv > 1e-12 ? 1.0 : 0.0;                                            // :263
      const double bh = bv - p.lambda_ * (pen * meaningful);                                      // :266-270
      const double r = bh > 0.0 ? bh / (double)cnt[a] : bh;                                        // :132-138
      local_ret += r;
      if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
      if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
      p.prev[(size_t)b * n + i] = a;  // :160
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local_ret += __shfl_xor_sync(SAP_FULL_MASK, local_ret, off);
    if (lane == 0) red[warp] = local_ret;
    if (vw.actions_onehot.ptr) {
      const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
      for (int e = tid; e < n * m; e += kT) {
        const int i = e / m, j = e - i * m;
        sap_store_int(vw.actions_onehot.ptr, base + e, vw.actions_onehot.dtype, act[i] == j ? 1 : 0);
      }
    }
    if (p.counts_out)
      for (int j = tid; j < m; j += kT) p.counts_out[(size_t)b * m + j] = cnt[j];
    env_sync();
    k_new = k_old + 1;
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kEnvWarps; ++w) t += red[w];
      p.ep_return[b] += t;
      p.k[b] = k_new;  // :145
      if (vw.terminated.ptr)
        sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, k_new >= T);  // :154
    }
  } else {
    for (int i = tid; i < n; i += kT) {
      act[i] = -1;  // curr_assignment = 0 (:96)
      int pv = (int)p.prev0[(size_t)b * n + i];
      p.prev[(size_t)b * n + i] = min(max(pv, 0), m - 1);  // :105 (injected draw)
    }
    if (tid == 0) {
      p.k[b] = 0;
      p.ep_return[b] = 0.0;
    }
    env_sync();
  }

  // ------------------------------------------------------------------ slot t = k_new
  const int t_slot = k_new;
  if (tid == 0 && vw.filled.ptr) sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, t_slot), vw.filled.dtype, 1);
  if (vw.avail_actions.ptr) {
    const int64_t base = sap_field_off(vw.avail_actions, b, t_slot);
    for (int e = tid; e < n * m; e += kT) sap_store_int(vw.avail_actions.ptr, base + e, vw.avail_actions.dtype, 1);
  }
  if (vw.beta.ptr) {  // beta = S[:,:,k] or zeros when done (:156-159)
    const int64_t bb = sap_field_off(vw.beta, b, t_slot);
    const float* cur = env_planes + (size_t)k_new * n * m;
    for (int e = tid; e < n * m; e += kT)
      sap_store_real(vw.beta.ptr, bb + e, vw.beta.dtype, k_new < T ? (double)cur[e] : 0.0);
  }
  const int64_t obs_base = sap_field_off(vw.obs, b, t_slot);
  float* ain = vw.agent_in.ptr ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride : nullptr;
  const int64_t ain_row = vw.agent_in.t_stride;
  if (p.vec4) {
    // obs row i = (L+1) segments of m floats = (L+1)*m/4 float4; all rows of the env are contiguous.
    float* out = reinterpret_cast<float*>(vw.obs.ptr) + obs_base;
    // Warp per agent row, lane per 4 tasks: no index arithmetic beyond adds, up to 4 independent 128-bit loads in
    // flight per lane, every segment of a row leaves as one contiguous run.
    const int m4 = m >> 2, row4 = (L + 1) * m4;
    for (int i = warp; i < n; i += kEnvWarps) {
      float4* orow = reinterpret_cast<float4*>(out) + (size_t)i * row4;
      float* arow = ain ? ain + i * ain_row : nullptr;
      const int ai = act[i];
      for (int j4 = lane; j4 < m4; j4 += 32) {
        const int a = ai - (j4 << 2);
        const float4 hot = make_float4(a == 0 ? 1.f : 0.f, a == 1 ? 1.f : 0.f, a == 2 ? 1.f : 0.f, a == 3 ? 1.f : 0.f);
        stg_stream4(orow + j4, hot);  // :107, :147  curr_assignment
        if (arow) {
          float* dst = arow + (j4 << 2);
          if (p.ain_vec4) stg_stream4(reinterpret_cast<float4*>(dst), hot);
          else { dst[0] = hot.x; dst[1] = hot.y; dst[2] = hot.z; dst[3] = hot.w; }
        }
        for (int seg0 = 1; seg0 <= L; seg0 += 4) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int seg = seg0 + u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (seg <= L && k_new + seg - 1 < T)  // :108-112  zero rows past T
              v[u] = ldg_stream4(reinterpret_cast<const float4*>(env_planes + ((size_t)(k_new + seg - 1) * n + i) * m) + j4);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int seg = seg0 + u;
            if (seg <= L) {
              stg_stream4(orow + seg * m4 + j4, v[u]);
              if (arow) {
                float* dst = arow + ((seg * m4 + j4) << 2);
                if (p.ain_vec4) stg_stream4(reinterpret_cast<float4*>(dst), v[u]);
                else { dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w; }
              }
            }
          }
        }
      }
    }
  } else {
    const int total = n * obs_size;
    for (int e = tid; e < total; e += kT) {
      const int i = e / obs_size, r = e - i * obs_size;
      const int seg = r / m, j = r - seg * m;
      double v;
      if (seg == 0) v = (act[i] == j) ? 1.0 : 0.0;                                                   // :107,:147
      else v = (k_new + seg - 1 < T) ? (double)env_planes[((size_t)(k_new + seg - 1) * n + i) * m + j] : 0.0;  // :108-112
      sap_store_real(vw.obs.ptr, obs_base + e, vw.obs.dtype, v);
      if (ain) ain[i * ain_row + r] = sap_round_real(vw.obs.dtype, v);
    }
  }
}

int validate(const SapEnvDims* d, const SapBatchView* view) {
  SAP_REQUIRE(d && view, SAP_E_NULL, "sap_mock: dims/view is null");
  SAP_REQUIRE(d->B > 0 && d->n > 0 && d->m > 0 && d->T > 0 && d->L > 0, SAP_E_DIMS,
              "sap_mock: non-positive dimension (B=%d n=%d m=%d T=%d L=%d)", d->B, d->n, d->m, d->T, d->L);
  SAP_REQUIRE(view->obs.ptr, SAP_E_NULL, "sap_mock: view.obs is required");
  SAP_REQUIRE(view->obs.dtype == SAP_F32 || view->obs.dtype == SAP_F16, SAP_E_DTYPE, "sap_mock: obs must be f32|f16");
  SAP_REQUIRE((size_t)(d->m + d->n) * 4 <= 200 * 1024, SAP_E_SMEM, "sap_mock: n + m too large for shared memory");
  return SAP_OK;
}

int launch(MockParams& p, void* stream) {
  const SapEnvDims& d = p.d;
  const SapField& o = p.view.obs;
  p.vec4 = (d.m % 4 == 0) && o.dtype == SAP_F32 && sap_aligned16(o.ptr) && sap_aligned16(p.planes) &&
           (o.env_stride % 4 == 0) && (o.t_stride % 4 == 0);
  const SapField& ai = p.view.agent_in;
  p.ain_vec4 = ai.ptr && sap_aligned16(ai.ptr) && (ai.env_stride % 4 == 0) && (ai.t_stride % 4 == 0);
  SAP_REQUIRE(!ai.ptr || ai.dtype == SAP_F32, SAP_E_DTYPE, "sap_mock: agent_in must be f32");
  // tiny envs: one warp per env, eight envs per CTA
  const bool tiny = (int64_t)d.n * (d.L + 1) * d.m <= 2048 && d.B >= 64;
  size_t bytes = sizeof(int32_t) * (size_t)(d.m + d.n) * (tiny ? kWarps : 1);
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sap_mock_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      sap_set_error("sap_mock: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  if (tiny) sap_mock_kernel<1><<<(d.B + kWarps - 1) / kWarps, kThreads, bytes, (cudaStream_t)stream>>>(p);
  else sap_mock_kernel<8><<<d.B, kThreads, bytes, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_mock_kernel");
  return SAP_OK;
}

}  // namespace

extern "C" int sap_mock_reset(const SapEnvDims* dims, const float* planes, const int64_t* prev0, int32_t* k,
                              int32_t* prev, double* ep_return, const SapBatchView* view, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && prev0 && k && prev && ep_return, SAP_E_NULL, "sap_mock_reset: planes/prev0/k/prev/ep_return is null");
  MockParams p{};
  p.d = *dims;
  p.planes = planes;
  p.prev0 = prev0;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.view = *view;
  p.is_reset = 1;
  return launch(p, stream);
}

extern "C" int sap_mock_step(const SapEnvDims* dims, const float* planes, const float* T_trans, double lambda_,
                             const int64_t* actions, int32_t* k, int32_t* prev, double* ep_return, int32_t* counts_out,
                             const SapBatchView* view, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && actions && k && prev && ep_return, SAP_E_NULL,
              "sap_mock_step: planes/actions/k/prev/ep_return is null");
  MockParams p{};
  p.d = *dims;
  p.planes = planes;
  p.ttrans = T_trans;
  p.lambda_ = lambda_;
  p.actions = actions;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.counts_out = counts_out;
  p.view = *view;
  p.is_reset = 0;
  return launch(p, stream);
}
