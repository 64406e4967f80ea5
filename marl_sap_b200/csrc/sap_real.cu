// RealConstellationEnv step + observation build, one CTA per environment (v1: warp-per-row selection).
//
// Replaces /root/reference/src/envs/real_constellation_env.py:
//   step :135-175, beta_hat :282-328 (evaluated at the chosen entries only), _build_obs :177-230,
//   reset :116-133, get_pretransition_data :232-244, and the buffer writes of
//   runners/episode_runner.py:71-100 (slots t and t+1 of the EpisodeBatch).
//
// Arithmetic: every benefit is an fp32 input; beta = S * prio, window sums, rewards are computed in
// float64 exactly like the reference's numpy, then rounded once to the scheme dtype.  All top-k use the
// stable total orders of sap_common.cuh, so results equal the reference's on fp32-representable inputs.
#include <stdlib.h>

#include <atomic>

#include "sap_real.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct Smem {
  double* tot;
  double* score;   // [kWarps][n]
  double* red;     // [kWarps]
  int32_t* cnt;    // [m]
  uint16_t* D;     // [n][M]   top-M tasks, (value desc, idx asc)
  uint16_t* E;     // [n][M+H] top-(M+H) tasks, (value desc, idx desc)
  uint16_t* nbr;   // [n][N]
  uint16_t* other; // [n][N][H]
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t smem_layout(const SapEnvDims& d, int ms, bool tot_in_smem, Smem* s,
                                              unsigned char* base) {
  const int H = d.M / 2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 16);
    return o;
  };
  size_t o_tot = take(tot_in_smem ? sizeof(double) * (size_t)d.n * ms : 0);
  size_t o_score = take(sizeof(double) * (size_t)kWarps * d.n);
  size_t o_red = take(sizeof(double) * kWarps);
  size_t o_cnt = take(sizeof(int32_t) * (size_t)d.m);
  size_t o_D = take(sizeof(uint16_t) * (size_t)d.n * d.M);
  size_t o_E = take(sizeof(uint16_t) * (size_t)d.n * (d.M + H));
  size_t o_nbr = take(sizeof(uint16_t) * (size_t)d.n * d.N);
  size_t o_other = take(sizeof(uint16_t) * (size_t)d.n * d.N * H);
  if (s) {
    s->tot = reinterpret_cast<double*>(base + o_tot);
    s->score = reinterpret_cast<double*>(base + o_score);
    s->red = reinterpret_cast<double*>(base + o_red);
    s->cnt = reinterpret_cast<int32_t*>(base + o_cnt);
    s->D = reinterpret_cast<uint16_t*>(base + o_D);
    s->E = reinterpret_cast<uint16_t*>(base + o_E);
    s->nbr = reinterpret_cast<uint16_t*>(base + o_nbr);
    s->other = reinterpret_cast<uint16_t*>(base + o_other);
  }
  return off;
}

__global__ void __launch_bounds__(kThreads) sap_real_kernel(RealParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SapEnvDims d = p.d;
  const int b = blockIdx.x;
  const int n = d.n, m = d.m, T = d.T, L = d.L, M = d.M, N = d.N, H = d.M / 2, K2 = d.M + d.M / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int obs_size = M * L + N * M * L + N * H * L + M;
  const int obs_row = p.obs_row > 0 ? p.obs_row : obs_size;  // power envs append N + 1 columns to every row
  Smem s;
  smem_layout(d, p.ms, true, &s, smem_raw);
  double* tot = s.tot;
  const int ms = p.ms;
  const float* env_planes = p.planes + (d.shared_planes ? (size_t)0 : (size_t)b * T * n * m);
  const SapBatchView& vw = p.view;

  int k_new = 0;
  // ------------------------------------------------------------------ step: rewards at the old window
  if (!p.is_reset) {
    const int k_old = p.k[b];
    if (k_old >= T) return;  // finished episode: nothing to do
    for (int j = tid; j < m; j += kThreads) s.cnt[j] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) {
      int a = (int)p.actions[(size_t)b * n + i];
      a = min(max(a, 0), m - 1);
      atomicAdd(&s.cnt[a], 1);  // real_constellation_env.py:145-147
    }
    __syncthreads();
    double local_ret = 0.0;
    for (int i = tid; i < n; i += kThreads) {
      int a = (int)p.actions[(size_t)b * n + i];
      a = min(max(a, 0), m - 1);
      const int pv = p.prev[(size_t)b * n + i];
      const double pr = p.prios ? (double)p.prios[a] : 1.0;
      double sum = 0.0, b0 = 0.0;
      for (int l = 0; l < L; ++l) {
        if (k_old + l < T) {
          double v = (double)env_planes[((size_t)(k_old + l) * n + i) * m + a] * pr;
          if (l == 0) b0 = v;
          sum += v;
        }
      }
      const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);  // :304-314
      const double meaningful = sum > 1e-12 ? 1.0 : 0.0;                                           // :317
      const double bh = b0 - p.lambda_ * (pen * meaningful);                                      // :320-324
      double r = bh > 0.0 ? bh / (double)s.cnt[a] : bh;                                            // :154-160
      if (p.dead && p.dead[(size_t)b * n + i]) r = 0.0;  // real_power_constellation_env.py:157-165, :351-355
      local_ret += r;
      if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
      if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
      p.prev[(size_t)b * n + i] = a;  // :171
    }
    // episode return accumulator (runners: episode_return += sum(rewards))
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local_ret += __shfl_xor_sync(SAP_FULL_MASK, local_ret, off);
    if (lane == 0) s.red[warp] = local_ret;
    if (vw.actions_onehot.ptr) {  // OneHot preprocess, transforms.py:16-19
      const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
      for (int e = tid; e < n * m; e += kThreads) {
        int i = e / m, j = e - i * m;
        int a = (int)p.actions[(size_t)b * n + i];
        a = min(max(a, 0), m - 1);
        sap_store_int(vw.actions_onehot.ptr, base + e, vw.actions_onehot.dtype, a == j ? 1 : 0);
      }
    }
    if (p.counts_out)
      for (int j = tid; j < m; j += kThreads) p.counts_out[(size_t)b * m + j] = s.cnt[j];
    __syncthreads();
    k_new = k_old + 1;
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kWarps; ++w) t += s.red[w];
      p.ep_return[b] += t;
      p.k[b] = k_new;                                                                               // :162
      if (vw.terminated.ptr)
        sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, k_new >= T);  // :164
    }
  } else {
    for (int i = tid; i < n; i += kThreads)   // :129 (power envs: a random permutation drawn by the caller)
      p.prev[(size_t)b * n + i] = p.prev0 ? (int)p.prev0[(size_t)b * n + i] : i;
    if (tid == 0) {
      p.k[b] = 0;
      p.ep_return[b] = 0.0;
    }
  }
  __syncthreads();  // new prev visible to the whole block

  // ------------------------------------------------------------------ pre-transition slot t = k_new
  const int t_slot = k_new;
  const bool done = k_new >= T;
  if (tid == 0 && vw.filled.ptr) sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, t_slot), vw.filled.dtype, 1);
  if (vw.prev_assigns.ptr) {
    const int64_t base = sap_field_off(vw.prev_assigns, b, t_slot);
    for (int i = tid; i < n; i += kThreads)
      sap_store_int(vw.prev_assigns.ptr, base + i, vw.prev_assigns.dtype, p.prev[(size_t)b * n + i]);
  }
  if (vw.avail_actions.ptr) {  // :267-273, always all ones
    const int64_t base = sap_field_off(vw.avail_actions, b, t_slot);
    for (int e = tid; e < n * m; e += kThreads) sap_store_int(vw.avail_actions.ptr, base + e, vw.avail_actions.dtype, 1);
  }
  const int64_t obs_base = sap_field_off(vw.obs, b, t_slot);
  float* ain = vw.agent_in.ptr ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride : nullptr;
  const int64_t ain_row = vw.agent_in.t_stride;
  if (done) {  // :226-228  beta := 0, obs := 0
    for (int e = tid; e < n * obs_row; e += kThreads) sap_store_real(vw.obs.ptr, obs_base + e, vw.obs.dtype, 0.0);
    if (ain)
      for (int i = warp; i < n; i += kWarps)
        for (int c = lane; c < obs_row; c += 32) ain[i * ain_row + c] = 0.f;
    if (vw.beta.ptr) {
      const int64_t bb = sap_field_off(vw.beta, b, t_slot);
      for (int e = tid; e < n * m * L; e += kThreads) sap_store_real(vw.beta.ptr, bb + e, vw.beta.dtype, 0.0);
    }
    return;
  }

  // ------------------------------------------------------------------ window sums  (:167-170, :190)
  const float* win = env_planes + (size_t)k_new * n * m;
  const int Leff = min(L, T - k_new);
  for (int e = tid; e < n * m; e += kThreads) {
    const int i = e / m, j = e - i * m;
    const double pr = p.prios ? (double)p.prios[j] : 1.0;
    double sum = 0.0;
    for (int l = 0; l < Leff; ++l) sum += (double)win[(size_t)l * n * m + e] * pr;
    tot[(size_t)i * ms + j] = sum;
    if (vw.beta.ptr) {
      const int64_t bb = sap_field_off(vw.beta, b, t_slot) + (int64_t)e * L;
      for (int l = 0; l < L; ++l)
        sap_store_real(vw.beta.ptr, bb + l, vw.beta.dtype, l < Leff ? (double)win[(size_t)l * n * m + e] * pr : 0.0);
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ per-row top lists (:198, :217)
  for (int i = warp; i < n; i += kWarps) {
    const double* row = tot + (size_t)i * ms;
    warp_select(m, M, false, lane, [&](int j) { return row[j]; },
                [&](int r, int j) { s.D[i * M + r] = (uint16_t)j; });
    warp_select(m, K2, true, lane, [&](int j) { return row[j]; },
                [&](int r, int j) { s.E[i * K2 + r] = (uint16_t)j; });
  }
  __syncthreads();

  // ------------------------------------------------------------------ rivals (:203-206)
  double* score = s.score + (size_t)warp * n;
  for (int i = warp; i < n; i += kWarps) {
    for (int a = lane; a < n; a += 32) {
      double best = -INFINITY;
      const double* row = tot + (size_t)a * ms;
      for (int q = 0; q < M; ++q) best = fmax(best, row[s.D[i * M + q]]);
      score[a] = (a == i) ? -INFINITY : best;
    }
    __syncwarp();
    warp_select(n, N, false, lane, [&](int a) { return score[a]; },
                [&](int r, int a) { s.nbr[i * N + r] = (uint16_t)a; });
    __syncwarp();
  }
  __syncthreads();

  // ------------------------------------------------------------------ rivals' other top tasks (:212-217)
  // The M/2 best tasks of row r outside D[i], under (value desc, idx desc), are the first M/2 entries
  // of E[r] that are not in D[i]; the reference lists them in ascending order, so reverse.
  for (int it = tid; it < n * N; it += kThreads) {
    const int i = it / N;
    const int r = s.nbr[it];
    int c = 0;
    for (int e = 0; e < K2 && c < H; ++e) {
      const uint16_t j = s.E[r * K2 + e];
      bool in_top = false;
      for (int q = 0; q < M; ++q) in_top |= (s.D[i * M + q] == j);
      if (!in_top) {
        s.other[(size_t)it * H + (H - 1 - c)] = j;
        ++c;
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ gather + write obs (:199-225)
  const int npairs = M + N * M + N * H;
  for (int i = warp; i < n; i += kWarps) {
    const int64_t out = obs_base + (int64_t)i * obs_row;
    if (p.nbr_out)
      for (int q = lane; q < N; q += 32) p.nbr_out[((size_t)b * n + i) * N + q] = s.nbr[i * N + q];
    for (int pp = lane; pp < npairs; pp += 32) {
      int a, j;
      if (pp < M) {
        a = i;
        j = s.D[i * M + pp];
      } else if (pp < M + N * M) {
        const int x = pp - M;
        a = s.nbr[i * N + x / M];
        j = s.D[i * M + x % M];
      } else {
        const int x = pp - M - N * M;
        a = s.nbr[i * N + x / H];
        j = s.other[(size_t)i * N * H + x];
      }
      const double pr = p.prios ? (double)p.prios[j] : 1.0;
      for (int l = 0; l < L; ++l) {
        const double v = l < Leff ? (double)win[((size_t)l * n + a) * m + j] * pr : 0.0;
        sap_store_real(vw.obs.ptr, out + (int64_t)pp * L + l, vw.obs.dtype, v);
        if (ain) ain[i * ain_row + pp * L + l] = sap_round_real(vw.obs.dtype, v);
      }
    }
    const int pv = p.prev[(size_t)b * n + i];
    for (int q = lane; q < M; q += 32) {
      const int j = s.D[i * M + q];
      sap_store_real(vw.obs.ptr, out + (int64_t)npairs * L + q, vw.obs.dtype, j == pv ? 1.0 : 0.0);  // :222
      if (ain) ain[i * ain_row + npairs * L + q] = j == pv ? 1.f : 0.f;
      if (p.top_out) p.top_out[((size_t)b * n + i) * M + q] = j;
    }
  }
}

int validate(const SapEnvDims* d, const SapBatchView* view) {
  SAP_REQUIRE(d && view, SAP_E_NULL, "sap_real: dims/view is null");
  SAP_REQUIRE(d->B > 0 && d->n > 0 && d->m > 0 && d->T > 0 && d->L > 0 && d->M > 0 && d->N > 0, SAP_E_DIMS,
              "sap_real: non-positive dimension (B=%d n=%d m=%d T=%d L=%d M=%d N=%d)", d->B, d->n, d->m, d->T, d->L,
              d->M, d->N);
  SAP_REQUIRE(d->L <= d->T, SAP_E_DIMS, "sap_real: L=%d must be <= T=%d (reference clamps L=min(L,T))", d->L, d->T);
  SAP_REQUIRE(d->M % 2 == 0, SAP_E_CONSTRAINT, "sap_real: M=%d must be even (reference -M//2 slice)", d->M);
  SAP_REQUIRE(d->m >= d->n, SAP_E_CONSTRAINT, "sap_real: need m >= n (prev_assigns = arange(n)), got m=%d n=%d", d->m,
              d->n);
  SAP_REQUIRE(d->n > d->N, SAP_E_CONSTRAINT, "sap_real: need n > N, got n=%d N=%d", d->n, d->N);
  SAP_REQUIRE(d->m >= d->M + d->M / 2, SAP_E_CONSTRAINT, "sap_real: need m >= M + M/2, got m=%d M=%d", d->m, d->M);
  SAP_REQUIRE(d->m <= 65535 && d->n <= 65535, SAP_E_DIMS, "sap_real: n, m must be < 65536");
  SAP_REQUIRE(view->obs.ptr, SAP_E_NULL, "sap_real: view.obs is required");
  SAP_REQUIRE(view->obs.dtype == SAP_F32 || view->obs.dtype == SAP_F16, SAP_E_DTYPE, "sap_real: obs must be f32|f16");
  if (view->rewards.ptr)
    SAP_REQUIRE(view->rewards.dtype == SAP_F32 || view->rewards.dtype == SAP_F16, SAP_E_DTYPE,
                "sap_real: rewards must be f32|f16");
  if (view->beta.ptr)
    SAP_REQUIRE(view->beta.dtype == SAP_F32 || view->beta.dtype == SAP_F16, SAP_E_DTYPE, "sap_real: beta must be f32|f16");
  return SAP_OK;
}

constexpr size_t kMaxSmem = 227 * 1024;

// kernel selection override (sap_real_select_kernel): 0 = automatic
std::atomic<int> g_real_path{0};

int launch(RealParams& p, void* stream) {
  const SapEnvDims& d = p.d;
  const int path = g_real_path.load(std::memory_order_relaxed);
#ifdef SAP_ABLATE
  const char* skip = getenv("SAP_DEBUG_SKIP_REDO");
  p.debug_skip_redo = skip ? atoi(skip) : 0;  // timing ablations (profiling builds only: -DSAP_ABLATE)
#else
  p.debug_skip_redo = 0;
#endif
  const bool extras = p.dead || p.nbr_out || p.prev0 || p.obs_row > 0;  // power / interference envs: generic kernel only
  if (extras) {
    p.ms = (d.m & 1) ? d.m : d.m + 1;
    SAP_REQUIRE(smem_layout(d, p.ms, true, nullptr, nullptr) <= kMaxSmem, SAP_E_CONSTRAINT,
                "sap_real: the power / interference envs run on the one-CTA-per-env generic kernel, which needs the window "
                "sums of an env in shared memory (n=%d m=%d does not fit)", d.n, d.m);
  }
  if (!extras && path != SAP_REAL_PATH_AUTO && path != SAP_REAL_PATH_FAST_GEN1 && path != SAP_REAL_PATH_FAST_RUNTIME_SHAPE)  // every other kernel widens into an fp32 agent_in
    SAP_REQUIRE(!p.view.agent_in.ptr || p.view.agent_in.dtype == SAP_F32, SAP_E_DTYPE, "sap_real: agent_in must be f32");
  if (!extras && (path == SAP_REAL_PATH_LARGE_KEYED || path == SAP_REAL_PATH_LARGE_EXACT)) {
    p.large_exact = path == SAP_REAL_PATH_LARGE_EXACT;
    return sap_real_large_launch(p, stream);
  }
  if (!extras && path != SAP_REAL_PATH_GENERIC) {
    int handled = 0;
    const int rc = sap_real_fast_try(p, stream, &handled, path == SAP_REAL_PATH_FAST_GEN1);
    if (rc != SAP_OK || handled) return rc;
  }
  SAP_REQUIRE(!p.view.agent_in.ptr || p.view.agent_in.dtype == SAP_F32, SAP_E_DTYPE, "sap_real: agent_in must be f32");
  p.ms = (d.m & 1) ? d.m : d.m + 1;
  size_t with_tot = smem_layout(d, p.ms, true, nullptr, nullptr);
  p.tot_in_smem = with_tot <= kMaxSmem;
  if (!p.tot_in_smem && !extras) return sap_real_large_launch(p, stream);  // one env over many CTAs (sap_real_large.cu)
  const size_t bytes = with_tot;
  static thread_local size_t configured = 0;
  if (bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(sap_real_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) {
      sap_set_error("sap_real: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = kMaxSmem;
  }
  sap_real_kernel<<<d.B, kThreads, bytes, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_kernel");
  return SAP_OK;
}

}  // namespace

int sap_real_path_override() { return g_real_path.load(std::memory_order_relaxed); }

extern "C" int32_t sap_real_select_kernel(int32_t which) {
  if (which < SAP_REAL_PATH_AUTO || which > SAP_REAL_PATH_FAST_RUNTIME_SHAPE) return -1;
  return g_real_path.exchange(which);
}

extern "C" int64_t sap_real_scratch_doubles(const SapEnvDims* d) {
  if (!d) return 0;
  const int path = g_real_path.load(std::memory_order_relaxed);
  if (path == SAP_REAL_PATH_LARGE_KEYED || path == SAP_REAL_PATH_LARGE_EXACT) return sap_real_large_scratch_doubles(*d);
  int ms = (d->m & 1) ? d->m : d->m + 1;
  if (smem_layout(*d, ms, true, nullptr, nullptr) <= kMaxSmem) return 0;
  return sap_real_large_scratch_doubles(*d);
}

extern "C" int sap_real_reset(const SapEnvDims* dims, const float* planes, const float* plane_stats,
                              const float* task_prios, int32_t* k, int32_t* prev, double* ep_return,
                              const SapBatchView* view, int32_t* top_out, double* scratch, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && k && prev && ep_return, SAP_E_NULL, "sap_real_reset: planes/k/prev/ep_return is null");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.prios = task_prios;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.view = *view;
  p.top_out = top_out;
  p.scratch = scratch;
  p.is_reset = 1;
  return launch(p, stream);
}

extern "C" int sap_real_reset_ex(const SapEnvDims* dims, const float* planes, const float* plane_stats,
                                 const float* task_prios, int32_t* k, int32_t* prev, double* ep_return,
                                 const SapBatchView* view, int32_t* top_out, const int64_t* prev0, int32_t* nbr_out,
                                 int32_t obs_row, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && k && prev && ep_return, SAP_E_NULL, "sap_real_reset_ex: planes/k/prev/ep_return is null");
  SAP_REQUIRE(obs_row >= 0, SAP_E_DIMS, "sap_real_reset_ex: negative obs_row");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.prios = task_prios;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.view = *view;
  p.top_out = top_out;
  p.prev0 = prev0;
  p.nbr_out = nbr_out;
  // a positive obs_row marks the call as one of the power-type envs: always the generic kernel (launch())
  p.obs_row = obs_row > 0 ? obs_row : dims->M * dims->L + dims->N * dims->M * dims->L + dims->N * (dims->M / 2) * dims->L + dims->M;
  p.is_reset = 1;
  return launch(p, stream);
}

extern "C" int sap_real_step_ex(const SapEnvDims* dims, const float* planes, const float* plane_stats,
                                const float* task_prios, const float* T_trans, double lambda_, const int64_t* actions,
                                int32_t* k, int32_t* prev, double* ep_return, int32_t* counts_out,
                                const SapBatchView* view, int32_t* top_out, const uint8_t* dead, int32_t* nbr_out,
                                int32_t obs_row, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && k && prev && ep_return && actions, SAP_E_NULL,
              "sap_real_step_ex: planes/k/prev/ep_return/actions is null");
  SAP_REQUIRE(obs_row >= 0, SAP_E_DIMS, "sap_real_step_ex: negative obs_row");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.prios = task_prios;
  p.ttrans = T_trans;
  p.lambda_ = lambda_;
  p.actions = actions;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.counts_out = counts_out;
  p.view = *view;
  p.top_out = top_out;
  p.dead = dead;
  p.nbr_out = nbr_out;
  p.obs_row = obs_row > 0 ? obs_row : dims->M * dims->L + dims->N * dims->M * dims->L + dims->N * (dims->M / 2) * dims->L + dims->M;
  p.is_reset = 0;
  return launch(p, stream);
}

extern "C" int sap_real_step(const SapEnvDims* dims, const float* planes, const float* plane_stats,
                             const float* task_prios, const float* T_trans, double lambda_, const int64_t* actions,
                             int32_t* k, int32_t* prev, double* ep_return, int32_t* counts_out,
                             const SapBatchView* view, int32_t* top_out, double* scratch, void* stream) {
  int rc = validate(dims, view);
  if (rc) return rc;
  SAP_REQUIRE(planes && k && prev && ep_return && actions, SAP_E_NULL,
              "sap_real_step: planes/k/prev/ep_return/actions is null");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.prios = task_prios;
  p.ttrans = T_trans;
  p.lambda_ = lambda_;
  p.actions = actions;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.counts_out = counts_out;
  p.view = *view;
  p.top_out = top_out;
  p.scratch = scratch;
  p.is_reset = 0;
  return launch(p, stream);
}
