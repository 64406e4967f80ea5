// RealConstellationEnv step + observation build for shapes whose per-env state does not fit one SM's shared memory
// (e.g. the reference's real_constellation_env.yaml: 324 agents x 450 tasks).  Same contract and results as
// sap_real.cu / sap_real_fast.cu (reference: real_constellation_env.py step :135-175, beta_hat :282-328,
// _build_obs :177-230), decomposed so that one environment is spread over many CTAs:
//
//   K1 sap_real_large_tot    grid (tiles, B)   float64 window sums -> L2-resident scratch in BOTH layouts
//                                              tot[b][agent][task] and totT[b][task][agent]; snapshot of k[b]
//   K2 sap_real_large_lists  grid (n/8, B)     warp per agent: exact top-M (idx asc) and top-(M+M/2) (idx desc)
//                                              lists of its row -> scratch
//   K3 sap_real_large_main   grid (n/8, B)     warp per agent: rival scores from coalesced totT rows, exact top-N,
//                                              rivals' other tasks, gather, obs / agent-input rows;
//                                              chunk 0 also does the reward phase and advances k, prev
//
// Every selection is the exact float64 warp selection (values cached in registers), so there is no certificate /
// redo logic here.  Float64 semantics as in the other kernels.
#include "sap_real.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct LargeScratch {
  double* tot;      // [B][n][m]
  double* totT;     // [B][m][n]
  uint16_t* D;      // [B][n][M]
  uint16_t* E;      // [B][n][K2]
  int32_t* ksnap;   // [B]
};

__host__ __device__ inline size_t dbl_of_bytes(size_t bytes) { return (bytes + 7) / 8; }

__host__ __device__ inline size_t large_layout(const SapEnvDims& d, LargeScratch* s, double* base) {
  const int K2 = d.M + d.M / 2;
  size_t off = 0;
  const size_t o_tot = off;  off += (size_t)d.B * d.n * d.m;
  const size_t o_totT = off; off += (size_t)d.B * d.m * d.n;
  const size_t o_D = off;    off += dbl_of_bytes(sizeof(uint16_t) * (size_t)d.B * d.n * d.M);
  const size_t o_E = off;    off += dbl_of_bytes(sizeof(uint16_t) * (size_t)d.B * d.n * K2);
  const size_t o_k = off;    off += dbl_of_bytes(sizeof(int32_t) * (size_t)d.B);
  if (s) {
    s->tot = base + o_tot;
    s->totT = base + o_totT;
    s->D = reinterpret_cast<uint16_t*>(base + o_D);
    s->E = reinterpret_cast<uint16_t*>(base + o_E);
    s->ksnap = reinterpret_cast<int32_t*>(base + o_k);
  }
  return off;
}

__device__ __forceinline__ int new_step(const RealParams& p, const int32_t* ksnap, int b) {
  return p.is_reset ? 0 : ksnap[b] + 1;
}

// ---------------------------------------------------------------------------------------------------- K1
__global__ void __launch_bounds__(kThreads) sap_real_large_tot(RealParams p) {
  __shared__ double tile[32][33];
  const SapEnvDims d = p.d;
  const int b = blockIdx.y, n = d.n, m = d.m, T = d.T, L = d.L;
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  const int k_old = p.is_reset ? -1 : p.k[b];
  if (blockIdx.x == 0 && threadIdx.x == 0) s.ksnap[b] = k_old;
  const int k_new = k_old + 1;
  if (k_new >= T || k_old >= T) return;  // done: no window
  const int Leff = min(L, T - k_new);
  const float* win = p.planes + ((d.shared_planes ? (size_t)0 : (size_t)b * T) + k_new) * n * m;
  const int tiles_j = (m + 31) / 32;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x - ti * tiles_j;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += kWarps) {
    const int i = ti * 32 + r, j = tj * 32 + tx;
    if (i < n && j < m) {
      const double pr = p.prios ? (double)p.prios[j] : 1.0;
      double sum = 0.0;
      for (int l = 0; l < Leff; ++l) sum += (double)win[((size_t)l * n + i) * m + j] * pr;  // :167-170, :190
      s.tot[((size_t)b * n + i) * m + j] = sum;
      tile[r][tx] = sum;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += kWarps) {
    const int j = tj * 32 + r, i = ti * 32 + tx;
    if (i < n && j < m) s.totT[((size_t)b * m + j) * n + i] = tile[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------------- K2
__global__ void __launch_bounds__(kThreads) sap_real_large_lists(RealParams p) {
  const SapEnvDims d = p.d;
  const int b = blockIdx.y, n = d.n, m = d.m, M = d.M, K2 = d.M + d.M / 2;
  const int lane = threadIdx.x & 31, i = blockIdx.x * kWarps + (threadIdx.x >> 5);
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  if (i >= n || new_step(p, s.ksnap, b) >= d.T || (!p.is_reset && s.ksnap[b] >= d.T)) return;
  const double* row = s.tot + ((size_t)b * n + i) * m;
  double vals[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) vals[c] = (lane + 32 * c < m) ? row[lane + 32 * c] : 0.0;
  uint16_t* Dr = s.D + ((size_t)b * n + i) * M;
  uint16_t* Er = s.E + ((size_t)b * n + i) * K2;
  warp_select_cached(m, M, false, lane, vals, [&](int r, int j) { Dr[r] = (uint16_t)j; });   // :198
  warp_select_cached(m, K2, true, lane, vals, [&](int r, int j) { Er[r] = (uint16_t)j; });   // :217 (pre-masking)
}

// ---------------------------------------------------------------------------------------------------- K3
__global__ void __launch_bounds__(kThreads) sap_real_large_main(RealParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SapEnvDims d = p.d;
  const int b = blockIdx.y;
  const int n = d.n, m = d.m, T = d.T, L = d.L, M = d.M, N = d.N, H = d.M / 2, K2 = d.M + d.M / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int obs_size = M * L + N * M * L + N * H * L + M;
  const int npairs = M + N * M + N * H;
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  const SapBatchView& vw = p.view;
  const float* env_planes = p.planes + (d.shared_planes ? (size_t)0 : (size_t)b * T * n * m);
  const int k_old = s.ksnap[b];
  if (!p.is_reset && k_old >= T) return;
  const int k_new = k_old + 1;
  const bool done = k_new >= T;
  // per-warp scratch: D of my agent, rivals, other tasks
  uint16_t* wD = reinterpret_cast<uint16_t*>(smem_raw) + (size_t)warp * (M + N + N * H);
  uint16_t* wN = wD + M;
  uint16_t* wO = wN + N;
  int32_t* cnt = reinterpret_cast<int32_t*>(smem_raw + sizeof(uint16_t) * (size_t)kWarps * (M + N + N * H) + 16);
  __shared__ double red[kWarps];

  // ------------------------------------------------------------------ chunk 0: rewards at the old window
  if (blockIdx.x == 0) {
    if (!p.is_reset) {
      for (int j = tid; j < m; j += kThreads) cnt[j] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kThreads) {
        const int a = min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
        atomicAdd(&cnt[a], 1);  // :145-147
      }
      __syncthreads();
      double local_ret = 0.0;
      for (int i = tid; i < n; i += kThreads) {
        const int a = min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
        const int pv = p.prev[(size_t)b * n + i];
        const double pr = p.prios ? (double)p.prios[a] : 1.0;
        double sum = 0.0, b0 = 0.0;
        for (int l = 0; l < L; ++l)
          if (k_old + l < T) {
            const double v = (double)env_planes[((size_t)(k_old + l) * n + i) * m + a] * pr;
            if (l == 0) b0 = v;
            sum += v;
          }
        const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);  // :304-314
        const double bh = b0 - p.lambda_ * (pen * (sum > 1e-12 ? 1.0 : 0.0));                         // :317-324
        const double r = bh > 0.0 ? bh / (double)cnt[a] : bh;                                         // :154-160
        local_ret += r;
        if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
        if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
        p.prev[(size_t)b * n + i] = a;  // :171 (other chunks take the new prev from `actions`)
        if (vw.prev_assigns.ptr)
          sap_store_int(vw.prev_assigns.ptr, sap_field_off(vw.prev_assigns, b, k_new) + i, vw.prev_assigns.dtype, a);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) local_ret += __shfl_xor_sync(SAP_FULL_MASK, local_ret, off);
      if (lane == 0) red[warp] = local_ret;
      if (vw.actions_onehot.ptr) {
        const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
        for (int i = warp; i < n; i += kWarps) {
          const int a = min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
          for (int j = lane; j < m; j += 32)
            sap_store_int(vw.actions_onehot.ptr, base + (int64_t)i * m + j, vw.actions_onehot.dtype, a == j ? 1 : 0);
        }
      }
      if (p.counts_out)
        for (int j = tid; j < m; j += kThreads) p.counts_out[(size_t)b * m + j] = cnt[j];
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kWarps; ++w) t += red[w];
        p.ep_return[b] += t;
        p.k[b] = k_new;
        if (vw.terminated.ptr)
          sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, done);
      }
    } else {
      for (int i = tid; i < n; i += kThreads) {
        p.prev[(size_t)b * n + i] = i;  // :129
        if (vw.prev_assigns.ptr)
          sap_store_int(vw.prev_assigns.ptr, sap_field_off(vw.prev_assigns, b, 0) + i, vw.prev_assigns.dtype, i);
      }
      if (tid == 0) {
        p.k[b] = 0;
        p.ep_return[b] = 0.0;
      }
    }
    if (tid == 0 && vw.filled.ptr)
      sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, k_new), vw.filled.dtype, 1);
    if (vw.avail_actions.ptr) {
      const int64_t base = sap_field_off(vw.avail_actions, b, k_new);
      for (int e = tid; e < n * m; e += kThreads) sap_store_int(vw.avail_actions.ptr, base + e, vw.avail_actions.dtype, 1);
    }
    if (vw.beta.ptr) {  // eager `beta` field (off the hot path)
      const int64_t bb = sap_field_off(vw.beta, b, k_new);
      const int Leff = done ? 0 : min(L, T - k_new);
      for (int e = tid; e < n * m; e += kThreads) {
        const double pr = p.prios ? (double)p.prios[e % m] : 1.0;
        for (int l = 0; l < L; ++l)
          sap_store_real(vw.beta.ptr, bb + (int64_t)e * L + l, vw.beta.dtype,
                         l < Leff ? (double)env_planes[((size_t)(k_new + l) * n) * m + e] * pr : 0.0);
      }
    }
  }

  // ------------------------------------------------------------------ my agent's observation row at slot k_new
  const int i = blockIdx.x * kWarps + warp;
  if (i >= n) return;
  const int64_t out = sap_field_off(vw.obs, b, k_new) + (int64_t)i * obs_size;
  float* arow = vw.agent_in.ptr ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride +
                                      (int64_t)i * vw.agent_in.t_stride
                                : nullptr;
  if (done) {  // :226-228
    for (int c = lane; c < obs_size; c += 32) {
      sap_store_real(vw.obs.ptr, out + c, vw.obs.dtype, 0.0);
      if (arow) arow[c] = 0.f;
    }
    return;
  }
  const int Leff = min(L, T - k_new);
  const float* win = env_planes + (size_t)k_new * n * m;
  const uint16_t* Dg = s.D + ((size_t)b * n + i) * M;
  for (int q = lane; q < M; q += 32) wD[q] = Dg[q];
  __syncwarp();
  // rivals: score[a] = max_q tot[a, D_i[q]], read as M coalesced rows of totT (:203-206)
  double vals[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int a = lane + 32 * c;
    double best = -INFINITY;
    if (a < n) {
      for (int q = 0; q < M; ++q) best = fmax(best, s.totT[((size_t)b * m + wD[q]) * n + a]);
      if (a == i) best = -INFINITY;
    }
    vals[c] = best;
  }
  warp_select_cached(n, N, false, lane, vals, [&](int r, int a) { wN[r] = (uint16_t)a; });
  __syncwarp();
  // rivals' other top tasks (:212-217): first M/2 entries of E[r] outside D[i], stored ascending
  for (int ps = lane; ps < N; ps += 32) {
    const uint16_t* Er = s.E + ((size_t)b * n + wN[ps]) * K2;
    int c = 0;
    for (int e = 0; e < K2 && c < H; ++e) {
      const uint16_t j = Er[e];
      bool in_top = false;
      for (int q = 0; q < M; ++q) in_top |= (wD[q] == j);
      if (!in_top) {
        wO[ps * H + (H - 1 - c)] = j;
        ++c;
      }
    }
  }
  __syncwarp();
  // gather + store (:199-225)
  for (int pp = lane; pp < npairs; pp += 32) {
    int a, j;
    if (pp < M) {
      a = i;
      j = wD[pp];
    } else if (pp < M + N * M) {
      const int x = pp - M;
      a = wN[x / M];
      j = wD[x % M];
    } else {
      const int x = pp - M - N * M;
      a = wN[x / H];
      j = wO[x];
    }
    const double pr = p.prios ? (double)p.prios[j] : 1.0;
    for (int l = 0; l < L; ++l) {
      const double v = l < Leff ? (double)win[((size_t)l * n + a) * m + j] * pr : 0.0;
      sap_store_real(vw.obs.ptr, out + (int64_t)pp * L + l, vw.obs.dtype, v);
      if (arow) arow[pp * L + l] = sap_round_real(vw.obs.dtype, v);
    }
  }
  const int pv = p.is_reset ? i : min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);  // the NEW prev_assigns
  for (int q = lane; q < M; q += 32) {
    const int j = wD[q];
    sap_store_real(vw.obs.ptr, out + (int64_t)npairs * L + q, vw.obs.dtype, j == pv ? 1.0 : 0.0);  // :222
    if (arow) arow[npairs * L + q] = j == pv ? 1.f : 0.f;
    if (p.top_out) p.top_out[((size_t)b * n + i) * M + q] = j;
  }
}

}  // namespace

int64_t sap_real_large_scratch_doubles(const SapEnvDims& d) { return (int64_t)large_layout(d, nullptr, nullptr); }

int sap_real_large_launch(RealParams& p, void* stream) {
  const SapEnvDims& d = p.d;
  SAP_REQUIRE(p.scratch, SAP_E_SMEM, "sap_real: this problem size needs scratch of sap_real_scratch_doubles() doubles");
  SAP_REQUIRE(d.n <= 512 && d.m <= 512, SAP_E_DIMS, "sap_real (large path): n, m must be <= 512");
  SAP_REQUIRE(d.B <= 65535, SAP_E_DIMS, "sap_real (large path): B must be <= 65535");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = d.M / 2;
  const dim3 g1(((d.n + 31) / 32) * ((d.m + 31) / 32), d.B), g2((d.n + kWarps - 1) / kWarps, d.B);
  sap_real_large_tot<<<g1, kThreads, 0, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_tot");
  sap_real_large_lists<<<g2, kThreads, 0, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_lists");
  const size_t smem = sizeof(uint16_t) * (size_t)kWarps * (d.M + d.N + d.N * H) + 16 + sizeof(int32_t) * (size_t)d.m;
  sap_real_large_main<<<g2, kThreads, smem, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_main");
  return SAP_OK;
}
