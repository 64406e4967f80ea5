// Action selectors: masked epsilon-greedy / argmax, one warp per (env, agent) row.
//
// Replaces /root/reference/src/action_selectors/classic_selectors.py:37-54
// (EpsilonGreedyActionSelector.select_action) and filtered_classic_selectors.py:17-63
// (FilteredEpsilonGreedyActionSelector.select_action).  Random draws are either injected
// (parity with the reference under patched torch.rand_like / Categorical.sample) or Philox4x32-10.
#include "sap_select.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// first-index argmax over lanes (value desc, idx asc) on fp32 keys; idx < 0 = empty
__device__ __forceinline__ void warp_argmax_f32(float& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float ov = __shfl_xor_sync(SAP_FULL_MASK, v, off);
    int oi = __shfl_xor_sync(SAP_FULL_MASK, i, off);
    bool take = (oi >= 0) && (i < 0 || ov > v || (ov == v && oi < i));
    if (take) {
      v = ov;
      i = oi;
    }
  }
}

__global__ void __launch_bounds__(kThreads) sap_select_classic_kernel(SelParams p, int vec4) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)p.B * p.n;
  const int64_t base = ((int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5)) * 32;
  if (base >= rows) return;
  const int nrows = (int)min((int64_t)32, rows - base);
  const int my_action = sap_classic_select_rows(p, base, nrows, vec4, lane);
  if (lane < nrows) p.out[base + lane] = (int64_t)my_action;
}

template <bool kFiltered>
__global__ void __launch_bounds__(kThreads) sap_select_kernel(SelParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)p.B * p.n;
  if (row >= rows) return;
  const int b = (int)(row / p.n);
  const int A = kFiltered ? p.m : p.A;  // width of the action space
  const uint8_t* av = p.avail ? p.avail + row * A : nullptr;
  const uint32_t ep_lo = p.episode_ctr ? (uint32_t)(*p.episode_ctr) : 0u;
  const uint32_t step = p.k ? (uint32_t)p.k[b] : 0u;
  const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);

  // ---- greedy branch: first-index argmax
  float bv = 0.f;
  int bi = -1;
  int n_avail = 0;
  if (!kFiltered) {
    const float* qr = p.q + row * A;
    for (int j = lane; j < A; j += 32) {
      const bool ok = av ? av[j] != 0 : true;
      n_avail += ok;
      const float v = ok ? qr[j] : -INFINITY;  // classic_selectors.py:46-47
      if (bi < 0 || v > bv) {
        bv = v;
        bi = j;
      }
    }
  } else {
    const float* qr = p.q + row * (p.M + 1);
    const int32_t* tp = p.top + row * p.M;
    const float base = qr[p.M];  // filtered_classic_selectors.py:42
    for (int j = lane; j < A; j += 32) {
      n_avail += av ? (av[j] != 0) : 1;
      float u;
      if (p.u_tie) {
        u = p.u_tie[row * A + j];
      } else {
        SapPhilox4 r = sap_philox4x32_10((uint32_t)row, ep_lo, step, 1u + (uint32_t)(j >> 2), k0, k1);
        const uint32_t bits = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
        u = sap_u01(bits);
      }
      float v = __fadd_rn(base, __fmul_rn(u, 1e-8f));  // :46-47, no FMA contraction
      for (int qi = 0; qi < p.M; ++qi)
        if (tp[qi] == j) v = qr[qi];  // :53-54
      if (bi < 0 || v > bv) {  // no availability mask on the greedy branch of this selector
        bv = v;
        bi = j;
      }
    }
  }
  warp_argmax_f32(bv, bi);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n_avail += __shfl_xor_sync(SAP_FULL_MASK, n_avail, off);

  // ---- explore branch
  float ue, ua;
  if (p.u_explore && p.u_action) {
    ue = p.u_explore[row];
    ua = p.u_action[row];
  } else {
    SapPhilox4 r = sap_philox4x32_10((uint32_t)row, ep_lo, step, 0u, k0, k1);
    ue = sap_u01(r.x);
    ua = sap_u01(r.y);
  }
  int action = bi;
  if (ue < (p.eps_dev ? *p.eps_dev : p.eps) && n_avail > 0) {  // :49-50 ; Categorical over the 0/1 mask = uniform over available
    int rank = (int)floorf(__fmul_rn(ua, (float)n_avail));
    rank = min(rank, n_avail - 1);
    action = warp_rank_select(av, A, rank, lane);
  }
  if (lane == 0) p.out[row] = (int64_t)action;
}

// top-M task indices per (env, agent) row from a beta tensor, stable (value desc, idx asc)
template <typename TB>
__global__ void __launch_bounds__(kThreads) sap_topm_kernel(const TB* beta, int64_t rows, int m, int L, int M,
                                                            int32_t* top) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TB* br = beta + row * (int64_t)m * L;
  double lastv = 0.0;
  int lasti = -1;
  for (int r = 0; r < M; ++r) {
    double bv = 0.0;
    int bi = -1;
    for (int j = lane; j < m; j += 32) {
      double v = 0.0;
      for (int l = 0; l < L; ++l) v += (double)(float)br[(int64_t)j * L + l];
      if (lasti >= 0 && !sap_better(lastv, lasti, v, j, false)) continue;
      if (sap_better(v, j, bv, bi, false)) {
        bv = v;
        bi = j;
      }
    }
    sap_warp_argbest(bv, bi, false);
    lastv = bv;
    lasti = bi;
    if (lane == 0) top[row * M + r] = bi;
  }
}

}  // namespace

extern "C" int sap_select_epsilon_greedy(const float* q, const uint8_t* avail, int32_t B, int32_t n, int32_t A, float eps,
                                         const float* eps_dev,
                                         uint64_t seed, const uint64_t* episode_ctr, const int32_t* k,
                                         const float* u_explore, const float* u_action, int64_t* actions_out,
                                         void* stream) {
  SAP_REQUIRE(q && actions_out, SAP_E_NULL, "sap_select_epsilon_greedy: q/actions_out is null");
  SAP_REQUIRE(B > 0 && n > 0 && A > 0, SAP_E_DIMS, "sap_select_epsilon_greedy: bad dims B=%d n=%d A=%d", B, n, A);
  SAP_REQUIRE((u_explore == nullptr) == (u_action == nullptr), SAP_E_NULL,
              "sap_select_epsilon_greedy: u_explore and u_action must be given together");
  SelParams p{};
  p.q = q; p.avail = avail; p.B = B; p.n = n; p.A = A; p.eps = eps; p.eps_dev = eps_dev; p.seed = seed;
  p.episode_ctr = episode_ctr; p.k = k; p.u_explore = u_explore; p.u_action = u_action; p.out = actions_out;
  const int64_t rows = (int64_t)B * n;
  const int vec4 = (A % 4 == 0) && sap_aligned16(q) && (!avail || (reinterpret_cast<uintptr_t>(avail) & 3) == 0);
  const int64_t warps = (rows + 31) / 32;
  sap_select_classic_kernel<<<(unsigned)((warps + kWarps - 1) / kWarps), kThreads, 0, (cudaStream_t)stream>>>(p, vec4);
  SAP_CUDA_LAUNCH_CHECK("sap_select_classic_kernel");
  return SAP_OK;
}

extern "C" int sap_select_filtered_epsilon_greedy(const float* q, const int32_t* top, const uint8_t* avail, int32_t B,
                                                  int32_t n, int32_t m, int32_t M, float eps, const float* eps_dev,
                                                  uint64_t seed,
                                                  const uint64_t* episode_ctr, const int32_t* k, const float* u_tie,
                                                  const float* u_explore, const float* u_action, int64_t* actions_out,
                                                  void* stream) {
  SAP_REQUIRE(q && top && actions_out, SAP_E_NULL, "sap_select_filtered_epsilon_greedy: q/top/actions_out is null");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && M > 0 && M <= m, SAP_E_DIMS,
              "sap_select_filtered_epsilon_greedy: bad dims B=%d n=%d m=%d M=%d", B, n, m, M);
  SAP_REQUIRE((u_explore == nullptr) == (u_action == nullptr), SAP_E_NULL,
              "sap_select_filtered_epsilon_greedy: u_explore and u_action must be given together");
  SelParams p{};
  p.q = q; p.top = top; p.avail = avail; p.B = B; p.n = n; p.m = m; p.M = M; p.eps = eps; p.eps_dev = eps_dev; p.seed = seed;
  p.episode_ctr = episode_ctr; p.k = k; p.u_tie = u_tie; p.u_explore = u_explore; p.u_action = u_action;
  p.out = actions_out;
  const int64_t rows = (int64_t)B * n;
  sap_select_kernel<true><<<(unsigned)((rows + kWarps - 1) / kWarps), kThreads, 0, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_select_kernel<filtered>");
  return SAP_OK;
}

// Categorical sampling by inverse CDF, one warp per row: out = first k with cdf[k] > u * cdf[A-1], cdf accumulated in
// float64 over probs * avail.  Lane l owns the contiguous chunk [l * per, (l + 1) * per).
__global__ void __launch_bounds__(kThreads) sap_sample_categorical_kernel(const float* __restrict__ probs,
                                                                          const uint8_t* __restrict__ avail, int64_t rows,
                                                                          int A, const float* __restrict__ u,
                                                                          int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* pr = probs + row * A;
  const uint8_t* av = avail ? avail + row * A : nullptr;
  const int per = (A + 31) / 32, j0 = lane * per, j1 = min(A, j0 + per);
  double part = 0.0;
  for (int j = j0; j < j1; ++j) part += (av && !av[j]) ? 0.0 : (double)pr[j];
  double incl = part;  // inclusive scan of the lane sums
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double o = __shfl_up_sync(SAP_FULL_MASK, incl, off);
    if (lane >= off) incl += o;
  }
  const double total = __shfl_sync(SAP_FULL_MASK, incl, 31);
  const double target = (double)u[row] * total;
  const unsigned hit = __ballot_sync(SAP_FULL_MASK, incl > target);  // lanes whose chunk end passes the target
  int pick = A - 1;
  if (hit) {
    const int owner = __ffs(hit) - 1;
    if (lane == owner) {
      double c = incl - part;
      for (int j = j0; j < j1; ++j) {
        c += (av && !av[j]) ? 0.0 : (double)pr[j];
        if (c > target) {
          pick = j;
          break;
        }
      }
    }
    pick = __shfl_sync(SAP_FULL_MASK, pick, owner);
  }
  if (lane == 0) out[row] = (int64_t)pick;
}

extern "C" int sap_sample_categorical(const float* probs, const uint8_t* avail, int64_t rows, int32_t A, const float* u,
                                      int64_t* out, void* stream) {
  SAP_REQUIRE(probs && u && out, SAP_E_NULL, "sap_sample_categorical: probs/u/out is null");
  SAP_REQUIRE(rows >= 0 && A > 0, SAP_E_DIMS, "sap_sample_categorical: bad dims rows=%lld A=%d", (long long)rows, A);
  if (rows == 0) return SAP_OK;
  sap_sample_categorical_kernel<<<(unsigned)((rows + kWarps - 1) / kWarps), kThreads, 0, (cudaStream_t)stream>>>(
      probs, avail, rows, A, u, out);
  SAP_CUDA_LAUNCH_CHECK("sap_sample_categorical_kernel");
  return SAP_OK;
}

extern "C" int sap_topm_from_beta(const void* beta, int32_t dtype, int32_t B, int32_t n, int32_t m, int32_t L, int32_t M,
                                  int32_t* top_out, void* stream) {
  SAP_REQUIRE(beta && top_out, SAP_E_NULL, "sap_topm_from_beta: beta/top_out is null");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0 && L > 0 && M > 0 && M <= m, SAP_E_DIMS,
              "sap_topm_from_beta: bad dims B=%d n=%d m=%d L=%d M=%d", B, n, m, L, M);
  const int64_t rows = (int64_t)B * n;
  const unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
  if (dtype == SAP_F32) {
    sap_topm_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const float*)beta, rows, m, L, M, top_out);
  } else if (dtype == SAP_F16) {
    sap_topm_kernel<__half><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const __half*)beta, rows, m, L, M, top_out);
  } else {
    SAP_REQUIRE(false, SAP_E_DTYPE, "sap_topm_from_beta: beta must be f32|f16");
  }
  SAP_CUDA_LAUNCH_CHECK("sap_topm_kernel");
  return SAP_OK;
}
