// Batched linear-sum-assignment (maximise), one WARP per environment.
//
// Replaces the per-env loop of /root/reference/src/action_selectors/sap_selectors.py:52-97
// (SequentialAssignmentProblemSelector.select_action): for every env of the batch
//     benefit = Q[b] + normal(0, std_b)                        (:84-86, std_b = mean|Q[b]| * eps * 2)
//     _, col_ind = scipy.optimize.linear_sum_assignment(benefit, maximize=True)     (:88)
// and of EpsilonGreedySAPTestActionSelector's test branch (:26-33, no noise).
//
// scipy's solver (scipy 1.x, `rectangular_lsap.cpp`, pinned nowhere by the reference) is the shortest-augmenting-path
// algorithm of Crouse, "On implementing 2D rectangular assignment algorithms" (IEEE TAES 2016).  This kernel restates
// that published algorithm on float64 duals, so it returns AN optimal assignment; when the optimum is unique (no exact
// ties between assignments - always the case with the Gaussian perturbation) it is the assignment scipy returns.
//
// Layout: lane l owns columns l, l + 32, ... (kCM register slots): dual v[j], shortest-path cost, predecessor row and
// row4col[j] live in registers; per-row state (dual u, col4row, the visited-row list) lives in shared memory.  One
// Dijkstra step = one coalesced read of a benefit row (L2), kCM relaxations per lane, and three redux.sync reductions
// (value high word, value low word, tie key) for the arg-min.  No block-level barrier anywhere.
#include "sap_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct LsaParams {
  const float* q;    // [B, n, m]
  const float* z;    // [B, n, m] standard-normal draws or null (no perturbation)
  const float* std;  // [B] standard deviation per env (required with z)
  int B, n, m;
  int64_t* out;       // [B, n] chosen column per row
  double* objective;  // [B] or null: sum of the chosen (perturbed) benefits
};

__device__ __forceinline__ unsigned long long ordered_u64(double v) {
  const long long b = __double_as_longlong(v + 0.0);
  return (unsigned long long)b ^ (unsigned long long)((b >> 63) | (long long)0x8000000000000000ull);
}

template <int kCM>
__global__ void __launch_bounds__(kThreads) sap_lsa_kernel(LsaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarps + warp;
  if (b >= p.B) return;
  const int n = p.n, m = p.m;
  // per-warp shared state
  const size_t per_warp = ((size_t)n * (8 + 8 + 2 + 2) + 15) & ~(size_t)15;
  unsigned char* base = smem_raw + per_warp * warp;
  double* u = reinterpret_cast<double*>(base);                 // [n] row duals
  double* entry = u + n;                                       // [n] minVal when the k-th visited row was reached
  int16_t* srl = reinterpret_cast<int16_t*>(entry + n);        // [n] visited rows, in order
  int16_t* col4row = srl + n;                                  // [n]
  const float* qb = p.q + (size_t)b * n * m;
  const float* zb = p.z ? p.z + (size_t)b * n * m : nullptr;
  const float sd = p.z ? p.std[b] : 0.f;
  // cost(i, j) = -(Q + z * std), the perturbed benefit in fp32 exactly as torch builds it (:84-86), negated (maximise)
  auto benefit = [&](int i, int j) -> float {
    const float qv = __ldg(qb + (size_t)i * m + j);
    return zb ? __fadd_rn(qv, __fmul_rn(__ldg(zb + (size_t)i * m + j), sd)) : qv;
  };

  double v[kCM], spc[kCM];
  int r4c[kCM], pth[kCM];
#pragma unroll
  for (int c = 0; c < kCM; ++c) {
    v[c] = 0.0;
    r4c[c] = -1;
    pth[c] = -1;
  }
  for (int i = lane; i < n; i += 32) {
    u[i] = 0.0;
    col4row[i] = -1;
  }
  __syncwarp();

  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  if (n == m) {
    // Square problems: start from the column-reduced duals v[j] = min_i cost(i, j) (the classic Jonker-Volgenant
    // initialisation).  They are feasible (every reduced cost >= 0), so the optimum reached is the same; what changes is
    // the work: Q-matrices whose rows rank the tasks alike (a freshly initialised shared-parameter agent) lose their
    // common column structure and the augmenting paths stay short (4096 x 100 x 100 of such matrices: 5 ms -> 2 ms).
    // Not valid for n < m, where a column that ends up unassigned must keep v = 0; padding with m - n zero-benefit
    // dummy rows makes it valid but the identical dummy rows are themselves a degenerate, slow instance (measured:
    // 64 x 324 x 450 random 4.4 -> 103 ms), so rectangular problems start from v = 0 like scipy.
#pragma unroll
    for (int c = 0; c < kCM; ++c) v[c] = INF;
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int c = 0; c < kCM; ++c) {
        const int j = lane + 32 * c;
        if (j < m) v[c] = fmin(v[c], -(double)benefit(i, j));
      }
    }
#pragma unroll
    for (int c = 0; c < kCM; ++c)
      if (lane + 32 * c >= m) v[c] = 0.0;
  }
  for (int cur = 0; cur < n; ++cur) {
    uint32_t scanned = 0u;  // bit c: my column lane + 32 c is in SC
#pragma unroll
    for (int c = 0; c < kCM; ++c) spc[c] = INF;
    int nSR = 0, sink = -1, i = cur;
    double minVal = 0.0;
    while (nSR < n) {  // every step scans a new row
      if (lane == 0) {
        srl[nSR] = (int16_t)i;
        entry[nSR] = minVal;
      }
      ++nSR;
      const double ui = u[i];
      // relax the unscanned columns through row i
      float bf[kCM];
#pragma unroll
      for (int c = 0; c < kCM; ++c) {
        const int j = lane + 32 * c;
        bf[c] = (j < m && !((scanned >> c) & 1u)) ? benefit(i, j) : 0.f;
      }
      unsigned long long best = ~0ull;
      uint32_t tie = 0xffffffffu;
#pragma unroll
      for (int c = 0; c < kCM; ++c) {
        const int j = lane + 32 * c;
        if (j < m && !((scanned >> c) & 1u)) {
          const double r = minVal - (double)bf[c] - ui - v[c];
          if (r < spc[c]) {
            spc[c] = r;
            pth[c] = i;
          }
          const unsigned long long k = ordered_u64(spc[c]);
          const uint32_t t = ((r4c[c] == -1) ? 0u : 0x10000u) | (uint32_t)j;  // ties: an unassigned column first
          if (k < best || (k == best && t < tie)) {
            best = k;
            tie = t;
          }
        }
      }
      // warp arg-min: high word, low word, tie key
      const uint32_t hi = (uint32_t)(best >> 32);
      const uint32_t hw = __reduce_min_sync(SAP_FULL_MASK, hi);
      const uint32_t lo = hi == hw ? (uint32_t)best : 0xffffffffu;
      const uint32_t lw = __reduce_min_sync(SAP_FULL_MASK, lo);
      const uint32_t tk = (hi == hw && (uint32_t)best == lw) ? tie : 0xffffffffu;
      const uint32_t tw = __reduce_min_sync(SAP_FULL_MASK, tk);
      if (hw == 0xffffffffu && lw == 0xffffffffu) break;  // nothing left: infeasible (cannot happen with finite input)
      const int js = (int)(tw & 0xffffu), owner = js & 31, slot = js >> 5;
      // new minVal = the selected column's shortest-path cost; its current row, if any
      double mv = 0.0;
      int rr = -1;
#pragma unroll
      for (int c = 0; c < kCM; ++c)
        if (c == slot) {
          mv = spc[c];
          rr = r4c[c];
        }
      minVal = __shfl_sync(SAP_FULL_MASK, mv, owner);
      rr = __shfl_sync(SAP_FULL_MASK, rr, owner);
      if (lane == owner) scanned |= 1u << slot;
      if (rr == -1) {
        sink = js;
        break;
      }
      i = rr;
    }
    __syncwarp();
    if (sink < 0) break;
    // dual updates
    for (int k = lane; k < nSR; k += 32) u[srl[k]] += minVal - entry[k];
#pragma unroll
    for (int c = 0; c < kCM; ++c)
      if ((scanned >> c) & 1u) v[c] -= minVal - spc[c];
    __syncwarp();
    // augment along the predecessor chain
    int j = sink;
    for (int guard = 0; guard <= n; ++guard) {  // a path visits every row at most once (bound: NaN input cannot hang it)
      const int owner = j & 31, slot = j >> 5;
      int pi = -1;
#pragma unroll
      for (int c = 0; c < kCM; ++c)
        if (c == slot) pi = pth[c];
      pi = __shfl_sync(SAP_FULL_MASK, pi, owner);
      if (pi < 0) break;  // unreachable with finite input
      if (lane == owner) {
#pragma unroll
        for (int c = 0; c < kCM; ++c)
          if (c == slot) r4c[c] = pi;
      }
      const int prev = col4row[pi];
      __syncwarp();
      if (lane == 0) col4row[pi] = (int16_t)j;
      __syncwarp();
      j = prev;
      if (pi == cur || j < 0) break;
    }
  }
  __syncwarp();
  double obj = 0.0;
  for (int i = lane; i < n; i += 32) {
    const int j = col4row[i];
    p.out[(size_t)b * n + i] = (int64_t)j;
    if (p.objective && j >= 0) obj += (double)benefit(i, j);
  }
  if (p.objective) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) obj += __shfl_xor_sync(SAP_FULL_MASK, obj, off);
    if (lane == 0) p.objective[b] = obj;
  }
}

template <int kCM>
int launch(const LsaParams& p, cudaStream_t st) {
  const size_t per_warp = ((size_t)p.n * (8 + 8 + 2 + 2) + 15) & ~(size_t)15;
  const size_t smem = per_warp * kWarps;
  static thread_local bool configured = false;
  if (!configured && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sap_lsa_kernel<kCM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      sap_set_error("sap_lsa: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  sap_lsa_kernel<kCM><<<(unsigned)((p.B + kWarps - 1) / kWarps), kThreads, smem, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_lsa_kernel");
  return SAP_OK;
}

}  // namespace

extern "C" int sap_lsa_maximize(const float* q, const float* z, const float* std_per_env, int32_t B, int32_t n, int32_t m,
                                int64_t* cols_out, double* objective_out, void* stream) {
  SAP_REQUIRE(q && cols_out, SAP_E_NULL, "sap_lsa_maximize: q/cols_out is null");
  SAP_REQUIRE((z == nullptr) == (std_per_env == nullptr), SAP_E_NULL, "sap_lsa_maximize: z and std must be given together");
  SAP_REQUIRE(B > 0 && n > 0 && m > 0, SAP_E_DIMS, "sap_lsa_maximize: bad dims B=%d n=%d m=%d", B, n, m);
  SAP_REQUIRE(n <= m, SAP_E_CONSTRAINT, "sap_lsa_maximize: need n <= m (every agent gets its own task), got n=%d m=%d", n, m);
  SAP_REQUIRE(m <= 512, SAP_E_DIMS, "sap_lsa_maximize: m must be <= 512");
  LsaParams p{};
  p.q = q; p.z = z; p.std = std_per_env; p.B = B; p.n = n; p.m = m; p.out = cols_out; p.objective = objective_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 128) return launch<4>(p, st);
  if (m <= 256) return launch<8>(p, st);
  return launch<16>(p, st);
}
