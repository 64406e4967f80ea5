// Parameters shared by the generic (sap_real.cu) and the fast (sap_real_fast.cu) RealConstellationEnv kernels.
#pragma once
#include "sap_common.cuh"
#include "sap_select.cuh"

struct RealParams {
  SapEnvDims d;
  const float* planes;   // [B,T,n,m]
  const float* plane_stats;  // [B,T,2] per-plane {min,max} or null
  const float* prios;    // [m] or null
  const float* ttrans;   // [m,m] or null (default 1 - I)
  double lambda_;
  const int64_t* actions;  // [B,n]
  int32_t* k;
  int32_t* prev;
  double* ep_return;
  int32_t* counts_out;
  SapBatchView view;
  int32_t* top_out;
  double* scratch;  // sap_real_scratch_doubles() doubles when tot does not fit shared memory (sap_real_large.cu)
  int is_reset;
  int debug_skip_redo;  // timing experiments only (SAP_DEBUG_SKIP_REDO=1): accept uncertified lists
  int tot_in_smem;
  int ms;  // row stride of tot (odd -> conflict-free column walks)
  int lookahead;  // sap_real_fast2: blocks of L2 look-ahead for the benefit window (0 = off)
  // power / interference envs (generic kernel only)
  const uint8_t* dead;    // [B,n] agents whose reward is 0 this step (out of power), or null
  int32_t* nbr_out;       // [B,n,N] rival indices of the new observation, or null
  const int64_t* prev0;   // reset only: [B,n] initial prev_assigns instead of arange(n), or null
  int obs_row;            // elements between consecutive agents' observation rows (0 = obs_size: packed)
  int obs_only;     // sap_real_fast2: build the observation of slot k + 1 only (no rewards, no counters, flags = 0)
  int large_exact;
  // sap_rollout_step: when sel.q is set the CTA of an env first selects its agents' actions (classic epsilon-greedy,
  // everything available) into sel.out == actions, then steps; only the one-CTA-per-env kernels of sap_real_fast2.cu
  SelParams sel;
  int sel_vec4;  // multi-CTA path: exact float64 selection even where the keyed lists would apply (selector override)
};


// Successive selection of the `count` best of x[0..len) under a stable total order, one warp.
// Round r picks the best element that ranks strictly after round r-1's winner, so no removal
// flags are needed.  get(j) returns the float64 key of element j.
template <typename Get, typename Put>
__device__ __forceinline__ void warp_select(int len, int count, bool idx_desc, int lane, Get get, Put put) {
  double lastv = 0.0;
  int lasti = -1;
  for (int r = 0; r < count; ++r) {
    double bv = 0.0;
    int bi = -1;
    for (int j = lane; j < len; j += 32) {
      double v = get(j);
      if (lasti >= 0 && !sap_better(lastv, lasti, v, j, idx_desc)) continue;  // already taken
      if (sap_better(v, j, bv, bi, idx_desc)) {
        bv = v;
        bi = j;
      }
    }
    sap_warp_argbest(bv, bi, idx_desc);
    lastv = bv;
    lasti = bi;
    if (lane == 0) put(r, bi);
  }
}

// Same selection on values cached in registers: lane holds elements lane, lane + 32, ... (at most 16 of them).
// The float64 values are mapped once to order-preserving 64-bit integers, so a round is three redux.sync
// reductions (high word, low word, index) over integer compares, with no float64 compare or 64-bit shuffle on the
// critical path: ~8x less latency than the compare-and-shuffle version, which matters because the exact path is what
// the certified kernels fall back to and a single slow warp sets the tail of a short launch.
template <typename Put>
__device__ __forceinline__ void warp_select_cached(int len, int count, bool idx_desc, int lane, const double (&vals)[16],
                                                   Put put) {
  uint32_t hi[16], lo[16];
  uint32_t live = 0u;  // bit c: element lane + 32 c exists and has not been selected yet
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    // +0.0 canonicalises -0.0 (they compare equal as doubles); then flip so that unsigned order == double order
    const long long b = __double_as_longlong(vals[c] + 0.0);
    const unsigned long long k = (unsigned long long)b ^ (unsigned long long)((b >> 63) | (long long)0x8000000000000000ull);
    hi[c] = (uint32_t)(k >> 32);
    lo[c] = (uint32_t)k;
    if (lane + 32 * c < len) live |= 1u << c;
  }
  for (int r = 0; r < count; ++r) {
    uint32_t h = 0u;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if ((live >> c) & 1u) h = max(h, hi[c]);
    const uint32_t hw = __reduce_max_sync(SAP_FULL_MASK, h);
    uint32_t cand = 0u, l = 0u;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (((live >> c) & 1u) && hi[c] == hw) {
        cand |= 1u << c;
        l = max(l, lo[c]);
      }
    const uint32_t lw = __reduce_max_sync(SAP_FULL_MASK, l);
    // among the elements equal to the maximum: smallest index (argsort(-x) order) or largest (reversed argsort(x))
    int best = idx_desc ? -1 : 0x7fffffff;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (((cand >> c) & 1u) && lo[c] == lw) best = idx_desc ? max(best, lane + 32 * c) : min(best, lane + 32 * c);
    const int bi = idx_desc ? __reduce_max_sync(SAP_FULL_MASK, best) : __reduce_min_sync(SAP_FULL_MASK, best);
    if ((bi & 31) == lane) live &= ~(1u << (bi >> 5));
    if (lane == 0) put(r, bi);
  }
}

// Launches the shared-memory-resident fast kernel when the problem fits it.
// Returns SAP_OK / error like every entry point; *handled = 0 means "not eligible, use the generic kernel".
int sap_real_fast_try(RealParams& p, void* stream, int* handled, bool gen1_only = false);
// Second-generation kernel for the shipped configuration (M = N = 10, L = 3, fp16) at 64 < n <= 128 (sap_real_fast2.cu).
int sap_real_fast2_try(RealParams& p, void* stream, int* handled);
// current sap_real_select_kernel override (sap_real.cu)
int sap_real_path_override();

// One environment spread over many CTAs, for shapes whose window sums do not fit shared memory (sap_real_large.cu).
int sap_real_large_launch(RealParams& p, void* stream);
int64_t sap_real_large_scratch_doubles(const SapEnvDims& d);
