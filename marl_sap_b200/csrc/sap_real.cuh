// Parameters shared by the generic (sap_real.cu) and the fast (sap_real_fast.cu) RealConstellationEnv kernels.
#pragma once
#include "sap_common.cuh"

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
template <typename Put>
__device__ __forceinline__ void warp_select_cached(int len, int count, bool idx_desc, int lane, const double (&vals)[16],
                                                   Put put) {
  double lastv = 0.0;
  int lasti = -1;
  for (int r = 0; r < count; ++r) {
    double bv = 0.0;
    int bi = -1;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int j = lane + 32 * c;
      if (j < len) {
        const double v = vals[c];
        const bool taken = lasti >= 0 && !sap_better(lastv, lasti, v, j, idx_desc);
        if (!taken && sap_better(v, j, bv, bi, idx_desc)) {
          bv = v;
          bi = j;
        }
      }
    }
    sap_warp_argbest(bv, bi, idx_desc);
    lastv = bv;
    lasti = bi;
    if (lane == 0) put(r, bi);
  }
}

// Launches the shared-memory-resident fast kernel when the problem fits it.
// Returns SAP_OK / error like every entry point; *handled = 0 means "not eligible, use the generic kernel".
int sap_real_fast_try(RealParams& p, void* stream, int* handled);

// One environment spread over many CTAs, for shapes whose window sums do not fit shared memory (sap_real_large.cu).
int sap_real_large_launch(RealParams& p, void* stream);
int64_t sap_real_large_scratch_doubles(const SapEnvDims& d);
