// Epsilon-greedy selection shared by the stand-alone selector kernel (sap_select.cu) and the fused selection + env step
// launch (sap_rollout_step, sap_real_fast2.cu).  Reference: action_selectors/classic_selectors.py:37-54.
#pragma once
#include "sap_common.cuh"

// the rank-th available action of a row (rank counted over set bits in index order), one warp
__device__ __forceinline__ int warp_rank_select(const uint8_t* avail_row, int A, int rank, int lane) {
  int seen = 0, found = -1;
  for (int base = 0; base < A && found < 0; base += 32) {
    const int j = base + lane;
    const bool av = (j < A) && (avail_row ? avail_row[j] != 0 : true);
    const unsigned bal = __ballot_sync(SAP_FULL_MASK, av);
    const int c = __popc(bal);
    if (rank < seen + c) {
      // the (rank-seen)-th set bit of bal
      const int want = rank - seen;
      const int before = __popc(bal & ((1u << lane) - 1u));
      const unsigned hit = __ballot_sync(SAP_FULL_MASK, av && before == want);
      found = base + (__ffs(hit) - 1);
    }
    seen += c;
  }
  return found;
}

struct SelParams {  // (the C ABI's SapSelectArgs plus the launch dimensions)
  const float* q;
  const int32_t* top;    // filtered only
  const uint8_t* avail;  // nullable
  int B, n, A, m, M;
  float eps;
  const float* eps_dev;  // when non-null the kernel reads epsilon from device memory (CUDA-graph replays)
  uint64_t seed;
  const uint64_t* episode_ctr;
  const int32_t* k;
  const float* u_tie;
  const float* u_explore;
  const float* u_action;
  int64_t* out;
};

// Classic epsilon-greedy, batched 32 rows per warp: lane r draws the Philox uniforms of row (base + r) once, then
// the warp walks the 32 rows, each a 128-bit coalesced read of the Q row and two warp-wide reductions
// (max of order-preserving keys, then min index among the maxima = torch's first-index argmax).
__device__ __forceinline__ uint32_t f32_ordered(float v) {
  v += 0.0f;  // -0.0 -> +0.0 so that equal floats get equal keys
  const uint32_t u = __float_as_uint(v);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}

// Narrow rows (32 < A <= 16 kLPR), everything available - the rollout case: kLPR lanes per row, 32 / kLPR rows per
// pass.  A lane reads up to four 128-bit pieces of its row and keeps its best (key, first index); log2(kLPR)
// xor-shuffle steps reduce the row.  ~2x fewer instructions per row than two warp-wide redux per row, and several
// rows of loads in flight.  Returns the action of row (base + lane).
template <int kLPR>
__device__ __forceinline__ int narrow_rows(const SelParams& p, int64_t base, int nrows, int A, float eps, float ue,
                                           float ua, int lane) {
  constexpr int kRPP = 32 / kLPR;  // rows per pass
  const int g = lane / kLPR, l = lane % kLPR;
  int my_action = 0;
  for (int it = 0; it < kLPR; ++it) {
    const int rr = it * kRPP + g;
    const bool row_ok = rr < nrows;
    const float* qr = p.q + (base + (row_ok ? rr : 0)) * A;
    float4 v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = 4 * (l + kLPR * t);
      v[t] = (row_ok && j < A) ? __ldg(reinterpret_cast<const float4*>(qr + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t bk = 0u;
    int bi = 0x7fffffff;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = 4 * (l + kLPR * t);
      if (j < A) {
        const float x[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t key = f32_ordered(x[c]);
          if (key > bk) {  // ascending j within the lane: strict > keeps the first index
            bk = key;
            bi = j + c;
          }
        }
      }
    }
#pragma unroll
    for (int off = 1; off < kLPR; off <<= 1) {
      const uint32_t ok = __shfl_xor_sync(SAP_FULL_MASK, bk, off);
      const int oi = __shfl_xor_sync(SAP_FULL_MASK, bi, off);
      if (ok > bk || (ok == bk && oi < bi)) {
        bk = ok;
        bi = oi;
      }
    }
    int action = bi;  // first-index argmax (classic_selectors.py:52-54)
    const float ue_r = __shfl_sync(SAP_FULL_MASK, ue, rr & 31);
    const float ua_r = __shfl_sync(SAP_FULL_MASK, ua, rr & 31);
    if (ue_r < eps) {  // explore (:49-51): the floor(u * A)-th action, all of them being available
      const int rank = (int)floorf(__fmul_rn(ua_r, (float)A));
      action = min(rank, A - 1);
    }
    // row `lane` is row (lane % kRPP) of pass (lane / kRPP)
    const int got = __shfl_sync(SAP_FULL_MASK, action, (lane % kRPP) * kLPR);
    if (lane / kRPP == it) my_action = got;
  }
  return my_action;
}

// The classic selection of the 32 rows [base, base + nrows) by ONE warp; returns the action of row base + lane.
// Philox draws are keyed by the GLOBAL row index, so how rows are grouped into warps and CTAs does not matter.
__device__ __forceinline__ int sap_classic_select_rows(const SelParams& p, int64_t base, int nrows, int vec4, int lane) {
  const int A = p.A;
  const float eps = p.eps_dev ? *p.eps_dev : p.eps;
  const uint32_t ep_lo = p.episode_ctr ? (uint32_t)(*p.episode_ctr) : 0u;
  const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  // per-lane draws for row base + lane
  float ue = 2.f, ua = 0.f;
  {
    const int64_t row = base + lane;
    if (lane < nrows) {
      if (p.u_explore) {
        ue = p.u_explore[row];
        ua = p.u_action[row];
      } else {
        const uint32_t step = p.k ? (uint32_t)p.k[row / p.n] : 0u;
        const SapPhilox4 r = sap_philox4x32_10((uint32_t)row, ep_lo, step, 0u, k0, k1);
        ue = sap_u01(r.x);
        ua = sap_u01(r.y);
      }
    }
  }
  int my_action = 0;
  if (vec4 && A <= 128 && A > 32 && !p.avail) {
    if (A > 64) my_action = narrow_rows<8>(p, base, nrows, A, eps, ue, ua, lane);
    else my_action = narrow_rows<4>(p, base, nrows, A, eps, ue, ua, lane);
    return my_action;
  }
  for (int rr = 0; rr < nrows; ++rr) {
    const int64_t row = base + rr;
    const float* qr = p.q + row * A;
    const uint8_t* av = p.avail ? p.avail + row * A : nullptr;
    uint32_t bk = 0u;  // best ordered key of this lane (0 is below every real key, even -inf)
    int bi = 0x7fffffff, n_avail = 0;
    if (vec4) {
      for (int j = lane * 4; j < A; j += 128) {
        const float4 v = *reinterpret_cast<const float4*>(qr + j);
        const float x[4] = {v.x, v.y, v.z, v.w};
        uchar4 a4 = make_uchar4(1, 1, 1, 1);
        if (av) a4 = *reinterpret_cast<const uchar4*>(av + j);
        const uint8_t ok[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          n_avail += ok[c] != 0;
          const uint32_t key = f32_ordered(ok[c] ? x[c] : -INFINITY);  // classic_selectors.py:46-47
          if (key > bk) {
            bk = key;
            bi = j + c;
          }
        }
      }
    } else {
      for (int j = lane; j < A; j += 32) {
        const bool ok = av ? av[j] != 0 : true;
        n_avail += ok;
        const uint32_t key = f32_ordered(ok ? qr[j] : -INFINITY);
        if (key > bk) {
          bk = key;
          bi = j;
        }
      }
    }
    const uint32_t gmax = __reduce_max_sync(SAP_FULL_MASK, bk);
    int action = (int)__reduce_min_sync(SAP_FULL_MASK, (uint32_t)(bk == gmax ? bi : 0x7fffffff));
    const float ue_r = __shfl_sync(SAP_FULL_MASK, ue, rr);
    if (ue_r < eps) {  // warp-uniform: explore (:49-51)
      if (av) n_avail = (int)__reduce_add_sync(SAP_FULL_MASK, (uint32_t)n_avail);
      else n_avail = A;
      if (n_avail > 0) {
        const float ua_r = __shfl_sync(SAP_FULL_MASK, ua, rr);
        int rank = (int)floorf(__fmul_rn(ua_r, (float)n_avail));
        rank = min(rank, n_avail - 1);
        action = av ? warp_rank_select(av, A, rank, lane) : rank;
      }
    }
    if (lane == rr) my_action = action;
  }
  return my_action;
}

