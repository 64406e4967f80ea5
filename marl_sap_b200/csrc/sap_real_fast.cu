// RealConstellationEnv step + observation build, shared-memory-resident fast path (one CTA per environment).
//
// Same contract and same results as the generic kernel in sap_real.cu (reference:
// /root/reference/src/envs/real_constellation_env.py step :135-175, beta_hat :282-328, _build_obs :177-230),
// restructured for throughput:
//   1. ONE pass over the L-plane benefit window (128-bit loads): every (agent, task) pair gets its float64
//      window sum mapped to a 32-bit selection key, and its benefits are parked in shared memory already
//      rounded to the observation dtype (fp16 for the reference's real scheme, so the tile is 60 KB at 100x100
//      and two CTAs share an SM);
//   2. key = fixed-point image of the float64 sum (scale = power of two chosen from per-plane {min,max}
//      metadata, so no extra pass) + one "inexact" bit.  The map is monotone, and two equal keys with the bit
//      clear are PROVEN equal sums;
//   3. all top-k's (agent's top-M tasks, top-(M+M/2) for the rivals' other tasks, top-N rivals) run as
//      register-resident sorting networks on packed (key | index) 32-bit words, TPL threads per list,
//      partial lists merged with warp shuffles: one VIMNMX pair per compare-exchange;
//   4. a list is accepted only if it is PROVABLY the exact float64 answer under the reference's stable tie
//      rules (strictly decreasing keys, or ties between exact keys); otherwise it is queued and redone by the
//      exact float64 warp selection.  Results never depend on the key resolution;
//   5. observation rows are assembled in shared memory and leave with 128-bit coalesced stores, together with
//      the fp32 copy the agent network reads (agent_in).
#include "sap_real.cuh"
#include "sap_sortnet.cuh"
#include <stdlib.h>

namespace {

// threads cooperating on one list: 2, or 4 when there are too few lists to occupy the CTA with 2 (n <= 64)
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr size_t kMaxSmem = 227 * 1024 - 1024;

// timing-ablation hooks (profiling builds only: -DSAP_ABLATE); in release builds the tests below are constant false
#ifdef SAP_ABLATE
#define SAP_DBG(p) ((p).debug_skip_redo)
#else
#define SAP_DBG(p) 0
#endif

#define SAP_CE(a, b)            \
  {                             \
    uint32_t hi__ = max(a, b);  \
    b = min(a, b);              \
    a = hi__;                   \
  }

// Shared-memory plan.  Two overlays share one region: while the lists are built it holds the key tile and the
// scratch lists; for the gather it is refilled with the benefits (rounded to the obs dtype) and the staging rows.
// At 100 x 100 (fp16 obs, 8-bit indices) that is 7.3 KB persistent + max(45 KB, 68 KB) = 75 KB: three CTAs per SM.
struct FastLayout {
  size_t D, nbr, other, lut, prios, pv;        // live through the whole kernel
  size_t k32, E, dmask, cnt, red, queue;       // overlay 1: list building
  size_t tile, stage;                          // overlay 2: gather
  size_t total;
  int ms, mw, rows_per_pass;
};

__host__ __device__ inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline FastLayout fast_layout(const SapEnvDims& d, int out_esz, int idx_esz, bool prios) {
  FastLayout f;
  const int H = d.M / 2, K2 = d.M + H;
  f.ms = (d.m + 3) & ~3;
  f.mw = (d.m + 31) / 32;
  size_t off = 0;
  f.D = off;     off = up16(off + (size_t)idx_esz * d.n * d.M);
  f.nbr = off;   off = up16(off + (size_t)idx_esz * d.n * d.N);
  f.other = off; off = up16(off + (size_t)idx_esz * d.n * d.N * H);
  f.lut = off;   off = up16(off + sizeof(uint16_t) * (size_t)(d.M + d.N * d.M + d.N * H));
  f.prios = off; off = up16(off + (prios ? sizeof(double) * (size_t)d.m : 0));
  f.pv = off;    off = up16(off + sizeof(uint16_t) * (size_t)d.n);
  const size_t base = off;
  f.k32 = off;   off = up16(off + sizeof(uint32_t) * (size_t)d.n * f.ms);
  f.E = off;     off = up16(off + (size_t)idx_esz * d.n * K2);
  f.dmask = off; off = up16(off + sizeof(uint32_t) * (size_t)d.n * f.mw);
  f.cnt = off;   off = up16(off + sizeof(int32_t) * (size_t)d.m);
  f.red = off;   off = up16(off + sizeof(double) * 80);
  f.queue = off; off = up16(off + sizeof(int32_t) * (2 * (size_t)d.n + 4));
  const size_t end1 = off;
  const size_t row_bytes = (size_t)out_esz * (d.M * d.L + d.N * d.M * d.L + d.N * H * d.L + d.M);
  f.rows_per_pass = kWarps;
  off = base;
  f.tile = off;  off = up16(off + (size_t)out_esz * d.L * d.n * d.m);
  f.stage = off; off = up16(off + (size_t)kWarps * row_bytes + 16);  // +16: staging alignment shift
  const size_t end2 = off;
  f.total = end1 > end2 ? end1 : end2;
  return f;
}

// ---------------------------------------------------------------------------------------------------------
// top-16 of a list under the packed order, TPL adjacent lanes per list.  key(e) returns the packed word of
// element e (0 = padding).  On return every lane of the group holds the same descending top[16].
template <int TPL, typename KeyFn>
__device__ __forceinline__ void group_top16(int len, int s, KeyFn key, uint32_t (&top)[16]) {
  const int per = (len + TPL - 1) / TPL;  // elements per thread
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int e = s + TPL * c;
    top[c] = (e < len) ? key(e) : 0u;
  }
  SAP_SORT16(top);
  for (int base = 16; base < per; base += 16) {
    uint32_t ch[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int e = s + TPL * (base + c);
      ch[c] = (e < len) ? key(e) : 0u;
    }
    SAP_SORT16(ch);
#pragma unroll
    for (int c = 0; c < 16; ++c) top[c] = max(top[c], ch[15 - c]);
    SAP_BITONIC_MERGE16(top);
  }
#pragma unroll
  for (int stride = 1; stride < TPL; stride <<= 1) {
    uint32_t ot[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) ot[c] = __shfl_xor_sync(SAP_FULL_MASK, top[15 - c], stride);
#pragma unroll
    for (int c = 0; c < 16; ++c) top[c] = max(top[c], ot[c]);
    SAP_BITONIC_MERGE16(top);
  }
}

// The first `need` entries of a sorted packed list are the exact float64 answer (value desc, index asc) when every
// adjacent pair among entries 0..need is strictly decreasing in key, or ties between two EXACT keys (bit 0 clear).
__device__ __forceinline__ bool certified(const uint32_t (&top)[16], int need, int ib) {
  bool ok = true;
#pragma unroll
  for (int t = 0; t < 15; ++t) {
    const uint32_t a = top[t] >> ib, b = top[t + 1] >> ib;
    if (t < need) ok = ok && (a > b || (a == b && !(a & 1u)));
  }
  return ok;
}

template <typename OutT>
__device__ __forceinline__ OutT to_out(double v);
template <>
__device__ __forceinline__ float to_out<float>(double v) { return (float)v; }
template <>
__device__ __forceinline__ __half to_out<__half>(double v) { return __double2half(v); }
template <typename OutT>
__device__ __forceinline__ OutT to_out_f(float v);
template <>
__device__ __forceinline__ float to_out_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half to_out_f<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float out_to_f(float v) { return v; }
__device__ __forceinline__ float out_to_f(__half v) { return __half2float(v); }

// L2 residency control: the window is read twice by the same CTA a few microseconds apart (keys, then gather
// tile), while the obs / agent-input stores in between are never re-read by this kernel.  First read = evict_last,
// second read and all bulk stores = evict_first, so the re-read hits L2 instead of going back to HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldg_hint4(const float* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_hint4(void* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

template <typename OutT_, typename IdxT_, bool kPrios_, bool kCommon_, int kTPL_ = 2, int kFixed_ = 0>
struct Cfg {
  static constexpr int kTPL = kTPL_;
  static constexpr int kFixed = kFixed_;  // n = m = kFixed compiled in (BASELINE config 50 x 50); 0 = read at run time
  static constexpr int kMinBlocks = kFixed_ == 50 ? 7 : (kTPL_ == 4 ? 5 : 3);  // small shapes need little shared memory: more CTAs per SM;
                                        // 50 x 50 compiled in fits 32 registers (20 bytes of spill): 7 CTAs per SM = 1036 slots, one wave for 1024 envs
  using OutT = OutT_;
  using IdxT = IdxT_;
  static constexpr bool kPrios = kPrios_;
  static constexpr bool kCommon = kCommon_;  // M = 10, N = 10, L = 3 known at compile time
};

template <typename C>
__device__ __forceinline__ void process_env(const RealParams& p, const int b, unsigned char* smem) {
  using OutT = typename C::OutT;
  using IdxT = typename C::IdxT;
  constexpr bool kPrios = C::kPrios;
  constexpr int TPL = C::kTPL;
  SapEnvDims d = p.d;
  if (C::kFixed) d.n = d.m = C::kFixed;
  if (C::kFixed && C::kCommon) {  // the layout below then folds into constants
    d.L = 3;
    d.M = 10;
    d.N = 10;
  }
  const int n = d.n, m = d.m, T = d.T;
  const int L = C::kCommon ? 3 : d.L, M = C::kCommon ? 10 : d.M, N = C::kCommon ? 10 : d.N;
  const int H = M / 2, K2 = M + H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nm = n * m;
  const int obs_size = M * L + N * M * L + N * H * L + M;
  const int npairs = M + N * M + N * H;
  const FastLayout f = fast_layout(d, (int)sizeof(OutT), (int)sizeof(IdxT), kPrios);
  const int ms = f.ms, mw = f.mw;
  OutT* tile = reinterpret_cast<OutT*>(smem + f.tile);              // [L][n][m], already in the obs dtype
  uint32_t* K32 = reinterpret_cast<uint32_t*>(smem + f.k32);         // [n][ms] selection keys
  IdxT* sD = reinterpret_cast<IdxT*>(smem + f.D);                    // [n][M]   top-M tasks (value desc, idx asc)
  IdxT* sE = reinterpret_cast<IdxT*>(smem + f.E);                    // [n][K2]  top-(M+H) tasks (value desc, idx DESC)
  IdxT* sNbr = reinterpret_cast<IdxT*>(smem + f.nbr);                // [n][N]   rivals
  IdxT* sOther = reinterpret_cast<IdxT*>(smem + f.other);            // [n][N][H]
  uint32_t* sMask = reinterpret_cast<uint32_t*>(smem + f.dmask);     // [n][mw] membership bits of D[i]
  int32_t* sCnt = reinterpret_cast<int32_t*>(smem + f.cnt);
  double* sPrio = reinterpret_cast<double*>(smem + f.prios);
  double* sRed = reinterpret_cast<double*>(smem + f.red);
  int32_t* sQ = reinterpret_cast<int32_t*>(smem + f.queue);          // [0]=rows count, [1]=nbr count, then ids
  uint16_t* sLut = reinterpret_cast<uint16_t*>(smem + f.lut);        // output pair -> (rival slot << 8 | column slot)
  uint16_t* sPrev = reinterpret_cast<uint16_t*>(smem + f.pv);        // [n] the NEW previous task of every agent (flag columns, :222)
  int32_t* qRows = sQ + 4;
  int32_t* qNbr = sQ + 4 + n;

  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  const size_t env_plane0 = d.shared_planes ? (size_t)0 : (size_t)b * T;
  const float* env_planes = p.planes + env_plane0 * nm;
  const SapBatchView& vw = p.view;

  // The reward phase is a chain of dependent global loads (k -> actions -> chosen benefits -> ...); everything that
  // does not depend on an earlier load is issued up front so the chain costs two DRAM round trips, not five.
  int a_mine = 0, pv_mine = 0;  // action / previous task of agent `tid`
  if (!p.is_reset && tid < n) {
    a_mine = min(max((int)p.actions[(size_t)b * n + tid], 0), m - 1);
    pv_mine = p.prev[(size_t)b * n + tid];
  }
  const int k_old = p.is_reset ? -1 : p.k[b];
  if (!p.is_reset && k_old >= T) return;
  const int k_new = k_old + 1;
  const bool done = k_new >= T;
  const int Leff = done ? 0 : min(L, T - k_new);
  const float* win = env_planes + (size_t)k_new * nm;
  // per-plane {min, max} of the new window, also requested now (used by step 3 after the reward phase)
  const bool own_scale = p.plane_stats && !kPrios && (reinterpret_cast<uintptr_t>(p.plane_stats) & 7) == 0;
  float2 pst[4];
#pragma unroll
  for (int l = 0; l < 4; ++l)
    if (own_scale && l < Leff) pst[l] = __ldg(reinterpret_cast<const float2*>(p.plane_stats) + (env_plane0 + k_new + l));

  if (tid == 0) {
    sQ[0] = 0;
    sQ[1] = 0;
  }
  if (kPrios)
    for (int j = tid; j < m; j += kThreads) sPrio[j] = (double)p.prios[j];
  for (int j = tid; j < m; j += kThreads) sCnt[j] = 0;
  for (int pp = tid; pp < npairs; pp += kThreads) {  // obs layout (:225): own | rivals on my tasks | rivals' other tasks
    uint32_t code;
    if (pp < M) code = (0xffu << 8) | pp;
    else if (pp < M + N * M) code = (((pp - M) / M) << 8) | ((pp - M) % M);
    else code = (((pp - M - N * M) / H) << 8) | (0x80u + (pp - M - N * M));
    sLut[pp] = (uint16_t)code;
  }
  __syncthreads();

  // ------------------------------------------------------------------ 1. rewards at the old window (:135-164)
  if (!p.is_reset) {
    float mine[4] = {0.f, 0.f, 0.f, 0.f};  // chosen benefits of agent `tid` over the old window, loaded before the barrier
    if (tid < n) {
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (l < L && k_old + l < T) mine[l] = __ldg(env_planes + ((size_t)(k_old + l) * n + tid) * m + a_mine);
    }
    for (int i = tid; i < n; i += kThreads) {
      const int a = i == tid ? a_mine : min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
      atomicAdd(&sCnt[a], 1);
    }
    __syncthreads();
    double local_ret = 0.0;
    for (int i = tid; i < n; i += kThreads) {
      const int a = i == tid ? a_mine : min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
      const int pv = i == tid ? pv_mine : p.prev[(size_t)b * n + i];
      const double pr = kPrios ? sPrio[a] : 1.0;
      double sum = 0.0, b0 = 0.0;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (l < L && k_old + l < T) {
          const double v = (double)(i == tid ? mine[l] : env_planes[((size_t)(k_old + l) * n + i) * m + a]) * pr;
          if (l == 0) b0 = v;
          sum += v;
        }
      }
      const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);
      const double bh = b0 - p.lambda_ * (pen * (sum > 1e-12 ? 1.0 : 0.0));
      const double r = bh > 0.0 ? bh / (double)sCnt[a] : bh;
      local_ret += r;
      if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
      if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
      p.prev[(size_t)b * n + i] = a;
      sPrev[i] = (uint16_t)a;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local_ret += __shfl_xor_sync(SAP_FULL_MASK, local_ret, off);
    if (lane == 0) sRed[warp] = local_ret;
    if (vw.actions_onehot.ptr) {
      const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
      for (int i = warp; i < n; i += kWarps) {
        int a = (int)p.actions[(size_t)b * n + i];
        a = min(max(a, 0), m - 1);
        for (int j = lane; j < m; j += 32)
          sap_store_int(vw.actions_onehot.ptr, base + (int64_t)i * m + j, vw.actions_onehot.dtype, a == j ? 1 : 0);
      }
    }
    if (p.counts_out)
      for (int j = tid; j < m; j += kThreads) p.counts_out[(size_t)b * m + j] = sCnt[j];
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kWarps; ++w) t += sRed[w];
      atomicAdd(&p.ep_return[b], t);  // one add per env and launch: same result as +=, without waiting for the load
      p.k[b] = k_new;
      if (vw.terminated.ptr)
        sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, done);
    }
  } else {
    for (int i = tid; i < n; i += kThreads) {
      p.prev[(size_t)b * n + i] = i;
      sPrev[i] = (uint16_t)i;
    }
    if (tid == 0) {
      p.k[b] = 0;
      p.ep_return[b] = 0.0;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ 2. pre-transition scalars of slot k_new
  const int t_slot = k_new;
  if (tid == 0 && vw.filled.ptr) sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, t_slot), vw.filled.dtype, 1);
  if (vw.prev_assigns.ptr) {
    const int64_t base = sap_field_off(vw.prev_assigns, b, t_slot);
    for (int i = tid; i < n; i += kThreads)
      sap_store_int(vw.prev_assigns.ptr, base + i, vw.prev_assigns.dtype,
                    p.is_reset ? i : (i == tid ? a_mine : p.prev[(size_t)b * n + i]));
  }
  if (vw.avail_actions.ptr) {
    const int64_t base = sap_field_off(vw.avail_actions, b, t_slot);
    for (int e = tid; e < nm; e += kThreads) sap_store_int(vw.avail_actions.ptr, base + e, vw.avail_actions.dtype, 1);
  }
  OutT* obs_out = reinterpret_cast<OutT*>(vw.obs.ptr) + sap_field_off(vw.obs, b, t_slot);
  float* ain = vw.agent_in.ptr ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride : nullptr;
  const int64_t ain_row = vw.agent_in.t_stride;
  if (done) {  // :226-228
    for (int e = tid; e < n * obs_size; e += kThreads) obs_out[e] = to_out_f<OutT>(0.f);
    if (ain)
      for (int i = warp; i < n; i += kWarps)
        for (int c = lane; c < obs_size; c += 32) ain[i * ain_row + c] = 0.f;
    if (vw.beta.ptr) {
      const int64_t bb = sap_field_off(vw.beta, b, t_slot);
      for (int e = tid; e < nm * L; e += kThreads) sap_store_real(vw.beta.ptr, bb + e, vw.beta.dtype, 0.0);
    }
    return;
  }

  // ------------------------------------------------------------------ 3. bounds of the window sums -> key scale
  // origin, power-of-two scale and sign class of the keys from (lo, sum of max |value|, max |prio|, any prio < 0)
  auto key_scale = [&](double lo, double habs, double pabs, bool pneg, double& o_lo, double& o_scale, bool& o_nonneg) {
    const bool nonneg = lo >= 0.0 && !pneg;
    const double hi = habs * pabs * 1.0000001;  // >= every |window sum| (margin covers fp64 rounding)
    int e2 = 0;
    if (hi > 0.0) (void)frexp(hi, &e2);  // hi < 2^e2
    const int ib0 = 32 - __clz(max(n, m));  // max(n, m) <= 2^ib - 1: index code 0 is never a real element
    const int vb1 = 31 - ib0;               // bits of the fixed-point part
    o_lo = nonneg ? 0.0 : -ldexp(1.0, e2);                               // origin
    o_scale = hi > 0.0 ? ldexp(1.0, vb1 - e2 - (nonneg ? 0 : 1)) : 1.0;  // scale (power of two)
    o_nonneg = nonneg;
  };
  double k_lo, k_scale;
  bool k_nonneg;
  if (own_scale) {
    // common case: every thread derives the scale from the per-plane {min, max} metadata itself (L broadcast loads):
    // no serial section, no barrier
    float vmin = INFINITY;
    double vabs = 0.0;  // float64: a true bound of the sum of the per-plane maxima
#pragma unroll
    for (int l = 0; l < 4; ++l)
      if (l < Leff) {
        vmin = fminf(vmin, pst[l].x);
        vabs += (double)fmaxf(fabsf(pst[l].x), fabsf(pst[l].y));
      }
    key_scale((double)vmin, vabs, 1.0, false, k_lo, k_scale, k_nonneg);
  } else {
    float vmin = INFINITY;
    double vabs = 0.0;  // min value and sum over planes of max |value| (float64: a true bound)
    if (p.plane_stats) {
      if (tid == 0) {
        for (int l = 0; l < Leff; ++l) {
          const float lo = p.plane_stats[2 * (env_plane0 + k_new + l)], hi = p.plane_stats[2 * (env_plane0 + k_new + l) + 1];
          vmin = fminf(vmin, lo);
          vabs += (double)fmaxf(fabsf(lo), fabsf(hi));
        }
        sRed[64] = (double)vmin;
        sRed[65] = vabs;
      }
    } else {  // no metadata: one extra read of the window
      float amax = 0.f;
      for (int e = tid; e < Leff * nm; e += kThreads) {
        const float v = win[e];
        vmin = fminf(vmin, v);
        amax = fmaxf(amax, fabsf(v));
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(SAP_FULL_MASK, vmin, off));
        amax = fmaxf(amax, __shfl_xor_sync(SAP_FULL_MASK, amax, off));
      }
      if (lane == 0) {
        sRed[warp] = (double)vmin;
        sRed[32 + warp] = (double)amax;
      }
      __syncthreads();
      if (tid == 0) {
        double lo = sRed[0], am = sRed[32];
        for (int w = 1; w < kWarps; ++w) {
          lo = fmin(lo, sRed[w]);
          am = fmax(am, sRed[32 + w]);
        }
        sRed[64] = lo;
        sRed[65] = am * Leff;
      }
    }
    if (tid == 0) {
      double pabs = 1.0;
      bool pneg = false;
      if (kPrios) {
        pabs = 0.0;
        for (int j = 0; j < m; ++j) {
          pabs = fmax(pabs, fabs(sPrio[j]));
          pneg = pneg || sPrio[j] < 0.0;
        }
      }
      double o_lo, o_scale;
      bool o_nonneg;
      key_scale(sRed[64], sRed[65], pabs, pneg, o_lo, o_scale, o_nonneg);
      sRed[66] = o_lo;
      sRed[67] = o_scale;
      sRed[68] = o_nonneg ? 1.0 : 0.0;
    }
    __syncthreads();
    k_lo = sRed[66];
    k_scale = sRed[67];
    k_nonneg = sRed[68] != 0.0;
  }
  const int ib = 32 - __clz(max(n, m));
  const uint32_t imask = (1u << ib) - 1u;
  const uint32_t fixed_max = (1u << (31 - ib)) - 1u;

  // key of a float64 window sum: monotone; bit 0 = "conversion was not exact"
  auto make_key = [&](double tot) -> uint32_t {
    // floor(y) for y >= 0 without float<->int conversion instructions (they run at a quarter of the fp64 rate and
    // used to bound this pass): adding 2^52 rounds to the nearest integer, one compare turns that into the floor.
    // (Floor, not nearest: an inexact key must mean "strictly between fx and fx + 1" for the certificates.)
    const double y = (tot - k_lo) * k_scale;  // exact when k_lo == 0 (power-of-two scale)
    const double t = y + 4503599627370496.0;
    const double r = t - 4503599627370496.0;
    uint32_t fx = (uint32_t)__double2loint(t) - (r > y ? 1u : 0u);
    bool inexact = !k_nonneg || (r != y);
    if (fx > fixed_max) {
      fx = fixed_max;
      inexact = true;
    }
    return (fx << 1) | (inexact ? 1u : 0u);
  };

  // ------------------------------------------------------------------ 4. one pass over the window: keys + rounded tile
  const bool vec4 = (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(win) & 15) == 0);
  if (vec4) {
    const int m4 = m >> 2, total4 = nm >> 2;
    // 4 adjacent tasks of one agent: L float4 in, 4 keys + L x 4 rounded benefits out
    auto finish = [&](int e4, int i, int j, const float4 (&v)[4]) {
      double tot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        if (l < Leff) {
          const float x[4] = {v[l].x, v[l].y, v[l].z, v[l].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) tot[c] += kPrios ? (double)x[c] * sPrio[j + c] : (double)x[c];
        }
      }
      uint4 kk;
      kk.x = make_key(tot[0]);
      kk.y = make_key(tot[1]);
      kk.z = make_key(tot[2]);
      kk.w = make_key(tot[3]);
      *reinterpret_cast<uint4*>(K32 + i * ms + j) = kk;
    };
    // (i, j/4) advance incrementally: no division in the loop; two groups per iteration keep 2L loads in flight
    const int step_i = kThreads / m4, step_j = kThreads - step_i * m4;
    int i0 = tid / m4, j0 = tid - i0 * m4;
    for (int e4 = tid; e4 < total4; e4 += 2 * kThreads) {
      int i1 = i0 + step_i, j1 = j0 + step_j;
      if (j1 >= m4) {
        j1 -= m4;
        ++i1;
      }
      const int e4b = e4 + kThreads;
      const bool has_b = e4b < total4;
      float4 va[4], vb[4];
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (l < Leff) va[l] = ldg_hint4(win + (size_t)l * nm + (size_t)e4 * 4, pol_keep);
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (l < Leff && has_b) vb[l] = ldg_hint4(win + (size_t)l * nm + (size_t)e4b * 4, pol_keep);
      finish(e4, i0, j0 << 2, va);
      if (has_b) finish(e4b, i1, j1 << 2, vb);
      i0 = i1 + step_i;
      j0 = j1 + step_j;
      if (j0 >= m4) {
        j0 -= m4;
        ++i0;
      }
    }
  } else {
    for (int i = warp; i < n; i += kWarps)
      for (int j = lane; j < m; j += 32) {
        const double pr = kPrios ? sPrio[j] : 1.0;
        double tot = 0.0;
        for (int l = 0; l < Leff; ++l) tot += (double)win[(size_t)l * nm + i * m + j] * pr;
        K32[i * ms + j] = make_key(tot);
      }
  }
  if (vw.beta.ptr) {  // eager `beta` buffer field (off the hot path: the runners keep it lazy)
    const int64_t bb = sap_field_off(vw.beta, b, t_slot);
    for (int i = warp; i < n; i += kWarps)
      for (int j = lane; j < m; j += 32) {
        const double pr = kPrios ? sPrio[j] : 1.0;
        for (int l = 0; l < L; ++l)
          sap_store_real(vw.beta.ptr, bb + ((int64_t)i * m + j) * L + l, vw.beta.dtype,
                         l < Leff ? (double)win[(size_t)l * nm + i * m + j] * pr : 0.0);
      }
  }
  __syncthreads();

  if (SAP_DBG(p) & 256) return;  // timing ablation: stop after the key pass
  // exact float64 window sum (the reference's beta.sum(-1), :190) straight from global memory: only the rare
  // lists that cannot be certified use it
  auto tot64 = [&](int a, int j) {
    const double pr = kPrios ? sPrio[j] : 1.0;
    double s = 0.0;
    for (int l = 0; l < Leff; ++l) s += (double)win[(size_t)l * nm + a * m + j] * pr;
    return s;
  };

  // ------------------------------------------------------------------ 5. per-agent task lists (:198, :217)
  for (int base = 0; base < n * TPL; base += kThreads) {
    const int g = base + tid;
    const int i = g / TPL, s = g % TPL;
    const bool live = i < n;
    if (!__any_sync(SAP_FULL_MASK, live)) continue;  // whole warp beyond the last list
    const uint32_t* row = K32 + (live ? i : 0) * ms;
    uint32_t top[16];
    group_top16<TPL>(live ? m : 0, s, [&](int e) { return (row[e] << ib) | (imask - (uint32_t)e); }, top);
    if (live && s == 0) {
      if (!certified(top, K2, ib) && !SAP_DBG(p)) {
        qRows[atomicAdd(&sQ[0], 1)] = i;
      } else {
        // D: first M entries as they are (ties are proven ties, already in index-ascending order)
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < M) sD[i * M + t] = (IdxT)(imask - (top[t] & imask));
        // E: same values, but ties in index-DESCENDING order and, when the tie group of the K2-th entry extends
        // past the cut, its LARGEST indices
        uint32_t kk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) kk[t] = top[t] >> ib;
        bool ties = false, ext = false;
        uint32_t vstar = 0u;
        int pfx = K2;
#pragma unroll
        for (int t = 0; t < 15; ++t) {
          if (t < K2 - 1) ties = ties || (kk[t] == kk[t + 1]);
          if (t == K2 - 1) {
            ext = kk[t] == kk[t + 1];
            vstar = kk[t];
          }
        }
        if (ext) {
          pfx = 0;
#pragma unroll
          for (int t = 0; t < 16; ++t)
            if (t < K2 && kk[t] > vstar) pfx = t + 1;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < pfx) sE[i * K2 + t] = (IdxT)(imask - (top[t] & imask));
        if (ties) {  // reverse every run of equal keys inside the prefix (rare: duplicate values)
          int rs = 0;
          while (rs < pfx) {
            int re = rs + 1;
            const uint32_t kv = row[sE[i * K2 + rs]];
            while (re < pfx && row[sE[i * K2 + re]] == kv) ++re;
            for (int x = rs, y = re - 1; x < y; ++x, --y) {
              const IdxT tmp = sE[i * K2 + x];
              sE[i * K2 + x] = sE[i * K2 + y];
              sE[i * K2 + y] = tmp;
            }
            rs = re;
          }
        }
        if (ext)
          for (int j = m - 1; j >= 0 && pfx < K2; --j)
            if (row[j] == vstar) sE[i * K2 + pfx++] = (IdxT)j;
      }
    }
  }
  __syncthreads();
  // exact float64 redo of the lists that could not be certified (near-ties, negative benefits)
  for (int qi = warp; qi < sQ[0]; qi += kWarps) {
    const int i = qRows[qi];
    double vals[16];  // this lane's window sums (m <= 511 -> at most 16 per lane), fetched once from L2
#pragma unroll
    for (int c = 0; c < 16; ++c) vals[c] = (lane + 32 * c < m) ? tot64(i, lane + 32 * c) : 0.0;
    warp_select_cached(m, M, false, lane, vals, [&](int r, int j) { sD[i * M + r] = (IdxT)j; });
    warp_select_cached(m, K2, true, lane, vals, [&](int r, int j) { sE[i * K2 + r] = (IdxT)j; });
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {  // membership mask of D[i], used to filter the rivals' lists
    for (int w = 0; w < mw; ++w) sMask[i * mw + w] = 0u;
    for (int q = 0; q < M; ++q) {
      const int j = sD[i * M + q];
      sMask[i * mw + (j >> 5)] |= 1u << (j & 31);
    }
  }

  if (SAP_DBG(p) & 512) return;  // timing ablation: stop after the task lists
  // ------------------------------------------------------------------ 6. rivals (:203-206)
  for (int base = 0; base < n * TPL && !(SAP_DBG(p) & 4); base += kThreads) {
    const int g = base + tid;
    const int i = g / TPL, s = g % TPL;
    const bool live = i < n;
    if (!__any_sync(SAP_FULL_MASK, live)) continue;
    int dcol[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) dcol[q] = (live && q < M) ? (int)sD[i * M + q] : 0;
    uint32_t top[16];
    group_top16<TPL>(live ? n : 0, s,
                [&](int a) {
                  const uint32_t* row = K32 + a * ms;
                  uint32_t best = 0u;
#pragma unroll
                  for (int q = 0; q < 16; ++q)
                    if (q < M) best = max(best, row[dcol[q]]);
                  return a == i ? 0u : (best << ib) | (imask - (uint32_t)a);
                },
                top);
    if (live && s == 0) {
      if (!certified(top, N, ib) && !SAP_DBG(p)) {
        qNbr[atomicAdd(&sQ[1], 1)] = i;
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < N) sNbr[i * N + t] = (IdxT)(imask - (top[t] & imask));
      }
    }
  }
  __syncthreads();
  for (int qi = warp; qi < sQ[1]; qi += kWarps) {
    const int i = qNbr[qi];
    // exact scores once into scratch (the key tile is dead after the barrier above), then the exact selection
    double* sc = reinterpret_cast<double*>(K32) + (size_t)warp * n;
    for (int a = lane; a < n; a += 32) {
      double best = -INFINITY;
      for (int q = 0; q < M; ++q) best = fmax(best, tot64(a, sD[i * M + q]));
      sc[a] = (a == i) ? -INFINITY : best;
    }
    __syncwarp();
    warp_select(n, N, false, lane, [&](int a) { return sc[a]; }, [&](int r, int a) { sNbr[i * N + r] = (IdxT)a; });
    __syncwarp();
  }
  __syncthreads();

  if (SAP_DBG(p) & 128) return;  // timing ablation: stop after the rival lists
  // ------------------------------------------------------------------ 7. rivals' other top tasks (:212-217)
  // The M/2 best tasks of rival r outside D[i] under (value desc, idx desc) are the first M/2 entries of E[r]
  // not in D[i]; the reference lists them ascending, so they are stored reversed.
  for (int it = tid; it < n * N && !(SAP_DBG(p) & 64); it += kThreads) {
    const int i = it / N;
    const int r = sNbr[it];
    int c = 0;
    for (int e = 0; e < K2 && c < H; ++e) {
      const int j = sE[r * K2 + e];
      if (!((sMask[i * mw + (j >> 5)] >> (j & 31)) & 1u)) {
        sOther[(size_t)it * H + (H - 1 - c)] = (IdxT)j;
        ++c;
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ 8a. benefits of the window -> shared memory
  // Second read of the window (it is L2-resident: this CTA streamed it a few microseconds ago).  It lands over
  // the key tile and the scratch lists, which are dead now, already rounded to the obs dtype (:199-219 gathers).
  if (SAP_DBG(p) & 32) {
  } else if (vec4) {
    const int total4 = nm >> 2;
    auto put4 = [&](int l, int e4, const float4& v) {
      OutT o[4];
      if (kPrios) {
        const int j = (e4 * 4) % m;
        o[0] = to_out<OutT>((double)v.x * sPrio[j]);
        o[1] = to_out<OutT>((double)v.y * sPrio[j + 1]);
        o[2] = to_out<OutT>((double)v.z * sPrio[j + 2]);
        o[3] = to_out<OutT>((double)v.w * sPrio[j + 3]);
      } else {
        o[0] = to_out_f<OutT>(v.x);
        o[1] = to_out_f<OutT>(v.y);
        o[2] = to_out_f<OutT>(v.z);
        o[3] = to_out_f<OutT>(v.w);
      }
      OutT* dst = tile + (size_t)l * nm + (size_t)e4 * 4;
      if (sizeof(OutT) == 2) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(&o[0]);
      else *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&o[0]);
    };
    // 2 x L independent 128-bit loads in flight per thread
    for (int e4 = tid; e4 < total4; e4 += 2 * kThreads) {
      const int e4b = e4 + kThreads;
      const bool has_b = e4b < total4;
      float4 va[4], vb[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        va[l] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[l] = va[l];
        if (l < Leff) va[l] = ldg_hint4(win + (size_t)l * nm + (size_t)e4 * 4, pol_drop);
        if (l < Leff && has_b) vb[l] = ldg_hint4(win + (size_t)l * nm + (size_t)e4b * 4, pol_drop);
      }
      if (C::kCommon) {
        // L = 3, fp16: planes 0 and 1 interleaved as half2 per (agent, task), plane 2 behind them: the gather reads
        // a pair's three benefits with two loads (one 32-bit, one 16-bit) instead of three
        auto put_pair = [&](int e4, const float4 (&v)[4]) {
          __half2 h[4] = {__floats2half2_rn(v[0].x, v[1].x), __floats2half2_rn(v[0].y, v[1].y),
                          __floats2half2_rn(v[0].z, v[1].z), __floats2half2_rn(v[0].w, v[1].w)};
          *reinterpret_cast<uint4*>(reinterpret_cast<__half2*>(tile) + (size_t)e4 * 4) = *reinterpret_cast<const uint4*>(h);
          __half2 g2[2] = {__floats2half2_rn(v[2].x, v[2].y), __floats2half2_rn(v[2].z, v[2].w)};
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(tile) + 2 * (size_t)nm + (size_t)e4 * 4) =
              *reinterpret_cast<const uint2*>(g2);
        };
        put_pair(e4, va);
        if (has_b) put_pair(e4b, vb);
      } else {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (l < L) {
            put4(l, e4, va[l]);
            if (has_b) put4(l, e4b, vb[l]);
          }
      }
    }
  } else {
    for (int e = tid; e < L * nm; e += kThreads) {
      const int l = e / nm, x = e - l * nm;
      const double pr = kPrios ? sPrio[x % m] : 1.0;
      const OutT o = to_out<OutT>(l < Leff ? (double)win[e] * pr : 0.0);
      if (C::kCommon) tile[l < 2 ? 2 * x + l : 2 * nm + x] = o;  // same paired layout as the vector path
      else tile[e] = o;
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ 8b. gather rows into shared memory, store 128-bit
  // The key tile is dead now; it becomes the staging area.  Rows are staged at the same 16-byte phase as their
  // global destination so the aligned interior goes out as uint4.
  unsigned char* stage = smem + f.stage;
  const size_t row_bytes = sizeof(OutT) * (size_t)obs_size;
  const int rpp = f.rows_per_pass;
  const bool ain_flat = ain && ain_row == obs_size && ((reinterpret_cast<uintptr_t>(ain) & 15) == 0);
  if (p.top_out)  // the top-M lists are final: one coalesced pass instead of M lanes of every row
    for (int e = tid; e < n * M; e += kThreads) p.top_out[(size_t)b * n * M + e] = sD[e];
  for (int r0 = 0; r0 < n; r0 += rpp) {
    const int rows = min(rpp, n - r0);
    unsigned char* gdst = reinterpret_cast<unsigned char*>(obs_out) + (size_t)r0 * row_bytes;
    const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15);
    if (warp < rows && !(SAP_DBG(p) & 8)) {
      const int i = r0 + warp;
      OutT* srow = reinterpret_cast<OutT*>(stage + phase + (size_t)warp * row_bytes);
      const IdxT* myD = sD + i * M;
      const IdxT* myN = sNbr + i * N;
      const IdxT* myO = sOther + (size_t)i * N * H;
      // one pair (agent a, task j) per lane and slot: L benefits each.  Branch-free decode through the LUT; kGB
      // independent pairs per lane are loaded before any is stored so their shared-memory latencies overlap.
      constexpr int kGB = 5;
      for (int pp0 = lane; pp0 < npairs; pp0 += 32 * kGB) {
        uint32_t code[kGB];
#pragma unroll
        for (int g = 0; g < kGB; ++g) code[g] = (pp0 + 32 * g < npairs) ? (uint32_t)sLut[pp0 + 32 * g] : 0xff00u;
        int a_r[kGB], j_d[kGB], j_o[kGB];
#pragma unroll
        for (int g = 0; g < kGB; ++g) {
          const uint32_t ps = code[g] >> 8, qs = code[g] & 0xffu;
          a_r[g] = myN[min(ps, (uint32_t)(N - 1))];
          j_d[g] = myD[min(qs, (uint32_t)(M - 1))];
          j_o[g] = myO[min(qs & 0x7fu, (uint32_t)(N * H - 1))];
        }
        OutT val[kGB][4];
#pragma unroll
        for (int g = 0; g < kGB; ++g) {
          const uint32_t ps = code[g] >> 8, qs = code[g] & 0xffu;
          const int a = ps == 0xffu ? i : a_r[g];
          const int j = (qs & 0x80u) ? j_o[g] : j_d[g];
          if (C::kCommon) {
            const int x = a * m + j;
            const uint32_t p01 = reinterpret_cast<const uint32_t*>(tile)[x];
            const __half2 h01 = *reinterpret_cast<const __half2*>(&p01);
            val[g][0] = *reinterpret_cast<const OutT*>(&h01.x);
            val[g][1] = *reinterpret_cast<const OutT*>(&h01.y);
            val[g][2] = tile[2 * nm + x];
          } else {
            const OutT* src = tile + a * m + j;
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (l < L) val[g][l] = src[(size_t)l * nm];
          }
        }
#pragma unroll
        for (int g = 0; g < kGB; ++g) {
          const int pp = pp0 + 32 * g;
          if (pp < npairs) {
            OutT* dst = srow + pp * L;
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (l < L) dst[l] = val[g][l];
          }
        }
      }
      // "is my previous task among my top-M" flags (:222)
      const int pv = sPrev[i];  // shared-memory copy: no global load per row
      for (int q = lane; q < M; q += 32) srow[npairs * L + q] = to_out_f<OutT>((int)myD[q] == pv ? 1.f : 0.f);
    }
    // generic-proxy writes of the staged rows must be visible to the async proxy (TMA) before the barrier
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t bytes = (SAP_DBG(p) & 16) ? 0 : (size_t)rows * row_bytes;
    const unsigned char* ssrc = stage + phase;
    if (phase == 0 && (bytes & 15) == 0 && (!ain || ain_flat)) {
      // aligned block.  The obs rows leave shared memory with ONE TMA bulk store (cp.async.bulk.global.shared::cta)
      // issued by thread 0; meanwhile all threads widen the same rows to fp32 for the agent network (128-bit stores).
      if (tid == 0 && bytes > 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                     "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"((uint32_t)bytes), "l"(pol_drop)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      float* adst = ain ? ain + (size_t)r0 * obs_size : nullptr;  // rows are contiguous when ain_row == obs_size
      const int chunks = adst ? (int)(bytes >> 4) : 0;
#pragma unroll 4
      for (int c = tid; c < chunks; c += kThreads) {
        const uint4 v = *reinterpret_cast<const uint4*>(ssrc + (size_t)c * 16);
        if (sizeof(OutT) == 2) {
          const __half2* h = reinterpret_cast<const __half2*>(&v);
          const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]),
                       f3 = __half22float2(h[3]);
          float* o = adst + (size_t)c * 8;
          const float4 lo4 = make_float4(f0.x, f0.y, f1.x, f1.y), hi4 = make_float4(f2.x, f2.y, f3.x, f3.y);
          stg_hint4(o, *reinterpret_cast<const uint4*>(&lo4), pol_drop);
          stg_hint4(o + 4, *reinterpret_cast<const uint4*>(&hi4), pol_drop);
        } else {
          stg_hint4(adst + (size_t)c * 4, v, pol_drop);
        }
      }
      // the staging rows may be overwritten once the bulk store has finished READING them
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
      size_t head = (16 - phase) & 15;
      if (head > bytes) head = bytes;
      const size_t body = (bytes - head) & ~(size_t)15;
      for (size_t o = tid * sizeof(OutT); o < head; o += kThreads * sizeof(OutT))
        *reinterpret_cast<OutT*>(gdst + o) = *reinterpret_cast<const OutT*>(ssrc + o);
      for (size_t o = head + (size_t)tid * 16; o < head + body; o += (size_t)kThreads * 16)
        *reinterpret_cast<uint4*>(gdst + o) = *reinterpret_cast<const uint4*>(ssrc + o);
      for (size_t o = head + body + tid * sizeof(OutT); o < bytes; o += kThreads * sizeof(OutT))
        *reinterpret_cast<OutT*>(gdst + o) = *reinterpret_cast<const OutT*>(ssrc + o);
      if (ain) {  // fp32 copy for the agent network: float(obs rounded to the buffer dtype)
        for (int r = warp; r < rows; r += kWarps) {
          const OutT* srow = reinterpret_cast<const OutT*>(ssrc + (size_t)r * row_bytes);
          float* drow = ain + (int64_t)(r0 + r) * ain_row;
          for (int c = lane; c < obs_size; c += 32) drow[c] = out_to_f(srow[c]);
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores complete before exit
}

// One CTA per environment; the hardware block scheduler balances the (slightly uneven) per-env durations better
// than a static persistent partition did (measured: 0.95 ms vs 1.18 ms at 4096 x 100 x 100).
template <typename C>
__global__ void __launch_bounds__(kThreads, C::kMinBlocks) sap_real_fast_kernel(RealParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  process_env<C>(p, blockIdx.x, smem);
}

template <typename C>
int launch_fast(RealParams& p, void* stream, size_t bytes) {
  static thread_local bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(sap_real_fast_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(sap_real_fast_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) {
      sap_set_error("sap_real_fast: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  sap_real_fast_kernel<C><<<p.d.B, kThreads, bytes, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_fast_kernel");
  return SAP_OK;
}

}  // namespace

int sap_real_fast_try(RealParams& p, void* stream, int* handled, bool gen1_only) {
  *handled = 0;
  if (!gen1_only) {  // the second-generation kernel takes the shipped configuration at 64 < n <= 128
    const int rc = sap_real_fast2_try(p, stream, handled);
    if (rc != SAP_OK || *handled) return rc;
  }
  const SapEnvDims& d = p.d;
  const int H = d.M / 2;
  const bool half_out = p.view.obs.dtype == SAP_F16;
  // eligibility: lists fit the 16-wide networks, indices fit the packed words, tiles fit shared memory
  if (d.M + H + 1 > 16 || d.N + 1 > 16 || d.L > 4) return SAP_OK;
  if (d.n > 511 || d.m > 511 || d.N * H > 127) return SAP_OK;
  const bool common = d.M == 10 && d.N == 10 && d.L == 3 && half_out && !p.prios;
  const bool idx8 = common && d.n <= 256 && d.m <= 256;
  const FastLayout f = fast_layout(d, half_out ? 2 : 4, idx8 ? 1 : 2, p.prios != nullptr);
  if (f.total > kMaxSmem || f.ms * 4 < kWarps * 8) return SAP_OK;
  if (p.view.agent_in.ptr && p.view.agent_in.dtype != SAP_F32) {
    sap_set_error("sap_real: agent_in must be f32");
    return SAP_E_DTYPE;
  }
  *handled = 1;
  if (common) {
    if (idx8 && d.n == 50 && d.m == 50 && sap_real_path_override() != SAP_REAL_PATH_FAST_RUNTIME_SHAPE)
      return launch_fast<Cfg<__half, uint8_t, false, true, 4, 50>>(p, stream, f.total);
    if (idx8 && d.n <= 64) return launch_fast<Cfg<__half, uint8_t, false, true, 4>>(p, stream, f.total);
    if (idx8) return launch_fast<Cfg<__half, uint8_t, false, true>>(p, stream, f.total);
    return launch_fast<Cfg<__half, uint16_t, false, true>>(p, stream, f.total);
  }
  if (half_out) {
    if (p.prios) return launch_fast<Cfg<__half, uint16_t, true, false>>(p, stream, f.total);
    return launch_fast<Cfg<__half, uint16_t, false, false>>(p, stream, f.total);
  }
  if (p.prios) return launch_fast<Cfg<float, uint16_t, true, false>>(p, stream, f.total);
  return launch_fast<Cfg<float, uint16_t, false, false>>(p, stream, f.total);
}
