// RealConstellationEnv step + observation build for the reference's shipped configuration (M = N = 10, L = 3, fp16
// scheme, no priorities) at 64 < n <= 128 agents, m <= 128 tasks: one CTA per environment, second generation.
//
// Same contract and same bytes as sap_real_fast.cu / sap_real.cu (reference: envs/real_constellation_env.py step
// :135-175, beta_hat :282-328, _build_obs :177-230).  What changed against sap_real_fast.cu is where the shared-memory
// wavefronts and the instructions went (profiles/r01_ncu_real_fast.txt: rival scoring 44 % and the gather 37 % of the
// wavefronts, half of them bank conflicts):
//   * the selection keys live TRANSPOSED, KT[task][agent] (pitch 128 words, 16-byte column groups XOR-swizzled by
//     task & 3).  The rival score of agent a for agent i, max_q key[a][D_i[q]], is then a max over 10 ROWS: one
//     128-bit load covers four rivals, VIMNMX3 folds two rows per instruction (10 LDS.32 + 10 adds + 5 max per
//     score before; 2.5 LDS.128 + 5 max per score now), and the two lists that share a quarter-warp read opposite
//     128-byte halves of their rows, so the loads are conflict-free whatever the tasks are;
//   * the key pass transposes 4 x 4 key blocks in registers (four shuffles) and stores them with one 128-bit store;
//   * the gather works on 30-byte SLOTS of the observation row: lane = (rival, half of my top tasks | its other
//     tasks), i.e. 32 slots = 32 lanes per agent row, five (agent, task) pairs each.  A lane reads one rival index
//     and five task indices instead of one LUT word and three indices PER PAIR, and writes its 15 halves as
//     7 words + 1 half;
//   * gather and stores run per warp on a private staging row, with no block barrier inside the phase;
//   * the bench shape 100 x 100 has its own instantiation with n and m compiled in (kFixed).
// Lists are certified exactly as before (keys are a monotone fixed-point image of the float64 window sum plus an
// "inexact" bit; an uncertified list is redone by the exact float64 warp selection), so results never depend on the
// key resolution.
#include "sap_real.cuh"
#include "sap_sortnet.cuh"
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kM = 10, kN = 10, kL = 3, kH = 5, kK2 = 15;
constexpr int kPairs = kM + kN * kM + kN * kH;  // 160 (agent, task) pairs per observation row
constexpr int kObs = kPairs * kL + kM;          // 490
constexpr int kRowBytes = kObs * 2;             // 980
constexpr int kKP = 128;                        // key tile pitch in words
constexpr int kStagePitch = 1008;                // per-warp staging row: 980 bytes + up to 12 of misalignment, 16-byte pitch
constexpr size_t kMaxSmem = 227 * 1024 - 1024;

#ifdef SAP_ABLATE
// profiling builds only: SAP_DEBUG_SKIP_REDO == k ends the kernel after phase k (timing ablation); == 99 makes thread 0 of
// every CTA record %globaltimer at the phase boundaries into p.scratch[b * 16 + k] (phase timeline, profiles/phase_timeline.py)
__device__ __forceinline__ unsigned long long sap_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SAP_TS(k)                                                                                        \
  if (p.debug_skip_redo == 99 && threadIdx.x == 0 && p.scratch)                                          \
    reinterpret_cast<unsigned long long*>(p.scratch)[(size_t)blockIdx.x * 16 + (k)] = sap_globaltimer();
#define SAP_STOP_AFTER(k)                      \
  SAP_TS(k)                                    \
  if (p.debug_skip_redo == (k)) return;
#else
#define SAP_TS(k)
#define SAP_STOP_AFTER(k)
#endif

#define SAP_CE(a, b)            \
  {                             \
    uint32_t hi__ = max(a, b);  \
    b = min(a, b);              \
    a = hi__;                   \
  }

struct F2Layout {
  uint32_t D, nbr, other, pv;                // live through the whole kernel
  uint32_t KT, E, dmask, cnt, red, queue, stage1;  // overlay 1: list building
  uint32_t t01, t2, stage;                   // overlay 2: gather
  uint32_t total;
  int p01, p2;                               // pitches of the two tile planes (words / halves)
};

__host__ __device__ inline uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline F2Layout f2_layout(int n, int m) {
  F2Layout f;
  uint32_t off = 0;
  f.D = off;     off = up16(off + n * kM);
  f.nbr = off;   off = up16(off + n * kN);
  f.other = off; off = up16(off + n * kN * kH);
  f.pv = off;    off = up16(off + n);
  const uint32_t base = (off + 127u) & ~127u;
  off = base;
  f.KT = off;    off += 4u * ((m + 7) & ~7) * kKP;  // rows m .. 8 ceil(m / 8) hold zero keys (list padding)
  f.cnt = off;   off = up16(off + 4u * m);
  f.red = off;   off += 8u * 80;
  f.queue = off; off = up16(off + 4u * (2 * n + 4));
  f.stage1 = off;                                    // key pass only: per-warp staging tiles, dead before E is written
  f.E = off;     off += 16u * n;
  f.dmask = off; off += 16u * n;
  const uint32_t end1s = f.stage1 + (uint32_t)kWarps * 16u * m;
  const uint32_t end1 = off > end1s ? off : end1s;
  f.p01 = m;
  f.p2 = m;
  off = base;
  f.t01 = off;   off = up16(off + 4u * n * f.p01);
  f.t2 = off;    off = up16(off + 2u * n * f.p2);
  f.stage = off; off = up16(off + (uint32_t)kWarps * kStagePitch);
  const uint32_t end2 = off;
  f.total = end1 > end2 ? end1 : end2;
  return f;
}

__device__ __forceinline__ bool certified(const uint32_t (&top)[16], int need, int ib) {
  bool ok = true;
#pragma unroll
  for (int t = 0; t < 15; ++t) {
    const uint32_t a = top[t] >> ib, b = top[t + 1] >> ib;
    if (t < need) ok = ok && (a > b || (a == b && !(a & 1u)));
  }
  return ok;
}

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
__device__ __forceinline__ float ldg_hint1(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_hint4(void* p, const uint4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return max(max(a, b), c); }

// sap_rollout_step: the env's CTA selects the actions of its own n agents before stepping (the selector's launch and the
// actions round trip through HBM between two launches disappear).  Warp c handles rows [32 c, 32 c + 32) of the env; the
// Philox draws are keyed by the global row, the env's step counter and the episode counter exactly as in the stand-alone
// selector kernel, so the two produce the same actions.  Ends with a barrier: the CTA reads the actions back from global
// memory (written by its own threads).
template <int kCtaWarps>
__device__ __forceinline__ void select_own_actions(const RealParams& p, int b, int n, int T, int tid) {
  if (p.k[b] < T) {
    const int lane = tid & 31;
    for (int c = tid >> 5; c * 32 < n; c += kCtaWarps) {
      const int64_t base = (int64_t)b * n + c * 32;
      const int nrows = min(32, n - c * 32);
      const int a = sap_classic_select_rows(p.sel, base, nrows, p.sel_vec4, lane);
      if (lane < nrows) p.sel.out[base + lane] = (int64_t)a;
    }
  }
  __syncthreads();
}

// kSelect: the sap_rollout_step instantiation (selection first); a separate instantiation so that the plain step keeps
// its register allocation
// kFixed: n = m = kFixed known at compile time (the bench shape 100 x 100: every tile pitch, loop bound and layout offset
// folds into an immediate); 0 = any eligible shape
template <bool kSelect, int kFixed>
__global__ void __launch_bounds__(kThreads, 3) sap_real_fast2_kernel(RealParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  const SapEnvDims d = p.d;
  const int n = kFixed ? kFixed : d.n, m = kFixed ? kFixed : d.m, T = d.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nm = n * m;
  const F2Layout f = f2_layout(n, m);
  uint8_t* sD = smem + f.D;                                     // [n][10] top-M tasks (value desc, idx asc)
  uint8_t* sNbr = smem + f.nbr;                                 // [n][10] rivals
  uint8_t* sOther = smem + f.other;                             // [n][10][5] rivals' other tasks (ascending value)
  uint8_t* sPrev = smem + f.pv;                                 // [n] previous task of every agent (0xff: none, obs_only)
  uint32_t* KT = reinterpret_cast<uint32_t*>(smem + f.KT);      // [m][128] keys, column group g of row j at g ^ (j & 3)
  uint8_t* sE = smem + f.E;                                     // [n][16] top-15 tasks (value desc, idx DESC)
  uint32_t* sMask = reinterpret_cast<uint32_t*>(smem + f.dmask);  // [n][4] membership bits of D[i]
  int32_t* sCnt = reinterpret_cast<int32_t*>(smem + f.cnt);
  double* sRed = reinterpret_cast<double*>(smem + f.red);
  int32_t* sQ = reinterpret_cast<int32_t*>(smem + f.queue);     // [0] = task lists to redo, [1] = rival lists, then ids
  int32_t* qRows = sQ + 4;
  int32_t* qNbr = sQ + 4 + n;
  uint32_t* t01 = reinterpret_cast<uint32_t*>(smem + f.t01);    // [n][p01] half2 {plane 0, plane 1}
  __half* t2 = reinterpret_cast<__half*>(smem + f.t2);          // [n][p2]  plane 2
  unsigned char* stage = smem + f.stage;

  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  const size_t env_plane0 = d.shared_planes ? (size_t)0 : (size_t)b * T;
  const float* env_planes = p.planes + env_plane0 * nm;
  const SapBatchView& vw = p.view;

  // reward-phase loads that do not depend on each other are issued up front (two DRAM round trips instead of five)
  // obs_only: the observation of slot k + 1 ahead of the step (it does not depend on the actions of step k apart from the
  // "previous task in my top-M" flags, which sap_real_step_after_obs sets): no rewards, no counters, flags = 0
  const bool stepping = !p.is_reset && !p.obs_only;
  SAP_TS(0)
  if (kSelect && stepping) select_own_actions<kWarps>(p, b, n, T, tid);
  int a_mine = 0, pv_mine = 0;
  if (stepping && tid < n) {
    a_mine = min(max((int)p.actions[(size_t)b * n + tid], 0), m - 1);
    pv_mine = p.prev[(size_t)b * n + tid];
  }
  // what the flag columns of the new rows compare with (:222): the task just taken, i at reset, nothing when the
  // observation is built ahead of the step (sap_real_step_after_obs sets those flags)
  if (tid < n) sPrev[tid] = (uint8_t)(p.obs_only ? 0xff : (p.is_reset ? tid : a_mine));
  const int k_old = p.is_reset ? -1 : p.k[b];
  if (!p.is_reset && k_old >= T) return;
  const int k_new = k_old + 1;
  const bool done = k_new >= T;
  const int Leff = done ? 0 : min(kL, T - k_new);
  const float* win = env_planes + (size_t)k_new * nm;
  float2 pst[kL];
#pragma unroll
  for (int l = 0; l < kL; ++l)
    if (l < Leff) pst[l] = __ldg(reinterpret_cast<const float2*>(p.plane_stats) + (env_plane0 + k_new + l));
  // L2 look-ahead: the CTA that will run `lookahead` blocks after this one reads the same window of its own env (all
  // envs of a batch step in lockstep; a wrong guess only wastes the prefetch).  By then the window sits in L2, so the
  // key pass of that CTA pays L2 latency instead of DRAM latency.
  if (tid == 0 && p.lookahead > 0 && !d.shared_planes && b + p.lookahead < d.B && Leff > 0) {
    const float* nxt = p.planes + ((size_t)(b + p.lookahead) * T + k_new) * nm;
    for (int l = 0; l < Leff; ++l)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt + (size_t)l * nm), "r"((uint32_t)(nm * 4)) : "memory");
  }

  if (tid == 0) {
    sQ[0] = 0;
    sQ[1] = 0;
  }
  for (int j = tid; j < m; j += kThreads) sCnt[j] = 0;
  __syncthreads();

  // ------------------------------------------------------------------ 1. rewards at the old window (:135-164)
  if (stepping) {
    float mine[kL] = {0.f, 0.f, 0.f};
    if (tid < n) {
#pragma unroll
      for (int l = 0; l < kL; ++l)
        if (k_old + l < T) mine[l] = __ldg(env_planes + ((size_t)(k_old + l) * n + tid) * m + a_mine);
      atomicAdd(&sCnt[a_mine], 1);
    }
    __syncthreads();
    double local_ret = 0.0;
    if (tid < n) {
      const int i = tid, a = a_mine, pv = pv_mine;
      double sum = 0.0;
#pragma unroll
      for (int l = 0; l < kL; ++l)
        if (k_old + l < T) sum += (double)mine[l];
      const double b0 = (double)mine[0];
      const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);
      const double bh = b0 - p.lambda_ * (pen * (sum > 1e-12 ? 1.0 : 0.0));
      const double r = bh > 0.0 ? bh / (double)sCnt[a] : bh;
      local_ret = r;
      if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
      if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
      p.prev[(size_t)b * n + i] = a;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local_ret += __shfl_xor_sync(SAP_FULL_MASK, local_ret, off);
    if (lane == 0) sRed[warp] = local_ret;
    if (vw.actions_onehot.ptr) {
      const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
      const bool v16 = vw.actions_onehot.dtype == SAP_I16 &&
                       ((reinterpret_cast<uintptr_t>(vw.actions_onehot.ptr) + 2 * base) & 7) == 0;
      for (int i = warp; i < n; i += kWarps) {
        const int a = min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
        if (v16) {  // four int16 flags per store (m % 4 == 0)
          uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<int16_t*>(vw.actions_onehot.ptr) + base + (int64_t)i * m);
          for (int j4 = lane; j4 < m >> 2; j4 += 32) {
            const int d = a - 4 * j4;
            o[j4] = make_uint2(d == 0 ? 1u : (d == 1 ? 0x10000u : 0u), d == 2 ? 1u : (d == 3 ? 0x10000u : 0u));
          }
        } else {
          for (int j = lane; j < m; j += 32)
            sap_store_int(vw.actions_onehot.ptr, base + (int64_t)i * m + j, vw.actions_onehot.dtype, a == j ? 1 : 0);
        }
      }
    }
    if (p.counts_out)
      for (int j = tid; j < m; j += kThreads) p.counts_out[(size_t)b * m + j] = sCnt[j];
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kWarps; ++w) t += sRed[w];
      atomicAdd(&p.ep_return[b], t);
      p.k[b] = k_new;
      if (vw.terminated.ptr)
        sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, done);
    }
  } else if (p.is_reset) {
    for (int i = tid; i < n; i += kThreads) p.prev[(size_t)b * n + i] = i;
    if (tid == 0) {
      p.k[b] = 0;
      p.ep_return[b] = 0.0;
    }
  }

  // ------------------------------------------------------------------ 2. pre-transition scalars of slot k_new
  const int t_slot = k_new;
  if (tid == 0 && vw.filled.ptr && !p.obs_only)
    sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, t_slot), vw.filled.dtype, 1);
  if (vw.prev_assigns.ptr && tid < n && !p.obs_only)
    sap_store_int(vw.prev_assigns.ptr, sap_field_off(vw.prev_assigns, b, t_slot) + tid, vw.prev_assigns.dtype,
                  p.is_reset ? tid : a_mine);
  if (vw.avail_actions.ptr) {  // eager field: every action available (:267-273)
    const int64_t base = sap_field_off(vw.avail_actions, b, t_slot);
    if (vw.avail_actions.dtype == SAP_U8 && ((reinterpret_cast<uintptr_t>(vw.avail_actions.ptr) + base) & 15) == 0) {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(vw.avail_actions.ptr) + base);  // nm % 16 == 0
      for (int e = tid; e < nm >> 4; e += kThreads) o[e] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    } else {
      for (int e = tid; e < nm; e += kThreads) sap_store_int(vw.avail_actions.ptr, base + e, vw.avail_actions.dtype, 1);
    }
  }
  __half* obs_out = reinterpret_cast<__half*>(vw.obs.ptr) + sap_field_off(vw.obs, b, t_slot);
  // agent-input staging rows of this env: fp32 (widened here) or fp16 (the same bytes, stored a second time)
  const bool ain_half = vw.agent_in.ptr && vw.agent_in.dtype == SAP_F16;
  float* ain = vw.agent_in.ptr && !ain_half ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride
                                            : nullptr;
  __half* ain16 = ain_half ? reinterpret_cast<__half*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride : nullptr;
  // fp16 staging rows are either packed (pitch = obs size: the same 128-bit pieces once more) or padded to a pitch
  // that a tensor-core GEMM can read with 128-bit loads (496 halves for 490): then the rows are copied word by word
  const int ain16_pitch = ain_half ? (int)vw.agent_in.t_stride : 0;
  const bool ain16_bulk = ain_half && ain16_pitch == kObs;
  // fp32 rows are packed (pitch = obs size: 128-bit stores across row boundaries) or pitched (the MAC appends last-action /
  // agent-id columns behind the observation: float2 stores row by row)
  const int ain32_pitch = ain ? (int)vw.agent_in.t_stride : 0;
  const bool ain32_flat = ain && ain32_pitch == kObs;
  if (done) {  // :226-228
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* o4 = reinterpret_cast<uint4*>(obs_out);
    for (int e = tid; e < (n * kRowBytes) >> 4; e += kThreads) o4[e] = z;
    if (ain32_flat) {
      uint4* a4 = reinterpret_cast<uint4*>(ain);
      for (int e = tid; e < (n * kObs * 4) >> 4; e += kThreads) a4[e] = z;
    } else if (ain) {
      for (int i = warp; i < n; i += kWarps)
        for (int c = lane; c < kObs; c += 32) ain[(size_t)i * ain32_pitch + c] = 0.f;
    }
    if (ain16) {  // packed or padded rows: only the observation columns (the MAC owns whatever follows them)
      uint32_t* a2 = reinterpret_cast<uint32_t*>(ain16);
      for (int i = warp; i < n; i += kWarps)
        for (int w = lane; w < kRowBytes / 4; w += 32) a2[(size_t)i * (ain16_pitch >> 1) + w] = 0u;
    }
    if (vw.beta.ptr) {
      const int64_t bb = sap_field_off(vw.beta, b, t_slot);
      for (int e = tid; e < nm * kL; e += kThreads) sap_store_real(vw.beta.ptr, bb + e, vw.beta.dtype, 0.0);
    }
    return;
  }

  SAP_STOP_AFTER(1)
  // ------------------------------------------------------------------ 3. key scale from the per-plane {min, max}
  double k_lo, k_scale;
  bool k_nonneg;
  {
    float vmin = INFINITY;
    double vabs = 0.0;  // float64: a true bound of the sum of the per-plane maxima
#pragma unroll
    for (int l = 0; l < kL; ++l)
      if (l < Leff) {
        vmin = fminf(vmin, pst[l].x);
        vabs += (double)fmaxf(fabsf(pst[l].x), fabsf(pst[l].y));
      }
    k_nonneg = vmin >= 0.f;
    const double hi = vabs * 1.0000001;  // >= every |window sum|
    int e2 = 0;
    if (hi > 0.0) (void)frexp(hi, &e2);  // hi < 2^e2
    const int ib0 = 32 - __clz(max(n, m));
    const int vb1 = 31 - ib0;
    k_lo = k_nonneg ? 0.0 : -ldexp(1.0, e2);
    k_scale = hi > 0.0 ? ldexp(1.0, vb1 - e2 - (k_nonneg ? 0 : 1)) : 1.0;
  }
  const int ib = 32 - __clz(max(n, m));
  const uint32_t imask = (1u << ib) - 1u;
  const uint32_t fixed_max = (1u << (31 - ib)) - 1u;
  const uint32_t kmul = 2u << ib, kflag = 1u << ib;
  // Key of a float64 window sum: fixed-point floor(sum * 2^s) in the high bits, then one "conversion was not exact" bit,
  // then ib zero bits (KT holds the keys already shifted into the position they have in a packed (key | ~index) word).
  // Monotone in the sum; two equal keys with the flag clear are equal sums (same map as sap_real_fast.cu).
  auto make_key_generic = [&](double tot) -> uint32_t {
    const double y = (tot - k_lo) * k_scale;
    const double t = y + 4503599627370496.0;
    const double r = t - 4503599627370496.0;
    uint32_t fx = (uint32_t)__double2loint(t) - (r > y ? 1u : 0u);
    bool inexact = !k_nonneg || (r != y);
    if (fx > fixed_max) {
      fx = fixed_max;
      inexact = true;
    }
    return ((fx << 1) | (inexact ? 1u : 0u)) << ib;
  };
  // Non-negative window (the usual case): 0 <= y < 2^(31 - ib) by the choice of the scale, so no origin and no clamp;
  // adding 2^52 with round-toward-minus-infinity IS the floor (no compare-and-correct step).
  auto make_key_fast = [&](double tot) -> uint32_t {
    const double y = tot * k_scale;
    const double t = __dadd_rd(y, 4503599627370496.0);
    const double r = t - 4503599627370496.0;
    return (uint32_t)__double2loint(t) * kmul + (r != y ? kflag : 0u);
  };
  // key of (agent a, task j) in the transposed, swizzled tile
  auto key_at = [&](int a, int j) -> uint32_t { return KT[j * kKP + ((((a >> 2) ^ (j & 3)) << 2) | (a & 3))]; };

  // ------------------------------------------------------------------ 4. key pass: window -> KT
  // A warp takes a group of 4 consecutive agents (4 I .. 4 I + 3): their rows are ONE contiguous 4 m float run per
  // plane, read with fully coalesced 128-bit loads, all 3 x 4 loads of a lane in flight at once (the pass is bound
  // by how many bytes an SM keeps in flight, not by arithmetic: row-wise 128-byte pieces of 4 separate rows were
  // 0.045 ms slower at 4096 x 100 x 100).  The keys go through a per-warp staging tile [4][m] so that they can be
  // re-read per TASK (4 agents of one task = one 128-bit store into the transposed tile).
  {
    const int m4 = m >> 2, n4 = n >> 2, m8 = (m + 7) & ~7;
    // zero keys for the agents n .. 127 of every task and for the padding tasks m .. m8 - 1: the lists below read them
    // as (key 0 | ~index) words, which rank behind every real entry, so no list needs a bounds test
    for (int row = tid >> 3; row < m8; row += kThreads / 8) {
      const int g0 = row < m ? n4 : 0;
      for (int g = g0 + (tid & 7); g < kKP / 4; g += 8)
        *reinterpret_cast<uint4*>(KT + row * kKP + ((g ^ (row & 3)) << 2)) = make_uint4(0u, 0u, 0u, 0u);
    }
    uint32_t* Sw = reinterpret_cast<uint32_t*>(smem + f.stage1) + warp * 4 * m;  // this warp's staging tile
    const uint32_t inv_m4 = (uint32_t)((0x100000000ull + (uint32_t)m4 - 1u) / (uint32_t)m4);  // x / m4 = umulhi(x, inv)
    auto keys_of = [&](auto tag, const float4 (&v)[kL]) -> uint4 {
      constexpr bool kFast = decltype(tag)::value;
      double tot[4] = {(double)v[0].x, (double)v[0].y, (double)v[0].z, (double)v[0].w};  // 0.0 + x == x
#pragma unroll
      for (int l = 1; l < kL; ++l) {  // (planes past the horizon were loaded as zeros)
        tot[0] += (double)v[l].x;
        tot[1] += (double)v[l].y;
        tot[2] += (double)v[l].z;
        tot[3] += (double)v[l].w;
      }
      uint4 kk;
      kk.x = kFast ? make_key_fast(tot[0]) : make_key_generic(tot[0]);
      kk.y = kFast ? make_key_fast(tot[1]) : make_key_generic(tot[1]);
      kk.z = kFast ? make_key_fast(tot[2]) : make_key_generic(tot[2]);
      kk.w = kFast ? make_key_fast(tot[3]) : make_key_generic(tot[3]);
      return kk;
    };
    // Whole agent group 4 I .. 4 I + 3 (the hot unit).  Its four rows are ONE run of 4 m floats per plane, so position
    // fpos = lane + 32 t is simply float4 number fpos of that run - and of the staging tile: every address below is a
    // per-lane base plus a compile-time offset.
    const uint32_t* sw_rd = Sw + lane;            // staging read: task lane + 32 u of row r at sw_rd[r m + 32 u]
    uint32_t* kt_wr = KT + lane * kKP;            // transposed tile: row of task lane + 32 u at kt_wr + 32 u kKP
    auto unit_rows = [&](auto tag, int I) {
      constexpr bool kFast = decltype(tag)::value;  // full window of non-negative benefits: no predicates on the loads
      const float* src = win + (4 * I) * m + 4 * lane;
      float4 v[4][kL];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (lane + 32 * t < m) {  // m float4 positions (4 rows x m / 4); divergent only in the last one
#pragma unroll
          for (int l = 0; l < kL; ++l) {
            if (kFast) {
              v[t][l] = ldg_hint4(src + l * nm + 128 * t, pol_keep);
            } else {
              v[t][l] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (l < Leff) v[t][l] = ldg_hint4(src + l * nm + 128 * t, pol_keep);
            }
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < m) *reinterpret_cast<uint4*>(Sw + 4 * lane + 128 * t) = keys_of(tag, v[t]);
      __syncwarp();
      const int col = (I ^ (lane & 3)) << 2;  // task & 3 == lane & 3 for every task lane + 32 u
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (lane + 32 * u < m) {
          const uint4 o = make_uint4(sw_rd[32 * u], sw_rd[m + 32 * u], sw_rd[2 * m + 32 * u], sw_rd[3 * m + 32 * u]);
          *reinterpret_cast<uint4*>(kt_wr + 32 * u * kKP + col) = o;
        }
      __syncwarp();
    };
    // part of an agent group: task groups [ja, ja + len) of agents 4 I .. 4 I + 3 (the groups left over when the number
    // of agent groups is not a multiple of the warp count are split by task range, so that no warp works an extra round)
    auto unit_part = [&](auto tag, int I, int ja, int len) {
      const int cnt4 = 4 * len;
      for (int fpos = lane; fpos < cnt4; fpos += 32) {
        const int r = fpos / len, j4 = ja + fpos - r * len;
        const float* src = win + (4 * I + r) * m + 4 * j4;
        float4 v[kL];
#pragma unroll
        for (int l = 0; l < kL; ++l) {
          v[l] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (l < Leff) v[l] = ldg_hint4(src + l * nm, pol_keep);
        }
        *reinterpret_cast<uint4*>(Sw + r * m + 4 * j4) = keys_of(tag, v);
      }
      __syncwarp();
      for (int x = lane; x < cnt4; x += 32) {
        const int jt = 4 * ja + x;
        const uint4 o = make_uint4(Sw[jt], Sw[m + jt], Sw[2 * m + jt], Sw[3 * m + jt]);
        *reinterpret_cast<uint4*>(KT + jt * kKP + ((I ^ (jt & 3)) << 2)) = o;
      }
      __syncwarp();
    };
    auto run = [&](auto tag) {
      const int full = n4 / kWarps;  // rounds in which every warp has a whole agent group
      for (int r = 0; r < full; ++r) unit_rows(tag, r * kWarps + warp);
      const int len = (m4 + kWarps - 1) / kWarps, ja = warp * len, mine = min(len, m4 - ja);
      for (int I = full * kWarps; I < n4; ++I)
        if (mine > 0) unit_part(tag, I, ja, mine);
    };
    if (Leff == kL && k_nonneg) run(std::true_type{});
    else run(std::false_type{});
  }
  const bool beta_fast = vw.beta.ptr && vw.beta.dtype == SAP_F16 &&
                         ((reinterpret_cast<uintptr_t>(vw.beta.ptr) + 2 * sap_field_off(vw.beta, b, t_slot)) & 7) == 0;
  if (vw.beta.ptr && !beta_fast) {  // eager `beta` buffer field in another dtype than the scheme's fp16: scalar stores
    const int64_t bb = sap_field_off(vw.beta, b, t_slot);
    for (int i = warp; i < n; i += kWarps)
      for (int j = lane; j < m; j += 32)
        for (int l = 0; l < kL; ++l)
          sap_store_real(vw.beta.ptr, bb + ((int64_t)i * m + j) * kL + l, vw.beta.dtype,
                         l < Leff ? (double)win[(size_t)l * nm + i * m + j] : 0.0);
  }
  __syncthreads();

  SAP_STOP_AFTER(2)
#ifdef SAP_ABLATE
  if (p.debug_skip_redo >= 20 && p.debug_skip_redo < 30) return;
#endif
  // exact float64 window sum (the reference's beta.sum(-1), :190) from global memory, for uncertified lists only
  auto tot64 = [&](int a, int j) {
    double s = 0.0;
    for (int l = 0; l < Leff; ++l) s += (double)win[(size_t)l * nm + a * m + j];
    return s;
  };

  // ------------------------------------------------------------------ 5. per-agent task lists (:198, :217)
  // Two lanes per list; lane s takes the tasks e with (e & 4) == 4 s (a column walk of KT; element c of a lane is
  // task 8 (c >> 2) + (c & 3) + 4 s).  Packed word = key + (imask - e): one add per element.
  for (int base = 0; base < n * 2; base += kThreads) {
    const int g = base + tid;
    const int i = g >> 1, s = g & 1;
    const bool live = i < n;
    if (!__any_sync(SAP_FULL_MASK, live)) continue;
    const int ii = live ? i : 0;
    const uint32_t* colp[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) colp[t] = KT + s * 4 * kKP + ((((ii >> 2) ^ t) << 2) | (ii & 3));
    uint32_t lowb = imask - 4u * (uint32_t)s;  // imask - e for e0 = 0
    const int per = ((m + 7) >> 3) * 4;        // elements per lane (rows up to m8 exist and hold zero keys)
    uint32_t top[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int e0 = 8 * (c >> 2) + (c & 3);
      top[c] = colp[c & 3][e0 * kKP] + (lowb - (uint32_t)e0);
    }
    SAP_SORT16(top);
    int cb = 16;
    for (; cb + 16 <= per; cb += 16) {
#pragma unroll
      for (int t = 0; t < 4; ++t) colp[t] += 32 * kKP;
      lowb -= 32u;
      uint32_t ch[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int e0 = 8 * (c >> 2) + (c & 3);
        ch[c] = colp[c & 3][e0 * kKP] + (lowb - (uint32_t)e0);
      }
      SAP_SORT16(ch);
#pragma unroll
      for (int c = 0; c < 16; ++c) top[c] = max(top[c], ch[15 - c]);
      SAP_BITONIC_MERGE16(top);
    }
    if (cb < per) {  // 4, 8 or 12 elements left
#pragma unroll
      for (int t = 0; t < 4; ++t) colp[t] += 32 * kKP;
      lowb -= 32u;
      const int rem = per - cb;
      if (rem == 4) {
        uint32_t x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = colp[c][c * kKP] + (lowb - (uint32_t)c);
        SAP_CE(x[0], x[1]); SAP_CE(x[2], x[3]); SAP_CE(x[0], x[2]); SAP_CE(x[1], x[3]); SAP_CE(x[1], x[2]);
#pragma unroll
        for (int c = 0; c < 4; ++c) top[12 + c] = max(top[12 + c], x[3 - c]);
      } else {
        uint32_t ch[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int e0 = 8 * (c >> 2) + (c & 3);
          ch[c] = c < rem ? colp[c & 3][e0 * kKP] + (lowb - (uint32_t)e0) : 0u;
        }
        SAP_SORT16(ch);
#pragma unroll
        for (int c = 0; c < 16; ++c) top[c] = max(top[c], ch[15 - c]);
      }
      SAP_BITONIC_MERGE16(top);
    }
    {
      uint32_t ot[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) ot[c] = __shfl_xor_sync(SAP_FULL_MASK, top[15 - c], 1);
#pragma unroll
      for (int c = 0; c < 16; ++c) top[c] = max(top[c], ot[c]);
      SAP_BITONIC_MERGE16(top);
    }
    if (live && s == 0) {
      if (!certified(top, kK2, ib)) {
        qRows[atomicAdd(&sQ[0], 1)] = i;
      } else {
        // D: first M entries as they are (ties are proven ties, already in index-ascending order)
#pragma unroll
        for (int t = 0; t < kM; ++t) sD[i * kM + t] = (uint8_t)(imask - (top[t] & imask));
        // E: same values, ties in index-DESCENDING order and, when the tie group of the 15th entry extends past the
        // cut, its LARGEST indices
        uint32_t kk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) kk[t] = top[t] >> ib;
        bool ties = false;
#pragma unroll
        for (int t = 0; t < kK2 - 1; ++t) ties = ties || (kk[t] == kk[t + 1]);
        const bool ext = kk[kK2 - 1] == kk[kK2];
        const uint32_t vstar = kk[kK2 - 1] << ib;  // as stored in KT
        int pfx = kK2;
        if (ext) {
          pfx = 0;
#pragma unroll
          for (int t = 0; t < kK2; ++t)
            if (kk[t] > kk[kK2 - 1]) pfx = t + 1;
        }
#pragma unroll
        for (int t = 0; t < kK2; ++t)
          if (t < pfx) sE[i * 16 + t] = (uint8_t)(imask - (top[t] & imask));
        if (ties) {  // reverse every run of equal keys inside the prefix (rare: duplicate values)
          int rs = 0;
          while (rs < pfx) {
            int re = rs + 1;
            const uint32_t kv = key_at(i, sE[i * 16 + rs]);
            while (re < pfx && key_at(i, sE[i * 16 + re]) == kv) ++re;
            for (int x = rs, y = re - 1; x < y; ++x, --y) {
              const uint8_t tmp = sE[i * 16 + x];
              sE[i * 16 + x] = sE[i * 16 + y];
              sE[i * 16 + y] = tmp;
            }
            rs = re;
          }
        }
        if (ext)
          for (int j = m - 1; j >= 0 && pfx < kK2; --j)
            if (key_at(i, j) == vstar) sE[i * 16 + pfx++] = (uint8_t)j;
      }
    }
  }
  __syncthreads();
  // exact float64 redo of the lists that could not be certified (near-ties, negative benefits)
  for (int qi = warp; qi < sQ[0]; qi += kWarps) {
    const int i = qRows[qi];
    double vals[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) vals[c] = (lane + 32 * c < m) ? tot64(i, lane + 32 * c) : 0.0;
    warp_select_cached(m, kM, false, lane, vals, [&](int r, int j) { sD[i * kM + r] = (uint8_t)j; });
    warp_select_cached(m, kK2, true, lane, vals, [&](int r, int j) { sE[i * 16 + r] = (uint8_t)j; });
  }
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {  // membership mask of D[i], used to filter the rivals' lists
    uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
#pragma unroll
    for (int q = 0; q < kM; ++q) {
      const uint32_t j = sD[i * kM + q];
      const uint32_t bit = 1u << (j & 31u), wi = j >> 5;
      w0 |= wi == 0u ? bit : 0u;
      w1 |= wi == 1u ? bit : 0u;
      w2 |= wi == 2u ? bit : 0u;
      w3 |= wi == 3u ? bit : 0u;
    }
    *reinterpret_cast<uint4*>(sMask + i * 4) = make_uint4(w0, w1, w2, w3);
  }
  if (p.top_out)  // the top-M lists are final: one coalesced pass instead of ten lanes per row in the gather
    for (int e = tid; e < n * kM; e += kThreads) p.top_out[(size_t)b * n * kM + e] = sD[e];

  SAP_STOP_AFTER(3)
  // ------------------------------------------------------------------ 6. rivals (:203-206)
  // Four lanes per list, 64 lists per pass.  Lane s of list i owns the rival columns s + 4 c (agents 4 s + 16 c .. + 3):
  // for every top task of i (a row of KT) it loads those columns with 128-bit loads and keeps the running maximum.
  // Lists of odd slot walk the column pairs in the opposite order, so the two lists of a quarter-warp always read
  // opposite halves of a 128-byte bank line.  Columns of agents >= n hold zero keys (padding, ranks last).
  for (int base_i = 0; base_i < n; base_i += kThreads / 4) {
    const int li = tid >> 2, s = tid & 3, par = li & 1;
    const int i = base_i + li;
    const bool live = i < n;
    if (!__any_sync(SAP_FULL_MASK, live)) continue;
    const uint8_t* myD = sD + (live ? i : 0) * kM;
    const int dy = 16 - 32 * par;
    const bool seven = n > 96 && n <= 112;  // column 6 is the last one with agents: load it alone (column 7 is padding)
    const int pairs = n > 112 ? 4 : 3;
    uint4 bx[4], by[4];  // bx[u]: column 2u + par, by[u]: column 2u + 1 - par; bx[3] = column 6 when `seven`
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bx[u] = make_uint4(0u, 0u, 0u, 0u);
      by[u] = bx[u];
    }
#pragma unroll
    for (int qp = 0; qp < kM / 2; ++qp) {
      const int j0 = myD[2 * qp], j1 = myD[2 * qp + 1];
      const uint32_t* r0 = KT + j0 * kKP + ((s ^ (j0 & 3)) << 2);
      const uint32_t* r1 = KT + j1 * kKP + ((s ^ (j1 & 3)) << 2);
      const uint32_t* x0 = r0 + 16 * par;
      const uint32_t* x1 = r1 + 16 * par;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (u < pairs) {
          const uint4 a0 = *reinterpret_cast<const uint4*>(x0 + 32 * u);
          const uint4 a1 = *reinterpret_cast<const uint4*>(x1 + 32 * u);
          const uint4 c0 = *reinterpret_cast<const uint4*>(x0 + 32 * u + dy);
          const uint4 c1 = *reinterpret_cast<const uint4*>(x1 + 32 * u + dy);
          bx[u].x = umax3(bx[u].x, a0.x, a1.x);
          bx[u].y = umax3(bx[u].y, a0.y, a1.y);
          bx[u].z = umax3(bx[u].z, a0.z, a1.z);
          bx[u].w = umax3(bx[u].w, a0.w, a1.w);
          by[u].x = umax3(by[u].x, c0.x, c1.x);
          by[u].y = umax3(by[u].y, c0.y, c1.y);
          by[u].z = umax3(by[u].z, c0.z, c1.z);
          by[u].w = umax3(by[u].w, c0.w, c1.w);
        }
      }
      if (seven && 96 + 4 * s < n) {  // quads of column 6 past the last agent hold zero keys: bx[3] stays zero
        const uint4 a0 = *reinterpret_cast<const uint4*>(r0 + 16 * 6);
        const uint4 a1 = *reinterpret_cast<const uint4*>(r1 + 16 * 6);
        bx[3].x = umax3(bx[3].x, a0.x, a1.x);
        bx[3].y = umax3(bx[3].y, a0.y, a1.y);
        bx[3].z = umax3(bx[3].z, a0.z, a1.z);
        bx[3].w = umax3(bx[3].w, a0.w, a1.w);
      }
    }
    // packed words (score + ~agent); agent i itself becomes padding
    uint32_t top[16], ch[16];
    {
      const uint32_t offX = 4u * s + 16u * par, offY = 4u * s + 16u - 16u * par;
      const uint32_t lowX = imask - offX, lowY = imask - offY;
      const uint32_t selfX = (uint32_t)i - offX, selfY = (uint32_t)i - offY;  // wraps when i is not in this lane's columns
      auto pk = [](uint32_t key, uint32_t low, uint32_t c, uint32_t self) -> uint32_t {
        return c == self ? 0u : key + (low - c);
      };
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t c = 32u * u;
        const bool single = u == 3 && seven;       // bx[3] holds column 6 for both parities
        const uint32_t lo_x = single ? imask - 4u * s - 96u + c : lowX;   // agent = 4 s + 96 + t
        const uint32_t se_x = single ? (uint32_t)i - 4u * s - 96u + c : selfX;
        top[4 * u + 0] = pk(bx[u].x, lo_x, c + 0, se_x);
        top[4 * u + 1] = pk(bx[u].y, lo_x, c + 1, se_x);
        top[4 * u + 2] = pk(bx[u].z, lo_x, c + 2, se_x);
        top[4 * u + 3] = pk(bx[u].w, lo_x, c + 3, se_x);
        if (u < 3 || pairs == 4) {
          ch[4 * u + 0] = pk(by[u].x, lowY, c + 0, selfY);
          ch[4 * u + 1] = pk(by[u].y, lowY, c + 1, selfY);
          ch[4 * u + 2] = pk(by[u].z, lowY, c + 2, selfY);
          ch[4 * u + 3] = pk(by[u].w, lowY, c + 3, selfY);
        } else {
          ch[12] = ch[13] = ch[14] = ch[15] = 0u;
        }
      }
      if (pairs == 3 && !seven) top[12] = top[13] = top[14] = top[15] = 0u;  // n <= 96: no column 6
    }
    SAP_SORT16(top);
    SAP_SORT16(ch);
#pragma unroll
    for (int c = 0; c < 16; ++c) top[c] = max(top[c], ch[15 - c]);
    SAP_BITONIC_MERGE16(top);
#pragma unroll
    for (int stride = 1; stride < 4; stride <<= 1) {
      uint32_t ot[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) ot[c] = __shfl_xor_sync(SAP_FULL_MASK, top[15 - c], stride);
#pragma unroll
      for (int c = 0; c < 16; ++c) top[c] = max(top[c], ot[c]);
      SAP_BITONIC_MERGE16(top);
    }
    if (live && s == 0) {
      if (!certified(top, kN, ib)) {
        qNbr[atomicAdd(&sQ[1], 1)] = i;
      } else {
#pragma unroll
        for (int t = 0; t < kN; ++t) sNbr[i * kN + t] = (uint8_t)(imask - (top[t] & imask));
      }
    }
  }
  __syncthreads();
  for (int qi = warp; qi < sQ[1]; qi += kWarps) {
    const int i = qNbr[qi];
    // exact scores once into scratch (the key tile is dead after the barrier above), then the exact selection
    double* sc = reinterpret_cast<double*>(KT) + (size_t)warp * n;
    for (int a = lane; a < n; a += 32) {
      double best = -INFINITY;
      for (int q = 0; q < kM; ++q) best = fmax(best, tot64(a, sD[i * kM + q]));
      sc[a] = (a == i) ? -INFINITY : best;
    }
    __syncwarp();
    warp_select(n, kN, false, lane, [&](int a) { return sc[a]; }, [&](int r, int a) { sNbr[i * kN + r] = (uint8_t)a; });
    __syncwarp();
  }
  __syncthreads();

  SAP_STOP_AFTER(4)
  // ------------------------------------------------------------------ 7. rivals' other top tasks (:212-217)
  // The 5 best tasks of rival r outside D[i] under (value desc, idx desc) are the first 5 entries of E[r] not in
  // D[i]; the reference lists them ascending, so they are stored reversed.
  for (int it0 = 0; it0 < n * kN; it0 += kThreads) {
    const int it = it0 + tid;
    const bool act = it < n * kN;
    const int itc = act ? it : 0;
    const int i = itc / kN;
    const int r = sNbr[itc];
    const uint4 e4 = *reinterpret_cast<const uint4*>(sE + r * 16);
    const uint32_t ew[4] = {e4.x, e4.y, e4.z, e4.w};
    const uint4 mk = *reinterpret_cast<const uint4*>(sMask + i * 4);  // membership bits of D[i]: registers, not one LDS per entry
    uint8_t* dst = sOther + itc * kH + (kH - 1);
    int c = act ? 0 : kH;
#pragma unroll
    for (int e = 0; e < kK2; ++e) {
      if (e >= kH + 1 && (e & 1) == 0 && __all_sync(SAP_FULL_MASK, c >= kH)) break;  // whole warp has its five
      const uint32_t j = (ew[e >> 2] >> (8 * (e & 3))) & 0xffu;
      const uint32_t mw = (j & 64u) ? ((j & 32u) ? mk.w : mk.z) : ((j & 32u) ? mk.y : mk.x);
      const bool in_d = (mw >> (j & 31u)) & 1u;
      if (c < kH && !in_d) {
        *dst = (uint8_t)j;
        --dst;
        ++c;
      }
    }
  }
  __syncthreads();

  SAP_STOP_AFTER(5)
  // ------------------------------------------------------------------ 8a. benefits of the window -> shared memory
  // Second read of the window (L2-resident: this CTA streamed it a few microseconds ago), rounded to fp16:
  // planes 0 and 1 of a pair side by side in one word, plane 2 in a separate plane.
  {
    const int total4 = nm >> 2;
    auto put = [&](int e4, const float4 (&v)[kL]) {
      const __half2 h[4] = {__floats2half2_rn(v[0].x, v[1].x), __floats2half2_rn(v[0].y, v[1].y),
                            __floats2half2_rn(v[0].z, v[1].z), __floats2half2_rn(v[0].w, v[1].w)};
      *reinterpret_cast<uint4*>(t01 + (size_t)e4 * 4) = *reinterpret_cast<const uint4*>(h);
      const __half2 g2[2] = {__floats2half2_rn(v[2].x, v[2].y), __floats2half2_rn(v[2].z, v[2].w)};
      *reinterpret_cast<uint2*>(t2 + (size_t)e4 * 4) = *reinterpret_cast<const uint2*>(g2);
      if (beta_fast) {  // eager `beta` field [n][m][3] fp16: the same rounded values, 4 pairs = 24 bytes
        const uint32_t* h32 = reinterpret_cast<const uint32_t*>(h);
        const uint32_t* g32 = reinterpret_cast<const uint32_t*>(g2);
        uint2* o = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(vw.beta.ptr) + sap_field_off(vw.beta, b, t_slot) +
                                            (int64_t)e4 * 12);
        // halves: p0.0 p0.1 p0.2 p1.0 | p1.1 p1.2 p2.0 p2.1 | p2.2 p3.0 p3.1 p3.2
        o[0] = make_uint2(h32[0], __byte_perm(g32[0], h32[1], 0x5410));
        o[1] = make_uint2(__byte_perm(h32[1], g32[0], 0x7632), h32[2]);
        o[2] = make_uint2(__byte_perm(g32[1], h32[3], 0x5410), __byte_perm(h32[3], g32[1], 0x7632));
      }
    };
    auto run = [&](auto tag) {
      constexpr bool kFull = decltype(tag)::value;
      for (int e4 = tid; e4 < total4; e4 += 2 * kThreads) {
        const int e4b = e4 + kThreads;
        const bool has_b = e4b < total4;
        float4 va[kL], vb[kL];
#pragma unroll
        for (int l = 0; l < kL; ++l) {
          if (!kFull) {
            va[l] = make_float4(0.f, 0.f, 0.f, 0.f);
            vb[l] = va[l];
          }
          if (kFull || l < Leff) va[l] = ldg_hint4(win + l * nm + e4 * 4, pol_drop);
          if ((kFull || l < Leff) && has_b) vb[l] = ldg_hint4(win + l * nm + e4b * 4, pol_drop);
        }
        put(e4, va);
        if (has_b) put(e4b, vb);
      }
    };
    if (Leff == kL) run(std::true_type{});
    else run(std::false_type{});
  }
  __syncthreads();

  SAP_STOP_AFTER(6)
  // ------------------------------------------------------------------ 8b. gather rows into shared memory, store
  // Row layout (:225): [own top tasks 10 x 3 | rival p on my top tasks 10 x (10 x 3) | rival p's other tasks
  // 10 x (5 x 3) | flags 10] = 32 slots of 15 halves + 10 flags.  Lane -> slot:
  //   lanes 0, 1: my own benefits on D[0..5), D[5..10)           -> slots 0, 1
  //   lane 2 + 3p + h (h = 0, 1): rival p on D[5h .. 5h + 5)     -> slot 2 + 2p + h
  //   lane 2 + 3p + 2: rival p on its other tasks                -> slot 22 + p
  const int g3 = lane < 2 ? 0 : lane - 2;
  const int gp = g3 / 3, part = lane < 2 ? lane : g3 - 3 * gp;
  const bool is_own = lane < 2, is_other = !is_own && part == 2;
  const int slot = is_own ? lane : (is_other ? 22 + gp : 2 + 2 * gp + part);
  const uint32_t idx_base = is_other ? f.other + gp * kH : f.D + (part * kH);  // + i * idx_stride
  const uint32_t idx_stride = is_other ? kN * kH : kM;
  const bool odd = slot & 1;
  const uint32_t slot_off = 30u * slot;
  const bool ain_on = ain != nullptr;
#ifdef SAP_ABLATE
  long long dbg_gather = 0, dbg_store = 0, dbg_t0 = 0;
#endif
  // Every warp owns ONE staging row and works through rows warp, warp + 8, ... on its own: gather the row, read it back
  // in 128-bit pieces, store it as the fp16 observation row and widened to fp32 for the agent network.  No block barrier
  // and no wait for a bulk store inside the phase: while one warp gathers, others store (the phase used to cost two
  // barriers and one TMA read-completion wait per 8 rows, profiles/r02_phase_timeline.txt).
  // A row is 980 bytes, so row i starts 4 (i & 3) bytes past a 16-byte boundary in global memory; the staging row is
  // placed at the same offset, which makes every 16-byte piece of the row aligned on both sides.  2 x 245 floats per row
  // keep the same property for the packed fp32 rows.
  constexpr int kRowWords = kRowBytes / 4;  // 245
  unsigned char* wstage = stage + warp * kStagePitch;
  const bool ain_on_rt =
#ifdef SAP_ABLATE
      ain_on && p.debug_skip_redo != 8;  // 8: fp16 obs store only, no fp32 agent-input copy
#else
      ain_on;
#endif
  for (int i = warp; i < n; i += kWarps) {
#ifdef SAP_ABLATE
    dbg_t0 = clock64();
#endif
    const int mis = i & 3;
    unsigned char* srow = wstage + 4 * mis;
    {
      const int a = is_own ? i : (int)sNbr[i * kN + gp];
      // my five task indices: two aligned words around the 5 bytes + a funnel shift instead of five byte loads (the
      // index arrays are followed by other shared-memory arrays, so the second word is always inside the allocation)
      const uint32_t jaddr = idx_base + i * idx_stride;
      const uint2 jw = make_uint2(*reinterpret_cast<const uint32_t*>(smem + (jaddr & ~3u)),
                                  *reinterpret_cast<const uint32_t*>(smem + (jaddr & ~3u) + 4));
      const uint32_t jsh = (jaddr & 3u) * 8u;
      const uint32_t j03 = __funnelshift_r(jw.x, jw.y, jsh), j4 = (jw.y >> jsh) & 0xffu;
      const uint32_t* r01 = t01 + a * f.p01;
      const __half* r2 = t2 + a * f.p2;
      uint32_t w[kH], h[kH];  // w: planes 0, 1 of pair q; h: plane 2 (low 16 bits)
#pragma unroll
      for (int q = 0; q < kH; ++q) {
        const int j = q < 4 ? (int)((j03 >> (8 * q)) & 0xffu) : (int)j4;
        w[q] = r01[j];
        h[q] = (uint32_t)__half_as_ushort(r2[j]);
      }
      // 15 halves: w0.lo w0.hi h0 w1.lo w1.hi h1 ... ; even slots start word-aligned, odd slots 2 bytes later
      const uint32_t E1 = __byte_perm(h[0], w[1], 0x5410), E2 = __byte_perm(w[1], h[1], 0x5432);
      const uint32_t E4 = __byte_perm(h[2], w[3], 0x5410), E5 = __byte_perm(w[3], h[3], 0x5432);
      const uint32_t O0 = __byte_perm(w[0], h[0], 0x5432), O2 = __byte_perm(h[1], w[2], 0x5410);
      const uint32_t O3 = __byte_perm(w[2], h[2], 0x5432), O5 = __byte_perm(h[3], w[4], 0x5410);
      const uint32_t O6 = __byte_perm(w[4], h[4], 0x5432);
      uint32_t* wp = reinterpret_cast<uint32_t*>(srow + slot_off + (odd ? 2u : 0u));
      wp[0] = odd ? O0 : w[0];
      wp[1] = odd ? w[1] : E1;
      wp[2] = odd ? O2 : E2;
      wp[3] = odd ? O3 : w[2];
      wp[4] = odd ? w[3] : E4;
      wp[5] = odd ? O5 : E5;
      wp[6] = odd ? O6 : w[4];
      *reinterpret_cast<uint16_t*>(srow + slot_off + (odd ? 0u : 28u)) = (uint16_t)(odd ? w[0] : h[4]);
      // "is my previous task among my top-M" flags (:222)
      if (lane < kM) {
        const int pv = sPrev[i];
        const int j = sD[i * kM + lane];
        reinterpret_cast<uint16_t*>(srow)[kPairs * kL + lane] = j == pv ? (uint16_t)0x3c00u : (uint16_t)0u;
      }
    }
    __syncwarp();
#ifdef SAP_ABLATE
    const long long dbg_t1 = clock64();
    dbg_gather += dbg_t1 - dbg_t0;
    if (p.debug_skip_redo == 7) continue;  // gather only: no stores
#endif
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(srow);
    uint32_t* gobs = reinterpret_cast<uint32_t*>(obs_out) + (size_t)i * kRowWords;
    const int w0 = (4 - mis) & 3;             // first word of the row that starts a 16-byte piece
    const int ng = (kRowWords - w0) >> 2;     // whole 16-byte pieces (60 or 61)
    const int tail0 = w0 + 4 * ng, nfrag = w0 + (kRowWords - tail0);  // <= 6 words outside the pieces
    const int fw = lane < w0 ? lane : tail0 + lane - w0;  // this lane's head / tail word (lanes < nfrag)
    const bool flat32 = ain_on_rt && ain32_flat;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int g = lane + 32 * u;
      if (g < ng) {
        const int wq = w0 + 4 * g;
        const uint4 v = *reinterpret_cast<const uint4*>(sw + wq);
        stg_hint4(gobs + wq, v, pol_drop);
        if (flat32) {
          const __half2* hh = reinterpret_cast<const __half2*>(&v);
          const float2 f0 = __half22float2(hh[0]), f1 = __half22float2(hh[1]), f2 = __half22float2(hh[2]),
                       f3 = __half22float2(hh[3]);
          float* o = ain + (size_t)i * kObs + 2 * wq;
          const float4 lo4 = make_float4(f0.x, f0.y, f1.x, f1.y), hi4 = make_float4(f2.x, f2.y, f3.x, f3.y);
          stg_hint4(o, *reinterpret_cast<const uint4*>(&lo4), pol_drop);
          stg_hint4(o + 4, *reinterpret_cast<const uint4*>(&hi4), pol_drop);
        } else if (ain16_bulk) {  // packed fp16 staging rows: the same bytes a second time
          *reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(ain16) + (size_t)i * kRowWords + wq) = v;
        }
      }
    }
    if (lane < nfrag) {
      const uint32_t hw = sw[fw];
      gobs[fw] = hw;
      if (flat32)
        *reinterpret_cast<float2*>(ain + (size_t)i * kObs + 2 * fw) = __half22float2(*reinterpret_cast<const __half2*>(&hw));
      else if (ain16_bulk)
        (reinterpret_cast<uint32_t*>(ain16) + (size_t)i * kRowWords)[fw] = hw;
    }
    if (ain_on_rt && !ain32_flat) {  // pitched fp32 rows (the MAC appends columns): two floats per word of the fp16 row
      float2* drow = reinterpret_cast<float2*>(ain + (size_t)i * ain32_pitch);
#pragma unroll
      for (int w = lane; w < kRowWords; w += 32) {
        const uint32_t hw = sw[w];
        drow[w] = __half22float2(*reinterpret_cast<const __half2*>(&hw));
      }
    }
    if (ain16 && !ain16_bulk) {  // padded fp16 rows: 245 words per row, coalesced 32-bit stores
      uint32_t* adst = reinterpret_cast<uint32_t*>(ain16 + (size_t)i * ain16_pitch);
#pragma unroll
      for (int w = lane; w < kRowWords; w += 32) adst[w] = sw[w];
    }
    __syncwarp();  // the row is rewritten by this warp's next gather
#ifdef SAP_ABLATE
    dbg_store += clock64() - dbg_t1;
#endif
  }
  SAP_TS(7)
#ifdef SAP_ABLATE
  if (p.debug_skip_redo == 99 && tid == 0 && p.scratch) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* ts = reinterpret_cast<unsigned long long*>(p.scratch) + (size_t)blockIdx.x * 16;
    ts[8] = smid;
    ts[9] = (unsigned long long)dbg_gather;   // SM cycles spent gathering (sum over the row blocks)
    ts[10] = (unsigned long long)dbg_store;   // ... and storing
  }
#endif
}

}  // namespace

// The rest of an env step once the observation of slot k + 1 is in place (sap_real_obs_ahead): conflict counts, beta_hat at
// the chosen entries, rewards, counters, prev_assigns, and the "previous task in my top-M" flags of the new rows (:222).
__global__ void __launch_bounds__(128) sap_real_step_only_kernel(RealParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  int32_t* sCnt = reinterpret_cast<int32_t*>(smem);
  __shared__ double sRed[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const SapEnvDims d = p.d;
  const int n = d.n, m = d.m, T = d.T, L = d.L, M = d.M, N = d.N, H = d.M / 2;
  const SapBatchView& vw = p.view;
  if (p.sel.q) select_own_actions<4>(p, b, n, T, tid);
  const int k_old = p.k[b];
  if (k_old >= T) return;
  const int k_new = k_old + 1;
  const bool done = k_new >= T;
  const size_t env_plane0 = d.shared_planes ? (size_t)0 : (size_t)b * T;
  const float* env_planes = p.planes + env_plane0 * (size_t)n * m;
  for (int j = tid; j < m; j += 128) sCnt[j] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += 128) atomicAdd(&sCnt[min(max((int)p.actions[(size_t)b * n + i], 0), m - 1)], 1);
  __syncthreads();
  const int flag_col = (M + N * M + N * H) * L;  // first flag column of an observation row
  const int obs_size = flag_col + M;
  double local = 0.0;
  for (int i = tid; i < n; i += 128) {
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
    const double pen = p.ttrans ? (double)p.ttrans[(size_t)pv * m + a] : (a != pv ? 1.0 : 0.0);
    const double bh = b0 - p.lambda_ * (pen * (sum > 1e-12 ? 1.0 : 0.0));
    const double r = bh > 0.0 ? bh / (double)sCnt[a] : bh;
    local += r;
    if (vw.rewards.ptr) sap_store_real(vw.rewards.ptr, sap_field_off(vw.rewards, b, k_old) + i, vw.rewards.dtype, r);
    if (vw.actions.ptr) sap_store_int(vw.actions.ptr, sap_field_off(vw.actions, b, k_old) + i, vw.actions.dtype, a);
    p.prev[(size_t)b * n + i] = a;
    if (vw.prev_assigns.ptr) sap_store_int(vw.prev_assigns.ptr, sap_field_off(vw.prev_assigns, b, k_new) + i, vw.prev_assigns.dtype, a);
    if (!done && p.top_out) {  // flags of the new row: 1 where my top-M task q is the task I just took
      for (int q = 0; q < M; ++q)
        if (p.top_out[((size_t)b * n + i) * M + q] == a) {
          sap_store_real(vw.obs.ptr, sap_field_off(vw.obs, b, k_new) + (int64_t)i * obs_size + flag_col + q, vw.obs.dtype, 1.0);
          if (vw.agent_in.ptr) {
            const int64_t at = (int64_t)b * vw.agent_in.env_stride + (int64_t)i * vw.agent_in.t_stride + flag_col + q;
            if (vw.agent_in.dtype == SAP_F16) reinterpret_cast<__half*>(vw.agent_in.ptr)[at] = __float2half(1.f);
            else reinterpret_cast<float*>(vw.agent_in.ptr)[at] = 1.f;
          }
        }
    }
  }
  if (vw.actions_onehot.ptr) {
    const int64_t base = sap_field_off(vw.actions_onehot, b, k_old);
    for (int e = tid; e < n * m; e += 128) {
      const int i = e / m, j = e - i * m;
      const int a = min(max((int)p.actions[(size_t)b * n + i], 0), m - 1);
      sap_store_int(vw.actions_onehot.ptr, base + e, vw.actions_onehot.dtype, a == j ? 1 : 0);
    }
  }
  if (p.counts_out)
    for (int j = tid; j < m; j += 128) p.counts_out[(size_t)b * m + j] = sCnt[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(SAP_FULL_MASK, local, off);
  if ((tid & 31) == 0) sRed[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&p.ep_return[b], sRed[0] + sRed[1] + sRed[2] + sRed[3]);
    p.k[b] = k_new;
    if (vw.terminated.ptr) sap_store_int(vw.terminated.ptr, sap_field_off(vw.terminated, b, k_old), vw.terminated.dtype, done);
    if (vw.filled.ptr) sap_store_int(vw.filled.ptr, sap_field_off(vw.filled, b, k_new), vw.filled.dtype, 1);
  }
}

static bool fast2_dims_ok(const SapEnvDims& d) {
  return d.M == kM && d.N == kN && d.L == kL && d.n > 64 && d.n <= 128 && d.m <= 128 && !(d.n & 3) && !(d.m & 3) && d.m >= d.n &&
         f2_layout(d.n, d.m).total <= kMaxSmem;
}

extern "C" int sap_real_agent_in_f16_ok(const SapEnvDims* dims) { return dims && fast2_dims_ok(*dims) ? 1 : 0; }

extern "C" int sap_real_obs_ahead_ok(const SapEnvDims* dims) { return dims && fast2_dims_ok(*dims) ? 1 : 0; }

extern "C" int sap_real_obs_ahead(const SapEnvDims* dims, const float* planes, const float* plane_stats, const int32_t* k,
                                  const SapBatchView* view, int32_t* top_out, void* stream) {
  SAP_REQUIRE(dims && planes && plane_stats && k && view && view->obs.ptr && top_out, SAP_E_NULL, "sap_real_obs_ahead: null pointer");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.k = const_cast<int32_t*>(k);
  p.view = *view;
  p.top_out = top_out;
  p.obs_only = 1;
  int handled = 0;
  const int rc = sap_real_fast2_try(p, stream, &handled);
  if (rc != SAP_OK) return rc;
  SAP_REQUIRE(handled, SAP_E_CONSTRAINT,
              "sap_real_obs_ahead: only the shipped configuration (M = N = 10, L = 3, fp16 obs, no priorities, 64 < n <= 128, "
              "m <= 128, per-plane stats) builds observations ahead of the step; use sap_real_step");
  return SAP_OK;
}

extern "C" int sap_real_step_after_obs(const SapEnvDims* dims, const float* planes, const float* task_prios,
                                       const float* T_trans, double lambda_, const int64_t* actions, int32_t* k, int32_t* prev,
                                       double* ep_return, int32_t* counts_out, const SapBatchView* view, const int32_t* top,
                                       void* stream) {
  SAP_REQUIRE(dims && planes && actions && k && prev && ep_return && view && view->obs.ptr && top, SAP_E_NULL,
              "sap_real_step_after_obs: null pointer");
  SAP_REQUIRE(dims->B > 0 && dims->n > 0 && dims->m > 0 && dims->T > 0 && dims->L > 0 && dims->L <= 8, SAP_E_DIMS,
              "sap_real_step_after_obs: bad dims");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.prios = task_prios;
  p.ttrans = T_trans;
  p.lambda_ = lambda_;
  p.actions = actions;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.counts_out = counts_out;
  p.view = *view;
  p.top_out = const_cast<int32_t*>(top);
  sap_real_step_only_kernel<<<dims->B, 128, sizeof(int32_t) * (size_t)dims->m, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_step_only_kernel");
  return SAP_OK;
}

extern "C" int sap_rollout_step_ok(const SapEnvDims* dims) { return dims && fast2_dims_ok(*dims) ? 1 : 0; }

extern "C" int sap_rollout_step(const SapSelectArgs* sel, const SapEnvDims* dims, const float* planes, const float* plane_stats,
                                const float* T_trans, double lambda_, int64_t* actions_out, int32_t* k, int32_t* prev,
                                double* ep_return, int32_t* counts_out, const SapBatchView* view, int32_t* top,
                                const int32_t* top_ahead, void* stream) {
  SAP_REQUIRE(sel && sel->q && dims && planes && actions_out && k && prev && ep_return && view && view->obs.ptr, SAP_E_NULL,
              "sap_rollout_step: null pointer");
  SAP_REQUIRE(dims->B > 0 && dims->T > 0 && fast2_dims_ok(*dims), SAP_E_CONSTRAINT,
              "sap_rollout_step: only the shipped configuration (M = N = 10, L = 3, 64 < n <= 128, m <= 128, n and m multiples "
              "of 4) selects and steps in one launch; call sap_select_epsilon_greedy + sap_real_step");
  SAP_REQUIRE((sel->u_explore == nullptr) == (sel->u_action == nullptr), SAP_E_NULL,
              "sap_rollout_step: u_explore and u_action are injected together");
  SAP_REQUIRE(top_ahead || (top && plane_stats), SAP_E_NULL, "sap_rollout_step: top / plane_stats is null");
  RealParams p{};
  p.d = *dims;
  p.planes = planes;
  p.plane_stats = plane_stats;
  p.ttrans = T_trans;
  p.lambda_ = lambda_;
  p.actions = actions_out;
  p.k = k;
  p.prev = prev;
  p.ep_return = ep_return;
  p.counts_out = counts_out;
  p.view = *view;
  p.sel.q = sel->q;
  p.sel.B = dims->B;
  p.sel.n = dims->n;
  p.sel.A = p.sel.m = dims->m;
  p.sel.eps = sel->eps;
  p.sel.eps_dev = sel->eps_dev;
  p.sel.seed = sel->seed;
  p.sel.episode_ctr = sel->episode_ctr;
  p.sel.k = k;
  p.sel.u_explore = sel->u_explore;
  p.sel.u_action = sel->u_action;
  p.sel.out = actions_out;
  p.sel_vec4 = sap_aligned16(sel->q) ? 1 : 0;  // m is a multiple of 4 here
  if (top_ahead) {  // the observation of the new slot is in place (sap_real_obs_ahead): selection + the light step
    p.top_out = const_cast<int32_t*>(top_ahead);
    sap_real_step_only_kernel<<<dims->B, 128, sizeof(int32_t) * (size_t)dims->m, (cudaStream_t)stream>>>(p);
    SAP_CUDA_LAUNCH_CHECK("sap_real_step_only_kernel");
    return SAP_OK;
  }
  p.top_out = top;
  int handled = 0;
  const int rc = sap_real_fast2_try(p, stream, &handled);
  if (rc != SAP_OK) return rc;
  SAP_REQUIRE(handled, SAP_E_CONSTRAINT,
              "sap_rollout_step: the buffers do not meet the one-CTA-per-env kernel's layout (fp16 obs, 16-byte aligned rows, "
              "per-plane stats); call sap_select_epsilon_greedy + sap_real_step");
  return SAP_OK;
}

int sap_real_fast2_try(RealParams& p, void* stream, int* handled) {
  *handled = 0;
  const SapEnvDims& d = p.d;
  const SapBatchView& vw = p.view;
  if (!fast2_dims_ok(d) || p.prios || vw.obs.dtype != SAP_F16) return SAP_OK;
  if (!p.plane_stats || (reinterpret_cast<uintptr_t>(p.plane_stats) & 7)) return SAP_OK;
  if (!sap_aligned16(p.planes)) return SAP_OK;
  // the rows of an env must start 16-byte aligned (128-bit observation and agent-input stores, see the gather)
  if (!sap_aligned16(vw.obs.ptr) || ((vw.obs.env_stride * 2) & 15) || ((vw.obs.t_stride * 2) & 15)) return SAP_OK;
  if (vw.agent_in.ptr) {
    if (vw.agent_in.dtype != SAP_F32 && vw.agent_in.dtype != SAP_F16) {
      sap_set_error("sap_real: agent_in must be f32 (or f16 on the one-CTA-per-env kernel of the shipped configuration)");
      return SAP_E_DTYPE;
    }
    const int esz = vw.agent_in.dtype == SAP_F32 ? 4 : 2;
    const bool padded = vw.agent_in.t_stride > kObs && !(vw.agent_in.t_stride & 1);  // rows with extra / pad columns
    if ((vw.agent_in.t_stride != kObs && !padded) || !sap_aligned16(vw.agent_in.ptr) || ((vw.agent_in.env_stride * esz) & 15)) {
      if (esz == 2) {
        sap_set_error("sap_real: an f16 agent_in must be [B, n, pitch >= obs] with an even pitch and 16-byte aligned envs");
        return SAP_E_CONSTRAINT;
      }
      return SAP_OK;
    }
  }
  const F2Layout f = f2_layout(d.n, d.m);
  static thread_local bool configured = false;
  if (!configured) {
    for (const void* fn : {(const void*)sap_real_fast2_kernel<false, 0>, (const void*)sap_real_fast2_kernel<true, 0>,
                           (const void*)sap_real_fast2_kernel<false, 100>, (const void*)sap_real_fast2_kernel<true, 100>}) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (e != cudaSuccess) {
        sap_set_error("sap_real_fast2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return (int)e;
      }
    }
    configured = true;
  }
  *handled = 1;
#ifdef SAP_ABLATE
  {
    const char* la = getenv("SAP_F2_LOOKAHEAD");  // blocks of L2 look-ahead (measured: no gain at 4096 x 100 x 100)
    p.lookahead = la ? atoi(la) : 0;
  }
#else
  p.lookahead = 0;
#endif
  const bool sel = p.sel.q && !p.is_reset && !p.obs_only;
  const bool fixed100 = d.n == 100 && d.m == 100 && sap_real_path_override() != SAP_REAL_PATH_FAST_RUNTIME_SHAPE;
  if (sel && fixed100) sap_real_fast2_kernel<true, 100><<<d.B, kThreads, f.total, (cudaStream_t)stream>>>(p);
  else if (sel) sap_real_fast2_kernel<true, 0><<<d.B, kThreads, f.total, (cudaStream_t)stream>>>(p);
  else if (fixed100) sap_real_fast2_kernel<false, 100><<<d.B, kThreads, f.total, (cudaStream_t)stream>>>(p);
  else sap_real_fast2_kernel<false, 0><<<d.B, kThreads, f.total, (cudaStream_t)stream>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_fast2_kernel");
  return SAP_OK;
}
