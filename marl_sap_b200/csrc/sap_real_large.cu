// RealConstellationEnv step + observation build for shapes whose per-env state does not fit one SM's shared memory
// (e.g. the reference's real_constellation_env.yaml: 324 agents x 450 tasks).  Same contract and results as
// sap_real.cu / sap_real_fast.cu (reference: real_constellation_env.py step :135-175, beta_hat :282-328,
// _build_obs :177-230), decomposed so that one environment is spread over many CTAs:
//
//   K0 sap_real_large_prep   grid (B)          snapshot of k[b]; origin / power-of-two scale of the selection keys
//   K1 sap_real_large_keys   grid (tiles, B)   float64 window sums -> L2-resident scratch in BOTH layouts,
//                                              [agent][task] and [task][agent]
//   K2 sap_real_large_lists  grid (n/64, B)    per agent: top-M (idx asc) and top-(M+M/2) (idx desc) task lists
//   K3 sap_real_large_main   grid (n/8, B)     warp per agent: rival scores from coalesced [task][agent] rows,
//                                              top-N rivals, rivals' other tasks, gather, obs / agent-input rows;
//                                              chunk 0 also does the reward phase and advances k, prev
//
// Two modes.  Keyed (used when the lists fit the 16-wide networks, like the fast kernel): the scratch holds 32-bit
// fixed-point keys of the float64 sums (+ an "inexact" bit); lists come from sorting networks on packed
// (key | index) words (K2) or from N+1 rounds of redux.max over register-resident packed scores (K3), and are
// accepted only when PROVABLY equal to the float64 answer under the stable tie rules (strictly decreasing keys or
// ties between exact keys); anything else is redone by the exact float64 warp selection.  Exact (any M, N): the
// scratch holds the float64 sums and every selection is the exact warp selection.
#include <stdlib.h>

#include "sap_real.cuh"
#include "sap_sortnet.cuh"

#ifdef SAP_ABLATE
#define SAP_DBG(p) ((p).debug_skip_redo)
#else
#define SAP_DBG(p) 0
#endif

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxRowsPerCta = kThreads / 4;  // K2: 4 or 8 lanes per task list -> 64 or 32 lists per CTA
constexpr int kTileI = 64;                    // agents per tile of the key kernel (x 32 tasks)
constexpr int kES = 16;                       // stride of an E row (uint16 entries; K2 <= 15 in keyed mode)

#define SAP_CE(a, b)            \
  {                             \
    uint32_t hi__ = max(a, b);  \
    b = min(a, b);              \
    a = hi__;                   \
  }

struct LargeScratch {
  double* scale;    // [B][4]  k_lo, k_scale, nonneg, -
  int32_t* ksnap;   // [B]
  uint16_t* D;      // [B][n][M]
  uint16_t* E;      // [B][n][es]  es = 16 in keyed mode (aligned 32-byte rows), K2 otherwise
  uint32_t* K;      // keyed: [B][n][m]
  uint32_t* KT;     // keyed: [B][m][n]
  double* tot;      // exact: [B][n][m]
  double* totT;     // exact: [B][m][n]
  void* G;          // keyed: [B][n][m][4] benefits of the window in the obs dtype (one 8/16-byte read per gathered pair)
};

__host__ __device__ inline size_t dbl_of_bytes(size_t bytes) { return (bytes + 7) / 8; }

__host__ __device__ inline bool large_keyed(const SapEnvDims& d) {
  return d.M + d.M / 2 + 1 <= 16 && d.N + 1 <= 16 && d.n <= 511 && d.m <= 511 && d.L <= 4;
}

// sized for the exact mode (float64 sums); the keyed mode uses half of the two big arrays
__host__ __device__ inline int e_stride(const SapEnvDims& d) { return large_keyed(d) ? kES : d.M + d.M / 2; }

__host__ __device__ inline size_t large_layout(const SapEnvDims& d, LargeScratch* s, double* base) {
  const int K2 = e_stride(d);
  size_t off = 0;
  const size_t o_scale = off; off += (size_t)d.B * 4;
  const size_t o_k = off;     off += dbl_of_bytes(sizeof(int32_t) * (size_t)d.B);
  const size_t o_D = off;     off += dbl_of_bytes(sizeof(uint16_t) * (size_t)d.B * d.n * d.M);
  off = (off + 1) & ~(size_t)1;  // E rows 16-byte aligned
  const size_t o_E = off;     off += dbl_of_bytes(sizeof(uint16_t) * (size_t)d.B * d.n * K2);
  const size_t nm = (size_t)d.B * d.n * d.m;
  const size_t o_a = off;     off += nm;
  const size_t o_b = off;     off += nm;
  off = (off + 1) & ~(size_t)1;  // 16-byte aligned
  const size_t o_g = off;     off += large_keyed(d) ? 2 * nm : 0;
  if (s) {
    s->G = base + o_g;
    s->scale = base + o_scale;
    s->ksnap = reinterpret_cast<int32_t*>(base + o_k);
    s->D = reinterpret_cast<uint16_t*>(base + o_D);
    s->E = reinterpret_cast<uint16_t*>(base + o_E);
    s->K = reinterpret_cast<uint32_t*>(base + o_a);
    s->KT = reinterpret_cast<uint32_t*>(base + o_b);
    s->tot = base + o_a;
    s->totT = base + o_b;
  }
  return off;
}

// exact float64 window sum (the reference's beta.sum(-1), :190) straight from the planes
__device__ __forceinline__ double tot64(const RealParams& p, const float* win, int Leff, int a, int j) {
  const double pr = p.prios ? (double)p.prios[j] : 1.0;
  const size_t nm = (size_t)p.d.n * p.d.m;
  double s = 0.0;
  for (int l = 0; l < Leff; ++l) s += (double)win[(size_t)l * nm + (size_t)a * p.d.m + j] * pr;
  return s;
}

// top-16 of a list under the packed order, kTPL adjacent lanes per list (same scheme as sap_real_fast.cu)
template <int kTPL, typename KeyFn>
__device__ __forceinline__ void group_top16(int len, int s, KeyFn key, uint32_t (&top)[16]) {
  const int per = (len + kTPL - 1) / kTPL;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int e = s + kTPL * c;
    top[c] = (e < len) ? key(e) : 0u;
  }
  SAP_SORT16(top);
  for (int base = 16; base < per; base += 16) {
    uint32_t ch[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int e = s + kTPL * (base + c);
      ch[c] = (e < len) ? key(e) : 0u;
    }
    SAP_SORT16(ch);
#pragma unroll
    for (int c = 0; c < 16; ++c) top[c] = max(top[c], ch[15 - c]);
    SAP_BITONIC_MERGE16(top);
  }
#pragma unroll
  for (int stride = 1; stride < kTPL; stride <<= 1) {
    uint32_t ot[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) ot[c] = __shfl_xor_sync(SAP_FULL_MASK, top[15 - c], stride);
#pragma unroll
    for (int c = 0; c < 16; ++c) top[c] = max(top[c], ot[c]);
    SAP_BITONIC_MERGE16(top);
  }
}

// env index of a CTA: grid = (chunks, min(B, 32768), ceil(B / 32768))
constexpr int kEnvFold = 32768;
// kShape: what the kernel knows at compile time.  0 = nothing (any M / N / L / n / m); 1 = the shipped configuration
// M = N = 10, L = 3 (every list length and the observation layout fold into constants, the per-pair loops unroll);
// 2 = that configuration at the constellation shape 324 x 450 of BASELINE config 3 (pitches and tile counts too).
template <int kShape>
__device__ __forceinline__ SapEnvDims shaped_dims(const SapEnvDims& in) {
  SapEnvDims d = in;
  if (kShape >= 1) {
    d.L = 3;
    d.M = 10;
    d.N = 10;
  }
  if (kShape == 2) {
    d.n = 324;
    d.m = 450;
  }
  return d;
}

__device__ __forceinline__ int env_of_block() { return (int)(blockIdx.z * kEnvFold + blockIdx.y); }

// adjacent keys strictly decreasing, or ties between two EXACT keys (bit 0 clear): then the order is proven
__device__ __forceinline__ bool pair_ok(uint32_t a, uint32_t b) { return a > b || (a == b && !(a & 1u)); }

// ---------------------------------------------------------------------------------------------------- K0
__global__ void __launch_bounds__(kThreads) sap_real_large_prep(RealParams p, int keyed) {
  __shared__ double sv[4][kWarps];
  const SapEnvDims d = p.d;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  const int k_old = p.is_reset ? -1 : p.k[b];
  if (tid == 0) s.ksnap[b] = k_old;
  const int k_new = k_old + 1;
  if (!keyed || k_new >= d.T || k_old >= d.T) return;
  const int Leff = min(d.L, d.T - k_new);
  const size_t nm = (size_t)d.n * d.m;
  const size_t plane0 = (d.shared_planes ? (size_t)0 : (size_t)b * d.T) + k_new;
  float vmin = INFINITY;
  double vabs = 0.0;  // sum over planes of max |value|, in float64 so that it is a true bound
  if (p.plane_stats) {
    for (int l = 0; l < Leff; ++l) {
      const float lo = p.plane_stats[2 * (plane0 + l)], hi = p.plane_stats[2 * (plane0 + l) + 1];
      vmin = fminf(vmin, lo);
      vabs += (double)fmaxf(fabsf(lo), fabsf(hi));
    }
  } else {  // no metadata: one extra read of the window
    const float* win = p.planes + plane0 * nm;
    float amax = 0.f;
    for (size_t e = tid; e < (size_t)Leff * nm; e += kThreads) {
      const float v = win[e];
      vmin = fminf(vmin, v);
      amax = fmaxf(amax, fabsf(v));
    }
    vabs = (double)amax * Leff;
  }
  float pabs = 1.f, pneg = 0.f;
  if (p.prios) {
    pabs = 0.f;
    for (int j = tid; j < d.m; j += kThreads) {
      pabs = fmaxf(pabs, fabsf(p.prios[j]));
      if (p.prios[j] < 0.f) pneg = 1.f;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    vmin = fminf(vmin, __shfl_xor_sync(SAP_FULL_MASK, vmin, off));
    vabs = fmax(vabs, __shfl_xor_sync(SAP_FULL_MASK, vabs, off));
    pabs = fmaxf(pabs, __shfl_xor_sync(SAP_FULL_MASK, pabs, off));
    pneg = fmaxf(pneg, __shfl_xor_sync(SAP_FULL_MASK, pneg, off));
  }
  if (lane == 0) {
    sv[0][warp] = vmin;
    sv[1][warp] = vabs;
    sv[2][warp] = pabs;
    sv[3][warp] = pneg;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kWarps; ++w) {
      vmin = fminf(vmin, (float)sv[0][w]);
      vabs = fmax(vabs, sv[1][w]);
      pabs = fmaxf(pabs, (float)sv[2][w]);
      pneg = fmaxf(pneg, (float)sv[3][w]);
    }
    const bool nonneg = vmin >= 0.f && pneg == 0.f;
    const double hi = vabs * (double)pabs * 1.0000001;  // >= every |window sum|
    int e2 = 0;
    if (hi > 0.0) (void)frexp(hi, &e2);  // hi < 2^e2
    const int ib = 32 - __clz(max(d.n, d.m));
    const int vb1 = 31 - ib;  // bits of the fixed-point part
    s.scale[4 * b + 0] = nonneg ? 0.0 : -ldexp(1.0, e2);
    s.scale[4 * b + 1] = hi > 0.0 ? ldexp(1.0, vb1 - e2 - (nonneg ? 0 : 1)) : 1.0;
    s.scale[4 * b + 2] = nonneg ? 1.0 : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------- K1
template <bool kKeyed, int kShape = 0>
__global__ void __launch_bounds__(kThreads) sap_real_large_keys(RealParams p) {
  // tile = kTileI agents x 32 tasks: 8 agent rows (x L planes) in flight per thread
  __shared__ double tile_d[kKeyed ? 1 : kTileI][kKeyed ? 1 : 33];
  __shared__ uint32_t tile_k[kKeyed ? kTileI : 1][kKeyed ? 33 : 1];
  const SapEnvDims d = shaped_dims<kShape>(p.d);
  const int b = env_of_block(), n = d.n, m = d.m, T = d.T, L = d.L;
  if (b >= d.B) return;
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  const int k_old = s.ksnap[b];
  const int k_new = k_old + 1;
  if (k_new >= T || k_old >= T) return;  // done: no window
  const int Leff = min(L, T - k_new);
  const float* win = p.planes + ((d.shared_planes ? (size_t)0 : (size_t)b * T) + k_new) * n * m;
  const int tiles_j = (m + 31) / 32;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x - ti * tiles_j;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ib = 32 - __clz(max(n, m));
  const uint32_t fixed_max = (1u << (31 - ib)) - 1u;
  const double k_lo = kKeyed ? s.scale[4 * b] : 0.0, k_scale = kKeyed ? s.scale[4 * b + 1] : 1.0;
  const bool k_nonneg = kKeyed ? s.scale[4 * b + 2] != 0.0 : true;
#pragma unroll 4
  for (int r = ty; r < kTileI; r += kWarps) {
    const int i = ti * kTileI + r, j = tj * 32 + tx;
    if (i < n && j < m) {
      const double pr = p.prios ? (double)p.prios[j] : 1.0;
      double sum = 0.0;
      double x[4] = {0.0, 0.0, 0.0, 0.0};
      if (kKeyed) {
#pragma unroll
        for (int l = 0; l < 4; ++l)
          if (l < Leff) {
            x[l] = (double)win[((size_t)l * n + i) * m + j] * pr;
            sum += x[l];
          }
        // the L benefits of this (agent, task), rounded once to the obs dtype, side by side: K3 gathers a pair with
        // ONE 8- or 16-byte read instead of L reads in L different planes
        const size_t e = ((size_t)b * n + i) * m + j;
        if (p.view.obs.dtype == SAP_F16) {
          const __half2 h01 = __halves2half2(__double2half(x[0]), __double2half(x[1]));
          const __half2 h23 = __halves2half2(__double2half(x[2]), __double2half(x[3]));
          reinterpret_cast<uint2*>(s.G)[e] = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
        } else {
          reinterpret_cast<float4*>(s.G)[e] = make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]);
        }
      } else {
        for (int l = 0; l < Leff; ++l) sum += (double)win[((size_t)l * n + i) * m + j] * pr;  // :167-170, :190
      }
      if (kKeyed) {
        // monotone 32-bit image of the float64 sum: floor(sum * 2^s) and an "inexact" bit (see sap_real_fast.cu)
        const double y = (sum - k_lo) * k_scale;
        const double t = y + 4503599627370496.0, rr = t - 4503599627370496.0;
        uint32_t fx = (uint32_t)__double2loint(t) - (rr > y ? 1u : 0u);
        bool inexact = !k_nonneg || (rr != y);
        if (fx > fixed_max) {
          fx = fixed_max;
          inexact = true;
        }
        const uint32_t key = (fx << 1) | (inexact ? 1u : 0u);
        s.K[((size_t)b * n + i) * m + j] = key;
        tile_k[r][tx] = key;
      } else {
        s.tot[((size_t)b * n + i) * m + j] = sum;
        tile_d[r][tx] = sum;
      }
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += kWarps) {
    const int j = tj * 32 + r;
#pragma unroll
    for (int h = 0; h < kTileI / 32; ++h) {
      const int i = ti * kTileI + 32 * h + tx;
      if (i < n && j < m) {
        if (kKeyed) s.KT[((size_t)b * m + j) * n + i] = tile_k[32 * h + tx][r];
        else s.totT[((size_t)b * m + j) * n + i] = tile_d[32 * h + tx][r];
      }
    }
  }
}

// Reward phase at the OLD window and the scalar fields of slot k_new (one CTA per env: chunk 0 of K2, in the shadow
// of the list building).  K1..K3 never read k / prev (they use the snapshot and `actions`), so the order is free.
__device__ __forceinline__ void reward_phase(const RealParams& p, int b, int k_old, int32_t* cnt, double* red) {
  const SapEnvDims d = p.d;
  const int n = d.n, m = d.m, T = d.T, L = d.L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const SapBatchView& vw = p.view;
  const float* env_planes = p.planes + (d.shared_planes ? (size_t)0 : (size_t)b * T * n * m);
  const int k_new = k_old + 1;
  const bool done = k_new >= T;
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

// ---------------------------------------------------------------------------------------------------- K2
template <bool kKeyed, int kTPL, int kShape = 0>
__global__ void __launch_bounds__(kThreads, 4) sap_real_large_lists(RealParams p) {
  constexpr int kRowsPerCta = kThreads / kTPL;
  __shared__ int32_t q_cnt;
  __shared__ int32_t q_rows[kMaxRowsPerCta];
  const SapEnvDims d = shaped_dims<kShape>(p.d);
  const int b = env_of_block(), n = d.n, m = d.m, M = d.M, K2 = d.M + d.M / 2;
  if (b >= d.B) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  LargeScratch s;
  large_layout(d, &s, p.scratch);
  const int k_old = s.ksnap[b], k_new = k_old + 1;
  if (blockIdx.x == 0 && (p.is_reset || k_old < d.T)) {
    __shared__ int32_t cnt_s[512];
    __shared__ double red_s[kWarps];
    reward_phase(p, b, k_old, cnt_s, red_s);
  }
  if (k_new >= d.T || k_old >= d.T) return;
  const int Leff = min(d.L, d.T - k_new);
  const float* win = p.planes + ((d.shared_planes ? (size_t)0 : (size_t)b * d.T) + k_new) * n * m;
  const int row0 = blockIdx.x * kRowsPerCta, es = e_stride(d);
  if (tid == 0) q_cnt = 0;
  __syncthreads();
  if (kKeyed) {
    const int ib = 32 - __clz(max(n, m));
    const uint32_t imask = (1u << ib) - 1u;
    const int i = row0 + tid / kTPL, sl = tid % kTPL;
    const bool live = i < n;
    const uint32_t* row = s.K + ((size_t)b * n + (live ? i : 0)) * m;
    uint32_t top[16];
    group_top16<kTPL>(live ? m : 0, sl, [&](int e) { return (row[e] << ib) | (imask - (uint32_t)e); }, top);
    if (live && sl == 0) {
      bool ok = true;
#pragma unroll
      for (int t = 0; t < 15; ++t)
        if (t < K2) ok = ok && pair_ok(top[t] >> ib, top[t + 1] >> ib);
      if (!ok && !(SAP_DBG(p) & 1)) {
        q_rows[atomicAdd(&q_cnt, 1)] = i;
      } else {
        uint16_t* Dr = s.D + ((size_t)b * n + i) * M;
        uint16_t* Er = s.E + ((size_t)b * n + i) * es;
        // D (:198): first M entries as they are (ties are proven ties, already in index-ascending order)
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (t < M) Dr[t] = (uint16_t)(imask - (top[t] & imask));
        // E (:217): same values, ties in index-DESCENDING order and, when the tie group of the K2-th entry extends
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
          if (t < pfx) Er[t] = (uint16_t)(imask - (top[t] & imask));
        if (ties) {  // reverse every run of equal keys inside the prefix
          int rs = 0;
          while (rs < pfx) {
            int re = rs + 1;
            const uint32_t kv = row[Er[rs]];
            while (re < pfx && row[Er[re]] == kv) ++re;
            for (int x = rs, y = re - 1; x < y; ++x, --y) {
              const uint16_t tmp = Er[x];
              Er[x] = Er[y];
              Er[y] = tmp;
            }
            rs = re;
          }
        }
        if (ext)
          for (int j = m - 1; j >= 0 && pfx < K2; --j)
            if (row[j] == vstar) Er[pfx++] = (uint16_t)j;
      }
    }
    __syncthreads();
  }
  // exact float64 warp selection: every row in exact mode, the uncertified rows in keyed mode
  const int todo = kKeyed ? q_cnt : min(kRowsPerCta, n - row0);
  for (int qi = warp; qi < todo; qi += kWarps) {
    const int i = kKeyed ? q_rows[qi] : row0 + qi;
    double vals[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int j = lane + 32 * c;
      vals[c] = j < m ? (kKeyed ? tot64(p, win, Leff, i, j) : s.tot[((size_t)b * n + i) * m + j]) : 0.0;
    }
    uint16_t* Dr = s.D + ((size_t)b * n + i) * M;
    uint16_t* Er = s.E + ((size_t)b * n + i) * es;
    warp_select_cached(m, M, false, lane, vals, [&](int r, int j) { Dr[r] = (uint16_t)j; });   // :198
    warp_select_cached(m, K2, true, lane, vals, [&](int r, int j) { Er[r] = (uint16_t)j; });   // :217 (pre-masking)
  }
}

// Keyed rival selection of agent i (whole warp): kC register slots per lane cover the n <= 32 kC candidate agents.
// Returns whether the N winners are provably the float64 answer.
template <int kC>
__device__ __forceinline__ bool keyed_rivals(const uint32_t* kt, const uint16_t* wD, uint16_t* wN, int n, int m, int M, int N,
                                             int i, int lane) {
  const int ib = 32 - __clz(max(n, m));
  const uint32_t imask = (1u << ib) - 1u;
  uint32_t pk[kC];  // packed (score key | ~agent) of agents lane, lane + 32, ...
#pragma unroll
  for (int c = 0; c < kC; ++c) pk[c] = 0u;
  for (int q = 0; q < M; ++q) {  // score[a] = max_q key[a, D_i[q]] (:203-206): M coalesced [task][agent] rows
    const uint32_t* col = kt + (size_t)wD[q] * n;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const int a = lane + 32 * c;
      if (a < n) pk[c] = max(pk[c], __ldg(col + a));
    }
  }
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int a = lane + 32 * c;
    pk[c] = (a < n && a != i) ? (pk[c] << ib) | (imask - (uint32_t)a) : 0u;
  }
  // N + 1 rounds of "largest remaining packed word" (redux.max); equal keys come out in index-ascending order
  uint32_t prevw = 0u;
  bool ok = true;
  for (int r = 0; r <= N; ++r) {
    uint32_t loc = 0u;
#pragma unroll
    for (int c = 0; c < kC; ++c) loc = max(loc, pk[c]);
    const uint32_t w = __reduce_max_sync(SAP_FULL_MASK, loc);
#pragma unroll
    for (int c = 0; c < kC; ++c) pk[c] = pk[c] == w ? 0u : pk[c];
    if (r > 0) ok = ok && pair_ok(prevw >> ib, w >> ib);
    if (r < N && lane == 0) wN[r] = (uint16_t)(imask - (w & imask));
    prevw = w;
  }
  return ok;
}

// exact float64 rival scores and selection of agent i (whole warp): every agent in exact mode, the rare uncertified
// ones in keyed mode.  Kept out of line so that its 16 doubles per lane do not weigh on K3's register budget.
// (Plain arguments, not the parameter struct: a struct reference would force a local-memory copy of it.)
__device__ __noinline__ void exact_rivals(const float* win, const float* prios, const double* totT_env, int n, int m, int M,
                                          int N, int Leff, int i, int lane, const uint16_t* wD, uint16_t* wN) {
  double vals[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) vals[c] = -INFINITY;
  for (int q = 0; q < M; ++q) {
    const int j = wD[q];
    double v[16];  // 16 independent agents per lane: their loads overlap
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = 0.0;
    if (totT_env) {
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (lane + 32 * c < n) v[c] = totT_env[(size_t)j * n + lane + 32 * c];
    } else {  // the reference's beta.sum(-1) (:190) straight from the planes
      const double pr = prios ? (double)prios[j] : 1.0;
      for (int l = 0; l < Leff; ++l) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (lane + 32 * c < n) v[c] += (double)win[((size_t)l * n + lane + 32 * c) * m + j] * pr;
      }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) vals[c] = fmax(vals[c], v[c]);
  }
#pragma unroll
  for (int c = 0; c < 16; ++c)
    if (lane + 32 * c >= n || lane + 32 * c == i) vals[c] = -INFINITY;
  warp_select_cached(n, N, false, lane, vals, [&](int r, int a) { wN[r] = (uint16_t)a; });
}

// exact top-N (value desc, index asc) of n float64 scores staged in shared memory, one warp
__device__ __noinline__ void exact_select_scores(const double* sScore, int n, int N, int lane, uint16_t* wN) {
  double vals[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) vals[c] = (lane + 32 * c < n) ? sScore[lane + 32 * c] : -INFINITY;
  warp_select_cached(n, N, false, lane, vals, [&](int r, int a) { wN[r] = (uint16_t)a; });
}

// ---------------------------------------------------------------------------------------------------- K3
template <bool kKeyed, int kShape = 0>
__global__ void __launch_bounds__(kThreads, 5) sap_real_large_main(RealParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SapEnvDims d = shaped_dims<kShape>(p.d);
  const int b = env_of_block();
  if (b >= d.B) return;
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
  const size_t mask_off = (sizeof(uint16_t) * (size_t)kWarps * (M + N + N * H) + 15) & ~(size_t)15;
  uint32_t* sLut = reinterpret_cast<uint32_t*>(smem_raw + mask_off + sizeof(uint32_t) * 16 * kWarps);
  const size_t score_off = (mask_off + sizeof(uint32_t) * 16 * kWarps + sizeof(uint32_t) * (size_t)npairs + 15) & ~(size_t)15;
  if (!done)
    for (int pp = tid; pp < npairs; pp += kThreads) {  // obs layout (:225): own | rivals on my tasks | rivals' other tasks
      uint32_t code;
      if (pp < M) code = (0xffffu << 16) | pp;
      else if (pp < M + N * M) code = (((pp - M) / M) << 16) | ((pp - M) % M);
      else code = (((pp - M - N * M) / H) << 16) | 0x8000u | (pp - M - N * M);
      sLut[pp] = code;
    }
  __syncthreads();

  // ------------------------------------------------------------------ my agent's observation row at slot k_new
  const int i = blockIdx.x * kWarps + warp;
  const bool have = i < n;  // warps without an agent stay for the CTA-wide fallback below
  const int64_t out = sap_field_off(vw.obs, b, k_new) + (int64_t)i * obs_size;
  float* arow = vw.agent_in.ptr ? reinterpret_cast<float*>(vw.agent_in.ptr) + (int64_t)b * vw.agent_in.env_stride +
                                      (int64_t)i * vw.agent_in.t_stride
                                : nullptr;
  if (done) {  // :226-228
    for (int c = lane; have && c < obs_size; c += 32) {
      sap_store_real(vw.obs.ptr, out + c, vw.obs.dtype, 0.0);
      if (arow) arow[c] = 0.f;
    }
    return;
  }
  const int Leff = min(L, T - k_new);
  const float* win = env_planes + (size_t)k_new * n * m;
  const uint16_t* Dg = s.D + ((size_t)b * n + (have ? i : 0)) * M;
  for (int q = lane; q < M; q += 32) wD[q] = Dg[q];
  __syncwarp();

  // rivals (:203-206): score[a] = max_q tot[a, D_i[q]], read as M coalesced [task][agent] rows
  bool need_exact = !kKeyed;
  if (kKeyed && have) {
    const uint32_t* kt = s.KT + (size_t)b * m * n;
    bool ok;
    if (n <= 128) ok = keyed_rivals<4>(kt, wD, wN, n, m, M, N, i, lane);
    else if (n <= 256) ok = keyed_rivals<8>(kt, wD, wN, n, m, M, N, i, lane);
    else if (n <= 384) ok = keyed_rivals<12>(kt, wD, wN, n, m, M, N, i, lane);
    else ok = keyed_rivals<16>(kt, wD, wN, n, m, M, N, i, lane);
    need_exact = !ok && !(SAP_DBG(p) & 1);
    __syncwarp();
  }
  if (kKeyed) {
    // Uncertified agents (rare): the WHOLE CTA computes the agent's exact float64 scores (one candidate per thread
    // instead of sixteen per lane), then its warp runs the exact selection.  A lone warp doing all of it used to be the
    // tail of this short kernel.
    __shared__ int sFail[kWarps];
    double* sScore = reinterpret_cast<double*>(smem_raw + score_off);
    if (lane == 0) sFail[warp] = (have && need_exact) ? i : -1;
    __syncthreads();
    for (int w = 0; w < kWarps; ++w) {
      const int iF = sFail[w];
      if (iF < 0) continue;  // CTA-uniform
      const uint16_t* fD = reinterpret_cast<const uint16_t*>(smem_raw) + (size_t)w * (M + N + N * H);
      for (int a = tid; a < n; a += kThreads) {
        double best = -INFINITY;
        if (a != iF)
          for (int q = 0; q < M; ++q) best = fmax(best, tot64(p, win, Leff, a, fD[q]));
        sScore[a] = best;
      }
      __syncthreads();
      if (warp == w) exact_select_scores(sScore, n, N, lane, wN);
      __syncthreads();
    }
    if (!have) return;
  } else {
    if (!have) return;
    exact_rivals(win, p.prios, s.totT + (size_t)b * m * n, n, m, M, N, Leff, i, lane, wD, wN);
    __syncwarp();
  }
  // rivals' other top tasks (:212-217): first M/2 entries of E[r] outside D[i], stored ascending
  if (kKeyed) {
    uint32_t* wMask = reinterpret_cast<uint32_t*>(smem_raw + mask_off) + warp * 16;  // membership bits of D[i] (m <= 511)
    if (lane < 16) wMask[lane] = 0u;
    __syncwarp();
    if (lane < M) atomicOr(&wMask[wD[lane] >> 5], 1u << (wD[lane] & 31));
    __syncwarp();
    for (int ps = lane; ps < N; ps += 32) {
      const uint4* Er4 = reinterpret_cast<const uint4*>(s.E + ((size_t)b * n + wN[ps]) * kES);
      const uint4 e0 = __ldg(Er4), e1 = __ldg(Er4 + 1);  // the whole list in two 128-bit loads
      const uint32_t ew[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
      int c = 0;
#pragma unroll
      for (int e = 0; e < 15; ++e) {
        const uint32_t j = (ew[e >> 1] >> (16 * (e & 1))) & 0xffffu;
        const bool in_top = (wMask[(j >> 5) & 15] >> (j & 31)) & 1u;
        if (e < K2 && c < H && !in_top) {
          wO[ps * H + (H - 1 - c)] = (uint16_t)j;
          ++c;
        }
      }
    }
  } else {
    for (int ps = lane; ps < N; ps += 32) {
      const uint16_t* Er = s.E + ((size_t)b * n + wN[ps]) * e_stride(d);
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
  }
  __syncwarp();
  // gather + store (:199-225)
  auto pair_of = [&](int pp, int& a, int& j) {  // sLut[pp] = (rival slot or 0xffff) << 16 | other-task flag << 15 | slot
    a = i;
    j = 0;
    if (pp < npairs) {
      const uint32_t code = sLut[pp], ps = code >> 16, qs = code & 0x7fffu;
      if (ps != 0xffffu) a = wN[ps];
      j = (code & 0x8000u) ? wO[qs] : wD[qs];
    }
  };
  if (kKeyed) {  // one read per pair from the K1 image (already in the obs dtype), 4 pairs per lane in flight
    const bool f16 = vw.obs.dtype == SAP_F16;
    const size_t row0 = (size_t)b * n;
    for (int pp0 = lane; pp0 < npairs; pp0 += 128) {
      uint4 raw[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int pp = pp0 + 32 * g;
        int a, j;
        pair_of(pp, a, j);
        const size_t e = (row0 + a) * m + j;
        raw[g] = make_uint4(0u, 0u, 0u, 0u);
        if (pp < npairs) {
          if (f16) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(s.G) + e);
            raw[g].x = t.x;
            raw[g].y = t.y;
          } else {
            raw[g] = __ldg(reinterpret_cast<const uint4*>(s.G) + e);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int pp = pp0 + 32 * g;
        if (pp < npairs) {
          const int64_t o = out + (int64_t)pp * L;
          if (f16) {
            const __half* h = reinterpret_cast<const __half*>(&raw[g]);
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (l < L) {
                reinterpret_cast<__half*>(vw.obs.ptr)[o + l] = h[l];
                if (arow) arow[pp * L + l] = __half2float(h[l]);
              }
          } else {
            const float* f = reinterpret_cast<const float*>(&raw[g]);
#pragma unroll
            for (int l = 0; l < 4; ++l)
              if (l < L) {
                reinterpret_cast<float*>(vw.obs.ptr)[o + l] = f[l];
                if (arow) arow[pp * L + l] = f[l];
              }
          }
        }
      }
    }
  } else {
    for (int pp = lane; pp < npairs; pp += 32) {
      int a, j;
      pair_of(pp, a, j);
      const double pr = p.prios ? (double)p.prios[j] : 1.0;
      for (int l = 0; l < L; ++l) {
        const double v = l < Leff ? (double)win[((size_t)l * n + a) * m + j] * pr : 0.0;
        sap_store_real(vw.obs.ptr, out + (int64_t)pp * L + l, vw.obs.dtype, v);
        if (arow) arow[pp * L + l] = sap_round_real(vw.obs.dtype, v);
      }
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

template <bool kKeyed, int kShape = 0>
int launch_mode(RealParams& p, cudaStream_t st) {
  const SapEnvDims& d = p.d;
  const int H = d.M / 2;
  const unsigned gy = (unsigned)min(d.B, kEnvFold), gz = (unsigned)((d.B + kEnvFold - 1) / kEnvFold);
  const dim3 g1(((d.n + kTileI - 1) / kTileI) * ((d.m + 31) / 32), gy, gz);
  // 8 lanes per list when there are too few lists to fill the GPU with 4 (e.g. 64 envs x 324 agents)
  const bool wide = (int64_t)d.B * ((d.n + 63) / 64) < 8 * 148;
  const int rows_per_cta = kThreads / (wide ? 8 : 4);
  const dim3 g2((d.n + rows_per_cta - 1) / rows_per_cta, gy, gz), g3((d.n + kWarps - 1) / kWarps, gy, gz);
  sap_real_large_prep<<<d.B, kThreads, 0, st>>>(p, kKeyed ? 1 : 0);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_prep");
  sap_real_large_keys<kKeyed, kShape><<<g1, kThreads, 0, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_keys");
  if (wide) sap_real_large_lists<kKeyed, 8, kShape><<<g2, kThreads, 0, st>>>(p);
  else sap_real_large_lists<kKeyed, 4, kShape><<<g2, kThreads, 0, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_lists");
  const size_t smem = sizeof(uint16_t) * (size_t)kWarps * (d.M + d.N + d.N * H) + 16 + sizeof(uint32_t) * 16 * kWarps +
                      sizeof(uint32_t) * (size_t)(d.M + d.N * d.M + d.N * H) + 16 + sizeof(double) * (size_t)d.n;
  sap_real_large_main<kKeyed, kShape><<<g3, kThreads, smem, st>>>(p);
  SAP_CUDA_LAUNCH_CHECK("sap_real_large_main");
  return SAP_OK;
}

}  // namespace

int64_t sap_real_large_scratch_doubles(const SapEnvDims& d) { return (int64_t)large_layout(d, nullptr, nullptr); }

int sap_real_large_launch(RealParams& p, void* stream) {
  const SapEnvDims& d = p.d;
  SAP_REQUIRE(p.scratch, SAP_E_SMEM, "sap_real: this problem size needs scratch of sap_real_scratch_doubles() doubles");
  SAP_REQUIRE(d.n <= 512 && d.m <= 512, SAP_E_DIMS, "sap_real (large path): n, m must be <= 512");
  if (large_keyed(d) && !p.large_exact) {  // large_exact: selector override
    const bool common = d.M == 10 && d.N == 10 && d.L == 3 && sap_real_path_override() != SAP_REAL_PATH_FAST_RUNTIME_SHAPE;
    if (common && d.n == 324 && d.m == 450) return launch_mode<true, 2>(p, (cudaStream_t)stream);
    if (common) return launch_mode<true, 1>(p, (cudaStream_t)stream);
    return launch_mode<true>(p, (cudaStream_t)stream);
  }
  return launch_mode<false>(p, (cudaStream_t)stream);
}
