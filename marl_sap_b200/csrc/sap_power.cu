// Power state and interference reward of RealPowerConstellationEnv / InterferenceConstellationEnv (SURVEY.md 8f rank 2;
// /root/reference/src/envs/real_power_constellation_env.py :135-183, :243-250, :310-355 and
// interference_constellation_env.py :309-353).  These envs are the real env plus
//   * a float64 power state per agent.  It must stay float64: 1 - 5 * 0.2 leaves 5.55e-17, which the reward branch sees as
//     alive (`> 0`) while beta_hat sees it as out of power (`< 1e-12`), and the next step turns it into -0.2;
//   * a zero reward for agents out of power (sap_power_pre hands the env kernel a `dead` mask), or - interference env - a
//     reward rule of its own (sap_interference_rewards);
//   * N + 1 power values behind every observation row (sap_power_post, from the rival indices the env kernel exports).
// The env kernel itself (csrc/sap_real.cu, generic one-CTA-per-env path) is shared with the real env.
#include "sap_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work) {
  const int64_t blocks = (work + kThreads - 1) / kThreads;
  return (int)(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}

__device__ __forceinline__ double chosen_benefit0(const SapEnvDims& d, const float* planes, const float* prios, int b, int k,
                                                  int i, int a) {
  const size_t row = d.shared_planes ? (size_t)0 : (size_t)b * d.T;
  const double v = (double)planes[((row + k) * d.n + i) * d.m + a];
  return prios ? v * (double)prios[a] : v;  // beta[i, a, 0] = S[i, a, k] * prio[a]
}

// dead[b, i] = power < 1e-12 (reward 0 this step), then the power update of :172-180 - both from the OLD power and the OLD
// window, so this runs before the env kernel advances k
__global__ void __launch_bounds__(kThreads) sap_power_pre_kernel(SapEnvDims d, const float* __restrict__ planes,
                                                                 const float* __restrict__ prios,
                                                                 const int64_t* __restrict__ actions,
                                                                 const int32_t* __restrict__ k, double* __restrict__ power,
                                                                 uint8_t* __restrict__ dead) {
  const int64_t total = (int64_t)d.B * d.n;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int b = (int)(e / d.n), i = (int)(e - (int64_t)b * d.n);
    const int kb = k[b];
    if (kb >= d.T) continue;
    const double pw = power[e];
    if (dead) dead[e] = pw < 1e-12 ? 1 : 0;
    if (pw > 0.0) {
      const int a = min(max((int)actions[e], 0), d.m - 1);
      const double b0 = chosen_benefit0(d, planes, prios, b, kb, i, a);
      power[e] = b0 > 1e-12 ? pw - 0.2 : fmin(pw + 0.1, 1.0);
    }
  }
}

// N + 1 power columns behind every observation row of the slot the env kernel just wrote (k[b]), the power_states field of
// that slot, and the same columns in the agent-input staging rows.  Finished envs keep their all-zero rows (:252-254).
__global__ void __launch_bounds__(kThreads) sap_power_post_kernel(SapEnvDims d, const int32_t* __restrict__ k,
                                                                  const double* __restrict__ power,
                                                                  const int32_t* __restrict__ nbr, SapBatchView vw,
                                                                  SapField power_field, int base_cols, int obs_row) {
  const int cols = d.N + 1;
  const int64_t total = (int64_t)d.B * d.n * cols;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int c = (int)(e % cols);
    const int64_t bi = e / cols;
    const int b = (int)(bi / d.n), i = (int)(bi - (int64_t)b * d.n);
    const int slot = k[b];
    const double mine = power[bi];
    if (c == 0 && power_field.ptr)
      sap_store_real(power_field.ptr, sap_field_off(power_field, b, slot) + i, power_field.dtype, mine);
    if (slot >= d.T) continue;
    const double v = c == 0 ? mine : power[(int64_t)b * d.n + nbr[bi * d.N + (c - 1)]];
    sap_store_real(vw.obs.ptr, sap_field_off(vw.obs, b, slot) + (int64_t)i * obs_row + base_cols + c, vw.obs.dtype, v);
    if (vw.agent_in.ptr)
      reinterpret_cast<float*>(vw.agent_in.ptr)[(int64_t)b * vw.agent_in.env_stride + (int64_t)i * vw.agent_in.t_stride + base_cols + c] =
          sap_round_real(vw.obs.dtype, v);
  }
}

// interference_reward_function (:309-353), one CTA per env, from the OLD window / power / prev_assigns
__global__ void __launch_bounds__(kThreads) sap_interference_kernel(SapEnvDims d, const float* __restrict__ planes,
                                                                    const float* __restrict__ prios,
                                                                    const float* __restrict__ neighbor,
                                                                    const int32_t* __restrict__ bands, int bands_per_env,
                                                                    double lambda_, const int64_t* __restrict__ actions,
                                                                    const int32_t* __restrict__ k,
                                                                    const int32_t* __restrict__ prev,
                                                                    const double* __restrict__ power,
                                                                    double* __restrict__ ep_return, SapField rewards) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.x, n = d.n, m = d.m, tid = threadIdx.x;
  const int kb = k[b];
  if (kb >= d.T) return;
  int* sA = reinterpret_cast<int*>(smem);       // [n] chosen task
  int* sApp = sA + n;                           // [n] applicable
  int* sBand = sApp + n;                        // [n]
  int* sCnt = sBand + n;                        // [m] applicable agents per task
  double* sRet = reinterpret_cast<double*>(sCnt + ((m + 1) & ~1));  // [kThreads / 32]
  for (int j = tid; j < m; j += kThreads) sCnt[j] = 0;
  __syncthreads();
  double beta0 = 0.0;  // of agent tid (n <= kThreads is not assumed: loop)
  for (int i = tid; i < n; i += kThreads) {
    const int a = min(max((int)actions[(size_t)b * n + i], 0), m - 1);
    const double b0 = chosen_benefit0(d, planes, prios, b, kb, i, a);
    const int app = (power[(size_t)b * n + i] > 0.0 && !(b0 < 1e-12)) ? 1 : 0;  // :314-316
    sA[i] = a;
    sApp[i] = app;
    sBand[i] = bands[(size_t)(bands_per_env ? b : 0) * n + i];
    if (app) atomicAdd(&sCnt[a], 1);                                             // :318-321
  }
  __syncthreads();
  double local = 0.0;
  for (int i = tid; i < n; i += kThreads) {
    const int a = sA[i];
    beta0 = chosen_benefit0(d, planes, prios, b, kb, i, a);
    double conflicts = -1.0;                                                     // beams do not self-conflict (:331)
    for (int o = 0; o < n; ++o)
      if (sBand[o] == sBand[i] && sApp[o]) conflicts += (double)neighbor[(size_t)a * m + sA[o]];
    // beta * 0.5 ** conflicts: exact scaling for the 0/1 neighbour matrices of the reference, pow otherwise
    double r = conflicts == rint(conflicts) ? ldexp(beta0, -(int)conflicts) : beta0 * pow(0.5, conflicts);
    if (sCnt[a] > 0) r /= (double)sCnt[a];                                       // :344-345
    if (sApp[i] && prev[(size_t)b * n + i] != a) r -= lambda_;                   // :348-349
    local += r;
    if (rewards.ptr) sap_store_real(rewards.ptr, sap_field_off(rewards, b, kb) + i, rewards.dtype, r);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(SAP_FULL_MASK, local, off);
  if ((tid & 31) == 0) sRet[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += sRet[w];
    ep_return[b] += t;
  }
}

}  // namespace

extern "C" int sap_power_pre(const SapEnvDims* dims, const float* planes, const float* task_prios, const int64_t* actions,
                             const int32_t* k, double* power, uint8_t* dead_out, void* stream) {
  SAP_REQUIRE(dims && planes && actions && k && power, SAP_E_NULL, "sap_power_pre: null pointer");
  SAP_REQUIRE(dims->B > 0 && dims->n > 0 && dims->m > 0 && dims->T > 0, SAP_E_DIMS, "sap_power_pre: bad dims");
  sap_power_pre_kernel<<<grid_for((int64_t)dims->B * dims->n), kThreads, 0, (cudaStream_t)stream>>>(*dims, planes, task_prios,
                                                                                                   actions, k, power, dead_out);
  SAP_CUDA_LAUNCH_CHECK("sap_power_pre_kernel");
  return SAP_OK;
}

extern "C" int sap_power_post(const SapEnvDims* dims, const int32_t* k, const double* power, const int32_t* nbr,
                              const SapBatchView* view, const SapField* power_states, int32_t base_cols, int32_t obs_row,
                              void* stream) {
  SAP_REQUIRE(dims && k && power && nbr && view && view->obs.ptr, SAP_E_NULL, "sap_power_post: null pointer");
  SAP_REQUIRE(base_cols > 0 && obs_row >= base_cols + dims->N + 1, SAP_E_DIMS, "sap_power_post: obs_row %d < %d + N + 1", obs_row,
              base_cols);
  SAP_REQUIRE(!view->agent_in.ptr || view->agent_in.dtype == SAP_F32, SAP_E_DTYPE, "sap_power_post: agent_in must be f32");
  SapField pf{};
  if (power_states) pf = *power_states;
  sap_power_post_kernel<<<grid_for((int64_t)dims->B * dims->n * (dims->N + 1)), kThreads, 0, (cudaStream_t)stream>>>(
      *dims, k, power, nbr, *view, pf, base_cols, obs_row);
  SAP_CUDA_LAUNCH_CHECK("sap_power_post_kernel");
  return SAP_OK;
}

extern "C" int sap_interference_rewards(const SapEnvDims* dims, const float* planes, const float* task_prios,
                                        const float* neighbor_matrix, const int32_t* sat_freq_bands, int32_t bands_per_env,
                                        double lambda_, const int64_t* actions, const int32_t* k, const int32_t* prev,
                                        const double* power, double* ep_return, const SapField* rewards, void* stream) {
  SAP_REQUIRE(dims && planes && neighbor_matrix && sat_freq_bands && actions && k && prev && power && ep_return, SAP_E_NULL,
              "sap_interference_rewards: null pointer");
  SAP_REQUIRE(dims->B > 0 && dims->n > 0 && dims->m > 0 && dims->T > 0, SAP_E_DIMS, "sap_interference_rewards: bad dims");
  SapField rf{};
  if (rewards) rf = *rewards;
  const size_t smem = sizeof(int) * (3 * (size_t)dims->n + ((dims->m + 1) & ~1)) + sizeof(double) * (kThreads / 32);
  SAP_REQUIRE(smem <= 48 * 1024, SAP_E_CONSTRAINT, "sap_interference_rewards: n=%d m=%d too large", dims->n, dims->m);
  sap_interference_kernel<<<dims->B, kThreads, smem, (cudaStream_t)stream>>>(*dims, planes, task_prios, neighbor_matrix,
                                                                             sat_freq_bands, bands_per_env, lambda_, actions, k,
                                                                             prev, power, ep_return, rf);
  SAP_CUDA_LAUNCH_CHECK("sap_interference_kernel");
  return SAP_OK;
}
