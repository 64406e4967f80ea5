/*
 * marl_sap_b200 -- C ABI of the B200-native rollout hot path of josh-holder/marl_sap.
 *
 * The reference is 100 % Python and has no FFI; its boundary for this path is three Python
 * registries plus EpisodeBatch.  Each entry point below names the
 * reference function (file:line under /root/reference/src) it replaces.  The Python host
 * (marl_sap_b200/) mirrors the reference classes and calls these through ctypes; any other
 * host (C, C++, cffi ...) can bind the same symbols -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory; nothing is allocated, freed or retained by the library;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit sync;
 *     no host-side read of device data, so every call is CUDA-graph capturable
 *     (except *_host uploads, which issue a cudaMemcpyAsync from the given host buffer);
 *   - return value: 0 = ok, < 0 = argument error (SAP_E_*), > 0 = cudaError_t of the launch;
 *     sap_last_error() returns a thread-local message for the last non-zero return;
 *   - never throws, never exits, re-entrant.
 *
 * Benefit tensor layouts
 *   reference layout  sat_prox_mat[n, m, T]  (T innermost), batched as [B, n, m, T];
 *   device layout     planes[B, T, n, m]     (one contiguous n x m plane per time step) so the
 *                     look-ahead window beta_k = planes[k .. k+L) is L contiguous planes.
 */
#ifndef MARL_SAP_B200_H
#define MARL_SAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAP_ABI_VERSION 1

/* element types of episode-buffer fields (the reference schemes use exactly these:
 * real_constellation_env.py:95-107 fp16/int16/bool, mock_constellation_env.py:73-85 fp32/int64/bool) */
enum SapDtype { SAP_F32 = 0, SAP_F16 = 1, SAP_I64 = 2, SAP_I16 = 3, SAP_I32 = 4, SAP_U8 = 5 };

enum SapError {
  SAP_OK = 0,
  SAP_E_NULL = -1,      /* required pointer is null */
  SAP_E_DIMS = -2,      /* non-positive or inconsistent dimension */
  SAP_E_CONSTRAINT = -3,/* reference precondition violated: m >= n, n > N, m >= M + M/2, M even */
  SAP_E_DTYPE = -4,     /* dtype not supported for that field */
  SAP_E_SMEM = -5,      /* required scratch buffer missing for this problem size */
  SAP_E_ALIGN = -6      /* pointer not aligned for vector access */
};

typedef struct SapEnvDims {
  int32_t B; /* independent environments in this launch            */
  int32_t n; /* agents                                               */
  int32_t m; /* tasks (= actions per agent)                          */
  int32_t T; /* episode length = number of benefit planes            */
  int32_t L; /* look-ahead window (already clamped: min(L, T_ctor))  */
  int32_t M; /* top tasks per agent in the observation   (real env)  */
  int32_t N; /* rival agents per agent in the observation (real env) */
  int32_t shared_planes; /* 1 = all B envs read env 0's planes (one sat_prox_mat for every worker,
                            like ParallelRunner's identical env_args), 0 = planes[B,T,n,m] */
} SapEnvDims;

/* One field of an EpisodeBatch ([B, T+1, ...] tensor, episode_buffer.py:30-77).
 * Element (b, t, x) lives at ptr + (b*env_stride + t*t_stride + x) * sizeof(dtype). */
typedef struct SapField {
  void* ptr;          /* null = field absent / not materialised */
  int64_t env_stride; /* in elements */
  int64_t t_stride;   /* in elements */
  int32_t dtype;      /* SapDtype */
  int32_t reserved;
} SapField;

/* The episode-batch slots one env step writes (A.5 timeline of SURVEY.md):
 * at time t = k_old: actions, actions_onehot, rewards, terminated;
 * at time t+1:       obs, prev_assigns, beta, avail_actions, filled. */
typedef struct SapBatchView {
  SapField obs;            /* [n, obs_size]   f32|f16   required            */
  SapField rewards;        /* [n]             f32|f16   nullable            */
  SapField actions;        /* [n, 1]          i64|i16   nullable            */
  SapField actions_onehot; /* [n, m]          i64|i16|f32 nullable (OneHot) */
  SapField terminated;     /* [1]             u8        nullable            */
  SapField filled;         /* [1]             i64       nullable            */
  SapField prev_assigns;   /* [n]             i64|i16   nullable (real env) */
  SapField beta;           /* real [n,m,L] / mock [n,m]  f32|f16 nullable   */
  SapField avail_actions;  /* [n, m]          u8        nullable (all ones) */
  /* Not an EpisodeBatch field: the agent network's input staging [B, n, row] f32 (nullable).  The obs of
   * slot t+1, rounded to the obs dtype and widened back to f32 (what basic_controller.py:82
   * `batch["obs"][:, t].float()` would produce), is written to columns [0, obs_size) of every agent row.
   * env_stride = elements between envs, t_stride = elements between agent rows (>= obs_size). */
  SapField agent_in;
} SapBatchView;

int sap_abi_version(void);
const char* sap_last_error(void);

/* ---- benefit tensors ------------------------------------------------------------------
 * Replaces the per-step slicing sat_prox_mat[:, :, k:k+L] (real_constellation_env.py:127,
 * 167-170; mock_constellation_env.py:104,110,150,157) by a one-off re-layout
 * [B,n,m,T] -> [B,T,n,m]. */
int sap_benefit_ingest(const float* src_nmT, float* dst_Tnm, int32_t B, int32_t n, int32_t m, int32_t T,
                       void* stream);
/* Same, from a HOST buffer: cudaMemcpyAsync into staging_dev ([B,n,m,T]) then re-layout. */
int sap_benefit_upload_host(const float* src_nmT_host, float* staging_dev, float* dst_Tnm, int32_t B, int32_t n,
                            int32_t m, int32_t T, void* stream);

/* sap_benefit_generate = generate_benefits_over_time (envs/mock_constellation_env.py:276-299), the synthetic
 * "constellation-like" benefits MockConstellationEnv draws in its constructor (:34, widths 5..8) and at every reset
 * (:99-100, widths 3..6) when no sat_prox_mat is given, for B envs at once and straight into the planes layout.
 * numpy's stream cannot be reproduced: same law (scale in {1,1,1,10} per task, P(active) = 1/4, centre U(0,T),
 * width U(width_min, width_max), value scale * exp(-(t - centre)^2 / sigma_2 / 2)), Philox draws keyed by
 * (seed; element, episode). */
int sap_benefit_generate(float* planes_Tnm, int32_t B, int32_t n, int32_t m, int32_t T, float width_min, float width_max,
                         uint64_t seed, uint64_t episode, void* stream);

/* Per-plane range metadata, computed once per episode when the benefits are installed:
 * stats[b, t] = {min, max} over planes[b, t, :, :].  The real-env kernels use it to scale window sums into
 * 32-bit selection keys without an extra pass over the window (nullable there: bounds are then derived
 * inside the kernel with an extra read of the window). */
int sap_benefit_stats(const float* planes_Tnm, float* stats /*[B,T,2]*/, int32_t B, int32_t n, int32_t m, int32_t T,
                      void* stream);

/* ---- RealConstellationEnv ---------------------------------------------------------------
 * sap_real_reset  = RealConstellationEnv.reset + get_pretransition_data
 *                   (real_constellation_env.py:116-133, 232-244): k=0, prev=arange(n),
 *                   obs_0 via _build_obs (:177-230) written to slot t=0, filled[0]=1.
 * sap_real_step   = RealConstellationEnv.step (:135-175) incl. beta_hat (:282-328) at the
 *                   chosen entries, conflict counts, reward split, k+=1, done, next window,
 *                   _build_obs, + the EpisodeRunner buffer writes of episode_runner.py:86-100.
 * state:  k[B] int32 in/out, prev[B,n] int32 in/out, ep_return[B] f64 accumulators.
 * top_out [B,n,M] int32 (nullable): agent i's top-M task indices of the NEW observation
 *         (shared with the filtered selector so top-M is computed once).
 * plane_stats [B or 1, T, 2] from sap_benefit_stats (nullable).  When given it MUST bound the planes it
 *         describes (min <= every value <= max): the kernels trust it for the key scale.
 * scratch: sap_real_scratch_doubles(dims) doubles, 16-byte aligned; 0 (pass NULL) when an environment fits one
 *         SM's shared memory, about 4*B*n*m otherwise (shapes like 324 x 450 run as four launches over it). */
int sap_real_reset(const SapEnvDims* dims, const float* planes, const float* plane_stats, const float* task_prios,
                   int32_t* k, int32_t* prev, double* ep_return, const SapBatchView* view, int32_t* top_out,
                   double* scratch, void* stream);
int sap_real_step(const SapEnvDims* dims, const float* planes, const float* plane_stats, const float* task_prios,
                  const float* T_trans, double lambda_, const int64_t* actions, int32_t* k, int32_t* prev,
                  double* ep_return, int32_t* counts_out, const SapBatchView* view, int32_t* top_out, double* scratch,
                  void* stream);
/* number of scratch doubles sap_real_reset / sap_real_step need for these dims (0: none) */
int64_t sap_real_scratch_doubles(const SapEnvDims* dims);

/* The observation of slot k + 1 (real_constellation_env.py:177-225) depends on the benefit window only - NOT on the actions of
 * step k - except for its last M columns ("is my previous task among my top-M", :222).  So one env step can be issued as
 *     sap_real_obs_ahead     : window -> keys -> lists -> gather -> obs rows of slot k[b] + 1 (flags 0), agent-input rows,
 *                              top_out; reads k, writes no counter.  Runs NEXT TO the agent forward of step k (another stream).
 *     sap_real_step_after_obs: conflict counts, beta_hat at the chosen entries, rewards (:135-164), actions / rewards /
 *                              terminated of slot k, k += 1, prev_assigns, filled, and the flags of the new rows (from `top`).
 * Together they write exactly what sap_real_step writes.  sap_real_obs_ahead_ok(dims) = 1 when the ahead kernel exists for
 * these dims (the one-CTA-per-env kernel of the shipped configuration); callers fall back to sap_real_step otherwise. */
int sap_real_obs_ahead_ok(const SapEnvDims* dims);
int sap_real_obs_ahead(const SapEnvDims* dims, const float* planes, const float* plane_stats, const int32_t* k,
                       const SapBatchView* view, int32_t* top_out, void* stream);
int sap_real_step_after_obs(const SapEnvDims* dims, const float* planes, const float* task_prios, const float* T_trans,
                            double lambda_, const int64_t* actions, int32_t* k, int32_t* prev, double* ep_return,
                            int32_t* counts_out, const SapBatchView* view, const int32_t* top, void* stream);

/* Kernel selection of sap_real_reset / sap_real_step.  AUTO picks by shape: the one-CTA-per-env kernels when the env
 * state fits one SM's shared memory (second generation for the shipped configuration M = N = 10, L = 3, fp16 scheme at
 * 64 < n <= 128; first generation otherwise), the generic kernel for unusual M / N / L, the multi-CTA path for large
 * shapes.  All paths produce identical bytes; the override exists so that tests and profiles can run EVERY path on one
 * shape.  Process-wide; returns the previous value (-1 for an unknown value). */
enum {
  SAP_REAL_PATH_AUTO = 0,
  SAP_REAL_PATH_GENERIC = 1,
  SAP_REAL_PATH_LARGE_KEYED = 2,
  SAP_REAL_PATH_LARGE_EXACT = 3,
  SAP_REAL_PATH_FAST_GEN1 = 4,
  SAP_REAL_PATH_FAST_RUNTIME_SHAPE = 5 /* AUTO, but the one-CTA-per-env kernel reads n, m at run time also at 100 x 100,
                                          where AUTO launches the instantiation with the shape compiled in */
};
int32_t sap_real_select_kernel(int32_t which);

/* ---- RealPowerConstellationEnv / InterferenceConstellationEnv --------------------
 * The real env plus a float64 power state per agent (envs/real_power_constellation_env.py:135-183, :243-250, :310-355;
 * envs/interference_constellation_env.py:309-353).  One env step is
 *     [sap_interference_rewards]  ->  sap_power_pre  ->  sap_real_step_ex  ->  sap_power_post
 * sap_real_reset_ex / sap_real_step_ex = sap_real_reset / sap_real_step on the generic one-CTA-per-env kernel with
 *   prev0[B,n]   (reset) the caller's draw for `np.random.choice(m, n, replace=False)` (:130),
 *   dead[B,n]    (step) agents whose reward is forced to 0 (out of power: :157-165 and the zeroed beta_hat of :351-355),
 *   nbr_out[B,n,N] the rival indices of the new observation (the power columns are taken at these agents),
 *   obs_row      elements between consecutive agents' observation rows (base size + N + 1 for these envs).
 * sap_power_pre  = dead mask from the old power, then the power update of :172-180 (float64: 1 - 5 * 0.2 = 5.55e-17 is
 *   `> 0` for the reward branch but `< 1e-12` for beta_hat).  Must run before the env kernel (old k, old window).
 * sap_power_post = the N + 1 power columns of every new observation row (:243-247), the `power_states` buffer field and
 *   the same columns of the agent-input rows; rows of finished envs stay zero.
 * sap_interference_rewards = interference_reward_function (:309-353): rewards of the step into `rewards` (slot k[b]) and
 *   their sum into ep_return; the env kernel is then called with view.rewards unset and a scratch ep_return. */
int sap_real_reset_ex(const SapEnvDims* dims, const float* planes, const float* plane_stats, const float* task_prios,
                      int32_t* k, int32_t* prev, double* ep_return, const SapBatchView* view, int32_t* top_out,
                      const int64_t* prev0, int32_t* nbr_out, int32_t obs_row, void* stream);
int sap_real_step_ex(const SapEnvDims* dims, const float* planes, const float* plane_stats, const float* task_prios,
                     const float* T_trans, double lambda_, const int64_t* actions, int32_t* k, int32_t* prev,
                     double* ep_return, int32_t* counts_out, const SapBatchView* view, int32_t* top_out,
                     const uint8_t* dead, int32_t* nbr_out, int32_t obs_row, void* stream);
int sap_power_pre(const SapEnvDims* dims, const float* planes, const float* task_prios, const int64_t* actions,
                  const int32_t* k, double* power, uint8_t* dead_out, void* stream);
int sap_power_post(const SapEnvDims* dims, const int32_t* k, const double* power, const int32_t* nbr,
                   const SapBatchView* view, const SapField* power_states, int32_t base_cols, int32_t obs_row, void* stream);
int sap_interference_rewards(const SapEnvDims* dims, const float* planes, const float* task_prios,
                             const float* neighbor_matrix, const int32_t* sat_freq_bands, int32_t bands_per_env,
                             double lambda_, const int64_t* actions, const int32_t* k, const int32_t* prev,
                             const double* power, double* ep_return, const SapField* rewards, void* stream);

/* ---- constellation proximities -------------------------------------------------
 * sap_proximities_fov = calc_fov_based_proximities_fast (envs/HighPerformanceConstellationSim.py:308-327) over every
 *   (satellite, task, time step): what get_proximities_for_random_tasks / _for_coverage_tasks (:91-270) produce, given the
 *   propagated satellite positions sat_r[n,3,T] (km, float64, the simulator's sat_rs_over_time layout) and the task
 *   positions task_r[m,3].  Writes the env kernels' plane layout planes[T,n,m] (fp32) and / or the reference layout
 *   prox[n,m,T] (float64).  gaussian_sigma_2 = -(fov^2) / (2 ln 0.05) in the reference (:100-101). */
int sap_proximities_fov(const double* sat_r_n3T, const double* task_r_m3, int32_t n, int32_t m, int32_t T, double fov,
                        double gaussian_sigma_2, float* planes_Tnm, double* prox_nmT, void* stream);

/* ---- MockConstellationEnv ---------------------------------------------------------------
 * sap_mock_reset = MockConstellationEnv.reset (mock_constellation_env.py:94-114); prev0[B,n]
 *                  replaces the np.random.choice draw at :105 (injected).
 * sap_mock_step  = MockConstellationEnv.step (:116-162) + beta_hat (:228-274) + buffer writes. */
int sap_mock_reset(const SapEnvDims* dims, const float* planes, const int64_t* prev0, int32_t* k, int32_t* prev,
                   double* ep_return, const SapBatchView* view, void* stream);
int sap_mock_step(const SapEnvDims* dims, const float* planes, const float* T_trans, double lambda_,
                  const int64_t* actions, int32_t* k, int32_t* prev, double* ep_return, int32_t* counts_out,
                  const SapBatchView* view, void* stream);

/* ---- action selectors ---------------------------------------------------------------------
 * sap_select_epsilon_greedy = EpsilonGreedyActionSelector.select_action
 *   (action_selectors/classic_selectors.py:37-54): mask -> -inf, Bernoulli(eps) explore,
 *   uniform-over-available random action, else first-index argmax.
 * sap_select_filtered_epsilon_greedy = FilteredEpsilonGreedyActionSelector.select_action
 *   (action_selectors/filtered_classic_selectors.py:17-63); `top` replaces its th.topk(beta.sum(-1)).
 * Random draws: injected uniforms (u_* non-null, fp32 in [0,1)) or Philox4x32-10 keyed by
 *   (seed; b*n+i, *episode_ctr, k[b], draw id).  avail null = everything available.
 * eps_dev (nullable): device scalar that overrides `eps`, so a captured CUDA graph follows the schedule. */
int sap_select_epsilon_greedy(const float* q, const uint8_t* avail, int32_t B, int32_t n, int32_t A, float eps,
                              const float* eps_dev,
                              uint64_t seed, const uint64_t* episode_ctr, const int32_t* k, const float* u_explore,
                              const float* u_action, int64_t* actions_out, void* stream);
int sap_select_filtered_epsilon_greedy(const float* q, const int32_t* top, const uint8_t* avail, int32_t B, int32_t n,
                                       int32_t m, int32_t M, float eps, const float* eps_dev, uint64_t seed, const uint64_t* episode_ctr,
                                       const int32_t* k, const float* u_tie, const float* u_explore,
                                       const float* u_action, int64_t* actions_out, void* stream);
/* top-M task indices from a beta tensor [B,n,m,L] (dtype f32|f16), stable (value desc, index asc);
 * replaces th.topk(beta.sum(-1), M).indices (filtered_classic_selectors.py:50). */
int sap_topm_from_beta(const void* beta, int32_t dtype, int32_t B, int32_t n, int32_t m, int32_t L, int32_t M,
                       int32_t* top_out, void* stream);

/* ---- selection + env step in ONE launch (SURVEY.md 7.4 `sap_rollout_step`) -------------------
 * What runners/episode_runner.py:75-84 does per timestep after the agent forward - `action_selector.select_action`
 * (classic_selectors.py:37-54, every action available) followed by `env.step(actions)` (real_constellation_env.py:135-175)
 * - for B environments at once: the CTA of an env selects its n agents' actions (same Philox keys / injected draws as
 * sap_select_epsilon_greedy, so the actions are identical), writes them to `actions_out` [B, n] and steps.
 * top_ahead == null: the full step (sap_real_step; `top`, `plane_stats` required);
 * top_ahead != null: the observation of slot k + 1 was built by sap_real_obs_ahead into this top-M buffer
 *   (sap_real_step_after_obs).
 * Only the shipped configuration on the one-CTA-per-env kernel (sap_rollout_step_ok); other shapes make the two calls. */
typedef struct SapSelectArgs {
  const float* q;              /* [B, n, m] fp32, contiguous                                   */
  const float* eps_dev;        /* nullable: device scalar overriding `eps` (CUDA-graph replays) */
  const uint64_t* episode_ctr; /* nullable device counter                                       */
  const float* u_explore;      /* nullable [B*n] injected uniforms (with u_action)              */
  const float* u_action;
  uint64_t seed;
  float eps;
  int32_t reserved;
} SapSelectArgs;
int sap_rollout_step_ok(const SapEnvDims* dims);
int sap_rollout_step(const SapSelectArgs* sel, const SapEnvDims* dims, const float* planes, const float* plane_stats,
                     const float* T_trans, double lambda_, int64_t* actions_out, int32_t* k, int32_t* prev,
                     double* ep_return, int32_t* counts_out, const SapBatchView* view, int32_t* top,
                     const int32_t* top_ahead, void* stream);

/* sap_sample_categorical = `Categorical(probs).sample()` of the policy-sampling selectors
 *   (action_selectors/classic_selectors.py:15-27 MultinomialActionSelector, :56-64 SoftPoliciesSelector,
 *   filtered_classic_selectors.py:65-102): rows of `A` unnormalised probabilities (masked by `avail` when given).
 *   torch.multinomial's stream cannot be reproduced, so the draw is an input: out[r] = first k with
 *   cdf[k] > u[r] * cdf[A-1], cdf accumulated in float64.  u [rows] fp32 in [0, 1). */
int sap_sample_categorical(const float* probs, const uint8_t* avail, int64_t rows, int32_t A, const float* u, int64_t* out,
                           void* stream);

/* ---- episode buffer -------------------------------------------------------------------------
 * sap_buffer_insert  = ReplayBuffer.insert_episode_batch (components/episode_buffer.py:244-259):
 *   copies `count` episode rows of `row_bytes` each from src (starting at src_row0) into the ring
 *   dst of `ring_rows` rows starting at dst_row0, wrapping around.
 * sap_buffer_gather  = ReplayBuffer.sample's fancy-index copy (:264-271): dst[r] = src[ids[r]].
 * sap_onehot         = OneHot.transform (components/transforms.py:16-19) for materialising
 *   actions_onehot from actions after the fact.
 * sap_real_beta_window = the `beta` buffer field ([rows, n, m, L]) rebuilt from the planes:
 *   row r = (env b = r / (T+1), time t = r % (T+1)), zero for t >= T (real_constellation_env.py:167-170, 226-228). */
int sap_buffer_insert(void* dst, const void* src, int64_t row_bytes, int64_t ring_rows, int64_t dst_row0,
                      int64_t src_row0, int64_t count, void* stream);
int sap_buffer_gather(void* dst, const void* src, const int64_t* ids, int64_t row_bytes, int64_t count, void* stream);
int sap_onehot(const void* actions, int32_t actions_dtype, void* onehot, int32_t onehot_dtype, int64_t rows, int32_t m,
               void* stream);
int sap_real_beta_window(const SapEnvDims* dims, const float* planes, const float* task_prios, void* beta,
                         int32_t dtype, void* stream);
/* sap_real_beta_rows = the same field for a SAMPLED batch: dims->B batch rows, time steps [t0, t0 + t_count), batch row b
 *   reading plane row plane_rows[b] of `planes` [*, T, n, m] (null: b itself / the shared row).  This is what lets a replay
 *   buffer that does not store `beta` (components/episode_buffer.py:264-271 returns it as a stored field) hand it out for
 *   any episode it still holds. */
int sap_real_beta_rows(const SapEnvDims* dims, const float* planes, const float* task_prios, const int64_t* plane_rows,
                       int32_t t0, int32_t t_count, void* beta, int32_t dtype, void* stream);

/* ---- assignment selectors -------------------------------------------------
 * sap_lsa_maximize = the per-env body of SequentialAssignmentProblemSelector.select_action
 *   (action_selectors/sap_selectors.py:77-91) and of EpsilonGreedySAPTestActionSelector's test branch (:26-33):
 *   cols_out[b, :] = scipy.optimize.linear_sum_assignment(Q[b] + z[b] * std[b], maximize=True)[1].
 *   z [B,n,m] standard-normal draws (nullable = no perturbation; torch's th.normal stream cannot be reproduced, so the
 *   draws are an input), std [B] = mean|Q[b]| * eps * 2 (:84-85).  Shortest-augmenting-path solver on float64 duals
 *   (the algorithm scipy implements): an optimal assignment, scipy's own whenever the optimum is unique.
 *   objective_out [B] (nullable) = sum of the chosen perturbed benefits.  Requires n <= m <= 512. */
int sap_lsa_maximize(const float* q, const float* z, const float* std_per_env, int32_t B, int32_t n, int32_t m,
                     int64_t* cols_out, double* objective_out, void* stream);

/* ---- agent-input glue -----------------------------------------------------------------------
 * sap_bias_act = the bias (+ ReLU) epilogue of one agent layer, in place: x[r, c] = act(x[r, c] + bias[c]).
 *   The layer itself stays a torch matmul (modules/agents/rnn_agent.py:22-31: F.relu(self.fc1(inputs))); for the
 *   first layer `mm` + this kernel is one pass cheaper than cuBLAS's beta*C epilogue and bit-identical to it. */
int sap_bias_act(float* x, const float* bias, int64_t rows, int32_t cols, int32_t relu, void* stream);
/* sap_split_bias_act = epilogue of the OPT-IN split-precision first layer (args.agent_fc1 = "fp16_split"): the fp32
 *   weight of fc1 is split into `terms` fp16 pieces, W = W_0 + scale W_1 (+ scale^2 W_2), the fp16 observation rows
 *   (exact: the buffer holds fp16, basic_controller.py:82 only widens them) are multiplied by [W_0 | W_1 | W_2] in ONE
 *   fp16 tensor-core GEMM with fp32 accumulation, and this kernel folds the pieces:
 *   out[r, c] = act(sum_k scale^k y_cat[r, k * cols + c] + bias[c]).  Products are exact in fp32; the result differs
 *   from the fp32 sgemm of the default path only by accumulation order (tests hold it to 2e-6 relative). */
int sap_split_bias_act(const float* y_cat, int32_t terms, float scale, const float* bias, float* out, int64_t rows,
                       int32_t cols, int32_t relu, void* stream);
/* 1 when sap_real_step / sap_real_reset accept an fp16 `agent_in` staging field for these dims (the one-CTA-per-env
 * kernel of the shipped configuration does: it bulk-stores the fp16 rows a second time instead of widening them). */
int sap_real_agent_in_f16_ok(const SapEnvDims* dims);

#ifdef __cplusplus
}
#endif
#endif /* MARL_SAP_B200_H */
