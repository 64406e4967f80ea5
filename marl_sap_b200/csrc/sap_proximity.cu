// Field-of-view proximities of a constellation: the benefit tensor of the real envs when it is not
// given, sat_prox_mat[i, j, k] = calc_fov_based_proximities_fast(sat_r[i, :, k], task_r[j], fov, sigma_2)
// (/root/reference/src/envs/HighPerformanceConstellationSim.py:308-327, evaluated over (sat, task, time) by
// get_proximities_for_random_tasks :91-173 / get_proximities_for_coverage_tasks :175-270; the plane-skipping shortcuts of
// those loops only avoid work whose result is 0 or a time shift of another satellite's row).  Orbit propagation itself
// (poliastro) stays outside: the satellite positions are an input.  One thread per (time, sat, task), float64 like numpy,
// written straight in the env kernels' plane layout [T, n, m] (fp32) - no [n, m, T] detour through the host.
#include "sap_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) sap_proximity_kernel(const double* __restrict__ sat_r /* [n,3,T] */,
                                                                 const double* __restrict__ task_r /* [m,3] */, int n, int m,
                                                                 int T, double fov, double sigma_2, float* __restrict__ planes,
                                                                 double* __restrict__ prox_nmT) {
  const int64_t total = (int64_t)T * n * m;
  for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kThreads) {
    const int j = (int)(e % m);
    const int64_t ti = e / m;
    const int i = (int)(ti % n), k = (int)(ti / n);
    const double sx = sat_r[((int64_t)i * 3 + 0) * T + k], sy = sat_r[((int64_t)i * 3 + 1) * T + k],
                 sz = sat_r[((int64_t)i * 3 + 2) * T + k];
    const double tx = task_r[j * 3 + 0], ty = task_r[j * 3 + 1], tz = task_r[j * 3 + 2];
    double prox = 0.0;
    // can_see (:309-312): the task lies on the satellite's side of the tangent plane through the task
    if (tx * sx + ty * sy + tz * sz > tx * tx + ty * ty + tz * tz) {
      const double dx = tx - sx, dy = ty - sy, dz = tz - sz;  // sat -> task
      const double c = (-sx * dx - sy * dy - sz * dz) / (sqrt(sx * sx + sy * sy + sz * sz) * sqrt(dx * dx + dy * dy + dz * dz));
      const double ang = acos(c) * 57.2957795131;              // degrees, the reference's constant (:317)
      if (ang < fov) prox = exp(-(ang * ang) / (2.0 * sigma_2));
    }
    if (planes) planes[e] = (float)prox;
    if (prox_nmT) prox_nmT[((int64_t)i * m + j) * T + k] = prox;
  }
}

}  // namespace

extern "C" int sap_proximities_fov(const double* sat_r_n3T, const double* task_r_m3, int32_t n, int32_t m, int32_t T, double fov,
                                   double gaussian_sigma_2, float* planes_Tnm, double* prox_nmT, void* stream) {
  SAP_REQUIRE(sat_r_n3T && task_r_m3 && (planes_Tnm || prox_nmT), SAP_E_NULL, "sap_proximities_fov: null pointer");
  SAP_REQUIRE(n > 0 && m > 0 && T > 0, SAP_E_DIMS, "sap_proximities_fov: bad dims n=%d m=%d T=%d", n, m, T);
  SAP_REQUIRE(fov > 0.0 && gaussian_sigma_2 > 0.0, SAP_E_CONSTRAINT, "sap_proximities_fov: fov and sigma_2 must be positive");
  const int64_t total = (int64_t)T * n * m;
  const int64_t blocks = (total + kThreads - 1) / kThreads;
  sap_proximity_kernel<<<(int)(blocks > 148 * 32 ? 148 * 32 : blocks), kThreads, 0, (cudaStream_t)stream>>>(
      sat_r_n3T, task_r_m3, n, m, T, fov, gaussian_sigma_2, planes_Tnm, prox_nmT);
  SAP_CUDA_LAUNCH_CHECK("sap_proximity_kernel");
  return SAP_OK;
}
