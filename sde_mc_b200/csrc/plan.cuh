// plan.cuh -- sizing a run from its pilot ON THE DEVICE.
//
// The reference's run-to-tolerance drivers read the pilot's statistics on the host, compute a trial count and launch
// the main run (find_num_trials mc.py:418-427 -> run_mc :437-440, run_cv_mc :443-467; get_optimal_trials
// mlmc.py:77-97 -> mc_multilevel).  Here a one-thread kernel does that arithmetic on the (all-reduced) pilot moments
// where they lie and writes each launch's path range into device memory; the main-run kernels, already queued on the
// same stream, read their range when they start (sdemc_range.d_range).  Nothing returns to the host in between.
#pragma once
#include "engine.cuh"

namespace sdemc {

// this rank's share of n paths: contiguous, balanced (sde_mc_b200/_engine.py:shard)
__device__ __forceinline__ void shard_of(uint64_t n, int rank, int world, uint64_t& off, uint64_t& cnt) {
  const uint64_t base = n / (uint64_t)world, rem = n % (uint64_t)world;
  cnt = base + ((uint64_t)rank < rem ? 1u : 0u);
  off = (uint64_t)rank * base + ((uint64_t)rank < rem ? (uint64_t)rank : rem);
}

// unbiased sample variance from running sums, as E.mean_and_stderr / helpers.mc_estimates compute it in fp64
__device__ __forceinline__ double pilot_variance(const double* m, double n) {
  const double mean = m[0] / n;
  const double v = m[1] / n - mean * mean;
  return (v > 0.0 ? v : 0.0) * (n / (n - 1.0));
}

// plain MC / control-variate MC:  N = ceil((1.96 se / eps)^2 n_pilot), se^2 = var / n_pilot   (mc.py:418-427), rounded up
// to a multiple of `multiple_of` (ceil_mult, mc.py:459), capped at max_trials.
__global__ void plan_mc_kernel(const double* __restrict__ pilot, double n_pilot, double eps, uint64_t multiple_of,
                               uint64_t max_trials, uint64_t path_base, int rank, int world,
                               uint64_t* __restrict__ range_out, uint64_t* __restrict__ trials_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double se = sqrt(pilot_variance(pilot, n_pilot) / n_pilot);
  const double ratio = se * 1.96 / eps;
  double want = ceil(ratio * ratio * n_pilot);
  if (!(want >= 1.0)) want = 1.0;
  uint64_t n = want < 1.8e19 ? (uint64_t)want : ~0ull;
  if (multiple_of > 1) n = ((n + multiple_of - 1) / multiple_of) * multiple_of;
  if (max_trials && n > max_trials) n = max_trials;
  uint64_t off, cnt;
  shard_of(n, rank, world, off, cnt);
  range_out[0] = path_base + off;
  range_out[1] = cnt;
  trials_out[0] = n;
}

// MLMC:  N_l = ceil(1.96^2 / eps^2 sqrt(V_l h_l) sum_k sqrt(V_k / h_k)),  h_l = T / levels[l]   (mlmc.py:84-96).
// Level l's pairs take the global path ids after those of levels 0..l-1.
__global__ void plan_mlmc_kernel(const double* __restrict__ pilot, int n_levels, const int* __restrict__ levels,
                                 double n_pilot, double T, double eps, uint64_t max_trials, uint64_t path_base, int rank,
                                 int world, uint64_t* __restrict__ ranges_out, uint64_t* __restrict__ trials_out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double total = 0.0;
  for (int l = 0; l < n_levels; ++l) {
    const double* m = pilot + (size_t)l * kNumMoments;
    const double var = (m[1] - m[0] * m[0] / n_pilot) / (n_pilot - 1.0);  // helpers.mc_estimates
    total += sqrt(var / (T / (double)levels[l]));
  }
  uint64_t lo = path_base;
  for (int l = 0; l < n_levels; ++l) {
    const double* m = pilot + (size_t)l * kNumMoments;
    const double var = (m[1] - m[0] * m[0] / n_pilot) / (n_pilot - 1.0);
    double want = ceil((1.96 * 1.96 / (eps * eps)) * sqrt(var * (T / (double)levels[l])) * total);
    if (!(want >= 1.0)) want = 1.0;
    uint64_t n = want < 1.8e19 ? (uint64_t)want : ~0ull;
    if (max_trials && n > max_trials) n = max_trials;
    uint64_t off, cnt;
    shard_of(n, rank, world, off, cnt);
    ranges_out[2 * l] = lo + off;
    ranges_out[2 * l + 1] = cnt;
    trials_out[l] = n;
    lo += n;
  }
}

}  // namespace sdemc
