// launch_pair.cu -- instantiation + dispatch of the MLMC pair kernels (pair.cuh)
#include <cmath>
#include <type_traits>

#include "launch.cuh"
#include "pair.cuh"

namespace sdemc {
namespace {

template <class Kernel>
int run(Kernel kernel, const LaunchArgs& a, int fine, int coarse, float* d_terminal) {
  int grid = 0;
  int rc = pick_grid(kernel, 0, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  DevPairOut pout;
  pout.terminal = d_terminal;
  kernel<<<grid, kBlock, 0, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, fine, coarse, pout, a.d_moments,
                                        a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

// moments-only pairs with persistent lanes (pair.cuh): same rule as launch_jump.cu:want_flat with the pair's
// coarse + rate T outer iterations.  sdemc_sde.short_path overrides (OFF: never, anything else but AUTO: always).
bool want_flat_pair(const DevSde& s, int coarse, int short_path) {
  if (short_path != SDEMC_SHORT_AUTO) return short_path != SDEMC_SHORT_OFF;
  const double lam_T = (double)s.rate * (double)s.T;
  return 2.1 * std::sqrt(lam_T) > 0.2 * ((double)coarse + lam_T);
}

template <class C>
int run_flat(const LaunchArgs& a, int fine, int coarse) {
  auto kernel = jump_pair_flat_kernel<C>;
  int grid = 0;
  int rc = pick_grid(kernel, 0, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, kBlock, 0, a.stream>>>(a.sde, a.payoff, a.range, a.keys, fine, coarse, a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

template <class C>
int jump_by_mode(const LaunchArgs& a, int fine, int coarse, float* t) {
  if (!a.use_inject && t == nullptr && want_flat_pair(a.sde, coarse, a.short_path)) return run_flat<C>(a, fine, coarse);
  return a.use_inject ? run(jump_pair_kernel<C, true>, a, fine, coarse, t) : run(jump_pair_kernel<C, false>, a, fine, coarse, t);
}
template <class C>
int diff_by_mode(const LaunchArgs& a, int fine, int coarse, float* t) {
  return a.use_inject ? run(diffusion_pair_kernel<C, true>, a, fine, coarse, t)
                      : run(diffusion_pair_kernel<C, false>, a, fine, coarse, t);
}

template <int FAMILY, int M, int MARKS>
int jump_by_dim(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* t) {
  switch (s.dim) {
    case 1: return jump_by_mode<Cfg<FAMILY, 1, M, MARKS, false>>(a, fine, coarse, t);
    case 2: return jump_by_mode<Cfg<FAMILY, 2, M, MARKS, false>>(a, fine, coarse, t);
    case 3: return jump_by_mode<Cfg<FAMILY, 3, M, MARKS, false>>(a, fine, coarse, t);
    case 4: return jump_by_mode<Cfg<FAMILY, 4, M, MARKS, false>>(a, fine, coarse, t);
  }
  return SDEMC_ERR_UNSUPPORTED;
}
template <int FAMILY, int M>
int diff_by_dim(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* t) {
  switch (s.dim) {
    case 1: return diff_by_mode<Cfg<FAMILY, 1, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 2: return diff_by_mode<Cfg<FAMILY, 2, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 3: return diff_by_mode<Cfg<FAMILY, 3, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 4: return diff_by_mode<Cfg<FAMILY, 4, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace

int launch_pair(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* d_terminal) {
  if (s.family == SDEMC_FAMILY_HESTON || s.asian) return SDEMC_ERR_UNSUPPORTED;
  if (s.marks == SDEMC_MARKS_NONE) {
    if (s.family == SDEMC_FAMILY_GEOMETRIC)
      return s.m == 1 ? diff_by_dim<SDEMC_FAMILY_GEOMETRIC, 1>(s, a, fine, coarse, d_terminal)
                      : diff_by_dim<SDEMC_FAMILY_GEOMETRIC, 2>(s, a, fine, coarse, d_terminal);
    return s.m == 1 ? diff_by_dim<SDEMC_FAMILY_ARITHMETIC, 1>(s, a, fine, coarse, d_terminal)
                    : diff_by_dim<SDEMC_FAMILY_ARITHMETIC, 2>(s, a, fine, coarse, d_terminal);
  }
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 1 && s.marks == SDEMC_MARKS_LOGNORMAL)
    return jump_by_dim<SDEMC_FAMILY_GEOMETRIC, 1, SDEMC_MARKS_LOGNORMAL>(s, a, fine, coarse, d_terminal);
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return jump_by_dim<SDEMC_FAMILY_GEOMETRIC, 2, SDEMC_MARKS_ICDF>(s, a, fine, coarse, d_terminal);
  if (s.family == SDEMC_FAMILY_ARITHMETIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return jump_by_dim<SDEMC_FAMILY_ARITHMETIC, 2, SDEMC_MARKS_ICDF>(s, a, fine, coarse, d_terminal);
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
