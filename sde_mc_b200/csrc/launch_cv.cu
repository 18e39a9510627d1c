// launch_cv.cu -- instantiation + dispatch of the fused control-variate kernel (cv.cuh)
#include <type_traits>

#include "cv.cuh"
#include "launch.cuh"

namespace sdemc {
namespace {

template <class Kernel>
int run(Kernel kernel, const LaunchArgs& a, const DevMlp& f, const DevMlp& g, const DevCv& cv) {
  SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvSmemBytes));
  int dev = 0, sms = 0, per_sm = 0;
  SDEMC_CUDA_CHECK(cudaGetDevice(&dev));
  SDEMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // two CTAs per SM: 2 x ~105 KB of shared memory, 2 x 256 of the SM's 512 TMEM columns
  // (the occupancy API under-reports co-residency of large-shared-memory CTAs here; measured)
  per_sm = (227 * 1024) / (kCvSmemBytes + 4096);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 512 / kCvTmemCols) per_sm = 512 / kCvTmemCols;
  uint64_t grid = (uint64_t)sms * per_sm;
  const uint64_t tiles = (a.range.n_paths + kCvRows - 1) / kCvRows;
  const uint64_t pairs = (tiles + kCvTiles - 1) / kCvTiles;
  if (pairs < grid) grid = pairs;
  kernel<<<(unsigned)grid, kCvThreads, kCvSmemBytes, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, f, g, cv,
                                                                a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

}  // namespace

int launch_cv(const sdemc_sde& s, const LaunchArgs& a, const DevMlp& f, const DevMlp& g, const DevCv& cv) {
  if (s.asian || s.family != SDEMC_FAMILY_GEOMETRIC || s.scheme != SDEMC_SCHEME_EULER) return SDEMC_ERR_UNSUPPORTED;
  // the 2-D 'indep' exp-Levy SDE of levy_rainbow_cv_experiment.py:39-40: f = Mlp(3, .., 4), g = Mlp(3, .., 2)
  if (s.dim == 2 && s.m == 2 && s.marks == SDEMC_MARKS_ICDF) {
    using C = Cfg<SDEMC_FAMILY_GEOMETRIC, 2, 2, SDEMC_MARKS_ICDF, false>;
    return a.use_inject ? run(cv_kernel<C, true, true>, a, f, g, cv) : run(cv_kernel<C, true, false>, a, f, g, cv);
  }
  if (s.dim != 1 || s.m != 1) return SDEMC_ERR_UNSUPPORTED;
  if (s.marks == SDEMC_MARKS_LOGNORMAL) {
    using C = Cfg<SDEMC_FAMILY_GEOMETRIC, 1, 1, SDEMC_MARKS_LOGNORMAL, false>;
    return a.use_inject ? run(cv_kernel<C, true, true>, a, f, g, cv) : run(cv_kernel<C, true, false>, a, f, g, cv);
  }
  if (s.marks == SDEMC_MARKS_NONE) {
    using C = Cfg<SDEMC_FAMILY_GEOMETRIC, 1, 1, SDEMC_MARKS_NONE, false>;
    return a.use_inject ? run(cv_kernel<C, false, true>, a, f, g, cv) : run(cv_kernel<C, false, false>, a, f, g, cv);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
