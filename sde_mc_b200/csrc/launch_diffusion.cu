// launch_diffusion.cu -- instantiation + dispatch of the uniform-grid kernels (diffusion.cuh)
#include "diffusion.cuh"
#include "launch.cuh"

namespace sdemc {
namespace {

template <class C, bool HESTON, bool INJECT, bool STORE>
int run(const LaunchArgs& a) {
  auto kernel = diffusion_kernel<C, HESTON, INJECT, STORE>;
  const int block = STORE ? kDiffusionStoreBlock : kBlock;
  const size_t smem = STORE ? (size_t)(block / 32) * 2 * DiffusionStoreWriter<C>::kFloats * sizeof(float) : 0;
  if (smem > 48 * 1024)
    SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 0;
  int rc = pick_grid(kernel, smem, a.range.n_paths, &grid, block);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, block, smem, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, a.out, a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

template <class C, bool HESTON>
int by_mode(const LaunchArgs& a) {
  if (a.store) return a.use_inject ? run<C, HESTON, true, true>(a) : run<C, HESTON, false, true>(a);
  if (a.use_inject) return SDEMC_ERR_UNSUPPORTED;  // injected noise is only offered with stored outputs
  return run<C, HESTON, false, false>(a);
}

template <int FAMILY, int M>
int by_dim(const sdemc_sde& s, const LaunchArgs& a) {
  switch (s.dim) {
    case 1: return by_mode<Cfg<FAMILY, 1, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 2: return by_mode<Cfg<FAMILY, 2, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 3: return by_mode<Cfg<FAMILY, 3, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 4: return by_mode<Cfg<FAMILY, 4, M, SDEMC_MARKS_NONE, false>, false>(a);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace

int launch_diffusion(const sdemc_sde& s, const LaunchArgs& a) {
  if (s.family == SDEMC_FAMILY_HESTON) {
    if (s.dim != 2 || s.m != 1 || s.asian) return SDEMC_ERR_UNSUPPORTED;
    return by_mode<Cfg<SDEMC_FAMILY_HESTON, 2, 1, SDEMC_MARKS_NONE, false>, true>(a);
  }
  if (s.asian) {
    if (s.dim != 2 || s.m != 1) return SDEMC_ERR_UNSUPPORTED;
    if (s.family == SDEMC_FAMILY_GEOMETRIC)
      return by_mode<Cfg<SDEMC_FAMILY_GEOMETRIC, 2, 1, SDEMC_MARKS_NONE, true>, false>(a);
    return by_mode<Cfg<SDEMC_FAMILY_ARITHMETIC, 2, 1, SDEMC_MARKS_NONE, true>, false>(a);
  }
  if (s.family == SDEMC_FAMILY_GEOMETRIC) return s.m == 1 ? by_dim<SDEMC_FAMILY_GEOMETRIC, 1>(s, a) : by_dim<SDEMC_FAMILY_GEOMETRIC, 2>(s, a);
  if (s.family == SDEMC_FAMILY_ARITHMETIC) return s.m == 1 ? by_dim<SDEMC_FAMILY_ARITHMETIC, 1>(s, a) : by_dim<SDEMC_FAMILY_ARITHMETIC, 2>(s, a);
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
