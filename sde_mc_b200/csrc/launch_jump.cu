// launch_jump.cu -- instantiation + dispatch of the jump-adapted MOMENTS kernels (jump.cuh, jump1d.cuh, jump_flat.cuh);
// the path-storing kernels are instantiated in launch_jump_store.cu (a translation unit of its own: build time)
#include <type_traits>

#include <algorithm>
#include <cmath>

#include "jump1d.cuh"
#include "debug_draws.cuh"
#include "jump_flat.cuh"
#include "launch.cuh"

namespace sdemc {
namespace {

template <class C, int JSRC, bool STORE>
int run(const LaunchArgs& a) {
  auto kernel = jump_kernel<C, JSRC, STORE>;
  const int block = STORE ? kJumpStoreBlock : kBlock;
  const size_t smem = (JSRC == JSRC_QUEUE ? (size_t)a.qdepth * block * sizeof(float2) : 0) +
                      (STORE ? (size_t)(block / 32) * 5 * JumpStoreWriter<C>::kFloats * sizeof(float) : 0);
  if (smem > 48 * 1024) SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 0;
  int rc = pick_grid(kernel, smem, a.range.n_paths, &grid, block);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, block, smem, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, a.out, a.qdepth, a.d_moments,
                                           a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

// moments-only fast path for 1-D single-driver models with queued (sparse) jumps: jump1d.cuh
template <class C, bool EXACT, bool TERM, bool SAT>
int run_1d_inst(const LaunchArgs& a) {
  auto kernel = jump1d_kernel<C, EXACT, TERM, SAT>;
  const size_t smem = (size_t)(a.qdepth + kQueueSlack) * kBlock * sizeof(float2);
  // always opt in: the kernel also has ~17 KB of static shared memory, so the 48 KB default can be exceeded by
  // dynamic sizes below 48 KB
  SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 0;
  int rc = pick_grid(kernel, smem, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  // groups of six iterations that end at least 1.5 h0 before T (t never exceeds its nominal grid point, up to the
  // rounding of at most num_steps additions): dt needs no cap at T there (jump1d.cuh)
  const int n = a.sde.num_steps;
  int nb_uncapped = (int)std::floor(((double)n - 1.5) / kNormalsPerBlock);
  nb_uncapped = std::max(0, std::min(nb_uncapped, n / kNormalsPerBlock));
  kernel<<<grid, kBlock, smem, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.qdepth, nb_uncapped,
                                           per_path_of_out(a.out), a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}
template <class C, bool EXACT>
int run_1d(const LaunchArgs& a) {
  const bool term = a.payoff.index_mode == SDEMC_INDEX_TERMINAL;
  const bool sat = a.sde.h0 <= 1.0f;  // the dt >= 0 clamp as FADD.SAT needs h0 <= 1
  if (term) return sat ? run_1d_inst<C, EXACT, true, true>(a) : run_1d_inst<C, EXACT, true, false>(a);
  return sat ? run_1d_inst<C, EXACT, false, true>(a) : run_1d_inst<C, EXACT, false, false>(a);
}

// moments-only kernel for short paths with inline jumps: lanes are persistent workers (jump_flat.cuh).  The per-path
// hook and the device-resident range are compile-time variants here: a whole path is ~340 instructions.
template <class Kernel>
int launch_flat(Kernel kernel, const LaunchArgs& a) {
  int grid = 0;
  int rc = pick_grid(kernel, 0, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, kBlock, 0, a.stream>>>(a.sde, a.payoff, a.range, a.keys, per_path_of_out(a.out), a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}
template <class C>
int run_flat(const LaunchArgs& a) {
  if (per_path_of_out(a.out).any()) return launch_flat(jump_flat_kernel<C, true, RANGE_HOST>, a);
  if (a.range.dyn) return launch_flat(jump_flat_kernel<C, false, RANGE_DEVICE>, a);
  return launch_flat(jump_flat_kernel<C, false, RANGE_HOST>, a);
}

// 1-D lognormal-mark models: two iterations per Philox block, no alignment across the warp (jump_flat.cuh)
template <class C, bool FAST>
int run_flat1d(const LaunchArgs& a) {
  if (per_path_of_out(a.out).any()) return launch_flat(jump_flat1d_kernel<C, FAST, true, RANGE_HOST>, a);
  if (a.range.dyn && a.range.count_on_host) return launch_flat(jump_flat1d_kernel<C, FAST, false, RANGE_LO_DEVICE>, a);
  if (a.range.dyn) return launch_flat(jump_flat1d_kernel<C, FAST, false, RANGE_DEVICE>, a);
  return launch_flat(jump_flat1d_kernel<C, FAST, false, RANGE_HOST>, a);
}

// A warp of jump_kernel runs until its slowest lane is done: E[max of 32 Poisson(rate T)] exceeds the mean by about
// 2.1 sqrt(rate T) iterations.  When that is more than a fifth of a path's num_steps + rate T iterations the
// persistent-lane kernel wins (MLMC level 0: 2.1 * 1.7 / 4).  sdemc_sde.short_path overrides the rule.
bool want_flat(const sdemc_sde& s) {
  const double lam_T = (double)s.rate * (double)s.T;
  return 2.1 * std::sqrt(lam_T) > 0.2 * ((double)s.num_steps + lam_T);
}

template <class C>
int by_mode(const LaunchArgs& a) {
  if constexpr (C::DIM == 1 && C::M == 1 && !C::ASIAN) {
    if (!a.use_inject && !a.store && a.qdepth > 0 && !a.sde.milstein)
      return a.sde.exact_jumps ? run_1d<C, true>(a) : run_1d<C, false>(a);
  }
  if (a.use_inject) return SDEMC_ERR_UNSUPPORTED;  // injected noise is only offered with stored outputs
  if (a.short_path != SDEMC_SHORT_OFF && !a.store && a.qdepth == 0) {
    if (a.short_path == SDEMC_SHORT_ALIGNED) return run_flat<C>(a);
    // PACKED / PACKED_GENERIC: the stream of its own, 1-D lognormal-mark models only
    if constexpr (C::DIM == 1 && C::M == 1 && !C::ASIAN && C::MARKS == SDEMC_MARKS_LOGNORMAL) {
      if constexpr (C::FAMILY == SDEMC_FAMILY_GEOMETRIC) {
        if (a.short_path == SDEMC_SHORT_PACKED && !a.sde.milstein) return run_flat1d<C, true>(a);
      }
      return run_flat1d<C, false>(a);
    }
    return SDEMC_ERR_UNSUPPORTED;
  }
  if (a.qdepth > 0) return run<C, JSRC_QUEUE, false>(a);
  return run<C, JSRC_INLINE, false>(a);
}

template <int FAMILY, int M, int MARKS>
int by_dim(const sdemc_sde& s, const LaunchArgs& a) {
  switch (s.dim) {
    case 1: return by_mode<Cfg<FAMILY, 1, M, MARKS, false>>(a);
    case 2: return by_mode<Cfg<FAMILY, 2, M, MARKS, false>>(a);
    case 3: return by_mode<Cfg<FAMILY, 3, M, MARKS, false>>(a);
    case 4: return by_mode<Cfg<FAMILY, 4, M, MARKS, false>>(a);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace

int launch_debug_draws(const sdemc_sde& s, const DevRange& rg, const PhiloxKeys& keys, int kind, int count, float* a,
                       float* b, float* c, cudaStream_t stream) {
  const float inv_rate = s.rate > 0.0f ? 1.0f / s.rate : 0.0f;
  const unsigned grid = (unsigned)std::min<uint64_t>((rg.n_paths + 255) / 256, 148 * 8);
  if (s.marks == SDEMC_MARKS_ICDF)
    debug_draws_kernel<SDEMC_MARKS_ICDF><<<grid, 256, 0, stream>>>(rg, keys, kind, count, inv_rate, a, b, c);
  else
    debug_draws_kernel<SDEMC_MARKS_LOGNORMAL><<<grid, 256, 0, stream>>>(rg, keys, kind, count, inv_rate, a, b, c);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

int launch_jump(const sdemc_sde& s, const LaunchArgs& a_in) {
  LaunchArgs a = a_in;
  if (a.qdepth < 0 || (a.qdepth & 3) || a.qdepth > 64) return SDEMC_ERR_BAD_ARG;
  if (a.store) return launch_jump_store(s, a);
  // resolve AUTO.  mc_moments keeps the Philox streams of the path-storing kernels (ALIGNED), so the batched and the
  // one-shot estimators of one seed agree; the MLMC single-level call takes the packed stream where it exists.
  if (a.short_path == SDEMC_SHORT_AUTO) {
    const bool packable = s.dim == 1 && s.m == 1 && !s.asian && s.marks == SDEMC_MARKS_LOGNORMAL;
    a.short_path = !want_flat(s) ? SDEMC_SHORT_OFF
                                 : ((a.prefer_packed && packable) ? SDEMC_SHORT_PACKED : SDEMC_SHORT_ALIGNED);
  }
  if (s.asian) {
    if (s.dim == 2 && s.m == 1 && s.family == SDEMC_FAMILY_GEOMETRIC && s.marks == SDEMC_MARKS_LOGNORMAL)
      return by_mode<Cfg<SDEMC_FAMILY_GEOMETRIC, 2, 1, SDEMC_MARKS_LOGNORMAL, true>>(a);
    return SDEMC_ERR_UNSUPPORTED;
  }
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 1 && s.marks == SDEMC_MARKS_LOGNORMAL)
    return by_dim<SDEMC_FAMILY_GEOMETRIC, 1, SDEMC_MARKS_LOGNORMAL>(s, a);
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return by_dim<SDEMC_FAMILY_GEOMETRIC, 2, SDEMC_MARKS_ICDF>(s, a);
  if (s.family == SDEMC_FAMILY_ARITHMETIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return by_dim<SDEMC_FAMILY_ARITHMETIC, 2, SDEMC_MARKS_ICDF>(s, a);
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
