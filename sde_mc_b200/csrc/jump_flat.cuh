// jump_flat.cuh -- moments-only jump-adapted Euler loop for SHORT paths (JumpDiffusionSolver.solve
// solvers.py:164-226 with low_storage semantics + payoff + moments), e.g. the coarsest MLMC level: mlmc.py:44-53 runs
// 2.4e8 paths of ONE nominal step with ~3 jumps each for BASELINE config C5.
//
// jump.cuh gives every lane one path and lets the warp run until its slowest lane is done: with num_steps + #jumps
// iterations per path the warp pays max over 32 lanes of Poisson(rate T), which for a handful of nominal steps is
// two to three times the mean.  Here a lane is a persistent worker instead: at every group boundary (one group of
// Philox blocks = steps_per_group iterations, as in jump.cuh) a lane whose path reached T folds its payoff into the
// running sums and starts its next path (path ids i, i + stride, ...), while its neighbours carry on with theirs.
// Groups stay aligned across the warp, so the Philox calls are executed once per group for all lanes, and every
// path consumes exactly the counters (stream, block, path id) it consumes in jump_kernel: the two kernels produce
// bit-identical paths, only the order of the fp64 additions differs.
#pragma once
#include "jump.cuh"

namespace sdemc {

// 3 CTAs per SM: 80 registers keep the per-path setup free of spills (measured +5 % on the C5 pass against 4)
template <class C>
__global__ void __launch_bounds__(256, 3)
    jump_flat_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                     double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M, MARKS = C::MARKS;
  constexpr int NZ = BASE + (M == 2 ? 1 : 0);  // normals per iteration
  constexpr int SPB = steps_per_group(NZ);     // iterations served by one group of Philox blocks
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  using Src = InlineJumps<MARKS>;

  Accum acc;
  acc.zero();
  const int n = s.num_steps;
  const int kcap = 4 * (n + s.max_jumps) + 64;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = i < rg.n_paths;

  JumpState st;
  Src src;
  uint32_t plo = 0, phi = 0;
  int b = 0;          // group index inside the current path
  float xs[kMaxDim];  // state at array index num_steps ('terminal' payoff index)
  auto start_path = [&](uint64_t idx) {
    const uint64_t gp = rg.path_lo + idx;
    plo = (uint32_t)gp;
    phi = (uint32_t)(gp >> 32);
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d] = d < DIM ? s.x0[d] : 0.0f;
    st.t = 0.0f;
    st.k = 0;
    st.need_pop = true;
    src.init(plo, phi);
    b = 0;
  };
  start_path(live ? i : 0);

  StepRecord rec_unused;
  while (live) {
    float nrm[NBUF];
#pragma unroll
    for (int r = 0; r < BPS; ++r) {
      uint32_t o[4];
      philox4x32_10((uint32_t)(b * BPS + r), STREAM_DIFFUSION, plo, phi, keys, o);
      philox_normals6(o, nrm + kNormalsPerBlock * r);
    }
    ++b;
#pragma unroll
    for (int sp = 0; sp < SPB; ++sp) {
      if (st.t < s.T && st.k < kcap) {  // `while t < T` of the reference, per path (:182)
        jump_iteration<C, Src, false>(s, keys, st, src, nrm + sp * NZ, rec_unused);
        if (st.k == n) {
#pragma unroll
          for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
        }
      }
    }
    if (!(st.t < s.T) || st.k >= kcap) {
      float xp[kMaxDim];
      // a path that stopped before array index num_steps idles there with dt = 0: index num_steps holds the final state
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) xp[d] = (po.index_mode == SDEMC_INDEX_TERMINAL && st.k >= n) ? xs[d] : st.x[d];
      acc.add(eval_payoff<DIM>(po, xp), po.df * st.x[0] - s.x0[0], st.k);
      i += stride;
      live = i < rg.n_paths;
      if (live) start_path(i);
    }
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
