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
template <class C, bool PERPATH = false, int RMODE = RANGE_HOST>
__global__ void __launch_bounds__(256, 3)
    jump_flat_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                     const DevPerPath pp, double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M, MARKS = C::MARKS;
  constexpr int NZ = BASE + (M == 2 ? 1 : 0);  // normals per iteration
  constexpr int SPB = steps_per_group(NZ);     // iterations served by one group of Philox blocks
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  using Src = InlineJumps<MARKS>;
  range_stage<RMODE>(rg);

  Accum acc;
  acc.zero();
  const int n = s.num_steps;
  const int kcap = 4 * (n + s.max_jumps) + 64;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = i < range_n<RMODE>(rg);

  JumpState st;
  Src src;
  uint32_t plo = 0, phi = 0;
  int b = 0;          // group index inside the current path
  float xs[kMaxDim];  // state at array index num_steps ('terminal' payoff index)
  auto start_path = [&](uint64_t idx) {
    const uint64_t gp = range_lo<RMODE>(rg) + idx;
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
      const float pay = eval_payoff<DIM>(po, xp);
      acc.add(pay, po.df * st.x[0] - s.x0[0], st.k);
      if (PERPATH) write_per_path<DIM>(pp, i, pay, st.k, xp);
      i += stride;
      live = i < range_n<RMODE>(rg);
      if (live) start_path(i);
    }
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

// ------------------------------------------------------------------------------------------------------------------
// 1-D single-driver models with lognormal marks (Merton): everything two consecutive iterations need fits ONE Philox
// block, so the group shrinks from six iterations to two and nothing has to stay aligned across the warp.
//   o0[22:0], o3[15:0]   Box-Muller pair -> the two Brownian normals          (23-bit radius, 16-bit angle: as
//   o1[22:0], o3[31:16]  Box-Muller pair -> the two mark normals               philox_normals6)
//   o2[22:0]             Exp(1) gap candidate of the even iteration            (2^-23 spacing: as exp1_from_bits)
//   o0[31:23] o1[31:23] o2[27:23]  the 23 bits of the odd iteration's gap candidate
// A path of 1 + #jumps iterations then occupies ceil(k/2) groups (2.25 on average for rate T = 3: 89 % of the
// iteration slots do work, against 61 % with groups of six) and draws 2.25 Philox blocks instead of 4.3.
// The iteration itself is jump_iteration of jump.cuh (same hit rule, same fresh-candidate-per-iteration jump
// strategy as InlineJumps), so the law of the paths is that of jump_kernel; the STREAM differs (STREAM_PACKED), i.e.
// same seed, different -- equally distributed -- paths.  Tests: moments against the generic kernel within Monte
// Carlo error at 2e7 paths (mean, variance, iteration count; 1 / 2 / 3 nominal steps, both payoff indices) and the
// MLMC estimate against the Merton series.
template <int MARKS>
struct PackedJumps {
  float tau, J;
  float cand_gap[2], cand_raw[2];
  int parity;
  __device__ __forceinline__ void init() {
    tau = 0.0f;
    J = 0.0f;
  }
  __device__ __forceinline__ void begin_iter(const DevSde&, const PhiloxKeys&, int k) { parity = k & 1; }
  __device__ __forceinline__ void advance(const DevSde& s, const PhiloxKeys&, bool pop) {
    const float t2 = fmaf(parity ? cand_gap[1] : cand_gap[0], s.inv_rate, tau);
    const float j2 = mark_from_raw<MARKS>(s, parity ? cand_raw[1] : cand_raw[0]);
    tau = pop ? t2 : tau;
    J = pop ? j2 : J;
  }
  __device__ __forceinline__ float mark(const DevSde&, int) const { return J; }
};

// One block of STREAM_PACKED = the draws of iterations 2b, 2b + 1: log2 of the Brownian radius uniform and the (cos,
// sin) of its angle (z = sqrt(-2 ln2 lg_z) * cs_z[.]), two Exp(1) gap candidates, two N(0,1) mark candidates.
__device__ __forceinline__ void packed_block_draws(uint32_t b, uint32_t plo, uint32_t phi, const PhiloxKeys& keys,
                                                   float& lg_z, float (&cs_z)[2], float (&gap)[2], float (&raw)[2]) {
  uint32_t o[4];
  philox4x32_10(b, STREAM_PACKED, plo, phi, keys, o);
  const float ang_z = fmaf(angle_bits_to_12(o[3]), 804.247719318987f, -804.247719318987f);
  const float ang_m = fmaf(__uint_as_float((o[3] >> 16) | 0x3f800000u), 804.247719318987f, -804.247719318987f);
  lg_z = fast_lg2(bits_to_u01_open0(o[0]));
  const float r_m = fast_sqrt(fast_lg2(bits_to_u01_open0(o[1])) * -1.3862943611198906f);
  cs_z[0] = fast_cos(ang_z);
  cs_z[1] = fast_sin(ang_z);
  raw[0] = r_m * fast_cos(ang_m);
  raw[1] = r_m * fast_sin(ang_m);
  gap[0] = exp1_from_bits(o[2]);
  gap[1] = exp1_from_bits((o[0] >> 23) | ((o[1] >> 23) << 9) | ((o[2] >> 23) << 18));  // low 23 bits are used
}

// FAST: the iteration in the restated form of jump1d.cuh -- stateless mesh, sigma^2 and dt folded into the Box-Muller
// radius (one square root per iteration), the jump coefficient folded into the mark -- about 16
// instructions instead of the ~45 of the generic jump_iteration (geometric Euler only; Milstein takes the generic
// form).  Same draws, same mesh and hits; the state agrees to fp32 rounding (test: sums to 1e-5, iterations equal).
template <class C, bool FAST, bool PERPATH = false, int RMODE = RANGE_HOST>
__global__ void __launch_bounds__(256, 3)
    jump_flat1d_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                       const DevPerPath pp, double* __restrict__ d_moments, void* __restrict__ d_ws) {
  static_assert(C::DIM == 1 && C::M == 1 && !C::ASIAN && C::MARKS == SDEMC_MARKS_LOGNORMAL, "1-D lognormal-mark models");
  static_assert(!FAST || C::FAMILY == SDEMC_FAMILY_GEOMETRIC, "the restated iteration is the geometric Euler step");
  using Src = PackedJumps<C::MARKS>;
  range_stage<RMODE>(rg);
  Accum acc;
  acc.zero();
  const int n = s.num_steps;
  const int kcap = 4 * (n + s.max_jumps) + 64;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = i < range_n<RMODE>(rg);

  JumpState st;
  Src src;
  uint32_t plo = 0, phi = 0;
  float x_at_n = 0.0f;  // state at array index num_steps ('terminal' payoff index)
  auto start_path = [&](uint64_t idx) {
    const uint64_t gp = range_lo<RMODE>(rg) + idx;
    plo = (uint32_t)gp;
    phi = (uint32_t)(gp >> 32);
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) st.x[d] = d < 1 ? s.x0[d] : 0.0f;
    x_at_n = s.x0[0];
    st.t = 0.0f;
    st.k = 0;
    st.need_pop = true;
    src.init();
  };
  start_path(live ? i : 0);

  StepRecord rec_unused;
  while (live) {
    float lg_z, cs_z[2];
    packed_block_draws((uint32_t)(st.k >> 1), plo, phi, keys, lg_z, cs_z, src.cand_gap, src.cand_raw);  // st.k is even here
    if constexpr (FAST) {
      const float r2 = lg_z * (-1.3862943611198906f * s.b1[0] * s.b1[0]);  // sigma^2 (-2 ln u)
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        if (st.t < s.T && st.k < kcap) {  // `while t < T` of the reference, per path (:182)
          // fresh (gap, mark) candidate, taken when the previous jump has been consumed
          const float t2 = fmaf(src.cand_gap[sp], s.inv_rate, src.tau);
          const float j2 = s.c[0] * mark_from_raw<C::MARKS>(s, src.cand_raw[sp]);
          src.tau = st.need_pop ? t2 : src.tau;
          src.J = st.need_pop ? j2 : src.J;
          const float dt = fmaxf(fminf(s.h0, fminf(src.tau, s.T) - st.t), 0.0f);   // stateless mesh (jump1d.cuh)
          const float g = fmaf(fast_sqrt(r2 * dt), cs_z[sp], s.a[0] * dt);
          const float xn = fmaf(st.x[0], g, st.x[0]);
          st.t += dt;
          const bool hit = fabsf(src.tau - st.t) <= fmaf(st.t, 1e-5f, 1e-12f);    // isclose(tau, t) as in jump.cuh
          const float Jc = hit ? src.J : 0.0f;
          st.x[0] = fmaf(s.exact_jumps ? xn : st.x[0], Jc, xn);
          st.need_pop = hit;
          ++st.k;
          if (st.k == n) x_at_n = st.x[0];
        }
      }
    } else {
      const float r_z = fast_sqrt(lg_z * -1.3862943611198906f);
      const float z[2] = {r_z * cs_z[0], r_z * cs_z[1]};
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        if (st.t < s.T && st.k < kcap) {  // `while t < T` of the reference, per path (:182)
          jump_iteration<C, Src, false>(s, keys, st, src, z + sp, rec_unused);
          if (st.k == n) x_at_n = st.x[0];
        }
      }
    }
    if (!(st.t < s.T) || st.k >= kcap) {
      float xp[kMaxDim];
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) xp[d] = 0.0f;
      xp[0] = (po.index_mode == SDEMC_INDEX_TERMINAL && st.k >= n) ? x_at_n : st.x[0];
      const float pay = eval_payoff<1>(po, xp);
      acc.add(pay, po.df * st.x[0] - s.x0[0], st.k);
      if (PERPATH) write_per_path<1>(pp, i, pay, st.k, xp);
      i += stride;
      live = i < range_n<RMODE>(rg);
      if (live) start_path(i);
    }
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
