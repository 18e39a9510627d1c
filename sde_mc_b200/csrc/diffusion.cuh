// diffusion.cuh -- fused uniform-grid kernels: DiffusionSolver.solve (solvers.py:68-88) + payoff + moments.
#pragma once
#include "engine.cuh"
#include "store_tile.cuh"

namespace sdemc {

// One thread per path, grid-stride over the call's path range.
//   INJECT : unit normals come from DevInject.z (deterministic parity mode) instead of Philox
//   STORE  : write the solve() outputs (paths, normals, payoffs); otherwise accumulate moments only
#ifndef SDEMC_DIFF_MIN_BLOCKS
#define SDEMC_DIFF_MIN_BLOCKS 1
#endif
#ifndef SDEMC_DIFF_STORE_MINB
#define SDEMC_DIFF_STORE_MINB 1
#endif
// The 1-D single-driver moments kernel (GBM, the default benchmark) is bound by the XU pipe (two MUFU per normal).
// Measured, path-steps/s: plain loop 1.557e12 at 4 CTAs per SM (56 registers), 1.569e12 at 5, 1.586e12 at 6 (40
// registers), 1.58e12 at 8 (spills); with the Philox rounds of the next block software-pipelined against the
// Box-Muller transcendentals of the current one (below) 1.614e12 at 4, **1.657e12 at 5 (47 registers)**, 1.538e12 at 6.
#ifndef SDEMC_DIFF_1D_MIN_BLOCKS
#define SDEMC_DIFF_1D_MIN_BLOCKS 5
#endif
template <class C, bool HESTON, bool INJECT, bool STORE>
constexpr int diffusion_min_blocks() {
  if (STORE) return SDEMC_DIFF_STORE_MINB;
  if (INJECT) return 1;
  return (C::DIM == 1 && C::M == 1 && !HESTON && C::FAMILY != SDEMC_FAMILY_USER) ? SDEMC_DIFF_1D_MIN_BLOCKS : SDEMC_DIFF_MIN_BLOCKS;
}
// staging tile of the path-storing mode: 32 elements per path and flush; a step group stages up to
// steps_per_group * max(dim, increments per step) elements
#ifndef SDEMC_DIFF_STORE_TILE
#define SDEMC_DIFF_STORE_TILE 32
#endif
template <class C>
using DiffusionStoreWriter =
    WarpTileWriter<SDEMC_DIFF_STORE_TILE, steps_per_group(C::BASE * C::M) * (C::DIM > C::BASE * C::M + (C::ASIAN ? 1 : 0)
                                                              ? C::DIM
                                                              : C::BASE * C::M + (C::ASIAN ? 1 : 0))>;

constexpr int kDiffusionStoreBlock = 128;  // threads per CTA of the storing kernels (shared tiles limit residency)

// PERPATH (moments mode): also write each path's payoff / iteration count / terminal state (sdemc_mc_moments
// per_path).  A template parameter rather than a run-time test of the pointers because the 1-D kernel runs at its
// register limit (47 at 5 CTAs per SM): keeping the path index alive across the step loop for three predicated
// stores cost 4.6 % of the C2 throughput (measured, 1.586e12 vs 1.662e12).  Both instantiations are the same source;
// tests/test_gpu_fastpath.py ties them: the fp64 sums of the plain launch equal those of the PERPATH launch bit for bit.
template <class C, bool HESTON, bool INJECT, bool STORE, bool PERPATH = false, int RMODE = RANGE_HOST>
__global__ void __launch_bounds__(256, diffusion_min_blocks<C, HESTON, INJECT, STORE>()) diffusion_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const DevOut out,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, M = C::M, BASE = C::BASE;
  constexpr int NZ = BASE * M;                  // normals consumed per step
  constexpr int SPB = steps_per_group(NZ);      // steps served by one group of Philox blocks (6 normals per block)
  constexpr int BPS = blocks_per_group(NZ);     // Philox blocks per group
  constexpr int NBUF = BPS * kNormalsPerBlock;
  constexpr bool FAST1D = DIM == 1 && M == 1 && !HESTON && !INJECT && !STORE && C::FAMILY != SDEMC_FAMILY_USER;
  const int S = s.num_steps;
  range_stage<RMODE>(rg);

  extern __shared__ float diff_store_smem[];  // STORE: two staging tiles per warp (paths, increments)
  using Writer = DiffusionStoreWriter<C>;
  Writer wpaths, wnorm;

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // warp-uniform trip count: in STORE mode all 32 lanes stage their outputs in lock-step, so lanes past the end of
  // the range keep iterating (their rows are never written)
  for (uint64_t wbase = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wbase < range_n<RMODE>(rg); wbase += stride) {
    const uint64_t i = wbase + (threadIdx.x & 31);
    if (!STORE && i >= range_n<RMODE>(rg)) break;
    const bool valid = i < range_n<RMODE>(rg);
    const uint64_t gp = range_lo<RMODE>(rg) + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float x[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) x[d] = d < DIM ? s.x0[d] : 0.0f;
    if (STORE) {
      float* tiles = diff_store_smem + (threadIdx.x >> 5) * (2 * Writer::kFloats);
      wpaths.init(tiles, out.paths, out.pitch_state, wbase, range_n<RMODE>(rg));
      wnorm.init(tiles + Writer::kFloats, out.normals, out.pitch_normals, wbase, range_n<RMODE>(rg));
#pragma unroll
      for (int d = 0; d < DIM; ++d) wpaths.append(x[d]);
    }

    float t_user = 0.0f;  // fp32 clock of the uniform grid (t += h, solvers.py:83-87); only user coefficients read it
    int b_first = 0;
    if (FAST1D && !s.milstein) {
      // 1-D single-driver moments path (GBM / log-GBM): sigma sqrt(h) is folded into the Box-Muller radius and
      // full Philox blocks (6 steps) run without per-step predicates: 2 FFMA per step on top of the normal.
      const int nb_full = S / SPB;
      auto six_steps = [&](const uint32_t(&o)[4]) {
        float r[3], c[3], sn[3];
        philox_polar3(o, s.neg2ln2_b1s2, r, c, sn);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float g0 = fmaf(r[j], c[j], s.ah[0]), g1 = fmaf(r[j], sn[j], s.ah[0]);
          if (C::FAMILY == SDEMC_FAMILY_GEOMETRIC) {
            x[0] = fmaf(x[0], g0, x[0]);
            x[0] = fmaf(x[0], g1, x[0]);
          } else {
            x[0] += g0 + g1;
          }
        }
      };
#ifndef SDEMC_DIFF_NO_PIPELINE
      // software pipeline: the Philox rounds of block b+1 (ALU / FMA pipes) are independent of the Box-Muller
      // transcendentals of block b (XU pipe), so ptxas can interleave them inside one loop body
      uint32_t oa[4], ob[4];
      philox4x32_10(0u, STREAM_DIFFUSION, plo, phi, keys, oa);
      int b = 0;
      for (; b + 2 <= nb_full; b += 2) {
        philox4x32_10((uint32_t)(b + 1), STREAM_DIFFUSION, plo, phi, keys, ob);
        six_steps(oa);
        philox4x32_10((uint32_t)(b + 2), STREAM_DIFFUSION, plo, phi, keys, oa);
        six_steps(ob);
      }
      if (b < nb_full) six_steps(oa);
#else
      for (int b = 0; b < nb_full; ++b) {
        uint32_t o[4];
        philox4x32_10((uint32_t)b, STREAM_DIFFUSION, plo, phi, keys, o);
        six_steps(o);
      }
#endif
      b_first = nb_full;  // the generic loop below finishes the S % 6 remaining steps
    }

    for (int b = b_first; b * SPB < S; ++b) {
      float nrm[NBUF];
      float extra[SPB];  // injected normal of the asian integral component (recorded, never used)
      if (!INJECT) {
#pragma unroll
        for (int r = 0; r < BPS; ++r) {
          uint32_t o[4];
          philox4x32_10((uint32_t)(b * BPS + r), STREAM_DIFFUSION, plo, phi, keys, o);
          philox_normals6(o, nrm + kNormalsPerBlock * r);
        }
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) extra[sp] = 0.0f;
      } else {
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) {
          const int step = b * SPB + sp;
          extra[sp] = 0.0f;
          if (step < S && valid) {
            const float* zp = inj.z + (i * (uint64_t)S + step) * (DIM * M);
#pragma unroll
            for (int q = 0; q < NZ; ++q) nrm[sp * NZ + q] = zp[q];
            if (C::ASIAN) extra[sp] = zp[BASE * M];
          }
        }
      }
      // STORE: the group's outputs are staged at compile-time slots and committed once (store_tile.cuh); steps past
      // the end of the grid stage nothing that is committed.
      constexpr int NPS = BASE * M + (C::ASIAN ? 1 : 0);  // increments recorded per step
#pragma unroll
      for (int sp = 0; sp < SPB; ++sp) {
        const int step = b * SPB + sp;
        if (step < S) {
          float z1[kMaxDim], z2[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
          for (int k = 0; k < BASE; ++k) {
            z1[k] = nrm[sp * NZ + k * M];
            z2[k] = M == 2 ? nrm[sp * NZ + k * M + 1] : 0.0f;
          }
          correlate<C>(s, z1, w1);
          if (M == 2) correlate<C>(s, z2, w2);  // DiffusionSolver: every driver is a correlated dim-vector (:79-81)
          if (HESTON) heston_step_uniform(s, x, w1);
          else euler_step_uniform<C>(s, x, w1, w2, t_user);
          if (C::FAMILY == SDEMC_FAMILY_USER) t_user += s.h0;
          if (STORE) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) wpaths.stage(sp * DIM + d, x[d]);
#pragma unroll
            for (int d = 0; d < BASE; ++d) {
              wnorm.stage(sp * NPS + d * M, w1[d] * s.sqrt_h0);
              if (M == 2) wnorm.stage(sp * NPS + d * M + 1, w2[d] * s.sqrt_h0);
            }
            if (C::ASIAN) wnorm.stage(sp * NPS + BASE * M, extra[sp] * s.sqrt_h0);
          }
        }
      }
      if (STORE) {
        const int done = min(SPB, S - b * SPB);
        wpaths.commit(done * DIM);
        wnorm.commit(done * NPS);
      }
    }

    const float pay = eval_payoff<DIM>(po, x);
    if (STORE) {
      wpaths.flush();
      wnorm.flush();
      if (valid && out.payoffs) out.payoffs[i] = pay;
      if (valid && out.iters) out.iters[i] = S;
      if (valid && out.terminal) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) out.terminal[i * DIM + d] = x[d];
      }
    } else {
      acc.add(pay, po.df * x[0] - s.x0[0], S);  // terminal control  D(T) x_T[0] - x_0[0]  mc.py:337
      if (PERPATH) write_per_path<DIM>(per_path_of_out(out), i, pay, S, x);
    }
  }
  if (!STORE) block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
