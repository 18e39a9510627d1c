// debug_draws.cuh -- test hook: the noise EXACTLY as the kernels compute it from Philox.
//
// The parity tests compare the Philox-driven moments kernels with the CPU oracle path by path.  The oracle needs the
// noise as arrays (the reference's three sampling hooks, solvers.py:51-56,143-148).  Re-deriving it on the CPU from
// the Philox words (oracle/philox_streams.py) reproduces it only to the accuracy of the MUFU units (lg2 / sin / cos
// approximations, ~1e-7), and the jump-adapted loop amplifies a 1e-7 shift of a jump time without bound (a step of
// length dt -> 0 before the jump has sqrt(dt) in its increment; whether t + (T - t) rounds to T decides an extra
// iteration).  This kernel therefore writes the draws with the SAME device functions the kernels call
// (philox_normals6, queue_group_draws, InlineJumps::block_draws, packed_block_draws): the oracle then sees
// bit-identical normals, gaps and marks, and states must agree at 1e-5 with exact iteration counts.  The CPU
// restatement pins this hook in turn (counters and bit maps, to MUFU accuracy).
#pragma once
#include "jump_flat.cuh"

namespace sdemc {

// kind (sdemc_draws_kind): 0 BROWNIAN  a[n][count] = unit normals of STREAM_DIFFUSION in consumption order
//                          1 QUEUE     a = cumulative jump times (fmaf(gap, 1/rate, tau) like queue_refill), b = raw marks
//                          2 INLINE    a = Exp(1) gap candidate of iteration k, b = raw mark candidate
//                          3 PACKED    a = Brownian unit normal of iteration k, b = gap candidate, c = raw mark candidate
template <int MARKS>
__global__ void __launch_bounds__(256) debug_draws_kernel(const DevRange rg, const PhiloxKeys keys, const int kind,
                                                         const int count, const float inv_rate, float* __restrict__ a,
                                                         float* __restrict__ b, float* __restrict__ c) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rg.n_paths; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float* pa = a + i * (uint64_t)count;
    float* pb = b ? b + i * (uint64_t)count : nullptr;
    float* pc = c ? c + i * (uint64_t)count : nullptr;
    if (kind == 0) {
      for (int blk = 0; blk * kNormalsPerBlock < count; ++blk) {
        uint32_t o[4];
        float nrm[kNormalsPerBlock];
        philox4x32_10((uint32_t)blk, STREAM_DIFFUSION, plo, phi, keys, o);
        philox_normals6(o, nrm);
#pragma unroll
        for (int j = 0; j < kNormalsPerBlock; ++j)
          if (blk * kNormalsPerBlock + j < count) pa[blk * kNormalsPerBlock + j] = nrm[j];
      }
    } else if (kind == 1) {
      float tau = 0.0f;
      for (int grp = 0; grp * 4 < count; ++grp) {
        float gap[4], raw[4];
        queue_group_draws<MARKS>((uint32_t)grp, plo, phi, keys, gap, raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tau = fmaf(gap[j], inv_rate, tau);
          if (grp * 4 + j < count) {
            pa[grp * 4 + j] = tau;
            pb[grp * 4 + j] = raw[j];
          }
        }
      }
    } else if (kind == 2) {
      for (int blk = 0; blk * 2 < count; ++blk) {
        float g0, r0, g1, r1;
        InlineJumps<MARKS>::block_draws((uint32_t)blk, plo, phi, keys, g0, r0, g1, r1);
        pa[blk * 2] = g0;
        pb[blk * 2] = r0;
        if (blk * 2 + 1 < count) {
          pa[blk * 2 + 1] = g1;
          pb[blk * 2 + 1] = r1;
        }
      }
    } else {
      for (int blk = 0; blk * 2 < count; ++blk) {
        float lg_z, cs_z[2], gap[2], raw[2];
        packed_block_draws((uint32_t)blk, plo, phi, keys, lg_z, cs_z, gap, raw);
        const float r_z = fast_sqrt(lg_z * -1.3862943611198906f);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (blk * 2 + j < count) {
            pa[blk * 2 + j] = r_z * cs_z[j];
            pb[blk * 2 + j] = gap[j];
            pc[blk * 2 + j] = raw[j];
          }
      }
    }
  }
}

}  // namespace sdemc
