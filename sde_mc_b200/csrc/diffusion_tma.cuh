// diffusion_tma.cuh -- path-storing uniform-grid kernel whose trajectories leave the SM through TMA.
//
// Same contract and arithmetic as diffusion_kernel<.., STORE=true> (DiffusionSolver.solve solvers.py:68-88: paths
// (bs, S+1, dim) and increments (bs, S, dim[, m]), row-per-path), different data path.  The 16-byte-store flush of
// store_tile.cuh drains through the LSU queue it shares with the staging STS/LDS and does not overlap the step
// loop (2.82 ms for the 8 GB of a 4e6 x 252 GBM solve()).  Here every warp stages its 32 rows in swizzled tiles
// (tma_gang.cuh): four consecutive elements of a path are collected in registers and written with one 16-byte shared
// store, and one lane hands a full tile to the TMA engine.  No transposing loads, no per-row address arithmetic,
// nothing in the LSU but the staging stores.
//
// What bounds the kernel is how the write stream reaches HBM, not the SM (with the step
// arithmetic stubbed out the time barely moves; per-thread 32-byte stores hit the same ceiling): 128-byte pieces
// scattered over tens of thousands of open rows.  A tile therefore holds W = SDEMC_DIFF_TMA_W sub-tiles of 32
// elements that are issued back to back, W x 128 contiguous bytes per row (a probe that writes the pieces of a row
// back to back reaches 7.2 TB/s against 5.5 for scattered 128-byte pieces, tools/tma_issue_cost_probe.cu), single-
// buffered with the wait deferred to the next store into the tile (the puts of a super-group follow its steps).
//
// The tensor maps declare the whole PITCH as the row length: a box cut by the tensor bound inside a row costs the
// engine several full boxes (2.31 ms with maps that end at the row against 1.73).  Rows past the end of the call are
// clipped by the map, boxes wholly past the row are not issued, so the kernel has no tail paths: it runs whole
// super-groups of steps; columns between the end of a row and its pitch (the caller's padding) receive the surplus
// of the last tile.  Requires a row pitch that is a multiple of 16 bytes and 16-byte aligned bases (the Python layer
// pads the pitch to 128 bytes); launch_diffusion.cu falls back to the store_tile.cuh kernel otherwise.
#pragma once
#include <cuda.h>

#include "engine.cuh"
#include "tma_gang.cuh"

namespace sdemc {

#ifndef SDEMC_DIFF_TMA_W
#define SDEMC_DIFF_TMA_W 2   // sub-tiles of 32 elements per tile: 256 contiguous bytes per row and flush
#endif
#ifndef SDEMC_DIFF_TMA_BLOCK
#define SDEMC_DIFF_TMA_BLOCK 128
#endif
constexpr int kTmaStoreBlock = SDEMC_DIFF_TMA_BLOCK;  // threads per CTA
constexpr int kDiffTmaW = SDEMC_DIFF_TMA_W;

template <class C, bool HESTON, bool INJECT>
__global__ void __launch_bounds__(kTmaStoreBlock)
    diffusion_store_tma_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                               const DevInject inj, const DevOut out, const __grid_constant__ CUtensorMap map_paths,
                               const __grid_constant__ CUtensorMap map_normals, const TmaRows rows_paths,
                               const TmaRows rows_normals, unsigned int* __restrict__ d_sched) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M;
  constexpr int NZ = BASE * M;                        // normals consumed per step
  constexpr int SPB = steps_per_group(NZ);            // steps served by one group of Philox blocks
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  constexpr int NPS = BASE * M + (C::ASIAN ? 1 : 0);  // increments recorded per step
  // super-group: whole Philox groups producing a multiple of four elements in both arrays
  constexpr int SG = tma_lcm(SPB, tma_lcm(4 / tma_gcd(4, DIM), 4 / tma_gcd(4, NPS)));
  constexpr int CP = DIM % 4;                         // path elements carried between super-groups (x0 comes first)
  // both arrays receive the same number of vectors per super-group at the same program points: one gang
  constexpr bool SAME = DIM == NPS && DIM != 4;
  constexpr int W = kDiffTmaW;
  using Gang2 = TmaGang<2, W, true>;   // rare staging code out of line (tma_gang.cuh)
  using Gang1 = TmaGang<1, W, true>;
  const int S = s.num_steps;

  extern __shared__ __align__(1024) uint8_t tma_store_smem[];
  // swizzled tiles sit on kAlign-byte boundaries of the shared window (the launch adds that much slack)
  const uint32_t warp_tiles = (((uint32_t)__cvta_generic_to_shared(tma_store_smem) + Gang2::kAlign - 1u) & ~(Gang2::kAlign - 1u)) +
                              (threadIdx.x >> 5) * Gang2::kBytes;
  Gang2 g_both;
  Gang1 g_paths, g_norm;
  if (SAME) {
    g_both.init(warp_tiles, rows_paths.dcol);  // (the host enables direct columns only when both arrays agree)
    g_both.set_array(0, &map_paths, rows_paths);
    g_both.set_array(1, &map_normals, rows_normals);
  } else {
    g_paths.init(warp_tiles, rows_paths.dcol);
    g_paths.set_array(0, &map_paths, rows_paths);
    g_norm.init(warp_tiles + Gang1::kBytes, rows_normals.dcol);
    g_norm.set_array(0, &map_normals, rows_normals);
  }
  TmaGroups grp;
  grp.init();

  TmaWarpTasks tasks;  // groups of 32 consecutive paths, handed out through the caller's workspace
  for (tasks.init(d_sched, rg.n_paths); tasks.valid(); tasks.advance()) {
    const uint64_t wbase = tasks.first_row();
    const uint64_t i = wbase + (threadIdx.x & 31);
    const bool valid = i < rg.n_paths;
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    if (SAME) {
      g_both.begin_rows((int)wbase);
    } else {
      g_paths.begin_rows((int)wbase);
      g_norm.begin_rows((int)wbase);
    }

    float x[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) x[d] = d < DIM ? s.x0[d] : 0.0f;
    float pb[CP + SG * DIM];  // path elements of a super-group; the first CP entries are carried over (x0 comes first)
    float nb[SG * NPS];
    float t_user = 0.0f;      // fp32 clock of the grid, read by user coefficients only
    // vector v of the path rows / of the increment rows
    auto put_paths = [&](int v) {
      g_paths.begin(grp);
      g_paths.store(0, pb[4 * v], pb[4 * v + 1], pb[4 * v + 2], pb[4 * v + 3]);
      g_paths.end(grp);
    };
    auto put_norm = [&](int v) {
      g_norm.begin(grp);
      g_norm.store(0, nb[4 * v], nb[4 * v + 1], nb[4 * v + 2], nb[4 * v + 3]);
      g_norm.end(grp);
    };
    auto put_both = [&](int v, bool with_normals) {
      g_both.begin(grp);
      g_both.store(0, pb[4 * v], pb[4 * v + 1], pb[4 * v + 2], pb[4 * v + 3]);
      if (with_normals) g_both.store(1, nb[4 * v], nb[4 * v + 1], nb[4 * v + 2], nb[4 * v + 3]);
      g_both.end(grp);
    };
    if (DIM == 4) {
#pragma unroll
      for (int d = 0; d < 4; ++d) pb[d] = x[d];
      put_paths(0);
    }
#pragma unroll
    for (int d = 0; d < CP; ++d) pb[d] = x[d];

    for (int g0 = 0; g0 < S; g0 += SG) {
#pragma unroll
      for (int gi = 0; gi < SG / SPB; ++gi) {
        const int b = g0 / SPB + gi;
        float nrm[NBUF];
        float extra[SPB];
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
#pragma unroll
            for (int q = 0; q < NZ; ++q) nrm[sp * NZ + q] = 0.0f;
            if (step < S && valid) {
              const float* zp = inj.z + (i * (uint64_t)S + step) * (DIM * M);
#pragma unroll
              for (int q = 0; q < NZ; ++q) nrm[sp * NZ + q] = zp[q];
              if (C::ASIAN) extra[sp] = zp[BASE * M];
            }
          }
        }
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) {
          const int ls = gi * SPB + sp;      // step inside the super-group
          const int step = g0 + ls;
          float z1[kMaxDim], z2[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
          for (int k = 0; k < BASE; ++k) {
            z1[k] = nrm[sp * NZ + k * M];
            z2[k] = M == 2 ? nrm[sp * NZ + k * M + 1] : 0.0f;
          }
          correlate<C>(s, z1, w1);
          if (M == 2) correlate<C>(s, z2, w2);  // DiffusionSolver: every driver is a correlated dim-vector (:79-81)
          if (step < S) {                       // steps past the grid only produce columns past the row
            if (HESTON) heston_step_uniform(s, x, w1);
            else euler_step_uniform<C>(s, x, w1, w2, t_user);
            if (C::FAMILY == SDEMC_FAMILY_USER) t_user += s.h0;
          }
#pragma unroll
          for (int d = 0; d < DIM; ++d) pb[CP + ls * DIM + d] = x[d];
#pragma unroll
          for (int d = 0; d < BASE; ++d) {
            nb[ls * NPS + d * M] = w1[d] * s.sqrt_h0;
            if (M == 2) nb[ls * NPS + d * M + 1] = w2[d] * s.sqrt_h0;
          }
          if (C::ASIAN) nb[ls * NPS + BASE * M] = extra[sp] * s.sqrt_h0;
        }
      }
      if (SAME) {
#pragma unroll
        for (int v = 0; v < SG * DIM / 4; ++v) put_both(v, true);
      } else {
#pragma unroll
        for (int v = 0; v < SG * DIM / 4; ++v) put_paths(v);
#pragma unroll
        for (int v = 0; v < SG * NPS / 4; ++v) put_norm(v);
      }
#pragma unroll
      for (int d = 0; d < CP; ++d) pb[d] = pb[SG * DIM + d];
    }
    if (CP > 0) {  // the last elements of the row (+ surplus that lands in the padding or past the row)
#pragma unroll
      for (int d = CP; d < 4; ++d) pb[d] = 0.0f;
      if (SAME) put_both(0, false);
      else put_paths(0);
    }
    if (SAME) {
      if (g_both.direct()) {  // a short last tile: whole sectors written by the lanes themselves (tma_gang.cuh)
        g_both.write_tail(0, out.paths, out.pitch_state, valid);
        g_both.write_tail(1, out.normals, out.pitch_normals, valid);
      }
      g_both.finish(grp);
    } else {
      if (g_paths.direct()) g_paths.write_tail(0, out.paths, out.pitch_state, valid);
      g_paths.finish(grp);
      if (g_norm.direct()) g_norm.write_tail(0, out.normals, out.pitch_normals, valid);
      g_norm.finish(grp);
    }

    const float pay = eval_payoff<DIM>(po, x);
    if (valid && out.payoffs) out.payoffs[i] = pay;
    if (valid && out.iters) out.iters[i] = S;
    if (valid && out.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) out.terminal[i * DIM + d] = x[d];
    }
  }
  tasks.finish();
  grp.drain();  // all bulk stores of this warp complete before the CTA retires
}

}  // namespace sdemc
