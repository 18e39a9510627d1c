// diffusion_tma.cuh -- path-storing uniform-grid kernel whose trajectories leave the SM through TMA.
//
// Same contract and arithmetic as diffusion_kernel<.., STORE=true> (DiffusionSolver.solve solvers.py:68-88: paths
// (bs, S+1, dim) and increments (bs, S, dim[, m]), row-per-path), different data path.  The 16-byte-store flush of
// store_tile.cuh drains through the LSU queue it shares with the staging STS/LDS and does not overlap the step
// loop (2.82 ms for the 8 GB of a 4e6 x 252 GBM solve()).  Through TMA: 1.73 ms (4.7 TB/s, 0.72 of the measured copy
// bandwidth); with the step arithmetic stubbed out 1.66 ms -- what remains is how fast 128-byte pieces scattered over
// ~57 000 open rows reach HBM (per-thread 32-byte stores hit the same ~4.8 TB/s, tools/thread_row_store_probe.cu),
// not the step loop.  The tensor maps declare the whole PITCH as the row length: a box cut by the tensor bound
// inside a row costs the engine several full boxes (2.31 ms with maps that end at the row).
// Here every warp keeps a
// [32 paths][32 elements] tile per output array in shared memory in the layout TMA reads (128-byte rows, 128B
// swizzle), fills it with 16-byte vector stores -- four consecutive elements of a path are collected in registers;
// lane q writes chunk (v ^ (q & 7)) of row q, conflict free per quarter warp -- and one lane hands the full tile to
// the TMA engine with a single cp.async.bulk.tensor.2d.global.shared::cta.  No transposing loads, no per-row
// address arithmetic, nothing in the LSU but the staging stores; tiles are double buffered and reused after
// cp.async.bulk.wait_group.read.  Rows past the end of the call and columns past the end of a row are clipped by
// the tensor map's bounds, so the kernel has no tail paths: it runs whole super-groups of steps and lets TMA drop
// what does not exist; columns between the end of a row and its pitch (the caller's padding) receive the surplus of
// the last tile.  Requires a row pitch that is a multiple of 16 bytes and 16-byte aligned bases (the Python layer
// pads the pitch to 128 bytes); launch_diffusion.cu falls back to the store_tile.cuh kernel otherwise.
#pragma once
#include <cuda.h>

#include "engine.cuh"

namespace sdemc {

#ifndef SDEMC_TMA_TILE_ELEMS
#define SDEMC_TMA_TILE_ELEMS 32
#endif
constexpr int kTmaStoreBlock = 128;                     // threads per CTA
constexpr int kTmaTileElems = SDEMC_TMA_TILE_ELEMS;     // elements per path and tile: 32 (128-byte rows, 128B swizzle)
                                                        // or 16 (64-byte rows, 64B swizzle)
static_assert(kTmaTileElems == 32 || kTmaTileElems == 16, "tile rows of 128 or 64 bytes");
constexpr int kTmaTileBytes = 32 * kTmaTileElems * 4;

__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, uint32_t smem, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col),
               "r"(row), "r"(smem)
               : "memory");
}

// One output array of one warp: two swizzled tiles, filled four elements at a time, flushed through TMA.
struct TmaRowWriter {
  uint32_t cur;        // shared-window address of the tile being filled
  uint32_t toggle;     // buf0 ^ buf1
  uint32_t lane_row;   // this lane's row inside a tile (q * row bytes)
  uint32_t swz;        // chunk swizzle of the row: q & 7 (128B mode) or (q >> 1) & 3 (64B mode)
  const CUtensorMap* map;
  int seq_other;       // bulk-group number of the last copy out of the tile NOT being filled (0: none yet)
  int vec;             // 16-byte vectors staged in the current tile (warp-uniform)
  int col;             // first element (column) of the current tile
  int row0;            // row of lane 0

  // once per kernel
  __device__ __forceinline__ void init(uint32_t tiles_s, const CUtensorMap* m) {
    const uint32_t q = threadIdx.x & 31;
    cur = tiles_s;
    toggle = tiles_s ^ (tiles_s + kTmaTileBytes);
    lane_row = q * (kTmaTileElems * 4u);
    swz = kTmaTileElems == 32 ? (q & 7u) : ((q >> 1) & 3u);
    map = m;
    seq_other = 0;
    vec = 0;
    col = 0;
    row0 = 0;
  }
  // start of the rows of the next group of 32 paths (the previous rows were finish()ed)
  __device__ __forceinline__ void begin_rows(int first_row) {
    col = 0;
    row0 = first_row;
  }
  // `issued`: bulk groups this thread has committed so far, shared by all writers of the warp.  The tile we switch
  // to was last copied out as group seq_other; only groups newer than that may still be reading shared memory.
  __device__ __forceinline__ void flush(int& issued) {
#ifndef SDEMC_TMA_NO_FENCE   // (timing experiments only: the copy may then read stale shared memory)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my staging stores -> visible to the TMA engine
#endif
    __syncwarp();
    const int my_seq = ++issued;
#ifdef SDEMC_TMA_NO_COPY     // (timing experiments only: nothing is written)
    if (false) {
#else
    if ((threadIdx.x & 31) == 0) {
#endif
      tma_store_tile(map, cur, col, row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      const int allowed = my_seq - seq_other;  // wait_group takes an immediate; fewer pending than allowed is safe
#ifdef SDEMC_TMA_NO_WAIT      // (timing experiments only: tiles may be overwritten while still being copied)
      if (false) {}
      else
#endif
      if (allowed >= 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      else if (allowed == 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      else if (allowed == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    __syncwarp();
    seq_other = my_seq;
    cur ^= toggle;
    vec = 0;
    col += kTmaTileElems;
  }
  // four consecutive elements of this lane's path
  __device__ __forceinline__ void put4(float a, float b, float c, float d, int& issued) {
    const uint32_t addr = cur + lane_row + ((((uint32_t)vec) ^ swz) << 4);
#ifdef SDEMC_TMA_NO_STS      // (timing experiments only)
    if (a == 123.456f && b == c && d == 7.0f)
#endif
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    if (++vec == kTmaTileElems / 4) flush(issued);
  }
  // end of the rows of this group of 32 paths
  __device__ __forceinline__ void finish(int& issued) {
    if (vec > 0) flush(issued);
  }
};

constexpr int tma_gcd(int a, int b) { return b == 0 ? a : tma_gcd(b, a % b); }
constexpr int tma_lcm(int a, int b) { return a / tma_gcd(a, b) * b; }

template <class C, bool HESTON, bool INJECT>
__global__ void __launch_bounds__(kTmaStoreBlock)
    diffusion_store_tma_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                               const DevInject inj, const DevOut out, const __grid_constant__ CUtensorMap map_paths,
                               const __grid_constant__ CUtensorMap map_normals) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M;
  constexpr int NZ = BASE * M;                        // normals consumed per step
  constexpr int SPB = steps_per_group(NZ);            // steps served by one group of Philox blocks
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  constexpr int NPS = BASE * M + (C::ASIAN ? 1 : 0);  // increments recorded per step
  // super-group: whole Philox groups producing a multiple of four elements in both arrays
  constexpr int SG = tma_lcm(SPB, tma_lcm(4 / tma_gcd(4, DIM), 4 / tma_gcd(4, NPS)));
  constexpr int CP = DIM % 4;                         // path elements carried between super-groups (x0 comes first)
  const int S = s.num_steps;

  extern __shared__ __align__(1024) uint8_t tma_store_smem[];
  // 128B-swizzled tiles must sit on 1024-byte boundaries of the shared window (the launch adds 1 KB of slack)
  const uint32_t warp_tiles = (((uint32_t)__cvta_generic_to_shared(tma_store_smem) + 1023u) & ~1023u) +
                              (threadIdx.x >> 5) * (4u * kTmaTileBytes);
  TmaRowWriter wp, wn;
  wp.init(warp_tiles, &map_paths);
  wn.init(warp_tiles + 2u * kTmaTileBytes, &map_normals);
  int issued = 0;  // bulk groups committed by this thread

  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t wbase = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wbase < rg.n_paths; wbase += stride) {
    const uint64_t i = wbase + (threadIdx.x & 31);
    const bool valid = i < rg.n_paths;
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    wp.begin_rows((int)wbase);
    wn.begin_rows((int)wbase);

    float x[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) x[d] = d < DIM ? s.x0[d] : 0.0f;
    float carry[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // CP pending path elements
    float t_user = 0.0f;                          // fp32 clock of the grid, read by user coefficients only
    if (DIM == 4) wp.put4(x[0], x[1], x[2], x[3], issued);
#pragma unroll
    for (int d = 0; d < CP; ++d) carry[d] = x[d];

    for (int g0 = 0; g0 < S; g0 += SG) {
      float pb[CP + SG * DIM];
      float nb[SG * NPS];
#pragma unroll
      for (int d = 0; d < CP; ++d) pb[d] = carry[d];
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
          if (step < S) {                       // steps past the grid only produce columns the tensor map clips
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
#pragma unroll
      for (int v = 0; v < SG * DIM / 4; ++v) wp.put4(pb[4 * v], pb[4 * v + 1], pb[4 * v + 2], pb[4 * v + 3], issued);
#pragma unroll
      for (int v = 0; v < SG * NPS / 4; ++v) wn.put4(nb[4 * v], nb[4 * v + 1], nb[4 * v + 2], nb[4 * v + 3], issued);
#pragma unroll
      for (int d = 0; d < CP; ++d) carry[d] = pb[SG * DIM + d];
    }
    if (CP > 0) wp.put4(carry[0], carry[1], carry[2], carry[3], issued);  // the last elements of the row (+ clipped slack)
    wp.finish(issued);
    wn.finish(issued);

    const float pay = eval_payoff<DIM>(po, x);
    if (valid && out.payoffs) out.payoffs[i] = pay;
    if (valid && out.iters) out.iters[i] = S;
    if (valid && out.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) out.terminal[i * DIM + d] = x[d];
    }
  }
  // all bulk stores of this thread complete before the CTA retires
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace sdemc
