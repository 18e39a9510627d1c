// jump_tma.cuh -- path-storing jump-adapted kernel whose five arrays leave the SM through TMA.
//
// Same contract, arithmetic and Philox streams as jump_kernel<.., STORE=true> (JumpDiffusionSolver.solve
// solvers.py:164-226: paths / left_paths / jump_paths (bs, S+1, dim), time_paths (bs, S+1), normals (bs, S, dim[, m]),
// row-per-path), different data path.  The store_tile.cuh kernel spends 113 warp instructions per stored slot (the
// moments kernel: 40 per iteration) on five stage / commit / transpose / flush sequences and is issue-bound at 0.37 of
// the copy bandwidth (profiles/r02_ncu_merton_store.*).  Here a warp keeps one [32 paths][32 elements] tile per array
// in the layout TMA reads (128-byte rows, 128B swizzle), every lane collects four consecutive elements of its path in
// registers and writes them with one 16-byte shared store, and one lane hands a full tile to the engine with a single
// cp.async.bulk.tensor.2d.
//
// Gangs.  Arrays that receive the same number of elements per iteration at the same phase (paths, left_paths and
// jump_paths; for 1-D models also time_paths) fill their tiles in lock-step: they share ONE vector counter, one
// full-tile test and one wait, and a staging store is `tile address of the lane ^ (counter << 4)` (the 128B swizzle
// of chunk v of row q is (v ^ (q & 7)) << 4, and the counter only touches those three bits): LOP3 + STS.128.
//
// Buffering.  The iterations run in super-groups (whole Philox blocks producing whole 16-byte vectors in every array:
// 12 iterations for 1-D models) and a vector is staged as soon as its fourth element exists, so between the flush of
// a tile and the next store into it lie the instructions of four iterations: tiles are single-buffered (20 KB per
// warp; shared memory bounds the residency, 9 one-warp CTAs per SM) and the wait for the engine's read of a tile
// (cp.async.bulk.wait_group.read) is deferred to that next store.  The copies issued at one staging point share one
// bulk group, committed lazily by the first wait that needs it.  Double-buffered tiles (4-5 warps per SM) and
// 64-element tiles flushed as two boxes back to back measured 1.5-2.5x slower: the kernel lives on resident paths.
//
// Control flow.  A super-group whose 12 iterations all lie inside the allocation runs as straight-line code (both
// Philox blocks first, then the iterations; the only branches are the rarely taken full-tile / wait tests every four
// iterations), so the scheduler sees blocks of ~200 instructions; the guarded form (iterations past the allocation
// skipped) serves the last partial super-group, and once all paths of the warp are idle (t = T, no jump pending) the
// remaining super-groups only stage the constants -- like the reference's batch, the warp runs in lock-step over the
// whole allocation (:182).
//
// As in diffusion_tma.cuh the tensor maps declare the PITCH as the row length (a box cut by the tensor bound inside a
// row costs the engine several full boxes), so a row's last tile spills into its padding -- unless it is short (at
// most 16 elements: 134 slots = 4 x 32 + 6 for 100 nominal steps), in which case the lanes write those columns
// themselves as whole 32-byte sectors (tma_gang.cuh: 18 % less DRAM traffic and a fifth fewer boxes for the engine);
// boxes wholly outside the row are not issued and rows past the end of the call are clipped by the map.
// Merton full storage 2e6 x 100: 2.2 ms (store_tile.cuh kernel) -> 1.11 ms, 0.74 of the measured copy bandwidth;
// what bounds it now is in DESIGN.md section 6.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "jump.cuh"
#include "tma_gang.cuh"

namespace sdemc {

#ifndef SDEMC_JUMP_TMA_BLOCK
#define SDEMC_JUMP_TMA_BLOCK 32   // threads per CTA: one warp = 20 KB of tiles (+ queue), nine CTAs per SM
#endif
#ifndef SDEMC_JUMP_TMA_OOL
#define SDEMC_JUMP_TMA_OOL 1      // rare staging code (wait, flush) as out-of-line calls (tma_gang.cuh)
#endif
#ifndef SDEMC_JUMP_TMA_PAD
#define SDEMC_JUMP_TMA_PAD 0      // A/B builds only: unused shared memory per CTA (lowers the residency)
#endif
constexpr int kJumpTmaBlock = SDEMC_JUMP_TMA_BLOCK;

enum { SG_FAST = 0, SG_GUARDED = 1, SG_IDLE = 2 };

// FULL: all five arrays (solve()); otherwise only `paths` (low_storage, solvers.py:152-153)
template <class C, int JSRC, bool FULL>
__global__ void __launch_bounds__(kJumpTmaBlock)
    jump_store_tma_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys,
                          const DevInject inj, const DevOut out, const int qdepth,
                          const __grid_constant__ CUtensorMap map_paths, const __grid_constant__ CUtensorMap map_left,
                          const __grid_constant__ CUtensorMap map_jumps, const __grid_constant__ CUtensorMap map_times,
                          const __grid_constant__ CUtensorMap map_normals, const TmaRows rows_state,
                          const TmaRows rows_times, const TmaRows rows_normals, unsigned int* __restrict__ d_sched) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M, MARKS = C::MARKS;
  constexpr int NZ = BASE + (M == 2 ? 1 : 0);  // normals per iteration
  constexpr int SPB = steps_per_group(NZ);     // iterations served by one group of Philox blocks
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  constexpr int NPS = DIM * M;                 // increments recorded per iteration
  // super-group: whole Philox groups producing a multiple of four elements in every array (times: one per iteration)
  constexpr int SG = tma_lcm(tma_lcm(SPB, 4), tma_lcm(4 / tma_gcd(4, DIM), 4 / tma_gcd(4, NPS)));
  constexpr int NG = SG / SPB;                 // Philox groups per super-group
  constexpr int CP = DIM % 4;                  // state elements carried between super-groups (x0 comes first)
  constexpr bool INJECT = JSRC == JSRC_INJECT;
  // the state gang: paths [, left, jumps [, times when they share the phase]]; times and normals otherwise on their own
  constexpr bool TIMES_IN_STATE = FULL && DIM == 1;
  constexpr int NSTATE = FULL ? (TIMES_IN_STATE ? 4 : 3) : 1;
  using Src = typename std::conditional<JSRC == JSRC_INJECT, InjectJumps<MARKS>,
                                        typename std::conditional<JSRC == JSRC_QUEUE, QueueJumps<MARKS>,
                                                                  InlineJumps<MARKS>>::type>::type;
  if (JSRC == JSRC_QUEUE) {
    if (threadIdx.x == 0) {
      g_sh_sde = s;
      g_sh_keys = keys;
    }
    __syncthreads();
  }
  // tiles after the jump queue, on a 1024-byte boundary of the shared window (the launch adds 1 KB of slack)
  const uint32_t queue_bytes = JSRC == JSRC_QUEUE ? (uint32_t)qdepth * blockDim.x * (uint32_t)sizeof(float2) : 0u;
  const uint32_t warp_tiles = (((uint32_t)__cvta_generic_to_shared(jump_queue_smem) + queue_bytes + 1023u) & ~1023u) +
                              (threadIdx.x >> 5) * ((FULL ? 5u : 1u) * kTmaTileBytes);
  TmaGang<NSTATE, 1, SDEMC_JUMP_TMA_OOL != 0> g_state;
  TmaGang<1, 1, SDEMC_JUMP_TMA_OOL != 0> g_times, g_norm;
  g_state.init(warp_tiles, rows_state.dcol);
  g_state.set_array(0, &map_paths, rows_state);
  if (FULL) {
    g_state.set_array(1, &map_left, rows_state);
    g_state.set_array(2, &map_jumps, rows_state);
    if (TIMES_IN_STATE) {  // (the host enables direct columns only when the time rows agree with the state rows)
      g_state.set_array(3, &map_times, rows_times);
    } else {
      g_times.init(warp_tiles + 3u * kTmaTileBytes, rows_times.dcol);
      g_times.set_array(0, &map_times, rows_times);
    }
    g_norm.init(warp_tiles + 4u * kTmaTileBytes, rows_normals.dcol);
    g_norm.set_array(0, &map_normals, rows_normals);
  }
  TmaGroups grp;
  grp.init();

  int local_max_iters = 0;
  const int n = s.num_steps;
  const int S = out.S;
  TmaWarpTasks tasks;  // groups of 32 consecutive paths, handed out through the caller's workspace
  for (tasks.init(d_sched, rg.n_paths); tasks.valid(); tasks.advance()) {
    const uint64_t wbase = tasks.first_row();
    const uint64_t i = wbase + (threadIdx.x & 31);
    const bool valid = i < rg.n_paths;
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    JumpState st;
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) st.x[d] = d < DIM ? s.x0[d] : 0.0f;
    st.t = 0.0f;
    st.k = 0;
    st.need_pop = true;
    Src src;
    if constexpr (JSRC == JSRC_INJECT) src.init(s, inj, valid ? i : 0);
    else if constexpr (JSRC == JSRC_QUEUE) src.init(qdepth, plo, phi);
    else src.init(plo, phi);

    // the unit normals of SPB consecutive iterations starting at iteration b * SPB (as in jump_kernel)
    auto load_normals = [&](int b, float(&nrm)[NBUF], float(&extra)[SPB]) {
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
          const int k = b * SPB + sp;
          extra[sp] = 0.0f;
#pragma unroll
          for (int q = 0; q < NZ; ++q) nrm[sp * NZ + q] = 0.0f;
          if (k < inj.K && valid) {
            const float* zp = inj.z + (i * (uint64_t)inj.K + k) * DIM;
#pragma unroll
            for (int q = 0; q < BASE; ++q) nrm[sp * NZ + q] = zp[q];
            if (C::ASIAN) extra[sp] = zp[BASE];
            if (M == 2) nrm[sp * NZ + BASE] = inj.zc[i * (uint64_t)inj.K + k];
          }
        }
      }
    };

    g_state.begin_rows((int)wbase);
    if (FULL) {
      if (!TIMES_IN_STATE) g_times.begin_rows((int)wbase);
      g_norm.begin_rows((int)wbase);
    }
    // register row buffers of one super-group; the first CP (state) / 1 (times) entries are carried over
    float pb[CP + SG * DIM], lb[CP + SG * DIM], jb[CP + SG * DIM], tb[1 + SG], nb[SG * NPS];
    // vector v of the state rows / the time row / the increment row
    auto emit_state = [&](int v) {
      g_state.begin(grp);  // (first store into tiles the engine may still be reading)
      g_state.store(0, pb[4 * v], pb[4 * v + 1], pb[4 * v + 2], pb[4 * v + 3]);
      if (FULL) {
        g_state.store(1, lb[4 * v], lb[4 * v + 1], lb[4 * v + 2], lb[4 * v + 3]);
        g_state.store(2, jb[4 * v], jb[4 * v + 1], jb[4 * v + 2], jb[4 * v + 3]);
        if (TIMES_IN_STATE) g_state.store(3, tb[4 * v], tb[4 * v + 1], tb[4 * v + 2], tb[4 * v + 3]);
      }
      g_state.end(grp);
    };
    auto emit_times = [&](int v) {
      g_times.begin(grp);
      g_times.store(0, tb[4 * v], tb[4 * v + 1], tb[4 * v + 2], tb[4 * v + 3]);
      g_times.end(grp);
    };
    auto emit_norm = [&](int v) {
      g_norm.begin(grp);
      g_norm.store(0, nb[4 * v], nb[4 * v + 1], nb[4 * v + 2], nb[4 * v + 3]);
      g_norm.end(grp);
    };
    if (DIM == 4) {  // element 0 of the state rows: the initial value (:172-176)
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        pb[d] = lb[d] = st.x[d];
        jb[d] = 0.0f;
      }
      emit_state(0);
    }
#pragma unroll
    for (int d = 0; d < CP; ++d) {
      pb[d] = st.x[d];
      lb[d] = st.x[d];
      jb[d] = 0.0f;
    }
    tb[0] = 0.0f;

    float xs[kMaxDim];  // state at array index num_steps ('terminal' payoff index)
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
    int own_iters = 0;

    // one super-group of SG iterations starting at iteration g0
    auto super_group = [&](auto mode_c, int g0) {
      constexpr int MODE = decltype(mode_c)::value;
      float nrm[NG][NBUF], extra[NG][SPB];
      if (MODE == SG_FAST) {
#pragma unroll
        for (int gi = 0; gi < NG; ++gi) load_normals(g0 / SPB + gi, nrm[gi], extra[gi]);
      }
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        if (MODE == SG_GUARDED) load_normals(g0 / SPB + gi, nrm[gi], extra[gi]);
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) {
          const int ls = gi * SPB + sp;  // iteration inside the super-group
          StepRecord rec;
#pragma unroll
          for (int d = 0; d < kMaxDim; ++d) {
            rec.left[d] = st.x[d];
            rec.dw1[d] = 0.0f;
          }
          rec.dw2 = 0.0f;
          rec.Jc = 0.0f;
          rec.sq = 0.0f;
          float ex = 0.0f;
          // GUARDED: iterations past the allocation are skipped (warp-uniform), their columns are clipped or padding
          if (MODE == SG_FAST || (MODE == SG_GUARDED && g0 + ls < S)) {
            if (st.t < s.T) own_iters = st.k + 1;
            jump_iteration<C, Src, true>(s, keys, st, src, nrm[gi] + sp * NZ, rec);
            ex = extra[gi][sp];
            if (st.k == n) {
#pragma unroll
              for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
            }
          }
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            pb[CP + ls * DIM + d] = st.x[d];
            lb[CP + ls * DIM + d] = rec.left[d];
            jb[CP + ls * DIM + d] = rec.Jc;
            if (d < BASE) {
              nb[(ls * DIM + d) * M] = rec.dw1[d];
              if (M == 2) nb[(ls * DIM + d) * M + 1] = rec.dw2;
            } else {
              nb[(ls * DIM + d) * M] = ex * rec.sq;  // injected normal of the asian integral component
            }
          }
          tb[1 + ls] = st.t;
          // stage every vector whose fourth element now exists
#pragma unroll
          for (int v = (CP + ls * DIM) / 4; v < (CP + (ls + 1) * DIM) / 4; ++v) emit_state(v);
          if (FULL) {
            if (!TIMES_IN_STATE) {
#pragma unroll
              for (int v = (1 + ls) / 4; v < (2 + ls) / 4; ++v) emit_times(v);
            }
#pragma unroll
            for (int v = (ls * NPS) / 4; v < ((ls + 1) * NPS) / 4; ++v) emit_norm(v);
          }
        }
      }
#pragma unroll
      for (int d = 0; d < CP; ++d) {
        pb[d] = pb[SG * DIM + d];
        lb[d] = lb[SG * DIM + d];
        jb[d] = jb[SG * DIM + d];
      }
      tb[0] = tb[SG];
    };

    for (int g0 = 0; g0 < S; g0 += SG) {
      // Once every path of the warp has reached T and none has a jump pending at T, the remaining iterations are
      // idle (dt = 0: state, time and the zero increments repeat): stage them without running the loop body.
      const bool idle = !(st.t < s.T) && !st.need_pop && !(fabsf(src.tau - st.t) <= fmaf(fabsf(st.t), 1e-5f, 1e-12f));
      if (__all_sync(0xffffffffu, idle)) super_group(std::integral_constant<int, SG_IDLE>{}, g0);
      else if (g0 + SG <= S) super_group(std::integral_constant<int, SG_FAST>{}, g0);
      else super_group(std::integral_constant<int, SG_GUARDED>{}, g0);
    }
    // the last elements of the rows when S is a whole number of super-groups (else: padding / clipped columns)
    if (CP > 0 || TIMES_IN_STATE) {
#pragma unroll
      for (int d = CP; d < 4; ++d) pb[d] = lb[d] = jb[d] = 0.0f;
#pragma unroll
      for (int d = 1; d < 4; ++d) tb[d] = 0.0f;
      emit_state(0);
    }
    if (FULL && !TIMES_IN_STATE) {
#pragma unroll
      for (int d = 1; d < 4; ++d) tb[d] = 0.0f;
      emit_times(0);
    }
    if (g_state.direct()) {  // a short last tile: whole sectors written by the lanes themselves (tma_gang.cuh)
      g_state.write_tail(0, out.paths, out.pitch_state, valid);
      if (FULL) {
        g_state.write_tail(1, out.left, out.pitch_state, valid);
        g_state.write_tail(2, out.jumps, out.pitch_state, valid);
        if (TIMES_IN_STATE) g_state.write_tail(3, out.times, out.pitch_times, valid);
      }
    }
    g_state.finish(grp);
    if (FULL) {
      if (!TIMES_IN_STATE) {
        if (g_times.direct()) g_times.write_tail(0, out.times, out.pitch_times, valid);
        g_times.finish(grp);
      }
      if (g_norm.direct()) g_norm.write_tail(0, out.normals, out.pitch_normals, valid);
      g_norm.finish(grp);
    }

    float xp[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xp[d] = po.index_mode == SDEMC_INDEX_TERMINAL ? xs[d] : st.x[d];
    const float pay = eval_payoff<DIM>(po, xp);
    if (valid && out.payoffs) out.payoffs[i] = pay;
    if (valid && out.iters) out.iters[i] = own_iters;
    if (valid) local_max_iters = max(local_max_iters, own_iters);
    if (valid && out.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) out.terminal[i * DIM + d] = xp[d];
    }
  }
  tasks.finish();
  if (out.total_steps) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      local_max_iters = max(local_max_iters, __shfl_xor_sync(0xffffffffu, local_max_iters, off));
    if ((threadIdx.x & 31) == 0 && local_max_iters > 0) atomicMax(out.total_steps, local_max_iters);
  }
  grp.drain();  // all bulk stores of this warp complete before the CTA retires
}

}  // namespace sdemc
