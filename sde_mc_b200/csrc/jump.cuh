// jump.cuh -- fused jump-adapted Euler kernels: JumpDiffusionSolver.solve (solvers.py:164-226) + payoff + moments.
//
// Semantics kept from the reference loop:
//   h  = min(h, max(T - t, 0))            the mesh only shrinks, the grid restarts at every jump   :190
//   dt = min(h, tau - t)                  evaluated statelessly as min(h0, min(tau, T) - t)        :191
//   hit <=> |tau - t| <= 1e-12 + 1e-5 |t|   (torch.isclose with its default rtol)                :212,225
//   the jump acts on the pre-step state unless exact_jumps                                       :214-217
//   second Brownian driver of 'indep' models = ONE scalar normal shared by all components        :198-201
//   payoff at array index num_steps ('terminal') or at the last state ('adapted')                mc.py:84-91
// Deviation: dt is clamped at 0 where the reference would trip `assert next_jump_time >= t` (:193).
#pragma once
#include "engine.cuh"
#include "store_tile.cuh"

namespace sdemc {

enum { JSRC_INJECT = 0, JSRC_QUEUE = 1, JSRC_INLINE = 2 };

// Shared memory of the sparse-jump queue.  Always addressed as an offset from the symbol (plain LDS/STS with a
// register index); holding a generic pointer to it would make ptxas rebuild the shared-window address
// (S2R SR_CgaCtaId + LEA) on every loop iteration.
extern __shared__ float2 jump_queue_smem[];
__device__ __forceinline__ float2 lds_float2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
// Block-shared copies of the problem for the out-of-line refill (a __noinline__ function cannot see kernel params).
__shared__ DevSde g_sh_sde;
__shared__ PhiloxKeys g_sh_keys;

// ------------------------------------------------------------------------------------------------------------
// Jump sources: where (tau, J) of the next jump comes from.
//   begin_iter(k) : called at the top of loop iteration k
//   advance()     : move to the following jump (called once before the first iteration and after every hit)
//   mark(k)       : jump mark to apply on a hit in iteration k
// ------------------------------------------------------------------------------------------------------------

// Deterministic parity mode: sample_jump_times / sample_one_jump replaced by arrays (solvers.py:143-148).
template <int MARKS>
struct InjectJumps {
  const float* jt;
  const float* mk;
  int jidx, max_jumps, K;
  float tau;
  __device__ __forceinline__ void init(const DevSde& s, const DevInject& inj, uint64_t i) {
    jt = inj.jump_times + i * (uint64_t)s.max_jumps;
    mk = inj.marks + i * (uint64_t)inj.K;
    jidx = -1;
    max_jumps = s.max_jumps;
    K = inj.K;
    tau = 0.0f;
  }
  __device__ __forceinline__ void begin_iter(const DevSde&, const PhiloxKeys&, int) {}
  __device__ __forceinline__ void advance(const DevSde&, const PhiloxKeys&, bool pop) {
    if (pop) {
      ++jidx;
      tau = jidx < max_jumps ? jt[jidx] : __int_as_float(0x7f800000);
    }
  }
  __device__ __forceinline__ float mark(const DevSde& s, int k) const {
    return mark_from_raw<MARKS>(s, k < K ? mk[k] : 0.0f);
  }
};

// Cold path of the sparse-jump queue, deliberately out of line (one copy per kernel instead of one per unrolled
// iteration): draws `qd` more (tau, J) pairs into this thread's queue column and returns the new running jump time.
// Two Philox blocks give 4 jumps: 4 gap uniforms + 4 mark draws (2 Box-Muller pairs or 4 uniforms).
// PREMUL: queue c[0] * J instead of J (1-D moments kernel, jump1d.cuh).
// The draws of queue group `grp` (global index: chunk * groups + r), i.e. jumps 4 grp .. 4 grp + 3 of the path: four
// Exp(1) gaps from block 2 grp, four raw mark draws from block 2 grp + 1 (two Box-Muller pairs for lognormal marks,
// four uniforms for icdf marks).  Shared by queue_refill and the draw-reporting test hook (debug_draws.cuh).
template <int MARKS>
__device__ __forceinline__ void queue_group_draws(uint32_t grp, uint32_t plo, uint32_t phi, const PhiloxKeys& keys,
                                                  float (&gap)[4], float (&raw)[4]) {
  uint32_t g[4], m[4];
  philox4x32_10(grp * 2u, STREAM_JUMP_QUEUE, plo, phi, keys, g);
  philox4x32_10(grp * 2u + 1u, STREAM_JUMP_QUEUE, plo, phi, keys, m);
  if (MARKS == SDEMC_MARKS_LOGNORMAL) {
    box_muller(m[0], m[1], raw[0], raw[1]);
    box_muller(m[2], m[3], raw[2], raw[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) raw[j] = bits_to_u01(m[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) gap[j] = exp1_from_bits(g[j]);
}

template <int MARKS, bool PREMUL = false>
__device__ __noinline__ float queue_refill(int qd, uint32_t chunk, uint32_t plo, uint32_t phi, float tau_acc) {
  const DevSde& s = g_sh_sde;
  const PhiloxKeys& keys = g_sh_keys;
  const int groups = qd >> 2;
  for (int r = 0; r < groups; ++r) {
    float gap[4], raw[4];
    queue_group_draws<MARKS>(chunk * (uint32_t)groups + (uint32_t)r, plo, phi, keys, gap, raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tau_acc = fmaf(gap[j], s.inv_rate, tau_acc);
      const float mark = mark_from_raw<MARKS>(s, raw[j]);
      jump_queue_smem[(r * 4 + j) * blockDim.x + threadIdx.x] = make_float2(tau_acc, PREMUL ? s.c[0] * mark : mark);
    }
  }
  return tau_acc;
}

// Sparse jumps (rate * h << 1, e.g. Merton): a per-thread queue of QD pre-drawn (tau, J) pairs in shared
// memory, filled with full lane utilisation.  The hot loop is branch-free: qi += hit, one LDS.64 per iteration.
template <int MARKS>
struct QueueJumps {
  // this thread's column: slot j lives at jump_queue_smem[j * blockDim.x + threadIdx.x]; the hot loop reads it
  // through a 32-bit shared-window address kept in a register (base + qi * stride, one IMAD + one LDS.64)
  int qi, qd;
  uint32_t chunk, plo, phi, q_base, q_stride;
  float tau_acc, tau, J;
  __device__ __forceinline__ void init(int qdepth, uint32_t plo_, uint32_t phi_) {
    q_base = (uint32_t)__cvta_generic_to_shared(&jump_queue_smem[threadIdx.x]);
    q_stride = blockDim.x * (uint32_t)sizeof(float2);
    qd = qdepth;
    qi = qdepth - 1;  // the first advance(pop = true) triggers the initial fill
    chunk = 0;
    plo = plo_;
    phi = phi_;
    tau_acc = 0.0f;
    tau = 0.0f;
    J = 0.0f;
  }
  __device__ __forceinline__ void begin_iter(const DevSde&, const PhiloxKeys&, int) {}
  __device__ __forceinline__ void advance(const DevSde&, const PhiloxKeys&, bool pop) {
    qi += pop ? 1 : 0;
    if (qi == qd) {
      tau_acc = queue_refill<MARKS>(qd, chunk, plo, phi, tau_acc);
      ++chunk;
      qi = 0;
    }
    const float2 e = lds_float2(q_base + (uint32_t)qi * q_stride);
    tau = e.x;
    J = e.y;
  }
  __device__ __forceinline__ float mark(const DevSde&, int) const { return J; }
};

// Dense jumps (rate * h ~ 1, e.g. the Levy models): every iteration draws a fresh (gap, mark) candidate in
// registers and a hit consumes it -- branch-free, no divergence, no queue.
template <int MARKS>
struct InlineJumps {
  uint32_t plo, phi;
  float tau, J, cand_gap, cand_raw;
  float next_gap, next_raw;  // second (gap, mark) candidate of the current Philox block, for the odd iteration
  __device__ __forceinline__ void init(uint32_t plo_, uint32_t phi_) {
    plo = plo_;
    phi = phi_;
    tau = 0.0f;
    J = 0.0f;
    next_gap = 0.0f;
    next_raw = 0.0f;
  }
  // one Philox block serves two consecutive iterations: (gap, mark) from words (0, 1) and (2, 3) for uniform
  // marks; two gaps and one Box-Muller pair for lognormal marks.  k is the same for all active lanes of a warp.
  // the (gap, raw mark) candidates of iterations 2b and 2b + 1 from block b of STREAM_JUMP_INLINE
  static __device__ __forceinline__ void block_draws(uint32_t b, uint32_t plo, uint32_t phi, const PhiloxKeys& keys,
                                                     float& gap0, float& raw0, float& gap1, float& raw1) {
    uint32_t o[4];
    philox4x32_10(b, STREAM_JUMP_INLINE, plo, phi, keys, o);
    gap0 = exp1_from_bits(o[0]);
    if (MARKS == SDEMC_MARKS_LOGNORMAL) {
      gap1 = exp1_from_bits(o[1]);
      box_muller(o[2], o[3], raw0, raw1);
    } else {
      raw0 = bits_to_u01(o[1]);
      gap1 = exp1_from_bits(o[2]);
      raw1 = bits_to_u01(o[3]);
    }
  }
  __device__ __forceinline__ void begin_iter(const DevSde&, const PhiloxKeys& keys, int k) {
    if ((k & 1) == 0) {
      block_draws((uint32_t)(k >> 1), plo, phi, keys, cand_gap, cand_raw, next_gap, next_raw);
    } else {
      cand_gap = next_gap;
      cand_raw = next_raw;
    }
  }
  __device__ __forceinline__ void advance(const DevSde& s, const PhiloxKeys&, bool pop) {
    const float t2 = fmaf(cand_gap, s.inv_rate, tau);
    const float j2 = mark_from_raw<MARKS>(s, cand_raw);
    tau = pop ? t2 : tau;
    J = pop ? j2 : J;
  }
  __device__ __forceinline__ float mark(const DevSde&, int) const { return J; }
};

// Sparse jumps without a queue (fused control-variate kernel, whose shared memory is taken by operand tiles): the
// next (gap, mark) pair is drawn only when the previous jump has been consumed -- a divergent branch, but with
// rate * h << 1 most warp iterations skip it.  Counter = index of the jump, so the stream is independent of the grid.
template <int MARKS>
struct LazyJumps {
  uint32_t plo, phi, njumps;
  float tau, J;
  __device__ __forceinline__ void init(uint32_t plo_, uint32_t phi_) {
    plo = plo_;
    phi = phi_;
    njumps = 0;
    tau = 0.0f;
    J = 0.0f;
  }
  __device__ __forceinline__ void begin_iter(const DevSde&, const PhiloxKeys&, int) {}
  __device__ __forceinline__ void advance(const DevSde& s, const PhiloxKeys& keys, bool pop) {
    if (pop) {
      uint32_t o[4];
      philox4x32_10(njumps++, STREAM_JUMP_INLINE, plo, phi, keys, o);
      tau = fmaf(exp1_from_bits(o[0]), s.inv_rate, tau);
      float raw, unused;
      if (MARKS == SDEMC_MARKS_LOGNORMAL) box_muller(o[1], o[2], raw, unused);
      else raw = bits_to_u01(o[1]);
      J = mark_from_raw<MARKS>(s, raw);
    }
  }
  __device__ __forceinline__ float mark(const DevSde&, int) const { return J; }
};

// ------------------------------------------------------------------------------------------------------------
// per-path state and one loop iteration
// ------------------------------------------------------------------------------------------------------------
struct JumpState {
  float x[kMaxDim];
  float t;
  int k;
  bool need_pop;
};

// what one iteration produced, for the path-storing mode (solvers.py:205-222)
struct StepRecord {
  float left[kMaxDim];  // state after the diffusion step, before the jump
  float dw1[kMaxDim];   // first-driver increments actually used
  float dw2;            // common second-driver increment
  float Jc;             // applied jump mark (0 if none)
  float sq;             // sqrt(dt) of the iteration
};

// one iteration of the while-loop solvers.py:182-225.  zn: this iteration's unit normals
// (BASE correlated-driver normals, then the common second-driver normal if M == 2).
template <class C, class Src, bool STORE>
__device__ __forceinline__ void jump_iteration(const DevSde& s, const PhiloxKeys& keys, JumpState& st, Src& src,
                                               const float* zn, StepRecord& rec) {
  constexpr int BASE = C::BASE, M = C::M;
  src.begin_iter(s, keys, st.k);
  src.advance(s, keys, st.need_pop);
  const float tau = src.tau;
  // h = min(h, max(T - t, 0)) (:190) only ever equals min(h0, max(T - t, 0)) because t never decreases, and fp32
  // subtraction of the same t is monotone: dt = min(h, tau - t) = min(h0, min(tau, T) - t), clamped at 0 (:191-193)
  const float dt = fmaxf(fminf(s.h0, fminf(tau, s.T) - st.t), 0.0f);
  const float sq = fast_sqrt(dt);
  float z1[kMaxDim], w1[kMaxDim], w2[kMaxDim], xo[kMaxDim];
#pragma unroll
  for (int i = 0; i < BASE; ++i) z1[i] = zn[i];
  correlate<C>(s, z1, w1);
#pragma unroll
  for (int i = 0; i < BASE; ++i) w2[i] = M == 2 ? zn[BASE] : 0.0f;
#pragma unroll
  for (int i = 0; i < kMaxDim; ++i) xo[i] = st.x[i];
  euler_step<C>(s, st.x, dt, sq, w1, w2, st.t);  // user coefficients see t before the step (:204)
  st.t += dt;
  const bool hit = fabsf(tau - st.t) <= fmaf(fabsf(st.t), 1e-5f, 1e-12f);
  // branch-free: a zero mark leaves the state untouched (x + c x_base * 0)
  const float Jc = hit ? src.mark(s, st.k) : 0.0f;
  if (STORE) {
#pragma unroll
    for (int i = 0; i < kMaxDim; ++i) {
      rec.left[i] = st.x[i];
      rec.dw1[i] = i < BASE ? w1[i] * sq : 0.0f;
    }
    rec.dw2 = M == 2 ? zn[BASE] * sq : 0.0f;
    rec.Jc = Jc;
    rec.sq = sq;
  }
  if (s.exact_jumps) {
#pragma unroll
    for (int i = 0; i < kMaxDim; ++i) xo[i] = st.x[i];
  }
  add_jump<C>(s, st.x, xo, Jc, st.t);            // ... and the jump time after it (:214-217)
  st.need_pop = hit;
  ++st.k;
}

// ------------------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------------------
// staging tiles of the path-storing mode: 16 elements per path and flush, five arrays; a group of iterations
// stages up to steps_per_group * dim * m elements per array
#ifndef SDEMC_JUMP_STORE_TILE
#define SDEMC_JUMP_STORE_TILE 16
#endif
#ifndef SDEMC_JUMP_STORE_MINB
#define SDEMC_JUMP_STORE_MINB 1
#endif
template <class C>
using JumpStoreWriter = WarpTileWriter<SDEMC_JUMP_STORE_TILE, steps_per_group(C::BASE + (C::M == 2 ? 1 : 0)) * C::DIM * C::M>;
constexpr int kJumpStoreBlock = 128;  // threads per CTA of the storing kernels (their shared tiles limit residency)

#ifndef SDEMC_JUMP_MIN_BLOCKS
#define SDEMC_JUMP_MIN_BLOCKS 4  // <= 64 registers per thread: 32 resident warps per SM
#endif
// Resident CTAs the register budget is sized for.  The dense multi-dimensional models (Levy 2-D: ~25 live parameters,
// three normals and a four-branch inverse cdf per iteration) need ~110 registers; capped at 64 ptxas re-reads its
// parameters with ~15 LDC per iteration (ADU pipe 71 % busy, profiles/r01_ncu_levy2d.*).  Measured on Levy 2-D:
// 8.2e10 path-steps/s at 4 or 3 CTAs per SM, 8.9e10 at 2.
template <class C, int JSRC, bool STORE>
constexpr int jump_min_blocks() {
  return STORE ? SDEMC_JUMP_STORE_MINB : ((JSRC == JSRC_INLINE && C::DIM >= 2) ? 2 : SDEMC_JUMP_MIN_BLOCKS);
}
template <class C, int JSRC, bool STORE>
__global__ void __launch_bounds__(256, jump_min_blocks<C, JSRC, STORE>()) jump_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                   const PhiloxKeys keys, const DevInject inj, const DevOut out,
                                                   const int qdepth, double* __restrict__ d_moments,
                                                   void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M, MARKS = C::MARKS;
  constexpr int NZ = BASE + (M == 2 ? 1 : 0);  // normals per iteration
  constexpr int SPB = steps_per_group(NZ);     // iterations served by one group of Philox blocks (6 normals each)
  constexpr int BPS = blocks_per_group(NZ);
  constexpr int NBUF = BPS * kNormalsPerBlock;
  constexpr bool INJECT = JSRC == JSRC_INJECT;
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

  using Writer = JumpStoreWriter<C>;
  Writer w_paths, w_left, w_jumps, w_times, w_norm;  // STORE: five staging tiles per warp, after the jump queue
  float* store_tiles = reinterpret_cast<float*>(jump_queue_smem + (JSRC == JSRC_QUEUE ? qdepth * blockDim.x : 0)) +
                       (threadIdx.x >> 5) * (5 * Writer::kFloats);

  Accum acc;
  acc.zero();
  int local_max_iters = 0;
  const int n = s.num_steps;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // warp-uniform trip count (see diffusion.cuh): STORE stages its outputs with all 32 lanes in lock-step
  for (uint64_t wbase = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wbase < range_n(rg); wbase += stride) {
    const uint64_t i = wbase + (threadIdx.x & 31);
    if (!STORE && i >= range_n(rg)) break;
    const bool valid = i < range_n(rg);
    const uint64_t gp = range_lo(rg) + i;
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

    // fetch the unit normals of SPB consecutive iterations starting at iteration b * SPB
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

    float xs[kMaxDim];  // state at array index num_steps ('terminal' payoff index)
    int own_iters = 0;
    if (STORE) {
      // lock-step over the whole allocation: finished paths idle with dt = 0 exactly as in the reference,
      // where the loop runs until the slowest path of the batch is done (:182).
      w_paths.init(store_tiles + 0 * Writer::kFloats, out.paths, out.pitch_state, wbase, range_n(rg));
      w_left.init(store_tiles + 1 * Writer::kFloats, out.left, out.pitch_state, wbase, range_n(rg));
      w_jumps.init(store_tiles + 2 * Writer::kFloats, out.jumps, out.pitch_state, wbase, range_n(rg));
      w_times.init(store_tiles + 3 * Writer::kFloats, out.times, out.pitch_times, wbase, range_n(rg));
      w_norm.init(store_tiles + 4 * Writer::kFloats, out.normals, out.pitch_normals, wbase, range_n(rg));
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        w_paths.append(st.x[d]);
        w_left.append(st.x[d]);
        w_jumps.append(0.0f);
      }
      w_times.append(0.0f);
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
      for (int b = 0; b * SPB < out.S; ++b) {
        // Once every path of the warp has reached T and none has a jump pending at T, the remaining iterations are
        // idle (dt = 0: state, time and the zero increments repeat); write them without running the loop body.
        const bool idle = !(st.t < s.T) && !st.need_pop && !(fabsf(src.tau - st.t) <= fmaf(fabsf(st.t), 1e-5f, 1e-12f));
        if (__all_sync(0xffffffffu, idle)) break;
        float nrm[NBUF], extra[SPB];
        load_normals(b, nrm, extra);
        const int done = min(SPB, out.S - st.k);  // iterations of this group inside the allocation (warp-uniform)
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) {
          if (sp < done) {
            if (st.t < s.T) own_iters = st.k + 1;
            StepRecord rec;
            jump_iteration<C, Src, true>(s, keys, st, src, nrm + sp * NZ, rec);
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
              w_left.stage(sp * DIM + d, rec.left[d]);
              w_paths.stage(sp * DIM + d, st.x[d]);
              w_jumps.stage(sp * DIM + d, rec.Jc);
              if (d < BASE) {
                w_norm.stage((sp * DIM + d) * M, rec.dw1[d]);
                if (M == 2) w_norm.stage((sp * DIM + d) * M + 1, rec.dw2);
              } else {
                w_norm.stage((sp * DIM + d) * M, extra[sp] * rec.sq);  // injected normal of the asian integral component
              }
            }
            w_times.stage(sp, st.t);
            if (st.k == n) {
#pragma unroll
              for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
            }
          }
        }
        w_left.commit(done * DIM);
        w_paths.commit(done * DIM);
        w_jumps.commit(done * DIM);
        w_norm.commit(done * DIM * M);
        w_times.commit(done);
      }
      for (; st.k < out.S; ++st.k) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          w_left.append(st.x[d]);
          w_paths.append(st.x[d]);
          w_jumps.append(0.0f);
#pragma unroll
          for (int q = 0; q < M; ++q) w_norm.append(0.0f);
        }
        w_times.append(st.t);
      }
      w_paths.flush();
      w_left.flush();
      w_jumps.flush();
      w_times.flush();
      w_norm.flush();
    } else {
      // phase 1: the first num_steps iterations can never reach T (each advances by at most T/num_steps),
      // so full Philox blocks run without the loop-exit test.
      StepRecord rec_unused;
      const int nb_full = n / SPB;
      for (int b = 0; b < nb_full; ++b) {
        float nrm[NBUF], extra[SPB];
        load_normals(b, nrm, extra);
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) jump_iteration<C, Src, false>(s, keys, st, src, nrm + sp * NZ, rec_unused);
      }
      if (st.k == n) {
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
      }
      // phase 2: the few remaining iterations (those forced by jumps) with the exit test; same block structure, so
      // the Philox stream stays identical to STORE mode.
      const int kcap = INJECT ? inj.K : 4 * (n + s.max_jumps) + 64;
      bool done = !(st.t < s.T) || st.k >= kcap;
      for (int b = nb_full; !done; ++b) {
        float nrm[NBUF], extra[SPB];
        load_normals(b, nrm, extra);
#pragma unroll
        for (int sp = 0; sp < SPB; ++sp) {
          if (!done) {
            jump_iteration<C, Src, false>(s, keys, st, src, nrm + sp * NZ, rec_unused);
            if (st.k == n) {
#pragma unroll
              for (int d = 0; d < kMaxDim; ++d) xs[d] = st.x[d];
            }
            done = !(st.t < s.T) || st.k >= kcap;
          }
        }
      }
      own_iters = st.k;
    }

    float xp[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xp[d] = po.index_mode == SDEMC_INDEX_TERMINAL ? xs[d] : st.x[d];
    const float pay = eval_payoff<DIM>(po, xp);
    if (STORE) {
      if (valid && out.payoffs) out.payoffs[i] = pay;
      if (valid && out.iters) out.iters[i] = own_iters;
      if (valid) local_max_iters = max(local_max_iters, own_iters);
      if (valid && out.terminal) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) out.terminal[i * DIM + d] = xp[d];
      }
    } else {
      acc.add(pay, po.df * st.x[0] - s.x0[0], own_iters);
      write_per_path<DIM>(per_path_of_out(out), i, pay, own_iters, xp);
    }
  }
  if (STORE) {
    if (out.total_steps) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        local_max_iters = max(local_max_iters, __shfl_xor_sync(0xffffffffu, local_max_iters, off));
      if ((threadIdx.x & 31) == 0 && local_max_iters > 0) atomicMax(out.total_steps, local_max_iters);
    }
  } else {
    block_reduce_and_publish(acc, d_moments, d_ws);
  }
}

}  // namespace sdemc
