// jump1d.cuh -- moments-only fast path of the jump-adapted Euler loop for ONE-dimensional geometric jump diffusions
// with sparse lognormal jumps (Merton, the north-star workload): JumpDiffusionSolver.solve solvers.py:164-226 with
// low_storage semantics + payoff + (sum, sum^2).  Same arithmetic rules as jump.cuh (Q1-Q4 of SURVEY.md), restated
// so that one loop iteration is 14 instructions on top of its normal:
//
//   * the mesh is stateless.  The reference keeps h = min(h, max(T - t, 0)) (:190); t never decreases, so
//     h_k = min(h0, max(T - t_k, 0)) and  dt = min(h, tau - t) = max(min(h0, min(tau, T) - t), 0)  (fp32 subtraction
//     of the same t is monotone, so min(T - t, tau - t) == min(T, tau) - t bit for bit; the outer max is the dt >= 0
//     clamp that replaces the reference's assert :193).
//   * for h0 <= 1 the clamp is the saturate modifier of the subtraction: min(h0, sat(m - t)) == max(min(h0, m - t), 0)
//     (FADD.SAT clamps to [0, 1], and values in (h0, 1] lose against h0 anyway).
//   * while t <= T - h0 -- i.e. in all but the last group or two of phase 1 -- min(tau, T) - t and tau - t give the
//     same dt (both exceed h0 when tau >= T), so the cap at T is dropped there (bit-identical, one FMNMX less).
//   * sigma^2 and dt are folded into the Box-Muller radius (one square root per iteration instead of sqrt(dt) plus
//     the radius root), the jump coefficient into the queued mark (c J).
//   * the hit test |tau - t| <= 1e-12 + 1e-5 |t| is evaluated bit for bit as in jump.cuh, so iteration counts and
//     hit iterations are those of the path-storing kernel (and of the reference on the same draws); a hit applies
//     the jump with ONE predicated FFMA (x + base * 0 == x exactly, so skipping it is the same arithmetic).
//   * the queue of pre-drawn (tau, c J) pairs is popped branch-free by bumping a shared-memory address on a hit.
//     Whether a path ran out of queued jumps is checked once per group of 6 iterations (one Philox block of
//     normals): the group runs speculatively from a saved (x, t, queue head); if the head left the filled part of
//     the queue the group is replayed from the saved state with the per-iteration refill test (rare: a path needs
//     more than `qdepth` jumps).  The Philox counters of normals and jumps are those of jump.cuh, so this kernel
//     and the path-storing kernel simulate identical paths for the same seed.
//   * phase 2 -- the num_steps % 6 last nominal iterations plus those forced by jumps -- runs in PAIRS (the two
//     iterations that share one Box-Muller pair) with the `t < T` test between pairs: a warp stops within two
//     iterations of its slowest lane instead of six (round 1: 2.3 masked groups of six per warp, 16 % of all
//     instructions of the kernel).
#pragma once
#include "jump.cuh"

namespace sdemc {

constexpr int kQueueSlack = kNormalsPerBlock;  // slots a speculative group may read past the filled queue

#ifndef SDEMC_JUMP1D_MIN_BLOCKS
#define SDEMC_JUMP1D_MIN_BLOCKS 3  // 80 registers, no spills in the step loop: measured 5% faster than 4 CTAs with spills
#endif
// EXACT: exact_jumps (:214-217).  TERM: the payoff reads the state at array index num_steps ('terminal', quirk Q1)
// instead of the last state.  SAT: h0 <= 1, the dt >= 0 clamp rides on the subtraction (see above).
template <class C, bool EXACT, bool TERM, bool SAT>
__global__ void __launch_bounds__(256, SDEMC_JUMP1D_MIN_BLOCKS)
    jump1d_kernel(const DevSde s, const DevPayoff po, const DevRange rg, const PhiloxKeys keys, const int qdepth,
                  const int nb_uncapped, const DevPerPath pp, double* __restrict__ d_moments, void* __restrict__ d_ws) {
  static_assert(C::DIM == 1 && C::M == 1 && !C::ASIAN, "1-D single-driver models only");
  constexpr int MARKS = C::MARKS;
  constexpr int G = kNormalsPerBlock;  // iterations per group
  constexpr bool GEO = C::FAMILY == SDEMC_FAMILY_GEOMETRIC;
  if (threadIdx.x == 0) {
    g_sh_sde = s;
    g_sh_keys = keys;
  }
  __syncthreads();

  const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(&jump_queue_smem[threadIdx.x]);
  constexpr uint32_t q_stride = kBlock * (uint32_t)sizeof(float2);  // launched with kBlock threads (launch_jump.cu)
  const uint32_t q_end = q_base + (uint32_t)qdepth * q_stride;
  const float T = s.T, h0 = s.h0, a = s.a[0];
  const float neg2ln2_b2 = -1.3862943611198906f * s.b1[0] * s.b1[0];
  const int n = s.num_steps;
  const int kcap = 4 * (n + s.max_jumps) + 64;

  // fp64 running sums live in shared memory (one column per thread, touched once per path): 14 registers less in
  // the step loop than a register-resident Accum
  __shared__ double acc_sh[kNumMoments - 1][kBlock];
#pragma unroll
  for (int m = 0; m < kNumMoments - 1; ++m) acc_sh[m][threadIdx.x] = 0.0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range_n(rg); i += stride) {
    const uint64_t gp = range_lo(rg) + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float x = s.x0[0], t = 0.0f;
    uint32_t chunk = 1;
    float tau_acc = queue_refill<MARKS, true>(qdepth, 0u, plo, phi, 0.0f);  // initial fill
    uint32_t q = q_base;

    // one loop iteration.  CHECKED: refill test before the read (replay path, phase 2); CAPPED: dt is capped at T
    // (needed once t can come within h0 of T); returns whether the jump was hit.
    auto iteration = [&](float r2, float cs, auto checked, auto capped) {
      if (decltype(checked)::value) {
        if (q >= q_end) {
          tau_acc = queue_refill<MARKS, true>(qdepth, chunk, plo, phi, tau_acc);
          ++chunk;
          q = q_base;
        }
      }
      const float2 e = lds_float2(q);  // (tau, c J)
      const float m = decltype(capped)::value ? fminf(e.x, T) : e.x;
      const float dt = SAT ? fminf(h0, __saturatef(m - t)) : fmaxf(fminf(h0, m - t), 0.0f);
      // sigma z sqrt(dt) = sqrt(sigma^2 (-2 ln u) dt) * (cos | sin): the Box-Muller root and sqrt(dt) are one MUFU
      const float g = fmaf(fast_sqrt(r2 * dt), cs, a * dt);
      float xn = GEO ? fmaf(x, g, x) : x + g;
      t += dt;
      // torch.isclose(tau, t, atol=1e-12) with its default rtol=1e-5 (:212,225), evaluated exactly as jump.cuh and
      // the reference do (t >= 0).  The one-FMA form t (1 + 1e-5) + 1e-12 >= tau used here in round 1 rounds its
      // threshold differently: about one path in 1e6 then hit a jump one iteration early -- or, for a jump 1e-5 T
      // after T, hit a jump the reference never applies (found by tests/test_gpu_fastpath.py).
      const bool hit = fabsf(e.x - t) <= fmaf(t, 1e-5f, 1e-12f);
      if (hit) {  // one predicated FFMA / FADD and the queue pop
        xn = GEO ? fmaf(EXACT ? xn : x, e.y, xn) : xn + e.y;
        q += q_stride;
      }
      x = xn;
    };
    // one Philox block -> squared radii (sigma^2 folded in) and directions of 6 normals; normal 2j uses
    // (r2[j], cos), normal 2j+1 uses (r2[j], sin)
    auto group_normals = [&](int b, float(&r2)[3], float(&cs)[G]) {
      uint32_t o[4];
      philox4x32_10((uint32_t)b, STREAM_DIFFUSION, plo, phi, keys, o);
      float c[3], sn[3];
      philox_polar3_sq(o, neg2ln2_b2, r2, c, sn);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        cs[2 * j] = c[j];
        cs[2 * j + 1] = sn[j];
      }
    };
    // a full group of phase 1: speculative on the queue, replayed with the refill test if the head ran past it
    auto phase1_group = [&](int b, auto capped) {
      float r2[3], cs[G];
      group_normals(b, r2, cs);
      const float xs = x, ts = t;
      const uint32_t qs = q;
#pragma unroll
      for (int sp = 0; sp < G; ++sp) iteration(r2[sp / 2], cs[sp], std::false_type(), capped);
      if (q >= q_end) {  // ran out of queued jumps inside the group: replay it with the refill test
        x = xs;
        t = ts;
        q = qs;
        group_normals(b, r2, cs);  // recomputed rather than kept live across the speculative group
#pragma unroll
        for (int sp = 0; sp < G; ++sp) iteration(r2[sp / 2], cs[sp], std::true_type(), std::true_type());
        if (q >= q_end) {  // the next group must start on a valid head
          tau_acc = queue_refill<MARKS, true>(qdepth, chunk, plo, phi, tau_acc);
          ++chunk;
          q = q_base;
        }
      }
    };

    // phase 1: the first num_steps iterations can never reach T (each advances by at most T / num_steps): whole
    // groups without the exit test; the first nb_uncapped of them stay h0 away from T and skip the cap as well
    const int nb_full = n / G;
    int b = 0;
    for (; b < nb_uncapped; ++b) phase1_group(b, std::false_type());
    for (; b < nb_full; ++b) phase1_group(b, std::true_type());
    int k = nb_full * G;
    float x_at_n = x;
    // phase 2: the remaining num_steps % 6 iterations and those forced by jumps, until t reaches T -- in pairs
    while (t < T && k < kcap) {
      float r2[3], cs[G];
      group_normals(b++, r2, cs);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (t < T && k < kcap) {
          iteration(r2[j], cs[2 * j], std::true_type(), std::true_type());
          ++k;
          if (TERM && k == n) x_at_n = x;
          if (t < T) {
            iteration(r2[j], cs[2 * j + 1], std::true_type(), std::true_type());
            ++k;
            if (TERM && k == n) x_at_n = x;
          }
        }
      }
    }

    float xp[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xp[d] = 0.0f;
    xp[0] = TERM ? x_at_n : x;
    const float pay = eval_payoff<1>(po, xp);
    write_per_path<1>(pp, i, pay, k, xp);
    {
      Accum one;  // this path's contribution, folded into the thread's shared column
      one.zero();
      one.add(pay, po.df * x - s.x0[0], k);
#pragma unroll
      for (int m = 0; m < kNumMoments - 1; ++m) acc_sh[m][threadIdx.x] += one.v[m];
    }
  }
  Accum acc;
  acc.zero();
#pragma unroll
  for (int m = 0; m < kNumMoments - 1; ++m) acc.v[m] = acc_sh[m][threadIdx.x];
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
