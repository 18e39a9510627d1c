// pair.cuh -- fused MLMC level kernels: a thread carries the FINE and the COARSE path of one coupled pair,
// sharing Brownian increments and jumps in registers, and only D(T) (P(fine) - P(coarse)) is reduced.
//   jump-adapted pair : JumpDiffusionSolver.multilevel_solve solvers.py:228-307
//   uniform-grid pair : DiffusionSolver.multilevel_solve     solvers.py:90-119
// Work per level is tiny next to the plain MC kernels (C5 totals ~1e9 fine steps), so these kernels favour
// simplicity: Brownian normals from a six-per-block shift register, inline jump candidates per outer iteration.
// Deviation from the reference: dt is clamped at 0 where its fp32 run asserts (solvers.py:264), which is what
// makes the fp32 pair usable at all.
#pragma once
#include "engine.cuh"
#include "jump.cuh"

namespace sdemc {

// Philox normals of the Brownian stream, consumed NZ at a time in loop order: one block of six normals serves
// 6 / NZ consecutive sub-steps (a shift register, so no dynamically indexed registers).
template <int NZ>
struct NormalStream {
  static constexpr int PER = NZ <= 3 ? kNormalsPerBlock / NZ : 1;          // draws served by one refill
  static constexpr int BLOCKS = NZ <= 3 ? 1 : (NZ + kNormalsPerBlock - 1) / kNormalsPerBlock;
  float buf[BLOCKS * kNormalsPerBlock];
  uint32_t blk, plo, phi;
  int left;
  __device__ __forceinline__ void init(uint32_t plo_, uint32_t phi_) {
    plo = plo_;
    phi = phi_;
    blk = 0;
    left = 0;
  }
  __device__ __forceinline__ void next(const PhiloxKeys& keys, float (&zn)[NZ]) {
    if (left == 0) {
#pragma unroll
      for (int r = 0; r < BLOCKS; ++r) {
        uint32_t o[4];
        philox4x32_10(blk++, STREAM_DIFFUSION, plo, phi, keys, o);
        philox_normals6(o, buf + kNormalsPerBlock * r);
      }
      left = PER;
    }
#pragma unroll
    for (int e = 0; e < NZ; ++e) zn[e] = buf[e];
    if (PER > 1) {
#pragma unroll
      for (int j = 0; j + NZ < kNormalsPerBlock; ++j) buf[j] = buf[j + NZ];
    }
    --left;
  }
};

struct DevPairOut {
  float* terminal;  // (n, 2, dim): fine, coarse terminal states (parity mode) or nullptr
};

// state of one coupled (fine, coarse) pair of jump-adapted paths
template <class C, bool INJECT>
struct PairPath {
  static constexpr int NZ = C::BASE + (C::M == 2 ? 1 : 0);
  float xf[kMaxDim], xc[kMaxDim];
  float tf, tc;
  int k;
  bool need_pop;
  typename std::conditional<INJECT, InjectJumps<C::MARKS>, InlineJumps<C::MARKS>>::type src;
  NormalStream<NZ> normals;

  __device__ __forceinline__ void start(const DevSde& s, const DevInject& inj, uint64_t i, uint32_t plo, uint32_t phi) {
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xf[d] = xc[d] = d < C::DIM ? s.x0[d] : 0.0f;
    tf = 0.0f;
    tc = 0.0f;
    k = 0;
    need_pop = true;
    if constexpr (INJECT) src.init(s, inj, i);
    else src.init(plo, phi);
    normals.init(plo, phi);
  }
};

// one outer iteration of the coupled loop solvers.py:254-305: `factor` fine sub-steps, one coarse step on the summed
// increments, the shared jump
template <class C, bool INJECT>
__device__ __forceinline__ void pair_iteration(const DevSde& s, const PhiloxKeys& keys, const DevInject& inj,
                                               PairPath<C, INJECT>& p, uint64_t i, int factor, float hf0, float hc0) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M;
  constexpr int NZ = PairPath<C, INJECT>::NZ;
  float xof[kMaxDim], xoc[kMaxDim];
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) xof[d] = p.xf[d];
  p.src.begin_iter(s, keys, p.k);
  p.src.advance(s, keys, p.need_pop);
  const float tau = p.src.tau;
  float s1[kMaxDim], s2[kMaxDim];
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) s1[d] = s2[d] = 0.0f;
  for (int q = 0; q < factor; ++q) {                         // :259-278
    const int sub = p.k * factor + q;
    float zn[NZ];
    if constexpr (!INJECT) {
      p.normals.next(keys, zn);
    } else {
      const uint64_t zi = i * (uint64_t)inj.K * factor + sub;
#pragma unroll
      for (int d = 0; d < BASE; ++d) zn[d] = inj.z[zi * DIM + d];
      if (M == 2) zn[BASE] = inj.zc[zi];
    }
    const float dt = fmaxf(fminf(hf0, fminf(tau, s.T) - p.tf), 0.0f);  // stateless mesh, see jump.cuh
    const float sq = fast_sqrt(dt);
    float z1[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
    for (int d = 0; d < BASE; ++d) z1[d] = zn[d];
    correlate<C>(s, z1, w1);
#pragma unroll
    for (int d = 0; d < BASE; ++d) w2[d] = M == 2 ? zn[BASE] : 0.0f;
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xof[d] = p.xf[d];      // state before the LAST fine sub-step (:275)
    euler_step<C>(s, p.xf, dt, sq, w1, w2);
    p.tf += dt;
#pragma unroll
    for (int d = 0; d < BASE; ++d) {
      s1[d] = fmaf(w1[d], sq, s1[d]);
      if (M == 2) s2[d] = fmaf(w2[d], sq, s2[d]);
    }
  }
  const float dtc = fmaxf(fminf(hc0, fminf(tau, s.T) - p.tc), 0.0f);   // :282-286
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) xoc[d] = p.xc[d];
  euler_step<C>(s, p.xc, dtc, 1.0f, s1, s2);
  p.tc += dtc;
  const bool hit = fabsf(tau - p.tf) <= fmaf(fabsf(p.tf), 1e-5f, 1e-12f);   // :291 (on the fine clock)
  const float Jc = hit ? p.src.mark(s, p.k) : 0.0f;
  if (s.exact_jumps) {
    add_jump<C>(s, p.xf, p.xf, Jc);
    add_jump<C>(s, p.xc, p.xc, Jc);
  } else {
    add_jump<C>(s, p.xf, xof, Jc);
    add_jump<C>(s, p.xc, xoc, Jc);
  }
  p.need_pop = hit;
  ++p.k;
}

template <class C, bool INJECT>
__global__ void __launch_bounds__(256) jump_pair_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const int fine,
                                                        const int coarse, const DevPairOut pout,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM;
  const int factor = fine / coarse;                              // :231
  const float hf0 = (float)((double)s.T / (double)fine);         // :233
  const float hc0 = (float)factor * hf0;                         // :234
  const int kcap = INJECT ? inj.K : 4 * (coarse + s.max_jumps) + 64;

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range_n(rg); i += stride) {
    const uint64_t gp = range_lo(rg) + i;
    PairPath<C, INJECT> p;
    p.start(s, inj, i, (uint32_t)gp, (uint32_t)(gp >> 32));
    while (p.tf < s.T && p.k < kcap) pair_iteration<C, INJECT>(s, keys, inj, p, i, factor, hf0, hc0);   // :254

    if (pout.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        pout.terminal[(i * 2 + 0) * DIM + d] = p.xf[d];
        pout.terminal[(i * 2 + 1) * DIM + d] = p.xc[d];
      }
    }
    const float diff = eval_payoff<DIM>(po, p.xf) - eval_payoff<DIM>(po, p.xc);
    acc.add(diff, 0.0f, p.k * factor);
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

// The same pairs with persistent lanes (see jump_flat.cuh): a warp of jump_pair_kernel runs until its slowest lane is
// done, which for the coarse levels (coarse + #jumps outer iterations with #jumps ~ Poisson(rate T)) is about twice
// the mean.  Here a lane whose pair reached T starts its next pair at the next group boundary.  A group is six outer
// iterations: an even number (a Philox block of jump candidates serves two iterations) that consumes whole Philox
// blocks of normals for every factor and driver count, so all lanes refill together and every pair reads exactly the
// counters it reads in jump_pair_kernel -- bit-identical pairs, only the order of the fp64 additions differs.
template <class C>
__global__ void __launch_bounds__(256) jump_pair_flat_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                             const PhiloxKeys keys, const int fine, const int coarse,
                                                             double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM;
  constexpr int G = 6;
  const int factor = fine / coarse;
  const float hf0 = (float)((double)s.T / (double)fine);
  const float hc0 = (float)factor * hf0;
  const int kcap = 4 * (coarse + s.max_jumps) + 64;
  DevInject no_inject;
  no_inject.z = no_inject.zc = no_inject.jump_times = no_inject.marks = nullptr;
  no_inject.K = 0;

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool live = i < range_n(rg);
  PairPath<C, false> p;
  {
    const uint64_t gp = range_lo(rg) + (live ? i : 0);
    p.start(s, no_inject, i, (uint32_t)gp, (uint32_t)(gp >> 32));
  }
  while (live) {
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      if (p.tf < s.T && p.k < kcap) pair_iteration<C, false>(s, keys, no_inject, p, i, factor, hf0, hc0);
    }
    if (!(p.tf < s.T) || p.k >= kcap) {
      const float diff = eval_payoff<DIM>(po, p.xf) - eval_payoff<DIM>(po, p.xc);
      acc.add(diff, 0.0f, p.k * factor);
      i += stride;
      live = i < range_n(rg);
      if (live) {
        const uint64_t gp = range_lo(rg) + i;
        p.start(s, no_inject, i, (uint32_t)gp, (uint32_t)(gp >> 32));
      }
    }
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

// HESTON: the steps of both paths are HestonScheme.step (schemes.py:16-22; HestonSolver inherits multilevel_solve)
template <class C, bool INJECT, bool HESTON = false>
__global__ void __launch_bounds__(256) diffusion_pair_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                             const PhiloxKeys keys, const DevInject inj,
                                                             const int fine, const int coarse, const DevPairOut pout,
                                                             double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M;
  constexpr int NZ = BASE * M;
  constexpr int BPS = (NZ + 3) / 4;
  const int factor = fine / coarse;
  const float hf = (float)((double)s.T / (double)fine), hc = (float)factor * hf, sq = sqrtf(hf);

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range_n(rg); i += stride) {
    const uint64_t gp = range_lo(rg) + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float xf[kMaxDim], xc[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xf[d] = xc[d] = d < DIM ? s.x0[d] : 0.0f;
    NormalStream<NZ> normals;
    normals.init(plo, phi);
    for (int k = 0; k < coarse; ++k) {
      float s1[kMaxDim], s2[kMaxDim];
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) s1[d] = s2[d] = 0.0f;
      for (int q = 0; q < factor; ++q) {
        const int step = k * factor + q;
        float zn[NZ];
        if constexpr (!INJECT) {
          normals.next(keys, zn);
        } else {
          const float* zp = inj.z + (i * (uint64_t)fine + step) * (DIM * M);
#pragma unroll
          for (int e = 0; e < NZ; ++e) zn[e] = zp[e];
        }
        float z1[kMaxDim], z2[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
        for (int d = 0; d < BASE; ++d) {
          z1[d] = zn[d * M];
          z2[d] = M == 2 ? zn[d * M + 1] : 0.0f;
          w2[d] = 0.0f;
        }
        correlate<C>(s, z1, w1);
        if (M == 2) correlate<C>(s, z2, w2);
        if (HESTON) heston_step(s, xf, hf, w1[0] * sq, w1[1] * sq);
        else euler_step<C>(s, xf, hf, sq, w1, w2);
#pragma unroll
        for (int d = 0; d < BASE; ++d) {
          s1[d] = fmaf(w1[d], sq, s1[d]);
          if (M == 2) s2[d] = fmaf(w2[d], sq, s2[d]);
        }
      }
      // coarse step driven by the summed fine increments (:114-116)
      if (HESTON) heston_step(s, xc, hc, s1[0], s1[1]);
      else euler_step<C>(s, xc, hc, 1.0f, s1, s2);
    }
    if (pout.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        pout.terminal[(i * 2 + 0) * DIM + d] = xf[d];
        pout.terminal[(i * 2 + 1) * DIM + d] = xc[d];
      }
    }
    const float diff = eval_payoff<DIM>(po, xf) - eval_payoff<DIM>(po, xc);
    acc.add(diff, 0.0f, fine);
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
