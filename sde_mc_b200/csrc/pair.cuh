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

template <class C, bool INJECT>
__global__ void __launch_bounds__(256) jump_pair_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const int fine,
                                                        const int coarse, const DevPairOut pout,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  constexpr int DIM = C::DIM, BASE = C::BASE, M = C::M, MARKS = C::MARKS;
  constexpr int NZ = BASE + (M == 2 ? 1 : 0);
  constexpr int BPS = (NZ + 3) / 4;
  const int factor = fine / coarse;                              // :231
  const float hf0 = (float)((double)s.T / (double)fine);         // :233
  const float hc0 = (float)factor * hf0;                         // :234
  const int kcap = INJECT ? inj.K : 4 * (coarse + s.max_jumps) + 64;

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rg.n_paths; i += stride) {
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float xf[kMaxDim], xc[kMaxDim], xof[kMaxDim], xoc[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xf[d] = xc[d] = xof[d] = xoc[d] = d < DIM ? s.x0[d] : 0.0f;
    float tf = 0.0f, tc = 0.0f;
    int k = 0;
    bool need_pop = true;
    typename std::conditional<INJECT, InjectJumps<MARKS>, InlineJumps<MARKS>>::type src;
    if constexpr (INJECT) src.init(s, inj, i);
    else src.init(plo, phi);
    NormalStream<NZ> normals;
    normals.init(plo, phi);

    while (tf < s.T && k < kcap) {                               // :254
      src.begin_iter(s, keys, k);
      src.advance(s, keys, need_pop);
      const float tau = src.tau;
      float s1[kMaxDim], s2[kMaxDim];
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) s1[d] = s2[d] = 0.0f;
      for (int q = 0; q < factor; ++q) {                         // :259-278
        const int sub = k * factor + q;
        float zn[NZ];
        if constexpr (!INJECT) {
          normals.next(keys, zn);
        } else {
          const uint64_t zi = i * (uint64_t)inj.K * factor + sub;
#pragma unroll
          for (int d = 0; d < BASE; ++d) zn[d] = inj.z[zi * DIM + d];
          if (M == 2) zn[BASE] = inj.zc[zi];
        }
        const float dt = fmaxf(fminf(hf0, fminf(tau, s.T) - tf), 0.0f);  // stateless mesh, see jump.cuh
        const float sq = fast_sqrt(dt);
        float z1[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
        for (int d = 0; d < BASE; ++d) z1[d] = zn[d];
        correlate<C>(s, z1, w1);
#pragma unroll
        for (int d = 0; d < BASE; ++d) w2[d] = M == 2 ? zn[BASE] : 0.0f;
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) xof[d] = xf[d];        // state before the LAST fine sub-step (:275)
        euler_step<C>(s, xf, dt, sq, w1, w2);
        tf += dt;
#pragma unroll
        for (int d = 0; d < BASE; ++d) {
          s1[d] = fmaf(w1[d], sq, s1[d]);
          if (M == 2) s2[d] = fmaf(w2[d], sq, s2[d]);
        }
      }
      const float dtc = fmaxf(fminf(hc0, fminf(tau, s.T) - tc), 0.0f);   // :282-286
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) xoc[d] = xc[d];
      euler_step<C>(s, xc, dtc, 1.0f, s1, s2);
      tc += dtc;
      const bool hit = fabsf(tau - tf) <= fmaf(fabsf(tf), 1e-5f, 1e-12f);   // :291 (on the fine clock)
      const float Jc = hit ? src.mark(s, k) : 0.0f;
      if (s.exact_jumps) {
        add_jump<C>(s, xf, xf, Jc);
        add_jump<C>(s, xc, xc, Jc);
      } else {
        add_jump<C>(s, xf, xof, Jc);
        add_jump<C>(s, xc, xoc, Jc);
      }
      need_pop = hit;
      ++k;
    }

    if (pout.terminal) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        pout.terminal[(i * 2 + 0) * DIM + d] = xf[d];
        pout.terminal[(i * 2 + 1) * DIM + d] = xc[d];
      }
    }
    const float diff = eval_payoff<DIM>(po, xf) - eval_payoff<DIM>(po, xc);
    acc.add(diff, 0.0f, k * factor);
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

template <class C, bool INJECT>
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
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rg.n_paths; i += stride) {
    const uint64_t gp = rg.path_lo + i;
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
        euler_step<C>(s, xf, hf, sq, w1, w2);
#pragma unroll
        for (int d = 0; d < BASE; ++d) {
          s1[d] = fmaf(w1[d], sq, s1[d]);
          if (M == 2) s2[d] = fmaf(w2[d], sq, s2[d]);
        }
      }
      euler_step<C>(s, xc, hc, 1.0f, s1, s2);   // coarse step driven by the summed fine increments (:114-116)
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
