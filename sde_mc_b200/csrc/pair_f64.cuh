// pair_f64.cuh -- the coupled jump-adapted MLMC pair with its path state in fp64.
//
// JumpDiffusionSolver.multilevel_solve (solvers.py:228-307) only runs in fp64 in the reference: in fp32 its
// `assert next_jump_time >= t` (:264) trips, and the level corrections P_f - P_c (1e-3 and smaller on the fine levels)
// drown in the 1e-7 relative rounding of states of order one.  pair.cuh is the fp32 pair (dt clamped at 0 instead of
// the assert); this is the same loop with double state, double coefficients (sdemc_coeffs_f64: the float fields of
// sdemc_sde cannot carry 0.02 to 1e-12), double marks / inverse cdf and a double payoff, so that with injected fp64
// noise it reproduces the reference's fp64 goldens to 1e-12.  With Philox noise it consumes exactly the counters of
// the fp32 pair kernel (fp32 normals, gaps and raw marks, widened to double).
// Generic in (family, dim <= 4, m, marks) at run time: the fine levels it matters for are a small share of an MLMC
// pass, so one kernel serves all models.
#pragma once
#include "pair.cuh"

namespace sdemc {

struct DevSde64 {
  int family, dim, m, marks, exact_jumps, max_jumps;
  double T, rate, inv_rate;
  double x0[kMaxDim], chol[kMaxDim * kMaxDim], a[kMaxDim], b1[kMaxDim], b2[kMaxDim], c[kMaxDim];
  double ln_alpha, ln_gamma;  // lognormal marks J = exp(gamma z + alpha) - 1 (sde.py:325-326)
  // inverse cdf (levy.py:10-30)
  double ic_y1, ic_y2, ic_y3, ic_mulda_cm, ic_inv_mu, ic_alpha, ic_lda_cm, ic_neg_inv_alpha, ic_malpha_cp, ic_lda,
      ic_x3_off, ic_eps_ma, ic_mulda_cp, ic_tol;
};
struct DevPayoff64 {
  int kind, log;
  double strike, tdisc, aux, df;
};
struct DevInject64 {
  const double* z;
  const double* zc;
  const double* jump_times;
  const double* marks;
  int K;
};

// InverseCdf.__call__ levy.py:19-30
__device__ inline double icdf_f64(const DevSde64& s, double y) {
  if (y <= s.ic_y1) return log(s.ic_mulda_cm * y) * s.ic_inv_mu - 1.0;
  if (y < s.ic_y2) return -pow(s.ic_alpha * (s.ic_lda_cm * y - s.ic_inv_mu) + 1.0, s.ic_neg_inv_alpha);
  if (y < s.ic_y3) return pow(s.ic_malpha_cp * (s.ic_lda * y - s.ic_x3_off) + s.ic_eps_ma, s.ic_neg_inv_alpha);
  return 1.0 - s.ic_inv_mu * log(s.ic_mulda_cp * (1.0 - y));
}
__device__ inline double mark_f64(const DevSde64& s, double raw) {
  if (s.marks == SDEMC_MARKS_LOGNORMAL) return exp(raw * s.ln_gamma + s.ln_alpha) - 1.0;
  return icdf_f64(s, raw + s.ic_tol);  // levy.py:86
}

// EulerScheme.step schemes.py:5-13 with increments w1 (per component) and w2 (second driver)
__device__ inline void euler_f64(const DevSde64& s, double (&x)[kMaxDim], double dt, const double (&w1)[kMaxDim],
                                 const double (&w2)[kMaxDim]) {
  const bool geo = s.family == SDEMC_FAMILY_GEOMETRIC;
  for (int i = 0; i < s.dim; ++i) {
    const double xi = x[i];
    double v = xi + (geo ? s.a[i] * xi : s.a[i]) * dt;
    if (s.m == 1) v = v + (geo ? s.b1[i] * xi : s.b1[i]) * w1[i];
    else v = v + ((geo ? s.b1[i] * xi : s.b1[i]) * w1[i] + (geo ? s.b2[i] * xi : s.b2[i]) * w2[i]);
    x[i] = v;
  }
}
__device__ inline void add_jump_f64(const DevSde64& s, double (&x)[kMaxDim], const double (&xb)[kMaxDim], double J) {
  const bool geo = s.family == SDEMC_FAMILY_GEOMETRIC;
  for (int i = 0; i < s.dim; ++i) x[i] = x[i] + (geo ? (s.c[i] * xb[i]) * J : s.c[i] * J);
}

// Option.__call__ options.py:167-176 and the payoffs :196-321, in double (same cases as eval_payoff in engine.cuh)
__device__ inline double eval_payoff_f64(const DevPayoff64& po, int dim, const double (&xin)[kMaxDim]) {
  double x[kMaxDim];
  for (int i = 0; i < dim; ++i) x[i] = po.tdisc * (po.log ? exp(xin[i]) : xin[i]);
  const double K = po.strike;
  double sp, r;
  switch (po.kind) {
    case SDEMC_PAYOFF_EURO_CALL: r = x[0] > K ? x[0] - K : 0.0; break;
    case SDEMC_PAYOFF_EURO_PUT: r = x[0] < K ? K - x[0] : 0.0; break;
    case SDEMC_PAYOFF_BINARY_AON: r = x[0] >= K ? x[0] : 0.0; break;
    case SDEMC_PAYOFF_BASKET_ARITH:
      sp = 0.0;
      for (int i = 0; i < dim; ++i) sp += x[i];
      sp = sp / (double)dim;
      r = sp > K ? sp - K : 0.0;
      break;
    case SDEMC_PAYOFF_BASKET_GEOM:
      sp = 0.0;
      for (int i = 0; i < dim; ++i) sp += log(x[i]);
      sp = exp(sp / (double)dim);
      r = sp > K ? sp - K : 0.0;
      break;
    case SDEMC_PAYOFF_RAINBOW:
      sp = x[0];
      for (int i = 1; i < dim; ++i) sp = fmax(sp, x[i]);
      r = sp > K ? sp - K : 0.0;
      break;
    case SDEMC_PAYOFF_DIGITAL: r = x[0] > K ? 1.0 : 0.0; break;
    case SDEMC_PAYOFF_BEST_OF:
      sp = x[0];
      for (int i = 1; i < dim; ++i) sp = fmax(sp, x[i]);
      r = fmax(sp, K);
      break;
    default: r = 0.0;  // AsianCall / HestonRainbow: no pair kernels for AsianWrapper / Heston (launch_pair.cu)
  }
  return r * po.df;
}

template <bool INJECT>
__global__ void __launch_bounds__(256) jump_pair_f64_kernel(const DevSde64 s, const DevPayoff64 po, const DevRange rg,
                                                            const PhiloxKeys keys, const DevInject64 inj, const int fine,
                                                            const int coarse, double* __restrict__ terminal,
                                                            double* __restrict__ d_moments, void* __restrict__ d_ws) {
  const int factor = fine / coarse;               // :231
  const double hf0 = s.T / (double)fine;           // :233
  const double hc0 = (double)factor * hf0;         // :234
  const int kcap = INJECT ? inj.K : 4 * (coarse + s.max_jumps) + 64;
  const int nz = s.dim + (s.m == 2 ? 1 : 0);       // normals per fine sub-step: one per component + the common one
  const int per_block = nz <= 3 ? kNormalsPerBlock / nz : 1, blocks = nz <= 3 ? 1 : (nz + 5) / 6;

  Accum acc;
  acc.zero();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < range_n(rg); i += stride) {
    const uint64_t gp = range_lo(rg) + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    double xf[kMaxDim], xc[kMaxDim];
    for (int d = 0; d < kMaxDim; ++d) xf[d] = xc[d] = d < s.dim ? s.x0[d] : 0.0;
    double tf = 0.0, tc = 0.0, tau = 0.0, Jcur = 0.0;
    int k = 0, jidx = -1;
    bool need_pop = true;
    // Philox mode: the streams of pair.cuh (NormalStream shift register over blocks of six normals; one block of
    // jump candidates per two outer iterations, InlineJumps::block_draws)
    float nbuf[2 * kNormalsPerBlock];
    int nleft = 0;
    uint32_t nblk = 0;
    float g0 = 0.f, r0 = 0.f, g1 = 0.f, r1 = 0.f;

    while (tf < s.T && k < kcap) {                 // :254
      double cand_gap = 0.0, cand_raw = 0.0;
      if (!INJECT) {
        if ((k & 1) == 0) {
          if (s.marks == SDEMC_MARKS_LOGNORMAL) InlineJumps<SDEMC_MARKS_LOGNORMAL>::block_draws((uint32_t)(k >> 1), plo, phi, keys, g0, r0, g1, r1);
          else InlineJumps<SDEMC_MARKS_ICDF>::block_draws((uint32_t)(k >> 1), plo, phi, keys, g0, r0, g1, r1);
        }
        cand_gap = (k & 1) ? (double)g1 : (double)g0;
        cand_raw = (k & 1) ? (double)r1 : (double)r0;
      }
      if (need_pop) {
        if (INJECT) {
          ++jidx;
          tau = jidx < s.max_jumps ? inj.jump_times[i * (uint64_t)s.max_jumps + jidx] : 1e300;
        } else {
          tau = fma(cand_gap, s.inv_rate, tau);
          Jcur = mark_f64(s, cand_raw);
        }
      }
      double s1[kMaxDim], s2[kMaxDim], xof[kMaxDim], xoc[kMaxDim];
      for (int d = 0; d < kMaxDim; ++d) { s1[d] = s2[d] = 0.0; xof[d] = xf[d]; }
      for (int q = 0; q < factor; ++q) {           // :259-278
        double zn[kMaxDim + 1];
        if (INJECT) {
          const uint64_t zi = i * (uint64_t)inj.K * factor + (uint64_t)k * factor + q;
          for (int d = 0; d < s.dim; ++d) zn[d] = inj.z[zi * s.dim + d];
          if (s.m == 2) zn[s.dim] = inj.zc[zi];
        } else {
          if (nleft == 0) {
            for (int r = 0; r < blocks; ++r) {
              uint32_t o[4];
              philox4x32_10(nblk++, STREAM_DIFFUSION, plo, phi, keys, o);
              philox_normals6(o, nbuf + kNormalsPerBlock * r);
            }
            nleft = per_block;
          }
          const int base = (per_block - nleft) * nz;
          for (int e = 0; e < nz; ++e) zn[e] = (double)nbuf[base + e];
          --nleft;
        }
        const double dt = fmax(fmin(hf0, fmin(tau, s.T) - tf), 0.0);
        const double sq = sqrt(dt);
        double w1[kMaxDim], w2[kMaxDim];
        for (int d = 0; d < s.dim; ++d) {          // torch.matmul(lower_cholesky, normals) :54, increments = z sqrt(dt)
          double accd = 0.0;
          for (int e = 0; e < s.dim; ++e) accd += s.chol[d * kMaxDim + e] * (zn[e] * sq);
          w1[d] = accd;
          w2[d] = s.m == 2 ? zn[s.dim] * sq : 0.0;
        }
        for (int d = 0; d < kMaxDim; ++d) xof[d] = xf[d];   // state before the LAST fine sub-step (:275)
        euler_f64(s, xf, dt, w1, w2);
        tf = tf + dt;
        for (int d = 0; d < s.dim; ++d) { s1[d] = s1[d] + w1[d]; s2[d] = s2[d] + w2[d]; }
      }
      const double dtc = fmax(fmin(hc0, fmin(tau, s.T) - tc), 0.0);   // :282-286
      for (int d = 0; d < kMaxDim; ++d) xoc[d] = xc[d];
      euler_f64(s, xc, dtc, s1, s2);
      tc = tc + dtc;
      const bool hit = fabs(tau - tf) <= 1e-12 + fabs(1e-5 * tf);      // :291 torch.isclose on the fine clock
      double Jc = 0.0;
      if (hit) Jc = INJECT ? mark_f64(s, inj.marks[i * (uint64_t)inj.K + k]) : Jcur;
      if (s.exact_jumps) {
        add_jump_f64(s, xf, xf, Jc);
        add_jump_f64(s, xc, xc, Jc);
      } else {
        add_jump_f64(s, xf, xof, Jc);
        add_jump_f64(s, xc, xoc, Jc);
      }
      need_pop = hit;
      ++k;
    }
    if (terminal) {
      for (int d = 0; d < s.dim; ++d) {
        terminal[(i * 2 + 0) * s.dim + d] = xf[d];
        terminal[(i * 2 + 1) * s.dim + d] = xc[d];
      }
    }
    const double diff = eval_payoff_f64(po, s.dim, xf) - eval_payoff_f64(po, s.dim, xc);
    acc.v[0] += diff;
    acc.v[1] = fma(diff, diff, acc.v[1]);
    acc.v[5] += 1.0;
    acc.v[6] += (double)(k * factor);
  }
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
