// engine.cuh -- device-side building blocks of the fused path kernels.
//
// One thread owns one path: state, time, step size, next jump and RNG counters live in registers for all
// time steps (reference: the Python loops DiffusionSolver.solve solvers.py:68-88 and
// JumpDiffusionSolver.solve solvers.py:164-226, which issue 20-244 ATen ops per step on (bs, dim) tensors).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/sdemc_b200.h"
#include "philox.cuh"

namespace sdemc {

constexpr int kMaxDim = SDEMC_MAX_DIM;
constexpr int kBlock = 256;  // threads per CTA of the moments kernels

// ------------------------------------------------------------------------------------------------------------
// Device copies of the problem (passed by value as kernel parameters -> constant bank / uniform registers)
// ------------------------------------------------------------------------------------------------------------
struct DevSde {
  int num_steps, max_jumps, exact_jumps;
  int milstein;  // 1: Milstein correction 1/2 b b' (dW^2 - dt) on top of the Euler step (extension, 'diag' noise)
  float T, h0, sqrt_h0;
  float x0[kMaxDim];
  float chol[kMaxDim * kMaxDim];
  float a[kMaxDim], b1[kMaxDim], b2[kMaxDim], c[kMaxDim];
  // uniform-grid constants: a h ; b * sqrt(h)
  float ah[kMaxDim], b1s[kMaxDim], b2s[kMaxDim];
  float neg2ln2_b1s2;  // -2 ln2 (b1 sqrt h)^2 for the 1-D path that folds the volatility into the Box-Muller radius
  float rate, inv_rate;
  // lognormal marks  J = 2^(z * g2 + a2) - 1
  float ln_a2, ln_g2;
  // icdf marks (levy.py:10-30), pre-combined constants
  float ic_y1, ic_y2, ic_y3, ic_mulda_cm, ic_inv_mu_ln2, ic_alpha, ic_lda_cm, ic_inv_mu, ic_neg_inv_alpha,
      ic_malpha_cp, ic_lda, ic_x3_off, ic_eps_ma, ic_mulda_cp, ic_tol;
  // heston
  float hes_r, hes_kappa, hes_xi, hes_kth, hes_halfxi2;
  // user-defined coefficient expressions (SDEMC_FAMILY_USER): their parameters
  float user_p[16];
};

struct DevPayoff {
  int kind, log, index_mode;
  float strike, tdisc, aux, df;
};

struct DevRange {
  uint64_t path_lo, n_paths;
  const uint64_t* dyn;  // device-resident {path_lo, n_paths} overriding the two above (sdemc_range.d_range), or nullptr
  bool count_on_host;   // with dyn: n_paths above is exact, only path_lo comes from device memory (SDEMC_RANGE_COUNT_ON_HOST)
};
// The range a moments kernel works on.  With `dyn` set it is read from device memory at run time: the launch was
// queued before the range was known (pilot -> size the run -> run without a host read in between, mc.py:418-440,
// mlmc.py:77-97).  Read where it is used (once per path) through a volatile load, so the values never occupy
// registers across the step loops; without `dyn` these are the constant-bank reads they always were.
__device__ __forceinline__ uint64_t range_dyn_word(const DevRange& rg, int w) {
  uint64_t v;
  asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(rg.dyn + w));
  return v;
}
__device__ __forceinline__ uint64_t range_n(const DevRange& rg) { return rg.dyn ? range_dyn_word(rg, 1) : rg.n_paths; }
__device__ __forceinline__ uint64_t range_lo(const DevRange& rg) { return rg.dyn ? range_dyn_word(rg, 0) : rg.path_lo; }
// The same with the choice made at compile time, for the kernels in which a few instructions per path or two live
// registers are measurable: the 1-D uniform-grid kernel sits at its register limit (the run-time form cost 3 % of the
// C2 throughput: 1.608e12 vs 1.659e12) and the short-path kernels spend ~340 instructions on a whole path (3 % of an
// MLMC pass).  RANGE_HOST: the constant-bank reads these kernels always had; RANGE_DEVICE: device memory.
// RANGE_LO_DEVICE: only the first path id comes from device memory, the count is the launch parameter (a captured
// launch replayed on new path ids): loop control keeps its constant-bank operand -- with the count loaded from memory
// ptxas compiles the persistent-lane loop of jump_flat1d_kernel into 1360 instructions instead of 800 (+6 % run time).
enum { RANGE_HOST = 0, RANGE_DEVICE = 1, RANGE_LO_DEVICE = 2 };
// RANGE_DEVICE: the CTA copies the two words into shared memory once (range_stage, first statement of the kernel) and
// every path reads them back with a volatile LDS: a fresh global load at the start of every path sat on the critical
// path of the persistent-lane kernels (MLMC level 0: 4.59 ms against 4.26 ms with a host range).
__shared__ uint64_t g_sh_range[2];
template <int MODE>
__device__ __forceinline__ void range_stage(const DevRange& rg) {
  if (MODE != RANGE_HOST) {
    if (threadIdx.x < 2) g_sh_range[threadIdx.x] = range_dyn_word(rg, (int)threadIdx.x);
    __syncthreads();
  }
}
// a plain shared read, not volatile asm: inside the divergent persistent-lane loop a volatile access pins the control
// flow around it (no predication of the path restart; 1360 instead of 800 instructions in jump_flat1d_kernel)
__device__ __forceinline__ uint64_t range_staged_word(int w) { return g_sh_range[w]; }
template <int MODE>
__device__ __forceinline__ uint64_t range_n(const DevRange& rg) { return MODE == RANGE_DEVICE ? range_staged_word(1) : rg.n_paths; }
template <int MODE>
__device__ __forceinline__ uint64_t range_lo(const DevRange& rg) { return MODE != RANGE_HOST ? range_staged_word(0) : rg.path_lo; }

struct DevInject {
  const float* z;
  const float* zc;
  const float* jump_times;
  const float* marks;
  int K;
};

struct DevOut {
  float* paths;
  float* left;
  float* times;
  float* jumps;
  float* normals;
  float* payoffs;
  int* iters;
  float* terminal;  // (n, dim) state the payoff is applied to
  int* total_steps;
  int S;  // allocated iterations per path (rows are S+1 / S long)
  uint64_t pitch_state, pitch_times, pitch_normals;  // floats between consecutive path rows
};

// per-path outputs of the MOMENTS kernels (sdemc_mc_moments per_path): what every path contributed, so the fast
// kernels can be compared path by path with the path-storing kernel and the oracle.  All NULL in production runs.
struct DevPerPath {
  float* payoffs;
  int* iters;
  float* terminal;
  __host__ __device__ __forceinline__ bool any() const { return payoffs != nullptr || iters != nullptr || terminal != nullptr; }
};
template <int DIM, class X>
__device__ __forceinline__ void write_per_path(const DevPerPath& pp, uint64_t i, float pay, int iters, const X& xp) {
#ifdef SDEMC_NO_PER_PATH  // A/B builds only (tools/build_variant.sh): what the per-path hook costs the hot loops
  return;
#endif
  if (pp.payoffs) pp.payoffs[i] = pay;
  if (pp.iters) pp.iters[i] = iters;
  if (pp.terminal) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) pp.terminal[i * DIM + d] = xp[d];
  }
}

__host__ __device__ __forceinline__ DevPerPath per_path_of_out(const DevOut& o) {
  DevPerPath pp;
  pp.payoffs = o.payoffs;
  pp.iters = o.iters;
  pp.terminal = o.terminal;
  return pp;
}

// compile-time configuration of a kernel instance
template <int FAMILY_, int DIM_, int M_, int MARKS_, bool ASIAN_>
struct Cfg {
  static constexpr int FAMILY = FAMILY_, DIM = DIM_, M = M_, MARKS = MARKS_;
  static constexpr bool ASIAN = ASIAN_;
  static constexpr int BASE = ASIAN_ ? DIM_ - 1 : DIM_;  // components driven by noise
};

// How a stream of NZ normals per step maps onto Philox blocks of 6 normals: NZ in {1,2,3} -> 6/NZ steps share a
// block; larger NZ -> ceil(NZ/6) blocks per step (slots beyond NZ unused).
constexpr int steps_per_group(int nz) { return nz <= 3 ? kNormalsPerBlock / nz : 1; }
constexpr int blocks_per_group(int nz) { return nz <= 3 ? 1 : (nz + kNormalsPerBlock - 1) / kNormalsPerBlock; }

// ------------------------------------------------------------------------------------------------------------
// user-defined coefficients (SDEMC_FAMILY_USER): a(t, x)_i, b(t, x)_i ('diag' noise) and c(t, x_base, J)_i, the
// abstract methods of the reference's Sde base class (sde.py:63-152).  A JIT translation unit (user_model.cu.in)
// defines them from the expressions of an Sde subclass BEFORE including this header; the stock library never
// instantiates the USER family, the stubs only keep the templates well-formed.
// ------------------------------------------------------------------------------------------------------------
#ifndef SDEMC_USER_MODEL
__device__ __forceinline__ float sdemc_user_drift(int, float, const float*, const float*) { return 0.0f; }
__device__ __forceinline__ float sdemc_user_diffusion(int, float, const float*, const float*) { return 0.0f; }
__device__ __forceinline__ float sdemc_user_jump(int, float, const float*, float, const float*) { return 0.0f; }
#endif

// ------------------------------------------------------------------------------------------------------------
// marks
// ------------------------------------------------------------------------------------------------------------
// InverseCdf.__call__ levy.py:19-30 with lg2/ex2 on the XU pipe; pow(x, p) = ex2(p * lg2(x)).
__device__ __forceinline__ float icdf_mark(const DevSde& s, float u) {
  const float y = u + s.ic_tol;  // levy.py:86
  // The two power-law branches (|x| < 1: all but (c- + c+)/(mu lda) of the mass, < 1 % for the example models) are
  // one expression with per-side constants, so a warp does not diverge on the sign of the jump ...
  const bool right = !(y < s.ic_y2);
  const float A = right ? s.ic_malpha_cp : s.ic_alpha;
  const float Lc = right ? s.ic_lda : s.ic_lda_cm;
  const float off = right ? s.ic_x3_off : s.ic_inv_mu;
  const float B = right ? s.ic_eps_ma : 1.0f;
  float r = fast_ex2(s.ic_neg_inv_alpha * fast_lg2(fmaf(A, fmaf(Lc, y, -off), B)));
  r = right ? r : -r;
  // ... and only the rare exponential tails branch
  if (y <= s.ic_y1) r = fmaf(fast_lg2(s.ic_mulda_cm * y), s.ic_inv_mu_ln2, -1.0f);
  else if (!(y < s.ic_y3)) r = fmaf(-s.ic_inv_mu_ln2, fast_lg2(s.ic_mulda_cp * (1.0f - y)), 1.0f);
  return r;
}

// raw draw -> jump mark.  LOGNORMAL: raw ~ N(0,1), sde.py:325-326.  ICDF: raw ~ U[0,1), levy.py:85-87.
template <int MARKS>
__device__ __forceinline__ float mark_from_raw(const DevSde& s, float raw) {
  if (MARKS == SDEMC_MARKS_LOGNORMAL) return fast_ex2(fmaf(raw, s.ln_g2, s.ln_a2)) - 1.0f;
  if (MARKS == SDEMC_MARKS_ICDF) return icdf_mark(s, raw);
  return 0.0f;
}

// ------------------------------------------------------------------------------------------------------------
// schemes
// ------------------------------------------------------------------------------------------------------------
// w[j][i] = sum_k L[i][k] z[k][j]   (torch.matmul(lower_cholesky, normals) solvers.py:54), unit variance
template <class C>
__device__ __forceinline__ void correlate(const DevSde& s, const float (&z)[kMaxDim], float (&w)[kMaxDim]) {
  if (C::BASE == 1) {  // lower_cholesky is [[1.]] for a single driven component (solvers.py:33-36)
    w[0] = z[0];
    return;
  }
#pragma unroll
  for (int i = 0; i < C::BASE; ++i) {
    float acc = s.chol[i * kMaxDim] * z[0];
#pragma unroll
    for (int k = 1; k <= i; ++k) acc = fmaf(s.chol[i * kMaxDim + k], z[k], acc);
    w[i] = acc;
  }
}

// EulerScheme.step schemes.py:5-13 for a step (dt, sqrt(dt)) with unit-variance correlated normals w1 (per
// component) and w2 (second driver).  Algebraically regrouped per family:
//   geometric  x (1 + a dt + b1 sq w1 + b2 sq w2)      arithmetic  x + a dt + b1 sq w1 + b2 sq w2
template <class C>
__device__ __forceinline__ void euler_step(const DevSde& s, float (&x)[kMaxDim], float dt, float sq,
                                           const float (&w1)[kMaxDim], const float (&w2)[kMaxDim], float t = 0.0f) {
  if (C::FAMILY == SDEMC_FAMILY_USER) {  // x + a(t, x) dt + b(t, x) dW, coefficients at the pre-step state
    float xin[kMaxDim];
#pragma unroll
    for (int i = 0; i < kMaxDim; ++i) xin[i] = x[i];
#pragma unroll
    for (int i = 0; i < C::BASE; ++i)
      x[i] = xin[i] + sdemc_user_drift(i, t, xin, s.user_p) * dt + sdemc_user_diffusion(i, t, xin, s.user_p) * (sq * w1[i]);
    return;
  }
  const float x_first = x[0];
#pragma unroll
  for (int i = 0; i < C::BASE; ++i) {
    // relative increment g = a dt + b1 sq w1 (+ b2 sq w2).  Keeping g small and adding it with one FMA
    // (x + x g) avoids the systematic rounding of a pre-added (1 + a dt) constant over hundreds of steps.
    float g = s.a[i] * dt;
    g = fmaf(s.b1[i] * sq, w1[i], g);
    if (C::M == 2) g = fmaf(s.b2[i] * sq, w2[i], g);
    // Milstein (geometric: b b' = b1^2 x): + 1/2 b1^2 (dW^2 - dt) with dW = sq w, relative to x like the rest of g.
    // Written on dW so that callers passing an accumulated increment (sq = 1, w = sum dW: the coarse path of an
    // MLMC pair) get the right correction too.
    if (C::M == 1 && C::FAMILY == SDEMC_FAMILY_GEOMETRIC && s.milstein) {
      const float dw = sq * w1[i];
      g = fmaf(0.5f * s.b1[i] * s.b1[i], fmaf(dw, dw, -dt), g);
    }
    if (C::FAMILY == SDEMC_FAMILY_GEOMETRIC) x[i] = fmaf(x[i], g, x[i]);
    else x[i] += g;
  }
  if (C::ASIAN) x[C::DIM - 1] = fmaf(x_first, dt, x[C::DIM - 1]);  // AsianWrapper.drift sde.py:396-397
}

// same, on the uniform grid of DiffusionSolver (h, sqrt(h) folded into per-launch constants ah = a h, b*s = b sqrt(h))
template <class C>
__device__ __forceinline__ void euler_step_uniform(const DevSde& s, float (&x)[kMaxDim], const float (&w1)[kMaxDim],
                                                   const float (&w2)[kMaxDim], float t = 0.0f) {
  if (C::FAMILY == SDEMC_FAMILY_USER) {
    euler_step<C>(s, x, s.h0, s.sqrt_h0, w1, w2, t);
    return;
  }
  const float x_first = x[0];
#pragma unroll
  for (int i = 0; i < C::BASE; ++i) {
    float g = fmaf(s.b1s[i], w1[i], s.ah[i]);
    if (C::M == 2) g = fmaf(s.b2s[i], w2[i], g);
    if (C::M == 1 && C::FAMILY == SDEMC_FAMILY_GEOMETRIC && s.milstein)
      g = fmaf(0.5f * s.b1s[i] * s.b1s[i], fmaf(w1[i], w1[i], -1.0f), g);
    if (C::FAMILY == SDEMC_FAMILY_GEOMETRIC) x[i] = fmaf(x[i], g, x[i]);
    else x[i] += g;
  }
  if (C::ASIAN) x[C::DIM - 1] = fmaf(x_first, s.h0, x[C::DIM - 1]);
}

// HestonScheme.step schemes.py:16-22 + Heston.quadratic_parameters sde.py:275-279 + solve_quadratic
// helpers.py:36-48 over a step of length h with the correlated increments (dw0, dw1).
__device__ __forceinline__ void heston_step(const DevSde& s, float (&x)[kMaxDim], float h, float dw0, float dw1) {
  const float S = x[0], v = x[1];
  x[0] = fmaf(sqrtf(v) * S, dw0, fmaf(s.hes_r * S, h, S));
  const float qa = -1.0f - s.hes_kappa * h;
  const float qb = s.hes_xi * dw1;
  const float qc = fmaf(s.hes_kth, h, v) - s.hes_halfxi2 * h;
  const float disc = sqrtf(fmaf(qb, qb, -4.0f * qa * qc));
  const float inv = 1.0f / (2.0f * qa);
  const float y = fmaxf((-qb + disc) * inv, (-qb - disc) * inv);
  x[1] = y * y;
}
// the same on the uniform grid; w = correlated unit normals, increments are w * sqrt(h)
__device__ __forceinline__ void heston_step_uniform(const DevSde& s, float (&x)[kMaxDim], const float (&w)[kMaxDim]) {
  heston_step(s, x, s.h0, w[0] * s.sqrt_h0, w[1] * s.sqrt_h0);
}

// sde.jumps(t, x_base, J): geometric c x J (sde.py:374-375, levy.py:154-155), arithmetic c J (levy.py:120-121)
template <class C>
__device__ __forceinline__ void add_jump(const DevSde& s, float (&x)[kMaxDim], const float (&xb)[kMaxDim], float J,
                                         float t = 0.0f) {
#pragma unroll
  for (int i = 0; i < C::BASE; ++i) {
    if (C::FAMILY == SDEMC_FAMILY_USER) x[i] += J != 0.0f ? sdemc_user_jump(i, t, xb, J, s.user_p) : 0.0f;
    else if (C::FAMILY == SDEMC_FAMILY_GEOMETRIC) x[i] = fmaf(s.c[i] * xb[i], J, x[i]);
    else x[i] = fmaf(s.c[i], J, x[i]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// payoff  (Option.__call__ options.py:167-176; payoffs options.py:196-321)
// ------------------------------------------------------------------------------------------------------------
// exp / log are the full-precision expf / logf (1 ulp): the payoff runs once per path, so their cost is nil, and the
// 54 reference vectors hold at 2e-6 (the 2-ulp intrinsics do not, for the geometric basket and log-price payoffs).
template <int DIM>
__device__ __forceinline__ float eval_payoff(const DevPayoff& po, const float (&xin)[kMaxDim]) {
  float x[kMaxDim];
#pragma unroll
  for (int i = 0; i < DIM; ++i) x[i] = po.tdisc * (po.log ? expf(xin[i]) : xin[i]);
  const float K = po.strike;
  float sp, r;
  switch (po.kind) {
    case SDEMC_PAYOFF_EURO_CALL: r = x[0] > K ? x[0] - K : 0.0f; break;
    case SDEMC_PAYOFF_EURO_PUT: r = x[0] < K ? K - x[0] : 0.0f; break;
    case SDEMC_PAYOFF_BINARY_AON: r = x[0] >= K ? x[0] : 0.0f; break;
    case SDEMC_PAYOFF_BASKET_ARITH:
      sp = x[0];
#pragma unroll
      for (int i = 1; i < DIM; ++i) sp += x[i];
      sp = sp / (float)DIM;
      r = sp > K ? sp - K : 0.0f;
      break;
    case SDEMC_PAYOFF_BASKET_GEOM:
      sp = logf(x[0]);
#pragma unroll
      for (int i = 1; i < DIM; ++i) sp += logf(x[i]);
      sp = expf(sp / (float)DIM);
      r = sp > K ? sp - K : 0.0f;
      break;
    case SDEMC_PAYOFF_RAINBOW:
      sp = x[0];
#pragma unroll
      for (int i = 1; i < DIM; ++i) sp = fmaxf(sp, x[i]);
      r = sp > K ? sp - K : 0.0f;
      break;
    case SDEMC_PAYOFF_DIGITAL: r = x[0] > K ? 1.0f : 0.0f; break;
    case SDEMC_PAYOFF_ASIAN_CALL:
      sp = DIM > 1 ? x[DIM > 1 ? 1 : 0] / po.aux : 0.0f;
      if (po.log) sp = expf(sp);
      r = sp > K ? sp - K : 0.0f;
      break;
    case SDEMC_PAYOFF_HESTON_RAINBOW:
      sp = x[0];
#pragma unroll
      for (int i = 2; i < DIM; i += 2) sp = fmaxf(sp, x[i]);
      r = sp > K ? sp - K : 0.0f;
      break;
    case SDEMC_PAYOFF_BEST_OF:
      sp = x[0];
#pragma unroll
      for (int i = 1; i < DIM; ++i) sp = fmaxf(sp, x[i]);
      r = fmaxf(sp, K);
      break;
    default: r = 0.0f;
  }
  return r * po.df;
}

// ------------------------------------------------------------------------------------------------------------
// moment accumulation: per-thread fp64 -> warp shuffle -> shared -> per-CTA partial in the workspace ->
// the last CTA to finish folds all partials (fixed order: deterministic) into the caller's sdemc_moments.
// Replaces the fp32 tensor accumulators of mc.py:116-120 / mlmc.py:49-53.
// ------------------------------------------------------------------------------------------------------------
constexpr int kNumMoments = 8;

struct Accum {
  double v[kNumMoments];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < kNumMoments; ++i) v[i] = 0.0;
  }
  // p: discounted payoff (or correction), c: control value, iters: executed iterations
  __device__ __forceinline__ void add(float p, float c, int iters) {
    const double dp = (double)p, dc = (double)c;
    v[0] += dp;
    v[1] = fma(dp, dp, v[1]);
    v[2] += dc;
    v[3] = fma(dc, dc, v[3]);
    v[4] = fma(dp, dc, v[4]);
    v[5] += 1.0;
    v[6] += (double)iters;
  }
};

// workspace layout: [0, 64) bytes: uint32 ticket of this reduction at 0, the two scheduling words of the TMA storing
// kernels at 16 (tma_gang.cuh: WarpTasks) -- all zero between calls; then gridDim.x * kNumMoments doubles
__device__ __forceinline__ void block_reduce_and_publish(const Accum& acc, double* d_moments, void* d_ws) {
  __shared__ double sh[32][kNumMoments];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  double v[kNumMoments];
#pragma unroll
  for (int i = 0; i < kNumMoments; ++i) {
    v[i] = acc.v[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kNumMoments; ++i) sh[warp][i] = v[i];
  }
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(d_ws);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(d_ws) + 64);
  if (threadIdx.x < kNumMoments) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += sh[w][threadIdx.x];
    partials[(size_t)blockIdx.x * kNumMoments + threadIdx.x] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(ticket, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < kNumMoments) {
      double t = 0.0;
      for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(&partials[(size_t)b * kNumMoments + threadIdx.x]);
      d_moments[threadIdx.x] += t;  // accumulate across launches of one estimator (same stream => ordered)
    }
    if (threadIdx.x == 0) *ticket = 0u;  // re-arm for the next launch
  }
}

}  // namespace sdemc
