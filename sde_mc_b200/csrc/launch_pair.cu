// launch_pair.cu -- instantiation + dispatch of the MLMC pair kernels (pair.cuh)
#include <cmath>
#include <type_traits>

#include "launch.cuh"
#include "pair.cuh"
#include "pair_f64.cuh"

namespace sdemc {
namespace {

template <class Kernel>
int run(Kernel kernel, const LaunchArgs& a, int fine, int coarse, float* d_terminal) {
  int grid = 0;
  int rc = pick_grid(kernel, 0, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  DevPairOut pout;
  pout.terminal = d_terminal;
  kernel<<<grid, kBlock, 0, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, fine, coarse, pout, a.d_moments,
                                        a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

// moments-only pairs with persistent lanes (pair.cuh): same rule as launch_jump.cu:want_flat with the pair's
// coarse + rate T outer iterations.  sdemc_sde.short_path overrides (OFF: never, anything else but AUTO: always).
bool want_flat_pair(const DevSde& s, int coarse, int short_path) {
  if (short_path != SDEMC_SHORT_AUTO) return short_path != SDEMC_SHORT_OFF;
  const double lam_T = (double)s.rate * (double)s.T;
  return 2.1 * std::sqrt(lam_T) > 0.2 * ((double)coarse + lam_T);
}

template <class C>
int run_flat(const LaunchArgs& a, int fine, int coarse) {
  auto kernel = jump_pair_flat_kernel<C>;
  int grid = 0;
  int rc = pick_grid(kernel, 0, a.range.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, kBlock, 0, a.stream>>>(a.sde, a.payoff, a.range, a.keys, fine, coarse, a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

template <class C>
int jump_by_mode(const LaunchArgs& a, int fine, int coarse, float* t) {
  if (!a.use_inject && t == nullptr && want_flat_pair(a.sde, coarse, a.short_path)) return run_flat<C>(a, fine, coarse);
  return a.use_inject ? run(jump_pair_kernel<C, true>, a, fine, coarse, t) : run(jump_pair_kernel<C, false>, a, fine, coarse, t);
}
template <class C>
int diff_by_mode(const LaunchArgs& a, int fine, int coarse, float* t) {
  return a.use_inject ? run(diffusion_pair_kernel<C, true>, a, fine, coarse, t)
                      : run(diffusion_pair_kernel<C, false>, a, fine, coarse, t);
}

template <int FAMILY, int M, int MARKS>
int jump_by_dim(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* t) {
  switch (s.dim) {
    case 1: return jump_by_mode<Cfg<FAMILY, 1, M, MARKS, false>>(a, fine, coarse, t);
    case 2: return jump_by_mode<Cfg<FAMILY, 2, M, MARKS, false>>(a, fine, coarse, t);
    case 3: return jump_by_mode<Cfg<FAMILY, 3, M, MARKS, false>>(a, fine, coarse, t);
    case 4: return jump_by_mode<Cfg<FAMILY, 4, M, MARKS, false>>(a, fine, coarse, t);
  }
  return SDEMC_ERR_UNSUPPORTED;
}
template <int FAMILY, int M>
int diff_by_dim(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* t) {
  switch (s.dim) {
    case 1: return diff_by_mode<Cfg<FAMILY, 1, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 2: return diff_by_mode<Cfg<FAMILY, 2, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 3: return diff_by_mode<Cfg<FAMILY, 3, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
    case 4: return diff_by_mode<Cfg<FAMILY, 4, M, SDEMC_MARKS_NONE, false>>(a, fine, coarse, t);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace

int launch_pair_f64(const sdemc_sde& s, const sdemc_coeffs_f64& co, const sdemc_payoff* payoff, int fine, int coarse,
                    const DevRange& rg, const PhiloxKeys& keys, const sdemc_inject_f64* inject, double* d_moments,
                    double* d_terminal, void* d_ws, cudaStream_t stream) {
  DevSde64 d;
  std::memset(&d, 0, sizeof d);
  d.family = s.family; d.dim = s.dim; d.m = s.m; d.marks = s.marks; d.exact_jumps = s.exact_jumps; d.max_jumps = s.max_jumps;
  d.T = co.T;
  d.rate = co.rate;
  d.inv_rate = co.rate > 0.0 ? 1.0 / co.rate : 0.0;
  for (int i = 0; i < kMaxDim; ++i) { d.x0[i] = co.x0[i]; d.a[i] = co.a[i]; d.b1[i] = co.b1[i]; d.b2[i] = co.b2[i]; d.c[i] = co.c[i]; }
  for (int i = 0; i < kMaxDim * kMaxDim; ++i) d.chol[i] = co.chol[i];
  if (s.marks == SDEMC_MARKS_LOGNORMAL) {
    d.ln_alpha = co.mark_p[0];
    d.ln_gamma = co.mark_p[1];
  } else {  // levy.py:10-18, pre-combined as in host_common.cuh (fp32) and the oracle
    const double cm = co.mark_p[0], cp = co.mark_p[1], mu = co.mark_p[2], al = co.mark_p[3], eps = co.mark_p[4], lda = co.mark_p[5];
    d.ic_y1 = co.mark_p[6]; d.ic_y2 = co.mark_p[7]; d.ic_y3 = co.mark_p[8];
    d.ic_mulda_cm = mu * lda / cm; d.ic_inv_mu = 1.0 / mu; d.ic_alpha = al; d.ic_lda_cm = lda / cm;
    d.ic_neg_inv_alpha = -1.0 / al; d.ic_malpha_cp = -al / cp; d.ic_lda = lda;
    d.ic_x3_off = cm / mu + cm * ((std::pow(eps, -al) - 1.0) / al);
    d.ic_eps_ma = std::pow(eps, -al); d.ic_mulda_cp = mu * lda / cp;
    d.ic_tol = 5.960464477539063e-08 / 3.0;
  }
  DevPayoff64 po;
  po.kind = payoff ? payoff->kind : SDEMC_PAYOFF_EURO_CALL;
  po.log = payoff ? payoff->log : 0;
  po.strike = co.strike; po.tdisc = co.transform_discount; po.aux = co.aux; po.df = co.df;
  DevInject64 inj;
  std::memset(&inj, 0, sizeof inj);
  if (inject) { inj.z = inject->d_z; inj.zc = inject->d_zc; inj.jump_times = inject->d_jump_times; inj.marks = inject->d_marks; inj.K = inject->K; }
  int grid = 0;
  int rc = inject ? pick_grid(jump_pair_f64_kernel<true>, 0, rg.n_paths, &grid) : pick_grid(jump_pair_f64_kernel<false>, 0, rg.n_paths, &grid);
  if (rc != SDEMC_OK) return rc;
  if (inject) jump_pair_f64_kernel<true><<<grid, kBlock, 0, stream>>>(d, po, rg, keys, inj, fine, coarse, d_terminal, d_moments, d_ws);
  else jump_pair_f64_kernel<false><<<grid, kBlock, 0, stream>>>(d, po, rg, keys, inj, fine, coarse, d_terminal, d_moments, d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

int launch_pair(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* d_terminal) {
  if (s.asian) return SDEMC_ERR_UNSUPPORTED;
  if (s.family == SDEMC_FAMILY_HESTON) {  // HestonSolver.multilevel_solve: the uniform-grid pair on HestonScheme steps
    if (s.dim != 2 || s.m != 1 || s.marks != SDEMC_MARKS_NONE) return SDEMC_ERR_UNSUPPORTED;
    using HC = Cfg<SDEMC_FAMILY_HESTON, 2, 1, SDEMC_MARKS_NONE, false>;
    return a.use_inject ? run(diffusion_pair_kernel<HC, true, true>, a, fine, coarse, d_terminal)
                        : run(diffusion_pair_kernel<HC, false, true>, a, fine, coarse, d_terminal);
  }
  if (s.marks == SDEMC_MARKS_NONE) {
    if (s.family == SDEMC_FAMILY_GEOMETRIC)
      return s.m == 1 ? diff_by_dim<SDEMC_FAMILY_GEOMETRIC, 1>(s, a, fine, coarse, d_terminal)
                      : diff_by_dim<SDEMC_FAMILY_GEOMETRIC, 2>(s, a, fine, coarse, d_terminal);
    return s.m == 1 ? diff_by_dim<SDEMC_FAMILY_ARITHMETIC, 1>(s, a, fine, coarse, d_terminal)
                    : diff_by_dim<SDEMC_FAMILY_ARITHMETIC, 2>(s, a, fine, coarse, d_terminal);
  }
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 1 && s.marks == SDEMC_MARKS_LOGNORMAL)
    return jump_by_dim<SDEMC_FAMILY_GEOMETRIC, 1, SDEMC_MARKS_LOGNORMAL>(s, a, fine, coarse, d_terminal);
  if (s.family == SDEMC_FAMILY_GEOMETRIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return jump_by_dim<SDEMC_FAMILY_GEOMETRIC, 2, SDEMC_MARKS_ICDF>(s, a, fine, coarse, d_terminal);
  if (s.family == SDEMC_FAMILY_ARITHMETIC && s.m == 2 && s.marks == SDEMC_MARKS_ICDF)
    return jump_by_dim<SDEMC_FAMILY_ARITHMETIC, 2, SDEMC_MARKS_ICDF>(s, a, fine, coarse, d_terminal);
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
