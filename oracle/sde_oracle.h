/*
 * sde_oracle.h -- TEST INFRASTRUCTURE.  CPU restatement of the reference's hot path
 * (piers-hinds/sde_mc, /root/reference/sde_mc).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (sde_mc_b200/, libsdemc_b200.so) never does.
 *
 * One scalar loop per path -- the loop a CUDA thread runs -- with every arithmetic operation rounded to the
 * working precision in the reference's own operation order (compiled with -ffp-contract=off).
 * Pinned against tests/golden/*.npz, which were produced by the unmodified reference under noise injection
 * (tests/golden/make_golden.py), so parity is PINNED for every function in this file.
 */
#ifndef SDE_ORACLE_H
#define SDE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_GEOMETRIC = 0, ORACLE_ARITHMETIC = 1, ORACLE_HESTON = 2 };
enum { ORACLE_MARKS_NONE = 0, ORACLE_MARKS_LOGNORMAL = 1, ORACLE_MARKS_ICDF = 2 };

/* Same fields as sdemc_sde (include/sdemc_b200.h) but in double, so the fp64 run of the reference
 * (needed for the jump MLMC) can be matched to 1e-12. */
typedef struct {
  int32_t family, scheme, dim, m, marks, num_steps, max_jumps, exact_jumps, asian, pad_;
  double T;
  double x0[4];
  double chol[16];
  double a[4], b1[4], b2[4], c[4];
  double rate;
  double mark_p[12];
  double heston[4];
} oracle_sde;

typedef struct {
  int32_t kind, log, index_mode, pad_;
  double strike, transform_discount, aux, df;
} oracle_payoff;

typedef struct {
  const float* w[4];
  const float* b[4];
  int32_t in_dim, hidden, out_dim, n_hidden_layers;
} oracle_mlp;

/* DiffusionSolver.solve solvers.py:68-88 (EulerScheme/HestonScheme schemes.py:5-22).
 * z (n,steps,dim,m) unit normals; paths (n,steps+1,dim); normals (n,steps,dim,m) = increments used (or NULL). */
void oracle_diffusion_f32(const oracle_sde* s, int64_t n, const float* z, float* paths, float* normals);
void oracle_diffusion_f64(const oracle_sde* s, int64_t n, const double* z, double* paths, double* normals);

/* DiffusionSolver.multilevel_solve solvers.py:90-119; z (n,fine,dim,m). */
void oracle_diffusion_pair_f32(const oracle_sde* s, int64_t n, int fine, int coarse, const float* z,
                               float* paths_fine, float* paths_coarse);

/* JumpDiffusionSolver.solve solvers.py:164-226.  K iterations of noise are available per path:
 * z (n,K,dim), zc (n,K) or NULL, jump_times (n,max_jumps), marks (n,K).
 * Outputs (any may be NULL) sized for K iterations: paths/left/jumps (n,K+1,dim), times (n,K+1),
 * normals (n,K,dim,m), iters (n).  Slots past a path's own last iteration hold what the reference's idle
 * iterations write there (frozen state, zero increments).  Returns total_steps = max_i iters[i], or -1 if
 * some path needed more than K iterations / max_jumps jumps. */
int oracle_jump_f32(const oracle_sde* s, int64_t n, int K, const float* z, const float* zc, const float* jump_times,
                    const float* marks, float* paths, float* left, float* times, float* jumps, float* normals,
                    int32_t* iters);
int oracle_jump_f64(const oracle_sde* s, int64_t n, int K, const double* z, const double* zc,
                    const double* jump_times, const double* marks, double* paths, double* left, double* times,
                    double* jumps, double* normals, int32_t* iters);

/* JumpDiffusionSolver.multilevel_solve solvers.py:228-307.  K_outer outer iterations available:
 * z (n,K_outer*factor,dim), zc (n,K_outer*factor) or NULL, marks (n,K_outer).  fine_last/coarse_last (n,dim).
 * The f32 variant clamps dt >= 0 where the reference's fp32 run would assert (solvers.py:264). */
int oracle_jump_pair_f32(const oracle_sde* s, int64_t n, int fine, int coarse, int K_outer, const float* z,
                         const float* zc, const float* jump_times, const float* marks, float* fine_last,
                         float* coarse_last, int32_t* iters);
int oracle_jump_pair_f64(const oracle_sde* s, int64_t n, int fine, int coarse, int K_outer, const double* z,
                         const double* zc, const double* jump_times, const double* marks, double* fine_last,
                         double* coarse_last, int32_t* iters);

/* Option.__call__ options.py:167-321 on (n,dim) states; out (n) WITHOUT the discount factor df. */
void oracle_payoff_f32(const oracle_payoff* p, int64_t n, int dim, const float* x, float* out);
void oracle_payoff_f64(const oracle_payoff* p, int64_t n, int dim, const double* x, double* out);

/* InverseCdf.__call__ levy.py:19-30 on raw uniforms (UNIFORM_TOL/3 already added by the caller or not). */
void oracle_icdf_f32(const double* mark_p, int64_t n, const float* u, float* out);

/* apply_adapted_control_variates varred.py:98-131 / apply_diffusion_control_variate varred.py:75-95, per path.
 * Consumes the arrays of oracle_jump_f32 (K slots) / oracle_diffusion_f32; total_steps = number of valid
 * iterations (global).  g may be NULL.  gamma (n) = payoff + brownian cv + jump cv + compensator. */
void oracle_cv_gamma_jump_f32(const oracle_sde* s, int64_t n, int K, int total_steps, int brownian_steps, double disc_rate,
                              double jump_mean, const oracle_mlp* f, const oracle_mlp* g, const float* paths,
                              const float* left, const float* times, const float* jumps, const float* normals,
                              const float* payoffs, float* gamma);
void oracle_cv_gamma_diffusion_f32(const oracle_sde* s, int64_t n, int brownian_steps, double disc_rate, const oracle_mlp* f,
                                   const float* paths, const float* normals, const float* payoffs, float* gamma);

/* Philox4x32-10 replay of the kernels' in-register noise (same counter layout, libm transcendentals instead
 * of MUFU approximations).  Fills unit normals / jump data exactly as the device draws them so the oracle can
 * re-run a Philox-driven kernel call deterministically.  See sde_mc_b200/csrc/philox.cuh for the layout. */
void oracle_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
