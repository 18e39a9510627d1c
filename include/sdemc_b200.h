/*
 * sdemc_b200.h -- C-ABI of libsdemc_b200.so, the B200 (sm_100a) Monte Carlo engine behind
 * the sde_mc Python API.
 *
 * The reference (piers-hinds/sde_mc) has no FFI of its own: its boundary is the Python API
 * (SdeSolver.solve / multilevel_solve, mc_simple, mc_multilevel, mc_apply_cvs ...).  These entry
 * points are what a ctypes binding on the reference side would call in place of the Python step
 * loops; each one cites the reference code it replaces.  See INTEGRATION.md for the binding.
 *
 * Conventions
 *   - plain C, POD structs, no exceptions; every function returns 0 or a negative sdemc_status
 *   - all `d_*` pointers are DEVICE pointers owned by the caller (e.g. torch tensors)
 *   - `stream` is a cudaStream_t (NULL = legacy default stream); calls are asynchronous w.r.t. it
 *   - the library keeps no mutable global state and reads no environment variables: every choice of kernel
 *     is a function of the structs below; the caller provides scratch (`d_workspace`)
 *   - all path arithmetic is fp32 like the reference (torch default dtype); the moment
 *     accumulators are fp64; the coupled MLMC pair also exists with fp64 state (sdemc_mlmc_pair_f64)
 *   - every struct the caller fills starts with `struct_size` = sizeof(that struct): an entry point handed a struct
 *     of another size (a binding written against another header) returns SDEMC_ERR_BAD_ARG instead of reading past
 *     it; sdemc_abi_layout() reports the sizes this build expects
 */
#ifndef SDEMC_B200_H
#define SDEMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDEMC_ABI_VERSION 4
#define SDEMC_MAX_DIM 4
#define SDEMC_MAX_LEVELS 16

typedef enum {
  SDEMC_OK = 0,
  SDEMC_ERR_BAD_ARG = -1,      /* NULL pointer, negative size, inconsistent struct */
  SDEMC_ERR_UNSUPPORTED = -2,  /* (model, scheme, dim, m) combination without a kernel */
  SDEMC_ERR_CUDA = -3,         /* CUDA runtime error; text via sdemc_last_cuda_error() */
  SDEMC_ERR_NO_DEVICE = -4,    /* no sm_100 device */
  SDEMC_ERR_WORKSPACE = -5     /* workspace too small */
} sdemc_status;

/* Coefficient families.  Each reference Sde class maps to one (sde.py / levy.py):
 *   GEOMETRIC : drift a_i x_i, diffusion b1_i x_i (+ b2_i x_i), jump c_i x_i J
 *               Gbm sde.py:179-207, DoubleGbm :222-255, Merton :335-375,
 *               LevySde(ExpExampleLevy) levy.py:65-96,132-160
 *   ARITHMETIC: drift a_i, diffusion b1_i (+ b2_i), jump c_i J   (log-price models)
 *               LogGbm sde.py:210-219, LevySde(ExampleLevy) levy.py:99-129, LevySde(Levy2d) :163-192
 *   HESTON    : sde.py:258-279 with the drift-implicit square-root scheme schemes.py:16-22
 */
typedef enum {
  SDEMC_FAMILY_GEOMETRIC = 0, SDEMC_FAMILY_ARITHMETIC = 1, SDEMC_FAMILY_HESTON = 2,
  /* USER: coefficients given as CUDA expressions by an Sde subclass (the reference's plugin point, the abstract
   * drift / diffusion / jumps of sde.py:63-152).  Only libraries JIT-built from csrc/user_model.cu.in accept it
   * (entry points sdemc_user_mc_moments / sdemc_user_solve_paths, same signatures as the two below). */
  SDEMC_FAMILY_USER = 3
} sdemc_family;

/* schemes.py:5-22.  MILSTEIN is an extension (absent from the reference, parity unpinned). */
typedef enum { SDEMC_SCHEME_EULER = 0, SDEMC_SCHEME_HESTON = 1, SDEMC_SCHEME_MILSTEIN = 2 } sdemc_scheme;

/* Jump mark distributions: LogNormalJumpsSde.sample_jumps sde.py:325-326, LevySde.sample_jumps levy.py:85-87 */
typedef enum { SDEMC_MARKS_NONE = 0, SDEMC_MARKS_LOGNORMAL = 1, SDEMC_MARKS_ICDF = 2 } sdemc_marks;

/* How the kernels draw compound-Poisson jumps from Philox (no reference counterpart: the reference pre-samples
 * max_jumps exponential gaps per path, solvers.py:143-144,178).  QUEUE: per-thread shared-memory queue of
 * pre-drawn (time, mark) pairs -- for sparse jumps.  INLINE: one candidate per iteration in registers -- for
 * dense jumps.  The two strategies consume different Philox counters, so estimates agree statistically only. */
typedef enum { SDEMC_JUMPS_AUTO = 0, SDEMC_JUMPS_QUEUE = 1, SDEMC_JUMPS_INLINE = 2 } sdemc_jump_strategy;

/* Moments-only kernels for SHORT paths (few nominal steps, several jumps: MLMC level 0), where a warp of the plain
 * kernel idles until its slowest lane is done.  Lanes become persistent workers (csrc/jump_flat.cuh).
 *   AUTO    : sdemc_mc_moments picks ALIGNED when 2.1 sqrt(rate T) > 0.2 (num_steps + rate T), else the plain kernel;
 *             the single-level call of sdemc_mlmc_pair (coarse == 0) picks PACKED under the same rule
 *   OFF     : always the plain kernel
 *   ALIGNED : persistent lanes, every path consumes exactly the Philox counters of the plain / path-storing kernels
 *             (same seed => same paths as sdemc_solve_paths)
 *   PACKED  : 1-D lognormal-mark models only: normals, gaps and marks of two iterations from ONE Philox block
 *             (a stream of its own: same law, different paths than sdemc_solve_paths)
 *   PACKED_GENERIC : PACKED's stream driven through the generic iteration body (what PACKED is tested against) */
typedef enum {
  SDEMC_SHORT_AUTO = 0, SDEMC_SHORT_OFF = 1, SDEMC_SHORT_ALIGNED = 2, SDEMC_SHORT_PACKED = 3,
  SDEMC_SHORT_PACKED_GENERIC = 4
} sdemc_short_path;

/* options.py:179-321 */
typedef enum {
  SDEMC_PAYOFF_EURO_CALL = 0, SDEMC_PAYOFF_EURO_PUT = 1, SDEMC_PAYOFF_BINARY_AON = 2,
  SDEMC_PAYOFF_BASKET_ARITH = 3, SDEMC_PAYOFF_BASKET_GEOM = 4, SDEMC_PAYOFF_RAINBOW = 5,
  SDEMC_PAYOFF_DIGITAL = 6, SDEMC_PAYOFF_ASIAN_CALL = 7, SDEMC_PAYOFF_HESTON_RAINBOW = 8,
  SDEMC_PAYOFF_BEST_OF = 9
} sdemc_payoff_kind;

/* mc.py:84-91 -- which stored state the payoff is applied to */
typedef enum { SDEMC_INDEX_TERMINAL = 0 /* array index num_steps */, SDEMC_INDEX_ADAPTED = 1 /* last state */ } sdemc_index_mode;

/* The SDE + discretisation, extracted from (Sde, SdeSolver) objects. Replaces the attribute reads in
 * solvers.py:10-37,131-134 and the coefficient callbacks sde.py:63-152. */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_sde) */
  int32_t family;       /* sdemc_family */
  int32_t scheme;       /* sdemc_scheme */
  int32_t dim;          /* state dimension, 1..SDEMC_MAX_DIM (Sde.dim) */
  int32_t m;            /* Brownian drivers per component: 1 = 'diag', 2 = 'indep' (brown_dim/dim) */
  int32_t marks;        /* sdemc_marks; NONE => pure diffusion */
  int32_t num_steps;    /* SdeSolver.num_steps */
  int32_t max_jumps;    /* JumpDiffusionSolver.max_jumps solvers.py:133 (sizes the storage arrays) */
  int32_t exact_jumps;  /* solvers.py:214-217 */
  int32_t asian;        /* 1: AsianWrapper sde.py:378-406 -- component dim-1 integrates component 0 */
  int32_t jump_strategy; /* sdemc_jump_strategy: how Philox jump draws are organised (AUTO picks by rate*T/num_steps) */
  int32_t queue_depth;  /* QUEUE strategy: pre-drawn jumps per refill, a multiple of 4 in [4, 64]; 0 = sized from rate*T */
  int32_t short_path;   /* sdemc_short_path */
  float T;              /* SdeSolver.time_interval */
  float x0[SDEMC_MAX_DIM];
  float chol[SDEMC_MAX_DIM * SDEMC_MAX_DIM]; /* lower Cholesky of corr_matrix, row-major, stride SDEMC_MAX_DIM */
  float a[SDEMC_MAX_DIM];   /* drift coefficients        */
  float b1[SDEMC_MAX_DIM];  /* first-driver diffusion    */
  float b2[SDEMC_MAX_DIM];  /* second-driver diffusion   */
  float c[SDEMC_MAX_DIM];   /* jump coefficients         */
  float rate;               /* sde.jump_rate().sum()     */
  /* mark parameters: LOGNORMAL {alpha, gamma};  ICDF {cm, cp, mu, alpha, eps, lda, y1, y2, y3} levy.py:10-30 */
  float mark_p[12];
  /* Heston {r, kappa, theta, xi} sde.py:258-279 */
  float heston[4];
  /* USER family: parameters p[0..15] of the coefficient expressions */
  float user_p[16];
} sdemc_sde;

/* Option + discounter: options.py:156-176 (transform), :179-321 (payoffs), :324-337 (ConstantShortRate) */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_payoff) */
  int32_t kind;        /* sdemc_payoff_kind */
  int32_t log;         /* Option.log */
  int32_t index_mode;  /* sdemc_index_mode */
  float strike;
  float transform_discount; /* Option.discount (multiplies the spot before the payoff) */
  float aux;           /* AsianCall.time_interval */
  float df;            /* discounter(T): factor multiplying the payoff, mc.py:93 */
} sdemc_payoff;

/* Which paths, and where their noise comes from.  Philox4x32-10 keyed by `seed`, countered by the GLOBAL
 * path id, so results do not depend on grid shape or on how a range is split over GPUs. */
/* A path range that lives in DEVICE memory, written by sdemc_plan_mc / sdemc_plan_mlmc (below) and read by the kernels
 * of a later launch on the same stream: this is what lets "pilot run -> size the main run -> main run" (mc.py:418-440,
 * mlmc.py:77-97) be queued as ONE submission without a host read in between. */
typedef struct {
  uint64_t path_lo;
  uint64_t n_paths;
} sdemc_dev_range;

#define SDEMC_RANGE_COUNT_ON_HOST 1u /* with d_range: n_paths below is exact; only path_lo is taken from *d_range */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_range) */
  uint32_t flags;       /* SDEMC_RANGE_* */
  uint64_t seed;
  uint64_t path_lo;    /* first global path id of this call */
  uint64_t n_paths;    /* number of paths in this call */
  /* NULL, or a device pointer: the kernels then take (path_lo, n_paths) from *d_range when they RUN, and the two host
   * fields above only bound the grid (n_paths = 0: unknown, a full persistent grid).  Moments entry points only
   * (sdemc_mc_moments, sdemc_mlmc_pair, sdemc_mc_cv), without injected noise or per-path outputs.
   * SDEMC_RANGE_COUNT_ON_HOST: the caller knows the count and (*d_range).n_paths equals n_paths; what lives on the
   * device is only WHERE the ids start -- a launch captured in a CUDA graph and replayed on new path ids. */
  const sdemc_dev_range* d_range;
} sdemc_range;

/* Injected noise for the deterministic-parity mode (replaces the three overridable sampling methods
 * sample_corr_normals solvers.py:51-56, sample_jump_times :143-144, sample_one_jump :146-148).
 * All arrays are row-major, one row per path. K = number of loop iterations available. */
typedef struct {
  uint32_t struct_size;      /* sizeof(sdemc_inject) */
  int32_t K;
  const float* d_z;          /* (n, K, dim, m') unit normals; m' = m for the diffusion solver, 1 for the jump solver */
  const float* d_zc;         /* (n, K) common unit normal of the 2nd driver (jump solver, m == 2), else NULL */
  const float* d_jump_times; /* (n, max_jumps) cumulative jump times, else NULL */
  const float* d_marks;      /* (n, K) raw mark draw per iteration: N(0,1) for LOGNORMAL, U[0,1) for ICDF */
  int32_t total_steps;       /* sdemc_mc_cv only: the batch's total_steps; the reference's compensator sum drops the
                                interval with index total_steps-1 (varred.py:104,126-127).  0 = drop nothing. */
  int32_t reserved;
} sdemc_inject;

/* fp64 running moments; layout of the device array handed to the kernels (8 doubles). */
typedef struct {
  double sum;       /* sum of discounted payoffs (or MLMC corrections, or CV-corrected payoffs) */
  double sumsq;
  double sum_c;     /* terminal control  D(T) x_T[0] - x_0[0]  (mc.py:337) */
  double sumsq_c;
  double sum_pc;
  double n;         /* paths accumulated */
  double iters;     /* executed loop iterations over all paths */
  double reserved;
} sdemc_moments;

#define SDEMC_OUT_NO_TMA 1u /* path-storing kernels: 16-byte LSU stores even where the layout allows TMA tiles */

/* Optional trajectory outputs, layouts exactly as the reference allocates them (solvers.py:64-66,150-162).
 * S = num_steps for the diffusion solver, num_steps + max_jumps for the jump solver.  NULL = skip.
 * Rows (one per path) may be padded: pitch_* = floats between consecutive rows, 0 = dense (the reference's
 * contiguous layout).  With a pitch that is a multiple of 4 floats and 16-byte aligned bases the kernels write
 * 16-byte vectors covering whole 128-byte lines; any other pitch takes the 4-byte store path.  The padding between
 * the end of a row and its pitch belongs to the library: its contents after a call are unspecified (the TMA kernel
 * writes whole 128-byte tiles into it rather than have the tensor bound cut a tile inside a row). */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_paths_out) */
  uint32_t flags;       /* SDEMC_OUT_* */
  float* d_paths;       /* (n, S+1, dim) */
  float* d_left;        /* (n, S+1, dim)  state before the jump            (jump solver) */
  float* d_times;       /* (n, S+1)       time after each iteration        (jump solver) */
  float* d_jumps;       /* (n, S+1, dim)  applied jump mark                (jump solver) */
  float* d_normals;     /* (n, S, dim[, m]) Brownian increments dW actually used */
  float* d_payoffs;     /* (n)            discounted payoff per path */
  int32_t* d_iters;     /* (n)            executed iterations per path */
  float* d_terminal;    /* (n, dim)       the state the payoff is applied to (index_mode; ADAPTED without a payoff) */
  int32_t* d_total_steps; /* scalar: max over paths of executed iterations (atomicMax; caller zeroes it) */
  int64_t pitch_state;    /* row pitch of d_paths / d_left / d_jumps, 0 = (S+1)*dim */
  int64_t pitch_times;    /* row pitch of d_times, 0 = S+1 */
  int64_t pitch_normals;  /* row pitch of d_normals, 0 = S*dim*m */
} sdemc_paths_out;

/* Control-variate networks for the fused CV kernel: the BN-free Mlp of nets.py:39-93,
 * Linear(d+1,H) ReLU Linear(H,H) ReLU Linear(H,H) ReLU Linear(H,out), H <= 63. Weights are torch Linear
 * layouts (out_features, in_features) row-major fp32, device pointers.  out = dim * m for f (one output per Brownian
 * driver of every component, integrate_cv varred.py:202-214) and dim for g (varred.py:124). */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_mlp) */
  uint32_t cv_steps;    /* f only: the Brownian sum keeps the terms of the first cv_steps loop iterations -- the `tol`
                           trimming of integrate_cv varred.py:202-209 with cv_steps = remove_steps(tol, steps, T)
                           (helpers.py:71-74; steps = num_steps for diffusions, the batch's total_steps for the jump
                           solver).  0 = keep all.  The jump and compensator sums are never trimmed (varred.py:124-127). */
  const float* d_w[4];
  const float* d_b[4];
  int32_t in_dim, hidden, out_dim, n_hidden_layers; /* n_hidden_layers == 3 */
} sdemc_mlp;

int sdemc_version(void);
const char* sdemc_strerror(int rc);
const char* sdemc_last_cuda_error(void);
/* SM count, SM clock (kHz) and global memory of `device`; any pointer may be NULL. */
int sdemc_device_info(int device, int* sm_count, int* clock_khz, uint64_t* mem_bytes);
/* Bytes of device scratch every entry point needs (per concurrent call).  The caller zeroes it ONCE after allocation;
 * the kernels keep a few counters in it and leave them zero when they finish.  sdemc_solve_paths accepts NULL (it then
 * takes the LSU storing kernels: the TMA kernels hand their work out warp by warp through the workspace). */
uint64_t sdemc_workspace_bytes(void);
/* Layout handshake: writes up to `n` of the sizes {sdemc_sde, sdemc_payoff, sdemc_range, sdemc_inject, sdemc_moments,
 * sdemc_paths_out, sdemc_mlp, sdemc_coeffs_f64, sdemc_inject_f64} this library was compiled with and returns how many
 * there are (9).  A binding compares
 * them with its own struct definitions once at load time (sde_mc_b200/_lib.py does). */
int sdemc_abi_layout(uint32_t* sizes, int n);

/* E1/H3/H10 fused: time-stepping + payoff + (sum, sumsq, ...) reduction; nothing per-path touches HBM.
 * Replaces the bodies of mc_simple mc.py:101-123, mc_terminal_cv mc.py:352-375 and the solve() they call.
 * d_moments (sdemc_moments) is ACCUMULATED into (caller zeroes it once per estimator).
 * per_path (may be NULL) asks the SAME kernels to also write what each path contributed -- only d_payoffs, d_iters
 * and d_terminal are honoured, any other output pointer must be NULL (trajectories: sdemc_solve_paths).  This is how
 * the moments kernels are compared path by path with the path-storing kernel and the oracle. */
int sdemc_mc_moments(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                     const sdemc_paths_out* per_path, sdemc_moments* d_moments, void* d_workspace, void* stream);

/* P1-P3: the device payoff function of all kernels (Option.__call__ options.py:167-176 + payoffs :196-321) applied to
 * n states d_x (n, dim) -> d_out (n), times payoff->df.  index_mode is ignored. */
int sdemc_eval_payoff(const sdemc_payoff* payoff, int32_t dim, const float* d_x, uint64_t n, float* d_out, void* stream);

/* H3/H10 with storage: the solve() contract (solvers.py:68-88,164-226).  Noise is Philox (inject == NULL)
 * or injected (deterministic parity mode).  Any subset of outputs may be requested. */
int sdemc_solve_paths(const sdemc_sde* sde, const sdemc_payoff* payoff /* may be NULL */, const sdemc_range* range,
                      const sdemc_inject* inject /* may be NULL */, const sdemc_paths_out* out,
                      void* d_workspace, void* stream);

/* H11/E4 fused: coupled fine/coarse jump-adapted (or uniform-grid) pair sharing increments and jumps,
 * accumulating D(T) (P(fine) - P(coarse)).  coarse == 0 runs the single level `fine` (mlmc.py:44-53).
 * fp32 path state (the reference's fp32 jump pair asserts, solvers.py:264: dt is clamped at 0 here instead).
 * Uniform-grid pairs: the GEOMETRIC / ARITHMETIC families (EulerScheme) and HESTON (HestonScheme steps for both paths,
 * DiffusionSolver.multilevel_solve solvers.py:90-119 as inherited by HestonSolver).
 * With inject != NULL and d_pair_out != NULL writes (n, 2, dim) fp32 terminal (fine, coarse) states instead. */
int sdemc_mlmc_pair(const sdemc_sde* sde, const sdemc_payoff* payoff, int32_t fine, int32_t coarse,
                    const sdemc_range* range, const sdemc_inject* inject, sdemc_moments* d_moments,
                    float* d_pair_out, void* d_workspace, void* stream);

/* The same coupled jump-adapted pair with the path state, the coefficients, the marks and the payoff in fp64 -- the
 * precision the reference's jump MLMC runs in (solvers.py:228-307 under torch.set_default_dtype(float64)).
 * The float fields of sdemc_sde cannot carry 0.02 to 1e-12, so the coefficients come as doubles; sde supplies the
 * model shape (family, dim, m, marks, max_jumps, exact_jumps).  Jump models only (marks != NONE), coarse >= 1. */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_coeffs_f64) */
  uint32_t reserved;
  double T;
  double x0[SDEMC_MAX_DIM];
  double chol[SDEMC_MAX_DIM * SDEMC_MAX_DIM];
  double a[SDEMC_MAX_DIM], b1[SDEMC_MAX_DIM], b2[SDEMC_MAX_DIM], c[SDEMC_MAX_DIM];
  double rate;
  double mark_p[12];    /* as sdemc_sde.mark_p */
  double strike, transform_discount, aux, df; /* the payoff's real-valued fields (sdemc_payoff gives kind and log) */
} sdemc_coeffs_f64;
/* injected fp64 noise, layouts as sdemc_inject with K = outer iterations: d_z (n, K * fine/coarse, dim),
 * d_zc (n, K * fine/coarse), d_jump_times (n, max_jumps), d_marks (n, K) */
typedef struct {
  uint32_t struct_size; /* sizeof(sdemc_inject_f64) */
  int32_t K;
  const double* d_z;
  const double* d_zc;
  const double* d_jump_times;
  const double* d_marks;
} sdemc_inject_f64;
/* d_pair_out (may be NULL): (n, 2, dim) fp64 terminal (fine, coarse) states.  Moments are always accumulated. */
int sdemc_mlmc_pair_f64(const sdemc_sde* sde, const sdemc_coeffs_f64* coeffs, const sdemc_payoff* payoff, int32_t fine,
                        int32_t coarse, const sdemc_range* range, const sdemc_inject_f64* inject,
                        sdemc_moments* d_moments, double* d_pair_out, void* d_workspace, void* stream);

/* E5/E6/E7 fused: simulate + evaluate the control-variate MLPs f, g along each path on tensor cores and
 * accumulate gamma = payoff + sum f dW D + sum g D J - sum rate E[J] g D h  (varred.py:98-131).
 * g may be NULL for pure diffusions (varred.py:75-95). jump_mean = sde.jump_mean().
 * Model shapes: 1-D geometric 'diag' SDEs (Gbm, Merton: merton_cv_experiment.py) and the 2-D geometric 'indep'
 * inverse-cdf-mark SDE (LevySde(ExpExampleLevy, dim=2): levy_rainbow_cv_experiment.py:39-40); others UNSUPPORTED. */
int sdemc_mc_cv(const sdemc_sde* sde, const sdemc_payoff* payoff, float disc_rate, float jump_mean,
                const sdemc_mlp* f, const sdemc_mlp* g, const sdemc_range* range, const sdemc_inject* inject,
                sdemc_moments* d_moments, float* d_gamma_out /* (n) or NULL */, void* d_workspace, void* stream);

/* N2: size a run from its pilot on the device.  d_pilot = the pilot's (all-reduced) moments; writes the trial count
 *   N = ceil((1.96 se / eps)^2 pilot_trials)              find_num_trials mc.py:418-427 (se = standard error of the pilot)
 * rounded up to a multiple of `multiple_of` (ceil_mult mc.py:459; 0 or 1 = none) and capped at max_trials (0 = no cap)
 * to *d_trials_out, and rank `rank` of `world`'s contiguous share of the global path ids [path_base, path_base + N) to
 * *d_range_out -- to be handed to a moments entry point as sdemc_range.d_range on the same stream. */
int sdemc_plan_mc(const sdemc_moments* d_pilot, uint64_t pilot_trials, double eps, uint64_t multiple_of, uint64_t max_trials,
                  uint64_t path_base, int32_t rank, int32_t world, sdemc_dev_range* d_range_out, uint64_t* d_trials_out,
                  void* stream);
/* The MLMC allocation of get_optimal_trials mlmc.py:77-97 from n_levels pilot moments (d_pilot[l], `pilot_trials` pairs
 * each): N_l = ceil(1.96^2 / eps^2 sqrt(V_l h_l) sum_k sqrt(V_k / h_k)), h_l = T / d_levels[l] (device int32 array).
 * Writes N_l to d_trials_out[l] and this rank's share of level l's path ids to d_ranges_out[l]; the levels take
 * consecutive id ranges starting at path_base. */
int sdemc_plan_mlmc(const sdemc_moments* d_pilot, int32_t n_levels, const int32_t* d_levels, uint64_t pilot_trials, double T,
                    double eps, uint64_t max_trials, uint64_t path_base, int32_t rank, int32_t world,
                    sdemc_dev_range* d_ranges_out, uint64_t* d_trials_out, void* stream);

/* Test hook: the noise of `range`'s paths exactly as the kernels compute it from Philox (same device functions, same
 * fp32 / MUFU arithmetic), `count` values per path and array, row-major (n, count):
 *   BROWNIAN : d_a = unit normals of the Brownian stream in consumption order
 *   QUEUE    : d_a = cumulative jump times, d_b = raw mark draws of the QUEUE strategy (jump j of the path)
 *   INLINE   : d_a = Exp(1) gap candidate, d_b = raw mark candidate of iteration k (INLINE strategy)
 *   PACKED   : d_a = Brownian unit normal, d_b = gap candidate, d_c = raw mark candidate of iteration k (SHORT_PACKED)
 * Raw mark draws are N(0,1) for LOGNORMAL and U[0,1) for ICDF marks, as in sdemc_inject.d_marks.  Feeding these
 * arrays to the reference's loop (or the oracle) must reproduce what the moments kernels report per path. */
typedef enum { SDEMC_DRAWS_BROWNIAN = 0, SDEMC_DRAWS_QUEUE = 1, SDEMC_DRAWS_INLINE = 2, SDEMC_DRAWS_PACKED = 3 } sdemc_draws_kind;
int sdemc_debug_draws(const sdemc_sde* sde, const sdemc_range* range, int32_t kind, int32_t count, float* d_a,
                      float* d_b, float* d_c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDEMC_B200_H */
