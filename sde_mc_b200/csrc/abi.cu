// abi.cu -- the extern "C" surface of libsdemc_b200.so (include/sdemc_b200.h)
#include <algorithm>
#include <string>

#include "cv.cuh"
#include "launch.cuh"
#include "plan.cuh"

namespace sdemc {

static thread_local std::string g_last_cuda_error;

void set_cuda_error(cudaError_t e, const char* where) {
  g_last_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
}

// Queue depth for the sparse-jump strategy: enough pre-drawn jumps that a refill is rare
// (mean + 1.65 sd of Poisson(rate T), + the terminating jump beyond T), as a multiple of 4, capped at 32.
// 0 selects the inline (dense) strategy.
static int choose_qdepth(const sdemc_sde& s) {
  if (s.marks == SDEMC_MARKS_NONE) return 0;
  if (s.jump_strategy == SDEMC_JUMPS_QUEUE && s.queue_depth > 0) return s.queue_depth;
  const double lam_T = (double)s.rate * (double)s.T;
  const double per_step = lam_T / (double)s.num_steps;
  int strategy = s.jump_strategy;
  if (strategy == SDEMC_JUMPS_AUTO) strategy = per_step >= 0.25 ? SDEMC_JUMPS_INLINE : SDEMC_JUMPS_QUEUE;
  if (strategy == SDEMC_JUMPS_INLINE) return 0;
  int q = (int)std::ceil((lam_T + 1.65 * std::sqrt(lam_T) + 2.0) / 4.0) * 4;
  if (q < 4) q = 4;
  if (q > 32) q = 32;
  return q;
}

static int solve_common(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                        const sdemc_inject* inject, const sdemc_paths_out* out, sdemc_moments* d_moments,
                        void* d_ws, void* stream, bool store, bool prefer_packed = false) {
  if (!valid_sde(sde) || !sized(range)) return SDEMC_ERR_BAD_ARG;
  if (!sized_or_null(payoff) || !sized_or_null(inject) || !sized_or_null(out)) return SDEMC_ERR_BAD_ARG;
  if (!store && (!d_moments || !d_ws || !payoff)) return SDEMC_ERR_BAD_ARG;
  if (store && !out) return SDEMC_ERR_BAD_ARG;
  if (range->d_range && (store || inject || out)) return SDEMC_ERR_BAD_ARG;  // device-resident ranges: plain moments runs
  // moments kernels only report what a path contributed; trajectories are sdemc_solve_paths' business
  if (!store && out && (out->d_paths || out->d_left || out->d_times || out->d_jumps || out->d_normals || out->d_total_steps))
    return SDEMC_ERR_BAD_ARG;
  if (range->n_paths == 0 && !range->d_range) return SDEMC_OK;
  const bool jumps = sde->marks != SDEMC_MARKS_NONE;
  if (inject) {
    if (!inject->d_z || inject->K < 1) return SDEMC_ERR_BAD_ARG;
    if (jumps && (!inject->d_jump_times || !inject->d_marks)) return SDEMC_ERR_BAD_ARG;
    if (jumps && sde->m == 2 && !inject->d_zc) return SDEMC_ERR_BAD_ARG;
    if (!jumps && inject->K != sde->num_steps) return SDEMC_ERR_BAD_ARG;
  }
  LaunchArgs a;
  a.sde = to_dev(*sde, sde->num_steps);
  a.payoff = to_dev(payoff);
  a.range = to_dev(*range);
  a.keys = make_philox_keys(range->seed);
  a.inject = to_dev(inject);
  a.use_inject = inject != nullptr;
  a.store = store;
  a.d_moments = reinterpret_cast<double*>(d_moments);
  a.d_ws = d_ws;
  a.stream = reinterpret_cast<cudaStream_t>(stream);
  a.qdepth = choose_qdepth(*sde);
  a.short_path = sde->short_path;
  a.prefer_packed = prefer_packed;
  a.no_tma = out && (out->flags & SDEMC_OUT_NO_TMA);
  if (jumps) {
    const int S = inject ? inject->K : sde->num_steps + sde->max_jumps;
    if (!valid_pitches(out, S, sde->dim, sde->dim * sde->m)) return SDEMC_ERR_BAD_ARG;
    a.out = to_dev(out, S, sde->dim, sde->dim * sde->m);
    return launch_jump(*sde, a);
  }
  if (!valid_pitches(out, sde->num_steps, sde->dim, sde->dim * sde->m)) return SDEMC_ERR_BAD_ARG;
  a.out = to_dev(out, sde->num_steps, sde->dim, sde->dim * sde->m);
  return launch_diffusion(*sde, a);
}

}  // namespace sdemc

namespace sdemc {
template <int DIM>
__global__ void payoff_kernel(const DevPayoff po, const float* __restrict__ x, uint64_t n, float* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    float xi[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) xi[d] = d < DIM ? x[i * DIM + d] : 0.0f;
    out[i] = eval_payoff<DIM>(po, xi);
  }
}
}  // namespace sdemc

using namespace sdemc;

extern "C" {

int sdemc_version(void) { return SDEMC_ABI_VERSION; }

const char* sdemc_strerror(int rc) {
  switch (rc) {
    case SDEMC_OK: return "ok";
    case SDEMC_ERR_BAD_ARG: return "bad argument (null pointer, size or inconsistent struct)";
    case SDEMC_ERR_UNSUPPORTED: return "no kernel for this (family, scheme, dim, m, marks) combination";
    case SDEMC_ERR_CUDA: return "CUDA runtime error (see sdemc_last_cuda_error)";
    case SDEMC_ERR_NO_DEVICE: return "no CUDA device of compute capability 10.x";
    case SDEMC_ERR_WORKSPACE: return "workspace too small";
  }
  return "unknown sdemc status";
}

const char* sdemc_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int sdemc_device_info(int device, int* sm_count, int* clock_khz, uint64_t* mem_bytes) {
  cudaDeviceProp p;
  SDEMC_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return SDEMC_ERR_NO_DEVICE;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (clock_khz) {
    int khz = 0;
    SDEMC_CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    *clock_khz = khz;
  }
  if (mem_bytes) *mem_bytes = (uint64_t)p.totalGlobalMem;
  return SDEMC_OK;
}

uint64_t sdemc_workspace_bytes(void) { return kWorkspaceBytes; }

int sdemc_abi_layout(uint32_t* sizes, int n) {
  const uint32_t all[9] = {(uint32_t)sizeof(sdemc_sde),        (uint32_t)sizeof(sdemc_payoff),    (uint32_t)sizeof(sdemc_range),
                           (uint32_t)sizeof(sdemc_inject),     (uint32_t)sizeof(sdemc_moments),   (uint32_t)sizeof(sdemc_paths_out),
                           (uint32_t)sizeof(sdemc_mlp),        (uint32_t)sizeof(sdemc_coeffs_f64), (uint32_t)sizeof(sdemc_inject_f64)};
  for (int i = 0; sizes && i < n && i < 9; ++i) sizes[i] = all[i];
  return 9;
}

int sdemc_mc_moments(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                     const sdemc_paths_out* per_path, sdemc_moments* d_moments, void* d_workspace, void* stream) {
  return solve_common(sde, payoff, range, nullptr, per_path, d_moments, d_workspace, stream, false);
}


int sdemc_plan_mc(const sdemc_moments* d_pilot, uint64_t pilot_trials, double eps, uint64_t multiple_of, uint64_t max_trials,
                  uint64_t path_base, int32_t rank, int32_t world, sdemc_dev_range* d_range_out, uint64_t* d_trials_out,
                  void* stream) {
  if (!d_pilot || !d_range_out || !d_trials_out || pilot_trials < 2 || !(eps > 0.0) || world < 1 || rank < 0 || rank >= world)
    return SDEMC_ERR_BAD_ARG;
  plan_mc_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const double*>(d_pilot), (double)pilot_trials, eps, multiple_of, max_trials, path_base, rank, world,
      reinterpret_cast<uint64_t*>(d_range_out), d_trials_out);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

int sdemc_plan_mlmc(const sdemc_moments* d_pilot, int32_t n_levels, const int32_t* d_levels, uint64_t pilot_trials, double T,
                    double eps, uint64_t max_trials, uint64_t path_base, int32_t rank, int32_t world,
                    sdemc_dev_range* d_ranges_out, uint64_t* d_trials_out, void* stream) {
  if (!d_pilot || !d_levels || !d_ranges_out || !d_trials_out || n_levels < 1 || n_levels > SDEMC_MAX_LEVELS ||
      pilot_trials < 2 || !(eps > 0.0) || !(T > 0.0) || world < 1 || rank < 0 || rank >= world)
    return SDEMC_ERR_BAD_ARG;
  plan_mlmc_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const double*>(d_pilot), n_levels, d_levels, (double)pilot_trials, T, eps, max_trials, path_base, rank,
      world, reinterpret_cast<uint64_t*>(d_ranges_out), d_trials_out);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

int sdemc_debug_draws(const sdemc_sde* sde, const sdemc_range* range, int32_t kind, int32_t count, float* d_a,
                      float* d_b, float* d_c, void* stream) {
  if (!valid_sde(sde) || !sized(range) || kind < SDEMC_DRAWS_BROWNIAN || kind > SDEMC_DRAWS_PACKED || count < 1 || !d_a)
    return SDEMC_ERR_BAD_ARG;
  if (kind != SDEMC_DRAWS_BROWNIAN && (!d_b || sde->marks == SDEMC_MARKS_NONE)) return SDEMC_ERR_BAD_ARG;
  if (kind == SDEMC_DRAWS_PACKED && !d_c) return SDEMC_ERR_BAD_ARG;
  if (range->n_paths == 0 && !range->d_range) return SDEMC_OK;
  DevRange rg;
  rg = to_dev(*range);
  if (rg.dyn) return SDEMC_ERR_BAD_ARG;
  return launch_debug_draws(*sde, rg, make_philox_keys(range->seed), kind, count, d_a, d_b, d_c,
                            reinterpret_cast<cudaStream_t>(stream));
}

int sdemc_eval_payoff(const sdemc_payoff* payoff, int32_t dim, const float* d_x, uint64_t n, float* d_out, void* stream) {
  if (!sized(payoff) || dim < 1 || dim > SDEMC_MAX_DIM || (n > 0 && (!d_x || !d_out))) return SDEMC_ERR_BAD_ARG;
  if (n == 0) return SDEMC_OK;
  const DevPayoff po = to_dev(payoff);
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dim) {
    case 1: payoff_kernel<1><<<grid, 256, 0, st>>>(po, d_x, n, d_out); break;
    case 2: payoff_kernel<2><<<grid, 256, 0, st>>>(po, d_x, n, d_out); break;
    case 3: payoff_kernel<3><<<grid, 256, 0, st>>>(po, d_x, n, d_out); break;
    default: payoff_kernel<4><<<grid, 256, 0, st>>>(po, d_x, n, d_out); break;
  }
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

int sdemc_solve_paths(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                      const sdemc_inject* inject, const sdemc_paths_out* out, void* d_workspace, void* stream) {
  return solve_common(sde, payoff, range, inject, out, nullptr, d_workspace, stream, true);
}

}  // extern "C"

extern "C" {

int sdemc_mlmc_pair(const sdemc_sde* sde, const sdemc_payoff* payoff, int32_t fine, int32_t coarse,
                    const sdemc_range* range, const sdemc_inject* inject, sdemc_moments* d_moments, float* d_pair_out,
                    void* d_workspace, void* stream) {
  if (!valid_sde(sde) || !sized(range) || !d_moments || !d_workspace) return SDEMC_ERR_BAD_ARG;
  if (!sized_or_null(payoff) || !sized_or_null(inject)) return SDEMC_ERR_BAD_ARG;
  if (fine < 1 || coarse < 0) return SDEMC_ERR_BAD_ARG;
  if (coarse == 0) {
    // single level (mlmc.py:44-53): the plain fused moments kernel on `fine` steps, payoff at the last state
    if (inject || d_pair_out || !payoff) return SDEMC_ERR_BAD_ARG;
    sdemc_sde lvl = *sde;
    lvl.num_steps = fine;
    sdemc_payoff po = *payoff;
    po.index_mode = SDEMC_INDEX_ADAPTED;
    return solve_common(&lvl, &po, range, nullptr, nullptr, d_moments, d_workspace, stream, false, /*prefer_packed=*/true);
  }
  if (fine % coarse != 0 || fine == coarse) return SDEMC_ERR_BAD_ARG;
  if (range->n_paths == 0 && !range->d_range) return SDEMC_OK;
  const bool jumps = sde->marks != SDEMC_MARKS_NONE;
  if (inject) {
    if (!inject->d_z || inject->K < 1) return SDEMC_ERR_BAD_ARG;
    if (jumps && (!inject->d_jump_times || !inject->d_marks)) return SDEMC_ERR_BAD_ARG;
    if (jumps && sde->m == 2 && !inject->d_zc) return SDEMC_ERR_BAD_ARG;
  }
  LaunchArgs a;
  a.sde = to_dev(*sde, fine);
  a.payoff = to_dev(payoff);
  a.payoff.index_mode = SDEMC_INDEX_ADAPTED;
  a.range = to_dev(*range);
  a.keys = make_philox_keys(range->seed);
  a.inject = to_dev(inject);
  a.use_inject = inject != nullptr;
  a.store = false;
  a.qdepth = 0;
  a.short_path = sde->short_path;
  a.d_moments = reinterpret_cast<double*>(d_moments);
  a.d_ws = d_workspace;
  a.stream = reinterpret_cast<cudaStream_t>(stream);
  a.out = to_dev(nullptr, 0);
  return launch_pair(*sde, a, fine, coarse, d_pair_out);
}

int sdemc_mlmc_pair_f64(const sdemc_sde* sde, const sdemc_coeffs_f64* co, const sdemc_payoff* payoff, int32_t fine,
                        int32_t coarse, const sdemc_range* range, const sdemc_inject_f64* inject, sdemc_moments* d_moments,
                        double* d_pair_out, void* d_workspace, void* stream) {
  if (!valid_sde(sde) || !sized(co) || !sized(range) || !d_moments || !d_workspace) return SDEMC_ERR_BAD_ARG;
  if (!sized_or_null(payoff) || !sized_or_null(inject)) return SDEMC_ERR_BAD_ARG;
  if (fine < 1 || coarse < 1 || fine % coarse != 0 || fine == coarse) return SDEMC_ERR_BAD_ARG;
  if (sde->marks == SDEMC_MARKS_NONE || sde->asian || sde->family == SDEMC_FAMILY_HESTON || sde->family == SDEMC_FAMILY_USER ||
      sde->scheme != SDEMC_SCHEME_EULER)
    return SDEMC_ERR_UNSUPPORTED;
  if (inject && (!inject->d_z || !inject->d_jump_times || !inject->d_marks || inject->K < 1 || (sde->m == 2 && !inject->d_zc)))
    return SDEMC_ERR_BAD_ARG;
  if (range->d_range && inject) return SDEMC_ERR_BAD_ARG;
  if (range->n_paths == 0 && !range->d_range) return SDEMC_OK;
  return launch_pair_f64(*sde, *co, payoff, fine, coarse, to_dev(*range), make_philox_keys(range->seed), inject,
                         reinterpret_cast<double*>(d_moments), d_pair_out, d_workspace, reinterpret_cast<cudaStream_t>(stream));
}

// f has dim * m outputs (one per driver of every component), g has dim; both see (t, x): dim + 1 inputs
static bool mlp_ok(const sdemc_mlp* m, int in_dim, int out_dim) {
  if (!sized(m)) return false;
  for (int i = 0; i < 4; ++i)
    if (!m->d_w[i] || !m->d_b[i]) return false;
  return m->in_dim == in_dim && m->out_dim == out_dim && m->n_hidden_layers == 3 && m->hidden >= 1 && m->hidden <= 63;
}
static DevMlp mlp_dev(const sdemc_mlp* m) {
  DevMlp d;
  std::memset(&d, 0, sizeof d);
  if (m) {
    for (int i = 0; i < 4; ++i) { d.w[i] = m->d_w[i]; d.b[i] = m->d_b[i]; }
    d.in_dim = m->in_dim; d.hidden = m->hidden; d.out_dim = m->out_dim;
  }
  return d;
}

int sdemc_mc_cv(const sdemc_sde* sde, const sdemc_payoff* payoff, float disc_rate, float jump_mean, const sdemc_mlp* f,
                const sdemc_mlp* g, const sdemc_range* range, const sdemc_inject* inject, sdemc_moments* d_moments,
                float* d_gamma_out, void* d_workspace, void* stream) {
  if (!valid_sde(sde) || !sized(payoff) || !sized(range) || !d_moments || !d_workspace) return SDEMC_ERR_BAD_ARG;
  if (!sized_or_null(inject) || !sized_or_null(f) || !sized_or_null(g)) return SDEMC_ERR_BAD_ARG;
  const bool jumps = sde->marks != SDEMC_MARKS_NONE;
  if (!mlp_ok(f, sde->dim + 1, sde->dim * sde->m) || (jumps && !mlp_ok(g, sde->dim + 1, sde->dim))) return SDEMC_ERR_UNSUPPORTED;
  if (range->n_paths == 0 && !range->d_range) return SDEMC_OK;
  if (inject) {
    if (!inject->d_z || inject->K < 1) return SDEMC_ERR_BAD_ARG;
    if (jumps && (!inject->d_jump_times || !inject->d_marks)) return SDEMC_ERR_BAD_ARG;
    if (jumps && sde->m == 2 && !inject->d_zc) return SDEMC_ERR_BAD_ARG;
    if (!jumps && inject->K != sde->num_steps) return SDEMC_ERR_BAD_ARG;
  }
  LaunchArgs a;
  a.sde = to_dev(*sde, sde->num_steps);
  a.payoff = to_dev(payoff);
  a.payoff.index_mode = SDEMC_INDEX_ADAPTED;
  a.range = to_dev(*range);
  a.keys = make_philox_keys(range->seed);
  a.inject = to_dev(inject);
  a.use_inject = inject != nullptr;
  a.store = false;
  a.qdepth = 0;
  a.d_moments = reinterpret_cast<double*>(d_moments);
  a.d_ws = d_workspace;
  a.stream = reinterpret_cast<cudaStream_t>(stream);
  a.out = to_dev(nullptr, 0);
  DevCv cv;
  cv.disc_rate_l2e = (float)((double)disc_rate * 1.4426950408889634);
  cv.comp_c = (float)(-(double)sde->rate * (double)jump_mean);
  cv.last_interval = (inject && inject->total_steps > 0) ? inject->total_steps - 1 : 0x7fffffff;
  cv.brownian_steps = f->cv_steps > 0 ? (int)f->cv_steps : 0x7fffffff;
  cv.gamma_out = d_gamma_out;
  return launch_cv(*sde, a, mlp_dev(f), mlp_dev(g), cv);
}

}  // extern "C"
