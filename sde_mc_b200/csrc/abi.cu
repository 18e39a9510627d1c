// abi.cu -- the extern "C" surface of libsdemc_b200.so (include/sdemc_b200.h)
#include <mutex>
#include <string>

#include "launch.cuh"

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
                        void* d_ws, void* stream, bool store) {
  if (!valid_sde(sde) || !range) return SDEMC_ERR_BAD_ARG;
  if (!store && (!d_moments || !d_ws || !payoff)) return SDEMC_ERR_BAD_ARG;
  if (store && !out) return SDEMC_ERR_BAD_ARG;
  if (range->n_paths == 0) return SDEMC_OK;
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
  a.range.path_lo = range->path_lo;
  a.range.n_paths = range->n_paths;
  a.keys = make_philox_keys(range->seed);
  a.inject = to_dev(inject);
  a.use_inject = inject != nullptr;
  a.store = store;
  a.d_moments = reinterpret_cast<double*>(d_moments);
  a.d_ws = d_ws;
  a.stream = reinterpret_cast<cudaStream_t>(stream);
  a.qdepth = choose_qdepth(*sde);
  if (jumps) {
    const int S = inject ? inject->K : sde->num_steps + sde->max_jumps;
    a.out = to_dev(out, S);
    return launch_jump(*sde, a);
  }
  a.out = to_dev(out, sde->num_steps);
  return launch_diffusion(*sde, a);
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

int sdemc_mc_moments(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                     sdemc_moments* d_moments, void* d_workspace, void* stream) {
  return solve_common(sde, payoff, range, nullptr, nullptr, d_moments, d_workspace, stream, false);
}

int sdemc_solve_paths(const sdemc_sde* sde, const sdemc_payoff* payoff, const sdemc_range* range,
                      const sdemc_inject* inject, const sdemc_paths_out* out, void* d_workspace, void* stream) {
  return solve_common(sde, payoff, range, inject, out, nullptr, d_workspace, stream, true);
}

}  // extern "C"

// ---- not built yet in this revision: explicit, loud status codes (never a silent fallback) ----------------------
extern "C" {

int sdemc_mlmc_pair(const sdemc_sde*, const sdemc_payoff*, int32_t, int32_t, int32_t, const sdemc_range*,
                    const sdemc_inject*, sdemc_moments*, void*, void*, void*) {
  return SDEMC_ERR_UNSUPPORTED;
}

int sdemc_mc_cv(const sdemc_sde*, const sdemc_payoff*, float, float, const sdemc_mlp*, const sdemc_mlp*,
                const sdemc_range*, const sdemc_inject*, sdemc_moments*, float*, void*, void*) {
  return SDEMC_ERR_UNSUPPORTED;
}

}  // extern "C"
