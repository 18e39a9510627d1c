// launch.cuh -- entry points of the per-solver translation units (internal, C++)
#pragma once
#include "host_common.cuh"

namespace sdemc {

struct LaunchArgs {
  DevSde sde;
  DevPayoff payoff;
  DevRange range;
  PhiloxKeys keys;
  DevInject inject;
  DevOut out;
  bool use_inject;
  bool store;
  int qdepth;          // jump queue depth (multiple of 4), 0 => inline jump strategy
  int short_path = SDEMC_SHORT_AUTO;  // sdemc_short_path as resolved by the entry point (AUTO: see launch_jump.cu)
  bool prefer_packed = false;         // AUTO resolves to PACKED instead of ALIGNED (single-level call of sdemc_mlmc_pair)
  bool no_tma = false;                // path-storing: SDEMC_OUT_NO_TMA
  double* d_moments;
  void* d_ws;
  cudaStream_t stream;
};

int launch_diffusion(const sdemc_sde& s, const LaunchArgs& a);
int launch_jump(const sdemc_sde& s, const LaunchArgs& a);
int launch_jump_store(const sdemc_sde& s, const LaunchArgs& a);  // launch_jump's storing half (launch_jump_store.cu)
int launch_pair(const sdemc_sde& s, const LaunchArgs& a, int fine, int coarse, float* d_terminal);
int launch_debug_draws(const sdemc_sde& s, const DevRange& rg, const PhiloxKeys& keys, int kind, int count, float* a,
                       float* b, float* c, cudaStream_t stream);
int launch_pair_f64(const sdemc_sde& s, const sdemc_coeffs_f64& co, const sdemc_payoff* payoff, int fine, int coarse,
                    const DevRange& rg, const PhiloxKeys& keys, const sdemc_inject_f64* inject, double* d_moments,
                    double* d_terminal, void* d_ws, cudaStream_t stream);
struct DevMlp;
struct DevCv;
int launch_cv(const sdemc_sde& s, const LaunchArgs& a, const DevMlp& f, const DevMlp& g, const DevCv& cv);

}  // namespace sdemc
