// host_common.cuh -- host-side helpers shared by the translation units of libsdemc_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>

#include "engine.cuh"

namespace sdemc {

constexpr uint64_t kWorkspaceBytes = 1u << 20;  // ticket (64 B) + up to 16383 per-CTA partials

void set_cuda_error(cudaError_t e, const char* where);

#define SDEMC_CUDA_CHECK(expr)                       \
  do {                                               \
    cudaError_t e__ = (expr);                        \
    if (e__ != cudaSuccess) {                        \
      ::sdemc::set_cuda_error(e__, #expr);           \
      return SDEMC_ERR_CUDA;                         \
    }                                                \
  } while (0)

// struct_size handshake (include/sdemc_b200.h): a struct laid out by another header is rejected, not read
template <class T>
inline bool sized(const T* p) { return p != nullptr && p->struct_size == (uint32_t)sizeof(T); }
template <class T>
inline bool sized_or_null(const T* p) { return p == nullptr || p->struct_size == (uint32_t)sizeof(T); }

inline bool valid_sde(const sdemc_sde* s) {
  if (!sized(s)) return false;
  if (s->dim < 1 || s->dim > SDEMC_MAX_DIM) return false;
  if (s->m < 1 || s->m > 2) return false;
  if (s->num_steps < 1) return false;
  if (!(s->T > 0.0f)) return false;
  if (s->marks != SDEMC_MARKS_NONE && !(s->rate > 0.0f)) return false;
  if (s->asian && s->dim < 2) return false;
  if (s->scheme == SDEMC_SCHEME_MILSTEIN && (s->m != 1 || s->family == SDEMC_FAMILY_HESTON)) return false;
  if (s->queue_depth < 0 || s->queue_depth > 64 || (s->queue_depth & 3)) return false;
  if (s->short_path < SDEMC_SHORT_AUTO || s->short_path > SDEMC_SHORT_PACKED_GENERIC) return false;
  return true;
}

// sdemc_sde -> device constants.  h0 follows solvers.py:70,166 (torch.tensor(T / num_steps) in fp32).
inline DevSde to_dev(const sdemc_sde& s, int num_steps) {
  DevSde d;
  std::memset(&d, 0, sizeof d);
  d.num_steps = num_steps;
  d.max_jumps = s.max_jumps;
  d.exact_jumps = s.exact_jumps;
  d.milstein = s.scheme == SDEMC_SCHEME_MILSTEIN ? 1 : 0;
  d.T = s.T;
  d.h0 = (float)((double)s.T / (double)num_steps);
  d.sqrt_h0 = sqrtf(d.h0);
  for (int i = 0; i < kMaxDim; ++i) {
    d.x0[i] = s.x0[i];
    d.a[i] = s.a[i];
    d.b1[i] = s.b1[i];
    d.b2[i] = s.b2[i];
    d.c[i] = s.c[i];
    d.ah[i] = s.a[i] * d.h0;
    d.b1s[i] = s.b1[i] * d.sqrt_h0;
    d.b2s[i] = s.b2[i] * d.sqrt_h0;
  }
  d.neg2ln2_b1s2 = -1.3862943611198906f * d.b1s[0] * d.b1s[0];
  for (int i = 0; i < kMaxDim * kMaxDim; ++i) d.chol[i] = s.chol[i];
  d.rate = s.rate;
  d.inv_rate = s.rate > 0.0f ? 1.0f / s.rate : 0.0f;
  const double log2e = 1.4426950408889634, ln2 = 0.6931471805599453;
  if (s.marks == SDEMC_MARKS_LOGNORMAL) {
    d.ln_a2 = (float)((double)s.mark_p[0] * log2e);
    d.ln_g2 = (float)((double)s.mark_p[1] * log2e);
  } else if (s.marks == SDEMC_MARKS_ICDF) {
    const double cm = s.mark_p[0], cp = s.mark_p[1], mu = s.mark_p[2], al = s.mark_p[3], eps = s.mark_p[4],
                 lda = s.mark_p[5];
    d.ic_y1 = s.mark_p[6];
    d.ic_y2 = s.mark_p[7];
    d.ic_y3 = s.mark_p[8];
    d.ic_mulda_cm = (float)(mu * lda / cm);
    d.ic_inv_mu = (float)(1.0 / mu);
    d.ic_inv_mu_ln2 = (float)(ln2 / mu);
    d.ic_alpha = (float)al;
    d.ic_lda_cm = (float)(lda / cm);
    d.ic_neg_inv_alpha = (float)(-1.0 / al);
    d.ic_malpha_cp = (float)(-al / cp);
    d.ic_lda = (float)lda;
    d.ic_x3_off = (float)(cm / mu + cm * ((std::pow(eps, -al) - 1.0) / al));
    d.ic_eps_ma = (float)std::pow(eps, -al);
    d.ic_mulda_cp = (float)(mu * lda / cp);
    d.ic_tol = (float)(5.960464477539063e-08 / 3.0);
  }
  for (int i = 0; i < 16; ++i) d.user_p[i] = s.user_p[i];
  d.hes_r = s.heston[0];
  d.hes_kappa = s.heston[1];
  d.hes_xi = s.heston[3];
  d.hes_kth = (float)((double)s.heston[1] * (double)s.heston[2]);
  d.hes_halfxi2 = (float)(0.5 * (double)s.heston[3] * (double)s.heston[3]);
  return d;
}

inline DevPayoff to_dev(const sdemc_payoff* p) {
  DevPayoff d;
  if (p) {
    d.kind = p->kind; d.log = p->log; d.index_mode = p->index_mode;
    d.strike = p->strike; d.tdisc = p->transform_discount; d.aux = p->aux; d.df = p->df;
  } else {
    d.kind = SDEMC_PAYOFF_EURO_CALL; d.log = 0; d.index_mode = SDEMC_INDEX_ADAPTED;
    d.strike = 0.0f; d.tdisc = 1.0f; d.aux = 1.0f; d.df = 1.0f;
  }
  return d;
}

// sdemc_range -> DevRange.  With a device-resident range the host field n_paths is only an upper bound used to size
// the grid (0 = unknown: a full persistent grid).
inline DevRange to_dev(const sdemc_range& r) {
  DevRange d;
  d.path_lo = r.path_lo;
  d.n_paths = r.n_paths;
  d.dyn = reinterpret_cast<const uint64_t*>(r.d_range);
  d.count_on_host = d.dyn != nullptr && (r.flags & SDEMC_RANGE_COUNT_ON_HOST) != 0 && r.n_paths > 0;
  if (d.dyn && d.n_paths == 0) d.n_paths = ~0ull >> 1;
  return d;
}

inline DevInject to_dev(const sdemc_inject* j) {
  DevInject d;
  std::memset(&d, 0, sizeof d);
  if (j) { d.z = j->d_z; d.zc = j->d_zc; d.jump_times = j->d_jump_times; d.marks = j->d_marks; d.K = j->K; }
  return d;
}

// normals_per_step: floats of the increments array per iteration (dim * m; dim for the 'diag' jump solver)
inline DevOut to_dev(const sdemc_paths_out* o, int S, int dim = 1, int normals_per_step = 1) {
  DevOut d;
  std::memset(&d, 0, sizeof d);
  d.S = S;
  d.pitch_state = (uint64_t)(S + 1) * dim;
  d.pitch_times = (uint64_t)(S + 1);
  d.pitch_normals = (uint64_t)S * normals_per_step;
  if (o) {
    d.paths = o->d_paths; d.left = o->d_left; d.times = o->d_times; d.jumps = o->d_jumps;
    d.normals = o->d_normals; d.payoffs = o->d_payoffs; d.iters = o->d_iters; d.terminal = o->d_terminal; d.total_steps = o->d_total_steps;
    if (o->pitch_state > 0) d.pitch_state = (uint64_t)o->pitch_state;
    if (o->pitch_times > 0) d.pitch_times = (uint64_t)o->pitch_times;
    if (o->pitch_normals > 0) d.pitch_normals = (uint64_t)o->pitch_normals;
  }
  return d;
}
inline bool valid_pitches(const sdemc_paths_out* o, int S, int dim, int normals_per_step) {
  if (!o) return true;
  if (o->pitch_state != 0 && o->pitch_state < (int64_t)(S + 1) * dim) return false;
  if (o->pitch_times != 0 && o->pitch_times < (int64_t)(S + 1)) return false;
  if (o->pitch_normals != 0 && o->pitch_normals < (int64_t)S * normals_per_step) return false;
  return true;
}

// persistent-style grid: a whole number of waves of resident CTAs, never more threads than paths
template <class Kernel>
inline int pick_grid(Kernel kernel, size_t dyn_smem, uint64_t n_paths, int* out_grid, int block = kBlock) {
  int dev = 0, sms = 0, per_sm = 0;
  SDEMC_CUDA_CHECK(cudaGetDevice(&dev));
  SDEMC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SDEMC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, dyn_smem));
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = (uint64_t)sms * (uint64_t)per_sm;
  const uint64_t need = (n_paths + block - 1) / block;
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  const uint64_t cap = (kWorkspaceBytes - 64) / (kNumMoments * sizeof(double));
  if (grid > cap) grid = cap;
  *out_grid = (int)grid;
  return SDEMC_OK;
}

}  // namespace sdemc
