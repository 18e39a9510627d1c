/*
 * sde_oracle.c -- TEST INFRASTRUCTURE (see sde_oracle.h).  CPU restatement of the hot path of
 * piers-hinds/sde_mc; never linked into, imported by or executed from the product path.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -shared)
 */
#include "sde_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- fp32 instantiation: the reference's default dtype ------------------------------------------------ */
#define REAL float
#define SUF(name) name##_f32
#define RSQRT sqrtf
#define REXP expf
#define RLOG logf
#define RPOW powf
#define RFABS fabsf
#define RFMIN fminf
#define RFMAX fmaxf
#define CLAMP_DT 1 /* only differs from the reference where its fp32 run would hit the assert solvers.py:193,264 */
#include "sde_oracle_impl.inc"
#undef REAL
#undef SUF
#undef RSQRT
#undef REXP
#undef RLOG
#undef RPOW
#undef RFABS
#undef RFMIN
#undef RFMAX
#undef CLAMP_DT

/* ---- fp64 instantiation: torch.set_default_dtype(float64) run of the reference (jump MLMC) ------------- */
#define REAL double
#define SUF(name) name##_f64
#define RSQRT sqrt
#define REXP exp
#define RLOG log
#define RPOW pow
#define RFABS fabs
#define RFMIN fmin
#define RFMAX fmax
#define CLAMP_DT 0
#include "sde_oracle_impl.inc"
#undef REAL
#undef SUF

/* DiffusionSolver.multilevel_solve solvers.py:90-119 (fp32).  Quirk kept: the coarse step is taken with the
 * time AFTER the fine sub-steps (:114-116) -- irrelevant for the time-homogeneous built-in models. */
void oracle_diffusion_pair_f32(const oracle_sde* s, int64_t n, int fine, int coarse, const float* z,
                               float* paths_fine, float* paths_coarse) {
  ctx_f32 c; make_ctx_f32(s, &c);
  const int d = c.dim, m = c.m, factor = fine / coarse;
  const float hf = (float)(s->T / (double)fine), hc = (float)factor * hf, sq = sqrtf(hf);
  for (int64_t p = 0; p < n; ++p) {
    float xf[4], xc[4], xn[4];
    for (int i = 0; i < d; ++i) {
      xf[i] = xc[i] = c.x0[i];
      paths_fine[(p * (fine + 1)) * d + i] = xf[i];
      paths_coarse[(p * (coarse + 1)) * d + i] = xc[i];
    }
    for (int k = 0; k < coarse; ++k) {
      float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
      for (int q = 0; q < factor; ++q) {
        const int step = k * factor + q;
        float w[2][4];
        for (int j = 0; j < m; ++j) {
          float nz[4];
          for (int i = 0; i < d; ++i) nz[i] = z[((p * fine + step) * d + i) * m + j] * sq;
          correlate_f32(&c, nz, w[j]);
        }
        if (c.heston) heston_step_f32(&c, xf, hf, w[0], xn);   /* HestonSolver: schemes.py:16-22 */
        else euler_f32(&c, xf, hf, w[0], w[1], xn);
        for (int i = 0; i < d; ++i) {
          xf[i] = xn[i];
          paths_fine[(p * (fine + 1) + step + 1) * d + i] = xf[i];
          s1[i] += w[0][i];
          if (m == 2) s2[i] += w[1][i];
        }
      }
      if (c.heston) heston_step_f32(&c, xc, hc, s1, xn);
      else euler_f32(&c, xc, hc, s1, s2, xn);
      for (int i = 0; i < d; ++i) { xc[i] = xn[i]; paths_coarse[(p * (coarse + 1) + k + 1) * d + i] = xc[i]; }
    }
  }
}

void oracle_icdf_f32(const double* mark_p, int64_t n, const float* u, float* out) {
  oracle_sde s; memset(&s, 0, sizeof s);
  s.marks = ORACLE_MARKS_ICDF; s.num_steps = 1; s.T = 1.0; s.dim = 1; s.m = 1;
  memcpy(s.mark_p, mark_p, sizeof s.mark_p);
  ctx_f32 c; make_ctx_f32(&s, &c);
  for (int64_t i = 0; i < n; ++i) out[i] = icdf_f32(&c, u[i]);
}

/* ---- control variates ---------------------------------------------------------------------------------- */
/* Mlp.forward nets.py:92-93 for the BN-free net Linear/ReLU x3 + Linear (nets.py:73-88) */
static void mlp_forward(const oracle_mlp* net, const float* in, float* out) {
  float a[256], b[256];
  const int H = net->hidden;
  for (int o = 0; o < H; ++o) {
    float acc = net->b[0][o];
    for (int i = 0; i < net->in_dim; ++i) acc += net->w[0][o * net->in_dim + i] * in[i];
    a[o] = acc > 0.f ? acc : 0.f;
  }
  for (int l = 1; l < net->n_hidden_layers; ++l) {
    for (int o = 0; o < H; ++o) {
      float acc = net->b[l][o];
      for (int i = 0; i < H; ++i) acc += net->w[l][o * H + i] * a[i];
      b[o] = acc > 0.f ? acc : 0.f;
    }
    memcpy(a, b, sizeof(float) * H);
  }
  const int L = net->n_hidden_layers;
  for (int o = 0; o < net->out_dim; ++o) {
    float acc = net->b[L][o];
    for (int i = 0; i < H; ++i) acc += net->w[L][o * H + i] * a[i];
    out[o] = acc;
  }
}

/* apply_adapted_control_variates varred.py:98-131 with the AdaptedPathData truncation nets.py:192-200
 * (the datasets drop the last time index) and integrate_cv varred.py:202-214: the Brownian sum keeps the first
 * brownian_steps = remove_steps(tol, total_steps, T) steps (helpers.py:71-74; <= 0: all, tol = 0). */
void oracle_cv_gamma_jump_f32(const oracle_sde* s, int64_t n, int K, int total_steps, int brownian_steps, double disc_rate,
                              double jump_mean, const oracle_mlp* f, const oracle_mlp* g, const float* paths,
                              const float* left, const float* times, const float* jumps, const float* normals,
                              const float* payoffs, float* gamma) {
  const int d = s->dim, m = s->m, S = total_steps;
  const float r = (float)disc_rate;
  const float comp_c = (float)(-(double)(float)s->rate * jump_mean); /* - rate * jump_mean varred.py:126 */
  for (int64_t p = 0; p < n; ++p) {
    double bcv = 0.0, jcv = 0.0, comp = 0.0;
    for (int k = 0; k < S; ++k) {
      const float t = times[p * (K + 1) + k];
      const float D = expf(-t * r);                                   /* options.py:334-337 */
      float in[5], fo[8], go[4];
      in[0] = t;
      for (int i = 0; i < d; ++i) in[1 + i] = paths[(p * (K + 1) + k) * d + i];
      mlp_forward(f, in, fo);
      float acc = 0.f;
      for (int i = 0; i < d * m; ++i) acc += normals[(p * K + k) * d * m + i] * fo[i];
      if (brownian_steps <= 0 || k < brownian_steps) bcv += (double)(acc * D);          /* varred.py:203-209 */
      if (g) {
        for (int i = 0; i < d; ++i) in[1 + i] = left[(p * (K + 1) + k) * d + i];
        mlp_forward(g, in, go);
        for (int i = 0; i < d; ++i) jcv += (double)(go[i] * D * jumps[(p * (K + 1) + k) * d + i]);
        if (k < S - 1) {
          const float h = times[p * (K + 1) + k + 1] - t;             /* torch.diff(time_paths) :104 */
          for (int i = 0; i < d; ++i) comp += (double)(comp_c * go[i] * D * h);
        }
      }
    }
    gamma[p] = (float)((double)payoffs[p] + bcv + jcv + comp);
  }
}

/* apply_diffusion_control_variate varred.py:75-95: time points partition(T, steps, 'left') helpers.py:6-33; the sum
 * keeps the first brownian_steps = remove_steps(tol, num_steps, T) steps (<= 0: all) */
void oracle_cv_gamma_diffusion_f32(const oracle_sde* s, int64_t n, int brownian_steps, double disc_rate, const oracle_mlp* f,
                                   const float* paths, const float* normals, const float* payoffs, float* gamma) {
  const int d = s->dim, m = s->m, S = s->num_steps;
  const float r = (float)disc_rate;
  for (int64_t p = 0; p < n; ++p) {
    double bcv = 0.0;
    for (int k = 0; k < S; ++k) {
      const float t = (float)(s->T * (double)k / (double)S);
      const float D = expf(-t * r);
      float in[5], fo[8];
      in[0] = t;
      for (int i = 0; i < d; ++i) in[1 + i] = paths[(p * (S + 1) + k) * d + i];
      mlp_forward(f, in, fo);
      float acc = 0.f;
      for (int i = 0; i < d * m; ++i) acc += normals[(p * S + k) * d * m + i] * fo[i];
      if (brownian_steps <= 0 || k < brownian_steps) bcv += (double)(acc * D);
    }
    gamma[p] = (float)((double)payoffs[p] + bcv);
  }
}

/* ---- Philox4x32-10 (Salmon et al. 2011), as in sde_mc_b200/csrc/philox.cuh -------------------------------- */
void oracle_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)c0 * 0xD2511F53u, p1 = (uint64_t)c2 * 0xCD9E8D57u;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
