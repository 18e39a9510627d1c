// philox.cuh -- counter-based noise generated in registers (no increment tensor ever touches HBM).
//
// Replaces the reference's draws from torch's global RNG:
//   torch.randn            solvers.py:52, sde.py:326      -> Philox4x32-10 + Box-Muller
//   tensor.exponential_    solvers.py:144                 -> -log2(u) * ln2 / rate
//   torch.rand             levy.py:86                     -> 23-bit uniform
//
// Counter layout (128 bit)  : (block index, stream id, global path id lo, global path id hi)
// Key (64 bit)              : the solver seed; the 10 round keys are precomputed on the host and live in
//                             uniform registers, so a round is 2 IMAD.WIDE + 2 LOP3.
// Because the counter is the GLOBAL path id, results do not depend on grid shape, batch split or GPU count.
#pragma once
#include <cstdint>

namespace sdemc {

struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};

enum : uint32_t {
  STREAM_DIFFUSION = 0,   // Brownian unit normals, consumed in loop order
  STREAM_JUMP_QUEUE = 1,  // (gap, mark) pairs of the sparse-jump queue
  STREAM_JUMP_INLINE = 2, // per-iteration (gap, mark) candidates of the dense-jump strategy
  STREAM_PACKED = 3       // jump_flat.cuh, 1-D short paths: normals, gaps and marks of two iterations in one block
};

__host__ inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys K;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.k0[r] = a;
    K.k1[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return K;
}

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& K, uint32_t (&o)[4]) {
#ifndef SDEMC_EXP_PHILOX_ROUNDS   // (timing experiments only)
#define SDEMC_EXP_PHILOX_ROUNDS 10
#endif
#pragma unroll
  for (int r = 0; r < SDEMC_EXP_PHILOX_ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)c0 * 0xD2511F53u;
    const uint64_t p1 = (uint64_t)c2 * 0xCD9E8D57u;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// 23 random mantissa bits -> float in [1, 2).  No I2F (that would go to the XU pipe).
// One LOP3 ((u & mask) | one, LUT 0xEA) with the exponent pattern hoisted into a register by ptxas.
__device__ __forceinline__ float bits_to_12(uint32_t u) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x007fffff, 0x3f800000, 0xEA;" : "=r"(r) : "r"(u));
  return __uint_as_float(r);
}
// uniform on [0, 1) with 2^-23 spacing
__device__ __forceinline__ float bits_to_u01(uint32_t u) { return bits_to_12(u) - 1.0f; }
// uniform on (0, 1] with 2^-23 spacing (safe for log)
__device__ __forceinline__ float bits_to_u01_open0(uint32_t u) { return 2.0f - bits_to_12(u); }

__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Box-Muller on two 32-bit words: radius*(cos, sin).  `scale2` multiplies the squared radius, so callers can
// fold a constant factor (e.g. sigma^2 h) into the square root: r = sqrt(scale2 * (-2 ln u)).
// 4 MUFU (lg2, sqrt, sin, cos) + 6 FP32 per pair.
// polar form: radius r = sqrt(neg2ln2_scale2 * log2(u)) with neg2ln2_scale2 = -2 ln2 * scale^2, and (cos, sin)
__device__ __forceinline__ void box_muller_polar(uint32_t wa, uint32_t wb, float neg2ln2_scale2, float& r, float& c,
                                                 float& s) {
  r = fast_sqrt(fast_lg2(bits_to_u01_open0(wa)) * neg2ln2_scale2);
  const float ang = bits_to_12(wb) * 6.283185307179586f;  // [2pi, 4pi): same law as [0, 2pi)
  c = fast_cos(ang);
  s = fast_sin(ang);
}
__device__ __forceinline__ void box_muller_scaled(uint32_t wa, uint32_t wb, float scale2, float& n0, float& n1) {
  float r, c, s;
  box_muller_polar(wa, wb, -1.3862943611198906f * scale2, r, c, s);
  n0 = r * c;
  n1 = r * s;
}
__device__ __forceinline__ void box_muller(uint32_t wa, uint32_t wb, float& n0, float& n1) {
  box_muller_scaled(wa, wb, 1.0f, n0, n1);
}

// Six normals from ONE Philox block.  IMAD.WIDE runs at quarter rate on sm_100 (measured: 20 of them = 80 of the 83
// SMSP-cycles a Philox4x32-10 block costs), so the 128 bits are spent frugally: three Box-Muller pairs, each with a
// 23-bit radius uniform (low bits of o0, o1, o2: tail reach 5.6 sigma, variance deficit 2e-6) and a 16-bit angle
// (o3 low half, o3 high half, top bytes of o0 and o1: 65536 directions).  Polar form so callers can fold a scale
// into the radius: r = sqrt(neg2ln2_scale2 * log2 u), unit normals are r*c and r*s.
__device__ __forceinline__ float angle_bits_to_12(uint32_t k16_in_low_bits) {
  uint32_t r;
  asm("lop3.b32 %0, %1, 0x0000ffff, 0x3f800000, 0xEA;" : "=r"(r) : "r"(k16_in_low_bits));
  return __uint_as_float(r);
}
__device__ __forceinline__ void philox_polar3(const uint32_t (&o)[4], float neg2ln2_scale2, float (&r)[3],
                                              float (&c)[3], float (&s)[3]) {
  const float f0 = angle_bits_to_12(o[3]);
  const float f1 = __uint_as_float((o[3] >> 16) | 0x3f800000u);
  const float f2 = angle_bits_to_12(__byte_perm(o[0], o[1], 0x0073));  // bytes {o0[31:24], o1[31:24]}
  const float fa[3] = {f0, f1, f2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    r[j] = fast_sqrt(fast_lg2(bits_to_u01_open0(o[j])) * neg2ln2_scale2);
    // f in [1, 1 + 2^-7): (f - 1) * 128 * 2pi is the angle in [0, 2pi)
    const float ang = fmaf(fa[j], 804.247719318987f, -804.247719318987f);
    c[j] = fast_cos(ang);
    s[j] = fast_sin(ang);
  }
}
// same, without the square root: r2[j] = neg2ln2_scale2 * log2(u_j) is the SQUARED radius, for callers that fold a
// per-use factor into it before taking the root (sqrt(r2 * dt) = radius * sqrt(dt): one MUFU instead of two)
__device__ __forceinline__ void philox_polar3_sq(const uint32_t (&o)[4], float neg2ln2_scale2, float (&r2)[3],
                                                 float (&c)[3], float (&s)[3]) {
  const float f0 = angle_bits_to_12(o[3]);
  const float f1 = __uint_as_float((o[3] >> 16) | 0x3f800000u);
  const float f2 = angle_bits_to_12(__byte_perm(o[0], o[1], 0x0073));
  const float fa[3] = {f0, f1, f2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    r2[j] = fast_lg2(bits_to_u01_open0(o[j])) * neg2ln2_scale2;
    const float ang = fmaf(fa[j], 804.247719318987f, -804.247719318987f);
    c[j] = fast_cos(ang);
    s[j] = fast_sin(ang);
  }
}
// six unit normals in slot order (r0 c0, r0 s0, r1 c1, r1 s1, r2 c2, r2 s2)
__device__ __forceinline__ void philox_normals6(const uint32_t (&o)[4], float* n) {
  float r[3], c[3], s[3];
  philox_polar3(o, -1.3862943611198906f, r, c, s);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    n[2 * j] = r[j] * c[j];
    n[2 * j + 1] = r[j] * s[j];
  }
}
constexpr int kNormalsPerBlock = 6;

// Exp(1) draw from one word: -ln(u), u in (0,1]
__device__ __forceinline__ float exp1_from_bits(uint32_t w) { return fast_lg2(bits_to_u01_open0(w)) * -0.6931471805599453f; }

}  // namespace sdemc
