// cv.cuh -- fused control-variate Monte Carlo (E5/E6/E7): simulate + evaluate the control-variate MLPs f and g
// along every path + accumulate gamma = payoff + sum f dW D + sum g D J - sum rate E[J] g D dt + moments, in ONE
// kernel.  Replaces mc_apply_cvs mc.py:195-242 -> simulate_adapted_data mc.py:391-398 (full trajectory storage,
// ~20 B per path-step through HBM) -> apply_adapted_control_variates varred.py:98-131 (two MLP forwards over
// bs*S rows), resp. the diffusion variant varred.py:75-95.
//
// Tensor cores (tcgen05, accumulators in TMEM): a CTA of 128 threads owns a tile of 128 paths; thread r is path r
// is row r of every activation matrix and lane r of the TMEM accumulators.  Each time step evaluates both nets
//   Linear(2,H) ReLU Linear(H,H) ReLU Linear(H,H) ReLU Linear(H,1)         (nets.py:39-93, BN-free, H <= 63)
// as three rounds of tcgen05.mma (M=128, N=64, K=16|64, bf16 in / fp32 accumulate) whose A operand the threads
// write themselves into shared memory in the canonical K-major no-swizzle UMMA layout, plus a SIMT dot product for
// the last layer.  Biases are folded into the contraction: every padded activation vector carries a constant 1 in
// slot 63 (W[n][63] = b[n], W[63][63] = 1).  The inputs (t, x) are split into bf16 hi + lo parts (two K slots each
// with the same weight) so the nets are evaluated at fp32-accurate inputs; weights and hidden activations are bf16.
// Any adapted f, g gives an unbiased estimator, so the reduced precision only perturbs the variance reduction.
#pragma once
#include <cuda_bf16.h>

#include "engine.cuh"
#include "jump.cuh"

namespace sdemc {

struct DevMlp {
  const float* w[4];
  const float* b[4];
  int in_dim, hidden, out_dim;
};

constexpr int kCvThreads = 128;
constexpr int kCvOne = 63;  // index of the constant-one unit in every padded (64-wide) activation vector

// shared-memory carve-up (bytes).  Operand tiles: 16-byte chunk c = k/8 of row r lives at c * (rows*16) + r * 16,
// i.e. UMMA descriptors with LBO = rows*16 (K direction) and SBO = 128 (next group of 8 rows).
constexpr int kCvW1Bytes = 64 * 16 * 2;
constexpr int kCvWBytes = 64 * 64 * 2;
constexpr int kCvABytes = 128 * 64 * 2;
constexpr int kCvOffW1F = 0;
constexpr int kCvOffW1G = kCvOffW1F + kCvW1Bytes;
constexpr int kCvOffW2F = kCvOffW1G + kCvW1Bytes;
constexpr int kCvOffW3F = kCvOffW2F + kCvWBytes;
constexpr int kCvOffW2G = kCvOffW3F + kCvWBytes;
constexpr int kCvOffW3G = kCvOffW2G + kCvWBytes;
constexpr int kCvOffAF = kCvOffW3G + kCvWBytes;
constexpr int kCvOffAG = kCvOffAF + kCvABytes;
constexpr int kCvOffW4F = kCvOffAG + kCvABytes;  // fp32[64]
constexpr int kCvOffW4G = kCvOffW4F + 256;
constexpr int kCvOffBar = kCvOffW4G + 256;
constexpr int kCvOffTmem = kCvOffBar + 8;
constexpr int kCvSmemBytes = kCvOffTmem + 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kCvIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(kCvIdesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
      "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// relu + round-to-nearest bf16 of two fp32 values in one instruction: low half <- lo, high half <- hi
__device__ __forceinline__ uint32_t relu_pack_bf16x2(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// epilogue of a hidden layer: this thread's 64 accumulator columns -> ReLU -> bf16 -> its row of the next A operand
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr, uint32_t a_row_addr) {
  uint32_t v[64];
  tmem_ld64(taddr, v);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    sts128(a_row_addr + c * (128 * 16), relu_pack_bf16x2(v[8 * c + 0], v[8 * c + 1]),
           relu_pack_bf16x2(v[8 * c + 2], v[8 * c + 3]), relu_pack_bf16x2(v[8 * c + 4], v[8 * c + 5]),
           relu_pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
  }
}
// last layer on the SIMT pipes: sum_k w4[k] relu(h3[k]); slot 63 carries the bias (h3[63] == 1)
__device__ __forceinline__ float head_epilogue(uint32_t taddr, const float* w4) {
  uint32_t v[64];
  tmem_ld64(taddr, v);
  float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float4 w = *reinterpret_cast<const float4*>(w4 + 4 * c);
    acc0 = fmaf(w.x, fmaxf(__uint_as_float(v[4 * c + 0]), 0.0f), acc0);
    acc1 = fmaf(w.y, fmaxf(__uint_as_float(v[4 * c + 1]), 0.0f), acc1);
    acc0 = fmaf(w.z, fmaxf(__uint_as_float(v[4 * c + 2]), 0.0f), acc0);
    acc1 = fmaf(w.w, fmaxf(__uint_as_float(v[4 * c + 3]), 0.0f), acc1);
  }
  return acc0 + acc1;
}
// first-layer A row: [t_hi, t_lo, x_hi, x_lo, 1, 0, 0, 0 | 0 x 8]
__device__ __forceinline__ void write_input_row(uint32_t a_row_addr, float t, float x) {
  const __nv_bfloat16 th = __float2bfloat16_rn(t), xh = __float2bfloat16_rn(x);
  const __nv_bfloat16 tl = __float2bfloat16_rn(t - __bfloat162float(th)), xl = __float2bfloat16_rn(x - __bfloat162float(xh));
  const uint32_t p0 = (uint32_t)__bfloat16_as_ushort(th) | ((uint32_t)__bfloat16_as_ushort(tl) << 16);
  const uint32_t p1 = (uint32_t)__bfloat16_as_ushort(xh) | ((uint32_t)__bfloat16_as_ushort(xl) << 16);
  sts128(a_row_addr, p0, p1, 0x00003f80u /* (1.0bf16, 0) */, 0u);
  sts128(a_row_addr + 128 * 16, 0u, 0u, 0u, 0u);
}

// weights -> bf16 canonical operand tiles with folded biases (see header comment)
__device__ __forceinline__ void load_mlp(const DevMlp& net, uint8_t* w1, uint8_t* w2, uint8_t* w3, float* w4) {
  const int H = net.hidden;
  for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
    const int n = idx >> 4, k = idx & 15;
    float v = 0.0f;
    if (n < H) {
      if (k < 2) v = net.w[0][n * 2 + 0];
      else if (k < 4) v = net.w[0][n * 2 + 1];
      else if (k == 4) v = net.b[0][n];
    } else if (n == kCvOne && k == 4) {
      v = 1.0f;
    }
    *reinterpret_cast<__nv_bfloat16*>(w1 + (k >> 3) * (64 * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  for (int l = 1; l <= 2; ++l) {
    uint8_t* dst = l == 1 ? w2 : w3;
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
      const int n = idx >> 6, k = idx & 63;
      float v = 0.0f;
      if (n < H) {
        if (k < H) v = net.w[l][n * H + k];
        else if (k == kCvOne) v = net.b[l][n];
      } else if (n == kCvOne && k == kCvOne) {
        v = 1.0f;
      }
      *reinterpret_cast<__nv_bfloat16*>(dst + (k >> 3) * (64 * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(v);
    }
  }
  for (int k = threadIdx.x; k < 64; k += blockDim.x) w4[k] = k < H ? net.w[3][k] : (k == kCvOne ? net.b[3][0] : 0.0f);
}

struct DevCv {
  float disc_rate_l2e;  // r * log2(e):  D(t) = 2^(-t r log2 e)   (ConstantShortRate options.py:334-337)
  float comp_c;         // - rate * E[J]                            (varred.py:126)
  int last_interval;    // compensator intervals with index >= last_interval are dropped (varred.py:104,126-127)
  float* gamma_out;     // (n) per-path gamma or nullptr
};

// 1-D 'diag' SDE (dim == 1, m == 1): Merton-type jump diffusion (JUMPS) or GBM-type diffusion (!JUMPS)
template <class C, bool JUMPS, bool INJECT>
__global__ void __launch_bounds__(kCvThreads) cv_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const DevMlp f,
                                                        const DevMlp g, const DevCv cv,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  extern __shared__ __align__(1024) uint8_t cv_smem[];
  constexpr int MARKS = C::MARKS;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(cv_smem);
  const uint32_t bar = sbase + kCvOffBar;
  const float* w4f = reinterpret_cast<const float*>(cv_smem + kCvOffW4F);
  const float* w4g = reinterpret_cast<const float*>(cv_smem + kCvOffW4G);

  load_mlp(f, cv_smem + kCvOffW1F, cv_smem + kCvOffW2F, cv_smem + kCvOffW3F, reinterpret_cast<float*>(cv_smem + kCvOffW4F));
  if (JUMPS) load_mlp(g, cv_smem + kCvOffW1G, cv_smem + kCvOffW2G, cv_smem + kCvOffW3G, reinterpret_cast<float*>(cv_smem + kCvOffW4G));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(sbase + kCvOffTmem));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(cv_smem + kCvOffTmem);
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  const uint32_t tacc_f = tlane, tacc_g = tlane + 64;           // f accumulators: columns 0-63, g: 64-127
  const uint32_t a_row_f = sbase + kCvOffAF + tid * 16, a_row_g = sbase + kCvOffAG + tid * 16;
  uint32_t phase = 0;

  // one round: A operands are in place -> sync -> MMAs -> wait for their completion
  auto mma_round = [&](int layer, int active_any_in, int* any_out) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my st.shared -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    const int any = __syncthreads_or(active_any_in);
    if (any_out) *any_out = any;
    if (!any) return;
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ksteps = layer == 0 ? 1 : 4;
      const uint32_t wf = sbase + (layer == 0 ? kCvOffW1F : layer == 1 ? kCvOffW2F : kCvOffW3F);
      const uint32_t wg = sbase + (layer == 0 ? kCvOffW1G : layer == 1 ? kCvOffW2G : kCvOffW3G);
      for (int ks = 0; ks < ksteps; ++ks) {
        umma_bf16(tmem, umma_desc(sbase + kCvOffAF + ks * 2 * (128 * 16), 128 * 16, 128),
                  umma_desc(wf + ks * 2 * (64 * 16), 64 * 16, 128), ks > 0);
        if (JUMPS)
          umma_bf16(tmem + 64, umma_desc(sbase + kCvOffAG + ks * 2 * (128 * 16), 128 * 16, 128),
                    umma_desc(wg + ks * 2 * (64 * 16), 64 * 16, 128), ks > 0);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  Accum acc;
  acc.zero();
  const uint64_t n_tiles = (rg.n_paths + kCvThreads - 1) / kCvThreads;
  const int n = s.num_steps;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t i = tile * kCvThreads + tid;
    const bool valid = i < rg.n_paths;
    const uint64_t gp = rg.path_lo + i;
    const uint32_t plo = (uint32_t)gp, phi = (uint32_t)(gp >> 32);
    float x[kMaxDim], xo[kMaxDim];
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) x[d] = d < 1 ? s.x0[d] : 0.0f;
    float t = 0.0f, h = s.h0, left = s.x0[0], Jprev = 0.0f, cvsum = 0.0f;
    bool need_pop = true;
    typename std::conditional<INJECT, InjectJumps<MARKS>, InlineJumps<MARKS>>::type src;
    if constexpr (JUMPS) {
      if constexpr (INJECT) src.init(s, inj, valid ? i : 0);
      else src.init(plo, phi);
    }
    const int kcap = JUMPS ? (INJECT ? inj.K : 4 * (n + s.max_jumps) + 64) : n;
    float zbuf[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int own_iters = 0;

    for (int k = 0;; ++k) {
      const bool active = valid && k < kcap && (JUMPS ? t < s.T : true);
      // ---- both nets at the state of index k --------------------------------------------------------------
      const float t_in = JUMPS ? t : (float)((double)s.T * (double)k / (double)n);  // partition(T, n, 'left')
      write_input_row(a_row_f, t_in, x[0]);
      if (JUMPS) write_input_row(a_row_g, t_in, left);
      int any = 0;
      mma_round(0, active ? 1 : 0, &any);
      if (!any) break;
      hidden_epilogue(tacc_f, a_row_f);
      if (JUMPS) hidden_epilogue(tacc_g, a_row_g);
      mma_round(1, 1, nullptr);
      hidden_epilogue(tacc_f, a_row_f);
      if (JUMPS) hidden_epilogue(tacc_g, a_row_g);
      mma_round(2, 1, nullptr);
      const float fval = head_epilogue(tacc_f, w4f);
      const float gval = JUMPS ? head_epilogue(tacc_g, w4g) : 0.0f;

      // ---- this thread's Brownian normal for iteration k (one Philox block serves 4 iterations) -----------
      if ((k & 3) == 0) {
        if constexpr (!INJECT) {
          uint32_t o[4];
          philox4x32_10((uint32_t)(k >> 2), STREAM_DIFFUSION, plo, phi, keys, o);
          box_muller(o[0], o[1], zbuf[0], zbuf[1]);
          box_muller(o[2], o[3], zbuf[2], zbuf[3]);
        }
      }
      float z;
      if constexpr (INJECT) {
        z = (valid && k < inj.K) ? inj.z[i * (uint64_t)inj.K + k] : 0.0f;
      } else {
        z = zbuf[0];
        zbuf[0] = zbuf[1]; zbuf[1] = zbuf[2]; zbuf[2] = zbuf[3];
      }

      // ---- advance the path by one iteration and accumulate the control variates -------------------------
      if (active) {
        const float D = fast_ex2(-t_in * cv.disc_rate_l2e);
        float dt, sq;
        float tau = 0.0f;
        if constexpr (JUMPS) {
          src.begin_iter(s, keys, k);
          src.advance(s, keys, need_pop);
          tau = src.tau;
          h = fminf(h, fmaxf(s.T - t, 0.0f));
          dt = fmaxf(fminf(h, tau - t), 0.0f);
          sq = fast_sqrt(dt);
        } else {
          dt = s.h0;
          sq = s.sqrt_h0;
        }
        const float dW = z * sq;
        float w1[kMaxDim] = {z, 0.0f, 0.0f, 0.0f}, w2[kMaxDim] = {0.0f, 0.0f, 0.0f, 0.0f};
        xo[0] = x[0];
        euler_step<C>(s, x, dt, sq, w1, w2);
        float c = fval * dW;                                     // f dW           (integrate_cv varred.py:202-214)
        if constexpr (JUMPS) {
          c = fmaf(gval, Jprev, c);                              // g J            (varred.py:124)
          if (k < cv.last_interval) c = fmaf(cv.comp_c * gval, dt, c);   // - rate E[J] g dt (varred.py:126-127)
          t += dt;
          left = x[0];
          const bool hit = fabsf(tau - t) <= fmaf(fabsf(t), 1e-5f, 1e-12f);
          const float Jc = hit ? src.mark(s, k) : 0.0f;
          if (s.exact_jumps) xo[0] = x[0];
          add_jump<C>(s, x, xo, Jc);
          Jprev = Jc;
          need_pop = hit;
        }
        cvsum = fmaf(c, D, cvsum);
        own_iters = k + 1;
      }
    }

    if (valid) {
      const float pay = eval_payoff<1>(po, x);
      const float gamma = pay + cvsum;
      if (cv.gamma_out) cv.gamma_out[i] = gamma;
      acc.add(gamma, pay, own_iters);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
