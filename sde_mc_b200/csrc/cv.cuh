// cv.cuh -- fused control-variate Monte Carlo (E5/E6/E7): simulate + evaluate the control-variate MLPs f and g
// along every path + accumulate gamma = payoff + sum f dW D + sum g D J - sum rate E[J] g D dt + moments, in ONE
// kernel.  Replaces mc_apply_cvs mc.py:195-242 -> simulate_adapted_data mc.py:391-398 (full trajectory storage,
// ~20 B per path-step through HBM) -> apply_adapted_control_variates varred.py:98-131 (two MLP forwards over
// bs*S rows), resp. the diffusion variant varred.py:75-95.
//
// Tensor cores (tcgen05, accumulators AND the activation operand in TMEM): path r of a tile of 128 paths is lane r of
// the tile's 128 TMEM columns -- 64 fp32 accumulator columns D shared by both nets, 32 columns A_f and 32 columns
// A_g holding the nets' current activation vectors as bf16 pairs (column c = units 2c, 2c+1).  Each time step evaluates
//   Linear(2,H) ReLU Linear(H,H) ReLU Linear(H,H) ReLU Linear(H,1)         (nets.py:39-93, BN-free, H <= 63)
// for both nets as SEVEN phases per tile, f and g alternating on D: (layer 1, f) (layer 1, g) (layer 2, f) (layer 2, g)
// (layer 3, f) (layer 3, g) and the two heads together (D columns 0-15 and 16-31); a phase is one `tcgen05.mma.cta_group::1.kind::f16 [D], [A_net], b_desc` chain (M=128; N=64,K=16 | N=64,K=64 | N=64,K=64 |
// N=16,K=64; bf16 in, fp32 accumulate) with the weights (B) in shared memory.  The four worker warps of a tile (one
// per TMEM lane quadrant) read D with tcgen05.ld, release it, and while the tensor core already runs the OTHER
// net's phase into D they apply ReLU, pack to bf16 and write the next activation vector back with tcgen05.st.
// Why the activations live in TMEM: with A in shared memory (rounds of the first version of this kernel) every
// K=16 MMA pulls 4 KB of A and 2 KB of B through the tensor core's shared-memory port at 128 B/clk -- 48 clk for an
// instruction whose math takes 32 -- and the epilogues write the same bytes through the LSU: ncu showed
// l1tex__data_pipe_tc_wavefronts_mem_shared at 65 % and sm__pipe_tc_cycles_active at 79 % with the tensor math only
// 36 % busy (profiles/r01_ncu_merton_cv_smemA.*).  From TMEM the A operand costs no shared-memory bandwidth at all, the
// generic->async proxy fence of every round disappears, and sharing D between the nets keeps a tile at 128 columns,
// so four tiles (2 CTAs x 2) are still resident per SM to hide each other's MMA / commit / wake-up latency.
// Every tile is an independent chain with its own issuer warp; there is no CTA barrier in the step loop.
// Biases are folded into the contraction: every padded activation vector carries a constant 1 in slot 63
// (W[n][63] = b[n]); for H <= 56 the epilogue writes that constant itself and reads only 56 accumulator columns
// (TMEM reads are the tightest floor of a step), otherwise W[63][63] = 1 carries it through the MMA.
// The inputs (t, x) are split into bf16 hi + lo parts (two K slots each with the same weight) so the nets are
// evaluated at fp32-accurate inputs; weights and hidden activations are bf16.
// Any adapted f, g gives an unbiased estimator, so the reduced precision only perturbs the variance reduction.
#pragma once
#include <cuda_bf16.h>
#include <cstdio>

#include "engine.cuh"
#include "jump.cuh"

namespace sdemc {

struct DevMlp {
  const float* w[4];
  const float* b[4];
  int in_dim, hidden, out_dim;
};

constexpr int kCvRows = 128;      // paths per tile = TMEM lanes
constexpr int kCvTiles = 2;       // path tiles per CTA, independent chains
constexpr int kCvWorkerWarps = 4 * kCvTiles;  // warps 4*tl .. 4*tl+3: paths and epilogues of tile tl
constexpr int kCvIssuerWarps = kCvTiles;      // warp kCvWorkerWarps + tl issues the MMAs of tile tl (one thread)
constexpr int kCvThreads = (kCvWorkerWarps + kCvIssuerWarps) * 32;
constexpr int kCvOne = 63;  // index of the constant-one unit in every padded (64-wide) activation vector

// TMEM columns of a tile
constexpr int kCvColD = 0;      // 64 fp32 accumulators (the head uses the first 16)
constexpr int kCvColAf = 64;    // activation vector of f: 32 columns = 64 bf16
constexpr int kCvColAg = 96;    // activation vector of g
constexpr int kCvTileCols = 128;
constexpr int kCvTmemCols = kCvTileCols * kCvTiles;

// shared-memory carve-up (bytes).  Weight (B operand) tiles in the canonical K-major no-swizzle UMMA layout: 16-byte
// chunk c = k/8 of row n lives at c * (rows*16) + n * 16, i.e. descriptors with LBO = rows*16 (K direction) and
// SBO = 128 (next group of 8 rows).
constexpr int kCvHeadN = 16;                // N of the head MMA (smallest N for M = 128); only column 0 is used
constexpr int kCvW1Bytes = 64 * 16 * 2;
constexpr int kCvWBytes = 64 * 64 * 2;
constexpr int kCvW4Bytes = kCvHeadN * 64 * 2;
constexpr int kCvOffW1F = 0;
constexpr int kCvOffW1G = kCvOffW1F + kCvW1Bytes;
constexpr int kCvOffW2F = kCvOffW1G + kCvW1Bytes;
constexpr int kCvOffW3F = kCvOffW2F + kCvWBytes;
constexpr int kCvOffW2G = kCvOffW3F + kCvWBytes;
constexpr int kCvOffW3G = kCvOffW2G + kCvWBytes;
constexpr int kCvOffW4F = kCvOffW3G + kCvWBytes;
constexpr int kCvOffW4G = kCvOffW4F + kCvW4Bytes;
constexpr int kCvOffBar = kCvOffW4G + kCvW4Bytes;
constexpr int kCvOffTmem = kCvOffBar + 16 * kCvTiles;  // full[tile] then done[tile]
constexpr int kCvOffFlags = kCvOffTmem + 8;            // int any_active[tile][parity], int live[tile]
constexpr int kCvSmemBytes = kCvOffFlags + 4 * 3 * kCvTiles + 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t cv_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem_d] (+)= A[tmem_a] * B[db]^T : A operand in TMEM (lane = row, 32-bit column c = bf16 pair k = 2c, 2c+1; layout
// verified by tools/umma_ts_f16_test.cu), B operand in shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate));
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
// relu + round-to-nearest bf16 of two fp32 values in one instruction: low half <- lo, high half <- hi
__device__ __forceinline__ uint32_t relu_pack_bf16x2(uint32_t lo_bits, uint32_t hi_bits) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return r;
}
// tcgen05.ld / tcgen05.st of this warp's 32 lanes x N consecutive 32-bit columns (asynchronous: tmem_wait_ld / _st)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// N consecutive columns (N = 1, 2 or 4) of this warp's lanes: the head outputs of a net
template <int N>
__device__ __forceinline__ void tmem_ld_small(uint32_t taddr, uint32_t (&v)[4]) {
  static_assert(N == 1 || N == 2 || N == 4, "head widths");
  if constexpr (N == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr));
  else if constexpr (N == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
  else asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// The accumulator row of a hidden layer, read out of D.  NARROW (H <= 56): only columns 0-55 (32 + 16 + 8).
template <bool NARROW>
struct CvAccRow {
  uint32_t lo[32];
  uint32_t mid[16];
  uint32_t hi[NARROW ? 8 : 16];
  __device__ __forceinline__ void load(uint32_t tacc) {
    tmem_ld32(tacc, lo);
    tmem_ld16(tacc + 32, mid);
    if constexpr (NARROW) tmem_ld8(tacc + 48, hi);
    else tmem_ld16(tacc + 48, hi);
    tmem_wait_ld();
  }
  // ReLU -> bf16 pairs -> the 32 columns of the next activation vector; slot 63 = the constant 1 (NARROW: written here,
  // otherwise computed by the MMA through W[63][63] = 1)
  __device__ __forceinline__ void store_relu(uint32_t ta) const {
    uint32_t a[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) a[c] = relu_pack_bf16x2(lo[2 * c], lo[2 * c + 1]);
    tmem_st16(ta, a);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = relu_pack_bf16x2(mid[2 * c], mid[2 * c + 1]);
    if constexpr (NARROW) {
#pragma unroll
      for (int c = 0; c < 4; ++c) a[8 + c] = relu_pack_bf16x2(hi[2 * c], hi[2 * c + 1]);
      a[12] = a[13] = a[14] = 0u;
      a[15] = 0x3f800000u;  // (unit 62 = 0, unit 63 = 1.0bf16)
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) a[8 + c] = relu_pack_bf16x2(hi[2 * c], hi[2 * c + 1]);
    }
    tmem_st16(ta + 16, a);
  }
};
// bf16 (hi, lo) split of an fp32 input: hi + lo carries 16 mantissa bits, both K slots use the same weight
__device__ __forceinline__ uint32_t split_bf16(float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  return (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
}
// first-layer activation vector (K = 16 -> 8 columns): [t_hi, t_lo | x0_hi, x0_lo | ... | 1, 0 | 0 ...]; NX = state
// dimension (inputs (t, x_0 .. x_{NX-1}), nets.py:39-93 with input_dim = dim + 1), at most 6
template <int NX>
__device__ __forceinline__ void write_input_row(uint32_t ta, float t, const float (&x)[kMaxDim]) {
  static_assert(NX >= 1 && NX <= 6, "first-layer K = 16 holds (t, x) as hi/lo pairs plus the bias slot");
  uint32_t a[8];
  a[0] = split_bf16(t);
#pragma unroll
  for (int j = 0; j < NX; ++j) a[1 + j] = split_bf16(x[j]);
  a[1 + NX] = 0x00003f80u;  // (1.0bf16, 0): the bias slot
#pragma unroll
  for (int c = 2 + NX; c < 8; ++c) a[c] = 0u;
  tmem_st8(ta, a);
}

// weights -> bf16 canonical operand tiles with folded biases (see header comment)
__device__ __forceinline__ void load_mlp(const DevMlp& net, uint8_t* w1, uint8_t* w2, uint8_t* w3, uint8_t* w4) {
  const int H = net.hidden;
  for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
    const int n = idx >> 4, k = idx & 15;
    float v = 0.0f;
    const int nin = net.in_dim;  // inputs (t, x_0, ...): slots 2j, 2j+1 = hi, lo of input j; slot 2 nin = bias
    if (n < H) {
      if (k < 2 * nin) v = net.w[0][n * nin + (k >> 1)];
      else if (k == 2 * nin) v = net.b[0][n];
    } else if (n == kCvOne && k == 2 * nin) {
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
  // head: B operand of kCvHeadN rows, rows 0 .. out_dim-1 = (w4 row, bias in the constant-one slot), the rest zero
  for (int idx = threadIdx.x; idx < kCvHeadN * 64; idx += blockDim.x) {
    const int n = idx >> 6, k = idx & 63;
    float v = 0.0f;
    if (n < net.out_dim) v = k < H ? net.w[3][n * H + k] : (k == kCvOne ? net.b[3][n] : 0.0f);
    *reinterpret_cast<__nv_bfloat16*>(w4 + (k >> 3) * (kCvHeadN * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
}

struct DevCv {
  float disc_rate_l2e;  // r * log2(e):  D(t) = 2^(-t r log2 e)   (ConstantShortRate options.py:334-337)
  float comp_c;         // - rate * E[J]                            (varred.py:126)
  int last_interval;    // compensator intervals with index >= last_interval are dropped (varred.py:104,126-127)
  int brownian_steps;   // f dW terms of iterations >= brownian_steps are dropped (integrate_cv tol, varred.py:203-209)
  float* gamma_out;     // (n) per-path gamma or nullptr
};

// ---- mbarrier helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Model shapes (launch_cv.cu): 1-D 'diag' SDEs (dim == 1, m == 1: Merton-type jump diffusion or GBM-type
// diffusion, nets Mlp(2, [H, H, H], 1): merton_cv_experiment.py:37-38) and the 2-D 'indep' exp-Levy SDE of
// levy_rainbow_cv_experiment.py:39-40 (dim == 2, m == 2: f = Mlp(3, [H, H, H], 4), g = Mlp(3, [H, H, H], 2)).  In
// general f has dim * m outputs -- f_{d,j} multiplies the increment of driver j of component d, integrate_cv
// varred.py:202-214 -- and g has dim outputs, all multiplying the ONE common jump mark of the path (varred.py:124).
//
// Warp roles: warps 4*tl .. 4*tl+3 are the workers of tile tl (thread = path = TMEM lane; a warp can only access the
// TMEM lanes of its quadrant, warp % 4); warp kCvWorkerWarps + tl is the MMA issuer of tile tl (one thread).
// Hand-off per tile is by two mbarriers, no CTA barrier in the loop:
//   full[tl]  (4 arrivals)    workers -> issuer : D has been read (the next phase may overwrite it) and the
//                                                 activation vector that phase consumes is in TMEM
//   done[tl]  (1 arrival)     issuer  -> workers: tcgen05.commit of a phase's MMAs (or a plain arrive when the tile
//                                                 has no active path left; live[tl] says which)
// With two nets the phases alternate f, g, f, g ...: a worker arrives on full right after its tcgen05.ld of D and
// converts / stores the activations of the net just read while the tensor core runs the other net -- the vector the
// next phase consumes was stored during the PREVIOUS phase, i.e. before this arrival in program order.  With one net
// (diffusions) the arrival follows the store.
template <class C, bool JUMPS, bool INJECT>
__global__ void __launch_bounds__(kCvThreads, 2) cv_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const DevMlp f,
                                                        const DevMlp g, const DevCv cv,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  extern __shared__ __align__(1024) uint8_t cv_smem[];
  constexpr int MARKS = C::MARKS;
  constexpr int NETS = JUMPS ? 2 : 1;
  using Src = typename std::conditional<INJECT, InjectJumps<MARKS>, LazyJumps<MARKS>>::type;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool worker = warp < kCvWorkerWarps;
  const int quad = warp & 3;        // TMEM lane quadrant this warp may access (lanes 32*quad .. 32*quad+31)
  const int mine = worker ? (warp >> 2) : (warp - kCvWorkerWarps);  // the tile this warp works for
  const int row = quad * 32 + (tid & 31);
  const uint32_t sbase = smem_u32(cv_smem);
  const uint32_t bar_full = sbase + kCvOffBar + 8 * mine, bar_done = sbase + kCvOffBar + 8 * kCvTiles + 8 * mine;
  volatile int* flags = reinterpret_cast<volatile int*>(cv_smem + kCvOffFlags);  // [tile][parity] any-active, then live[tile]
  const bool narrow = f.hidden <= 56 && (!JUMPS || g.hidden <= 56);  // CTA-uniform: 56-column epilogues

  load_mlp(f, cv_smem + kCvOffW1F, cv_smem + kCvOffW2F, cv_smem + kCvOffW3F, cv_smem + kCvOffW4F);
  if (JUMPS) load_mlp(g, cv_smem + kCvOffW1G, cv_smem + kCvOffW2G, cv_smem + kCvOffW3G, cv_smem + kCvOffW4G);
  if (tid == 0) {
#pragma unroll
    for (int tl = 0; tl < kCvTiles; ++tl) {
      mbar_init(sbase + kCvOffBar + 8 * tl, 4);  // one arrival per worker warp
      mbar_init(sbase + kCvOffBar + 8 * kCvTiles + 8 * tl, 1);
    }
    for (int q = 0; q < 3 * kCvTiles; ++q) flags[q] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kCvOffTmem),
                 "n"(kCvTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the weight tiles -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(cv_smem + kCvOffTmem);
  const uint32_t tile_cols = tmem + (uint32_t)mine * kCvTileCols;

  const int n = s.num_steps;
  const int kcap = JUMPS ? (INJECT ? inj.K : 4 * (n + s.max_jumps) + 64) : n;
  const uint64_t n_tiles = (range_n(rg) + kCvRows - 1) / kCvRows;
  Accum acc;
  acc.zero();

  if (!worker) {
    // ========================= MMA issuer of tile `mine` (one thread of warp 8 + mine) ==========================
    if ((tid & 31) == 0) {
      uint32_t ph_full = 0;
      uint64_t bdesc[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t brows = r == 3 ? kCvHeadN : 64;
        const uint32_t wf = sbase + (r == 0 ? kCvOffW1F : r == 1 ? kCvOffW2F : r == 2 ? kCvOffW3F : kCvOffW4F);
        const uint32_t wg = sbase + (r == 0 ? kCvOffW1G : r == 1 ? kCvOffW2G : r == 2 ? kCvOffW3G : kCvOffW4G);
        bdesc[r][0] = umma_desc(wf, brows * 16, 128);
        bdesc[r][1] = umma_desc(wg, brows * 16, 128);
      }
      // phase (layer r, net): D[dcol ..] = A_net * W_r,net^T.  The B descriptor is affine in the k-step (the
      // start-address field counts 16-byte units and never carries out of its 14 bits below 256 KB); A advances 8
      // columns per k-step.
      auto issue = [&](int r, int net, uint32_t dcol, bool commit) {
        const int ksteps = r == 0 ? 1 : 4;
        const uint32_t idesc = r == 3 ? cv_idesc(kCvHeadN) : cv_idesc(64);
        const uint64_t b_step = (2u * ((r == 3 ? kCvHeadN : 64) * 16)) >> 4;
        uint64_t db = bdesc[r][net];
        uint32_t ta = tile_cols + (net ? kCvColAg : kCvColAf);
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_bf16_ts(tile_cols + kCvColD + dcol, ta, db, idesc, ks > 0);
          ta += 8;
          db += b_step;
        }
        if (commit) umma_commit(bar_done);
      };
      auto wait_workers = [&]() {
        mbar_wait(bar_full, ph_full);
        ph_full ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      };
      // the first phase of a step starts only if some path of the tile is still active (flag written by its owners)
      auto start_step = [&](int par) -> bool {
        wait_workers();
        const bool any = flags[mine * 2 + par] != 0;
        flags[mine * 2 + par] = 0;
        flags[2 * kCvTiles + mine] = any ? 1 : 0;
        __threadfence_block();  // flag writes before the arrival (commit or plain) the workers synchronise on
        if (any) issue(0, 0, 0, true);
        else mbar_arrive(bar_done);
        return any;
      };
      for (uint64_t pair = blockIdx.x; pair * kCvTiles < n_tiles; pair += gridDim.x) {
        bool live = start_step(0);
        for (int k = 0; live; ++k) {
#pragma unroll
          for (int ph = 1; ph < 3 * NETS; ++ph) {  // hidden layers, f and g alternating
            wait_workers();
            issue(ph / NETS, ph % NETS, 0, true);
          }
          wait_workers();                            // heads of both nets in one phase: D columns 0-15 and 16-31
          issue(3, 0, 0, !JUMPS);
          if (JUMPS) issue(3, 1, kCvHeadN, true);
          live = start_step((k + 1) & 1);
        }
      }
    }
  } else {
    // ======================================== workers ==========================================================
    constexpr int DIM = C::DIM, M = C::M, BASE = C::BASE;
    constexpr int NF = DIM * M;                      // outputs of f: one per (component, driver)
    constexpr int NG = DIM;                          // outputs of g: one per component
    constexpr int NZ = BASE + (M == 2 ? 1 : 0);      // unit normals per iteration (jump solver: ONE common 2nd driver)
    constexpr int LF = NF <= 1 ? 1 : (NF <= 2 ? 2 : 4), LG = NG <= 1 ? 1 : (NG <= 2 ? 2 : 4);
    static_assert(NF <= 4 && NZ <= 4 && !C::ASIAN, "head widths / one Philox block of normals per iteration");
    constexpr bool SHARE4 = NZ == 1;                 // 1-D: one Philox block serves four iterations
    struct Path {
      float x[kMaxDim], left[kMaxDim];
      float t, Jprev, cvsum;
      float zbuf[4];
      uint64_t i;
      uint32_t plo, phi;
      int own_iters;
      bool valid, need_pop;
      Src src;
    };
    Path p;
    uint32_t ph_done = 0;
#ifdef SDEMC_CV_PROFILE
    long long prof_wait = 0, prof_epi = 0, prof_adv = 0, prof_t0 = clock64(), prof_steps = 0;
#define CVP_BEGIN long long cvp_t = clock64();
#define CVP_END(acc) { const long long cvp_n = clock64(); acc += cvp_n - cvp_t; cvp_t = cvp_n; }
#else
#define CVP_BEGIN
#define CVP_END(acc)
#endif
    auto t_input = [&](const Path& q, int k) {
      return JUMPS ? q.t : (float)((double)s.T * (double)k / (double)n);  // partition(T, n, 'left')
    };
    auto is_active = [&](const Path& q, int k) { return q.valid && k < kcap && (JUMPS ? q.t < s.T : true); };
    // my tcgen05.ld / tcgen05.st have completed (tmem_wait_*) -> ordered before the issuer's next MMA.  One arrival
    // per warp (128 arrivals on one mbarrier word serialise): every lane fences, the warp converges, lane 0 arrives.
    auto release = [&]() {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(bar_full);
    };
    // phase finished?
    auto wait_phase = [&]() {
      mbar_wait(bar_done, ph_done);
      ph_done ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    // first phase of a step: false when the issuer retired the tile instead (no active path left)
    auto wait_step_start = [&]() -> bool {
      wait_phase();
      return flags[2 * kCvTiles + mine] != 0;
    };
    const uint32_t tl_lane = tile_cols + ((uint32_t)(quad * 32) << 16);  // my lanes, my tile's columns
    const uint32_t t_d = tl_lane + kCvColD, t_af = tl_lane + kCvColAf, t_ag = tl_lane + kCvColAg;
    // hidden-layer epilogue of one net: D -> registers, ReLU/bf16 -> the net's next activation vector.  EARLY: D is
    // released right after the load, so the tensor core runs the other net's phase during the conversion (the vector
    // that phase consumes was stored one phase ago); otherwise the arrival follows the store.
    auto hidden_epilogue = [&](uint32_t ta, bool early) {
      if (narrow) {
        CvAccRow<true> d;
        d.load(t_d);
        if (early) release();
        d.store_relu(ta);
      } else {
        CvAccRow<false> d;
        d.load(t_d);
        if (early) release();
        d.store_relu(ta);
      }
      tmem_wait_st();
      if (!early) release();
    };
    // One iteration of the path (solvers.py:182-225) and the weights its control-variate terms carry:
    //   cvsum += sum_{d,j} cf[d,j] f_{d,j}(t_k, x_k) + cg sum_d g_d(t_k, left_k)
    // with cf[d,j] = D dW_{d,j} and cg = D (J_prev - rate E[J] dt)   (integrate_cv varred.py:202-214, :124, :126-127;
    // the jump mark is common to all components, solvers.py:146-148).  Nothing here depends on the nets, so it runs
    // while the tensor core works on the first phases of the step.
    float cf[4] = {0.0f, 0.0f, 0.0f, 0.0f}, cg = 0.0f;
    bool cv_live = false;  // the path was active in the iteration just advanced (finished paths add nothing)
    auto advance_path = [&](int k) {
      const bool active = is_active(p, k);
      const float t_in = t_input(p, k);
      // this thread's unit normals for iteration k: one Philox block = four normals (two full Box-Muller pairs) serves
      // four iterations of a single-driver path, or one iteration of up to four drivers
      float zn[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if constexpr (INJECT) {
        if (p.valid && k < inj.K) {
#pragma unroll
          for (int d = 0; d < BASE; ++d) zn[d] = inj.z[(p.i * (uint64_t)inj.K + k) * DIM + d];
          if (M == 2) zn[BASE] = inj.zc[p.i * (uint64_t)inj.K + k];
        }
      } else if constexpr (SHARE4) {
        if ((k & 3) == 0) {
          uint32_t o[4];
          philox4x32_10((uint32_t)(k >> 2), STREAM_DIFFUSION, p.plo, p.phi, keys, o);
          box_muller(o[0], o[1], p.zbuf[0], p.zbuf[1]);
          box_muller(o[2], o[3], p.zbuf[2], p.zbuf[3]);
        }
        zn[0] = p.zbuf[0];
        p.zbuf[0] = p.zbuf[1]; p.zbuf[1] = p.zbuf[2]; p.zbuf[2] = p.zbuf[3];
      } else {
        uint32_t o[4];
        philox4x32_10((uint32_t)k, STREAM_DIFFUSION, p.plo, p.phi, keys, o);
        box_muller(o[0], o[1], zn[0], zn[1]);
        box_muller(o[2], o[3], zn[2], zn[3]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) cf[q] = 0.0f;
      cg = 0.0f;
      cv_live = active;
      if (active) {
        const float D = fast_ex2(-t_in * cv.disc_rate_l2e);
        float dt, sq;
        float tau = 0.0f;
        if constexpr (JUMPS) {
          p.src.begin_iter(s, keys, k);
          p.src.advance(s, keys, p.need_pop);
          tau = p.src.tau;
          dt = fmaxf(fminf(s.h0, fminf(tau, s.T) - p.t), 0.0f);  // stateless mesh, see jump.cuh
          sq = fast_sqrt(dt);
        } else {
          dt = s.h0;
          sq = s.sqrt_h0;
        }
        float xv[kMaxDim], xo[kMaxDim], z1[kMaxDim], w1[kMaxDim], w2[kMaxDim];
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) {
          xv[d] = xo[d] = p.x[d];
          z1[d] = d < BASE ? zn[d] : 0.0f;
          w2[d] = (M == 2 && d < BASE) ? zn[BASE] : 0.0f;   // the second driver is one scalar normal (:198-201)
        }
        correlate<C>(s, z1, w1);
        euler_step<C>(s, xv, dt, sq, w1, w2);
        const float Df = k < cv.brownian_steps ? D : 0.0f;           // tol > 0: the last steps carry no f dW term
#pragma unroll
        for (int d = 0; d < BASE; ++d) {
          cf[d * M] = Df * (w1[d] * sq);                             // f_{d,0} dW_{d,0}
          if (M == 2) cf[d * M + 1] = Df * (zn[BASE] * sq);          // f_{d,1} dW_{.,1}: the common driver
        }
        if constexpr (JUMPS) {
          cg = D * (k < cv.last_interval ? fmaf(cv.comp_c, dt, p.Jprev) : p.Jprev);  // g J - rate E[J] g dt
          p.t += dt;
#pragma unroll
          for (int d = 0; d < kMaxDim; ++d) p.left[d] = xv[d];
          const bool hit = fabsf(tau - p.t) <= fmaf(fabsf(p.t), 1e-5f, 1e-12f);
          const float Jc = hit ? p.src.mark(s, k) : 0.0f;
          if (s.exact_jumps) {
#pragma unroll
            for (int d = 0; d < kMaxDim; ++d) xo[d] = xv[d];
          }
          add_jump<C>(s, xv, xo, Jc);
          p.Jprev = Jc;
          p.need_pop = hit;
        }
#pragma unroll
        for (int d = 0; d < kMaxDim; ++d) p.x[d] = xv[d];
        p.own_iters = k + 1;
      }
    };

    for (uint64_t pair = blockIdx.x; pair * kCvTiles < n_tiles; pair += gridDim.x) {
      // ---- state 0 of my path -> first-layer activation vectors of my tile --------------------------------------
      p.i = (pair * kCvTiles + mine) * kCvRows + row;
      p.valid = p.i < range_n(rg);
      {
        const uint64_t gp = range_lo(rg) + p.i;
        p.plo = (uint32_t)gp;
        p.phi = (uint32_t)(gp >> 32);
      }
#pragma unroll
      for (int d = 0; d < kMaxDim; ++d) p.x[d] = p.left[d] = d < DIM ? s.x0[d] : 0.0f;
      p.t = 0.0f;
      p.Jprev = 0.0f;
      p.cvsum = 0.0f;
      p.own_iters = 0;
      p.need_pop = true;
#pragma unroll
      for (int q = 0; q < 4; ++q) p.zbuf[q] = 0.0f;
      if constexpr (JUMPS) {
        if constexpr (INJECT) p.src.init(s, inj, p.valid ? p.i : 0);
        else p.src.init(p.plo, p.phi);
      }
      write_input_row<DIM>(t_af, t_input(p, 0), p.x);
      if (JUMPS) write_input_row<DIM>(t_ag, t_input(p, 0), p.left);
      tmem_wait_st();
      if (is_active(p, 0)) flags[mine * 2 + 0] = 1;
      release();

      for (int k = 0;; ++k) {
        CVP_BEGIN
        if (!wait_step_start()) break;  // no active path was left in the tile: the issuer retired it
        CVP_END(prof_wait)
        // ---- hidden layers 1-3 of f (and g): six (three) phases; the last one keeps D until its vector is stored,
        // because the head phase consumes the vectors of BOTH nets
#pragma unroll
        for (int ph = 0; ph < 3 * NETS; ++ph) {
          hidden_epilogue(ph % NETS ? t_ag : t_af, JUMPS && ph < 3 * NETS - 1);
          if (ph == 0) advance_path(k);  // state k+1, off the critical path
          CVP_END(prof_epi)
          wait_phase();
          CVP_END(prof_wait)
        }
        // ---- heads: both nets evaluated at the state of index k ---------------------------------------------------
        {
          uint32_t hf[4] = {0u, 0u, 0u, 0u}, hg[4] = {0u, 0u, 0u, 0u};
          tmem_ld_small<LF>(t_d, hf);
          if (JUMPS) tmem_ld_small<LG>(t_d + kCvHeadN, hg);
          tmem_wait_ld();
          if (cv_live) {
#pragma unroll
            for (int q = 0; q < NF; ++q) p.cvsum = fmaf(cf[q], __uint_as_float(hf[q]), p.cvsum);
            if (JUMPS) {
#pragma unroll
              for (int q = 0; q < NG; ++q) p.cvsum = fmaf(cg, __uint_as_float(hg[q]), p.cvsum);
            }
          }
        }
        // first-layer activation vectors of state k+1; the issuer starts the next step if any path of the tile goes on
        write_input_row<DIM>(t_af, t_input(p, k + 1), p.x);
        if (JUMPS) write_input_row<DIM>(t_ag, t_input(p, k + 1), p.left);
        tmem_wait_st();
        if (is_active(p, k + 1)) flags[mine * 2 + ((k + 1) & 1)] = 1;
        release();
        CVP_END(prof_adv)
#ifdef SDEMC_CV_PROFILE
        ++prof_steps;
#endif
      }

      if (p.valid) {
        const float pay = eval_payoff<DIM>(po, p.x);
        const float gamma = pay + p.cvsum;
        if (cv.gamma_out) cv.gamma_out[p.i] = gamma;
        acc.add(gamma, pay, p.own_iters);
      }
    }
#ifdef SDEMC_CV_PROFILE
    if (blockIdx.x == 0 && (tid & 31) == 0)
      printf("cvprof warp %d: total %lld clk, steps %lld; per step: wait %.0f epilogues %.0f advance %.0f\n", warp,
             clock64() - prof_t0, prof_steps, (double)prof_wait / prof_steps, (double)prof_epi / prof_steps,
             (double)prof_adv / prof_steps);
#endif
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCvTmemCols));
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
