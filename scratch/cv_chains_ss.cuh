// cv.cuh -- fused control-variate Monte Carlo (E5/E6/E7): simulate + evaluate the control-variate MLPs f and g
// along every path + accumulate gamma = payoff + sum f dW D + sum g D J - sum rate E[J] g D dt + moments, in ONE
// kernel.  Replaces mc_apply_cvs mc.py:195-242 -> simulate_adapted_data mc.py:391-398 (full trajectory storage,
// ~20 B per path-step through HBM) -> apply_adapted_control_variates varred.py:98-131 (two MLP forwards over
// bs*S rows), resp. the diffusion variant varred.py:75-95.
//
// Tensor cores (tcgen05, accumulators in TMEM): a CTA owns TWO tiles of 128 paths; path r of a tile is row r of its
// activation matrices and lane r of its TMEM accumulators.  Every tile is an independent chain: four worker warps
// (one per TMEM lane quadrant; warps 0-3 tile 0, warps 4-7 tile 1) simulate its paths and run its epilogues, and an
// issuer warp of its own (warp 8 + tile) feeds it to the tensor core, so the four chains resident on an SM (2 CTAs x 2
// tiles = all 512 TMEM columns) interleave on the tensor pipe, the TMEM read port and the issue slots at their own
// pace instead of walking one fixed (round, tile) order.  Each time step evaluates
//   Linear(2,H) ReLU Linear(H,H) ReLU Linear(H,H) ReLU Linear(H,1)         (nets.py:39-93, BN-free, H <= 63)
// for both nets as FOUR rounds of tcgen05.mma per tile (M=128; N=64,K=16 | N=64,K=64 | N=64,K=64 | N=16,K=64; bf16
// in, fp32 accumulate) whose A operand the threads write themselves into shared memory in the canonical K-major
// no-swizzle UMMA layout.  While the tensor core runs a round of one tile, the workers of the other tiles run their
// epilogues (TMEM -> ReLU -> bf16 -> next A operand): MMA, commit and mbarrier wake-up latency of one chain hide
// behind the SIMT work of the others; there is no CTA barrier in the step loop.
// Biases are folded into the contraction: every padded activation vector carries a constant 1 in slot 63
// (W[n][63] = b[n]); for H <= 56 the epilogue writes that constant itself and reads only 56 accumulator columns
// (TMEM reads are the tightest floor of a step), otherwise W[63][63] = 1 carries it through the MMA.  The inputs (t, x) are split into bf16 hi + lo parts (two K slots each with the
// same weight) so the nets are evaluated at fp32-accurate inputs; weights and hidden activations are bf16.
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

constexpr int kCvRows = 128;     // paths per tile = rows of the activation matrices = TMEM lanes
constexpr int kCvWorkerWarps = 8; // warps 0-3: paths and epilogues of tile 0, warps 4-7: tile 1
constexpr int kCvIssuerWarps = 2; // warp 8 + tl issues the MMAs of tile tl (one elected thread)
constexpr int kCvThreads = (kCvWorkerWarps + kCvIssuerWarps) * 32;
constexpr int kCvOne = 63;  // index of the constant-one unit in every padded (64-wide) activation vector

// shared-memory carve-up (bytes).  Operand tiles: 16-byte chunk c = k/8 of row r lives at c * (rows*16) + r * 16,
// i.e. UMMA descriptors with LBO = rows*16 (K direction) and SBO = 128 (next group of 8 rows).
constexpr int kCvTiles = 2;                 // path tiles per CTA, pipelined against each other
constexpr int kCvHeadN = 16;                // N of the head MMA (smallest N for M = 128); only column 0 is used
constexpr int kCvW1Bytes = 64 * 16 * 2;
constexpr int kCvWBytes = 64 * 64 * 2;
constexpr int kCvW4Bytes = kCvHeadN * 64 * 2;
constexpr int kCvABytes = 128 * 64 * 2;
constexpr int kCvOffW1F = 0;
constexpr int kCvOffW1G = kCvOffW1F + kCvW1Bytes;
constexpr int kCvOffW2F = kCvOffW1G + kCvW1Bytes;
constexpr int kCvOffW3F = kCvOffW2F + kCvWBytes;
constexpr int kCvOffW2G = kCvOffW3F + kCvWBytes;
constexpr int kCvOffW3G = kCvOffW2G + kCvWBytes;
constexpr int kCvOffW4F = kCvOffW3G + kCvWBytes;
constexpr int kCvOffW4G = kCvOffW4F + kCvW4Bytes;
constexpr int kCvOffA = kCvOffW4G + kCvW4Bytes;  // per tile: A_f then A_g
constexpr int kCvOffBar = kCvOffA + kCvTiles * 2 * kCvABytes;
constexpr int kCvOffTmem = kCvOffBar + 16 * kCvTiles;  // full[tile] then done[tile]
constexpr int kCvOffFlags = kCvOffTmem + 8;            // int any_active[tile][parity], int live[tile]
constexpr int kCvSmemBytes = kCvOffFlags + 4 * 3 * kCvTiles + 8;
constexpr int kCvTmemCols = 128 * kCvTiles;  // per tile: f accumulators in columns 0-63, g in 64-127

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

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
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
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
// eight accumulator columns -> ReLU -> bf16 -> one 16-byte chunk of the next A operand
__device__ __forceinline__ void relu_store_chunk(uint32_t addr, const uint32_t* v) {
  sts128(addr, relu_pack_bf16x2(v[0], v[1]), relu_pack_bf16x2(v[2], v[3]), relu_pack_bf16x2(v[4], v[5]),
         relu_pack_bf16x2(v[6], v[7]));
}
// epilogue of a hidden layer for this thread's row: accumulator columns -> ReLU -> bf16 -> the eight 16-byte chunks
// of the row of the next A operand.  NARROW (H <= 56): only columns 0-55 are read (32 + 16 + 8); chunk 7 is the
// constant (0 x 7, 1.0) that the full-width path gets from W[63][63] = 1.
template <bool NARROW>
__device__ __forceinline__ void hidden_epilogue_row(uint32_t taddr, uint32_t a_row_addr) {
  constexpr uint32_t kChunk = kCvRows * 16;
  {
    uint32_t v[32];
    tmem_ld32(taddr, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) relu_store_chunk(a_row_addr + c * kChunk, v + 8 * c);
  }
  if constexpr (NARROW) {
    uint32_t v[16], w[8];
    tmem_ld16(taddr + 32, v);
    tmem_ld8(taddr + 48, w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    relu_store_chunk(a_row_addr + 4 * kChunk, v);
    relu_store_chunk(a_row_addr + 5 * kChunk, v + 8);
    relu_store_chunk(a_row_addr + 6 * kChunk, w);
    sts128(a_row_addr + 7 * kChunk, 0u, 0u, 0u, 0x3f800000u /* (0, 1.0bf16): slot 63 */);
  } else {
    uint32_t v[32];
    tmem_ld32(taddr + 32, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) relu_store_chunk(a_row_addr + (4 + c) * kChunk, v + 8 * c);
  }
}
// output of the head MMA: column 0 of this thread's accumulator row
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(v);
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
__device__ __forceinline__ void load_mlp(const DevMlp& net, uint8_t* w1, uint8_t* w2, uint8_t* w3, uint8_t* w4) {
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
  // head: B operand of kCvHeadN rows, row 0 = (w4, bias in the constant-one slot), the other rows zero
  for (int idx = threadIdx.x; idx < kCvHeadN * 64; idx += blockDim.x) {
    const int n = idx >> 6, k = idx & 63;
    float v = 0.0f;
    if (n == 0) v = k < H ? net.w[3][k] : (k == kCvOne ? net.b[3][0] : 0.0f);
    *reinterpret_cast<__nv_bfloat16*>(w4 + (k >> 3) * (kCvHeadN * 16) + n * 16 + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
}

struct DevCv {
  float disc_rate_l2e;  // r * log2(e):  D(t) = 2^(-t r log2 e)   (ConstantShortRate options.py:334-337)
  float comp_c;         // - rate * E[J]                            (varred.py:126)
  int last_interval;    // compensator intervals with index >= last_interval are dropped (varred.py:104,126-127)
  float* gamma_out;     // (n) per-path gamma or nullptr
};

// ---- mbarrier helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// 1-D 'diag' SDE (dim == 1, m == 1): Merton-type jump diffusion (JUMPS) or GBM-type diffusion (!JUMPS)
//
// Warp roles: warps 0-3 are the workers of tile 0, warps 4-7 those of tile 1 (thread = path = TMEM lane; a warp can
// only read the TMEM lanes of its quadrant, warp % 4); warp 8 + tl is the MMA issuer of tile tl (one thread).
// Hand-off per tile is by two mbarriers, no CTA barrier in the loop:
//   full[tl]  (128 arrivals)  workers -> issuer : the A operands of the next round are in shared memory
//   done[tl]  (1 arrival)     issuer  -> workers: tcgen05.commit of that round's MMAs (or a plain arrive when the tile
//                                                 has no active path left; tile_live[tl] says which)
// The tiles of a CTA never wait for each other: a tile whose paths finished early is retired and its warps start the
// next pair's tile while the other one is still stepping.
template <class C, bool JUMPS, bool INJECT>
__global__ void __launch_bounds__(kCvThreads, 2) cv_kernel(const DevSde s, const DevPayoff po, const DevRange rg,
                                                        const PhiloxKeys keys, const DevInject inj, const DevMlp f,
                                                        const DevMlp g, const DevCv cv,
                                                        double* __restrict__ d_moments, void* __restrict__ d_ws) {
  extern __shared__ __align__(1024) uint8_t cv_smem[];
  constexpr int MARKS = C::MARKS;
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
      mbar_init(sbase + kCvOffBar + 8 * tl, (kCvWorkerWarps / kCvTiles) * 32);
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(cv_smem + kCvOffTmem);
  const uint32_t tacc_f = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)mine * 128u;  // my lanes, my tile's columns
  const uint32_t tacc_g = tacc_f + 64;

  const int n = s.num_steps;
  const int kcap = JUMPS ? (INJECT ? inj.K : 4 * (n + s.max_jumps) + 64) : n;
  const uint64_t n_tiles = (rg.n_paths + kCvRows - 1) / kCvRows;
  Accum acc;
  acc.zero();

  if (!worker) {
    // ========================= MMA issuer of tile `mine` (one thread of warp 8 + mine) ==========================
    if ((tid & 31) == 0) {
      uint32_t ph_full = 0;
      // Descriptors are affine in the k-step (the start-address field counts 16-byte units and never carries out of
      // its 14 bits for addresses below 256 KB), so the issue loop is one 64-bit add per operand and MMA: the
      // issuing thread sits on its tile's critical path.
      uint64_t adesc[2], bdesc[4][2];
      {
        const uint32_t af = sbase + kCvOffA + (uint32_t)mine * (2 * kCvABytes);
        adesc[0] = umma_desc(af, 128 * 16, 128);
        adesc[1] = umma_desc(af + kCvABytes, 128 * 16, 128);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t brows = r == 3 ? kCvHeadN : 64;
        const uint32_t wf = sbase + (r == 0 ? kCvOffW1F : r == 1 ? kCvOffW2F : r == 2 ? kCvOffW3F : kCvOffW4F);
        const uint32_t wg = sbase + (r == 0 ? kCvOffW1G : r == 1 ? kCvOffW2G : r == 2 ? kCvOffW3G : kCvOffW4G);
        bdesc[r][0] = umma_desc(wf, brows * 16, 128);
        bdesc[r][1] = umma_desc(wg, brows * 16, 128);
      }
      const uint32_t accum = tmem + (uint32_t)mine * 128u;
      auto issue = [&](int r) {
        const int ksteps = r == 0 ? 1 : 4;
        const uint32_t idesc = r == 3 ? cv_idesc(kCvHeadN) : cv_idesc(64);
        const uint64_t a_step = (2u * (128 * 16)) >> 4;                                   // two 16-byte chunks of A
        const uint64_t b_step = (2u * ((r == 3 ? kCvHeadN : 64) * 16)) >> 4;
        uint64_t daf = adesc[0], dag = adesc[1], dbf = bdesc[r][0], dbg = bdesc[r][1];
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_bf16(accum, daf, dbf, idesc, ks > 0);
          if (JUMPS) umma_bf16(accum + 64, dag, dbg, idesc, ks > 0);
          daf += a_step; dag += a_step; dbf += b_step; dbg += b_step;
        }
        umma_commit(bar_done);
      };
      auto wait_operands = [&]() {
        mbar_wait(bar_full, ph_full);
        ph_full ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      };
      // round 0 of a step starts only if some path of the tile is still active (flag written by its owners)
      auto start_step = [&](int par) -> bool {
        wait_operands();
        const bool any = flags[mine * 2 + par] != 0;
        flags[mine * 2 + par] = 0;
        flags[2 * kCvTiles + mine] = any ? 1 : 0;
        __threadfence_block();  // flag writes before the arrival (commit or plain) the workers synchronise on
        if (any) issue(0);
        else mbar_arrive(bar_done);
        return any;
      };
      for (uint64_t pair = blockIdx.x; pair * kCvTiles < n_tiles; pair += gridDim.x) {
        bool live = start_step(0);
        for (int k = 0; live; ++k) {
#pragma unroll
          for (int r = 1; r < 4; ++r) {
            wait_operands();
            issue(r);
          }
          live = start_step((k + 1) & 1);
        }
      }
    }
  } else {
    // ======================================== workers ==========================================================
    struct Path {
      float x, t, left, Jprev, cvsum;
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
    long long prof_wait = 0, prof_epi = 0, prof_ready = 0, prof_adv = 0, prof_t0 = clock64(), prof_rounds = 0;
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
    // my st.shared -> visible to the tensor core; my tcgen05.ld -> ordered before the next MMA; tell the issuer
    auto operands_ready = [&]() {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_full);
    };
    // round finished?  returns false when the issuer retired the tile instead
    auto wait_round = [&]() -> bool {
      mbar_wait(bar_done, ph_done);
      ph_done ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      return flags[2 * kCvTiles + mine] != 0;
    };
    const uint32_t a_row_f = sbase + kCvOffA + (uint32_t)mine * (2 * kCvABytes) + row * 16, a_row_g = a_row_f + kCvABytes;

    for (uint64_t pair = blockIdx.x; pair * kCvTiles < n_tiles; pair += gridDim.x) {
      // ---- state 0 of my path -> first-layer operands of my tile ------------------------------------------------
      p.i = (pair * kCvTiles + mine) * kCvRows + row;
      p.valid = p.i < rg.n_paths;
      {
        const uint64_t gp = rg.path_lo + p.i;
        p.plo = (uint32_t)gp;
        p.phi = (uint32_t)(gp >> 32);
      }
      p.x = s.x0[0];
      p.t = 0.0f;
      p.left = s.x0[0];
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
      write_input_row(a_row_f, t_input(p, 0), p.x);
      if (JUMPS) write_input_row(a_row_g, t_input(p, 0), p.left);
      if (is_active(p, 0)) flags[mine * 2 + 0] = 1;
      operands_ready();

      for (int k = 0;; ++k) {
        CVP_BEGIN
        bool ok = wait_round();
        CVP_END(prof_wait)
        if (!ok) break;  // no active path was left in the tile: the issuer retired it
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          // hidden layer r+1: accumulators -> ReLU -> bf16 -> A operand of the next round
          if (narrow) {
            hidden_epilogue_row<true>(tacc_f, a_row_f);
            if (JUMPS) hidden_epilogue_row<true>(tacc_g, a_row_g);
          } else {
            hidden_epilogue_row<false>(tacc_f, a_row_f);
            if (JUMPS) hidden_epilogue_row<false>(tacc_g, a_row_g);
          }
          CVP_END(prof_epi)
          operands_ready();
          CVP_END(prof_ready)
          wait_round();
          CVP_END(prof_wait)
        }
#ifdef SDEMC_CV_PROFILE
        prof_rounds += 4;
#endif
        // ---- head round done: both nets evaluated at the state of index k -> advance my path by one iteration ----
        const float fval = tmem_ld1(tacc_f);
        const float gval = JUMPS ? tmem_ld1(tacc_g) : 0.0f;
        const bool active = is_active(p, k);
        const float t_in = t_input(p, k);
        // this thread's Brownian normal for iteration k (one Philox block serves 4 iterations)
        if ((k & 3) == 0) {
          if constexpr (!INJECT) {
            uint32_t o[4];
            philox4x32_10((uint32_t)(k >> 2), STREAM_DIFFUSION, p.plo, p.phi, keys, o);
            box_muller(o[0], o[1], p.zbuf[0], p.zbuf[1]);
            box_muller(o[2], o[3], p.zbuf[2], p.zbuf[3]);
          }
        }
        float z;
        if constexpr (INJECT) {
          z = (p.valid && k < inj.K) ? inj.z[p.i * (uint64_t)inj.K + k] : 0.0f;
        } else {
          z = p.zbuf[0];
          p.zbuf[0] = p.zbuf[1]; p.zbuf[1] = p.zbuf[2]; p.zbuf[2] = p.zbuf[3];
        }
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
          const float dW = z * sq;
          float xv[kMaxDim] = {p.x, 0.0f, 0.0f, 0.0f}, xo[kMaxDim] = {p.x, 0.0f, 0.0f, 0.0f};
          float w1[kMaxDim] = {z, 0.0f, 0.0f, 0.0f}, w2[kMaxDim] = {0.0f, 0.0f, 0.0f, 0.0f};
          euler_step<C>(s, xv, dt, sq, w1, w2);
          float c = fval * dW;                                     // f dW           (integrate_cv varred.py:202-214)
          if constexpr (JUMPS) {
            c = fmaf(gval, p.Jprev, c);                            // g J            (varred.py:124)
            if (k < cv.last_interval) c = fmaf(cv.comp_c * gval, dt, c);   // - rate E[J] g dt (varred.py:126-127)
            p.t += dt;
            p.left = xv[0];
            const bool hit = fabsf(tau - p.t) <= fmaf(fabsf(p.t), 1e-5f, 1e-12f);
            const float Jc = hit ? p.src.mark(s, k) : 0.0f;
            if (s.exact_jumps) xo[0] = xv[0];
            add_jump<C>(s, xv, xo, Jc);
            p.Jprev = Jc;
            p.need_pop = hit;
          }
          p.x = xv[0];
          p.cvsum = fmaf(c, D, p.cvsum);
          p.own_iters = k + 1;
        }
        // first-layer operands of state k+1; the issuer starts the next step if any path of the tile goes on
        write_input_row(a_row_f, t_input(p, k + 1), p.x);
        if (JUMPS) write_input_row(a_row_g, t_input(p, k + 1), p.left);
        if (is_active(p, k + 1)) flags[mine * 2 + ((k + 1) & 1)] = 1;
        CVP_END(prof_adv)
        operands_ready();
        CVP_END(prof_ready)
      }

      if (p.valid) {
        float xp[kMaxDim] = {p.x, 0.0f, 0.0f, 0.0f};
        const float pay = eval_payoff<1>(po, xp);
        const float gamma = pay + p.cvsum;
        if (cv.gamma_out) cv.gamma_out[p.i] = gamma;
        acc.add(gamma, pay, p.own_iters);
      }
    }
#ifdef SDEMC_CV_PROFILE
    if (blockIdx.x == 0 && (tid & 31) == 0)
      printf("cvprof warp %d: total %lld clk, rounds %lld; per round: wait %.0f epi %.0f ready %.0f adv %.0f\n", warp,
             clock64() - prof_t0, prof_rounds, (double)prof_wait / prof_rounds, (double)prof_epi / prof_rounds,
             (double)prof_ready / prof_rounds, (double)prof_adv / prof_rounds);
#endif
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCvTmemCols));
  block_reduce_and_publish(acc, d_moments, d_ws);
}

}  // namespace sdemc
