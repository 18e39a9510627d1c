// Standalone check of two tcgen05 layouts the CV kernel wants to rely on (no documentation offline):
//   (1) A operand in TMEM (tcgen05.mma ... [d], [a_tmem], b_desc): row r of A = TMEM lane r, 32-bit column c holds
//       the f16 pair (k = 2c in the low half, k = 2c+1 in the high half); one K=16 step = 8 columns
//   (2) f16 accumulators (idesc D format 0): row r = lane r, column c holds (n = 2c low, n = 2c+1 high)
// C[128 x 64] = A[128 x 64] * B[64 x 64]^T, f16 inputs.  Prints the error under these hypotheses.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
#define LD32(taddr, v) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr))
#define ST32(taddr, v) asm volatile( \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
      :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")

template <bool D_F16>
__global__ void __launch_bounds__(128) k(const __half* A, const __half* B, float* C) {
  __shared__ __align__(1024) uint8_t sB[N * K * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < N * (K / 8); idx += 128) {
    int r = idx % N, c = idx / N;
    *reinterpret_cast<uint4*>(sB + c * (N * 16) + r * 16) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t tA = 64;  // A at columns 64..95, D at columns 0..63
  // my row of A -> TMEM: column c = (A[r][2c], A[r][2c+1])
  {
    uint32_t v[32];
    const uint32_t* row = reinterpret_cast<const uint32_t*>(A + tid * K);
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = row[c];
    ST32(tlane + tA, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = ((D_F16 ? 0u : 1u) << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      const uint32_t accumulate = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem),
          "r"(tmem + tA + ks * 8), "l"(db), "r"(idesc), "r"(accumulate));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0));
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32], w[32];
  LD32(tlane, v);
  LD32(tlane + 32, w);
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  if (D_F16) {
    for (int c = 0; c < 32; ++c) {
      const __half2 h = *reinterpret_cast<const __half2*>(&v[c]);
      C[tid * N + 2 * c] = __low2float(h);
      C[tid * N + 2 * c + 1] = __high2float(h);
    }
    for (int c = 0; c < 32; ++c) C[M * N + tid * 32 + c] = __uint_as_float(w[c]);  // raw dump of columns 32..63
  } else {
    for (int c = 0; c < 32; ++c) { C[tid * N + c] = __uint_as_float(v[c]); C[tid * N + 32 + c] = __uint_as_float(w[c]); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  std::vector<__half> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N + M * 32);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.0f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.0f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double a = 0; for (int kk = 0; kk < K; ++kk) a += (double)fA[m * K + kk] * fB[n * K + kk]; ref[m * N + n] = (float)a; }
  __half *dA, *dB; float* dC;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dC, out.size() * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dC, 0, out.size() * 4);
    if (variant == 0) k<false><<<1, 128>>>(dA, dB, dC); else k<true><<<1, 128>>>(dA, dB, dC);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0; const double tol = variant ? 3e-2 : 1e-3;
    for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > tol) ++bad; }
    printf("%s: %s  max abs err %.3e, mismatches %d / %d; C[0][0..3] = %f %f %f %f (ref %f %f %f %f); C[77][40]=%f ref %f\n",
           variant ? "A in TMEM, D f16" : "A in TMEM, D f32", cudaGetErrorString(e), maxerr, bad, M * N, out[0], out[1], out[2], out[3],
           ref[0], ref[1], ref[2], ref[3], out[77 * N + 40], ref[77 * N + 40]);
    if (variant) { int nz = 0; for (int i = 0; i < M * 32; ++i) nz += out[M * N + i] != 0.0f; printf("  non-zero raw words in columns 32..63 after an f16 accumulate: %d\n", nz); }
  }
  return 0;
}
