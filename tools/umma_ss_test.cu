// standalone check of hand-written tcgen05 descriptors: C[128x64] = A[128xK] * B[64xK]^T, bf16 in, fp32 out
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

__global__ void __launch_bounds__(128) umma_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C) {
  __shared__ __align__(1024) uint8_t sA[M * K * 2];
  __shared__ __align__(1024) uint8_t sB[N * K * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  // canonical no-swizzle K-major layout: 16-byte chunk c = k/8 of row r at c * (rows*16) + r * 16
  for (int idx = tid; idx < M * (K / 8); idx += 128) {
    int r = idx % M, c = idx / M;
    *reinterpret_cast<uint4*>(sA + c * (M * 16) + r * 16) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
  }
  for (int idx = tid; idx < N * (K / 8); idx += 128) {
    int r = idx % N, c = idx / N;
    *reinterpret_cast<uint4*>(sB + c * (N * 16) + r * 16) = *reinterpret_cast<const uint4*>(B + r * K + c * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * (M * 16), M * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
      const uint32_t accumulate = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
  }
  // wait for the MMA (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&bar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[64];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  for (int j = 0; j < 64; ++j) C[tid * N + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double a = 0; for (int k = 0; k < K; ++k) a += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)a; }
  __nv_bfloat16 *dA, *dB; float* dC;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dC, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0, M * N * 4);
  umma_kernel<<<1, 128>>>(dA, dB, dC);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dC, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
  printf("max abs err %.3e, mismatches %d / %d; C[0][0..3] = %f %f %f %f (ref %f %f %f %f)\n", maxerr, bad, M * N, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  printf("C[5][7]=%f ref %f; C[100][63]=%f ref %f\n", out[5*N+7], ref[5*N+7], out[100*N+63], ref[100*N+63]);
  return bad ? 1 : 0;
}
