// Why does the TMA engine move the path-storing kernel's boxes at 3.5 TB/s when tools/tma_box_probe.cu measured
// 5.4-5.6 TB/s for the same boxes?  Differences to the probe: the kernel alternates between TWO tensor maps (paths,
// increments: two 4 GB arrays) and spaces its copies by compute.  This probe adds the second map / array, and an
// optional delay loop between copies, one factor at a time.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_two_map_probe.cu -o tools/tma_two_map_probe.bin
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static bool make_map(CUtensorMap* map, float* base, unsigned long long n_rows, unsigned long long row_len,
                     unsigned long long pitch, unsigned C, unsigned R) {
  auto enc = encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {32, row_len / 32, n_rows};
  const cuuint64_t strides[2] = {32 * sizeof(float), pitch * sizeof(float)};
  const cuuint32_t box[3] = {32, C, R};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// W warps per CTA; each warp owns 32 rows (R rows per copy, 32/R copies per chunk group) and alternates between the
// two arrays like the kernel (tile of array A, tile of array B, next columns ...)
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                         long n_rows, int chunks, int C, int R, int two, int delay, int stage, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned base = (((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u) + warp * 16384u;
  for (int i = threadIdx.x; i < 4096 * (blockDim.x >> 5); i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  int buf = 0;
  float acc = (float)lane;
  const long stride = (long)gridDim.x * (blockDim.x >> 5);
  for (long g = (long)blockIdx.x * (blockDim.x >> 5) + warp; g * 32 < n_rows; g += stride) {
    for (int c = 0; c < chunks; c += C) {
      for (int rr = 0; rr < 32; rr += R) {
        for (int a = 0; a <= two; ++a) {
          for (int d = 0; d < delay; ++d) acc = fmaf(acc, 1.0001f, 0.5f);   // stand-in for the step loop
          if (stage) {   // stage the tile like the kernel: 8 swizzled 16-byte stores per lane, then the proxy fence
            const unsigned t0 = base + (a * 2 + buf) * 4096u + lane * 128u;
#pragma unroll
            for (int v = 0; v < 8; ++v)
              asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(t0 + (((unsigned)v ^ (lane & 7u)) << 4)), "f"(acc) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
          }
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(a ? &mapB : &mapA),
                         "r"(0), "r"(c), "r"((int)(g * 32 + rr)), "r"(base + (a * 2 + buf) * 4096u)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (two) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          }
          __syncwarp();
        }
        buf ^= 1;
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const long n_rows = 4000000;
  const int row_len = 256, pitch = 256;
  float *dA, *dB, *sink;
  cudaMalloc(&dA, (size_t)n_rows * pitch * 4);
  cudaMalloc(&dB, (size_t)n_rows * pitch * 4);
  cudaMalloc(&sink, 4);
  const size_t smem = 4 * 16384 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int stage : {1})
  for (int two : {1}) {
    for (int C : {1, 2, 4, 8}) {
      for (int delay : {0, 400, 600}) {
        const int R = 32 / C;
        CUtensorMap mA, mB;
        if (!make_map(&mA, dA, n_rows, row_len, pitch, C, R) || !make_map(&mB, dB, n_rows, row_len, pitch, C, R)) continue;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k<<<148 * 3, 128, smem>>>(mA, mB, n_rows, row_len / 32, C, R, two, delay, stage, sink);
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) k<<<148 * 3, 128, smem>>>(mA, mB, n_rows, row_len / 32, C, R, two, delay, stage, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 3;
        printf("stage %d arrays %d  box {32, %d, %2d} (%4d B/row)  delay %4d FMA/copy : %.3f ms  %.0f GB/s  (%s)\n", stage, two + 1, C, R,
               C * 128, delay, ms, (double)n_rows * row_len * 4 * (two + 1) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
