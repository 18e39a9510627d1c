// Probe for the next step of the path-storing kernels ("longer rows at constant shared memory").
// tools/bulk_row_probe.cu showed that the copy engine is bound by row REQUESTS; diffusion_tma.cuh hands it boxes of
// 32 rows x 128 bytes.  Question: when a 4 KB box covers FEWER rows with MORE contiguous bytes each -- a 3-D tensor
// map {32 floats, chunk, row} with box {32, C, 32 / C}: C adjacent 128-byte lines of the same row -- does the engine
// merge the adjacent lines into one request (rate grows with C) or still issue one per line (rate flat)?
// Every warp owns a double-buffered 4 KB tile (128B swizzle, contents irrelevant) and walks its rows chunk by chunk
// like the kernel does; only the copy engine is measured.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_box_probe.cu -o tools/tma_box_probe.bin
// Measured (B200): 6 warps/SM 5.6 / 6.1 / 6.1 / 6.2 TB/s for 128 / 256 / 512 / 1024 contiguous bytes per row; 12 warps/SM
// 5.4 / 6.1 / 6.1 / 6.2; 24 warps/SM 5.4 / 6.4 / 6.5 / 6.7 -- the engine alone is NOT the 3.5 TB/s limit of the kernel.
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

// (n_rows, row_len) fp32, pitch floats between rows, viewed as {32 floats, row_len / 32 chunks, n_rows}; box {32, C, R}
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

// one warp per CTA (residency = 8 KB of tiles per warp); a warp owns R rows at a time and C chunks per copy
__global__ void __launch_bounds__(32) k(const __grid_constant__ CUtensorMap map, long n_rows, int chunks, int C, int R) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned base = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int lane = threadIdx.x;
  for (int i = lane; i < 2048; i += 32) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  int buf = 0;
  for (long g = blockIdx.x; g * R < n_rows; g += gridDim.x) {
    for (int c = 0; c < chunks; c += C) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map),
                     "r"(0), "r"(c), "r"((int)(g * R)), "r"(base + buf * 4096u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      __syncwarp();
      buf ^= 1;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const long n_rows = 4000000;
  const int row_len = 256, pitch = 256;  // 1 KB rows: the GBM solve() layout
  float* d;
  cudaMalloc(&d, (size_t)n_rows * pitch * 4);
  const size_t smem = 2 * 4096 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int warps : {6, 12, 24}) {
    for (int C : {1, 2, 4, 8}) {
      const int R = 32 / C;
      CUtensorMap map;
      if (!make_map(&map, d, n_rows, row_len, pitch, C, R)) {
        printf("box {32, %d, %d}: tensor map rejected\n", C, R);
        continue;
      }
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      k<<<148 * warps, 32, smem>>>(map, n_rows, row_len / 32, C, R);
      cudaEventRecord(e0);
      for (int r = 0; r < 3; ++r) k<<<148 * warps, 32, smem>>>(map, n_rows, row_len / 32, C, R);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 3;
      printf("warps/SM %2d  box {32 floats, %d chunks, %2d rows} = %4d contiguous bytes per row : %.3f ms  %.0f GB/s  (%s)\n",
             warps, C, R, C * 128, ms, (double)n_rows * row_len * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
