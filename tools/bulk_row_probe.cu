// Probe: what bounds the TMA / bulk-copy engine in the path-storing pattern -- bytes or row requests?
// Each warp owns 32 rows of a (n_rows, row_len) fp32 array (pitch floats apart) and writes them chunk by chunk from a
// shared-memory tile with ONE bulk copy per row and chunk (cp.async.bulk.global.shared::cta, issued by the lane that
// owns the row: 32 copies per warp instruction), CHUNK bytes each.  If the engine were limited by bytes the rate
// would not depend on CHUNK; if it is limited by requests (what diffusion_tma.cuh's 32 x 128-byte boxes suggest) the
// rate scales with CHUNK.  The tile contents are never rewritten (only the copy engine is measured); tiles are double
// buffered and reused after cp.async.bulk.wait_group.read like in the kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bulk_row_probe.cu -o gpurun_out/bulk_row_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int CHUNK_BYTES>
__global__ void __launch_bounds__(32) k(float* out, long n_rows, int row_len, int pitch) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = 0;  // one warp per CTA: residency is set by the tile size alone
  constexpr int ROW = CHUNK_BYTES + 16;  // padded rows: conflict-free 16-byte staging stores in a real kernel
  unsigned char* tile = smem + (size_t)w * 2 * 32 * ROW;
  for (int i = lane; i < 2 * 32 * ROW / 4; i += 32) reinterpret_cast<float*>(tile)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const long warp = (long)(blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (long)(gridDim.x * blockDim.x) >> 5;
  const int chunks = row_len * 4 / CHUNK_BYTES;
  int buf = 0;
  for (long g = warp; g * 32 < n_rows; g += nwarps) {
    const long row = g * 32 + lane;
    for (int t = 0; t < chunks; ++t) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(tile + (size_t)(buf * 32 + lane) * ROW);
      float* dst = out + row * pitch + (long)t * (CHUNK_BYTES / 4);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "n"(CHUNK_BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the other buffer may be refilled
      buf ^= 1;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int CHUNK_BYTES>
void run(float* d, long n_rows, int row_len, int pitch, int ctas_per_sm) {
  const size_t smem = 2 * 32 * (CHUNK_BYTES + 16);
  if (smem * ctas_per_sm > 220 * 1024) return;
  cudaFuncSetAttribute(k<CHUNK_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(148 * ctas_per_sm), block(32);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CHUNK_BYTES><<<grid, block, smem>>>(d, n_rows, row_len, pitch);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<CHUNK_BYTES><<<grid, block, smem>>>(d, n_rows, row_len, pitch);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  printf("row chunk %4d B  warps/SM %2d  : %.3f ms  %.0f GB/s  (%s)\n", CHUNK_BYTES, ctas_per_sm, ms, (double)n_rows * row_len * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long n_rows = 4000000;
  const int row_len = 256, pitch = 256;  // 1 KB rows, the GBM solve() layout (253 steps padded to 256 floats)
  float* d;
  cudaMalloc(&d, n_rows * pitch * 4);
  for (int c : {3, 6, 12, 24}) {
    run<64>(d, n_rows, row_len, pitch, c);
    run<128>(d, n_rows, row_len, pitch, c);
    run<256>(d, n_rows, row_len, pitch, c);
    run<512>(d, n_rows, row_len, pitch, c);
    run<1024>(d, n_rows, row_len, pitch, c);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
