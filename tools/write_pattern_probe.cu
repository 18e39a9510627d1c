// Probe: HBM write bandwidth of the path-storing access pattern.  Each warp owns 32 rows (pitch floats apart) and
// writes them tile by tile: CHUNK floats of every row per tile (16-byte stores, CHUNK/4 lanes per row), with an
// optional ALU delay between tiles standing in for the step computation.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int CHUNK>
__global__ void k(float* out, long n_rows, int row_len, int pitch, int delay) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)(blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (long)(gridDim.x * blockDim.x) >> 5;
  constexpr int G = CHUNK / 4, R = 32 / G;  // lanes per row, rows per instruction
  float acc = lane;
  for (long g = warp; g * 32 < n_rows; g += nwarps) {
    for (int t = 0; t * CHUNK < row_len; ++t) {
      for (int d = 0; d < delay; ++d) acc = fmaf(acc, 1.0001f, 0.5f);
      if (G <= 32) {
#pragma unroll
        for (int i = 0; i < 32 / R; ++i) {
          const long row = g * 32 + i * R + lane / G;
          float4 v = make_float4(acc, acc, acc, acc);
          *reinterpret_cast<float4*>(out + row * pitch + t * CHUNK + 4 * (lane % G)) = v;
        }
      }
    }
  }
}
template <int CHUNK>
void run(float* d, long n_rows, int row_len, int pitch, int warps_per_sm, int delay) {
  int sms = 148;
  dim3 grid(sms * warps_per_sm / 4), block(128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHUNK><<<grid, block>>>(d, n_rows, row_len, pitch, delay);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<CHUNK><<<grid, block>>>(d, n_rows, row_len, pitch, delay);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  printf("chunk %4d B  warps/SM %2d  delay %4d : %.3f ms  %.0f GB/s\n", CHUNK * 4, warps_per_sm, delay, ms,
         (double)n_rows * row_len * 4 / ms / 1e6);
}
int main() {
  const long n_rows = 8000000; const int row_len = 256, pitch = 256;
  float* d; cudaMalloc(&d, n_rows * pitch * 4);
  for (int delay : {0, 200, 800})
    for (int w : {8, 16, 32}) {
      run<16>(d, n_rows, row_len, pitch, w, delay);
      run<32>(d, n_rows, row_len, pitch, w, delay);
      run<64>(d, n_rows, row_len, pitch, w, delay);
      run<128>(d, n_rows, row_len, pitch, w, delay);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
