// Probe: every THREAD owns one row (one path) and writes it front to back with 16-byte (st.global.v4.f32) or 32-byte
// (st.global.v8.f32, sm_100+) stores -- no shared-memory staging, no transposition: what does HBM take when a warp
// instruction touches 32 different rows with one half / one whole 32-byte sector each?  `delay` dependent FMAs
// between stores stand in for the step loop; arrays = 1, 2 or 5 rows per thread (paths + normals, or the five arrays
// of the jump solver).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/thread_row_store_probe.cu -o tools/thread_row_store_probe.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int VEC, int ARR>
__global__ void __launch_bounds__(256) k(float* base, long n_rows, int row_len, int pitch, long arr_stride, int delay) {
  float acc = threadIdx.x;
  for (long r = (long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long)gridDim.x * blockDim.x) {
    float* row = base + r * pitch;
    for (int c = 0; c < row_len; c += VEC) {
      for (int d = 0; d < delay; ++d) acc = fmaf(acc, 1.0001f, 0.5f);
#pragma unroll
      for (int a = 0; a < ARR; ++a) {
        float* p = row + a * arr_stride + c;
        if (VEC == 4) {
          asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(acc) : "memory");
        } else {
          asm volatile("st.global.v8.f32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"l"(p), "f"(acc) : "memory");
        }
      }
    }
  }
}

template <int VEC, int ARR>
void run(float* d, long n_rows, int row_len, int pitch, int ctas_per_sm, int delay) {
  const long arr_stride = n_rows * pitch;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<VEC, ARR><<<148 * ctas_per_sm, 256>>>(d, n_rows, row_len, pitch, arr_stride, delay);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<VEC, ARR><<<148 * ctas_per_sm, 256>>>(d, n_rows, row_len, pitch, arr_stride, delay);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  printf("st.v%d  arrays %d  row %4d floats (pitch %4d)  %d CTAs/SM  delay %4d : %.3f ms  %.0f GB/s  (%s)\n", VEC, ARR, row_len, pitch,
         ctas_per_sm, delay, ms, (double)n_rows * row_len * 4 * ARR / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* d;
  const size_t bytes = (size_t)9 << 30;
  if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  for (int ctas : {2, 4, 8}) {
    for (int delay : {0, 64}) {
      // GBM solve(): 4e6 rows of 256 floats, two arrays
      run<4, 2>(d, 4000000, 256, 256, ctas, delay);
      run<8, 2>(d, 4000000, 256, 256, ctas, delay);
      // Merton solve(): 2e6 rows of 160 floats (134 used), five arrays
      run<4, 5>(d, 2000000, 136, 160, ctas, delay);
      run<8, 5>(d, 2000000, 136, 160, ctas, delay);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
