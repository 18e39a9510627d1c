// micro-benchmarks of the RNG building blocks: where do the issue slots go?
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../sde_mc_b200/csrc/philox.cuh"
using namespace sdemc;

template <int MODE>
__global__ void __launch_bounds__(256) k(PhiloxKeys keys, int iters, float c0, float c1, float* out) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  float x = 1.0f; uint32_t acc = 0;
  for (int b = 0; b < iters; ++b) {
    uint32_t o[4];
    if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4) {            // philox
      philox4x32_10(b, 0, tid, 0, keys, o);
    } else {                                  // cheap bits: no philox
      o[0] = tid * 2654435761u + b; o[1] = o[0] ^ 0x9e3779b9u; o[2] = o[0] + 0x7f4a7c15u; o[3] = o[1] + b;
    }
    if (MODE == 0) { acc ^= o[0] ^ o[1] ^ o[2] ^ o[3]; }
    if (MODE == 1 || MODE == 2) {             // box-muller polar + 2 ffma per normal
      float r0, cc0, s0, r1, cc1, s1;
      box_muller_polar(o[0], o[1], c1, r0, cc0, s0);
      box_muller_polar(o[2], o[3], c1, r1, cc1, s1);
      x = fmaf(x, fmaf(r0, cc0, c0), x); x = fmaf(x, fmaf(r0, s0, c0), x);
      x = fmaf(x, fmaf(r1, cc1, c0), x); x = fmaf(x, fmaf(r1, s1, c0), x);
    }
    if (MODE == 3) {                          // philox + only the MUFU-free part (uniform -> fma)
      x = fmaf(x, fmaf(bits_to_12(o[0]), c1, c0), x); x = fmaf(x, fmaf(bits_to_12(o[1]), c1, c0), x);
      x = fmaf(x, fmaf(bits_to_12(o[2]), c1, c0), x); x = fmaf(x, fmaf(bits_to_12(o[3]), c1, c0), x);
    }
    if (MODE == 4) {                          // philox + lg2/sqrt only (2 MUFU per pair instead of 4)
      float r0 = fast_sqrt(fast_lg2(bits_to_u01_open0(o[0])) * c1), r1 = fast_sqrt(fast_lg2(bits_to_u01_open0(o[2])) * c1);
      x = fmaf(x, fmaf(r0, bits_to_12(o[1]), c0), x); x = fmaf(x, fmaf(r1, bits_to_12(o[3]), c0), x);
    }
  }
  if (x == 12345.f || acc == 0x1234567u) out[tid] = x + acc;
}

template <int MODE> void run(const char* name, int iters) {
  PhiloxKeys keys = make_philox_keys(1234);
  float* out; cudaMalloc(&out, 4 << 20);
  int grid = 148 * 5;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(keys, iters, 1e-4f, -1e-4f, out);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(keys, iters, 1e-4f, -1e-4f, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double blocks = (double)grid * 256 * iters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc_per_warp_block = ms * 1e-3 * clk * 1e3 * 148 * 4 / (blocks / 32);
  printf("%-28s %8.3f ms  %.3e blocks(4 words)/s  %.1f SMSP-cycles per warp-block\n", name, ms, blocks / (ms * 1e-3), cyc_per_warp_block);
}
int main() {
  int iters = 4000;
  run<0>("philox only", iters);
  run<1>("box-muller+euler only", iters);
  run<2>("philox+box-muller+euler", iters);
  run<3>("philox+uniform fma", iters);
  run<4>("philox+lg2/sqrt only", iters);
  return 0;
}
