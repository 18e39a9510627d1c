// What does ISSUING a TMA tile store cost the issuing warp?  One lane issues N cp.async.bulk.tensor stores of 4 KB
// boxes back to back (different rows each time, fresh commit groups) and timestamps itself with clock64: cycles per
// issue with W warps per SM doing the same.  (tools/tma_two_map_probe.cu: with compute between copies the kernel time
// is compute + ~1200 cycles per copy and warp, not max(compute, copy).)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_issue_cost_probe.cu -o tools/tma_issue_cost_probe.bin
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
}
static bool make_map(CUtensorMap* map, float* base, unsigned long long n_rows, unsigned long long pitch) {
  const cuuint64_t dims[2] = {256, n_rows};
  const cuuint64_t strides[1] = {pitch * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// mode 0: issue + commit only; 1: + wait_group.read 8 (never blocks on recent copies); 2: gap cycles of dependent FMAs between issues
__global__ void __launch_bounds__(32) k(const __grid_constant__ CUtensorMap map, int n, int mode, int gap, long long* out, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned base = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  float acc = (float)lane;
  long long t_issue = 0, t_total0 = clock64();
  for (int i = 0; i < n; ++i) {
    for (int d = 0; d < gap; ++d) acc = fmaf(acc, 1.0001f, 0.5f);
    if (lane == 0) {
      const long long t0 = clock64();
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"((i & 7) * 32),
                   "r"((int)((blockIdx.x * (n / 8 + 1) + i / 8) * 32)), "r"(base) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (mode == 1) asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
      t_issue += clock64() - t0;
    }
    __syncwarp();
  }
  const long long t_total = clock64() - t_total0;
  if (lane == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    out[2 * blockIdx.x] = t_issue;
    out[2 * blockIdx.x + 1] = t_total;
  }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const long n_rows = 8000000;
  float *d, *sink;
  long long* out;
  cudaMalloc(&d, (size_t)n_rows * 256 * 4);
  cudaMalloc(&sink, 4);
  cudaMallocManaged(&out, 148 * 24 * 2 * sizeof(long long));
  CUtensorMap m;
  if (!make_map(&m, d, n_rows, 256)) { printf("map failed\n"); return 1; }
  const size_t smem = 4096 + 1024;
  const int n = 256;
  for (int warps : {1, 3, 12}) {
    for (int mode : {0, 1}) {
      for (int gap : {0, 500, 2000}) {
        k<<<148 * warps, 32, smem>>>(m, n, mode, gap, out, sink);
        cudaDeviceSynchronize();
        double si = 0, st = 0;
        for (int b = 0; b < 148 * warps; ++b) { si += out[2 * b]; st += out[2 * b + 1]; }
        printf("warps/SM %2d  %s  gap %4d FMA: cycles per copy in issue(+commit%s) %.0f, loop total %.0f (gap alone ~%d)  %s\n", warps,
               mode ? "wait.read 8" : "no wait    ", gap, mode ? "+wait" : "", si / (148.0 * warps * n), st / (148.0 * warps * n), gap * 4,
               cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
