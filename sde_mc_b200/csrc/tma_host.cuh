// tma_host.cuh -- host side of the TMA path-storing kernels: tensor maps over row-per-path arrays (internal, C++)
#pragma once
#include <cudaTypedefs.h>

#include "tma_gang.cuh"

namespace sdemc {
namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// (n_rows, row_len) fp32 array with `pitch` floats between rows -> 2-D tensor map with a [32 rows][32 elements]
// box and the 128-byte swizzle of the staging tiles (diffusion_tma.cuh)
bool make_row_map(CUtensorMap* map, float* base, uint64_t n_rows, uint64_t row_len, uint64_t pitch) {
  auto encode = tensor_map_encoder();
  if (!encode) return false;
  const cuuint64_t dims[2] = {row_len, n_rows};
  const cuuint64_t strides[1] = {pitch * sizeof(float)};
  const cuuint32_t box[2] = {kTmaTileElems, 32};
  const cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                kTmaTileElems == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// row length the tensor map declares: the pitch when the rows are padded to whole tiles (the padding belongs to the
// allocation, sdemc_paths_out), else the row itself
inline uint64_t tma_map_row_len(uint64_t row_len, uint64_t pitch) {
  return (pitch >= row_len && pitch % kTmaTileElems == 0) ? pitch : row_len;
}

// What a kernel needs to know about rows of `elems` elements: the declared length and, for a short last tile (at most
// 16 elements), the columns it writes with direct stores, whole 32-byte sectors (tma_gang.cuh)
inline TmaRows tma_rows_of(uint64_t elems, uint64_t len, uint64_t pitch, bool allow_direct = true) {
  TmaRows r{(int)len, 0x7fffffff, 0};
  const uint64_t tail = elems % kTmaTileElems;
  if (allow_direct && tail > 0 && tail <= 16 && pitch % kTmaTileElems == 0 && pitch >= elems) {
    r.dcol = (int)(elems - tail);
    r.dend = (int)((elems + 7) / 8 * 8);
  }
  return r;
}
// arrays that share a gang switch to direct stores together or not at all
inline void tma_rows_agree(TmaRows& a, TmaRows& b) {
  if (a.dcol != b.dcol) {
    a.dcol = b.dcol = 0x7fffffff;
    a.dend = b.dend = 0;
  }
}

// the two scheduling words of the storing kernels inside the caller's workspace (engine.cuh: bytes [16, 24) of the
// 64-byte header; zero between calls).  Without a workspace the TMA kernels are not used.
inline unsigned int* tma_sched_words(void* d_ws) {
  return d_ws ? reinterpret_cast<unsigned int*>(static_cast<char*>(d_ws) + 16) : nullptr;
}

inline bool tma_rows_ok(const float* base, uint64_t pitch) {
  return base != nullptr && (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0;
}

}  // namespace
}  // namespace sdemc
