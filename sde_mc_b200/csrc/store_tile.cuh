// store_tile.cuh -- coalesced, vectorised trajectory output for the path-storing kernels (the solve() contract).
//
// The reference layouts are row-per-path: (bs, S+1, dim) etc. (solvers.py:64-66,150-162).  With one thread per path a
// direct store touches 32 different rows per warp instruction (4 useful bytes per 32-byte sector).  Instead every
// warp stages TILE consecutive elements of its 32 paths in shared memory, element-major ([element][lane], row stride
// 32 + 32/TILE words), and flushes the tile transposed:
//   * vector path (row pitch a multiple of 4 floats, 16-byte aligned base -- the Python layer pads the pitch to 128
//     bytes): TILE/4 lanes cover one path's TILE*4 contiguous bytes with one 16-byte store each, 128/TILE paths per
//     warp instruction.  The 4 shared loads per store are bank-conflict free: element 4j+e of path q sits in bank
//     (4j+e)*(32/TILE) + q (mod 32), distinct over the (j, q) pairs of one instruction.
//   * scalar path (any pitch, partial tiles, the last rows of a call): one 4-byte store per lane, TILE contiguous
//     elements of one path per group of TILE lanes.
// All lanes of a warp append in lock-step (same element index), which the storing kernels guarantee.
#pragma once
#include <cstdint>

namespace sdemc {

template <int TILE>
struct TileGeom {
  static_assert(TILE == 8 || TILE == 16 || TILE == 32, "tile of 8, 16 or 32 elements");
  static constexpr int kStride = 32 + 32 / TILE;  // words between consecutive elements of the tile
  static constexpr int G = TILE / 4;              // 16-byte granules per path and tile
  static constexpr int R = 32 / G;                // paths per vector store instruction
};

// Scalar flush (any pitch, partial tiles, last rows of a call): writes the first n staged elements of the warp's 32
// paths with 4-byte stores, TILE lanes per path.  Out of line: rare, one copy per kernel.
// tile: this warp's staging tile.  rows: global base of the array, pitch: floats between rows,
// first_row: row of lane 0, n_rows: rows that exist, pos: elements of each row already written.
template <int TILE>
__device__ __noinline__ void flush_tile_scalar(const float* tile, float* rows, uint64_t pitch, uint64_t first_row,
                                               uint64_t n_rows, uint64_t pos, int n) {
  using Geo = TileGeom<TILE>;
  constexpr int PPI = 32 / TILE;  // paths per store instruction
  const int lane = threadIdx.x & 31;
  const int e = lane % TILE, q0 = lane / TILE;
  for (int p = 0; p < 32; p += PPI) {
    const int q = p + q0;
    const uint64_t r = first_row + q;
    if (r < n_rows && e < n) rows[r * pitch + pos + e] = tile[e * Geo::kStride + q];
  }
}

// All lanes of a warp stage in lock-step.  Hot loops stage a group of elements at compile-time slots and commit
// them once (one shared store per element plus a handful of instructions per group); append() is stage + commit
// of a single element.  SLACK: the largest number of elements one step group stages (a group may start with
// TILE - 1 elements staged).
template <int TILE, int SLACK>
struct WarpTileWriter {
  using Geo = TileGeom<TILE>;
  static constexpr int kCap = TILE + SLACK;               // staged elements the tile can hold
  static constexpr int kFloats = kCap * Geo::kStride;     // shared floats per warp and array
  float* tile;       // this warp's staging tile (shared memory)
  float* wp;         // this lane's slot for the next element
  const float* vsrc; // vector flush: this lane's first element (granule j of path q0)
  float* rows;       // global base of the array (nullptr: output not requested)
  float* vdst;       // vector flush: this lane's destination in the first group of rows, advanced by every flush
  uint64_t pitch;    // floats between consecutive path rows
  uint64_t first_row;  // path index handled by lane 0 of this warp
  uint64_t n_rows;   // rows that exist (paths in the call)
  uint64_t pos;      // elements of the row already flushed
  int cnt;           // elements staged (warp-uniform)
  bool vec;          // 16-byte store path usable for full tiles (warp-uniform)

  __device__ __forceinline__ void init(float* smem_tile, float* global, uint64_t pitch_, uint64_t first_row_,
                                       uint64_t n_rows_) {
    const int lane = threadIdx.x & 31;
    tile = smem_tile;
    wp = smem_tile + lane;
    rows = global;
    pitch = pitch_;
    first_row = first_row_;
    n_rows = n_rows_;
    pos = 0;
    cnt = 0;
    vec = global != nullptr && first_row_ + 32 <= n_rows_ && (pitch_ & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(global) & 15) == 0;
    const int j = lane % Geo::G, q0 = lane / Geo::G;
    vsrc = smem_tile + (4 * j) * Geo::kStride + q0;
    vdst = global + (first_row_ + q0) * pitch_ + 4 * j;
  }
  // element i (< SLACK) of the current group
  __device__ __forceinline__ void stage(int i, float v) { wp[i * Geo::kStride] = v; }

  // full tile -> global.  Vector path: TILE/4 lanes cover one path's TILE*4 contiguous bytes with a 16-byte store
  // each, 128/TILE paths per instruction; rows are reached through `pitch` only.
  __device__ __forceinline__ void flush_full() {
    __syncwarp();
    if (vec) {
      float* dst = vdst;
      const uint64_t dst_step = (uint64_t)Geo::R * pitch;
#pragma unroll
      for (int i = 0; i < Geo::G; ++i) {
        float4 v;
        v.x = vsrc[i * Geo::R + 0 * Geo::kStride];
        v.y = vsrc[i * Geo::R + 1 * Geo::kStride];
        v.z = vsrc[i * Geo::R + 2 * Geo::kStride];
        v.w = vsrc[i * Geo::R + 3 * Geo::kStride];
        *reinterpret_cast<float4*>(dst) = v;
        dst += dst_step;
      }
      vdst += TILE;
    } else if (rows != nullptr) {
      flush_tile_scalar<TILE>(tile, rows, pitch, first_row, n_rows, pos, TILE);
    }
    __syncwarp();
    pos += TILE;
    cnt -= TILE;
    // every lane moves the leftovers of its own column to the front of the tile
    float* col = tile + (threadIdx.x & 31);
#pragma unroll 1
    for (int k = 0; k < cnt; ++k) col[k * Geo::kStride] = col[(k + TILE) * Geo::kStride];
    wp = col + cnt * Geo::kStride;
  }
  // the first n staged elements of the group become part of the row
  __device__ __forceinline__ void commit(int n) {
    cnt += n;
    wp += n * Geo::kStride;
    if (cnt >= TILE) flush_full();
  }
  __device__ __forceinline__ void append(float v) {
    stage(0, v);
    commit(1);
  }
  // end of the row: write what is left (partial tile, scalar stores)
  __device__ __forceinline__ void flush() {
    __syncwarp();
    if (cnt > 0 && rows != nullptr) flush_tile_scalar<TILE>(tile, rows, pitch, first_row, n_rows, pos, cnt);
    __syncwarp();
    pos += cnt;
    cnt = 0;
    wp = tile + (threadIdx.x & 31);
  }
};

}  // namespace sdemc
