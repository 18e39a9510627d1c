// store_tile.cuh -- coalesced trajectory output for the path-storing kernels (the solve() contract).
//
// The reference layouts are row-per-path: (bs, S+1, dim) etc. (solvers.py:64-66,150-162).  With one thread per path a
// direct store touches 32 different rows per warp instruction (4 useful bytes per 32-byte sector).  Instead every
// warp stages TILE consecutive elements of its 32 paths in shared memory ([element][lane], padded to 33 columns so
// both the per-lane appends and the transposed reads are bank-conflict free) and flushes them path by path: 32 lanes
// write TILE*4 contiguous bytes of one row.  All lanes of a warp append in lock-step (same element index), which the
// storing kernels guarantee.
#pragma once
#include <cstdint>

namespace sdemc {

template <int TILE>
struct WarpTileWriter {
  static constexpr int kFloats = TILE * 33;  // shared floats per warp
  float* tile;        // this warp's staging tile
  float* rows;        // global base of the array
  uint64_t row_len;   // floats per path row
  uint64_t first_row; // path index handled by lane 0 of this warp
  uint64_t n_rows;    // rows that exist (paths in the call)
  uint64_t pos;       // elements of the row already flushed
  int cnt;            // elements staged (warp-uniform)

  __device__ __forceinline__ void init(float* smem_tile, float* global, uint64_t row_len_, uint64_t first_row_,
                                       uint64_t n_rows_) {
    tile = smem_tile;
    rows = global;
    row_len = row_len_;
    first_row = first_row_;
    n_rows = n_rows_;
    pos = 0;
    cnt = 0;
  }
  __device__ __forceinline__ void flush() {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (rows != nullptr) {
      if (TILE == 32) {
        for (int p = 0; p < 32; ++p) {
          const uint64_t r = first_row + p;
          if (r < n_rows && lane < cnt) rows[r * row_len + pos + lane] = tile[lane * 33 + p];
        }
      } else {
        // TILE == 16: two paths per instruction (lanes 0-15 serve path p, lanes 16-31 path p + 16)
        const int e = lane & 15;
        for (int p = 0; p < 16; ++p) {
          const int q = p + (lane >> 4) * 16;
          const uint64_t r = first_row + q;
          if (r < n_rows && e < cnt) rows[r * row_len + pos + e] = tile[e * 33 + q];
        }
      }
    }
    __syncwarp();
    pos += cnt;
    cnt = 0;
  }
  // every lane appends the next element of ITS path; all lanes call this together
  __device__ __forceinline__ void append(float v) {
    tile[cnt * 33 + (threadIdx.x & 31)] = v;
    ++cnt;
    if (cnt == TILE) flush();
  }
};

}  // namespace sdemc
