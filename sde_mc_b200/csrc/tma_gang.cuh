// tma_gang.cuh -- staging tiles of the path-storing kernels that leave the SM through TMA (jump_tma.cuh,
// diffusion_tma.cuh).
//
// A warp owns 32 consecutive paths (rows of the row-per-path output arrays).  Per output array it keeps a tile of
// W sub-tiles of [32 rows][32 elements] in shared memory in the layout TMA reads (128-byte rows, 128B swizzle);
// every lane collects four consecutive elements of its row in registers and stages them with one 16-byte shared
// store; a full tile is handed to the engine by one lane as W cp.async.bulk.tensor.2d boxes issued back to back, so
// W x 128 consecutive bytes of every row reach the memory system together.
//
// Gangs.  Arrays that receive the same number of elements per step at the same phase fill their tiles in lock-step:
// they share ONE vector counter, one full-tile test and one wait, and a staging store is `tile address of the lane ^
// mask(counter)` -- the 128B swizzle of chunk v of row q is (v ^ (q & 7)) << 4, the sub-tile index sits in address
// bits the lane address leaves zero, and the counter only touches those bits: LOP3 + STS.128.
//
// Tiles are single-buffered: the wait for the engine's read of a tile (cp.async.bulk.wait_group.read) is deferred to
// the next store into it, which the kernels place several steps after the flush.  The copies issued at one staging
// point share one bulk group, committed lazily by the first wait that needs it.
#pragma once
#include <cuda.h>

#include <cstdint>

namespace sdemc {

constexpr int kTmaTileElems = 32;                      // elements per row of a sub-tile (128-byte rows, 128B swizzle)
constexpr int kTmaTileBytes = 32 * kTmaTileElems * 4;  // one [32][32] fp32 sub-tile

__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, uint32_t smem, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col),
               "r"(row), "r"(smem)
               : "memory");
}

constexpr int tma_gcd(int a, int b) { return b == 0 ? a : tma_gcd(b, a % b); }
constexpr int tma_lcm(int a, int b) { return a / tma_gcd(a, b) * b; }

// The rare parts of staging -- waiting for the engine, handing tiles over -- exist inline and as out-of-line functions
// with by-value arguments; a kernel picks one (template parameter OOL of the gang).  Inlined at every staging point
// of an unrolled step loop they put ~60 cold instructions and a taken branch over them every few steps into the hot
// path.  Measured: out of line the uniform-grid kernel gains 6 % (GBM solve() 1.54 -> 1.45 ms) and the jump kernel's
// instruction-fetch stalls fall from 1.3 to 0.3 warps per issue cycle.  (With a static partition of the paths the jump
// kernel was slower out of line -- its phase-locked warps reached the next store into a tile sooner and waited for
// the engine, 1.28 -> 1.41 ms; with the warp tasks handed out dynamically both forms take 1.11 ms and the smaller one
// is kept.)

// commits the open bulk group if there is one and waits until the engine has read the shared memory of every copy in
// groups 1 .. seq; returns the number of committed groups
__device__ __forceinline__ int tma_acquire_body(bool open, int committed, int seq) {
  const bool lane0 = (threadIdx.x & 31) == 0;
  if (open) {
    if (lane0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    ++committed;
  }
  const int allowed = committed - seq;  // newer groups that may stay pending (wait_group takes an immediate)
  if (lane0) {
    if (allowed >= 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    else if (allowed == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    else if (allowed == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
  return committed;
}
static __device__ __noinline__ int tma_acquire_ool(bool open, int committed, int seq) {
  return tma_acquire_body(open, committed, seq);
}

// Bulk-group bookkeeping of a warp.  Every lane carries the same values; only lane 0 talks to the engine.
struct TmaGroups {
  int committed;  // bulk groups committed so far
  bool open;      // copies issued since the last commit
  __device__ __forceinline__ void init() {
    committed = 0;
    open = false;
  }
  // returns once the engine has read the shared memory of every copy in groups 1 .. seq
  template <bool OOL>
  __device__ __forceinline__ void acquire(int seq) {
    committed = OOL ? tma_acquire_ool(open, committed, seq) : tma_acquire_body(open, committed, seq);
    open = false;
  }
  // before the CTA retires: everything written
  __device__ __forceinline__ void drain() {
    if ((threadIdx.x & 31) == 0) {
      if (open) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
};

// the arrays of a gang as the flush sees them
template <int N>
struct TmaGangArrays {
  const CUtensorMap* map[N];
  int row_len[N];
};
// hands the first `vec4 / 16` staged vectors of every array's tile to the engine: sub-tile after sub-tile of one
// array, array after array
template <int N, int W>
__device__ __forceinline__ void tma_flush_body(const TmaGangArrays<N>& arr, uint32_t tiles, int vec4, int col, int row0) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's staging stores -> visible to the TMA engine
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < N; ++a) {
#pragma unroll
      for (int w = 0; w < W; ++w)
        if (vec4 > w * 128 && col + 32 * w < arr.row_len[a])
          tma_store_tile(arr.map[a], tiles + (uint32_t)(a * W + w) * kTmaTileBytes, col + 32 * w, row0);
    }
  }
}
template <int N, int W>
__device__ __noinline__ void tma_flush_ool(TmaGangArrays<N> arr, uint32_t tiles, int vec4, int col, int row0) {
  tma_flush_body<N, W>(arr, tiles, vec4, col, row0);
}

// Which group of 32 rows a warp works on next: the groups are handed out one by one through a counter in the caller's
// workspace (requested one group ahead, so the atomic's round trip hides behind the step loop).  The storing kernels
// run ~47 groups per warp; with a static grid-stride partition the SMs finish up to a group apart
// (sm__cycles_active 95.7 % of elapsed in the jump kernel), handed out dynamically within a fraction of one
// (measured: GBM solve() 1.45 -> 1.35 ms, Merton full storage 1.28 -> 1.24 ms).  The last warp to finish re-arms the words, so the workspace stays zeroed
// between calls.  Which warp writes a group has no influence on what is written.  (STATIC = true: the grid-stride
// partition, A/B builds only -- a run-time switch makes the compiler clone the whole step loop.)
template <bool STATIC = false>
struct WarpTasks {
  unsigned int* sched;  // [0] next group, [1] warps done
  uint32_t n_groups, cur, stride;
  unsigned int nxt_raw;  // lane 0: the group after `cur`, requested one group ahead (its atomic may still be in flight)
  __device__ __forceinline__ void request() {
    if ((threadIdx.x & 31) == 0) nxt_raw = atomicAdd(&sched[0], 1u);
  }
  __device__ __forceinline__ void init(unsigned int* sched_, uint64_t n_rows) {
    sched = sched_;
    n_groups = (uint32_t)((n_rows + 31) / 32);
    stride = gridDim.x * (blockDim.x >> 5);
    nxt_raw = 0;
    if (STATIC) {
      cur = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    } else {
      request();
      cur = __shfl_sync(0xffffffffu, nxt_raw, 0);
      if (cur < n_groups) request();
    }
  }
  __device__ __forceinline__ bool valid() const { return cur < n_groups; }
  __device__ __forceinline__ uint64_t first_row() const { return (uint64_t)cur * 32u; }
  // after a group: take the one requested a group ago, request the one after it
  __device__ __forceinline__ void advance() {
    if (STATIC) {
      cur += stride;
    } else {
      cur = __shfl_sync(0xffffffffu, nxt_raw, 0);
      if (cur < n_groups) request();
    }
  }
  __device__ __forceinline__ void finish() {
    if (!STATIC && (threadIdx.x & 31) == 0) {
      __threadfence();
      if (atomicAdd(&sched[1], 1u) == stride - 1u) {  // every warp has stopped requesting: re-arm
        sched[0] = 0u;
        sched[1] = 0u;
      }
    }
  }
};
#ifdef SDEMC_TMA_STATIC_TASKS
using TmaWarpTasks = WarpTasks<true>;
#else
using TmaWarpTasks = WarpTasks<false>;
#endif

// what the host tells the kernel about the rows of one gang of arrays
struct TmaRows {
  int len;   // columns the tensor maps declare (the pitch when rows are padded to whole tiles)
  int dcol;  // first column written with direct stores instead of a tile (INT_MAX: none); a multiple of 32
  int dend;  // end of the directly written columns (the row's elements rounded up to whole 32-byte sectors)
};

// N output arrays of one warp that fill in lock-step: one single-buffered tile of W swizzled [32][32] sub-tiles each,
// consecutive in shared memory (the first on a W x 4 KB boundary), filled four elements per array at a time.  OOL: the
// rare staging code (wait, flush) as out-of-line calls.
template <int N, int W = 1, bool OOL = false>
struct TmaGang {
  static_assert(W == 1 || W == 2 || W == 4, "1, 2 or 4 sub-tiles");
  static constexpr uint32_t kTileBytes = (uint32_t)W * kTmaTileBytes;
  static constexpr uint32_t kBytes = (uint32_t)N * kTileBytes;  // shared memory per warp
  static constexpr uint32_t kAlign = W == 1 ? 1024u : kTileBytes;  // alignment of the first tile in the shared window
  uint32_t tiles;         // shared-window address of the first tile
  uint32_t lane_addr[N];  // this lane's chunk 0 of its row in every tile, swizzle applied: + row * 128 + ((row & 7) << 4)
  const CUtensorMap* map[N];
  int row_len[N];         // columns the tensor map of every array declares
  int dcol;               // from this column on the rows leave through 32-byte sector stores (short last tile;
                          // INT_MAX: none) -- a property of the gang, the host only enables it when all arrays agree
  int dend[N];            // end of the directly written columns of every array
  int seq;                // bulk group of the last copies out of the tiles (0: none pending)
  int vec4;               // 16-byte vectors staged in the tiles, times 16 (warp-uniform)
  int col;                // first column of the tiles
  int row0;               // row of lane 0

  // dcol_: the gang's switch to direct stores; every array is then described with set_array
  __device__ __forceinline__ void init(uint32_t tiles_s, int dcol_) {
    const uint32_t q = threadIdx.x & 31;
    tiles = tiles_s;
#pragma unroll
    for (int a = 0; a < N; ++a) lane_addr[a] = tiles_s + (uint32_t)a * kTileBytes + q * 128u + ((q & 7u) << 4);
    dcol = dcol_;
    seq = 0;
    vec4 = 0;
    col = 0;
    row0 = 0;
  }
  __device__ __forceinline__ void set_array(int a, const CUtensorMap* m, const TmaRows& rows) {
    map[a] = m;
    row_len[a] = rows.len;
    dend[a] = rows.dend;
  }
  __device__ __forceinline__ void begin_rows(int first_row) {
    col = 0;
    row0 = first_row;
  }
  // before the stores of a vector: the first store into tiles the engine may still be reading waits for it
  __device__ __forceinline__ void begin(TmaGroups& g) {
    if (vec4 == 0 && seq != 0) {
      g.template acquire<OOL>(seq);
      seq = 0;
    }
  }
  // four consecutive elements of this lane's path in array a
  __device__ __forceinline__ void store(int a, float x0, float x1, float x2, float x3) {
    // chunk bits (4..6) swizzled by XOR; the sub-tile index (vec4 >> 7) moves to bits 12.., zero in lane_addr
    const uint32_t mask = W == 1 ? (uint32_t)vec4 : (((uint32_t)vec4 & 0x70u) | (((uint32_t)vec4 & ~0x7fu) << 5));
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lane_addr[a] ^ mask), "f"(x0), "f"(x1), "f"(x2),
                 "f"(x3)
                 : "memory");
  }
  // The short last tile of a row (at most 16 elements: 134 = 4 x 32 + 6 slots for 100 nominal steps) does not go
  // through the engine: a box costs it the same ~185 cycles whether 6 or 32 of its columns exist, and the kernels are
  // bound by its box rate.  The columns are staged in the tile like any others (end() does not flush once the tile's
  // first column has reached dcol) and at the end of the row every lane reads its own row back and writes its last one or two 32-byte sectors with one 32-byte store
  // each (write_tail).  Whole sectors, one request per sector: 16-byte stores issued as the vectors appear were
  // measured 6-20 % SLOWER than the tile (two half-sector writes per sector).  The surplus lands in the row's padding.
  __device__ __forceinline__ bool direct() const { return col >= dcol; }
  // end of the row: the staged tail columns of array a (this lane's own stores: no synchronisation needed)
  __device__ __forceinline__ void write_tail(int a, float* base, uint64_t pitch, bool row_ok) {
    float* dst = base + (uint64_t)(row0 + (int)(threadIdx.x & 31)) * pitch + col;
#pragma unroll
    for (int j = 0; j < 2; ++j) {  // sector j = vectors 2j, 2j + 1 of the tile
      if (row_ok && col + 8 * j < dend[a]) {
        float4 lo, hi;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w)
                     : "r"(lane_addr[a] ^ (uint32_t)(32 * j)));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                     : "r"(lane_addr[a] ^ (uint32_t)(32 * j + 16)));
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * j), "f"(lo.x), "f"(lo.y),
                     "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                     : "memory");
      }
    }
  }
  // hands the staged columns to the engine
  __device__ __forceinline__ void flush(TmaGroups& g) {
    TmaGangArrays<N> arr;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      arr.map[a] = map[a];
      arr.row_len[a] = row_len[a];
    }
    if (OOL) tma_flush_ool<N, W>(arr, tiles, vec4, col, row0);
    else tma_flush_body<N, W>(arr, tiles, vec4, col, row0);
    g.open = true;
    seq = g.committed + 1;  // the group the next commit closes
    col += vec4 >> 2;       // (a partial flush only happens at the end of a row or in front of the tail columns)
    vec4 = 0;
  }
  // after the stores of a vector.  A full tile is flushed -- unless it holds the tail columns (direct()): those wait
  // for write_tail, and the surplus vectors a kernel stages past the end of the row (at most SG / 4 + 1, the tail has
  // at most four) pile up in the tile's last slot.
  __device__ __forceinline__ void end(TmaGroups& g) {
    vec4 += 16;
    if (vec4 == W * 128 || (W > 1 && col + (vec4 >> 2) == dcol)) {
      if (direct()) vec4 -= 16;
      else flush(g);
    }
  }
  // end of the rows of this group of 32 paths (tail columns: write_tail of every array first)
  __device__ __forceinline__ void finish(TmaGroups& g) {
    if (vec4 > 0 && !direct()) flush(g);
    vec4 = 0;
  }
};

}  // namespace sdemc
