// launch_diffusion.cu -- instantiation + dispatch of the uniform-grid kernels (diffusion.cuh)
#include "diffusion.cuh"
#include "diffusion_tma.cuh"
#include "launch.cuh"
#include "tma_host.cuh"

namespace sdemc {
namespace {

// path-storing launch through TMA (diffusion_tma.cuh); returns 1 when the layout does not qualify
template <class C, bool HESTON, bool INJECT>
int run_store_tma(const LaunchArgs& a) {
  constexpr int NPS = C::BASE * C::M + (C::ASIAN ? 1 : 0);
  if (!tma_rows_ok(a.out.paths, a.out.pitch_state) || !tma_rows_ok(a.out.normals, a.out.pitch_normals)) return 1;
  if (a.range.n_paths >= (1ull << 31)) return 1;
  if (a.no_tma || a.d_ws == nullptr) return 1;  // (the TMA kernel hands its warp tasks out through the workspace)
  CUtensorMap mp, mn;
  const uint64_t S = (uint64_t)a.sde.num_steps;
#ifdef SDEMC_TMA_CLIP_AT_ROW_END
  const uint64_t len_p = (S + 1) * C::DIM, len_n = S * NPS;
#else
  // The maps span the whole pitch, not just the row: a box the tensor bound cuts inside a row costs the engine
  // several times a full box (measured), so the last tile of a row spills into the row's padding
  const uint64_t len_p = tma_map_row_len((S + 1) * C::DIM, a.out.pitch_state), len_n = tma_map_row_len(S * NPS, a.out.pitch_normals);
#endif
  if (!make_row_map(&mp, a.out.paths, a.range.n_paths, len_p, a.out.pitch_state)) return 1;
  if (!make_row_map(&mn, a.out.normals, a.range.n_paths, len_n, a.out.pitch_normals)) return 1;
  TmaRows rp = tma_rows_of((S + 1) * C::DIM, len_p, a.out.pitch_state), rn = tma_rows_of(S * NPS, len_n, a.out.pitch_normals);
  if (C::DIM == NPS && C::DIM != 4) tma_rows_agree(rp, rn);  // one gang (diffusion_tma.cuh)
  auto kernel = diffusion_store_tma_kernel<C, HESTON, INJECT>;
  using Gang = TmaGang<2, kDiffTmaW>;
  const size_t smem = (size_t)(kTmaStoreBlock / 32) * Gang::kBytes + Gang::kAlign;
  SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 0;
  int rc = pick_grid(kernel, smem, a.range.n_paths, &grid, kTmaStoreBlock);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, kTmaStoreBlock, smem, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, a.out, mp, mn, rp, rn, tma_sched_words(a.d_ws));
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

template <class C, bool HESTON, bool INJECT, bool STORE, bool PERPATH = false, int RMODE = RANGE_HOST>
int run(const LaunchArgs& a) {
  auto kernel = diffusion_kernel<C, HESTON, INJECT, STORE, PERPATH, RMODE>;
  const int block = STORE ? kDiffusionStoreBlock : kBlock;
  const size_t smem = STORE ? (size_t)(block / 32) * 2 * DiffusionStoreWriter<C>::kFloats * sizeof(float) : 0;
  if (smem > 48 * 1024)
    SDEMC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = 0;
  int rc = pick_grid(kernel, smem, a.range.n_paths, &grid, block);
  if (rc != SDEMC_OK) return rc;
  kernel<<<grid, block, smem, a.stream>>>(a.sde, a.payoff, a.range, a.keys, a.inject, a.out, a.d_moments, a.d_ws);
  SDEMC_CUDA_CHECK(cudaGetLastError());
  return SDEMC_OK;
}

template <class C, bool HESTON>
int by_mode(const LaunchArgs& a) {
  if (a.store) {
    // 16-byte aligned padded rows leave the SM through TMA; any other layout takes the store_tile.cuh kernel
    const int rc = a.use_inject ? run_store_tma<C, HESTON, true>(a) : run_store_tma<C, HESTON, false>(a);
    if (rc <= 0) return rc;
    return a.use_inject ? run<C, HESTON, true, true>(a) : run<C, HESTON, false, true>(a);
  }
  if (a.use_inject) return SDEMC_ERR_UNSUPPORTED;  // injected noise is only offered with stored outputs
  if (per_path_of_out(a.out).any()) return run<C, HESTON, false, false, true>(a);
  if (a.range.dyn) return run<C, HESTON, false, false, false, RANGE_DEVICE>(a);   // range read from device memory
  return run<C, HESTON, false, false>(a);
}

template <int FAMILY, int M>
int by_dim(const sdemc_sde& s, const LaunchArgs& a) {
  switch (s.dim) {
    case 1: return by_mode<Cfg<FAMILY, 1, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 2: return by_mode<Cfg<FAMILY, 2, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 3: return by_mode<Cfg<FAMILY, 3, M, SDEMC_MARKS_NONE, false>, false>(a);
    case 4: return by_mode<Cfg<FAMILY, 4, M, SDEMC_MARKS_NONE, false>, false>(a);
  }
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace

int launch_diffusion(const sdemc_sde& s, const LaunchArgs& a) {
  if (s.family == SDEMC_FAMILY_HESTON) {
    if (s.dim != 2 || s.m != 1 || s.asian) return SDEMC_ERR_UNSUPPORTED;
    return by_mode<Cfg<SDEMC_FAMILY_HESTON, 2, 1, SDEMC_MARKS_NONE, false>, true>(a);
  }
  if (s.asian) {
    if (s.dim != 2 || s.m != 1) return SDEMC_ERR_UNSUPPORTED;
    if (s.family == SDEMC_FAMILY_GEOMETRIC)
      return by_mode<Cfg<SDEMC_FAMILY_GEOMETRIC, 2, 1, SDEMC_MARKS_NONE, true>, false>(a);
    return by_mode<Cfg<SDEMC_FAMILY_ARITHMETIC, 2, 1, SDEMC_MARKS_NONE, true>, false>(a);
  }
  if (s.family == SDEMC_FAMILY_GEOMETRIC) return s.m == 1 ? by_dim<SDEMC_FAMILY_GEOMETRIC, 1>(s, a) : by_dim<SDEMC_FAMILY_GEOMETRIC, 2>(s, a);
  if (s.family == SDEMC_FAMILY_ARITHMETIC) return s.m == 1 ? by_dim<SDEMC_FAMILY_ARITHMETIC, 1>(s, a) : by_dim<SDEMC_FAMILY_ARITHMETIC, 2>(s, a);
  return SDEMC_ERR_UNSUPPORTED;
}

}  // namespace sdemc
