// Explicit instantiations of the TMA-pipelined window kernel for one NTX (separate TUs compile in parallel).
#pragma once
#include "fading_inst.cuh"
#include "fading_tma.cuh"

namespace hb {

template <int NTX, int P, bool LIN, bool ZMODE>
static int launch_tma_z(const FadingArgs& a, const TmaPlan& tp, const CUtensorMap& xmap, int grid, size_t smem,
                        cudaStream_t st) {
  auto kern = tdl_tma_kernel<NTX, P, LIN, ZMODE>;
  if (int e = ensure_smem(kern, smem)) return e;
  kern<<<grid, kTmaThreads, smem, st>>>(a, tp, xmap);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

// z mode (large arrays, spatial product on the tensor cores) always runs in chunks of 4 antennas
template <int NTX, int P, bool LIN>
static int launch_tma_one(const FadingArgs& a, const TmaPlan& tp, const CUtensorMap& xmap, int grid, size_t smem,
                          cudaStream_t st) {
  if constexpr (NTX == 4) {
    if (a.z_mode) return launch_tma_z<NTX, P, LIN, true>(a, tp, xmap, grid, smem, st);
  }
  if (a.z_mode) {
    set_error("z mode is compiled for 4-antenna chunks only");
    return HB_ERR_UNSUPPORTED;
  }
  return launch_tma_z<NTX, P, LIN, false>(a, tp, xmap, grid, smem, st);
}

template <int NTX>
int launch_tdl_tma(int P, bool lin, bool /*zmode: carried by a.z_mode*/, const FadingArgs& a, const TmaPlan& tp, const CUtensorMap& xmap, int grid,
                   size_t smem, cudaStream_t st) {
  switch (P) {
    case 1: return launch_tma_one<NTX, 1, false>(a, tp, xmap, grid, smem, st);
    case 2: return launch_tma_one<NTX, 2, false>(a, tp, xmap, grid, smem, st);
    case 3:
      return lin ? launch_tma_one<NTX, 3, true>(a, tp, xmap, grid, smem, st)
                 : launch_tma_one<NTX, 3, false>(a, tp, xmap, grid, smem, st);
    case 4:
      return lin ? launch_tma_one<NTX, 4, true>(a, tp, xmap, grid, smem, st)
                 : launch_tma_one<NTX, 4, false>(a, tp, xmap, grid, smem, st);
    case 5: case 6: return launch_tma_one<NTX, 6, false>(a, tp, xmap, grid, smem, st);
    case 7: case 8: return launch_tma_one<NTX, 8, false>(a, tp, xmap, grid, smem, st);
    default: set_error("polynomial order %d outside the compiled set", P); return HB_ERR_UNSUPPORTED;
  }
}

#define HB_INSTANTIATE_FADING_TMA(NTX) \
  template int launch_tdl_tma<NTX>(int, bool, bool, const FadingArgs&, const TmaPlan&, const CUtensorMap&, int, size_t, cudaStream_t);

}  // namespace hb
