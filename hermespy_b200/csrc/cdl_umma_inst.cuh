// Explicit instantiations of the tensor-core CDL kernel for one I/O element type (separate TUs compile in parallel).
#pragma once
#include <algorithm>

#include "cdl_umma.cuh"

namespace hb {

template <int NRX, int P, typename IO>
static int launch_cu_one(const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  auto kern = cdl_umma_kernel<NRX, P, IO>;
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfileScope prof(KIND_CDL_PROPAGATE, st);
  const long long nitems = (long long)a.B * a.ntiles;
  const int grid = (int)std::min<long long>(nitems, persistent_sm_count());
  kern<<<grid, kCuThreads, smem, st>>>(a, tb);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NRX, typename IO>
static int launch_cu_p(int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  switch (P) {
    case 1: return launch_cu_one<NRX, 1, IO>(a, tb, smem, st);
    case 2: return launch_cu_one<NRX, 2, IO>(a, tb, smem, st);
    case 3: return launch_cu_one<NRX, 3, IO>(a, tb, smem, st);
    default: return launch_cu_one<NRX, 4, IO>(a, tb, smem, st);
  }
}

template <typename IO>
int launch_cdl_umma_io(int nrx_tpl, int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  switch (nrx_tpl) {
    case 1: return launch_cu_p<1, IO>(P, a, tb, smem, st);
    case 2: return launch_cu_p<2, IO>(P, a, tb, smem, st);
    case 4: return launch_cu_p<4, IO>(P, a, tb, smem, st);
    default: return launch_cu_p<8, IO>(P, a, tb, smem, st);
  }
}

}  // namespace hb
