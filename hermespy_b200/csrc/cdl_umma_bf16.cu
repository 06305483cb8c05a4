// Explicit instantiations and launcher of the BF16x3 tensor-core CDL kernel (cdl_umma_bf16.cuh).
#include <algorithm>

#include "cdl_umma_bf16.cuh"

namespace hb {

template <int NRX, int P, typename IO>
static int cb_one(const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  auto kern = cdl_umma_bf16_kernel<NRX, P, IO>;
  HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfileScope prof(KIND_CDL_PROPAGATE, st);
  const long long nitems = (long long)a.B * a.ntiles;
  const int grid = (int)std::min<long long>(nitems, persistent_sm_count());
  kern<<<grid, kCuThreads, smem, st>>>(a, tb);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NRX, typename IO>
static int cb_p(int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  if constexpr (NRX <= 4) {
    switch (P) {
      case 1: return cb_one<NRX, 1, IO>(a, tb, smem, st);
      case 2: return cb_one<NRX, 2, IO>(a, tb, smem, st);
      case 3: return cb_one<NRX, 3, IO>(a, tb, smem, st);
      default: return cb_one<NRX, 4, IO>(a, tb, smem, st);
    }
  } else {
    switch (P) {
      case 1: return cb_one<NRX, 1, IO>(a, tb, smem, st);
      case 2: return cb_one<NRX, 2, IO>(a, tb, smem, st);
    }
    set_error("BF16x3 CDL kernel: 8 receive antennas only up to two Taylor terms");
    return HB_ERR_UNSUPPORTED;
  }
}

template <typename IO>
static int cb_nrx(int nrx_tpl, int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  switch (nrx_tpl) {
    case 1: return cb_p<1, IO>(P, a, tb, smem, st);
    case 2: return cb_p<2, IO>(P, a, tb, smem, st);
    case 4: return cb_p<4, IO>(P, a, tb, smem, st);
    default: return cb_p<8, IO>(P, a, tb, smem, st);
  }
}

int launch_cdl_umma_bf16(int nrx_tpl, int P, bool io128, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st) {
  return io128 ? cb_nrx<double2>(nrx_tpl, P, a, tb, smem, st) : cb_nrx<float2>(nrx_tpl, P, a, tb, smem, st);
}

}  // namespace hb
