// Explicit instantiations and launchers of the float64 Taylor path (fading_poly64.cuh).
#include "fading_poly64.cuh"

namespace hb {

template <int P>
static int coef64_one(const FadingArgs& a, const DelayTable& dt, cudaStream_t st) {
  const size_t items = (size_t)a.ntiles * a.B * dt.num_groups;
  sos_poly_coef64_kernel<P><<<(unsigned)((items + 3) / 4), 128, 0, st>>>(a, dt);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

int launch_coef64(int P, const FadingArgs& a, const DelayTable& dt, cudaStream_t st) {
  switch (P) {
    case 4: return coef64_one<4>(a, dt, st);
    case 6: return coef64_one<6>(a, dt, st);
    case 8: return coef64_one<8>(a, dt, st);
  }
  set_error("float64 Taylor path: polynomial order %d outside {4, 6, 8}", P);
  return HB_ERR_UNSUPPORTED;
}

template <int NTX, int P, typename IO>
static int poly64_one(const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  constexpr int R = NTX <= 4 ? 2 : 1;
  auto kern = tdl_poly64_kernel<NTX, P, R, IO>;
  if (smem > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)((size_t)a.ntiles * a.B), kThreads, smem, st>>>(a, dt);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NTX, typename IO>
static int poly64_p(int P, const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  switch (P) {
    case 4: return poly64_one<NTX, 4, IO>(a, dt, smem, st);
    case 6: return poly64_one<NTX, 6, IO>(a, dt, smem, st);
    case 8: return poly64_one<NTX, 8, IO>(a, dt, smem, st);
  }
  set_error("float64 Taylor path: polynomial order %d outside {4, 6, 8}", P);
  return HB_ERR_UNSUPPORTED;
}

template <typename IO>
static int poly64_ntx(int ntx_tpl, int P, const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  switch (ntx_tpl) {
    case 1: return poly64_p<1, IO>(P, a, dt, smem, st);
    case 2: return poly64_p<2, IO>(P, a, dt, smem, st);
    case 4: return poly64_p<4, IO>(P, a, dt, smem, st);
    default: return poly64_p<8, IO>(P, a, dt, smem, st);
  }
}

int launch_tdl_poly64(int ntx_tpl, int P, bool io128, const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  return io128 ? poly64_ntx<double2>(ntx_tpl, P, a, dt, smem, st) : poly64_ntx<float2>(ntx_tpl, P, a, dt, smem, st);
}

}  // namespace hb
