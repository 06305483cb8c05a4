// Explicit instantiations and launcher of the single-antenna kernel (fading_siso.cuh).
#include <algorithm>

#include "fading_siso.cuh"

namespace hb {

template <int P, int KP, typename IO>
static int launch_siso_one(const FadingArgs& a, const DelayTable& dt, int poly_tile, int npoly, size_t smem, cudaStream_t st) {
  auto kern = tdl_siso_kernel<P, KP, IO>;
  if (smem > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSisoThreads, smem));
  const long long total = (long long)a.B * a.ntiles;
  const int grid = (int)std::min<long long>(total, (long long)std::max(1, per_sm) * persistent_sm_count());
  kern<<<grid, kSisoThreads, smem, st>>>(a, dt, poly_tile, npoly);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int P, typename IO>
static int launch_siso_kp(int tile, const FadingArgs& a, const DelayTable& dt, int poly_tile, int npoly, size_t smem, cudaStream_t st) {
  switch (tile) {
    case 512: return launch_siso_one<P, 1, IO>(a, dt, poly_tile, npoly, smem, st);
    case 1024: return launch_siso_one<P, 2, IO>(a, dt, poly_tile, npoly, smem, st);
    case 2048: return launch_siso_one<P, 4, IO>(a, dt, poly_tile, npoly, smem, st);
  }
  set_error("single-antenna kernel: tile %d outside {512, 1024, 2048}", tile);
  return HB_ERR_UNSUPPORTED;
}

template <typename IO>
static int launch_siso_p(int P, int tile, const FadingArgs& a, const DelayTable& dt, int poly_tile, int npoly, size_t smem,
                         cudaStream_t st) {
  switch (P) {
    case 1: return launch_siso_kp<1, IO>(tile, a, dt, poly_tile, npoly, smem, st);
    case 2: return launch_siso_kp<2, IO>(tile, a, dt, poly_tile, npoly, smem, st);
    case 3: return launch_siso_kp<3, IO>(tile, a, dt, poly_tile, npoly, smem, st);
    case 4: return launch_siso_kp<4, IO>(tile, a, dt, poly_tile, npoly, smem, st);
    case 6: return launch_siso_kp<6, IO>(tile, a, dt, poly_tile, npoly, smem, st);
    case 8: return launch_siso_kp<8, IO>(tile, a, dt, poly_tile, npoly, smem, st);
  }
  set_error("polynomial order %d outside the compiled set", P);
  return HB_ERR_UNSUPPORTED;
}

int launch_tdl_siso(int P, int tile, bool io128, const FadingArgs& a, const DelayTable& dt, int poly_tile, int npoly, size_t smem,
                    cudaStream_t st) {
  return io128 ? launch_siso_p<double2>(P, tile, a, dt, poly_tile, npoly, smem, st)
               : launch_siso_p<float2>(P, tile, a, dt, poly_tile, npoly, smem, st);
}

}  // namespace hb
