// Explicit instantiations of the fading kernels for one NTX (compiled as separate TUs in parallel).
#pragma once
#include "fading_window.cuh"

namespace hb {

template <typename Kern>
static int ensure_smem(Kern kern, size_t smem) {
  if (smem > 48 * 1024) {
    HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  return HB_OK;
}

template <int NTX, int P, typename IO>
static int launch_poly_one(const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  constexpr int R = poly_samples_per_thread<NTX>();
  auto kern = tdl_poly_kernel<NTX, P, R, IO>;
  if (int e = ensure_smem(kern, smem)) return e;
  dim3 grid((unsigned)((size_t)a.ntiles * a.B));
  kern<<<grid, kThreads, smem, st>>>(a, dt);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NTX, typename IO>
static int launch_poly_io(int P, const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st) {
  switch (P) {
    case 1: return launch_poly_one<NTX, 1, IO>(a, dt, smem, st);
    case 2: return launch_poly_one<NTX, 2, IO>(a, dt, smem, st);
    case 3: return launch_poly_one<NTX, 3, IO>(a, dt, smem, st);
    case 4: return launch_poly_one<NTX, 4, IO>(a, dt, smem, st);
    case 5: case 6: return launch_poly_one<NTX, 6, IO>(a, dt, smem, st);
    case 7: case 8: return launch_poly_one<NTX, 8, IO>(a, dt, smem, st);
    default: set_error("polynomial order %d outside the compiled set", P); return HB_ERR_UNSUPPORTED;
  }
}

template <int NTX>
int launch_tdl_poly(int P, bool io128, const FadingArgs& a, const DelayTable& dt, size_t smem,
                    cudaStream_t st) {
  return io128 ? launch_poly_io<NTX, double2>(P, a, dt, smem, st)
               : launch_poly_io<NTX, float2>(P, a, dt, smem, st);
}

template <int NTX, int P, int HALO, bool LIN, typename IO>
static int launch_window_one(const FadingArgs& a, const WindowPlan& wp, int threads, size_t smem,
                             cudaStream_t st) {
  constexpr int R = window_samples_per_thread<NTX>();
  auto kern = tdl_window_kernel<NTX, P, R, HALO, LIN, window_split<NTX>(), IO>;
  if (int e = ensure_smem(kern, smem)) return e;
  dim3 grid((unsigned)((size_t)a.ntiles * a.B));
  kern<<<grid, threads, smem, st>>>(a, wp);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NTX, int HALO, typename IO>
static int launch_window_io(int P, bool lin, const FadingArgs& a, const WindowPlan& wp, int threads, size_t smem,
                            cudaStream_t st) {
  switch (P) {
    case 1: return launch_window_one<NTX, 1, HALO, false, IO>(a, wp, threads, smem, st);
    case 2: return launch_window_one<NTX, 2, HALO, false, IO>(a, wp, threads, smem, st);
    case 3:
      return lin ? launch_window_one<NTX, 3, HALO, true, IO>(a, wp, threads, smem, st)
                 : launch_window_one<NTX, 3, HALO, false, IO>(a, wp, threads, smem, st);
    case 4:
      return lin ? launch_window_one<NTX, 4, HALO, true, IO>(a, wp, threads, smem, st)
                 : launch_window_one<NTX, 4, HALO, false, IO>(a, wp, threads, smem, st);
    case 5: case 6: return launch_window_one<NTX, 6, HALO, false, IO>(a, wp, threads, smem, st);
    case 7: case 8: return launch_window_one<NTX, 8, HALO, false, IO>(a, wp, threads, smem, st);
    default: set_error("polynomial order %d outside the compiled set", P); return HB_ERR_UNSUPPORTED;
  }
}

template <int NTX>
int launch_tdl_window(int P, bool io128, bool large_halo, bool lin, const FadingArgs& a, const WindowPlan& wp,
                      int threads, size_t smem, cudaStream_t st) {
  if (large_halo) {
    return io128 ? launch_window_io<NTX, kWindowHaloLarge, double2>(P, lin, a, wp, threads, smem, st)
                 : launch_window_io<NTX, kWindowHaloLarge, float2>(P, lin, a, wp, threads, smem, st);
  }
  return io128 ? launch_window_io<NTX, kWindowHaloSmall, double2>(P, lin, a, wp, threads, smem, st)
               : launch_window_io<NTX, kWindowHaloSmall, float2>(P, lin, a, wp, threads, smem, st);
}

template <int NTX, typename REAL, typename IO, bool STAGED = true>
static int launch_direct_one(const FadingArgs& a, const DelayTable& dt, int tpc, size_t smem,
                             cudaStream_t st) {
  auto kern = tdl_direct_kernel<NTX, REAL, IO, STAGED>;
  if (int e = ensure_smem(kern, smem)) return e;
  dim3 grid((unsigned)((size_t)a.ntiles * a.B));
  kern<<<grid, kThreads, smem, st>>>(a, dt, tpc);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

template <int NTX>
int launch_tdl_direct(bool f64, bool io128, bool unstaged, const FadingArgs& a, const DelayTable& dt, int tpc, size_t smem,
                      cudaStream_t st) {
  if (unstaged) {  // any delay spread: FP64 evaluation, x from global memory (both precision modes)
    return io128 ? launch_direct_one<NTX, double, double2, false>(a, dt, tpc, smem, st)
                 : launch_direct_one<NTX, double, float2, false>(a, dt, tpc, smem, st);
  }
  if (f64) {
    return io128 ? launch_direct_one<NTX, double, double2>(a, dt, tpc, smem, st)
                 : launch_direct_one<NTX, double, float2>(a, dt, tpc, smem, st);
  }
  return io128 ? launch_direct_one<NTX, float, double2>(a, dt, tpc, smem, st)
               : launch_direct_one<NTX, float, float2>(a, dt, tpc, smem, st);
}

#define HB_INSTANTIATE_FADING(NTX)                                                                        \
  template int launch_tdl_poly<NTX>(int, bool, const FadingArgs&, const DelayTable&, size_t, cudaStream_t); \
  template int launch_tdl_direct<NTX>(bool, bool, bool, const FadingArgs&, const DelayTable&, int, size_t,  \
                                      cudaStream_t);                                                        \
  template int launch_tdl_window<NTX>(int, bool, bool, bool, const FadingArgs&, const WindowPlan&, int, size_t,  \
                                      cudaStream_t);

}  // namespace hb
