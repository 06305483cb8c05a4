// C-ABI entry of the tensor-core spatial GEMM (K4 for large arrays).
#include <algorithm>

#include "fading_fused.cuh"

namespace hb {

template <int P>
static int launch_fused_p(const FusedArgs& a, unsigned grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA(cudaFuncSetAttribute(fused_gemm_tdl_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes));
    attr_set = true;
  }
  ProfileScope prof(KIND_SPATIAL_GEMM, st);
  fused_gemm_tdl_kernel<P><<<grid, kFusedThreads, kFusedSmemBytes, st>>>(a);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

// Fused spatial GEMM + tap delay lines (fading_fused.cuh).  `a` comes with everything but the work split filled in.
int launch_fused_gemm_tdl(int P, const FusedArgs& args, cudaStream_t st) {
  FusedArgs a = args;
  if (a.B == 0 || a.T + a.D == 0 || a.nrx == 0) return HB_OK;
  const int sms = persistent_sm_count();
  a.ntiles = (a.T + a.D + kGemmTileSamples - 1) / kGemmTileSamples;
  // work items: enough per SM to balance the tail, long enough to amortize the conversion of S and the two history tiles
  const long long want_items = 4ll * sms;
  int seg = (int)std::max<long long>(32, ((long long)a.ntiles * a.B + want_items - 1) / want_items);
  seg = std::min(seg, a.ntiles);
  a.seg_tiles = seg;
  a.nseg = (a.ntiles + seg - 1) / seg;
  const long long items = (long long)a.B * a.nseg;
  const unsigned grid = (unsigned)std::min<long long>(items, sms);
  switch (P) {
    case 1: return launch_fused_p<1>(a, grid, st);
    case 2: return launch_fused_p<2>(a, grid, st);
    case 3: return launch_fused_p<3>(a, grid, st);
    case 4: return launch_fused_p<4>(a, grid, st);
  }
  set_error("fused large-array kernel: polynomial order %d outside the compiled set", P);
  return HB_ERR_UNSUPPORTED;
}

int launch_spatial_gemm(const double2* S, const float2* z, float2* y, int B, int nrx, int ntx, int T, cudaStream_t st,
                        int z_pitch) {
  if (B == 0 || T == 0 || nrx == 0) return HB_OK;
  static bool attr_set = false;
  if (!attr_set) {
    HB_CUDA(cudaFuncSetAttribute(spatial_gemm_3xtf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kGemmSmemBytes));
    HB_CUDA(cudaFuncSetAttribute(spatial_gemm_3xtf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kGemmSmemBytes));
    attr_set = true;
  }
  const int sms = persistent_sm_count();
  GemmArgs a;
  a.S = S;
  a.z = z;
  a.ldz = z_pitch > 0 ? z_pitch : T;
  a.y = y;
  a.B = B;
  a.T = T;
  a.nrx_total = nrx;
  a.ntx_total = ntx;
  a.ntiles = (T + kGemmTileSamples - 1) / kGemmTileSamples;
  // work items: enough per SM to balance the tail, long enough to amortize the conversion of S (64 KB per item)
  long long want_items = 4ll * sms;
  int seg = (int)std::max<long long>(16, ((long long)a.ntiles * B + want_items - 1) / want_items);
  seg = std::min(seg, a.ntiles);
  a.seg_tiles = seg;
  a.nseg = (a.ntiles + seg - 1) / seg;
  const long long items = (long long)B * a.nseg;
  const unsigned grid = (unsigned)std::min<long long>(items, sms);
  for (int rx0 = 0; rx0 < nrx; rx0 += kGemmMaxAnt) {
    for (int tx0 = 0; tx0 < ntx; tx0 += kGemmMaxAnt) {
      a.rx0 = rx0;
      a.nrx = std::min(kGemmMaxAnt, nrx - rx0);
      a.tx0 = tx0;
      a.ntx = std::min(kGemmMaxAnt, ntx - tx0);
      a.accumulate = tx0 > 0;
      ProfileScope prof(KIND_SPATIAL_GEMM, st);
      if (a.nrx == kGemmMaxAnt && a.ntx == kGemmMaxAnt && !a.accumulate)
        spatial_gemm_3xtf32_kernel<true><<<grid, kGemmLaunchThreads, kGemmSmemBytes, st>>>(a);
      else
        spatial_gemm_3xtf32_kernel<false><<<grid, kGemmLaunchThreads, kGemmSmemBytes, st>>>(a);
      HB_CUDA(cudaGetLastError());
    }
  }
  return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_spatial_gemm_3xtf32(const void* spatial, const void* z, void* y, int32_t batch, int32_t num_rx,
                                      int32_t num_tx, int32_t num_samples, void* stream) {
  if (batch < 0 || num_rx < 0 || num_tx < 1 || num_samples < 0) {
    set_error("invalid spatial GEMM shape (B=%d Nrx=%d Ntx=%d T=%d)", batch, num_rx, num_tx, num_samples);
    return HB_ERR_INVALID;
  }
  if (batch == 0 || num_rx == 0 || num_samples == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!spatial || !z || !y) {
    set_error("NULL device pointer in spatial GEMM request");
    return HB_ERR_INVALID;
  }
  return launch_spatial_gemm((const double2*)spatial, (const float2*)z, (float2*)y, batch, num_rx, num_tx, num_samples,
                             (cudaStream_t)stream, 0);
}
