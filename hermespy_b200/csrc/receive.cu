// Receive superposition + additive white Gaussian noise, fused (sm_100a).  SURVEY 8(f)-3.
//
//   out[b, i, m] = sum_k in_k[b, i, m - off_k]   (0 <= m - off_k < T_k)  +  scale[b] (n_re[b, i, m] + j n_im[b, i, m])
//
// Replaces, for signals of one sampling rate and carrier frequency whose delays are whole samples,
//   * the superposition loop of SimulatedDevice.process_input      hermespy/simulation/simulated_device.py:1899-1915
//     (SparseSignal.Empty(...).superimpose(signal) for every impinging signal, then to_dense())
//   * AWGNRealization.add_to                                       hermespy/simulation/rf/noise/model.py:140-160
//     noise = sqrt(P / 2) (rng.standard_normal(shape) + 1j rng.standard_normal(shape)); block += noise
// The standard normals are drawn by the caller with the reference's generator (parity: noise is part of the drop's random
// stream) and shipped as two float64 planes.  Arithmetic order is the reference's -- inputs summed in the order given,
// the scaled noise added last, every product and sum rounded on its own (no FMA contraction) -- so the float64 /
// complex128 result is bit-identical to numpy's.
//
// One pass over HBM: reads sum_k T_k + 2 Tout (noise planes), writes Tout per stream; grid-stride, 16-byte accesses.
#include "hb_common.cuh"

namespace hb {

constexpr int kReceiveMaxInputs = 8;

struct ReceiveArgs {
  const void* in[kReceiveMaxInputs];
  int32_t len[kReceiveMaxInputs];
  int32_t off[kReceiveMaxInputs];
  int32_t K;
  const double* noise_re;
  const double* noise_im;
  const double* scale;
  void* out;
  int32_t B, nrx, Tout;
};

template <typename IO>
__global__ void __launch_bounds__(256) receive_combine_kernel(const __grid_constant__ ReceiveArgs a) {
  const long long rows = (long long)a.B * a.nrx;
  const long long total = rows * a.Tout;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / a.Tout;
    const int m = (int)(e - row * a.Tout);
    double re = 0.0, im = 0.0;
    bool first = true;
#pragma unroll
    for (int k = 0; k < kReceiveMaxInputs; ++k) {
      if (k < a.K) {
        const int n = m - a.off[k];
        if (n >= 0 && n < a.len[k]) {
          const IO v = ldg_stream(reinterpret_cast<const IO*>(a.in[k]) + row * a.len[k] + n);
          // the reference starts from an all-zero block and adds: 0 + v == v exactly, so the first term is a copy
          re = first ? (double)v.x : __dadd_rn(re, (double)v.x);
          im = first ? (double)v.y : __dadd_rn(im, (double)v.y);
          first = false;
        }
      }
    }
    if (a.noise_re != nullptr) {
      const double s = a.scale[row / a.nrx];
      re = __dadd_rn(re, __dmul_rn(s, a.noise_re[e]));
      im = __dadd_rn(im, __dmul_rn(s, a.noise_im[e]));
    }
    stg_stream(reinterpret_cast<IO*>(a.out) + e, IoConv<IO>::make(re, im));
  }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_receive_combine(const hb_receive_input* inputs, int32_t num_inputs, const double* noise_re,
                                  const double* noise_im, const double* noise_scale, void* out, int32_t batch,
                                  int32_t num_rx, int32_t num_out_samples, int32_t io_complex128, void* stream) {
  if (batch < 0 || num_rx < 0 || num_out_samples < 0 || num_inputs < 0 || (num_inputs > 0 && !inputs)) {
    set_error("invalid receive shape (B=%d Nrx=%d T=%d inputs=%d)", batch, num_rx, num_out_samples, num_inputs);
    return HB_ERR_INVALID;
  }
  if (num_inputs > kReceiveMaxInputs) {
    set_error("at most %d impinging signals per call (got %d): combine in stages", kReceiveMaxInputs, num_inputs);
    return HB_ERR_UNSUPPORTED;
  }
  if ((noise_re == nullptr) != (noise_im == nullptr) || (noise_re != nullptr && noise_scale == nullptr)) {
    set_error("noise needs both normal planes and the per-link scale");
    return HB_ERR_INVALID;
  }
  ReceiveArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < num_inputs; ++k) {
    if (inputs[k].num_samples < 0 || inputs[k].offset < 0 || inputs[k].offset + inputs[k].num_samples > num_out_samples ||
        (inputs[k].num_samples > 0 && !inputs[k].samples)) {
      set_error("impinging signal %d: %d samples at offset %d do not fit %d output samples", k, inputs[k].num_samples,
                inputs[k].offset, num_out_samples);
      return HB_ERR_INVALID;
    }
    a.in[k] = inputs[k].samples;
    a.len[k] = inputs[k].num_samples;
    a.off[k] = inputs[k].offset;
  }
  a.K = num_inputs;
  a.noise_re = noise_re;
  a.noise_im = noise_im;
  a.scale = noise_scale;
  a.out = out;
  a.B = batch;
  a.nrx = num_rx;
  a.Tout = num_out_samples;
  const long long total = (long long)batch * num_rx * num_out_samples;
  if (total == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!out) {
    set_error("NULL output pointer");
    return HB_ERR_INVALID;
  }
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)device_sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_MISC, st);
  if (io_complex128) receive_combine_kernel<double2><<<grid, 256, 0, st>>>(a);
  else receive_combine_kernel<float2><<<grid, 256, 0, st>>>(a);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}
