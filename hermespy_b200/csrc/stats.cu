// K7 evaluator statistics and K2 Kronecker antenna-correlation mixing (sm_100a).
//
//   hb_bit_errors        BitErrorEvaluator.evaluate / artifact            hermespy/modem/evaluators.py:231-259
//   hb_stats_accumulate  ScalarEvaluationResult.add_artifact (sum, sum^2, count per grid cell)
//                                                                         hermespy/core/pymonte/scalar.py:101-125
//   hb_kron_mix          S <- R_rx S R_tx of MultipathFadingRealization._sample   hermespy/channel/fading/fading.py:476-489
//
// The statistics buffers are the ones the evaluator all-reduce reads: the local reduction lands
// directly in the tensors handed to NCCL.  All sums use a fixed partition and a fixed order, so a rank's
// contribution is reproducible bit for bit; the integer counters are exact.
#include "hb_common.cuh"

namespace hb {

// One CTA per drop: errors = sum_k |int8(tx_k) - int8(rx_k)| over max(tx_len, rx_len) bits, the shorter sequence padded
// with zeros (evaluators.py:246-256).
__global__ void __launch_bounds__(128) bit_error_kernel(const uint8_t* __restrict__ tx, const uint8_t* __restrict__ rx,
                                                        const int32_t* __restrict__ tx_len,
                                                        const int32_t* __restrict__ rx_len, int num_bits,
                                                        int64_t* __restrict__ errors, int64_t* __restrict__ bits,
                                                        double* __restrict__ artifact) {
  const int i = blockIdx.x;
  const int lt = tx_len ? min(max(tx_len[i], 0), num_bits) : num_bits;
  const int lr = rx_len ? min(max(rx_len[i], 0), num_bits) : num_bits;
  const int n = max(lt, lr);
  const uint8_t* t = tx + (size_t)i * num_bits;
  const uint8_t* r = rx + (size_t)i * num_bits;
  int acc = 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int a = k < lt ? (int)(int8_t)t[k] : 0;
    const int b = k < lr ? (int)(int8_t)r[k] : 0;
    acc += abs(a - b);
  }
  __shared__ int part[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t e = (int64_t)part[0] + part[1] + part[2] + part[3];
    errors[i] = e;
    bits[i] = n;
    if (artifact) artifact[i] = n > 0 ? (double)e / (double)n : 0.0;  // np.mean of the error indicators
  }
}

// One warp per grid cell.  Lane l visits drops l, l+32, ... in order, the 32 partial sums are combined by a fixed
// butterfly: the result does not depend on scheduling.
__global__ void __launch_bounds__(128) stats_accumulate_kernel(const double* __restrict__ artifact,
                                                               const int32_t* __restrict__ cell,
                                                               const int64_t* __restrict__ errors,
                                                               const int64_t* __restrict__ bits, int n, int num_cells,
                                                               double* __restrict__ stats, int64_t* __restrict__ counts) {
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= num_cells) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0, s2 = 0.0;
  long long cnt = 0, err = 0, nb = 0;
  for (int i = lane; i < n; i += 32) {
    if (cell[i] == c) {
      const double a = artifact[i];
      s += a;
      s2 = fma(a, a, s2);
      ++cnt;
      if (errors) err += errors[i];
      if (bits) nb += bits[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    err += __shfl_xor_sync(0xffffffffu, err, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }
  if (lane == 0) {
    stats[3 * c + 0] += s;
    stats[3 * c + 1] += s2;
    stats[3 * c + 2] += (double)cnt;
    if (counts) {
      counts[2 * c + 0] += err;
      counts[2 * c + 1] += nb;
    }
  }
}

// out[b] = R_rx @ S[b] @ R_tx (complex128, FP64 accumulate, row-major).  One CTA per link; the two products go through
// shared memory.  The reference multiplies by the correlation matrices themselves, not their square roots
// (fading.py:480-489).  R_rx / R_tx are shared by the batch.
__global__ void __launch_bounds__(256) kron_mix_kernel(const double2* __restrict__ Rrx, const double2* __restrict__ S,
                                                       const double2* __restrict__ Rtx, double2* __restrict__ out,
                                                       int nrx, int ntx, int has_rx, int has_tx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* Sb = reinterpret_cast<double2*>(smem_raw);  // [nrx, ntx]
  double2* Tb = Sb + nrx * ntx;                        // [nrx, ntx]  R_rx @ S
  const size_t off = (size_t)blockIdx.x * nrx * ntx;
  for (int i = threadIdx.x; i < nrx * ntx; i += blockDim.x) Sb[i] = S[off + i];
  __syncthreads();
  for (int i = threadIdx.x; i < nrx * ntx; i += blockDim.x) {
    const int r = i / ntx, c = i - r * ntx;
    double2 acc = Sb[i];
    if (has_rx) {
      acc = make_double2(0.0, 0.0);
      for (int k = 0; k < nrx; ++k) cmac<double>(acc, Rrx[r * nrx + k], Sb[k * ntx + c]);
    }
    Tb[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrx * ntx; i += blockDim.x) {
    const int r = i / ntx, c = i - r * ntx;
    double2 acc = Tb[i];
    if (has_tx) {
      acc = make_double2(0.0, 0.0);
      for (int k = 0; k < ntx; ++k) cmac<double>(acc, Tb[r * ntx + k], Rtx[k * ntx + c]);
    }
    out[off + i] = acc;
  }
}

// a3 on the device: the standard normals a static realization drew (host numpy generator: draw order is parity) ->
// kernel parameter block of every link.  MultipathFadingRealization._sample (fading.py:468-515) with
// ConsistentUniform.sample = norm.cdf (consistent.py:475-485):  u = Phi(g) = erfc(-g / sqrt 2) / 2,
//   S[i, j] = exp(2j pi u_ant[i, j]),  los_angle = 2 pi u,  nlos_angle = los_phase = nlos_phase = -pi + 2 pi u,
//   omega[l, 0] = w_los cos(los_angle_l) / fs,  omega[l, k] = w_nlos cos((2 pi k + nlos_angle_lk) / N) / fs  (fading.py:326-339)
// Variable layout of the normals: antenna (dim x dim) | los angles (L) | nlos angles (L, N) | los phases (L) | nlos phases (L, N)
// (declaration order, fading.py:742-754).  One thread per output element; trivially parallel, FP64 throughout.
struct SampleArgs {
  const double* normals;  // [B, S]
  const double* amp_tab;  // [L, 2] los / nlos amplitude incl. sqrt(gain * power_l), shared by the batch
  double* omega;          // [B, L, N + 1]
  double* phi;            // [B, L, N + 1]
  double* amp;            // [B, L, 2]
  double2* spatial;       // [B, nrx, ntx]
  double w_los, w_nlos;   // Doppler rates per SAMPLE (rad): doppler / fs
  int B, S, dim, nrx, ntx, L, N, transpose;
};

__device__ __forceinline__ double std_normal_cdf(double g) { return 0.5 * erfc(-g * 0.70710678118654752440); }

__global__ void __launch_bounds__(256) fading_sample_kernel(const SampleArgs a) {
  const int K = a.N + 1;
  const int per_link = a.L * K + a.nrx * a.ntx;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)a.B * per_link) return;
  const int b = (int)(e / per_link), r = (int)(e - (long long)b * per_link);
  const double* g = a.normals + (size_t)b * a.S;
  const int o_los_ang = a.dim * a.dim, o_nlos_ang = o_los_ang + a.L, o_los_ph = o_nlos_ang + a.L * a.N,
            o_nlos_ph = o_los_ph + a.L;
  if (r < a.L * K) {
    const int l = r / K, k = r - l * K;
    double om, ph;
    if (k == 0) {
      const double ang = kTwoPi * std_normal_cdf(g[o_los_ang + l]);
      om = a.w_los * cos(ang);
      ph = -M_PI + kTwoPi * std_normal_cdf(g[o_los_ph + l]);
      a.amp[((size_t)b * a.L + l) * 2 + 0] = a.amp_tab[2 * l];
      a.amp[((size_t)b * a.L + l) * 2 + 1] = a.amp_tab[2 * l + 1];
    } else {
      const double ang = -M_PI + kTwoPi * std_normal_cdf(g[o_nlos_ang + l * a.N + (k - 1)]);
      om = a.w_nlos * cos((kTwoPi * (double)k + ang) / (double)a.N);
      ph = -M_PI + kTwoPi * std_normal_cdf(g[o_nlos_ph + l * a.N + (k - 1)]);
    }
    a.omega[(size_t)b * a.L * K + r] = om;
    a.phi[(size_t)b * a.L * K + r] = ph;
  } else {
    const int s = r - a.L * K, i = s / a.ntx, j = s - i * a.ntx;
    // reciprocal direction: the transposed antenna matrix (fading.py:517-538)
    const double u = std_normal_cdf(a.transpose ? g[j * a.dim + i] : g[i * a.dim + j]);
    double sn, cs;
    sincospi(2.0 * u, &sn, &cs);
    a.spatial[(size_t)b * a.nrx * a.ntx + s] = make_double2(cs, sn);
  }
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb_fading_sample(const double* normals, int32_t batch, int32_t num_scalars, int32_t antenna_dim, int32_t num_rx,
                     int32_t num_tx, int32_t num_taps, int32_t num_sinusoids, const double* amp_table, double los_rate,
                     double nlos_rate, int32_t reciprocal, double* omega, double* phi, double* amp, void* spatial,
                     void* stream) {
  if (batch < 0 || num_taps < 1 || num_sinusoids < 0 || num_rx < 1 || num_tx < 1 || antenna_dim < 1) {
    set_error("invalid sampling shape (B=%d L=%d N=%d Nrx=%d Ntx=%d dim=%d)", batch, num_taps, num_sinusoids, num_rx, num_tx,
              antenna_dim);
    return HB_ERR_INVALID;
  }
  const long long need = (long long)antenna_dim * antenna_dim + 2ll * num_taps + 2ll * num_taps * num_sinusoids;
  if (num_scalars != need) {
    set_error("a realization of this channel holds %lld normals, got %d", need, num_scalars);
    return HB_ERR_INVALID;
  }
  if (num_rx > antenna_dim || num_tx > antenna_dim) {
    set_error("antenna variable is %d x %d, link needs %d x %d", antenna_dim, antenna_dim, num_rx, num_tx);
    return HB_ERR_INVALID;
  }
  if (batch == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!normals || !amp_table || !omega || !phi || !amp || !spatial) {
    set_error("NULL device pointer in sampling request");
    return HB_ERR_INVALID;
  }
  SampleArgs a;
  a.normals = normals;
  a.amp_tab = amp_table;
  a.omega = omega;
  a.phi = phi;
  a.amp = amp;
  a.spatial = (double2*)spatial;
  a.w_los = los_rate;
  a.w_nlos = nlos_rate;
  a.B = batch;
  a.S = num_scalars;
  a.dim = antenna_dim;
  a.nrx = num_rx;
  a.ntx = num_tx;
  a.L = num_taps;
  a.N = num_sinusoids;
  a.transpose = reciprocal ? 1 : 0;
  const long long total = (long long)batch * (num_taps * (num_sinusoids + 1) + num_rx * num_tx);
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_MISC, st);
  fading_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_bit_errors(const uint8_t* tx_bits, const uint8_t* rx_bits, const int32_t* tx_len, const int32_t* rx_len,
                  int32_t num_drops, int32_t num_bits, int64_t* errors, int64_t* bits, double* artifact, void* stream) {
  if (num_drops < 0 || num_bits < 0) {
    set_error("invalid bit-error shape (drops=%d, bits=%d)", num_drops, num_bits);
    return HB_ERR_INVALID;
  }
  if (num_drops == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!tx_bits || !rx_bits || !errors || !bits) {
    set_error("NULL device pointer in bit-error request");
    return HB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_STATS, st);
  bit_error_kernel<<<num_drops, 128, 0, st>>>(tx_bits, rx_bits, tx_len, rx_len, num_bits, errors, bits, artifact);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_stats_accumulate(const double* artifact, const int32_t* cell, const int64_t* errors, const int64_t* bits,
                        int32_t num_drops, int32_t num_cells, double* stats, int64_t* counts, void* stream) {
  if (num_drops < 0 || num_cells < 0) {
    set_error("invalid statistics shape (drops=%d, cells=%d)", num_drops, num_cells);
    return HB_ERR_INVALID;
  }
  if (num_drops == 0 || num_cells == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!artifact || !cell || !stats) {
    set_error("NULL device pointer in statistics request");
    return HB_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_STATS, st);
  stats_accumulate_kernel<<<(num_cells + 3) / 4, 128, 0, st>>>(artifact, cell, errors, bits, num_drops, num_cells, stats,
                                                              counts);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

int hb_kron_mix(const void* r_rx, const void* spatial, const void* r_tx, void* out, int32_t batch, int32_t num_rx,
                int32_t num_tx, void* stream) {
  if (batch < 0 || num_rx < 1 || num_tx < 1) {
    set_error("invalid Kronecker mixing shape (B=%d Nrx=%d Ntx=%d)", batch, num_rx, num_tx);
    return HB_ERR_INVALID;
  }
  if (batch == 0) return HB_OK;
  if (int e = require_device()) return e;
  if (!spatial || !out) {
    set_error("NULL device pointer in Kronecker mixing request");
    return HB_ERR_INVALID;
  }
  const size_t smem = 2 * sizeof(double2) * (size_t)num_rx * num_tx;
  if (smem > 200 * 1024) {
    set_error("%d x %d spatial matrix exceeds the shared-memory tile of hb_kron_mix", num_rx, num_tx);
    return HB_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kron_mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = (cudaStream_t)stream;
  ProfileScope prof(KIND_MISC, st);
  kron_mix_kernel<<<batch, 256, smem, st>>>((const double2*)r_rx, (const double2*)spatial, (const double2*)r_tx,
                                            (double2*)out, num_rx, num_tx, r_rx != nullptr, r_tx != nullptr);
  HB_CUDA(cudaGetLastError());
  return HB_OK;
}

}  // extern "C"
