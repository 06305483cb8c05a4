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

}  // namespace hb

using namespace hb;

extern "C" {

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
