// Float64 parity mode on the Taylor path (sm_100a): K1 and the gather form of K3 + K4 entirely in FP64.
//
// The parity mode (HB_F64: <= 1e-12 against the reference, bit-exact BER counts) used to be direct evaluation only: one
// FP64 sincos per (tap, sinusoid, sample) -- C2: 483 sincos per link sample, 0.79 G samples/s, 0.015 of the HBM roofline.
// B200's FP64 pipe runs at half the FP32 rate, so the same factorization as the complex64 path pays here too:
//   K1' sos_poly_coef64_kernel   Taylor moments of every delay group about the window centre, phase, sincos and moment
//                                recurrence in FP64; one warp per (link, window, group)
//   K3' tdl_poly64_kernel        y = S sum_g h_g(m) x[m - d_g], h_g by Horner, all FP64; x tile + halo in shared memory
// with the truncation bound of the expansion held below 1e-14 (P in {4, 6, 8}, planner).  What differs from the
// reference's own rounding is then the ARGUMENT rounding of its cos / sin (half an ulp of omega n + phi), so AUTO takes
// this path only while the largest phase of the frame stays below 1000 rad (1e-13); beyond, direct evaluation reproduces the
// reference's argument bit for bit.
#pragma once
#include "fading_kernels.cuh"

namespace hb {

constexpr double kPolyTarget64 = 1e-14;     // truncation bound, relative to the RMS tap gain
constexpr double kPoly64MaxPhase = 1000.0;  // AUTO: omega_max (T + D) above this stays on direct evaluation

template <int P>
__global__ void __launch_bounds__(128) sos_poly_coef64_kernel(const FadingArgs a, const __grid_constant__ DelayTable dt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, G = dt.num_groups;
  const long long item = (long long)blockIdx.x * 4 + warp;
  if (item >= (long long)a.B * a.ntiles * G) return;
  const long long bq = item / G;
  const int g = (int)(item - bq * G);
  const int b = (int)(bq / a.ntiles), q = (int)(bq - (long long)b * a.ntiles);
  const double shift = (double)q * a.tile + 0.5 * a.tile - (double)dt.group_delay[g];
  const double* om_b = a.omega + (size_t)b * a.L * K;
  const double* ph_b = a.phi + (size_t)b * a.L * K;
  const double* am_b = a.amp + (size_t)b * a.L * 2;
  double vr[P], vi[P];
#pragma unroll
  for (int p = 0; p < P; ++p) vr[p] = vi[p] = 0.0;
  const int l0 = dt.group_start[g], l1 = dt.group_start[g + 1];
  for (int idx = l0 * K + lane; idx < l1 * K; idx += 32) {  // fixed pair -> lane assignment: deterministic sums
    const int l = idx / K, k = idx - l * K;
    const double om = om_b[idx];
    double s, c;
    sincos(fma(om, shift, ph_b[idx]), &s, &c);
    const double am = am_b[2 * l + (k != 0 ? 1 : 0)];
    double tr = am * c, ti = am * s;
    const double u = om * (double)a.tile;
    vr[0] += tr;
    vi[0] += ti;
#pragma unroll
    for (int p = 1; p < P; ++p) {
      const double f = u / (double)p;
      const double nr = -ti * f, ni = tr * f;  // times (j u / p)
      tr = nr;
      ti = ni;
      vr[p] += tr;
      vi[p] += ti;
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      vr[p] += __shfl_xor_sync(0xffffffffu, vr[p], off);
      vi[p] += __shfl_xor_sync(0xffffffffu, vi[p], off);
    }
  }
  if (lane == 0) {
    double2* out = reinterpret_cast<double2*>(const_cast<float2*>(a.coef)) + ((size_t)b * a.ntiles + q) * a.coef_stride + g * P;
#pragma unroll
    for (int p = 0; p < P; ++p) out[p] = make_double2(vr[p], vi[p]);
  }
}

// grid = B * ntiles, 256 threads, R outputs per thread (stride 256).  smem: xs[NTX][W] | cs[G P] | Ss[nrx NTX] (double2)
template <int NTX, int P, int R, typename IO>
__global__ void __launch_bounds__(kThreads) tdl_poly64_kernel(const FadingArgs a, const __grid_constant__ DelayTable dt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int W = a.tile + a.Dpad;
  const int G = dt.num_groups;
  double2* xs = reinterpret_cast<double2*>(smem_raw);
  double2* cs = xs + NTX * W;
  double2* Ss = cs + G * P;
  stage_x_tile<NTX, double2, IO>(xs, a, b, q, W);
  {
    const double2* cb = reinterpret_cast<const double2*>(a.coef) + ((size_t)b * a.ntiles + q) * a.coef_stride;
    for (int c = tid; c < G * P; c += kThreads) cs[c] = cb[c];
  }
  stage_spatial<NTX, double2>(Ss, a, b);
  __syncthreads();

  const int Tout = a.T + a.D;
  const double inv_tile = 1.0 / (double)a.tile, half = 0.5 * (double)a.tile;
  for (int base = 0; base < a.tile; base += kThreads * R) {
    if ((long long)q * a.tile + base >= Tout) break;
    const int nu = min(R, (a.tile - base) / kThreads);  // tile is a multiple of 256: uniform over the CTA
    double2 z[R][NTX];
    double rr[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      rr[u] = ((double)(base + u * kThreads + tid) - half) * inv_tile;
#pragma unroll
      for (int j = 0; j < NTX; ++j) z[u][j] = make_double2(0.0, 0.0);
    }
    for (int g = 0; g < G; ++g) {
      const int off = base + tid + a.Dpad - dt.group_delay[g];
      double2 c[P];
#pragma unroll
      for (int p = 0; p < P; ++p) c[p] = cs[g * P + p];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (u < nu) {
          double2 h = c[P - 1];
#pragma unroll
          for (int p = P - 2; p >= 0; --p) {
            h.x = fma(h.x, rr[u], c[p].x);
            h.y = fma(h.y, rr[u], c[p].y);
          }
#pragma unroll
          for (int j = 0; j < NTX; ++j) cmac<double>(z[u][j], xs[j * W + off + u * kThreads], h);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int m = q * a.tile + base + u * kThreads + tid;
      if (u < nu && m < Tout) spatial_store<NTX, double2, IO>(Ss, z[u], a, b, m);
    }
  }
}

inline size_t poly64_smem(int ntx_tpl, int tile, int Dpad, int G, int P, int nrx) {
  return sizeof(double2) * ((size_t)ntx_tpl * (tile + Dpad) + (size_t)G * P + (size_t)nrx * ntx_tpl);
}

// defined in fading_poly64.cu
int launch_coef64(int P, const FadingArgs& a, const DelayTable& dt, cudaStream_t st);
int launch_tdl_poly64(int ntx_tpl, int P, bool io128, const FadingArgs& a, const DelayTable& dt, size_t smem, cudaStream_t st);

}  // namespace hb
