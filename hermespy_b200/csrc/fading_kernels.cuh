// Fading hot path kernels (sm_100a).
//
//   K1  sos_poly_coef_kernel   per (link, tile, delay group): Taylor moments of the Rician
//                              sum-of-sinusoids tap gain about the tile centre      (fading.py:293-343)
//   K3+K4 tdl_poly_kernel      y = S @ sum_g shift_{d_g}(x * h_g), h_g from the K1 moments by Horner;
//                              x tile + delay halo staged in shared memory            (fading.py:371-406)
//   K1+K3+K4 tdl_direct_kernel same result with one sincos per sinusoid per sample: MUFU pipe (float,
//                              32-bit fixed-point phase accumulators) or FP64 parity mode (double)
//   sos_state_kernel           SISO tap gains h[B, G, T] for the channel state        (fading.py:345-358)
//
// Algebra used by the gather form: the reference scatters  z[:, d_l + n] += x[:, n] h_l[n]; each output
// sample m gathers  z[:, m] = sum_l x[:, m - d_l] h_l[m - d_l]  over taps with 0 <= m - d_l < T.  Taps whose
// rounded delays coincide share x[:, m - d] and only ever appear as their sum, so they are merged into
// one "delay group" whose gain is the sum of the member taps' gains.
#pragma once
#include "hb_common.cuh"

namespace hb {

struct FadingArgs {
  const void* x;
  void* y;
  const double* omega;     // [B, L, K]
  const double* phi;       // [B, L, K]
  const double* amp;       // [B, L, 2]
  const double2* spatial;  // [B, Nrx, Ntx]
  const float2* coef;      // [B, ntiles, coef_stride >= G * P]  (poly mode)
  float2* spatial32;       // [B, nchunks, s32_stride >= Nrx * s32_tpl] FP32 copy of `spatial` in antenna chunks of
                           // s32_tpl transmit antennas (zero padded), written by K1 when not NULL
  unsigned int* tile_counters;  // [num_counters] work counters of the persistent kernels, zeroed by K1 when not NULL
  int num_counters;
  int coef_flat;  // K1 work split: 1 = one warp per (link, window, group) with flattened (tap, sinusoid) lanes (G < 4)
  int coef_stride, s32_tpl, s32_stride;
  int B, ntx, nrx, T, D, L, K;
  int tile, ntiles, Dpad;
  int tx0, ntx_chunk, accumulate;
  int ypitch;   // z mode: samples between consecutive rows of the z workspace (even, >= T + D); 0 = T + D
  int z_mode;   // large arrays: skip the spatial mix, store the chunk's tap-delay-line outputs z[b, tx0 + j, :] (y has ntx rows)
  int dbg;  // attribution builds only (-DHB_ATTRIBUTION, env HB_DBG: 1 skip staging, 2 skip stores, 4 skip walk)
};

// ------------------------------------------------------------------------------------------------
// K1: Taylor moments.  For output index m = q*tile + i, r = (i - tile/2) / tile in [-1/2, 1/2):
//   h_g(m - d_g) = sum_{l in g} sum_k amp_lk exp(j(omega_lk (m - d_g) + phi_lk))
//                = sum_p r^p * coef[g][p],
//   coef[g][p]   = sum_{l,k} amp_lk e^{j theta_lk} (j u_lk)^p / p!,
//   theta_lk = phi_lk + omega_lk (q*tile + tile/2 - d_g),   u_lk = omega_lk * tile.
// theta is formed and range-reduced in FP64 (it reaches 1e3..1e6 rad for long frames), the sincos and the
// moment recurrence run in FP32.  One warp per delay group, lanes over the (tap, sinusoid) pairs of the
// group in a fixed order -> deterministic sums.
template <int P>
__global__ void __launch_bounds__(128) sos_poly_coef_kernel(const FadingArgs a,
                                                            const __grid_constant__ DelayTable dt) {
  // Two work splits (a.coef_flat, chosen by the launcher):
  //  * G >= 4: CTA = (link, Taylor window), its 4 warps stride over the delay groups -- the per-CTA set-up (pointers,
  //    FP32 spatial matrix) is paid once per window;
  //  * G < 4 (C1: 23 taps share ONE delay): warp = (link, window, group), 4 consecutive items per CTA, the (tap, sinusoid)
  //    pairs of the group flattened over the lanes -- otherwise three warps in four would idle and K = 21 sinusoids would
  //    leave a third of the remaining lanes empty.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, G = dt.num_groups;
  int b, q, g_first, g_step;
  if (a.coef_flat) {
    const long long item = (long long)blockIdx.x * 4 + warp;
    if (item >= (long long)a.B * a.ntiles * G) return;
    const long long bq = item / G;
    g_first = (int)(item - bq * G);
    g_step = G;  // one group per warp
    b = (int)(bq / a.ntiles);
    q = (int)(bq - (long long)b * a.ntiles);
  } else {
    b = blockIdx.x / a.ntiles;
    q = blockIdx.x - b * a.ntiles;
    g_first = warp;
    g_step = 4;
  }
  const double centre = (double)q * a.tile + 0.5 * a.tile;
  const double* om_b = a.omega + (size_t)b * a.L * K;
  const double* ph_b = a.phi + (size_t)b * a.L * K;
  const double* am_b = a.amp + (size_t)b * a.L * 2;
  if (blockIdx.x == 0 && a.tile_counters != nullptr)
    for (int i = threadIdx.x; i < a.num_counters; i += blockDim.x) a.tile_counters[i] = 0u;
  if (q == 0 && a.spatial32 != nullptr && a.s32_tpl > 0 && (!a.coef_flat || g_first == 0)) {
    // FP32 spatial matrix, chunked, for the bulk-copy staged kernel (flat split: the warp that owns group 0 converts it)
    const int tpl = a.s32_tpl, nch = (a.ntx + tpl - 1) / tpl, per = a.nrx * tpl;
    const int i0 = a.coef_flat ? lane : (int)threadIdx.x, di = a.coef_flat ? 32 : (int)blockDim.x;
    for (int i = i0; i < nch * per; i += di) {
      const int c = i / per, r = i - c * per, irx = r / tpl, j = c * tpl + (r - irx * tpl);
      float2 v = make_float2(0.f, 0.f);
      if (j < a.ntx) v = to_c32(a.spatial[((size_t)b * a.nrx + irx) * a.ntx + j]);
      a.spatial32[((size_t)b * nch + c) * a.s32_stride + r] = v;
    }
  }
  constexpr int NV = 2 * P <= 2 ? 2 : (2 * P <= 4 ? 4 : (2 * P <= 8 ? 8 : 16));  // values to reduce, padded to 2^k
  for (int g = g_first; g < G; g += g_step) {
    const int l0 = dt.group_start[g], l1 = dt.group_start[g + 1];
    const double shift = centre - (double)dt.group_delay[g];
    float v[NV];  // v[2p] = Re, v[2p+1] = Im of moment p
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
    // one (tap, sinusoid) pair: theta reduced in FP64, sincos + moment recurrence in FP32.  Fixed pair -> lane assignment
    // in both splits: deterministic sums.
    auto pair = [&](int l, int k) {
      const int idx = l * K + k;
      const double om = om_b[idx];
      const double th = fma(om, shift, ph_b[idx]);
      double t = th * kInvTwoPi;
      t -= rint(t);
      float s, c;
      sincosf((float)(t * kTwoPi), &s, &c);
      const float am = (float)am_b[2 * l + (k != 0 ? 1 : 0)];
      float tr = am * c, ti = am * s;
      const float u = (float)(om * (double)a.tile);
      v[0] += tr;
      v[1] += ti;
#pragma unroll
      for (int p = 1; p < P; ++p) {
        const float f = u * (1.0f / (float)p);
        const float nr = -ti * f, ni = tr * f;  // times (j u / p)
        tr = nr;
        ti = ni;
        v[2 * p] += tr;
        v[2 * p + 1] += ti;
      }
    };
    if (a.coef_flat) {
      for (int idx = l0 * K + lane; idx < l1 * K; idx += 32) pair(idx / K, idx % K);
    } else {
      for (int l = l0; l < l1; ++l)
        for (int k = lane; k < K; k += 32) pair(l, k);
    }
    // Transposing butterfly: at every step a lane keeps one half of its values and sends the other half, so the NV
    // sums cost NV - 1 + (5 - log2 NV) shuffles instead of 5 NV.  Value i ends up in the lanes whose top log2(NV)
    // bits spell i (bit 4 = most significant).
    int n = NV, off = 16;
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      if (n > 1) {
        const int half = n >> 1;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) {
          if (i < half) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        n = half;
        off >>= 1;
      }
    }
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    constexpr int LOG = NV == 2 ? 1 : (NV == 4 ? 2 : (NV == 8 ? 3 : 4));
    if ((lane & ((32 >> LOG) - 1)) == 0) {
      // index of the value this lane holds: lane bit 4 is the most significant index bit
      int vi = 0;
#pragma unroll
      for (int sft = 0; sft < LOG; ++sft) vi |= ((lane >> (4 - sft)) & 1) << (LOG - 1 - sft);
      if (vi < 2 * P) {
        float* out = reinterpret_cast<float*>(const_cast<float2*>(a.coef) + ((size_t)b * a.ntiles + q) * a.coef_stride + g * P);
        out[vi] = v[0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1, rotating form (links spanning several Taylor windows).  The phase of a sinusoid advances by omega * tile from one
// window centre to the next, so only the FIRST window of a warp's range pays the range reduction and the sincos; every
// further window is one FP64 complex rotation of the running phasor (4 FP64 operations; rounding 1e-16 per step, i.e.
// 5e-14 after the 513 windows of a 2^20-sample frame) followed by the FP32 moment recurrence.  Per (pair, window) that
// is ~22 issue slots instead of ~100 (FP64 reduction + full-range sincosf).
//   warp = (link, delay group, chunk of `wchunk` consecutive windows); lanes over the (tap, sinusoid) pairs of the group,
//   four pairs per lane and pass; the per-window sums use the same transposing butterfly and lane assignment every time
//   (deterministic).  grid = ceil(B G nchunk / 4) CTAs of 128 threads.
template <int P>
__global__ void __launch_bounds__(128) sos_poly_coef_rot_kernel(const FadingArgs a, const __grid_constant__ DelayTable dt,
                                                                const int wchunk, const int nchunk) {
  constexpr int kSlots = 4;
  constexpr int NV = 2 * P <= 2 ? 2 : (2 * P <= 4 ? 4 : (2 * P <= 8 ? 8 : 16));
  constexpr int LOG = NV == 2 ? 1 : (NV == 4 ? 2 : (NV == 8 ? 3 : 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, G = dt.num_groups;
  if (blockIdx.x == 0 && a.tile_counters != nullptr)
    for (int i = threadIdx.x; i < a.num_counters; i += blockDim.x) a.tile_counters[i] = 0u;
  const long long item = (long long)blockIdx.x * 4 + warp;
  if (item >= (long long)a.B * G * nchunk) return;
  const int c = (int)(item % nchunk);
  const long long bg = item / nchunk;
  const int g = (int)(bg % G), b = (int)(bg / G);
  const int q0 = c * wchunk, q1 = min(a.ntiles, q0 + wchunk);
  if (g == 0 && c == 0 && a.spatial32 != nullptr && a.s32_tpl > 0) {
    // FP32 spatial matrix, chunked, for the bulk-copy staged kernel
    const int tpl = a.s32_tpl, nch = (a.ntx + tpl - 1) / tpl, per = a.nrx * tpl;
    for (int i = lane; i < nch * per; i += 32) {
      const int cc = i / per, r = i - cc * per, irx = r / tpl, j = cc * tpl + (r - irx * tpl);
      float2 v = make_float2(0.f, 0.f);
      if (j < a.ntx) v = to_c32(a.spatial[((size_t)b * a.nrx + irx) * a.ntx + j]);
      a.spatial32[((size_t)b * nch + cc) * a.s32_stride + r] = v;
    }
  }
  const double* om_b = a.omega + (size_t)b * a.L * K;
  const double* ph_b = a.phi + (size_t)b * a.L * K;
  const double* am_b = a.amp + (size_t)b * a.L * 2;
  const int i0 = dt.group_start[g] * K, i1 = dt.group_start[g + 1] * K;
  const double shift0 = (double)q0 * a.tile + 0.5 * a.tile - (double)dt.group_delay[g];
  // which of the NV reduced values this lane ends up holding (lane bit 4 is the most significant index bit)
  int vi = 0;
#pragma unroll
  for (int sft = 0; sft < LOG; ++sft) vi |= ((lane >> (4 - sft)) & 1) << (LOG - 1 - sft);
  const bool writer = (lane & ((32 >> LOG) - 1)) == 0 && vi < 2 * P;

  for (int s0 = i0; s0 < i1; s0 += 32 * kSlots) {
    double xr[kSlots], xi[kSlots], rr[kSlots], ri[kSlots];
    float uu[kSlots];
#pragma unroll
    for (int j = 0; j < kSlots; ++j) {
      const int idx = s0 + 32 * j + lane;
      xr[j] = xi[j] = ri[j] = 0.0;
      rr[j] = 1.0;
      uu[j] = 0.f;
      if (idx < i1) {
        const int l = idx / K, k = idx - l * K;
        const double om = om_b[idx];
        double t = fma(om, shift0, ph_b[idx]) * kInvTwoPi;
        t -= rint(t);
        double sn, cs;
        sincospi(2.0 * t, &sn, &cs);
        const double am = am_b[2 * l + (k != 0 ? 1 : 0)];
        xr[j] = am * cs;
        xi[j] = am * sn;
        const double step = om * (double)a.tile;
        uu[j] = (float)step;
        if (q1 - q0 > 1) {
          double ts = step * kInvTwoPi;
          ts -= rint(ts);
          sincospi(2.0 * ts, &ri[j], &rr[j]);
        }
      }
    }
    const int nslot = min(kSlots, (i1 - s0 + 31) >> 5);  // warp-uniform: C2's one-tap groups fill a single slot
    for (int q = q0; q < q1; ++q) {
      float v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = 0.f;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        if (j < nslot) {
        float tr = (float)xr[j], ti = (float)xi[j];
        v[0] += tr;
        v[1] += ti;
#pragma unroll
        for (int p = 1; p < P; ++p) {
          const float f = uu[j] * (1.0f / (float)p);
          const float nr = -ti * f, ni = tr * f;  // times (j u / p)
          tr = nr;
          ti = ni;
          v[2 * p] += tr;
          v[2 * p + 1] += ti;
        }
        const double nx = xr[j] * rr[j] - xi[j] * ri[j];  // advance the phasor to the next window centre
        xi[j] = fma(xr[j], ri[j], xi[j] * rr[j]);
        xr[j] = nx;
        }
      }
      int n = NV, off = 16;
#pragma unroll
      for (int stepi = 0; stepi < 4; ++stepi) {
        if (n > 1) {
          const int half = n >> 1;
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < NV / 2; ++i) {
            if (i < half) {
              const float send = upper ? v[i] : v[i + half];
              const float keep = upper ? v[i + half] : v[i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          n = half;
          off >>= 1;
        }
      }
      for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
      if (writer) {
        float* out = reinterpret_cast<float*>(const_cast<float2*>(a.coef) + ((size_t)b * a.ntiles + q) * a.coef_stride + g * P) + vi;
        *out = s0 == i0 ? v[0] : *out + v[0];  // groups of more than 128 pairs: further passes add, same lane, same order
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage x[b, tx0 : tx0+NTX, q*tile - Dpad : q*tile + tile) into shared memory, zero outside [0, T).
template <int NTX, typename C, typename IO>
__device__ __forceinline__ void stage_x_tile(C* xs, const FadingArgs& a, int b, int q, int W) {
  const IO* xb = reinterpret_cast<const IO*>(a.x) + ((size_t)b * a.ntx + a.tx0) * a.T;
  const int n0 = q * a.tile - a.Dpad;
  using R = decltype(C::x);
#pragma unroll
  for (int j = 0; j < NTX; ++j) {
    const bool live = j < a.ntx_chunk;
    const IO* row = xb + (size_t)j * a.T;
#pragma unroll 4
    for (int c = threadIdx.x; c < W; c += kThreads) {
      const int n = n0 + c;
      C v;
      v.x = (R)0;
      v.y = (R)0;
      if (live && n >= 0 && n < a.T) v = Conv<R>::from(ldg_stream(row + n));
      xs[j * W + c] = v;
    }
  }
}

template <int NTX, typename C>
__device__ __forceinline__ void stage_spatial(C* Ss, const FadingArgs& a, int b) {
  using R = decltype(C::x);
  const double2* Sb = a.spatial + (size_t)b * a.nrx * a.ntx;
  for (int c = threadIdx.x; c < a.nrx * NTX; c += kThreads) {
    const int irx = c / NTX, j = c - irx * NTX;
    C v;
    v.x = (R)0;
    v.y = (R)0;
    if (j < a.ntx_chunk) v = Conv<R>::from(Sb[irx * a.ntx + a.tx0 + j]);
    Ss[c] = v;
  }
}

// y[b, irx, m] (=|+=) sum_j S[irx, j] z[j]
template <int NTX, typename C, typename IO>
__device__ __forceinline__ void spatial_store(const C* Ss, const C (&z)[NTX], const FadingArgs& a, int b,
                                              int m) {
  using R = decltype(C::x);
  const int Tout = a.T + a.D;
  IO* yb = reinterpret_cast<IO*>(a.y) + (size_t)b * a.nrx * Tout + m;
  for (int irx = 0; irx < a.nrx; ++irx) {
    C acc;
    acc.x = (R)0;
    acc.y = (R)0;
#pragma unroll
    for (int j = 0; j < NTX; ++j) cmac<R>(acc, Ss[irx * NTX + j], z[j]);
    IO* dst = yb + (size_t)irx * Tout;
    if (a.accumulate) {
      const IO old = *dst;
      acc.x += (R)old.x;
      acc.y += (R)old.y;
    }
    stg_stream(dst, IoConv<IO>::make(acc.x, acc.y));
  }
}

// ------------------------------------------------------------------------------------------------
// K3 + K4 (polynomial tap gains).  grid (ntiles, B), 256 threads, R samples per thread per pass.
template <int NTX, int P, int R, typename IO>
__global__ void __launch_bounds__(kThreads) tdl_poly_kernel(const FadingArgs a,
                                                            const __grid_constant__ DelayTable dt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int W = a.tile + a.Dpad;
  const int G = dt.num_groups;
  float2* xs = reinterpret_cast<float2*>(smem_raw);
  float2* cs = xs + NTX * W;
  float2* Ss = cs + G * P;

  stage_x_tile<NTX, float2, IO>(xs, a, b, q, W);
  {
    const float2* cb = a.coef + ((size_t)b * a.ntiles + q) * a.coef_stride;
    for (int c = tid; c < G * P; c += kThreads) cs[c] = cb[c];
  }
  stage_spatial<NTX, float2>(Ss, a, b);
  __syncthreads();

  const int Tout = a.T + a.D;
  const float inv_tile = 1.0f / (float)a.tile;
  const float half = 0.5f * (float)a.tile;
  for (int base = 0; base < a.tile; base += kThreads * R) {
    if ((long long)q * a.tile + base >= Tout) break;
    // tile is a multiple of 256, so the number of live sample slots is uniform over the CTA
    const int nu = min(R, (a.tile - base) / kThreads);
    float2 z[R][NTX];
    float rr[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      rr[u] = ((float)(base + u * kThreads + tid) - half) * inv_tile;
#pragma unroll
      for (int j = 0; j < NTX; ++j) z[u][j] = make_float2(0.f, 0.f);
    }
    for (int g = 0; g < G; ++g) {
      const int off = base + tid + a.Dpad - dt.group_delay[g];
      float2 c[P];
#pragma unroll
      for (int p = 0; p < P; ++p) c[p] = cs[g * P + p];
      float2 h[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        float2 v = c[P - 1];
#pragma unroll
        for (int p = P - 2; p >= 0; --p) {
          v.x = fmaf(v.x, rr[u], c[p].x);
          v.y = fmaf(v.y, rr[u], c[p].y);
        }
        h[u] = v;
      }
#pragma unroll
      for (int j = 0; j < NTX; ++j) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          if (u < nu) {  // keeps off + u*256 < W; samples outside the frame are zeros in xs
            const float2 xv = xs[j * W + off + u * kThreads];
            cmac<float>(z[u][j], xv, h[u]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int m = q * a.tile + base + u * kThreads + tid;
      if (u < nu && m < Tout) spatial_store<NTX, float2, IO>(Ss, z[u], a, b, m);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Direct evaluation: one sincos per (tap, sinusoid, sample).  tile == 256, one sample per thread.
//   REAL = float : phases kept as 32-bit binary angles (2^32 = one turn).  The start phase of each
//                  sinusoid at the first sample of the tile is formed in FP64 and rounded once; stepping by
//                  the rounded increment w for < 256 samples adds < 256 * 2^-33 turns = 1.9e-7 rad.
//                  Wrap-around of the unsigned accumulator IS the range reduction; __sincosf then sees
//                  |angle| <= pi where its absolute error is 2^-21.4.
//   REAL = double: theta = omega * n + phi, sincos in FP64 -- the parity mode.
template <typename REAL> struct SinParam;
template <> struct SinParam<float> { using type = uint2; };     // (phase0, increment) binary angles
template <> struct SinParam<double> { using type = double2; };  // (phi, omega)

constexpr int kDirectParamBytes = 32 * 1024;  // per tap-chunk staging budget in shared memory

// STAGED = false: x is read from global memory (through L1 / L2) instead of a staged tile + delay halo -- the form that
// takes ANY delay spread (a 60 000-sample halo of four antennas would need 2 MB of shared memory); the planner's last
// resort before refusing a link, instantiated for REAL = double only.
template <int NTX, typename REAL, typename IO, bool STAGED = true>
__global__ void __launch_bounds__(kThreads) tdl_direct_kernel(const FadingArgs a,
                                                              const __grid_constant__ DelayTable dt,
                                                              const int taps_per_chunk) {
  using C = typename Cplx<REAL>::type;
  using SP = typename SinParam<REAL>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int W = a.tile + a.Dpad;
  const int K = a.K;
  C* xs = reinterpret_cast<C*>(smem_raw);
  C* Ss = STAGED ? xs + NTX * W : xs;
  SP* sp = reinterpret_cast<SP*>(Ss + a.nrx * NTX);
  REAL* am = reinterpret_cast<REAL*>(sp + taps_per_chunk * K);

  if constexpr (STAGED) stage_x_tile<NTX, C, IO>(xs, a, b, q, W);
  stage_spatial<NTX, C>(Ss, a, b);
  const IO* xg = reinterpret_cast<const IO*>(a.x) + ((size_t)b * a.ntx + a.tx0) * a.T;  // unstaged reads

  const double* om_b = a.omega + (size_t)b * a.L * K;
  const double* ph_b = a.phi + (size_t)b * a.L * K;
  const double* am_b = a.amp + (size_t)b * a.L * 2;
  const int m = q * a.tile + tid;  // output sample of this thread

  C z[NTX];
#pragma unroll
  for (int j = 0; j < NTX; ++j) {
    z[j].x = (REAL)0;
    z[j].y = (REAL)0;
  }

  for (int l0 = 0; l0 < a.L; l0 += taps_per_chunk) {
    const int lc = min(taps_per_chunk, a.L - l0);
    __syncthreads();  // previous chunk fully consumed (also orders the staging above on the first pass)
    for (int idx = tid; idx < lc * K; idx += kThreads) {
      const int l = l0 + idx / K;
      const double om = om_b[(size_t)l0 * K + idx];
      const double ph = ph_b[(size_t)l0 * K + idx];
      if constexpr (sizeof(REAL) == 4) {
        // phase at the first output sample of the tile, i.e. input index n = q*tile - d_l
        const double n_first = (double)q * a.tile - (double)dt.tap_delay[l];
        double t = fma(om, n_first, ph) * kInvTwoPi;
        t -= floor(t);
        double wq = om * kInvTwoPi;
        wq -= rint(wq);
        uint2 v;
        v.x = (uint32_t)(unsigned long long)(t * 4294967296.0);
        v.y = (uint32_t)(long long)rint(wq * 4294967296.0);
        sp[idx] = v;
      } else {
        sp[idx] = make_double2(ph, om);
      }
    }
    for (int idx = tid; idx < lc * 2; idx += kThreads) am[idx] = (REAL)am_b[(size_t)l0 * 2 + idx];
    __syncthreads();

    for (int ll = 0; ll < lc; ++ll) {
      const int d = dt.tap_delay[l0 + ll];
      const SP* spl = sp + ll * K;
      REAL sr = 0, si = 0, lr = 0, li = 0;
      if constexpr (sizeof(REAL) == 4) {
        const float scale = 1.4629180792671596e-9f;  // 2*pi / 2^32
        {
          const uint2 v = spl[0];
          const int ang = (int)(v.x + v.y * (uint32_t)tid);
          __sincosf((float)ang * scale, &li, &lr);
        }
#pragma unroll 4
        for (int k = 1; k < K; ++k) {
          const uint2 v = spl[k];
          const int ang = (int)(v.x + v.y * (uint32_t)tid);
          float s, c;
          __sincosf((float)ang * scale, &s, &c);
          sr += c;
          si += s;
        }
      } else {
        const double n = (double)(m - d);
        {
          const double2 v = spl[0];
          sincos(fma(v.y, n, v.x), &li, &lr);
        }
        for (int k = 1; k < K; ++k) {
          const double2 v = spl[k];
          double s, c;
          sincos(fma(v.y, n, v.x), &s, &c);
          sr += c;
          si += s;
        }
      }
      C h;
      h.x = am[2 * ll] * lr + am[2 * ll + 1] * sr;
      h.y = am[2 * ll] * li + am[2 * ll + 1] * si;
      if constexpr (STAGED) {
        const int off = tid + a.Dpad - d;
#pragma unroll
        for (int j = 0; j < NTX; ++j) cmac<REAL>(z[j], xs[j * W + off], h);
      } else {
        const int n = m - d;
        if (n >= 0 && n < a.T) {
#pragma unroll
          for (int j = 0; j < NTX; ++j)
            if (j < a.ntx_chunk) cmac<REAL>(z[j], Conv<REAL>::from(xg[(size_t)j * a.T + n]), h);
        }
      }
    }
  }
  if (m < a.T + a.D) spatial_store<NTX, C, IO>(Ss, z, a, b, m);
}

// ------------------------------------------------------------------------------------------------
// SISO tap gains for the channel state: h[b, g, n] = sum_{l in g} h_l[n], n = 0..T-1 (input time).
template <typename REAL, typename IO>
__global__ void __launch_bounds__(kThreads) sos_state_kernel(const FadingArgs a,
                                                             const __grid_constant__ DelayTable dt) {
  // grid.x = B * G * ntiles with ntiles = ceil(T / 256)
  const int bg = blockIdx.x / a.ntiles;
  const int b = bg / dt.num_groups, g = bg - b * dt.num_groups;
  const int n = (blockIdx.x - bg * a.ntiles) * kThreads + threadIdx.x;
  if (n >= a.T) return;
  const int K = a.K;
  const int l0 = dt.group_start[g], l1 = dt.group_start[g + 1];
  const double* om_b = a.omega + (size_t)b * a.L * K;
  const double* ph_b = a.phi + (size_t)b * a.L * K;
  const double* am_b = a.amp + (size_t)b * a.L * 2;
  REAL hr = 0, hi = 0;
  for (int l = l0; l < l1; ++l) {
    REAL sr = 0, si = 0, lr = 0, li = 0;
    for (int k = 0; k < K; ++k) {
      const double th = fma(om_b[l * K + k], (double)n, ph_b[l * K + k]);
      REAL s, c;
      if constexpr (sizeof(REAL) == 4) {
        double t = th * kInvTwoPi;
        t -= rint(t);
        sincosf((float)(t * kTwoPi), &s, &c);
      } else {
        sincos(th, &s, &c);
      }
      if (k == 0) {
        lr = c;
        li = s;
      } else {
        sr += c;
        si += s;
      }
    }
    hr += (REAL)am_b[2 * l] * lr + (REAL)am_b[2 * l + 1] * sr;
    hi += (REAL)am_b[2 * l] * li + (REAL)am_b[2 * l + 1] * si;
  }
  IO* out = reinterpret_cast<IO*>(a.y) + ((size_t)b * dt.num_groups + g) * a.T + n;
  *out = IoConv<IO>::make(hr, hi);
}

// ---- per-NTX launchers implemented in fading_inst_ntx*.cu -----------------------------------------
template <int NTX>
int launch_tdl_poly(int P, bool io128, const FadingArgs& a, const DelayTable& dt, size_t smem,
                    cudaStream_t st);
template <int NTX>
int launch_tdl_direct(bool f64, bool io128, bool unstaged, const FadingArgs& a, const DelayTable& dt, int taps_per_chunk,
                      size_t smem, cudaStream_t st);
template <int NTX> constexpr int poly_samples_per_thread() { return NTX <= 2 ? 4 : (NTX <= 4 ? 2 : 1); }

}  // namespace hb
