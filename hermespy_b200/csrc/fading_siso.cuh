// K3 for single-antenna links (sm_100a): y[m] = s * sum_g h_g(m) x[m - d_g]           (fading.py:385-393, Ntx = Nrx = 1)
//
// The MIMO kernels (fading_window.cuh, fading_tma.cuh) pack (re, im) of one sample into an FFMA2 and share the tap gain
// of a (group, output) among the antennas; with one antenna nothing is shared and the pack / swap / negate moves around
// every complex MAC outnumber the FMAs (C5, ncu: 32 % FFMA2, 19 % MOV, 12 % LOP3).  This kernel packs along TIME instead:
//
// * x is staged PLANAR (re plane, im plane), each plane twice: as is, and shifted by one sample.  A thread owns pairs of
//   consecutive outputs (o, o + 1), o even; the inputs of a pair at delay d are one aligned 8-byte load from the copy whose
//   shift equals the parity of d -- (xr[k], xr[k+1]) arrive as the two halves of one FFMA2 operand, no move.
// * tap gains of the pair by Horner in the window coordinate, packed over the two outputs (coefficients are staged
//   pre-duplicated (re, re, im, im), one broadcast LDS.128 each);
// * four FFMA2 per pair and group into xr*hr, xi*hi, xr*hi, xi*hr; the sign of Re = xr*hr - xi*hi is applied once at the end.
// Per (group, output): 4 + 2 (P - 1) lane-FMAs... (P = 3: 4 FFMA2 per output), 8 bytes of shared memory, no other instruction
// but the address of the pair.  The 32 lanes of a warp read 256 contiguous bytes: conflict-free.
//
// Persistent CTAs walk (link, tile of 512 KP outputs) items, 256 threads, KP pairs per thread 512 samples apart; the next
// item's inputs are loaded into registers before the current one is computed (ncu of the one-tile-per-CTA form: 49 % of the
// stall samples sat in the staging phase).  Shared memory:
// 4 planes x (tile + Dpad + 2) floats + G P float4.  Algorithmic bytes: 8 (T + T + D) per link (c64).  Bounds for C5
// (15 delay groups): HBM 0.71 clk / sample / SM, shared memory 0.94, FP32 pipe 0.94.
#pragma once
#include "fading_kernels.cuh"
#include "fading_window.cuh"

namespace hb {

constexpr int kSisoThreads = 256;

template <int P, int KP, typename IO>
__global__ void __launch_bounds__(kSisoThreads) tdl_siso_kernel(const FadingArgs a, const __grid_constant__ DelayTable dt,
                                                                const int poly_tile, const int npoly) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TS = 512 * KP;
  constexpr int NPRE = 2 * KP + 2;  // register-prefetched input samples per thread: covers Dpad <= 512
  const int tid = threadIdx.x;
  const int G = dt.num_groups;
  const int W = TS + a.Dpad;  // Dpad is even
  const int WP = W + 2;       // plane length in floats (even: every plane starts 8-byte aligned)
  float* planes = reinterpret_cast<float*>(smem_raw);  // [copy 0: re | im][copy 1 (shifted by one sample): re | im]
  float4* cs = reinterpret_cast<float4*>(planes + 4 * WP);
  const int Tout = a.T + a.D;
  const int total = a.B * a.ntiles;
  const float inv = 1.0f / (float)poly_tile;
  const uint32_t p0 = smem_u32(planes);
  if (tid == 0) {
    planes[2 * WP] = 0.f;
    planes[3 * WP] = 0.f;
  }

  // Persistent CTA: the inputs of the NEXT tile are in flight (registers) while this one is computed, so the HBM latency
  // of the staging phase is hidden without a second shared-memory buffer.
  float2 pre[NPRE], cpre;
  auto load_x = [&](const IO* xb, int n) {
    float2 v = make_float2(0.f, 0.f);
    if (n >= 0 && n < a.T) v = to_c32(ldg_stream(xb + n));
    return v;
  };
  auto prefetch = [&](int t) {
    const int b = t / a.ntiles, q = t - b * a.ntiles;
    const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.T;
    const int n0 = q * TS - a.Dpad;
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int i = tid + kSisoThreads * k;
      pre[k] = i < W ? load_x(xb, n0 + i) : make_float2(0.f, 0.f);
    }
    const int qp = (q * TS) / poly_tile;
    cpre = tid < G * P ? a.coef[((size_t)b * npoly + qp) * a.coef_stride + tid] : make_float2(0.f, 0.f);
  };
  auto put = [&](int i, float2 v) {
    planes[i] = v.x;
    planes[WP + i] = v.y;
    planes[2 * WP + i + 1] = v.x;
    planes[3 * WP + i + 1] = v.y;
  };

  int t = blockIdx.x;
  if (t < total) prefetch(t);
  for (; t < total; t += gridDim.x) {
    const int b = t / a.ntiles, q = t - b * a.ntiles;
    const int qp = (q * TS) / poly_tile;
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int i = tid + kSisoThreads * k;
      if (i < W) put(i, pre[k]);
    }
    if (tid < G * P) cs[tid] = make_float4(cpre.x, cpre.x, cpre.y, cpre.y);
    if (W > kSisoThreads * NPRE || G * P > kSisoThreads) {  // long delay spreads / many groups: the rest, unhidden
      const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.T;
      for (int i = tid + kSisoThreads * NPRE; i < W; i += kSisoThreads) put(i, load_x(xb, q * TS - a.Dpad + i));
      const float2* cb = a.coef + ((size_t)b * npoly + qp) * a.coef_stride;
      for (int c = tid + kSisoThreads; c < G * P; c += kSisoThreads) {
        const float2 v = cb[c];
        cs[c] = make_float4(v.x, v.x, v.y, v.y);
      }
    }
    __syncthreads();
    if (t + (int)gridDim.x < total) prefetch(t + gridDim.x);

    u64 r2[KP], arr[KP], aii[KP], ari[KP], air[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int m = q * TS + 2 * tid + 512 * k;
      const float r = ((float)(m - qp * poly_tile) - 0.5f * (float)poly_tile) * inv;
      r2[k] = pk2(r, r + inv);
      arr[k] = aii[k] = ari[k] = air[k] = 0ull;
    }
    for (int g = 0; g < G; ++g) {
      const int d = dt.group_delay[g];
      const int cpy = d & 1;
      // pair (o, o + 1) at delay d: inputs at i = o + Dpad - d; odd d reads the shifted copy at i + 1
      const uint32_t ar = p0 + 4u * (uint32_t)(cpy * 2 * WP + 2 * tid + a.Dpad - d + cpy);
      const uint32_t ai = ar + 4u * (uint32_t)WP;
      u64 crr[P], cii[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float4 c = cs[g * P + p];
        crr[p] = pk2(c.x, c.y);
        cii[p] = pk2(c.z, c.w);
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        u64 xr, xi;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(xr) : "r"(ar + 2048u * k));
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(xi) : "r"(ai + 2048u * k));
        u64 hr = crr[P - 1], hi = cii[P - 1];
#pragma unroll
        for (int p = P - 2; p >= 0; --p) {
          hr = fma2(hr, r2[k], crr[p]);
          hi = fma2(hi, r2[k], cii[p]);
        }
        arr[k] = fma2(xr, hr, arr[k]);
        aii[k] = fma2(xi, hi, aii[k]);
        ari[k] = fma2(xr, hi, ari[k]);
        air[k] = fma2(xi, hr, air[k]);
      }
    }

    const double2 sd = a.spatial[b];
    const float sr = (float)sd.x, si = (float)sd.y;
    IO* yb = reinterpret_cast<IO*>(a.y) + (size_t)b * Tout;
    const bool pair_ok = (((size_t)b * Tout) & 1) == 0 && (reinterpret_cast<uintptr_t>(a.y) & 15) == 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int m = q * TS + 2 * tid + 512 * k;
      if (m < Tout) {
        const float2 rr = upk2(arr[k]), ii = upk2(aii[k]), ri = upk2(ari[k]), ir = upk2(air[k]);
        const float re0 = rr.x - ii.x, im0 = ri.x + ir.x, re1 = rr.y - ii.y, im1 = ri.y + ir.y;
        const float y0r = sr * re0 - si * im0, y0i = sr * im0 + si * re0;
        const float y1r = sr * re1 - si * im1, y1i = sr * im1 + si * re1;
        if constexpr (sizeof(IO) == 8) {
          if (pair_ok && m + 1 < Tout) {
            stg_stream4(reinterpret_cast<float2*>(yb + m), pk2(y0r, y0i), pk2(y1r, y1i));
          } else {
            stg_stream(yb + m, IoConv<IO>::make(y0r, y0i));
            if (m + 1 < Tout) stg_stream(yb + m + 1, IoConv<IO>::make(y1r, y1i));
          }
        } else {
          stg_stream(yb + m, IoConv<IO>::make(y0r, y0i));
          if (m + 1 < Tout) stg_stream(yb + m + 1, IoConv<IO>::make(y1r, y1i));
        }
      }
    }
    __syncthreads();  // every warp is done with this tile's planes before the next one overwrites them
  }
}

inline size_t siso_smem_bytes(int tile, int Dpad, int G, int P) {
  return 16 * (size_t)(tile + Dpad + 2) + 16 * (size_t)G * P;
}

// defined in fading_siso.cu
int launch_tdl_siso(int P, int tile, bool io128, const FadingArgs& a, const DelayTable& dt, int poly_tile, int npoly, size_t smem,
                    cudaStream_t st);

}  // namespace hb
