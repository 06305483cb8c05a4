// K3 + K4, sliding-window form (sm_100a).
//
//   y[:, m] = S @ sum_g h_g(m) x[:, m - d_g]        (gather form of fading.py:385-393, see fading_kernels.cuh)
//
// The gather kernel (tdl_poly_kernel) re-reads one x element from shared memory per (delay group, antenna,
// output) and is bound by shared-memory wavefronts (profiles/r01_ncu_tdl_poly_ntx4_p3_r2.md).  Here every
// thread owns R CONSECUTIVE outputs m0..m0+R-1 and walks the delay axis d = 0..Dmax once, keeping the R inputs
// x[m0 - d .. m0 - d + R - 1] of every antenna in registers: stepping d -> d+1 shifts ONE new element in
// (one LDS.64 per antenna).  Shared-memory reads per output drop from G to (Dmax + R) / R per antenna
// (C2: 16 -> 6.5).  Static register indexing is obtained by unrolling the walk R-fold: at delay d = R c + s the
// input of output u sits in window slot (u - s) mod R.
//
// Shared-memory layout ("polyphase"): element e of the staged tile [q tile - Dpad, q tile + tile) of antenna j
// lives in plane e mod R at index e / R, so that the 32 lanes of a warp (outputs R apart) read consecutive 8-byte
// words -- conflict-free -- and the plane pitch PL = 16/R (mod 16) keeps the coalesced staging writes
// conflict-free as well.  The same layout, reused after the delay walk, transposes the results for fully
// coalesced 8-byte global stores.
//
// Arithmetic: packed FFMA2 (fma.rn.f32x2).  A complex MAC  acc += x * h  is
//   acc(re,im) += x(re,im) * h.re            (scalar-broadcast operand)
//   acc(re,im) += x(im,re) * (-h.im, h.im)   (ptxas folds the swap into the .LO_HI operand modifier)
// i.e. 2 issue slots instead of 4 -- the FP32 pipe does the same flops, but the issue port is left free for the
// shared-memory loads and the address arithmetic (measured: tools/microbench/pipes.cu, FFMA2 = 64 lanes/clk/SM).
#pragma once
#include "fading_kernels.cuh"

namespace hb {

typedef unsigned long long u64;

constexpr int kWindowMaxDelay = 1023;  // delay walk masks cover d = 0..1023

// Launch-uniform plan of the delay walk, by value in kernel parameter space.
struct WindowPlan {
  int32_t num_groups;
  int32_t dmax;       // largest group delay
  int32_t nblk;       // dmax / R + 1 blocks of R delays
  int32_t plane;      // plane pitch PL in elements
  int32_t poly_tile;  // samples per Taylor expansion window (multiple of the CTA tile)
  int32_t npoly;      // expansion windows per link
  uint32_t present[(kWindowMaxDelay + 1) / 32];  // bit d: some tap has rounded delay d
  uint32_t load[(kWindowMaxDelay + 1) / 32];     // bit d: x[m0 - d] is needed by a present delay in [d, d+R)
};

__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 upk2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 swap2(u64 v) {
  const float2 f = upk2(v);
  return pk2(f.y, f.x);
}
// acc += x * h with h given as (h.re broadcast, (-h.im, h.im))
__device__ __forceinline__ void cmac2(u64& acc, u64 x, u64 hre, u64 him) {
  acc = fma2(x, hre, acc);
  acc = fma2(swap2(x), him, acc);
}

__device__ __forceinline__ u64 lds_pair(const float2* p) {
  const float2 v = *p;
  return pk2(v.x, v.y);
}

template <int NTX, int P, int R, typename IO>
__global__ void __launch_bounds__(128, (NTX * R >= 32) ? 3 : 4)
    tdl_window_kernel(const FadingArgs a, const __grid_constant__ WindowPlan wp) {
  static_assert(R == 4 || R == 8 || R == 16, "R must divide 32");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NT = blockDim.x;
  const int tile = NT * R;
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int G = wp.num_groups, PL = wp.plane;
  const int Dq = wp.nblk;  // halo planes-worth: Dpad = R * nblk
  const int W = tile + R * Dq;
  const int Tout = a.T + a.D;

  float2* xs = reinterpret_cast<float2*>(smem_raw);  // [NTX][R][PL]
  float2* cs = xs + (size_t)NTX * R * PL;             // [G][P]
  float2* Ss = cs + G * P;                            // [nrx][NTX]

  // ---- stage: x tile (+ halo) into the polyphase planes, Taylor coefficients, spatial matrix ------------
  {
    const IO* xb = reinterpret_cast<const IO*>(a.x) + ((size_t)b * a.ntx + a.tx0) * a.T;
    const int n0 = q * tile - R * Dq;
#pragma unroll
    for (int j = 0; j < NTX; ++j) {
      const bool live = j < a.ntx_chunk;
      const IO* row = xb + (size_t)j * a.T;
      float2* xj = xs + (size_t)j * R * PL;
#pragma unroll 8
      for (int e = tid; e < W; e += NT) {
        const int n = n0 + e;
        float2 v = make_float2(0.f, 0.f);
        if (live && n >= 0 && n < a.T) v = to_c32(ldg_stream(row + n));
        xj[(e & (R - 1)) * PL + (e / R)] = v;
      }
    }
    const int qp = (q * tile) / wp.poly_tile;
    const float2* cb = a.coef + ((size_t)b * wp.npoly + qp) * G * P;
    for (int c = tid; c < G * P; c += NT) cs[c] = cb[c];
    const double2* Sb = a.spatial + (size_t)b * a.nrx * a.ntx;
    for (int c = tid; c < a.nrx * NTX; c += NT) {
      const int irx = c / NTX, j = c - irx * NTX;
      float2 v = make_float2(0.f, 0.f);
      if (j < a.ntx_chunk) v = to_c32(Sb[irx * a.ntx + a.tx0 + j]);
      Ss[c] = v;
    }
  }
  __syncthreads();

  const int m0 = q * tile + R * tid;  // first output of this thread
  const bool active = m0 < Tout;      // warp-uniform except in one warp of the last tile

  u64 acc[R][NTX];
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int j = 0; j < NTX; ++j) acc[u][j] = 0ull;

  if (active) {
    // normalized expansion coordinate of each owned output
    float rr[R];
    {
      const int qp = (q * tile) / wp.poly_tile;
      const float inv = 1.0f / (float)wp.poly_tile;
      const float r0 = ((float)(m0 - qp * wp.poly_tile) - 0.5f * (float)wp.poly_tile) * inv;
#pragma unroll
      for (int u = 0; u < R; ++u) rr[u] = fmaf((float)u, inv, r0);
    }
    // window at d = 0: slot u holds element e = Dpad + R tid + u  (plane u, index tid + Dq)
    u64 w[NTX][R];
    const float2* xt = xs + tid + Dq;
#pragma unroll
    for (int j = 0; j < NTX; ++j)
#pragma unroll
      for (int u = 0; u < R; ++u) w[j][u] = lds_pair(xt + (j * R + u) * PL);

    int g = 0;
    for (int c = 0; c < wp.nblk; ++c) {
      const int bit0 = c * R;
      const uint32_t pm = (wp.present[bit0 >> 5] >> (bit0 & 31)) & ((1u << R) - 1u);
      const uint32_t lm = (wp.load[bit0 >> 5] >> (bit0 & 31)) & ((1u << R) - 1u);
      if ((pm | lm) == 0u) continue;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        if (((lm >> s) & 1u) && (c | s)) {
          // new element x[m0 - d], d = R c + s: plane (R - s) % R, index tid + Dq - c - (s > 0)
          const int k = (R - s) % R;
          const float2* src = xt + k * PL - c - (s > 0 ? 1 : 0);
#pragma unroll
          for (int j = 0; j < NTX; ++j) w[j][k] = lds_pair(src + j * R * PL);
        }
        if ((pm >> s) & 1u) {
          u64 cf[P];
#pragma unroll
          for (int p = 0; p < P; ++p) cf[p] = lds_pair(cs + g * P + p);
          ++g;
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const u64 rb = pk2(rr[u], rr[u]);
            u64 hv = cf[P - 1];
#pragma unroll
            for (int p = P - 2; p >= 0; --p) hv = fma2(hv, rb, cf[p]);
            const float2 h = upk2(hv);
            const u64 hre = pk2(h.x, h.x);
            const u64 him = pk2(-h.y, h.y);
#pragma unroll
            for (int j = 0; j < NTX; ++j) cmac2(acc[u][j], w[j][(u - s + R) % R], hre, him);
          }
        }
      }
    }
  }
  __syncthreads();  // every thread is done with the x planes: reuse them for the output transpose

  // ---- spatial mix  y[irx] = sum_j S[irx][j] z[j], NTX receive streams at a time, transposed through the
  //      planes so that the global stores are coalesced 8-byte (16-byte for complex128) accesses ------------
  IO* yb = reinterpret_cast<IO*>(a.y) + (size_t)b * a.nrx * Tout;
  const int mbase = q * tile;
  const int live = min(tile, Tout - mbase);
  for (int irx0 = 0; irx0 < a.nrx; irx0 += NTX) {
    const int nr = min(NTX, a.nrx - irx0);
    if (irx0 > 0) __syncthreads();  // previous chunk stored
    if (active) {
      for (int i = 0; i < nr; ++i) {
        u64 yv[R];
#pragma unroll
        for (int u = 0; u < R; ++u) yv[u] = 0ull;
#pragma unroll
        for (int j = 0; j < NTX; ++j) {
          const float2 s = Ss[(irx0 + i) * NTX + j];
          const u64 sre = pk2(s.x, s.x), sim = pk2(-s.y, s.y);
#pragma unroll
          for (int u = 0; u < R; ++u) cmac2(yv[u], acc[u][j], sre, sim);
        }
        float2* yt = xs + (size_t)i * R * PL + tid;
#pragma unroll
        for (int u = 0; u < R; ++u) yt[u * PL] = upk2(yv[u]);
      }
    }
    __syncthreads();
    for (int i = 0; i < nr; ++i) {
      const float2* yr = xs + (size_t)i * R * PL;
      IO* dst = yb + (size_t)(irx0 + i) * Tout + mbase;
#pragma unroll 4
      for (int e = tid; e < live; e += NT) {
        float2 v = yr[(e & (R - 1)) * PL + (e / R)];
        if (a.accumulate) {
          const IO old = dst[e];
          v.x += (float)old.x;
          v.y += (float)old.y;
        }
        stg_stream(dst + e, IoConv<IO>::make(v.x, v.y));
      }
    }
  }
}

template <int NTX> constexpr int window_samples_per_thread() { return NTX <= 4 ? 8 : 4; }

template <int NTX>
int launch_tdl_window(int P, bool io128, const FadingArgs& a, const WindowPlan& wp, int threads, size_t smem,
                      cudaStream_t st);

}  // namespace hb
