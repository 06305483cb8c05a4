// K3 + K4, sliding-window form (sm_100a).
//
//   y[:, m] = S @ sum_g h_g(m) x[:, m - d_g]        (gather form of fading.py:385-393, see fading_kernels.cuh)
//
// The gather kernel (tdl_poly_kernel) re-reads one x element from shared memory per (delay group, antenna,
// output) and is bound by shared-memory wavefronts (profiles/r01_ncu_tdl_poly_ntx4_p3_r2.md).  Here every
// thread owns R CONSECUTIVE outputs m0..m0+R-1 and walks the delay axis once, keeping the R inputs
// x[m0 - d .. m0 - d + R - 1] of every antenna in registers: stepping d -> d+1 shifts ONE new element in.
// Shared-memory reads per output drop from G to (Dmax + R) / R per antenna (C2: 16 -> 6.5).
//
// Static register indexing.  The walk is unrolled R-fold: at delay d = R c + s (c: block, s: phase) the input of
// output u sits in window slot (u - s) mod R and the element entering (u = 0) goes to slot (R - s) mod R --
// always the slot e mod R of its index e in the staged tile, so a slot is simply overwritten in place.
//
// Shared-memory layout ("polyphase planes of antenna pairs"): the staged tile holds elements e = 0 .. R (NT + Dq)
// - 1 of every antenna, element e <-> input sample n = q tile - R Dq + e.  Row i = e / R belongs to thread i - Dq
// (+ Dq halo rows), plane k = e mod R.  Antennas (2p, 2p+1) are interleaved, xs[p][k][i] = one float4
// (re_2p, im_2p, re_2p+1, im_2p+1): the 32 lanes of a warp read consecutive 16-byte words (conflict-free
// LDS.128, one per antenna pair and delay step), and a row is 64 contiguous bytes of global memory per antenna,
// copied by its thread with 8-byte cp.async (zero-filled outside the frame) -- no index arithmetic, no
// alignment requirement beyond the element size.  Single-antenna chunks use 8-byte planes xs[k][i].
//
// Arithmetic: packed FFMA2 (fma.rn.f32x2).  A complex MAC  acc += x * h  is
//   acc(re,im) += x(re,im) * h.re            (scalar-broadcast operand)
//   acc(re,im) += x(im,re) * (-h.im, h.im)   (ptxas folds swap and sign into the .LO_HI / .NP operand modifiers)
// i.e. 2 issue slots instead of 4; the FP32 pipe does the same flops (tools/microbench/pipes.cu: FFMA2 issues at
// half rate), but the issue port is left free for shared-memory loads and address arithmetic.
//
// Results leave the registers directly: every thread stores its R consecutive outputs of each receive stream
// with 16-byte stores (64 contiguous bytes per thread and stream).
#pragma once
#include "fading_kernels.cuh"

namespace hb {

typedef unsigned long long u64;

constexpr int kWindowMaxDelay = 1023;   // walk masks cover d = 0..1023
constexpr int kWindowThreads = 128;     // maximum CTA size (32 / 64 for short frames)
constexpr int kWindowHaloSmall = 16;    // halo rows of the small-delay instantiation (d' < 16 R)
constexpr int kWindowHaloLarge = 1024 / 4 + 2;

// Launch-uniform plan of the delay walk, by value in kernel parameter space.
struct WindowPlan {
  int32_t num_groups;
  int32_t nblk;       // blocks of R delays: dmax / R + 1
  int32_t poly_tile;  // samples per Taylor expansion window (multiple of the CTA tile)
  int32_t npoly;      // expansion windows per link
  // per block c: bits 0..R-1 "a tap has delay d = R c + s", bits 8..8+R "x[m0 - d] is needed (by a present delay in
  // [d, d + R))" for d = R c + s, s = 0..R (bit 8+R repeats bit 8 of the next block: that element is fetched
  // from within this block)
  uint32_t mask[kWindowMaxDelay / 4 + 2];
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

// 32-bit shared-window addressing with immediate offsets (no generic-address arithmetic in the hot loop)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int IMM>
__device__ __forceinline__ void lds_pair(uint32_t addr, u64& lo, u64& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2 + %3];" : "=l"(lo), "=l"(hi) : "r"(addr), "n"(IMM));
}
template <int IMM>
__device__ __forceinline__ u64 lds_one(uint32_t addr) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1 + %2];" : "=l"(v) : "r"(addr), "n"(IMM));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void stg_stream4(float2* p, u64 a, u64 b) {
  asm volatile("st.global.L1::no_allocate.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

__device__ __forceinline__ void cp_async8_full(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// Stage the rows of one antenna pair (or single antenna, ES == 8) of the tile starting at sample n0.  Work item
// = (row, antenna of the pair): consecutive lanes write consecutive 8-byte words of a plane (conflict-free) and
// read 64-byte spans of two global rows.  `dst0` is the shared address of (plane 0, row 0) of the pair.
template <int R, int PL, int ES, typename IO>
__device__ __forceinline__ void stage_pair(uint32_t dst0, const IO* rowA, const IO* rowB, bool liveA, bool liveB,
                                           int n0, int T, int nrows, int tid, int NT) {
  constexpr int PAR = ES / 8;
  const bool interior = n0 >= 0 && n0 + R * nrows <= T;  // CTA-uniform: no frame edge inside the tile
  for (int it = tid; it < nrows * PAR; it += NT) {
    const int i = PAR == 2 ? (it >> 1) : it;
    const int par = PAR == 2 ? (it & 1) : 0;
    const IO* row = par ? rowB : rowA;
    const bool live = par ? liveB : liveA;
    const int n = n0 + R * i;  // first sample of the row
    const uint32_t dst = dst0 + (uint32_t)i * ES + par * 8;
    if constexpr (sizeof(IO) == 8) {
      if (interior && live) {
        const IO* src = row + n;
#pragma unroll
        for (int k = 0; k < R; ++k) cp_async8_full(dst + k * PL * ES, src + k);
      } else {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int nn = n + k;
          const bool ok = live && nn >= 0 && nn < T;
          cp_async8(dst + k * PL * ES, ok ? (const void*)(row + nn) : (const void*)row, ok ? 8 : 0);
        }
      }
    } else {  // complex128 host layout: convert on the way in
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int nn = n + k;
        float2 v = make_float2(0.f, 0.f);
        if (live && nn >= 0 && nn < T) v = to_c32(ldg_stream(row + nn));
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst + k * PL * ES), "f"(v.x), "f"(v.y) : "memory");
      }
    }
  }
}

// LIN: the tap gain polynomial is evaluated once per thread and delay group (value and slope at the centre of the
// thread's R outputs) and extended linearly over them; hb_fading_plan enables it when the neglected curvature
// sqrt(N+1) ((R-1)/2 omega_max)^2 / 2 stays below the truncation target.
//
// SPLIT: threads per output row.  With SPLIT = 2 the two antenna pairs of a 4-antenna chunk belong to two lanes of
// the same warp (lane, lane ^ 16): each walks the delays for its own pair with half the accumulators and half the
// window (~100 registers instead of 160 -> 5 CTAs per SM instead of 3), forms the partial spatial mix of its pair
// for every receive stream, and the two exchange halves with one shuffle per value (reduce-scatter: the first lane
// finishes and stores the first half of the receive streams, its partner the second half).
template <int NTX, int P, int R, int HALO, bool LIN, int SPLIT, typename IO>
__global__ void __launch_bounds__(kWindowThreads, (NTX * R / SPLIT >= 32) ? 3 : (SPLIT == 2 ? 5 : 4))
    tdl_window_kernel(const FadingArgs a, const __grid_constant__ WindowPlan wp) {
  static_assert(R == 4 || R == 8, "R must be 4 or 8");
  static_assert(NTX == 1 || NTX % 2 == 0, "antennas are staged in pairs");
  static_assert(SPLIT == 1 || (SPLIT == 2 && NTX == 4), "SPLIT = 2 is written for 4-antenna chunks");
  constexpr int PL = kWindowThreads / SPLIT + HALO;  // rows per plane
  constexpr int ES = NTX == 1 ? 8 : 16;              // bytes per staged element (antenna pair)
  constexpr int NP = NTX == 1 ? 1 : NTX / 2;         // antenna pairs of the chunk
  constexpr int NA = NTX / SPLIT;                    // antennas per thread
  constexpr int NPT = NTX == 1 ? 1 : NA / 2;         // antenna pairs per thread
  constexpr int PS = R * PL * ES;                    // bytes per antenna pair
  constexpr int RW = 32 / SPLIT;                     // rows per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NT = blockDim.x;
  const int NR = NT / SPLIT;  // rows (threads along time) per CTA
  const int tile = NR * R;
  const int b = blockIdx.x / a.ntiles, q = blockIdx.x - b * a.ntiles, tid = threadIdx.x;
  const int lane = tid & 31;
  const int pidx = SPLIT == 1 ? 0 : lane / RW;                         // antenna pair group of this thread
  const int trow = SPLIT == 1 ? tid : (tid >> 5) * RW + (lane % RW);   // row of this thread
  const int G = wp.num_groups;
  const int Dq = wp.nblk;  // halo rows: Dp = R * nblk
  const int Tout = a.T + a.D;

  float2* cs = reinterpret_cast<float2*>(smem_raw + NP * PS);  // [G][P]
  float2* Ss = cs + (G + 1) * P;                                // [nrx][NTX]; one group of padding before it
  const uint32_t xs = smem_u32(smem_raw);

  // ---- stage: x tile (+ halo), Taylor coefficients, spatial matrix ------------------------------------------
  {
    const IO* xb = reinterpret_cast<const IO*>(a.x) + ((size_t)b * a.ntx + a.tx0) * a.T;
    const int n0 = q * tile - R * Dq;
#ifdef HB_ATTRIBUTION
    if (!(a.dbg & 1))
#endif
#pragma unroll
      for (int pr = 0; pr < NP; ++pr)
        stage_pair<R, PL, ES, IO>(xs + pr * PS, xb + (size_t)(2 * pr) * a.T, xb + (size_t)(2 * pr + 1) * a.T,
                                  2 * pr < a.ntx_chunk, 2 * pr + 1 < a.ntx_chunk, n0, a.T, NR + Dq, tid, NT);
    // coefficients ride the same asynchronous copy group as the tile; the spatial matrix needs a conversion, its
    // loads are issued before anything waits so that the CTA pays ONE memory round trip, not three
    const int qp = (q * tile) / wp.poly_tile;
    const float2* cb = a.coef + ((size_t)b * wp.npoly + qp) * a.coef_stride;
    const uint32_t csa0 = smem_u32(cs);
    for (int c = tid; c < G * P; c += NT) cp_async8_full(csa0 + c * 8, cb + c);
    const double2* Sb = a.spatial + (size_t)b * a.nrx * a.ntx;
    for (int c = tid; c < a.nrx * NTX; c += NT) {
      const int irx = c / NTX, j = c - irx * NTX;
      float2 v = make_float2(0.f, 0.f);
      if (j < a.ntx_chunk) v = to_c32(Sb[irx * a.ntx + a.tx0 + j]);
      Ss[c] = v;
    }
    cp_async_wait_all();
  }
  __syncthreads();

  const int m0 = q * tile + R * trow;  // first output of this thread
  const unsigned live_mask = __ballot_sync(0xffffffffu, m0 < Tout);
  if (m0 >= Tout) return;  // no barrier below; a row's lanes leave together

  u64 acc[R][NA];
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int j = 0; j < NA; ++j) acc[u][j] = 0ull;

  {
    // normalized expansion coordinate: centre of the R outputs (LIN) or every output
    const int qp = (q * tile) / wp.poly_tile;
    const float inv = 1.0f / (float)wp.poly_tile;
    const float r0 = ((float)(m0 - qp * wp.poly_tile) - 0.5f * (float)wp.poly_tile) * inv;
    const float rc = fmaf(0.5f * (float)(R - 1), inv, r0);
    float rr[LIN ? 1 : R];
    if constexpr (!LIN) {
#pragma unroll
      for (int u = 0; u < R; ++u) rr[u] = fmaf((float)u, inv, r0);
    }
    // window at d = 0: slot u holds element Dp + R trow + u (plane u, row trow + Dq)
    u64 w[NA][R];
    uint32_t xa = xs + (uint32_t)(trow + Dq) * ES + pidx * NPT * PS;
#pragma unroll
    for (int pr = 0; pr < NPT; ++pr)
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if constexpr (NTX == 1)
          w[0][u] = lds_one<0>(xa + u * PL * ES);
        else
          lds_pair<0>(xa + pr * PS + u * PL * ES, w[2 * pr][u], w[2 * pr + 1][u]);
      }

    // Software pipeline over the delay groups: the tap-gain polynomial of group g+1 is fetched (and, with LIN,
    // reduced to value + slope at the thread's centre) while the MACs of group g issue, and the element entering
    // at the next delay is loaded right after the MACs of output R-1 -- the last reader of its window slot.
    constexpr int NH = LIN ? 2 : P;
    uint32_t csa = smem_u32(cs);
    u64 hp[NH];
    auto prep = [&]() {  // consume the coefficients at csa (one group past the end is padding)
      u64 cf[P];
#pragma unroll
      for (int p = 0; p < P; ++p) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(cf[p]) : "r"(csa + p * 8));
      csa += P * 8;
      if constexpr (LIN) {
        static_assert(!LIN || P >= 3, "LIN only pays for P >= 3");
        const u64 rcb = pk2(rc, rc);
        const float fp = (float)(P - 1) * inv;
        u64 hs = fma2(cf[P - 1], pk2(fp, fp), 0ull);  // slope per output sample: inv * sum_p p c_p rc^(p-1)
#pragma unroll
        for (int p = P - 2; p >= 1; --p) {
          const float fq = (float)p * inv;
          hs = fma2(hs, rcb, fma2(cf[p], pk2(fq, fq), 0ull));
        }
        u64 hc = cf[P - 1];
#pragma unroll
        for (int p = P - 2; p >= 0; --p) hc = fma2(hc, rcb, cf[p]);
        hp[0] = hc;
        hp[1] = hs;
      } else {
#pragma unroll
        for (int p = 0; p < P; ++p) hp[p] = cf[p];
      }
    };
    prep();

#ifdef HB_ATTRIBUTION
    const int nblk = (a.dbg & 4) ? 0 : wp.nblk;
#else
    const int nblk = wp.nblk;
#endif
    for (int c = 0; c < nblk; ++c, xa -= ES) {
      const uint32_t mk = wp.mask[c];
      if (mk == 0u) continue;
      const uint32_t pm = mk & 0xffu, lm = mk >> 8;  // lm bit s: the element entering at phase s is needed
#pragma unroll
      for (int s = 0; s < R; ++s) {
        // element entering at the NEXT phase, x[m0 - (R c + s + 1)]: plane (R - s - 1), row trow + Dq - c - 1
        auto load_next = [&]() {
          if ((lm >> (s + 1)) & 1u) {
#pragma unroll
            for (int pr = 0; pr < NPT; ++pr) {
              const uint32_t ad = xa + pr * PS + (R - s - 1) * PL * ES;
              if constexpr (NTX == 1)
                w[0][R - s - 1] = lds_one<-ES>(ad);
              else
                lds_pair<-ES>(ad, w[2 * pr][R - s - 1], w[2 * pr + 1][R - s - 1]);
            }
          }
        };
        if ((pm >> s) & 1u) {
          u64 hq[NH];
#pragma unroll
          for (int i = 0; i < NH; ++i) hq[i] = hp[i];
          prep();  // next group
          auto mac_u = [&](int u) {
            u64 hv;
            if constexpr (LIN) {
              const float ku = (float)u - 0.5f * (float)(R - 1);  // immediate operand of the FFMA2
              hv = fma2(hq[1], pk2(ku, ku), hq[0]);
            } else {
              const u64 rb = pk2(rr[u], rr[u]);
              hv = hq[P - 1];
#pragma unroll
              for (int p = P - 2; p >= 0; --p) hv = fma2(hv, rb, hq[p]);
            }
            const float2 h = upk2(hv);
            const u64 hre = pk2(h.x, h.x);
            const u64 him = pk2(-h.y, h.y);
#pragma unroll
            for (int j = 0; j < NA; ++j) cmac2(acc[u][j], w[j][(u - s + R) % R], hre, him);
          };
          mac_u(R - 1);
          load_next();
#pragma unroll
          for (int u = 0; u < R - 1; ++u) mac_u(u);
        } else {
          load_next();
        }
      }
    }
  }

  // ---- spatial mix  y[irx] = sum_j S[irx][j] z[j]  and direct stores of the thread's R consecutive outputs ----
  IO* yb = reinterpret_cast<IO*>(a.y) + (size_t)b * a.nrx * Tout + m0;
  const bool vec_ok = sizeof(IO) == 8 && ((Tout & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0) &&
                      (m0 + R <= Tout) && !a.accumulate;
  const uint32_t ssa = smem_u32(Ss) + pidx * NA * 8;
  // partial mix over this thread's antennas for receive stream irx
  auto mix = [&](int irx, u64 (&yv)[R]) {
#pragma unroll
    for (int u = 0; u < R; ++u) yv[u] = 0ull;
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      u64 sv;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(sv) : "r"(ssa + (irx * NTX + j) * 8));
      const float2 sj = upk2(sv);
      const u64 sre = pk2(sj.x, sj.x), sim = pk2(-sj.y, sj.y);
#pragma unroll
      for (int u = 0; u < R; ++u) cmac2(yv[u], acc[u][j], sre, sim);
    }
  };
  auto store_row = [&](int irx, const u64 (&yv)[R]) {
    IO* dst = yb + (size_t)irx * Tout;
#ifdef HB_ATTRIBUTION
    if ((a.dbg & 2) && irx + tid + m0 != -12345) return;
#endif
    if (vec_ok) {
      if constexpr (sizeof(IO) == 8) {
#pragma unroll
        for (int u = 0; u < R; u += 2) stg_stream4(reinterpret_cast<float2*>(dst) + u, yv[u], yv[u + 1]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        if (m0 + u < Tout) {
          float2 v = upk2(yv[u]);
          if (a.accumulate) {
            const IO old = dst[u];
            v.x += (float)old.x;
            v.y += (float)old.y;
          }
          stg_stream(dst + u, IoConv<IO>::make(v.x, v.y));
        }
      }
    }
  };
  if constexpr (SPLIT == 1) {
    for (int irx = 0; irx < a.nrx; ++irx) {
      u64 yv[R];
      mix(irx, yv);
      store_row(irx, yv);
    }
  } else {
    // reduce-scatter between the two lanes of a row: lane group 0 finishes streams [0, half), group 1 the rest
    const int half = (a.nrx + 1) >> 1;
    for (int i = 0; i < half; ++i) {
      const int ia = i, ib = i + half;  // ib may be past the end for odd nrx (its partial is zero work, unused)
      u64 ya[R], yb2[R];
      mix(ia, ya);
      if (ib < a.nrx) {
        mix(ib, yb2);
      } else {
#pragma unroll
        for (int u = 0; u < R; ++u) yb2[u] = 0ull;
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const u64 mine = pidx == 0 ? ya[u] : yb2[u];
        const u64 send = pidx == 0 ? yb2[u] : ya[u];
        const u64 recv = __shfl_xor_sync(live_mask, send, RW);
        const float2 m2 = upk2(mine), r2 = upk2(recv);
        ya[u] = pk2(m2.x + r2.x, m2.y + r2.y);
      }
      const int irx = pidx == 0 ? ia : ib;
      if (irx < a.nrx) store_row(irx, ya);
    }
  }
}

// SPLIT = 2 measured slower on C2 (0.85 ms against 0.74 ms: the duplicated tap-gain evaluation costs more than the
// extra warps buy -- profiles/r01_window_kernel.md), so every chunk width runs one thread per row.
template <int NTX> constexpr int window_split() { return 1; }
template <int NTX> constexpr int window_samples_per_thread() { return NTX <= 4 ? 8 : 4; }

template <int NTX>
int launch_tdl_window(int P, bool io128, bool large_halo, bool lin, const FadingArgs& a, const WindowPlan& wp, int threads,
                      size_t smem, cudaStream_t st);

}  // namespace hb
