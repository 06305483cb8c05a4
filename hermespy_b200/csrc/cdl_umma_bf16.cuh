// K6 on the tensor cores, BF16x3 form (tcgen05.mma kind::f16 with BF16 operands, FP32 accumulators in TMEM), sm_100a only.
//
// Same GEMM as cdl_umma.cuh -- delay groups folded into K through shifted descriptor start addresses, time on M -- with the
// FP32 operands split into THREE bfloat16 terms instead of two TF32 terms:  x = a0 + a1 + a2,  M = c0 + c1 + c2 (8 + 8 + 8
// mantissa bits, FP32 exponent range, so no scaling of the Taylor moments is needed).  Products kept (error ~2^-23):
//   a0 c0 | a0 c1 + a1 c0 | a0 c2 + a1 c1 + a2 c0          dropped: a1 c2, a2 c1, a2 c2 (<= 2^-24)
// as three MMAs per (delay group, 8 antennas, 128 samples) with the B operand stacked along N, [c0 ; c1 ; c2] rows:
//   D[:, 0 : 3 N1P]     += A0 [c0 ; c1 ; c2]^T    (N = 3 N1P)
//   D[:, N1P : 3 N1P]   += A1 [c0 ; c1]^T          (N = 2 N1P)
//   D[:, 2 N1P : 3 N1P] += A2 [c0]^T               (N = N1P)
// so the three column blocks hold terms of magnitude 1, 2^-8, 2^-16 and the epilogue adds them.  Why: small-N MMAs are bound
// by the shared-memory operand reads (tools/microbench/umma_bf16_probe.cu: 144 clk for this triple = 18 KB / 128 B per clock,
// against 2 x 88 clk for the two TF32 pairs that cover the same 8 antennas): -18 % of the binding traffic.  K = 16 per
// instruction means 8 antennas per K stage, so a stage's operand images are twice as large: tile = 256 outputs (two M-tiles,
// two issuing warps -- the probe reaches the operand-read bound with two), 2-slot ring.
// One accumulation chain: 88 MMAs of the leading term for C3, the same length the TF32 kernel reaches with its two chains.
//
// Eligible for 2 P NRX <= 32 columns (4 receive antennas up to P = 4); everything else stays on cdl_umma_kernel.
#pragma once
#include <cuda_bf16.h>

#include "cdl_umma.cuh"

namespace hb {

constexpr int kCbTile = 256;     // outputs per item: two M-tiles
constexpr int kCbMaxAIt = 3;     // register-prefetched (chunk, row) tasks per producer thread: 2 (256 + Dpad) <= 960
constexpr int kCbBBatch = 3;     // B-operand tasks whose loads are in flight together
constexpr int kCbProducers = 320;  // warps 6..15: the BF16x3 operand images cost ~3x the conversions of the TF32 ones

inline bool cb_eligible(int nrx_tpl, int P) { return cu_n1p(nrx_tpl, P) <= 32; }
inline size_t cb_smem_bytes(int nrx_tpl, int P, int Dpad, int G) {
  const size_t W = (size_t)kCbTile + Dpad;
  return 2 * (96 * W + 96 * (size_t)G * cu_n1p(nrx_tpl, P)) + 256;
}

namespace umma {
__device__ __forceinline__ uint32_t instr_desc_bf16(int M, int N) {  // D = f32, A = B = bf16, both K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// x = t0 + t1 + t2 with bfloat16 terms taken by TRUNCATION (t = upper 16 bits of the running remainder): every remainder is
// exact in FP32, so the three terms carry the leading 24 bits of x exactly -- and a term costs one AND and one subtraction.
// Returns the three terms of (lo, hi) packed lo | hi << 16 (PRMT picks the upper halves).
__device__ __forceinline__ void split_bf16x3(float lo, float hi, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  const uint32_t l0 = __float_as_uint(lo) & 0xffff0000u, h0 = __float_as_uint(hi) & 0xffff0000u;
  const float lr = lo - __uint_as_float(l0), hr = hi - __uint_as_float(h0);
  const uint32_t l1 = __float_as_uint(lr) & 0xffff0000u, h1 = __float_as_uint(hr) & 0xffff0000u;
  const uint32_t l2 = __float_as_uint(lr - __uint_as_float(l1)), h2 = __float_as_uint(hr - __uint_as_float(h1));
  p0 = __byte_perm(l0, h0, 0x7632);
  p1 = __byte_perm(l1, h1, 0x7632);
  p2 = __byte_perm(l2, h2, 0x7632);
}
}  // namespace umma

template <int NRX, int P, typename IO>
__global__ void __launch_bounds__(kCuThreads, 1) cdl_umma_bf16_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  using namespace umma;
  constexpr int N1 = 2 * P * NRX, N1P = (N1 + 15) & ~15, MT = 2, COLS = 3 * N1P;
  static_assert(N1P <= 32, "BF16x3 kernel: at most 32 accumulator columns per term");
  extern __shared__ unsigned char cb_smem_raw[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_acc_full[MT], bar_acc_empty[MT];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = tb.num_groups;
  const int W = kCbTile + a.Dpad;                      // staged rows (Dpad % 8 == 0)
  const uint32_t a_plane = (uint32_t)W * 16u;          // one 16-byte K chunk (4 antennas, re / im, bf16) of all rows
  const uint32_t a_split = 2u * a_plane;               // the two chunks of one split term
  const uint32_t a_bytes = 3u * a_split;               // a0 | a1 | a2
  const uint32_t b_chunk = (uint32_t)COLS * 16u;       // rows c0[N1P] | c1[N1P] | c2[N1P] of one K chunk
  const uint32_t b_group = 2u * b_chunk;
  const uint32_t stage_bytes = a_bytes + (uint32_t)G * b_group;
  const uint32_t smem0 = (smem_addr(cb_smem_raw) + 127u) & ~127u;
  const int NS = (a.ntx + 7) >> 3;                     // K stages of 8 transmit antennas
  const int nitems = a.B * a.ntiles;
  const int Tout = a.T + a.D;
  const int nij = a.nrx * a.ntx;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], kCbProducers);
      mbar_init(&bar_empty[s], MT);
    }
    for (int mt = 0; mt < MT; ++mt) {
      mbar_init(&bar_acc_full[mt], 1);
      mbar_init(&bar_acc_empty[mt], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows the producers never write (N padding) must hold finite values
  for (uint32_t o = (uint32_t)tid * 16u; o < 2u * stage_bytes; o += kCuThreads * 16u) sts128(smem0 + o, 0.f, 0.f, 0.f, 0.f);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp >= 6) {
    // ================================ producers ====================================================================
    const int ptid = tid - (kCuThreads - kCbProducers);
    float2 xa[kCbMaxAIt][4];  // task e = c W' + r: the four antennas of chunk c at row r
    const int btasks = 2 * G * N1;  // (group, chunk, row) entries of the B operand per stage
    auto rows_needed = [&](int q) { return ((min(kCbTile, Tout - q * kCbTile) + 127) & ~127) + a.Dpad; };
    auto load_a = [&](const IO* xb, int n0, int j0, int e, int wn, float2 (&v)[4]) {
      const int c = e >= wn ? 1 : 0, r = e - c * wn;
      const int n = n0 + r;
      const bool ok = e < 2 * wn && n >= 0 && n < a.T;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = j0 + 4 * c + i;
        v[i] = (ok && j < a.ntx) ? to_c32(ldg_stream(xb + (size_t)j * a.T + n)) : make_float2(0.f, 0.f);
      }
    };
    auto store_a = [&](uint32_t sA, int e, int wn, const float2 (&v)[4]) {
      const int c = e >= wn ? 1 : 0, r = e - c * wn;
      uint32_t p0[4], p1[4], p2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16x3(v[i].x, v[i].y, p0[i], p1[i], p2[i]);
      const uint32_t o = sA + (uint32_t)c * a_plane + (uint32_t)r * 16u;
      sts128u(o, p0[0], p0[1], p0[2], p0[3]);
      sts128u(o + a_split, p1[0], p1[1], p1[2], p1[3]);
      sts128u(o + 2u * a_split, p2[0], p2[1], p2[2], p2[3]);
    };
    // B task e = (g, c, row): operand row (p, i, re / im) of the 4 antennas of chunk c of delay group g; the lanes of a warp write
    // consecutive 16-byte rows (conflict-free).  Loads of kCbBBatch tasks are issued before the first conversion (the
    // moments come from L2; the volatile stores would otherwise serialize one round trip per task).
    auto load_b = [&](const float2* mq, int j0, int e, float2 (&m)[4]) {
      const int row = e % N1, c = (e / N1) & 1, g = e / (2 * N1);
      const int pi = row >> 1, p = pi / NRX, i = pi - p * NRX, j = j0 + 4 * c;
      const bool ok = e < btasks && i < a.nrx_chunk;
      const float2* src = mq + (size_t)(g * P + p) * nij + (size_t)(a.rx0 + i) * a.ntx + j;
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = (ok && j + k < a.ntx) ? src[k] : make_float2(0.f, 0.f);
    };
    auto store_b = [&](uint32_t sB, int e, const float2 (&m)[4]) {
      const int row = e % N1, c = (e / N1) & 1, g = e / (2 * N1);
      const bool im = row & 1;
      uint32_t p0[4], p1[4], p2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)  // Re row: [Mr, -Mi] (Mr x_r - Mi x_i);  Im row: [Mi, Mr] (Mi x_r + Mr x_i)
        split_bf16x3(im ? m[k].y : m[k].x, im ? m[k].x : -m[k].y, p0[k], p1[k], p2[k]);
      const uint32_t o = sB + (uint32_t)g * b_group + (uint32_t)c * b_chunk + (uint32_t)row * 16u;
      sts128u(o, p0[0], p0[1], p0[2], p0[3]);
      sts128u(o + (uint32_t)N1P * 16u, p1[0], p1[1], p1[2], p1[3]);
      sts128u(o + 2u * (uint32_t)N1P * 16u, p2[0], p2[1], p2[2], p2[3]);
    };
    auto prefetch = [&](int item, int s) {
      const int b = item / a.ntiles, q = item - b * a.ntiles;
      const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
      const int wn = rows_needed(q);
#pragma unroll
      for (int it = 0; it < kCbMaxAIt; ++it) load_a(xb, q * kCbTile - a.Dpad, 8 * s, ptid + kCbProducers * it, wn, xa[it]);
    };
    uint32_t sc = 0;
    int item = blockIdx.x;
    if (item < nitems) prefetch(item, 0);
    for (; item < nitems; item += gridDim.x) {
      const int b = item / a.ntiles, q = item - b * a.ntiles;
      const int wn = rows_needed(q);
      const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
      const float2* mq = a.moments + (((size_t)b * a.nwin + (q * kCbTile) / a.ptile) * G) * P * nij;
      for (int s = 0; s < NS; ++s, ++sc) {
        const int slot = sc & 1;
        if (sc >= 2) mbar_wait(&bar_empty[slot], ((sc >> 1) - 1u) & 1u);  // MMAs that read this slot have completed
        const uint32_t sA = smem0 + (uint32_t)slot * stage_bytes, sB = sA + a_bytes;
#pragma unroll
        for (int it = 0; it < kCbMaxAIt; ++it) {
          const int e = ptid + kCbProducers * it;
          if (e < 2 * wn) store_a(sA, e, wn, xa[it]);
        }
        for (int e = ptid + kCbProducers * kCbMaxAIt; e < 2 * wn; e += kCbProducers) {  // long delay spreads: unhidden rest
          float2 v[4];
          load_a(xb, q * kCbTile - a.Dpad, 8 * s, e, wn, v);
          store_a(sA, e, wn, v);
        }
        for (int e0 = ptid; e0 < btasks; e0 += kCbProducers * kCbBBatch) {
          float2 m[kCbBBatch][4];
#pragma unroll
          for (int it = 0; it < kCbBBatch; ++it) load_b(mq, 8 * s, e0 + kCbProducers * it, m[it]);
#pragma unroll
          for (int it = 0; it < kCbBBatch; ++it)
            if (e0 + kCbProducers * it < btasks) store_b(sB, e0 + kCbProducers * it, m[it]);
        }
        fence_async_smem();
        mbar_arrive(&bar_full[slot]);
        if (s + 1 < NS) prefetch(item, s + 1);
        else if (item + (int)gridDim.x < nitems) prefetch(item + gridDim.x, 0);
      }
    }
  } else if (warp >= 4) {
    // ================================ MMA issue: warp 4 + mt owns M-tile mt ==========================================
    const int mt = warp - 4;
    if (mt < MT) {
      const uint32_t id3 = instr_desc_bf16(128, 3 * N1P), id2 = instr_desc_bf16(128, 2 * N1P), id1 = instr_desc_bf16(128, N1P);
      uint32_t sc = 0, ic = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ic) {
        const int q = item % a.ntiles;
        const bool active = mt * 128 < min(kCbTile, Tout - q * kCbTile);
        int gd0 = 0, gd1 = 0;  // heterogeneous batch: this link's group delays, lane l holds groups l and l + 32
        if (a.link_tab) {
          const int32_t* gdl = a.link_tab[item / a.ntiles].group_delay;
          gd0 = lane < G ? gdl[lane] : 0;
          gd1 = lane + 32 < G ? gdl[lane + 32] : 0;
        }
        if (ic >= 1) {  // the epilogue has drained this M-tile's accumulator of the previous item
          mbar_wait(&bar_acc_empty[mt], (ic - 1u) & 1u);
          fence_after_sync();
        }
        const uint32_t d = tmem + (uint32_t)mt * COLS;
        for (int s = 0; s < NS; ++s, ++sc) {
          const int slot = sc & 1;
          mbar_wait(&bar_full[slot], (sc >> 1) & 1u);
          fence_after_sync();
          if (active) {
            const uint32_t sA = smem0 + (uint32_t)slot * stage_bytes + (uint32_t)(mt * 128 + a.Dpad) * 16u;
            const uint64_t da0 = smem_desc(sA, a_plane, 128u), da1 = smem_desc(sA + a_split, a_plane, 128u),
                           da2 = smem_desc(sA + 2u * a_split, a_plane, 128u);
            uint64_t db = smem_desc(smem0 + (uint32_t)slot * stage_bytes + a_bytes, b_chunk, 128u);
            if (a.link_tab == nullptr) {  // the uniform loop stays free of the per-link select (it is the critical path)
              for (int g = 0; g < G; ++g) {
                const uint64_t k = (uint64_t)(uint32_t)tb.group_delay[g];  // start-address field counts 16-byte rows
                mma_bf16_elect(d, da0 - k, db, id3, (uint32_t)(s | g));
                mma_bf16_elect(d + N1P, da1 - k, db, id2, 1u);
                mma_bf16_elect(d + 2 * N1P, da2 - k, db, id1, 1u);
                db += (uint64_t)(b_group >> 4);
              }
            } else {
              for (int g = 0; g < G; ++g) {
                const uint64_t k = (uint64_t)(uint32_t)__shfl_sync(0xffffffffu, g < 32 ? gd0 : gd1, g & 31);
                mma_bf16_elect(d, da0 - k, db, id3, (uint32_t)(s | g));
                mma_bf16_elect(d + N1P, da1 - k, db, id2, 1u);
                mma_bf16_elect(d + 2 * N1P, da2 - k, db, id1, 1u);
                db += (uint64_t)(b_group >> 4);
              }
            }
          }
          commit_elect(&bar_empty[slot]);
        }
        commit_elect(&bar_acc_full[mt]);
      }
    }
  } else {
    // ================================ epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 ============================
    const float inv_win = 1.0f / (float)a.ptile, half = 0.5f * (float)a.ptile;
    uint32_t ic = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ic) {
      const int b = item / a.ntiles, q = item - b * a.ntiles;
      const int valid = min(kCbTile, Tout - q * kCbTile);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        mbar_wait(&bar_acc_full[mt], ic & 1u);
        fence_after_sync();
        float v[N1P];
        if (mt * 128 < valid) {  // warp-uniform
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)mt * COLS;
#pragma unroll
          for (int c0 = 0; c0 < N1P; c0 += 16) {
            uint32_t t0[16], t1[16], t2[16];
            tmem_ld16(taddr + c0, t0);
            tmem_ld16(taddr + N1P + c0, t1);
            tmem_ld16(taddr + 2 * N1P + c0, t2);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k)
              v[c0 + k] = __uint_as_float(t0[k]) + (__uint_as_float(t1[k]) + __uint_as_float(t2[k]));
          }
        }
        fence_before_sync();
        mbar_arrive(&bar_acc_empty[mt]);
        const int il = mt * 128 + warp * 32 + lane;
        if (mt * 128 < valid && il < valid) {
          const int m = q * kCbTile + il;
          const float r = ((float)(m % a.ptile) - half) * inv_win;  // coordinate inside the Taylor window
#pragma unroll
          for (int i = 0; i < NRX; ++i) {
            if (i < a.nrx_chunk) {
              float re = v[((P - 1) * NRX + i) * 2], im = v[((P - 1) * NRX + i) * 2 + 1];
#pragma unroll
              for (int p = P - 2; p >= 0; --p) {
                re = fmaf(re, r, v[(p * NRX + i) * 2]);
                im = fmaf(im, r, v[(p * NRX + i) * 2 + 1]);
              }
              IO* dst = reinterpret_cast<IO*>(a.y) + ((size_t)b * a.nrx + a.rx0 + i) * Tout + m;
              stg_stream(dst, IoConv<IO>::make(re, im));
            }
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// explicit instantiations: cdl_umma_bf16.cu
int launch_cdl_umma_bf16(int nrx_tpl, int P, bool io128, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st);

}  // namespace hb
