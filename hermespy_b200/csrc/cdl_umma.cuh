// K6 on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), sm_100a only:
//   y[i, m] = sum_p r(m)^p  sum_g sum_j  M[g, p, i, j] x[j, m - k_g]          (cluster_delay_lines.py:526-558)
// as ONE real-valued GEMM per tile of 128 MT output samples (the Taylor window of the moments is one or several tiles)
// with the delay groups folded into K:
//   A (M = 128 time rows, K = (g, j, re/im)):  A[m, (g, j, c)] = x_c[j, m - k_g]
//   B (N = (p, i, re/im) rows, same K):        Re row = [Mr, -Mi],  Im row = [Mi, Mr]
//   D[m, (p, i, c')] in TMEM; epilogue = Horner over p in the window coordinate r, then one complex store per thread.
//
// The delayed copies of x cost nothing.  Operands use the no-swizzle K-major canonical layout with SBO = 128 bytes,
// in which the rows of one 16-byte K chunk are linear in shared memory (row r at 16 r): the staged x tile
// [2 chunks][tile + Dpad rows][16 B = two antennas' complex samples] is addressed for delay group g by advancing
// the descriptor START ADDRESS by (Dpad - k_g) rows -- validated bit-exactly for arbitrary row shifts by
// tools/microbench/umma_shift_probe.cu.
//
// 3xTF32 with two MMAs per (delay group, 4 antennas, 128 samples): x = hi + lo, M = hi + lo,
//   D[:, 0:N1P | N1P:2 N1P] += A_hi [B_hi ; B_lo]^T   (N = 2 N1P)       D[:, N1P:2 N1P] += A_lo B_hi^T   (N = N1P)
// and the epilogue adds the column halves (lo.lo is below 2^-22).  The tensor core adds into its FP32 accumulator with
// truncation, so the error of a long accumulation chain grows linearly with its length (measured: 6.9e-6 relative for the
// 352 additions of config C3 in one accumulator).  Therefore (i) the small correction products (hi.lo, lo.hi) go to their
// own columns [N1P, 2 N1P) and never lengthen the chain of the hi.hi products, and (ii) even and odd K stages accumulate
// in two independent column sets that the epilogue adds with round-to-nearest: chain length 88 for C3.
//
// Measured (umma_shift_probe, B200): such small-N MMAs are bound by the shared-memory operand reads (128 B/clk:
// 48 + 40 clk per pair at N1P = 32) once FOUR warps issue them in parallel; one issuing thread alone needs ~94 clk
// per MMA.  Hence the roles of the 16 warps of the persistent CTA (one per SM):
//   warps 0-3   epilogue (TMEM lane quadrants); accumulators are handed over per M-tile, so only the issuing warp of the
//               M-tile being drained waits while the other three keep the tensor pipe busy
//   warps 4-7   MMA issue, warp 4 + mt owns M-tile mt of the window (its own accumulator columns)
//   warps 8-15  producers: x tile (+ delay halo) and the moment block of 4 transmit antennas -> hi / lo operand
//               images in a 2-slot shared-memory ring; the next slot's global loads are in registers while the
//               previous one is consumed.  K is walked antenna-chunk by antenna-chunk, so every operand byte is
//               staged once per window and the accumulators stay in TMEM for the whole contraction.
#pragma once
#include "cdl_types.cuh"
#include "spatial_gemm.cuh"

namespace hb {

constexpr int kCuThreads = 512;
constexpr int kCuProducers = 256;
constexpr int kCuMaxAIt = 3;  // register-prefetched rows per producer thread and antenna pair (tile + Dpad <= 768)
constexpr int kCuMaxBIt = 6;  // register-prefetched (group, antenna pair, row) entries per producer thread (4 G P NRX <= 1536)

template <int NRX, int P>
struct CuShape {
  static constexpr int N1 = 2 * P * NRX;           // real output columns (p, i, re/im)
  static constexpr int N1P = (N1 + 15) & ~15;      // padded to the N granularity of M = 128
  static constexpr int MT = (128 / N1P) >= 4 ? 4 : (128 / N1P);  // M-tiles per window: 2 chains x MT x 2 N1P <= 512 columns
  static constexpr int TILE = 128 * MT;            // output samples per Taylor window
  static constexpr int COLS = 2 * N1P;             // TMEM columns per M-tile
  static_assert(N1P <= 128 && MT >= 1, "receive chunk x Taylor order too large for one accumulator");
};

// host-side mirror of CuShape for the planner
inline int cu_n1p(int nrx_tpl, int P) { return (2 * P * nrx_tpl + 15) & ~15; }
inline int cu_tile(int nrx_tpl, int P) {
  const int mt = 128 / cu_n1p(nrx_tpl, P);
  return 128 * (mt >= 4 ? 4 : mt);
}
inline size_t cu_smem_bytes(int nrx_tpl, int P, int Dpad, int G) {
  const size_t W = (size_t)cu_tile(nrx_tpl, P) + Dpad;
  return 2 * (64 * W + 64 * (size_t)G * cu_n1p(nrx_tpl, P)) + 256;
}

namespace umma {
// 32 TMEM lanes x 16 consecutive columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
}  // namespace umma

template <int NRX, int P, typename IO>
__global__ void __launch_bounds__(kCuThreads, 1) cdl_umma_kernel(const CdlArgs a, const __grid_constant__ CdlTable tb) {
  using S = CuShape<NRX, P>;
  using namespace umma;
  extern __shared__ unsigned char cu_smem_raw[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_acc_full[4], bar_acc_empty[4];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = tb.num_groups;
  const int W = S::TILE + a.Dpad;                       // staged rows: window + delay halo (Dpad % 8 == 0)
  const uint32_t a_plane = (uint32_t)W * 16u;           // one 16-byte K chunk (antenna pair) of all rows
  const uint32_t a_bytes = 4u * a_plane;                // hi[2 chunks] | lo[2 chunks]
  const uint32_t b_chunk = (uint32_t)S::COLS * 16u;     // rows hi[N1P] | lo[N1P] of one K chunk
  const uint32_t b_group = 2u * b_chunk;
  const uint32_t stage_bytes = a_bytes + (uint32_t)G * b_group;
  const uint32_t smem0 = (smem_addr(cu_smem_raw) + 127u) & ~127u;
  const int NS = (a.ntx + 3) >> 2;                      // K stages: 4 transmit antennas = one K = 8 step per delay group
  const int nitems = a.B * a.ntiles;
  const int Tout = a.T + a.D;
  const int nij = a.nrx * a.ntx;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_full[s], kCuProducers);
      mbar_init(&bar_empty[s], S::MT);
    }
    for (int mt = 0; mt < 4; ++mt) {
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

  if (warp >= 8) {
    // ================================ producers ====================================================================
    const int ptid = tid - (kCuThreads - kCuProducers);
    float2 xa[2][kCuMaxAIt][2];
    float4 mb[kCuMaxBIt];
    const int btasks = 2 * G * S::N1;  // (group, antenna pair, row) entries of the B operand per stage
    // x[b, ja .. ja + 1, n0 + r] of row r (zero outside the frame / past the last antenna)
    // rows the MMAs of window q read: its M-tiles inside the frame plus the delay halo (the last window of a frame is short)
    auto rows_needed = [&](int q) { return ((min(S::TILE, Tout - q * S::TILE) + 127) & ~127) + a.Dpad; };
    auto load_a = [&](const IO* xb, int n0, int ja, int r, int wn, float2& v0, float2& v1) {
      const int n = n0 + r;
      const bool ok = r < wn && n >= 0 && n < a.T;
      v0 = make_float2(0.f, 0.f);
      v1 = v0;
      if (ok && ja < a.ntx) v0 = to_c32(ldg_stream(xb + (size_t)ja * a.T + n));
      if (ok && ja + 1 < a.ntx) v1 = to_c32(ldg_stream(xb + (size_t)(ja + 1) * a.T + n));
    };
    auto store_a = [&](uint32_t sA, int c, int r, float2 v0, float2 v1) {
      float h0, h1, h2, h3, l0, l1, l2, l3;
      split_tf32(v0.x, h0, l0);
      split_tf32(v0.y, h1, l1);
      split_tf32(v1.x, h2, l2);
      split_tf32(v1.y, h3, l3);
      const uint32_t o = sA + (uint32_t)c * a_plane + (uint32_t)r * 16u;
      sts128(o, h0, h1, h2, h3);
      sts128(o + 2u * a_plane, l0, l1, l2, l3);
    };
    // B task e = (g, c, row): operand row (p, i, re/im) of antenna pair c of delay group g; the 32 lanes of a warp write 32
    // consecutive 16-byte rows (conflict-free); the two rows of one (p, i) read the same pair of moments
    auto load_b = [&](const float2* mq, int j0, int e) {
      const int row = e % S::N1, c = (e / S::N1) & 1, g = e / (2 * S::N1);
      const int pi = row >> 1, p = pi / NRX, i = pi - p * NRX, j = j0 + 2 * c;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < G && i < a.nrx_chunk) {
        const float2* src = mq + (size_t)(g * P + p) * nij + (size_t)(a.rx0 + i) * a.ntx + j;
        if (j < a.ntx) {
          const float2 m0 = src[0];
          v.x = m0.x;
          v.y = m0.y;
        }
        if (j + 1 < a.ntx) {
          const float2 m1 = src[1];
          v.z = m1.x;
          v.w = m1.y;
        }
      }
      return v;
    };
    auto store_b = [&](uint32_t sB, int e, float4 m) {
      const int row = e % S::N1, c = (e / S::N1) & 1, g = e / (2 * S::N1);
      // Re row: [Mr, -Mi] (Mr x_r - Mi x_i);  Im row: [Mi, Mr] (Mi x_r + Mr x_i)
      const bool im = row & 1;
      const float k0 = im ? m.y : m.x, k1 = im ? m.x : -m.y, k2 = im ? m.w : m.z, k3 = im ? m.z : -m.w;
      float h0, h1, h2, h3, l0, l1, l2, l3;
      split_tf32(k0, h0, l0);
      split_tf32(k1, h1, l1);
      split_tf32(k2, h2, l2);
      split_tf32(k3, h3, l3);
      const uint32_t o = sB + (uint32_t)g * b_group + (uint32_t)c * b_chunk + (uint32_t)row * 16u;
      sts128(o, h0, h1, h2, h3);
      sts128(o + (uint32_t)S::N1P * 16u, l0, l1, l2, l3);
    };
    auto prefetch = [&](int item, int s) {
      const int b = item / a.ntiles, q = item - b * a.ntiles;
      const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
      const int wn = rows_needed(q);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int it = 0; it < kCuMaxAIt; ++it)
          load_a(xb, q * S::TILE - a.Dpad, 4 * s + 2 * c, ptid + kCuProducers * it, wn, xa[c][it][0], xa[c][it][1]);
      const float2* mq = a.moments + (((size_t)b * a.nwin + (q * S::TILE) / a.ptile) * G) * P * nij;
#pragma unroll
      for (int it = 0; it < kCuMaxBIt; ++it) mb[it] = load_b(mq, 4 * s, ptid + kCuProducers * it);
    };
    auto store = [&](int slot, int item, int s) {
      const uint32_t sA = smem0 + (uint32_t)slot * stage_bytes;
      const uint32_t sB = sA + a_bytes;
      const int wn = rows_needed(item % a.ntiles);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int it = 0; it < kCuMaxAIt; ++it) {
          const int r = ptid + kCuProducers * it;
          if (r < wn) store_a(sA, c, r, xa[c][it][0], xa[c][it][1]);
        }
#pragma unroll
      for (int it = 0; it < kCuMaxBIt; ++it) {
        const int e = ptid + kCuProducers * it;
        if (e < btasks) store_b(sB, e, mb[it]);
      }
      // long delay spreads / many (group, order, antenna) blocks: the part beyond the register prefetch is loaded here
      if (wn > kCuProducers * kCuMaxAIt || btasks > kCuProducers * kCuMaxBIt) {
        const int b = item / a.ntiles, q = item - b * a.ntiles;
        const IO* xb = reinterpret_cast<const IO*>(a.x) + (size_t)b * a.ntx * a.T;
        for (int c = 0; c < 2; ++c)
          for (int r = ptid + kCuProducers * kCuMaxAIt; r < wn; r += kCuProducers) {
            float2 v0, v1;
            load_a(xb, q * S::TILE - a.Dpad, 4 * s + 2 * c, r, wn, v0, v1);
            store_a(sA, c, r, v0, v1);
          }
        const float2* mq = a.moments + (((size_t)b * a.nwin + (q * S::TILE) / a.ptile) * G) * P * nij;
        for (int e = ptid + kCuProducers * kCuMaxBIt; e < btasks; e += kCuProducers) store_b(sB, e, load_b(mq, 4 * s, e));
      }
    };
    uint32_t sc = 0;
    int item = blockIdx.x;
    if (item < nitems) prefetch(item, 0);
    for (; item < nitems; item += gridDim.x) {
      for (int s = 0; s < NS; ++s, ++sc) {
        const int slot = sc & 1;
        if (sc >= 2) mbar_wait(&bar_empty[slot], ((sc >> 1) - 1u) & 1u);  // MMAs that read this slot have completed
        store(slot, item, s);
        fence_async_smem();
        mbar_arrive(&bar_full[slot]);
        if (s + 1 < NS) prefetch(item, s + 1);
        else if (item + (int)gridDim.x < nitems) prefetch(item + gridDim.x, 0);
      }
    }
  } else if (warp >= 4) {
    // ================================ MMA issue: warp 4 + mt owns M-tile mt ==========================================
    const int mt = warp - 4;
    if (mt < S::MT) {
      const uint32_t idesc_full = instr_desc_tf32(128, S::COLS), idesc_hi = instr_desc_tf32(128, S::N1P);
      uint32_t sc = 0, ic = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ic) {
        const int q = item % a.ntiles;
        const bool active = mt * 128 < min(S::TILE, Tout - q * S::TILE);
        int gd0 = 0, gd1 = 0;  // heterogeneous batch: this link's group delays, lane l holds groups l and l + 32
        if (a.link_tab) {
          const int32_t* gdl = a.link_tab[item / a.ntiles].group_delay;
          gd0 = lane < G ? gdl[lane] : 0;
          gd1 = lane + 32 < G ? gdl[lane + 32] : 0;
        }
        if (ic >= 1) {  // the epilogue has drained this M-tile's accumulators of the previous window
          mbar_wait(&bar_acc_empty[mt], (ic - 1u) & 1u);
          fence_after_sync();
        }
        for (int s = 0; s < NS; ++s, ++sc) {
          const int slot = sc & 1;
          mbar_wait(&bar_full[slot], (sc >> 1) & 1u);
          fence_after_sync();
          if (active) {
            const uint32_t d = tmem + (uint32_t)(s & 1) * 256u + (uint32_t)mt * S::COLS;  // accumulation chain s & 1
            const uint32_t sA = smem0 + (uint32_t)slot * stage_bytes + (uint32_t)(mt * 128 + a.Dpad) * 16u;
            const uint64_t da_hi = smem_desc(sA, a_plane, 128u), da_lo = smem_desc(sA + 2u * a_plane, a_plane, 128u);
            uint64_t db = smem_desc(smem0 + (uint32_t)slot * stage_bytes + a_bytes, b_chunk, 128u);
            if (a.link_tab == nullptr) {  // the uniform loop stays free of the per-link select (it is the critical path)
              for (int g = 0; g < G; ++g) {
                const uint64_t k = (uint64_t)(uint32_t)tb.group_delay[g];  // start-address field counts 16-byte rows
                mma_tf32_elect(d, da_hi - k, db, idesc_full, (uint32_t)((s >> 1) | g));
                mma_tf32_elect(d + S::N1P, da_lo - k, db, idesc_hi, 1u);
                db += (uint64_t)(b_group >> 4);
              }
            } else {
              for (int g = 0; g < G; ++g) {
                const uint64_t k = (uint64_t)(uint32_t)__shfl_sync(0xffffffffu, g < 32 ? gd0 : gd1, g & 31);
                mma_tf32_elect(d, da_hi - k, db, idesc_full, (uint32_t)((s >> 1) | g));
                mma_tf32_elect(d + S::N1P, da_lo - k, db, idesc_hi, 1u);
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
      const int valid = min(S::TILE, Tout - q * S::TILE);
#pragma unroll 1
      for (int mt = 0; mt < S::MT; ++mt) {
        mbar_wait(&bar_acc_full[mt], ic & 1u);
        fence_after_sync();
        if (mt * 128 < valid) {  // warp-uniform
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)mt * S::COLS;
          float v[S::N1P];
#pragma unroll
          for (int c0 = 0; c0 < S::N1P; c0 += 16) {
            uint32_t hi[16], lo[16];
            tmem_ld16(taddr + c0, hi);
            tmem_ld16(taddr + S::N1P + c0, lo);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) v[c0 + k] = __uint_as_float(hi[k]) + __uint_as_float(lo[k]);
            if (NS > 1) {  // the chain of the odd K stages
              tmem_ld16(taddr + 256u + c0, hi);
              tmem_ld16(taddr + 256u + S::N1P + c0, lo);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; ++k) v[c0 + k] += __uint_as_float(hi[k]) + __uint_as_float(lo[k]);
            }
          }
          const int il = mt * 128 + warp * 32 + lane;
          if (il < valid) {
            const int m = q * S::TILE + il;
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
        fence_before_sync();
        mbar_arrive(&bar_acc_empty[mt]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// explicit instantiations: cdl_umma_c64.cu (IO = float2), cdl_umma_c128.cu (IO = double2)
template <typename IO>
int launch_cdl_umma_io(int nrx_tpl, int P, const CdlArgs& a, const CdlTable& tb, size_t smem, cudaStream_t st);

}  // namespace hb
