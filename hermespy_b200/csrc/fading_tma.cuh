// K3 + K4, persistent TMA-pipelined sliding-window form (sm_100a).
//
//   y[:, m] = S @ sum_g h_g(m) x[:, m - d_g]        (gather form of fading.py:385-393, see fading_kernels.cuh)
//
// Same arithmetic as tdl_window_kernel (fading_window.cuh): every thread owns R = 8 consecutive outputs and walks
// the delay axis once with the R inputs of every antenna in a register window.  What changes is how the signal
// reaches shared memory and how the CTA is scheduled:
//
// * The x tile (+ delay halo) of every antenna of the chunk arrives by `cp.async.bulk.tensor.4d` (TMA),
//   in its natural time-major layout, written by the copy engine -- no LSU instructions, no shared-memory
//   wavefronts spent on staging (the cp.async staging of the window kernel costs 2.1 of its 4.4 wavefronts per
//   sample, profiles/r01_window_kernel.md).  The frame is described to the TMA unit as a 4-D tensor
//   (32 floats = 16 samples, T / 16 rows, Ntx, B): a tile is one box {32, 64 + H, 1, 1} per antenna at row 64 q - H,
//   and the copy engine zero-fills rows before the frame start / past its end and antennas past the chunk.
// * SWIZZLE_128B: the 16-byte chunk index of every 128-byte row is XORed with (row & 7).  A thread reads TIME
//   PAIRS (x[e], x[e+1]) of one antenna with LDS.128; the 8 lanes of a quarter warp are 64 bytes apart in the
//   natural layout (4-way bank conflict) and land on 8 distinct chunk positions under the swizzle: conflict-free.
// * Persistent CTAs (3 per SM) with a two-slot ring: the next unclaimed tile (global atomic counter, so that CTAs
//   sharing an SM unevenly still finish together) is requested right after the walk of tile i released its slot, so every walk starts on data that is already resident, and the streaming stores of
//   the epilogue overlap the loads of the following tiles.  Taylor coefficients and the FP32 spatial matrix ride
//   the same mbarrier as 1-D bulk copies (three aux slots: the spatial matrix is still read by the epilogue of
//   tile i when the request for tile i+2 goes out).
//
// Pair walk.  Outputs m0 .. m0+7 (m0 = 8 row), window slot of x[k] = k mod 8.  Stepping to delay d needs x[m0 - d].
// For odd d the aligned pair (x[m0-d-1], x[m0-d]) holds the elements entering at d and d+1; its slots are free
// once output 7 of delay d has been accumulated (x[m0-d+7] shares the slot of x[m0-d-1]), and the upper element is
// first read by output 0 of delay d.  Order at an odd delay: MAC(7), load pair, MAC(1..6), MAC(0).
#pragma once
#include <cuda.h>

#include "fading_window.cuh"

namespace hb {

constexpr int kTmaThreads = 384;   // ONE persistent CTA per SM: 12 independent warps
constexpr int kTmaR = 8;
constexpr int kTmaTile = 1024;     // outputs per tile = 4 warp parts of 32 rows x 8 outputs
constexpr int kTmaParts = 4;
constexpr int kTmaMaxSlots = 6;    // ring depth (tiles resident or in flight per SM)
constexpr int kTmaMaxHaloRows = 8;             // 16-sample rows of delay halo: d <= 127
constexpr int kTmaMaxBlocks = 2 * kTmaMaxHaloRows;
constexpr uint32_t kTmaPlaneBytes = (kTmaTile / 16 + kTmaMaxHaloRows) * 128;  // 9216: antenna plane stride, a multiple of 1024

struct TmaPlan {
  int32_t num_groups;
  int32_t nblk;        // blocks of 8 delays
  int32_t hrows;       // halo rows (16 samples each) = (nblk + 1) / 2
  int32_t rows;        // 64 + hrows
  int32_t poly_tile, npoly;
  int32_t ntiles;      // tiles per link
  int32_t total_tiles; // B * ntiles
  int32_t coef_stride; // float2 elements per (link, Taylor window) block, even
  int32_t s_stride;    // float2 elements per (link, antenna chunk) block of the FP32 spatial matrix, even
  int32_t chunk;       // antenna chunk index (tx0 / NTX)
  int32_t nchunks;
  unsigned int* tile_counter;  // zeroed by K1; tiles past the first two of every CTA are claimed from it
  uint32_t stage_bytes;  // NTX * rows * 128 rounded up to 1024
  uint32_t aux_bytes;    // coefficients (+ one group of padding) + spatial matrix (+ one row), multiple of 16
  uint32_t slot_bytes;   // one ring slot: x stage + aux, multiple of 1024
  int32_t num_slots;     // ring depth, <= kTmaMaxSlots
  uint32_t coef_bytes, s_bytes;  // bulk copy sizes (multiples of 16)
  uint32_t s_off;                // offset of the spatial matrix inside an aux slot
  // per block c: bits 0..7 "a tap has delay d = 8 c + s"; bits 8..15, odd s only: "the pair entering at d is
  // needed by a present delay in [d, d + 8]"
  uint32_t mask[kTmaMaxBlocks + 1];
};

namespace tma {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// SWIZZLE_128B of an absolute shared address inside a 1024-byte aligned stage
__device__ __forceinline__ uint32_t swz(uint32_t addr) { return addr ^ ((addr >> 3) & 0x70u); }
template <int IMM>
__device__ __forceinline__ void lds_pair_at(uint32_t addr, u64& lo, u64& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2 + %3];" : "=l"(lo), "=l"(hi) : "r"(addr), "n"(IMM));
}

}  // namespace tma

template <int NTX, int P, bool LIN, bool ZMODE>
__global__ void __launch_bounds__(kTmaThreads, 1)
    tdl_tma_kernel(const FadingArgs a, const __grid_constant__ TmaPlan tp, const __grid_constant__ CUtensorMap xmap) {
  constexpr int R = kTmaR;
  constexpr int NH = LIN ? 2 : P;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // Output row (unit of 8 samples) of a lane inside its 32-row part.  The lanes of a quarter warp take every other
  // row of a 16-row span: their 64-byte half rows then share a parity for EVERY block shift of the walk, which is
  // what makes the swizzled LDS.128 conflict-free (8 consecutive half rows starting at an odd one collide).
  const int lrow = ((lane >> 4) << 4) + 2 * (lane & 7) + ((lane >> 3) & 1);
  const int NS = tp.num_slots;
  const uint32_t ring0 = smem_u32(smem_raw);                     // NS slots of (x stage | aux)
  const uint32_t bar0 = ring0 + (uint32_t)NS * tp.slot_bytes;    // NS full barriers
  const uint32_t cnt0 = bar0 + 8u * kTmaMaxSlots;                // NS release counters (parts finished)
  const uint32_t tid0 = cnt0 + 4u * kTmaMaxSlots;                // tile index held by each slot (-1: no more work)
  const uint32_t next0 = tid0 + 4u * kTmaMaxSlots;               // next (tile, part) ticket of this CTA
  const uint32_t done0 = next0 + 16u;                            // NS flags "all parts of the slot's tile are stored"
  const uint32_t rptr0 = done0 + 4u * kTmaMaxSlots;              // ring position that retires next (in order)
  const uint32_t lock0 = rptr0 + 4u;                             // spin lock of the retire loop
  const int Tout = a.T + a.D;
  const int step = (int)gridDim.x;
  const uint32_t plane = (uint32_t)tp.rows * 128u;

  auto request = [&](int t, int slot) {  // one thread: all copies of tile t into ring slot `slot`, one barrier
    const uint32_t bar = bar0 + 8u * slot;
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(tid0 + 4u * slot), "r"(t < tp.total_tiles ? t : -1) : "memory");
    if (t >= tp.total_tiles) {  // no more work: complete the phase so that the waiting warps see the end marker
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      return;
    }
    int b = t / tp.ntiles, chunk = tp.chunk;
    const int q = t - b * tp.ntiles;
    if constexpr (ZMODE) {  // one launch covers every antenna chunk: t = (link, chunk, tile)
      chunk = b % tp.nchunks;
      b /= tp.nchunks;
    }
    const uint32_t xs = ring0 + (uint32_t)slot * tp.slot_bytes;
    tma::mbar_expect_tx(bar, (uint32_t)NTX * plane + tp.coef_bytes + tp.s_bytes);
    // one box per antenna, each plane on its own 1024-byte boundary: the swizzle phase of a (row, chunk) is then the
    // same in every plane and the walk addresses all antennas from ONE swizzled pointer plus immediate offsets
#pragma unroll
    for (int j = 0; j < NTX; ++j)
      tma::load_4d(xs + j * kTmaPlaneBytes, &xmap, 0, q * (kTmaTile / 16) - tp.hrows, chunk * NTX + j, b, bar);
    const uint32_t ax = xs + tp.stage_bytes;
    const int qp = (q * kTmaTile) / tp.poly_tile;
    tma::load_1d(ax, a.coef + ((size_t)b * tp.npoly + qp) * tp.coef_stride, tp.coef_bytes, bar);
    tma::load_1d(ax + tp.s_off, a.spatial32 + ((size_t)b * tp.nchunks + chunk) * tp.s_stride, tp.s_bytes, bar);
  };

  if (tid == 0) {
    if (ring0 & 1023u) __trap();  // the swizzle phase is derived from absolute shared addresses
    tma::prefetch_map(&xmap);
    for (int s = 0; s < NS; ++s) {
      tma::mbar_init(bar0 + 8u * s, 1);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt0 + 4u * s), "r"(0u) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(done0 + 4u * s), "r"(0u) : "memory");
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(next0), "r"(0u) : "memory");
    asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(rptr0), "r"(0u) : "memory");
    tma::fence_barrier_init();
    // the first NS tiles of every CTA are static, the rest are claimed from the global counter as slots free up
    for (int s = 0; s < NS; ++s) request((int)blockIdx.x + s * step, s);
  }
  __syncthreads();

  const float inv = 1.0f / (float)tp.poly_tile;
  // ticket = (ring position k, part): every warp of the CTA pulls the next 256-output part on its own.  The ticket
  // of the NEXT part is drawn right after the walk, so the shared-memory atomic completes under the epilogue.
  uint32_t ticket = 0;
  if (lane == 0) asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(ticket) : "r"(next0) : "memory");
  for (;;) {
    const uint32_t kw = __shfl_sync(0xffffffffu, ticket, 0);
    const uint32_t k = kw / kTmaParts, part = kw % kTmaParts;
    const uint32_t use = k / (uint32_t)NS, slot = k - use * (uint32_t)NS;
    tma::mbar_wait(bar0 + 8u * slot, use & 1u);
    int t;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(tid0 + 4u * slot) : "memory");
    if (t < 0) break;
    int b = t / tp.ntiles, tx0 = a.tx0;
    const int q = t - b * tp.ntiles;
    if constexpr (ZMODE) {
      tx0 = (b % tp.nchunks) * NTX;
      b /= tp.nchunks;
    }
    const int trow = 32 * (int)part + lrow;
    const int m0 = q * kTmaTile + R * trow;
    const bool active = m0 < Tout;
    const uint32_t xs = ring0 + slot * tp.slot_bytes;
    const uint32_t csa0 = xs + tp.stage_bytes;

    u64 acc[R][NTX];
#pragma unroll
    for (int u = 0; u < R; ++u)
#pragma unroll
      for (int j = 0; j < NTX; ++j) acc[u][j] = 0ull;

    if (active) {
      const int qp = (q * kTmaTile) / tp.poly_tile;
      const float r0 = ((float)(m0 - qp * tp.poly_tile) - 0.5f * (float)tp.poly_tile) * inv;
      const float rc = fmaf(0.5f * (float)(R - 1), inv, r0);
      float rr[LIN ? 1 : R];
      if constexpr (!LIN) {
#pragma unroll
        for (int u = 0; u < R; ++u) rr[u] = fmaf((float)u, inv, r0);
      }
      // window at d = 0: x[m0 .. m0+7] = pairs pe0 .. pe0+3, pe0 = 8 hrows + 4 trow (natural offset 16 pe0)
      u64 w[NTX][R];
      uint32_t nat = xs + (uint32_t)tp.hrows * 128u + 64u * (uint32_t)trow;  // antenna 0, first pair of the window
      {
        const uint32_t s0 = tma::swz(nat);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
          const uint32_t ak = s0 ^ (16u * k);
#pragma unroll
          for (int j = 0; j < NTX; ++j) tma::lds_pair_at<0>(ak + j * kTmaPlaneBytes, w[j][2 * k], w[j][2 * k + 1]);
        }
      }

      uint32_t csa = csa0;
      u64 hp[NH];
      auto prep = [&]() {  // consume the coefficients of the next delay group (one group past the end is padding)
        u64 cf[P];
#pragma unroll
        for (int p = 0; p < P; ++p) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(cf[p]) : "r"(csa + p * 8));
        csa += P * 8;
        if constexpr (LIN) {
          const u64 rcb = pk2(rc, rc);
          const float fp = (float)(P - 1) * inv;
          u64 hs = fma2(cf[P - 1], pk2(fp, fp), 0ull);
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

      for (int c = 0; c < tp.nblk; ++c) {
        nat -= 64u;  // pairs pe0 - 4 (c + 1) .. + 3: the half row entering during this block
        const uint32_t mk = tp.mask[c];
        if (mk == 0u) continue;
        const uint32_t pm = mk & 0xffu, lm = mk >> 8;
        const uint32_t s0 = tma::swz(nat);
#pragma unroll
        for (int s = 0; s < R; ++s) {
          auto load_pair = [&]() {  // odd s: (x[m0-d-1], x[m0-d]) -> slots 7 - s, 8 - s; chunk 3 - (s - 1) / 2 of the half row
            if ((lm >> s) & 1u) {
              const uint32_t ak = s0 ^ (16u * (3 - (s >> 1)));
#pragma unroll
              for (int j = 0; j < NTX; ++j) tma::lds_pair_at<0>(ak + j * kTmaPlaneBytes, w[j][7 - s], w[j][(8 - s) & 7]);
            }
          };
          if ((pm >> s) & 1u) {
            u64 hq[NH];
#pragma unroll
            for (int k = 0; k < NH; ++k) hq[k] = hp[k];
            prep();
            auto mac_u = [&](int u) {
              u64 hv;
              if constexpr (LIN) {
                const float ku = (float)u - 0.5f * (float)(R - 1);
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
              for (int j = 0; j < NTX; ++j) cmac2(acc[u][j], w[j][(u - s + R) % R], hre, him);
            };
            if (s & 1) {
              mac_u(R - 1);
              load_pair();
#pragma unroll
              for (int u = 1; u < R - 1; ++u) mac_u(u);
              mac_u(0);
            } else {
              // output 0 reads the element that entered with the pair loaded at the previous (odd) phase: last
#pragma unroll
              for (int u = 1; u < R; ++u) mac_u(u);
              mac_u(0);
            }
          } else if (s & 1) {
            load_pair();
          }
        }
      }
    }

    if (lane == 0) asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(ticket) : "r"(next0) : "memory");

    if (active) {
      const int pitch = (ZMODE && a.ypitch > 0) ? a.ypitch : Tout;  // rows of the z workspace start 16-byte aligned
      const bool vec_ok = ((pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0) && (m0 + R <= Tout) &&
                          !a.accumulate;
      auto store_row = [&](float2* dst, const u64 (&yv)[R]) {
        if (vec_ok) {
#pragma unroll
          for (int u = 0; u < R; u += 2) stg_stream4(dst + u, yv[u], yv[u + 1]);
        } else {
#pragma unroll
          for (int u = 0; u < R; ++u) {
            if (m0 + u < Tout) {
              float2 v = upk2(yv[u]);
              if (a.accumulate) {
                const float2 old = dst[u];
                v.x += old.x;
                v.y += old.y;
              }
              stg_stream(dst + u, v);
            }
          }
        }
      };
      if constexpr (ZMODE) {
        // ---- large arrays: the spatial product runs on the tensor cores (spatial_gemm.cuh); store z itself ----------
        float2* zb = reinterpret_cast<float2*>(a.y) + ((size_t)b * a.ntx + tx0) * pitch + m0;
#pragma unroll
        for (int j = 0; j < NTX; ++j) {
          if (tx0 + j < a.ntx) {
            u64 yv[R];
#pragma unroll
            for (int u = 0; u < R; ++u) yv[u] = acc[u][j];
            store_row(zb + (size_t)j * pitch, yv);
          }
        }
      } else {
        // ---- spatial mix  y[irx] = sum_j S[irx][j] z[j]  and direct stores of the thread's R consecutive outputs ----
        float2* yb = reinterpret_cast<float2*>(a.y) + (size_t)b * a.nrx * Tout + m0;
        const uint32_t ssa = csa0 + tp.s_off;
        u64 snext[NTX];  // row irx + 1 of S is fetched while row irx is applied (the aux slot has padding past the end)
#pragma unroll
        for (int j = 0; j < NTX; ++j) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(snext[j]) : "r"(ssa + j * 8));
        for (int irx = 0; irx < a.nrx; ++irx) {
          u64 yv[R], scur[NTX];
#pragma unroll
          for (int u = 0; u < R; ++u) yv[u] = 0ull;
#pragma unroll
          for (int j = 0; j < NTX; ++j) {
            scur[j] = snext[j];
            asm volatile("ld.shared.b64 %0, [%1];" : "=l"(snext[j]) : "r"(ssa + ((irx + 1) * NTX + j) * 8));
          }
#pragma unroll
          for (int j = 0; j < NTX; ++j) {
            const float2 sc = upk2(scur[j]);
            const u64 sre = pk2(sc.x, sc.x), sim = pk2(-sc.y, sc.y);
#pragma unroll
            for (int u = 0; u < R; ++u) cmac2(yv[u], acc[u][j], sre, sim);
          }
          store_row(yb + (size_t)irx * Tout, yv);
        }
      }
    }

    // Release.  The warp that finishes the LAST part of a tile marks its slot done and runs the retire loop: ring
    // positions retire IN ORDER, and each retirement claims the next unprocessed tile from the global counter (SMs that
    // run slower simply claim fewer) and requests it into the freed slot.  In-order requests make the end markers a
    // suffix of the ring sequence, so a warp may leave at the first marker it meets without stranding a later tile.
    // No barrier anywhere in the loop: warps drift apart by up to NS - 1 tiles; loads, walks and stores overlap.
    __syncwarp();
    if (lane == 0) {
      uint32_t old;
      asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(cnt0 + 4u * slot) : "memory");
      if (old == kTmaParts - 1) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt0 + 4u * slot), "r"(0u) : "memory");
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(done0 + 4u * slot), "r"(1u) : "memory");
        uint32_t busy;
        do {
          asm volatile("atom.acquire.cta.shared::cta.cas.b32 %0, [%1], 0, 1;" : "=r"(busy) : "r"(lock0) : "memory");
        } while (busy != 0u);
        for (;;) {
          uint32_t pos, flag;
          asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(pos) : "r"(rptr0) : "memory");
          const uint32_t rs = pos % (uint32_t)NS;
          asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(flag) : "r"(done0 + 4u * rs) : "memory");
          if (flag == 0u) break;
          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(done0 + 4u * rs), "r"(0u) : "memory");
          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(rptr0), "r"(pos + 1u) : "memory");
          request((int)atomicAdd(tp.tile_counter, 1u) + NS * step, (int)rs);
        }
        asm volatile("atom.release.cta.shared::cta.exch.b32 %0, [%1], 0;" : "=r"(busy) : "r"(lock0) : "memory");
      }
    }
  }
}

template <int NTX>
int launch_tdl_tma(int P, bool lin, bool zmode, const FadingArgs& a, const TmaPlan& tp, const CUtensorMap& xmap, int grid,
                   size_t smem, cudaStream_t st);

}  // namespace hb
