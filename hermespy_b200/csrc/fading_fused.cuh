// K4 + K3 fused for large arrays (16..64 antennas per side): spatial GEMM on the tensor cores, tap delay lines on its
// accumulator -- the intermediate never touches HBM (sm_100a).
//
//   reference:  z[:, d_l + n] += x[:, n] h_l[n];  y = S z                           hermespy/channel/fading/fading.py:385-393
//   here:       u = S x  (tcgen05, 3xTF32);  y[:, m] = sum_g h~_g(m) u[:, m - d_g]
//
// The tap gains are scalars common to every antenna, so the spatial matrix commutes with the delay line: applying S
// FIRST makes the GEMM's A operand the raw signal (staged exactly like spatial_gemm_3xtf32_kernel does) and leaves the
// delay line to run on the GEMM's output while it is still on chip.  The two-kernel path (tdl_tma_kernel in z mode, then
// the GEMM) writes and re-reads z[B, Ntx, T + D]: 2 x the algorithmic bytes.  Here HBM sees x once and y once.
//
// One persistent CTA per SM, work item = (link, segment of consecutive 64-sample tiles).  Per tile t:
//   1. all threads: x tile -> hi / lo TF32 halves -> A operand in shared memory (loads prefetched one tile ahead)
//   2. one thread: 3 x K/8 tcgen05.mma into the TMEM accumulator (u tile = 64 samples x Nrx streams)
//   3. all threads, WHILE the tensor core works: delay line of tile t-1 out of the u ring (shared memory, the last three
//      tiles = 192 samples >= 64 + max delay), tap gains by Horner from the K1 Taylor coefficients, y stored to HBM
//   4. all threads: accumulator -> registers -> u ring slot t mod 3
// A segment starts two tiles early (u only, no output) so that its first outputs find their delay history in the ring.
//
// Shared memory: S hi/lo 64 KB + A hi/lo 64 KB + u ring 96 KB = 224 KB.  Bound: shared-memory bandwidth (per tile 192 KB
// of MMA operand reads + G x 32 KB of ring reads + 96 KB of stores), see DESIGN.md section 4.
#pragma once
#include "fading_kernels.cuh"
#include "spatial_gemm.cuh"

namespace hb {

constexpr int kFusedThreads = 512;  // 16 warps at <= 128 registers: the delay line is latency bound, warps hide it
constexpr int kFusedRingTiles = 3;
constexpr int kFusedRing = kFusedRingTiles * kGemmTileSamples;  // 192 samples of u history per receive stream
constexpr int kFusedMaxDelay = kFusedRing - kGemmTileSamples;   // 128
constexpr int kFusedMaxGroups = 64;
constexpr int kFusedMaxOrder = 4;
constexpr size_t kFusedSmemBytes = 4 * (size_t)kGemmOperandBytes + (size_t)kGemmMaxAnt * kFusedRing * 8 + 128;

struct FusedArgs {
  const double2* S;   // [B, nrx, ntx] complex128
  const float2* x;    // [B, ntx, T]
  float2* y;          // [B, nrx, T + D]
  const float2* coef; // [B, npoly, coef_stride >= G * P]  (K1 output)
  int B, T, D, ntx, nrx;
  int ntiles;         // 64-sample tiles covering T + D
  int seg_tiles, nseg;
  int poly_tile, npoly, coef_stride;
  int num_groups;
  int group_delay[kFusedMaxGroups];
};

template <int P>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_gemm_tdl_kernel(const __grid_constant__ FusedArgs a) {
  using namespace umma;
  extern __shared__ unsigned char fused_smem_raw[];
  __shared__ uint64_t mma_done;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float2 coef_s[kFusedMaxGroups * kFusedMaxOrder];  // Taylor coefficients of the tile the delay line works on

  const uint32_t smem0 = (smem_addr(fused_smem_raw) + 127u) & ~127u;  // no swizzle: operand tiles need 16-byte alignment only
  const uint32_t sB_hi = smem0, sB_lo = smem0 + kGemmOperandBytes;
  const uint32_t sA_hi = smem0 + 2 * kGemmOperandBytes, sA_lo = smem0 + 3 * kGemmOperandBytes;
  float2* ring = reinterpret_cast<float2*>(fused_smem_raw + (smem0 - smem_addr(fused_smem_raw)) + 4 * kGemmOperandBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = (a.ntx + 7) & ~7;        // K padded to the MMA depth
  const int Np = (2 * a.nrx + 15) & ~15;  // N padded to the M = 128 granularity
  const int ksteps = Kp >> 3;
  const int kchunks = Kp >> 2;            // 16-byte K chunks (4 antennas)
  const int Tout = a.T + a.D;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t idesc = instr_desc_tf32(128, Np);

  // staging task of this thread in pass p: K chunk kc = 8 p + (warp >> 1), sample ms = 32 (warp & 1) + lane
  const int ms = ((warp & 1) << 5) + lane;
  const int kc0 = warp >> 1;
  constexpr int kChunksPerPass = kFusedThreads / 64;                  // 8
  constexpr int kMaxPass = kGemmMaxAnt / 4 / kChunksPerPass;          // 2 passes cover 64 antennas
  const uint32_t a_off = (uint32_t)(ms >> 2) * kGemmSbo + (uint32_t)(ms & 3) * 32;  // rows 2 ms, 2 ms + 1

  // accumulator read-out role: TMEM lane quadrant (warp % 4, as the hardware demands) and 32-column group
  const int quad = warp & 3, cgrp = warp >> 2;
  const int em = (quad << 4) + (lane >> 1);  // sample of this thread's TMEM lane inside the tile
  const int comp = lane & 1;                 // 0: real row, 1: imaginary row

  // delay-line role: output sample dm of the tile, receive streams [8 rq, 8 rq + 8)
  constexpr int kRxPerThread = kGemmMaxAnt / (kFusedThreads / 64);  // 8
  const int dm = tid & 63, rq = tid >> 6;
  const float inv_poly = 1.0f / (float)a.poly_tile;

  uint32_t phase = 0u;

  for (int item = blockIdx.x; item < a.B * a.nseg; item += gridDim.x) {
    const int b = item / a.nseg, seg = item - b * a.nseg;
    const int t_begin = seg * a.seg_tiles, t_end = min(a.ntiles, t_begin + a.seg_tiles);
    const int t_first = max(0, t_begin - (kFusedRingTiles - 1));  // history tiles: u only
    const float2* xb = a.x + (size_t)b * a.ntx * a.T;
    float2* yb = a.y + (size_t)b * a.nrx * Tout;
    const float2* cb = a.coef + (size_t)b * a.npoly * a.coef_stride;

    // ---- B operand: S, converted, split, zero padded; u ring cleared (no MMA in flight, no reader of the ring) -------
    {
      const double2* Sb = a.S + (size_t)b * a.nrx * a.ntx;
      for (int i = tid; i < (Np >> 1) * Kp; i += kFusedThreads) {
        const int j = i / Kp, k = i - j * Kp;
        float re = 0.f, im = 0.f;
        if (j < a.nrx && k < a.ntx) {
          const double2 s = Sb[(size_t)j * a.ntx + k];
          re = (float)s.x;
          im = (float)s.y;
        }
        float h, l;
        const uint32_t o = gemm_operand_offset(2 * j, k);
        split_tf32(re, h, l);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB_hi + o), "f"(h) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB_lo + o), "f"(l) : "memory");
        split_tf32(im, h, l);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB_hi + o + 16), "f"(h) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sB_lo + o + 16), "f"(l) : "memory");
      }
      float4* r4 = reinterpret_cast<float4*>(ring);
      for (int i = tid; i < kGemmMaxAnt * kFusedRing / 2; i += kFusedThreads) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    float2 zr[kMaxPass][4];  // prefetched x elements of the next tile to stage
    auto load_tile = [&](int t) {
      const int n = t * kGemmTileSamples + ms;
      if (a.ntx == kGemmMaxAnt && (t + 1) * kGemmTileSamples <= a.T) {  // CTA-uniform: interior tile of a full block
        const float2* xp = xb + (size_t)(4 * kc0) * a.T + n;
#pragma unroll
        for (int p = 0; p < kMaxPass; ++p)
#pragma unroll
          for (int i = 0; i < 4; ++i) zr[p][i] = ldg_stream(xp + (size_t)(4 * kChunksPerPass * p + i) * a.T);
        return;
      }
#pragma unroll
      for (int p = 0; p < kMaxPass; ++p) {
        const int kc = kChunksPerPass * p + kc0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = 4 * kc + i;
          float2 v = make_float2(0.f, 0.f);
          if (kc < kchunks && k < a.ntx && n < a.T) v = ldg_stream(xb + (size_t)k * a.T + n);  // u = 0 past the frame
          zr[p][i] = v;
        }
      }
    };
    auto stage_tile = [&]() {
      const uint32_t hi0 = sA_hi + a_off, lo0 = sA_lo + a_off;
#pragma unroll
      for (int p = 0; p < kMaxPass; ++p) {
        const int kc = kChunksPerPass * p + kc0;
        if (kc < kchunks) {
          float hr[4], lr[4], hi_[4], li[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            split_tf32(zr[p][i].x, hr[i], lr[i]);
            split_tf32(zr[p][i].y, hi_[i], li[i]);
          }
          const uint32_t o = (uint32_t)kc * kGemmLbo;
          sts128(hi0 + o, hr[0], hr[1], hr[2], hr[3]);
          sts128(hi0 + o + 16, hi_[0], hi_[1], hi_[2], hi_[3]);
          sts128(lo0 + o, lr[0], lr[1], lr[2], lr[3]);
          sts128(lo0 + o + 16, li[0], li[1], li[2], li[3]);
        }
      }
    };
    auto issue_mma = [&]() {  // one thread
      uint32_t acc = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t adv = (uint32_t)ks * 2 * kGemmLbo;
        const uint64_t dah = smem_desc(sA_hi + adv, kGemmLbo, kGemmSbo), dal = smem_desc(sA_lo + adv, kGemmLbo, kGemmSbo);
        const uint64_t dbh = smem_desc(sB_hi + adv, kGemmLbo, kGemmSbo), dbl = smem_desc(sB_lo + adv, kGemmLbo, kGemmSbo);
        mma_tf32(tmem, dal, dbh, idesc, acc);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dah, dbh, idesc, 1u);
        acc = 1u;
      }
      commit(&mma_done);
    };
    // accumulator of tile t -> ring slot t mod 3.  Even TMEM lane (Re x row): P - V = Re u; odd lane (Im x row): U + Q = Im u;
    // the 32 lanes of a warp write 32 consecutive floats of one receive stream: conflict-free.
    auto drain = [&](int t) {
      mbar_wait(&mma_done, phase);
      phase ^= 1u;
      fence_after_sync();
      float* rbase = reinterpret_cast<float*>(ring) + ((t % kFusedRingTiles) * kGemmTileSamples + em) * 2 + comp;
      if (cgrp * 32 < Np) {  // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)cgrp * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float keep = __uint_as_float(v[2 * jj]);
          const float recv = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * jj + 1]), 1);
          const float out = comp ? keep + recv : keep - recv;
          const int j = cgrp * 16 + jj;
          if (j < a.nrx) rbase[(size_t)j * (kFusedRing * 2)] = out;
        }
      }
    };
    // delay line of tile t out of the ring:  y[i, m] = sum_g h~_g(m) u[i, m - d_g],  h~_g by Horner in the window coordinate
    auto delay_line = [&](int t) {
      const int m = t * kGemmTileSamples + dm;
      if (m >= Tout) return;
      const int qp = (t * kGemmTileSamples) / a.poly_tile;  // a tile never straddles a Taylor window (64 | poly_tile)
      const float r = ((float)(m - qp * a.poly_tile) - 0.5f * (float)a.poly_tile) * inv_poly;
      float2 acc[kRxPerThread];
#pragma unroll
      for (int j = 0; j < kRxPerThread; ++j) acc[j] = make_float2(0.f, 0.f);
      const int slot = (t % kFusedRingTiles) * kGemmTileSamples + dm + kFusedRing;  // + ring: keeps the difference positive
      const float2* rrow = ring + (size_t)(kRxPerThread * rq) * kFusedRing;
#pragma unroll 2
      for (int g = 0; g < a.num_groups; ++g) {
        float2 hv = coef_s[g * P + (P - 1)];
#pragma unroll
        for (int p = P - 2; p >= 0; --p) {
          const float2 c = coef_s[g * P + p];
          hv.x = fmaf(hv.x, r, c.x);
          hv.y = fmaf(hv.y, r, c.y);
        }
        // samples before the frame read ring slots no tile of this item has written yet: cleared at item start = zeros
        const int idx = (slot - a.group_delay[g]) % kFusedRing;
#pragma unroll
        for (int j = 0; j < kRxPerThread; ++j) cmac<float>(acc[j], rrow[(size_t)j * kFusedRing + idx], hv);
      }
      float2* yp = yb + (size_t)(kRxPerThread * rq) * Tout + m;
#pragma unroll
      for (int j = 0; j < kRxPerThread; ++j)
        if (kRxPerThread * rq + j < a.nrx) stg_stream(yp + (size_t)j * Tout, acc[j]);
    };
    // Taylor coefficients of tile t into shared memory (made visible by the barrier that precedes delay_line(t))
    auto load_coef = [&](int t) {
      const int qp = (t * kGemmTileSamples) / a.poly_tile;
      if (tid < a.num_groups * P) coef_s[tid] = __ldg(cb + (size_t)qp * a.coef_stride + tid);
    };

    // ---- pipeline over the tiles of this item ----------------------------------------------------------------
    load_tile(t_first);
    __syncthreads();  // B operand and cleared ring complete
    for (int t = t_first; t < t_end; ++t) {
      stage_tile();  // A operand of tile t (the MMAs of tile t - 1 were waited for in its drain)
      if (t + 1 < t_end) load_tile(t + 1);
      if (t - 1 >= t_begin) load_coef(t - 1);  // the previous delay_line finished before the barrier that ended tile t - 1
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (tid == 0) issue_mma();
      if (t - 1 >= t_begin) delay_line(t - 1);  // overlaps the MMAs of tile t
      __syncthreads();                          // every reader of ring slot (t - 3) mod 3 is done before it is overwritten
      drain(t);
      fence_before_sync();
      __syncthreads();  // ring slot t visible; accumulator and A operand free
      fence_after_sync();
    }
    if (t_end - 1 >= t_begin) {
      load_coef(t_end - 1);
      __syncthreads();
      delay_line(t_end - 1);
    }
    __syncthreads();  // ring, coefficients and B operand are rewritten by the next item
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int launch_fused_gemm_tdl(int P, const FusedArgs& a, cudaStream_t st);

}  // namespace hb
