// K4 for large arrays: the per-link spatial GEMM  Y[b] = S[b] @ Z[b]  (fading.py:395, `spatial_response @ propagated`)
// on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), sm_100a only.
//
// Real-valued form.  One MMA tile covers 64 complex time samples and up to 64 x 64 antennas:
//   A (M = 128, K = Ntx): row 2m = Re z[:, m], row 2m+1 = Im z[:, m]              (time is the M axis)
//   B (N = 2 Nrx, K = Ntx): row 2j = Re S[j, :], row 2j+1 = Im S[j, :]
//   D = A B^T:  D[2m, 2j] = P = Re z.Re S   D[2m, 2j+1] = Q = Re z.Im S
//               D[2m+1, 2j] = U = Im z.Re S D[2m+1, 2j+1] = V = Im z.Im S
//   y_re[j, m] = P - V,  y_im[j, m] = Q + U: the two TMEM lanes of a sample are neighbouring lanes of one warp, so
//   the combination is ONE shuffle per output, and the warp's store of one receive stream is 128 contiguous bytes.
//
// 3xTF32.  Every FP32 operand is split x = hi + lo with hi = cvt.rna.tf32(x); D accumulates lo.hi + hi.lo + hi.hi
// in FP32 (the lo.lo term is below 2^-22).  Measured against an FP64 GEMM: relative L2 error ~1e-7.
//
// Shared-memory operands use the no-swizzle K-major canonical layout (8 rows x 16 bytes core matrices, 128 bytes
// each; row groups SBO = 128 bytes apart, 16-byte K chunks LBO = 2048 bytes apart), validated bit-exactly by
// tools/microbench/umma_probe.cu, which also measured the issue cost: 99 clk per 128x128x8 MMA from shared memory.
//
// Schedule: persistent CTA (one per SM), work item = (link, segment of consecutive tiles): S is converted and split once
// per item.  256 WORKER threads stage z (global -> registers -> hi/lo -> shared, loads prefetched one tile ahead) and run
// the epilogue of the previous tile (tcgen05.ld -> shuffle -> global); two ISSUER warps (tiles of even / odd index, each
// with its own TMEM accumulator and A buffer) issue the 3 x K/8 MMAs.  Round 1 issued from thread 0 of worker warp 0:
// its 24 MMAs (~99 clk each from one thread) stalled that warp's epilogue and, through the CTA barrier, everyone else --
// 5286 clk per tile against 2376 of issue time.  Workers and issuers now meet only on mbarriers (`staged` / `mma_done`).
#pragma once
#include "hb_common.cuh"

namespace hb {

constexpr int kGemmThreads = 512;                 // worker threads (staging + epilogue): the CUDA-core side of a tile is ~5400
                                                  // warp instructions (hi / lo splits, pair shuffles), 16 warps hide its latencies
constexpr int kGemmChunksPerPass = kGemmThreads / 64;    // K chunks (4 antennas) staged per pass: 2 warps cover 64 samples
constexpr int kGemmColParts = kGemmThreads / 128;        // epilogue: column parts per TMEM lane quadrant
constexpr int kGemmPiecesPerWarp = 4 / kGemmColParts;    // 32-column pieces (16 receive streams) per epilogue warp
constexpr int kGemmLaunchThreads = kGemmThreads + 64;  // + two MMA-issuing warps
constexpr int kGemmTileSamples = 64;              // complex samples per MMA tile (M = 128 real rows)
constexpr int kGemmMaxAnt = 64;                   // antennas per block on either side
constexpr uint32_t kGemmLbo = 2048;               // bytes between 16-byte K chunks (16 row groups x 128 B)
constexpr uint32_t kGemmSbo = 128;                // bytes between 8-row groups
constexpr uint32_t kGemmOperandBytes = 128 * kGemmMaxAnt * 4;  // one 128 x 64 fp32 operand = 32 KB
constexpr size_t kGemmSmemBytes = 6 * (size_t)kGemmOperandBytes + 1024;  // B hi/lo + 2 x A hi/lo (+ alignment slack)

struct GemmArgs {
  const double2* S;  // [B, nrx_total, ntx_total] complex128
  const float2* z;   // [B, ntx_total, ldz]  (ldz >= T samples per row)
  int ldz;
  float2* y;         // [B, nrx_total, T]
  int B, T;
  int nrx_total, ntx_total;
  int rx0, nrx;      // receive block [rx0, rx0 + nrx), nrx <= 64
  int tx0, ntx;      // transmit block
  int accumulate;    // y += (transmit blocks after the first)
  int ntiles;        // tiles per link
  int seg_tiles;     // tiles per work item
  int nseg;          // segments per link
};

namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, leading / stride byte offsets (16-byte units), version 1, no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// issued from warp-convergent code by the elected lane: the same lane for a given member mask every time
__device__ __forceinline__ void mma_tf32_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 TMEM lanes (this warp's quadrant) x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = hi + lo, hi rounded to TF32 (10 explicit mantissa bits); the tensor core ignores the low 13 bits of lo
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace umma

// byte offset of element (row r, column k) inside a K-major operand tile
__device__ __forceinline__ uint32_t gemm_operand_offset(int r, int k) {
  return (uint32_t)(r >> 3) * kGemmSbo + (uint32_t)(k >> 2) * kGemmLbo + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 3) * 4;
}

// FULL: a full 64 x 64 antenna block without accumulation -- interior tiles (all 64 samples inside the frame) then run
// without a single bounds predicate: plain loads, pair-shuffled 8-byte stores.
template <bool FULL>
__global__ void __launch_bounds__(kGemmLaunchThreads, 1) spatial_gemm_3xtf32_kernel(const GemmArgs a) {
  using namespace umma;
  extern __shared__ unsigned char gemm_smem_raw[];
  __shared__ uint64_t mma_done[2], staged[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem0 = (smem_addr(gemm_smem_raw) + 1023u) & ~1023u;
  const uint32_t sB_hi = smem0, sB_lo = smem0 + kGemmOperandBytes;
  // A buffers: [buf][hi, lo]
  const uint32_t sA0 = smem0 + 2 * kGemmOperandBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = (a.ntx + 7) & ~7;          // K padded to the MMA depth
  const int Np = (2 * a.nrx + 15) & ~15;    // N padded to the M = 128 granularity
  const int ksteps = Kp >> 3;
  const int kchunks = Kp >> 2;              // 16-byte K chunks (4 antennas)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_init(&staged[0], kGemmThreads);
    mbar_init(&staged[1], kGemmThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t idesc = instr_desc_tf32(128, Np);

  // staging task of this thread in pass p: K chunk kc = kGemmChunksPerPass p + (warp >> 1), sample ms = 32 (warp & 1) + lane
  const int ms = ((warp & 1) << 5) + lane;
  const int kc0 = warp >> 1;
  constexpr int kMaxPass = kGemmMaxAnt / 4 / kGemmChunksPerPass;  // passes of kGemmChunksPerPass chunks cover 64 antennas
  const uint32_t a_off = (uint32_t)(ms >> 2) * kGemmSbo + (uint32_t)(ms & 3) * 32;  // rows 2 ms, 2 ms + 1

  // epilogue role: TMEM lane quadrant and column part
  const int quad = warp & 3, cpart = warp >> 2;
  const int em = (quad << 4) + (lane >> 1);  // sample of this thread's TMEM lane inside the tile
  const int comp = lane & 1;                 // 0: real row, 1: imaginary row

  uint32_t phase_bits = 0u;  // bit s: parity the next wait on accumulator slot s expects

  if (warp >= kGemmThreads / 32) {
    // ================================ MMA issue: warp 8 + w takes the tiles of local index i = w (mod 2) ================
    const int w = warp - kGemmThreads / 32;
    const uint32_t ahi = sA0 + (uint32_t)w * 2 * kGemmOperandBytes, alo = ahi + kGemmOperandBytes;
    const uint32_t d = tmem + (uint32_t)w * 128;
    uint32_t use = 0;  // tiles this warp has issued: phase of staged[w]
    for (int item = blockIdx.x; item < a.B * a.nseg; item += gridDim.x) {
      const int seg = item % a.nseg;
      const int ntile = min(a.ntiles, (seg + 1) * a.seg_tiles) - seg * a.seg_tiles;
      for (int i = w; i < ntile; i += 2, ++use) {
        mbar_wait(&staged[w], use & 1u);  // all workers: A buffer w (and S) staged, accumulator w drained
        fence_after_sync();
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t adv = (uint32_t)ks * 2 * kGemmLbo;
          const uint64_t dah = smem_desc(ahi + adv, kGemmLbo, kGemmSbo), dal = smem_desc(alo + adv, kGemmLbo, kGemmSbo);
          const uint64_t dbh = smem_desc(sB_hi + adv, kGemmLbo, kGemmSbo), dbl = smem_desc(sB_lo + adv, kGemmLbo, kGemmSbo);
          mma_tf32_elect(d, dal, dbh, idesc, (uint32_t)ks);
          mma_tf32_elect(d, dah, dbl, idesc, 1u);
          mma_tf32_elect(d, dah, dbh, idesc, 1u);
        }
        commit_elect(&mma_done[w]);
      }
    }
  } else

  for (int item = blockIdx.x; item < a.B * a.nseg; item += gridDim.x) {
    const int b = item / a.nseg, seg = item - b * a.nseg;
    const int t_begin = seg * a.seg_tiles, t_end = min(a.ntiles, t_begin + a.seg_tiles);
    const int ntile = t_end - t_begin;
    const float2* zb = a.z + ((size_t)b * a.ntx_total + a.tx0) * a.ldz;
    float* yb = reinterpret_cast<float*>(a.y + ((size_t)b * a.nrx_total + a.rx0) * a.T);

    // ---- B operand: S block, converted, split, zero padded (no MMA is in flight here) -----------------------------
    {
      const double2* Sb = a.S + ((size_t)b * a.nrx_total + a.rx0) * a.ntx_total + a.tx0;
      for (int i = tid; i < (Np >> 1) * Kp; i += kGemmThreads) {
        const int j = i / Kp, k = i - j * Kp;
        float re = 0.f, im = 0.f;
        if (j < a.nrx && k < a.ntx) {
          const double2 s = Sb[(size_t)j * a.ntx_total + k];
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
    }

    float2 zr[kMaxPass][4];  // prefetched z elements of the next tile to stage
    auto load_tile = [&](int t) {
      const int n = t * kGemmTileSamples + ms;
      if (FULL && (t + 1) * kGemmTileSamples <= a.T) {  // CTA-uniform
        const float2* zp = zb + (size_t)(4 * kc0) * a.ldz + n;
#pragma unroll
        for (int p = 0; p < kMaxPass; ++p)
#pragma unroll
          for (int i = 0; i < 4; ++i) zr[p][i] = ldg_stream(zp + (size_t)(4 * kGemmChunksPerPass * p + i) * a.ldz);
        return;
      }
#pragma unroll
      for (int p = 0; p < kMaxPass; ++p) {
        const int kc = kGemmChunksPerPass * p + kc0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = 4 * kc + i;
          float2 v = make_float2(0.f, 0.f);
          if (kc < kchunks && k < a.ntx && n < a.T) v = ldg_stream(zb + (size_t)k * a.ldz + n);
          zr[p][i] = v;
        }
      }
    };
    auto stage_tile = [&](int buf) {
      const uint32_t hi0 = sA0 + (uint32_t)buf * 2 * kGemmOperandBytes + a_off, lo0 = hi0 + kGemmOperandBytes;
#pragma unroll
      for (int p = 0; p < kMaxPass; ++p) {
        const int kc = kGemmChunksPerPass * p + kc0;
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
    // hand buffer `buf` (A operand staged, accumulator drained by this thread's earlier epilogue) to its issuer warp
    auto hand_over = [&](int buf) {
      fence_async_smem();
      fence_before_sync();
      mbar_arrive(&staged[buf]);
    };
    auto epilogue = [&](int t, int buf) {
      mbar_wait(&mma_done[buf], (phase_bits >> buf) & 1u);
      phase_bits ^= 1u << buf;
      fence_after_sync();
      const int n = t * kGemmTileSamples + em;
      const int piece0 = cpart * kGemmPiecesPerWarp;  // first 32-column piece (16 receive streams) of this warp
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * 128 + (uint32_t)piece0 * 32;
      if (FULL && (t + 1) * kGemmTileSamples <= a.T) {  // CTA-uniform
        // Pair shuffle: for receive streams (j0, j0 + 1) the even lane (Re z row: P, Q) finishes stream j0 and the odd
        // lane (Im z row: U, V) finishes j0 + 1; each receives the partner's two coefficients of ITS stream, so every
        // thread stores one complete complex sample (8 bytes); a warp writes two 128-byte row segments per instruction.
        float2* yrow = reinterpret_cast<float2*>(yb) + (size_t)(piece0 * 16 + comp) * a.T + n;
#pragma unroll
        for (int h = 0; h < kGemmPiecesPerWarp; ++h) {
          uint32_t v[32];
          tmem_ld32(taddr + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) {
            // columns: (2 jj, 2 jj + 1) = stream j0, (2 jj + 2, 2 jj + 3) = stream j0 + 1
            const float a0 = __uint_as_float(v[2 * jj]), b0 = __uint_as_float(v[2 * jj + 1]);
            const float a1 = __uint_as_float(v[2 * jj + 2]), b1 = __uint_as_float(v[2 * jj + 3]);
            // even lane keeps (P0, Q0) and sends (P1, Q1); odd lane keeps (U1, V1) and sends (U0, V0)
            const float ra = __shfl_xor_sync(0xffffffffu, comp ? a0 : a1, 1);
            const float rb = __shfl_xor_sync(0xffffffffu, comp ? b0 : b1, 1);
            float2 out;
            if (comp) {  // own (U1, V1), received (P1, Q1): y = (P - V, Q + U)
              out.x = ra - b1;
              out.y = rb + a1;
            } else {     // own (P0, Q0), received (U0, V0)
              out.x = a0 - rb;
              out.y = b0 + ra;
            }
            stg_stream(yrow + (size_t)(h * 16 + jj) * a.T, out);
          }
        }
        return;
      }
#pragma unroll
      for (int h = 0; h < kGemmPiecesPerWarp; ++h) {
        if ((piece0 + h) * 32 < Np) {  // warp-uniform
          uint32_t v[32];
          tmem_ld32(taddr + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            // even lane (Re z row): keeps P, sends Q, receives V -> y_re = P - V
            // odd lane  (Im z row): keeps U, sends V, receives Q -> y_im = U + Q
            const float keep = __uint_as_float(v[2 * jj]);
            const float recv = __shfl_xor_sync(0xffffffffu, __uint_as_float(v[2 * jj + 1]), 1);
            float out = comp ? keep + recv : keep - recv;
            const int j = (piece0 + h) * 16 + jj;
            if (j < a.nrx && n < a.T) {
              float* dst = yb + ((size_t)j * a.T + n) * 2 + comp;
              if (a.accumulate) out += *dst;
              asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(dst), "f"(out) : "memory");
            }
          }
        }
      }
    };

    // ---- software pipeline over the tiles of this item -------------------------------------------------------
    load_tile(t_begin);
    stage_tile(0);
    if (ntile > 1) load_tile(t_begin + 1);
    hand_over(0);
    for (int i = 0; i < ntile; ++i) {
      if (i + 1 < ntile) {
        stage_tile((i + 1) & 1);  // buffer last read by the MMAs of tile i - 1 (waited for in its epilogue)
        if (i + 2 < ntile) load_tile(t_begin + i + 2);
        hand_over((i + 1) & 1);   // accumulator last read by this thread's epilogue of tile i - 1
      }
      epilogue(t_begin + i, i & 1);
    }
    // every MMA of the item has completed (each worker waited for them in its epilogues): S and the A buffers may be
    // overwritten by this thread; the accumulators are re-used only after all 256 workers handed the next tile over
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

}  // namespace hb
