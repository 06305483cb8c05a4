// tcgen05.mma kind::f16 (BF16 operands, FP32 accumulate) variant of umma_shift_probe.cu: K = 16 per instruction, rows of two
// 16-byte chunks of 8 bf16; validates the shifted start address and measures the N = 96 / 64 / 32 triple of a BF16x3 scheme.
//
// Question 1 (correctness): in the no-swizzle K-major canonical layout with SBO = 128 bytes the rows of one 16-byte
// K chunk are LINEAR in shared memory (row r at 16 r).  Does a matrix descriptor whose start address is advanced by
// an arbitrary number of rows (16-byte granularity, not a multiple of the 8-row core matrix) address rows
// s .. s + 127?  If yes, the delayed copies x[:, m - k_g] of a delay line are free: one staged tile, one
// descriptor per delay.
// Question 2 (rate): issue cost of 128 x N x 8 MMAs with N = 32 / 64 and a different A start address per MMA
// (the A operand is re-read from shared memory for every delay group; N is small, so A bandwidth may bind).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_shift_probe umma_shift_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // D f32, A = B = bf16
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// MMA issued from warp-convergent code by one elected lane: operands stay in uniform registers (no R2UR per issue)
__device__ __forceinline__ void mma_tf32_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
               "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

constexpr int kRows = 640;  // staged rows of A (time samples incl. halo)
constexpr int kK = 16;      // one MMA K step: two 16-byte chunks of 8 bf16

// mode 0: D = A[shift : shift + 128] B^T, one MMA (N columns).
// mode 1: rate.  reps x groups MMAs; MMA i uses A start row (7 i) % 500 and B block i % groups; per group one
//         N = n1 MMA and (if n2) one N = n2 MMA, into 4 rotating accumulators.
struct Cfg { int mode, N, shift, reps, groups, n1, n2, nw, n3; };

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;                               // [chunk 2][row kRows][16 B]
  uint8_t* sB = smem + 2 * kRows * 16;              // [group][chunk 2][row 64][16 B]
  const uint32_t a_lbo = kRows * 16, b_lbo = 96 * 16, b_group = 2 * 96 * 16;
  for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
    int r = i / kK, k = i % kK;
    *(__nv_bfloat16*)(sA + (k >> 3) * a_lbo + r * 16 + (k & 7) * 2) = __float2bfloat16(A[i]);
  }
  for (int i = threadIdx.x; i < c.groups * 96 * kK; i += blockDim.x) {
    int g = i / (96 * kK), r = (i / kK) % 96, k = i % kK;
    *(__nv_bfloat16*)(sB + g * b_group + (k >> 3) * b_lbo + r * 16 + (k & 7) * 2) = __float2bfloat16(B[i]);
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) mbar_init(&bar, c.mode == 0 ? 1 : c.nw);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  long long t0 = 0, t1 = 0;
  if (c.mode == 0) {
    if (threadIdx.x == 0) {
      mma_tf32(tm, make_desc(smem_u32(sA) + c.shift * 16, a_lbo, 128), make_desc(smem_u32(sB), b_lbo, 128), make_idesc(128, c.N), 0u);
      mma_commit(&bar);
    }
  } else if (warp < c.nw) {
    {
      const uint32_t id1 = make_idesc(128, c.n1), id2 = make_idesc(128, c.n2 ? c.n2 : 32);
      const uint64_t a0 = make_desc(smem_u32(sA), a_lbo, 128), b0 = make_desc(smem_u32(sB), b_lbo, 128);
      t0 = clock64();
      uint32_t i = 0;
      for (int rep = 0; rep < c.reps; ++rep) {
        for (int g = 0; g < c.groups; ++g, ++i) {
          const uint64_t ad = a0 + (uint64_t)((7u * i + 128u * warp) % 500u);          // start address field counts 16-byte rows
          const uint64_t bd = b0 + (uint64_t)(g * (b_group >> 4));
          const uint32_t d = tm + (uint32_t)warp * 128u;
          mma_tf32_elect(d, ad, bd, id1, rep > 0 || g > 0 ? 1u : 0u);
          if (c.n2) mma_tf32_elect(d, ad + 3, bd, id2, 1u);
          if (c.n3) mma_tf32_elect(d, ad + 5, bd, make_idesc(128, c.n3), 1u);
        }
      }
    }
    if ((threadIdx.x & 31) == 0) mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (threadIdx.x == 0) { t1 = clock64(); *cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (c.mode == 0) {
    for (int c0 = 0; c0 < c.N; c0 += 32) {
      uint32_t v[32];
      uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                     "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                     "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      int row = warp * 32 + (threadIdx.x & 31);
      for (int j = 0; j < 32; ++j) D[row * c.N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}

int main() {
  const int G = 22;
  float *dA, *dB, *dD; long long* dcyc;
  CK(cudaMalloc(&dA, kRows * kK * 4)); CK(cudaMalloc(&dB, G * 96 * kK * 4)); CK(cudaMalloc(&dD, 128 * 96 * 4)); CK(cudaMalloc(&dcyc, 8));
  const size_t smem = 2 * kRows * 16 + G * 2 * 96 * 16 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> A(kRows * kK), B(G * 96 * kK);
  srand(99);
  for (auto& v : A) v = (float)(rand() % 9 - 4);
  for (auto& v : B) v = (float)(rand() % 9 - 4);
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  int fails = 0;
  for (int N : {32, 64, 96}) for (int shift : {0, 1, 2, 3, 5, 8, 13, 77, 200, 511}) {
    std::vector<float> D(128 * N), R(128 * N);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < kK; ++k) s += A[(m + shift) * kK + k] * B[n * kK + k]; R[m * N + n] = s; }
    CK(cudaMemset(dD, 0, 128 * 96 * 4));
    Cfg c{0, N, shift, 1, 1, 0, 0, 1, 0};
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d shift=%d: CUDA error %s\n", N, shift, cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; for (int i = 0; i < 128 * N; ++i) bad += (D[i] != R[i]);
    printf("shifted start: N=%2d row shift %3d : %s (%d / %d mismatches)\n", N, shift, bad ? "FAIL" : "ok", bad, 128 * N);
    fails += bad != 0;
  }
  struct R { int n1, n2, n3; };
  for (int nw : {1, 2, 4}) for (R r : {R{32, 0, 0}, R{64, 0, 0}, R{96, 0, 0}, R{96, 64, 32}}) {
    Cfg c{1, 0, 0, 400, G, r.n1, r.n2, nw, r.n3};
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c, dcyc);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const int n = 400 * G * nw;
    printf("rate: %d issuing warp(s), per delay group [128 x %d x 16%s], new A start address each: %.1f clk per group (%d groups)\n", nw,
           r.n1, r.n3 ? " + 128 x 64 x 16 + 128 x 32 x 16" : "", (double)cyc / n, n);
  }
  printf(fails ? "BF16 SHIFT PROBE FAILED\n" : "BF16 SHIFT PROBE OK\n");
  return fails ? 1 : 0;
}
