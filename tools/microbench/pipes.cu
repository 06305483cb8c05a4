// Pipe-rate microbenchmarks on sm_100a: FFMA vs FFMA2 (packed / scalar-broadcast operand), LDS.64/LDS.128.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }

constexpr int ITERS = 4096;
constexpr int CH = 16;

__global__ void k_ffma(float* out, float a, float b) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
  float x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// acc(pair) = x(pair) * h(scalar broadcast) + acc(pair)
__global__ void k_ffma2_bcast(float* out, float a, float b) {
  u64 acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = pk(threadIdx.x + i, i);
  u64 x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = pk(a + i * threadIdx.x, b + threadIdx.x);
  float h = b * threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      u64 hh = pk(h, h);
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(x[i & 3]), "l"(hh));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float2 v = upk(acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_packed(float* out, float a, float b) {
  u64 acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = pk(threadIdx.x + i, i);
  u64 x[4], h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] = pk(a + i * threadIdx.x, b + threadIdx.x); h[i] = pk(b + i * threadIdx.x, a - i * threadIdx.x); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(x[i & 3]), "l"(h[(i >> 2) & 3]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float2 v = upk(acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// LDS: each thread reads VEC bytes at (tid + k + shift) elements
template <typename V>
__global__ void k_lds(float* out, int shift, int stride) {
  extern __shared__ __align__(16) unsigned char sm[];
  V* s = reinterpret_cast<V*>(sm);
  const int n = 4096;
  for (int i = threadIdx.x; i < n + 512; i += blockDim.x) { V v; memset(&v, 0, sizeof(v)); s[i] = v; }
  __syncthreads();
  float acc = 0;
  int base = threadIdx.x * stride + shift;
  for (int it = 0; it < ITERS / 16; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      V v = s[(base + k * 33 + it) & (n - 1)];
      acc += reinterpret_cast<float*>(&v)[0];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}



template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

// Mixed issue: NF2 packed FFMA2 and NF1 scalar FFMA per iteration on independent accumulators.  Answers whether the
// scalar FFMA can use FMA-pipe capacity that FFMA2 leaves idle (total lane-FMAs/clk/SM above 128 would say yes).
template <int NF2, int NF1>
__global__ void k_mix(float* out, float a, float b) {
  u64 acc2[NF2 > 0 ? NF2 : 1];
  float acc1[NF1 > 0 ? NF1 : 1];
#pragma unroll
  for (int i = 0; i < NF2; ++i) acc2[i] = pk(threadIdx.x + i, i);
#pragma unroll
  for (int i = 0; i < NF1; ++i) acc1[i] = threadIdx.x + i;
  u64 x2 = pk(a + threadIdx.x, b + threadIdx.x), h2 = pk(b, b);
  float x1 = a + threadIdx.x, h1 = b;
  constexpr int M = NF2 > NF1 ? NF2 : NF1;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (i < NF2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(x2), "l"(h2));
      if (i < NF1) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc1[i]) : "f"(x1), "f"(h1));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NF2; ++i) { float2 v = upk(acc2[i]); s += v.x + v.y; }
#pragma unroll
  for (int i = 0; i < NF1; ++i) s += acc1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF2, int NF1>
void run_mix(float* out, int clk, int blocks, int thr) {
  float ms = timeit([&] { k_mix<NF2, NF1><<<blocks, thr>>>(out, 1.0001f, 0.5f); });
  const double fmas = (double)blocks * thr * ITERS * (2.0 * NF2 + NF1);
  const double instr = (double)blocks * thr * ITERS * (NF2 + NF1);
  printf("mix FFMA2 x%2d + FFMA x%2d, %d x %d thr: %.3f ms  %.1f lane-FMA/clk/SM, %.1f instr-lanes/clk/SM\n", NF2, NF1, blocks, thr, ms,
         fmas / (ms * 1e-3) / 148 / (clk * 1e3), instr / (ms * 1e-3) / 148 / (clk * 1e3));
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int blocks = 148 * 4, thr = 512;
  const double lanes = (double)blocks * thr * ITERS * CH;
  float ms;
  ms = timeit([&] { k_ffma<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
  printf("FFMA          : %.3f ms  %.1f lane-FMA/clk/SM @%d kHz -> %.1f TFLOP/s\n", ms, lanes / (ms * 1e-3) / 148 / (clk * 1e3), clk, 2 * lanes / ms * 1e-9);
  ms = timeit([&] { k_ffma2_bcast<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2 bcast   : %.3f ms  %.1f instr-lanes/clk/SM -> %.1f TFLOP/s\n", ms, lanes / (ms * 1e-3) / 148 / (clk * 1e3), 4 * lanes / ms * 1e-9);
  ms = timeit([&] { k_ffma2_packed<<<blocks, thr>>>(out, 1.0001f, 0.5f); });
  printf("FFMA2 packed  : %.3f ms  %.1f instr-lanes/clk/SM -> %.1f TFLOP/s\n", ms, lanes / (ms * 1e-3) / 148 / (clk * 1e3), 4 * lanes / ms * 1e-9);
  run_mix<16, 0>(out, clk, 148 * 4, 512);
  run_mix<0, 16>(out, clk, 148 * 4, 512);
  run_mix<8, 8>(out, clk, 148 * 4, 512);
  run_mix<12, 6>(out, clk, 148 * 4, 512);
  run_mix<8, 16>(out, clk, 148 * 4, 512);
  run_mix<16, 8>(out, clk, 148 * 4, 512);
  // the window kernels' occupancy: 12 warps per SM
  run_mix<16, 0>(out, clk, 148 * 3, 128);
  run_mix<32, 0>(out, clk, 148 * 3, 128);
  run_mix<16, 8>(out, clk, 148 * 3, 128);
  run_mix<24, 12>(out, clk, 148 * 3, 128);
  cudaFuncSetAttribute(k_lds<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_lds<float4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_lds<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const double loads = (double)148 * 2 * 512 * ITERS;
  for (int shift = 0; shift < 4; ++shift) {
    ms = timeit([&] { k_lds<float><<<148 * 2, 512, 5000 * 4>>>(out, shift, 1); });
    printf("LDS.32  shift %d: %.3f ms  %.1f B/clk/SM\n", shift, ms, loads * 4 / (ms * 1e-3) / 148 / (clk * 1e3));
    ms = timeit([&] { k_lds<float2><<<148 * 2, 512, 5000 * 8>>>(out, shift, 1); });
    printf("LDS.64  shift %d: %.3f ms  %.1f B/clk/SM\n", shift, ms, loads * 8 / (ms * 1e-3) / 148 / (clk * 1e3));
    ms = timeit([&] { k_lds<float4><<<148 * 2, 512, 5000 * 16>>>(out, shift, 1); });
    printf("LDS.128 shift %d: %.3f ms  %.1f B/clk/SM\n", shift, ms, loads * 16 / (ms * 1e-3) / 148 / (clk * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
