// Shared device/host helpers of libhermes_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "hermes_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhermes_b200 is written for sm_100a (B200) only"
#endif

namespace hb {

constexpr int kThreads = 256;  // CTA size of the streaming kernels
constexpr double kTwoPi = 6.283185307179586476925286766559;
constexpr double kInvTwoPi = 0.15915494309189533576888376337251;

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int require_device();
int device_sm_count();
int persistent_sm_count();  // device_sm_count() minus the SMs set aside by hb_reserve_sms()

// ---- per-kernel accounting (hb_profile_begin/end, hb_launch_counts) -------------------------------
enum KernelKind {
  KIND_SOS_COEF = 0,
  KIND_TDL_POLY = 1,
  KIND_TDL_DIRECT = 2,
  KIND_SOS_STATE = 3,
  KIND_CDL_RAYS = 4,
  KIND_CDL_PROPAGATE = 5,
  KIND_SPATIAL_GEMM = 6,
  KIND_STATS = 7,
  KIND_MISC = 8,
  KIND_COUNT = HB_NUM_KERNEL_KINDS
};

// RAII marker around ONE kernel launch on `stream`: always counts the launch; when profiling is enabled it
// also brackets the launch with CUDA events recorded on the launching stream.
class ProfileScope {
 public:
  ProfileScope(int kind, cudaStream_t stream);
  ~ProfileScope();

 private:
  int kind_;
  cudaStream_t stream_;
  cudaEvent_t start_;
};

#define HB_CUDA(call)                                   \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return hb::cuda_fail(_e, #call); \
  } while (0)

// ---- host-buffer pipeline shared by the *_host entry points ------------------------------------------
// kSlots device staging buffers, each with its own non-blocking stream: chunk c uses slot c % kSlots, so the H2D
// copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernels of chunk c.
constexpr int kSlots = 3;
struct HostPipe {
  cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
  void* buf[kSlots] = {nullptr, nullptr, nullptr};
  size_t cap[kSlots] = {0, 0, 0};
  int device = -1;
  std::mutex mu;
};
extern HostPipe g_pipe;
int pipe_prepare(size_t bytes_per_slot);  // call with g_pipe.mu held
void pipe_release();
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Launch-uniform delay tables, passed by value in kernel parameter (constant) space so that no
// host->device copy with lifetime concerns is needed for them.
struct DelayTable {
  int32_t num_taps;                       // L
  int32_t num_groups;                     // G: distinct integer delays
  int32_t tap_delay[HB_MAX_TAPS];         // d_l, ascending
  int32_t group_delay[HB_MAX_TAPS];       // distinct delays, ascending
  uint16_t group_start[HB_MAX_TAPS + 1];  // taps [group_start[g], group_start[g+1]) share group_delay[g]
};

// ---- complex helpers ----------------------------------------------------------------------------
template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

__device__ __forceinline__ float2 to_c32(float2 v) { return v; }
__device__ __forceinline__ float2 to_c32(double2 v) { return make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ double2 to_c64(float2 v) { return make_double2((double)v.x, (double)v.y); }
__device__ __forceinline__ double2 to_c64(double2 v) { return v; }

template <typename R> struct Conv;
template <> struct Conv<float> {
  template <typename V> static __device__ __forceinline__ float2 from(V v) { return to_c32(v); }
};
template <> struct Conv<double> {
  template <typename V> static __device__ __forceinline__ double2 from(V v) { return to_c64(v); }
};

template <typename IO> struct IoConv;
template <> struct IoConv<float2> {
  static __device__ __forceinline__ float2 make(float re, float im) { return make_float2(re, im); }
  static __device__ __forceinline__ float2 make(double re, double im) { return make_float2((float)re, (float)im); }
};
template <> struct IoConv<double2> {
  static __device__ __forceinline__ double2 make(float re, float im) { return make_double2((double)re, (double)im); }
  static __device__ __forceinline__ double2 make(double re, double im) { return make_double2(re, im); }
};

// acc += a * b (complex), 4 FMAs
template <typename R, typename C>
__device__ __forceinline__ void cmac(C& acc, const C a, const C b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// Streaming global accesses: the signal is touched exactly once, keep it out of L1.
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float2* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(double2* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

}  // namespace hb
