// Library-wide runtime: error strings, device checks, per-kernel launch accounting.
#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "hb_common.cuh"

namespace hb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  cudaGetLastError();
  return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? HB_ERR_NO_DEVICE : HB_ERR_CUDA;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (%s); libhermes_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return HB_ERR_NO_DEVICE;
  }
  // Workspaces come from the stream-ordered pool (cudaMallocAsync).  Its default release threshold is 0: every
  // synchronization hands the memory back to the driver and the next call pays a fresh ~1 ms allocation.  Keep it.
  static bool pool_kept[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !pool_kept[dev]) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    pool_kept[dev] = true;
  }
  return HB_OK;
}

static std::atomic<int> g_reserved_sms{0};

int persistent_sm_count() { return std::max(1, device_sm_count() - g_reserved_sms.load(std::memory_order_relaxed)); }

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- host pipeline -----------------------------------------------------------------------------------
HostPipe g_pipe;

void pipe_release() {
  for (int s = 0; s < kSlots; ++s) {
    if (g_pipe.buf[s]) cudaFree(g_pipe.buf[s]);
    if (g_pipe.st[s]) cudaStreamDestroy(g_pipe.st[s]);
    g_pipe.buf[s] = nullptr;
    g_pipe.cap[s] = 0;
    g_pipe.st[s] = nullptr;
  }
  g_pipe.device = -1;
}

int pipe_prepare(size_t bytes) {
  int dev = 0;
  HB_CUDA(cudaGetDevice(&dev));
  if (g_pipe.device != dev) {
    pipe_release();
    g_pipe.device = dev;
  }
  for (int s = 0; s < kSlots; ++s) {
    if (!g_pipe.st[s]) HB_CUDA(cudaStreamCreateWithFlags(&g_pipe.st[s], cudaStreamNonBlocking));
    if (g_pipe.cap[s] < bytes) {
      if (g_pipe.buf[s]) HB_CUDA(cudaFree(g_pipe.buf[s]));
      g_pipe.buf[s] = nullptr;
      g_pipe.cap[s] = 0;
      HB_CUDA(cudaMalloc(&g_pipe.buf[s], bytes));
      g_pipe.cap[s] = bytes;
    }
  }
  return HB_OK;
}

// ---- accounting -----------------------------------------------------------------------------------
static std::atomic<long long> g_counts[KIND_COUNT];
static std::atomic<bool> g_profiling{false};
struct Bracket {
  int kind;
  cudaEvent_t start, stop;
};
static std::mutex g_prof_mu;
static std::vector<Bracket> g_brackets;

ProfileScope::ProfileScope(int kind, cudaStream_t stream) : kind_(kind), stream_(stream), start_(nullptr) {
  g_counts[kind_].fetch_add(1, std::memory_order_relaxed);
  if (g_profiling.load(std::memory_order_relaxed)) {
    if (cudaEventCreate(&start_) == cudaSuccess) {
      cudaEventRecord(start_, stream_);
    } else {
      start_ = nullptr;
      cudaGetLastError();
    }
  }
}

ProfileScope::~ProfileScope() {
  if (!start_) return;
  cudaEvent_t stop = nullptr;
  if (cudaEventCreate(&stop) != cudaSuccess) {
    cudaGetLastError();
    cudaEventDestroy(start_);
    return;
  }
  cudaEventRecord(stop, stream_);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_brackets.push_back({kind_, start_, stop});
}

}  // namespace hb

using namespace hb;

extern "C" {

int hb_version(void) { return HB_VERSION; }

const char* hb_last_error(void) { return g_err; }

int hb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int hb_set_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible; libhermes_b200 has no CPU fallback");
    return HB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("device %d outside [0, %d)", device, n);
    return HB_ERR_INVALID;
  }
  HB_CUDA(cudaSetDevice(device));
  return HB_OK;
}

int hb_reserve_sms(int num_sms) {
  const int n = num_sms < 0 ? 0 : num_sms;
  return g_reserved_sms.exchange(n);
}

void hb_release(void) {
  std::lock_guard<std::mutex> lock(g_pipe.mu);
  pipe_release();
}

void hb_launch_counts(int64_t* counts) {
  if (!counts) return;
  for (int k = 0; k < KIND_COUNT; ++k) counts[k] = (int64_t)g_counts[k].load();
}

int hb_profile_begin(void) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& b : g_brackets) {
    cudaEventDestroy(b.start);
    cudaEventDestroy(b.stop);
  }
  g_brackets.clear();
  g_profiling.store(true);
  return HB_OK;
}

int hb_profile_end(hb_profile_report* report) {
  g_profiling.store(false);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  int rc = HB_OK;
  if (report) memset(report, 0, sizeof(*report));
  for (auto& b : g_brackets) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(b.stop);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, b.start, b.stop);
    if (e != cudaSuccess && rc == HB_OK) rc = cuda_fail(e, "profile event");
    if (report && e == cudaSuccess) {
      report->ms[b.kind] += (double)ms;
      report->launches[b.kind] += 1;
    }
    cudaEventDestroy(b.start);
    cudaEventDestroy(b.stop);
  }
  g_brackets.clear();
  return rc;
}

}  // extern "C"
