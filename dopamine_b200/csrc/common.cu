// Error plumbing and the pinned bounce buffer.
#include "common.cuh"

#include <cstdlib>

namespace b2r {

std::atomic<int64_t> g_launches{0};

bool pdl_enabled() {
  static const bool on = std::getenv("B2R_NO_PDL") == nullptr;
  return on;
}

int chain_priority() {
  static const int value = [] {
    if (std::getenv("B2R_NO_PRIORITY") != nullptr) return 0;
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return 0;
    return greatest;
  }();
  return value;
}

TreeWindow &tree_window() {
  static thread_local TreeWindow w;
  return w;
}

void set_tree_window(void *base, size_t bytes) {
  static const bool enabled = std::getenv("B2R_NO_L2_PERSIST") == nullptr;
  static size_t reserved = 0;
  TreeWindow &w = tree_window();
  if (!enabled || base == nullptr) {
    w.base = nullptr;
    w.bytes = 0;
    return;
  }
  int device = 0, max_persist = 0, max_window = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
  if (max_persist <= 0 || max_window <= 0) {
    w.base = nullptr;
    return;
  }
  if (bytes > (size_t)max_window) bytes = (size_t)max_window;
  const size_t want = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  if (want > reserved) {  // grow the set-aside (a device-wide limit) when needed
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
      reserved = want;
    else
      cudaGetLastError();
  }
  w.base = base;
  w.bytes = bytes;
}

std::string &last_error_slot() {
  static thread_local std::string slot;
  return slot;
}

int fail(int code, const char *fmt, ...) {
  char text[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(text, sizeof(text), fmt, ap);
  va_end(ap);
  last_error_slot() = text;
  return code;
}

int Bounce::mark_busy(cudaStream_t stream) {
  if (!busy) B2R_CUDA(cudaEventCreateWithFlags(&busy, cudaEventDisableTiming));
  B2R_CUDA(cudaEventRecord(busy, stream));
  pending = true;
  return B2R_OK;
}

int Bounce::reserve(size_t bytes) {
  if (pending) {
    B2R_CUDA(cudaEventSynchronize(busy));
    pending = false;
  }
  if (bytes <= cap) return B2R_OK;
  size_t want = cap ? cap : 4096;
  while (want < bytes) want *= 2;
  release();
  B2R_CUDA(cudaMallocHost(reinterpret_cast<void **>(&host), want));
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&dev), want));
  cap = want;
  return B2R_OK;
}

void Bounce::release() {
  if (pending && busy) cudaEventSynchronize(busy);
  pending = false;
  if (host) cudaFreeHost(host);
  if (dev) cudaFree(dev);
  host = dev = nullptr;
  cap = 0;
}

}  // namespace b2r

extern "C" {

const char *b2r_last_error(void) { return b2r::last_error_slot().c_str(); }

int b2r_abi_version(void) { return B2R_ABI_VERSION; }

int64_t b2r_launch_count(void) {
  return b2r::g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
