// Actor-side frame stack: replaces DQNAgent._record_observation and _reset_state
// (dopamine/agents/dqn/dqn_agent.py:444-458, 474-476):
//   self.state = np.roll(self.state, -1, axis=-1); self.state[0, ..., -1] = observation
// The agent's state is a (1, H, W, S) tensor with the stack index innermost, so the
// roll moves every pixel's S elements one place down and the new frame lands in the
// last place.  Here the state lives in HBM (it is the network's input) and ONE
// launch per environment step does the roll and the insert: the frame is read
// straight from a pinned host slot (zero-copy over PCIe / C2C; 7 056 B for Atari),
// so the step needs neither the host-side np.roll of 28 KB nor a separate H2D copy.
//
// Traffic per step (Atari): 28 224 B read + 28 224 B written in HBM, 7 056 B from the
// host.  Latency-bound (one small launch); the point is what it removes from the host.
#include "common.cuh"

#include <cuda_fp16.h>

#include <cstring>
#include <new>

namespace b2r {
namespace {

constexpr int kMaxSlots = 16;

// S * elem_size == 4: a pixel's stack is one 32-bit word (uint8 x 4: Atari).
__global__ void __launch_bounds__(256)
record_u8x4_kernel(uint32_t *__restrict__ state, const uint8_t *__restrict__ frame,
                   int64_t pixels) {
  // 4 pixels per thread: one 16-byte word of state, 4 bytes of frame
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p = g * 4;
  if (p + 3 < pixels) {
    uint4 w = reinterpret_cast<uint4 *>(state)[g];
    const uint32_t f = reinterpret_cast<const uint32_t *>(frame)[g];
    w.x = (w.x >> 8) | ((f & 0xffu) << 24);
    w.y = (w.y >> 8) | (((f >> 8) & 0xffu) << 24);
    w.z = (w.z >> 8) | (((f >> 16) & 0xffu) << 24);
    w.w = (w.w >> 8) | ((f >> 24) << 24);
    reinterpret_cast<uint4 *>(state)[g] = w;
  } else {
    for (int64_t q = p; q < pixels; ++q)
      state[q] = (state[q] >> 8) | ((uint32_t)frame[q] << 24);
  }
}

// Any element size / stack size: one thread per pixel moves its S elements in
// ascending order (each source byte is read before its place is overwritten).
__global__ void __launch_bounds__(256)
record_generic_kernel(uint8_t *__restrict__ state, const uint8_t *__restrict__ frame,
                      int64_t pixels, int stack, int elem) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  uint8_t *row = state + p * stack * elem;
  const int keep = (stack - 1) * elem;
  for (int k = 0; k < keep; ++k) row[k] = row[k + elem];
  for (int k = 0; k < elem; ++k) row[keep + k] = frame[p * elem + k];
}

}  // namespace
}  // namespace b2r

struct b2r_actor {
  int64_t pixels = 0;
  int32_t elem = 1, stack = 4, slots = 0;
  size_t frame_bytes = 0;
  uint8_t *host[b2r::kMaxSlots] = {};      // pinned frame slots
  uint8_t *host_dev[b2r::kMaxSlots] = {};  // the same slots as the device sees them
  uint8_t *dev = nullptr;                  // fallback when the slots are not mapped
  cudaEvent_t done[b2r::kMaxSlots] = {};
  bool pending[b2r::kMaxSlots] = {};
  int64_t recorded = 0;
};

using b2r::as_stream;
using b2r::fail;

namespace b2r {

// Network input: uint8 frame stacks (B, pixels, S) with the stack axis innermost — what
// the replay gather and the actor state hold, the layout the reference feeds its
// network — to (B, S, pixels) planes of float(u8) / 255 (atari_lib.py:124-125:
// tf.cast(state, tf.float32) then tf.div(net, 255.), true division), the layout cuDNN's
// first convolution reads.  One pass, 16-byte loads and stores: 4 B read and
// 4 x sizeof(T) B written per pixel, where the eager tensor ops (permute, to, div_) move
// about three times as much.  T = float (the reference's arithmetic, bit for bit) or
// __half (the f32 quotient rounded once to nearest: the input of a half-precision conv).
template <typename T>
struct Quad;
template <>
struct Quad<float> {
  using type = float4;
  static __device__ __forceinline__ float4 make(float a, float b, float c, float d) {
    return make_float4(a, b, c, d);
  }
};
template <>
struct Quad<__half> {
  using type = uint2;
  static __device__ __forceinline__ uint2 make(float a, float b, float c, float d) {
    const __half2 lo = __halves2half2(__float2half_rn(a), __float2half_rn(b));
    const __half2 hi = __halves2half2(__float2half_rn(c), __float2half_rn(d));
    uint2 o;
    o.x = *reinterpret_cast<const unsigned int *>(&lo);
    o.y = *reinterpret_cast<const unsigned int *>(&hi);
    return o;
  }
};

__device__ __forceinline__ float unit_byte(uint32_t word, int k) {
  return __fdiv_rn((float)((word >> (8 * k)) & 0xffu), 255.0f);
}

// S == 4, pixels % 4 == 0: thread = 4 pixels (one 16-byte load, four 4-element stores).
template <typename T>
__global__ void __launch_bounds__(256)
stack4_to_planes_kernel(const uint4 *__restrict__ in, T *__restrict__ out, int64_t quads,
                        int64_t pixels) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const uint4 v = __ldg(in + q);
  const int64_t per_image = pixels / 4;
  const int64_t image = q / per_image, p0 = (q - image * per_image) * 4;
  T *base = out + image * 4 * pixels + p0;
  using Q = Quad<T>;
#pragma unroll
  for (int k = 0; k < 4; ++k)  // plane k = stack position k of the 4 pixels
    *reinterpret_cast<typename Q::type *>(base + k * pixels) =
        Q::make(unit_byte(v.x, k), unit_byte(v.y, k), unit_byte(v.z, k), unit_byte(v.w, k));
}

template <typename T>
__global__ void stack_to_planes_generic_kernel(const uint8_t *__restrict__ in,
                                               T *__restrict__ out, int64_t total,
                                               int64_t pixels, int stack) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // output element
  if (e >= total) return;
  const int64_t image = e / (pixels * stack), r = e - image * pixels * stack;
  const int64_t k = r / pixels, p = r - k * pixels;
  const float v = __fdiv_rn((float)in[(image * pixels + p) * stack + k], 255.0f);
  out[e] = (T)v;
}

}  // namespace b2r

extern "C" {

int b2r_actor_create(int64_t pixels, int32_t elem_size, int32_t stack_size,
                     int32_t slots, b2r_actor **out) {
  if (out == nullptr) return fail(B2R_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (pixels <= 0 || elem_size <= 0 || stack_size <= 0)
    return fail(B2R_ERR_INVALID_ARGUMENT,
                "the actor state needs pixels, elem_size and stack_size > 0");
  if (slots <= 0) slots = 4;
  if (slots > b2r::kMaxSlots) slots = b2r::kMaxSlots;
  b2r_actor *a = new (std::nothrow) b2r_actor();
  if (a == nullptr) return fail(B2R_ERR_CUDA, "out of host memory");
  a->pixels = pixels;
  a->elem = elem_size;
  a->stack = stack_size;
  a->slots = slots;
  a->frame_bytes = (size_t)pixels * elem_size;
  bool mapped = true;
  for (int k = 0; k < slots; ++k) {
    if (cudaMallocHost(reinterpret_cast<void **>(&a->host[k]), a->frame_bytes) !=
            cudaSuccess ||
        cudaEventCreateWithFlags(&a->done[k], cudaEventDisableTiming) != cudaSuccess) {
      b2r_actor_destroy(a);
      return fail(B2R_ERR_CUDA, "cannot allocate the pinned frame slots");
    }
    void *as_device = nullptr;
    if (cudaHostGetDevicePointer(&as_device, a->host[k], 0) == cudaSuccess)
      a->host_dev[k] = static_cast<uint8_t *>(as_device);
    else
      mapped = false;
  }
  if (!mapped) {
    cudaGetLastError();
    if (cudaMalloc(reinterpret_cast<void **>(&a->dev), a->frame_bytes) != cudaSuccess) {
      b2r_actor_destroy(a);
      return fail(B2R_ERR_CUDA, "cannot allocate the device frame");
    }
  }
  *out = a;
  return B2R_OK;
}

int b2r_actor_destroy(b2r_actor *a) {
  if (a == nullptr) return B2R_OK;
  for (int k = 0; k < b2r::kMaxSlots; ++k) {
    if (a->done[k]) {
      if (a->pending[k]) cudaEventSynchronize(a->done[k]);
      cudaEventDestroy(a->done[k]);
    }
    if (a->host[k]) cudaFreeHost(a->host[k]);
  }
  if (a->dev) cudaFree(a->dev);
  delete a;
  return B2R_OK;
}

int b2r_actor_reset(b2r_actor *a, void *state, b2r_stream stream) {
  if (a == nullptr || state == nullptr)
    return fail(B2R_ERR_INVALID_ARGUMENT, "actor or state is NULL");
  B2R_CUDA(cudaMemsetAsync(state, 0, a->frame_bytes * a->stack, as_stream(stream)));
  return B2R_OK;
}

int b2r_actor_record(b2r_actor *a, void *state, const void *observation,
                     b2r_stream stream) {
  if (a == nullptr || state == nullptr || observation == nullptr)
    return fail(B2R_ERR_INVALID_ARGUMENT, "actor, state or observation is NULL");
  cudaStream_t s = as_stream(stream);
  const int k = (int)(a->recorded % a->slots);
  if (a->pending[k]) {  // the launch that read this slot `slots` steps ago
    B2R_CUDA(cudaEventSynchronize(a->done[k]));
    a->pending[k] = false;
  }
  std::memcpy(a->host[k], observation, a->frame_bytes);
  const uint8_t *frame = a->host_dev[k];
  if (a->dev != nullptr) {
    B2R_CUDA(cudaMemcpyAsync(a->dev, a->host[k], a->frame_bytes, cudaMemcpyHostToDevice, s));
    frame = a->dev;
  }
  if (a->elem * a->stack == 4 && a->elem == 1 &&
      (reinterpret_cast<uintptr_t>(state) & 15) == 0) {
    const int64_t groups = (a->pixels + 3) / 4;
    b2r::record_u8x4_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, s>>>(
        static_cast<uint32_t *>(state), frame, a->pixels);
  } else {
    b2r::record_generic_kernel<<<(unsigned)((a->pixels + 255) / 256), 256, 0, s>>>(
        static_cast<uint8_t *>(state), frame, a->pixels, a->stack, a->elem);
  }
  B2R_CUDA(cudaGetLastError());
  B2R_LAUNCHED();
  B2R_CUDA(cudaEventRecord(a->done[k], s));
  a->pending[k] = true;
  ++a->recorded;
  return B2R_OK;
}

int b2r_stack_to_planes_device(const void *stacks, void *planes, int64_t images,
                                int64_t pixels, int32_t stack_size, int32_t half_out,
                                b2r_stream stream) {
  if (!stacks || !planes || images < 0 || pixels <= 0 || stack_size <= 0)
    return fail(B2R_ERR_INVALID_ARGUMENT, "stack_to_planes: bad argument");
  if (images == 0) return B2R_OK;
  cudaStream_t s = as_stream(stream);
  const bool aligned = (reinterpret_cast<uintptr_t>(stacks) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(planes) & 15) == 0;
  if (stack_size == 4 && pixels % 4 == 0 && aligned) {
    const int64_t quads = images * pixels / 4;
    const unsigned grid = (unsigned)((quads + 255) / 256);
    if (half_out)
      b2r::stack4_to_planes_kernel<__half><<<grid, 256, 0, s>>>(
          static_cast<const uint4 *>(stacks), static_cast<__half *>(planes), quads, pixels);
    else
      b2r::stack4_to_planes_kernel<float><<<grid, 256, 0, s>>>(
          static_cast<const uint4 *>(stacks), static_cast<float *>(planes), quads, pixels);
  } else {
    const int64_t total = images * pixels * stack_size;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (half_out)
      b2r::stack_to_planes_generic_kernel<__half><<<grid, 256, 0, s>>>(
          static_cast<const uint8_t *>(stacks), static_cast<__half *>(planes), total, pixels,
          stack_size);
    else
      b2r::stack_to_planes_generic_kernel<float><<<grid, 256, 0, s>>>(
          static_cast<const uint8_t *>(stacks), static_cast<float *>(planes), total, pixels,
          stack_size);
  }
  B2R_CUDA(cudaGetLastError());
  B2R_LAUNCHED();
  return B2R_OK;
}

}  // extern "C"
