// Replay buffer internals shared by replay.cu / sample.cu / gather.cu.
#pragma once

#include <vector>

#include "common.cuh"
#include "tree.cuh"

namespace b2r {

constexpr int kMaxColumns = 4 + B2R_MAX_EXTRAS;
constexpr int kMaxInvalid = 64;  // stack_size + update_horizon entries

// What is_valid_transition (circular_replay_buffer.py:381-414) needs, by value in
// kernel parameters: the host owns add_count / invalid_range (adds are host-driven).
struct ValidCtx {
  int64_t capacity;
  int64_t add_count;
  int64_t cursor;
  int32_t stack;
  int32_t horizon;
  int32_t n_invalid;
  int64_t invalid[kMaxInvalid];
  const uint8_t *term_flag;  // 1 byte per slot: terminal != 0
};

__device__ __forceinline__ bool is_valid_transition(const ValidCtx &c,
                                                    int64_t index) {
  if (index < 0 || index >= c.capacity) return false;
  bool ok = true;
  if (c.add_count < c.capacity)  // not full
    ok = index < c.cursor - c.horizon && index >= c.stack - 1;
  for (int k = 0; k < c.n_invalid; ++k) ok = ok && c.invalid[k] != index;
  // get_terminal_stack(index)[:-1].any(): all flags are loaded before any is
  // tested, so the check costs one memory round trip.
  unsigned any = 0;
#pragma unroll 4
  for (int k = 1; k < c.stack; ++k) {
    int64_t s = index - k;
    if (s < 0) s += c.capacity;
    any |= c.term_flag[s];
  }
  return ok && any == 0;
}

// Block-wide exclusive scan of a 0/1 flag (warp ballots + one shared array).
// Every thread of the block must call it.  Returns this thread's exclusive
// prefix; *total receives the block total.
__device__ __forceinline__ int block_scan_flag(bool flag, int *warp_counts,
                                               int *total) {
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int within = __popc(ballot & ((1u << lane) - 1u));
  __syncthreads();  // protect warp_counts from the previous call's readers
  if (lane == 0) warp_counts[warp] = __popc(ballot);
  __syncthreads();
  int before = 0, all = 0;
  const int warps = (blockDim.x + 31) >> 5;
  for (int w = 0; w < warps; ++w) {
    const int c = warp_counts[w];
    if (w < warp) before += c;
    all += c;
  }
  *total = all;
  return before + within;
}

struct Column {
  int64_t row_bytes = 0;
  int64_t queue_offset = 0;  // byte offset inside a staged row
  uint8_t *dev = nullptr;
};

struct Staging {
  uint8_t *host = nullptr;  // pinned
  uint8_t *dev = nullptr;
  cudaEvent_t done = nullptr;
  bool in_flight = false;
};

}  // namespace b2r

struct b2r_buffer {
  b2r_config cfg;
  int num_columns = 0;
  b2r::Column col[b2r::kMaxColumns];
  uint8_t *term_flag = nullptr;  // aliases the terminal column when it is 1 byte
  bool term_flag_owned = false;
  float *discounts = nullptr;    // f32(pow(gamma, k)), k < update_horizon
  std::vector<float> discounts_host;
  b2r_tree *tree = nullptr;

  // host bookkeeping (circular_replay_buffer.py:177-178, 284-287)
  int64_t add_count = 0;
  std::vector<int64_t> invalid_range;
  std::vector<uint8_t> term_is_one;  // host mirror of `terminal == 1` (CRB:255)

  // deferred-add queue: rows staged in pinned memory, applied by one kernel
  int queue_cap = 0;        // entries (real rows + zero pads)
  int64_t row_stride = 0;   // bytes of one staged row (16-byte aligned)
  int64_t header_bytes = 0; // SoA entry table in front of the rows
  b2r::Staging staging[2];
  int active = 0;
  int q_entries = 0, q_rows = 0;

  // scratch
  int32_t *inv_slots = nullptr;
  int64_t inv_slots_cap = 0;
  int32_t *info = nullptr;      // device [4]: status, fail slot, draws used, count
  int64_t *status = nullptr;    // device [2]: latched asynchronous error
  unsigned int *ticket = nullptr;   // device: last-CTA election of the sample kernel
  uint64_t *shard_counter = nullptr;  // device: like draw_counter, for sharded draws
                                      // (advances in lockstep on every rank)
  uint64_t *draw_counter = nullptr;  // device: bumps per Philox sample launch, so a
                                     // replayed CUDA graph draws fresh uniforms
  b2r::Bounce bounce;           // HOST-array calls
  uint8_t *out_scratch = nullptr;  // device outputs of b2r_gather (HOST variant)
  size_t out_scratch_cap = 0;
};

namespace b2r {
int flush_queue(b2r_buffer *buf, cudaStream_t stream);
void fill_valid_ctx(const b2r_buffer *buf, ValidCtx *ctx);
int ensure_inv_slots(b2r_buffer *buf, int64_t n);
int launch_gather(b2r_buffer *buf, int32_t batch, const int32_t *indices_dev,
                  const b2r_batch *out, cudaStream_t stream,
                  const int32_t *count_dev = nullptr);
int launch_sample(b2r_buffer *buf, int32_t batch, bool philox, uint64_t seed,
                  uint64_t offset, const double *strat_dev,
                  const double *retry_dev, int32_t n_retry, int32_t *out_idx_dev,
                  int32_t *info_dev, cudaStream_t stream);
}  // namespace b2r
