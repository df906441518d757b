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

// Completion hand-shake between the first half of the C51 loss (c51.cu: c51_pre_*,
// launched on a forked stream beside the sampler because it needs the network outputs
// only) and the sampler of the same step.  The loss tail must follow BOTH; a kernel node
// with two parents loses its programmatic early launch (measured: 2.4 us from the end of
// the sampler to the first instruction of the tail, against 0.5 us with one parent), so
// the dependency on the first half goes through memory instead of through the graph:
// its last CTA bumps `done` (release), the sampler's closing thread waits for that bump
// (acquire) before the kernel ends, and the tail has the sampler as its only parent.
// The first half is a few microseconds of work that starts together with the sampler,
// so the wait normally finds the bump already there.
struct PreSync {
  unsigned int *done;    // first-half launches completed (monotone); nullptr: no hand-shake
  unsigned int *seen;    // ... consumed by a sampler (touched by its closing thread only)
  unsigned int *ticket;  // CTAs finished in the running first-half launch
};

// Row-by-row hand-over from the sampling kernel to the kernels that consume its rows in
// the same step (the frame copies; see per_sample_warp_kernel and
// gather_stack4_u8_kernel).  Those kernels are programmatic dependents of the sampler:
// the hardware starts their CTAs once every sampler CTA is running, and instead of
// waiting for the sampler to END they wait for the row they need.  The sampler leaves
// one 8-byte descriptor per row, written with a single store:
//   [63:40] the step's tag (device draw counter + 1, folded to 24 bits, never 0)
//   [39:32] trajectory length L (circular_replay_buffer.py:517-527)
//   [31:0]  the sampled index
// so a consumer that sees the tag has everything it needs to address the frames: no
// acquire, no second load of the index, no terminal look-up.  The sampler's first CTA
// publishes the tag in tag_word before it lets the dependents start; final_word == tag
// says that the sampler is finished (rows that never got a descriptor — a latched
// failure — are then skipped).
struct RowFlags {
  uint64_t *desc;        // [rows]; nullptr: no hand-over (wait for the whole kernel)
  uint32_t *tag_word;
  uint32_t *final_word;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Last statement of a first-half CTA (every thread calls it).
__device__ __forceinline__ void pre_sync_signal(const PreSync &p) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(p.ticket, 1u) == gridDim.x - 1) {
    *p.ticket = 0u;  // ready for the next launch
    __threadfence();
    st_release_u32(p.done, *p.done + 1u);
  }
}
// Called by the sampler's closing thread (ONE thread) before the kernel ends.  Gives up
// after 2 s: a first half that never ran must not hang the GPU (the tail then reads stale
// rows; the launch error that caused it has been reported to the caller).
__device__ __forceinline__ void pre_sync_consume(const PreSync &p) {
  if (p.done == nullptr) return;
  const unsigned int seen = *p.seen;
  const long long t0 = clock64();
  while (ld_acquire_u32(p.done) == seen)
    if (clock64() - t0 > 4000000000ll) break;
  *p.seen = seen + 1u;
}
// The same in two steps, so that the two loads are in flight long before the closing
// thread needs their answer (the first half is normally done by then: no spin at all).
struct PreSyncPeek {
  unsigned int seen, done;
};
__device__ __forceinline__ PreSyncPeek pre_sync_peek(const PreSync &p) {
  PreSyncPeek k = {0u, 1u};
  if (p.done != nullptr) {
    k.seen = *p.seen;
    k.done = ld_acquire_u32(p.done);
  }
  return k;
}
__device__ __forceinline__ void pre_sync_consume(const PreSync &p, const PreSyncPeek &k) {
  if (p.done == nullptr) return;
  if (k.done == k.seen) {  // not yet when we looked: wait now
    const long long t0 = clock64();
    while (ld_acquire_u32(p.done) == k.seen)
      if (clock64() - t0 > 4000000000ll) break;
  }
  *p.seen = k.seen + 1u;
}
// 24-bit form of the step tag, never 0 (a descriptor that was never written is 0)
__device__ __forceinline__ uint32_t row_tag24(uint32_t tag) { return tag % 0xffffffu + 1u; }
__device__ __forceinline__ uint64_t row_descriptor(uint32_t tag, int length, int64_t idx) {
  return ((uint64_t)row_tag24(tag) << 40) | ((uint64_t)(length & 0xff) << 32) |
         (uint64_t)(uint32_t)idx;
}
// Called by ONE thread: the descriptor of row `row` of this step once it is there, 0 if
// the sampler finished without it (or after 2 s: a hung producer must not hang the GPU;
// *timed_out says which).
__device__ __forceinline__ uint64_t row_flags_wait(const RowFlags &f, int row,
                                                   bool *timed_out) {
  *timed_out = false;
  uint64_t d = ld_volatile_u64(f.desc + row);   // (both loads in flight together)
  const uint32_t tag = ld_volatile_u32(f.tag_word);
  const long long t0 = clock64();
  while (true) {
    if ((uint32_t)(d >> 40) == row_tag24(tag)) return d;
    if (ld_volatile_u32(f.final_word) == tag) {
      d = ld_volatile_u64(f.desc + row);
      return (uint32_t)(d >> 40) == row_tag24(tag) ? d : 0ull;
    }
    if (clock64() - t0 > 4000000000ll) {
      *timed_out = true;
      return 0ull;
    }
    d = ld_volatile_u64(f.desc + row);
  }
}
#endif

// Staged rows -> ring (replay.cu: flush_queue).  In a header because the fused flush
// kernel of tree.cu runs it beside the priority update.
struct AddParams {
  int n_entries;
  int num_columns;
  int64_t row_stride;
  const int64_t *slots;     // device, per entry
  const int32_t *src_rows;  // device, per entry; -1 = all-zero padding transition
  const uint8_t *rows;      // device staged rows
  uint8_t *col_dev[kMaxColumns];
  int64_t col_bytes[kMaxColumns];
  int64_t col_qoff[kMaxColumns];
  uint8_t *term_flag;       // nullptr when it aliases the 1-byte terminal column
  int term_itemsize;
  const uint64_t *ctx_src;  // ValidCtx image in the staged header
  uint64_t *ctx_dst;        // the buffer's device ValidCtx
};

// One CTA of the row writer: entry e, chunk-block bx of nbx.
#ifdef __CUDACC__
__device__ __forceinline__ void add_rows_body(const AddParams &p, int e, int bx, int nbx) {
  const int64_t slot = p.slots[e];
  const int src = p.src_rows[e];
  const uint8_t *row = src >= 0 ? p.rows + (int64_t)src * p.row_stride : nullptr;

  // observation: coalesced 16-byte stores when the frame size allows it.
  const int64_t obs_bytes = p.col_bytes[0];
  uint8_t *dst = p.col_dev[0] + slot * obs_bytes;
  const int64_t tid = bx * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = nbx * (int64_t)blockDim.x;
  if ((obs_bytes & 15) == 0) {
    const int64_t chunks = obs_bytes >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(row);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    for (int64_t c = tid; c < chunks; c += nthreads)
      d4[c] = row ? s4[c] : make_uint4(0, 0, 0, 0);
  } else {
    for (int64_t c = tid; c < obs_bytes; c += nthreads) dst[c] = row ? row[c] : 0;
  }

  // the validity context that goes with these rows
  if (bx == 0 && e == 0)
    for (int w = threadIdx.x; w < (int)(sizeof(ValidCtx) / 8); w += blockDim.x)
      p.ctx_dst[w] = p.ctx_src[w];

  // scalar columns: a handful of bytes, first block of the entry only.
  if (bx == 0) {
    for (int c = 1; c < p.num_columns; ++c) {
      const int64_t nb = p.col_bytes[c];
      uint8_t *d = p.col_dev[c] + slot * nb;
      const uint8_t *s = row ? row + p.col_qoff[c] : nullptr;
      for (int64_t b = threadIdx.x; b < nb; b += blockDim.x) d[b] = s ? s[b] : 0;
    }
    if (p.term_flag != nullptr && threadIdx.x == 0) {
      uint8_t any = 0;
      if (row)
        for (int b = 0; b < p.term_itemsize; ++b) any |= row[p.col_qoff[3] + b];
      p.term_flag[slot] = any ? 1 : 0;
    }
  }
}

#endif  // __CUDACC__

struct Column {
  int64_t row_bytes = 0;
  int64_t queue_offset = 0;  // byte offset inside a staged row
  uint8_t *dev = nullptr;
};

struct Staging {
  uint8_t *host = nullptr;  // pinned
  uint8_t *host_dev = nullptr;  // the same memory as the device sees it (zero-copy)
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
  // fused step: the frame copies run on `side`, forked/joined with these events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // ... and the write-back's grouping pass (needs the indices only) on `side2`, beside
  // the loss kernel
  cudaStream_t side2 = nullptr;
  cudaEvent_t ev_join2 = nullptr;
  cudaEvent_t ev_pre = nullptr, ev_h2d = nullptr, ev_rows = nullptr;  // split flush
  // ... and the half of the C51 loss that needs the network outputs only on `side3`,
  // beside the sampler (c51.cu: c51_pre_kernel)
  cudaStream_t side3 = nullptr;
  cudaEvent_t ev_c51_fork = nullptr, ev_c51_pre = nullptr;
  float *c51_bestp = nullptr;  // device [2 halves][c51_bestp_rows][64 probabilities | 64 stats]
  int64_t c51_bestp_rows = 0;
  int c51_half = 0;            // half the next step's first half writes
  unsigned int *pre_sync = nullptr;  // device [4]: PreSync done, seen, ticket
  // Deferred frame copies (b2r_set_deferred_frames): the copies of a fused step are not
  // joined into the caller's stream when the call returns but when b2r_join_frames is
  // called (or staged adds are flushed: they overwrite ring slots the copies may still
  // read), so the next step's sampler -> loss -> write-back chain runs BESIDE them.  The
  // copies read their indices from a private two-slot ring: the next sampler overwrites
  // the caller's `indices` while they are still running.
  bool deferred_frames = false;
  bool frames_pending = false;   // a copy was queued on `side` since the last join
  int frame_parity = 0;
  int32_t *idx_ring = nullptr;   // device [2][idx_ring_cap]
  int64_t idx_ring_cap = 0;
  // copies that read ring slot p have finished (a sampler may then overwrite it: the
  // chain runs at most one step ahead of the copies)
  cudaEvent_t ev_slot_free[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  float *min_prob = nullptr;    // device: min sampling probability of the last batch
  uint32_t *row_flags = nullptr;  // device: tag word, final word, desc[row] (RowFlags)
  int64_t row_flags_cap = 0;
  // Device copy of the validity context (add_count, cursor, invalid_range): the
  // prioritized sampler reads it from here, so a captured CUDA graph stays valid
  // while adds move the cursor.  Refreshed by every flush (the image rides in the
  // staging header) and lazily after b2r_set_state.
  b2r::ValidCtx *ctx_dev = nullptr;
  bool ctx_dirty = true;
  b2r::Bounce bounce;           // HOST-array calls
  uint8_t *out_scratch = nullptr;  // device outputs of b2r_gather (HOST variant)
  size_t out_scratch_cap = 0;
};

struct b2r_exchange {
  int world = 0, rank = 0;
  uint64_t *mailbox = nullptr;  // device, written by the peers
  uint64_t *seq = nullptr;      // device
  uint64_t *peer[b2r::kMaxShards] = {nullptr};
  bool opened[b2r::kMaxShards] = {false};  // peer[g] came from cudaIpcOpenMemHandle
  bool connected = false;
  int64_t timeout_ns = 2000000000ll;
  uint64_t *pub = nullptr;              // device [2]: ExchangeArgs::pub
  b2r::ExchangeArgs *args_dev = nullptr;  // device copy of the kernel arguments (the tree
                                          // kernels' publish hook reads it from there)
  // b2r_exchange_set_early_publish: the kernel that leaves the tree final for the next
  // sharded step publishes the shard total itself, and the staged adds of a step call are
  // flushed at the END of that call (they become visible to the next step's sampler).
  bool early_publish = false;
};

namespace b2r {
// Applies the staged adds.  split: the staged rows are copied and written to the
// ring on the buffer's side stream while the priorities go into the tree on
// `stream`; `stream` then waits for the rows, so callers see no difference.
// publish (nullable): the flush leaves the tree final for the next sharded step (see
// tree_apply).
int flush_queue(b2r_buffer *buf, cudaStream_t stream, bool split = false,
                const b2r_exchange *publish = nullptr);
// Makes `stream` wait for every deferred frame copy queued so far (no-op otherwise).
int join_frames(b2r_buffer *buf, cudaStream_t stream);
void fill_exchange_args(const b2r_exchange *x, ExchangeArgs *out);
int launch_exchange_publish(const b2r_exchange *x, const b2r_buffer *buf,
                            cudaStream_t stream);
// Sharded sampling.  Totals come from `shard_totals` (device array) or, when `x` is
// set, from the peer-memory exchange.  `scalars` / `min_prob_out` as in launch_sample;
// out_count (nullable, device) receives this rank's row count.  max_rows (> 0): rows
// the caller's outputs hold — a rank whose share of a Philox batch is larger serves the
// first max_rows and latches B2R_ERR_UNSUPPORTED (0: global_batch rows).
int launch_sample_sharded(b2r_buffer *buf, int32_t global_batch, int32_t num_shards,
                          int32_t rank, const double *shard_totals,
                          const b2r_exchange *x, const double *query01,
                          int32_t n_retry, const double *retry_u01, uint64_t seed,
                          uint64_t offset, int32_t *out_slots, int32_t *out_indices,
                          int32_t *out_count, cudaStream_t stream,
                          const b2r_batch *scalars = nullptr,
                          float *min_prob_out = nullptr, int32_t max_rows = 0,
                          const PreSync *pre = nullptr);
void fill_valid_ctx(const b2r_buffer *buf, ValidCtx *ctx);
int ensure_ctx(b2r_buffer *buf, cudaStream_t stream);
// tree.cu: largest batch the one-CTA tree kernel takes, and the fused flush launch
// (priorities of n new rows + their row writes, staging read zero-copy).
int tree_small_max();
int flush_fused(b2r_tree *tree, int n, const int64_t *slots, const double *prio,
                const uint8_t *mode, const AddParams &rows, int row_blocks_per_entry,
                cudaStream_t stream, const b2r_exchange *publish = nullptr);
int ensure_inv_slots(b2r_buffer *buf, int64_t n);
// Row flags for `rows` rows (RowFlags; all fields nullptr when the hand-over is off:
// B2R_ROW_FLAGS=0, the thread sampler).
int row_flags_for(b2r_buffer *buf, int64_t rows, RowFlags *out);
bool sampler_hands_over_rows(int strata);  // sample.cu: that sampler publishes RowFlags
bool gather_takes_row_flags(const b2r_buffer *buf);  // gather.cu
int launch_gather(b2r_buffer *buf, int32_t batch, const int32_t *indices_dev,
                  const b2r_batch *out, cudaStream_t stream,
                  const int32_t *count_dev = nullptr, bool frames_only = false,
                  const RowFlags *flags = nullptr);
int launch_sample(b2r_buffer *buf, int32_t batch, bool philox, uint64_t seed,
                  uint64_t offset, const double *strat_dev,
                  const double *retry_dev, int32_t n_retry, int32_t *out_idx_dev,
                  int32_t *info_dev, cudaStream_t stream,
                  const b2r_batch *scalars = nullptr, float *min_prob_out = nullptr,
                  const RowFlags *flags = nullptr, const PreSync *pre = nullptr);
}  // namespace b2r
