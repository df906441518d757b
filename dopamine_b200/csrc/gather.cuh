// Per-transition scalar outputs of sample_transition_batch
// (circular_replay_buffer.py:516-556, prioritized_replay_buffer.py:193-200), shared
// by the gather kernels (gather.cu) and by the sampling kernel (sample.cu), which
// can emit them itself so that the loss kernel does not have to wait for the frame
// copies.
#pragma once

#include "replay.cuh"

namespace b2r {

constexpr int kMaxRowCopies = 3 + B2R_MAX_EXTRAS;

struct RowCopy {
  const uint8_t *src;
  uint8_t *dst;
  int32_t row_bytes;
  int32_t at_next;  // 0: row i, 1: row (i + L) mod C
};

struct ScalarArgs {
  int64_t capacity;
  int32_t horizon;
  const uint8_t *term_flag;
  const void *reward;   // f32 or f64 column
  int32_t reward_itemsize;
  const float *discounts;
  void *ret;            // n-step return, reward dtype
  uint8_t *terminal_out;
  int32_t terminal_itemsize;
  int32_t *indices_out;
  int32_t n_copies;
  RowCopy copies[kMaxRowCopies];
  const double *leaves;  // tree leaf level (nullable)
  float *prio_out;
  // Fast form (fast != 0): update_horizon <= 4, f32 rewards, 1-byte terminals,
  // 4-byte action rows, no extras — the Atari configuration.  Every load of a row
  // is then issued in one batch (see ScalarLoads).
  int32_t fast;
  const uint32_t *action_col;
  uint32_t *action_out, *next_action_out;  // nullable
  float *next_reward_out;                  // nullable
};

constexpr int kFastHorizon = 4;

// Everything a row's scalar outputs depend on, loaded before any of it is used:
// one memory round trip instead of the dependent chain index -> terminals ->
// trajectory length -> rewards / next row.
struct ScalarLoads {
  uint8_t term[kFastHorizon];
  float reward[kFastHorizon + 1];
  uint32_t action[kFastHorizon + 1];
  float disc[kFastHorizon];
  double leaf;
};

__device__ __forceinline__ void load_scalars(const ScalarArgs &a, int64_t i,
                                             ScalarLoads *r) {
  if (i >= a.capacity) i = 0;  // padded leaf: invalid pick, the values are unused
  const float *reward = static_cast<const float *>(a.reward);
#pragma unroll
  for (int k = 0; k <= kFastHorizon; ++k) {
    int64_t s = i + k;
    if (s >= a.capacity) s -= a.capacity;
    if (k < kFastHorizon) {
      r->term[k] = k < a.horizon ? a.term_flag[s] : (uint8_t)0;
      r->disc[k] = k < a.horizon ? a.discounts[k] : 0.f;
    }
    r->reward[k] = k <= a.horizon ? reward[s] : 0.f;
    r->action[k] = k <= a.horizon ? a.action_col[s] : 0u;
  }
  r->leaf = a.leaves ? a.leaves[i] : 0.0;
}

// Writes row b from the loaded values; returns f32(leaf) (+inf without a tree).
__device__ __forceinline__ float finish_scalars(const ScalarArgs &a, int b, int64_t i,
                                                const ScalarLoads &r,
                                                int *length_out = nullptr) {
  int length = a.horizon;
  bool ends = false;
#pragma unroll
  for (int k = kFastHorizon - 1; k >= 0; --k)
    if (k < a.horizon && r.term[k]) { length = k + 1; ends = true; }
  // np.sum(discount[:L] * reward[i:i+L]): +0.0f, then left to right (L < 8)
  float acc = 0.f;
  float next_reward = r.reward[0];
  uint32_t next_action = r.action[0];
#pragma unroll
  for (int k = 0; k < kFastHorizon; ++k) {
    if (k < length) {
      acc = __fadd_rn(acc, __fmul_rn(r.disc[k], r.reward[k]));
      next_reward = r.reward[k + 1];
      next_action = r.action[k + 1];
    }
  }
  const float prio = a.leaves ? (float)r.leaf : INFINITY;
  if (length_out) *length_out = length;
  if (a.ret) static_cast<float *>(a.ret)[b] = acc;
  if (a.prio_out) a.prio_out[b] = prio;
  if (a.indices_out) a.indices_out[b] = (int32_t)i;
  if (a.terminal_out) a.terminal_out[b] = ends ? 1 : 0;
  if (a.action_out) a.action_out[b] = r.action[0];
  if (a.next_action_out) a.next_action_out[b] = next_action;
  if (a.next_reward_out) a.next_reward_out[b] = next_reward;
  return prio;
}

// Trajectory length and terminal flag (circular_replay_buffer.py:517-527).  The
// flags of up to 8 steps are loaded together (one round trip) before any is tested.
__device__ __forceinline__ int trajectory_length(const uint8_t *__restrict__ term,
                                                 int64_t i, int horizon,
                                                 int64_t cap, bool *ends) {
  for (int base = 0; base < horizon; base += 8) {
    unsigned flags = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (base + k < horizon) {
        int64_t s = i + base + k;
        if (s >= cap) s -= cap;
        flags |= (term[s] ? 1u : 0u) << k;
      }
    }
    if (flags) {
      *ends = true;
      return base + __ffs(flags);
    }
  }
  *ends = false;
  return horizon;
}

// np.sum(discount[:L] * reward[i:i+L]) in numpy's evaluation order (probed on
// numpy 2.3.5; DESIGN.md "n-step return"): L < 8: +0.0f then left to right;
// 8 <= L <= 128: 8-lane unrolled block, pairwise combine, sequential tail.
template <typename R>
__device__ __forceinline__ R mul_rn(float d, R r);
template <>
__device__ __forceinline__ float mul_rn<float>(float d, float r) { return __fmul_rn(d, r); }
template <>
__device__ __forceinline__ double mul_rn<double>(float d, double r) { return __dmul_rn((double)d, r); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename R>
__device__ __forceinline__ R nstep_return(const R *__restrict__ reward,
                          const float *__restrict__ disc, int64_t i, int length,
                          int64_t cap) {
  auto term = [&](int k) {
    int64_t s = i + k;
    if (s >= cap) s -= cap;
    return mul_rn<R>(disc[k], reward[s]);
  };
  if (length < 8) {
    R acc = (R)0;
#pragma unroll 1
    for (int k = 0; k < length; ++k) acc = add_rn(acc, term(k));
    return acc;
  }
  R r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int k = 8;
#pragma unroll 1
  for (; k < length - (length % 8); k += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = add_rn(r[j], term(k + j));
  }
  R acc = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])),
                 add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
  for (; k < length; ++k) acc = add_rn(acc, term(k));
  return acc;
}

// All scalar outputs of batch row `b` for sampled index `i`.  Returns f32(leaf)
// (the row's sampling_probability; +inf when there is no tree).
static __device__ __noinline__ float write_scalars(const ScalarArgs &a, int b, int64_t i) {
  bool ends;
  const int length = trajectory_length(a.term_flag, i, a.horizon, a.capacity, &ends);
  int64_t nxt = i + length;
  if (nxt >= a.capacity) nxt -= a.capacity;
  float prio = INFINITY;
  if (a.leaves) prio = (float)a.leaves[i];  // PRB:231-235
  if (a.ret) {
    if (a.reward_itemsize == 4)
      static_cast<float *>(a.ret)[b] = nstep_return<float>(
          static_cast<const float *>(a.reward), a.discounts, i, length, a.capacity);
    else
      static_cast<double *>(a.ret)[b] = nstep_return<double>(
          static_cast<const double *>(a.reward), a.discounts, i, length, a.capacity);
  }
  if (a.prio_out) a.prio_out[b] = prio;
  if (a.indices_out) a.indices_out[b] = (int32_t)i;
  if (a.terminal_out) {
    uint8_t *t = a.terminal_out + (int64_t)b * a.terminal_itemsize;
    t[0] = ends ? 1 : 0;
    for (int k = 1; k < a.terminal_itemsize; ++k) t[k] = 0;
  }
#pragma unroll 1
  for (int c = 0; c < a.n_copies; ++c) {
    const RowCopy rc = a.copies[c];
    const uint8_t *s = rc.src + (rc.at_next ? nxt : i) * (int64_t)rc.row_bytes;
    uint8_t *d = rc.dst + (int64_t)b * rc.row_bytes;
    if ((rc.row_bytes & 3) == 0) {  // word rows (int32 actions, f32 rewards, ...)
      for (int k = 0; k < rc.row_bytes; k += 4)
        *reinterpret_cast<uint32_t *>(d + k) = *reinterpret_cast<const uint32_t *>(s + k);
    } else {
      for (int k = 0; k < rc.row_bytes; ++k) d[k] = s[k];
    }
  }
  return prio;
}

// Fills the scalar part of a gather from the buffer and the caller's outputs.
void fill_scalar_args(const b2r_buffer *buf, const b2r_batch *out, ScalarArgs *sc);

}  // namespace b2r
