// GPU sum tree internals shared by tree.cu and replay.cu.
#pragma once

#include "common.cuh"

// Nodes live in ONE fp64 heap array in HBM: node (level l, position i) at
// 2^l - 1 + i, so level l is the contiguous slice [2^l - 1, 2^(l+1) - 1) and is
// exactly the reference's `nodes[l]` (sum_tree.py:79-87).  16 MB for capacity 1M:
// L2-resident on B200 (126 MB).
struct b2r_tree {
  int64_t capacity = 0;
  int depth = 0;          // levels = depth + 1
  int64_t leaves = 0;     // 2^depth
  double *heap = nullptr;
  double *max_rec = nullptr;   // device scalar: max_recorded_priority
  int64_t *status = nullptr;   // device [2]: latched error code, offending position
  double *delta = nullptr;     // device scratch: per-element leaf deltas of a chunk
  int32_t *n_eff = nullptr;    // device scalar: elements of the chunk to apply
  b2r::Bounce bounce;
};

namespace b2r {

constexpr int kTreeChunk = 4096;   // elements sorted per CTA (64 KB of smem)
constexpr int kTopLevels = 10;     // levels 0..10 (2047 nodes, 16 KB) staged in smem

// sum_tree.py:126-141 — strict `<` against the stored left child, subtract when
// going right.  __dsub_rn pins the rounding (no contraction).
__device__ __forceinline__ int64_t tree_descend(const double *__restrict__ heap,
                                                int depth, double q) {
  int64_t node = 0;
  for (int l = 1; l <= depth; ++l) {
    const double left = heap[(((int64_t)1) << l) - 1 + 2 * node];
    if (q < left) {
      node = 2 * node;
    } else {
      node = 2 * node + 1;
      q = __dsub_rn(q, left);
    }
  }
  return node;
}

// Same descent, with levels 0..top_depth read from a shared-memory copy.
__device__ __forceinline__ int64_t tree_descend_staged(
    const double *__restrict__ heap, const double *top, int top_depth, int depth,
    double q) {
  int64_t node = 0;
  int l = 1;
  for (; l <= top_depth; ++l) {
    const double left = top[(1 << l) - 1 + 2 * (int)node];
    if (q < left) {
      node = 2 * node;
    } else {
      node = 2 * node + 1;
      q = __dsub_rn(q, left);
    }
  }
  for (; l <= depth; ++l) {
    const double left = heap[(((int64_t)1) << l) - 1 + 2 * node];
    if (q < left) {
      node = 2 * node;
    } else {
      node = 2 * node + 1;
      q = __dsub_rn(q, left);
    }
  }
  return node;
}

// Cooperative copy of levels 0..min(depth, kTopLevels) into shared memory.
__device__ __forceinline__ int stage_top_levels(const double *__restrict__ heap,
                                                int depth, double *top) {
  const int top_depth = depth < kTopLevels ? depth : kTopLevels;
  const int count = (1 << (top_depth + 1)) - 1;
  for (int i = threadIdx.x; i < count; i += blockDim.x) top[i] = heap[i];
  __syncthreads();
  return top_depth;
}

// Applies n sets in array order (device arrays).  mode (nullable): 1 = use the
// running max_recorded_priority instead of values[k].
template <typename I, typename V>
int tree_apply(b2r_tree *t, int64_t n, const I *indices, const V *values,
               const uint8_t *mode, cudaStream_t stream);

}  // namespace b2r
