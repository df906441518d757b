// GPU sum tree internals shared by tree.cu and replay.cu.
#pragma once

#include "common.cuh"

// Nodes live in ONE fp64 heap array in HBM, 1-based: node (level l, position i)
// at h = 2^l + i (root at 1, element 0 unused).  Level l is the contiguous slice
// [2^l, 2^(l+1)) and is exactly the reference's `nodes[l]` (sum_tree.py:79-87);
// the children of h are 2h and 2h+1, and the 4 grandchildren / 8
// great-grandchildren of a node start on 32 B / 64 B boundaries, which the
// speculative descent below relies on.  16 MB for capacity 1M: L2-resident.
struct b2r_tree {
  int64_t capacity = 0;
  int depth = 0;          // levels = depth + 1
  int64_t leaves = 0;     // 2^depth; leaf i sits at heap[leaves + i]
  double *heap = nullptr;
  double *max_rec = nullptr;   // device scalar: max_recorded_priority
  int64_t *status = nullptr;   // device [2]: latched error code, offending position
  double *delta = nullptr;     // device scratch: per-element leaf deltas of a chunk
  b2r::Bounce bounce;
};

namespace b2r {

constexpr int kTreeChunk = 4096;   // elements sorted per CTA (64 KB of smem)
constexpr int kTopLevels = 10;     // levels 0..10 (heap[1..2048), 16 KB) staged in smem

// One step of sum_tree.py:128-139: strict `<` against the stored left child,
// subtract when going right.  __dsub_rn pins the rounding (no contraction).
__device__ __forceinline__ void descend_step(int64_t &h, double &q, double left) {
  if (q < left) {
    h = 2 * h;
  } else {
    h = 2 * h + 1;
    q = __dsub_rn(q, left);
  }
}

// Root-to-leaf descent (sum_tree.py:126-141).  `top` is a shared-memory copy of
// heap[0 .. 2^(top_depth+1)).  Below it, the dependent chain of loads is cut by 3:
// the left-child values of the next three levels (1 + 2 + 4 candidates, three
// sectors thanks to the 1-based layout) are fetched in ONE round trip and the
// three decisions are then taken exactly as the reference takes them.
__device__ __forceinline__ int64_t tree_descend_staged(
    const double *__restrict__ heap, const double *top, int top_depth, int depth,
    double q) {
  int64_t h = 1;
  int level = 0;
  for (; level < top_depth; ++level) descend_step(h, q, top[2 * h]);
  while (depth - level >= 3) {
    const double c1 = heap[2 * h];
    const double c2a = heap[4 * h], c2b = heap[4 * h + 2];
    const double c3a = heap[8 * h], c3b = heap[8 * h + 2];
    const double c3c = heap[8 * h + 4], c3d = heap[8 * h + 6];
    const int64_t h0 = h;
    descend_step(h, q, c1);
    descend_step(h, q, (h == 2 * h0) ? c2a : c2b);
    const int which = (int)(h - 4 * h0);
    descend_step(h, q, which == 0 ? c3a : which == 1 ? c3b : which == 2 ? c3c : c3d);
    level += 3;
  }
  for (; level < depth; ++level) descend_step(h, q, heap[2 * h]);
  return h - (((int64_t)1) << depth);
}

// Cooperative copy of heap[0 .. 2^(min(depth, kTopLevels)+1)) into shared memory.
__device__ __forceinline__ int stage_top_levels(const double *__restrict__ heap,
                                                int depth, double *top) {
  const int top_depth = depth < kTopLevels ? depth : kTopLevels;
  const int count = 2 << top_depth;
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
