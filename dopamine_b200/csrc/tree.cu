// GPU sum tree: replaces dopamine/replay_memory/sum_tree.py (SumTree).
//
// Bit-exactness contract.  The reference's `set` (sum_tree.py:178-205) computes
// delta = value - leaf once and adds it to the node on every level, so internal
// nodes are history-dependent fp64 sums, not left+right.  A batch of sets
// (prioritized_replay_buffer.py:213-214) is a sequential loop, so for every node
// the deltas of the batch elements below it must be added IN BATCH ORDER.
//
//   kernel 1  tree_leaf_pass      one CTA: sorts (leaf, k) in shared memory,
//                                 resolves duplicate leaves as chains, emits
//                                 delta[k], updates leaves + max_recorded.
//   kernel 2  tree_internal_pass  one CTA per internal level, all levels
//                                 concurrently: sorts (node, k) and runs one
//                                 ordered fp64 add-chain per touched node.
//
// The critical path is the root's chain of n dependent DADDs; everything else
// overlaps with it.  Bandwidth is irrelevant here (n * depth * 16 bytes).
#include "tree.cuh"

#include <new>

namespace b2r {
namespace {

constexpr uint64_t kPadKey = ~0ull;

__device__ __forceinline__ void bitonic_sort(uint64_t *keys, int padded) {
  for (int k = 2; k <= padded; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < padded; t += blockDim.x) {
        const int partner = t ^ j;
        if (partner > t) {
          const uint64_t a = keys[t], b = keys[partner];
          const bool ascending = (t & k) == 0;
          if ((a > b) == ascending) {
            keys[t] = b;
            keys[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <typename I, typename V>
__global__ void __launch_bounds__(1024)
tree_leaf_pass(double *__restrict__ heap, int depth, int64_t leaves, int n,
               int padded, const I *__restrict__ indices,
               const V *__restrict__ values, const uint8_t *__restrict__ mode,
               int64_t k_base, double *__restrict__ delta_out,
               int32_t *__restrict__ n_eff_out, double *__restrict__ max_rec,
               int64_t *__restrict__ status) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
  double *vals = reinterpret_cast<double *>(smem_raw) + padded;
  __shared__ int s_stop;       // first position that must not be applied
  __shared__ int s_stop_code;
  __shared__ double s_max[32];

  if (status[0] != 0) {  // an earlier chunk failed: the sequence stopped there
    if (threadIdx.x == 0) *n_eff_out = 0;
    return;
  }
  if (threadIdx.x == 0) {
    s_stop = n;
    s_stop_code = 0;
  }
  __syncthreads();

  // 1. stage values; find the first element the reference would have raised on.
  if (mode != nullptr) {
    // add-path batches may ask for "current max_recorded_priority": needs the
    // running maximum in order.  These batches are tiny; one thread walks them.
    if (threadIdx.x == 0) {
      double running = *max_rec;
      for (int k = 0; k < n; ++k) {
        double v = mode[k] ? running : (double)values[k];
        const int64_t idx = (int64_t)indices[k];
        if (v < 0.0) { s_stop = k; s_stop_code = B2R_ERR_NEGATIVE_PRIORITY; break; }
        if (idx < 0 || idx >= leaves) { s_stop = k; s_stop_code = B2R_ERR_INDEX_RANGE; break; }
        if (v > running) running = v;
        vals[k] = v;
      }
    }
  } else {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const double v = (double)values[k];
      const int64_t idx = (int64_t)indices[k];
      vals[k] = v;
      if (v < 0.0 || idx < 0 || idx >= leaves) atomicMin(&s_stop, k);
    }
  }
  __syncthreads();
  const int n_eff = s_stop;
  if (mode == nullptr && n_eff < n && threadIdx.x == 0) {
    s_stop_code = (vals[n_eff] < 0.0) ? B2R_ERR_NEGATIVE_PRIORITY
                                      : B2R_ERR_INDEX_RANGE;
  }

  // 2. max_recorded_priority = max(value, current) over the applied prefix.
  double local_max = 0.0;
  for (int k = threadIdx.x; k < n_eff; k += blockDim.x)
    local_max = fmax(local_max, vals[k]);
  for (int off = 16; off > 0; off >>= 1)
    local_max = fmax(local_max, __shfl_xor_sync(0xffffffffu, local_max, off));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = local_max;

  // 3. keys = (leaf, k): sorting groups duplicates and keeps batch order inside.
  for (int k = threadIdx.x; k < padded; k += blockDim.x)
    keys[k] = k < n_eff ? (((uint64_t)(int64_t)indices[k]) << 32) | (uint32_t)k
                        : kPadKey;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = *max_rec;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w)
      if (s_max[w] > m) m = s_max[w];
    if (n_eff > 0) *max_rec = m;
    *n_eff_out = n_eff;
    if (n_eff < n) {
      status[0] = s_stop_code;
      status[1] = k_base + n_eff;
    }
  }
  bitonic_sort(keys, padded);

  // 4. one thread per distinct leaf walks its chain in batch order:
  //    delta = value - leaf; leaf += delta   (sum_tree.py:196-202, last level).
  const int64_t leaf_base = leaves - 1;
  for (int p = threadIdx.x; p < n_eff; p += blockDim.x) {
    const uint32_t node = (uint32_t)(keys[p] >> 32);
    if (p > 0 && (uint32_t)(keys[p - 1] >> 32) == node) continue;
    double leaf = heap[leaf_base + node];
    for (int q = p; q < n_eff && (uint32_t)(keys[q] >> 32) == node; ++q) {
      const uint32_t k = (uint32_t)keys[q];
      const double d = __dsub_rn(vals[k], leaf);
      leaf = __dadd_rn(leaf, d);
      vals[k] = d;
    }
    heap[leaf_base + node] = leaf;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_eff; k += blockDim.x) delta_out[k] = vals[k];
}

template <typename I>
__global__ void __launch_bounds__(1024)
tree_internal_pass(double *__restrict__ heap, int depth, int padded,
                   const I *__restrict__ indices,
                   const double *__restrict__ delta_in,
                   const int32_t *__restrict__ n_eff_in) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t *keys = reinterpret_cast<uint64_t *>(smem_raw);
  double *delta = reinterpret_cast<double *>(smem_raw) + padded;

  const int n_eff = *n_eff_in;
  if (n_eff == 0) return;
  const int level = blockIdx.x;  // 0 .. depth-1
  const int shift = depth - level;
  // Shrink the sort to the next power of two >= n_eff.
  int p2 = 32;
  while (p2 < n_eff) p2 <<= 1;
  if (p2 > padded) p2 = padded;
  for (int k = threadIdx.x; k < p2; k += blockDim.x) {
    if (k < n_eff) {
      keys[k] = (((uint64_t)((int64_t)indices[k] >> shift)) << 32) | (uint32_t)k;
      delta[k] = delta_in[k];
    } else {
      keys[k] = kPadKey;
    }
  }
  __syncthreads();
  bitonic_sort(keys, p2);

  const int64_t base = (((int64_t)1) << level) - 1;
  for (int p = threadIdx.x; p < n_eff; p += blockDim.x) {
    const uint32_t node = (uint32_t)(keys[p] >> 32);
    if (p > 0 && (uint32_t)(keys[p - 1] >> 32) == node) continue;
    double acc = heap[base + node];
    for (int q = p; q < n_eff && (uint32_t)(keys[q] >> 32) == node; ++q)
      acc = __dadd_rn(acc, delta[(uint32_t)keys[q]]);
    heap[base + node] = acc;
  }
}

__global__ void tree_get_kernel(const double *__restrict__ heap, int64_t leaves,
                                int64_t n, const int64_t *__restrict__ indices,
                                double *__restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t i = indices[k];
  out[k] = (i >= 0 && i < leaves) ? heap[leaves - 1 + i] : 0.0;
}

__global__ void tree_query_kernel(const double *__restrict__ heap, int depth,
                                  int64_t n, const double *__restrict__ query01,
                                  int64_t *__restrict__ out,
                                  double *__restrict__ total_out) {
  extern __shared__ double top[];
  const int top_depth = stage_top_levels(heap, depth, top);
  const double total = top[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) *total_out = total;
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  // sum_tree.py:123-124: query_value *= total
  out[k] = tree_descend_staged(heap, top, top_depth, depth,
                               __dmul_rn(query01[k], total));
}

__global__ void tree_set_scalar_kernel(double *dst, double v) { *dst = v; }

int padded_size(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

template <typename K>
int allow_big_smem(K kernel) {
  static bool done = false;  // per instantiation
  if (!done) {
    B2R_CUDA(cudaFuncSetAttribute(kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kTreeChunk * 16));
    done = true;
  }
  return B2R_OK;
}

}  // namespace

template <typename I, typename V>
int tree_apply(b2r_tree *t, int64_t n, const I *indices, const V *values,
               const uint8_t *mode, cudaStream_t stream) {
  B2R_TRY(allow_big_smem(tree_leaf_pass<I, V>));
  B2R_TRY(allow_big_smem(tree_internal_pass<I>));
  for (int64_t base = 0; base < n; base += kTreeChunk) {
    const int len = (int)((n - base) < kTreeChunk ? (n - base) : kTreeChunk);
    const int padded = padded_size(len);
    int threads = padded / 2;
    if (threads < 32) threads = 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = (size_t)padded * 16;
    tree_leaf_pass<I, V><<<1, threads, smem, stream>>>(
        t->heap, t->depth, t->leaves, len, padded, indices + base,
        values + base, mode ? mode + base : nullptr, base, t->delta, t->n_eff,
        t->max_rec, t->status);
    B2R_LAUNCHED();
    if (t->depth > 0) {
      tree_internal_pass<I><<<t->depth, threads, smem, stream>>>(
          t->heap, t->depth, padded, indices + base, t->delta, t->n_eff);
      B2R_LAUNCHED();
    }
  }
  return B2R_OK;
}

template int tree_apply<int64_t, double>(b2r_tree *, int64_t, const int64_t *,
                                         const double *, const uint8_t *,
                                         cudaStream_t);
template int tree_apply<int32_t, float>(b2r_tree *, int64_t, const int32_t *,
                                        const float *, const uint8_t *,
                                        cudaStream_t);
template int tree_apply<int32_t, double>(b2r_tree *, int64_t, const int32_t *,
                                         const double *, const uint8_t *,
                                         cudaStream_t);

}  // namespace b2r

using b2r::as_stream;
using b2r::fail;

extern "C" {

int b2r_tree_create(int64_t capacity, b2r_tree **out) {
  if (out == nullptr) return fail(B2R_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (capacity <= 0)
    return fail(B2R_ERR_INVALID_ARGUMENT,
                "Sum tree capacity should be positive. Got: %lld",
                (long long)capacity);
  if (capacity > (1ll << 30))
    return fail(B2R_ERR_UNSUPPORTED, "capacity above 2^30 is not supported");
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count == 0)
    return fail(B2R_ERR_CUDA,
                "no CUDA device: libb200replay has no CPU fallback");
  b2r_tree *t = new (std::nothrow) b2r_tree();
  if (!t) return fail(B2R_ERR_INVALID_ARGUMENT, "out of host memory");
  t->capacity = capacity;
  int depth = 0;
  while ((1ll << depth) < capacity) ++depth;  // ceil(log2(capacity)), ST:81
  t->depth = depth;
  t->leaves = 1ll << depth;
  const size_t nodes = (size_t)(2 * t->leaves - 1);
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->heap), nodes * 8));
  B2R_CUDA(cudaMemset(t->heap, 0, nodes * 8));
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->max_rec), 8));
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->status), 16));
  B2R_CUDA(cudaMemset(t->status, 0, 16));
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->delta), b2r::kTreeChunk * 8));
  B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&t->n_eff), 4));
  const double one = 1.0;  // ST:89
  B2R_CUDA(cudaMemcpy(t->max_rec, &one, 8, cudaMemcpyHostToDevice));
  *out = t;
  return B2R_OK;
}

int b2r_tree_destroy(b2r_tree *t) {
  if (!t) return B2R_OK;
  cudaFree(t->heap);
  cudaFree(t->max_rec);
  cudaFree(t->status);
  cudaFree(t->delta);
  cudaFree(t->n_eff);
  t->bounce.release();
  delete t;
  return B2R_OK;
}

int b2r_tree_depth(const b2r_tree *t) { return t ? t->depth : -1; }

int b2r_tree_check(b2r_tree *t, b2r_stream stream) {
  int64_t st[2];
  B2R_CUDA(cudaMemcpyAsync(st, t->status, 16, cudaMemcpyDeviceToHost,
                           as_stream(stream)));
  B2R_CUDA(cudaStreamSynchronize(as_stream(stream)));
  if (st[0] == 0) return B2R_OK;
  B2R_CUDA(cudaMemsetAsync(t->status, 0, 16, as_stream(stream)));
  if (st[0] == B2R_ERR_NEGATIVE_PRIORITY)
    return fail(B2R_ERR_NEGATIVE_PRIORITY,
                "Sum tree values should be nonnegative (element %lld)",
                (long long)st[1]);
  return fail((int)st[0], "sum tree index out of range (element %lld)",
              (long long)st[1]);
}

int b2r_tree_set(b2r_tree *t, int64_t n, const int64_t *indices,
                 const double *values, int64_t *bad_pos, b2r_stream stream) {
  if (bad_pos) *bad_pos = -1;
  if (n <= 0) return B2R_OK;
  cudaStream_t s = as_stream(stream);
  B2R_TRY(t->bounce.reserve((size_t)n * 16 + 16));
  memcpy(t->bounce.host, indices, (size_t)n * 8);
  memcpy(t->bounce.host + (size_t)n * 8, values, (size_t)n * 8);
  B2R_CUDA(cudaMemcpyAsync(t->bounce.dev, t->bounce.host, (size_t)n * 16,
                           cudaMemcpyHostToDevice, s));
  B2R_TRY((b2r::tree_apply<int64_t, double>(
      t, n, reinterpret_cast<const int64_t *>(t->bounce.dev),
      reinterpret_cast<const double *>(t->bounce.dev + (size_t)n * 8), nullptr,
      s)));
  int64_t st[2];
  B2R_CUDA(cudaMemcpyAsync(st, t->status, 16, cudaMemcpyDeviceToHost, s));
  B2R_CUDA(cudaStreamSynchronize(s));
  if (st[0] != 0) {
    B2R_CUDA(cudaMemsetAsync(t->status, 0, 16, s));
    if (bad_pos) *bad_pos = st[1];
    if (st[0] == B2R_ERR_NEGATIVE_PRIORITY)
      return fail(B2R_ERR_NEGATIVE_PRIORITY,
                  "Sum tree values should be nonnegative. Got %g",
                  values[st[1]]);
    return fail((int)st[0], "index %lld is out of range for a tree of %lld leaves",
                (long long)indices[st[1]], (long long)t->leaves);
  }
  return B2R_OK;
}

int b2r_tree_set_device(b2r_tree *t, int64_t n, const int32_t *indices,
                        const float *values, b2r_stream stream) {
  if (n <= 0) return B2R_OK;
  return b2r::tree_apply<int32_t, float>(t, n, indices, values, nullptr,
                                         as_stream(stream));
}

int b2r_tree_get(b2r_tree *t, int64_t n, const int64_t *indices, double *out,
                 b2r_stream stream) {
  if (n <= 0) return B2R_OK;
  cudaStream_t s = as_stream(stream);
  for (int64_t k = 0; k < n; ++k)
    if (indices[k] < 0 || indices[k] >= t->leaves)
      return fail(B2R_ERR_INDEX_RANGE,
                  "index %lld is out of bounds for a tree of %lld leaves",
                  (long long)indices[k], (long long)t->leaves);
  B2R_TRY(t->bounce.reserve((size_t)n * 16));
  memcpy(t->bounce.host, indices, (size_t)n * 8);
  B2R_CUDA(cudaMemcpyAsync(t->bounce.dev, t->bounce.host, (size_t)n * 8,
                           cudaMemcpyHostToDevice, s));
  double *dout = reinterpret_cast<double *>(t->bounce.dev + (size_t)n * 8);
  b2r::tree_get_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      t->heap, t->leaves, n, reinterpret_cast<const int64_t *>(t->bounce.dev),
      dout);
  B2R_LAUNCHED();
  B2R_CUDA(cudaMemcpyAsync(t->bounce.host + (size_t)n * 8, dout, (size_t)n * 8,
                           cudaMemcpyDeviceToHost, s));
  B2R_CUDA(cudaStreamSynchronize(s));
  memcpy(out, t->bounce.host + (size_t)n * 8, (size_t)n * 8);
  return B2R_OK;
}

int b2r_tree_total(b2r_tree *t, double *out, b2r_stream stream) {
  B2R_CUDA(cudaMemcpyAsync(out, t->heap, 8, cudaMemcpyDeviceToHost,
                           as_stream(stream)));
  B2R_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return B2R_OK;
}

int b2r_tree_max_recorded(b2r_tree *t, double *out, b2r_stream stream) {
  B2R_CUDA(cudaMemcpyAsync(out, t->max_rec, 8, cudaMemcpyDeviceToHost,
                           as_stream(stream)));
  B2R_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return B2R_OK;
}

int b2r_tree_set_max_recorded(b2r_tree *t, double value, b2r_stream stream) {
  b2r::tree_set_scalar_kernel<<<1, 1, 0, as_stream(stream)>>>(t->max_rec, value);
  B2R_LAUNCHED();
  return B2R_OK;
}

int b2r_tree_sample(b2r_tree *t, int64_t n, const double *query01, int64_t *out,
                    b2r_stream stream) {
  if (n <= 0) return B2R_OK;
  cudaStream_t s = as_stream(stream);
  B2R_TRY(t->bounce.reserve((size_t)n * 16 + 8));
  memcpy(t->bounce.host, query01, (size_t)n * 8);
  B2R_CUDA(cudaMemcpyAsync(t->bounce.dev, t->bounce.host, (size_t)n * 8,
                           cudaMemcpyHostToDevice, s));
  int64_t *dout = reinterpret_cast<int64_t *>(t->bounce.dev + (size_t)n * 8);
  double *dtotal = reinterpret_cast<double *>(t->bounce.dev + (size_t)n * 16);
  const size_t smem = ((size_t)2 << b2r::kTopLevels) * 8;
  b2r::tree_query_kernel<<<(unsigned)((n + 255) / 256), 256, smem, s>>>(
      t->heap, t->depth, n, reinterpret_cast<const double *>(t->bounce.dev),
      dout, dtotal);
  B2R_LAUNCHED();
  B2R_CUDA(cudaMemcpyAsync(t->bounce.host + (size_t)n * 8, dout,
                           (size_t)n * 8 + 8, cudaMemcpyDeviceToHost, s));
  B2R_CUDA(cudaStreamSynchronize(s));
  double total;
  memcpy(&total, t->bounce.host + (size_t)n * 16, 8);
  if (total == 0.0)
    return fail(B2R_ERR_EMPTY_TREE, "Cannot sample from an empty sum tree.");
  memcpy(out, t->bounce.host + (size_t)n * 8, (size_t)n * 8);
  return B2R_OK;
}

int b2r_tree_read_level(b2r_tree *t, int level, double *out, b2r_stream stream) {
  if (level < 0 || level > t->depth)
    return fail(B2R_ERR_INVALID_ARGUMENT, "level %d out of range", level);
  const size_t count = (size_t)1 << level;
  B2R_CUDA(cudaMemcpyAsync(out, t->heap + (count - 1), count * 8,
                           cudaMemcpyDeviceToHost, as_stream(stream)));
  B2R_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return B2R_OK;
}

int b2r_tree_write_level(b2r_tree *t, int level, const double *in,
                         b2r_stream stream) {
  if (level < 0 || level > t->depth)
    return fail(B2R_ERR_INVALID_ARGUMENT, "level %d out of range", level);
  const size_t count = (size_t)1 << level;
  B2R_CUDA(cudaMemcpyAsync(t->heap + (count - 1), in, count * 8,
                           cudaMemcpyHostToDevice, as_stream(stream)));
  B2R_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return B2R_OK;
}

}  // extern "C"
