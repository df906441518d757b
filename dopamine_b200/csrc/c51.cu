// C51 distributional Bellman target, projection, cross-entropy, priorities and IS
// weights: replaces the TensorFlow graph built by RainbowAgent
// (rainbow_agent.py:200-305) and project_distribution (rainbow_agent.py:340-494).
//
// One warp per batch row; atoms are strided over lanes, reductions are warp
// shuffles, the (source support, probability) pairs of a row sit in shared memory
// so that lane i accumulates output atom i over all j exactly as the dense
// [B, N, N] form does — without ever materialising it.  All arithmetic is f32
// with explicit round-to-nearest intrinsics where TF evaluates separate ops
// (no FMA contraction), true division and IEEE sqrt.
//
// Traffic per row (A=18, N=51): 3 672 B of target logits + 204 B of online logits
// read, <= 204 B target + 12 B scalars written (+3 672 B if grad_logits is asked).
#include "common.cuh"

namespace b2r {
namespace {

constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out[i] = sum_j clip(1 - |clip(s_j, z0, zlast) - z_i| / dz, 0, 1) * w_j
// (rainbow_agent.py:424-492); s, w in shared memory, i strided over lanes.
__device__ __forceinline__ float project_atom(const float *s, const float *w,
                                              int n, float z_i, float z0,
                                              float zlast, float dz) {
  float acc = 0.f;
  for (int j = 0; j < n; ++j) {
    const float clipped = fminf(fmaxf(s[j], z0), zlast);
    const float gap = fabsf(__fsub_rn(clipped, z_i));
    float hat = __fsub_rn(1.0f, __fdiv_rn(gap, dz));
    hat = fminf(fmaxf(hat, 0.f), 1.f);
    acc = __fadd_rn(acc, __fmul_rn(hat, w[j]));
  }
  return acc;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
c51_project_kernel(int batch, int n, const float *__restrict__ supports,
                   const float *__restrict__ weights,
                   const float *__restrict__ z, float *__restrict__ out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerBlock + warp;
  if (b >= batch) return;
  float *s = smem + (size_t)warp * 2 * n;
  float *w = s + n;
  for (int j = lane; j < n; j += 32) {
    s[j] = supports[(size_t)b * n + j];
    w[j] = weights[(size_t)b * n + j];
  }
  __syncwarp();
  const float z0 = z[0], zlast = z[n - 1];
  const float dz = __fsub_rn(z[1], z[0]);  // rainbow_agent.py:381-383
  for (int i = lane; i < n; i += 32)
    out[(size_t)b * n + i] = project_atom(s, w, n, z[i], z0, zlast, dz);
}

struct LossArgs {
  b2r_c51_args u;
  float *raw_weights;  // scratch (B,)
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) c51_loss_kernel(LossArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerBlock + warp;
  const int N = a.u.num_atoms, A = a.u.num_actions;
  if (b >= a.u.batch) return;
  float *cur = smem + (size_t)warp * 3 * N;  // exp(x - max) of the action scanned
  float *best_p = cur + N;                   // probabilities of the argmax action
  float *sup = cur + 2 * N;                  // Bellman support r + g*z_j
  const float *z = a.u.support;

  // ---- target network head: softmax, q = sum z*p, first argmax
  //      (atari_lib.py:141-143, rainbow_agent.py:238-248)
  float best_q = 0.f;
  int best_a = -1;
  for (int act = 0; act < A; ++act) {
    const float *x = a.u.target_logits + ((size_t)b * A + act) * N;
    float m = -INFINITY;
    for (int i = lane; i < N; i += 32) m = fmaxf(m, x[i]);
    m = warp_max(m);
    float part = 0.f;
    for (int i = lane; i < N; i += 32) {
      const float e = expf(__fsub_rn(x[i], m));
      cur[i] = e;
      part = __fadd_rn(part, e);
    }
    const float denom = warp_sum(part);
    float qpart = 0.f;
    for (int i = lane; i < N; i += 32) {
      const float p = __fdiv_rn(cur[i], denom);
      cur[i] = p;
      qpart = __fadd_rn(qpart, __fmul_rn(z[i], p));
    }
    const float q = warp_sum(qpart);
    if (best_a < 0 || q > best_q) {  // strict > keeps the first maximum
      best_q = q;
      best_a = act;
      for (int i = lane; i < N; i += 32) best_p[i] = cur[i];
    }
    __syncwarp();
  }

  // ---- Bellman support (rainbow_agent.py:229-235)
  const float live = __fsub_rn(1.0f, (float)a.u.terminals[b]);
  const float gwt = __fmul_rn(a.u.cumulative_gamma, live);
  const float r = a.u.rewards[b];
  for (int j = lane; j < N; j += 32) sup[j] = __fadd_rn(r, __fmul_rn(gwt, z[j]));
  __syncwarp();

  // ---- projection + cross entropy against the chosen online logits
  //      (rainbow_agent.py:250, 262-271)
  const float z0 = z[0], zlast = z[N - 1];
  const float dz = __fsub_rn(z[1], z[0]);
  const int chosen = a.u.actions[b];
  const float *x = a.u.online_logits + ((size_t)b * A + chosen) * N;
  float m = -INFINITY;
  for (int i = lane; i < N; i += 32) m = fmaxf(m, x[i]);
  m = warp_max(m);
  float part = 0.f;
  for (int i = lane; i < N; i += 32) part = __fadd_rn(part, expf(__fsub_rn(x[i], m)));
  const float denom = warp_sum(part);
  const float lse = logf(denom);
  float ce_part = 0.f, tsum_part = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float t = project_atom(sup, best_p, N, z[i], z0, zlast, dz);
    cur[i] = t;  // `cur` is free again: keep the target for the gradient pass
    if (a.u.target) a.u.target[(size_t)b * N + i] = t;
    const float logp = __fsub_rn(__fsub_rn(x[i], m), lse);
    ce_part = __fadd_rn(ce_part, __fmul_rn(t, logp));
    tsum_part = __fadd_rn(tsum_part, t);
  }
  const float ce = -warp_sum(ce_part);
  const float tsum = warp_sum(tsum_part);
  if (lane == 0) {
    a.u.loss[b] = ce;
    a.u.priorities[b] = sqrtf(__fadd_rn(ce, 1e-10f));  // rainbow_agent.py:290
    if (a.u.sampling_probabilities)  // rainbow_agent.py:279
      a.raw_weights[b] =
          __fdiv_rn(1.0f, sqrtf(__fadd_rn(a.u.sampling_probabilities[b], 1e-10f)));
  }
  if (a.u.grad_logits) {
    // d ce / d x_i = softmax(x)_i * sum(t) - t_i ; other actions get zero.
    // Scaled by w_b / B in c51_finalize_kernel once max(w) is known.
    __syncwarp();
    float *g = a.u.grad_logits + (size_t)b * A * N;
    for (int k = lane; k < A * N; k += 32) {
      const int act = k / N, i = k - act * N;
      float v = 0.f;
      if (act == chosen) {
        const float p = __fdiv_rn(expf(__fsub_rn(x[i], m)), denom);
        v = __fsub_rn(__fmul_rn(p, tsum), cur[i]);
      }
      g[k] = v;
    }
  }
}

// One CTA: IS weights / max, mean weighted loss, gradient scaling
// (rainbow_agent.py:279-280, 293, 305).
__global__ void __launch_bounds__(1024) c51_finalize_kernel(LossArgs a) {
  __shared__ float red[32];
  __shared__ float s_max;
  const int B = a.u.batch, N = a.u.num_atoms, A = a.u.num_actions;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool weighted = a.u.sampling_probabilities != nullptr;
  float wmax = 1.f;
  if (weighted) {
    float m = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) m = fmaxf(m, a.raw_weights[b]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmaxf(mm, red[w]);
      s_max = mm;
    }
    __syncthreads();
    wmax = s_max;
  }
  float part = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float w = weighted ? __fdiv_rn(a.raw_weights[b], wmax) : 1.f;
    if (a.u.weights) a.u.weights[b] = w;
    part = __fadd_rn(part, __fmul_rn(w, a.u.loss[b]));
  }
  part = warp_sum(part);
  __syncthreads();
  if (lane == 0) red[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0 && a.u.mean_weighted_loss) {
    float total = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total = __fadd_rn(total, red[w]);
    *a.u.mean_weighted_loss = __fdiv_rn(total, (float)B);
  }
  if (a.u.grad_logits) {
    const float inv_b = __fdiv_rn(1.0f, (float)B);
    for (int k = threadIdx.x; k < B * N; k += blockDim.x) {
      const int b = k / N, i = k - b * N;
      const float w = weighted ? __fdiv_rn(a.raw_weights[b], wmax) : 1.f;
      float *g = a.u.grad_logits + ((size_t)b * A + a.u.actions[b]) * N + i;
      *g = __fmul_rn(*g, __fmul_rn(w, inv_b));
    }
  }
}

float *g_raw_weights = nullptr;
int g_raw_weights_cap = 0;

}  // namespace
}  // namespace b2r

using b2r::as_stream;
using b2r::fail;

extern "C" {

int b2r_c51_project(int32_t batch, int32_t num_atoms, const float *supports,
                    const float *weights, const float *target_support, float *out,
                    b2r_stream stream) {
  if (batch <= 0 || num_atoms < 2)
    return fail(B2R_ERR_INVALID_ARGUMENT,
                "project_distribution needs batch > 0 and at least 2 atoms");
  const size_t smem = (size_t)b2r::kWarpsPerBlock * 2 * num_atoms * 4;
  if (smem > 48 * 1024)
    return fail(B2R_ERR_UNSUPPORTED, "num_atoms too large");
  const int blocks = (batch + b2r::kWarpsPerBlock - 1) / b2r::kWarpsPerBlock;
  b2r::c51_project_kernel<<<blocks, b2r::kWarpsPerBlock * 32, smem,
                            as_stream(stream)>>>(batch, num_atoms, supports,
                                                 weights, target_support, out);
  B2R_LAUNCHED();
  return B2R_OK;
}

int b2r_c51_loss(const b2r_c51_args *args, b2r_stream stream) {
  if (!args || args->batch <= 0 || args->num_atoms < 2 || args->num_actions <= 0)
    return fail(B2R_ERR_INVALID_ARGUMENT, "bad C51 shape");
  if (!args->support || !args->target_logits || !args->online_logits ||
      !args->actions || !args->rewards || !args->terminals || !args->loss ||
      !args->priorities)
    return fail(B2R_ERR_INVALID_ARGUMENT, "a required C51 pointer is NULL");
  const size_t smem = (size_t)b2r::kWarpsPerBlock * 3 * args->num_atoms * 4;
  if (smem > 48 * 1024) return fail(B2R_ERR_UNSUPPORTED, "num_atoms too large");
  cudaStream_t s = as_stream(stream);
  if (args->batch > b2r::g_raw_weights_cap) {
    if (b2r::g_raw_weights) cudaFree(b2r::g_raw_weights);
    b2r::g_raw_weights = nullptr;
    int cap = 4096;
    while (cap < args->batch) cap *= 2;
    B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&b2r::g_raw_weights), (size_t)cap * 4));
    b2r::g_raw_weights_cap = cap;
  }
  b2r::LossArgs a;
  a.u = *args;
  a.raw_weights = b2r::g_raw_weights;
  const int blocks = (args->batch + b2r::kWarpsPerBlock - 1) / b2r::kWarpsPerBlock;
  b2r::c51_loss_kernel<<<blocks, b2r::kWarpsPerBlock * 32, smem, s>>>(a);
  B2R_LAUNCHED();
  if (args->weights || args->mean_weighted_loss || args->grad_logits) {
    int threads = 32;
    while (threads < args->batch && threads < 1024) threads <<= 1;
    b2r::c51_finalize_kernel<<<1, threads, 0, s>>>(a);
    B2R_LAUNCHED();
  }
  return B2R_OK;
}

}  // extern "C"
