// C51 distributional Bellman target, projection, cross-entropy, priorities and IS
// weights: replaces the TensorFlow graph built by RainbowAgent
// (rainbow_agent.py:200-305) and project_distribution (rainbow_agent.py:340-494).
//
// Two kernels share the arithmetic.  Batches below 128 rows (latency-bound): one CTA per
// row, one warp per action (c51_loss_kernel).  From 128 rows (bound by instruction
// issue): one WARP per row, 8 lanes per action (c51_loss_rows_kernel) — a third of the
// instructions per row.  In both, atoms are strided over lanes, reductions are warp
// shuffles, the (Bellman support, probability) pairs of a row sit in shared memory so
// that each output atom is accumulated over all j exactly as the dense [B, N, N] form
// does — without ever materialising it.  All arithmetic is f32 with explicit
// round-to-nearest intrinsics where TF evaluates separate ops (no FMA contraction),
// true division and IEEE sqrt.
//
// Traffic per row (A=18, N=51): 3 672 B of target logits + 204 B of online logits
// read, <= 204 B target + 12 B scalars written (+3 672 B if grad_logits is asked).
#include "tree.cuh"

#include <cstdlib>

namespace b2r {
namespace {

B2R_TRACE_DECL

constexpr int kWarpsPerBlock = 4;

// UNROLLED = false keeps the reductions as 5-trip loops (compact code for the
// latency-bound small-batch instance); true unrolls them: the large-batch instance
// is bound by instruction issue and the loop control is 2/3 of a rolled reduction.
template <bool UNROLLED = false>
__device__ __forceinline__ float warp_max(float v) {
  if (UNROLLED) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  } else {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}
template <bool UNROLLED = false>
__device__ __forceinline__ float warp_sum(float v) {
  if (UNROLLED) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  } else {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
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
  float *weighted;        // scratch (B,): w_b * loss_b, reduced by the last CTA
  unsigned int *ticket;   // scratch: CTAs finished (self-resetting)
  int warps;              // warps per CTA = min(num_actions, 32)
  // Priority write-back at the tail of the loss kernel (batches of at most 32 rows:
  // the fused step at the agent's batch size).  The last CTA to finish its row runs the
  // match-based tree update over the priorities the grid has just produced — one kernel
  // boundary less on a chain of three short kernels.
  int fuse_tree;
  UpdateArgs<int32_t, float> tree;
};

// One CTA per batch row, one warp per action (rainbow_agent.py:200-293):
//   A. every warp: softmax of its action's target logits, q = sum z*p
//      (atari_lib.py:141-143); meanwhile the CTA finds min(sampling_probabilities)
//      for the IS-weight normalisation.
//   B. first argmax over actions (RA:238-248); Bellman support r + g^n(1-t) z
//      (RA:229-235); projection, `parts` threads per output atom (RA:381-494).
//   C. warp 0: cross-entropy vs the chosen online logits (RA:262-271), priority
//      sqrt(loss + 1e-10) (RA:290), weight 1/sqrt(p + 1e-10) / max (RA:279-280).
//   D. all threads: gradient row; the last CTA to finish sums w*loss in a fixed
//      order for mean_weighted_loss (RA:293, 305).
constexpr int kMaxAtomsPerLane = 4;  // num_atoms <= 128

// PL = atoms per lane (2 covers C51's 51 atoms; fewer unrolled copies = less
// straight-line code to fetch on a cold instruction cache).
// The large-batch instance runs a third of the warps per row (<= 11 warps) and wants
// many rows resident per SM: bounds that keep it at 32 registers (10 CTAs of 192
// threads per SM); the small-batch instance runs one warp per action.
template <int PL, bool FAST>
__global__ void __launch_bounds__(FAST ? 384 : 1024, FAST ? 5 : 1)
c51_loss_kernel(LossArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int N = a.u.num_atoms, A = a.u.num_actions, W = a.warps;
  float *bestp = smem + (size_t)warp * N;        // [W][N] best action's probs
  float *sup = smem + (size_t)W * N;             // [N] Bellman support
  float *tgt = sup + N;                          // [N] projected target
  float *onl = tgt + N;                          // [N] chosen online logits
  float *zs = onl + N;                           // [N] the support, staged
  float *lgp = zs + N;                           // [N] log_softmax of the chosen logits
  __shared__ float s_q[32];
  __shared__ int s_a[32];
  __shared__ float s_red[32];
  __shared__ float s_scalar[4];  // tsum, w_b, m, denom
  __shared__ float s_ce;
  __shared__ bool s_last;
  const float *z = a.u.support;
  B2R_MARK(0);
  pdl_release();
  pdl_acquire();
  B2R_MARK(1);
  const int rows = a.u.batch_count ? min(*a.u.batch_count, a.u.batch) : a.u.batch;
  if (b >= rows) return;  // (mean_weighted_loss is not supported with batch_count)

  // ---- every global load the row needs is issued here, before the first use:
  // the kernel is latency-bound at batch 32 and this keeps it to one round trip.
  const int chosen = a.u.actions[b];
  const float r = a.u.rewards[b];
  const float term = (float)a.u.terminals[b];
  const float my_prob = a.u.sampling_probabilities ? a.u.sampling_probabilities[b] : 1.f;
  // min over the batch of the sampling probabilities (IS-weight normaliser):
  // handed in by the sampler when it produced the batch, else reduced here.
  float pmin = INFINITY;
  if (a.u.min_probability)
    pmin = *a.u.min_probability;
  else if (a.u.sampling_probabilities)
    for (int k = threadIdx.x; k < rows; k += blockDim.x)
      pmin = fminf(pmin, a.u.sampling_probabilities[k]);
  float zl[PL], xt[PL], xo[PL];
  const int act0 = warp;  // first (usually only) action of this warp
  const float *__restrict__ trow = a.u.target_logits + (size_t)b * A * N;
  const float *__restrict__ orow = a.u.online_logits + (size_t)b * A * N;
#pragma unroll
  for (int t = 0; t < PL; ++t) {
    const int i = lane + 32 * t;
    const bool ok = i < N && act0 < A;
    zl[t] = i < N ? z[i] : 0.f;
    xt[t] = ok ? trow[act0 * N + i] : -INFINITY;
    xo[t] = ok ? orow[act0 * N + i] : 0.f;
  }

  B2R_MARK(2);
  // ---- A. per-action softmax + q-value (atari_lib.py:141-143)
  float best_q = 0.f;
  int best_a = -1;
  for (int act = warp; act < A; act += W) {
    if (act != act0) {  // more actions than warps: later rounds load as they go
#pragma unroll
      for (int t = 0; t < PL; ++t) {
        const int i = lane + 32 * t;
        xt[t] = i < N ? trow[act * N + i] : -INFINITY;
        // only the chosen action's online logits are ever used
        xo[t] = (i < N && act == chosen) ? orow[act * N + i] : 0.f;
      }
    }
    if (act == chosen) {
      // log_softmax of the chosen action's online logits (rainbow_agent.py:262-271),
      // by the warp that holds them, while the other warps do their softmaxes
      float mo = -INFINITY;
#pragma unroll
      for (int t = 0; t < PL; ++t)
        if (lane + 32 * t < N) mo = fmaxf(mo, xo[t]);
      mo = warp_max<FAST>(mo);
      float ps = 0.f;
#pragma unroll
      for (int t = 0; t < PL; ++t)
        if (lane + 32 * t < N) ps = __fadd_rn(ps, expf(__fsub_rn(xo[t], mo)));
      const float den_o = warp_sum<FAST>(ps);
      const float lse = logf(den_o);
#pragma unroll
      for (int t = 0; t < PL; ++t) {
        if (lane + 32 * t < N) {
          onl[lane + 32 * t] = xo[t];
          lgp[lane + 32 * t] = __fsub_rn(__fsub_rn(xo[t], mo), lse);
        }
      }
      if (lane == 0) {
        s_scalar[2] = mo;
        s_scalar[3] = den_o;
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < PL; ++t) m = fmaxf(m, xt[t]);
    m = warp_max<FAST>(m);
    float e[PL];
    float psum = 0.f;
#pragma unroll
    for (int t = 0; t < PL; ++t) {
      const bool ok = lane + 32 * t < N;
      e[t] = ok ? expf(__fsub_rn(xt[t], m)) : 0.f;
      if (ok) psum = __fadd_rn(psum, e[t]);
    }
    const float denom = warp_sum<FAST>(psum);
    float qpart = 0.f;
#pragma unroll
    for (int t = 0; t < PL; ++t) {
      if (lane + 32 * t < N) {
        e[t] = __fdiv_rn(e[t], denom);
        qpart = __fadd_rn(qpart, __fmul_rn(zl[t], e[t]));
      }
    }
    const float q = warp_sum<FAST>(qpart);
    if (best_a < 0 || q > best_q) {  // strict > keeps the first maximum
      best_q = q;
      best_a = act;
#pragma unroll
      for (int t = 0; t < PL; ++t)
        if (lane + 32 * t < N) bestp[lane + 32 * t] = e[t];
    }
  }
  B2R_MARK(3);
  if (lane == 0) {
    s_q[warp] = best_q;
    s_a[warp] = best_a;
  }
  // min over the batch of the sampling probabilities: 1/sqrt(p + 1e-10) is
  // monotone under round-to-nearest, so max_b w_b == w(min_b p_b) exactly.
  if (a.u.sampling_probabilities) {
    pmin = -warp_max(-pmin);
    if (lane == 0) s_red[warp] = pmin;
  }
  // Bellman support (rainbow_agent.py:229-235)
  const float live = __fsub_rn(1.0f, term);
  const float gwt = __fmul_rn(a.u.cumulative_gamma, live);
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float zj = z[j];
    zs[j] = zj;
    sup[j] = __fadd_rn(r, __fmul_rn(gwt, zj));
  }
  __syncthreads();

  B2R_MARK(4);
  // ---- B. argmax action (first maximum, RA:238-248), projection (RA:381-494)
  // every warp reduces the per-warp bests itself (5 shuffle steps instead of a
  // serial scan of shared memory): highest q, ties to the smaller action index
  int win = lane;
  {
    float q = lane < W ? s_q[lane] : 0.f;
    int act = lane < W ? s_a[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float q2 = __shfl_xor_sync(0xffffffffu, q, o);
      const int act2 = __shfl_xor_sync(0xffffffffu, act, o);
      const int win2 = __shfl_xor_sync(0xffffffffu, win, o);
      if (act2 >= 0 && (act < 0 || q2 > q || (q2 == q && act2 < act))) {
        q = q2;
        act = act2;
        win = win2;
      }
    }
  }
  const float *next_p = smem + (size_t)win * N;
  const float z0 = zs[0], zlast = zs[N - 1];
  const float dz = __fsub_rn(zs[1], zs[0]);
  // The dense form sums hat(i, j) * p_j over all j, but hat is exactly 0 unless
  // |clip(s_j) - z_i| < dz.  The Bellman atoms s_j = r + g * z_j are non-decreasing
  // in j (g = gamma^n (1 - terminal) >= 0), so the j that can reach atom i form one
  // interval: those with s_j inside (z_i - dz, z_i + dz), open-ended below for the
  // first atom and above for the last one (clipping).  Each thread evaluates the
  // reference's expression over a conservative superset of its interval (two extra
  // positions either side cover any rounding of the bounds; terminal rows, where all
  // s_j coincide, take the whole range) in ascending j.  Skipped terms are exact
  // zeros, so the sum is the dense form's.
#pragma unroll 1
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float zi = zs[i];
    int jl = 0, jh = N - 1;
    if (gwt > 0.f) {
      // s_j > zi - dz  <=>  j > ((zi - dz - r) / g - z0) / dz
      const float inv = __fdividef(1.0f, gwt * dz);
      const float lo = (zi - dz - r - gwt * z0) * inv;
      const float hi = (zi + dz - r - gwt * z0) * inv;
      if (i > 0 && lo > 2.f) jl = min(N - 1, (int)fminf(lo, 1e6f) - 2);
      if (i < N - 1 && hi < (float)(N - 3)) jh = max(0, (int)fmaxf(hi, -1e6f) + 3);
    }
    float acc = 0.f;
#pragma unroll 1
    for (int j = jl; j <= jh; ++j) {
      const float clipped = fminf(fmaxf(sup[j], z0), zlast);
      const float gap = fabsf(__fsub_rn(clipped, zi));
      if (gap < dz) {
        float hat = __fsub_rn(1.0f, __fdiv_rn(gap, dz));
        hat = fminf(fmaxf(hat, 0.f), 1.f);
        acc = __fadd_rn(acc, __fmul_rn(hat, next_p[j]));
      }
    }
    tgt[i] = acc;
    if (a.u.target) a.u.target[(size_t)b * N + i] = acc;
  }
  __syncthreads();

  B2R_MARK(5);
  // ---- C. cross entropy (RA:262-271), priority (RA:290), weight (RA:279-280)
  const float *x = onl;
  if (warp == 0) {
    float ce_part = 0.f, tsum_part = 0.f;
#pragma unroll 1
    for (int i = lane; i < N; i += 32) {
      const float t = tgt[i];
      const float logp = lgp[i];
      ce_part = __fadd_rn(ce_part, __fmul_rn(t, logp));
      tsum_part = __fadd_rn(tsum_part, t);
    }
    const float ce = -warp_sum(ce_part);
    const float tsum = warp_sum(tsum_part);
    if (lane == 0) {
      a.u.loss[b] = ce;
      a.u.priorities[b] = sqrtf(__fadd_rn(ce, 1e-10f));
      s_scalar[0] = tsum;
      s_ce = ce;
    }
  }
  // importance weight (RA:279-280), by another warp beside the cross entropy
  if (warp == (W > 1 ? 1 : 0) && lane == 0) {
    float w = 1.f;
    if (a.u.sampling_probabilities) {
      float mn = pmin;  // handed over by the sampler, or reduced per warp above
      if (!a.u.min_probability) {
        mn = s_red[0];
#pragma unroll 1
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mn = fminf(mn, s_red[k]);
      }
      const float wmax = __fdiv_rn(1.0f, sqrtf(__fadd_rn(mn, 1e-10f)));
      const float raw = __fdiv_rn(1.0f, sqrtf(__fadd_rn(my_prob, 1e-10f)));
      w = __fdiv_rn(raw, wmax);
    }
    if (a.u.weights) a.u.weights[b] = w;
    s_scalar[1] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) a.weighted[b] = __fmul_rn(s_scalar[1], s_ce);

  B2R_MARK(6);
  // ---- D. gradient of mean(w * ce) w.r.t. the online logits
  if (a.u.grad_logits) {
    const float tsum = s_scalar[0], w = s_scalar[1], m = s_scalar[2], denom = s_scalar[3];
    const float scale = __fmul_rn(w, __fdiv_rn(1.0f, (float)rows));
    float *g = a.u.grad_logits + (size_t)b * A * N;
    for (int k = threadIdx.x; k < A * N; k += blockDim.x) {
      const int act = k / N, i = k - act * N;
      float v = 0.f;
      if (act == chosen) {
        const float p = __fdiv_rn(expf(__fsub_rn(x[i], m)), denom);
        v = __fmul_rn(__fsub_rn(__fmul_rn(p, tsum), tgt[i]), scale);
      }
      g[k] = v;
    }
  }

  B2R_MARK(7);
  B2R_MARK_END(8);
  // ---- the last CTA to finish: mean weighted loss, reduced in a fixed order, and
  // (fuse_tree) the priority write-back.
  if (a.u.mean_weighted_loss == nullptr && !a.fuse_tree) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.ticket, 1u);
    s_last = (done == gridDim.x - 1);
    if (s_last) *a.ticket = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (a.fuse_tree) {
    tree_update_tiny_body<false>(a.tree);
    B2R_MARK_END(9);
  }
  if (a.u.mean_weighted_loss == nullptr) return;
  float acc = 0.f;
  for (int k = threadIdx.x; k < a.u.batch; k += blockDim.x)
    acc = __fadd_rn(acc, __ldcg(a.weighted + k));
  acc = warp_sum(acc);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total = __fadd_rn(total, s_red[w]);
    *a.u.mean_weighted_loss = __fdiv_rn(total, (float)a.u.batch);
  }
}

// ---- throughput instance (batches above 256 rows, num_atoms <= 64) ----------
// One WARP per batch row, kRowWarps rows per CTA.  The CTA-per-row kernel above
// spends most of its issue slots on 5-step shuffle reductions whose lanes are 80 %
// idle on the second atom pass (51 atoms over 32 lanes) and on block barriers: 6 300
// warp instructions per row at A = 18, N = 51.  Here G lanes share an action and hold
// 64 / G atoms each, so one warp instruction serves 32 / G actions, a reduction is
// log2(G) shuffle steps, and nothing but __syncwarp separates the phases.  The
// arithmetic is the same f32 sequence per element (exp of the max-shifted logit,
// probabilities, q = sum z p, first maximum), except that a probability is
// e * (1 / denom) instead of e / denom (<= 1 ulp apart; parity bound 1e-6 relative).
constexpr int kRowWarps = 4;
constexpr int kRowAtoms = 64;

template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Pad value of the lanes beyond num_atoms (and of the groups beyond num_actions):
// finite, so max-shifting never forms inf - inf, and exp(pad - max) is exactly 0.
constexpr float kLogitPad = -1e30f;

// NC = num_atoms at compile time (51: the lanes' bounds tests fold away), 0 = runtime.
template <int G, int NC>
__global__ void __launch_bounds__(kRowWarps * 32)
c51_loss_rows_kernel(LossArgs a) {
  constexpr int PL = NC ? (NC + G - 1) / G : kRowAtoms / G;  // atoms per lane
  constexpr int GROUPS = 32 / G;     // actions per warp instruction
  __shared__ float s_bestp[kRowWarps][kRowAtoms];  // greedy action's probabilities
  __shared__ float s_sup[kRowWarps][kRowAtoms];    // Bellman support
  __shared__ float s_red[kRowWarps];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / G, l = lane % G;
  const int N = NC ? NC : a.u.num_atoms, A = a.u.num_actions;
  const float *__restrict__ z = a.u.support;
  B2R_MARK(10);
  pdl_release();
  pdl_acquire();
  B2R_MARK(11);
  const int rows = a.u.batch_count ? min(*a.u.batch_count, a.u.batch) : a.u.batch;
  if ((int)blockIdx.x * kRowWarps >= rows) return;  // (no mean loss with batch_count)
  const int b = blockIdx.x * kRowWarps + warp;
  const bool row_ok = b < rows;

  // min over the batch of the sampling probabilities (IS-weight normaliser): handed
  // in by the sampler when it produced the batch, else reduced by the CTA.
  float pmin = INFINITY;
  if (a.u.sampling_probabilities) {
    if (a.u.min_probability) {
      pmin = *a.u.min_probability;
    } else {
      for (int k = threadIdx.x; k < rows; k += blockDim.x)
        pmin = fminf(pmin, a.u.sampling_probabilities[k]);
      pmin = -group_max<32>(-pmin);
      if (lane == 0) s_red[warp] = pmin;
      __syncthreads();
      pmin = s_red[0];
#pragma unroll
      for (int k = 1; k < kRowWarps; ++k) pmin = fminf(pmin, s_red[k]);
      __syncthreads();  // s_red is reused by the mean-loss reduction
    }
  }

  if (row_ok) {
    const int chosen = a.u.actions[b];
    const float r = a.u.rewards[b];
    const float term = (float)a.u.terminals[b];
    const float my_prob = a.u.sampling_probabilities ? a.u.sampling_probabilities[b] : 1.f;
    const float *__restrict__ trow = a.u.target_logits + (size_t)b * A * N;
    const float *__restrict__ orow = a.u.online_logits + (size_t)b * A * N;
    float zl[PL], xn[PL];
#pragma unroll
    for (int t = 0; t < PL; ++t) {
      const int i = l + G * t;
      zl[t] = i < N ? z[i] : 0.f;
      xn[t] = (i < N && grp < A) ? trow[grp * N + i] : kLogitPad;
    }
    float xo[2];  // the chosen action's online logits, atoms lane and lane + 32
#pragma unroll
    for (int t = 0; t < 2; ++t)
      xo[t] = lane + 32 * t < N ? orow[chosen * N + lane + 32 * t] : kLogitPad;

    // ---- A. per-action softmax + q-value (atari_lib.py:141-143), GROUPS actions per
    // round; the next round's logits are in flight while this one is reduced
    float best_q = 0.f;
    int best_a = -1;
    float bp[PL];
#pragma unroll
    for (int t = 0; t < PL; ++t) bp[t] = 0.f;
    const int rounds = (A + GROUPS - 1) / GROUPS;
#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      const int act = rd * GROUPS + grp;
      const bool valid = act < A;
      float xt[PL];
#pragma unroll
      for (int t = 0; t < PL; ++t) xt[t] = xn[t];
      const int nact = act + GROUPS;
#pragma unroll
      for (int t = 0; t < PL; ++t) {
        const int i = l + G * t;
        xn[t] = (i < N && nact < A) ? trow[nact * N + i] : kLogitPad;
      }
      float m = xt[0];
#pragma unroll
      for (int t = 1; t < PL; ++t) m = fmaxf(m, xt[t]);
      m = group_max<G>(m);
      float e[PL];
      float psum = 0.f;
#pragma unroll
      for (int t = 0; t < PL; ++t) {
        e[t] = expf(__fsub_rn(xt[t], m));  // pads: exactly 0
        psum = __fadd_rn(psum, e[t]);
      }
      const float denom = group_sum<G>(psum);  // >= 1: the maximum contributes 1
      const float inv = __frcp_rn(denom);
      float qpart = 0.f;
#pragma unroll
      for (int t = 0; t < PL; ++t) {
        e[t] = __fmul_rn(e[t], inv);
        qpart = __fadd_rn(qpart, __fmul_rn(zl[t], e[t]));
      }
      const float q = group_sum<G>(qpart);
      if (valid && (best_a < 0 || q > best_q)) {  // strict > keeps the first maximum
        best_q = q;
        best_a = act;
#pragma unroll
        for (int t = 0; t < PL; ++t) bp[t] = e[t];
      }
    }
    // first maximum over the groups (RA:238-248): highest q, ties to the smaller action
    const int own_a = best_a;
#pragma unroll
    for (int o = G; o < 32; o <<= 1) {
      const float q2 = __shfl_xor_sync(0xffffffffu, best_q, o);
      const int a2 = __shfl_xor_sync(0xffffffffu, best_a, o);
      if (a2 >= 0 && (best_a < 0 || q2 > best_q || (q2 == best_q && a2 < best_a))) {
        best_q = q2;
        best_a = a2;
      }
    }
    if (own_a == best_a && own_a >= 0) {
#pragma unroll
      for (int t = 0; t < PL; ++t)
        if (l + G * t < N) s_bestp[warp][l + G * t] = bp[t];
    }

    // log_softmax of the chosen action's online logits (rainbow_agent.py:262-271):
    // atoms lane, lane + 32 — the layout of the projection below
    float mo = fmaxf(xo[0], xo[1]);
    mo = group_max<32>(mo);
    const float eo0 = expf(__fsub_rn(xo[0], mo)), eo1 = expf(__fsub_rn(xo[1], mo));
    const float den_o = group_sum<32>(__fadd_rn(eo0, eo1));  // pads add exactly 0
    const float lse = logf(den_o);
    const float lgp[2] = {__fsub_rn(__fsub_rn(xo[0], mo), lse),
                          __fsub_rn(__fsub_rn(xo[1], mo), lse)};
    // Bellman support (rainbow_agent.py:229-235)
    const float live = __fsub_rn(1.0f, term);
    const float gwt = __fmul_rn(a.u.cumulative_gamma, live);
    for (int j = lane; j < N; j += 32) s_sup[warp][j] = __fadd_rn(r, __fmul_rn(gwt, z[j]));
    __syncwarp();

    // ---- B. projection (RA:381-494).  hat(i, j) is exactly 0 unless
    // |clip(s_j) - z_i| < dz, and the Bellman atoms s_j = r + g z_j are non-decreasing
    // in j (g >= 0), so the j that reach atom i form one interval: s_j inside
    // (z_i - dz, z_i + dz), i.e. lo < j < hi in units of atoms, open-ended below for
    // the first atom and above for the last (clipping).  Each lane sums the
    // reference's expression in ascending j over [floor(lo), floor(hi) + 1]: one
    // spare position either side, against a rounding error of lo / hi below 1e-4.
    // Skipped terms are exact zeros, so the sum is the dense form's.
    const float *sup = s_sup[warp], *next_p = s_bestp[warp];
    const float z0 = z[0], zlast = z[N - 1];
    const float dz = __fsub_rn(z[1], z0);
    float tg[2] = {0.f, 0.f};
    float ce_part = 0.f, tsum_part = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = lane + 32 * t;
      if (i < N) {
        const float zi = z[i];
        float acc = 0.f;
        if (gwt == 0.f) {
          // terminal row: every s_j is r, so hat(i, .) is one number, and it is 0 for
          // all but the (at most two) atoms next to clip(r)
          const float clipped = fminf(fmaxf(sup[0], z0), zlast);
          const float gap = fabsf(__fsub_rn(clipped, zi));
          if (gap < dz) {
            float hat = __fsub_rn(1.0f, __fdiv_rn(gap, dz));
            hat = fminf(fmaxf(hat, 0.f), 1.f);
#pragma unroll 1
            for (int j = 0; j < N; ++j) acc = __fadd_rn(acc, __fmul_rn(hat, next_p[j]));
          }
        } else {
          int jl = 0, jh = N - 1;
          if (gwt > 0.f) {
            const float inv = __fdividef(1.0f, gwt * dz);
            const float lo = (zi - dz - r - gwt * z0) * inv;
            const float hi = (zi + dz - r - gwt * z0) * inv;
            if (i > 0 && lo > 0.f) jl = min(N - 1, (int)fminf(lo, 1e6f));
            if (i < N - 1 && hi < (float)(N - 2)) jh = max(0, (int)fmaxf(hi, -1e6f) + 1);
          }
#pragma unroll 1
          for (int j = jl; j <= jh; ++j) {
            const float clipped = fminf(fmaxf(sup[j], z0), zlast);
            const float gap = fabsf(__fsub_rn(clipped, zi));
            if (gap < dz) {
              float hat = __fsub_rn(1.0f, __fdiv_rn(gap, dz));
              hat = fminf(fmaxf(hat, 0.f), 1.f);
              acc = __fadd_rn(acc, __fmul_rn(hat, next_p[j]));
            }
          }
        }
        tg[t] = acc;
        if (a.u.target) a.u.target[(size_t)b * N + i] = acc;
        ce_part = __fadd_rn(ce_part, __fmul_rn(acc, lgp[t]));
        tsum_part = __fadd_rn(tsum_part, acc);
      }
    }

    // ---- C. cross entropy (RA:262-271), priority (RA:290), weight (RA:279-280)
    const float ce = -group_sum<32>(ce_part);
    const float tsum = group_sum<32>(tsum_part);
    float w = 1.f;
    if (a.u.sampling_probabilities) {
      const float wmax = __fdiv_rn(1.0f, sqrtf(__fadd_rn(pmin, 1e-10f)));
      const float raw = __fdiv_rn(1.0f, sqrtf(__fadd_rn(my_prob, 1e-10f)));
      w = __fdiv_rn(raw, wmax);
    }
    if (lane == 0) {
      a.u.loss[b] = ce;
      a.u.priorities[b] = sqrtf(__fadd_rn(ce, 1e-10f));
      if (a.u.weights) a.u.weights[b] = w;
      a.weighted[b] = __fmul_rn(w, ce);
    }

    // ---- D. gradient of mean(w * ce) w.r.t. the online logits
    if (a.u.grad_logits) {
      const float scale = __fmul_rn(w, __fdiv_rn(1.0f, (float)rows));
      float *g = a.u.grad_logits + (size_t)b * A * N;
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = lane + 32 * t;
        if (i < N) {
          const float p = __fdiv_rn(t == 0 ? eo0 : eo1, den_o);
          v[t] = __fmul_rn(__fsub_rn(__fmul_rn(p, tsum), tg[t]), scale);
        }
      }
      for (int act = 0; act < A; ++act) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int i = lane + 32 * t;
          if (i < N) g[act * N + i] = act == chosen ? v[t] : 0.f;
        }
      }
    }
  }

  B2R_MARK_END(12);
  // ---- mean weighted loss: the last CTA to finish reduces in a fixed order.
  if (a.u.mean_weighted_loss == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.ticket, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float acc = 0.f;
  for (int k = threadIdx.x; k < a.u.batch; k += blockDim.x)
    acc = __fadd_rn(acc, __ldcg(a.weighted + k));
  acc = group_sum<32>(acc);
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int k = 0; k < kRowWarps; ++k) total = __fadd_rn(total, s_red[k]);
    *a.u.mean_weighted_loss = __fdiv_rn(total, (float)a.u.batch);
    *a.ticket = 0u;  // ready for the next launch
  }
}

float *g_weighted = nullptr;
unsigned int *g_ticket = nullptr;
int g_weighted_cap = 0;

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

}  // extern "C"

namespace b2r {

// Can the write-back of this batch ride at the tail of its loss kernel?
bool c51_can_fuse_writeback(const b2r_c51_args *args, const b2r_tree *tree) {
  // Off unless B2R_FUSE_WRITEBACK=1: measured on one B200 in one process
  // (profiles/r2/README.md), the fence + ticket by which the last CTA finds out that it
  // is the last costs what the kernel boundary it replaces costs (17.4 vs 17.2 us per
  // step at batch 32).
  static const bool on = [] {
    const char *e = std::getenv("B2R_FUSE_WRITEBACK");
    return e != nullptr && std::atoi(e) != 0;
  }();
  return on && tree != nullptr && args->batch <= kTinyBatch && args->batch_count == nullptr &&
         tree_tiny_enabled() && tree->depth + 1 <= 32 && args->num_atoms <= 64;
}

int c51_loss_launch(const b2r_c51_args *args, cudaStream_t s, b2r_tree *tree,
                    const int32_t *indices);

}  // namespace b2r

extern "C" {

int b2r_c51_loss(const b2r_c51_args *args, b2r_stream stream) {
  return b2r::c51_loss_launch(args, as_stream(stream), nullptr, nullptr);
}

}  // extern "C"

namespace b2r {

// tree != nullptr (c51_can_fuse_writeback): also set_priority(indices, priorities).
int c51_loss_launch(const b2r_c51_args *args, cudaStream_t s, b2r_tree *tree,
                    const int32_t *indices) {
  if (!args || args->batch <= 0 || args->num_atoms < 2 || args->num_actions <= 0)
    return fail(B2R_ERR_INVALID_ARGUMENT, "bad C51 shape");
  if (args->batch_count && args->mean_weighted_loss)
    return fail(B2R_ERR_UNSUPPORTED,
                "mean_weighted_loss cannot be combined with batch_count");
  if (!args->support || !args->target_logits || !args->online_logits ||
      !args->actions || !args->rewards || !args->terminals || !args->loss ||
      !args->priorities)
    return fail(B2R_ERR_INVALID_ARGUMENT, "a required C51 pointer is NULL");
  b2r::LossArgs a;
  a.u = *args;
  a.fuse_tree = 0;
  if (tree != nullptr) {
    if (!c51_can_fuse_writeback(args, tree))
      return fail(B2R_ERR_INVALID_ARGUMENT, "this batch cannot fuse its write-back");
    set_tree_window(tree->heap, (size_t)tree->leaves * 16);
    a.fuse_tree = 1;
    a.tree.heap = tree->heap;
    a.tree.depth = tree->depth;
    a.tree.leaves = tree->leaves;
    a.tree.n = args->batch;
    a.tree.padded = 32;
    a.tree.indices = indices;
    a.tree.values = args->priorities;
    a.tree.mode = nullptr;
    a.tree.k_base = 0;
    a.tree.delta = tree->delta;
    a.tree.max_rec = tree->max_rec;
    a.tree.status = tree->status;
    a.tree.n_dev = nullptr;
  }
  // Small batches are latency-bound: one warp per action.  Large batches are
  // throughput-bound: a third of that, so more rows are resident per SM.
  a.warps = args->num_actions < 32 ? args->num_actions : 32;
  if (args->batch > 256) a.warps = (a.warps + 2) / 3;
  int threads = a.warps * 32;
  if (a.fuse_tree && threads < 32 * (tree->depth + 1)) threads = 32 * (tree->depth + 1);
  const size_t smem = ((size_t)a.warps + 5) * args->num_atoms * sizeof(float);
  if (smem > 48 * 1024 || args->num_atoms > 32 * b2r::kMaxAtomsPerLane)
    return fail(B2R_ERR_UNSUPPORTED, "num_atoms above 128 is not supported");
  if (args->batch > b2r::g_weighted_cap) {
    if (b2r::g_weighted) cudaFree(b2r::g_weighted);
    b2r::g_weighted = nullptr;
    int cap = 4096;
    while (cap < args->batch) cap *= 2;
    B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&b2r::g_weighted), (size_t)cap * 4));
    b2r::g_weighted_cap = cap;
  }
  if (!b2r::g_ticket) {
    B2R_CUDA(cudaMalloc(reinterpret_cast<void **>(&b2r::g_ticket), 4));
    B2R_CUDA(cudaMemset(b2r::g_ticket, 0, 4));
  }
  a.weighted = b2r::g_weighted;
  a.ticket = b2r::g_ticket;
  // From 128 rows up: one warp per row (B2R_C51_ROWS_MIN overrides the threshold,
  // B2R_C51_GROUP = 4 | 16 picks another lanes-per-action instance).  Measured with
  // 10 launches per graph (profiles/r1/README.md): at batch 32 the CTA-per-row kernel
  // and the G = 4 instance tie (5.95 us), at 256 the rows kernel wins once the logits
  // no longer sit in L2 (10.2 -> 7.0 us).
  static const int rows_min = [] {
    const char *e = std::getenv("B2R_C51_ROWS_MIN");
    return e ? std::atoi(e) : 128;
  }();
  static const int rows_group = [] {
    const char *e = std::getenv("B2R_C51_GROUP");
    return e ? std::atoi(e) : 8;
  }();
  if (args->batch >= rows_min && args->num_atoms <= b2r::kRowAtoms) {
    const int blocks = (args->batch + b2r::kRowWarps - 1) / b2r::kRowWarps;
    const dim3 grid(blocks), block(b2r::kRowWarps * 32);
    if (rows_group == 16)
      B2R_CUDA(b2r::launch(b2r::c51_loss_rows_kernel<16, 0>, grid, block, 0, s, a));
    else if (rows_group == 4 && args->num_atoms == 51)
      B2R_CUDA(b2r::launch(b2r::c51_loss_rows_kernel<4, 51>, grid, block, 0, s, a));
    else if (args->num_atoms == 51)
      B2R_CUDA(b2r::launch(b2r::c51_loss_rows_kernel<8, 51>, grid, block, 0, s, a));
    else
      B2R_CUDA(b2r::launch(b2r::c51_loss_rows_kernel<8, 0>, grid, block, 0, s, a));
    B2R_LAUNCHED();
    return B2R_OK;
  }
  static const int force_fast = [] {
    const char *e = std::getenv("B2R_C51_FAST");
    return e ? std::atoi(e) : -1;
  }();
  // (the unrolled instance is compiled for at most 384 threads per CTA)
  const bool fast = (force_fast >= 0 ? force_fast != 0 : args->batch > 256) &&
                    threads <= 384;
  if (args->num_atoms <= 64) {
    if (fast)
      B2R_CUDA(b2r::launch(b2r::c51_loss_kernel<2, true>, dim3(args->batch),
                           dim3(threads), smem, s, a));
    else
      B2R_CUDA(b2r::launch(b2r::c51_loss_kernel<2, false>, dim3(args->batch),
                           dim3(threads), smem, s, a));
  } else {
    B2R_CUDA(b2r::launch(b2r::c51_loss_kernel<4, false>, dim3(args->batch),
                         dim3(threads), smem, s, a));
  }
  B2R_LAUNCHED();
  return B2R_OK;
}

}  // namespace b2r

#ifdef B2R_TRACE
extern "C" int b2r_debug_trace_c51(long long *out) {
  return (int)cudaMemcpyFromSymbol(out, b2r::g_trace, sizeof(long long) * 32);
}
#endif
