/*
 * Plain-C restatement of the reference replay hot path, for parity checks at
 * full size (capacity 1M, batch 4096) where the Python port takes minutes.
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  The product never loads this.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no fast-math: fp64 node
 * arithmetic must round exactly like numpy's).
 *
 * Follows (relative to /root/reference/dopamine/replay_memory/):
 *   fo_tree_set_seq ....... sum_tree.py:178-205 applied element by element, the way
 *                           prioritized_replay_buffer.py:213-214 loops over a batch
 *   fo_tree_descend ....... sum_tree.py:126-141
 *   fo_is_valid ........... circular_replay_buffer.py:381-414
 *   fo_gather_u8 .......... circular_replay_buffer.py:516-556 (+338-375 for stacks)
 * Pinned against the Python port and the imported reference in
 * tests/test_oracle_golden.py.
 */
#include <stdint.h>
#include <string.h>

/* heap layout: node (level l, position i) lives at (1<<l) - 1 + i. */

int fo_tree_set_seq(double *heap, int depth, int64_t n, const int64_t *idx,
                    const double *val, double *max_recorded) {
  for (int64_t k = 0; k < n; ++k) {
    double v = val[k];
    if (v < 0.0) return (int)(k + 1); /* ValueError raised at element k */
    if (v > *max_recorded) *max_recorded = v; /* max(value, current) */
    int64_t node = idx[k];
    double delta = v - heap[(((int64_t)1) << depth) - 1 + node];
    for (int l = depth; l >= 0; --l) {
      heap[(((int64_t)1) << l) - 1 + node] += delta;
      node >>= 1;
    }
  }
  return 0;
}

void fo_tree_descend(const double *heap, int depth, int64_t n,
                     const double *mass, int64_t *out) {
  for (int64_t k = 0; k < n; ++k) {
    double q = mass[k];
    int64_t node = 0;
    for (int l = 1; l <= depth; ++l) {
      double left = heap[(((int64_t)1) << l) - 1 + 2 * node];
      if (q < left) {
        node = 2 * node;
      } else {
        node = 2 * node + 1;
        q -= left;
      }
    }
    out[k] = node;
  }
}

static int64_t wrap(int64_t i, int64_t cap) {
  int64_t r = i % cap;
  return r < 0 ? r + cap : r;
}


/* np.sum over `length` f32 products, in numpy's order (probed, numpy 2.3.5):
 * fewer than 8 terms: acc = +0.0f, then left to right; 8..128 terms: numpy's
 * 8-lane unrolled pairwise block (r[j] += a[i+j]; ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7));
 * then the tail left to right). */
static float nstep_sum_f32(const float *discounts, const float *reward, int64_t i,
                           int length, int64_t capacity) {
  if (length < 8) {
    float acc = 0.0f;
    for (int k = 0; k < length; ++k)
      acc = acc + discounts[k] * reward[wrap(i + k, capacity)];
    return acc;
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = discounts[j] * reward[wrap(i + j, capacity)];
  int k = 8;
  for (; k < length - (length % 8); k += 8)
    for (int j = 0; j < 8; ++j)
      r[j] = r[j] + discounts[k + j] * reward[wrap(i + k + j, capacity)];
  float acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; k < length; ++k) acc = acc + discounts[k] * reward[wrap(i + k, capacity)];
  return acc;
}

/* invalid_range is passed explicitly (n_inv entries), exactly as the attribute. */
int fo_is_valid(int64_t index, int64_t capacity, int64_t add_count, int stack,
                int horizon, const int64_t *invalid_range, int n_inv,
                const uint8_t *terminal_nonzero) {
  if (index < 0 || index >= capacity) return 0;
  int64_t cursor = add_count % capacity;
  if (add_count < capacity) {
    if (index >= cursor - horizon) return 0;
    if (index < stack - 1) return 0;
  }
  for (int k = 0; k < n_inv; ++k)
    if (invalid_range[k] == index) return 0;
  for (int k = 1; k < stack; ++k)
    if (terminal_nonzero[wrap(index - k, capacity)]) return 0;
  return 1;
}

/* uint8 frames, scalar int32 action, f32 reward, uint8 terminal. */
void fo_gather_u8(int64_t capacity, int64_t frame_bytes, int stack, int horizon,
                  const float *discounts, const uint8_t *obs,
                  const int32_t *action, const float *reward,
                  const uint8_t *terminal, int64_t n, const int32_t *indices,
                  uint8_t *state, int32_t *out_action, float *out_return,
                  uint8_t *next_state, int32_t *next_action, float *next_reward,
                  uint8_t *out_terminal, int32_t *out_indices) {
  for (int64_t b = 0; b < n; ++b) {
    int64_t i = indices[b];
    int length = horizon, ends = 0;
    for (int k = 0; k < horizon; ++k) {
      if (terminal[wrap(i + k, capacity)]) {
        ends = 1;
        length = k + 1;
        break;
      }
    }
    float acc = nstep_sum_f32(discounts, reward, i, length, capacity);
    int64_t nxt = wrap(i + length, capacity);
    for (int s = 0; s < stack; ++s) {
      const uint8_t *fa = obs + wrap(i - stack + 1 + s, capacity) * frame_bytes;
      const uint8_t *fb = obs + wrap(nxt - stack + 1 + s, capacity) * frame_bytes;
      uint8_t *da = state + b * frame_bytes * stack + s;
      uint8_t *db = next_state + b * frame_bytes * stack + s;
      for (int64_t p = 0; p < frame_bytes; ++p) {
        da[p * stack] = fa[p];
        db[p * stack] = fb[p];
      }
    }
    out_action[b] = action[i];
    out_return[b] = acc;
    next_action[b] = action[nxt];
    next_reward[b] = reward[nxt];
    out_terminal[b] = (uint8_t)ends;
    out_indices[b] = (int32_t)i;
  }
}
