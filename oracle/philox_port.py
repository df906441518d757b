"""numpy restatement of the device RNG of the throughput mode.  TEST INFRASTRUCTURE.

The reference draws from Python's `random` (sum_tree.py:123, 162-165); the
`rng='device'` mode of this build draws the same quantities from Philox4x32-10
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) keyed by the
seed, with the counter (draw number, stream offset).  This port lets the tests feed
the SAME uniforms to the CPU statement of the sampling rules, so the device-RNG path
is checked for exact indices, not only for properties.
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
  c0, c1, c2, c3 = [int(x) & MASK for x in counter]
  k0, k1 = [int(x) & MASK for x in key]
  for _ in range(10):
    p0, p1 = M0 * c0, M1 * c2
    c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, \
                     ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
    k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
  return c0, c1, c2, c3


def uniform53(seed, offset, n):
  """53-bit uniform in [0, 1) of draw `n` of stream (seed, offset), as
  philox_uniform53 in dopamine_b200/csrc/common.cuh."""
  seed, offset, n = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), int(n)
  o = philox4x32_10((n & MASK, n >> 32, offset & MASK, offset >> 32),
                    (seed & MASK, seed >> 32))
  a, b = o[0] >> 5, o[1] >> 6
  return np.float64((a << 26) | b) * np.float64(1.0 / 9007199254740992.0)


def stratified_queries(seed, draw_offset, batch):
  """query01[i] = lo_i + (hi_i - lo_i) * u_i with lo_i = i * (1 / batch) (the last
  hi is 1.0): random.uniform over np.linspace(0, 1, batch + 1) strata."""
  step = np.float64(1.0) / np.float64(batch)
  out = np.empty(batch, dtype=np.float64)
  for i in range(batch):
    lo = np.float64(i) * step
    hi = np.float64(1.0) if i + 1 == batch else np.float64(i + 1) * step
    out[i] = lo + (hi - lo) * uniform53(seed, draw_offset, i)
  return out


def retry_uniforms(seed, rank, draw_offset, batch, count):
  """The rank-private retry stream: draws batch .. batch + count - 1 of the stream
  keyed by seed + golden * (rank + 1)."""
  key = (int(seed) + 0x9E3779B97F4A7C15 * (rank + 1)) & (2**64 - 1)
  return np.array([uniform53(key, draw_offset, batch + r) for r in range(count)],
                  dtype=np.float64)
