"""CPU restatement of the reference sum tree.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this; the product (`dopamine_b200`) never does.

Restates `/root/reference/dopamine/replay_memory/sum_tree.py`:
  * constructor / level sizes ......... sum_tree.py:65-89
  * total priority .................... sum_tree.py:91-97
  * sample (root-to-leaf descent) ..... sum_tree.py:99-141
  * stratified_sample ................. sum_tree.py:143-166
  * get / set (delta propagation) ..... sum_tree.py:168-205

Parity is PINNED: `tests/test_oracle_golden.py` checks this port against the
reference's own known-answer tests (tests/dopamine/replay_memory/sum_tree_test.py)
and against fixtures produced by running the unmodified reference
(`oracle/make_golden.py` -> `tests/golden/*.npz`).

Storage differs from the reference on purpose (one flat fp64 heap array, node
(l, i) at 2**l - 1 + i, the same layout the CUDA tree uses); the arithmetic per
node is identical: fp64, `leaf_delta = value - leaf`, `node += leaf_delta` on
every level, strict `<` against the left child during descent.
"""
import math
import random as _py_random

import numpy as np


class PortSumTree(object):
  """fp64 sum tree over a flat heap array."""

  def __init__(self, capacity, rng=None):
    if not isinstance(capacity, int):
      raise AssertionError('capacity must be an int')
    if capacity <= 0:
      raise ValueError(
          'Sum tree capacity should be positive. Got: {}'.format(capacity))
    self.depth = int(math.ceil(np.log2(capacity)))  # sum_tree.py:81
    self.heap = np.zeros((1 << (self.depth + 1)) - 1, dtype=np.float64)
    self.max_recorded_priority = 1.0  # sum_tree.py:89
    self._rng = rng if rng is not None else _py_random

  # -- layout helpers -------------------------------------------------------
  def level(self, l):
    base = (1 << l) - 1
    return self.heap[base:base + (1 << l)]

  @property
  def nodes(self):
    return [self.level(l) for l in range(self.depth + 1)]

  def total(self):
    return self.heap[0]

  # -- sum_tree.py:168-205 ---------------------------------------------------
  def get(self, index):
    return self.level(self.depth)[index]

  def set(self, index, value):
    if value < 0.0:
      raise ValueError(
          'Sum tree values should be nonnegative. Got {}'.format(value))
    self.max_recorded_priority = max(value, self.max_recorded_priority)
    leaf_base = (1 << self.depth) - 1
    delta = value - self.heap[leaf_base + index]
    node = index
    for l in range(self.depth, -1, -1):
      self.heap[(1 << l) - 1 + node] += delta
      node //= 2

  # -- sum_tree.py:99-141 ----------------------------------------------------
  def descend(self, mass):
    """Leaf reached by a query already scaled to [0, root)."""
    node = 0
    for l in range(1, self.depth + 1):
      left = self.heap[(1 << l) - 1 + 2 * node]
      if mass < left:
        node = 2 * node
      else:
        node = 2 * node + 1
        mass -= left
    return node

  def sample(self, query_value=None):
    if self.total() == 0.0:
      raise Exception('Cannot sample from an empty sum tree.')
    if query_value and (query_value < 0. or query_value > 1.):
      raise ValueError('query_value must be in [0, 1].')
    u = self._rng.random() if query_value is None else query_value
    return self.descend(u * self.total())

  # -- sum_tree.py:143-166 ---------------------------------------------------
  def stratified_queries(self, batch_size):
    """The batch_size query values in [0, 1] the reference would draw."""
    edges = np.linspace(0., 1., batch_size + 1)
    return [edges[i] + (edges[i + 1] - edges[i]) * self._rng.random()
            for i in range(batch_size)]

  def stratified_sample(self, batch_size):
    if self.total() == 0.0:
      raise Exception('Cannot sample from an empty sum tree.')
    return [self.sample(query_value=q)
            for q in self.stratified_queries(batch_size)]
