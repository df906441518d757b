"""GPU sum tree behind the reference's `SumTree` interface.

Drop-in for `dopamine/replay_memory/sum_tree.py` (class SumTree, sum_tree.py:30):
same constructor, `set`, `get`, `sample`, `stratified_sample`,
`max_recorded_priority`, `nodes`, `_total_priority`, same exceptions and messages.
The nodes are fp64 in HBM and every operation is a CUDA kernel reached through the
C ABI (`include/b200_replay.h`); node values are bit-identical to the reference's
because batched updates apply their deltas per node in the reference's order.

Random draws come from Python's `random` module exactly as in the reference
(sum_tree.py:123, 162-165), so seeding `random` reproduces its samples.
"""
import ctypes
import random

import numpy as np

from dopamine_b200 import _native


class SumTree(object):
  """A sum tree over `capacity` leaves stored on the GPU."""

  def __init__(self, capacity, _handle=None, _before_read=None):
    assert isinstance(capacity, int)
    if capacity <= 0:
      raise ValueError('Sum tree capacity should be positive. Got: {}'.
                       format(capacity))
    self._lib = _native.lib()
    self._capacity = capacity
    self._owned = _handle is None
    self._before_read = _before_read
    if _handle is None:
      handle = ctypes.c_void_p()
      _native.check(self._lib.b2r_tree_create(capacity, ctypes.byref(handle)))
      _handle = handle
    self._h = _handle
    self._depth = self._lib.b2r_tree_depth(self._h)

  def __del__(self):
    if getattr(self, '_owned', False) and getattr(self, '_h', None):
      self._lib.b2r_tree_destroy(self._h)
      self._h = None

  # -- plumbing ---------------------------------------------------------------
  def _sync_point(self):
    if self._before_read is not None:
      self._before_read()

  @staticmethod
  def _stream():
    return _native.current_stream()

  # -- reference attributes ---------------------------------------------------
  @property
  def nodes(self):
    """List of per-level fp64 arrays (a host copy), like `SumTree.nodes`."""
    self._sync_point()
    levels = []
    for level in range(self._depth + 1):
      out = np.empty(1 << level, dtype=np.float64)
      _native.check(self._lib.b2r_tree_read_level(
          self._h, level, _native.ptr(out), self._stream()))
      levels.append(out)
    return levels

  @property
  def max_recorded_priority(self):
    self._sync_point()
    out = ctypes.c_double()
    _native.check(self._lib.b2r_tree_max_recorded(
        self._h, ctypes.byref(out), self._stream()))
    return out.value

  @max_recorded_priority.setter
  def max_recorded_priority(self, value):
    self._sync_point()
    _native.check(self._lib.b2r_tree_set_max_recorded(
        self._h, float(value), self._stream()))

  def _total_priority(self):
    """sum_tree.py:91-97."""
    self._sync_point()
    out = ctypes.c_double()
    _native.check(self._lib.b2r_tree_total(
        self._h, ctypes.byref(out), self._stream()))
    return np.float64(out.value)

  # -- sampling (sum_tree.py:99-166) --------------------------------------------
  def _descend(self, queries):
    queries = np.ascontiguousarray(queries, dtype=np.float64)
    out = np.empty(len(queries), dtype=np.int64)
    status = self._lib.b2r_tree_sample(self._h, len(queries),
                                       _native.ptr(queries), _native.ptr(out),
                                       self._stream())
    if status == _native.ERR_EMPTY_TREE:
      raise Exception('Cannot sample from an empty sum tree.')
    _native.check(status)
    return out

  def sample(self, query_value=None):
    self._sync_point()
    if query_value and (query_value < 0. or query_value > 1.):
      # The reference tests emptiness first; an invalid query on an empty tree
      # reports the empty tree.
      if self._total_priority() == 0.0:
        raise Exception('Cannot sample from an empty sum tree.')
      raise ValueError('query_value must be in [0, 1].')
    state = random.getstate()
    query = random.random() if query_value is None else query_value
    try:
      return int(self._descend([query])[0])
    except Exception:
      random.setstate(state)  # the reference raises before drawing
      raise

  def stratified_sample(self, batch_size):
    self._sync_point()
    state = random.getstate()
    bounds = np.linspace(0., 1., batch_size + 1)
    assert len(bounds) == batch_size + 1
    queries = [random.uniform(bounds[i], bounds[i + 1])
               for i in range(batch_size)]
    try:
      return [int(i) for i in self._descend(queries)]
    except Exception:
      random.setstate(state)
      raise

  # -- get / set (sum_tree.py:168-205) -------------------------------------------
  def get(self, node_index):
    self._sync_point()
    idx = np.array([node_index], dtype=np.int64)
    out = np.empty(1, dtype=np.float64)
    status = self._lib.b2r_tree_get(self._h, 1, _native.ptr(idx),
                                    _native.ptr(out), self._stream())
    if status == _native.ERR_INDEX_RANGE:
      raise IndexError(_native.last_error())
    _native.check(status)
    return out[0]

  def set(self, node_index, value):
    if value < 0.0:
      raise ValueError('Sum tree values should be nonnegative. Got {}'.
                       format(value))
    self.set_batch([node_index], [value])

  def set_batch(self, indices, values):
    """`for i, v in zip(indices, values): self.set(i, v)` in one call.

    Sequential semantics are kept (duplicates chain, ancestors accumulate the
    deltas in order), as prioritized_replay_buffer.py:213-214 relies on.
    """
    self._sync_point()
    idx = _native.as_i64(indices)
    val = np.ascontiguousarray(values, dtype=np.float64)
    assert idx.shape == val.shape and idx.ndim == 1
    bad = ctypes.c_int64(-1)
    status = self._lib.b2r_tree_set(self._h, len(idx), _native.ptr(idx),
                                    _native.ptr(val), ctypes.byref(bad),
                                    self._stream())
    if status == _native.ERR_NEGATIVE_PRIORITY:
      raise ValueError('Sum tree values should be nonnegative. Got {}'.
                       format(values[bad.value]))
    if status == _native.ERR_INDEX_RANGE:
      raise IndexError(_native.last_error())
    _native.check(status)
