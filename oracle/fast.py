"""ctypes front-end of oracle/fast_oracle.c.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
  subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
  global _LIB
  if _LIB is None:
    path = os.path.join(_HERE, '_build', 'libfast_oracle.so')
    if not os.path.exists(path):
      build()
    _LIB = ctypes.CDLL(path)
    _LIB.fo_tree_set_seq.restype = ctypes.c_int
    _LIB.fo_is_valid.restype = ctypes.c_int
  return _LIB


def _p(a):
  return a.ctypes.data_as(ctypes.c_void_p)


class FastTree(object):
  """Flat-heap fp64 sum tree driven by the C restatement."""

  def __init__(self, capacity):
    self.depth = int(np.ceil(np.log2(capacity)))
    self.heap = np.zeros((1 << (self.depth + 1)) - 1, dtype=np.float64)
    self.max_recorded = np.array([1.0], dtype=np.float64)

  def level(self, l):
    return self.heap[(1 << l) - 1:(1 << (l + 1)) - 1]

  def set_seq(self, idx, val):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    return lib().fo_tree_set_seq(_p(self.heap), ctypes.c_int(self.depth),
                                 ctypes.c_int64(len(idx)), _p(idx), _p(val),
                                 _p(self.max_recorded))

  def descend(self, mass):
    mass = np.ascontiguousarray(mass, dtype=np.float64)
    out = np.empty(len(mass), dtype=np.int64)
    lib().fo_tree_descend(_p(self.heap), ctypes.c_int(self.depth),
                          ctypes.c_int64(len(mass)), _p(mass), _p(out))
    return out


def is_valid(index, capacity, add_count, stack, horizon, invalid_range,
             terminal_nonzero):
  inv = np.ascontiguousarray(invalid_range, dtype=np.int64)
  return bool(lib().fo_is_valid(
      ctypes.c_int64(int(index)), ctypes.c_int64(capacity),
      ctypes.c_int64(int(add_count)), ctypes.c_int(stack),
      ctypes.c_int(horizon), _p(inv), ctypes.c_int(len(inv)),
      _p(terminal_nonzero)))


def gather_u8(capacity, frame_bytes, stack, horizon, discounts, obs, action,
              reward, terminal, indices):
  """Returns the 8 reference outputs for uint8 frames / scalar columns."""
  n = len(indices)
  indices = np.ascontiguousarray(indices, dtype=np.int32)
  discounts = np.ascontiguousarray(discounts, dtype=np.float32)
  state = np.empty((n, frame_bytes, stack), dtype=np.uint8)
  nstate = np.empty((n, frame_bytes, stack), dtype=np.uint8)
  act = np.empty(n, dtype=np.int32)
  ret = np.empty(n, dtype=np.float32)
  nact = np.empty(n, dtype=np.int32)
  nrew = np.empty(n, dtype=np.float32)
  term = np.empty(n, dtype=np.uint8)
  idx = np.empty(n, dtype=np.int32)
  lib().fo_gather_u8(
      ctypes.c_int64(capacity), ctypes.c_int64(frame_bytes),
      ctypes.c_int(stack), ctypes.c_int(horizon), _p(discounts), _p(obs),
      _p(action), _p(reward), _p(terminal), ctypes.c_int64(n), _p(indices),
      _p(state), _p(act), _p(ret), _p(nstate), _p(nact), _p(nrew), _p(term),
      _p(idx))
  return state, act, ret, nstate, nact, nrew, term, idx
