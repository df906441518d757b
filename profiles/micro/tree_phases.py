"""Phases of the early write-back kernel run ALONE (b2r_tree_set with
B2R_TREE_SET_PHASE=3, nothing else on the GPU): what its index-only half costs without
neighbours.

  B2R_TRACE_GT=1 python -m dopamine_b200.csrc.build --trace
  B2R_LIB=profiles/micro/libb200replay_trace.so python profiles/micro/tree_phases.py 1024
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ['B2R_TREE_SET_PHASE'] = '3'

NAMES = ['start', 'indices in', 'analysed', 'grouped by leaf', 'duplicates counted',
         'list made', 'grouped by node', 'nodes fetched', 'leaves landed', 'said so',
         'before the wait', 'parent ended']


def main():
  batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  from dopamine_b200 import _native
  from dopamine_b200.replay_memory import sum_tree
  lib = _native.lib()
  cap = 1 << 20
  rng = np.random.RandomState(0)
  tree = sum_tree.SumTree(cap)
  for lo in range(0, cap, 65536):
    tree.set_batch(np.arange(lo, lo + 65536), (0.5 + rng.rand(65536)).astype(np.float32))
  for rep in range(20):
    idx = np.sort(rng.randint(0, cap, size=batch)).astype(np.int64)
    idx[rng.choice(batch, size=max(1, batch // 300), replace=False)] = rng.randint(
        0, cap, size=max(1, batch // 300))
    tree.set_batch(idx, np.abs(rng.randn(batch)).astype(np.float32))
  if not hasattr(lib, 'b2r_debug_trace_tree_phases'):
    return  # (not a trace build: the launches were the point, e.g. under ncu)
  ph = (ctypes.c_longlong * 32)()
  lib.b2r_debug_trace_tree_phases.argtypes = [ctypes.c_void_p]
  lib.b2r_debug_trace_tree_phases(ph)
  for who, label in ((0, 'leaf CTA'), (1, 'CTA of level 10')):
    t0 = ph[who * 16]
    print('%s (us since its start): %s' % (label, ', '.join(
        '%s %.2f' % (NAMES[i], (ph[who * 16 + i] - t0) / float(os.environ.get('B2R_TICKS_PER_US', '1e3')))
        for i in range(1, len(NAMES)) if ph[who * 16 + i])))
  cnt = (ctypes.c_ulonglong * 8)()
  lib.b2r_debug_trace_tree_counts.argtypes = [ctypes.c_void_p]
  lib.b2r_debug_trace_tree_counts(cnt)
  print('early write-backs: %d launched, %d with own lists, %d without a hand-over' % (
      cnt[0], cnt[1], cnt[2]))


if __name__ == '__main__':
  main()
