"""Phase timestamps (clock64, CTA 0 / thread 0) of the four step kernels.

  B2R_LIB=profiles/micro/libb200replay_trace.so python profiles/micro/trace_step.py
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
  import torch
  from dopamine_b200 import _native
  wl = bench.GpuWorkload(200000, batch, 0)
  wl.fused = os.environ.get('B2R_UNFUSED') is None
  lib = _native.lib()
  g = torch.cuda.CUDAGraph()
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for _ in range(5):
      wl.step(batch)
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
      wl.step(batch)
    for _ in range(20):
      g.replay()
    s.synchronize()
  for name in ('sample', 'gather', 'c51', 'tree'):
    out = (ctypes.c_longlong * 32)()
    fn = getattr(lib, 'b2r_debug_trace_' + name)
    fn.argtypes = [ctypes.c_void_p]
    fn(out)
    marks = {i: v for i, v in enumerate(out) if v}
    if marks:
      t0 = min(marks.values())
      print(name, 'marks {index: cycles since the first}:',
            {i: int(v - t0) for i, v in marks.items()})


if __name__ == '__main__':
  main()
