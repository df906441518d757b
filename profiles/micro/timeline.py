"""One timeline of the fused step: %globaltimer marks of all its kernels.

  B2R_TRACE_GT=1 python -m dopamine_b200.csrc.build --trace --force
  B2R_LIB=profiles/micro/libb200replay_trace.so python profiles/micro/timeline.py 1024

Prints the marks of the LAST replayed step in microseconds since the sampler started."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = {
    'sample': {10: 'S start (CTA 0)', 11: 'S acquired', 12: 'S prologue loads in',
               13: 'S draw + owner', 14: 'S descent done', 15: 'S validity done',
               16: 'S rows written (CTA 0)', 17: 'S last CTA: alone', 18: 'S invalid list',
               19: 'S fix-ups done', 20: 'S END'},
    'gather': {0: 'G start (CTA 0)', 1: 'G acquired', 5: 'G CTA 0 stored', 6: 'G END'},
    'c51': {0: 'L-pre start', 1: 'L-pre acquired', 3: 'L softmaxes', 5: 'L projection',
            8: 'L-pre END', 10: 'L-tail start', 11: 'L-tail acquired', 12: 'L-tail loss END',
            9: 'L-tail tree END', 13: 'L-tail row scalars in', 14: 'L-tail log-softmax done'},
    'tree': {30: 'T tiny start', 7: 'T tiny / one-CTA END',
             0: 'T one-CTA start', 1: 'T one-CTA acquired', 5: 'T one-CTA leaf level done',
             12: 'P presort start',
             14: 'P presort END', 13: 'T apply start', 10: 'T leaf deltas published',
             11: 'T levels released', 15: 'T apply END'},
}


def main():
  batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
  capacity = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
  import torch
  from dopamine_b200 import _native
  wl = bench.GpuWorkload(capacity, batch, 0)
  lib = _native.lib()
  g = torch.cuda.CUDAGraph()
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for _ in range(5):
      wl.step(batch)
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
      wl.step(batch)
    for _ in range(300):
      g.replay()
    s.synchronize()
  events = []
  for name in ('sample', 'gather', 'c51', 'tree'):
    out = (ctypes.c_longlong * 32)()
    fn = getattr(lib, 'b2r_debug_trace_' + name)
    fn.argtypes = [ctypes.c_void_p]
    fn(out)
    for i, v in enumerate(out):
      if v and i in NAMES[name]:
        events.append((int(v), NAMES[name][i]))
  t0 = dict((n, t) for t, n in events).get('S start (CTA 0)', min(t for t, _ in events))
  recent = [(t, n) for t, n in events if t >= t0 - 1000]
  for t, n in sorted(recent):
    print('%9.2f us  %s' % ((t - t0) / 1e3, n))
  stale = [n for t, n in events if t < t0 - 1000]
  if stale:
    print('(not in the last step: %s)' % ', '.join(stale))


if __name__ == '__main__':
  main()
