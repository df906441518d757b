"""Timeline of ONE shard's fused step (%globaltimer marks), ranks emulated on one GPU.

  B2R_TRACE_GT=1 python -m dopamine_b200.csrc.build --trace --force
  B2R_LIB=profiles/micro/libb200replay_trace.so python profiles/micro/timeline_sharded.py 32 2

Two (or `world`) shards on one device, wired to each other by raw pointers
(PeerExchange.emulated); every replay publishes all totals and then runs the ranks' steps
one after the other, so the marks left behind are those of the LAST rank's step: the
sharded kernels as they run at N > 1, minus the NVLink wait."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import timeline  # noqa: E402


def main():
  per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 32
  world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
  capacity = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
  import torch
  from dopamine_b200 import _native
  from dopamine_b200.replay_memory import sharded_replay
  lib = _native.lib()
  wls = [bench.GpuWorkload(capacity, per_gpu * world, g) for g in range(world)]
  xs = sharded_replay.PeerExchange.emulated(world)
  steps = [sharded_replay.ShardedStep(wls[g], per_gpu * world, world, g, None, exchange=xs[g])
           for g in range(world)]

  def one():
    for g in range(world):
      xs[g].publish(wls[g].mem)
    for g in range(world):
      steps[g].step()

  g = torch.cuda.CUDAGraph()
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    for _ in range(5):
      one()
    s.synchronize()
    with torch.cuda.graph(g, stream=s):
      one()
    for _ in range(200):
      g.replay()
    s.synchronize()
  for w in wls:
    _native.check(lib.b2r_check(w.h, _native.current_stream()))
  events = []
  for name in ('sample', 'gather', 'c51', 'tree'):
    out = (ctypes.c_longlong * 32)()
    fn = getattr(lib, 'b2r_debug_trace_' + name)
    fn.argtypes = [ctypes.c_void_p]
    fn(out)
    for i, v in enumerate(out):
      if v and i in timeline.NAMES[name]:
        events.append((int(v), timeline.NAMES[name][i]))
  t0 = dict((n, t) for t, n in events).get('S start (CTA 0)', min(t for t, _ in events))
  for t, n in sorted((t, n) for t, n in events if t >= t0 - 1000):
    print('%9.2f us  %s' % ((t - t0) / 1e3, n))
  print('rows of the last rank:', int(steps[-1].sharded._count.item()),
        'max_rows', steps[-1].max_rows)


if __name__ == '__main__':
  main()
