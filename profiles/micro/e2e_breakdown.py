"""Where does an end-to-end update go?  Times the pieces of bench.py's e2e loop
(batch 32): adds only, trainer steps only (pipelined / synchronous), both."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import bench  # noqa: E402


def main():
  import torch
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
  wl = bench.GpuWorkload(1000000, 32, 0)
  mem, ra = wl.mem, wl.ra
  frames = np.random.RandomState(3).randint(0, 256, size=(64, 84, 84)).astype(np.uint8)
  online_h = wl.online[:32].cpu().pin_memory()
  target_h = wl.target[:32].cpu().pin_memory()
  op, tp = online_h.data_ptr(), target_h.data_ptr()
  stream = wl.native.current_stream()
  sentinel = prb.MAX_RECORDED_PRIORITY
  k = [0]

  def adds(n=4):
    for _ in range(n):
      i = k[0]
      k[0] += 1
      mem.add(frames[i & 63], i % 18, 0.5, int(i % 1000 == 999), sentinel)

  def timed(name, fn, steps=3000, after=None):
    for _ in range(50):
      fn()
    if after:
      after()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
      fn()
    host = time.perf_counter() - t0
    if after:
      after()
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print('%-44s host %.1f us/iter, with drain %.1f us/iter' % (
        name, host / steps * 1e6, total / steps * 1e6), flush=True)

  timed('4 x add() + flush', lambda: (adds(), mem._flush()))
  timed('4 x add() only (flush every 32 iters)', adds)
  import gc
  for depth in (2, 0):
    # one trainer per loop: with B2R_HOST_TRACE=1 each prints its own host segments
    tr = ra.ReplayTrainer(mem, 18, 51, 10., batch_size=32, pipeline_depth=depth, seed=1)
    timed('trainer.step only, depth %d' % depth,
          lambda: tr.step_pointers(op, tp, stream), after=tr.drain)
    del tr
    gc.collect()
    tr = ra.ReplayTrainer(mem, 18, 51, 10., batch_size=32, pipeline_depth=depth, seed=1)
    timed('4 x add() + trainer.step, depth %d' % depth,
          lambda: (adds(), tr.step_pointers(op, tp, stream)), after=tr.drain)
    del tr
    gc.collect()


if __name__ == '__main__':
  main()
