"""Per-kernel device times of the hot path, each kernel replayed alone in a CUDA
graph (CUDA events on the launching stream), beside the fused / unfused step.

  python profiles/micro/kernel_times.py [--batches 32,256,1024,4096] [--capacity N]

Prints one JSON line per batch size.  The sum of the parts is an upper bound for the
unfused step; the fused step overlaps the frame copies with loss + write-back.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--batches', default='32,256,1024,4096')
  p.add_argument('--capacity', type=int, default=1000000)
  p.add_argument('--reps', type=int, default=200)
  p.add_argument('--per-graph', type=int, default=1,
                 help='calls captured per CUDA graph (1: each replay is one launch and '
                      'the time is quantised by the ~2 us graph-launch pacing)')
  a = p.parse_args()
  batches = [int(b) for b in a.batches.split(',')]
  import torch
  wl = bench.GpuWorkload(a.capacity, max(batches), 0)
  nat, lib = wl.native, wl.lib
  for batch in batches:
    t, b, c = wl.plan(batch)
    stream = nat.current_stream

    def sample():
      nat.check(lib.b2r_sample_indices_device(wl.h, batch, wl.seed, 0,
                                              t['indices'].data_ptr(), stream()))

    def gather():
      nat.check(lib.b2r_gather_device(wl.h, batch, t['indices'].data_ptr(),
                                      ctypes.byref(b), stream()))

    def loss():
      nat.check(lib.b2r_c51_loss(ctypes.byref(c), stream()))

    def write_back():
      nat.check(lib.b2r_set_priority_device(wl.h, batch, t['indices'].data_ptr(),
                                            t['priorities'].data_ptr(), stream()))

    def fused():
      wl.fused = True
      wl.step(batch)

    def unfused():
      wl.fused = False
      wl.step(batch)

    def fused_deferred():
      wl.fused = True
      wl.step(batch)

    # the latency chain alone (fused call without the frame outputs) and the sampler
    # followed by the frame copies alone
    b_noframes = type(b).from_buffer_copy(b)
    b_noframes.state = None
    b_noframes.next_state = None

    def chain_only():
      nat.check(lib.b2r_train_step_device(wl.h, batch, wl.seed, 0,
                                          ctypes.byref(b_noframes), ctypes.byref(c),
                                          stream()))

    def sample_gather():
      nat.check(lib.b2r_sample_transition_batch_device(wl.h, batch, wl.seed, 0,
                                                       ctypes.byref(b), stream()))

    wl.fused = False
    wl.step(batch)  # valid indices / priorities in the plan's buffers
    torch.cuda.synchronize()
    row = {'batch': batch}
    reps = max(20, a.reps // max(1, batch // 256))
    for name, fn in (('sample_us', sample), ('gather_us', gather),
                     ('c51_loss_us', loss), ('write_back_us', write_back),
                     ('step_unfused_us', unfused), ('step_fused_us', fused),
                     ('chain_only_us', chain_only), ('sample_gather_us', sample_gather),
                     ('step_fused_deferred_us', fused_deferred)):
      deferred = fn is fused_deferred
      if deferred:
        wl.set_deferred(True)
      ms = bench.time_graph_or_eager(torch, fn, reps, 5, True,
                                     per_graph=bench.steps_per_graph(reps, a.per_graph),
                                     finish=wl.join if deferred else None)
      if deferred:
        wl.set_deferred(False)
      row[name] = round(ms * 1e3 / reps, 2)
    row['parts_sum_us'] = round(row['sample_us'] + row['gather_us'] +
                                row['c51_loss_us'] + row['write_back_us'], 2)
    nat.check(lib.b2r_check(wl.h, stream()))
    print(json.dumps(row), flush=True)


if __name__ == '__main__':
  main()
