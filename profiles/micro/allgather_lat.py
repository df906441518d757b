"""Latency of the shard-total all-gather vs the rest of the sharded step (torchrun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('NCCL_DEBUG', 'WARN')
import torch
import torch.distributed as dist
import bench


def timed(fn, reps=200, graph=True):
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(s):
    for _ in range(5):
      fn()
    s.synchronize()
    run = fn
    if graph:
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=s):
        fn()
      run = g.replay
      for _ in range(3):
        run()
      s.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
      run()
    b.record(s)
    b.synchronize()
    return a.elapsed_time(b) * 1000 / reps


def main():
  rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
  torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
  dist.init_process_group('nccl')
  from dopamine_b200.replay_memory import sharded_replay
  wl = bench.GpuWorkload(200000, 32 * world, rank)
  step = sharded_replay.ShardedStep(wl, 32 * world, world, rank, dist)
  send = torch.ones(1, dtype=torch.float64, device='cuda') * (rank + 1)
  out = torch.empty(world, dtype=torch.float64, device='cuda')
  t_ag = timed(lambda: dist.all_gather_into_tensor(out, send))
  t_step = timed(step.step)
  totals = step.sharded.totals().clone()
  step.sharded.totals = lambda: totals  # no collective
  t_nocoll = timed(step.step)
  cnt = int(step.sharded._count.cpu()[0])
  if rank == 0:
    print('world %d: all_gather %.1f us, sharded step %.1f us, step without collective '
          '%.1f us, local count %d' % (world, t_ag, t_step, t_nocoll, cnt))
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
