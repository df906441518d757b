"""Where the time of one e2e_host_batch update goes (bench.py:measure_e2e_host_batch):
each stage timed on the host clock with a device synchronize behind it, so the stages
add up to MORE than the pipelined loop — the ranking is what matters.

  python profiles/micro/host_batch_breakdown.py [batch]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
  import torch
  wl = bench.GpuWorkload(1000000, batch, 0)
  mem, ra = wl.mem, wl.ra
  online_h = wl.online[:batch].cpu().pin_memory()
  target_h = wl.target[:batch].cpu().pin_memory()
  online_d = torch.empty_like(wl.online[:batch])
  target_d = torch.empty_like(wl.target[:batch])
  acc = {}

  def timed(name, fn):
    t0 = time.perf_counter()
    out = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += t1 - t0
    a[1] += t2 - t0
    return out

  reps = 300
  for k in range(reps + 20):
    if k == 20:
      acc.clear()
    idx = timed('sample_index_batch', lambda: mem.sample_index_batch(batch))
    bnp = timed('gather to host (numpy views)',
                lambda: mem.sample_transition_batch(batch, indices=None)
                if False else mem._gather_to_host(batch, None, idx))
    timed('logits H2D x2', lambda: (online_d.copy_(online_h, non_blocking=True),
                                    target_d.copy_(target_h, non_blocking=True)))
    dev = lambda a: torch.as_tensor(a).cuda(non_blocking=True)
    cols = timed('4 scalar columns H2D', lambda: [dev(bnp[i]) for i in (1, 2, 6, 8)])
    out = timed('c51_loss', lambda: ra.c51_loss(online_d, target_d, cols[0], cols[1], cols[2],
                                                cols[3], wl.support, wl.gamma_n))
    prio = timed('priorities D2H', lambda: out['priorities'].cpu().numpy())
    timed('set_priority (numpy)', lambda: mem.set_priority(bnp[7], prio))
  print('stage: host-only us / with device sync us (per update, batch %d)' % batch)
  tot = [0.0, 0.0]
  for name, (h, d) in acc.items():
    print('  %-32s %7.1f %7.1f' % (name, h / reps * 1e6, d / reps * 1e6))
    tot[0] += h
    tot[1] += d
  print('  %-32s %7.1f %7.1f' % ('sum', tot[0] / reps * 1e6, tot[1] / reps * 1e6))


if __name__ == '__main__':
  main()
