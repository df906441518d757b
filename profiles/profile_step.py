"""Runs a few eager hot-path steps between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` sees only the step kernels.

  python profiles/profile_step.py --batch 32 --steps 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--batch', type=int, default=32)
  p.add_argument('--steps', type=int, default=3)
  p.add_argument('--capacity', type=int, default=1000000)
  a = p.parse_args()
  import torch
  wl = bench.GpuWorkload(a.capacity, a.batch, 0)
  for _ in range(3):
    wl.step(a.batch)
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  for _ in range(a.steps):
    wl.step(a.batch)
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()
  print('profiled', a.steps, 'steps of batch', a.batch)


if __name__ == '__main__':
  main()
