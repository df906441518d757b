"""Benchmark of the replay-and-update hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one batch: prioritized stratified sample
-> fused frame-stack / n-step gather -> fused C51 target + cross-entropy + new
priorities -> batched priority write-back, on synthetic Atari-shaped transitions
(SURVEY.md section 8d).  Workload at N=1: BASELINE.json configs[1] — prioritized
replay capacity 1M, update_horizon 3, gamma 0.99, batch 32, 51 atoms, 18 actions.
With N>1 (torchrun) every rank owns a 1M shard and the global batch is 32*N
(weak scaling); the only collective is the all-gather of shard totals.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU
algorithm (oracle port, Python) on the host cores for the same workload.
"""
import argparse
import ctypes
import json
import math
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
  os.environ['NCCL_DEBUG'] = 'WARN'  # keep stdout to the one JSON line

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

FRAME = 84 * 84
NUM_ACTIONS = 18
NUM_ATOMS = 51
VMAX = 10.0
GAMMA = 0.99
HORIZON = 3
STACK = 4
METRIC = ('sampled transitions/sec (PER sample+gather+C51 target+priority '
          'update)')
UNIT = 'transitions/s'
# batches whose frame copies are deferred (see main): below, the chain bounds the step and
# the copies hide behind it either way (measured: 14.5 us joined, 15.0 deferred at 32);
# above, the copies saturate HBM and slow the chain beside them
DEFER_MIN_BATCH = 128
DEFER_MAX_BATCH = 2048
# shortest timed work the headline number may rest on (see time_graph_or_eager)
MIN_TIMED_MS = 50.0


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=20000)
  p.add_argument('--warmup', type=int, default=100)
  p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  p.add_argument('--batch', type=int, default=32)
  p.add_argument('--capacity', type=int, default=1000000)
  p.add_argument('--no-graph', action='store_true',
                 help='launch eagerly instead of replaying a CUDA graph')
  p.add_argument('--steps-per-graph', type=int, default=10,
                 help='consecutive steps captured in one CUDA graph (the largest '
                      'divisor of --steps up to this is used); 1 = one graph launch '
                      'per step')
  p.add_argument('--no-sweep', action='store_true')
  p.add_argument('--no-early-publish', action='store_true',
                 help='N > 1: publish shard totals from the sampler only')
  p.add_argument('--no-defer', action='store_true',
                 help='join the frame copies into the stream after every step')
  p.add_argument('--no-cpu-baseline', action='store_true')
  p.add_argument('--no-e2e', action='store_true')
  p.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'],
                 help='N>1: shard totals over peer memory inside the sampling '
                      'kernel (default) or an NCCL all-gather before it')
  p.add_argument('--unfused', action='store_true',
                 help='four separate calls (sample+gather, loss, write-back) on one '
                      'stream instead of b2r_train_step_device')
  return p.parse_args()


def workload_name(batch, capacity, n_gpus):
  name = ('Rainbow prioritized replay capacity {}, update_horizon 3, gamma 0.99, '
          'batch {}, 51-atom C51 projection'.format(capacity, batch))
  if n_gpus > 1:
    name += ', {} shards (one per GPU), global batch {}'.format(
        n_gpus, batch * n_gpus)
  return name


# --------------------------------------------------------------------------- #
# clocks sampling (B200_PROFILING.md)
# --------------------------------------------------------------------------- #
class ClockSampler(object):
  """Samples SM clocks and throttle reasons of one GPU during the timed region.

  Uses NVML in-process (no fork: spawning nvidia-smi from several ranks stalls the
  driver for milliseconds and shows up in a 30 us step); falls back to nvidia-smi."""

  def __init__(self, index, period_s=0.05):
    self.index = index
    self.period = period_s
    self.sm, self.sm_max, self.reasons = [], [], set()
    self._stop = threading.Event()
    self._thread = threading.Thread(target=self._run, daemon=True)
    self._nvml = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self._nvml = pynvml
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = index
      if visible:
        ids = [v for v in visible.split(',') if v.strip()]
        if index < len(ids) and ids[index].strip().isdigit():
          phys = int(ids[index])
      self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
    except Exception:  # pylint: disable=broad-except
      self._nvml = None

  def _sample_nvml(self):
    n = self._nvml
    self.sm.append(int(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
    self.sm_max.append(int(n.nvmlDeviceGetMaxClockInfo(self._handle, n.NVML_CLOCK_SM)))
    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
               if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') else
               n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
    names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown',
             0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
    for bit, name in names.items():
      if mask & bit:
        self.reasons.add(name)

  def _sample_smi(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    out = subprocess.run(
        ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
         '--format=csv,noheader,nounits'], capture_output=True, text=True,
        timeout=5).stdout.strip()
    r = [x.strip() for x in out.split(',')]
    if len(r) >= 6 and r[0].isdigit():
      self.sm.append(int(r[0]))
      self.sm_max.append(int(r[1]))
      for k, name in enumerate(['hw_slowdown', 'hw_thermal_slowdown',
                                'sw_thermal_slowdown', 'sw_power_cap']):
        if r[2 + k].lower().startswith('active'):
          self.reasons.add(name)

  def _run(self):
    while not self._stop.is_set():
      try:
        if self._nvml is not None:
          self._sample_nvml()
        else:
          self._sample_smi()
      except Exception:  # pylint: disable=broad-except
        pass
      self._stop.wait(self.period if self._nvml is not None else 0.5)

  def __enter__(self):
    self._thread.start()
    return self

  def __exit__(self, *exc):
    self._stop.set()
    self._thread.join(timeout=6)

  def summary(self):
    sm = sorted(self.sm)
    return {'sm_mhz': sm[len(sm) // 2] if sm else None,
            'sm_max_mhz': max(self.sm_max) if self.sm_max else None,
            'reasons': sorted(self.reasons), 'samples': len(sm),
            'source': 'nvml' if self._nvml is not None else 'nvidia-smi'}


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
class GpuWorkload(object):
  """Capacity-`capacity` prioritized buffer filled with synthetic transitions on
  the device, plus resident network outputs for the C51 step."""

  def __init__(self, capacity, max_batch, rank, seed=1234):
    import torch
    from dopamine_b200 import _native
    from dopamine_b200.agents.rainbow import rainbow_agent
    from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
    self.torch, self.native, self.ra = torch, _native, rainbow_agent
    self.lib = _native.lib()
    self.capacity = capacity
    self.mem = prb.OutOfGraphPrioritizedReplayBuffer(
        (84, 84), STACK, capacity, 32, update_horizon=HORIZON, gamma=GAMMA,
        output='numpy', rng='device', seed=seed + rank)
    self.h = self.mem._h  # pylint: disable=protected-access
    gen = torch.Generator(device='cuda')
    gen.manual_seed(seed + rank)
    stream = _native.current_stream()
    chunk = 65536
    terminal_host = np.zeros(capacity, dtype=np.uint8)
    for lo in range(0, capacity, chunk):
      n = min(chunk, capacity - lo)
      frames = torch.randint(0, 256, (n, FRAME), dtype=torch.uint8,
                             device='cuda', generator=gen)
      actions = torch.randint(0, NUM_ACTIONS, (n,), dtype=torch.int32,
                              device='cuda', generator=gen)
      rewards = torch.randn(n, device='cuda', generator=gen).clamp_(-1, 1)
      terms = (torch.rand(n, device='cuda', generator=gen) < 1e-3).to(
          torch.uint8)
      for col, t in ((0, frames), (1, actions), (2, rewards), (3, terms)):
        _native.check(self.lib.b2r_store_write(self.h, col, lo, n, t.data_ptr(),
                                               stream))
      terminal_host[lo:lo + n] = terms.cpu().numpy()
    self.terminal_host = terminal_host
    add_count = capacity + 500  # full and wrapped
    self.mem.add_count = add_count
    cursor = add_count % capacity
    self.mem.invalid_range = [(cursor - HORIZON + i) % capacity
                              for i in range(STACK + HORIZON)]
    for lo in range(0, capacity, chunk):  # non-uniform priorities
      n = min(chunk, capacity - lo)
      idx = torch.arange(lo, lo + n, dtype=torch.int32, device='cuda')
      pr = (torch.randn(n, device='cuda', generator=gen).abs() + 1e-10).sqrt()
      self.mem.set_priority(idx, pr)
    torch.cuda.synchronize()
    _native.check(self.lib.b2r_check(self.h, stream))
    gen.manual_seed(7)
    self.online = torch.randn(max_batch, NUM_ACTIONS, NUM_ATOMS, device='cuda',
                              generator=gen)
    self.target = torch.randn(max_batch, NUM_ACTIONS, NUM_ATOMS, device='cuda',
                              generator=gen)
    self.support = rainbow_agent.make_support(VMAX, NUM_ATOMS)
    self.gamma_n = float(np.float32(math.pow(GAMMA, HORIZON)))
    self.seed = seed + rank
    self.fused = True
    self._plans = {}
    self.deferred = False

  def plan(self, batch):
    """Preallocated outputs + argument structs for a batch size (reused across
    steps, like the reference's optional output reuse for benchmarks)."""
    if batch in self._plans:
      return self._plans[batch]
    torch, nat = self.torch, self.native
    t = {
        'state': torch.empty(batch, 84, 84, STACK, dtype=torch.uint8, device='cuda'),
        'action': torch.empty(batch, dtype=torch.int32, device='cuda'),
        'reward': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'next_state': torch.empty(batch, 84, 84, STACK, dtype=torch.uint8, device='cuda'),
        'next_action': torch.empty(batch, dtype=torch.int32, device='cuda'),
        'next_reward': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'terminal': torch.empty(batch, dtype=torch.uint8, device='cuda'),
        'indices': torch.empty(batch, dtype=torch.int32, device='cuda'),
        'sampling_probabilities': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'loss': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'priorities': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'weights': torch.empty(batch, dtype=torch.float32, device='cuda'),
        'mean': torch.empty((), dtype=torch.float32, device='cuda'),
    }
    b = nat.Batch()
    for name in ('state', 'action', 'reward', 'next_state', 'next_action',
                 'next_reward', 'terminal', 'indices', 'sampling_probabilities'):
      setattr(b, name, t[name].data_ptr())
    c = nat.C51Args()
    c.batch, c.num_actions, c.num_atoms = batch, NUM_ACTIONS, NUM_ATOMS
    c.cumulative_gamma = self.gamma_n
    c.support = self.support.data_ptr()
    c.target_logits = self.target.data_ptr()
    c.online_logits = self.online.data_ptr()
    c.actions = t['action'].data_ptr()
    c.rewards = t['reward'].data_ptr()
    c.terminals = t['terminal'].data_ptr()
    c.sampling_probabilities = t['sampling_probabilities'].data_ptr()
    c.target = None
    c.loss = t['loss'].data_ptr()
    c.priorities = t['priorities'].data_ptr()
    c.weights = t['weights'].data_ptr()
    # The scalar mean(w * loss) only feeds summaries (RA:298-301); the gradient
    # path needs loss / weights per row, which are produced.
    c.mean_weighted_loss = None
    c.grad_logits = None
    self._plans[batch] = (t, b, c)
    return self._plans[batch]

  def set_deferred(self, on):
    """Deferred frame copies (b2r_set_deferred_frames): the copies of step n run beside
    the sampler -> loss -> write-back chain of step n + 1; `join` is then part of every
    timed group of steps."""
    self.native.check(self.lib.b2r_join_frames(self.h, self.native.current_stream()))
    self.native.check(self.lib.b2r_set_deferred_frames(self.h, 1 if on else 0))
    self.deferred = bool(on)

  def join(self):
    self.native.check(self.lib.b2r_join_frames(self.h, self.native.current_stream()))

  def step(self, batch):
    """sample -> gather -> C51 loss/priorities -> write-back, all in HBM."""
    t, b, c = self.plan(batch)
    nat, lib = self.native, self.lib
    stream = nat.current_stream()
    if self.fused:
      # one call: the sampler writes the scalar columns, the frame copies run on a
      # forked stream beside loss + write-back and rejoin before the call returns
      nat.check(lib.b2r_train_step_device(
          self.h, batch, self.seed, 0, ctypes.byref(b), ctypes.byref(c), stream))
      return
    nat.check(lib.b2r_sample_transition_batch_device(
        self.h, batch, self.seed, 0, ctypes.byref(b), stream))
    nat.check(lib.b2r_c51_loss(ctypes.byref(c), stream))
    nat.check(lib.b2r_set_priority_device(
        self.h, batch, t['indices'].data_ptr(), t['priorities'].data_ptr(),
        stream))

  def gather_only(self, batch, idx_tensor):
    _, b, _ = self.plan(batch)
    self.native.check(self.lib.b2r_gather_device(
        self.h, batch, idx_tensor.data_ptr(), ctypes.byref(b),
        self.native.current_stream()))

  def algorithmic_bytes(self, idx_host):
    """unique frames read once + outputs written once, per index."""
    cap = self.capacity
    total = 0
    for i in idx_host:
      length = HORIZON
      for k in range(HORIZON):
        if self.terminal_host[(int(i) + k) % cap]:
          length = k + 1
          break
      frames_read = STACK + min(length, STACK)
      total += FRAME * (frames_read + 2 * STACK)
      total += 64 + 25  # scalar columns in, scalar outputs
    return total


def steps_per_graph(steps, limit):
  """Largest divisor of `steps` that is <= limit (so that exactly `steps` run)."""
  g = max(1, min(int(limit), int(steps)))
  while steps % g:
    g -= 1
  return g


def time_graph_or_eager(torch, fn, steps, warmup, use_graph, dist=None, per_graph=1,
                        min_ms=0.0, info=None, finish=None):
  """Times `steps` calls of fn with CUDA events on the launching stream.

  per_graph > 1 captures that many consecutive calls in one CUDA graph (it must
  divide `steps`): consecutive graph launches are paced by the front end in units of
  about 2 us on B200 (a one-kernel graph replayed back to back reports 6.16, 8.21,
  10.26 ... us whatever the kernel does), and programmatic dependent launch cannot
  overlap a step's first kernel with the previous graph's last one.

  min_ms > 0: a timed region of `steps` steps that is shorter than min_ms is REPEATED
  (each repetition is again exactly `steps` steps between two events, barrier and
  synchronize on both sides) until min_ms of timed work has run, and the MEDIAN region
  is returned — so that `--steps 20` does not rest on one 0.4 ms sample.  With `dist`
  every rank runs the same number of regions and each region counts with its maximum
  over the ranks.  info (dict, optional) receives what was done.

  finish (optional): called after every group of per_graph calls (inside the captured
  graph; after each call when eager) — the join of work that fn leaves running on other
  streams (deferred frame copies), so that every timed region contains ALL the work of
  its steps."""
  side = torch.cuda.Stream()
  side.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(side):
    for _ in range(max(3, warmup)):
      fn()
    if finish is not None:
      finish()
    side.synchronize()
    runner = fn
    if finish is not None:
      def runner():  # eager: one call, then the join
        fn()
        finish()
    if use_graph:
      graph = torch.cuda.CUDAGraph()
      assert steps % per_graph == 0, (steps, per_graph)
      with torch.cuda.graph(graph, stream=side):
        for _ in range(per_graph):
          fn()
        if finish is not None:
          finish()
      runner = graph.replay
      for _ in range(3):
        runner()
      side.synchronize()
    else:
      per_graph = 1

    def region():
      if dist is not None:
        dist.barrier()
      torch.cuda.synchronize()
      start = torch.cuda.Event(enable_timing=True)
      end = torch.cuda.Event(enable_timing=True)
      start.record(side)
      for _ in range(steps // per_graph):
        runner()
      end.record(side)
      end.synchronize()
      torch.cuda.synchronize()
      if dist is not None:
        dist.barrier()
      return start.elapsed_time(end)

    regions = [region()]
    if min_ms > 0.0:
      first = torch.tensor([regions[0]], dtype=torch.float64, device='cuda')
      if dist is not None:
        dist.all_reduce(first, op=dist.ReduceOp.MAX)  # every rank repeats equally often
      more = 0
      if float(first.item()) < min_ms:
        more = min(2000, int(np.ceil(min_ms / max(float(first.item()), 1e-3))) - 1)
        more += (more + 1) % 2 == 0  # odd number of regions: the median is one of them
      regions += [region() for _ in range(more)]
    ms_all = torch.tensor(regions, dtype=torch.float64, device='cuda')
    if dist is not None and len(regions) > 1:
      dist.all_reduce(ms_all, op=dist.ReduceOp.MAX)
    ms = float(ms_all.median().item()) if len(regions) > 1 else regions[0]
    if info is not None:
      info.update({'timed_regions': len(regions), 'steps_per_region': steps,
                   'timed_total_ms': round(float(ms_all.sum().item()), 3),
                   'region_ms_min': round(float(ms_all.min().item()), 5),
                   'region_ms_median': round(ms, 5),
                   'region_ms_max': round(float(ms_all.max().item()), 5)})
  torch.cuda.current_stream().wait_stream(side)
  return ms


def measure_gather_roofline(torch, wl, batch, peak_gbs, launches=200):
  """Average duration of the gather kernel alone over pre-sampled index batches
  (distinct every launch: the 7 GB ring defeats L2) -> achieved algorithmic GB/s."""
  nat = wl.native
  nbuf = 20
  idx_bufs = []
  for k in range(nbuf):
    idx = torch.empty(batch, dtype=torch.int32, device='cuda')
    nat.check(wl.lib.b2r_sample_indices_device(
        wl.h, batch, wl.seed, 1000 + k, idx.data_ptr(), nat.current_stream()))
    idx_bufs.append(idx)
  torch.cuda.synchronize()
  bytes_per_launch = np.mean(
      [wl.algorithmic_bytes(i.cpu().numpy()) for i in idx_bufs])

  def body():
    for idx in idx_bufs:
      wl.gather_only(batch, idx)

  reps = max(1, launches // nbuf)
  ms = time_graph_or_eager(torch, body, reps, 3, True)
  sec_per_launch = ms * 1e-3 / (reps * nbuf)
  achieved = bytes_per_launch / sec_per_launch / 1e9
  # which of the two frame-copy kernels a launch of this size takes (gather.cu)
  variant = int(wl.lib.b2r_gather_variant(wl.h, batch))
  kernel = {0: 'gather_stack4_u8_kernel', 1: 'gather_stack4_u8_tma_kernel'}.get(
      variant, 'gather_generic_kernel')
  traffic, traffic_note = None, None
  try:  # per-launch DRAM bytes of this kernel from the committed ncu capture
    t = json.load(open(os.path.join(ROOT, 'profiles', 'r2', 'traffic.json')))
    t = t[kernel].get(str(batch))
    if t:
      traffic, traffic_note = t['traffic_bytes'], t['note']
  except Exception:  # pylint: disable=broad-except
    pass
  return {
      'bound': 'hbm', 'kernel': kernel,
      'achieved': round(achieved, 1), 'peak': peak_gbs, 'unit': 'GB/s',
      'frac': round(achieved / peak_gbs, 4), 'traffic': traffic,
      'traffic_note': traffic_note,
      'us_per_launch': round(sec_per_launch * 1e6, 3),
      'algorithmic_bytes_per_launch': int(bytes_per_launch),
      'peak_source': 'MEASURED_PEAKS.json hbm_gbs (burst, kernel timed alone)',
  }


def measure_e2e(torch, wl, batch, steps, update_period=4, pipeline_depth=2,
                dist=None, world=1, rank=0):
  """The same metric end to end through the public host-facing API.

  Per step, as the agent drives the replay (dqn_agent.py:359-442): `update_period`
  new transitions are add()-ed from HOST frames (staged and copied to HBM), then
  `ReplayTrainer.step` takes both network outputs from pinned HOST memory (H2D),
  runs sample -> gather -> C51 loss -> priority write-back on the device and
  hands the per-row losses back in HOST memory (D2H).  With pipeline_depth d the
  losses returned by a call are those of the step queued d calls earlier, so the
  host does not wait for the step it has just queued (d = 0: fully synchronous).
  Every step's copies and kernels are inside the timed region either way."""
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
  ra, mem = wl.ra, wl.mem
  # N > 1: one trainer per rank over its own shard; `batch` is the GLOBAL batch, the
  # shard totals travel through a peer-memory exchange of the trainer's own, every
  # rank adds its own rows and gets the losses of the rows it served.
  # a rank's network produces logits for the rows the rank serves (its share of the
  # global batch, here capped at twice the mean share: +6 sigma at 8 ranks)
  logit_rows = batch if world == 1 else min(batch, 2 * (batch // world))
  trainer = ra.ReplayTrainer(mem, NUM_ACTIONS, NUM_ATOMS, VMAX, batch_size=batch,
                             pipeline_depth=pipeline_depth,
                             seed=wl.seed if world == 1 else 4321,
                             logit_rows=logit_rows)
  if world > 1:
    from dopamine_b200.replay_memory import sharded_replay
    e2e_exchange = sharded_replay.PeerExchange(rank=rank, world_size=world)
    e2e_exchange.set_early_publish(True)
    trainer.set_exchange(e2e_exchange)
  rng = np.random.RandomState(3)
  frames = rng.randint(0, 256, size=(64, 84, 84)).astype(np.uint8)
  online_h = wl.online[:logit_rows].cpu().pin_memory()
  target_h = wl.target[:logit_rows].cpu().pin_memory()
  online_p, target_p = online_h.data_ptr(), target_h.data_ptr()
  counter = [0]
  most_rows = [0]
  sentinel = prb.MAX_RECORDED_PRIORITY
  stream = wl.native.current_stream()

  def one():
    for _ in range(update_period):
      k = counter[0]
      counter[0] += 1
      mem.add(frames[k & 63], k % NUM_ACTIONS, 0.5, int(k % 1000 == 999), sentinel)
    done = trainer.step_pointers(online_p, target_p, stream)
    if world > 1:
      most_rows[0] = max(most_rows[0], trainer.last_rows)
    return done

  for _ in range(20):
    one()
  trainer.drain()
  torch.cuda.synchronize()
  if dist is not None:
    dist.barrier()
  t0 = time.perf_counter()
  for _ in range(steps):
    one()
  _, last = trainer.drain()
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  if dist is not None:  # the slowest rank's clock
    slowest = torch.tensor([dt], device='cuda', dtype=torch.float64)
    dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    dt = float(slowest.item())
  assert last == steps + 20 - 1, (last, steps)
  assert most_rows[0] <= logit_rows, (most_rows[0], logit_rows)
  wl.native.check(wl.lib.b2r_check(wl.h, stream))
  row = 7056 + 16  # staged row: frame + action + reward + terminal (padded)
  h2d = world * (update_period * row + (online_h.numel() + target_h.numel()) * 4)
  d2h = world * (logit_rows + 1) * 4
  return {'value': round(batch * steps / dt, 1), 'unit': UNIT,
          'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
          'ms_per_step': round(dt * 1e3 / steps, 4), 'steps': steps,
          'pipeline_depth': pipeline_depth,
          'what': ('public host API per step: %d x add() of host frames + '
                   'ReplayTrainer.step(): H2D of both logits tensors from pinned '
                   'memory, sample+gather+C51+set_priority on the device, D2H of '
                   'the per-row losses (returned %d steps later)'
                   % (update_period, pipeline_depth))}


def measure_full_train_step(torch, wl, batch, steps, ddp=False, cuda_graph=False):
  """BASELINE config 5: the whole Rainbow update on the synthetic Atari shape —
  fused sample + gather feeding the cuDNN Nature-DQN distribution network (online on
  state, target on next_state), fused C51 loss, backward, Adam, priority write-back
  (dopamine_b200/agents/rainbow/agent.py).  Eager PyTorch around the kernels."""
  from dopamine_b200.agents.rainbow import agent
  mem = wl.mem
  saved = (mem._output, mem._reuse_outputs, mem._batch_size)  # pylint: disable=protected-access
  mem._output, mem._reuse_outputs, mem._batch_size = 'torch', True, batch  # pylint: disable=protected-access
  learner = agent.RainbowLearner(NUM_ACTIONS, batch_size=batch, memory=mem, ddp=ddp,
                                 update_horizon=HORIZON, gamma=GAMMA, vmax=VMAX,
                                 cuda_graph=cuda_graph)
  for _ in range(10):
    learner.train_step()
  torch.cuda.synchronize()
  start = torch.cuda.Event(enable_timing=True)
  end = torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  start.record()
  for _ in range(steps):
    loss = learner.train_step()
  end.record()
  end.synchronize()
  wall = time.perf_counter() - t0
  ms = start.elapsed_time(end)
  mem._output, mem._reuse_outputs, mem._batch_size = saved  # pylint: disable=protected-access
  return {'updates_per_s': round(steps / (ms * 1e-3), 1),
          'transitions_per_s': round(batch * steps / (ms * 1e-3), 1),
          'ms_per_update': round(ms / steps, 4), 'wall_ms_per_update':
          round(wall * 1e3 / steps, 4), 'batch': batch, 'steps': steps,
          'last_loss': round(float(loss.detach()), 5),
          'what': 'sample+gather -> conv nets (cuDNN, fp32/TF32) -> fused C51 loss '
                  '-> backward -> Adam -> set_priority; ' + (
                      'whole update replayed as one CUDA graph' if cuda_graph else
                      'eager PyTorch host loop')}


def measure_config1_dqn(torch, steps, batch=32, capacity=100000, cpu_budget_s=4.0):
  """BASELINE configs[0]: DQN's OutOfGraphReplayBuffer, 84x84 uint8 frames, capacity
  100k, stack 4, batch 32, uniform sampling, n = 1 — uniform sample + gather + DQN
  Bellman target / Huber loss (dqn_agent.py:283-322), device-timed as a CUDA graph,
  beside the CPU port of the same step."""
  from dopamine_b200 import _native
  from dopamine_b200.replay_memory import circular_replay_buffer as crb
  from oracle import dqn_port
  from oracle.replay_port import PortReplay, cursor_window
  lib = _native.lib()
  mem = crb.OutOfGraphReplayBuffer((84, 84), STACK, capacity, batch, update_horizon=1,
                                   gamma=GAMMA, output='torch', rng='device', seed=11,
                                   reuse_outputs=True)
  gen = torch.Generator(device='cuda')
  gen.manual_seed(11)
  stream = _native.current_stream()
  frames = torch.randint(0, 256, (capacity, FRAME), dtype=torch.uint8, device='cuda',
                         generator=gen)
  actions = torch.randint(0, NUM_ACTIONS, (capacity,), dtype=torch.int32,
                          device='cuda', generator=gen)
  rewards = torch.randn(capacity, device='cuda', generator=gen).clamp_(-1, 1)
  terms = (torch.rand(capacity, device='cuda', generator=gen) < 1e-3).to(torch.uint8)
  for col, t in ((0, frames), (1, actions), (2, rewards), (3, terms)):
    _native.check(lib.b2r_store_write(mem._h, col, 0, capacity, t.data_ptr(), stream))  # pylint: disable=protected-access
  mem.add_count = capacity + 500  # full and wrapped
  mem.invalid_range = [(500 - 1 + i) % capacity for i in range(STACK + 1)]
  _, arrays, b = mem._alloc_outputs(batch, True)  # pylint: disable=protected-access
  online_q = torch.randn(batch, NUM_ACTIONS, device='cuda', generator=gen)
  target_q = torch.randn(batch, NUM_ACTIONS, device='cuda', generator=gen)
  loss = torch.empty(batch, dtype=torch.float32, device='cuda')
  a = _native.DqnArgs()
  a.batch, a.num_actions = batch, NUM_ACTIONS
  a.cumulative_gamma = float(np.float32(GAMMA))
  a.target_q, a.online_q = target_q.data_ptr(), online_q.data_ptr()
  a.actions, a.rewards = arrays[1].data_ptr(), arrays[2].data_ptr()
  a.terminals, a.loss = arrays[6].data_ptr(), loss.data_ptr()

  def step():
    s = _native.current_stream()
    _native.check(lib.b2r_sample_transition_batch_device(
        mem._h, batch, 11, 0, ctypes.byref(b), s))  # pylint: disable=protected-access
    _native.check(lib.b2r_dqn_loss(ctypes.byref(a), s))

  ms = time_graph_or_eager(torch, step, steps, 20, True,
                           per_graph=steps_per_graph(steps, 10))
  _native.check(lib.b2r_check(mem._h, _native.current_stream()))  # pylint: disable=protected-access
  # CPU: the oracle port of the same step, one core
  port = PortReplay((84, 84), STACK, capacity, batch, update_horizon=1, gamma=GAMMA)
  rng = np.random.RandomState(11)
  pattern = rng.randint(0, 256, size=(4096, 84, 84)).astype(np.uint8)
  for lo in range(0, capacity, 4096):
    n = min(4096, capacity - lo)
    port.store['observation'][lo:lo + n] = pattern[:n]
  port.store['action'][:] = rng.randint(0, NUM_ACTIONS, size=capacity)
  port.store['reward'][:] = np.clip(rng.randn(capacity), -1, 1)
  port.store['terminal'][:] = rng.rand(capacity) < 1e-3
  port.add_count = np.array(capacity + 500)
  port.invalid_range = cursor_window(500, capacity, STACK, 1)
  oq = rng.randn(batch, NUM_ACTIONS).astype(np.float32)
  tq = rng.randn(batch, NUM_ACTIONS).astype(np.float32)
  np.random.seed(0)
  done, t0 = 0, time.perf_counter()
  while time.perf_counter() - t0 < cpu_budget_s or done < 3:
    bt = port.sample_transition_batch(batch)
    dqn_port.dqn_update(bt[2], bt[6], bt[1], oq, tq, GAMMA, 1)
    done += 1
  cpu_dt = time.perf_counter() - t0
  return {'workload': 'DQN OutOfGraphReplayBuffer capacity {}, stack 4, batch {}, '
                      'uniform sampling, n = 1 (BASELINE configs[0])'.format(
                          capacity, batch),
          'value': round(batch * steps / (ms * 1e-3), 1), 'unit': UNIT,
          'ms_per_step': round(ms / steps, 6), 'steps': steps,
          'cpu_port_value': round(batch * done / cpu_dt, 1), 'cpu_cores': 1,
          'what': 'uniform sample + gather + DQN target / Huber loss; CUDA graph replay, '
                  '10 steps per graph launch'}


def measure_next_rows(torch, batch=32):
  """SURVEY 8f rows 3-4, one number each: IQN's quantile-Huber loss (64 x 64 tau
  samples, 32 action samples, 18 actions) as one device-timed launch beside its numpy
  port, and the actor's per-environment-step record_observation (host call + kernel)
  beside the reference's np.roll + assignment."""
  from dopamine_b200.agents.dqn import dqn_agent
  from dopamine_b200.agents.implicit_quantile import implicit_quantile_agent as iqa
  from oracle import iqn_port
  out = {}
  n, n_prime, k = 64, 64, 32
  rng = np.random.RandomState(5)
  case = dict(
      rewards=np.clip(rng.randn(batch), -1, 1).astype(np.float32),
      terminals=(rng.rand(batch) < 0.05).astype(np.uint8),
      actions=rng.randint(0, NUM_ACTIONS, size=batch).astype(np.int32),
      online_quantile_values=rng.randn(n * batch, NUM_ACTIONS).astype(np.float32),
      quantiles=rng.rand(n * batch, 1).astype(np.float32),
      target_quantile_values=rng.randn(n_prime * batch, NUM_ACTIONS).astype(np.float32),
      action_quantile_values=rng.randn(k * batch, NUM_ACTIONS).astype(np.float32))
  dev = {key: torch.as_tensor(v, device='cuda') for key, v in case.items()}
  res = iqa.quantile_huber_loss(
      dev['online_quantile_values'], dev['quantiles'], dev['target_quantile_values'],
      dev['action_quantile_values'], dev['actions'], dev['rewards'], dev['terminals'],
      GAMMA ** 3, 1.0, want_grad=True)

  def iqn():
    iqa.quantile_huber_loss(
        dev['online_quantile_values'], dev['quantiles'], dev['target_quantile_values'],
        dev['action_quantile_values'], dev['actions'], dev['rewards'],
        dev['terminals'], GAMMA ** 3, 1.0, want_grad=True, out=res)

  reps = 500
  ms = time_graph_or_eager(torch, iqn, reps, 10, True, per_graph=10)
  t0, done = time.perf_counter(), 0
  while time.perf_counter() - t0 < 1.0 or done < 3:
    iqn_port.iqn_update(num_tau_samples=n, num_tau_prime_samples=n_prime,
                        num_quantile_samples=k, gamma=GAMMA, update_horizon=3, **case)
    done += 1
  out['iqn_loss'] = {
      'batch': batch, 'tau_samples': [n, n_prime, k],
      'us_per_launch': round(ms * 1e3 / reps, 2),
      'cpu_port_us': round((time.perf_counter() - t0) / done * 1e6, 1),
      'what': 'greedy next action + target quantiles + quantile-Huber loss + '
              'gradient, one launch (b2r_iqn_loss)'}
  # actor: wall clock per call (host memcpy into a pinned slot + one launch)
  actor = dqn_agent.ActorState((84, 84), STACK, np.uint8, slots=8)
  frames = rng.randint(0, 256, size=(64, 84, 84)).astype(np.uint8)
  for i in range(50):
    actor.record(frames[i % 64])
  torch.cuda.synchronize()
  calls = 5000
  t0 = time.perf_counter()
  for i in range(calls):
    actor.record(frames[i % 64])
  torch.cuda.synchronize()
  record_us = (time.perf_counter() - t0) / calls * 1e6
  state = np.zeros((1, 84, 84, STACK), np.uint8)
  t0 = time.perf_counter()
  for i in range(calls):
    state = np.roll(state, -1, axis=-1)
    state[0, ..., -1] = frames[i % 64]
  out['record_observation'] = {
      'us_per_call': round(record_us, 2),
      'numpy_roll_us': round((time.perf_counter() - t0) / calls * 1e6, 2),
      'what': 'ActorState.record: frame through a pinned slot, roll + insert of the '
              '(1, 84, 84, 4) state in HBM in one launch; beside the reference\'s '
              'np.roll + assignment on the host (which still has to ship the state to '
              'the device)'}
  actor.close()
  return out


def measure_e2e_host_batch(torch, wl, batch, steps):
  """Variant that also ships the whole sampled batch to host numpy arrays, i.e. the
  reference's OutOfGraph* return convention (1.8 MB of D2H per step at batch 32)."""
  mem, ra = wl.mem, wl.ra
  online_h = wl.online[:batch].cpu().pin_memory()
  target_h = wl.target[:batch].cpu().pin_memory()
  online_d = torch.empty_like(wl.online[:batch])
  target_d = torch.empty_like(wl.target[:batch])

  def one():
    batch_np = mem.sample_transition_batch(batch)  # host numpy tuple
    online_d.copy_(online_h, non_blocking=True)
    target_d.copy_(target_h, non_blocking=True)
    dev = lambda a: torch.as_tensor(a).cuda(non_blocking=True)
    out = ra.c51_loss(online_d, target_d, dev(batch_np[1]), dev(batch_np[2]),
                      dev(batch_np[6]), dev(batch_np[8]), wl.support, wl.gamma_n)
    prio = out['priorities'].cpu().numpy()
    mem.set_priority(batch_np[7], prio)
    return batch_np

  for _ in range(5):
    sample = one()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(steps):
    one()
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  d2h = sum(a.nbytes for a in sample) + batch * 4
  h2d = (online_h.numel() + target_h.numel()) * 4 + batch * (4 + 4 + 1 + 4 + 4 + 8)
  return {'value': round(batch * steps / dt, 1), 'unit': UNIT,
          'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
          'ms_per_step': round(dt * 1e3 / steps, 4), 'steps': steps}


# --------------------------------------------------------------------------- #
# CPU arm: the oracle port (the reference's algorithm, Python/numpy)
# --------------------------------------------------------------------------- #
def cpu_arm_kind():
  """'reference': the reference's own replay classes are importable (from
  /root/reference, or from the copy `make -C oracle _ref` made for the GPU box);
  'port': only the oracle restatement is."""
  from oracle import refshim
  return 'reference' if refshim.reference_available() else 'port'


def build_cpu_port(capacity, batch, seed=1234, kind=None):
  """A full, wrapped prioritized replay memory on the host with the bench's synthetic
  contents: the reference's OutOfGraphPrioritizedReplayBuffer itself (kind
  'reference'; prioritized_replay_buffer.py, imported unmodified behind
  oracle/refshim.py's tensorflow / gin stubs) or the oracle port of it."""
  kind = kind or cpu_arm_kind()
  rng = np.random.RandomState(seed)
  if kind == 'reference':
    from oracle import refshim
    _, crb, prb = refshim.load_reference()
    port = prb.OutOfGraphPrioritizedReplayBuffer(
        (84, 84), STACK, capacity, batch, update_horizon=HORIZON, gamma=GAMMA)
    store = port._store  # pylint: disable=protected-access
    levels = port.sum_tree.nodes
    window = lambda cursor: crb.invalid_range(cursor, capacity, STACK, HORIZON)
  else:
    from oracle.replay_port import PortPrioritizedReplay, cursor_window
    port = PortPrioritizedReplay((84, 84), STACK, capacity, batch,
                                 update_horizon=HORIZON, gamma=GAMMA)
    store = port.store
    levels = [port.sum_tree.level(l) for l in range(port.sum_tree.depth + 1)]
    window = lambda cursor: cursor_window(cursor, capacity, STACK, HORIZON)
  pattern = rng.randint(0, 256, size=(4096, 84, 84)).astype(np.uint8)
  obs = store['observation']
  for lo in range(0, capacity, 4096):
    n = min(4096, capacity - lo)
    obs[lo:lo + n] = pattern[:n]
  store['action'][:] = rng.randint(0, NUM_ACTIONS, size=capacity)
  store['reward'][:] = np.clip(rng.randn(capacity), -1, 1)
  store['terminal'][:] = rng.rand(capacity) < 1e-3
  port.add_count = np.array(capacity + 500)
  port.invalid_range = window(500 % capacity)
  # tree: leaves = priorities, parents = child sums (timing only needs the shape)
  depth = len(levels) - 1
  leaves = np.zeros(1 << depth)
  leaves[:capacity] = np.sqrt(np.abs(rng.randn(capacity)) + 1e-10).astype(
      np.float32)
  level = leaves
  for l in range(depth, -1, -1):
    levels[l][:] = level
    level = level.reshape(-1, 2).sum(axis=1) if l else level
  port.sum_tree.max_recorded_priority = float(leaves.max())
  return port


CPU_ARM_WHAT = {
    'reference': 'the reference itself: dopamine.replay_memory (sum_tree, circular_'
                 'replay_buffer, prioritized_replay_buffer) imported unmodified behind '
                 'tensorflow/gin stubs for add / sample_transition_batch / set_priority; '
                 'its C51 loss is TensorFlow-1.x graph code (absent), so that part is '
                 'the numpy restatement oracle/c51_port.py; single thread as the '
                 'reference is',
    'port': 'oracle port of the reference: Python/numpy, single thread as the '
            'reference is',
}


def cpu_steps(port, batch, budget_s, max_steps, seed=7, update_period=4):
  """The e2e workload on the CPU: `update_period` add()s, then one pass of the path
  (sample_transition_batch -> C51 loss / priorities -> set_priority)."""
  from oracle import c51_port
  rng = np.random.RandomState(seed)
  online = rng.randn(batch, NUM_ACTIONS, NUM_ATOMS).astype(np.float32)
  target = rng.randn(batch, NUM_ACTIONS, NUM_ATOMS).astype(np.float32)
  frames = rng.randint(0, 256, size=(64, 84, 84)).astype(np.uint8)
  random.seed(0)
  done = 0
  k = 0
  t0 = time.perf_counter()
  while done < max_steps and (time.perf_counter() - t0 < budget_s or done < 3):
    for _ in range(update_period):
      port.add(frames[k & 63], k % NUM_ACTIONS, 0.5, int(k % 1000 == 999),
               port.sum_tree.max_recorded_priority)
      k += 1
    b = port.sample_transition_batch(batch)
    out = c51_port.rainbow_update(b[2], b[6], b[1], b[8], online, target,
                                  vmax=VMAX, num_atoms=NUM_ATOMS, gamma=GAMMA,
                                  update_horizon=HORIZON)
    port.set_priority(b[7], out['priorities'])
    done += 1
  return done, time.perf_counter() - t0


def cpu_baseline(batch, capacity, budget_s=12.0):
  kind = cpu_arm_kind()
  port = build_cpu_port(capacity, batch, kind=kind)
  cpu_steps(port, batch, 0.0, 3)  # warm-up
  steps, dt = cpu_steps(port, batch, budget_s, 100000)
  return {
      'value': round(batch * steps / dt, 1), 'unit': UNIT, 'cores': 1,
      'kind': kind,
      'sample': '{} steps (4 x add() + sample + C51 + set_priority each) of batch {} '
                'in {:.1f} s, capacity {} ({})'.format(steps, batch, dt, capacity,
                                                       CPU_ARM_WHAT[kind]),
  }


def _replica(args):
  batch, capacity, budget_s, seed = args
  port = build_cpu_port(capacity, batch, seed=seed)
  cpu_steps(port, batch, 0.0, 2)
  steps, dt = cpu_steps(port, batch, budget_s, 100000, seed=seed)
  return steps, dt


def _usable_replicas(per_replica_bytes, cap=64):
  """Replica processes the host can carry: one per core, bounded by free memory."""
  cores = os.cpu_count() or 1
  try:
    import psutil
    free = psutil.virtual_memory().available
  except Exception:  # pylint: disable=broad-except
    free = 16 << 30
  by_mem = int(free * 0.4 // per_replica_bytes)
  return max(1, min(cores, cap, by_mem))


def run_reference(args):
  """--impl reference: the reference's CPU implementation of the path (oracle
  port; the reference itself is TF-1.x Python and cannot travel) on the host.

  The reference is single-threaded by construction (one Python process, the GIL,
  a sequential `random` stream feeding one replay memory), so one core is every
  host thread it can use for THIS workload (one replay memory of `capacity`
  transitions): that is `value`.  What the same cores deliver as independent
  replica processes, each with its own smaller replay memory — a different job,
  the CPU analogue of running one shard per GPU — is reported beside it in
  `all_cores_replicas`."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  import multiprocessing as mp
  budget = 12.0
  kind = cpu_arm_kind()
  t0 = time.perf_counter()
  port = build_cpu_port(args.capacity, args.batch, kind=kind)
  cpu_steps(port, args.batch, 0.0, max(3, min(args.warmup, 20)))
  steps, dt = cpu_steps(port, args.batch, budget, 100000)
  del port
  rate = args.batch * steps / dt
  per_cap = 65536
  replicas = _usable_replicas(per_cap * (FRAME + 64) * 1.2)
  with mp.get_context('fork').Pool(replicas) as pool:
    res = pool.map(_replica, [(args.batch, per_cap, 8.0, 100 + r)
                              for r in range(replicas)])
  wall = time.perf_counter() - t0
  rate_all = sum(args.batch * s / d for s, d in res)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': round(rate, 1),
      'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'cpu_steps_timed': steps,
      'ms_per_step': round(1e3 * args.batch / rate, 4),
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'u8', 'data': 'synthetic',
      'config': {'workload': workload_name(args.batch, args.capacity, 1)},
      'cpu_baseline': {
          'value': round(rate, 1), 'unit': UNIT, 'cores': 1, 'kind': kind,
          'sample': '{} steps (4 x add() + sample + C51 + set_priority each) of '
                    'batch {} in {:.1f} s, one process, capacity {} ({})'.format(
                        steps, args.batch, dt, args.capacity, CPU_ARM_WHAT[kind]),
          'all_cores_replicas': {
              'value': round(rate_all, 1), 'unit': UNIT, 'cores': replicas,
              'host_cores': os.cpu_count(),
              'sample': '{} independent replica processes x ~8 s, capacity {} '
                        'each: not one replay memory, reported for scale'.format(
                            replicas, per_cap)},
          'wall_s': round(wall, 1),
      },
      'e2e': {'value': round(rate, 1), 'unit': UNIT, 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def ordered_line(line):
  """Key order of the JSON line: the contract's keys first, then a compact copy of the
  sweep (whole-step fraction of the HBM peak and microseconds per step, per batch), the
  roofline / e2e / baseline objects, and the long descriptive objects last; the compact
  sweep is repeated as the very last key, so that a record which keeps only one end of
  the line still has it."""
  compact = None
  if isinstance(line.get('sweep'), dict):
    compact = {}
    for b, r in line['sweep'].items():
      compact[b] = {k: r[k] for k in ('whole_step_frac', 'gather_frac', 'value')
                    if k in r}
      if 'ms_per_step' in r:
        compact[b]['us_per_step'] = round(r['ms_per_step'] * 1e3, 2)
  front = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
           'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data']
  middle = ['roofline', 'e2e', 'cpu_baseline', 'gpu_launches', 'clocks', 'shard_check',
            'timing', 'e2e_sync', 'e2e_host_batch']
  out = {k: line[k] for k in front if k in line}
  if compact is not None:
    out['sweep_summary'] = compact
  for k in middle:
    if k in line:
      out[k] = line[k]
  for k, v in line.items():
    if k not in out:
      out[k] = v
  if compact is not None:
    out['sweep_summary_again'] = compact
  return out


def bind_to_gpu_numa_node(local_rank):
  """One process per GPU: run (and first-touch pinned memory) on the CPU cores the
  GPU is attached to, as NCCL does for its own threads.  The host loop of a rank
  reads and writes pinned buffers that its GPU accesses over PCIe; from the far
  socket every such access crosses the inter-socket link as well."""
  try:
    import pynvml
    pynvml.nvmlInit()
    visible = os.environ.get('CUDA_VISIBLE_DEVICES')
    phys = local_rank
    if visible:
      ids = [v for v in visible.split(',') if v.strip()]
      if local_rank < len(ids) and ids[local_rank].strip().isdigit():
        phys = int(ids[local_rank])
    handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
    cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64)
            if (int(word) >> b) & 1]
    cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
    if cpus:
      os.sched_setaffinity(0, cpus)
      return len(cpus)
  except Exception:  # pylint: disable=broad-except
    pass
  return None


# --------------------------------------------------------------------------- #
def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
    return
  import torch
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback')
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
  torch.cuda.set_device(local_rank)
  dist = None
  if world > 1:
    import torch.distributed as dist_mod
    dist_mod.init_process_group('nccl')
    dist = dist_mod
  from dopamine_b200 import _native
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
  except Exception:  # pylint: disable=broad-except
    pass
  peak_gbs = float(peaks.get('hbm_gbs', 6650.0))

  sweep_batches = [] if args.no_sweep else [256, 1024, 4096]
  wl = GpuWorkload(args.capacity,
                   max([args.batch * world] + [b * world for b in sweep_batches]),
                   rank)
  wl.fused = not args.unfused
  launches_before = _native.lib().b2r_launch_count()
  wl.step(args.batch)
  launches_per_step = _native.lib().b2r_launch_count() - launches_before
  torch.cuda.synchronize()

  if world > 1:
    from dopamine_b200.replay_memory import sharded_replay
    exchange = None
    if args.exchange == 'p2p':
      exchange = sharded_replay.PeerExchange(rank=rank, world_size=world)
      # the write-back publishes the shard total for the next step the moment the root
      # is written (one-CTA tree kernels): the wire latency hides behind its tail
      exchange.set_early_publish(not args.no_early_publish)
    sharded = sharded_replay.ShardedStep(wl, args.batch * world, world, rank, dist,
                                         exchange=exchange)
    step_fn = sharded.step
    launches_per_step = sharded.launches_per_step()
    use_graph = not args.no_graph  # NCCL all-gather is captured with the kernels
  else:
    step_fn = lambda: wl.step(args.batch)
    use_graph = not args.no_graph
  # Fused step, batches 128..2048: the frame copies of step n run beside the chain of
  # step n + 1 and are joined once per timed group of steps (b2r_set_deferred_frames).
  defer = lambda b: (wl.fused and not args.no_defer and
                     DEFER_MIN_BATCH <= b <= DEFER_MAX_BATCH and
                     (world == 1 or args.exchange == 'p2p'))
  finish_of = lambda b: wl.join if defer(b) else None
  wl.set_deferred(defer(args.batch))

  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.__enter__()
  per_graph = steps_per_graph(args.steps, args.steps_per_graph) if use_graph else 1
  timing = {}
  ms = time_graph_or_eager(torch, step_fn, args.steps, args.warmup, use_graph, dist,
                           per_graph=per_graph, min_ms=MIN_TIMED_MS, info=timing,
                           finish=finish_of(args.batch))
  wl.set_deferred(False)
  if rank == 0:
    clocks.__exit__()
  if dist is not None:
    t = torch.tensor([ms], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
  _native.check(_native.lib().b2r_check(wl.h, _native.current_stream()))
  shard_check = None
  if world > 1 and not os.environ.get('B2R_DEBUG_XCHG_NOWAIT'):
    # every rank has run the same steps: the last one's rows must partition the global
    # batch across the ranks (asserts; see ShardedStep.check_partition)
    counts = sharded.check_partition()
    shard_check = {'partition_of_global_batch': True, 'rows_per_rank': counts,
                   'what': 'after the timed region: all ranks\' strata of the last step '
                           'all-gathered, asserted to be a partition of range(%d), rows '
                           'asserted valid transitions of their shard' % (
                               args.batch * world)}
  transitions = args.batch * world * args.steps
  value = transitions / (ms * 1e-3)

  line = {
      'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': round(ms / args.steps, 6), 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
      'data': 'synthetic',
      'config': {
          'workload': workload_name(args.batch, args.capacity, world),
          'l2': 'inputs larger than L2: 7.06 GB frame ring per GPU, fresh random '
                'indices every step (device Philox); no flush needed',
          'launch': ('CUDA graph replay, %d consecutive steps per graph launch' % per_graph
                     if use_graph else 'eager launches') + (
              '; one b2r_train_step_device call per step: frame-stack copies on a '
              'forked stream beside loss + write-back' + (
                  ', joined once per graph launch (deferred: they also run beside the '
                  'NEXT step\'s sampler -> loss -> write-back)' if defer(args.batch)
                  else ', joined every step')
              if wl.fused and world == 1 else '') + (
                  '; shard totals exchanged over peer memory (NVLink) inside the '
                  'sampling kernel, no NCCL call on the path'
                  if world > 1 and args.exchange == 'p2p' else
                  '; NCCL all-gather of shard totals' if world > 1 else ''),
          'rng': 'device Philox4x32-10',
          'logits': 'inputs of the step (random, resident in HBM): the path is timed '
                    'without a network between gather and loss, as the metric names it; '
                    'full_train_step is the same kernels with the cuDNN networks between '
                    'them',
          'cpu_affinity': ('GPU-local NUMA node (%d cores per rank)' % numa_cpus
                           if numa_cpus else 'unbound'),
      },
      'gpu_launches': int(launches_per_step * args.steps),
      'clocks': clocks.summary(),
      'timing': dict(timing, what=(
          'exactly --steps steps per timed region (CUDA events, barrier + synchronize on '
          'both sides); regions shorter than %d ms are repeated until that much timed work '
          'has run and the median region is reported (max over ranks per region)'
          % MIN_TIMED_MS)),
  }
  # N > 1: the larger batches of config 3 / 4 (per-GPU batch b, global b * N)
  sweep_multi = None
  if world > 1 and sweep_batches:
    sweep_multi = {}
    for b in sweep_batches:
      st = sharded_replay.ShardedStep(wl, b * world, world, rank, dist,
                                      exchange=exchange)
      k = max(20, min(args.steps, 300))
      wl.set_deferred(defer(b))
      ms_b = time_graph_or_eager(torch, st.step, k, 5, use_graph, dist,
                                 per_graph=steps_per_graph(k, args.steps_per_graph),
                                 min_ms=MIN_TIMED_MS, finish=finish_of(b))
      wl.set_deferred(False)
      t = torch.tensor([ms_b], device='cuda')
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms_b = float(t.item())
      sweep_multi[str(b)] = {'global_batch': b * world, 'rows_per_rank_bound': st.max_rows,
                             'frame_copies': 'deferred' if defer(b) else 'joined every step',
                             'value': round(b * world * k / (ms_b * 1e-3), 1),
                             'ms_per_step': round(ms_b / k, 5)}
  # N > 1: the end-to-end loop and the full train step run on every rank together
  e2e_multi, full_multi = None, None
  if world > 1 and not args.no_e2e:
    # deeper pipeline than at N = 1: every step waits for the slowest rank, so the
    # queue has to absorb the host jitter of all ranks
    e2e_multi = measure_e2e(torch, wl, args.batch * world,
                            max(50, min(args.steps, 3000)), pipeline_depth=8,
                            dist=dist, world=world, rank=rank)
    e2e_multi['what'] += ('; %d ranks, each over its own shard, global batch %d, shard '
                          'totals over peer memory' % (world, args.batch * world))
    full = measure_full_train_step(torch, wl, args.batch, 100, ddp=True)
    full['transitions_per_s'] = round(full['transitions_per_s'] * world, 1)
    full['what'] += '; DistributedDataParallel over %d ranks (NCCL), one shard each' % world
    full_multi = {str(args.batch): full}
  if rank == 0:
    line['roofline'] = measure_gather_roofline(torch, wl, args.batch, peak_gbs)
    # the whole step (sampler + copies + loss + write-back) against the same peak:
    # algorithmic bytes of one step / device time of one step (N = 1: this rank's)
    try:
      step_gbs = (line['roofline']['algorithmic_bytes_per_launch'] /
                  (ms / args.steps * 1e-3) / 1e9)
      line['roofline']['whole_step_GBps'] = round(step_gbs, 1)
      line['roofline']['whole_step_frac'] = round(step_gbs / peak_gbs, 4)
    except (KeyError, ZeroDivisionError, TypeError):  # reporting only
      pass
    if shard_check is not None:
      line['shard_check'] = shard_check
    if sweep_multi is not None:
      line['sweep'] = sweep_multi
    if e2e_multi is not None:
      line['e2e'] = e2e_multi
      line['full_train_step'] = full_multi
    if sweep_batches and world == 1:
      sweep = {}
      for b in sweep_batches:
        k = max(20, min(args.steps, 400))
        wl.set_deferred(defer(b))
        ms_b = time_graph_or_eager(torch, lambda: wl.step(b), k, 5, use_graph,
                                   per_graph=steps_per_graph(k, args.steps_per_graph),
                                   min_ms=MIN_TIMED_MS, finish=finish_of(b))
        wl.set_deferred(False)
        roof = measure_gather_roofline(torch, wl, b, peak_gbs, launches=60)
        sweep[str(b)] = {'value': round(b * k / (ms_b * 1e-3), 1),
                         'ms_per_step': round(ms_b / k, 5),
                         'frame_copies': 'deferred' if defer(b) else 'joined every step',
                         'gather_GBps': roof['achieved'],
                         'gather_frac': roof['frac'],
                         'gather_us': roof['us_per_launch']}
        try:
          step_gbs = roof['algorithmic_bytes_per_launch'] / (ms_b / k * 1e-3) / 1e9
          sweep[str(b)]['whole_step_GBps'] = round(step_gbs, 1)
          sweep[str(b)]['whole_step_frac'] = round(step_gbs / peak_gbs, 4)
        except (KeyError, ZeroDivisionError, TypeError):  # reporting only
          pass
      line['sweep'] = sweep
    if not args.no_e2e and world == 1:
      line['e2e_host_batch'] = measure_e2e_host_batch(
          torch, wl, args.batch, max(50, min(args.steps, 300)))
      line['e2e_sync'] = measure_e2e(torch, wl, args.batch,
                                     max(50, min(args.steps, 2000)),
                                     pipeline_depth=0)
      line['e2e'] = measure_e2e(torch, wl, args.batch,
                                max(50, min(args.steps, 5000)))
    if not args.no_e2e and world == 1:
      line['full_train_step'] = {
          str(b): measure_full_train_step(torch, wl, b, 200 if b <= 256 else 50)
          for b in ([args.batch] + ([256] if sweep_batches else []))}
      line['full_train_step_graph'] = {
          str(b): measure_full_train_step(torch, wl, b, 200 if b <= 256 else 50,
                                          cuda_graph=True)
          for b in ([args.batch] + ([256] if sweep_batches else []))}
    if not args.no_cpu_baseline and world == 1:
      line['config1_dqn_uniform'] = measure_config1_dqn(
          torch, max(50, min(args.steps, 2000)))
      line['next_rows'] = measure_next_rows(torch)
      line['cpu_baseline'] = cpu_baseline(args.batch, args.capacity)
    print(json.dumps(ordered_line(line)))
  if dist is not None:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
