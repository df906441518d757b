"""Parity of the fused hot-path step (b2r_train_step_device / b2r_trainer_*) against
the oracle and against the separately tested sample / gather / loss / write-back
entry points.  Needs a B200: run with `pytest -m gpu`.
"""
import ctypes

import numpy as np
import pytest

from oracle import c51_port
from oracle import fast

pytestmark = pytest.mark.gpu

FRAME = 84 * 84
ACTIONS, ATOMS = 18, 51


@pytest.fixture(scope='module')
def gpu():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200 import _native
  from dopamine_b200.agents.rainbow import rainbow_agent
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb

  class Mods(object):
    pass

  m = Mods()
  m.torch, m.prb, m.ra, m.native = torch, prb, rainbow_agent, _native
  return m


def _filled(gpu, cap, batch, seed, horizon=3, term_p=0.01, hot=True):
  """A full, wrapped prioritized buffer + host mirrors of its columns and tree."""
  rng = np.random.RandomState(seed)
  mem = gpu.prb.OutOfGraphPrioritizedReplayBuffer(
      (84, 84), 4, cap, batch, update_horizon=horizon, gamma=0.99, output='torch',
      rng='device', seed=seed)
  pattern = rng.randint(0, 256, size=(1031, FRAME)).astype(np.uint8)
  obs = np.empty((cap, FRAME), dtype=np.uint8)
  for start in range(0, cap, 1031):
    n = min(1031, cap - start)
    obs[start:start + n] = pattern[:n]
  obs[:, :8] = np.arange(cap, dtype=np.int64).view(np.uint8).reshape(cap, 8)
  action = rng.randint(0, ACTIONS, size=cap).astype(np.int32)
  reward = np.clip(rng.randn(cap), -1, 1).astype(np.float32)
  terminal = (rng.rand(cap) < term_p).astype(np.uint8)
  mem._store['observation'] = obs.reshape(cap, 84, 84)
  mem._store['action'] = action
  mem._store['reward'] = reward
  mem._store['terminal'] = terminal
  add_count = cap + 4242
  mem.add_count = add_count
  inv = np.array([(4242 - horizon + i) % cap for i in range(4 + horizon)])
  mem.invalid_range = inv
  tree = fast.FastTree(cap)
  step = 50000
  for lo in range(0, cap, step):
    n = min(step, cap - lo)
    ids = np.arange(lo, lo + n, dtype=np.int32)
    pr = np.sqrt(np.abs(rng.randn(n)) + 1e-10).astype(np.float32)
    mem.set_priority(ids, pr)
    tree.set_seq(ids, pr.astype(np.float64))
  if hot:  # high-priority slots inside the invalid window: retries must happen
    ids = np.array(inv[:4], dtype=np.int32)
    mem.set_priority(ids, np.full(4, 0.02 * cap, dtype=np.float32))
    tree.set_seq(ids, np.full(4, 0.02 * cap))
  cols = dict(obs=obs, action=action, reward=reward, terminal=terminal, inv=inv,
              add_count=add_count)
  return mem, tree, cols


def _check_step(gpu, mem, tree, cols, cap, horizon, got, out, online, target):
  """One fused step's outputs against the oracles; applies the write-back to `tree`."""
  idx = got[7].cpu().numpy()
  assert all(fast.is_valid(int(i), cap, cols['add_count'], 4, horizon, cols['inv'],
                           cols['terminal']) for i in idx)
  want = fast.gather_u8(cap, FRAME, 4, horizon, mem._cumulative_discount_vector,
                        cols['obs'], cols['action'], cols['reward'],
                        cols['terminal'], idx)
  names = ['state', 'action', 'reward', 'next_state', 'next_action',
           'next_reward', 'terminal', 'indices']
  for nm, w, g in zip(names, want, got[:8]):
    assert w.tobytes() == g.cpu().numpy().reshape(w.shape).tobytes(), nm
  leaves = tree.level(tree.depth)
  probs = leaves[idx].astype(np.float32)
  assert got[8].cpu().numpy().tobytes() == probs.tobytes()
  ref = c51_port.rainbow_update(want[2], want[6], want[1], probs, online, target,
                                update_horizon=horizon)
  for key in ('loss', 'priorities', 'weights'):
    np.testing.assert_allclose(out[key].cpu().numpy(), ref[key], rtol=1e-6,
                               atol=1e-7, err_msg=key)
  pr = out['priorities'].cpu().numpy()
  tree.set_seq(idx, pr.astype(np.float64))
  return idx


@pytest.mark.parametrize('cap,batch', [(100000, 32), (100000, 48), (100000, 300),
                                       (200000, 512), (200000, 1024), (1000000, 4096)])
def test_fused_step_matches_oracles(gpu, cap, batch):
  """sample -> scalars -> C51 -> write-back in one call (frames on the forked
  stream): every batch column bit-exact against the C restatement at the sampled
  indices, loss / priorities / weights within 1e-6 of the numpy restatement, every
  fp64 tree node bit-exact after the write-backs.  (32: tail + write-back as one cluster;
  48: cluster sampler, write-back that groups ahead of its values; 300, 512: that
  write-back behind the many-CTA sampler; 1024: grouping on a side stream; 4096: thread
  sampler, unsplit loss, plain write-back.)"""
  torch = gpu.torch
  mem, tree, cols = _filled(gpu, cap, batch, seed=cap // 1000 + batch)
  rng = np.random.RandomState(5)
  support = gpu.ra.make_support(10., ATOMS)
  out = None
  seen = set()
  for step in range(3):
    online = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    target = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    got, out = gpu.ra.train_step(mem, torch.as_tensor(online, device='cuda'),
                                 torch.as_tensor(target, device='cuda'), support,
                                 0.99 ** 3, out=out)
    torch.cuda.synchronize()
    idx = _check_step(gpu, mem, tree, cols, cap, 3, got, out, online, target)
    seen.add(tuple(idx[:8].tolist()))
  assert len(seen) == 3  # fresh strata every step
  gpu.native.check(gpu.native.lib().b2r_check(mem._h, gpu.native.current_stream()))
  for l, level in enumerate(mem.sum_tree.nodes):
    assert np.array_equal(level.view(np.uint64), tree.level(l).view(np.uint64)), l


@pytest.mark.parametrize('batch', [32, 1024])
def test_fused_step_equals_separate_calls(gpu, batch):
  """Same seed, twin buffers: the fused call and the sequence sample_transition_batch
  -> c51_loss -> set_priority produce identical bits (indices, batch, losses, tree)."""
  torch = gpu.torch
  cap = 100000
  mem_a, _, _ = _filled(gpu, cap, batch, seed=11)
  mem_b, _, _ = _filled(gpu, cap, batch, seed=11)
  rng = np.random.RandomState(2)
  support = gpu.ra.make_support(10., ATOMS)
  for step in range(3):
    online = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                             device='cuda')
    target = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                             device='cuda')
    got_a, out_a = gpu.ra.train_step(mem_a, online, target, support, 0.99 ** 3)
    got_b = mem_b.sample_transition_batch(batch)
    out_b = gpu.ra.c51_loss(online, target, got_b[1], got_b[2], got_b[6], got_b[8],
                            support, 0.99 ** 3, want_mean=False)
    mem_b.set_priority(got_b[7], out_b['priorities'])
    torch.cuda.synchronize()
    for a, b in zip(got_a, got_b):
      assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()
    for key in ('loss', 'priorities', 'weights'):
      assert (out_a[key].cpu().numpy().tobytes() ==
              out_b[key].cpu().numpy().tobytes()), key
  for la, lb in zip(mem_a.sum_tree.nodes, mem_b.sum_tree.nodes):
    assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))


@pytest.mark.parametrize('batch', [32, 600])
def test_deferred_frame_copies_equal_joined_steps(gpu, batch):
  """b2r_set_deferred_frames: the frame copies of a step are joined by b2r_join_frames
  (or by the next flush of staged adds) instead of by the call itself, so the next
  step's chain runs beside them.  Twin buffers, same seed: every step's batch — frames
  included, read after the join at the END of the run — losses and the final tree are
  bit-identical to the joined mode; an add() between two steps joins by itself."""
  torch, native = gpu.torch, gpu.native
  lib = native.lib()
  cap = 100000
  mem_a, _, _ = _filled(gpu, cap, batch, seed=12)
  mem_b, _, _ = _filled(gpu, cap, batch, seed=12)
  native.check(lib.b2r_set_deferred_frames(mem_a._h, 1))
  rng = np.random.RandomState(3)
  support = gpu.ra.make_support(10., ATOMS)
  got_a, got_b = [], []
  for step in range(5):
    online = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                             device='cuda')
    target = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                             device='cuda')
    got_a.append(gpu.ra.train_step(mem_a, online, target, support, 0.99 ** 3))
    got_b.append(gpu.ra.train_step(mem_b, online, target, support, 0.99 ** 3))
    if step == 2:  # staged adds: their flush waits for the copies still in flight
      for mem in (mem_a, mem_b):
        for k in range(3):
          mem.add(np.full((84, 84), 10 * step + k, np.uint8), k, 0.5, 0, 1.0)
  native.check(lib.b2r_join_frames(mem_a._h, native.current_stream()))
  torch.cuda.synchronize()
  for (batch_a, out_a), (batch_b, out_b) in zip(got_a, got_b):
    for a, b in zip(batch_a, batch_b):
      assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()
    for key in ('loss', 'priorities', 'weights'):
      assert out_a[key].cpu().numpy().tobytes() == out_b[key].cpu().numpy().tobytes()
  for la, lb in zip(mem_a.sum_tree.nodes, mem_b.sum_tree.nodes):
    assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))
  native.check(lib.b2r_check(mem_a._h, native.current_stream()))
  native.check(lib.b2r_set_deferred_frames(mem_a._h, 0))


def test_fused_step_in_cuda_graph(gpu):
  """The forked frame copies survive stream capture: a replayed graph keeps drawing
  fresh strata and stays equal to eager execution on a twin buffer."""
  torch = gpu.torch
  cap, batch = 100000, 64
  mem_a, _, _ = _filled(gpu, cap, batch, seed=3)
  mem_b, _, _ = _filled(gpu, cap, batch, seed=3)
  rng = np.random.RandomState(9)
  support = gpu.ra.make_support(10., ATOMS)
  online = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                           device='cuda')
  target = torch.as_tensor(rng.randn(batch, ACTIONS, ATOMS).astype(np.float32),
                           device='cuda')
  mem_a._reuse_outputs = mem_b._reuse_outputs = True
  side = torch.cuda.Stream()
  side.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(side):
    got_a, out_a = gpu.ra.train_step(mem_a, online, target, support, 0.99 ** 3)
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
      gpu.ra.train_step(mem_a, online, target, support, 0.99 ** 3, out=out_a)
    for _ in range(3):
      graph.replay()
    side.synchronize()
  torch.cuda.current_stream().wait_stream(side)
  # mem_a ran 1 eager step (host offset 1) + 3 replays of a graph captured with host
  # offset 2; the device draw counter advanced 0, 1, 2, 3 underneath.
  out_b = None
  for host_offset in (1, 2, 2, 2):
    mem_b._draw_counter = host_offset - 1
    got_b, out_b = gpu.ra.train_step(mem_b, online, target, support, 0.99 ** 3,
                                     out=out_b)
  torch.cuda.synchronize()
  for a, b in zip(got_a, got_b):
    assert a.cpu().numpy().tobytes() == b.cpu().numpy().tobytes()
  for la, lb in zip(mem_a.sum_tree.nodes, mem_b.sum_tree.nodes):
    assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))


@pytest.mark.parametrize('depth,graph,pinned,batch', [
    (0, False, False, 32), (2, False, False, 32), (2, True, False, 32),
    (0, False, True, 32), (1, False, True, 32), (2, False, True, 32), (3, False, True, 200)])
def test_trainer_host_steps_match_oracles(gpu, depth, graph, pinned, batch):
  """The pipelined host-facing trainer: adds between steps, logits from host
  memory, losses handed back `depth` calls later; batch, losses and tree checked
  against the oracles step by step.  pinned: the logits live in page-locked memory —
  the kernels then read them in place and write the losses into the result ring
  themselves (no copy calls: DirectIO in step.cu; default only at depth 0, forced here
  for every depth); pageable logits take the copies."""
  import os
  torch = gpu.torch
  if pinned:
    os.environ['B2R_TRAINER_DIRECT'] = '1'
  else:
    os.environ.pop('B2R_TRAINER_DIRECT', None)
  cap = 50000
  mem, tree, cols = _filled(gpu, cap, batch, seed=21, hot=False)
  trainer = gpu.ra.ReplayTrainer(mem, ACTIONS, ATOMS, 10., pipeline_depth=depth,
                                 seed=21, use_graph=graph)
  rng = np.random.RandomState(4)
  inputs, losses = [], {}
  # (page-locked buffers are the caller's until the step has run: one pair per step)
  pin = [(torch.empty(batch, ACTIONS, ATOMS, dtype=torch.float32).pin_memory(),
          torch.empty(batch, ACTIONS, ATOMS, dtype=torch.float32).pin_memory())
         for _ in range(6)] if pinned else None
  for step in range(6):
    online = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    target = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    inputs.append((online, target))
    if pinned:
      pin[step][0].copy_(torch.from_numpy(online))
      pin[step][1].copy_(torch.from_numpy(target))
      loss, done = trainer.step(pin[step][0], pin[step][1])
    else:
      loss, done = trainer.step(online, target)
    assert done == step - depth if step >= depth else done == -1
    if done >= 0:
      losses[done] = loss
    # Every step is checked as it completes (the views are only stable then).
    torch.cuda.synchronize()
    transition, out = trainer.views()
    got = [transition[k] for k in ('state', 'action', 'reward', 'next_state',
                                   'next_action', 'next_reward', 'terminal',
                                   'indices', 'sampling_probabilities')]
    _check_step(gpu, mem, tree, cols, cap, 3, got, out, online, target)
    losses.setdefault(('device', step), out['loss'].cpu().numpy().copy())
  loss, done = trainer.drain()
  assert done == 5
  losses[done] = loss
  for step in range(6):
    if step in losses:
      assert losses[step].tobytes() == losses[('device', step)].tobytes(), step
  assert 5 - depth in losses
  os.environ.pop('B2R_TRAINER_DIRECT', None)
  for l, level in enumerate(mem.sum_tree.nodes):
    assert np.array_equal(level.view(np.uint64), tree.level(l).view(np.uint64)), l


@pytest.mark.parametrize('graph', [False, True])
def test_trainer_applies_adds_between_steps(gpu, graph):
  """add() rows staged between trainer steps are in HBM (rows, priorities, validity
  window) before the next step samples."""
  from oracle.replay_port import PortPrioritizedReplay
  torch = gpu.torch
  cap, batch = 256, 16
  prb = gpu.prb
  mem = prb.OutOfGraphPrioritizedReplayBuffer((84, 84), 4, cap, batch,
                                              update_horizon=3, gamma=0.99,
                                              output='torch', rng='device', seed=1)
  port = PortPrioritizedReplay((84, 84), 4, cap, batch, update_horizon=3, gamma=0.99)
  rng = np.random.RandomState(0)
  trainer = gpu.ra.ReplayTrainer(mem, ACTIONS, ATOMS, 10., pipeline_depth=1, seed=1,
                                 use_graph=graph)
  added = 0

  def add_rows(n):
    nonlocal added
    for _ in range(n):
      row = (rng.randint(0, 256, size=(84, 84)).astype(np.uint8), rng.randint(18),
             np.float32(np.clip(rng.randn(), -1, 1)), int(rng.rand() < 0.05))
      port.add(*row, port.sum_tree.max_recorded_priority)
      mem.add(*row, prb.MAX_RECORDED_PRIORITY)
      added += 1

  add_rows(100)
  for step in range(40):
    add_rows(4)
    online = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    target = rng.randn(batch, ACTIONS, ATOMS).astype(np.float32)
    trainer.step(online, target)
    torch.cuda.synchronize()
    transition, out = trainer.views()
    idx = transition['indices'].cpu().numpy()
    assert all(port.is_valid_transition(int(i)) for i in idx), step
    want = port.sample_transition_batch(batch, indices=idx.tolist())
    names = ['state', 'action', 'reward', 'next_state', 'next_action',
             'next_reward', 'terminal', 'indices', 'sampling_probabilities']
    for nm, w in zip(names, want):
      assert w.tobytes() == transition[nm].cpu().numpy().reshape(w.shape).tobytes(), nm
    port.set_priority(idx, out['priorities'].cpu().numpy())
  trainer.drain()
  for a, b in zip(mem.sum_tree.nodes, port.sum_tree.nodes):
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


# ------------------------------------------------- peer-memory exchange ----
def _shard_pairs(gpu, num_shards, rng, cap=400, shape=(8, 8), attempts=64,
                 term_p=0.1):
  from oracle.replay_port import PortPrioritizedReplay
  from tests.test_gpu_parity import _fill_pair
  kw = dict(update_horizon=3, gamma=0.99, max_sample_attempts=attempts)
  ours, ports = [], []
  for g in range(num_shards):
    o = gpu.prb.OutOfGraphPrioritizedReplayBuffer(shape, 4, cap, 8, output='torch',
                                                  **kw)
    p = PortPrioritizedReplay(shape, 4, cap, 8, **kw)
    _fill_pair(rng, o, p, 300 + 57 * g, shape, True, term_p=term_p)
    ours.append(o)
    ports.append(p)
  return ours, ports


@pytest.mark.parametrize('num_shards,global_batch', [(2, 32), (8, 256), (4, 1024)])
def test_peer_exchange_sampling_matches_oracle(gpu, num_shards, global_batch):
  """Shard totals through the peer-memory mailboxes (ranks emulated on one device:
  every rank publishes, then every rank samples): over several steps with the
  priorities changing in between, each rank must serve exactly the strata and
  indices of the CPU statement of the rule — i.e. it saw every peer's CURRENT total
  (both mailbox parities are exercised)."""
  from dopamine_b200.replay_memory import sharded_replay
  from oracle import sharded_port
  torch = gpu.torch
  rng = np.random.RandomState(num_shards * 100 + global_batch)
  ours, ports = _shard_pairs(gpu, num_shards, rng)
  exchanges = sharded_replay.PeerExchange.emulated(num_shards)
  shards = [sharded_replay.ShardedPrioritizedReplay(
      ours[g], rank=g, world_size=num_shards, exchange=exchanges[g])
            for g in range(num_shards)]
  for step in range(5):
    for g in range(num_shards):  # priorities move, so the totals differ every step
      ids = rng.randint(0, 300, size=50).astype(np.int32)
      pr = (np.sqrt(np.abs(rng.randn(50)) + 1e-10) * (1 + g + step)).astype(np.float32)
      ours[g].set_priority(ids, pr)
      ports[g].set_priority(ids, pr)
    bounds = np.linspace(0., 1., global_batch + 1)
    queries = bounds[:-1] + (bounds[1:] - bounds[:-1]) * rng.rand(global_batch)
    d_queries = torch.as_tensor(queries, device='cuda')
    retries = [rng.rand(max(64, global_batch)) for _ in range(num_shards)]
    want = sharded_port.sharded_sample(ports, queries, retries)
    for g in range(num_shards):
      exchanges[g].publish(ours[g])
    served = []
    for g in range(num_shards):
      slots, idx, count = shards[g].sample_index_batch(
          global_batch, queries01=d_queries,
          retry_u01=torch.as_tensor(retries[g], device='cuda'))
      n = int(count.cpu()[0])
      w_slots, w_idx, _ = want[g]
      assert n == len(w_slots), (step, g)
      assert slots[:n].cpu().numpy().tolist() == w_slots
      assert idx[:n].cpu().numpy().tolist() == [int(i) for i in w_idx]
      served += w_slots
      gpu.native.check(gpu.native.lib().b2r_check(ours[g]._h,
                                                  gpu.native.current_stream()))
    assert sorted(served) == list(range(global_batch))


def test_peer_exchange_times_out_instead_of_hanging(gpu):
  """A peer that never publishes must latch B2R_ERR_EXCHANGE after the timeout, and
  later calls must not wait again."""
  import time
  from dopamine_b200.replay_memory import sharded_replay
  torch = gpu.torch
  rng = np.random.RandomState(1)
  ours, _ = _shard_pairs(gpu, 2, rng)
  exchanges = sharded_replay.PeerExchange.emulated(2)
  lib = gpu.native.lib()
  gpu.native.check(lib.b2r_exchange_set_timeout(exchanges[0]._h, 0.05))
  shard = sharded_replay.ShardedPrioritizedReplay(ours[0], rank=0, world_size=2,
                                                  exchange=exchanges[0])
  t0 = time.perf_counter()
  shard.sample_index_batch(32)   # rank 1 never published
  torch.cuda.synchronize()
  first = time.perf_counter() - t0
  assert 0.04 < first < 10.0
  t0 = time.perf_counter()
  for _ in range(20):
    shard.sample_index_batch(32)
  torch.cuda.synchronize()
  assert time.perf_counter() - t0 < 0.5 * 20 * 0.05 + 2.0  # latched: no further waits
  status = lib.b2r_check(ours[0]._h, gpu.native.current_stream())
  assert status == gpu.native.ERR_EXCHANGE
  assert 'did not publish' in gpu.native.last_error()


@pytest.mark.parametrize('global_batch,bounded,deferred,early', [
    (256, False, False, False), (2048, False, False, False), (256, True, False, False),
    (2048, True, False, False), (256, True, True, False), (2048, True, True, False),
    (128, True, True, True), (96, True, False, True), (64, True, True, True),
    (96, True, False, False), (128, False, False, False),
    # expected share 1024 rows, the hot rank's beyond it: the early write-back takes the
    # first 1024 entries, the plain kernel behind it the rest
    (4096, True, False, False), (4096, True, True, False)])
def test_sharded_fused_step_matches_oracles(gpu, global_batch, bounded, deferred, early):
  """b2r_train_step_sharded_device on 4 emulated ranks (one sampling CTA up to a
  global batch of 256, tiles over each rank's stratum range above): each rank's rows
  (batch columns at its indices, losses, write-back) against the oracles, the rows of
  all ranks partitioning the global batch.  bounded: outputs, logits and launches sized
  by a bound on the rank's share (max_rows) instead of the global batch.  deferred: frame
  copies joined by b2r_join_frames (the row count then travels through the ring slot).
  early: the write-back publishes the shard total for the next step itself
  (b2r_exchange_set_early_publish): only the first step needs the totals published ahead
  (the ranks are emulated one after the other), and the later steps must still see the
  partition of the global batch — i.e. every rank read the totals its peers left behind
  at the end of their previous step.  (Expected shares of at most 32 rows: the one-CTA
  tree kernels — and the loss tail's cluster, which applies the write-back of such a
  step itself — carry the publish hook; larger write-backs leave publishing to the
  sampler, which ranks emulated one after the other cannot wait for.)"""
  import ctypes
  from dopamine_b200.replay_memory import sharded_replay
  torch, native = gpu.torch, gpu.native
  lib = native.lib()
  num_shards, cap = 4, 50000
  shards = [_filled(gpu, cap, 32, seed=40 + g, hot=(g == 1)) for g in range(num_shards)]
  exchanges = sharded_replay.PeerExchange.emulated(num_shards)
  for x in exchanges:
    x.set_early_publish(early)
  rng = np.random.RandomState(8)
  support = gpu.ra.make_support(10., ATOMS)
  for mem, _, _ in shards:
    native.check(lib.b2r_set_deferred_frames(mem._h, 1 if deferred else 0))
  outs = []
  # (shard 1 is "hot": it serves well over its even share)
  rows = global_batch * 5 // 8 if bounded else global_batch
  for g in range(num_shards):
    mem = shards[g][0]
    _, arrays, batch = mem._alloc_outputs(rows, True)
    outs.append(dict(
        arrays=arrays, batch=batch,
        loss={k: torch.zeros(rows, dtype=torch.float32, device='cuda')
              for k in ('loss', 'priorities', 'weights')},
        slots=torch.zeros(rows, dtype=torch.int32, device='cuda'),
        count=torch.zeros(1, dtype=torch.int32, device='cuda')))
  for step in range(3):
    online = rng.randn(rows, ACTIONS, ATOMS).astype(np.float32)
    target = rng.randn(rows, ACTIONS, ATOMS).astype(np.float32)
    d_online = torch.as_tensor(online, device='cuda')
    d_target = torch.as_tensor(target, device='cuda')
    for g in range(num_shards):
      if not early or step == 0:
        exchanges[g].publish(shards[g][0])
    served = []
    for g in range(num_shards):
      mem, tree, cols = shards[g]
      o = outs[g]
      args = native.C51Args()
      args.batch, args.num_actions, args.num_atoms = global_batch, ACTIONS, ATOMS
      args.cumulative_gamma = float(np.float32(0.99 ** 3))
      args.support = support.data_ptr()
      args.online_logits, args.target_logits = d_online.data_ptr(), d_target.data_ptr()
      args.loss = o['loss']['loss'].data_ptr()
      args.priorities = o['loss']['priorities'].data_ptr()
      args.weights = o['loss']['weights'].data_ptr()
      native.check(lib.b2r_train_step_sharded_device(
          mem._h, exchanges[g]._h, global_batch, 77, step, ctypes.byref(o['batch']),
          ctypes.byref(args), o['slots'].data_ptr(), o['count'].data_ptr(),
          rows if bounded else 0, native.current_stream()))
      native.check(lib.b2r_join_frames(mem._h, native.current_stream()))
      torch.cuda.synchronize()
      n = int(o['count'].cpu()[0])
      assert n <= rows
      served += o['slots'][:n].cpu().numpy().tolist()
      if n == 0:
        continue
      got = [a[:n] for a in o['arrays']]
      loss = {k: v[:n] for k, v in o['loss'].items()}
      _check_step(gpu, mem, tree, cols, cap, 3, got, loss, online[:n], target[:n])
      native.check(lib.b2r_check(mem._h, native.current_stream()))
    assert sorted(served) == list(range(global_batch)), step
  for mem, tree, _ in shards:
    for l, level in enumerate(mem.sum_tree.nodes):
      assert np.array_equal(level.view(np.uint64), tree.level(l).view(np.uint64)), l


@pytest.mark.parametrize('global_batch', [64, 1024, 16384])
def test_sharded_step_share_beyond_max_rows_is_latched(gpu, global_batch):
  """A rank whose share of the global batch exceeds max_rows serves exactly max_rows
  rows (its first strata), writes nothing beyond its buffers (canary rows stay) and
  latches B2R_ERR_UNSUPPORTED.  64: one sampling CTA; 1024: warp sampler over the rank's
  range; 16384: tiled thread sampler."""
  import ctypes
  from dopamine_b200.replay_memory import sharded_replay
  torch, native = gpu.torch, gpu.native
  lib = native.lib()
  world, cap = 2, 20000
  shards = [_filled(gpu, cap, 32, seed=70 + g, hot=False) for g in range(world)]
  exchanges = sharded_replay.PeerExchange.emulated(world)
  rows = global_batch // 4  # each rank's share is about half of the batch
  support = gpu.ra.make_support(10., ATOMS)
  for g in range(world):
    exchanges[g].publish(shards[g][0])
  for g in range(world):
    mem = shards[g][0]
    _, arrays, batch = mem._alloc_outputs(rows + 8, True)
    canary = torch.full((rows + 8,), -7, dtype=torch.int32, device='cuda')
    arrays[7].copy_(canary)  # indices
    slots = canary.clone()
    count = torch.zeros(1, dtype=torch.int32, device='cuda')
    logits = torch.zeros(rows + 8, ACTIONS, ATOMS, dtype=torch.float32, device='cuda')
    loss = {k: torch.full((rows + 8,), -7., dtype=torch.float32, device='cuda')
            for k in ('loss', 'priorities', 'weights')}
    args = native.C51Args()
    args.batch, args.num_actions, args.num_atoms = global_batch, ACTIONS, ATOMS
    args.cumulative_gamma = float(np.float32(0.99 ** 3))
    args.support = support.data_ptr()
    args.online_logits = args.target_logits = logits.data_ptr()
    args.loss = loss['loss'].data_ptr()
    args.priorities = loss['priorities'].data_ptr()
    args.weights = loss['weights'].data_ptr()
    native.check(lib.b2r_train_step_sharded_device(
        mem._h, exchanges[g]._h, global_batch, 5, 0, ctypes.byref(batch),
        ctypes.byref(args), slots.data_ptr(), count.data_ptr(), rows,
        native.current_stream()))
    torch.cuda.synchronize()
    assert int(count.cpu()[0]) == rows
    got_slots = slots.cpu().numpy()
    first = 0 if g == 0 else int(got_slots[0])
    assert np.array_equal(got_slots[:rows], np.arange(first, first + rows))
    assert (got_slots[rows:] == -7).all()
    assert (arrays[7].cpu().numpy()[rows:] == -7).all()
    assert (loss['loss'].cpu().numpy()[rows:] == -7.).all()
    assert (loss['loss'].cpu().numpy()[:rows] > 0.).all()
    status = lib.b2r_check(mem._h, native.current_stream())
    assert status == native.ERR_UNSUPPORTED, (status, native.last_error())


def _sharded_step_args(gpu, mem, rows, support, logits):
  import ctypes
  torch, native = gpu.torch, gpu.native
  _, arrays, batch = mem._alloc_outputs(rows, True)
  loss = {k: torch.zeros(rows, dtype=torch.float32, device='cuda')
          for k in ('loss', 'priorities', 'weights')}
  args = native.C51Args()
  args.num_actions, args.num_atoms = ACTIONS, ATOMS
  args.cumulative_gamma = float(np.float32(0.99 ** 3))
  args.support = support.data_ptr()
  args.online_logits = args.target_logits = logits.data_ptr()
  args.loss = loss['loss'].data_ptr()
  args.priorities = loss['priorities'].data_ptr()
  args.weights = loss['weights'].data_ptr()
  return dict(arrays=arrays, batch=batch, loss=loss, args=args,
              slots=torch.zeros(rows, dtype=torch.int32, device='cuda'),
              count=torch.zeros(1, dtype=torch.int32, device='cuda'), ctypes=ctypes)


def test_early_publish_applies_staged_adds_behind_the_write_back(gpu):
  """b2r_exchange_set_early_publish: the kernel that writes the tree last in a step call
  publishes the shard total for the next step, so add()s staged before a call are
  applied at the END of that call, behind its write-back.  Twin shards, two emulated
  ranks each: A (early) stages its adds and THEN steps; B (default) steps, then adds the
  same rows and flushes.  The effects on the tree are the same sequence — sample,
  set_priority, add, sample, ... — so every step samples the same indices, the rings hold
  the same rows and every fp64 tree node is equal; in A only the first step's totals are
  published ahead (ranks emulated one after the other: the later steps find the totals
  their peers' write-backs / flushes left behind)."""
  from dopamine_b200.replay_memory import sharded_replay
  torch, native = gpu.torch, gpu.native
  lib = native.lib()
  world, cap, global_batch, rows = 2, 20000, 64, 64
  support = gpu.ra.make_support(10., ATOMS)
  logits = torch.as_tensor(np.random.RandomState(1).randn(rows, ACTIONS, ATOMS)
                           .astype(np.float32), device='cuda')
  runs = {}
  for name, early in (('A', True), ('B', False)):
    shards = [_filled(gpu, cap, 32, seed=80 + g, hot=False)[0] for g in range(world)]
    xs = sharded_replay.PeerExchange.emulated(world)
    for x in xs:
      x.set_early_publish(early)
    state = [_sharded_step_args(gpu, m, rows, support, logits) for m in shards]
    picked = []
    rng = np.random.RandomState(9)
    for step in range(4):
      new_rows = [(rng.randint(0, 256, size=(84, 84)).astype(np.uint8), int(rng.randint(18)),
                   0.25, int(rng.rand() < 0.2), float(1.0 + rng.rand())) for _ in range(6)]
      for g in range(world):
        if not early or step == 0:
          xs[g].publish(shards[g])
      if early:  # staged now, applied behind this call's write-back
        for m in shards:
          for row in new_rows:
            m.add(*row)
      for g in range(world):
        st = state[g]
        st['args'].batch = global_batch
        native.check(lib.b2r_train_step_sharded_device(
            shards[g]._h, xs[g]._h, global_batch, 31, step, st['ctypes'].byref(st['batch']),
            st['ctypes'].byref(st['args']), st['slots'].data_ptr(), st['count'].data_ptr(),
            rows, native.current_stream()))
        torch.cuda.synchronize()
        n = int(st['count'].cpu()[0])
        picked.append(st['arrays'][7][:n].cpu().numpy().copy())
      if not early:
        for m in shards:
          for row in new_rows:
            m.add(*row)
          m._flush()
    for m in shards:
      native.check(lib.b2r_check(m._h, native.current_stream()))
    runs[name] = (shards, picked)
  (shards_a, picked_a), (shards_b, picked_b) = runs['A'], runs['B']
  assert len(picked_a) == len(picked_b) == 8
  for ia, ib in zip(picked_a, picked_b):
    assert np.array_equal(ia, ib)
  for ma, mb in zip(shards_a, shards_b):
    assert ma.add_count == mb.add_count
    assert list(ma.invalid_range) == list(mb.invalid_range)
    for la, lb in zip(ma.sum_tree.nodes, mb.sum_tree.nodes):
      assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))
    lo = int(mb.cursor()) - 40
    assert (ma._read_column('observation', lo, 40).tobytes() ==
            mb._read_column('observation', lo, 40).tobytes())


def test_early_publish_latches_a_tree_change_behind_its_back(gpu):
  """With early publish, nothing but the sharded steps may change a shard's tree between
  two steps: a set_priority in between makes the next sampler find a root different from
  the total its peers were given, and it latches B2R_ERR_STALE_TOTAL instead of letting
  the ranks apportion the batch from different totals."""
  from dopamine_b200.replay_memory import sharded_replay
  torch, native = gpu.torch, gpu.native
  lib = native.lib()
  world, cap, global_batch, rows = 2, 20000, 64, 64
  support = gpu.ra.make_support(10., ATOMS)
  logits = torch.zeros(rows, ACTIONS, ATOMS, dtype=torch.float32, device='cuda')
  shards = [_filled(gpu, cap, 32, seed=90 + g, hot=False)[0] for g in range(world)]
  xs = sharded_replay.PeerExchange.emulated(world)
  for x in xs:
    x.set_early_publish(True)
  state = [_sharded_step_args(gpu, m, rows, support, logits) for m in shards]

  def step_all(k):
    for g in range(world):
      st = state[g]
      st['args'].batch = global_batch
      native.check(lib.b2r_train_step_sharded_device(
          shards[g]._h, xs[g]._h, global_batch, 31, k, st['ctypes'].byref(st['batch']),
          st['ctypes'].byref(st['args']), st['slots'].data_ptr(), st['count'].data_ptr(),
          rows, native.current_stream()))
    torch.cuda.synchronize()

  for g in range(world):
    xs[g].publish(shards[g])
  step_all(0)
  step_all(1)
  for m in shards:
    native.check(lib.b2r_check(m._h, native.current_stream()))
  shards[0].set_priority(np.array([5, 6], np.int32), np.array([3.0, 4.0], np.float32))
  step_all(2)
  status = lib.b2r_check(shards[0]._h, native.current_stream())
  assert status == native.ERR_STALE_TOTAL
  assert 'changed between the early publish' in native.last_error()
  native.check(lib.b2r_check(shards[1]._h, native.current_stream()))


@pytest.mark.parametrize('variant', ['tma', 'reg'])
def test_each_gather_variant_passes_the_gather_parity_suite(variant):
  """The frame copies have two kernels: the TMA-staged one (cp.async.bulk into shared
  memory on an mbarrier, the default) and the register one (LDG.128 -> PRMT -> STG.128;
  gather.cu: gather_variant).  B2R_GATHER forces one for a whole
  process, so the gather parity tests are re-run in a child process with each (bit-exact
  batches at 100k / 1M, wrap-around, terminals inside trajectories, fused and sharded
  steps) — every size goes through both kernels."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, B2R_GATHER=variant)
  out = subprocess.run(
      [sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu',
       'tests/test_gpu_parity.py', 'tests/test_gpu_step.py', '-k',
       'gather_full_size or prioritized_full_size or uniform_reference_fixture or '
       'prioritized_reference_fixture or fused_step_matches_oracles or '
       'sharded_fused_step or deferred_frame_copies or host_batches'],
      cwd=root, env=env, capture_output=True, text=True, timeout=900)
  assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
  assert ' passed' in out.stdout


def test_sharded_trainer_matches_oracles(gpu):
  """The host-facing trainer on two emulated shards: every update's local rows (batch
  columns, losses handed back through pinned memory, write-back) against the
  oracles; the two ranks' rows partition the global batch."""
  from dopamine_b200.replay_memory import sharded_replay
  torch = gpu.torch
  cap, global_batch, world = 50000, 64, 2
  shards = [_filled(gpu, cap, 32, seed=60 + g, hot=False) for g in range(world)]
  exchanges = sharded_replay.PeerExchange.emulated(world)
  trainers = []
  for g in range(world):
    t = gpu.ra.ReplayTrainer(shards[g][0], ACTIONS, ATOMS, 10., batch_size=global_batch,
                             pipeline_depth=0, seed=5)
    t.set_exchange(exchanges[g])
    trainers.append(t)
  rng = np.random.RandomState(6)
  for step in range(4):
    online = rng.randn(global_batch, ACTIONS, ATOMS).astype(np.float32)
    target = rng.randn(global_batch, ACTIONS, ATOMS).astype(np.float32)
    for g in range(world):
      exchanges[g].publish(shards[g][0])
    total_rows = 0
    for g in range(world):
      mem, tree, cols = shards[g]
      loss, done = trainers[g].step(online, target)
      assert done == step
      n = trainers[g].last_rows
      total_rows += n
      torch.cuda.synchronize()
      transition, out = trainers[g].views()
      got = [transition[k][:n] for k in (
          'state', 'action', 'reward', 'next_state', 'next_action', 'next_reward',
          'terminal', 'indices', 'sampling_probabilities')]
      dev = {k: v[:n] for k, v in out.items()}
      if n:
        _check_step(gpu, mem, tree, cols, cap, 3, got, dev, online[:n], target[:n])
        assert loss[:n].tobytes() == dev['loss'].cpu().numpy().tobytes()
      gpu.native.check(gpu.native.lib().b2r_check(mem._h, gpu.native.current_stream()))
    assert total_rows == global_batch, step
  for mem, tree, _ in shards:
    for l, level in enumerate(mem.sum_tree.nodes):
      assert np.array_equal(level.view(np.uint64), tree.level(l).view(np.uint64)), l


@pytest.mark.parametrize('num_shards,global_batch', [(4, 128), (4, 2048), (8, 4096),
                                                     (2, 300)])
def test_sharded_device_rng_matches_oracle(gpu, num_shards, global_batch):
  """Throughput mode (Philox strata and retries drawn on the device), single-CTA
  compaction (<= 256 strata) and the multi-CTA range search above: with the same
  uniforms from the numpy Philox port, every emulated rank must serve exactly the
  strata and indices of the CPU statement of the rule, step after step."""
  from dopamine_b200.replay_memory import sharded_replay
  from oracle import philox_port
  from oracle import sharded_port
  torch = gpu.torch
  rng = np.random.RandomState(num_shards + global_batch)
  ours, ports = _shard_pairs(gpu, num_shards, rng, cap=3000, attempts=1500,
                             term_p=0.03)
  for g in range(num_shards):  # uneven shards
    ids = rng.randint(0, 300, size=200).astype(np.int32)
    pr = (np.sqrt(np.abs(rng.randn(200)) + 1e-10) * (1 + 2 * g)).astype(np.float32)
    ours[g].set_priority(ids, pr)
    ports[g].set_priority(ids, pr)
  exchanges = sharded_replay.PeerExchange.emulated(num_shards)
  seed = 31
  shards = [sharded_replay.ShardedPrioritizedReplay(
      ours[g], rank=g, world_size=num_shards, exchange=exchanges[g], seed=seed)
            for g in range(num_shards)]
  budget = ours[0]._max_sample_attempts
  for step in range(3):
    queries = philox_port.stratified_queries(seed, step, global_batch)
    retries = [philox_port.retry_uniforms(seed, g, step, global_batch, budget)
               for g in range(num_shards)]
    want = sharded_port.sharded_sample(ports, queries, retries)
    for g in range(num_shards):
      exchanges[g].publish(ours[g])
    served = []
    for g in range(num_shards):
      slots, idx, count = shards[g].sample_index_batch(global_batch)
      n = int(count.cpu()[0])
      w_slots, w_idx, _ = want[g]
      assert n == len(w_slots), (step, g, n, len(w_slots))
      assert slots[:n].cpu().numpy().tolist() == w_slots
      assert idx[:n].cpu().numpy().tolist() == [int(i) for i in w_idx]
      served += w_slots
      gpu.native.check(gpu.native.lib().b2r_check(ours[g]._h,
                                                  gpu.native.current_stream()))
    assert sorted(served) == list(range(global_batch))
    for g in range(num_shards):  # the totals move between steps
      ids = rng.randint(0, 300, size=20).astype(np.int32)
      pr = np.sqrt(np.abs(rng.randn(20)) + 1e-10).astype(np.float32)
      ours[g].set_priority(ids, pr)
      ports[g].set_priority(ids, pr)


@pytest.mark.parametrize('batch', [32, 256, 1000, 4096])
def test_device_rng_sampling_matches_oracle(gpu, batch):
  """rng='device' on one buffer (one CTA up to 256 strata, tiles of 128 above): with
  the Philox port's uniforms the CPU statement of PRB:142-171 picks the same indices,
  retries included."""
  from oracle import philox_port
  from oracle import sharded_port
  rng = np.random.RandomState(batch)
  ours, ports = _shard_pairs(gpu, 1, rng, cap=3000, attempts=1500, term_p=0.03)
  mem, port = ours[0], ports[0]
  ids = rng.randint(0, 300, size=200).astype(np.int32)
  pr = np.sqrt(np.abs(rng.randn(200)) + 1e-10).astype(np.float32)
  mem.set_priority(ids, pr)
  port.set_priority(ids, pr)
  mem._rng, mem._seed = 'device', 77
  retried = 0
  for call in range(3):
    got = mem.sample_index_batch(batch).cpu().numpy()
    draw_offset = (call + 1) + call  # host offset (1, 2, 3) + device counter (0, 1, 2)
    queries = philox_port.stratified_queries(77, draw_offset, batch)
    retries = philox_port.retry_uniforms(77, 0, draw_offset, batch, 1500)
    (slots, want, used), = sharded_port.sharded_sample([port], queries, [retries])
    assert slots == list(range(batch))
    assert got.tolist() == [int(i) for i in want], call
    retried += used
    gpu.native.check(gpu.native.lib().b2r_check(mem._h, gpu.native.current_stream()))
  assert retried > 0 or batch < 100  # the retry stream was exercised
