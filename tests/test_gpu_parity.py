"""Parity of the CUDA path (through the C ABI) against the pinned oracle.

Everything here needs a B200: run with `pytest -m gpu`.  Nothing reads
/root/reference; the checker is `oracle/` plus the committed fixtures.
"""
import random

import numpy as np
import pytest

from oracle import c51_port
from oracle import fast
from oracle.replay_port import PortPrioritizedReplay, PortReplay
from oracle.sumtree_port import PortSumTree
from tests import golden_cases
from tests import reference_kats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gpu():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.rainbow import rainbow_agent
  from dopamine_b200.replay_memory import circular_replay_buffer as crb
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
  from dopamine_b200.replay_memory import sum_tree as st

  class Mods(object):
    pass

  m = Mods()
  m.torch, m.crb, m.prb, m.st, m.ra = torch, crb, prb, st, rainbow_agent
  return m


def _nodes(tree):
  return tree.nodes


# ---------------------------------------------------------------- sum tree ----
def test_tree_reference_known_answers(gpu):
  reference_kats.tree_kats(gpu.st.SumTree, _nodes)


@pytest.mark.parametrize('cap', golden_cases.TREE_CAPS)
def test_tree_reference_fixture(gpu, cap):
  golden_cases.check_tree(gpu.st.SumTree, cap, _nodes)


@pytest.mark.parametrize('cap,batch', [(1, 7), (2, 33), (100, 32), (1000, 256),
                                       (4096, 1024), (100000, 4096),
                                       (1 << 20, 4096), (50000, 10000)])
def test_tree_batched_sets_bit_exact(gpu, cap, batch):
  """Long mixed history of batched sets (duplicates, zeros) == sequential oracle."""
  rng = np.random.RandomState(cap % 1000 + batch)
  tree = gpu.st.SumTree(cap)
  want = fast.FastTree(cap)
  for rep in range(6):
    idx = rng.randint(0, cap, size=batch).astype(np.int64)
    if batch > 4:
      idx[rng.randint(0, batch, size=max(2, batch // 8))] = idx[0]  # duplicates
    val = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32).astype(
        np.float64) * (10.0 ** rng.randint(-3, 3))
    val[rng.rand(batch) < 0.05] = 0.0
    tree.set_batch(idx, val)
    assert want.set_seq(idx, val) == 0
  for l, level in enumerate(tree.nodes):
    assert np.array_equal(level.view(np.uint64), want.level(l).view(np.uint64)), (
        'level %d differs' % l)
  assert tree.max_recorded_priority == float(want.max_recorded[0])
  # descent parity on the drifted tree
  q = rng.rand(2000)
  got = tree._descend(q)
  assert got.tolist() == want.descend(q * want.heap[0]).tolist()


@pytest.mark.parametrize('batch', [300, 1024, 4096, 6000])
@pytest.mark.parametrize('odd', [0, 1, 2, 3, 9, 500])
def test_tree_long_chains_verified_scan_bit_exact(gpu, batch, odd):
  """The upper levels add their deltas as a prefix scan that is only kept when it
  satisfies the sequential recurrence bit for bit (tree.cu: chains_by_verified_scan).
  f32 priorities on a tree worth ~1e6 never round (scan accepted at once); `odd`
  entries with full 53-bit mantissas make adds round: 1-2 are repaired by re-rooting
  the scan, more exhaust the rounds and the serial chains run.  Every case must leave
  the oracle's bits on every level."""
  cap = 1 << 20
  rng = np.random.RandomState(batch + odd)
  tree = gpu.st.SumTree(cap)
  want = fast.FastTree(cap)
  fill_idx = np.arange(cap, dtype=np.int64)
  fill = (0.5 + rng.rand(cap)).astype(np.float32).astype(np.float64)
  tree.set_batch(fill_idx, fill)
  assert want.set_seq(fill_idx, fill) == 0
  for rep in range(3):
    idx = rng.randint(0, cap, size=batch).astype(np.int64)
    val = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32).astype(np.float64)
    if odd:
      val[rng.choice(batch, size=min(odd, batch), replace=False)] = rng.rand(
          min(odd, batch)) * np.pi
    tree.set_batch(idx, val)
    assert want.set_seq(idx, val) == 0
    for l, level in enumerate(tree.nodes):
      assert np.array_equal(level.view(np.uint64), want.level(l).view(np.uint64)), (
          'level %d differs after batch %d' % (l, rep))
  assert tree.max_recorded_priority == float(want.max_recorded[0])


@pytest.fixture
def early_tree_sets(monkeypatch):
  """Batched sets through the write-back that groups ahead of its values (tree.cu:
  tree_update_early_kernel, the fused step's kernel)."""
  monkeypatch.setenv('B2R_TREE_SET_PHASE', '3')


def _early_batch(rng, cap, batch, dups, order):
  """Indices of one batch: 'random'; 'sorted' (a stratified sample: rows in leaf order);
  'nearly' (the same with ~1 % of the rows drawn again, as invalid picks are);
  'shuffled-tail' (sorted but for a stretch of 100 rows: too many to move, so the kernel
  sorts)."""
  idx = rng.randint(0, cap, size=batch).astype(np.int64)
  if dups:
    idx[rng.choice(batch, size=dups, replace=False)] = idx[0]
  if order != 'random':
    idx = np.sort(idx)
  if order == 'nearly':
    again = rng.choice(batch, size=max(1, batch // 100), replace=False)
    idx[again] = rng.randint(0, cap, size=len(again))
    if batch > 8:  # two neighbours drawn again, one of them onto an existing leaf
      idx[5] = idx[batch - 2]
      idx[6] = rng.randint(0, cap)
  if order == 'shuffled-tail' and batch > 200:
    rng.shuffle(idx[batch - 150:batch - 50])
  return idx


@pytest.mark.parametrize('order', ['random', 'sorted', 'nearly', 'shuffled-tail'])
@pytest.mark.parametrize('cap,batch,dups', [(2, 33, 0), (1000, 256, 32), (4096, 1024, 128),
                                            (1 << 20, 1024, 2), (1 << 20, 700, 600),
                                            (1 << 20, 64, 0)])
def test_tree_early_grouping_bit_exact(gpu, early_tree_sets, cap, batch, dups, order):
  """Same contract as test_tree_batched_sets_bit_exact; `dups` entries share one leaf
  (600 of 700: more than the list of duplicate leaves takes, so the plain leaf pass
  runs)."""
  rng = np.random.RandomState(cap % 1000 + batch + dups + len(order))
  tree = gpu.st.SumTree(cap)
  want = fast.FastTree(cap)
  if cap > 4096:
    fill_idx = np.arange(cap, dtype=np.int64)
    fill = (0.5 + rng.rand(cap)).astype(np.float32).astype(np.float64)
    for lo in range(0, cap, 4096 * 16):
      tree.set_batch(fill_idx[lo:lo + 4096 * 16], fill[lo:lo + 4096 * 16])
    assert want.set_seq(fill_idx, fill) == 0
  for rep in range(5):
    idx = _early_batch(rng, cap, batch, dups, order)
    val = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32).astype(
        np.float64) * (10.0 ** rng.randint(-3, 3))
    val[rng.rand(batch) < 0.05] = 0.0
    if rep == 3:
      val[rng.choice(batch, size=3, replace=False)] = rng.rand(3) * np.pi  # adds that round
    tree.set_batch(idx, val)
    assert want.set_seq(idx, val) == 0
    for l, level in enumerate(tree.nodes):
      assert np.array_equal(level.view(np.uint64), want.level(l).view(np.uint64)), (
          'level %d differs after batch %d' % (l, rep))
  assert tree.max_recorded_priority == float(want.max_recorded[0])


@pytest.mark.parametrize('order', ['random', 'nearly'])
@pytest.mark.parametrize('bad', ['negative', 'index'])
def test_tree_early_grouping_stops_where_the_reference_raises(gpu, early_tree_sets, bad,
                                                              order):
  """sum_tree.py:178-205 in a loop (prioritized_replay_buffer.py:213-214): the entries
  before the offending one are applied, nothing behind it is."""
  cap, batch, stop = 5000, 600, 417
  rng = np.random.RandomState(11)
  tree = gpu.st.SumTree(cap)
  want = fast.FastTree(cap)
  idx = rng.randint(0, cap, size=batch).astype(np.int64)
  idx[5] = idx[500] = idx[100]
  val = (0.1 + rng.rand(batch)).astype(np.float32).astype(np.float64)
  tree.set_batch(idx, val)
  assert want.set_seq(idx, val) == 0
  idx2 = _early_batch(rng, cap, batch, 4, order)
  val2 = (0.1 + rng.rand(batch)).astype(np.float32).astype(np.float64)
  if bad == 'negative':
    val2[stop] = -1.5
    with pytest.raises(ValueError, match='nonnegative. Got -1.5'):
      tree.set_batch(idx2, val2)
  else:
    idx2[stop] = cap + (1 << 32)  # (its low 32 bits are a valid leaf)
    with pytest.raises((ValueError, IndexError, RuntimeError)):
      tree.set_batch(idx2, val2)
  assert want.set_seq(idx2[:stop], val2[:stop]) == 0
  for l, level in enumerate(tree.nodes):
    assert np.array_equal(level.view(np.uint64), want.level(l).view(np.uint64)), (
        'level %d differs' % l)
  assert tree.max_recorded_priority == float(want.max_recorded[0])
  tree.set_batch(idx, val)  # the latch is cleared; the next batch applies in full
  assert want.set_seq(idx, val) == 0
  assert np.array_equal(tree.nodes[0].view(np.uint64), want.level(0).view(np.uint64))


def test_tree_negative_value_stops_the_batch(gpu):
  tree = gpu.st.SumTree(64)
  with pytest.raises(ValueError, match='nonnegative. Got -2.0'):
    tree.set_batch([1, 2, 3, 4], [1.0, 3.0, -2.0, 5.0])
  leaves = tree.nodes[-1]
  assert leaves[1] == 1.0 and leaves[2] == 3.0 and leaves[3] == 0 and leaves[4] == 0
  assert tree.max_recorded_priority == 3.0
  assert tree._total_priority() == 4.0
  tree.set(9, 2.0)  # the error latch is cleared
  assert tree._total_priority() == 6.0


def test_tree_device_set_matches_host_set(gpu):
  torch = gpu.torch
  rng = np.random.RandomState(4)
  a, b = gpu.st.SumTree(5000), gpu.st.SumTree(5000)
  idx = rng.randint(0, 5000, size=3000).astype(np.int32)
  val = np.abs(rng.randn(3000)).astype(np.float32)
  a.set_batch(idx, val)
  from dopamine_b200 import _native
  d_idx = torch.as_tensor(idx, device='cuda')  # keep alive across the launch
  d_val = torch.as_tensor(val, device='cuda')
  _native.check(_native.lib().b2r_tree_set_device(
      b._h, 3000, d_idx.data_ptr(), d_val.data_ptr(), _native.current_stream()))
  for la, lb in zip(a.nodes, b.nodes):
    assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))


# ---------------------------------------------------------- uniform buffer ----
def test_uniform_reference_known_answers(gpu):
  reference_kats.uniform_kats(gpu.crb.OutOfGraphReplayBuffer,
                              gpu.crb.invalid_range, gpu.crb.ReplayElement)


@pytest.mark.parametrize('name', golden_cases.UNIFORM_CASES)
def test_uniform_reference_fixture(gpu, name):
  golden_cases.check_uniform(gpu.crb.OutOfGraphReplayBuffer, name)


@pytest.mark.parametrize('name', golden_cases.UNIFORM_CASES)
def test_uniform_reference_fixture_device_outputs(gpu, name):
  def make(*a, **k):
    return gpu.crb.OutOfGraphReplayBuffer(*a, output='torch', **k)
  golden_cases.check_uniform(make, name)


def _fill_pair(rng, ours, port, steps, shape, prioritized, term_p=0.03):
  for _ in range(steps):
    row = (rng.randint(0, 256, size=shape).astype(np.uint8), rng.randint(18),
           np.float32(np.clip(rng.randn(), -1, 1)), int(rng.rand() < term_p))
    if prioritized:
      p = port.sum_tree.max_recorded_priority
      ours.add(*row, p)
      port.add(*row, p)
    else:
      ours.add(*row)
      port.add(*row)


def test_uniform_sampling_stream_matches_port(gpu):
  rng = np.random.RandomState(2)
  args = ((8, 8), 4, 500, 32)
  kw = dict(update_horizon=3, gamma=0.99, max_sample_attempts=50)
  ours = gpu.crb.OutOfGraphReplayBuffer(*args, **kw)
  port = PortReplay(*args, **kw)
  for rounds in range(6):
    _fill_pair(rng, ours, port, 173, (8, 8), False, term_p=0.2)
    for bs in (1, 32, 300):
      np.random.seed(rounds * 10 + bs)
      try:
        want = port.sample_transition_batch(bs)
        err = None
      except RuntimeError as e:
        want, err = None, str(e)
      after = np.random.randint(1 << 30)
      np.random.seed(rounds * 10 + bs)
      if err:
        with pytest.raises(RuntimeError) as info:
          ours.sample_transition_batch(bs)
        assert str(info.value) == err
      else:
        got = ours.sample_transition_batch(bs)
        for w, g in zip(want, got):
          assert w.tobytes() == g.tobytes()
      assert np.random.randint(1 << 30) == after


def test_gairl_usage_of_the_uniform_buffer_matches_port(gpu):
  """GAIRL drives the uniform buffer directly (gairl_agent.py:300-316, 419, 453-455,
  604): its configured batch of 256 from `sample_transition_batch()` and single
  transitions from `sample_transition_batch(batch_size=1)` until a non-terminal one
  turns up.  CUDA path against the port (which
  test_oracle_golden.py::test_port_matches_reference_on_the_gairl_usage holds to the
  imported reference) on the same numpy global stream: batches bit-exact, stream
  position equal after the 40 single draws.  Both output conventions."""
  shape, stack, cap = (6, 4), 4, 600
  for output in ('numpy', 'torch'):
    ours = gpu.crb.OutOfGraphReplayBuffer(shape, stack, cap, 256, update_horizon=1,
                                          output=output)
    port = PortReplay(shape, stack, cap, 256, update_horizon=1)
    rng = np.random.RandomState(21)
    for _ in range(900):  # wraps once
      row = (rng.randint(0, 256, size=shape).astype(np.uint8), rng.randint(4),
             np.float32(rng.randn()), int(rng.rand() < 0.15))
      ours.add(*row)
      port.add(*row)
    as_bytes = lambda x: (x.cpu().numpy() if hasattr(x, 'cpu') else x).tobytes()
    for seed in range(3):
      np.random.seed(seed)
      want = port.sample_transition_batch()
      np.random.seed(seed)
      got = ours.sample_transition_batch()
      assert len(got[7]) == 256
      for w, g in zip(want, got):
        assert w.tobytes() == as_bytes(g)
    np.random.seed(8)
    picks_port = [port.sample_transition_batch(batch_size=1) for _ in range(40)]
    state_after_port = np.random.get_state()[1].copy()
    np.random.seed(8)
    picks_ours = [ours.sample_transition_batch(batch_size=1) for _ in range(40)]
    assert np.array_equal(state_after_port, np.random.get_state()[1])
    for w, g in zip(picks_port, picks_ours):
      for x, y in zip(w, g):
        assert x.tobytes() == as_bytes(y)
    terminals = [int(t[6][0]) for t in picks_port]
    assert any(terminals) and not all(terminals)


def test_host_batches_are_fresh_arrays_over_pinned_slabs(gpu):
  """output='numpy' ships the batch in one copy into a page-locked slab and returns views
  over it (b2r_gather_slab).  The reference returns fresh arrays per call (CRB:416-434):
  a batch the caller keeps is never overwritten by later calls, and slabs whose views
  were dropped are reused instead of allocated again."""
  shape, stack, cap, batch = (12, 10), 4, 300, 16
  ours = gpu.crb.OutOfGraphReplayBuffer(shape, stack, cap, batch, update_horizon=2)
  port = PortReplay(shape, stack, cap, batch, update_horizon=2)
  rng = np.random.RandomState(5)
  for _ in range(450):
    row = (rng.randint(0, 256, size=shape).astype(np.uint8), rng.randint(4),
           np.float32(rng.randn()), int(rng.rand() < 0.1))
    ours.add(*row)
    port.add(*row)
  kept, want = [], []
  for seed in range(6):
    np.random.seed(seed)
    want.append(port.sample_transition_batch())
    np.random.seed(seed)
    kept.append(ours.sample_transition_batch())
  for got, ref in zip(kept, want):  # nothing was overwritten by the later calls
    for g, w in zip(got, ref):
      assert g.dtype == w.dtype and g.shape == w.shape
      assert g.tobytes() == w.tobytes()
  slabs = sum(len(v) for v in ours._slab_pool.values())
  assert slabs == 6
  del kept, got, g
  for seed in range(6, 12):  # dropped batches: their slabs come round again
    np.random.seed(seed)
    ref = port.sample_transition_batch()
    np.random.seed(seed)
    got = ours.sample_transition_batch()
    for g, w in zip(got, ref):
      assert g.tobytes() == w.tobytes()
    del got, g
  assert sum(len(v) for v in ours._slab_pool.values()) == 6
  # explicit indices and another batch size go through the same path
  idx = [int(i) for i in want[0][7][:5]]
  a = ours.sample_transition_batch(batch_size=5, indices=idx)
  b = port.sample_transition_batch(batch_size=5, indices=idx)
  for g, w in zip(a, b):
    assert g.tobytes() == w.tobytes()


# ------------------------------------------------------ prioritized buffer ----
def test_prioritized_reference_known_answers(gpu):
  reference_kats.prioritized_kats(gpu.prb.OutOfGraphPrioritizedReplayBuffer)


@pytest.mark.parametrize('name', golden_cases.PER_CASES)
def test_prioritized_reference_fixture(gpu, name):
  golden_cases.check_prioritized(gpu.prb.OutOfGraphPrioritizedReplayBuffer, name,
                                 _nodes)


@pytest.mark.parametrize('attempts', [0, 1, 3, 1000])
def test_prioritized_train_loop_matches_port(gpu, attempts):
  """add / sample / set_priority interleaved like the agent does, incl. retries,
  budget exhaustion (Q10) and duplicate indices in the write-back."""
  rng = np.random.RandomState(attempts)
  args = ((6, 6), 4, 300, 16)
  kw = dict(update_horizon=3, gamma=0.99, max_sample_attempts=attempts)
  ours = gpu.prb.OutOfGraphPrioritizedReplayBuffer(*args, **kw)
  port = PortPrioritizedReplay(*args, **kw)
  errors = 0
  silent_invalid = 0
  for step in range(160):
    _fill_pair(rng, ours, port, 4, (6, 6), True, term_p=0.15)
    if step < 5:
      continue
    random.seed(step)
    try:
      want_idx = port.sample_index_batch(16)
      err = None
    except RuntimeError as e:
      want_idx, err = None, str(e)
    after = random.random()
    random.seed(step)
    if err:
      errors += 1
      with pytest.raises(RuntimeError) as info:
        ours.sample_index_batch(16)
      assert str(info.value) == err
    else:
      got_idx = ours.sample_index_batch(16)
      assert got_idx == [int(i) for i in want_idx]
    assert random.random() == after, 'retry draws consumed differ at %d' % step
    if err:
      continue
    if not all(port.is_valid_transition(i) for i in want_idx):
      silent_invalid += 1  # Q10: the last slot kept an invalid draw
      continue
    want = port.sample_transition_batch(16, want_idx)
    got = ours.sample_transition_batch(16, got_idx)
    for w, g in zip(want, got):
      assert w.tobytes() == g.tobytes()
    pr = np.sqrt(np.abs(rng.randn(16)) + 1e-10).astype(np.float32)
    port.set_priority(want[7], pr)
    ours.set_priority(got[7], pr)
  for lo, lp in zip(ours.sum_tree.nodes, port.sum_tree.nodes):
    assert np.array_equal(lo.view(np.uint64), lp.view(np.uint64))
  assert ours.sum_tree.max_recorded_priority == port.sum_tree.max_recorded_priority
  if attempts in (0, 1):
    assert errors > 0  # the tight budgets must actually exercise the error path


def test_max_recorded_priority_sentinel(gpu):
  prb = gpu.prb
  ours = prb.OutOfGraphPrioritizedReplayBuffer((4, 4), 4, 64, 8)
  port = PortPrioritizedReplay((4, 4), 4, 64, 8)
  rng = np.random.RandomState(0)
  for k in range(40):
    row = (rng.randint(0, 256, size=(4, 4)).astype(np.uint8), 1, 0.5, k % 9 == 8)
    port.add(*row, port.sum_tree.max_recorded_priority)
    ours.add(*row, prb.MAX_RECORDED_PRIORITY)
    if k % 7 == 3:
      ids = np.array([k, max(k - 1, 0)], dtype=np.int32)
      pr = np.array([2.0 + k, 0.25], dtype=np.float32)
      port.set_priority(ids, pr)
      ours.set_priority(ids, pr)
    if k == 20:  # an explicit priority above the running max, same flush
      port.add(*row, 99.0)
      ours.add(*row, 99.0)
  for lo, lp in zip(ours.sum_tree.nodes, port.sum_tree.nodes):
    assert np.array_equal(lo.view(np.uint64), lp.view(np.uint64))
  assert ours.sum_tree.max_recorded_priority == port.sum_tree.max_recorded_priority


# ------------------------------------------------------------- big gathers ----
def _stamp_fill(mem, cap, frame_bytes, rng, term_p=0.001):
  """Fills the HBM stores directly (store_write) with stamped pseudo-random frames."""
  pattern = rng.randint(0, 256, size=(4099, frame_bytes)).astype(np.uint8)
  obs = np.empty((cap, frame_bytes), dtype=np.uint8)
  for start in range(0, cap, 4099):
    n = min(4099, cap - start)
    obs[start:start + n] = pattern[:n]
  obs[:, :8] = np.arange(cap, dtype=np.int64).view(np.uint8).reshape(cap, 8)
  action = rng.randint(0, 18, size=cap).astype(np.int32)
  reward = np.clip(rng.randn(cap), -1, 1).astype(np.float32)
  terminal = (rng.rand(cap) < term_p).astype(np.uint8)
  mem._store['observation'] = obs.reshape((cap,) + mem._observation_shape)
  mem._store['action'] = action
  mem._store['reward'] = reward
  mem._store['terminal'] = terminal
  return obs, action, reward, terminal


@pytest.mark.parametrize('cap,batch,horizon', [(100000, 32, 1), (100000, 4096, 3),
                                               (1000000, 4096, 3)])
def test_gather_full_size_matches_c_oracle(gpu, cap, batch, horizon):
  """BASELINE configs 1-3 at full size: Atari frames, bit-exact against the C
  restatement, wrap-around and terminals inside the trajectory included."""
  rng = np.random.RandomState(cap // 1000 + batch)
  mem = gpu.crb.OutOfGraphReplayBuffer((84, 84), 4, cap, batch,
                                       update_horizon=horizon, gamma=0.99,
                                       output='torch')
  obs, action, reward, terminal = _stamp_fill(mem, cap, 7056, rng, term_p=0.01)
  mem.add_count = cap + 12345  # full and wrapped; cursor = 12345
  idx = rng.randint(0, cap, size=batch).astype(np.int32)
  idx[:8] = [0, 1, 2, 3, cap - 1, cap - 2, cap - 3, cap - 4]  # wrap both ways
  got = mem.sample_transition_batch(batch, indices=idx.tolist())
  want = fast.gather_u8(cap, 7056, 4, horizon, mem._cumulative_discount_vector,
                        obs, action, reward, terminal, idx)
  names = ['state', 'action', 'reward', 'next_state', 'next_action',
           'next_reward', 'terminal', 'indices']
  for nm, w, g in zip(names, want, got):
    g = g.cpu().numpy().reshape(w.shape)
    assert w.tobytes() == g.tobytes(), nm
  assert want[6].any() and not want[6].all()  # both kinds of trajectories seen


def test_validity_mask_full_size(gpu):
  cap = 100000
  rng = np.random.RandomState(1)
  mem = gpu.crb.OutOfGraphReplayBuffer((84, 84), 4, cap, 32, update_horizon=3)
  _, _, _, terminal = _stamp_fill(mem, cap, 7056, rng, term_p=0.01)
  for add_count in (cap + 777, 5000, cap, 2 * cap - 1):
    mem.add_count = add_count
    cursor = add_count % cap
    inv = [(cursor - 3 + i) % cap for i in range(7)]
    mem.invalid_range = inv
    probe = np.concatenate([rng.randint(-5, cap + 5, size=3000),
                            np.array(inv), np.arange(cursor - 10, cursor + 10)])
    from dopamine_b200 import _native
    out = np.zeros(len(probe), dtype=np.uint8)
    p64 = np.ascontiguousarray(probe, dtype=np.int64)
    _native.check(_native.lib().b2r_valid_mask(
        mem._h, len(p64), _native.ptr(p64), _native.ptr(out),
        _native.current_stream()))
    want = [fast.is_valid(int(i), cap, add_count, 4, 3, np.array(inv), terminal)
            for i in probe]
    assert out.astype(bool).tolist() == want


def test_prioritized_full_size_step_matches_oracles(gpu):
  """Config 2/3 shape: capacity 1M tree + sampling + gather + write-back, checked
  against the C tree/gather restatement with the same injected uniforms."""
  cap, batch = 1000000, 1024
  rng = np.random.RandomState(7)
  mem = gpu.prb.OutOfGraphPrioritizedReplayBuffer(
      (84, 84), 4, cap, batch, update_horizon=3, gamma=0.99, output='torch')
  obs, action, reward, terminal = _stamp_fill(mem, cap, 7056, rng)
  add_count = cap + 4242
  mem.add_count = add_count
  inv = np.array([(4242 - 3 + i) % cap for i in range(7)])
  mem.invalid_range = inv
  want_tree = fast.FastTree(cap)
  for lo in range(0, cap, 50000):  # non-uniform priorities over the whole ring
    ids = np.arange(lo, lo + 50000, dtype=np.int32)
    pr = np.sqrt(np.abs(rng.randn(50000)) + 1e-10).astype(np.float32)
    mem.set_priority(ids, pr)
    want_tree.set_seq(ids, pr.astype(np.float64))
  # make sure retries happen: a run of high-priority slots right at the cursor
  hot = np.array(inv[:4], dtype=np.int32)
  mem.set_priority(hot, np.full(4, 3000.0, dtype=np.float32))
  want_tree.set_seq(hot, np.full(4, 3000.0))
  retried = 0
  for step in range(3):
    random.seed(step)
    got = mem.sample_transition_batch()
    ours_next = random.random()
    random.seed(step)
    bounds = np.linspace(0., 1., batch + 1)
    q = np.array([random.uniform(bounds[i], bounds[i + 1]) for i in range(batch)])
    picks = want_tree.descend(q * want_tree.heap[0])
    fixed = []
    for i in picks:
      while not fast.is_valid(int(i), cap, add_count, 4, 3, inv, terminal):
        retried += 1
        i = int(want_tree.descend(np.array([random.random()]) *
                                  want_tree.heap[0])[0])
      fixed.append(int(i))
    assert ours_next == random.random(), 'retry draws consumed differ'
    idx = got[7].cpu().numpy()
    assert idx.tolist() == fixed
    want = fast.gather_u8(cap, 7056, 4, 3, mem._cumulative_discount_vector, obs,
                          action, reward, terminal, idx)
    for w, g in zip(want, got[:8]):
      assert w.tobytes() == g.cpu().numpy().reshape(w.shape).tobytes()
    leaves = want_tree.level(want_tree.depth)
    assert got[8].cpu().numpy().tobytes() == leaves[idx].astype(np.float32).tobytes()
    pr = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32)
    mem.set_priority(got[7], gpu.torch.as_tensor(pr, device='cuda'))
    want_tree.set_seq(idx, pr.astype(np.float64))
  assert retried > 0
  for l, level in enumerate(mem.sum_tree.nodes):
    assert np.array_equal(level.view(np.uint64), want_tree.level(l).view(np.uint64))


@pytest.mark.parametrize('prioritized', [False, True])
def test_add_batch_equals_the_loop_of_adds(gpu, prioritized):
  """b2r_add_batch: n consecutive adds in one native call leave the buffer exactly as
  the loop of add() calls does — stores, zero padding after terminals, add_count,
  invalid_range and (prioritized) every fp64 tree node, for explicit priorities and for
  the max-recorded sentinel — also when n exceeds the staging queue and the ring wraps.
  The loop itself is held to the reference elsewhere (fixtures, known answers)."""
  shape, stack, cap = (9, 7), 4, 257
  rng = np.random.RandomState(17)
  n = 700
  obs = rng.randint(0, 256, size=(n,) + shape).astype(np.uint8)
  act = rng.randint(0, 6, size=n).astype(np.int32)
  rew = rng.randn(n).astype(np.float32)
  term = (rng.rand(n) < 0.07).astype(np.uint8)
  prio = np.abs(rng.randn(n)) + 0.1
  make = (gpu.prb.OutOfGraphPrioritizedReplayBuffer if prioritized
          else gpu.crb.OutOfGraphReplayBuffer)
  one, many = make(shape, stack, cap, 8), make(shape, stack, cap, 8)
  chunks = [(0, 1), (1, 40), (40, 41), (41, 400), (400, 700)]
  for lo, hi in chunks:
    sentinel = prioritized and lo == 40
    for k in range(lo, hi):
      row = (obs[k], int(act[k]), float(rew[k]), int(term[k]))
      if prioritized:
        one.add(*row, gpu.prb.MAX_RECORDED_PRIORITY if sentinel else float(prio[k]))
      else:
        one.add(*row)
    cols = (obs[lo:hi], act[lo:hi], rew[lo:hi], term[lo:hi])
    if prioritized:
      many.add_batch(*cols, gpu.prb.MAX_RECORDED_PRIORITY if sentinel else prio[lo:hi])
    else:
      many.add_batch(*cols)
    assert one.add_count == many.add_count
    assert list(one.invalid_range) == list(many.invalid_range)
  for name in ('observation', 'action', 'reward', 'terminal'):
    assert one._store[name].tobytes() == many._store[name].tobytes(), name
  if prioritized:
    for la, lb in zip(one.sum_tree.nodes, many.sum_tree.nodes):
      assert np.array_equal(la.view(np.uint64), lb.view(np.uint64))
    assert one.sum_tree.max_recorded_priority == many.sum_tree.max_recorded_priority
    with pytest.raises(ValueError, match='nonnegative'):
      many.add_batch(obs[:3], act[:3], rew[:3], term[:3], np.array([1.0, -2.0, 1.0]))
    assert many.add_count >= one.add_count + 1  # the row before the bad one was added
  with pytest.raises(ValueError):
    many.add_batch(obs[:3, :5], act[:3], rew[:3], term[:3], *([prio[:3]] * prioritized))


# ------------------------------------------------------------------- C51 ----
def test_projection_reference_known_answers(gpu):
  reference_kats.projection_kats(gpu.ra.project_distribution)


def test_projection_shape_errors(gpu):
  pd = gpu.ra.project_distribution
  s = np.array([[0, 2, 4, 6, 8], [3, 4, 5, 6, 7]], np.float32)
  w4 = np.array([[0.1, 0.2, 0.3, 0.2]] * 2, np.float32)
  w5 = np.array([[0.1, 0.2, 0.3, 0.2, 0.2]] * 2, np.float32)
  with pytest.raises(ValueError, match='are incompatible'):
    pd(s, w4, np.array([4, 5, 6, 7, 8], np.float32))
  with pytest.raises(ValueError, match='are incompatible'):
    pd(s, w5, np.array([4, 5, 6], np.float32))
  with pytest.raises(ValueError, match='Index out of range'):
    pd(s, w5, np.float32(3))
  with pytest.raises(ValueError, match='out of bounds'):
    pd(s, w5, np.array([[3]], np.float32))
  with pytest.raises(ValueError, match='assertion failed'):
    pd(s, w5, np.array([8, 7, 6, 5, 4], np.float32), validate_args=True)
  with pytest.raises(ValueError, match='assertion failed'):
    pd(s, w5, np.array([3, 4, 6, 7, 8], np.float32), validate_args=True)


@pytest.mark.parametrize('batch,actions,atoms', [(32, 18, 51), (256, 4, 51),
                                                 (1024, 18, 51), (7, 3, 5),
                                                 (4096, 18, 51), (33, 6, 101),
                                                 (1000, 6, 51), (600, 32, 51),
                                                 (300, 40, 21)])
def test_c51_loss_matches_numpy_restatement(gpu, batch, actions, atoms):
  """Projection, loss, new priorities and IS weights within 1e-6 relative of the
  f32 numpy restatement (north_star tolerance; TF's own assertAllClose default)."""
  torch = gpu.torch
  rng = np.random.RandomState(7)
  online = rng.randn(batch, actions, atoms).astype(np.float32)
  target = rng.randn(batch, actions, atoms).astype(np.float32)
  act = rng.randint(0, actions, size=batch).astype(np.int32)
  rew = np.clip(rng.randn(batch), -1, 1).astype(np.float32)
  term = (rng.rand(batch) < 0.2).astype(np.uint8)
  probs = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32)
  want = c51_port.rainbow_update(rew, term, act, probs, online, target,
                                 vmax=10., num_atoms=atoms, gamma=0.99,
                                 update_horizon=3)
  support = gpu.ra.make_support(10., atoms)
  assert support.cpu().numpy().tobytes() == want['support'].tobytes()
  dev = lambda x: torch.as_tensor(x, device='cuda')
  got = gpu.ra.c51_loss(dev(online), dev(target), dev(act), dev(rew), dev(term),
                        dev(probs), support, 0.99 ** 3, want_target=True,
                        want_grad=True)
  # north_star: 1e-6 relative.  (Measured over 16 384 random rows, profiles/r2/README.md:
  # worst loss 4.4e-7, worst priority 2.3e-7 — of the size of the port's own distance
  # from a float64 evaluation, 6.7e-7.)  Projected atoms can be exact zeros on one side
  # and 1e-9 on the other: they keep TF's assertAllClose absolute slack.
  tol = dict(rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(got['target'].cpu().numpy(), want['target'], rtol=1e-6,
                             atol=1e-6)
  np.testing.assert_allclose(got['loss'].cpu().numpy(), want['loss'], **tol)
  np.testing.assert_allclose(got['priorities'].cpu().numpy(),
                             want['priorities'], **tol)
  np.testing.assert_allclose(got['weights'].cpu().numpy(), want['weights'], **tol)
  np.testing.assert_allclose(float(got['mean_weighted_loss']),
                             want['weighted_loss'].mean(), rtol=1e-5)
  # gradient of mean(w * ce) w.r.t. the online logits, against torch autograd
  x = torch.tensor(online, device='cuda', dtype=torch.float64, requires_grad=True)
  t = torch.tensor(want['target'], device='cuda', dtype=torch.float64)
  w = torch.tensor(want['weights'], device='cuda', dtype=torch.float64)
  chosen = x[torch.arange(batch), torch.as_tensor(act, device='cuda').long()]
  ce = -(t * torch.log_softmax(chosen, dim=1)).sum(1)
  (w * ce).mean().backward()
  np.testing.assert_allclose(got['grad_logits'].cpu().numpy(),
                             x.grad.cpu().numpy(), rtol=1e-4, atol=1e-7)
  # uniform scheme: weights are all ones
  uni = gpu.ra.c51_loss(dev(online), dev(target), dev(act), dev(rew), dev(term),
                        None, support, 0.99 ** 3)
  assert (uni['weights'].cpu().numpy() == 1.0).all()
  np.testing.assert_allclose(uni['loss'].cpu().numpy(), want['loss'], **tol)


# --------------------------------------------------------------- sharding ----
@pytest.mark.parametrize('num_shards,global_batch', [(2, 32), (4, 64), (8, 256),
                                                     (4, 2048)])
def test_sharded_sampling_matches_oracle(gpu, num_shards, global_batch):
  """All ranks of a sharded replay emulated on one GPU (one buffer per rank): the
  strata each rank serves, its local indices incl. retries, the counted gather and
  the counted write-back must match the CPU statement of the rule."""
  from dopamine_b200 import _native
  from dopamine_b200.replay_memory import sharded_replay
  from oracle import sharded_port
  import ctypes
  torch = gpu.torch
  rng = np.random.RandomState(num_shards * 1000 + global_batch)
  cap, shape = 400, (8, 8)
  kw = dict(update_horizon=3, gamma=0.99, max_sample_attempts=64)
  ours, ports = [], []
  for g in range(num_shards):
    o = gpu.prb.OutOfGraphPrioritizedReplayBuffer(shape, 4, cap, 8, output='torch', **kw)
    p = PortPrioritizedReplay(shape, 4, cap, 8, **kw)
    _fill_pair(rng, o, p, 300 + 57 * g, shape, True, term_p=0.1)
    ids = rng.randint(0, min(cap, int(p.add_count)), size=200).astype(np.int32)
    pr = (np.sqrt(np.abs(rng.randn(200)) + 1e-10) * (1 + g)).astype(np.float32)
    o.set_priority(ids, pr)
    p.set_priority(ids, pr)
    ours.append(o)
    ports.append(p)
  totals = np.array([p.sum_tree.total() for p in ports], dtype=np.float64)
  d_totals = torch.as_tensor(totals, device='cuda')
  bounds = np.linspace(0., 1., global_batch + 1)
  queries = bounds[:-1] + (bounds[1:] - bounds[:-1]) * rng.rand(global_batch)
  d_queries = torch.as_tensor(queries, device='cuda')
  retries = [rng.rand(max(64, global_batch)) for _ in range(num_shards)]
  want = sharded_port.sharded_sample(ports, queries, retries)
  served = []
  for g in range(num_shards):
    sh = sharded_replay.ShardedPrioritizedReplay(ours[g], rank=g,
                                                 world_size=num_shards)
    # the device total this rank would contribute to the all-gather
    assert float(sh.local_total().cpu()[0]) == totals[g]
    d_retry = torch.as_tensor(retries[g], device='cuda')
    slots, count, batch = sh.sample_transition_batch(
        global_batch, totals=d_totals, queries01=d_queries, retry_u01=d_retry)
    n = int(count.cpu()[0])
    w_slots, w_idx, _ = want[g]
    assert n == len(w_slots)
    assert slots[:n].cpu().numpy().tolist() == w_slots
    idx = batch[7][:n].cpu().numpy()
    assert idx.tolist() == [int(i) for i in w_idx]
    served += w_slots
    if n:
      ref = ports[g].sample_transition_batch(n, [int(i) for i in w_idx])
      for w, got in zip(ref, batch):
        assert w.tobytes() == got[:n].cpu().numpy().tobytes()
      pr = np.sqrt(np.abs(rng.randn(global_batch)) + 1e-10).astype(np.float32)
      sh.set_priority(batch[7], torch.as_tensor(pr, device='cuda'), count)
      ports[g].set_priority(idx, pr[:n])
      for lo, lp in zip(ours[g].sum_tree.nodes, ports[g].sum_tree.nodes):
        assert np.array_equal(lo.view(np.uint64), lp.view(np.uint64))
  assert sorted(served) == list(range(global_batch))


def test_sharded_philox_strata_are_shared_by_all_ranks(gpu):
  """Throughput mode: every rank draws the same strata from Philox(seed, step), so
  the ranks' slot sets partition the global batch without any exchange of draws."""
  from dopamine_b200.replay_memory import sharded_replay
  torch = gpu.torch
  rng = np.random.RandomState(5)
  num_shards, global_batch, cap = 4, 128, 500
  ours = []
  for g in range(num_shards):
    o = gpu.prb.OutOfGraphPrioritizedReplayBuffer((8, 8), 4, cap, 8, output='torch',
                                                  update_horizon=3)
    for _ in range(450):
      o.add(rng.randint(0, 256, size=(8, 8)).astype(np.uint8), 1, 0.5,
            int(rng.rand() < 0.05), float(rng.rand() + 0.1 * g))
    ours.append(o)
  totals = torch.cat([
      sharded_replay.ShardedPrioritizedReplay(o, rank=g, world_size=num_shards)
      .local_total().clone() for g, o in enumerate(ours)])
  served = []
  for g in range(num_shards):
    sh = sharded_replay.ShardedPrioritizedReplay(ours[g], rank=g,
                                                 world_size=num_shards, seed=99)
    slots, idx, count = sh.sample_index_batch(global_batch, totals=totals)
    n = int(count.cpu()[0])
    served += slots[:n].cpu().numpy().tolist()
    valid = [ours[g].is_valid_transition(int(i)) for i in idx[:n].cpu().numpy()]
    assert all(valid)
  assert sorted(served) == list(range(global_batch))
