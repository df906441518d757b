"""Pins the CPU oracle (oracle/*_port.py, oracle/fast_oracle.c) BEFORE it is
trusted as the checker for the CUDA path:

  1. against fixtures produced by running the unmodified reference
     (tests/golden/*.npz, generator: oracle/make_golden.py);
  2. against the reference's own known-answer tests (tests/reference_kats.py);
  3. when /root/reference is present (authoring container), side by side with the
     imported reference on fresh random histories.
"""
import random

import numpy as np
import pytest

from oracle import c51_port
from oracle import fast
from oracle import refshim
from oracle.replay_port import (Element, PortPrioritizedReplay, PortReplay,
                                cursor_window)
from oracle.sumtree_port import PortSumTree
from tests import golden_cases
from tests import reference_kats


def _nodes(tree):
  return tree.nodes


@pytest.mark.parametrize('cap', golden_cases.TREE_CAPS)
def test_port_tree_matches_reference_fixture(cap):
  golden_cases.check_tree(PortSumTree, cap, _nodes)


@pytest.mark.parametrize('name', golden_cases.UNIFORM_CASES)
def test_port_uniform_matches_reference_fixture(name):
  golden_cases.check_uniform(PortReplay, name)


@pytest.mark.parametrize('name', golden_cases.PER_CASES)
def test_port_prioritized_matches_reference_fixture(name):
  golden_cases.check_prioritized(PortPrioritizedReplay, name, _nodes)


def test_port_passes_reference_sum_tree_tests():
  reference_kats.tree_kats(PortSumTree, _nodes)


def test_port_passes_reference_uniform_buffer_tests():
  reference_kats.uniform_kats(PortReplay, cursor_window, Element)


def test_port_passes_reference_prioritized_buffer_tests():
  reference_kats.prioritized_kats(PortPrioritizedReplay)


def test_port_projection_known_answers():
  reference_kats.projection_kats(c51_port.project_distribution)


def test_discount_vector_bit_patterns():
  # gamma=.99, n=3.
  mem = PortReplay((2, 2), 4, 16, 2, update_horizon=3, gamma=0.99)
  assert mem.discounts.view(np.uint32).tolist() == [
      0x3F800000, 0x3F7D70A4, 0x3F7AE7D5]


def test_c_tree_matches_port():
  rng = np.random.RandomState(3)
  for cap in (1, 2, 37, 1000, 4096):
    port = PortSumTree(cap)
    ctree = fast.FastTree(cap)
    idx = rng.randint(0, cap, size=3000).astype(np.int64)
    val = np.sqrt(np.abs(rng.randn(3000))).astype(np.float32).astype(np.float64)
    for i, v in zip(idx, val):
      port.set(int(i), v)
    assert ctree.set_seq(idx, val) == 0
    assert np.array_equal(ctree.heap.view(np.uint64), port.heap.view(np.uint64))
    assert float(ctree.max_recorded[0]) == float(port.max_recorded_priority)
    mass = rng.rand(500) * port.total()
    assert ctree.descend(mass).tolist() == [port.descend(m) for m in mass]
  # negative value stops the sequence exactly where the reference raises.
  ctree = fast.FastTree(8)
  assert ctree.set_seq([1, 2, 3], [1.0, -1.0, 5.0]) == 2
  assert ctree.level(3).tolist() == [0, 1.0, 0, 0, 0, 0, 0, 0]


def test_c_gather_and_validity_match_port():
  rng = np.random.RandomState(5)
  for (shape, stack, cap, n, adds) in [((6, 8), 4, 50, 3, 137),
                                       ((4, 4), 2, 64, 10, 150),
                                       ((84, 84), 4, 24, 3, 40)]:
    mem = PortReplay(shape, stack, cap, 8, update_horizon=n, gamma=0.97)
    for _ in range(adds):
      mem.add(rng.randint(0, 256, size=shape).astype(np.uint8),
              rng.randint(18), np.float32(rng.randn()), int(rng.rand() < 0.1))
    term_nz = (mem.store['terminal'] != 0).astype(np.uint8)
    for i in range(-1, cap + 1):
      assert fast.is_valid(i, cap, mem.add_count, stack, n, mem.invalid_range,
                           term_nz) == mem.is_valid_transition(i)
    good = [i for i in range(cap) if mem.is_valid_transition(i)]
    want = mem.sample_transition_batch(len(good), good)
    fb = int(np.prod(shape))
    got = fast.gather_u8(cap, fb, stack, n, mem.discounts,
                         mem.store['observation'], mem.store['action'],
                         mem.store['reward'], mem.store['terminal'], good)
    for w, g in zip(want, got):
      assert w.tobytes() == g.tobytes()


needs_reference = pytest.mark.skipif(
    not refshim.reference_available(),
    reason='reference tree only exists in the authoring container')


@needs_reference
def test_port_side_by_side_with_imported_reference():
  _, crb, prb = refshim.load_reference()
  for seed in range(4):
    rng = np.random.RandomState(seed)
    shape, stack, cap, n = (5, 7), 4, 96, 3
    ref = prb.OutOfGraphPrioritizedReplayBuffer(shape, stack, cap, 16,
                                                update_horizon=n, gamma=0.99)
    port = PortPrioritizedReplay(shape, stack, cap, 16, update_horizon=n,
                                 gamma=0.99)
    for step in range(300):
      row = (rng.randint(0, 256, size=shape).astype(np.uint8), rng.randint(18),
             np.float32(np.clip(rng.randn(), -1, 1)), int(rng.rand() < 0.05))
      ref.add(*row, ref.sum_tree.max_recorded_priority)
      port.add(*row, port.sum_tree.max_recorded_priority)
      if step > 40 and step % 4 == 0:
        random.seed(step)
        a = ref.sample_transition_batch()
        random.seed(step)
        b = port.sample_transition_batch()
        for x, y in zip(a, b):
          assert x.tobytes() == y.tobytes()
        pr = np.sqrt(np.abs(rng.randn(16)) + 1e-10).astype(np.float32)
        ref.set_priority(a[7], pr)
        port.set_priority(b[7], pr)
    for lr, lp in zip(ref.sum_tree.nodes, port.sum_tree.nodes):
      assert np.array_equal(lr.view(np.uint64), lp.view(np.uint64))
    assert ref.sum_tree.max_recorded_priority == port.sum_tree.max_recorded_priority
  # uniform buffer, numpy global RNG
  ref = crb.OutOfGraphReplayBuffer((5, 7), 4, 64, 16, update_horizon=1)
  port = PortReplay((5, 7), 4, 64, 16, update_horizon=1)
  rng = np.random.RandomState(9)
  for step in range(200):
    row = (rng.randint(0, 256, size=(5, 7)).astype(np.uint8), rng.randint(18),
           np.float32(rng.randn()), int(rng.rand() < 0.1))
    ref.add(*row)
    port.add(*row)
  np.random.seed(4)
  a = ref.sample_transition_batch()
  np.random.seed(4)
  b = port.sample_transition_batch()
  for x, y in zip(a, b):
    assert x.tobytes() == y.tobytes()


@needs_reference
def test_port_matches_reference_on_the_gairl_usage():
  """GAIRL drives the uniform buffer directly (gairl_agent.py:300-316, 419, 453-455,
  604): the configured batch of 256 from `sample_transition_batch()` and single
  transitions from `sample_transition_batch(batch_size=1)` until a non-terminal one
  turns up.  Same numpy global stream on both sides."""
  _, crb, _ = refshim.load_reference()
  shape, stack, cap = (6, 4), 4, 600
  ref = crb.OutOfGraphReplayBuffer(shape, stack, cap, 256, update_horizon=1)
  port = PortReplay(shape, stack, cap, 256, update_horizon=1)
  rng = np.random.RandomState(21)
  for _ in range(900):  # wraps once
    row = (rng.randint(0, 256, size=shape).astype(np.uint8), rng.randint(4),
           np.float32(rng.randn()), int(rng.rand() < 0.15))
    ref.add(*row)
    port.add(*row)
  for seed in range(3):
    np.random.seed(seed)
    a = ref.sample_transition_batch()
    np.random.seed(seed)
    b = port.sample_transition_batch()
    assert len(a[7]) == 256
    for x, y in zip(a, b):
      assert x.tobytes() == y.tobytes()
  np.random.seed(8)
  picks_ref = [ref.sample_transition_batch(batch_size=1) for _ in range(40)]
  state_after_ref = np.random.get_state()[1].copy()
  np.random.seed(8)
  picks_port = [port.sample_transition_batch(batch_size=1) for _ in range(40)]
  assert np.array_equal(state_after_ref, np.random.get_state()[1])
  for a, b in zip(picks_ref, picks_port):
    for x, y in zip(a, b):
      assert x.tobytes() == y.tobytes()
  assert any(int(t[6][0]) for t in picks_ref) and not all(int(t[6][0]) for t in picks_ref)
