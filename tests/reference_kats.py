"""The reference's own known-answer tests for the hot path, restated so that they
can be run against the CPU port (oracle pin) AND the CUDA classes (-m gpu).

Each function cites the reference test it mirrors (paths relative to
/root/reference/tests/dopamine/).  TF-wrapper / session tests are out of scope.
"""
import random

import numpy as np
import pytest

from tests.golden_cases import to_np

OBS = (84, 84)
STACK = 4
BATCH = 32


# ---------------------------------------------------------------------------
# replay_memory/sum_tree_test.py
# ---------------------------------------------------------------------------
def tree_kats(make_tree, nodes_of):
  # :33-36, :38-41
  with pytest.raises(ValueError, match='Sum tree capacity should be positive'):
    make_tree(-1)
  tree = make_tree(100)
  with pytest.raises(ValueError, match='Sum tree values should be nonnegative'):
    tree.set(0, -1)
  # :43-47
  assert len(nodes_of(make_tree(1))) == 1
  assert len(nodes_of(make_tree(2))) == 2
  # :49-52
  t1 = make_tree(1)
  t1.set(0, 1.5)
  assert t1.get(0) == 1.5
  # :54-63
  tree = make_tree(100)
  tree.set(0, 1.0)
  assert tree.get(0) == 1.0
  for level in nodes_of(tree):
    level = np.asarray(level)
    assert level[0] == 1.0
    assert not level[1:].any()
  # :65-66
  assert len(nodes_of(tree)[-1]) >= 100
  # :68-71, :134-137
  empty = make_tree(100)
  with pytest.raises(Exception, match='Cannot sample from an empty sum tree.'):
    empty.sample()
  with pytest.raises(Exception, match='Cannot sample from an empty sum tree.'):
    empty.stratified_sample(5)
  # :73-78
  tree = make_tree(100)
  tree.set(5, 1.0)
  with pytest.raises(ValueError, match=r'query_value must be in \[0, 1\].'):
    tree.sample(query_value=-0.1)
  with pytest.raises(ValueError, match=r'query_value must be in \[0, 1\].'):
    tree.sample(query_value=1.1)
  # :80-84
  assert tree.sample() == 5
  # :86-99
  tree = make_tree(100)
  tree.set(2, 1.0)
  tree.set(3, 3.0)
  for _ in range(20):
    random.seed(1)
    assert tree.sample() == 2
    assert tree.sample(query_value=0.1) == 2
  # :101-132
  tree = make_tree(100)
  random.seed(1)
  r = random.random()
  total = 100 / (1 - r - 0.01)
  tree.set(2, r * total + 0.01)
  tree.set(3, 100)
  for _ in range(20):
    random.seed(1)
    assert tree.sample() == 2
  counts = {2: 0, 3: 0}
  for _ in range(300):
    counts[tree.sample()] += 1
  assert counts[2] < counts[3]
  # :139-146
  tree = make_tree(100)
  for i in range(32):
    tree.set(i, 1)
  assert list(tree.stratified_sample(32)) == list(range(32))
  # :148-154
  tree = make_tree(100)
  tree.set(0, 0)
  assert tree.max_recorded_priority == 1
  for i in range(1, 32):
    tree.set(i, i)
    assert tree.max_recorded_priority == i


# ---------------------------------------------------------------------------
# replay_memory/circular_replay_buffer_test.py (OutOfGraph cases)
# ---------------------------------------------------------------------------
def uniform_kats(make_buffer, invalid_range_fn, make_element):
  # :61-65
  with pytest.raises(AssertionError):
    make_buffer(84, STACK, 5, BATCH)
  # :67-89
  mem = make_buffer((4, 20), STACK, 5, BATCH)
  assert mem._observation_shape == (4, 20)
  assert mem.add_count == 0
  mem = make_buffer(OBS, STACK, 5, BATCH, terminal_dtype=np.int32)
  assert mem._terminal_dtype == np.int32
  # :91-101
  mem = make_buffer(OBS, STACK, 5, BATCH)
  assert mem.cursor() == 0
  mem.add(np.zeros(OBS), 0, 0, 0)
  assert mem.cursor() == STACK
  # :103-137
  extras = [make_element('extra1', [], np.float32),
            make_element('extra2', [2], np.int8)]
  mem = make_buffer(OBS, STACK, 5, BATCH, extra_storage_types=extras)
  mem.add(np.zeros(OBS), 0, 0, 0, 0, [0, 0])
  with pytest.raises(ValueError, match='Add expects'):
    mem.add(np.zeros(OBS), 0, 0, 0)
  assert mem.cursor() == STACK
  mem._check_add_types(np.zeros(OBS), 0, 0, 0, 0, [0, 0])
  with pytest.raises(ValueError, match='Add expects'):
    mem._check_add_types(np.zeros(OBS), 0, 0, 0)
  # :139-166
  with pytest.raises(ValueError, match='There is not enough capacity'):
    make_buffer(OBS, 10, 10, BATCH, update_horizon=1, gamma=1.0)
  with pytest.raises(ValueError, match='There is not enough capacity'):
    make_buffer(OBS, 5, 10, BATCH, update_horizon=10, gamma=1.0)
  make_buffer(OBS, 5, 10, BATCH, update_horizon=5, gamma=1.0)
  # :168-189
  mem = make_buffer(OBS, STACK, 10, BATCH, update_horizon=5, gamma=1.0)
  with pytest.raises(AssertionError,
                     match='end_index must be larger than start_index'):
    mem.get_range([], 2, 1)
  with pytest.raises(AssertionError):
    mem.get_range([], 1, -1)
  with pytest.raises(AssertionError):
    mem.get_range([], 10, 11)
  with pytest.raises(AssertionError, match='Index 1 has not been added.'):
    mem.get_range([], 1, 2)
  # :191-249
  for _ in range(10):
    mem.add(np.full(OBS, 0, dtype=np.uint8), 0, 2.0, 0)
  array = np.arange(10).reshape(10, 1) + np.ones(5)
  assert np.array_equal(mem.get_range(array, 2, 5), array[2:5])
  assert np.array_equal(mem.get_range(array, 8, 12),
                        np.roll(array, 2, axis=0)[:4])
  # :251-268
  mem = make_buffer(OBS, STACK, 10, BATCH, update_horizon=5, gamma=1.0)
  for i in range(50):
    mem.add(np.full(OBS, i, dtype=np.uint8), 0, 2.0, 0)
  for _ in range(10):
    batch = mem.sample_transition_batch()
    assert float(to_np(batch[2])[0]) == 10.0
  # :270-297
  mem = make_buffer(OBS, STACK, 50, BATCH)
  for i in range(11):
    mem.add(np.full(OBS, i, dtype=np.uint8), 0, 0, 0)
  for i in range(3, int(mem.cursor())):
    assert to_np(mem.get_observation_stack(i)).shape == OBS + (4,)
  assert not to_np(mem.get_observation_stack(3)).any()
  stack = to_np(mem.get_observation_stack(6))
  for i in range(4):
    assert np.array_equal(np.full(OBS, i), stack[:, :, i])
  # :299-350 and :352-410
  for with_extras in (False, True):
    kw = dict(extra_storage_types=extras) if with_extras else {}
    tail = (0, [0, 0]) if with_extras else ()
    cap = 10
    mem = make_buffer(OBS, 1, cap, 2, **kw)
    for i in range(50):
      mem.add(np.full(OBS, i, np.uint8), 0, 0, i % 4, *tail)
    for bs in (None, BATCH, None):
      for _ in range(20):
        batch = (mem.sample_transition_batch() if bs is None else
                 mem.sample_transition_batch(bs))
        assert to_np(batch[0]).shape[0] == (2 if bs is None else bs)
    indices = [1, 2, 3, 5, 8]
    want_states = np.array(
        [np.full(OBS + (1,), i, dtype=np.uint8) for i in indices])
    want_next = (want_states + 1) % cap
    want_states += 50 - cap
    want_next += 50 - cap
    want_term = np.array([min((x + 50 - cap) % 4, 1) for x in indices])
    batch = [to_np(x) for x in mem.sample_transition_batch(
        batch_size=len(indices), indices=indices)]
    assert np.array_equal(batch[0], want_states)
    assert not batch[1].any() and not batch[2].any()
    assert np.array_equal(batch[3], want_next)
    assert not batch[4].any() and not batch[5].any()
    assert np.array_equal(batch[6], want_term)
    assert np.array_equal(batch[7], indices)
    if with_extras:
      assert np.array_equal(batch[8], np.zeros(len(indices)))
      assert np.array_equal(batch[9], np.zeros([len(indices), 2]))
      assert batch[8].dtype == np.float32 and batch[9].dtype == np.int8
  # :412-450
  mem = make_buffer(OBS, 1, 10, 2, update_horizon=3, gamma=1.0)
  for i in range(10):
    mem.add(np.full(OBS, i, dtype=np.uint8), i * 2, i, 1 if i == 3 else 0)
  indices = [2, 3, 4]
  batch = [to_np(x) for x in mem.sample_transition_batch(
      batch_size=3, indices=indices)]
  assert np.array_equal(
      batch[0],
      np.array([np.full(OBS + (1,), i, dtype=np.uint8) for i in indices]))
  assert np.array_equal(batch[1], np.array(indices) * 2)
  assert np.array_equal(batch[2], [5, 3, 15])
  assert np.array_equal(batch[6], [1, 1, 0])
  assert np.array_equal(batch[7], indices)
  # :452-474
  assert list(invalid_range_fn(6, 10, 4, 1)) == [5, 6, 7, 8, 9]
  assert list(invalid_range_fn(9, 10, 4, 1)) == [8, 9, 0, 1, 2]
  assert list(invalid_range_fn(0, 10, 4, 1)) == [9, 0, 1, 2, 3]
  assert list(invalid_range_fn(6, 10, 4, 3)) == [3, 4, 5, 6, 7, 8, 9]
  # :476-496
  mem = make_buffer(OBS, STACK, 10, 2)
  mem.add(np.full(OBS, 0, dtype=np.uint8), 0, 0, 0)
  mem.add(np.full(OBS, 0, dtype=np.uint8), 0, 0, 0)
  mem.add(np.full(OBS, 0, dtype=np.uint8), 0, 0, 1)
  want = [0, 0, 0, 1, 1, 0, 0, 0, 0, 0]
  assert [int(bool(mem.is_valid_transition(i))) for i in range(10)] == want


# ---------------------------------------------------------------------------
# replay_memory/prioritized_replay_buffer_test.py (OutOfGraph cases)
# ---------------------------------------------------------------------------
def prioritized_kats(make_buffer):
  cap = 100

  def default_memory():
    return make_buffer(OBS, STACK, cap, BATCH, max_sample_attempts=10)

  def add_blank(mem, action=0, reward=0.0, terminal=0, priority=1.0):
    mem.add(np.zeros(OBS), action, reward, terminal, priority)
    return np.int64((int(mem.cursor()) - 1) % cap)

  # :63-75
  mem = default_memory()
  assert mem.cursor() == 0
  add_blank(mem)
  assert mem.cursor() == STACK and mem.add_count == STACK
  with pytest.raises(ValueError, match='Add expects'):
    mem.add(np.zeros(OBS), 0, 0, 0)
  # :77-81
  mem = default_memory()
  index = add_blank(mem)
  for i in range(index):
    assert mem.sum_tree.get(i) == 0.0
  # :83-90
  with pytest.raises(AssertionError):
    mem.get_priority(index)
  with pytest.raises(AssertionError):
    mem.get_priority(np.array([index]))
  # :92-104
  mem = default_memory()
  indices = np.zeros(7, dtype=np.int32)
  for k in range(7):
    indices[k] = add_blank(mem)
  priorities = np.arange(7)
  mem.set_priority(indices, priorities)
  fetched = to_np(mem.get_priority(np.flip(indices, 0)))
  for i in range(7):
    assert priorities[i] == fetched[7 - 1 - i]
  # :106-111
  mem = default_memory()
  index = add_blank(mem)
  assert to_np(mem.get_priority(np.array([index], dtype=np.int32)))[0] == 1.0
  # :113-125
  mem = default_memory()
  add_blank(mem, terminal=0, priority=0.0)
  for _ in range(3):
    add_blank(mem, terminal=1)
  for _ in range(30):
    batch = mem.sample_transition_batch(batch_size=2)
    assert len(batch) == 9
    assert (to_np(batch[6]) == 1).all()
  # :127-138
  mem = default_memory()
  add_blank(mem)
  with pytest.raises(RuntimeError, match='Max sample attempts: Tried 10 times '
                     'but only sampled 1 valid indices. Batch size is 2'):
    mem.sample_index_batch(2)
  # :140-157
  mem = make_buffer(OBS, STACK, cap, BATCH, max_sample_attempts=cap)
  for _ in range(cap - STACK + 2):
    add_blank(mem)
  assert mem.cursor() == 1
  for s in mem.sample_index_batch(cap):
    assert STACK <= s <= cap - 1


# ---------------------------------------------------------------------------
# agents/rainbow/rainbow_agent_test.py:178-285 (projection vectors)
# ---------------------------------------------------------------------------
PROJECTION_VECTORS = [
    # (supports, weights, target_support, expected)            reference lines
    ([[0, 1, 2, 3, 4]], [[0.1, 0.2, 0.1, 0.3, 0.3]], [0, 1, 2, 3, 4],
     [[0.1, 0.2, 0.1, 0.3, 0.3]]),                                  # :178-188
    ([[0, 1, 2, 3, 4]], [[0.1, 0.2, 0.1, 0.3, 0.3]], [3, 4, 5, 6, 7],
     [[0.7, 0.3, 0.0, 0.0, 0.0]]),                                  # :190-200
    ([[4, 3, 2, 1, 0]], [[0.1, 0.2, 0.1, 0.3, 0.3]], [3, 4, 5, 6, 7],
     [[0.9, 0.1, 0.0, 0.0, 0.0]]),                                  # :202-212
    ([[0, 2, 4, 6, 8], [1, 3, 4, 5, 6]],
     [[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.2, 0.5, 0.1, 0.1]], [4, 5, 6, 7, 8],
     [[0.8, 0.0, 0.1, 0.0, 0.1], [0.8, 0.1, 0.1, 0.0, 0.0]]),       # :214-227
    ([[0, 2, 4, 6, 8], [0, 1, 2, 3, 4], [3, 4, 5, 6, 7]],
     [[0.1, 0.2, 0.3, 0.2, 0.2], [0.1, 0.2, 0.1, 0.3, 0.3],
      [0.1, 0.2, 0.3, 0.2, 0.2]], [3, 4, 5, 6, 7],
     [[0.3, 0.3, 0.0, 0.2, 0.2], [0.7, 0.3, 0.0, 0.0, 0.0],
      [0.1, 0.2, 0.3, 0.2, 0.2]]),                                  # :229-269
    ([[0, 2, 4, 6, 8], [8, 9, 10, 12, 14]],
     [[0.1, 0.2, 0.2, 0.2, 0.3], [0.1, 0.2, 0.4, 0.1, 0.2]], [0, 4, 8, 12, 16],
     [[0.2, 0.4, 0.4, 0.0, 0.0], [0.0, 0.0, 0.45, 0.45, 0.1]]),     # :271-285
]


def projection_kats(project):
  """project(supports, weights, target_support) -> (B, N) array-like, f32."""
  for supports, weights, target, want in PROJECTION_VECTORS:
    got = to_np(project(np.array(supports, np.float32),
                        np.array(weights, np.float32),
                        np.array(target, np.float32)))
    # tf.test.TestCase.assertAllClose defaults: rtol=1e-6, atol=1e-6.
    np.testing.assert_allclose(got, np.array(want, np.float32), rtol=1e-6,
                               atol=1e-6)
