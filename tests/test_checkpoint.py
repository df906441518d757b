"""Checkpoint format (circular_replay_buffer.py:593-687): files written by the
UNMODIFIED reference (tests/golden/ckpt_ref/, made by oracle/make_golden.py) load
into the HBM-resident buffer and behave as they do in the reference afterwards; what
this build saves is byte-compatible with what the reference saves.
"""
import gzip
import io
import os
import pickle
import random
import shutil

import numpy as np
import pytest

from tests import golden_cases

GOLDEN = golden_cases.GOLDEN
CKPT = os.path.join(GOLDEN, 'ckpt_ref')


def _load_gz(path, array=True):
  with open(path, 'rb') as f:
    with gzip.GzipFile(fileobj=f) as g:
      return np.load(g, allow_pickle=False) if array else g.read()


# ------------------------------------------------------------------ CPU side ----
def test_reference_sum_tree_pickle_loads_without_the_reference():
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
  want = golden_cases.load('checkpoint')
  with open(os.path.join(CKPT, 'sum_tree_ckpt.7.gz'), 'rb') as f:
    with gzip.GzipFile(fileobj=f) as g:
      state = prb._TreeUnpickler(g).load()
  assert isinstance(state, prb.SumTreeState)
  for l, level in enumerate(state.nodes):
    assert np.array_equal(level.view(np.uint64), want['level%d' % l].view(np.uint64))
  assert float(state.max_recorded_priority) == float(want['max_recorded'])


def test_written_sum_tree_pickle_is_the_reference_pickle():
  """dump_reference_sum_tree() emits `SumTree.__new__()` + state for the class path
  the reference pickles; it round-trips through our reader and, when the reference
  is importable (authoring container), unpickles into its real SumTree."""
  from dopamine_b200.replay_memory import prioritized_replay_buffer as prb
  nodes = [np.arange(1 << l, dtype=np.float64) + 0.25 for l in range(4)]
  buf = io.BytesIO()
  prb.dump_reference_sum_tree(nodes, np.float32(3.5), buf)
  raw = buf.getvalue()
  assert b'dopamine.replay_memory.sum_tree\nSumTree\n' in raw
  state = prb._TreeUnpickler(io.BytesIO(raw)).load()
  assert all(np.array_equal(a, b) for a, b in zip(state.nodes, nodes))
  assert state.max_recorded_priority == 3.5
  from oracle import refshim
  if refshim.reference_available():
    st, _, _ = refshim.load_reference()
    tree = pickle.loads(raw)
    assert type(tree) is st.SumTree
    assert all(np.array_equal(a, b) for a, b in zip(tree.nodes, nodes))
    assert tree.max_recorded_priority == 3.5
    tree.set(2, 7.0)  # and it is a working tree
    assert tree.get(2) == 7.0


# ------------------------------------------------------------------ GPU side ----
@pytest.fixture(scope='module')
def prb():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.replay_memory import prioritized_replay_buffer
  return prioritized_replay_buffer


def _make(prb, want):
  stack, cap, batch, n = [int(x) for x in want['cfg']]
  shape = tuple(int(x) for x in want['shape'])
  return prb.OutOfGraphPrioritizedReplayBuffer(
      shape, stack, cap, batch, update_horizon=n, gamma=0.99,
      max_sample_attempts=100)


@pytest.mark.gpu
def test_reference_checkpoint_loads_and_behaves_like_the_reference(prb):
  want = golden_cases.load('checkpoint')
  mem = _make(prb, want)
  mem.load(CKPT, '7')
  assert int(mem.add_count) == int(want['add_count'])
  assert mem.invalid_range.tolist() == want['invalid_range'].tolist()
  for name in ('observation', 'action', 'reward', 'terminal'):
    assert mem._store[name].tobytes() == want['store_' + name].tobytes(), name
  for l, level in enumerate(mem.sum_tree.nodes):
    assert np.array_equal(level.view(np.uint64), want['level%d' % l].view(np.uint64))
  assert mem.sum_tree.max_recorded_priority == float(want['max_recorded'])
  random.seed(3)
  got = mem.sample_transition_batch()
  assert random.random() == float(want['next_u'])
  for e, g in zip(mem.get_transition_elements(), got):
    assert want['out_' + e.name].tobytes() == np.asarray(g).tobytes(), e.name
  for k in range(2):
    mem.add(want['more_obs'][k], want['more_act'][k], want['more_rew'][k],
            want['more_term'][k], 2.5)
  assert int(mem.add_count) == int(want['after_add_count'])
  for l, level in enumerate(mem.sum_tree.nodes):
    assert np.array_equal(level.view(np.uint64),
                          want['after_level%d' % l].view(np.uint64))


@pytest.mark.gpu
def test_saved_files_equal_the_reference_files(prb, tmp_path):
  """load(reference files) -> save(): same file names, same arrays, a sum-tree pickle
  with the same state; old iterations are garbage-collected like the reference's."""
  want = golden_cases.load('checkpoint')
  mem = _make(prb, want)
  mem.load(CKPT, '7')
  out = str(tmp_path)
  mem.save(out, 5)
  mem.save(out, 7)
  names = sorted(os.listdir(CKPT))
  assert sorted(f for f in os.listdir(out) if f.endswith('.7.gz')) == names
  assert all(os.path.exists(os.path.join(out, f.replace('.7.gz', '.5.gz')))
             for f in names)
  for f in names:
    if f.startswith('sum_tree'):
      with open(os.path.join(out, f), 'rb') as fh:
        with gzip.GzipFile(fileobj=fh) as g:
          state = prb._TreeUnpickler(g).load()
      for l, level in enumerate(state.nodes):
        assert np.array_equal(level, want['level%d' % l])
      assert float(state.max_recorded_priority) == float(want['max_recorded'])
    else:
      a, b = _load_gz(os.path.join(out, f)), _load_gz(os.path.join(CKPT, f))
      assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), f
  mem.save(out, 7 + 4)  # CHECKPOINT_DURATION later: iteration 7 is collected
  assert not any(f.endswith('.7.gz') for f in os.listdir(out))
  assert any(f.endswith('.5.gz') for f in os.listdir(out))
  from oracle import refshim
  if refshim.reference_available():  # the reference loads what we wrote
    _, _, ref_prb = refshim.load_reference()
    stack, cap, batch, n = [int(x) for x in want['cfg']]
    ref = ref_prb.OutOfGraphPrioritizedReplayBuffer(
        tuple(int(x) for x in want['shape']), stack, cap, batch, update_horizon=n,
        gamma=0.99, max_sample_attempts=100)
    ref.load(out, '11')
    random.seed(3)
    got = ref.sample_transition_batch()
    for e, g in zip(ref.get_transition_elements(), got):
      assert want['out_' + e.name].tobytes() == g.tobytes(), e.name


@pytest.mark.gpu
def test_reference_checkpoint_known_answers(prb, tmp_path):
  """circular_replay_buffer_test.py:498-647 (testSave, testSaveNonNDArrayAttributes,
  testLoadFromNonexistentDirectory, testPartialLoadFails, testLoad) restated."""
  from dopamine_b200.replay_memory import circular_replay_buffer as crb
  obs_shape, stack, batch = (84, 84), 4, 8
  test_obs = np.ones((5,) + obs_shape, dtype=np.uint8) * 1
  test_action = np.ones(5, dtype=np.int32) * 2
  test_reward = np.ones(5, dtype=np.float32) * 3
  test_terminal = np.ones(5, dtype=np.uint8) * 4
  test_add_count = np.array(7)
  test_invalid = np.array([2, 3, 4, 0, 1])  # length stack + update_horizon
  sub = str(tmp_path / 'ckpt')
  os.makedirs(sub)

  def fresh():
    return crb.OutOfGraphReplayBuffer(obs_shape, stack, 5, batch)

  # testSave + testSaveNonNDArrayAttributes
  mem = fresh()
  mem.observation, mem.action = test_obs, test_action
  mem.reward, mem.terminal = test_reward, test_terminal
  mem.dummy_attribute_1, mem.dummy_attribute_2 = 4753849, 'String data'
  mem.save(sub, 1)
  public = [a for a in mem.__dict__ if not a.startswith('_')]
  assert {'observation', 'dummy_attribute_1', 'dummy_attribute_2'} <= set(public)
  for attr in public + ['add_count', 'invalid_range', '$store$_observation']:
    assert os.path.exists(os.path.join(sub, '{}_ckpt.1.gz'.format(attr))), attr
  mem.save(sub, 5)
  for attr in public:
    assert os.path.exists(os.path.join(sub, '{}_ckpt.5.gz'.format(attr)))
    assert not os.path.exists(os.path.join(sub, '{}_ckpt.1.gz'.format(attr)))
  assert mem.save('/does/not/exist', 1) is None  # silently skipped (CRB:623-624)

  # testLoadFromNonexistentDirectory
  mem = fresh()
  with pytest.raises(FileNotFoundError):
    mem.load('/does/not/exist', '3')
  assert int(mem.add_count) == 0

  # testPartialLoadFails: everything but the reward store is there
  shutil.rmtree(sub)
  os.makedirs(sub)
  arrays = {'$store$_observation': test_obs, '$store$_action': test_action,
            '$store$_terminal': test_terminal, 'add_count': test_add_count,
            'invalid_range': test_invalid}

  def write(name, array):
    with open(os.path.join(sub, '{}_ckpt.3.gz'.format(name)), 'wb') as f:
      with gzip.GzipFile(fileobj=f, mode='wb') as g:
        np.save(g, array, allow_pickle=False)

  for name, array in arrays.items():
    write(name, array)
  mem = fresh()
  with pytest.raises(FileNotFoundError):
    mem.load(sub, '3')
  assert int(mem.add_count) == 0
  assert not mem._store['observation'].any()  # nothing was loaded

  # testLoad
  write('$store$_reward', test_reward)
  mem.load(sub, '3')
  assert np.array_equal(mem._store['observation'], test_obs)
  assert np.array_equal(mem._store['action'], test_action)
  assert np.array_equal(mem._store['reward'], test_reward)
  assert np.array_equal(mem._store['terminal'], test_terminal)
  assert int(mem.add_count) == 7
  assert mem.invalid_range.tolist() == test_invalid.tolist()
