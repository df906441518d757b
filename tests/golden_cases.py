"""Shared checks that replay the committed reference fixtures (tests/golden/*.npz,
made by oracle/make_golden.py from the unmodified reference) through ANY
implementation exposing the reference API: the CPU port (oracle pin test) and the
CUDA classes (-m gpu parity tests) run exactly the same assertions.
"""
import os
import random

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

UNIFORM_CASES = ['small_wrap', 'not_full', 'atari_tiny', 'long_horizon',
                 'stack1']
PER_CASES = ['per_wrap', 'per_not_full', 'per_tight_budget']
TREE_CAPS = [1, 2, 100, 1000]


def load(name):
  return np.load(os.path.join(GOLDEN, name + '.npz'))


def to_np(x):
  """Accepts numpy arrays or torch tensors (CUDA outputs)."""
  if hasattr(x, 'detach'):
    return x.detach().cpu().numpy()
  return np.asarray(x)


def check_tree(make_tree, cap, nodes_of):
  """make_tree(cap) -> tree with set/get/sample/stratified_sample;
  nodes_of(tree) -> list of per-level fp64 numpy arrays."""
  g = load('sum_tree')
  tree = make_tree(cap)
  for i, v in zip(g['cap%d_idx' % cap], g['cap%d_val' % cap]):
    tree.set(int(i), float(v))
  levels = nodes_of(tree)
  for l, level in enumerate(levels):
    want = g['cap%d_level%d' % (cap, l)]
    assert np.array_equal(np.asarray(level).view(np.uint64),
                          want.view(np.uint64)), 'level %d differs' % l
  assert float(tree.max_recorded_priority) == float(g['cap%d_max' % cap])
  picks = [tree.sample(query_value=float(q)) for q in g['cap%d_queries' % cap]]
  assert picks == g['cap%d_picks' % cap].tolist()
  random.seed(cap)
  strat = tree.stratified_sample(32)
  assert list(strat) == g['cap%d_strat_seed%d' % (cap, cap)].tolist()


def build_uniform(make_buffer, g, name):
  p = name + '_'
  stack, cap, n, adds = [int(x) for x in g[p + 'cfg']]
  shape = tuple(int(x) for x in g[p + 'shape'])
  mem = make_buffer(shape, stack, cap, 8, update_horizon=n,
                    gamma=float(g[p + 'gamma']))
  for k in range(adds):
    mem.add(g[p + 'obs'][k], g[p + 'act'][k], g[p + 'rew'][k], g[p + 'term'][k])
  return mem, cap


def check_uniform(make_buffer, name):
  g = load('uniform_replay')
  p = name + '_'
  mem, cap = build_uniform(make_buffer, g, name)
  assert int(mem.add_count) == int(g[p + 'add_count'])
  assert np.asarray(mem.invalid_range).tolist() == g[p + 'invalid_range'].tolist()
  valid = [int(bool(mem.is_valid_transition(i))) for i in range(-2, cap + 2)]
  assert valid == g[p + 'valid_m2_to_cap_p2'].tolist()
  good = g[p + 'good']
  batch = mem.sample_transition_batch(batch_size=len(good),
                                      indices=good.tolist())
  names = [e.name for e in mem.get_transition_elements(len(good))]
  assert names == ['state', 'action', 'reward', 'next_state', 'next_action',
                   'next_reward', 'terminal', 'indices']
  for nm, got in zip(names, batch):
    want = g[p + 'out_' + nm]
    got = to_np(got)
    assert got.dtype == want.dtype, (nm, got.dtype, want.dtype)
    assert got.shape == want.shape, (nm, got.shape, want.shape)
    assert got.tobytes() == want.tobytes(), 'output %r differs' % nm
  np.random.seed(11)
  drawn = mem.sample_index_batch(16)
  assert [int(x) for x in drawn] == g[p + 'uniform_seed11'].tolist()
  assert int(np.random.randint(0, 1 << 30)) == int(g[p + 'np_next_randint'])


def build_prioritized(make_buffer, g, name):
  p = name + '_'
  stack, cap, n, adds, attempts = [int(x) for x in g[p + 'cfg']]
  shape = tuple(int(x) for x in g[p + 'shape'])
  mem = make_buffer(shape, stack, cap, 8, update_horizon=n,
                    gamma=float(g[p + 'gamma']), max_sample_attempts=attempts)
  for k in range(adds):
    # The fixture recorded max_recorded_priority at each add; check ours agrees.
    assert float(mem.sum_tree.max_recorded_priority) == float(
        g[p + 'add_prio'][k]), k
    mem.add(g[p + 'obs'][k], g[p + 'act'][k], g[p + 'rew'][k],
            g[p + 'term'][k], g[p + 'add_prio'][k])
    key = '%s_set%d_ids' % (name, k)
    if key in g.files:
      mem.set_priority(g[key], g['%s_set%d_pr' % (name, k)])
  return mem, cap


def check_prioritized(make_buffer, name, nodes_of):
  g = load('prioritized_replay')
  p = name + '_'
  mem, _ = build_prioritized(make_buffer, g, name)
  assert int(mem.add_count) == int(g[p + 'add_count'])
  for l, level in enumerate(nodes_of(mem.sum_tree)):
    want = g[p + 'level%d' % l]
    assert np.array_equal(np.asarray(level).view(np.uint64),
                          want.view(np.uint64)), 'level %d differs' % l
  assert float(mem.sum_tree.max_recorded_priority) == float(
      g[p + 'max_recorded'])
  for rep in range(len(g[p + 'sample_idx'])):
    random.seed(1000 + rep)
    err = str(g[p + 'sample_err'][rep])
    if err:
      try:
        mem.sample_index_batch(8)
        raise AssertionError('expected RuntimeError: ' + err)
      except RuntimeError as e:
        assert str(e) == err
    else:
      got = mem.sample_index_batch(8)
      assert [int(x) for x in got] == g[p + 'sample_idx'][rep].tolist(), rep
    assert random.random() == float(g[p + 'sample_next_u'][rep]), (
        'draw consumption differs at rep %d' % rep)
  good = g[p + 'good']
  batch = mem.sample_transition_batch(batch_size=len(good),
                                      indices=good.tolist())
  names = [e.name for e in mem.get_transition_elements(len(good))]
  assert names[-1] == 'sampling_probabilities'
  for nm, got in zip(names, batch):
    want = g[p + 'out_' + nm]
    got = to_np(got)
    assert got.dtype == want.dtype, (nm, got.dtype, want.dtype)
    assert got.tobytes() == want.tobytes(), 'output %r differs' % nm
