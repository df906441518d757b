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


# ----------------------------------------------------------------- actor side ----
# tests/golden/actor_episodes.npz: oracle/make_golden.py:golden_actor ran the
# reference's own DQNAgent methods (dqn_agent.py:341-476) over these scripts.
ACTOR_CASES = {
    # name: (observation_shape, stack_size, num_actions)
    'atari': ((84, 84), 4, 6),
    'small': ((6, 5), 4, 4),
    'stack1': ((3, 4), 1, 3),
}

ACTOR_PARAMS = dict(min_replay_history=10, update_period=2, target_update_period=8,
                    epsilon_train=0.1, epsilon_eval=0.05, epsilon_decay_period=40)


def actor_script(name):
  """The episode script both sides replay (regenerated from the seed, not stored):
  per episode a list of (reward, observation); the last entry closes the episode."""
  shape, _, _ = ACTOR_CASES[name]
  rng = np.random.RandomState(sum(map(ord, name)))
  episodes = []
  for length in (9, 17, 12, 8):
    episodes.append([(float(np.float32(np.clip(rng.randn(), -1, 1))),
                      rng.randint(0, 256, size=shape + (1,)).astype(np.uint8))
                     for _ in range(length)])
  return episodes


def greedy_rule(state, num_actions):
  """What the stand-in network prefers: a function of the WHOLE frame stack, so a
  wrong roll or insert changes the actions."""
  return int(to_np(state).astype(np.int64).sum() % num_actions)


def check_actor(make_agent, name):
  """make_agent(shape, stack, num_actions, params, log) -> an agent with the
  reference's episode interface whose greedy action is greedy_rule(state), whose
  train op / target sync append `training_steps` to log['train'] / log['sync'] and
  whose replay adds append (crc32(obs), action, reward, terminal) to log['stored']
  before reaching a real replay memory (`agent.memory`)."""
  import zlib
  g = load('actor_episodes')
  p = name + '_'
  shape, stack, num_actions = ACTOR_CASES[name]
  log = dict(train=[], sync=[], stored=[])
  agent = make_agent(shape, stack, num_actions, dict(ACTOR_PARAMS), log)
  random.seed(2024)
  actions, state_crcs = [], []
  episodes = actor_script(name)
  for e, episode in enumerate(episodes):
    agent.eval_mode = e == len(episodes) - 1
    actions.append(agent.begin_episode(episode[0][1]))
    state_crcs.append(zlib.crc32(to_np(agent.state).tobytes()))
    for reward, observation in episode[1:]:
      actions.append(agent.step(reward, observation))
      state_crcs.append(zlib.crc32(to_np(agent.state).tobytes()))
    agent.end_episode(episode[-1][0])
  assert actions == g[p + 'actions'].tolist()
  assert state_crcs == g[p + 'state_crcs'].tolist()
  assert to_np(agent.state).tobytes() == g[p + 'final_state'].tobytes()
  assert [x[0] for x in log['stored']] == g[p + 'stored_crc'].tolist()
  assert [x[1] for x in log['stored']] == g[p + 'stored_action'].tolist()
  assert [x[2] for x in log['stored']] == g[p + 'stored_reward'].tolist()
  assert [x[3] for x in log['stored']] == g[p + 'stored_terminal'].tolist()
  assert log['train'] == g[p + 'train_steps'].tolist()
  assert log['sync'] == g[p + 'sync_steps'].tolist()
  assert agent.training_steps == int(g[p + 'training_steps'])
  assert int(agent.memory.add_count) == int(g[p + 'add_count'])
  assert random.random() == float(g[p + 'random_after'])  # same draws consumed
