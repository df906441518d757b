"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Runs only in the authoring container (needs
/root/reference); the fixtures it writes are committed, so the GPU box and the
CPU test-suite never need the reference tree.

  python -m oracle.make_golden            # rewrites tests/golden/

Every fixture stores the inputs (so a test can replay the same history through the
port / the CUDA library) and the reference's outputs.
"""
import math
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # pylint: disable=g-import-not-at-top

OUT = os.path.join(ROOT, 'tests', 'golden')


def synth_history(rng, num, obs_shape, terminal_p, num_actions=18):
  obs = rng.randint(0, 256, size=(num,) + obs_shape).astype(np.uint8)
  act = rng.randint(0, num_actions, size=num).astype(np.int32)
  rew = np.clip(rng.randn(num), -1, 1).astype(np.float32)
  term = (rng.rand(num) < terminal_p).astype(np.uint8)
  return obs, act, rew, term


def golden_sum_tree(sum_tree):
  out = {}
  for cap in (1, 2, 100, 1000):
    rng = np.random.RandomState(100 + cap)
    tree = sum_tree.SumTree(cap)
    n = 400
    idx = rng.randint(0, cap, size=n).astype(np.int64)
    # f32-valued priorities (what set_priority receives) + some exact dups/zeros
    val = np.sqrt(np.abs(rng.randn(n)) + 1e-10).astype(np.float32).astype(
        np.float64)
    val[rng.rand(n) < 0.05] = 0.0
    val[7] = 12.5
    for i, v in zip(idx, val):
      tree.set(int(i), float(v))
    queries = np.concatenate([[0.0, 1.0], rng.rand(64)])
    picks = np.array([tree.sample(query_value=float(q)) for q in queries],
                     dtype=np.int64)
    random.seed(cap)
    strat = np.array(tree.stratified_sample(32), dtype=np.int64)
    out['cap%d_idx' % cap] = idx
    out['cap%d_val' % cap] = val
    out['cap%d_queries' % cap] = queries
    out['cap%d_picks' % cap] = picks
    out['cap%d_strat_seed%d' % (cap, cap)] = strat
    out['cap%d_max' % cap] = np.float64(tree.max_recorded_priority)
    for l, level in enumerate(tree.nodes):
      out['cap%d_level%d' % (cap, l)] = level.copy()
  np.savez_compressed(os.path.join(OUT, 'sum_tree.npz'), **out)


def golden_uniform(crb):
  out = {}
  cases = {
      # name: (obs_shape, stack, capacity, horizon, gamma, num_adds, term_p)
      'small_wrap': ((6, 8), 4, 50, 3, 0.99, 137, 0.08),
      'not_full': ((6, 8), 4, 64, 1, 0.99, 40, 0.1),
      'atari_tiny': ((84, 84), 4, 24, 3, 0.99, 40, 0.1),
      'long_horizon': ((4, 4), 2, 64, 10, 0.9, 150, 0.04),
      'stack1': ((5, 3), 1, 20, 2, 1.0, 55, 0.2),
  }
  for name, (shape, stack, cap, n, gamma, adds, tp) in cases.items():
    rng = np.random.RandomState(len(name))
    obs, act, rew, term = synth_history(rng, adds, shape, tp)
    mem = crb.OutOfGraphReplayBuffer(shape, stack, cap, 8, update_horizon=n,
                                     gamma=gamma)
    for k in range(adds):
      mem.add(obs[k], act[k], rew[k], term[k])
    valid = np.array([mem.is_valid_transition(i) for i in range(-2, cap + 2)],
                     dtype=np.uint8)
    good = [i for i in range(cap) if mem.is_valid_transition(i)]
    batch = mem.sample_transition_batch(batch_size=len(good), indices=good)
    np.random.seed(11)
    drawn = np.array(mem.sample_index_batch(16), dtype=np.int64)
    after = np.random.randint(0, 1 << 30)  # pins how many draws were consumed
    p = name + '_'
    out[p + 'cfg'] = np.array([stack, cap, n, adds], dtype=np.int64)
    out[p + 'shape'] = np.array(shape, dtype=np.int64)
    out[p + 'gamma'] = np.float64(gamma)
    out[p + 'obs'], out[p + 'act'] = obs, act
    out[p + 'rew'], out[p + 'term'] = rew, term
    out[p + 'add_count'] = np.int64(mem.add_count)
    out[p + 'invalid_range'] = np.asarray(mem.invalid_range, dtype=np.int64)
    out[p + 'valid_m2_to_cap_p2'] = valid
    out[p + 'good'] = np.array(good, dtype=np.int32)
    for e, arr in zip(mem.get_transition_elements(len(good)), batch):
      out[p + 'out_' + e.name] = arr
    out[p + 'uniform_seed11'] = drawn
    out[p + 'np_next_randint'] = np.int64(after)
  np.savez_compressed(os.path.join(OUT, 'uniform_replay.npz'), **out)


def golden_prioritized(prb):
  out = {}
  cases = {
      'per_wrap': ((6, 8), 4, 100, 3, 0.99, 260, 0.05, 1000),
      'per_not_full': ((6, 8), 4, 128, 3, 0.99, 90, 0.05, 1000),
      'per_tight_budget': ((4, 4), 4, 32, 1, 0.99, 70, 0.45, 2),
  }
  for name, (shape, stack, cap, n, gamma, adds, tp, attempts) in cases.items():
    rng = np.random.RandomState(len(name) * 7)
    obs, act, rew, term = synth_history(rng, adds, shape, tp)
    mem = prb.OutOfGraphPrioritizedReplayBuffer(
        shape, stack, cap, 8, update_horizon=n, gamma=gamma,
        max_sample_attempts=attempts)
    add_prio = np.zeros(adds, dtype=np.float64)
    for k in range(adds):
      add_prio[k] = mem.sum_tree.max_recorded_priority
      mem.add(obs[k], act[k], rew[k], term[k], add_prio[k])
      if k % 17 == 5:  # interleave priority write-backs (with duplicates)
        ids = rng.randint(0, min(cap, int(mem.add_count)), size=6).astype(
            np.int32)
        ids[1] = ids[0]
        pr = np.sqrt(np.abs(rng.randn(6)) + 1e-10).astype(np.float32)
        mem.set_priority(ids, pr)
        out['%s_set%d_ids' % (name, k)] = ids
        out['%s_set%d_pr' % (name, k)] = pr
    p = name + '_'
    results, errors, states = [], [], []
    for rep in range(24):
      random.seed(1000 + rep)
      try:
        results.append(np.array(mem.sample_index_batch(8), dtype=np.int64))
        errors.append('')
      except RuntimeError as e:
        results.append(np.full(8, -1, dtype=np.int64))
        errors.append(str(e))
      states.append(random.random())  # pins how many draws were consumed
    out[p + 'sample_idx'] = np.stack(results)
    out[p + 'sample_err'] = np.array(errors)
    out[p + 'sample_next_u'] = np.array(states)
    good = [i for i in range(cap) if mem.is_valid_transition(i)][:40]
    batch = mem.sample_transition_batch(batch_size=len(good), indices=good)
    out[p + 'cfg'] = np.array([stack, cap, n, adds, attempts], dtype=np.int64)
    out[p + 'shape'] = np.array(shape, dtype=np.int64)
    out[p + 'gamma'] = np.float64(gamma)
    out[p + 'obs'], out[p + 'act'] = obs, act
    out[p + 'rew'], out[p + 'term'] = rew, term
    out[p + 'add_prio'] = add_prio
    out[p + 'add_count'] = np.int64(mem.add_count)
    out[p + 'good'] = np.array(good, dtype=np.int32)
    for e, arr in zip(mem.get_transition_elements(len(good)), batch):
      out[p + 'out_' + e.name] = arr
    out[p + 'max_recorded'] = np.float64(mem.sum_tree.max_recorded_priority)
    for l, level in enumerate(mem.sum_tree.nodes):
      out[p + 'level%d' % l] = level.copy()
  np.savez_compressed(os.path.join(OUT, 'prioritized_replay.npz'), **out)


def golden_checkpoint(prb):
  """A checkpoint WRITTEN BY THE REFERENCE's save() (circular_replay_buffer.py:
  612-653): the gzip files go to tests/golden/ckpt_ref/ verbatim; what the
  reference does after load()-ing them again goes to checkpoint.npz."""
  ckpt_dir = os.path.join(OUT, 'ckpt_ref')
  os.makedirs(ckpt_dir, exist_ok=True)
  for f in os.listdir(ckpt_dir):
    os.remove(os.path.join(ckpt_dir, f))
  rng = np.random.RandomState(77)
  shape, stack, cap, batch, n = (6, 6), 4, 50, 8, 3
  kw = dict(update_horizon=n, gamma=0.99, max_sample_attempts=100)
  mem = prb.OutOfGraphPrioritizedReplayBuffer(shape, stack, cap, batch, **kw)
  obs, act, rew, term = synth_history(rng, 83, shape, 0.08)
  for k in range(83):
    mem.add(obs[k], act[k], rew[k], term[k], mem.sum_tree.max_recorded_priority)
    if k % 9 == 4:
      ids = rng.randint(0, min(cap, int(mem.add_count)), size=5).astype(np.int32)
      mem.set_priority(ids, np.sqrt(np.abs(rng.randn(5)) + 1e-10).astype(np.float32))
  mem.save(ckpt_dir, 7)
  fresh = prb.OutOfGraphPrioritizedReplayBuffer(shape, stack, cap, batch, **kw)
  fresh.load(ckpt_dir, '7')
  out = {'cfg': np.array([stack, cap, batch, n], dtype=np.int64),
         'shape': np.array(shape, dtype=np.int64),
         'add_count': np.int64(fresh.add_count),
         'invalid_range': np.asarray(fresh.invalid_range, dtype=np.int64),
         'max_recorded': np.float64(fresh.sum_tree.max_recorded_priority)}
  for name, array in fresh._store.items():  # pylint: disable=protected-access
    out['store_' + name] = array.copy()
  for l, level in enumerate(fresh.sum_tree.nodes):
    out['level%d' % l] = level.copy()
  random.seed(3)
  sampled = fresh.sample_transition_batch()
  out['next_u'] = np.float64(random.random())
  for e, arr in zip(fresh.get_transition_elements(), sampled):
    out['out_' + e.name] = arr
  # and it keeps working: two more adds, one more priority batch
  more_obs, more_act, more_rew, more_term = synth_history(rng, 2, shape, 0.0)
  for k in range(2):
    fresh.add(more_obs[k], more_act[k], more_rew[k], more_term[k], 2.5)
  out['more_obs'], out['more_act'] = more_obs, more_act
  out['more_rew'], out['more_term'] = more_rew, more_term
  out['after_add_count'] = np.int64(fresh.add_count)
  for l, level in enumerate(fresh.sum_tree.nodes):
    out['after_level%d' % l] = level.copy()
  np.savez_compressed(os.path.join(OUT, 'checkpoint.npz'), **out)


from tests.golden_cases import (ACTOR_CASES, ACTOR_PARAMS, actor_script,  # pylint: disable=g-import-not-at-top
                                greedy_rule)


def golden_actor(dqn_agent, crb):
  """Drives the reference's own DQNAgent methods (dqn_agent.py:341-476), bound to a
  hand-made object whose `_sess.run` answers the three ops they evaluate: the greedy
  action (greedy_rule of the fed state), the train op and the target sync (recorded).
  The replay memory is the reference's OutOfGraphReplayBuffer.  The last episode runs
  in eval mode."""
  import types
  import zlib
  out = {}
  for name, (shape, stack, num_actions) in ACTOR_CASES.items():
    memory = crb.OutOfGraphReplayBuffer(shape, stack, 200, 8, update_horizon=1)
    log = dict(train=[], sync=[], stored=[])

    class Self(object):
      pass

    me = Self()
    me.observation_shape, me.stack_size, me.num_actions = shape, stack, num_actions
    me.state = np.zeros((1,) + shape + (stack,), dtype=np.uint8)
    me.eval_mode, me.training_steps = False, 0
    me.epsilon_fn = dqn_agent.linearly_decaying_epsilon
    for key, value in ACTOR_PARAMS.items():
      setattr(me, key, value)
    me.summary_writer = None
    me._q_argmax, me._train_op, me._sync_qt_ops, me.state_ph = 'q', 'train', 'sync', 'ph'

    def run(op, feed=None, me=me, log=log, num_actions=num_actions):
      if op == 'q':
        return greedy_rule(feed['ph'], num_actions)
      log[op].append(me.training_steps)
      return None

    def add(obs, action, reward, terminal, memory=memory, log=log):
      log['stored'].append((zlib.crc32(np.ascontiguousarray(obs).tobytes()),
                            int(action), float(reward), int(terminal)))
      memory.add(obs, action, reward, terminal)

    me._sess = types.SimpleNamespace(run=run)
    me._replay = types.SimpleNamespace(memory=memory, add=add)
    for method in ('begin_episode', 'step', 'end_episode', '_select_action',
                   '_train_step', '_record_observation', '_reset_state',
                   '_store_transition'):
      setattr(me, method, types.MethodType(getattr(dqn_agent.DQNAgent, method), me))
    random.seed(2024)
    actions, state_crcs = [], []
    episodes = actor_script(name)
    for e, episode in enumerate(episodes):
      me.eval_mode = e == len(episodes) - 1
      actions.append(me.begin_episode(episode[0][1]))
      state_crcs.append(zlib.crc32(me.state.tobytes()))
      for reward, observation in episode[1:]:
        actions.append(me.step(reward, observation))
        state_crcs.append(zlib.crc32(me.state.tobytes()))
      me.end_episode(episode[-1][0])
    p = name + '_'
    out[p + 'actions'] = np.array(actions, np.int64)
    out[p + 'state_crcs'] = np.array(state_crcs, np.uint32)
    out[p + 'final_state'] = me.state.copy()
    out[p + 'stored_crc'] = np.array([x[0] for x in log['stored']], np.uint32)
    out[p + 'stored_action'] = np.array([x[1] for x in log['stored']], np.int64)
    out[p + 'stored_reward'] = np.array([x[2] for x in log['stored']], np.float64)
    out[p + 'stored_terminal'] = np.array([x[3] for x in log['stored']], np.int64)
    out[p + 'train_steps'] = np.array(log['train'], np.int64)
    out[p + 'sync_steps'] = np.array(log['sync'], np.int64)
    out[p + 'training_steps'] = np.int64(me.training_steps)
    out[p + 'add_count'] = np.int64(memory.add_count)
    out[p + 'random_after'] = np.float64(random.random())
  np.savez_compressed(os.path.join(OUT, 'actor_episodes.npz'), **out)


def _bind(obj, cls, names):
  import types
  for name in names:
    setattr(obj, name, types.MethodType(getattr(cls, name), obj))


def golden_losses(dqn_agent, rainbow_agent, iq_agent):
  """tests/golden/losses.npz: the reference's own loss-building methods, executed
  eagerly on numpy arrays through oracle/tfshim.py (see its header for what that does
  and does not pin).  The "networks" are seeded arrays handed out by a stand-in for
  tf.make_template; everything downstream of them is the reference's code."""
  import collections
  import functools
  import types
  from oracle import tfshim
  tf = sys.modules['tensorflow']
  tfshim.install(tf)
  tf.make_template = lambda name, fn, **unused: functools.partial(fn, name)
  T = tfshim.tensor
  out = {}

  def replay(rng, batch, num_actions, with_probs):
    r = types.SimpleNamespace()
    r.batch_size = batch
    r.states, r.next_states = 'states', 'next_states'
    r.rewards = T(np.clip(rng.randn(batch), -1, 1).astype(np.float32))
    r.terminals = T((rng.rand(batch) < 0.25).astype(np.uint8))
    r.actions = T(rng.randint(0, num_actions, size=batch).astype(np.int32))
    r.indices = T(np.arange(batch, dtype=np.int32))
    r.transition = {}
    if with_probs:
      r.transition['sampling_probabilities'] = T(
          np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32))
    r.set_priority_calls = []
    r.tf_set_priority = lambda idx, pr: r.set_priority_calls.append(
        (np.asarray(idx).copy(), np.asarray(pr).copy())) or 'update_priorities'
    return r

  # ---- Rainbow / C51 (RA:200-305 on DQ:237-263, AL:141-143) ----------------------
  for case, (batch, num_actions, num_atoms, scheme, horizon) in {
      'c51_per': (16, 6, 51, 'prioritized', 3),
      'c51_uniform': (9, 4, 51, 'uniform', 1),
      'c51_atoms11': (12, 3, 11, 'prioritized', 3)}.items():
    rng = np.random.RandomState(sum(map(ord, case)))
    me = types.SimpleNamespace()
    me._replay = replay(rng, batch, num_actions, scheme == 'prioritized')
    me._num_atoms, me.num_actions = num_atoms, num_actions
    vmax = 10.
    me._support = tf.linspace(-vmax, vmax, num_atoms)  # rainbow_agent.py:126
    me.cumulative_gamma = math.pow(0.99, horizon)  # dqn_agent.py:175 (a Python float)
    me._replay_scheme = scheme
    me.summary_writer = None
    me.state_ph = 'state_ph'
    me.optimizer = types.SimpleNamespace(minimize=lambda loss: loss)
    logits = {('Online', 'states'): rng.randn(batch, num_actions, num_atoms),
              ('Target', 'next_states'): rng.randn(batch, num_actions, num_atoms),
              ('Online', 'state_ph'): rng.randn(1, num_actions, num_atoms)}
    net_type = collections.namedtuple('c51_network',
                                      ['q_values', 'logits', 'probabilities'])

    def template(name, state, me=me, logits=logits, net_type=net_type):
      lg = T(logits[(name, state)].astype(np.float32))
      probabilities = tf.softmax(lg)                      # atari_lib.py:142
      q_values = tf.reduce_sum(me._support * probabilities, axis=2)  # :143
      return net_type(q_values, lg, probabilities)

    me._network_template = template
    _bind(me, dqn_agent.DQNAgent, ['_build_networks'])
    _bind(me, rainbow_agent.RainbowAgent, ['_build_target_distribution',
                                           '_build_train_op'])
    me._build_networks()
    target = me._build_target_distribution()
    mean_loss, loss = me._build_train_op()
    p = case + '_'
    out[p + 'cfg'] = np.array([batch, num_actions, num_atoms, horizon,
                               int(scheme == 'prioritized')], np.int64)
    out[p + 'online_logits'] = logits[('Online', 'states')].astype(np.float32)
    out[p + 'target_logits'] = logits[('Target', 'next_states')].astype(np.float32)
    out[p + 'rewards'] = np.asarray(me._replay.rewards)
    out[p + 'terminals'] = np.asarray(me._replay.terminals)
    out[p + 'actions'] = np.asarray(me._replay.actions)
    if scheme == 'prioritized':
      out[p + 'probs'] = np.asarray(me._replay.transition['sampling_probabilities'])
      (_, priorities), = me._replay.set_priority_calls
      out[p + 'priorities'] = priorities
    out[p + 'support'] = np.asarray(me._support)
    out[p + 'target'] = np.asarray(target)
    out[p + 'weighted_loss'] = np.asarray(loss)
    out[p + 'mean_loss'] = np.float32(mean_loss)
    out[p + 'q_argmax'] = np.int64(me._q_argmax)

  # ---- DQN (DQ:237-322) -----------------------------------------------------------
  for case, (batch, num_actions, horizon) in {'dqn_a': (16, 6, 1),
                                              'dqn_b': (7, 3, 3)}.items():
    rng = np.random.RandomState(sum(map(ord, case)))
    me = types.SimpleNamespace()
    me._replay = replay(rng, batch, num_actions, False)
    me.num_actions = num_actions
    me.cumulative_gamma = math.pow(0.99, horizon)  # dqn_agent.py:175 (a Python float)
    me.summary_writer = None
    me.state_ph = 'state_ph'
    me.optimizer = types.SimpleNamespace(minimize=lambda loss: loss)
    q = {('Online', 'states'): rng.randn(batch, num_actions) * 2,
         ('Target', 'next_states'): rng.randn(batch, num_actions) * 2,
         ('Online', 'state_ph'): rng.randn(1, num_actions)}
    net_type = collections.namedtuple('DQN_network', ['q_values'])
    me._network_template = lambda name, state, q=q, net_type=net_type: net_type(
        T(q[(name, state)].astype(np.float32)))
    _bind(me, dqn_agent.DQNAgent, ['_build_networks', '_build_target_q_op',
                                   '_build_train_op'])
    me._build_networks()
    losses = []
    tf.losses.huber_loss = (lambda real: lambda *a, **k: losses.append(real(*a, **k))
                            or losses[-1])(tfshim._huber_loss)  # pylint: disable=protected-access
    mean_loss = me._build_train_op()
    tf.losses.huber_loss = tfshim._huber_loss  # pylint: disable=protected-access
    p = case + '_'
    out[p + 'cfg'] = np.array([batch, num_actions, horizon], np.int64)
    out[p + 'online_q'] = q[('Online', 'states')].astype(np.float32)
    out[p + 'target_q'] = q[('Target', 'next_states')].astype(np.float32)
    out[p + 'rewards'] = np.asarray(me._replay.rewards)
    out[p + 'terminals'] = np.asarray(me._replay.terminals)
    out[p + 'actions'] = np.asarray(me._replay.actions)
    out[p + 'target'] = np.asarray(me._build_target_q_op())
    out[p + 'loss'] = np.asarray(losses[0])
    out[p + 'mean_loss'] = np.float32(mean_loss)

  # ---- IQN (IQ:120-321) -----------------------------------------------------------
  for case, (batch, num_actions, n, n_prime, k, kappa, horizon, double_dqn) in {
      'iqn_a': (8, 5, 16, 12, 8, 1.0, 3, False),
      'iqn_kappa': (6, 3, 8, 8, 4, 0.5, 1, True),
      'iqn_paper': (4, 18, 64, 64, 32, 1.0, 3, False)}.items():
    rng = np.random.RandomState(sum(map(ord, case)))
    me = types.SimpleNamespace()
    me._replay = replay(rng, batch, num_actions, False)
    me.num_actions = num_actions
    me.num_tau_samples, me.num_tau_prime_samples = n, n_prime
    me.num_quantile_samples, me.kappa, me.double_dqn = k, kappa, double_dqn
    me.cumulative_gamma = math.pow(0.99, horizon)  # dqn_agent.py:175 (a Python float)
    me.summary_writer = None
    me.state_ph = 'state_ph'
    me.optimizer = types.SimpleNamespace(minimize=lambda loss: loss)
    nets = {}
    net_type = collections.namedtuple('iqn_network', ['quantile_values', 'quantiles'])

    def template(name, state, num_quantiles, nets=nets, rng=rng, batch=batch,
                 num_actions=num_actions, net_type=net_type):
      rows = num_quantiles * (1 if state == 'state_ph' else batch)
      key = (name, state, num_quantiles)
      nets[key] = net_type(T((rng.randn(rows, num_actions) * 1.5).astype(np.float32)),
                           T(rng.rand(rows, 1).astype(np.float32)))
      return nets[key]

    me._network_template = template
    _bind(me, iq_agent.ImplicitQuantileAgent, ['_build_networks',
                                               '_build_target_quantile_values_op',
                                               '_build_train_op'])
    me._build_networks()
    # the per-row loss is what the final tf.reduce_mean (no axis, IQ:315-321) is fed
    row_losses = []
    real_mean = tf.reduce_mean

    def spy_mean(x, *a, **k):
      if not a and not k:
        row_losses.append(np.asarray(x).copy())
      return real_mean(x, *a, **k)

    tf.reduce_mean = spy_mean
    _, mean_loss = me._build_train_op()
    tf.reduce_mean = real_mean
    action_net = 'Online' if double_dqn else 'Target'
    p = case + '_'
    out[p + 'cfg'] = np.array([batch, num_actions, n, n_prime, k, horizon], np.int64)
    out[p + 'kappa'] = np.float64(kappa)
    out[p + 'online_quantile_values'] = np.asarray(
        nets[('Online', 'states', n)].quantile_values)
    out[p + 'quantiles'] = np.asarray(nets[('Online', 'states', n)].quantiles)
    out[p + 'target_quantile_values'] = np.asarray(
        nets[('Target', 'next_states', n_prime)].quantile_values)
    out[p + 'action_quantile_values'] = np.asarray(
        nets[(action_net, 'next_states', k)].quantile_values)
    out[p + 'rewards'] = np.asarray(me._replay.rewards)
    out[p + 'terminals'] = np.asarray(me._replay.terminals)
    out[p + 'actions'] = np.asarray(me._replay.actions)
    out[p + 'next_action'] = np.asarray(me._replay_next_qt_argmax)
    out[p + 'target'] = np.asarray(me._build_target_quantile_values_op())
    out[p + 'mean_loss'] = np.float32(mean_loss)
    out[p + 'loss'] = row_losses[-1].reshape(batch)
  np.savez_compressed(os.path.join(OUT, 'losses.npz'), **out)


def main():
  st, crb, prb = refshim.load_reference()
  os.makedirs(OUT, exist_ok=True)
  golden_sum_tree(st)
  golden_uniform(crb)
  golden_prioritized(prb)
  golden_checkpoint(prb)
  dqn_agent, _ = refshim.load_reference_agents()
  golden_actor(dqn_agent, crb)
  from dopamine.agents.implicit_quantile import implicit_quantile_agent  # pylint: disable=g-import-not-at-top
  golden_losses(dqn_agent, refshim.load_reference_agents()[1], implicit_quantile_agent)
  for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
  main()
