"""Generates tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Runs only in the authoring container (needs
/root/reference); the fixtures it writes are committed, so the GPU box and the
CPU test-suite never need the reference tree.

  python -m oracle.make_golden            # rewrites tests/golden/

Every fixture stores the inputs (so a test can replay the same history through the
port / the CUDA library) and the reference's outputs.
"""
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


def main():
  st, crb, prb = refshim.load_reference()
  os.makedirs(OUT, exist_ok=True)
  golden_sum_tree(st)
  golden_uniform(crb)
  golden_prioritized(prb)
  golden_checkpoint(prb)
  dqn_agent, _ = refshim.load_reference_agents()
  golden_actor(dqn_agent, crb)
  for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
  main()
