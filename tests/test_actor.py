"""Actor side of the path: the agent's frame stack in
HBM (`_record_observation` / `_reset_state`, dqn_agent.py:444-458, 474-476), the
epsilon schedules and the episode interface of DQNAgent / RainbowAgent.

The agent tests restate tests/dopamine/agents/dqn/dqn_agent_test.py (testBeginEpisode
:103, testStepEval :139, testStepTrain :174, testStepTrainCustom* :282-295,
testLinearlyDecayingEpsilon :297, testBundling :346) and rainbow_agent_test.py
(testStoreTransitionWith*Sampling :493-520) against our classes."""
import random
import zlib

import numpy as np
import pytest

from tests import golden_cases


def test_linearly_decaying_epsilon_reference_schedule():
  """dqn_agent_test.py:297-311."""
  from dopamine_b200.agents.dqn import dqn_agent
  decay_period, warmup_steps, epsilon = 100, 6, 0.1
  for step, want in [(0, 1.0), (16, 0.91), (decay_period + warmup_steps + 1, epsilon)]:
    got = dqn_agent.linearly_decaying_epsilon(decay_period, step, warmup_steps, epsilon)
    assert abs(got - want) < 0.01
  assert dqn_agent.identity_epsilon(1, 2, 3, 0.25) == 0.25


@pytest.fixture(scope='module')
def mods():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.dqn import dqn_agent
  from dopamine_b200.agents.rainbow import agent

  class Mods(object):
    pass

  m = Mods()
  m.torch, m.dqn, m.agent = torch, dqn_agent, agent
  return m


@pytest.mark.gpu
@pytest.mark.parametrize('shape,stack,dtype', [
    ((84, 84), 4, np.uint8), ((84, 84), 1, np.uint8), ((7, 9), 4, np.uint8),
    ((84, 84), 3, np.uint8), ((4,), 1, np.float32), ((5, 3), 2, np.float64),
    ((2, 3, 5), 4, np.int32)])
def test_record_observation_is_roll_and_insert(mods, shape, stack, dtype):
  """The device frame stack against the reference's two numpy lines
  (dqn_agent.py:456-458), over more steps than there are pinned slots, bit for bit;
  float observations land in a uint8 stack by assignment (truncation)."""
  rng = np.random.RandomState(len(shape) * 10 + stack)
  actor = mods.dqn.ActorState(shape, stack, dtype, slots=3)
  want = np.zeros((1,) + shape + (stack,), dtype=dtype)
  want.fill(9)
  actor.tensor.fill_(9)
  actor.reset()
  want.fill(0)
  for step in range(11):
    if np.dtype(dtype) == np.uint8 and step % 3 == 2:
      obs = rng.rand(*shape) * 255.9  # float observation, cast on assignment
    elif np.dtype(dtype).kind == 'f':
      obs = rng.randn(*shape)
    else:
      obs = rng.randint(0, 256, size=shape)
    if step % 2:
      obs = np.reshape(obs, shape + (1,))  # environments without frame stacking
    frame = actor.record(obs)
    observation = np.reshape(obs, shape)
    want = np.roll(want, -1, axis=-1)
    want[0, ..., -1] = observation
    assert frame.dtype == np.dtype(dtype)
    assert frame.tobytes() == want[0, ..., -1].tobytes()
    if step in (0, 4, 10):
      assert actor.numpy().tobytes() == want.tobytes()
  assert actor.numpy().tobytes() == want.tobytes()
  actor.reset()
  assert not actor.numpy().any()
  actor.close()


def _test_agent(mods, **kw):
  """dqn_agent_test.py:53-90: a network that always prefers action 0, no
  exploration, eval mode."""
  torch = mods.torch

  class MockRainbowAgent(mods.agent.RainbowAgent):

    def q_values(self, state):
      q = torch.arange(self.num_actions, 0, -1, dtype=torch.float32, device='cuda')
      return (q + 1.0).repeat(state.shape[0], 1)

  args = dict(num_actions=4, min_replay_history=6, update_period=2,
              target_update_period=4, epsilon_fn=lambda w, x, y, z: 0.0,
              epsilon_eval=0.0, replay_capacity=1000, batch_size=8, update_horizon=3)
  args.update(kw)
  agent = MockRainbowAgent(**args)
  agent.eval_mode = True
  return agent


class _MockMemory(object):
  """test_utils.MockReplayBuffer: records the calls to add()."""

  def __init__(self):
    self.add_count = 0
    self.calls = []

  def add(self, *args):
    self.calls.append(args)


class _HostActorState(object):
  """Test double for ActorState (CPU tests): the reference's two numpy lines."""

  def __init__(self, observation_shape, stack_size, observation_dtype):
    self.observation_shape = tuple(observation_shape)
    self.tensor = np.zeros((1,) + self.observation_shape + (stack_size,),
                           dtype=observation_dtype)

  def reset(self):
    self.tensor.fill(0)

  def record(self, observation):
    frame = np.reshape(observation, self.observation_shape).astype(self.tensor.dtype)
    self.tensor = np.roll(self.tensor, -1, axis=-1)
    self.tensor[0, ..., -1] = frame
    return frame

  def numpy(self):
    return self.tensor


def _host_agent(**kw):
  """ActingLoop over stubs: the episode logic needs no device."""
  from dopamine_b200.agents.dqn import dqn_agent

  class HostAgent(dqn_agent.ActingLoop):

    def __init__(self, num_actions, **acting):
      self.num_actions = num_actions
      self.memory = _MockMemory()
      self.train_ops, self.syncs = [], []
      self._init_acting((5, 3), 4, np.uint8, **acting)

    def _make_actor_state(self, shape, stack, dtype):
      return _HostActorState(shape, stack, dtype)

    def q_values(self, state):
      class _Q(object):  # stands in for a (1, A) tensor whose argmax is action 2

        def argmax(self, dim):
          del dim
          return [2]
      return _Q()

    def train_step(self):
      self.train_ops.append(self.training_steps)

    def sync_target(self):
      self.syncs.append(self.training_steps)

  return HostAgent(**kw)


def test_acting_loop_cadence_on_the_host():
  """dqn_agent.py:341-442 without a device: what is stored when, when the train op and
  the target sync fire (add_count > min_replay_history, multiples of update_period /
  target_update_period), what eval mode suppresses."""
  agent = _host_agent(num_actions=4, min_replay_history=3, update_period=2,
                      target_update_period=4, epsilon_fn=lambda w, x, y, z: 0.0,
                      epsilon_eval=0.0)
  obs = lambda v: np.full((5, 3, 1), v)
  assert agent.begin_episode(obs(1)) == 2
  assert agent.training_steps == 1 and not agent.memory.calls
  for step in range(2, 9):
    agent.memory.add_count = len(agent.memory.calls)  # what add() would have counted
    assert agent.step(0.5, obs(step)) == 2
    stored = agent.memory.calls[-1]
    assert np.array_equal(stored[0], np.full((5, 3), step - 1))
    assert stored[1:] == (2, 0.5, False)
  assert agent.training_steps == 8
  # the cadence saw add_count = 0..6 before steps 1..7 (counters 1..7): add_count > 3
  # from counter 5 on; train op on even counters, sync on multiples of 4
  assert agent.train_ops == [6] and agent.syncs == []
  agent.memory.add_count = 7
  agent.step(0.5, obs(9))  # counter 8: both fire
  assert agent.train_ops == [6, 8] and agent.syncs == [8]
  agent.end_episode(1.0)
  assert agent.memory.calls[-1][1:] == (2, 1.0, True)
  assert np.array_equal(agent.memory.calls[-1][0], np.full((5, 3), 9))
  want = np.zeros((1, 5, 3, 4), np.uint8)
  for k, v in enumerate((6, 7, 8, 9)):
    want[..., k] = v
  assert np.array_equal(agent.state, want)
  # eval mode: nothing stored, nothing trained, the state still rolls
  agent.eval_mode = True
  calls, steps = len(agent.memory.calls), agent.training_steps
  agent.begin_episode(obs(3))
  agent.step(1.0, obs(4))
  agent.end_episode(1.0)
  assert len(agent.memory.calls) == calls and agent.training_steps == steps
  assert agent.state[0, 0, 0].tolist() == [0, 0, 3, 4]


def test_select_action_uses_the_reference_random_stream():
  """dqn_agent.py:411-413: one random.random() per decision, then random.randint."""
  from dopamine_b200.agents.dqn import dqn_agent
  agent = _host_agent(num_actions=6, epsilon_fn=dqn_agent.identity_epsilon,
                      epsilon_train=0.5, epsilon_eval=0.0)
  agent.begin_episode(np.zeros((5, 3)))
  random.seed(3)
  got = [agent._select_action() for _ in range(50)]
  random.seed(3)
  want = []
  for _ in range(50):
    want.append(random.randint(0, 5) if random.random() <= 0.5 else 2)
  assert got == want and len(set(got)) > 2
  agent.eval_mode = True
  assert [agent._select_action() for _ in range(5)] == [2] * 5


class _LoggedMemory(object):
  """Forwards add() to a real replay memory after logging what the agent stored."""

  def __init__(self, memory, log):
    self._memory, self._log = memory, log

  def add(self, obs, action, reward, terminal, *rest):
    self._log['stored'].append((zlib.crc32(np.ascontiguousarray(obs).tobytes()),
                                int(action), float(reward), int(terminal)))
    self._memory.add(obs, action, reward, terminal, *rest)

  def __getattr__(self, name):
    return getattr(self._memory, name)


@pytest.mark.parametrize('name', sorted(golden_cases.ACTOR_CASES))
def test_acting_loop_matches_reference_fixture_on_the_host(name):
  """tests/golden/actor_episodes.npz was written by the reference's own DQNAgent
  methods (oracle/make_golden.py:golden_actor).  ActingLoop over the numpy stand-in
  for the device state and the CPU port of the replay memory must reproduce it:
  actions (greedy and exploratory), every intermediate state, every stored
  transition, the steps on which the train op and the target sync ran, add_count
  and the position of Python's random stream afterwards."""
  from dopamine_b200.agents.dqn import dqn_agent
  from oracle.replay_port import PortReplay

  def make_agent(shape, stack, num_actions, params, log):

    class HostAgent(dqn_agent.ActingLoop):

      def __init__(self):
        self.num_actions = num_actions
        self.memory = _LoggedMemory(PortReplay(shape, stack, 200, 8, update_horizon=1),
                                    log)
        self._init_acting(shape, stack, np.uint8, **params)

      def _make_actor_state(self, s, k, d):
        return _HostActorState(s, k, d)

      def q_values(self, state):
        best = golden_cases.greedy_rule(state, num_actions)

        class _Q(object):

          def argmax(self, dim):
            del dim
            return [best]
        return _Q()

      def train_step(self):
        log['train'].append(self.training_steps)

      def sync_target(self):
        log['sync'].append(self.training_steps)

    return HostAgent()

  golden_cases.check_actor(make_agent, name)


@pytest.mark.gpu
def test_begin_episode(mods):
  """dqn_agent_test.py:103-137 / rainbow_agent_test.py:384-418."""
  agent = _test_agent(mods)
  shape = (84, 84)
  agent.state.fill_(9)
  first = np.ones(shape + (1,))
  assert agent.begin_episode(first) == 0
  want = np.zeros((1,) + shape + (4,), np.uint8)
  want[:, :, :, -1] = 1
  assert np.array_equal(agent.state.cpu().numpy(), want)
  assert np.array_equal(agent._observation, first[:, :, 0])
  assert agent.training_steps == 0  # no training in eval mode
  agent.eval_mode = False
  second = np.ones(shape + (1,)) * 2
  agent.begin_episode(second)
  want[:, :, :, -1] = 2
  assert np.array_equal(agent.state.cpu().numpy(), want)
  assert np.array_equal(agent._observation, second[:, :, 0])
  assert agent.training_steps == 1
  assert agent.updates == 0  # add_count below min_replay_history: no train op


@pytest.mark.gpu
def test_step_eval(mods):
  """dqn_agent_test.py:139-172."""
  agent = _test_agent(mods)
  shape = (84, 84)
  base = np.ones(shape + (1,))
  agent.begin_episode(base)
  agent.memory = _MockMemory()
  want = np.zeros((1,) + shape + (4,), np.uint8)
  num_steps = 10
  for step in range(1, num_steps + 1):
    observation = base * step
    assert agent.step(reward=1, observation=observation) == 0
    stack_pos = step - num_steps - 1
    if stack_pos >= -4:
      want[:, :, :, stack_pos] = step
  assert np.array_equal(agent.state.cpu().numpy(), want)
  assert np.array_equal(agent._last_observation, np.ones(shape) * (num_steps - 1))
  assert np.array_equal(agent._observation, observation[:, :, 0])
  assert agent.training_steps == 0
  assert not agent.memory.calls


@pytest.mark.gpu
@pytest.mark.parametrize('scheme,default_priority', [('uniform', 1.), ('prioritized', None)])
def test_step_train(mods, scheme, default_priority):
  """dqn_agent_test.py:174-223 with the priorities of rainbow_agent_test.py:493-520."""
  agent = _test_agent(mods, replay_scheme=scheme)
  agent.eval_mode = False
  shape = (84, 84)
  base = np.ones(shape + (1,))
  agent.memory = _MockMemory()
  agent.begin_episode(base)
  observation = base
  want = np.zeros((1,) + shape + (4,), np.uint8)
  num_steps = 10
  for step in range(1, num_steps + 1):
    last_observation = observation
    observation = base * step
    assert agent.step(reward=1, observation=observation) == 0
    stack_pos = step - num_steps - 1
    if stack_pos >= -4:
      want[:, :, :, stack_pos] = step
    assert len(agent.memory.calls) == step
    args = agent.memory.calls[-1]
    assert np.array_equal(last_observation[:, :, 0], args[0])
    assert args[1] == 0 and args[2] == 1 and not args[3]
    if default_priority is not None:
      assert args[4] == default_priority
  assert np.array_equal(agent.state.cpu().numpy(), want)
  assert np.array_equal(agent._last_observation, np.full(shape, num_steps - 1))
  assert agent.training_steps == num_steps + 1
  agent.end_episode(reward=1)
  assert len(agent.memory.calls) == num_steps + 1
  args = agent.memory.calls[-1]
  assert np.array_equal(observation[:, :, 0], args[0])
  assert args[1] == 0 and args[2] == 1 and args[3]


@pytest.mark.gpu
def test_train_cadence_exploration_and_bundle(mods, tmp_path):
  """A real agent over a few episodes: the train op runs every update_period steps
  once add_count passes min_replay_history (dqn_agent.py:430-441), the exploratory
  actions are Python's random stream (dqn_agent.py:411-413), and a bundle restores
  state, counters, replay and networks (dqn_agent.py:480-560)."""
  torch = mods.torch
  rng = np.random.RandomState(1)
  kw = dict(num_actions=5, min_replay_history=40, update_period=2,
            target_update_period=8, epsilon_fn=mods.dqn.identity_epsilon,
            epsilon_train=0.3, replay_capacity=500, batch_size=8, update_horizon=3,
            seed=1)
  agent = mods.agent.RainbowAgent(**kw)
  random.seed(11)
  actions = []
  for _ in range(3):
    actions.append(agent.begin_episode(rng.randint(0, 256, size=(84, 84, 1))))
    for _ in range(30):
      actions.append(agent.step(float(rng.randn()), rng.randint(0, 256, size=(84, 84, 1))))
    agent.end_episode(1.0)
  assert all(0 <= a < 5 for a in actions)
  assert len(set(actions)) > 1  # epsilon 0.3: some exploration
  assert agent.training_steps == 3 * 31
  # adds: 30 steps + 1 terminal per episode, + 3 pads at each episode start
  assert int(agent.memory.add_count) == 3 * (31 + 3)
  # train ops: steps with add_count > 40 and training_steps even
  assert 0 < agent.updates <= agent.training_steps // 2
  # the random stream is the reference's: epsilon test, then randint
  greedy = int(agent.q_values(agent.state).argmax(dim=1)[0])
  assert 0 <= greedy < 5
  random.seed(5)
  picked = agent._select_action()
  random.seed(5)
  if random.random() <= 0.3:
    assert picked == random.randint(0, 4)
  else:
    assert picked == greedy
  bundle = agent.bundle_and_checkpoint(str(tmp_path), 7)
  assert set(bundle) == {'state', 'training_steps'}
  assert agent.bundle_and_checkpoint(str(tmp_path / 'missing'), 7) is None
  clone = mods.agent.RainbowAgent(**dict(kw, seed=2))
  assert not clone.unbundle(str(tmp_path), 8, bundle)  # no such iteration
  assert clone.unbundle(str(tmp_path), 7, bundle)
  assert clone.training_steps == agent.training_steps
  assert torch.equal(clone.state, agent.state)
  assert int(clone.memory.add_count) == int(agent.memory.add_count)
  for a, b in zip(agent.online.parameters(), clone.online.parameters()):
    assert torch.equal(a, b)
  x = agent.state
  assert torch.equal(agent.q_values(x), clone.q_values(x))


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(golden_cases.ACTOR_CASES))
def test_agent_matches_reference_fixture(mods, name):
  """The same fixture through the real thing: RainbowAgent with its frame stack in
  HBM (b2r_actor_record / b2r_actor_reset), the CUDA replay memory behind it
  (uniform scheme: priority 1.0) and the reference's episode interface."""
  torch = mods.torch

  def make_agent(shape, stack, num_actions, params, log):

    class Agent(mods.agent.RainbowAgent):

      def q_values(self, state):
        q = torch.zeros(1, num_actions, device='cuda')
        q[0, golden_cases.greedy_rule(state, num_actions)] = 1.0
        return q

      def train_step(self):
        log['train'].append(self.training_steps)

      def sync_target(self):
        log['sync'].append(self.training_steps)

    agent = Agent(num_actions=num_actions, observation_shape=shape, stack_size=stack,
                  update_horizon=1, replay_scheme='uniform', replay_capacity=200,
                  batch_size=8, **params)
    agent.memory = _LoggedMemory(agent.memory, log)
    return agent

  golden_cases.check_actor(make_agent, name)
