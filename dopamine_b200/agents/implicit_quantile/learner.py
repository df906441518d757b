"""The learner half of ImplicitQuantileAgent on top of the B200 replay path.

What the reference builds around its replay memory
(dopamine/agents/implicit_quantile/implicit_quantile_agent.py:120-321,
implicit_quantile.gin): the implicit quantile network (atari_lib.py:147-199: state
features tiled over the quantile samples, times a cosine embedding of the sampled taus),
the target quantile values and the quantile-Huber loss as ONE kernel
(`quantile_huber_loss`, csrc/iqn.cu), TensorFlow's Adam with the reference's numbers, the
target sync.  Uniform replay through the prioritized buffer class, as the reference does
("IQN currently does not support prioritized replay", implicit_quantile.gin:23-24).
"""
import math

from dopamine_b200.agents.dqn import learner as dqn_learner
from dopamine_b200.agents.implicit_quantile import implicit_quantile_agent as iqn
from dopamine_b200.agents.rainbow import agent as conv
from dopamine_b200.replay_memory import prioritized_replay_buffer


def _torch():
  import torch  # pylint: disable=g-import-not-at-top
  return torch


def _variance_scaling_(torch, module):
  """variance_scaling_initializer(1 / sqrt(3), FAN_IN, uniform) (atari_lib.py:160-161)."""
  fan_in = module.weight[0].numel()
  limit = math.sqrt(3.0 * (1.0 / math.sqrt(3.0)) / fan_in)
  torch.nn.init.uniform_(module.weight, -limit, limit)
  torch.nn.init.zeros_(module.bias)


def make_implicit_quantile_network(num_actions, quantile_embedding_dim=64,
                                   observation_shape=(84, 84), stack_size=4):
  """atari_lib.implicit_quantile_network (atari_lib.py:147-199).  forward(state,
  num_quantiles) -> (quantile_values (num_quantiles * B, A), quantiles (num_quantiles * B,
  1)); rows are sample-major (tf.tile of the state features, :176), the layout
  `quantile_huber_loss` takes."""
  torch = _torch()
  nn = torch.nn

  class ImplicitQuantileNetwork(nn.Module):

    def __init__(self):
      super().__init__()
      self.convs, self.pads, features = dqn_learner.conv_trunk(
          torch, observation_shape, stack_size, _variance_scaling_)
      self.embed = nn.Linear(quantile_embedding_dim, features)
      self.fc1 = nn.Linear(features, 512)
      self.fc2 = nn.Linear(512, num_actions)
      for m in (self.embed, self.fc1, self.fc2):
        _variance_scaling_(torch, m)
      # tf.range(1, dim + 1) * pi  (atari_lib.py:184-186), float32
      self.register_buffer(
          'multiples', torch.arange(1, quantile_embedding_dim + 1, dtype=torch.float32) *
          torch.tensor(math.pi, dtype=torch.float32))

    def forward(self, state, num_quantiles, quantiles=None):
      features = dqn_learner.run_trunk(torch, self.convs, self.pads, state)
      batch = features.shape[0]
      tiled = features.repeat(num_quantiles, 1)                      # :176
      if quantiles is None:                                          # :180-181
        quantiles = torch.rand(num_quantiles * batch, 1, device=features.device)
      embedding = torch.relu(self.embed(torch.cos(self.multiples * quantiles)))  # :183-189
      x = torch.relu(self.fc1(tiled * embedding))                    # :191-194
      return self.fc2(x), quantiles

  return ImplicitQuantileNetwork()


class IQNLearner(object):
  """Replay + train op of ImplicitQuantileAgent (implicit_quantile_agent.py:120-321)."""

  def __init__(self, num_actions, observation_shape=(84, 84), stack_size=4, kappa=1.0,
               num_tau_samples=32, num_tau_prime_samples=32, num_quantile_samples=32,
               quantile_embedding_dim=64, double_dqn=False, gamma=0.99, update_horizon=3,
               replay_capacity=1000000, batch_size=32, target_update_period=8000,
               update_period=4, learning_rate=0.00005, adam_epsilon=0.0003125, seed=0,
               memory=None):
    torch = _torch()
    self.num_actions = num_actions
    self.kappa = kappa
    self.num_tau_samples = num_tau_samples
    self.num_tau_prime_samples = num_tau_prime_samples
    self.num_quantile_samples = num_quantile_samples
    self.double_dqn = double_dqn
    self.batch_size = batch_size
    self.update_period = update_period
    self.target_update_period = target_update_period
    self.memory = memory or prioritized_replay_buffer.OutOfGraphPrioritizedReplayBuffer(
        observation_shape, stack_size, replay_capacity, batch_size,
        update_horizon=update_horizon, gamma=gamma, output='torch', rng='device',
        seed=seed, reuse_outputs=True)
    self.cumulative_gamma = iqn.cumulative_gamma(gamma, update_horizon)
    torch.manual_seed(seed)
    make = lambda: make_implicit_quantile_network(
        num_actions, quantile_embedding_dim, observation_shape, stack_size).cuda()
    self.online, self.target = make(), make()
    self.target.load_state_dict(self.online.state_dict())
    for p in self.target.parameters():
      p.requires_grad_(False)
    self.optimizer = conv.make_tf_adam(self.online.parameters(), lr=learning_rate,
                                       epsilon=adam_epsilon)  # implicit_quantile.gin:26-29
    self.training_steps = 0
    self.updates = 0

  def store_transition(self, last_observation, action, reward, is_terminal):
    """rainbow_agent.py:307-337 with the uniform scheme: priority 1."""
    self.memory.add(last_observation, action, reward, is_terminal, 1.0)

  def q_values(self, state):
    """implicit_quantile_agent.py:148-164: mean over num_quantile_samples."""
    torch = _torch()
    with torch.no_grad():
      values, _ = self.online(state, self.num_quantile_samples)
      return values.view(self.num_quantile_samples, -1, self.num_actions).mean(0)

  def train_step(self):
    """implicit_quantile_agent.py:166-321: online net on `state` with N taus, target net
    on `next_state` with N' taus, the action-picking net with K taus (target net; online
    with double_dqn) -> greedy next action, target quantile values and quantile-Huber
    loss (one kernel) -> backward -> Adam.  Returns the mean loss (a CUDA tensor)."""
    torch = _torch()
    batch = self.memory.sample_transition_batch(self.batch_size)
    state, action, reward, next_state, _, _, terminal = batch[:7]
    with torch.no_grad():
      target_values, _ = self.target(next_state, self.num_tau_prime_samples)
      picker = self.online if self.double_dqn else self.target
      action_values, _ = picker(next_state, self.num_quantile_samples)
    online_values, quantiles = self.online(state, self.num_tau_samples)
    loss, _ = iqn.QuantileHuberLoss.apply(
        online_values, quantiles, target_values, action_values, action, reward, terminal,
        self.cumulative_gamma, self.kappa)
    self.optimizer.zero_grad(set_to_none=True)
    loss.backward()
    self.optimizer.step()
    self.updates += 1
    return loss

  def sync_target(self):
    self.target.load_state_dict(self.online.state_dict())

  def step_cadence(self):
    """dqn_agent.py:418-442."""
    loss = None
    if self.training_steps % self.update_period == 0:
      loss = self.train_step()
    if self.training_steps % self.target_update_period == 0:
      self.sync_target()
    self.training_steps += 1
    return loss
