"""DQN's and IQN's learners on top of the replay path: networks
against numpy restatements of atari_lib.py, optimizers against TensorFlow's update
rules, one train step of each against the loss oracles."""
import numpy as np
import pytest

from oracle import dqn_port


def test_rmsprop_is_tensorflows_centred_rmsprop():
  """dqn_agent.py:100-105 / dqn.gin:19-25: RMSPropOptimizer(0.00025, decay 0.95, momentum
  0, epsilon 1e-5, centered).  TensorFlow: ms (starting at ONE) and mg as exponential
  averages, theta -= lr g / sqrt(ms - mg^2 + epsilon) — restated in numpy float32."""
  import torch
  from dopamine_b200.agents.dqn import learner
  rng = np.random.RandomState(0)
  shapes = [(7, 5), (11,)]
  w0 = [rng.randn(*sh).astype(np.float32) for sh in shapes]
  params = [torch.nn.Parameter(torch.tensor(w.copy())) for w in w0]
  opt = learner.make_tf_rmsprop(params)
  w = [x.copy() for x in w0]
  ms = [np.ones_like(x) for x in w0]
  mg = [np.zeros_like(x) for x in w0]
  rho, lr, eps = np.float32(0.95), np.float32(0.00025), np.float32(1e-5)
  for _ in range(6):
    grads = [(1e-2 * rng.randn(*sh)).astype(np.float32) for sh in shapes]
    for p, g in zip(params, grads):
      p.grad = torch.tensor(g.copy())
    opt.step()
    for k, g in enumerate(grads):
      ms[k] = rho * ms[k] + (np.float32(1) - rho) * g * g
      mg[k] = rho * mg[k] + (np.float32(1) - rho) * g
      w[k] = w[k] - lr * g / np.sqrt(ms[k] - mg[k] * mg[k] + eps)
    for p, x, x0 in zip(params, w, w0):
      np.testing.assert_allclose(p.detach().numpy() - x0, x - x0, rtol=2e-5,
                                 atol=1e-6 * np.abs(x - x0).max())


@pytest.fixture(scope='module')
def cuda():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  return torch


def _trunk_numpy(net, state):
  """The three SAME-padded convolutions + NHWC flatten (atari_lib.py:95-100) in float64."""
  x = state.astype(np.float64) / 255.0
  for conv, (k, stride) in zip(net.convs, [(8, 4), (4, 2), (3, 1)]):
    w = conv.weight.detach().cpu().numpy().astype(np.float64)
    b = conv.bias.detach().cpu().numpy().astype(np.float64)
    size = x.shape[1]
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    lo, hi = total // 2, total - total // 2
    x = np.pad(x, ((0, 0), (lo, hi), (lo, hi), (0, 0)))
    win = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(1, 2))[:, ::stride, ::stride]
    x = np.maximum(np.einsum('bhwikl,oikl->bhwo', win, w) + b, 0.0)
  return x.reshape(x.shape[0], -1)


def _linear(layer, x, relu):
  w = layer.weight.detach().cpu().numpy().astype(np.float64)
  y = x @ w.T + layer.bias.detach().cpu().numpy().astype(np.float64)
  return np.maximum(y, 0.0) if relu else y


def _no_tf32(torch):
  class _Ctx(object):
    def __enter__(self):
      self.old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
      torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    def __exit__(self, *unused):
      torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.old
  return _Ctx()


@pytest.mark.gpu
def test_nature_dqn_network_matches_the_reference_architecture(cuda):
  """atari_lib.nature_dqn_network (atari_lib.py:85-105) restated in numpy over the
  module's weights; Xavier-uniform bounds of slim's default initialiser."""
  torch = cuda
  from dopamine_b200.agents.dqn import learner
  torch.manual_seed(1)
  net = learner.make_nature_dqn_network(6).cuda()
  limit = (6.0 / (4 * 64 + 32 * 64)) ** 0.5  # first conv: fan_in 4*8*8, fan_out 32*8*8
  w = net.convs[0].weight
  assert 0.9 * limit < float(w.abs().max()) <= limit
  assert float(net.fc2.bias.abs().max()) == 0.0
  with torch.no_grad():
    for p in net.parameters():
      if p.dim() == 1:
        p.uniform_(-0.05, 0.05)
  state = np.random.RandomState(3).randint(0, 256, size=(3, 84, 84, 4)).astype(np.uint8)
  with _no_tf32(torch), torch.no_grad():
    got = net(torch.as_tensor(state, device='cuda')).cpu().numpy()
  want = _linear(net.fc2, _linear(net.fc1, _trunk_numpy(net, state), True), False)
  np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_implicit_quantile_network_matches_the_reference_architecture(cuda):
  """atari_lib.implicit_quantile_network (atari_lib.py:147-199): state features tiled
  sample-major over the quantile samples, times relu(FC(cos(i pi tau))), FC 512, FC A —
  restated in numpy for given taus."""
  torch = cuda
  from dopamine_b200.agents.implicit_quantile import learner
  torch.manual_seed(2)
  net = learner.make_implicit_quantile_network(5, quantile_embedding_dim=16).cuda()
  with torch.no_grad():
    for p in net.parameters():
      if p.dim() == 1:
        p.uniform_(-0.05, 0.05)
  rng = np.random.RandomState(4)
  state = rng.randint(0, 256, size=(2, 84, 84, 4)).astype(np.uint8)
  n = 3
  taus = rng.rand(n * 2, 1).astype(np.float32)
  with _no_tf32(torch), torch.no_grad():
    got, got_taus = net(torch.as_tensor(state, device='cuda'), n,
                        torch.as_tensor(taus, device='cuda'))
    sampled, sampled_taus = net(torch.as_tensor(state, device='cuda'), n)
  assert got_taus.cpu().numpy().tobytes() == taus.tobytes()
  assert tuple(sampled.shape) == (n * 2, 5) and tuple(sampled_taus.shape) == (n * 2, 1)
  assert float(sampled_taus.min()) >= 0.0 and float(sampled_taus.max()) < 1.0
  features = np.tile(_trunk_numpy(net, state), (n, 1))
  multiples = (np.arange(1, 17, dtype=np.float32) * np.float32(np.pi)).astype(np.float64)
  embedding = _linear(net.embed, np.cos(multiples[None, :] * taus.astype(np.float64)), True)
  want = _linear(net.fc2, _linear(net.fc1, features * embedding, True), False)
  np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-4, atol=2e-5)


def _fill(learner, rng, steps, actions):
  for _ in range(steps):
    learner.store_transition(rng.randint(0, 256, size=(84, 84)).astype(np.uint8),
                             int(rng.randint(actions)), float(np.clip(rng.randn(), -1, 1)),
                             int(rng.rand() < 0.02))


@pytest.mark.gpu
def test_dqn_learner_trains_on_the_replay_path(cuda):
  """One update's loss is the Huber loss of THIS batch under the networks' outputs
  (numpy restatement of dqn_agent.py:283-322); the online network moves, the target
  network only at sync; the loss on a fixed replay goes down."""
  torch = cuda
  from dopamine_b200.agents.dqn import learner as dqn_learner
  rng = np.random.RandomState(0)
  learner = dqn_learner.DQNLearner(6, replay_capacity=2000, batch_size=16, seed=3)
  _fill(learner, rng, 600, 6)
  mem = learner.memory
  before = [p.detach().clone() for p in learner.online.parameters()]
  loss = learner.train_step()
  torch.cuda.synchronize()
  batch = mem._output_cache[(16, True)][1]
  state, action, reward, next_state, _, _, terminal = batch[:7]
  with torch.no_grad():
    # the update has moved the online network: evaluate the loss with the weights as
    # they were (`before`)
    probe = dqn_learner.make_nature_dqn_network(6).cuda()
    probe.load_state_dict(learner.target.state_dict())
    target_q = probe(next_state).cpu().numpy()
    for p, old in zip(probe.parameters(), before):
      p.copy_(old)
    online_q = probe(state).cpu().numpy()
  want = dqn_port.dqn_update(reward.cpu().numpy(), terminal.cpu().numpy(),
                             action.cpu().numpy(), online_q, target_q, gamma=0.99,
                             update_horizon=1)
  np.testing.assert_allclose(float(loss), want['mean_loss'], rtol=1e-4)
  assert any(not torch.equal(p, old) for p, old in zip(learner.online.parameters(), before))
  learner.sync_target()
  for a, b in zip(learner.online.parameters(), learner.target.parameters()):
    assert torch.equal(a, b)
  # regression onto fixed targets (no further sync) over a small replay: the loss falls
  small = dqn_learner.DQNLearner(6, replay_capacity=256, batch_size=32, seed=5)
  _fill(small, rng, 200, 6)
  losses = [float(small.train_step()) for _ in range(400)]
  assert all(np.isfinite(losses))
  assert np.mean(losses[-40:]) < 0.8 * np.mean(losses[:40])


@pytest.mark.gpu
def test_iqn_learner_trains_on_the_replay_path(cuda):
  """The IQN learner runs its update through the quantile-Huber kernel: finite losses
  that go down on a fixed replay, q-values of the right shape, target sync."""
  torch = cuda
  from dopamine_b200.agents.implicit_quantile import learner as iqn_learner
  rng = np.random.RandomState(1)
  learner = iqn_learner.IQNLearner(5, replay_capacity=256, batch_size=16, seed=4,
                                   num_tau_samples=8, num_tau_prime_samples=8,
                                   num_quantile_samples=4, learning_rate=3e-4)
  _fill(learner, rng, 200, 5)
  state = torch.as_tensor(rng.randint(0, 256, size=(3, 84, 84, 4)).astype(np.uint8),
                          device='cuda')
  assert tuple(learner.q_values(state).shape) == (3, 5)
  losses = [float(learner.train_step()) for _ in range(400)]
  assert all(np.isfinite(losses))
  assert np.mean(losses[-40:]) < 0.9 * np.mean(losses[:40])
  learner.sync_target()
  for a, b in zip(learner.online.parameters(), learner.target.parameters()):
    assert torch.equal(a, b)
