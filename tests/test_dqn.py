"""DQN target / Huber loss kernel (dqn_agent.py:283-322) and the uniform-replay step
of BASELINE config 1 against the numpy restatement and torch autograd."""
import numpy as np
import pytest

from oracle import dqn_port
from oracle import fast


def test_huber_restatement_known_values():
  # |e| <= 1: 0.5 e^2; above: |e| - 0.5 (tf.losses.huber_loss, delta = 1)
  out = dqn_port.dqn_update(
      rewards=[0., 0., 0.], terminals=[1, 1, 0], actions=[0, 1, 0],
      online_q=[[0.5, 9.], [9., 3.], [2., 0.]], target_q=[[7., 7.], [7., 7.], [1., 1.5]],
      gamma=0.5, update_horizon=1)
  np.testing.assert_allclose(out['target'], [0., 0., 0.75])
  np.testing.assert_allclose(out['loss'], [0.125, 2.5, 1.25 - 0.5])


@pytest.fixture(scope='module')
def mods():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.dqn import dqn_agent
  from dopamine_b200.replay_memory import circular_replay_buffer as crb

  class M(object):
    pass

  m = M()
  m.torch, m.dqn, m.crb = torch, dqn_agent, crb
  return m


@pytest.mark.gpu
@pytest.mark.parametrize('batch,actions', [(1, 2), (32, 18), (1000, 6), (4096, 18)])
def test_dqn_loss_matches_restatement(mods, batch, actions):
  torch = mods.torch
  rng = np.random.RandomState(batch + actions)
  online = (3 * rng.randn(batch, actions)).astype(np.float32)
  target = (3 * rng.randn(batch, actions)).astype(np.float32)
  act = rng.randint(0, actions, size=batch).astype(np.int32)
  rew = np.clip(rng.randn(batch), -1, 1).astype(np.float32)
  term = (rng.rand(batch) < 0.2).astype(np.uint8)
  want = dqn_port.dqn_update(rew, term, act, online, target, 0.99, 3)
  dev = lambda x: torch.as_tensor(x, device='cuda')
  got = mods.dqn.dqn_loss(dev(online), dev(target), dev(act), dev(rew), dev(term),
                          0.99 ** 3, want_target=True, want_grad=True)
  for key in ('target', 'loss', 'grad_q'):
    np.testing.assert_allclose(got[key].cpu().numpy(), want[key], rtol=1e-6, atol=1e-7,
                               err_msg=key)
  np.testing.assert_allclose(float(got['mean_loss']), want['mean_loss'], rtol=2e-6)
  # gradient against torch autograd of the same expression in fp64
  x = torch.tensor(online, device='cuda', dtype=torch.float64, requires_grad=True)
  tq = torch.tensor(want['target'], device='cuda', dtype=torch.float64)
  chosen = x.gather(1, dev(act).long()[:, None])[:, 0]
  torch.nn.functional.huber_loss(chosen, tq, delta=1.0).backward()
  np.testing.assert_allclose(got['grad_q'].cpu().numpy(), x.grad.cpu().numpy(),
                             rtol=1e-5, atol=1e-9)
  mean, _ = mods.dqn.DQNLoss.apply(dev(online).requires_grad_(True), dev(target),
                                   dev(act), dev(rew), dev(term), 0.99 ** 3)
  assert abs(float(mean) - float(want['mean_loss'])) <= 2e-6 * abs(float(want['mean_loss']))


@pytest.mark.gpu
def test_config1_uniform_step_matches_oracles(mods):
  """BASELINE config 1: capacity 100k, stack 4, batch 32, uniform device sampling,
  n = 1: batch bit-exact against the C restatement at the sampled indices, DQN loss
  against the numpy restatement."""
  torch = mods.torch
  cap, batch = 100000, 32
  rng = np.random.RandomState(4)
  mem = mods.crb.OutOfGraphReplayBuffer((84, 84), 4, cap, batch, update_horizon=1,
                                        gamma=0.99, output='torch', rng='device', seed=9)
  pattern = rng.randint(0, 256, size=(1031, 7056)).astype(np.uint8)
  obs = np.empty((cap, 7056), dtype=np.uint8)
  for start in range(0, cap, 1031):
    n = min(1031, cap - start)
    obs[start:start + n] = pattern[:n]
  obs[:, :8] = np.arange(cap, dtype=np.int64).view(np.uint8).reshape(cap, 8)
  action = rng.randint(0, 18, size=cap).astype(np.int32)
  reward = np.clip(rng.randn(cap), -1, 1).astype(np.float32)
  terminal = (rng.rand(cap) < 0.01).astype(np.uint8)
  mem._store['observation'] = obs.reshape(cap, 84, 84)
  mem._store['action'], mem._store['reward'] = action, reward
  mem._store['terminal'] = terminal
  add_count = cap + 500  # full and wrapped
  mem.add_count = add_count
  inv = np.array([(500 - 1 + i) % cap for i in range(5)])
  mem.invalid_range = inv
  seen = set()
  for step in range(3):
    got = mem.sample_transition_batch()
    idx = got[7].cpu().numpy()
    assert all(fast.is_valid(int(i), cap, add_count, 4, 1, inv, terminal) for i in idx)
    seen.add(tuple(idx[:6].tolist()))
    want = fast.gather_u8(cap, 7056, 4, 1, mem._cumulative_discount_vector, obs,
                          action, reward, terminal, idx)
    for w, g in zip(want, got[:8]):
      assert w.tobytes() == g.cpu().numpy().reshape(w.shape).tobytes()
    online = rng.randn(batch, 18).astype(np.float32)
    target = rng.randn(batch, 18).astype(np.float32)
    out = mods.dqn.dqn_loss(torch.as_tensor(online, device='cuda'),
                            torch.as_tensor(target, device='cuda'), got[1], got[2],
                            got[6], 0.99)
    ref = dqn_port.dqn_update(want[2], want[6], want[1], online, target, 0.99, 1)
    np.testing.assert_allclose(out['loss'].cpu().numpy(), ref['loss'], rtol=1e-6,
                               atol=1e-7)
  assert len(seen) == 3
