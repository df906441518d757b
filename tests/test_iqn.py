"""IQN's quantile-Huber loss.  The numpy port follows
implicit_quantile_agent.py:166-315 op by op; the reference has no numeric test of this
loss and TensorFlow is absent, so the port is cross-checked against an independent
float64 closed form here (parity unpinned by the reference: see oracle/iqn_port.py)."""
import numpy as np
import pytest

from oracle import iqn_port


def _case(rng, batch, actions, n, n_prime, k, terminal_p=0.2, scale=1.0):
  return dict(
      rewards=np.clip(rng.randn(batch), -1, 1).astype(np.float32),
      terminals=(rng.rand(batch) < terminal_p).astype(np.uint8),
      actions=rng.randint(0, actions, size=batch).astype(np.int32),
      online_quantile_values=(rng.randn(n * batch, actions) * scale).astype(np.float32),
      quantiles=rng.rand(n * batch, 1).astype(np.float32),
      target_quantile_values=(rng.randn(n_prime * batch, actions) * scale).astype(np.float32),
      action_quantile_values=(rng.randn(k * batch, actions) * scale).astype(np.float32),
      num_tau_samples=n, num_tau_prime_samples=n_prime, num_quantile_samples=k)


@pytest.mark.parametrize('batch,actions,n,n_prime,k,kappa', [
    (5, 3, 4, 6, 2, 1.0), (8, 4, 16, 8, 4, 0.5), (3, 2, 7, 5, 3, 2.0)])
def test_port_matches_closed_form(batch, actions, n, n_prime, k, kappa):
  rng = np.random.RandomState(batch * 7 + n)
  case = _case(rng, batch, actions, n, n_prime, k, scale=1.5)
  got = iqn_port.iqn_update(kappa=kappa, gamma=0.99, update_horizon=3, **case)
  # the greedy action is the first maximum of the sample mean
  q = case['action_quantile_values'].astype(np.float64).reshape(k, batch, actions).mean(0)
  assert got['next_action'].tolist() == np.argmax(q, axis=1).tolist()
  want = iqn_port.closed_form_f64(
      case['rewards'], case['terminals'], case['actions'],
      case['online_quantile_values'], case['quantiles'],
      case['target_quantile_values'], got['next_action'], n, n_prime, kappa=kappa,
      gamma=0.99, update_horizon=3)
  np.testing.assert_allclose(got['loss'], want, rtol=5e-6)
  np.testing.assert_allclose(got['mean_loss'], want.mean(), rtol=5e-6)
  # gradient by central differences of the closed form on a few entries
  eps = 1e-4
  for _ in range(6):
    b = rng.randint(batch)
    t = rng.randint(n)
    row, col = t * batch + b, case['actions'][b]
    hi = case['online_quantile_values'].astype(np.float64)
    lo = hi.copy()
    hi[row, col] += eps
    lo[row, col] -= eps
    f = lambda x: iqn_port.closed_form_f64(  # pylint: disable=g-long-lambda
        case['rewards'], case['terminals'], case['actions'], x, case['quantiles'],
        case['target_quantile_values'], got['next_action'], n, n_prime, kappa=kappa,
        gamma=0.99, update_horizon=3).mean()
    fd = (f(hi) - f(lo)) / (2 * eps)
    assert abs(got['grad'][row, col] - fd) <= 1e-4 * max(1.0, abs(fd)) + 1e-6
  # nothing but the chosen action's column carries gradient
  mask = np.ones_like(got['grad'], dtype=bool)
  mask[np.arange(n * batch), np.tile(case['actions'], n)] = False
  assert not got['grad'][mask].any()


def test_terminal_rows_ignore_the_target_network():
  rng = np.random.RandomState(3)
  case = _case(rng, 6, 3, 8, 8, 4, terminal_p=1.0)
  a = iqn_port.iqn_update(**case)
  case['target_quantile_values'] = case['target_quantile_values'] * 0 + 5
  b = iqn_port.iqn_update(**case)
  assert a['loss'].tobytes() == b['loss'].tobytes()
  assert np.array_equal(a['target'], np.tile(case['rewards'][:, None], [1, 8]))


@pytest.fixture(scope='module')
def iq():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.implicit_quantile import implicit_quantile_agent
  return implicit_quantile_agent


@pytest.mark.gpu
@pytest.mark.parametrize('batch,actions,n,n_prime,k,kappa', [
    (32, 18, 64, 64, 32, 1.0), (7, 3, 5, 9, 4, 1.0), (256, 6, 64, 64, 32, 0.5),
    (33, 40, 32, 8, 8, 2.0), (1, 2, 1, 1, 1, 1.0), (64, 4, 200, 130, 3, 1.0)])
def test_iqn_loss_matches_numpy_restatement(iq, batch, actions, n, n_prime, k, kappa):
  """Loss and mean within 2e-6 relative of the f32 port (the kernel forms every term
  in f32 as the reference does and sums in f64; the port sums in f32 pairwise), greedy
  next actions exact, gradient against torch autograd in float64."""
  import torch
  rng = np.random.RandomState(batch + n)
  case = _case(rng, batch, actions, n, n_prime, k, scale=1.5)
  want = iqn_port.iqn_update(kappa=kappa, gamma=0.99, update_horizon=3, **case)
  dev = lambda x: torch.as_tensor(x, device='cuda')
  got = iq.quantile_huber_loss(
      dev(case['online_quantile_values']), dev(case['quantiles']),
      dev(case['target_quantile_values']), dev(case['action_quantile_values']),
      dev(case['actions']), dev(case['rewards']), dev(case['terminals']),
      0.99 ** 3, kappa, want_grad=True)
  assert got['next_action'].cpu().numpy().tolist() == want['next_action'].tolist()
  np.testing.assert_allclose(got['loss'].cpu().numpy(), want['loss'], rtol=2e-6,
                             atol=1e-7)
  np.testing.assert_allclose(float(got['mean_loss']), want['mean_loss'], rtol=2e-6)
  # autograd of the reference's expression in float64
  x = torch.tensor(case['online_quantile_values'], device='cuda', dtype=torch.float64,
                   requires_grad=True)
  tgt = torch.tensor(want['target'], device='cuda', dtype=torch.float64)  # B x N'
  act = torch.as_tensor(np.tile(case['actions'], n), device='cuda').long()
  chosen = x[torch.arange(n * batch, device='cuda'), act].reshape(n, batch).t()  # B x N
  err = tgt[:, :, None] - chosen[:, None, :]
  huber = torch.where(err.abs() <= kappa, 0.5 * err ** 2,
                      kappa * (err.abs() - 0.5 * kappa))
  taus = torch.tensor(case['quantiles'], device='cuda',
                      dtype=torch.float64).reshape(n, batch).t()
  w = (taus[:, None, :] - (err < 0).double()).abs()
  loss = (w * huber / kappa).sum(2).mean(1)
  loss.mean().backward()
  np.testing.assert_allclose(got['grad'].cpu().numpy(), x.grad.cpu().numpy(),
                             rtol=1e-4, atol=1e-8)
  np.testing.assert_allclose(got['loss'].cpu().numpy(), loss.detach().cpu().numpy(),
                             rtol=2e-6, atol=1e-7)


@pytest.mark.gpu
def test_iqn_autograd_wrapper_and_errors(iq):
  import torch
  rng = np.random.RandomState(0)
  case = _case(rng, 16, 5, 8, 8, 4)
  dev = lambda x: torch.as_tensor(x, device='cuda')
  online = dev(case['online_quantile_values']).requires_grad_(True)
  mean, rows = iq.QuantileHuberLoss.apply(
      online, dev(case['quantiles']), dev(case['target_quantile_values']),
      dev(case['action_quantile_values']), dev(case['actions']), dev(case['rewards']),
      dev(case['terminals']), 0.99, 1.0)
  (2.0 * mean).backward()
  want = iqn_port.iqn_update(gamma=0.99, update_horizon=1, **case)
  np.testing.assert_allclose(online.grad.cpu().numpy(), 2.0 * want['grad'], rtol=1e-4,
                             atol=1e-8)
  np.testing.assert_allclose(rows.cpu().numpy(), want['loss'], rtol=2e-6, atol=1e-7)
  with pytest.raises(ValueError, match='samples \\* batch'):
    iq.quantile_huber_loss(online.detach()[:-1], dev(case['quantiles']),
                           dev(case['target_quantile_values']),
                           dev(case['action_quantile_values']), dev(case['actions']),
                           dev(case['rewards']), dev(case['terminals']), 0.99)
  with pytest.raises(ValueError, match='kappa must be positive'):
    iq.quantile_huber_loss(online.detach(), dev(case['quantiles']),
                           dev(case['target_quantile_values']),
                           dev(case['action_quantile_values']), dev(case['actions']),
                           dev(case['rewards']), dev(case['terminals']), 0.99, kappa=0.0)
