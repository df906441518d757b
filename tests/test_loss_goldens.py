"""The loss oracles against the reference's own loss-building code.

tests/golden/losses.npz was written by executing, unmodified, RainbowAgent.
_build_target_distribution / _build_train_op (+ project_distribution), DQNAgent.
_build_networks / _build_target_q_op / _build_train_op and ImplicitQuantileAgent.
_build_networks / _build_target_quantile_values_op / _build_train_op, with numpy
stand-ins for the TensorFlow ops they call (oracle/tfshim.py; generator:
oracle/make_golden.py:golden_losses).  That pins what the reference's code decides —
tiling, gathers, masks, reduction axes, operation order — not TensorFlow's kernels;
the ports reproduce it bit for bit, and the CUDA kernels (-m gpu) are held to the same
fixture at the tolerances of their port comparisons (north_star: 1e-6 relative)."""
import numpy as np
import pytest

from oracle import c51_port
from oracle import dqn_port
from oracle import iqn_port
from tests import golden_cases



def _same_bits(got, want, what):
  got, want = np.asarray(got), np.asarray(want)
  assert got.dtype == want.dtype and got.shape == want.shape, what
  assert got.tobytes() == want.tobytes(), what + ' differs from the reference code'


@pytest.mark.parametrize('case', ['c51_per', 'c51_uniform', 'c51_atoms11'])
def test_c51_port_matches_reference_code(case):
  g = golden_cases.load('losses')
  p = case + '_'
  _, _, num_atoms, horizon, prioritized = [int(x) for x in g[p + 'cfg']]
  # (the port always applies importance weights: all-equal probabilities give 1.0,
  # which is the uniform scheme, rainbow_agent.py:296-297)
  probs = g[p + 'probs'] if prioritized else np.ones(len(g[p + 'rewards']), np.float32)
  got = c51_port.rainbow_update(
      g[p + 'rewards'], g[p + 'terminals'], g[p + 'actions'], probs,
      g[p + 'online_logits'], g[p + 'target_logits'], vmax=10., num_atoms=num_atoms,
      gamma=0.99, update_horizon=horizon)
  _same_bits(got['support'], g[p + 'support'], 'support')
  _same_bits(got['target'], g[p + 'target'], 'target distribution')
  _same_bits(got['weighted_loss'], g[p + 'weighted_loss'], 'weighted loss')
  np.testing.assert_allclose(got['weighted_loss'].mean(), g[p + 'mean_loss'], rtol=1e-6)
  if prioritized:
    _same_bits(got['priorities'], g[p + 'priorities'], 'priorities')
  else:
    assert (got['weights'] == 1.0).all()


@pytest.mark.parametrize('case', ['dqn_a', 'dqn_b'])
def test_dqn_port_matches_reference_code(case):
  g = golden_cases.load('losses')
  p = case + '_'
  horizon = int(g[p + 'cfg'][2])
  got = dqn_port.dqn_update(g[p + 'rewards'], g[p + 'terminals'], g[p + 'actions'],
                            g[p + 'online_q'], g[p + 'target_q'], gamma=0.99,
                            update_horizon=horizon)
  _same_bits(got['target'], g[p + 'target'], 'target')
  _same_bits(got['loss'], g[p + 'loss'], 'loss')
  np.testing.assert_allclose(got['mean_loss'], g[p + 'mean_loss'], rtol=1e-6)


@pytest.mark.parametrize('case', ['iqn_a', 'iqn_kappa', 'iqn_paper'])
def test_iqn_port_matches_reference_code(case):
  g = golden_cases.load('losses')
  p = case + '_'
  _, _, n, n_prime, k, horizon = [int(x) for x in g[p + 'cfg']]
  got = iqn_port.iqn_update(
      g[p + 'rewards'], g[p + 'terminals'], g[p + 'actions'],
      g[p + 'online_quantile_values'], g[p + 'quantiles'],
      g[p + 'target_quantile_values'], g[p + 'action_quantile_values'], n, n_prime, k,
      kappa=float(g[p + 'kappa']), gamma=0.99, update_horizon=horizon)
  assert got['next_action'].tolist() == g[p + 'next_action'].tolist()
  # the reference keeps its targets tiled (N' * B, 1), sample-major
  want_target = g[p + 'target'].reshape(n_prime, -1).T
  _same_bits(got['target'], want_target, 'target quantile values')
  _same_bits(got['loss'], g[p + 'loss'], 'loss')
  np.testing.assert_allclose(got['mean_loss'], g[p + 'mean_loss'], rtol=1e-6)


# ------------------------------------------------------------------ CUDA path ----
@pytest.fixture(scope='module')
def cuda():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.dqn import dqn_agent
  from dopamine_b200.agents.implicit_quantile import implicit_quantile_agent
  from dopamine_b200.agents.rainbow import rainbow_agent

  class Mods(object):
    pass

  m = Mods()
  m.torch, m.ra, m.dqn, m.iq = torch, rainbow_agent, dqn_agent, implicit_quantile_agent
  m.dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), device='cuda')
  return m


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['c51_per', 'c51_uniform', 'c51_atoms11'])
def test_c51_kernel_matches_reference_code(cuda, case):
  g = golden_cases.load('losses')
  p = case + '_'
  _, _, num_atoms, horizon, prioritized = [int(x) for x in g[p + 'cfg']]
  support = cuda.ra.make_support(10., num_atoms)
  assert support.cpu().numpy().tobytes() == g[p + 'support'].tobytes()
  probs = cuda.dev(g[p + 'probs']) if prioritized else None
  got = cuda.ra.c51_loss(cuda.dev(g[p + 'online_logits']),
                         cuda.dev(g[p + 'target_logits']), cuda.dev(g[p + 'actions']),
                         cuda.dev(g[p + 'rewards']), cuda.dev(g[p + 'terminals']), probs,
                         support, 0.99 ** horizon, want_target=True)
  np.testing.assert_allclose(got['target'].cpu().numpy(), g[p + 'target'], rtol=1e-6,
                             atol=1e-6)
  weighted = (got['weights'] * got['loss']).cpu().numpy()
  np.testing.assert_allclose(weighted, g[p + 'weighted_loss'], rtol=1e-6, atol=1e-7)
  np.testing.assert_allclose(float(got['mean_weighted_loss']), g[p + 'mean_loss'],
                             rtol=1e-5)
  if prioritized:
    np.testing.assert_allclose(got['priorities'].cpu().numpy(), g[p + 'priorities'],
                               rtol=1e-6, atol=1e-7)
  else:
    assert (got['weights'].cpu().numpy() == 1.0).all()


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['dqn_a', 'dqn_b'])
def test_dqn_kernel_matches_reference_code(cuda, case):
  g = golden_cases.load('losses')
  p = case + '_'
  horizon = int(g[p + 'cfg'][2])
  got = cuda.dqn.dqn_loss(cuda.dev(g[p + 'online_q']), cuda.dev(g[p + 'target_q']),
                          cuda.dev(g[p + 'actions']), cuda.dev(g[p + 'rewards']),
                          cuda.dev(g[p + 'terminals']), 0.99 ** horizon,
                          want_target=True)
  np.testing.assert_allclose(got['target'].cpu().numpy(), g[p + 'target'], rtol=1e-6,
                             atol=1e-7)
  np.testing.assert_allclose(got['loss'].cpu().numpy(), g[p + 'loss'], rtol=1e-6,
                             atol=1e-7)
  np.testing.assert_allclose(float(got['mean_loss']), g[p + 'mean_loss'], rtol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['iqn_a', 'iqn_kappa', 'iqn_paper'])
def test_iqn_kernel_matches_reference_code(cuda, case):
  g = golden_cases.load('losses')
  p = case + '_'
  horizon = int(g[p + 'cfg'][5])
  got = cuda.iq.quantile_huber_loss(
      cuda.dev(g[p + 'online_quantile_values']), cuda.dev(g[p + 'quantiles']),
      cuda.dev(g[p + 'target_quantile_values']),
      cuda.dev(g[p + 'action_quantile_values']), cuda.dev(g[p + 'actions']),
      cuda.dev(g[p + 'rewards']), cuda.dev(g[p + 'terminals']), 0.99 ** horizon,
      float(g[p + 'kappa']))
  assert got['next_action'].cpu().numpy().tolist() == g[p + 'next_action'].tolist()
  np.testing.assert_allclose(got['loss'].cpu().numpy(), g[p + 'loss'], rtol=2e-6,
                             atol=1e-7)
  np.testing.assert_allclose(float(got['mean_loss']), g[p + 'mean_loss'], rtol=2e-6)
