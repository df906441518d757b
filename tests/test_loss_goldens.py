"""The loss oracles against the reference's own loss-building code.

tests/golden/losses.npz was written by executing, unmodified, RainbowAgent.
_build_target_distribution / _build_train_op (+ project_distribution), DQNAgent.
_build_networks / _build_target_q_op / _build_train_op and ImplicitQuantileAgent.
_build_networks / _build_target_quantile_values_op / _build_train_op, with numpy
stand-ins for the TensorFlow ops they call (oracle/tfshim.py; generator:
oracle/make_golden.py:golden_losses).  That pins what the reference's code decides —
tiling, gathers, masks, reduction axes, operation order — not TensorFlow's kernels;
the ports must agree to float32 rounding of a reduction (1e-6 relative)."""
import numpy as np
import pytest

from oracle import c51_port
from oracle import dqn_port
from oracle import iqn_port
from tests import golden_cases

TOL = dict(rtol=1e-6, atol=1e-7)


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
  assert got['support'].tobytes() == g[p + 'support'].tobytes()
  np.testing.assert_allclose(got['target'], g[p + 'target'], **TOL)
  np.testing.assert_allclose(got['weighted_loss'], g[p + 'weighted_loss'], **TOL)
  np.testing.assert_allclose(got['weighted_loss'].mean(), g[p + 'mean_loss'], rtol=1e-6)
  if prioritized:
    np.testing.assert_allclose(got['priorities'], g[p + 'priorities'], **TOL)
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
  np.testing.assert_allclose(got['target'], g[p + 'target'], **TOL)
  np.testing.assert_allclose(got['loss'], g[p + 'loss'], **TOL)
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
  np.testing.assert_allclose(got['target'], want_target, **TOL)
  np.testing.assert_allclose(got['loss'], g[p + 'loss'], **TOL)
  np.testing.assert_allclose(got['mean_loss'], g[p + 'mean_loss'], rtol=1e-6)
