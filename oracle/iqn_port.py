"""numpy f32 restatement of IQN's target and quantile-Huber loss.  TEST INFRASTRUCTURE ONLY.

Follows dopamine/agents/implicit_quantile/implicit_quantile_agent.py op by op, keeping
the reference's tiling conventions ((samples x batch) rows, sample-major):
  :166-188  greedy next action: reshape the (K*B, A) quantile values of the action
            network to (K, B, A), reduce_mean over K, argmax over A (first maximum);
  :190-231  target quantile values r + gamma^n (1 - terminal) * Z_target[t'*B + b, a*];
  :233-315  Bellman errors target[b, t'] - chosen[b, t], the two-case Huber loss, the
            |tau - 1[error < 0]| weighting, / kappa, reduce_sum over t, reduce_mean
            over t', and the scalar loss reduce_mean over the batch.

Parity: the reference holds no numeric test of this loss
(tests/dopamine/agents/implicit_quantile/implicit_quantile_agent_test.py checks shapes
and q-values only) and TensorFlow is not importable here.  The port is pinned to the
reference's CODE: ImplicitQuantileAgent._build_networks, _build_target_quantile_values_op
and _build_train_op executed unmodified over numpy stand-ins for the TensorFlow ops they
call (oracle/tfshim.py; generator oracle/make_golden.py:golden_losses) wrote
tests/golden/losses.npz, and tests/test_loss_goldens.py holds this port to it at 1e-6
relative (greedy actions exact).  That fixes the tiling / gather / transpose / mask /
reduction structure, not the rounding of TensorFlow's kernels.  `closed_form_f64` below
is an independent float64 statement of the same mathematics, a second check
(tests/test_iqn.py); tolerance of the CUDA path against this port: 2e-6 relative (f32
summation order differs).
"""
import math

import numpy as np

F32 = np.float32


def iqn_update(rewards, terminals, actions, online_quantile_values, quantiles,
               target_quantile_values, action_quantile_values, num_tau_samples,
               num_tau_prime_samples, num_quantile_samples, kappa=1.0, gamma=0.99,
               update_horizon=1):
  rewards = np.asarray(rewards, dtype=F32)
  batch = rewards.shape[0]
  online = np.asarray(online_quantile_values, dtype=F32)
  target_net = np.asarray(target_quantile_values, dtype=F32)
  action_net = np.asarray(action_quantile_values, dtype=F32)
  num_actions = online.shape[1]
  n, n_prime, k = num_tau_samples, num_tau_prime_samples, num_quantile_samples
  kappa = F32(kappa)
  gamma_n = F32(math.pow(gamma, update_horizon))  # dqn_agent.py:175

  # :176-188
  q_values = action_net.reshape(k, batch, num_actions).mean(axis=0, dtype=F32)
  next_action = np.argmax(q_values, axis=1)

  # :196-231 (everything tiled num_tau_prime_samples times, sample-major)
  tiled_rewards = np.tile(rewards[:, None], [n_prime, 1])
  live = (F32(1.) - np.asarray(terminals).astype(F32)).astype(F32)
  gamma_with_terminal = np.tile((gamma_n * live).astype(F32)[:, None], [n_prime, 1])
  tiled_argmax = np.tile(next_action[:, None], [n_prime, 1])
  rows = np.arange(n_prime * batch)
  gathered = target_net[rows, tiled_argmax[:, 0]][:, None]
  target = (tiled_rewards + (gamma_with_terminal * gathered).astype(F32)).astype(F32)

  # :245-276
  target = target.reshape(n_prime, batch, 1).transpose(1, 0, 2)  # B x N' x 1
  tiled_actions = np.tile(np.asarray(actions)[:, None], [n, 1])
  chosen = online[np.arange(n * batch), tiled_actions[:, 0]]
  chosen = chosen.reshape(n, batch, 1).transpose(1, 0, 2)  # B x N x 1

  # :278-290
  errors = (target[:, :, None, :] - chosen[:, None, :, :]).astype(F32)  # B x N' x N x 1
  abs_err = np.abs(errors)
  case_one = ((abs_err <= kappa).astype(F32) * F32(0.5) * (errors ** 2).astype(F32)
             ).astype(F32)
  case_two = ((abs_err > kappa).astype(F32) * kappa *
              (abs_err - F32(0.5) * kappa).astype(F32)).astype(F32)
  huber = (case_one + case_two).astype(F32)

  # :292-305
  taus = np.asarray(quantiles, dtype=F32).reshape(n, batch, 1).transpose(1, 0, 2)
  taus = np.tile(taus[:, None, :, :], [1, n_prime, 1, 1])
  weight = np.abs(taus - (errors < 0).astype(F32)).astype(F32)
  quantile_huber = ((weight * huber).astype(F32) / kappa).astype(F32)

  # :306-311
  loss = quantile_huber.sum(axis=2, dtype=F32).mean(axis=1, dtype=F32)[:, 0]
  mean_loss = F32(loss.mean(dtype=F32))

  # d mean(loss) / d online_quantile_values (what the optimizer's backward produces):
  # d huber / d error = error (case one) or kappa sign(error) (case two); the error
  # falls with the chosen value; the indicator is behind tf.stop_gradient (:304).
  dh = np.where(abs_err <= kappa, errors, kappa * np.sign(errors)).astype(np.float64)
  g = -(weight.astype(np.float64) * dh / np.float64(kappa)).sum(axis=1)[:, :, 0]  # B x N
  g = g / np.float64(n_prime) / np.float64(batch)
  grad = np.zeros((n * batch, num_actions), dtype=F32)
  grad[np.arange(n * batch), tiled_actions[:, 0]] = g.T.reshape(-1).astype(F32)
  return dict(next_action=next_action.astype(np.int32), q_values=q_values,
              target=target[:, :, 0], loss=loss, mean_loss=mean_loss, grad=grad)


def closed_form_f64(rewards, terminals, actions, online_quantile_values, quantiles,
                    target_quantile_values, next_action, num_tau_samples,
                    num_tau_prime_samples, kappa=1.0, gamma=0.99, update_horizon=1):
  """rho^kappa_tau(delta) of Dabney et al. 2018 (eq. 3-4) in float64 with plain
  loops over (b, t', t): the independent statement the port is checked against."""
  batch = len(rewards)
  n, n_prime = num_tau_samples, num_tau_prime_samples
  online = np.asarray(online_quantile_values, dtype=np.float64)
  target_net = np.asarray(target_quantile_values, dtype=np.float64)
  taus = np.asarray(quantiles, dtype=np.float64).reshape(-1)
  gamma_n = np.float64(F32(math.pow(gamma, update_horizon)))
  loss = np.zeros(batch)
  for b in range(batch):
    live = 1.0 - float(terminals[b])
    total = 0.0
    for tp in range(n_prime):
      z = float(rewards[b]) + gamma_n * live * target_net[tp * batch + b, next_action[b]]
      for t in range(n):
        delta = z - online[t * batch + b, actions[b]]
        if abs(delta) <= kappa:
          huber = 0.5 * delta * delta
        else:
          huber = kappa * (abs(delta) - 0.5 * kappa)
        total += abs(taus[t * batch + b] - (1.0 if delta < 0 else 0.0)) * huber / kappa
    loss[b] = total / n_prime
  return loss
