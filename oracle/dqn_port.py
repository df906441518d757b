"""numpy f32 restatement of DQN's target and Huber loss.  TEST INFRASTRUCTURE ONLY.

Follows dopamine/agents/dqn/dqn_agent.py:283-300 (_build_target_q_op: r + gamma^n *
max_a Q_target(s') * (1 - terminal)) and :302-322 (_build_train_op: chosen q through a
one-hot, tf.losses.huber_loss with delta 1.0 and Reduction.NONE, then reduce_mean),
with tf.losses.huber_loss as TF 1.x defines it: error = predictions - labels,
quadratic = min(|error|, delta), linear = |error| - quadratic,
loss = 0.5 quadratic^2 + delta * linear.  No known-answer test exists for these values
in tests/dopamine/agents/dqn/dqn_agent_test.py and TensorFlow is absent; the port is
pinned to the reference's CODE instead: DQNAgent._build_networks / _build_target_q_op /
_build_train_op executed unmodified over numpy stand-ins for the TensorFlow ops
(oracle/tfshim.py) wrote tests/golden/losses.npz, which tests/test_loss_goldens.py
compares with at 1e-6 relative (structure pinned; TensorFlow's kernel rounding not).
"""
import math

import numpy as np

F32 = np.float32


def dqn_update(rewards, terminals, actions, online_q, target_q, gamma=0.99,
               update_horizon=1):
  online_q = np.asarray(online_q, dtype=F32)
  target_q = np.asarray(target_q, dtype=F32)
  gamma_n = F32(math.pow(gamma, update_horizon))  # dqn_agent.py:175
  best = target_q.max(axis=1)
  live = (F32(1.0) - np.asarray(terminals).astype(F32)).astype(F32)
  target = (np.asarray(rewards, dtype=F32) + (gamma_n * best).astype(F32) * live).astype(F32)
  chosen = online_q[np.arange(len(actions)), np.asarray(actions)]
  err = (chosen - target).astype(F32)
  abs_err = np.abs(err)
  quad = np.minimum(abs_err, F32(1.0))
  lin = (abs_err - quad).astype(F32)
  loss = (F32(0.5) * (quad * quad).astype(F32) + lin).astype(F32)
  grad = np.zeros_like(online_q)
  grad[np.arange(len(actions)), np.asarray(actions)] = (
      np.clip(err, -1, 1) / F32(len(actions)))
  return dict(target=target, loss=loss, mean_loss=F32(loss.mean(dtype=np.float64)),
              grad_q=grad)
