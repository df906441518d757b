"""numpy f32 restatement of Rainbow's C51 target / loss math.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this; the product (`dopamine_b200`) never does.

The reference implements this path as TensorFlow-1.x graph code
(/root/reference/dopamine/agents/rainbow/rainbow_agent.py); TensorFlow is not
installed and cannot be, so this is a restatement, function by function:
  * support = linspace(-vmax, vmax, N) in f32 ....... rainbow_agent.py:124-126
  * softmax / q-values of the network head .......... discrete_domains/atari_lib.py:141-143
  * Bellman target support, argmax, gather .......... rainbow_agent.py:218-251
  * project_distribution (dense N x N form) ......... rainbow_agent.py:381-494
  * softmax cross-entropy, IS weights, priority ..... rainbow_agent.py:262-293
  * cumulative_gamma = math.pow(gamma, n) ........... agents/dqn/dqn_agent.py:175

Parity: `project_distribution` is PINNED by the six known-answer vectors of
tests/dopamine/agents/rainbow/rainbow_agent_test.py:178-285 (see
tests/test_oracle_golden.py).  Loss / priority / IS-weight values are not pinned
by any reference TEST; they are pinned to the reference's CODE:
tests/golden/losses.npz holds what RainbowAgent._build_target_distribution and
_build_train_op (rainbow_agent.py:200-305) compute when executed unmodified over numpy
stand-ins for the TensorFlow ops they call (oracle/tfshim.py, generator
oracle/make_golden.py:golden_losses), and tests/test_loss_goldens.py holds this port to
it at 1e-6 relative.  That fixes every structural decision of the reference (gathers,
masks, axes, operation order), not the rounding of TensorFlow's own kernels.
"""
import math

import numpy as np

F32 = np.float32


def make_support(vmax, num_atoms):
  """f32 `start + i*step` like tf.linspace (rainbow_agent.py:126)."""
  vmax = F32(vmax)
  if num_atoms == 1:
    return np.array([-vmax], dtype=F32)
  step = (vmax - (-vmax)) / F32(num_atoms - 1)
  return (-vmax + np.arange(num_atoms, dtype=F32) * step).astype(F32)


def softmax(logits):
  x = logits.astype(F32)
  x = x - x.max(axis=-1, keepdims=True)
  e = np.exp(x)
  return (e / e.sum(axis=-1, keepdims=True)).astype(F32)


def log_softmax(logits):
  x = logits.astype(F32)
  x = x - x.max(axis=-1, keepdims=True)
  return (x - np.log(np.exp(x).sum(axis=-1, keepdims=True))).astype(F32)


def project_distribution(supports, weights, target_support):
  """Dense form of Eq. 7 (rainbow_agent.py:381-494), all f32."""
  supports = np.asarray(supports, dtype=F32)
  weights = np.asarray(weights, dtype=F32)
  z = np.asarray(target_support, dtype=F32)
  if z.ndim != 1:
    raise ValueError('target_support must have rank 1')
  if supports.shape != weights.shape or supports.shape[-1] != z.shape[0]:
    raise ValueError('shapes are incompatible')
  dz = z[1] - z[0]
  clipped = np.clip(supports, z[0], z[-1])[:, None, :]       # (B, 1, N)
  gap = np.abs(clipped - z[None, :, None])                   # (B, N, N)
  hat = np.clip(F32(1) - gap / dz, F32(0), F32(1))
  return (hat * weights[:, None, :]).sum(axis=2).astype(F32)


def target_distribution(rewards, terminals, target_logits, support, gamma,
                        update_horizon):
  """rainbow_agent.py:218-251 given the target network's logits (B, A, N)."""
  rewards = np.asarray(rewards, dtype=F32)
  live = F32(1) - np.asarray(terminals).astype(F32)
  gamma_n = F32(math.pow(gamma, update_horizon))
  bellman = rewards[:, None] + (gamma_n * live)[:, None] * support[None, :]
  probs = softmax(target_logits)
  q = (support * probs).sum(axis=2)
  best = np.argmax(q, axis=1)
  next_probs = probs[np.arange(len(best)), best]
  return project_distribution(bellman, next_probs, support), best


def loss_and_priorities(target, online_logits, actions, sampling_probs):
  """rainbow_agent.py:262-293.

  Returns (per-row cross entropy, new priorities, normalised IS weights,
  weighted loss).
  """
  b = np.arange(len(actions))
  chosen = online_logits[b, np.asarray(actions)]
  ce = -(target * log_softmax(chosen)).sum(axis=1).astype(F32)
  new_prio = np.sqrt(ce + F32(1e-10)).astype(F32)
  w = (F32(1.0) / np.sqrt(np.asarray(sampling_probs, dtype=F32) + F32(1e-10)))
  w = (w / w.max()).astype(F32)
  return ce, new_prio, w, (w * ce).astype(F32)


def rainbow_update(rewards, terminals, actions, sampling_probs, online_logits,
                   target_logits, vmax=10., num_atoms=51, gamma=0.99,
                   update_horizon=3):
  support = make_support(vmax, num_atoms)
  target, best = target_distribution(rewards, terminals, target_logits, support,
                                     gamma, update_horizon)
  ce, prio, w, wl = loss_and_priorities(target, online_logits, actions,
                                        sampling_probs)
  return dict(target=target, argmax=best, loss=ce, priorities=prio, weights=w,
              weighted_loss=wl, support=support)
