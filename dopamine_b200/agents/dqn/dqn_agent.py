"""DQN's target / loss math as a fused CUDA kernel.

Drop-in for the hot-path pieces of `dopamine/agents/dqn/dqn_agent.py`:
`_build_target_q_op` (dqn_agent.py:283-300) and the Huber loss of `_build_train_op`
(dqn_agent.py:302-322), over torch CUDA tensors, in one launch (+ a one-CTA
fixed-order mean).  The replay side is `OutOfGraphReplayBuffer` /
`WrappedReplayBuffer` of `dopamine_b200.replay_memory.circular_replay_buffer`
(uniform sampling, BASELINE config 1).
"""
import ctypes
import math

import numpy as np

from dopamine_b200 import _native


def _torch():
  import torch  # pylint: disable=g-import-not-at-top
  return torch


def cumulative_gamma(gamma, update_horizon):
  """dqn_agent.py:175."""
  return math.pow(gamma, update_horizon)


def dqn_loss(online_q, target_q, actions, rewards, terminals, cumulative_gamma,  # pylint: disable=redefined-outer-name
             want_target=False, want_grad=False, want_mean=True, out=None):
  """Bellman target + Huber loss for one batch, on the device.

  Args:
    online_q: (B, A) f32 CUDA, online network on `state`.
    target_q: (B, A) f32 CUDA, target network on `next_state`.
    actions: (B,) int32; rewards: (B,) f32 n-step returns; terminals: (B,) uint8.
    cumulative_gamma: gamma ** update_horizon.
  Returns:
    dict with 'loss' (B,), optionally 'mean_loss' (scalar), 'target' (B,),
    'grad_q' (B, A) = d mean(loss) / d online_q.
  """
  torch = _torch()
  b, a = online_q.shape
  assert tuple(target_q.shape) == (b, a)
  assert online_q.dtype == torch.float32 and online_q.is_cuda
  assert target_q.dtype == torch.float32 and target_q.is_cuda
  assert actions.dtype == torch.int32 and rewards.dtype == torch.float32
  assert terminals.dtype == torch.uint8
  dev = online_q.device
  if out is None:
    out = {'loss': torch.empty(b, dtype=torch.float32, device=dev)}
    if want_mean:
      out['mean_loss'] = torch.empty((), dtype=torch.float32, device=dev)
    if want_target:
      out['target'] = torch.empty(b, dtype=torch.float32, device=dev)
    if want_grad:
      out['grad_q'] = torch.empty(b, a, dtype=torch.float32, device=dev)
  args = _native.DqnArgs()
  args.batch, args.num_actions = b, a
  args.cumulative_gamma = float(np.float32(cumulative_gamma))
  args.target_q = target_q.contiguous().data_ptr()
  args.online_q = online_q.contiguous().data_ptr()
  args.actions = actions.contiguous().data_ptr()
  args.rewards = rewards.contiguous().data_ptr()
  args.terminals = terminals.contiguous().data_ptr()
  args.loss = out['loss'].data_ptr()
  args.target = out['target'].data_ptr() if 'target' in out else None
  args.mean_loss = out['mean_loss'].data_ptr() if 'mean_loss' in out else None
  args.grad_q = out['grad_q'].data_ptr() if 'grad_q' in out else None
  _native.check(_native.lib().b2r_dqn_loss(ctypes.byref(args),
                                           _native.current_stream()))
  return out


class DQNLoss(object):
  """Differentiable wrapper: `mean_loss, loss = DQNLoss.apply(online_q, ...)`."""

  _fn = None

  @classmethod
  def apply(cls, online_q, target_q, actions, rewards, terminals, gamma_n):
    if cls._fn is None:
      torch = _torch()

      class _Fn(torch.autograd.Function):

        @staticmethod
        def forward(ctx, online_q, target_q, actions, rewards, terminals, gamma_n):  # pylint: disable=redefined-outer-name
          out = dqn_loss(online_q.detach(), target_q.detach(), actions, rewards,
                         terminals, gamma_n, want_grad=True)
          ctx.save_for_backward(out['grad_q'])
          ctx.mark_non_differentiable(out['loss'])
          return out['mean_loss'], out['loss']

        @staticmethod
        def backward(ctx, grad_mean, *unused):
          (grad_q,) = ctx.saved_tensors
          return grad_q * grad_mean, None, None, None, None, None

      cls._fn = _Fn
    return cls._fn.apply(online_q, target_q, actions, rewards, terminals, gamma_n)
