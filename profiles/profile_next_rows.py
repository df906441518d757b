"""One IQN loss launch (batch 32 and 256, 64 x 64 tau samples) and a few actor-state
records between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

  python profiles/profile_next_rows.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  import torch
  from dopamine_b200.agents.dqn import dqn_agent
  from dopamine_b200.agents.implicit_quantile import implicit_quantile_agent as iqa
  rng = np.random.RandomState(0)
  n = n_prime = 64
  k, actions = 32, 18
  cases = []
  for batch in (32, 256):
    dev = lambda x: torch.as_tensor(x, device='cuda')
    cases.append(dict(
        online_quantile_values=dev(rng.randn(n * batch, actions).astype(np.float32)),
        quantiles=dev(rng.rand(n * batch, 1).astype(np.float32)),
        target_quantile_values=dev(rng.randn(n_prime * batch, actions).astype(np.float32)),
        action_quantile_values=dev(rng.randn(k * batch, actions).astype(np.float32)),
        actions=dev(rng.randint(0, actions, size=batch).astype(np.int32)),
        rewards=dev(np.clip(rng.randn(batch), -1, 1).astype(np.float32)),
        terminals=dev((rng.rand(batch) < 0.05).astype(np.uint8))))
  actor = dqn_agent.ActorState((84, 84), 4)
  frame = rng.randint(0, 256, size=(84, 84)).astype(np.uint8)

  def run():
    for c in cases:
      iqa.quantile_huber_loss(c['online_quantile_values'], c['quantiles'],
                              c['target_quantile_values'], c['action_quantile_values'],
                              c['actions'], c['rewards'], c['terminals'], 0.99 ** 3,
                              1.0, want_grad=True)
    actor.record(frame)

  for _ in range(3):
    run()
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  run()
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()
  print('profiled: iqn_loss x2, record_observation x1')


if __name__ == '__main__':
  main()
