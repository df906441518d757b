"""Worst relative error of the C51 loss / priority kernels against the numpy port,
row by row, over many random batches — and of the port itself against a float64
evaluation of the same formulas (what rounding alone explains).

  python profiles/micro/c51_error.py [--rows 4096] [--seeds 4]

Prints one JSON line per (kernel instance, seed count)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def f64_reference(rewards, terminals, actions, probs, online, target, vmax=10.,
                  atoms=51, gamma=0.99, horizon=3):
  """The port's formulas in float64 from the same f32 inputs (f32 support and
  cumulative gamma as the reference builds them)."""
  from oracle import c51_port
  z = c51_port.make_support(vmax, atoms).astype(np.float64)
  gamma_n = np.float64(np.float32(gamma ** horizon))
  t = target.astype(np.float64)
  t = t - t.max(axis=-1, keepdims=True)
  p = np.exp(t)
  p /= p.sum(axis=-1, keepdims=True)
  q = (z * p).sum(axis=2)
  # the f32 argmax decides the row (ties aside): take the port's to compare like with like
  best = c51_port.target_distribution(rewards, terminals, target, c51_port.make_support(
      vmax, atoms), gamma, horizon)[1]
  nxt = p[np.arange(len(best)), best]
  live = 1.0 - terminals.astype(np.float64)
  s = rewards.astype(np.float64)[:, None] + (gamma_n * live)[:, None] * z[None, :]
  dz = z[1] - z[0]
  hat = np.clip(1.0 - np.abs(np.clip(s, z[0], z[-1])[:, None, :] - z[None, :, None]) / dz,
                0.0, 1.0)
  tgt = (hat * nxt[:, None, :]).sum(axis=2)
  x = online[np.arange(len(actions)), actions].astype(np.float64)
  x = x - x.max(axis=-1, keepdims=True)
  logp = x - np.log(np.exp(x).sum(axis=-1, keepdims=True))
  ce = -(tgt * logp).sum(axis=1)
  return ce, np.sqrt(ce + 1e-10)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rows', type=int, default=4096)
  ap.add_argument('--seeds', type=int, default=4)
  a = ap.parse_args()
  import torch
  from dopamine_b200.agents.rainbow import rainbow_agent
  from oracle import c51_port
  support = rainbow_agent.make_support(10., 51)
  for name, batch in (('cta_per_row', 32), ('warp_per_row', a.rows)):
    worst = dict(loss=0.0, prio=0.0, port_loss=0.0, port_prio=0.0, gpu64_loss=0.0,
                 gpu64_prio=0.0)
    over = dict(loss=0, prio=0)
    rows = 0
    for seed in range(a.seeds * (a.rows // batch if batch < a.rows else 1)):
      rng = np.random.RandomState(100 + seed)
      online = rng.randn(batch, 18, 51).astype(np.float32)
      target = rng.randn(batch, 18, 51).astype(np.float32)
      rewards = np.clip(rng.randn(batch), -1, 1).astype(np.float32)
      terminals = (rng.rand(batch) < 0.05).astype(np.uint8)
      actions = rng.randint(0, 18, size=batch).astype(np.int32)
      probs = np.sqrt(np.abs(rng.randn(batch)) + 1e-10).astype(np.float32)
      ref = c51_port.rainbow_update(rewards, terminals, actions, probs, online, target)
      ce64, pr64 = f64_reference(rewards, terminals, actions, probs, online, target)
      dev = lambda x: torch.as_tensor(x, device='cuda')
      out = rainbow_agent.c51_loss(dev(online), dev(target), dev(actions), dev(rewards),
                                   dev(terminals), dev(probs), support,
                                   float(np.float32(0.99 ** 3)))
      loss = out['loss'].cpu().numpy().astype(np.float64)
      prio = out['priorities'].cpu().numpy().astype(np.float64)
      rel = lambda x, y: np.abs(x - y) / np.abs(y)
      e_l, e_p = rel(loss, ref['loss']), rel(prio, ref['priorities'])
      worst['loss'] = max(worst['loss'], e_l.max())
      worst['prio'] = max(worst['prio'], e_p.max())
      worst['port_loss'] = max(worst['port_loss'], rel(ref['loss'], ce64).max())
      worst['port_prio'] = max(worst['port_prio'], rel(ref['priorities'], pr64).max())
      worst['gpu64_loss'] = max(worst['gpu64_loss'], rel(loss, ce64).max())
      worst['gpu64_prio'] = max(worst['gpu64_prio'], rel(prio, pr64).max())
      over['loss'] += int((e_l > 1e-6).sum())
      over['prio'] += int((e_p > 1e-6).sum())
      rows += batch
    print(json.dumps({'kernel': name, 'batch': batch, 'rows': rows,
                      'max_rel_vs_port': {k: float('%.3g' % worst[k]) for k in ('loss', 'prio')},
                      'rows_over_1e-6_vs_port': over,
                      'port_vs_f64': {'loss': float('%.3g' % worst['port_loss']),
                                      'prio': float('%.3g' % worst['port_prio'])},
                      'gpu_vs_f64': {'loss': float('%.3g' % worst['gpu64_loss']),
                                     'prio': float('%.3g' % worst['gpu64_prio'])}}),
          flush=True)


if __name__ == '__main__':
  main()
