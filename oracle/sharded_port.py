"""CPU statement of the sharded (multi-GPU) sampling rule.  TEST INFRASTRUCTURE ONLY.

The reference is single-process and has no notion of shards, so
there is nothing to restate from it beyond SumTree.sample: the cross-shard step is
DEFINED here as the natural extension of sum_tree.py:126-141 — the G shard totals
act as one more tree level scanned left to right in rank order with the same
"q < left ? take it : (q -= left, go on)" rule — and every shard then behaves
exactly like a reference OutOfGraphPrioritizedReplayBuffer of its own
(prioritized_replay_buffer.py:142-171) for the strata it owns.
"""
import numpy as np


def grand_total(totals):
  acc = np.float64(0.0)
  for t in totals:
    acc = acc + np.float64(t)
  return acc


def apportion(totals, queries01):
  """[(owner rank, residual mass inside that shard)] for each stratum query."""
  total = grand_total(totals)
  out = []
  for q in queries01:
    mass = np.float64(q) * total
    owner = 0
    while owner < len(totals) - 1:
      left = np.float64(totals[owner])
      if mass < left:
        break
      mass = mass - left
      owner += 1
    out.append((owner, mass))
  return out


def sharded_sample(shards, queries01, retry_streams):
  """shards: list of PortPrioritizedReplay; retry_streams[g]: uniforms shard g may
  consume for its retries.  Returns per shard (slots, indices, draws_used)."""
  totals = [s.sum_tree.total() for s in shards]
  owners = apportion(totals, queries01)
  result = []
  for g, shard in enumerate(shards):
    slots = [i for i, (o, _) in enumerate(owners) if o == g]
    picked = [shard.sum_tree.descend(owners[i][1]) for i in slots]
    stream = list(retry_streams[g])
    budget = len(stream)
    used = 0
    for k in range(len(picked)):
      if shard.is_valid_transition(picked[k]):
        continue
      if budget - used == 0:
        raise RuntimeError('shard {}: attempts exhausted at slot {}'.format(
            g, slots[k]))
      cand = picked[k]
      while not shard.is_valid_transition(cand) and used < budget:
        cand = shard.sum_tree.descend(np.float64(stream[used]) *
                                      shard.sum_tree.total())
        used += 1
      picked[k] = cand
    result.append((slots, picked, used))
  return result
