"""CPU model of the grouping shortcut of the early tree write-back (csrc/tree.cu:
nearly_sorted_analyse / nearly_sorted_insertions / nearly_sorted_positions).

The kernel groups a batch by (node on a level, batch position) without sorting when the
batch is in leaf order already but for a few entries — what a stratified sample is
(sum_tree.py:162-166: the queries grow with the stratum, the descent is monotone) apart
from the rows whose invalid pick was drawn again (prioritized_replay_buffer.py:155-170).
This file restates the rules in numpy and checks the ARGUMENT on the CPU: whenever the
rules accept a batch, the order they produce is the stable sort by (key >> shift, k), for
every shift; and they accept what they are meant for.  The kernel itself is held to the
sequential oracle bit for bit on the GPU (tests/test_gpu_parity.py::test_tree_early_*)."""
import numpy as np
import pytest

K_MAX_MOVED = 64  # tree.cu: kMaxMoved
PAD = 0xffffffff


def analyse(keys):
  """Suspects = both ends of every descent; the rest must be non-decreasing (checked
  against its running maximum).  Returns (moved mask, cleaned keys) or None."""
  keys = np.asarray(keys, dtype=np.uint64)
  n = len(keys)
  left = np.concatenate(([0], keys[:-1]))
  right = np.concatenate((keys[1:], [PAD]))
  moved = (keys < left) | (keys > right)
  if moved.sum() > K_MAX_MOVED:
    return None
  cleaned = keys.copy()
  running = 0
  for k in range(n):
    if moved[k]:
      cleaned[k] = running          # a suspect's slot carries the running maximum
    else:
      if keys[k] < running:
        return None                 # the rest is not in order: the kernel sorts
      running = keys[k]
  return moved, cleaned


def positions(keys, moved, cleaned, shift):
  """Where every entry goes in the order by (key >> shift, k)."""
  keys = np.asarray(keys, dtype=np.uint64)
  n = len(keys)
  c = cleaned >> np.uint64(shift)
  suspects = np.nonzero(moved)[0]
  # insertion point of a suspect: the first slot that sorts behind it (one monotone
  # predicate over the slots, found by binary search)
  insertion = {}
  for k in suspects:
    x = int(keys[k]) >> shift
    lo, hi = 0, n
    while lo < hi:
      mid = (lo + hi) // 2
      if int(c[mid]) > x or (int(c[mid]) == x and mid > k):
        hi = mid
      else:
        lo = mid + 1
    insertion[k] = lo
  before_moved = np.concatenate(([0], np.cumsum(moved)[:-1]))
  pos = np.empty(n, dtype=np.int64)
  for k in range(n):
    if not moved[k]:
      pos[k] = k - before_moved[k] + sum(1 for o in suspects if insertion[o] <= k)
    else:
      x = int(keys[k]) >> shift
      at = insertion[k]
      p = at
      for o in suspects:
        xo = int(keys[o]) >> shift
        p -= 1 if o < at else 0      # a suspect's slot in front of the insertion point
        p += 1 if (xo < x or (xo == x and o < k)) else 0
      pos[k] = p
  return pos


def stable_order(keys, shift):
  k = np.asarray(keys, dtype=np.uint64) >> np.uint64(shift)
  return np.argsort(k, kind='stable')


def stratified(rng, leaves, n, redrawn):
  idx = np.sort(rng.randint(0, leaves, size=n)).astype(np.uint64)
  if redrawn:
    where = rng.choice(n, size=redrawn, replace=False)
    idx[where] = rng.randint(0, leaves, size=redrawn)
  return idx


@pytest.mark.parametrize('n,redrawn', [(33, 0), (64, 1), (256, 2), (256, 9), (1024, 4),
                                       (1024, 25)])
def test_accepted_batches_come_out_in_stable_order_on_every_level(n, redrawn):
  rng = np.random.RandomState(n + redrawn)
  depth = 20
  accepted = 0
  for trial in range(30):
    idx = stratified(rng, 1 << depth, n, redrawn)
    if trial % 3 == 0 and n > 8:      # neighbours that share leaves, a redrawn duplicate
      idx[3] = idx[2]
      idx[n // 2] = idx[n - 1]
    padded = np.concatenate((idx, np.full(7, PAD, dtype=np.uint64)))  # the chunk's pads
    got = analyse(padded)
    if got is None:
      continue
    accepted += 1
    moved, cleaned = got
    for shift in (0, 1, 5, 10, 19, 20):
      pos = positions(padded, moved, cleaned, shift)
      order = np.empty(len(padded), dtype=np.int64)
      order[pos] = np.arange(len(padded))
      assert sorted(pos.tolist()) == list(range(len(padded))), 'not a permutation'
      assert order.tolist() == stable_order(padded, shift).tolist(), (trial, shift)
  # (declined now and then with many redrawn rows: two neighbours redrawn upwards leave a
  # rest that is not in order — the kernel sorts those batches)
  assert accepted >= (28 if redrawn <= 4 else 18), 'the shortcut is meant for these batches'


def test_batches_in_random_order_are_declined():
  rng = np.random.RandomState(5)
  for n in (64, 300, 1024):
    assert analyse(rng.randint(0, 1 << 20, size=n)) is None


def test_a_run_out_of_order_is_declined_not_misplaced():
  """More entries out of place than the list takes, or a rest that is not in order:
  the rules must say no (the kernel then sorts), never produce an order."""
  rng = np.random.RandomState(6)
  idx = np.sort(rng.randint(0, 1 << 20, size=512)).astype(np.uint64)
  tail = idx[300:420].copy()
  rng.shuffle(tail)
  idx[300:420] = tail
  got = analyse(idx)
  if got is not None:  # (accepted only if it still comes out right)
    moved, cleaned = got
    pos = positions(idx, moved, cleaned, 0)
    order = np.empty(len(idx), dtype=np.int64)
    order[pos] = np.arange(len(idx))
    assert order.tolist() == stable_order(idx, 0).tolist()
  # two blocks swapped: ONE descent, two suspects — but the rest is not non-decreasing
  ordered = np.sort(rng.randint(0, 1 << 20, size=512)).astype(np.uint64)
  assert analyse(np.concatenate((ordered[256:], ordered[:256]))) is None


def test_every_level_needs_no_more_suspects_than_the_leaves():
  """The kernel analyses the LEAF keys once and reuses the suspects on its own level: a
  descent of the shifted keys is a descent of the leaf keys, so the leaf level's suspects
  are a superset of any level's, and the rest stays in order under any shift."""
  rng = np.random.RandomState(7)
  for _ in range(50):
    idx = stratified(rng, 1 << 20, 256, 6)
    got = analyse(idx)
    if got is None:
      continue
    moved, _ = got
    for shift in (1, 4, 12, 20):
      k = idx >> np.uint64(shift)
      left = np.concatenate(([0], k[:-1]))
      right = np.concatenate((k[1:], [PAD]))
      level_moved = (k < left) | (k > right)
      assert not np.any(level_moved & ~moved)
      rest = k[~moved]
      assert np.all(rest[1:] >= rest[:-1])
