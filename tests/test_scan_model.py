"""A numpy model of `chains_by_verified_scan` (dopamine_b200/csrc/tree.cu): the argument
that lets the upper tree levels replace n dependent fp64 adds by a prefix scan without
giving up bit-exactness, executable on the CPU.

The kernel is checked on the GPU against the C oracle (tests/test_gpu_parity.py::
test_tree_long_chains_verified_scan_bit_exact); this file checks the ALGORITHM, with the
kernel's own structure (4 entries per thread, 32-lane Hillis-Steele warp scans, a scan
of the warp totals, per-thread re-walk from the carry, bit-for-bit verification of every
carry against the left neighbour's last value, promotion of failed carries to segment
heads, at most 3 rounds):

  * whenever the model accepts, its node values ARE the sequential chains, bit for bit
    — for exact data (f32 priorities under a large root), for data where a few adds
    round (repaired by re-rooting) and for fully inexact data (declined);
  * exact data is accepted in the first round (that is the performance claim).
"""
import numpy as np
import pytest

ITEMS, ROUNDS = 4, 3


def sequential_chains(node_vals, nodes, deltas):
  """The reference: per node, add its deltas in order, rounding after each add."""
  out = dict(node_vals)
  for node, d in zip(nodes, deltas):
    out[node] = np.float64(out[node] + d)
  return out


def _segscan(v, f, width):
  """Inclusive segmented Hillis-Steele scan over groups of `width` lanes, the kernel's
  loop: `if lane >= o: if not f: v = v[lane - o] + v; f |= f[lane - o]`."""
  v, f = v.copy(), f.copy()
  n = len(v)
  lane = np.arange(n) % width
  o = 1
  while o < width:
    pv = np.roll(v, o)
    pf = np.roll(f, o)
    act = lane >= o
    add = act & ~f
    v = np.where(add, pv + v, v)
    f = np.where(act, f | pf, f)
    o <<= 1
  return v, f


def verified_scan(node_vals, nodes, deltas, threads):
  """Returns (accepted, rounds_used, {node: value}); mirrors the kernel step by step."""
  n_eff = len(nodes)
  assert n_eff <= threads * ITEMS and threads % 32 == 0
  warps = threads // 32
  nodes = np.asarray(nodes, dtype=np.int64)
  d = np.zeros(threads * ITEMS)
  d[:n_eff] = deltas
  nd = np.full(threads * ITEMS, -1, dtype=np.int64)
  nd[:n_eff] = nodes
  inside = np.arange(threads * ITEMS) < n_eff
  before = np.concatenate([[-2], nd[:-1]])
  head = ~inside | (before != nd)
  hv = np.zeros(threads * ITEMS)
  for p in np.nonzero(head & inside)[0]:
    hv[p] = np.float64(node_vals[nd[p]] + d[p])
  d, nd, head, hv = (x.reshape(threads, ITEMS) for x in (d, nd, head, hv))
  head = head.copy()
  inside_first = inside.reshape(threads, ITEMS)[:, 0]
  checked = (np.arange(threads) > 0) & inside_first & ~head[:, 0]
  promoted = np.zeros(threads, dtype=bool)
  for rnd in range(ROUNDS):
    # thread aggregates
    v = np.where(head[:, 0], hv[:, 0], d[:, 0])
    f = head[:, 0].copy()
    for j in range(1, ITEMS):
      v = np.where(head[:, j], hv[:, j], v + d[:, j])
      f = f | head[:, j]
    iv, iflag = _segscan(v, f, 32)                       # warp scans
    wv, wf = iv[31::32].copy(), iflag[31::32].copy()     # warp totals
    wv = np.concatenate([wv, np.zeros(32 - warps)])
    wf = np.concatenate([wf, np.ones(32 - warps, dtype=bool)])
    wv, _ = _segscan(wv, wf, 32)
    lane = np.arange(threads) % 32
    warp = np.arange(threads) // 32
    ev = np.roll(iv, 1)
    ef = np.roll(iflag, 1)
    prefix = np.where(warp > 0, wv[np.maximum(warp - 1, 0)], 0.0)
    carry = np.where(lane > 0, np.where(ef | (warp == 0), ev, prefix + ev),
                     np.where(warp > 0, prefix, 0.0))
    # re-walk from the carry
    acc = carry.copy()
    P = np.zeros((threads, ITEMS))
    for j in range(ITEMS):
      acc = np.where(head[:, j], hv[:, j], acc + d[:, j])
      P[:, j] = acc
    last = acc
    left = np.roll(last, 1)
    need = left + d[:, 0]
    bad_promoted = checked & promoted & (need.view(np.uint64) != hv[:, 0].view(np.uint64))
    bad_plain = checked & ~promoted & (carry.view(np.uint64) != left.view(np.uint64))
    hv[:, 0] = np.where(bad_promoted | bad_plain, need, hv[:, 0])
    head[:, 0] = head[:, 0] | bad_plain
    promoted = promoted | bad_plain
    if not (bad_promoted | bad_plain).any():
      flat_p, flat_nd = P.reshape(-1), nd.reshape(-1)
      out = dict(node_vals)
      for p in range(n_eff):
        if p + 1 >= n_eff or flat_nd[p + 1] != flat_nd[p]:
          out[int(flat_nd[p])] = np.float64(flat_p[p])
      return True, rnd + 1, out
  return False, ROUNDS, None


def _case(rng, n, num_nodes, odd, root=2.0 ** 20):
  nodes = np.sort(rng.randint(0, num_nodes, size=n))
  deltas = (np.sqrt(np.abs(rng.randn(n))).astype(np.float32).astype(np.float64) -
            np.sqrt(np.abs(rng.randn(n))).astype(np.float32).astype(np.float64))
  if odd:
    deltas[rng.choice(n, size=min(odd, n), replace=False)] = rng.randn(min(odd, n)) * np.pi
  node_vals = {int(k): np.float64(np.float32(rng.rand() + 0.5) * root / num_nodes)
               for k in range(num_nodes)}
  return node_vals, nodes, deltas


@pytest.mark.parametrize('n,threads', [(4096, 1024), (3000, 1024), (1024, 256), (100, 256)])
@pytest.mark.parametrize('num_nodes', [1, 2, 7, 64])
def test_exact_data_is_accepted_at_once_and_is_the_sequential_chain(n, threads, num_nodes):
  rng = np.random.RandomState(n + num_nodes)
  node_vals, nodes, deltas = _case(rng, n, num_nodes, odd=0)
  ok, rounds, got = verified_scan(node_vals, nodes, deltas, threads)
  assert ok and rounds == 1
  want = sequential_chains(node_vals, nodes, deltas)
  assert all(np.float64(got[k]).tobytes() == np.float64(want[k]).tobytes() for k in want)


@pytest.mark.parametrize('odd', [1, 2, 3, 9, 200, 4096])
@pytest.mark.parametrize('num_nodes', [1, 3, 32])
def test_accepted_always_means_bit_exact(odd, num_nodes):
  """Adds that round: the model may need more rounds or decline, but it never accepts
  anything other than the sequential result."""
  accepted = declined = 0
  for seed in range(12):
    rng = np.random.RandomState(1000 * odd + 10 * num_nodes + seed)
    node_vals, nodes, deltas = _case(rng, 4096, num_nodes, odd=odd)
    ok, _, got = verified_scan(node_vals, nodes, deltas, 1024)
    if not ok:
      declined += 1
      continue
    accepted += 1
    want = sequential_chains(node_vals, nodes, deltas)
    assert all(np.float64(got[k]).tobytes() == np.float64(want[k]).tobytes()
               for k in want)
  if odd >= 200:
    assert declined > 0      # most adds round: the serial chains take over
  if odd <= 2 and num_nodes > 1:
    assert accepted > 0      # a couple of rounding adds are repaired by re-rooting
