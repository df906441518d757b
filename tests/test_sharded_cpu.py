"""N>1 host path on CPU: world_size-2 gloo run of the shard-total exchange, and the
apportioning rule every rank evaluates on the gathered totals."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sharded_port
from oracle.sumtree_port import PortSumTree


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, out_dir):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from dopamine_b200.replay_memory import sharded_replay
  rng = np.random.RandomState(rank)
  tree = PortSumTree(64)
  for i in range(40):
    tree.set(i, float(np.float32(abs(rng.randn()) + 0.1 * rank)))
  local = torch.tensor([tree.total()], dtype=torch.float64)
  totals = sharded_replay.all_gather_totals(local)
  assert totals.shape == (world,) and totals.dtype == torch.float64
  assert float(totals[rank]) == float(tree.total())
  # every rank derives the same ownership from the same totals + shared uniforms
  shared = np.random.RandomState(123).rand(32)
  bounds = np.linspace(0., 1., 33)
  queries = bounds[:-1] + (bounds[1:] - bounds[:-1]) * shared
  owners = sharded_port.apportion(totals.tolist(), queries)
  mine = [i for i, (o, _) in enumerate(owners) if o == rank]
  picks = [tree.descend(owners[i][1]) for i in mine]
  np.save(os.path.join(out_dir, 'rank%d.npy' % rank),
          np.array([mine, picks], dtype=np.int64))
  gidx = sharded_replay.global_index(rank, 64, np.array(picks, dtype=np.int64))
  assert ((gidx // 64) == rank).all()
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_gloo_exchange_partitions_the_batch(tmp_path):
  world = 2
  mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world,
           join=True)
  seen = []
  for r in range(world):
    mine, picks = np.load(str(tmp_path / ('rank%d.npy' % r)))
    assert (picks >= 0).all() and (picks < 64).all()
    seen += mine.tolist()
  assert sorted(seen) == list(range(32))  # every stratum served exactly once


def test_apportion_matches_one_big_tree():
  """Hanging G shard trees under one extra level == one tree over all leaves, when
  the shard totals are exact sums (powers of two, so no rounding is involved)."""
  g, cap = 4, 16
  rng = np.random.RandomState(0)
  shards = [PortSumTree(cap) for _ in range(g)]
  big = PortSumTree(g * cap)
  for s in range(g):
    for i in range(cap):
      v = float(2.0 ** rng.randint(-3, 4))
      shards[s].set(i, v)
      big.set(s * cap + i, v)
  queries = rng.rand(500)
  owners = sharded_port.apportion([t.total() for t in shards], queries)
  for q, (o, mass) in zip(queries, owners):
    assert o * cap + shards[o].descend(mass) == big.descend(q * big.total())


def test_philox_port_known_answers():
  """oracle/philox_port.py against the Random123 known-answer vectors of
  philox4x32_10 (kat_vectors: zero and all-ones counter/key, and the pi digits)."""
  from oracle import philox_port as p
  assert p.philox4x32_10((0, 0, 0, 0), (0, 0)) == (
      0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
  assert p.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (
      0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
  assert p.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
                         (0xa4093822, 0x299f31d0)) == (
      0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
  u = p.uniform53(1, 2, 3)
  assert 0.0 <= u < 1.0
  q = p.stratified_queries(7, 0, 16)
  assert all(i / 16 <= q[i] < (i + 1) / 16 for i in range(16))
