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
