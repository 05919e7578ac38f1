"""World-size-2 CPU test (gloo) of the multi-GPU host logic: shards cover the
game range exactly, per-rank seeds equal the global ones, and the all-reduced
episode statistics equal a single-process run over all games."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hanabi_b200 import sharding
from oracle import oracle as O
from tests.common import CONFIGS

TOTAL, STEPS, BASE_SEED = 37, 40, 11


def play_and_count(seeds):
  """Random-legal rollout on the CPU checker; returns the 32 statistics counters."""
  n = len(seeds)
  b = O.CpuBatch("oracle", CONFIGS["small2"], n, np.asarray(seeds, np.int32))
  b.reset()
  stats = np.zeros(sharding.STATS_LEN, np.int64)
  lengths = np.zeros(n, np.int64)
  rng = np.random.default_rng(int(seeds[0]))
  for _ in range(STEPS):
    mask = b.legal_mask()
    actions = (rng.random(mask.shape) * mask).argmax(axis=1).astype(np.int32)
    _, done, score, _ = b.step(actions)
    lengths += 1
    d = done != 0
    stats[0] += d.sum()
    stats[1] += score[d].sum()
    stats[2] += lengths[d].sum()
    np.add.at(stats, 3 + score[d], 1)
    lengths[d] = 0
  return stats


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                    LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    assert sharding.world_from_env() == (rank, rank, world)
    lo, hi = sharding.shard_range(TOTAL, rank, world)
    seeds = sharding.shard_seeds(BASE_SEED, TOTAL, rank, world)
    assert seeds == [BASE_SEED + i for i in range(lo, hi)]
    # each rank plays its own slice with a per-game policy stream so that the
    # result does not depend on how the games are grouped
    stats = np.zeros(sharding.STATS_LEN, np.int64)
    for s in seeds:
      stats += play_and_count([s])
    t = torch.from_numpy(stats.copy())
    sharding.reduce_episode_statistics(t)
    out.put((rank, lo, hi, t.tolist()))
    dist.barrier()
  finally:
    dist.destroy_process_group()


def test_shard_ranges_partition_the_games():
  for total in (1, 7, 37, 1 << 20):
    for world in (1, 2, 3, 8):
      edges = [sharding.shard_range(total, r, world) for r in range(world)]
      assert edges[0][0] == 0 and edges[-1][1] == total
      for (a, b), (c, d) in zip(edges, edges[1:]):
        assert b == c
      sizes = [b - a for a, b in edges]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    sharding.shard_range(8, 2, 2)


def test_two_rank_gloo_statistics_match_single_process():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  results = [out.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  results.sort()
  assert results[0][1] == 0 and results[0][2] == results[1][1] and results[1][2] == TOTAL
  want = np.zeros(sharding.STATS_LEN, np.int64)
  for i in range(TOTAL):
    want += play_and_count([BASE_SEED + i])
  assert results[0][3] == want.tolist() and results[1][3] == want.tolist()
  assert want[0] > 0
  d = sharding.describe_statistics(torch.from_numpy(want))
  assert d["episodes"] == want[0] and sum(d["score_hist"]) == want[0]
