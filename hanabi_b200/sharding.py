"""Multi-GPU layout of a batch of games: one process per GPU, games sharded by
contiguous index range, no collective on the step path.

Games never interact (a reference HanabiState touches only its own fields and
its own game's RNG, hanabi_lib/hanabi_state.h:131-142, hanabi_lib/hanabi_game.h:114),
so rank r simply owns games [lo, hi) with seeds base_seed + global index; the
only exchange is a sum all-reduce of the 32 int64 episode-statistics counters
(256 bytes) when somebody wants a report -- NCCL over NVLink for CUDA tensors,
gloo for the CPU tests.
"""
import os

STATS_LEN = 32  # [0] episodes [1] sum score [2] sum length [3..28] score histogram 0..25


def world_from_env():
  """(rank, local_rank, world_size) as torchrun exports them."""
  return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
          int(os.environ.get("WORLD_SIZE", "1")))


def _parse_cpulist(text):
  cpus = set()
  for part in text.strip().split(","):
    if not part:
      continue
    lo, _, hi = part.partition("-")
    cpus.update(range(int(lo), int(hi or lo) + 1))
  return cpus


def bind_rank_to_local_cores(local_rank, local_world):
  """Pins this process (and the host worker pool it creates later) to its share of the
  cores next to its GPU: the GPU's NUMA-local CPU list from sysfs, split evenly among the
  local ranks whose GPUs report the same list.  Pinned host buffers allocated afterwards
  are first touched on that node.  Host-side policy only; returns the chosen CPU set, or
  None when the topology cannot be read (nothing is changed then)."""
  try:
    import torch
    def cpus_of(i):
      pr = torch.cuda.get_device_properties(i)
      name = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
      with open("/sys/bus/pci/devices/%s/local_cpulist" % name) as f:
        return frozenset(_parse_cpulist(f.read()))
    allowed = os.sched_getaffinity(0)
    lists = [cpus_of(i) for i in range(local_world)]
    mine = lists[local_rank] & allowed
    if not mine:
      return None
    peers = [i for i in range(local_world) if lists[i] == lists[local_rank]]
    per = len(mine) // len(peers)
    if per < 1:
      return None
    k = peers.index(local_rank)
    chosen = set(sorted(mine)[k * per:(k + 1) * per])
    os.sched_setaffinity(0, chosen)
    return chosen
  except Exception:  # pylint: disable=broad-except
    return None


def shard_range(total_games, rank, world_size):
  """Contiguous, balanced slice [lo, hi) of the global game index range."""
  if not 0 <= rank < world_size:
    raise ValueError("rank %d outside world of %d" % (rank, world_size))
  base, extra = divmod(int(total_games), int(world_size))
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def shard_seeds(base_seed, total_games, rank, world_size):
  """Seeds of this rank's games: seed_i = base_seed + global index i, wrapped to
  int32 like the reference's std::stoi'd seed (hanabi_lib/util.cc:37-47)."""
  lo, hi = shard_range(total_games, rank, world_size)
  return [((base_seed + i + 2**31) % 2**32) - 2**31 for i in range(lo, hi)]


def reduce_episode_statistics(stats, group=None):
  """Sum all-reduce of a [32] int64 statistics tensor across ranks (in place).
  No-op without an initialised process group."""
  import torch.distributed as dist
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
  return stats


def describe_statistics(stats):
  vals = [int(v) for v in stats.tolist()]
  episodes = max(vals[0], 1)
  return {"episodes": vals[0], "mean_score": vals[1] / episodes, "mean_length": vals[2] / episodes,
          "score_hist": vals[3:29]}


def make_sharded_env(environment_name="Hanabi-Full", num_players=2, total_games=1 << 20,
                     base_seed=1, auto_reset=True):
  """This rank's slice of a `total_games` batch as a BatchedHanabiEnv on its GPU,
  with the statistics buffer ready for reduce_episode_statistics()."""
  import torch
  from hanabi_b200 import rl_env
  rank, local_rank, world = world_from_env()
  lo, hi = shard_range(total_games, rank, world)
  config = rl_env._preset(environment_name, num_players)  # pylint: disable=protected-access
  env = rl_env.BatchedHanabiEnv(config, hi - lo, device=local_rank,
                                seeds=shard_seeds(base_seed, total_games, rank, world),
                                auto_reset=auto_reset)
  env.stats = torch.zeros(STATS_LEN, dtype=torch.int64, device=env.device)
  env.batch.set_stats_buffer(env.stats)
  env.game_range = (lo, hi)
  return env
