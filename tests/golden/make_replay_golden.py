"""Records tests/golden/replay_golden.npz from the reference's own replay classes
(agents/rainbow_copy/replay_memory.py, prioritized_replay_memory.py, sum_tree.py).
tensorflow and gin are not installed; both are only used for the in-graph wrapper
and decorators, so empty stub modules stand in for them.  Build container only."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
RB = "/root/reference/agents/rainbow_copy"


def load_reference_replay():
  gin = types.ModuleType("gin")
  gin.configurable = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
  gin.tf = types.ModuleType("gin.tf")
  tf = types.ModuleType("tensorflow")
  sys.modules.setdefault("gin", gin)
  sys.modules.setdefault("gin.tf", gin.tf)
  sys.modules.setdefault("tensorflow", tf)
  sys.path.insert(0, RB)
  sys.path.insert(0, os.path.join(RB, "third_party", "dopamine"))
  third = types.ModuleType("third_party")
  third.__path__ = [os.path.join(RB, "third_party")]
  sys.modules.setdefault("third_party", third)
  import prioritized_replay_memory as prm  # noqa: E402
  import replay_memory as rm  # noqa: E402
  return rm, prm


def main():
  from oracle import replay_oracle as R
  rm, prm = load_reference_replay()
  rng = np.random.default_rng(0)
  out = {}
  for name, (cap, n, gamma, count) in {"wrap_n3": (50, 3, 0.99, 130), "fill_n1": (64, 1, 1.0, 40),
                                       "wrap_n5": (37, 5, 0.9, 200)}.items():
    D, A = 12, 7
    mem = prm.OutOfGraphPrioritizedReplayMemory(A, D, 1, cap, 8, update_horizon=n, gamma=gamma)
    obs, act, rew, term, legal = R.random_stream(rng, count, D, A)
    for i in range(count):
      mem.add(obs[i], act[i], rew[i], term[i], legal[i])
    pri_idx = rng.integers(0, cap, 25).astype(np.int32)
    pri_val = rng.random(25).astype(np.float32) * 10
    mem.set_priority(pri_idx, pri_val)
    valid = np.asarray([mem.is_valid_transition(i) for i in range(-1, cap + 1)], np.uint8)
    idx = [i for i in range(cap) if mem.is_valid_transition(i)]
    batch = mem.sample_transition_batch(batch_size=len(idx), indices=idx)
    queries = rng.random(64)
    picks = np.asarray([mem.sum_tree.sample(query_value=q) for q in queries], np.int32)
    leaves = np.asarray([mem.sum_tree.get(i) for i in range(cap)])
    out.update({name + "/cfg": np.asarray([cap, n, count, D, A]), name + "/gamma": np.asarray(gamma),
                name + "/obs": obs, name + "/act": act, name + "/rew": rew, name + "/term": term,
                name + "/legal": legal, name + "/pri_idx": pri_idx, name + "/pri_val": pri_val,
                name + "/valid": valid, name + "/idx": np.asarray(idx, np.int32),
                name + "/state": batch[0][:, :, 0], name + "/action": batch[1], name + "/reward": batch[2],
                name + "/next_state": batch[3][:, :, 0], name + "/terminal": batch[4],
                name + "/next_legal": batch[6], name + "/queries": queries, name + "/picks": picks,
                name + "/leaves": leaves, name + "/total": np.asarray(mem.sum_tree._total_priority())})
    print(name, "valid", int(valid.sum()), "terminals in batch", int(batch[4].sum()))
  np.savez_compressed(os.path.join(HERE, "replay_golden.npz"), **out)


if __name__ == "__main__":
  main()
