"""Replay memory: the numpy restatement (oracle/replay_oracle.py) against vectors
recorded from the reference's own classes, and the device replay (hb_replay.cu)
against the restatement."""
import os

import numpy as np
import pytest

from oracle import replay_oracle as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_golden.npz")
CASES = ["wrap_n3", "fill_n1", "wrap_n5"]


@pytest.fixture(scope="module")
def golden():
  return np.load(GOLDEN)


def _fill_oracle(g, name):
  cap, n, count, D, A = [int(x) for x in g[name + "/cfg"]]
  mem = R.ReplayOracle(A, D, cap, update_horizon=n, gamma=float(g[name + "/gamma"]))
  for i in range(count):
    mem.add(g[name + "/obs"][i], g[name + "/act"][i], g[name + "/rew"][i], g[name + "/term"][i],
            g[name + "/legal"][i])
  for i, v in zip(g[name + "/pri_idx"], g[name + "/pri_val"]):
    mem.tree.set(int(i), float(v))
  return mem, (cap, n, count, D, A)


@pytest.mark.parametrize("name", CASES)
def test_restatement_reproduces_reference_replay_vectors(golden, name):
  g = golden
  mem, (cap, n, count, D, A) = _fill_oracle(g, name)
  valid = np.asarray([mem.is_valid_transition(i) for i in range(-1, cap + 1)], np.uint8)
  assert (valid == g[name + "/valid"]).all()
  idx = g[name + "/idx"]
  batch = mem.sample_transition_batch(list(idx))
  assert (batch[0] == g[name + "/state"]).all() and (batch[1] == g[name + "/action"]).all()
  np.testing.assert_allclose(batch[2], g[name + "/reward"], rtol=1e-6, atol=1e-6)
  assert (batch[3] == g[name + "/next_state"]).all() and (batch[4] == g[name + "/terminal"]).all()
  assert (batch[6] == g[name + "/next_legal"]).all()
  assert [mem.tree.sample(q) for q in g[name + "/queries"]] == g[name + "/picks"].tolist()
  np.testing.assert_array_equal(np.asarray([mem.tree.get(i) for i in range(cap)]), g[name + "/leaves"])
  assert mem.tree.total() == pytest.approx(float(g[name + "/total"]), rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_replay_reproduces_reference_replay_vectors(golden, name):
  import torch
  from hanabi_b200 import replay
  g = golden
  cap, n, count, D, A = [int(x) for x in g[name + "/cfg"]]
  dev = torch.device("cuda:0")
  t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
  mem = replay.DeviceReplayMemory(A, D, cap, update_horizon=n, gamma=float(g[name + "/gamma"]))
  # appended in several batches of uneven size, wrapping around the ring
  edges = [0, 1, 9, count // 2, count]
  for lo, hi in zip(edges, edges[1:]):
    mem.add(t(g[name + "/obs"][lo:hi]), t(g[name + "/act"][lo:hi]), t(g[name + "/rew"][lo:hi]),
            t(g[name + "/term"][lo:hi]), t(g[name + "/legal"][lo:hi]))
  assert mem.add_count == count and mem.cursor() == count % cap
  mem.set_priority(t(g[name + "/pri_idx"]), t(g[name + "/pri_val"]))  # contains duplicate indices
  probe = t(np.arange(-1, cap + 1, dtype=np.int32))
  assert (mem.is_valid_transition(probe).cpu().numpy() == g[name + "/valid"]).all()
  idx = t(g[name + "/idx"])
  state, action, reward, next_state, terminal, indices, next_legal = mem.sample_transition_batch(indices=idx)
  assert (state.cpu().numpy() == g[name + "/state"]).all()
  assert (action.cpu().numpy() == g[name + "/action"]).all()
  np.testing.assert_allclose(reward.cpu().numpy(), g[name + "/reward"], rtol=1e-6, atol=1e-6)
  assert (next_state.cpu().numpy() == g[name + "/next_state"]).all()
  assert (terminal.cpu().numpy() == g[name + "/terminal"]).all()
  assert (next_legal.cpu().numpy() == g[name + "/next_legal"]).all()
  leaves = mem.get_priority(t(np.arange(cap, dtype=np.int32))).cpu().numpy()
  np.testing.assert_array_equal(leaves, g[name + "/leaves"].astype(np.float32))
  picks = mem.tree_sample(t(g[name + "/queries"])).cpu().numpy()
  assert (picks == g[name + "/picks"]).all()


@pytest.mark.gpu
def test_device_replay_against_restatement_at_size():
  """Larger ring, many wraps, batched adds and priority updates; sampling statistics."""
  import torch
  from hanabi_b200 import replay
  rng = np.random.default_rng(1)
  cap, n, D, A = 5000, 5, 658, 20
  dev = torch.device("cuda:0")
  t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
  ora = R.ReplayOracle(A, D, cap, update_horizon=n, gamma=0.99)
  mem = replay.DeviceReplayMemory(A, D, cap, update_horizon=n, gamma=0.99, seed=3)
  for batch in (700, 3000, 2500, 1234):
    obs, act, rew, term, legal = R.random_stream(rng, batch, D, A, terminal_rate=0.05)
    for i in range(batch):
      ora.add(obs[i], act[i], rew[i], term[i], legal[i])
    mem.add(t(obs), t(act), t(rew), t(term), t(legal))
    pidx = rng.integers(0, cap, 256).astype(np.int32)
    pval = (rng.random(256) * 5).astype(np.float32)
    for i, v in zip(pidx, pval):
      ora.tree.set(int(i), float(v))
    mem.set_priority(t(pidx), t(pval))
    valid = np.asarray([ora.is_valid_transition(i) for i in range(cap)], np.uint8)
    assert (mem.is_valid_transition(t(np.arange(cap, dtype=np.int32))).cpu().numpy() == valid).all()
    idx = np.flatnonzero(valid)[::7].astype(np.int32)
    want = ora.sample_transition_batch(list(idx))
    got = mem.sample_transition_batch(indices=t(idx))
    for k in (0, 1, 3, 4, 6):
      assert (got[k].cpu().numpy() == want[k]).all(), k
    np.testing.assert_allclose(got[2].cpu().numpy(), want[2], rtol=1e-6, atol=1e-6)
    q = rng.random(4096)
    assert (mem.tree_sample(t(q)).cpu().numpy() == np.asarray([ora.tree.sample(x) for x in q])).all()
  # prioritized sampling: only valid indices, frequencies proportional to priority
  picks = mem.sample_index_batch(200000).cpu().numpy()
  assert (picks >= 0).all() and valid[picks].all()
  leaves = np.asarray([ora.tree.get(i) for i in range(cap)]) * valid
  freq = np.bincount(picks, minlength=cap) / len(picks)
  heavy = np.argsort(leaves)[-50:]
  np.testing.assert_allclose(freq[heavy].sum(), (leaves / leaves.sum())[heavy].sum(), rtol=0.05)
  strat = mem.sample_index_batch(4096, stratified=True).cpu().numpy()
  assert (strat >= 0).all() and valid[strat].all()
  # uniform memory
  uni = replay.DeviceReplayMemory(A, D, 300, update_horizon=2, gamma=1.0, prioritized=False)
  obs, act, rew, term, legal = R.random_stream(rng, 120, D, A)
  uni.add(t(obs), t(act), t(rew), t(term), t(legal))
  u = uni.sample_index_batch(5000).cpu().numpy()
  assert u.min() >= 0 and u.max() < 120 - 2


@pytest.mark.gpu
def test_device_replay_failed_draws_raise_and_never_read_out_of_bounds():
  """sampling before any transition is valid, or from a tree whose total is
  zero, must raise like the reference (replay_memory.py:268,
  prioritized_replay_memory.py:128) -- and an index of -1 handed to the gather must not read
  outside the ring."""
  import torch
  from hanabi_b200 import replay
  dev = torch.device("cuda:0")
  A, D, cap, n = 5, 12, 64, 3
  mem = replay.DeviceReplayMemory(A, D, cap, update_horizon=n, gamma=0.99)
  t = lambda a, dt: torch.as_tensor(np.asarray(a), dtype=dt, device=dev)
  with pytest.raises(RuntimeError):
    mem.sample_index_batch(4)  # empty
  for k in range(n):  # still no valid transition: needs more than update_horizon adds
    mem.add(t(np.full((1, D), k), torch.uint8), t([1], torch.int32), t([1.0], torch.float32),
            t([0], torch.uint8), t(np.zeros((1, A)), torch.float32))
  with pytest.raises(RuntimeError):
    mem.sample_index_batch(4)
  for k in range(20):
    mem.add(t(np.full((1, D), 10 + k), torch.uint8), t([2], torch.int32), t([0.5], torch.float32),
            t([0], torch.uint8), t(np.ones((1, A)), torch.float32))
  idx = mem.sample_index_batch(16)
  assert int(idx.min()) >= 0
  # every priority zero: the tree cannot produce a valid index -> raises
  every = torch.arange(cap, dtype=torch.int32, device=dev)
  mem.set_priority(every, torch.zeros(cap, dtype=torch.float32, device=dev))
  if mem.get_priority(every).sum().item() == 0:
    idx0 = mem.sample_index_batch(8, check=False)
    if int(idx0.min()) < 0:
      assert mem.sample_failures(clear=False) > 0
      with pytest.raises(RuntimeError):
        mem.sample_index_batch(8)
  # the gather with failed / foreign indices: zero rows flagged terminal, no fault
  bad = torch.tensor([-1, 0, cap + 5, 3], dtype=torch.int32, device=dev)
  state, action, reward, nxt, term, _, legal = mem.sample_transition_batch(indices=bad)
  torch.cuda.synchronize()
  assert int(state[0].sum()) == 0 and int(state[2].sum()) == 0 and int(term[0]) == 1 and int(term[2]) == 1
  assert float(mem.get_priority(bad)[0]) == 0.0
