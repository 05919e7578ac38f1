"""The reference's hand-written policies: CPU restatement (oracle/agents_oracle.py)
and the CUDA batch kernel (BatchAgentAct) against traces recorded from the
unmodified reference agents (tests/golden/agent_traces.npz)."""
import os

import numpy as np
import pytest

from oracle import agents_oracle as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = dict(colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3, observation_type=1)
TEAMS = {  # must match tests/golden/make_agent_golden.py
    "adhoc4_shared": (4, "-RRR", [0, 0, 0, 0]),
    "rule2_separate": (2, "RR", [0, 1]),
    "rule3_mixed": (3, "R-R", [0, 0, 1]),
    "rule5_shared": (5, "RRRRR", [0, 0, 0, 0, 0]),
    "simple2": (2, "SS", [0, 0]),
    "simple4_mixed": (4, "S-SS", [0, 0, 0, 0]),
}


@pytest.fixture(scope="module")
def traces():
  return np.load(os.path.join(GOLDEN, "agent_traces.npz"))


@pytest.mark.parametrize("name", sorted(TEAMS))
@pytest.mark.parametrize("seed", [11, 12])
def test_restated_agents_reproduce_reference_agent_traces(built_lib, traces, name, seed):
  from hanabi_b200 import rl_env
  players, kinds, slots = TEAMS[name]
  cfg = dict(FULL, players=players, seed=seed)
  env = rl_env.HanabiEnv(cfg)
  objs = {}
  agents = []
  for seat, k in enumerate(kinds):
    if k == "R":
      agents.append(objs.setdefault(slots[seat], A.RuleBasedAgent(players)))
    elif k == "S":
      agents.append(A.SimpleAgent(cfg))
    else:
      agents.append(None)
  key = "%s/%d/" % (name, seed)
  cur, action, done = traces[key + "cur"], traces[key + "action"], traces[key + "done"]
  obs = env.reset()
  for t in range(len(action)):
    assert obs["current_player"] == cur[t]
    if agents[cur[t]] is not None:
      got = agents[cur[t]].act_uid(obs["player_observations"][cur[t]])
      assert got == action[t], (name, seed, t, got, int(action[t]))
    obs, _, d, _ = env.step(int(action[t]))
    assert bool(d) == bool(done[t])
    if d:
      obs = env.reset()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(TEAMS))
def test_gpu_agents_reproduce_reference_agent_traces(traces, name):
  """One batch holds both recorded games (seeds 11, 12) of a team."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  players, kinds, slots = TEAMS[name]
  seeds = [11, 12]
  n = len(seeds)
  b = ph.HanabiBatch(dict(FULL, players=players), n, seeds)
  b.enable_history()
  b.reset()
  dev = b.device
  num_slots = max(slots) + 1
  state = torch.zeros((n, num_slots), dtype=torch.int32, device=dev)
  act = torch.zeros(n, dtype=torch.int32, device=dev)
  cur_t = torch.zeros(n, dtype=torch.int32, device=dev)
  done_t = torch.zeros(n, dtype=torch.uint8, device=dev)
  rew_t = torch.zeros(n, dtype=torch.int32, device=dev)
  rule_mask = sum(1 << s for s, k in enumerate(kinds) if k == "R")
  simple_mask = sum(1 << s for s, k in enumerate(kinds) if k == "S")
  tr = [{k: traces["%s/%d/%s" % (name, s, k)] for k in ("cur", "action", "done")} for s in seeds]
  steps = len(tr[0]["action"])
  for t in range(steps):
    b.cur_player(cur_t)
    cur = cur_t.cpu().numpy()
    act.fill_(-7)
    if rule_mask:
      b.agent_act(0, rule_mask, act, agent_state=state, seat_slot=slots)
    if simple_mask:
      b.agent_act(1, simple_mask, act)
    got = act.cpu().numpy()
    want = np.asarray([tr[i]["action"][t] for i in range(n)], np.int32)
    for i in range(n):
      assert cur[i] == tr[i]["cur"][t], (name, t)
      if kinds[cur[i]] == "-":
        assert got[i] == -7  # scripted seat: the kernel leaves the entry alone
      else:
        assert got[i] == want[i], (name, seeds[i], t, int(got[i]), int(want[i]))
    act.copy_(torch.as_tensor(want))
    b.step(act, rew_t, done_t, auto_reset=True)
    d = done_t.cpu().numpy()
    for i in range(n):
      assert bool(d[i]) == bool(tr[i]["done"][t])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["adhoc4_shared", "rule2_separate", "simple4_mixed"])
def test_gpu_agents_lockstep_with_restated_agents(name):
  """More games than the golden traces hold: every seat's kernel action equals the
  (reference-pinned) Python restatement playing the same seeded games on the host."""
  import torch
  from hanabi_b200 import pyhanabi as ph, rl_env
  players, kinds, slots = TEAMS[name]
  n, steps = 24, 120
  seeds = [500 + i for i in range(n)]
  b = ph.HanabiBatch(dict(FULL, players=players), n, seeds)
  b.enable_history()
  b.reset()
  dev = b.device
  num_slots = max(slots) + 1
  state = torch.zeros((n, num_slots), dtype=torch.int32, device=dev)
  act = torch.zeros(n, dtype=torch.int32, device=dev)
  done_t = torch.zeros(n, dtype=torch.uint8, device=dev)
  rew_t = torch.zeros(n, dtype=torch.int32, device=dev)
  rule_mask = sum(1 << s for s, k in enumerate(kinds) if k == "R")
  simple_mask = sum(1 << s for s, k in enumerate(kinds) if k == "S")
  envs, teams, obs = [], [], []
  for s in seeds:
    cfg = dict(FULL, players=players, seed=s)
    e = rl_env.HanabiEnv(cfg)
    objs = {}
    team = []
    for seat, k in enumerate(kinds):
      team.append(objs.setdefault(slots[seat], A.RuleBasedAgent(players)) if k == "R"
                  else A.SimpleAgent(cfg) if k == "S" else None)
    envs.append(e); teams.append(team); obs.append(e.reset())
  for t in range(steps):
    want = np.zeros(n, np.int32)
    for i in range(n):
      cur = obs[i]["current_player"]
      po = obs[i]["player_observations"][cur]
      if teams[i][cur] is None:
        legal = po["legal_moves_as_int"]
        want[i] = legal[(t * 5 + i) % len(legal)]
      else:
        want[i] = teams[i][cur].act_uid(po)
    act.copy_(torch.as_tensor(want))  # scripted seats keep this entry
    if rule_mask:
      b.agent_act(0, rule_mask, act, agent_state=state, seat_slot=slots)
    if simple_mask:
      b.agent_act(1, simple_mask, act)
    got = act.cpu().numpy()
    assert (got == want).all(), (name, t, np.flatnonzero(got != want)[:4])
    b.step(act, rew_t, done_t, auto_reset=True)
    d = done_t.cpu().numpy()
    for i in range(n):
      obs[i], _, done, _ = envs[i].step(int(want[i]))
      assert bool(d[i]) == done
      if done:
        obs[i] = envs[i].reset()
