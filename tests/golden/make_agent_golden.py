"""Generates tests/golden/agent_traces.npz from the UNMODIFIED reference agents
(agents/rule_based_agent.py, agents/simple_agent.py) playing in the reference's own
rl_env.HanabiEnv over the reference library (oracle/_ref/libpyhanabi.so).

Run in the build container only:  make -C oracle ref && python tests/golden/make_agent_golden.py

`gin` is not installed here; the agent module only uses it as a decorator, so a
stub module stands in for it.  Each trace: a team of seats (rule-based agents --
possibly sharing one object like create_adhoc_team does,
training/run_adhoc_experiment.py:251-258 -- and scripted seats), several episodes
on one seeded environment, agent objects never reset between episodes (as in
run_one_episode, :394-469).  Recorded per step: player to act, move uid, done.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_HLE = os.path.join(REF, "hanabi_learning_environment")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, REF_HLE)

gin = types.ModuleType("gin")
gin.configurable = lambda x: x
gin.tf = types.ModuleType("gin.tf")
sys.modules["gin"] = gin
sys.modules["gin.tf"] = gin.tf

import pyhanabi  # noqa: E402  (the reference's binding)
assert pyhanabi.try_cdef(prefixes=[REF_HLE])
assert pyhanabi.try_load(prefixes=[os.path.join(ROOT, "oracle", "_ref")])
import rl_env  # noqa: E402
sys.path.insert(0, os.path.join(REF, "agents"))
import rule_based_agent  # noqa: E402
import simple_agent  # noqa: E402

FULL = dict(colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3, observation_type=1)
# name -> (players, seat kinds: 'R' rule-based / 'S' simple / '-' scripted, slot per seat)
TEAMS = {
    "adhoc4_shared": (4, "-RRR", [0, 0, 0, 0]),   # config 5: seat 0 = learner, 1-3 share one object
    "rule2_separate": (2, "RR", [0, 1]),
    "rule3_mixed": (3, "R-R", [0, 0, 1]),
    "rule5_shared": (5, "RRRRR", [0, 0, 0, 0, 0]),
    "simple2": (2, "SS", [0, 0]),
    "simple4_mixed": (4, "S-SS", [0, 0, 0, 0]),
}


def uid_of(env, obs, action):
  if not isinstance(action, dict):
    return int(action)
  for idx, legal in enumerate(obs["legal_moves"]):
    if legal == action:
      return int(obs["legal_moves_as_int"][idx])
  return int(obs["legal_moves_as_int"][-1])


def run(name, seed, steps):
  players, kinds, slots = TEAMS[name]
  cfg = dict(FULL, players=players, seed=seed)
  env = rl_env.HanabiEnv(cfg)
  objs = {}
  agents = []
  for seat, k in enumerate(kinds):
    if k == "R":
      key = ("R", slots[seat])
      if key not in objs:
        objs[key] = rule_based_agent.RuleBasedAgent(players=players)
      agents.append(objs[key])
    elif k == "S":
      agents.append(simple_agent.SimpleAgent(cfg))
    else:
      agents.append(None)
  rec = dict(cur=[], action=[], done=[])
  observations = env.reset()
  for t in range(steps):
    cur = observations["current_player"]
    po = observations["player_observations"][cur]
    if kinds[cur] == "R":
      a = uid_of(env, po, agents[cur].act_train(po))
    elif kinds[cur] == "S":
      a = uid_of(env, po, agents[cur].act(po))
    else:
      legal = po["legal_moves_as_int"]
      a = int(legal[(t * 7 + 3) % len(legal)])
    observations, reward, done, _ = env.step(a)
    rec["cur"].append(cur)
    rec["action"].append(a)
    rec["done"].append(int(done))
    if done:
      observations = env.reset()
  return {k: np.asarray(v, np.int32) for k, v in rec.items()}


def main():
  out = {}
  for name in TEAMS:
    for seed in (11, 12):
      rec = run(name, seed, 400)
      for k, v in rec.items():
        out["%s/%d/%s" % (name, seed, k)] = v
      print(name, seed, "episodes", int(rec["done"].sum()))
  np.savez_compressed(os.path.join(HERE, "agent_traces.npz"), **out)


if __name__ == "__main__":
  main()
