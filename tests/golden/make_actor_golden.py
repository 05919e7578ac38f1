"""Generates tests/golden/actor_golden.npz from the UNMODIFIED reference code of the three
rows that round 1 had pinned "by reading only":

  * a11  `linearly_decaying_epsilon`              agents/rainbow_copy/dqn_agent.py:44-59
         `DQNAgent._select_action`'s greedy branch semantics (first maximal legal index of
         q + legal, legal in {0, -inf})            dqn_agent.py:200, 387-420
  * f2   the agent's transition bookkeeping        dqn_agent.py:286-385
         driven by the reference's own episode loop  agents/run_experiment.py:344-403
         (`run_one_episode`, `parse_observations`, `ObservationStacker`)
  * f4   `PyhanabiEnvWrapper._reset/_step`          training/tf_agents_lib/pyhanabi_env_wrapper.py:41-90

Run in the build container only:  make -C oracle ref && python tests/golden/make_actor_golden.py

TensorFlow, gin and tf_agents are not installed.  What the three pieces need from them is
nothing numerical: decorators, a base class whose reset()/step() call _reset()/_step(), a
namedtuple and three integer constants.  Permissive stub modules stand in for them; the
reference files are imported (dqn_agent.py, pyhanabi_env_wrapper.py) or -- run_experiment.py
pulls in half the repository at import time -- the needed functions are compiled from the
file where it lies, unmodified, into a scratch namespace.  The DQNAgent object is created
without running its constructor (which builds a TF graph); only its network-free methods
run: begin_episode / step / end_episode / _record_transition / _post_transitions.  Its three
network hooks are replaced: _train_step (no-op), _select_action (a scripted legal move,
recorded), _store_transition (appends to the recorded stream).
"""
import ast
import collections
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_HLE = os.path.join(REF, "hanabi_learning_environment")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_HLE)

from tests.common import CONFIGS  # noqa: E402


class _Anything(object):
  """Stands in for every TensorFlow symbol touched at import / class-definition time."""

  def __getattr__(self, name):
    return _Anything()

  def __call__(self, *a, **k):
    return _Anything()


def install_stubs():
  tf = types.ModuleType("tensorflow")
  tf.__getattr__ = lambda name: _Anything()
  gin = types.ModuleType("gin")
  gin.configurable = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
  gin.tf = types.ModuleType("gin.tf")
  sys.modules["tensorflow"] = tf
  sys.modules["gin"] = gin
  sys.modules["gin.tf"] = gin.tf
  # tf_agents: the protocol only (py_environment.PyEnvironment.reset/step call _reset/_step)
  ta = types.ModuleType("tf_agents")
  env = types.ModuleType("tf_agents.environments")
  wrappers = types.ModuleType("tf_agents.environments.wrappers")

  class PyEnvironmentBaseWrapper(object):
    def __init__(self, *a, **k):
      pass

    def reset(self):
      return self._reset()

    def step(self, action):
      return self._step(action)

  wrappers.PyEnvironmentBaseWrapper = PyEnvironmentBaseWrapper
  traj = types.ModuleType("tf_agents.trajectories")
  ts = types.ModuleType("tf_agents.trajectories.time_step")
  ts.TimeStep = collections.namedtuple("TimeStep", ["step_type", "reward", "discount", "observation"])

  class StepType(object):  # tf_agents.trajectories.time_step.StepType
    FIRST, MID, LAST = 0, 1, 2

  ts.StepType = StepType
  specs = types.ModuleType("tf_agents.specs")
  aspec = types.ModuleType("tf_agents.specs.array_spec")
  aspec.BoundedArraySpec = lambda *a, **k: ("BoundedArraySpec", a, k)
  aspec.ArraySpec = lambda *a, **k: ("ArraySpec", a, k)
  for name, mod in {"tf_agents": ta, "tf_agents.environments": env,
                    "tf_agents.environments.wrappers": wrappers, "tf_agents.trajectories": traj,
                    "tf_agents.trajectories.time_step": ts, "tf_agents.specs": specs,
                    "tf_agents.specs.array_spec": aspec}.items():
    sys.modules[name] = mod


def load_reference():
  install_stubs()
  import pyhanabi  # the reference's binding
  assert pyhanabi.try_cdef(prefixes=[REF_HLE])
  assert pyhanabi.try_load(prefixes=[os.path.join(ROOT, "oracle", "_ref")])
  import rl_env
  sys.path.insert(0, os.path.join(REF, "agents", "rainbow_copy"))
  third = types.ModuleType("third_party")
  third.__path__ = [os.path.join(REF, "agents", "rainbow_copy", "third_party")]
  sys.modules.setdefault("third_party", third)
  import dqn_agent
  sys.path.insert(0, os.path.join(REF, "training", "tf_agents_lib"))
  import pyhanabi_env_wrapper
  # run_experiment.py: compile the episode loop from the file where it lies
  path = os.path.join(REF, "agents", "run_experiment.py")
  tree = ast.parse(open(path).read(), path)
  wanted = {"ObservationStacker", "create_obs_stacker", "format_legal_moves", "parse_observations",
            "run_one_episode"}
  body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in wanted]
  assert {n.name for n in body} == wanted
  ns = {"np": np, "tf": sys.modules["tensorflow"], "gin": sys.modules["gin"], "LENIENT_SCORE": False,
        "__name__": "run_experiment_excerpt"}
  exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)  # pylint: disable=exec-used
  return rl_env, dqn_agent, pyhanabi_env_wrapper, ns


# ------------------------------------------------------------------ a11
def epsilon_vectors(dqn_agent):
  rng = np.random.default_rng(0)
  rows = []
  for decay, warm, eps in ((1000, 500, 0.02), (250000, 20000, 0.01), (1, 0, 0.0), (1000.0, 0, 0.5),
                           (333, 77, 0.001)):
    steps = sorted(set([0, 1, warm - 1, warm, warm + 1, warm + int(decay) // 2, warm + int(decay) - 1,
                        warm + int(decay), warm + int(decay) + 1, 10 ** 7] +
                       [int(x) for x in rng.integers(0, warm + 2 * int(decay) + 2, 12)]))
    for s in steps:
      if s < 0:
        continue
      rows.append((decay, s, warm, eps, float(dqn_agent.linearly_decaying_epsilon(decay, s, warm, eps))))
  return np.asarray(rows, np.float64)


# ------------------------------------------------------------------ f2
def actor_loop(rl_env, dqn_agent, ns, params, seeds, min_steps):
  out = {}
  for i, seed in enumerate(seeds):
    env = rl_env.HanabiEnv(dict(params, seed=int(seed)))
    P, A, D = env.players, env.num_moves(), env.vectorized_observation_shape()[0]
    stacker = ns["create_obs_stacker"](env, history_size=1)
    agent = object.__new__(dqn_agent.DQNAgent)  # no constructor: it would build a TF graph
    agent.num_players = P
    agent.transitions = [[] for _ in range(P)]  # dqn_agent.py:181
    agent.eval_mode = False
    steps = dict(player=[], action=[], reward=[], done=[])
    posted = dict(obs=[], action=[], reward=[], terminal=[], legal=[], at_step=[])
    count = [0]

    def select(observation, legal_actions, _count=count):
      legal = np.where(legal_actions == 0.0)[0]
      a = legal[(_count[0] * 7 + 3) % len(legal)]
      _count[0] += 1
      return a  # numpy integer: run_one_episode calls action.item()

    def store(observation, action, reward, is_terminal, legal_actions, _posted=posted, _steps=steps):
      _posted["obs"].append(np.packbits(np.asarray(observation, np.uint8), bitorder="little"))
      _posted["action"].append(int(action))
      _posted["reward"].append(np.float32(reward))
      _posted["terminal"].append(int(bool(is_terminal)))
      _posted["legal"].append((np.asarray(legal_actions) == 0.0).astype(np.uint8))
      _posted["at_step"].append(len(_steps["action"]))  # env steps taken when the episode was posted

    agent._train_step = lambda: None
    agent._select_action = select
    agent._store_transition = store

    # the environment as run_one_episode sees it, with a tap on step() for the trace
    real_step = env.step

    def tapped(action, _env=env, _steps=steps, _real=real_step):
      player = _env.state.cur_player()
      o, r, d, info = _real(action)
      _steps["player"].append(player); _steps["action"].append(int(action))
      _steps["reward"].append(r); _steps["done"].append(int(d))
      return o, r, d, info

    env.step = tapped
    episodes = 0
    while len(steps["action"]) < min_steps:
      ns["run_one_episode"](agent, env, stacker)
      episodes += 1
    for k, v in steps.items():
      out["env%d/step_%s" % (i, k)] = np.asarray(v, np.int32)
    for k, v in posted.items():
      out["env%d/post_%s" % (i, k)] = np.asarray(v)
    out["env%d/shape" % i] = np.asarray([P, A, D, episodes, seed])
  return out


# ------------------------------------------------------------------ f4
def wrapper_traces(rl_env, wrapper_mod, params, seeds, steps):
  recs = dict(action=[], step_type=[], reward=[], discount=[], state=[], mask_legal=[], mask_min=[], info=[])
  for seed in seeds:
    env = rl_env.HanabiEnv(dict(params, seed=int(seed)))
    w = wrapper_mod.PyhanabiEnvWrapper(env)
    rng = np.random.default_rng(int(seed))
    row = {k: [] for k in recs}

    def emit(ts, action, _row=row):
      _row["action"].append(action)
      _row["step_type"].append(int(ts.step_type))
      _row["reward"].append(np.float32(ts.reward))
      _row["discount"].append(np.float32(ts.discount))
      _row["state"].append(np.packbits(np.asarray(ts.observation["state"], np.uint8), bitorder="little"))
      m = np.asarray(ts.observation["mask"])
      assert m.dtype == np.float32
      _row["mask_legal"].append((m == 0).astype(np.uint8))
      _row["mask_min"].append(np.float32(m.min()))
      _row["info"].append(int(ts.observation["info"]))

    ts = w.reset()
    emit(ts, -1)
    for _ in range(steps):
      legal = np.flatnonzero(np.asarray(ts.observation["mask"]) == 0)
      a = int(rng.choice(legal))
      ts = w.step(np.asarray(a))
      emit(ts, a)
    for k in recs:
      recs[k].append(np.asarray(row[k]))
  return {k: np.stack(v) for k, v in recs.items()}


def main():
  rl_env, dqn_agent, wrapper_mod, ns = load_reference()
  out = {"epsilon": epsilon_vectors(dqn_agent)}
  # greedy branch: tf.argmax(q + legal, axis=1)[0] -- first maximal index; numpy's argmax has
  # the same tie rule (TensorFlow itself is absent), recorded on rows with ties and -inf
  rng = np.random.default_rng(1)
  q = np.round(rng.normal(size=(64, 20)).astype(np.float32), 1)  # coarse values: many ties
  legal = np.where(rng.random((64, 20)) < 0.4, 0.0, -np.inf).astype(np.float32)
  legal[:, 3] = 0.0
  out["greedy/q"], out["greedy/legal"] = q, legal
  out["greedy/argmax"] = np.argmax(q + legal, axis=1).astype(np.int32)
  for name, seeds, min_steps in (("full2", range(300, 308), 150), ("small3", range(400, 408), 120),
                                 ("full4", range(500, 504), 150)):
    params = dict(CONFIGS["small2"], players=3) if name == "small3" else CONFIGS[name]
    for k, v in actor_loop(rl_env, dqn_agent, ns, params, list(seeds), min_steps).items():
      out["loop/%s/%s" % (name, k)] = v
    out["loop/%s/seeds" % name] = np.asarray(list(seeds))
    print("loop", name, "envs", len(list(seeds)))
  for name, seeds, steps in (("small2", range(100, 124), 70), ("full3", range(200, 208), 90)):
    params = dict(CONFIGS[name], observation_type=1)
    for k, v in wrapper_traces(rl_env, wrapper_mod, params, list(seeds), steps).items():
      out["wrapper/%s/%s" % (name, k)] = v
    out["wrapper/%s/seeds" % name] = np.asarray(list(seeds))
    print("wrapper", name, out["wrapper/%s/step_type" % name].shape,
          "LAST steps", int((out["wrapper/%s/step_type" % name] == 2).sum()))
  np.savez_compressed(os.path.join(HERE, "actor_golden.npz"), **out)
  print("epsilon rows", len(out["epsilon"]))


if __name__ == "__main__":
  main()
