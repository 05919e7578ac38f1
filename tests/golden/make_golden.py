"""Generates the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/rl_env_traces.npz   per config: seeded rl_env.HanabiEnv episodes driven through
                                   the reference's own Python stack (rl_env.py -> pyhanabi.py ->
                                   libpyhanabi.so): actions, current player, bit-packed
                                   'vectorized' observation of the player to act, legal-move
                                   mask, reward, done, score
  tests/golden/gui_fixture_2p.json the reference's own golden vector
                                   (gui/tests/tests_pyhanabi_to_gui.py:4-76): hands, the two
                                   658-element vectors and legal_moves_as_int
  tests/golden/c51_docstring.json  the worked example in project_distribution's docstring
                                   (agents/rainbow_copy/rainbow_agent.py:262-266,400-401)
"""
import ast
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
REF_HLE = os.path.join(REF, "hanabi_learning_environment")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF_HLE)

from tests.common import CONFIGS  # noqa: E402


def load_reference_env():
  import pyhanabi  # the reference's module
  assert pyhanabi.try_cdef(prefixes=[REF_HLE])
  assert pyhanabi.try_load(prefixes=[os.path.join(ROOT, "oracle", "_ref")])
  import rl_env
  return pyhanabi, rl_env


def trace_config(rl_env, params, seed, steps):
  cfg = dict(params)
  cfg["seed"] = seed
  env = rl_env.HanabiEnv(cfg)
  A = env.num_moves()
  D = env.vectorized_observation_shape()[0]
  rec = dict(action=[], cur=[], obs=[], mask=[], reward=[], done=[], score=[])

  def emit(o, action, reward, done):
    cur = o["current_player"]
    po = o["player_observations"][cur]
    m = np.zeros(A, np.uint8)
    m[po["legal_moves_as_int"]] = 1
    rec["action"].append(action)
    rec["cur"].append(cur)
    rec["obs"].append(np.packbits(np.asarray(po["vectorized"], np.uint8), bitorder="little"))
    rec["mask"].append(m)
    rec["reward"].append(reward)
    rec["done"].append(int(done))
    rec["score"].append(env.state.score())
    return po["legal_moves_as_int"]

  legal = emit(env.reset(), -1, 0, False)
  for t in range(steps):
    a = int(legal[(t * 7 + 3) % len(legal)])
    o, r, d, _ = env.step(a)
    if d:
      # the trainers reset here (agents/run_experiment.py:380-381, 355); the
      # batched engine's auto-reset does the same inside the step.
      score = env.state.score()
      o = env.reset()
      legal = emit(o, a, r, True)
      rec["score"][-1] = score
    else:
      legal = emit(o, a, r, False)
  return {k: np.asarray(v) for k, v in rec.items()}, D, A


def main():
  pyhanabi, rl_env = load_reference_env()
  out = {}
  for name, params in CONFIGS.items():
    steps = 160 if name.startswith("full") else 120
    for seed in (1234, 7):
      rec, D, A = trace_config(rl_env, params, seed, steps)
      for k, v in rec.items():
        out["%s/%d/%s" % (name, seed, k)] = v
      out["%s/%d/shape" % (name, seed)] = np.asarray([D, A])
    print(name, D, A, "episodes", int(rec["done"].sum()))
  np.savez_compressed(os.path.join(HERE, "rl_env_traces.npz"), **out)

  # the reference's own golden vector
  src = open(os.path.join(REF, "gui", "tests", "tests_pyhanabi_to_gui.py")).read()
  tree = ast.parse(src)
  observations = None
  for node in tree.body:
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "observations":
      observations = ast.literal_eval(node.value)
  assert observations is not None
  fixture = {"current_player": observations["current_player"], "players": []}
  for po in observations["player_observations"]:
    fixture["players"].append({
        "observed_hands": po["observed_hands"],
        "legal_moves_as_int": po["legal_moves_as_int"],
        "vectorized": po["vectorized"],
        "deck_size": po["deck_size"],
        "information_tokens": po["information_tokens"],
        "life_tokens": po["life_tokens"],
    })
  json.dump(fixture, open(os.path.join(HERE, "gui_fixture_2p.json"), "w"))

  c51 = {
      "supports": [[0, 2, 4, 6, 8], [1, 3, 4, 5, 6]],
      "weights": [[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.2, 0.5, 0.1, 0.1]],
      "target_support": [4, 5, 6, 7, 8],
      "projection": [[0.8, 0.0, 0.1, 0.0, 0.1], [0.8, 0.1, 0.1, 0.0, 0.0]],
      "source": "agents/rainbow_copy/rainbow_agent.py:262-266,400-401",
  }
  json.dump(c51, open(os.path.join(HERE, "c51_docstring.json"), "w"))


if __name__ == "__main__":
  main()
