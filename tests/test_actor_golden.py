"""Rows a11 / f2 / f4 against vectors recorded from the reference's own code
(tests/golden/make_actor_golden.py -> tests/golden/actor_golden.npz):

  * `linearly_decaying_epsilon`, agents/rainbow_copy/dqn_agent.py:44-59;
  * the greedy branch of `_select_action`: first maximal index of q + legal, dqn_agent.py:200;
  * the transition bookkeeping of `DQNAgent.begin_episode / step / end_episode /
    _post_transitions` (dqn_agent.py:286-385) run by the reference's `run_one_episode`
    (agents/run_experiment.py:344-403);
  * `PyhanabiEnvWrapper._reset / _step`, training/tf_agents_lib/pyhanabi_env_wrapper.py:41-90.

CPU tests pin the restatements (oracle/actor_oracle.py, hanabi_b200.actor) on the recorded
vectors; the gpu tests run the kernels (DeviceRecorder, BatchSelectAction) and the batched
adapter on the same episodes and compare with the recorded streams directly.
"""
import os

import numpy as np
import pytest

from tests import common

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "actor_golden.npz")
LOOPS = {"full2": ("Hanabi-Full", 2, common.CONFIGS["full2"]),
         "small3": ("Hanabi-Small", 3, dict(common.CONFIGS["small2"], players=3)),
         "full4": ("Hanabi-Full", 4, common.CONFIGS["full4"])}


@pytest.fixture(scope="module")
def golden():
  return np.load(GOLDEN)


def test_epsilon_schedule_equals_the_reference_function(golden):
  from hanabi_b200 import actor
  rows = golden["epsilon"]
  assert len(rows) >= 80
  for decay, step, warm, eps, want in rows:
    got = actor.linearly_decaying_epsilon(decay, int(step), int(warm), eps)
    assert got == want, (decay, step, warm, eps)  # float64, same operation order: exact


def test_masked_argmax_restatement_equals_reference_rule(golden):
  from oracle import c51_oracle
  q, legal, want = golden["greedy/q"], golden["greedy/legal"], golden["greedy/argmax"]
  got = c51_oracle.masked_argmax(q, (legal == 0).astype(np.uint8))
  assert (got == want).all()


def _expected_posts(g, name, i, upto):
  """Recorded posted stream of env i restricted to the episodes that ended within `upto`
  env steps: list of (obs bits, action, reward, terminal, legal bits, at_step)."""
  pre = "loop/%s/env%d/" % (name, i)
  at = g[pre + "post_at_step"]
  keep = at <= upto
  return [(g[pre + "post_obs"][k].tobytes(), int(g[pre + "post_action"][k]),
           np.float32(g[pre + "post_reward"][k]), int(g[pre + "post_terminal"][k]),
           g[pre + "post_legal"][k].astype(np.uint8).tobytes(), int(at[k]))
          for k in np.flatnonzero(keep)]


@pytest.mark.parametrize("name", sorted(LOOPS))
def test_episode_book_reproduces_the_reference_agent_bookkeeping(built_lib, golden, name):
  """oracle/actor_oracle.EpisodeBook (the checker of the gpu recorder tests) fed by this
  repo's single-game HanabiEnv posts exactly what the reference DQNAgent posted."""
  from hanabi_b200 import rl_env
  from oracle import actor_oracle
  g = golden
  _, _, params = LOOPS[name]
  for i, seed in enumerate(g["loop/%s/seeds" % name]):
    pre = "loop/%s/env%d/" % (name, i)
    P, A, D, episodes, seed_rec = [int(x) for x in g[pre + "shape"]]
    assert seed_rec == seed
    env = rl_env.HanabiEnv(dict(params, seed=int(seed)))
    book = actor_oracle.EpisodeBook(P)
    obs = env.reset()
    got = []
    acts, players = g[pre + "step_action"], g[pre + "step_player"]
    for t in range(len(acts)):
      cur = obs["current_player"]
      assert cur == players[t]
      po = obs["player_observations"][cur]
      mask = np.zeros(A, np.uint8)
      mask[po["legal_moves_as_int"]] = 1
      book.act(cur, np.asarray(po["vectorized"], np.uint8), mask, int(acts[t]))
      obs, reward, done, _ = env.step(int(acts[t]))
      assert reward == g[pre + "step_reward"][t] and int(done) == g[pre + "step_done"][t]
      posted = book.env_result(float(reward), bool(done))
      if posted:
        for o, a, r, term, legal in posted:
          bits = np.packbits(np.frombuffer(o, np.uint8), bitorder="little").tobytes()
          got.append((bits, a, np.float32(r), term, legal, t + 1))
      if done:
        obs = env.reset()
    want = _expected_posts(g, name, i, len(acts))
    assert len(want) > 20 and episodes >= 2
    assert got == want


@pytest.mark.gpu
def test_gpu_greedy_select_equals_reference_rule(golden):
  import torch
  from hanabi_b200 import pyhanabi as ph
  q = torch.as_tensor(golden["greedy/q"], device="cuda:0")
  mask = torch.as_tensor((golden["greedy/legal"] == 0).astype(np.uint8), device="cuda:0")
  act = ph.select_action(q, mask, 0.0, seed=1, step=0)
  assert (act.cpu().numpy() == golden["greedy/argmax"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LOOPS))
def test_device_recorder_reproduces_the_reference_agent_bookkeeping(golden, name):
  """The kernel recorder + device replay ring against the stream the reference DQNAgent
  posted for the same seeded games and the same actions: observation bytes, action, reward,
  terminal flag, legal vector, in the ring's order (step, then game, then seat, then time)."""
  import torch
  from hanabi_b200 import actor, replay, rl_env
  g = golden
  game, players, _ = LOOPS[name]
  seeds = g["loop/%s/seeds" % name]
  n = len(seeds)
  T = min(len(g["loop/%s/env%d/step_action" % (name, i)]) for i in range(n))
  actions = np.stack([g["loop/%s/env%d/step_action" % (name, i)][:T] for i in range(n)], axis=1)  # [T, n]
  env = rl_env.make_batched(game, num_players=players, num_games=n, seed=int(seeds[0]))
  assert (np.diff(seeds) == 1).all()
  D, A, P = env.batch.obs_size, env.batch.max_moves, env.players
  mem = replay.DeviceReplayMemory(A, D, 1 << 12, update_horizon=1, gamma=0.99)
  rec = actor.DeviceRecorder(n, P, D, A, mem)
  obs = env.reset()
  bits = torch.zeros((n, env.batch.packed_words), dtype=torch.int32, device=env.device)
  for t in range(T):
    env.batch.observe_packed(bits)
    act = torch.as_tensor(actions[t], dtype=torch.int32, device=env.device)
    cur = obs["current_player"].clone()
    for i in range(n):
      assert int(cur[i]) == int(g["loop/%s/env%d/step_player" % (name, i)][t])
    rec.before_step(cur, bits, obs["legal_moves_mask"], act)
    obs, reward, done, _ = env.step(act)
    rec.after_step(reward, done)
  # expected ring: episodes in the order they were posted -- by step, then game index
  want = []
  per_env = [_expected_posts(g, name, i, T) for i in range(n)]
  for t in range(1, T + 1):
    for i in range(n):
      want.extend(p for p in per_env[i] if p[5] == t)
  assert len(want) > 100 and rec.dropped == 0 and mem.add_count == len(want) < (1 << 12)
  idx = torch.arange(0, len(want) - 1, dtype=torch.int32, device=env.device)
  state, action, reward1, _, terminal, _, next_legal = mem.sample_transition_batch(indices=idx)
  state, action, reward1, terminal, next_legal = [x.cpu().numpy() for x in (state, action, reward1, terminal, next_legal)]
  for k in range(len(want) - 1):
    o, a, r, term, legal, _ = want[k]
    assert np.packbits(state[k], bitorder="little").tobytes() == o, k
    assert int(action[k]) == a and np.float32(reward1[k]) == r and int(terminal[k]) == term, k
    assert (next_legal[k] == 0).astype(np.uint8).tobytes() == want[k + 1][4], k


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small2", "full3"])
def test_batched_adapter_reproduces_the_reference_wrapper(golden, name):
  """BatchedPyhanabiEnvWrapper against TimeSteps recorded from the reference's
  PyhanabiEnvWrapper (one env per seed, the same actions; the step after LAST ignores its
  action and returns FIRST)."""
  import torch
  from hanabi_b200 import rl_env, tf_agents_adapter as T
  g = golden
  pre = "wrapper/%s/" % name
  seeds = g[pre + "seeds"]
  n, steps = g[pre + "action"].shape[0], g[pre + "action"].shape[1] - 1
  assert (np.diff(seeds) == 1).all()
  cfg = dict(common.CONFIGS[name], observation_type=1, seed=int(seeds[0]))
  benv = rl_env.BatchedHanabiEnv(cfg, n, auto_reset=False)
  w = T.BatchedPyhanabiEnvWrapper(benv)
  fmin = np.finfo(np.float32).min

  def check(ts, t):
    state = np.packbits(ts.observation["state"].cpu().numpy(), axis=1, bitorder="little")
    assert (state == g[pre + "state"][:, t]).all(), t
    mask = ts.observation["mask"].cpu().numpy()
    assert mask.dtype == np.float32
    assert ((mask == 0).astype(np.uint8) == g[pre + "mask_legal"][:, t]).all(), t
    assert (mask[mask != 0] == fmin).all() and (g[pre + "mask_min"][:, t] == fmin).all()
    assert (ts.observation["info"].cpu().numpy() == g[pre + "info"][:, t]).all(), t
    assert (ts.step_type.cpu().numpy() == g[pre + "step_type"][:, t]).all(), t
    assert (ts.reward.cpu().numpy() == g[pre + "reward"][:, t]).all(), t
    assert np.float32(ts.discount.item()) == g[pre + "discount"][0, t]

  ts = w.reset()
  check(ts, 0)
  for t in range(1, steps + 1):
    ts = w.step(torch.as_tensor(g[pre + "action"][:, t].astype(np.int64), device=benv.device))
    check(ts, t)
  assert int((g[pre + "step_type"] == T.LAST).sum()) > 20
  assert (T.FIRST, T.MID, T.LAST) == (0, 1, 2)
