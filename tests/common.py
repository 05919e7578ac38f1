"""Shared helpers for the parity tests."""
import numpy as np

from oracle import oracle as O

CONFIGS = {
    "full2": dict(players=2),
    "full3": dict(players=3),
    "full4": dict(players=4),
    "full5": dict(players=5),
    "small2": dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3,
                   max_life_tokens=1),
    "small4": dict(players=4, colors=2, ranks=5, hand_size=2, max_information_tokens=3,
                   max_life_tokens=1),
    "vsmall2": dict(players=2, colors=1, ranks=5, hand_size=2, max_information_tokens=3,
                    max_life_tokens=1),
    "minimal2": dict(players=2, observation_type=0),
    "seer3": dict(players=3, observation_type=2),
    "randstart4": dict(players=4, random_start_player="true"),
    "ppo_sweep": dict(players=4, colors=3, ranks=5, hand_size=5, max_information_tokens=2,
                      max_life_tokens=1),
    "dealt_out": dict(players=4, colors=2, ranks=5, hand_size=5, max_information_tokens=8,
                      max_life_tokens=1),
    "odd": dict(players=3, colors=4, ranks=3, hand_size=3, max_information_tokens=5,
                max_life_tokens=2),
    # hands of six to eight cards (hanabi_game.cc:38,58: hand_size * players <= deck; the
    # reference's 8-bit hand masks cap a hand at eight): the engine's 64-bit hand words
    "wide6": dict(players=2, hand_size=6),
    "wide8": dict(players=3, hand_size=8),
    "wide7seer": dict(players=5, hand_size=7, observation_type=2),
    "wide8dealt": dict(players=5, colors=4, ranks=5, hand_size=8, max_life_tokens=2),  # 40 of 40 cards dealt
}


def fuzz_rule_sets(count=40, seed=2024):
  """Deterministic random rule sets over everything the reference's HanabiGame constructor
  accepts (hanabi_game.cc:29-58: 2-5 players, 1-5 colours, 1-5 ranks, hand_size * players <=
  deck, any token maxima, the three observation types, random start player), plus the corners:
  one colour, one rank, no information tokens, hands of one and of eight cards, decks dealt out
  completely by the initial deal."""
  rng = np.random.default_rng(seed)
  out = [
      dict(players=2, colors=1, ranks=1, hand_size=1, max_information_tokens=1, max_life_tokens=1),
      dict(players=3, colors=1, ranks=5, hand_size=3, max_information_tokens=0, max_life_tokens=2),
      dict(players=5, colors=5, ranks=1, hand_size=3, max_information_tokens=4, max_life_tokens=1),
      dict(players=5, colors=5, ranks=5, hand_size=8, max_information_tokens=8, max_life_tokens=3,
           observation_type=2, random_start_player="true"),
      dict(players=2, colors=5, ranks=5, hand_size=8, observation_type=0),
      dict(players=4, colors=3, ranks=4, hand_size=6, max_information_tokens=3, max_life_tokens=4),  # 24 of 24 dealt
      dict(players=2, colors=2, ranks=2, hand_size=4, max_information_tokens=2, max_life_tokens=1),  # 8 of 8 dealt
  ]
  while len(out) < count:
    P, C, R = int(rng.integers(2, 6)), int(rng.integers(1, 6)), int(rng.integers(1, 6))
    per_color = sum(3 if r == 0 else (1 if r == R - 1 else 2) for r in range(R))
    hmax = min(8, per_color * C // P)
    if hmax < 1:
      continue
    out.append(dict(players=P, colors=C, ranks=R, hand_size=int(rng.integers(1, hmax + 1)),
                    max_information_tokens=int(rng.integers(0, 10)), max_life_tokens=int(rng.integers(1, 6)),
                    observation_type=int(rng.integers(0, 3)),
                    random_start_player="true" if rng.random() < 0.3 else "false"))
  return out


EXPECTED_SHAPES = {  # (obs_dim, num_moves)
    "full2": (658, 20), "full3": (956, 30), "full4": (1041, 38), "full5": (1280, 48),
    "small2": (171, 11), "small4": (281, 25), "vsmall2": (106, 10), "minimal2": (308, 20),
    "seer3": (956, 30),
}


def seeds_for(n, base=1):
  return (np.arange(n, dtype=np.int64) + base).astype(np.int32)


def choose_actions(dump, mask, rng, keep_alive, spec_R, hand_size):
  """Host policy over a batch: uniform random legal, or an omniscient
  'keep-alive' policy (play a playable card if there is one, otherwise hint or
  discard) that drives games to deck exhaustion and completed fireworks."""
  n, A = mask.shape
  u = rng.random((n, A)) * mask
  actions = u.argmax(axis=1).astype(np.int32)
  if not keep_alive:
    return actions
  D = O.DUMP
  cur = dump[:, D["cur"]]
  fw = dump[:, D["fireworks"]:D["fireworks"] + 5]
  for i in range(n):
    if rng.random() < 0.15:
      continue
    p = cur[i]
    cards = dump[i, D["cards"] + p * 8:D["cards"] + p * 8 + hand_size]
    playable = [s for s, c in enumerate(cards) if c >= 0 and fw[i, c // spec_R] == c % spec_R]
    if playable:
      actions[i] = hand_size + playable[int(rng.integers(len(playable)))]
      continue
    legal = np.flatnonzero(mask[i])
    safe = legal[(legal < hand_size) | (legal >= 2 * hand_size)]  # discard or hint, never play
    if len(safe):
      actions[i] = safe[int(rng.integers(len(safe)))]
  return actions


class GpuBatch(object):
  """The CUDA batch behind the verbs of oracle.CpuBatch (numpy in / numpy out),
  so that one replay routine drives both.  Calls go through the C-ABI via cffi."""

  def __init__(self, params, num_games, seeds, pitch=None, fused=True, device=0):
    import torch
    from hanabi_b200 import pyhanabi
    self.torch = torch
    self.b = pyhanabi.HanabiBatch(params, num_games, seeds, device=device)
    self.n = num_games
    self.obs_size, self.max_moves, self.num_players = (self.b.obs_size, self.b.max_moves,
                                                       self.b.num_players)
    dev = self.b.device
    self.pitch = self.obs_size if pitch is None else pitch
    self.fused = fused
    self.t_obs = torch.zeros((num_games, self.pitch), dtype=torch.uint8, device=dev)
    self.t_mask = torch.zeros((num_games, self.max_moves), dtype=torch.uint8, device=dev)
    self.t_cur = torch.zeros(num_games, dtype=torch.int32, device=dev)
    self.t_rew = torch.zeros(num_games, dtype=torch.int32, device=dev)
    self.t_done = torch.zeros(num_games, dtype=torch.uint8, device=dev)
    self.t_score = torch.zeros(num_games, dtype=torch.int32, device=dev)
    self.t_act = torch.zeros(num_games, dtype=torch.int32, device=dev)
    self._fresh = False

  def reset(self, which=None):
    w = None
    if which is not None:
      w = self.torch.as_tensor(np.asarray(which, np.uint8), device=self.b.device)
    self.b.reset(w)
    self._fresh = False

  def step(self, actions, auto_reset=True):
    self.t_act.copy_(self.torch.as_tensor(np.asarray(actions, np.int32)))
    if self.fused:
      self.t_obs.fill_(7)
      self.t_mask.fill_(7)
      self.b.step_encode(self.t_act, self.t_rew, self.t_done, self.t_obs, self.t_mask, self.t_cur,
                         scores=self.t_score, auto_reset=auto_reset)
      self._fresh = True
    else:
      self.b.step(self.t_act, self.t_rew, self.t_done, self.t_score, auto_reset=auto_reset)
      self._fresh = False
    return (self.t_rew.cpu().numpy(), self.t_done.cpu().numpy(), self.t_score.cpu().numpy(), None)

  def _observe(self):
    if not self._fresh:
      self.t_obs.fill_(7)
      self.b.observe(self.t_obs, self.t_mask, self.t_cur)
      self._fresh = True

  def encode(self, observer=-1):
    if observer >= 0:
      out = self.torch.full((self.n, self.pitch), 7, dtype=self.torch.uint8, device=self.b.device)
      self.b.encode(out, observer)
      return out.cpu().numpy()[:, :self.obs_size]
    self._observe()
    return self.t_obs.cpu().numpy()[:, :self.obs_size]

  def legal_mask(self):
    self._observe()
    return self.t_mask.cpu().numpy()

  def cur_player(self):
    self._observe()
    return self.t_cur.cpu().numpy()

  def dump(self):
    return self.b.unpack_state().cpu().numpy()


def assert_same(a, b, what):
  a = np.asarray(a)
  b = np.asarray(b)
  if a.shape != b.shape or not (a == b).all():
    bad = np.argwhere(a != b) if a.shape == b.shape else None
    raise AssertionError("%s differs; first mismatches at %s" %
                         (what, None if bad is None else bad[:5].tolist()))


def lockstep(leader, follower, steps, rng, keep_alive, params, dump_every=16,
             check_all_observers_every=0, auto_reset=True):
  """Drives `leader` (a CPU checker) and `follower` through the same actions and
  compares, every step: observation bytes of the player to act, legal mask,
  current player, reward, done, score; and the whole unpacked state every
  `dump_every` steps.  Returns the number of finished episodes."""
  R = int(params.get("ranks", 5))
  H = int(params.get("hand_size", 5 if int(params.get("players", 2)) < 4 else 4))
  episodes = 0
  max_score = 0
  for t in range(steps):
    mask = leader.legal_mask()
    assert_same(mask, follower.legal_mask(), "legal mask @%d" % t)
    assert_same(leader.encode(), follower.encode(), "observation @%d" % t)
    assert_same(leader.cur_player(), follower.cur_player(), "cur_player @%d" % t)
    dump = leader.dump()
    if dump_every and t % dump_every == 0:
      assert_same(dump, follower.dump(), "state dump @%d" % t)
    if check_all_observers_every and t % check_all_observers_every == 0:
      for o in range(leader.num_players):
        assert_same(leader.encode(o), follower.encode(o), "observation of player %d @%d" % (o, t))
    actions = choose_actions(dump, mask, rng, keep_alive, R, H)
    ra = leader.step(actions, auto_reset=auto_reset)
    rb = follower.step(actions, auto_reset=auto_reset)
    assert_same(ra[0], rb[0], "reward @%d" % t)
    assert_same(ra[1], rb[1], "done @%d" % t)
    assert_same(ra[2], rb[2], "score @%d" % t)
    episodes += int(ra[1].sum())
    if ra[1].any():
      max_score = max(max_score, int(ra[2][ra[1] != 0].max()))
  return episodes, max_score


def replay_golden(engine_factory, traces, name, seed):
  """Replays one committed rl_env trace (tests/golden/rl_env_traces.npz) on a
  single-game batch and compares every emitted record."""
  key = "%s/%d/" % (name, seed)
  D, A = [int(x) for x in traces[key + "shape"]]
  eng = engine_factory(CONFIGS[name], 1, np.asarray([seed], np.int32))
  assert (eng.obs_size, eng.max_moves) == (D, A)
  eng.reset()
  action, cur, obs, mask = (traces[key + k] for k in ("action", "cur", "obs", "mask"))
  reward, done, score = (traces[key + k] for k in ("reward", "done", "score"))
  for t in range(len(action)):
    if t > 0:
      r, d, s, _ = eng.step(np.asarray([action[t]], np.int32), auto_reset=True)
      assert int(r[0]) == int(reward[t]), ("reward", name, seed, t)
      assert int(d[0]) == int(done[t]), ("done", name, seed, t)
      if done[t]:
        assert int(s[0]) == int(score[t]), ("score", name, seed, t)
    want = np.unpackbits(obs[t], bitorder="little")[:D]
    assert_same(want, eng.encode()[0], "golden obs %s/%d @%d" % (name, seed, t))
    assert_same(mask[t], eng.legal_mask()[0], "golden mask %s/%d @%d" % (name, seed, t))
    assert int(eng.cur_player()[0]) == int(cur[t])
