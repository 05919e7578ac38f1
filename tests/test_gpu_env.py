"""GPU tests of the layers above the kernels: BatchedHanabiEnv, the host-buffer
entry points and the host twin of the random policy."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common
from tests.common import CONFIGS

pytestmark = pytest.mark.gpu


def test_batched_env_matches_reference_semantics():
  import torch
  from hanabi_b200 import rl_env
  n = 300
  env = rl_env.make_batched("Hanabi-Full", num_players=3, num_games=n, seed=41)
  cpu = O.CpuBatch("ref" if O.have_ref() else "oracle",
                   dict(players=3, colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3,
                        observation_type=1), n, common.seeds_for(n, base=41))
  assert env.vectorized_observation_shape() == [956] and env.num_moves() == 30
  obs = env.reset()
  cpu.reset()
  rng = np.random.default_rng(0)
  for t in range(60):
    common.assert_same(cpu.encode(), obs["vectorized"].cpu().numpy(), "vectorized @%d" % t)
    mask = cpu.legal_mask()
    common.assert_same(mask, obs["legal_moves_mask"].cpu().numpy(), "mask @%d" % t)
    common.assert_same(cpu.cur_player(), obs["current_player"].cpu().numpy(), "cur @%d" % t)
    legal_f = env.legal_moves_as_float().cpu().numpy()
    assert ((legal_f == 0) == (mask == 1)).all() and np.isneginf(legal_f[mask == 0]).all()
    a = common.choose_actions(None, mask, rng, False, 5, 5)
    obs, reward, done, info = env.step(torch.as_tensor(a, device=env.device))
    r, d, s, _ = cpu.step(a)
    common.assert_same(r, reward.cpu().numpy(), "reward")
    common.assert_same(d, done.cpu().numpy(), "done")
    common.assert_same(s, info["score"].cpu().numpy(), "score")
  st = env.episode_statistics()
  assert st["episodes"] > 0


@pytest.mark.parametrize("packed_link", [True, False])
def test_host_buffer_entry_points_match_device_path(packed_link):
  """packed_link: the uint8 rows cross PCIe bit-packed and the host pool expands them
  (BatchSetHostLink(1), the default) or cross as the kernel's bytes (0)."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 70000  # several chunks, ragged last tile
  params = CONFIGS["full2"]
  seeds = common.seeds_for(n, base=5)
  dev = ph.HanabiBatch(params, n, seeds)
  host = ph.HanabiBatch(params, n, seeds)
  assert host.host_link_packed  # the default
  host.set_host_link(packed_link)
  assert host.host_link_packed == packed_link
  D, A = dev.obs_size, dev.max_moves
  cuda = torch.device("cuda:0")
  d_obs = torch.zeros((n, D), dtype=torch.uint8, device=cuda)
  d_mask = torch.zeros((n, A), dtype=torch.uint8, device=cuda)
  d_cur = torch.zeros(n, dtype=torch.int32, device=cuda)
  d_rew = torch.zeros(n, dtype=torch.int32, device=cuda)
  d_done = torch.zeros(n, dtype=torch.uint8, device=cuda)
  d_score = torch.zeros(n, dtype=torch.int32, device=cuda)
  d_act = torch.zeros(n, dtype=torch.int32, device=cuda)
  h_obs = torch.zeros((n, D), dtype=torch.uint8).pin_memory()
  h_mask = torch.zeros((n, A), dtype=torch.uint8).pin_memory()
  h_cur = torch.zeros(n, dtype=torch.int32).pin_memory()
  h_rew = torch.zeros(n, dtype=torch.int32).pin_memory()
  h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
  h_score = torch.zeros(n, dtype=torch.int32).pin_memory()
  h_act = torch.zeros(n, dtype=torch.int32).pin_memory()
  dev.reset()
  dev.observe(d_obs, d_mask, d_cur)
  host.reset_host(h_obs, h_mask, h_cur)
  for t in range(12):
    assert torch.equal(d_obs.cpu(), h_obs) and torch.equal(d_mask.cpu(), h_mask)
    assert torch.equal(d_cur.cpu(), h_cur)
    # the host policy is the bit-identical twin of the device kernel
    ph.select_action(None, d_mask, 1.0, seed=7, step=t, out=d_act)
    ph.host_random_legal_actions(h_mask, seed=7, step=t, out=h_act)
    assert torch.equal(d_act.cpu(), h_act)
    dev.step_encode(d_act, d_rew, d_done, d_obs, d_mask, d_cur, scores=d_score)
    host.step_encode_host(h_act, h_rew, h_done, h_obs, h_mask, h_cur, scores=h_score)
    assert torch.equal(d_rew.cpu(), h_rew) and torch.equal(d_done.cpu(), h_done)
    assert torch.equal(d_score.cpu(), h_score)
  # padded host pitch
  h_pad = torch.full((n, D + 14), 9, dtype=torch.uint8).pin_memory()
  host.step_encode_host(h_act, h_rew, h_done, h_pad, h_mask, h_cur)
  dev.step_encode(d_act, d_rew, d_done, d_obs, d_mask, d_cur)
  assert torch.equal(h_pad[:, :D], d_obs.cpu()) and (h_pad[:, D:] == 9).all()
  # bit-packed rows in the host buffer unpack to the byte form
  W = dev.packed_words
  h_bits = torch.zeros((n, W), dtype=torch.int32).pin_memory()
  for t in range(12, 16):
    ph.host_random_legal_actions(h_mask, seed=7, step=t, out=h_act)
    d_act.copy_(h_act)
    host.step_encode_host_packed(h_act, h_rew, h_done, h_bits, h_mask, h_cur, scores=h_score)
    dev.step_encode(d_act, d_rew, d_done, d_obs, d_mask, d_cur, scores=d_score)
    raw = np.unpackbits(h_bits.numpy().view(np.uint8), axis=1, bitorder="little")
    common.assert_same(d_obs.cpu().numpy(), raw[:, :D], "host packed rows @%d" % t)
    assert not raw[:, D:].any()
    assert torch.equal(d_mask.cpu(), h_mask) and torch.equal(d_cur.cpu(), h_cur)
    assert torch.equal(d_rew.cpu(), h_rew) and torch.equal(d_done.cpu(), h_done)
    assert torch.equal(d_score.cpu(), h_score)


def test_new_batch_rejects_bad_parameters():
  from hanabi_b200 import pyhanabi as ph
  with pytest.raises(RuntimeError, match="hand_size"):
    ph.HanabiBatch(dict(players=2, hand_size=9), 4, [1, 2, 3, 4])  # 8-bit hand masks: hanabi_state.cc:30,43
  assert ph.HanabiBatch(dict(players=2, hand_size=8), 4, [1, 2, 3, 4]).max_moves == 26
  with pytest.raises(RuntimeError, match="players"):
    ph.HanabiBatch(dict(players=6), 4, [1, 2, 3, 4])
  with pytest.raises(RuntimeError, match="deck"):
    ph.HanabiBatch(dict(players=5, colors=1, ranks=5, hand_size=4), 4, [1, 2, 3, 4])


def test_tf_agents_style_adapter_follows_the_wrapper_protocol():
  """The batched TimeStep adapter vs the reference wrapper's logic
  (training/tf_agents_lib/pyhanabi_env_wrapper.py:41-90) replayed game by game on
  the single-game HanabiEnv."""
  import torch
  from hanabi_b200 import rl_env, tf_agents_adapter as T
  n, steps = 24, 70
  cfg = dict(players=2, colors=2, ranks=5, hand_size=2, max_information_tokens=3,
             max_life_tokens=1, observation_type=1)
  benv = rl_env.BatchedHanabiEnv(dict(cfg, seed=100), n, auto_reset=False)
  wrapped = T.BatchedPyhanabiEnvWrapper(benv)
  ts = wrapped.reset()
  envs = [rl_env.HanabiEnv(dict(cfg, seed=100 + i)) for i in range(n)]
  fmin = np.finfo(np.float32).min

  def ref_obs(env, observations):
    o = observations["player_observations"][observations["current_player"]]
    m = np.full(env.num_moves(), fmin, np.float32)
    m[o["legal_moves_as_int"]] = 0
    return np.asarray(o["vectorized"], np.uint8), m, env.state.score()

  ref = [ref_obs(e, e.reset()) for e in envs]
  ended = [False] * n
  rng = np.random.default_rng(4)
  assert (ts.step_type.cpu().numpy() == T.FIRST).all() and float(ts.discount) == pytest.approx(0.999)
  for t in range(steps):
    state = ts.observation["state"].cpu().numpy()
    mask = ts.observation["mask"].cpu().numpy()
    info = ts.observation["info"].cpu().numpy()
    for i in range(n):
      assert (state[i] == ref[i][0]).all() and (mask[i] == ref[i][1]).all() and info[i] == ref[i][2]
    actions = np.asarray([int(rng.choice(np.flatnonzero(mask[i] == 0))) for i in range(n)])
    ts = wrapped.step(torch.as_tensor(actions, device=benv.device))
    st, rw = ts.step_type.cpu().numpy(), ts.reward.cpu().numpy()
    for i, e in enumerate(envs):
      if ended[i]:  # wrapper :64-67: ignore the action, start a new episode
        ref[i] = ref_obs(e, e.reset())
        ended[i] = False
        assert st[i] == T.FIRST and rw[i] == 0
      else:
        observations, reward, done, _ = e.step(int(actions[i]))
        ref[i] = ref_obs(e, observations)
        ended[i] = done
        assert st[i] == (T.LAST if done else T.MID) and rw[i] == np.float32(reward)
  assert wrapped.action_spec()["maximum"] == benv.num_moves() - 1


def test_self_play_recorder_posts_the_reference_trajectories():
  """hanabi_b200.actor.SelfPlayRecorder vs the per-game restatement of the agent's
  transition bookkeeping (oracle/actor_oracle.py); posted runs land in the device
  replay memory as contiguous per-seat episodes."""
  import torch
  from hanabi_b200 import actor, replay, rl_env
  from hanabi_b200 import pyhanabi as ph
  from oracle import actor_oracle
  n, steps = 40, 90
  env = rl_env.make_batched("Hanabi-Small", num_players=3, num_games=n, seed=77)
  D, A, P = env.batch.obs_size, env.batch.max_moves, env.players
  mem = replay.DeviceReplayMemory(A, D, 1 << 14, update_horizon=3, gamma=0.99)
  rec = actor.SelfPlayRecorder(n, P, D, A, env.device, replay=mem)
  books = [actor_oracle.EpisodeBook(P) for _ in range(n)]
  obs = env.reset()
  act = torch.zeros(n, dtype=torch.int32, device=env.device)
  want_posted, got_posted = [], []
  for t in range(steps):
    ph.select_action(None, obs["legal_moves_mask"], 1.0, seed=5, step=t, out=act)
    cur = obs["current_player"].clone()
    rec.before_step(cur, obs["vectorized"], obs["legal_moves_mask"], act)
    o_h, m_h, c_h, a_h = (obs["vectorized"].cpu().numpy(), obs["legal_moves_mask"].cpu().numpy(),
                          cur.cpu().numpy(), act.cpu().numpy())
    for i in range(n):
      books[i].act(int(c_h[i]), o_h[i], m_h[i], int(a_h[i]))
    obs, reward, done, _ = env.step(act)
    out = rec.after_step(reward, done)
    r_h, d_h = reward.cpu().numpy(), done.cpu().numpy()
    for i in range(n):
      posted = books[i].env_result(float(r_h[i]), bool(d_h[i]))
      if posted:
        want_posted.extend(posted)
    if out is not None:
      o, a, r, term, legal = [x.cpu().numpy() for x in out]
      for k in range(len(a)):
        mask_bytes = (legal[k] == 0).astype(np.uint8).tobytes()
        got_posted.append((o[k].tobytes(), int(a[k]), np.float32(r[k]), int(term[k]), mask_bytes))
  assert len(want_posted) > 200 and rec.dropped == 0
  # same games finish in the same step on both sides and runs are ordered by game then seat
  assert got_posted == want_posted
  assert mem.add_count == len(want_posted)
  # the replay sees each seat's episode as one contiguous run: a 3-step window that
  # starts inside a run never mixes rewards of two runs unless it crosses a terminal
  idx = torch.arange(0, mem.add_count - 4, dtype=torch.int32, device=env.device)
  state, action, reward3, next_state, terminal, _, _ = mem.sample_transition_batch(indices=idx)
  flat_r = np.asarray([p[2] for p in want_posted], np.float32)
  flat_t = np.asarray([p[3] for p in want_posted], np.uint8)
  for i in range(0, len(idx), 17):
    acc, term = np.float32(0), 0
    for j in range(3):
      acc = np.float32(acc + np.float32(0.99 ** j) * flat_r[i + j])
      if flat_t[i + j]:
        term = 1
        break
    assert abs(float(reward3[i]) - float(acc)) < 1e-5 and int(terminal[i]) == term


def _move_item_tuple(m):
  mv = m.move()
  return (str(mv), m.player(), m.scored(), m.information_token(), m.color(), m.rank(),
          tuple(m.card_info_revealed()), m.deal_to_player())


@pytest.mark.parametrize("players", [2, 3, 5])
def test_exported_game_with_history_keeps_discard_order_and_last_moves(players):
  """With BatchEnableHistory the single-game view of a batched game has the discard pile in the
  order the cards were discarded (rl_env.py:416-418) and, for EVERY observer, exactly the
  last_moves the reference observation holds (hanabi_observation.cc:77-92: player moves and
  deals back to the observer's own previous move) -- against the single-game HanabiEnv."""
  import torch
  from hanabi_b200 import rl_env
  n, probe = 70, [0, 33, 69]
  env = rl_env.make_batched("Hanabi-Full", num_players=players, num_games=n, seed=300)
  env.batch.enable_history()
  singles = {}
  for i in probe:
    cfg = dict(colors=5, ranks=5, players=players, max_information_tokens=8, max_life_tokens=3,
               observation_type=1, seed=300 + i)
    singles[i] = rl_env.HanabiEnv(cfg)
    singles[i].reset()
  obs = env.reset()
  rng = np.random.default_rng(11)
  alive = set(probe)
  saw_discards = 0
  for t in range(60):
    mask = obs["legal_moves_mask"].cpu().numpy()
    for i in sorted(alive):
      got = env.game_observations(i)
      want = singles[i]._make_observation_all_players()
      for pg, pw in zip(got["player_observations"], want["player_observations"]):
        assert pg["discard_pile"] == pw["discard_pile"], (i, t)   # order included
        saw_discards = max(saw_discards, len(pw["discard_pile"]))
        mine = [_move_item_tuple(m) for m in pg["pyhanabi"].last_moves()]
        ref = [_move_item_tuple(m) for m in pw["pyhanabi"].last_moves()]
        assert mine == ref, (i, t)
    # a policy that discards and plays a lot, so that piles grow and deals follow
    a = common.choose_actions(None, mask, rng, False, 5, 5)
    low = (mask[:, :2 * (5 if players < 4 else 4)] * rng.random((n, 2 * (5 if players < 4 else 4)))).argmax(axis=1)
    use_low = (mask[:, :2 * (5 if players < 4 else 4)].sum(axis=1) > 0) & (rng.random(n) < 0.6)
    a = np.where(use_low, low, a).astype(np.int32)
    obs, reward, done, _ = env.step(torch.as_tensor(a, device=env.device))
    for i in sorted(alive):
      _, r, d, _ = singles[i].step(int(a[i]))
      assert r == int(reward[i]) and d == bool(done[i])
      if d:
        alive.discard(i)
    if not alive:
      break
  assert saw_discards >= 4


@pytest.mark.parametrize("players", [2, 4])
def test_exported_game_gives_the_reference_observation_dict(players):
  """BatchedHanabiEnv.game_observations(i) against the single-game HanabiEnv (same seed,
  same moves): every key of every player's dict, the discard pile as a multiset."""
  import torch
  from hanabi_b200 import rl_env
  n, probe = 97, [0, 31, 32, 96]
  env = rl_env.make_batched("Hanabi-Full", num_players=players, num_games=n, seed=200)
  singles = {}
  for i in probe:
    cfg = dict(colors=5, ranks=5, players=players, max_information_tokens=8, max_life_tokens=3,
               observation_type=1, seed=200 + i)
    singles[i] = rl_env.HanabiEnv(cfg)
    singles[i].reset()
  obs = env.reset()
  rng = np.random.default_rng(5)
  alive = set(probe)
  for t in range(45):
    mask = obs["legal_moves_mask"].cpu().numpy()
    for i in sorted(alive):
      got = env.game_observations(i)
      want = singles[i]._make_observation_all_players()
      assert got["current_player"] == want["current_player"]
      for pg, pw in zip(got["player_observations"], want["player_observations"]):
        for key in pw:
          if key == "pyhanabi":
            continue
          if key == "discard_pile":
            order = lambda cards: sorted((c["color"], c["rank"]) for c in cards)
            assert order(pg[key]) == order(pw[key])
          else:
            assert pg[key] == pw[key], (i, t, key)
      cur = got["current_player"]
      assert got["player_observations"][cur]["vectorized"] == obs["vectorized"][i].cpu().tolist()
      # without the history ring the newest player move (and the deal after it) is kept
      moves = [m for m in want["player_observations"][cur]["pyhanabi"].last_moves()
               if m.move().type() != 5]
      mine = [m for m in got["player_observations"][cur]["pyhanabi"].last_moves()
              if m.move().type() != 5]
      if moves:
        assert str(mine[0].move()) == str(moves[0].move()) and mine[0].player() == moves[0].player()
        assert mine[0].scored() == moves[0].scored()
        assert mine[0].information_token() == moves[0].information_token()
        assert mine[0].card_info_revealed() == moves[0].card_info_revealed()
    a = common.choose_actions(None, mask, rng, False, 5, 5)
    obs, reward, done, _ = env.step(torch.as_tensor(a, device=env.device))
    for i in sorted(alive):
      _, r, d, _ = singles[i].step(int(a[i]))
      assert r == int(reward[i]) and d == bool(done[i])
      if d:
        alive.discard(i)  # the batch re-deals on the same RNG stream; the single env would re-seed
    if not alive:
      break


@pytest.mark.parametrize("links", [(False, True), (True, False), (True, True)])
def test_ranged_async_host_step_matches_the_synchronous_call(links):
  """links = (synchronous call, ranged calls): packed host link or bytes on the link."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 40000
  params = CONFIGS["full2"]
  seeds = common.seeds_for(n, base=77)
  a, b = ph.HanabiBatch(params, n, seeds), ph.HanabiBatch(params, n, seeds)
  a.set_host_link(links[0])
  b.set_host_link(links[1])
  D, A = a.obs_size, a.max_moves
  pin = lambda *shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
  bufs = []
  for x in (a, b):
    t = dict(obs=pin(n, D, dt=torch.uint8), mask=pin(n, A, dt=torch.uint8), cur=pin(n, dt=torch.int32),
             rew=pin(n, dt=torch.int32), done=pin(n, dt=torch.uint8), act=pin(n, dt=torch.int32),
             score=pin(n, dt=torch.int32))
    x.reset_host(t["obs"], t["mask"], t["cur"])
    bufs.append(t)
  ta, tb = bufs
  launch, wait = b.bind_step_encode_host_range(tb["act"], tb["rew"], tb["done"], tb["obs"], tb["mask"],
                                               tb["cur"], scores=tb["score"])
  ranges = [(0, 12800), (12800, 12800), (25600, n - 25600)]  # ragged last range
  for t in range(10):
    ph.host_random_legal_actions(ta["mask"], seed=3, step=t, out=ta["act"])
    a.step_encode_host(ta["act"], ta["rew"], ta["done"], ta["obs"], ta["mask"], ta["cur"], scores=ta["score"])
    for lo, cnt in ranges:
      wait(lo)
      tb["act"][lo:lo + cnt].copy_(ta["act"][lo:lo + cnt])
      launch(lo, cnt)
    wait(-1)
    for k in ("obs", "mask", "cur", "rew", "done", "score"):
      assert torch.equal(ta[k], tb[k]), (k, t)
  with pytest.raises(RuntimeError, match="multiple of 128"):
    launch(64, 128)
  # the same ranges with bit-packed rows in the host buffer (BatchStepEncodeHostRangePacked)
  bits = pin(n, b.packed_words, dt=torch.int32)
  launch_p, wait_p = b.bind_step_encode_host_range(tb["act"], tb["rew"], tb["done"], None, tb["mask"],
                                                   tb["cur"], scores=tb["score"], obs_bits=bits)
  dev_bits = torch.zeros((n, b.packed_words), dtype=torch.int32, device="cuda")
  for t in range(10, 14):
    ph.host_random_legal_actions(ta["mask"], seed=3, step=t, out=ta["act"])
    a.step_encode_host(ta["act"], ta["rew"], ta["done"], ta["obs"], ta["mask"], ta["cur"], scores=ta["score"])
    for lo, cnt in ranges:
      wait_p(lo)
      tb["act"][lo:lo + cnt].copy_(ta["act"][lo:lo + cnt])
      launch_p(lo, cnt)
    wait_p(-1)
    for k in ("mask", "cur", "rew", "done", "score"):
      assert torch.equal(ta[k], tb[k]), (k, t)
    a.observe_packed(dev_bits)
    assert torch.equal(dev_bits.cpu(), bits), t


@pytest.mark.parametrize("players,game", [(3, "Hanabi-Small"), (2, "Hanabi-Full")])
def test_device_recorder_posts_the_reference_trajectories(players, game):
  """The kernel recorder (hanabi_b200.actor.DeviceRecorder) vs the per-game restatement of
  the agent's transition bookkeeping (oracle/actor_oracle.py): the replay ring holds exactly
  the posted stream -- observation bytes, action, reward, terminal, legal vector -- in the
  same order (games that finish in a step by index, then seat, then time)."""
  import torch
  from hanabi_b200 import actor, replay, rl_env
  from hanabi_b200 import pyhanabi as ph
  from oracle import actor_oracle
  n, steps = 1500, 70  # more than one scan block
  env = rl_env.make_batched(game, num_players=players, num_games=n, seed=77)
  D, A, P = env.batch.obs_size, env.batch.max_moves, env.players
  mem = replay.DeviceReplayMemory(A, D, 1 << 18, update_horizon=1, gamma=0.99)
  rec = actor.DeviceRecorder(n, P, D, A, mem)
  books = [actor_oracle.EpisodeBook(P) for _ in range(n)]
  obs = env.reset()
  bits = torch.zeros((n, env.batch.packed_words), dtype=torch.int32, device=env.device)
  act = torch.zeros(n, dtype=torch.int32, device=env.device)
  want = []
  for t in range(steps):
    env.batch.observe_packed(bits)
    ph.select_action(None, obs["legal_moves_mask"], 1.0, seed=5, step=t, out=act)
    cur = obs["current_player"].clone()
    rec.before_step(cur, bits, obs["legal_moves_mask"], act)
    o_h, m_h, c_h, a_h = (obs["vectorized"].cpu().numpy(), obs["legal_moves_mask"].cpu().numpy(),
                          cur.cpu().numpy(), act.cpu().numpy())
    for i in range(n):
      books[i].act(int(c_h[i]), o_h[i], m_h[i], int(a_h[i]))
    obs, reward, done, _ = env.step(act)
    before = mem.add_count
    posted = rec.after_step(reward, done)
    r_h, d_h = reward.cpu().numpy(), done.cpu().numpy()
    want_now = []
    for i in range(n):
      p = books[i].env_result(float(r_h[i]), bool(d_h[i]))
      if p:
        want_now.extend(p)
    assert posted == len(want_now) and mem.add_count == before + posted
    want.extend(want_now)
  assert len(want) > 2000 and rec.dropped == 0 and len(want) < (1 << 18)
  # read the ring back: with update_horizon 1, slot i gives (obs_i, a_i, r_i, obs_{i+1}, t_i, ., legal_{i+1})
  idx = torch.arange(0, len(want) - 1, dtype=torch.int32, device=env.device)
  state, action, reward1, next_state, terminal, _, next_legal = mem.sample_transition_batch(indices=idx)
  state, action, reward1, terminal, next_legal = [x.cpu().numpy() for x in (state, action, reward1, terminal, next_legal)]
  for i in range(len(want) - 1):
    o, a, r, term, legal = want[i]
    assert state[i].tobytes() == o and int(action[i]) == a and np.float32(reward1[i]) == r and int(terminal[i]) == term, i
    assert (next_legal[i] == 0).astype(np.uint8).tobytes() == want[i + 1][4], i
  pri = mem.get_priority(idx).cpu().numpy()
  assert (pri == 100.0).all()


@pytest.mark.parametrize("cfg", [dict(players=2, hand_size=7), dict(players=4, hand_size=8, observation_type=2)])
def test_wide_hands_through_every_view(cfg):
  """Hands of six to eight cards (the engine's 64-bit hand words): the single-game view of a
  batched game with history (dict keys, discard order, last_moves incl. reveal bitmasks of
  cards 5-7), the checkpoint blob, the bit-packed rows and the host-buffer path -- against
  the single-game HanabiEnv on the same seeds and moves."""
  import torch
  from hanabi_b200 import pyhanabi as ph, rl_env
  n, probe = 70, [0, 33, 69]
  base = dict(colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3, observation_type=1)
  base.update(cfg)
  env = rl_env.BatchedHanabiEnv(dict(base, seed=500), n)
  env.batch.enable_history()
  singles = {}
  for i in probe:
    singles[i] = rl_env.HanabiEnv(dict(base, seed=500 + i))
    singles[i].reset()
  obs = env.reset()
  H = base["hand_size"]
  D, A, W = env.batch.obs_size, env.batch.max_moves, env.batch.packed_words
  bits = torch.zeros((n, W), dtype=torch.int32, device=env.device)
  rng = np.random.default_rng(21)
  alive = set(probe)
  hints_seen = 0
  for t in range(50):
    mask = obs["legal_moves_mask"].cpu().numpy()
    env.batch.observe_packed(bits)
    raw = np.unpackbits(bits.cpu().numpy().view(np.uint8).reshape(n, W * 4), axis=1, bitorder="little")
    common.assert_same(obs["vectorized"].cpu().numpy(), raw[:, :D], "packed rows @%d" % t)
    for i in sorted(alive):
      got = env.game_observations(i)
      want = singles[i]._make_observation_all_players()
      assert got["current_player"] == want["current_player"]
      for pg, pw in zip(got["player_observations"], want["player_observations"]):
        for key in pw:
          if key != "pyhanabi":
            assert pg[key] == pw[key], (i, t, key)
        mine = [_move_item_tuple(m) for m in pg["pyhanabi"].last_moves()]
        ref = [_move_item_tuple(m) for m in pw["pyhanabi"].last_moves()]
        assert mine == ref, (i, t)
        hints_seen += sum(1 for m in pw["pyhanabi"].last_moves() if m.card_info_revealed())
    if t == 20:  # checkpoint round trip in the middle of the episodes
      blob = env.batch.get_state()
      again = ph.HanabiBatch(env.config, n, [1] * n)
      again.enable_history()
      again.set_state(blob)
      o2 = torch.zeros((n, D), dtype=torch.uint8, device=env.device)
      m2 = torch.zeros((n, A), dtype=torch.uint8, device=env.device)
      c2 = torch.zeros(n, dtype=torch.int32, device=env.device)
      again.observe(o2, m2, c2)
      assert torch.equal(o2, obs["vectorized"]) and torch.equal(m2, obs["legal_moves_mask"])
    a = common.choose_actions(None, mask, rng, False, 5, H)
    obs, reward, done, _ = env.step(torch.as_tensor(a, device=env.device))
    for i in sorted(alive):
      _, r, d, _ = singles[i].step(int(a[i]))
      assert r == int(reward[i]) and d == bool(done[i])
      if d:
        alive.discard(i)
    if not alive:
      break
  assert hints_seen > 0
  # the host-buffer path (packed link: HostExpandPacked on rows of this width) against the device rows
  hb = ph.HanabiBatch(env.config, n, [500 + i for i in range(n)])
  dv = ph.HanabiBatch(env.config, n, [500 + i for i in range(n)])
  pin = lambda *shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
  h_obs, h_mask, h_cur = pin(n, D, dt=torch.uint8), pin(n, A, dt=torch.uint8), pin(n, dt=torch.int32)
  h_rew, h_done, h_act = pin(n, dt=torch.int32), pin(n, dt=torch.uint8), pin(n, dt=torch.int32)
  d_obs = torch.zeros((n, D), dtype=torch.uint8, device=env.device)
  d_mask = torch.zeros((n, A), dtype=torch.uint8, device=env.device)
  d_cur = torch.zeros(n, dtype=torch.int32, device=env.device)
  d_rew = torch.zeros(n, dtype=torch.int32, device=env.device)
  d_done = torch.zeros(n, dtype=torch.uint8, device=env.device)
  hb.reset_host(h_obs, h_mask, h_cur)
  dv.reset(); dv.observe(d_obs, d_mask, d_cur)
  for t in range(12):
    assert torch.equal(d_obs.cpu(), h_obs) and torch.equal(d_mask.cpu(), h_mask)
    ph.host_random_legal_actions(h_mask, seed=9, step=t, out=h_act)
    hb.step_encode_host(h_act, h_rew, h_done, h_obs, h_mask, h_cur)
    dv.step_encode(h_act.to(env.device), d_rew, d_done, d_obs, d_mask, d_cur)
    assert torch.equal(d_rew.cpu(), h_rew) and torch.equal(d_done.cpu(), h_done)
  # illegal actions are flagged per game in the wide layout's scalar column too
  bad = torch.full((n,), 999, dtype=torch.int32, device=env.device)
  dv.step_encode(bad, d_rew, d_done, d_obs, d_mask, d_cur)
  assert bool(dv.error_flags(clear=True).all()) and not bool(dv.error_flags().any())
