"""GPU parity tests: the CUDA engine, called through the C-ABI (cffi), against the
CPU oracle on identical seeds and move sequences.  Bit-exact: observation bytes,
legal masks, rewards, done flags, scores, current player and the whole unpacked
state."""
import functools
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import common
from tests.common import CONFIGS, GpuBatch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def traces():
  return np.load(os.path.join(GOLDEN, "rl_env_traces.npz"))


def _checker(params, n, seeds):
  """The unmodified reference where it was built, else the C restatement."""
  return O.CpuBatch("ref" if O.have_ref() else "oracle", params, n, seeds)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_gpu_replays_reference_rl_env_traces(traces, name):
  for seed in (1234, 7):
    common.replay_golden(GpuBatch, traces, name, seed)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_random_policy_lockstep(name):
  n = 1000  # not a multiple of 32: exercises the ragged last tile
  seeds = common.seeds_for(n, base=1)
  cpu = _checker(CONFIGS[name], n, seeds)
  gpu = GpuBatch(CONFIGS[name], n, seeds)
  cpu.reset(); gpu.reset()
  eps, _ = common.lockstep(cpu, gpu, 96, np.random.default_rng(10), False, CONFIGS[name],
                           dump_every=16, check_all_observers_every=32)
  assert eps > 100


@pytest.mark.parametrize("name", ["full2", "full4", "full5", "small2", "seer3", "dealt_out", "wide6", "wide8",
                                  "wide8dealt"])
def test_keep_alive_policy_lockstep(name):
  n = 256
  seeds = common.seeds_for(n, base=9000)
  cpu = _checker(CONFIGS[name], n, seeds)
  gpu = GpuBatch(CONFIGS[name], n, seeds)
  cpu.reset(); gpu.reset()
  eps, max_score = common.lockstep(cpu, gpu, 300, np.random.default_rng(11), True, CONFIGS[name],
                                   dump_every=20, check_all_observers_every=50)
  assert eps > 0
  if name == "full2":
    assert max_score >= 15


@pytest.mark.parametrize("index", range(0, 40, 5))
def test_fuzzed_rule_sets_lockstep(index):
  """Random rule sets over everything the reference's constructor accepts (common.fuzz_rule_sets:
  1-5 colours and ranks, hands of 1-8 cards, token maxima from 0, all observation types, random
  start player, decks dealt out by the initial deal) in lock-step with the compiled reference:
  the runtime-parameter kernels (DView, WView) on rule sets nobody wrote a case for."""
  for k, params in enumerate(common.fuzz_rule_sets()[index:index + 5]):
    n = 200
    seeds = common.seeds_for(n, base=7000 + 10 * (index + k))
    cpu = _checker(params, n, seeds)
    gpu = GpuBatch(params, n, seeds)
    cpu.reset(); gpu.reset()
    common.lockstep(cpu, gpu, 64, np.random.default_rng(index + k), bool(k & 1), params, dump_every=8,
                    check_all_observers_every=16)


@pytest.mark.parametrize("name", ["small2", "full4"])
def test_rng_stream_across_block_wraps(name):
  """Long lock-step run: every game takes several thousand mt19937 words, i.e. its
  624-word block is regenerated in place several times over (slot 623 / slot 0 boundary,
  the 397-slot look-ahead wrapping), through in-game deals and re-deals alike."""
  n = 96
  seeds = common.seeds_for(n, base=31337)
  cpu = _checker(CONFIGS[name], n, seeds)
  gpu = GpuBatch(CONFIGS[name], n, seeds)
  cpu.reset(); gpu.reset()
  steps = 1200 if name == "small2" else 900
  eps, _ = common.lockstep(cpu, gpu, steps, np.random.default_rng(13), False, CONFIGS[name],
                           dump_every=150)
  assert eps > n


def test_full2_at_survey_size():
  """SURVEY.md section 8d parity run: 4096 games x 256 steps, Hanabi-Full 2p."""
  n = 4096
  seeds = common.seeds_for(n, base=1)
  cpu = O.CpuBatch("oracle", CONFIGS["full2"], n, seeds)
  gpu = GpuBatch(CONFIGS["full2"], n, seeds)
  cpu.reset(); gpu.reset()
  rng = np.random.default_rng(12)
  eps = 0
  for t in range(256):
    mask = cpu.legal_mask()
    keep = (t // 64) % 2 == 1
    if t % 8 == 0:
      common.assert_same(mask, gpu.legal_mask(), "mask @%d" % t)
      common.assert_same(cpu.encode(), gpu.encode(), "obs @%d" % t)
    if t % 64 == 0:
      common.assert_same(cpu.dump(), gpu.dump(), "dump @%d" % t)
    dump = cpu.dump() if keep else None
    actions = common.choose_actions(dump, mask, rng, False, 5, 5) if not keep else \
        common.choose_actions(dump, mask, rng, True, 5, 5)
    ra, rb = cpu.step(actions), gpu.step(actions)
    for k, w in enumerate(("reward", "done", "score")):
      common.assert_same(ra[k], rb[k], "%s @%d" % (w, t))
    eps += int(ra[1].sum())
  common.assert_same(cpu.dump(), gpu.dump(), "final dump")
  assert eps > 10000


@pytest.mark.parametrize("name", ["full2", "full5", "ppo_sweep", "wide8dealt"])
def test_split_kernels_and_padded_pitch(name):
  """BatchStep + BatchObserve (unfused) and obs rows with a 16-byte padded pitch
  and with an odd pitch give the same bytes as the fused dense path."""
  n = 200
  seeds = common.seeds_for(n, base=77)
  cpu = O.CpuBatch("oracle", CONFIGS[name], n, seeds)
  cpu.reset()
  obs_size = cpu.obs_size
  variants = [GpuBatch(CONFIGS[name], n, seeds, fused=False),
              GpuBatch(CONFIGS[name], n, seeds, pitch=(obs_size + 15) // 16 * 16 + 16),
              GpuBatch(CONFIGS[name], n, seeds, pitch=obs_size + 3)]
  for g in variants:
    g.reset()
  rng = np.random.default_rng(13)
  for t in range(40):
    mask = cpu.legal_mask()
    want = cpu.encode()
    for g in variants:
      common.assert_same(want, g.encode(), "obs @%d" % t)
      common.assert_same(mask, g.legal_mask(), "mask @%d" % t)
      if g.pitch != obs_size:  # padding bytes must be left untouched
        assert (g.t_obs.cpu().numpy()[:, obs_size:] == 7).all()
    actions = common.choose_actions(None, mask, rng, False, 5, 5)
    ra = cpu.step(actions)
    for g in variants:
      rb = g.step(actions)
      common.assert_same(ra[0], rb[0], "reward")
      common.assert_same(ra[1], rb[1], "done")


@pytest.mark.parametrize("name", ["full2", "small4"])
def test_generic_kernels_and_exact_sampler(name):
  """The runtime-parameter kernels (used for rule sets without a static
  instantiation) and the always-exact fp64 sampler reproduce the same games."""
  n = 160
  seeds = common.seeds_for(n, base=31)
  cpu = O.CpuBatch("oracle", CONFIGS[name], n, seeds)
  gen = GpuBatch(CONFIGS[name], n, seeds)
  gen.b.set_debug(sampler_mode=0, generic=True)
  exact = GpuBatch(CONFIGS[name], n, seeds)
  exact.b.set_debug(sampler_mode=1, generic=False)
  cpu.reset(); gen.reset(); exact.reset()
  rng = np.random.default_rng(14)
  for t in range(80):
    mask = cpu.legal_mask()
    want = cpu.encode()
    for g in (gen, exact):
      common.assert_same(want, g.encode(), "obs @%d" % t)
      common.assert_same(mask, g.legal_mask(), "mask @%d" % t)
    actions = common.choose_actions(None, mask, rng, False, 5, 5)
    ra = cpu.step(actions)
    for g in (gen, exact):
      rb = g.step(actions)
      common.assert_same(ra[0], rb[0], "reward")
      common.assert_same(ra[1], rb[1], "done")
  common.assert_same(cpu.dump(), gen.dump(), "dump (generic kernels)")
  common.assert_same(cpu.dump(), exact.dump(), "dump (exact sampler)")


def test_deal_sampler_at_boundaries():
  """Fast integer sampler + exact fallback vs libstdc++ (via oracle/_ref) or the
  C restatement, at and around every cumulative-probability boundary."""
  import torch
  from tests.test_oracle import adversarial_draws
  cases = adversarial_draws(np.random.default_rng(6), n_random=20000)
  kind = "ref" if O.have_ref() else "oracle"
  want = np.asarray([O.deal_from_counts(kind, c, lo, hi) for c, lo, hi in cases], np.int32)
  decks = np.zeros(len(cases), np.uint64)
  for k in range(25):
    decks |= np.asarray([int(c[k]) for c, _, _ in cases], np.uint64) << np.uint64(2 * k)
  sizes = np.asarray([int(c.sum()) for c, _, _ in cases], np.int32)
  lo = np.asarray([c[1] for c in cases], np.uint32)
  hi = np.asarray([c[2] for c in cases], np.uint32)
  g = GpuBatch(CONFIGS["full2"], 32, common.seeds_for(32))
  dev = g.b.device
  t = lambda a: torch.as_tensor(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32),
                                device=dev)
  for mode in (0, 1):
    got, amb = g.b.debug_deal(t(decks), t(sizes), t(lo), t(hi), mode)
    common.assert_same(want, got.cpu().numpy(), "deal sampler mode %d" % mode)
  assert int(amb.sum()) > 1000  # the boundary cases do take the exact fallback
  fast, _ = g.b.debug_deal(t(decks), t(sizes), t(lo), t(hi), 2)
  clear = amb.cpu().numpy() == 0
  common.assert_same(want[clear], fast.cpu().numpy()[clear], "fast path outside the guard band")


def test_reference_golden_vector_on_gpu():
  import torch
  from tests.test_oracle import _fixture_dump
  fix = json.load(open(os.path.join(GOLDEN, "gui_fixture_2p.json")))
  g = GpuBatch(CONFIGS["full2"], 1, [1])
  g.reset()
  dump = torch.as_tensor(_fixture_dump(fix)[None, :], device=g.b.device)
  extra = torch.as_tensor(np.asarray([[2, 0]], np.int32), device=g.b.device)
  g.b.pack_state(dump, extra)
  for p in range(2):
    want = np.asarray(fix["players"][p]["vectorized"], np.uint8)
    common.assert_same(want, g.encode(p)[0], "golden vector of player %d" % p)
  g._fresh = False
  assert np.flatnonzero(g.legal_mask()[0]).tolist() == fix["players"][0]["legal_moves_as_int"]
  common.assert_same(_fixture_dump(fix)[:6], g.dump()[0][:6], "dump header")


def test_state_roundtrip_and_determinism():
  """BatchGetState/BatchSetState restore games AND their RNG streams."""
  n = 300
  p = CONFIGS["full3"]
  g = GpuBatch(p, n, common.seeds_for(n, base=5))
  g.reset()
  rng = np.random.default_rng(15)
  for _ in range(20):
    g.step(common.choose_actions(None, g.legal_mask(), rng, False, 5, 5))
  blob = g.b.get_state()
  script = []
  for _ in range(30):
    a = common.choose_actions(None, g.legal_mask(), rng, False, 5, 5)
    script.append((a, g.step(a), g.encode().copy()))
  g.b.set_state(blob)
  g._fresh = False
  for a, res, obs in script:
    got = g.step(a)
    common.assert_same(res[0], got[0], "reward after restore")
    common.assert_same(res[1], got[1], "done after restore")
    common.assert_same(obs, g.encode(), "obs after restore")


def test_no_auto_reset_illegal_actions_and_partial_reset():
  import torch
  n = 100
  p = CONFIGS["small2"]
  seeds = common.seeds_for(n, base=3)
  cpu = O.CpuBatch("oracle", p, n, seeds)
  gpu = GpuBatch(p, n, seeds)
  cpu.reset(); gpu.reset()
  common.lockstep(cpu, gpu, 50, np.random.default_rng(16), False, p, auto_reset=False)
  r = gpu.step(np.zeros(n, np.int32), auto_reset=False)
  assert r[1].all() and (r[0] == 0).all()
  which = (np.arange(n) % 3 == 0).astype(np.uint8)
  cpu.reset(which); gpu.reset(which)
  common.assert_same(cpu.dump(), gpu.dump(), "dump after partial reset")
  cpu.reset(); gpu.reset()
  before = gpu.dump()
  bad = np.full(n, 999, np.int32)
  r = gpu.step(bad)
  assert (r[0] == 0).all() and (r[1] == 0).all()
  common.assert_same(before, gpu.dump(), "illegal actions leave the games unchanged")
  assert bool(gpu.b.error_flags(clear=False).all())
  half = bad.copy()
  half[::2] = p["hand_size"]  # uid H = play slot 0: always legal in a fresh game
  assert bool(gpu.b.error_flags(clear=True).all()) and not bool(gpu.b.error_flags().any())
  r = gpu.step(half)
  flags = gpu.b.error_flags().cpu().numpy()
  assert (flags[1::2] == 1).all() and (flags[::2] == 0).all()
  del torch


def test_episode_statistics():
  n = 512
  p = CONFIGS["small2"]
  seeds = common.seeds_for(n, base=21)
  cpu = O.CpuBatch("oracle", p, n, seeds)
  gpu = GpuBatch(p, n, seeds)
  cpu.reset(); gpu.reset()
  gpu.b.read_stats(clear=True)
  rng = np.random.default_rng(17)
  episodes = score_sum = 0
  hist = np.zeros(26, np.int64)
  for _ in range(60):
    a = common.choose_actions(None, cpu.legal_mask(), rng, False, 5, 2)
    r = cpu.step(a)
    gpu.step(a)
    d = r[1] != 0
    episodes += int(d.sum())
    score_sum += int(r[2][d].sum())
    np.add.at(hist, r[2][d], 1)
  st = gpu.b.read_stats()
  assert st["episodes"] == episodes and st["score_sum"] == score_sum
  assert st["score_hist"] == hist.tolist()
  assert st["length_sum"] > episodes


def test_unaligned_output_buffer_takes_the_byte_path():
  """An observation tensor that is not 16-byte aligned (dense pitch) still gets the
  right bytes, and nothing outside the rows is touched."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 70
  p = CONFIGS["full2"]
  seeds = common.seeds_for(n, base=8)
  cpu = O.CpuBatch("oracle", p, n, seeds)
  cpu.reset()
  b = ph.HanabiBatch(p, n, seeds)
  b.reset()
  D = b.obs_size
  store = torch.full((n * D + 64,), 9, dtype=torch.uint8, device=b.device)
  for shift in (1, 2, 8):
    store.fill_(9)
    view = store[shift:shift + n * D].view(n, D)
    assert view.data_ptr() % 16 != 0
    b.encode(view)
    common.assert_same(cpu.encode(), view.cpu().numpy(), "obs at byte offset %d" % shift)
    assert bool((store[:shift] == 9).all()) and bool((store[shift + n * D:] == 9).all())


@pytest.mark.parametrize("name", ["full2", "full5", "small2", "vsmall2", "odd", "wide6", "wide7seer"])
def test_packed_observations_match_byte_form(name):
  """Bit-packed rows (BatchStepEncodePacked / BatchObservePacked) unpack to exactly
  the byte-per-bit observations, padding bits are zero, for full and ragged tiles."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 333
  p = CONFIGS[name]
  seeds = common.seeds_for(n, base=61)
  a = ph.HanabiBatch(p, n, seeds)
  b = ph.HanabiBatch(p, n, seeds)
  dev = a.device
  W, D, A = a.packed_words, a.obs_size, a.max_moves
  assert W % 4 == 0 and W * 32 >= D and (W - 4) * 32 < D
  obs = torch.zeros((n, D), dtype=torch.uint8, device=dev)
  bits = torch.full((n, W), -1, dtype=torch.int32, device=dev)
  mask_a = torch.zeros((n, A), dtype=torch.uint8, device=dev)
  mask_b = torch.zeros((n, A), dtype=torch.uint8, device=dev)
  cur = torch.zeros(n, dtype=torch.int32, device=dev)
  rew_a = torch.zeros(n, dtype=torch.int32, device=dev); rew_b = torch.zeros_like(rew_a)
  done_a = torch.zeros(n, dtype=torch.uint8, device=dev); done_b = torch.zeros_like(done_a)
  act = torch.zeros(n, dtype=torch.int32, device=dev)
  a.reset(); b.reset()
  a.observe(obs, mask_a, cur)
  b.observe_packed(bits, mask_b, cur)

  def check():
    raw = bits.cpu().numpy().view(np.uint8).reshape(n, W * 4)
    unpacked = np.unpackbits(raw, axis=1, bitorder="little")
    common.assert_same(obs.cpu().numpy(), unpacked[:, :D], "packed vs byte observation")
    assert not unpacked[:, D:].any()
    assert torch.equal(mask_a, mask_b)

  check()
  for t in range(30):
    ph.select_action(None, mask_a, 1.0, seed=9, step=t, out=act)
    a.step_encode(act, rew_a, done_a, obs, mask_a, cur)
    bits.fill_(-1)
    b.step_encode_packed(act, rew_b, done_b, bits, mask_b, cur)
    assert torch.equal(rew_a, rew_b) and torch.equal(done_a, done_b)
    check()
  for observer in range(a.num_players):
    a.encode(obs, observer)
    b.observe_packed(bits, observer=observer)
    raw = np.unpackbits(bits.cpu().numpy().view(np.uint8).reshape(n, W * 4), axis=1, bitorder="little")
    common.assert_same(obs.cpu().numpy(), raw[:, :D], "packed observation of player %d" % observer)


@pytest.mark.parametrize("name", ["full2", "full5", "small4", "wide8"])
def test_no_write_outside_the_output_buffers(name):
  """compute-sanitizer is not available on this pool: every output buffer of every step
  flavour sits between canary regions inside one allocation, with a ragged number of games;
  the canaries must survive."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 32 * 9 + 5
  p = CONFIGS[name]
  b = ph.HanabiBatch(p, n, common.seeds_for(n, base=17))
  dev = torch.device("cuda:0")
  D, A, W = b.obs_size, b.max_moves, b.packed_words
  K = (D + 31) // 32 * 32
  gap = 4096
  sizes = dict(obs=n * D, mask=n * A, cur=4 * n, rew=4 * n, done=n, score=4 * n, act=4 * n,
               bits=4 * n * W, x16=2 * n * K, x32=4 * n * K, flags=n)
  total, offs = gap, {}
  for k, v in sizes.items():
    offs[k] = total
    total += (v + 255) // 256 * 256 + gap
  arena = torch.full((total,), 0xA5, dtype=torch.uint8, device=dev)
  view = lambda k, dt, shape: arena[offs[k]:offs[k] + sizes[k]].view(dt).view(shape)
  obs, mask = view("obs", torch.uint8, (n, D)), view("mask", torch.uint8, (n, A))
  cur, rew = view("cur", torch.int32, (n,)), view("rew", torch.int32, (n,))
  done, score = view("done", torch.uint8, (n,)), view("score", torch.int32, (n,))
  act, bits = view("act", torch.int32, (n,)), view("bits", torch.int32, (n, W))
  x16, x32 = view("x16", torch.bfloat16, (n, K)), view("x32", torch.float32, (n, K))
  b.reset()
  b.observe(obs, mask, cur)
  for t in range(30):
    ph.select_action(None, mask, 1.0, seed=4, step=t, out=act)
    kind = t % 6
    if kind == 0:
      b.step_encode(act, rew, done, obs, mask, cur, scores=score)
    elif kind == 1:
      b.step_encode_random(act, 4, t, rew, done, obs, mask, cur, scores=score)
      b.legal_mask(mask)
    elif kind == 2:
      b.step_encode_packed(act, rew, done, bits, mask, cur, scores=score)
    elif kind == 3:
      b.step_encode_net(act, x16, rew, done, obs_bits=bits, mask=mask, cur_player=cur, scores=score)
    elif kind == 4:
      b.step_encode_net(act, x32, rew, done, mask=mask, cur_player=cur)
    else:
      b.step(act, rew, done, score)
      b.observe(obs, mask, cur)
  b.observe_packed(bits, mask, cur)
  b.observe_net(x16, obs_bits=bits)
  ph.expand_packed(bits, D, x32)
  keep = torch.ones(total, dtype=torch.bool, device=dev)
  for k, v in sizes.items():
    keep[offs[k]:offs[k] + v] = False
  assert bool((arena[keep] == 0xA5).all()), "a kernel wrote outside its output buffers"
