"""Full-size checks (2^20 games, the bench configuration) through properties that do
not need a full-size CPU run:
  * sampled parity: games are independent, so game i of the big batch must equal a
    stand-alone CPU game with seed_i driven by the same actions -- checked for a
    random sample of game indices, every step;
  * conservation: deck + hands + discards + fireworks hold every card exactly once;
  * structure of the encoding: one-hot cards, thermometers consistent with the state;
  * shard invariance: two half batches produce the bytes of the one big batch
    (what the multi-GPU layout relies on);
  * idempotence of observe and of the state save/restore round trip.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import common

pytestmark = pytest.mark.gpu
N = 1 << 20
STEPS = 48


def _run(ph, torch, seeds, steps, sample=None, device=0):
  n = len(seeds)
  dev = torch.device("cuda", device)
  b = ph.HanabiBatch(dict(players=2), n, seeds, device=device)
  obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev)
  mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
  cur = torch.empty(n, dtype=torch.int32, device=dev)
  rew = torch.empty(n, dtype=torch.int32, device=dev)
  done = torch.empty(n, dtype=torch.uint8, device=dev)
  act = torch.empty(n, dtype=torch.int32, device=dev)
  b.reset()
  b.observe(obs, mask, cur)
  trace = []
  for t in range(steps):
    ph.select_action(None, mask, 1.0, seed=7, step=t, out=act)
    if sample is not None:
      trace.append((act[sample].cpu().numpy(), obs[sample].cpu().numpy(), mask[sample].cpu().numpy()))
    b.step_encode(act, rew, done, obs, mask, cur)
    if sample is not None:
      trace[-1] += (rew[sample].cpu().numpy(), done[sample].cpu().numpy())
  return b, obs, mask, cur, trace


def test_full_size_sampled_parity_and_invariants():
  import torch
  from hanabi_b200 import pyhanabi as ph
  rng = np.random.default_rng(0)
  seeds = (np.arange(N, dtype=np.int64) + 1).astype(np.int32)
  idx = np.sort(rng.choice(N, 768, replace=False))
  idx[:3] = [0, 1, 31]
  idx[-3:] = [N - 33, N - 2, N - 1]
  idx = np.unique(idx)
  sample = torch.as_tensor(idx, device="cuda:0")
  b, obs, mask, cur, trace = _run(ph, torch, seeds, STEPS, sample=sample)

  # -- sampled parity against stand-alone CPU games
  cpu = O.CpuBatch("ref" if O.have_ref() else "oracle", dict(players=2), len(idx), seeds[idx])
  cpu.reset()
  for t, (a, o, m, r, d) in enumerate(trace):
    common.assert_same(cpu.encode(), o, "sampled obs @%d" % t)
    common.assert_same(cpu.legal_mask(), m, "sampled mask @%d" % t)
    rr, dd, _, illegal = cpu.step(a)
    assert not illegal.any()
    common.assert_same(rr, r, "sampled reward @%d" % t)
    common.assert_same(dd, d, "sampled done @%d" % t)
  common.assert_same(cpu.encode(), obs[sample].cpu().numpy(), "sampled final obs")

  # -- conservation of cards and consistency of the whole batch
  dump = b.unpack_state()
  D = O.DUMP
  hands = dump[:, D["handsize"]:D["handsize"] + 2].sum(dim=1)
  deck = dump[:, D["deck"]]
  deckc = dump[:, D["deckcount"]:D["deckcount"] + 25].sum(dim=1)
  disc = dump[:, D["discardcount"]:D["discardcount"] + 25].sum(dim=1)
  fw = dump[:, D["fireworks"]:D["fireworks"] + 5].sum(dim=1)
  assert bool((deck == deckc).all())
  assert bool((deck + hands + disc + fw == 50).all())
  assert bool((dump[:, D["life"]] >= 1).all()) and bool((dump[:, D["end"]] == 0).all())  # auto-reset
  # per card type: deck + hands + discards + fireworks == multiplicity
  per_type = dump[:, D["deckcount"]:D["deckcount"] + 25] + dump[:, D["discardcount"]:D["discardcount"] + 25]
  cards = dump[:, D["cards"]:D["cards"] + 16].long()
  onehot = torch.zeros((N, 26), dtype=torch.int32, device=dump.device)
  onehot.scatter_add_(1, torch.where(cards >= 0, cards, torch.full_like(cards, 25)),
                      torch.ones_like(cards, dtype=torch.int32))
  played = (torch.arange(5, device=dump.device)[None, None, :] <
            dump[:, D["fireworks"]:D["fireworks"] + 5][:, :, None]).reshape(N, 25).int()
  mult = torch.tensor([3, 2, 2, 2, 1] * 5, dtype=torch.int32, device=dump.device)
  assert bool((per_type + onehot[:, :25] + played == mult[None, :]).all())

  # -- structure of the encoding (canonical_encoders.cc section layout, 2p Full)
  o = obs
  partner = o[:, :125].reshape(N, 5, 25).sum(dim=2)
  partner_n = dump.gather(1, (D["handsize"] + 1 - cur.long())[:, None])[:, 0]
  assert bool((partner.sum(dim=1) == partner_n).all()) and bool((partner <= 1).all())
  assert bool((o[:, 127:167].sum(dim=1) == deck).all())                       # deck thermometer
  thermo = o[:, 127:167].int()
  assert bool((thermo[:, 1:] <= thermo[:, :-1]).all())                         # sorted: ones first
  assert bool((o[:, 192:200].sum(dim=1) == dump[:, D["info"]]).all())
  assert bool((o[:, 200:203].sum(dim=1) == dump[:, D["life"]]).all())
  assert bool((o[:, 203:253].sum(dim=1) == disc).all())
  assert bool((mask.sum(dim=1) >= 1).all())

  # -- idempotence: observing again changes nothing; save/restore reproduces the state
  obs2 = torch.empty_like(obs); mask2 = torch.empty_like(mask); cur2 = torch.empty_like(cur)
  b.observe(obs2, mask2, cur2)
  assert torch.equal(obs, obs2) and torch.equal(mask, mask2) and torch.equal(cur, cur2)


def test_full_size_shard_invariance():
  """Two half batches (what two GPUs would hold) == one big batch, byte for byte."""
  import hashlib
  import torch
  from hanabi_b200 import pyhanabi as ph
  n = 1 << 18
  seeds = (np.arange(n, dtype=np.int64) + 1).astype(np.int32)
  _, obs, mask, cur, _ = _run(ph, torch, seeds, 24)
  whole = hashlib.sha256(obs.cpu().numpy().tobytes() + mask.cpu().numpy().tobytes()).hexdigest()
  parts = []
  for lo, hi in ((0, n // 2), (n // 2, n)):
    # the policy kernel is keyed by the game's index within its batch, so replay the
    # big batch's actions instead of re-drawing them: run the half with recorded actions
    parts.append((lo, hi))
  # record actions of the big batch, then replay them on the halves
  dev = torch.device("cuda:0")
  big = ph.HanabiBatch(dict(players=2), n, seeds)
  halves = [ph.HanabiBatch(dict(players=2), hi - lo, seeds[lo:hi]) for lo, hi in parts]
  bo = torch.empty((n, big.obs_size), dtype=torch.uint8, device=dev)
  bm = torch.empty((n, big.max_moves), dtype=torch.uint8, device=dev)
  bc = torch.empty(n, dtype=torch.int32, device=dev)
  act = torch.empty(n, dtype=torch.int32, device=dev)
  rew = torch.empty(n, dtype=torch.int32, device=dev)
  dn = torch.empty(n, dtype=torch.uint8, device=dev)
  big.reset(); big.observe(bo, bm, bc)
  ho = [torch.empty((hi - lo, big.obs_size), dtype=torch.uint8, device=dev) for lo, hi in parts]
  hm = [torch.empty((hi - lo, big.max_moves), dtype=torch.uint8, device=dev) for lo, hi in parts]
  hc = [torch.empty(hi - lo, dtype=torch.int32, device=dev) for lo, hi in parts]
  for h, o_, m_, c_ in zip(halves, ho, hm, hc):
    h.reset(); h.observe(o_, m_, c_)
  for t in range(24):
    ph.select_action(None, bm, 1.0, seed=7, step=t, out=act)
    big.step_encode(act, rew, dn, bo, bm, bc)
    for (lo, hi), h, o_, m_, c_ in zip(parts, halves, ho, hm, hc):
      a = act[lo:hi].contiguous()
      r = torch.empty(hi - lo, dtype=torch.int32, device=dev)
      d = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
      h.step_encode(a, r, d, o_, m_, c_)
      assert torch.equal(r, rew[lo:hi]) and torch.equal(d, dn[lo:hi])
  assert torch.equal(torch.cat(ho), bo) and torch.equal(torch.cat(hm), bm)
  assert hashlib.sha256(bo.cpu().numpy().tobytes() + bm.cpu().numpy().tobytes()).hexdigest() == whole


def _fused_rollout(ph, torch, params, seeds, base, steps, seed=7, init_step=1 << 40, device=0):
  """Fused random-policy rollout (BatchStepEncodeRandom) of one shard whose game 0 has the
  global index `base`; returns the final output tensors and the pending actions."""
  n = len(seeds)
  dev = torch.device("cuda", device)
  with torch.cuda.device(dev):
    b = ph.HanabiBatch(params, n, seeds, device=device)
    b.set_game_index_base(base)
    obs = torch.empty((n, b.obs_size), dtype=torch.uint8, device=dev)
    mask = torch.empty((n, b.max_moves), dtype=torch.uint8, device=dev)
    cur = torch.empty(n, dtype=torch.int32, device=dev)
    rew = torch.empty(n, dtype=torch.int32, device=dev)
    done = torch.empty(n, dtype=torch.uint8, device=dev)
    act = torch.empty(n, dtype=torch.int32, device=dev)
    b.reset()
    b.observe(obs, mask, cur)
    b.random_legal_actions(act, seed, init_step)
    for t in range(steps):
      b.step_encode_random(act, seed, t, rew, done, obs, mask, cur)
    torch.cuda.synchronize(dev)
  return obs, mask, cur, rew, done, act


@pytest.mark.parametrize("name", ["full2", "small4"])
def test_sharded_fused_rollout_equals_unsharded(name):
  """Shards with BatchSetGameIndexBase reproduce the unsharded fused rollout byte for
  byte (the policy is keyed by the GLOBAL game index), and sampled games equal the CPU
  reference driven by the numpy twin of the policy -- what bench.py checks on every rank."""
  import torch
  import bench
  from hanabi_b200 import pyhanabi as ph
  params = common.CONFIGS[name]
  n, steps = 1 << 14, 40
  seeds = common.seeds_for(n)
  whole = _fused_rollout(ph, torch, params, seeds, 0, steps)
  cuts = [0, 4096, 4096 + 32 * 101, n]
  parts = [_fused_rollout(ph, torch, params, seeds[lo:hi], lo, steps) for lo, hi in zip(cuts, cuts[1:])]
  for k in range(len(whole)):
    assert torch.equal(torch.cat([p[k] for p in parts]), whole[k])
  # the select kernel with epsilon >= 1 is the same draw for base 0
  twin = ph.select_action(None, whole[1], 1.0, seed=7, step=steps - 1)
  assert torch.equal(twin, whole[5])
  # CPU reference on sampled games of the last shard (global indices)
  lo = cuts[-2]
  local = np.array([0, 1, 31, 32, 777, n - lo - 1])
  kind = "ref" if O.have_ref() else "oracle"
  cpu = O.CpuBatch(kind, params, len(local), seeds[lo + local])
  cpu.reset()
  g = (lo + local).astype(np.int64)
  act = bench.policy_draw(cpu.legal_mask(), 7, 1 << 40, g)
  for t in range(steps):
    _, _, _, illegal = cpu.step(act)
    assert not illegal.any()
    act = bench.policy_draw(cpu.legal_mask(), 7, t, g)
  last = parts[-1]
  idx = torch.as_tensor(local, device="cuda:0")
  assert (last[0][idx].cpu().numpy() == cpu.encode()).all()
  assert (last[1][idx].cpu().numpy() == cpu.legal_mask()).all()
  assert (last[2][idx].cpu().numpy() == cpu.cur_player()).all()
  assert (last[5][idx].cpu().numpy() == act).all()


def test_two_gpu_shards_equal_one_gpu_run():
  """On a box with two GPUs: the halves stepped on cuda:0 and cuda:1 equal the one-GPU run."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  if torch.cuda.device_count() < 2:
    pytest.skip("needs two GPUs")
  params = common.CONFIGS["full2"]
  n, steps = 1 << 15, 32
  seeds = common.seeds_for(n)
  whole = _fused_rollout(ph, torch, params, seeds, 0, steps, device=0)
  a = _fused_rollout(ph, torch, params, seeds[:n // 2], 0, steps, device=0)
  b = _fused_rollout(ph, torch, params, seeds[n // 2:], n // 2, steps, device=1)
  for k in range(len(whole)):
    assert torch.equal(torch.cat([a[k], b[k].to("cuda:0")]), whole[k])


@pytest.mark.parametrize("name,steps", [("full2", 3000), ("small2", 3000), ("full4", 2000)])
def test_long_soak_lockstep(name, steps):
  """Soak: 1 024 games in lock-step with the compiled reference for thousands of steps -- several
  hundred episodes per game on the small rule set, every game's mt19937 block regenerated many
  times over, the re-deal queues of every CTA exercised at all fill levels (serial path for two
  players, job path with lazy regeneration for four)."""
  n = 1024
  seeds = common.seeds_for(n, base=424242)
  params = common.CONFIGS[name]
  cpu = O.CpuBatch("ref" if O.have_ref() else "oracle", params, n, seeds)
  gpu = common.GpuBatch(params, n, seeds)
  cpu.reset(); gpu.reset()
  eps, _ = common.lockstep(cpu, gpu, steps, np.random.default_rng(99), False, params, dump_every=250,
                           check_all_observers_every=500)
  assert eps > n * steps // 40
