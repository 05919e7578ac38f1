"""GPU tests for the fused actor/learner kernels (select action, C51)."""
import json
import os

import numpy as np
import pytest

from oracle import c51_oracle as C

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5  # north_star: C51 projection within 1e-5 relative in fp32


def _t(a, dtype=None):
  import torch
  return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


def test_project_distribution_docstring_known_answer():
  from hanabi_b200 import pyhanabi as ph
  kat = json.load(open(os.path.join(GOLDEN, "c51_docstring.json")))
  out = ph.project_distribution(_t(np.asarray(kat["supports"], np.float32)),
                                _t(np.asarray(kat["weights"], np.float32)),
                                _t(np.asarray(kat["target_support"], np.float32)))
  np.testing.assert_allclose(out.cpu().numpy(), np.asarray(kat["projection"], np.float32),
                             rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("B", [1, 32, 257])
def test_project_distribution_matches_restatement(B):
  from hanabi_b200 import pyhanabi as ph
  rng = np.random.default_rng(B)
  z = np.linspace(-25, 25, 51).astype(np.float32)
  r = rng.integers(-24, 2, size=(B, 1)).astype(np.float32)
  g = (0.99 ** rng.integers(1, 6, size=(B, 1))).astype(np.float32) * (rng.random((B, 1)) < 0.9)
  s = (r + g * z[None, :]).astype(np.float32)
  w = C.softmax(rng.normal(size=(B, 51)).astype(np.float32) * 3)
  want = C.project_distribution(s, w, z)
  got = ph.project_distribution(_t(s), _t(w), _t(z)).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-7)


def test_select_action_greedy_matches_masked_argmax():
  from hanabi_b200 import pyhanabi as ph
  rng = np.random.default_rng(1)
  n, A = 5000, 20
  q = rng.normal(size=(n, A)).astype(np.float32)
  q[:, 3] = q[:, 7]  # ties: first maximal index wins
  mask = (rng.random((n, A)) < 0.5).astype(np.uint8)
  mask[np.arange(n), rng.integers(0, A, n)] = 1
  got = ph.select_action(_t(q), _t(mask), 0.0, seed=7, step=0).cpu().numpy()
  assert (got == C.masked_argmax(q, mask)).all()


def test_select_action_c51_greedy_and_q_values():
  from hanabi_b200 import pyhanabi as ph
  rng = np.random.default_rng(2)
  n, A, N = 700, 20, 51
  logits = rng.normal(size=(n, A, N)).astype(np.float32) * 2
  z = np.linspace(-25, 25, N).astype(np.float32)
  mask = (rng.random((n, A)) < 0.6).astype(np.uint8)
  mask[:, 0] = 1
  q_want = C.c51_q_values(logits, z)
  q_got, greedy = ph.c51_q_values(_t(logits), _t(z), _t(mask))
  np.testing.assert_allclose(q_got.cpu().numpy(), q_want, rtol=1e-5, atol=5e-5)  # 2e-6 * v_max
  a = ph.select_action(_t(logits), _t(mask), 0.0, seed=1, step=0, support=_t(z)).cpu().numpy()
  assert (a == greedy.cpu().numpy()).all()
  # the device q-values decide the argmax; compare on the device values to avoid fp ties
  assert (a == C.masked_argmax(q_got.cpu().numpy(), mask)).all()
  agree = (a == C.masked_argmax(q_want, mask)).mean()
  assert agree > 0.999


def test_select_action_random_is_uniform_over_legal_moves():
  from hanabi_b200 import pyhanabi as ph
  n, A = 200000, 11
  mask = np.zeros((n, A), np.uint8)
  mask[:, [1, 4, 5, 9]] = 1
  a = ph.select_action(None, _t(mask), 1.0, seed=3, step=5).cpu().numpy()
  counts = np.bincount(a, minlength=A)
  assert counts[[0, 2, 3, 6, 7, 8, 10]].sum() == 0
  assert np.abs(counts[[1, 4, 5, 9]] / n - 0.25).max() < 0.01
  b = ph.select_action(None, _t(mask), 1.0, seed=3, step=6).cpu().numpy()
  assert (a != b).mean() > 0.5  # a new step draws new actions
  c = ph.select_action(None, _t(mask), 1.0, seed=3, step=5).cpu().numpy()
  assert (a == c).all()         # counter-based: reproducible
  # epsilon = 0.3: about 30 % + (chance agreement) differ from greedy
  q = np.tile(np.arange(A, dtype=np.float32), (n, 1))
  e = ph.select_action(_t(q), _t(mask), 0.3, seed=4, step=0).cpu().numpy()
  frac_non_greedy = (e != 9).mean()
  assert abs(frac_non_greedy - 0.3 * 0.75) < 0.01


def test_c51_target_distribution_fused():
  from hanabi_b200 import pyhanabi as ph
  rng = np.random.default_rng(3)
  B, A, N = 64, 20, 51
  z = np.linspace(-25, 25, N).astype(np.float32)
  logits = rng.normal(size=(B, A, N)).astype(np.float32) * 2
  mask = (rng.random((B, A)) < 0.5).astype(np.uint8)
  mask[:, 2] = 1
  rewards = rng.integers(-5, 2, size=B).astype(np.float32)
  terminals = (rng.random(B) < 0.2).astype(np.uint8)
  want, a_want = C.target_distribution(rewards, terminals, logits, mask, z, 0.99 ** 5)
  got, a_got = ph.c51_target_distribution(_t(rewards), _t(terminals), _t(logits), _t(mask), _t(z),
                                          0.99 ** 5)
  assert (a_got.cpu().numpy() == a_want).mean() > 0.98
  same = a_got.cpu().numpy() == a_want
  np.testing.assert_allclose(got.cpu().numpy()[same], want[same], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["full2", "small2", "full4", "wide8dealt"])
def test_fused_random_policy_equals_the_select_kernel(name):
  """BatchStepEncodeRandom leaves in `actions` exactly what BatchSelectAction (epsilon >= 1,
  same seed/step) picks from the mask of the same call; the rollout it drives is legal."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  from tests import common
  n = 5000  # ragged last tile
  params = common.CONFIGS[name]
  seeds = common.seeds_for(n, base=3)
  a, b = ph.HanabiBatch(params, n, seeds), ph.HanabiBatch(params, n, seeds)
  dev = torch.device("cuda:0")
  D, A = a.obs_size, a.max_moves
  mk = lambda *shape, dt: torch.zeros(shape, dtype=dt, device=dev)
  obs_a, obs_b = mk(n, D, dt=torch.uint8), mk(n, D, dt=torch.uint8)
  mask_a, mask_b = mk(n, A, dt=torch.uint8), mk(n, A, dt=torch.uint8)
  cur_a, cur_b = mk(n, dt=torch.int32), mk(n, dt=torch.int32)
  rew_a, rew_b = mk(n, dt=torch.int32), mk(n, dt=torch.int32)
  done_a, done_b = mk(n, dt=torch.uint8), mk(n, dt=torch.uint8)
  act_a, act_b = mk(n, dt=torch.int32), mk(n, dt=torch.int32)
  for x, o, m, c in ((a, obs_a, mask_a, cur_a), (b, obs_b, mask_b, cur_b)):
    x.reset()
    x.observe(o, m, c)
  ph.select_action(None, mask_a, 1.0, seed=11, step=0, out=act_a)
  act_b.copy_(act_a)
  run = a.bind_step_encode_random(act_a, 11, rew_a, done_a, obs_a, mask_a, cur_a)
  for t in range(1, 40):
    run(t)                                             # fused: step + next action (in place)
    b.step_encode(act_b, rew_b, done_b, obs_b, mask_b, cur_b)
    ph.select_action(None, mask_b, 1.0, seed=11, step=t, out=act_b)
    assert torch.equal(act_a, act_b), "actions differ at step %d" % t
    assert torch.equal(obs_a, obs_b) and torch.equal(mask_a, mask_b)
    assert torch.equal(rew_a, rew_b) and torch.equal(done_a, done_b)
    legal = mask_a.gather(1, act_a.long().clamp(min=0).view(-1, 1)).view(-1)
    assert bool(((legal == 1) | (act_a < 0)).all())
  # separate output buffer: `actions` stays untouched
  nxt = mk(n, dt=torch.int32)
  before = act_a.clone()
  a.step_encode_random(act_a, 11, 99, rew_a, done_a, obs_a, mask_a, cur_a, random_actions=nxt)
  assert torch.equal(act_a, before)
  assert torch.equal(nxt, ph.select_action(None, mask_a, 1.0, seed=11, step=99))


@pytest.mark.parametrize("dtype_name", ["float32", "bfloat16", "float16"])
def test_expand_packed_and_c51_select_on_gemm_outputs(dtype_name):
  """BatchExpandPacked gives the byte observation as 0/1 in the GEMM's input type with a
  zero-padded K; BatchSelectActionC51 on a padded, typed logits matrix picks what
  BatchSelectAction picks on the same values in float32."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  from tests import common
  dt = getattr(torch, dtype_name)
  n = 3001
  b = ph.HanabiBatch(common.CONFIGS["full2"], n, common.seeds_for(n, base=9))
  dev = torch.device("cuda:0")
  D, A, W = b.obs_size, b.max_moves, b.packed_words
  obs = torch.zeros((n, D), dtype=torch.uint8, device=dev)
  bits = torch.zeros((n, W), dtype=torch.int32, device=dev)
  mask = torch.zeros((n, A), dtype=torch.uint8, device=dev)
  b.reset()
  b.observe(obs, mask)
  b.observe_packed(bits)
  K = 672
  x = torch.full((n, K), 7, dtype=dt, device=dev)
  ph.expand_packed(bits, D, x)
  assert torch.equal(x[:, :D].float(), obs.float()) and not x[:, D:].any()
  # a wider pitch: every column from obs_size up to the pitch is zeroed
  wide = torch.full((n, K + 8), 7, dtype=dt, device=dev)  # pitch not a multiple of 32
  ph.expand_packed(bits, D, wide)
  assert torch.equal(wide[:, :D].float(), obs.float()) and not wide[:, D:].any()
  atoms = 51
  torch.manual_seed(1)
  full = (torch.randn((n, 1024), device=dev) * 2).to(dt)
  support = torch.linspace(-25.0, 25.0, atoms, device=dev)
  got = ph.select_action_c51(full, atoms, mask, 0.0, seed=5, step=3, support=support)
  ref_logits = full[:, :A * atoms].float().contiguous().view(n, A, atoms)
  want = ph.select_action(ref_logits, mask, 0.0, seed=5, step=3, support=support)
  # the two kernels subtract different maxima inside the softmax: near-ties may flip
  assert float((got == want).float().mean()) > 0.999
  legal_pick = mask.gather(1, got.long().view(-1, 1)).view(-1)
  assert bool((legal_pick == 1).all())
  # and against a torch restatement of q = sum softmax * support with the legal mask
  q = (torch.softmax(ref_logits, dim=2) * support).sum(dim=2)
  q = torch.where(mask.bool(), q, torch.full_like(q, float("-inf")))
  top = q.max(dim=1).values
  picked = q.gather(1, got.long().view(-1, 1)).view(-1)
  assert bool(((top - picked).abs() <= 1e-4 * top.abs().clamp(min=1)).all())


@pytest.mark.parametrize("dtype_name,name", [("bfloat16", "full2"), ("float32", "full2"), ("float16", "small2"),
                                             ("bfloat16", "full5"), ("bfloat16", "wide8"), ("float32", "wide7seer")])
def test_step_kernel_writes_the_network_input(dtype_name, name):
  """BatchStepEncodeNet / BatchObserveNet: the step kernel's own 0/1 rows in the GEMM's element
  type (K zero-padded) equal the byte observation of a twin batch, step after step; the packed
  rows written alongside are the ones BatchStepEncodePacked writes."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  from tests import common
  dt = getattr(torch, dtype_name)
  n = 2100  # ragged last tile
  params = common.CONFIGS[name]
  seeds = common.seeds_for(n, base=21)
  a, b = ph.HanabiBatch(params, n, seeds), ph.HanabiBatch(params, n, seeds)
  dev = torch.device("cuda:0")
  D, A, W = a.obs_size, a.max_moves, a.packed_words
  K = (D + 31) // 32 * 32
  obs = torch.zeros((n, D), dtype=torch.uint8, device=dev)
  mask_a = torch.zeros((n, A), dtype=torch.uint8, device=dev)
  mask_b = torch.zeros((n, A), dtype=torch.uint8, device=dev)
  cur_a = torch.zeros(n, dtype=torch.int32, device=dev)
  cur_b = torch.zeros(n, dtype=torch.int32, device=dev)
  rew_a, rew_b = torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev)
  done_a, done_b = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.uint8, device=dev)
  act = torch.zeros(n, dtype=torch.int32, device=dev)
  x = torch.full((n, K), 7, dtype=dt, device=dev)
  bits = torch.full((n, W), -1, dtype=torch.int32, device=dev)
  want_bits = torch.zeros((n, W), dtype=torch.int32, device=dev)
  a.reset(); b.reset()
  a.observe(obs, mask_a, cur_a)
  b.observe_net(x, obs_bits=bits, mask=mask_b, cur_player=cur_b)
  for t in range(25):
    assert torch.equal(x[:, :D].float(), obs.float()) and not x[:, D:].any(), t
    a.observe_packed(want_bits)
    assert torch.equal(bits, want_bits) and torch.equal(mask_a, mask_b) and torch.equal(cur_a, cur_b)
    ph.select_action(None, mask_a, 1.0, seed=2, step=t, out=act)
    a.step_encode(act, rew_a, done_a, obs, mask_a, cur_a)
    x.fill_(7)
    b.step_encode_net(act, x, rew_b, done_b, obs_bits=bits, mask=mask_b, cur_player=cur_b)
    assert torch.equal(rew_a, rew_b) and torch.equal(done_a, done_b)
  # a wider pitch (not a multiple of 32) and no packed rows
  wide = torch.full((n, K + 8), 7, dtype=dt, device=dev)
  b.observe_net(wide)
  assert torch.equal(wide[:, :D].float(), obs.float()) and not wide[:, D:].any()


def test_masked_categorical_policy_head():
  """BatchSampleMaskedCategorical vs torch: log-probabilities equal log_softmax(logits + mask)
  at the drawn action (masked_networks.py:45-50 with mask in {0, float32.min}), draws are
  always legal, the empirical distribution over many draws matches the softmax, greedy takes
  the mode."""
  import torch
  from hanabi_b200 import pyhanabi as ph
  dev = torch.device("cuda:0")
  torch.manual_seed(3)
  n, A = 4096, 11
  logits = torch.randn((n, A), device=dev) * 2
  mask = (torch.rand((n, A), device=dev) < 0.6).to(torch.uint8)
  mask[:, 0] |= (mask.sum(dim=1) == 0).to(torch.uint8)  # at least one legal action
  fmin = torch.finfo(torch.float32).min
  ref_logp = torch.log_softmax(logits + torch.where(mask.bool(), 0.0, fmin), dim=1)
  act, logp = ph.sample_masked_categorical(logits, mask, seed=1, step=0)
  assert bool(mask.gather(1, act.long().view(-1, 1)).all())
  want = ref_logp.gather(1, act.long().view(-1, 1)).view(-1)
  assert torch.allclose(logp, want, rtol=1e-5, atol=1e-5)
  g_act, _ = ph.sample_masked_categorical(logits, mask, seed=1, step=0, greedy=True)
  assert torch.equal(g_act.long(), ref_logp.argmax(dim=1))
  # one game replicated: frequencies of 200 000 draws against its softmax
  row, mrow = logits[:1].repeat(200000, 1).contiguous(), mask[:1].repeat(200000, 1).contiguous()
  draws, _ = ph.sample_masked_categorical(row, mrow, seed=9, step=4)
  freq = torch.bincount(draws.long(), minlength=A).float() / draws.numel()
  assert float((freq - ref_logp[0].exp()).abs().max()) < 5e-3
  # bfloat16 logits with a padded pitch, and a game without legal action
  wide = torch.zeros((n, 16), dtype=torch.bfloat16, device=dev)
  wide[:, :A] = logits.to(torch.bfloat16)
  mask2 = mask.clone()
  mask2[5] = 0
  a2, lp2 = ph.sample_masked_categorical(wide, mask2, seed=1, step=0)
  assert int(a2[5]) == -1
  ok = torch.ones(n, dtype=torch.bool, device=dev)
  ok[5] = False
  ref2 = torch.log_softmax(wide[:, :A].float() + torch.where(mask2.bool(), 0.0, fmin), dim=1)
  assert torch.allclose(lp2[ok], ref2.gather(1, a2.long().clamp(min=0).view(-1, 1)).view(-1)[ok], rtol=1e-4, atol=1e-4)
