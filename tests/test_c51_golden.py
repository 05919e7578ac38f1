"""C51 rows (a11 q-values / argmax, a12 project_distribution, the fused target distribution)
against tests/golden/c51_golden.npz, produced by executing the reference's own
`project_distribution`, `_reshape_networks` and `_build_target_distribution`
(agents/rainbow_copy/rainbow_agent.py:163-213, 252-404) on a numpy stand-in for the
TensorFlow ops they call (tests/golden/make_c51_golden.py).

Tolerance (north_star): 1e-5 relative in fp32.  Probabilities are compared with
rtol 1e-5 + atol 1e-6 (values in [0, 1]; two fp32 summation orders of 51 terms differ by
~1e-7); q-values, which are sums of 51 terms of magnitude up to v_max, with
rtol 1e-5 + atol 2e-6 * v_max (5e-5 for v_max = 25: the sum's own fp32 rounding).  Rows are compared where the greedy action agrees -- it must
agree everywhere the two best q-values are further apart than that tolerance.
"""
import os

import numpy as np
import pytest

from oracle import c51_oracle as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c51_golden.npz")
PROJ = ["p51", "p51_wide", "p11", "p2"]
TARGETS = ["t51", "t51_n3", "t7"]
RTOL, ATOL_P = 1e-5, 1e-6


@pytest.fixture(scope="module")
def golden():
  return np.load(GOLDEN)


def _clear_winner(q, legal, tol):
  """Rows whose best legal q-value beats the runner-up by more than tol."""
  v = np.sort(np.where(np.isfinite(legal), q, -np.inf), axis=1)
  return (v[:, -1] - v[:, -2]) > tol


@pytest.mark.parametrize("name", PROJ)
def test_projection_restatement_equals_reference_code(golden, name):
  g = golden
  got = C.project_distribution(g[name + "/supports"], g[name + "/weights"], g[name + "/target_support"])
  np.testing.assert_allclose(got, g[name + "/projection"], rtol=RTOL, atol=ATOL_P)


@pytest.mark.parametrize("name", TARGETS)
def test_target_restatement_equals_reference_code(golden, name):
  g = golden
  z = np.linspace(-g[name + "/vmax"], g[name + "/vmax"], int(g[name + "/cfg"][2])).astype(np.float32)
  mask = np.isfinite(g[name + "/next_legal"]).astype(np.uint8)
  got, a = C.target_distribution(g[name + "/rewards"], g[name + "/terminals"], g[name + "/next_logits"],
                                 mask, z, float(g[name + "/cumulative_gamma"]))
  vmax = float(g[name + "/vmax"])
  clear = _clear_winner(g[name + "/next_q"], g[name + "/next_legal"], 4e-5 * vmax)
  assert clear.mean() > 0.8 and (a[clear] == g[name + "/next_argmax"][clear]).all()
  same = a == g[name + "/next_argmax"]
  np.testing.assert_allclose(got[same], g[name + "/target"][same], rtol=RTOL, atol=ATOL_P)
  np.testing.assert_allclose(C.c51_q_values(g[name + "/next_logits"], z), g[name + "/next_q"],
                             rtol=RTOL, atol=2e-6 * vmax)
  act_mask = np.isfinite(g[name + "/act_legal"]).astype(np.uint8)[None]
  assert C.masked_argmax(g[name + "/act_q"][None], act_mask)[0] == g[name + "/act_argmax"]


def _t(a):
  import torch
  return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROJ)
def test_gpu_projection_equals_reference_code(golden, name):
  from hanabi_b200 import pyhanabi as ph
  g = golden
  got = ph.project_distribution(_t(g[name + "/supports"]), _t(g[name + "/weights"]),
                                _t(g[name + "/target_support"])).cpu().numpy()
  np.testing.assert_allclose(got, g[name + "/projection"], rtol=RTOL, atol=ATOL_P)


@pytest.mark.gpu
@pytest.mark.parametrize("name", TARGETS)
def test_gpu_c51_kernels_equal_reference_code(golden, name):
  """BatchC51QValues, BatchSelectAction (C51 head, epsilon 0) and BatchC51TargetDistribution
  against what the reference's graph-building code computes for the same inputs."""
  from hanabi_b200 import pyhanabi as ph
  g = golden
  B, A, N = [int(x) for x in g[name + "/cfg"]]
  vmax = float(g[name + "/vmax"])
  z = np.linspace(-vmax, vmax, N).astype(np.float32)
  logits, legal = g[name + "/next_logits"], g[name + "/next_legal"]
  mask = np.isfinite(legal).astype(np.uint8)
  q, greedy = ph.c51_q_values(_t(logits), _t(z), _t(mask))
  q, greedy = q.cpu().numpy(), greedy.cpu().numpy()
  np.testing.assert_allclose(q, g[name + "/next_q"], rtol=RTOL, atol=2e-6 * vmax)
  clear = _clear_winner(g[name + "/next_q"], legal, 4e-5 * vmax)
  assert (greedy[clear] == g[name + "/next_argmax"][clear]).all()
  a = ph.select_action(_t(logits), _t(mask), 0.0, seed=1, step=0, support=_t(z)).cpu().numpy()
  assert (a == greedy).all()
  got, arg = ph.c51_target_distribution(_t(g[name + "/rewards"]), _t(g[name + "/terminals"]), _t(logits),
                                        _t(mask), _t(z), float(g[name + "/cumulative_gamma"]))
  got, arg = got.cpu().numpy(), arg.cpu().numpy()
  assert (arg[clear] == g[name + "/next_argmax"][clear]).all()
  same = arg == g[name + "/next_argmax"]
  assert same.mean() > 0.8
  np.testing.assert_allclose(got[same], g[name + "/target"][same], rtol=RTOL, atol=ATOL_P)
  # the acting head: one row
  act_mask = np.isfinite(g[name + "/act_legal"]).astype(np.uint8)[None]
  q1, g1 = ph.c51_q_values(_t(g[name + "/act_logits"][None]), _t(z), _t(act_mask))
  np.testing.assert_allclose(q1.cpu().numpy()[0], g[name + "/act_q"], rtol=RTOL, atol=2e-6 * vmax)
  assert int(g1[0]) == int(g[name + "/act_argmax"])
