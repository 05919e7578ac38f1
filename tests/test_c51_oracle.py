"""CPU tests pinning the C51 restatement (oracle/c51_oracle.py)."""
import json
import os

import numpy as np

from oracle import c51_oracle as C

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_docstring_known_answer():
  """agents/rainbow_copy/rainbow_agent.py:262-266,400-401."""
  kat = json.load(open(os.path.join(GOLDEN, "c51_docstring.json")))
  out = C.project_distribution(kat["supports"], kat["weights"], kat["target_support"])
  np.testing.assert_allclose(out, np.asarray(kat["projection"], np.float32), rtol=1e-5, atol=1e-6)


def test_against_independent_scatter_formulation():
  rng = np.random.default_rng(0)
  z = np.linspace(-25, 25, 51).astype(np.float32)
  for _ in range(5):
    B = 32
    r = rng.integers(-3, 2, size=(B, 1)).astype(np.float32)
    g = (0.99 ** rng.integers(1, 6, size=(B, 1))).astype(np.float32) * (rng.random((B, 1)) < 0.9)
    s = (r + g * z[None, :]).astype(np.float32)
    w = C.softmax(rng.normal(size=(B, 51)).astype(np.float32))
    a = C.project_distribution(s, w, z)
    b = C.project_distribution_scatter(s, w, z)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(a.sum(axis=1), 1.0, rtol=1e-5)


def test_masked_argmax_ties_take_first_index():
  q = np.asarray([[1.0, 3.0, 3.0, 2.0], [5.0, 1.0, 1.0, 1.0]], np.float32)
  mask = np.asarray([[1, 1, 1, 1], [0, 1, 1, 0]], np.uint8)
  assert C.masked_argmax(q, mask).tolist() == [1, 1]


def test_linearly_decaying_epsilon_schedule():
  """agents/rainbow_copy/dqn_agent.py:44-59 (values worked by hand from the formula)."""
  from hanabi_b200.actor import linearly_decaying_epsilon as eps
  assert eps(1000, 0, 500, 0.02) == 1.0            # warm-up
  assert eps(1000, 500, 500, 0.02) == 1.0          # decay starts
  assert abs(eps(1000, 1000, 500, 0.02) - 0.51) < 1e-12
  assert abs(eps(1000, 1500, 500, 0.02) - 0.02) < 1e-12
  assert eps(1000, 10**6, 500, 0.02) == 0.02
