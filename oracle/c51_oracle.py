"""oracle/c51_oracle.py -- TEST INFRASTRUCTURE ONLY.

fp32 restatement (numpy) of the two Rainbow pieces on the hot path, following the
reference op by op.  Pinned (tests/test_c51_golden.py) by tests/golden/c51_golden.npz:
the outputs of the reference's own `project_distribution`, `_reshape_networks` and
`_build_target_distribution` (agents/rainbow_copy/rainbow_agent.py:163-213, 252-404)
executed unmodified on a numpy stand-in for the TensorFlow ops they call
(tests/golden/make_c51_golden.py; TensorFlow itself is not installed), plus the worked
example of the reference docstring (rainbow_agent.py:262-266,400-401,
tests/golden/c51_docstring.json) and an independent floor/ceil scatter formulation.
Tolerance 1e-5 relative in fp32 as north_star states.
"""
import numpy as np


def project_distribution(supports, weights, target_support):
  """agents/rainbow_copy/rainbow_agent.py:291-404, fp32."""
  s = np.asarray(supports, np.float32)
  w = np.asarray(weights, np.float32)
  z = np.asarray(target_support, np.float32)
  delta_z = z[1] - z[0]                                        # :293-295
  v_min, v_max = z[0], z[-1]                                   # :330
  clipped = np.clip(s, v_min, v_max)[:, None, :]               # :336  [B,1,N]
  numerator = np.abs(clipped - z[None, :, None])               # :370  [B,N,N]
  quotient = np.float32(1) - numerator / delta_z               # :371
  clipped_q = np.clip(quotient, 0, 1)                          # :383
  inner = clipped_q * w[:, None, :]                            # :399
  out = np.zeros((s.shape[0], z.shape[0]), np.float32)
  for j in range(z.shape[0]):                                  # :402 reduce_sum over j, index order
    out = out + inner[:, :, j]
  return out


def project_distribution_scatter(supports, weights, target_support):
  """Independent formulation (Bellemare et al. 2017, Algorithm 1): each source
  atom spreads its mass to the two neighbouring target atoms.  float64."""
  s = np.asarray(supports, np.float64)
  w = np.asarray(weights, np.float64)
  z = np.asarray(target_support, np.float64)
  dz = z[1] - z[0]
  B, N = s.shape
  out = np.zeros((B, N))
  b = (np.clip(s, z[0], z[-1]) - z[0]) / dz
  lo = np.floor(b).astype(int)
  hi = np.ceil(b).astype(int)
  for i in range(B):
    for j in range(N):
      if lo[i, j] == hi[i, j]:
        out[i, lo[i, j]] += w[i, j]
      else:
        out[i, lo[i, j]] += w[i, j] * (hi[i, j] - b[i, j])
        out[i, hi[i, j]] += w[i, j] * (b[i, j] - lo[i, j])
  return out


def softmax(x):
  x = np.asarray(x, np.float32)
  e = np.exp(x - x.max(axis=-1, keepdims=True))
  return e / e.sum(axis=-1, keepdims=True)


def c51_q_values(logits, support):
  """rainbow_agent.py:163-170: q = sum_i support_i * softmax(logits)_i."""
  return (softmax(logits) * np.asarray(support, np.float32)).sum(axis=-1)


def masked_argmax(q, mask):
  """rainbow_agent.py:172 / dqn_agent.py:200: argmax(q + legal), legal in {0,-inf}."""
  legal = np.where(np.asarray(mask) != 0, np.float32(0), np.float32(-np.inf))
  return np.argmax(np.asarray(q, np.float32) + legal, axis=1)


def target_distribution(rewards, terminals, next_logits, next_mask, support, cumulative_gamma):
  """rainbow_agent.py:181-213."""
  z = np.asarray(support, np.float32)
  r = np.asarray(rewards, np.float32)[:, None]
  g = np.float32(cumulative_gamma) * (np.float32(1) - np.asarray(terminals, np.float32))[:, None]
  target_support = r + g * z[None, :]
  probs = softmax(next_logits)
  a = masked_argmax((probs * z).sum(axis=2), next_mask)
  return project_distribution(target_support, probs[np.arange(len(a)), a], z), a
