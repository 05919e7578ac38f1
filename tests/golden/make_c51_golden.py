"""Generates tests/golden/c51_golden.npz by EXECUTING the reference's own C51 code
(agents/rainbow_copy/rainbow_agent.py: `project_distribution` :36-... / 252-404,
`RainbowAgent._reshape_networks` :163-178, `RainbowAgent._build_target_distribution`
:181-213) on a numpy stand-in for the dozen TensorFlow ops those three functions call.

TensorFlow is not installed in the build container.  Round 1 therefore pinned the C51 row
with a re-typed numpy restatement plus the docstring's 2x5 example.  This script removes the
re-typing: the reference module is imported unmodified (gin stubbed, `tf` replaced by the
shim below) and its functions are called as they stand.  What the shim contributes is
only the meaning of each op -- elementwise float32 arithmetic, clip, tile, reshape, abs,
reduce_sum over the last axis, softmax as exp(x - max) / sum, argmax with the first maximal
index, gather_nd -- which is what TensorFlow documents for them; none of the algorithm's
structure is re-stated.  Outputs are float32 like the TF graph's.

Run in the build container only:  python tests/golden/make_c51_golden.py
"""
import contextlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


class Shape(tuple):
  """tf.TensorShape's two assertions used by project_distribution."""

  def assert_is_compatible_with(self, other):
    assert tuple(self) == tuple(other), (self, other)

  def assert_has_rank(self, rank):
    assert len(self) == rank, (self, rank)


class T(np.ndarray):
  """ndarray whose .shape carries the TensorShape assertions."""

  @property
  def shape(self):
    return Shape(np.ndarray.shape.__get__(self))


def t(x, dtype=None):
  return np.asarray(x, dtype=dtype).view(T)


def make_tf_shim():
  tf = types.ModuleType("tensorflow_numpy_shim")
  tf.float32, tf.int64 = np.float32, np.int64
  tf.shape = lambda x: t(np.asarray(np.ndarray.shape.__get__(np.asarray(x)), np.int64))
  tf.size = lambda x: np.asarray(x).size
  tf.equal = lambda a, b: t(np.equal(a, b))
  tf.reduce_all = lambda x: bool(np.all(x))
  tf.Assert = lambda cond, data: None if cond else (_ for _ in ()).throw(AssertionError(data))
  tf.control_dependencies = lambda deps: contextlib.nullcontext()
  tf.clip_by_value = lambda x, lo, hi: t(np.minimum(np.maximum(x, np.float32(lo)), np.float32(hi)))
  tf.tile = lambda x, mult: t(np.tile(np.asarray(x), [int(m) for m in mult]))
  tf.reshape = lambda x, shape: t(np.reshape(np.asarray(x), [int(s) for s in shape]))
  tf.abs = lambda x: t(np.abs(x))
  tf.reduce_sum = lambda x, axis=None: t(np.sum(np.asarray(x, np.float32), axis=axis, dtype=np.float32))
  tf.cast = lambda x, dtype: t(np.asarray(x).astype(dtype))
  tf.argmax = lambda x, axis=None: t(np.argmax(np.asarray(x), axis=axis).astype(np.int64))
  tf.range = lambda n: t(np.arange(int(n), dtype=np.int64))
  tf.to_int64 = lambda x: np.int64(x)
  tf.concat = lambda xs, axis=0: t(np.concatenate([np.asarray(x) for x in xs], axis=axis))
  tf.gather_nd = lambda params, idx: t(np.asarray(params)[tuple(np.asarray(idx).T)])

  def softmax(logits):  # tf.contrib.layers.softmax: over the last axis
    x = np.asarray(logits, np.float32)
    e = np.exp(x - x.max(axis=-1, keepdims=True), dtype=np.float32)
    return t(e / e.sum(axis=-1, keepdims=True, dtype=np.float32))

  tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(softmax=softmax), slim=None)
  tf.train = types.SimpleNamespace(RMSPropOptimizer=lambda **k: None, AdamOptimizer=lambda **k: None)
  tf.logging = types.SimpleNamespace(info=lambda *a, **k: None)
  return tf


def load_reference_rainbow():
  tf = make_tf_shim()
  gin = types.ModuleType("gin")
  gin.configurable = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
  gin.tf = types.ModuleType("gin.tf")
  sys.modules["tensorflow"] = tf
  sys.modules["gin"] = gin
  sys.modules["gin.tf"] = gin.tf
  rb = os.path.join(REF, "agents", "rainbow_copy")
  sys.path.insert(0, rb)
  third = types.ModuleType("third_party")
  third.__path__ = [os.path.join(rb, "third_party")]
  sys.modules.setdefault("third_party", third)
  import rainbow_agent
  return rainbow_agent


def main():
  ra = load_reference_rainbow()
  rng = np.random.default_rng(0)
  out = {}

  # --- project_distribution on its own: the docstring example and random batches
  z = t(np.asarray([4, 5, 6, 7, 8], np.float32))
  s = t(np.asarray([[0, 2, 4, 6, 8], [1, 3, 4, 5, 6]], np.float32))
  w = t(np.asarray([[.1, .6, .1, .1, .1], [.1, .2, .5, .1, .1]], np.float32))
  doc = np.asarray(ra.project_distribution(s, w, z, validate_args=True))
  assert np.allclose(doc, [[.8, 0, .1, 0, .1], [.8, .1, .1, 0, 0]], atol=1e-6)
  for name, (B, n, vmax) in {"p51": (64, 51, 25.0), "p51_wide": (48, 51, 25.0), "p11": (33, 11, 10.0),
                             "p2": (9, 2, 1.0)}.items():
    support = np.linspace(-vmax, vmax, n).astype(np.float32)
    spread = 3.0 if name == "p51_wide" else 1.2
    sup = (rng.normal(size=(B, n)) * vmax * spread / 2).astype(np.float32)
    sup[0] = support                       # identity row
    sup[1] = support + np.float32(0.5) * (support[1] - support[0])  # half-way between atoms
    sup[2] = vmax * 4                      # everything clipped to v_max
    sup[3] = -vmax * 4
    wts = rng.random((B, n)).astype(np.float32)
    wts /= wts.sum(axis=1, keepdims=True)
    proj = np.asarray(ra.project_distribution(t(sup), t(wts), t(support)))
    assert proj.dtype == np.float32
    out[name + "/supports"], out[name + "/weights"], out[name + "/target_support"] = sup, wts, support
    out[name + "/projection"] = proj
    print(name, proj.shape, "row sums", float(proj.sum(axis=1).min()), float(proj.sum(axis=1).max()))

  # --- the agent's own graph pieces on a constructor-free RainbowAgent
  for name, (B, A, n, vmax, gamma_n) in {"t51": (48, 20, 51, 25.0, 0.99), "t51_n3": (24, 38, 51, 25.0, 0.99 ** 3),
                                         "t7": (40, 11, 7, 5.0, 0.9)}.items():
    agent = object.__new__(ra.RainbowAgent)
    agent.num_atoms = n
    agent.support = t(np.linspace(-vmax, vmax, n).astype(np.float32))  # rainbow_agent.py: tf.linspace(-vmax, vmax, num_atoms)
    agent.cumulative_gamma = gamma_n
    logits_act = (rng.normal(size=(1, A, n)) * 2).astype(np.float32)
    next_logits = (rng.normal(size=(B, A, n)) * 2).astype(np.float32)
    next_logits[:4] = np.round(next_logits[:4])  # ties between actions are possible
    legal_act = np.where(rng.random(A) < 0.5, 0.0, -np.inf).astype(np.float32)
    legal_act[int(rng.integers(A))] = 0.0
    next_legal = np.where(rng.random((B, A)) < 0.45, 0.0, -np.inf).astype(np.float32)
    next_legal[np.arange(B), rng.integers(0, A, B)] = 0.0
    rewards = rng.integers(-1, 3, B).astype(np.float32)
    terminals = (rng.random(B) < 0.25).astype(np.uint8)
    agent._q = t(logits_act)
    agent.legal_actions_ph = t(legal_act)
    agent._replay_qs = t(next_logits)
    agent._replay_next_qt = t(next_logits)
    agent._replay = types.SimpleNamespace(rewards=t(rewards), terminals=t(terminals),
                                          next_legal_actions=t(next_legal))
    target = np.asarray(agent._build_target_distribution())  # calls _reshape_networks itself
    assert target.dtype == np.float32 and target.shape == (B, n)
    out[name + "/cfg"] = np.asarray([B, A, n])
    out[name + "/vmax"], out[name + "/cumulative_gamma"] = np.float32(vmax), np.float32(gamma_n)
    out[name + "/act_logits"], out[name + "/act_legal"] = logits_act[0], legal_act
    out[name + "/act_q"] = np.asarray(agent._q)[0]
    out[name + "/act_argmax"] = np.int32(agent._q_argmax)
    out[name + "/rewards"], out[name + "/terminals"] = rewards, terminals
    out[name + "/next_logits"], out[name + "/next_legal"] = next_logits, next_legal
    out[name + "/target"] = target
    # per-row q-values / argmax of the next state, from the same reference ops
    probs = sys.modules["tensorflow"].contrib.layers.softmax(next_logits)
    q_next = np.asarray(sys.modules["tensorflow"].reduce_sum(agent.support * probs, 2))
    out[name + "/next_q"] = q_next
    out[name + "/next_argmax"] = np.argmax(q_next + next_legal, axis=1).astype(np.int32)
    print(name, target.shape, "row sums", float(target.sum(axis=1).min()), float(target.sum(axis=1).max()))
  np.savez_compressed(os.path.join(HERE, "c51_golden.npz"), **out)


if __name__ == "__main__":
  main()
