"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the two CPU checkers:

  kind="oracle": oracle/libhanabi_oracle.so, the plain-C restatement
                 (oracle/hanabi_oracle.c)
  kind="ref":    oracle/_ref/libhanabi_ref.so, the UNMODIFIED reference engine
                 driven by oracle/ref_driver.cc

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg import this module.  Nothing under hanabi_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libhanabi_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libhanabi_ref.so")
REF_PYHANABI_SO = os.path.join(HERE, "_ref", "libpyhanabi.so")
REFERENCE_ROOT = "/root/reference/hanabi_learning_environment"

DUMP_LEN = 272
DUMP = dict(cur=0, info=1, life=2, deck=3, score=4, end=5, fireworks=6, handsize=11, cards=16,
            cplaus=56, rplaus=96, chint=136, rhint=176, deckcount=216, discardcount=241)

_libs = {}

u8p = ctypes.POINTER(ctypes.c_uint8)
i32p = ctypes.POINTER(ctypes.c_int32)
i64p = ctypes.POINTER(ctypes.c_int64)
u32p = ctypes.POINTER(ctypes.c_uint32)
f64p = ctypes.POINTER(ctypes.c_double)


def build(want_ref=True):
  """Compile the checkers (oracle always; _ref only where the reference tree exists)."""
  subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
  if want_ref and os.path.isdir(REFERENCE_ROOT):
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_ref():
  return os.path.exists(REF_SO)


def _ptr(a, t):
  return None if a is None else a.ctypes.data_as(t)


def load(kind):
  if kind in _libs:
    return _libs[kind]
  path, prefix = (ORACLE_SO, "Orc") if kind == "oracle" else (REF_SO, "Ref")
  if not os.path.exists(path):
    if kind == "oracle":
      build(want_ref=False)
    else:
      raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
  lib = ctypes.CDLL(path)

  def fn(name, restype, *argtypes):
    f = getattr(lib, prefix + name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f

  api = dict(
      new=fn("BatchNew", ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
             ctypes.c_int, i32p),
      delete=fn("BatchDelete", None, ctypes.c_void_p),
      obs_size=fn("BatchObservationSize", ctypes.c_int, ctypes.c_void_p),
      max_moves=fn("BatchMaxMoves", ctypes.c_int, ctypes.c_void_p),
      num_players=fn("BatchNumPlayers", ctypes.c_int, ctypes.c_void_p),
      reset=fn("BatchReset", None, ctypes.c_void_p, u8p),
      step=fn("BatchStep", None, ctypes.c_void_p, i32p, i32p, u8p, i32p, ctypes.c_int, u8p),
      encode=fn("BatchEncode", None, ctypes.c_void_p, ctypes.c_int, u8p, ctypes.c_int64),
      legal_mask=fn("BatchLegalMask", None, ctypes.c_void_p, u8p),
      cur_player=fn("BatchCurPlayer", None, ctypes.c_void_p, i32p),
      dump=fn("BatchDumpState", None, ctypes.c_void_p, i32p),
      discrete=fn("DiscreteSample", ctypes.c_int, f64p, ctypes.c_int, ctypes.c_uint32,
                  ctypes.c_uint32),
      deal_from_counts=fn("DealFromCounts", ctypes.c_int, i32p, ctypes.c_int, ctypes.c_uint32,
                          ctypes.c_uint32),
      mt_words=fn("MtWords", None, ctypes.c_int32, ctypes.c_int, u32p),
      uniform_ints=fn("UniformInts", None, ctypes.c_int32, ctypes.c_int, ctypes.c_int, i32p),
  )
  if kind == "oracle":
    api["rollout"] = fn("BatchRollout", ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, u8p,
                        ctypes.c_int64, u8p, i64p, i64p)
    api["load_state"] = fn("BatchLoadState", None, ctypes.c_void_p, ctypes.c_int, i32p,
                           ctypes.c_int, ctypes.c_int)
  else:
    api["rollout"] = fn("BatchRollout", ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                        ctypes.c_int, u8p, ctypes.c_int64, u8p, i64p, i64p)
  _libs[kind] = api
  return api


def params_to_kv(params):
  """dict -> flat [k0, v0, k1, v1, ...] of bytes, like pyhanabi.HanabiGame does
  (hanabi_learning_environment/pyhanabi.py:700-712)."""
  flat = []
  for k, v in (params or {}).items():
    if k == "seed":
      continue  # seeds are per game
    flat += [str(k).encode("ascii"), str(v).encode("ascii")]
  return flat


class CpuBatch(object):
  """N independent CPU games behind the same verbs as the CUDA batch."""

  def __init__(self, kind, params, num_games, seeds):
    self.kind = kind
    self.api = load(kind)
    flat = params_to_kv(params)
    arr = (ctypes.c_char_p * max(len(flat), 1))(*flat)
    self.seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.int32))
    assert self.seeds.shape == (num_games,)
    self.n = int(num_games)
    self.handle = self.api["new"](len(flat), arr, self.n, _ptr(self.seeds, i32p))
    if not self.handle:
      raise ValueError("bad game parameters: %r" % (params,))
    self.obs_size = self.api["obs_size"](self.handle)
    self.max_moves = self.api["max_moves"](self.handle)
    self.num_players = self.api["num_players"](self.handle)

  def __del__(self):
    if getattr(self, "handle", None):
      self.api["delete"](self.handle)
      self.handle = None

  def reset(self, which=None):
    if which is not None:
      which = np.ascontiguousarray(which, dtype=np.uint8)
    self.api["reset"](self.handle, _ptr(which, u8p))

  def step(self, actions, auto_reset=True):
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    rewards = np.zeros(self.n, np.int32)
    dones = np.zeros(self.n, np.uint8)
    scores = np.zeros(self.n, np.int32)
    illegal = np.zeros(self.n, np.uint8)
    self.api["step"](self.handle, _ptr(actions, i32p), _ptr(rewards, i32p), _ptr(dones, u8p),
                     _ptr(scores, i32p), int(bool(auto_reset)), _ptr(illegal, u8p))
    return rewards, dones, scores, illegal

  def encode(self, observer=-1, pitch=None):
    pitch = self.obs_size if pitch is None else pitch
    obs = np.zeros((self.n, pitch), np.uint8)
    self.api["encode"](self.handle, int(observer), _ptr(obs, u8p), pitch)
    return obs

  def legal_mask(self):
    mask = np.zeros((self.n, self.max_moves), np.uint8)
    self.api["legal_mask"](self.handle, _ptr(mask, u8p))
    return mask

  def cur_player(self):
    cur = np.zeros(self.n, np.int32)
    self.api["cur_player"](self.handle, _ptr(cur, i32p))
    return cur

  def dump(self):
    out = np.zeros((self.n, DUMP_LEN), np.int32)
    self.api["dump"](self.handle, _ptr(out, i32p))
    return out

  def load_state(self, game, dump, next_player, turns_to_play):
    dump = np.ascontiguousarray(dump, dtype=np.int32)
    self.api["load_state"](self.handle, int(game), _ptr(dump, i32p), int(next_player),
                           int(turns_to_play))

  def rollout(self, steps, threads=1):
    """Random-legal rollout with auto-reset, obs+mask written each step.
    Returns (env_steps, episodes, score_sum)."""
    obs = np.zeros((self.n, self.obs_size), np.uint8)
    mask = np.zeros((self.n, self.max_moves), np.uint8)
    ep = ctypes.c_int64(0)
    sc = ctypes.c_int64(0)
    if self.kind == "oracle":
      total = self.api["rollout"](self.handle, int(steps), _ptr(obs, u8p), self.obs_size,
                                  _ptr(mask, u8p), ctypes.byref(ep), ctypes.byref(sc))
    else:
      total = self.api["rollout"](self.handle, int(steps), int(threads), _ptr(obs, u8p),
                                  self.obs_size, _ptr(mask, u8p), ctypes.byref(ep),
                                  ctypes.byref(sc))
    return int(total), int(ep.value), int(sc.value)


def deal_from_counts(kind, counts, lo, hi):
  counts = np.ascontiguousarray(counts, dtype=np.int32)
  return load(kind)["deal_from_counts"](_ptr(counts, i32p), len(counts), int(lo), int(hi))


def mt_words(kind, seed, n):
  out = np.zeros(n, np.uint32)
  load(kind)["mt_words"](int(seed), n, _ptr(out, u32p))
  return out


def uniform_ints(kind, seed, hi, n):
  out = np.zeros(n, np.int32)
  load(kind)["uniform_ints"](int(seed), int(hi), n, _ptr(out, i32p))
  return out
