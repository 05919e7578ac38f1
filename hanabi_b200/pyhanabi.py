"""cffi binding of the B200-native libpyhanabi.so.

Mirrors the loader of the reference binding
(hanabi_learning_environment/pyhanabi.py:42-104): the header text between the
`extern "C" {` / `} /* extern "C" */` marker lines is fed to ffi.cdef() and the
library is dlopen'ed in ABI mode from a list of prefixes.  On top of that sits
`HanabiBatch`, the thin wrapper over the batched entry points with torch
tensors as the device buffers.

There is no CPU fallback: if the CUDA library cannot be loaded the batch
constructor raises.
"""
import os
import re

import cffi

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CDEF_PREFIXES = (os.path.join(os.path.dirname(_HERE), "include"), _HERE, None, ".",
                         "/include")
DEFAULT_LIB_PREFIXES = (_HERE, None, ".", "/lib")
PYHANABI_HEADER = "pyhanabi.h"
PYHANABI_LIB = ["libpyhanabi.so"]
COLOR_CHAR = ["R", "Y", "G", "W", "B"]  # hanabi_lib/util.cc
CHANCE_PLAYER_ID = -1

ffi = cffi.FFI()
lib = None
cdef_loaded_flag = False
lib_loaded_flag = False
lib_path = None

DUMP_LEN = 272


def encode_ffi_string(x):
  return str(ffi.string(x), "ascii")


def try_cdef(header=PYHANABI_HEADER, prefixes=DEFAULT_CDEF_PREFIXES):
  """Parse the library header; same contract as the reference's try_cdef."""
  global cdef_loaded_flag
  if cdef_loaded_flag:
    return True
  for prefix in prefixes:
    try:
      cdef_file = header if prefix is None else prefix + "/" + header
      lines = open(cdef_file).readlines()
      reading_cdef = False
      cdef_string = ""
      for line in lines:
        line = line.rstrip()
        if re.match("extern *\"C\" *{", line):
          reading_cdef = True
          continue
        elif re.match("} */[*] *extern *\"C\" *[*]/", line):
          reading_cdef = False
          continue
        if reading_cdef:
          cdef_string = cdef_string + line + "\n"
      ffi.cdef(cdef_string)
      cdef_loaded_flag = True
      return True
    except IOError:
      pass
  return False


def try_load(library=None, prefixes=DEFAULT_LIB_PREFIXES):
  """dlopen the library; same contract as the reference's try_load."""
  global lib_loaded_flag, lib, lib_path
  if lib_loaded_flag:
    return True
  if library is None:
    libnames = PYHANABI_LIB
  elif type(library) in (list, tuple):
    libnames = library
  else:
    libnames = (library,)
  for prefix in prefixes:
    for libname in libnames:
      try:
        lib_file = libname if prefix is None else prefix + "/" + libname
        lib = ffi.dlopen(lib_file)
        lib_loaded_flag = True
        lib_path = lib_file
        return True
      except OSError:
        pass
  return False


def cdef_loaded():
  return cdef_loaded_flag


def lib_loaded():
  return lib_loaded_flag


def _require_lib():
  if not (try_cdef() and try_load()):
    raise RuntimeError(
        "libpyhanabi.so (CUDA build) not found next to %s; run "
        "`python -m hanabi_b200.build`. The batched engine has no CPU fallback." % __file__)


def _check(rc, what):
  if rc != 0:
    raise RuntimeError("%s failed: %s" % (what, encode_ffi_string(lib.BatchLastError())))


def _params_to_c(params):
  """dict -> (n, char*[]) exactly like HanabiGame.__init__ of the reference
  (pyhanabi.py:700-712); 'seed' is dropped because seeds are per game."""
  keep = []
  for key, value in (params or {}).items():
    if key == "seed":
      continue
    keep.append(ffi.new("char[]", str(key).encode("ascii")))
    keep.append(ffi.new("char[]", str(value).encode("ascii")))
  arr = ffi.new("char * [" + str(max(len(keep), 1)) + "]", keep if keep else [ffi.NULL])
  return len(keep), arr, keep


class HanabiBatch(object):
  """N independent Hanabi games resident on one GPU.

  All tensor arguments are torch CUDA tensors on the batch's device; outputs are
  written in place.  Calls are asynchronous on torch's current stream.
  """

  def __init__(self, params, num_games, seeds, device=0):
    import torch  # device memory / streams only
    _require_lib()
    self._torch = torch
    self.num_games = int(num_games)
    self.device = torch.device("cuda", int(device))
    seeds = [int(s) for s in seeds]
    if len(seeds) != self.num_games:
      raise ValueError("need one seed per game")
    c_seeds = ffi.new("int32_t[]", [((s + 2**31) % 2**32) - 2**31 for s in seeds])
    n, arr, keep = _params_to_c(params)
    self._c = ffi.new("pyhanabi_batch_t*")
    _check(lib.NewBatch(self._c, n, arr, self.num_games, int(device), c_seeds), "NewBatch")
    del keep
    self.obs_size = lib.BatchObservationSize(self._c)
    self.max_moves = lib.BatchMaxMoves(self._c)
    self.num_players = lib.BatchNumPlayers(self._c)

  def __del__(self):
    c = getattr(self, "_c", None)
    if c is not None and lib is not None:
      lib.DeleteBatch(c)
      self._c = None

  # -- helpers
  def _stream(self):
    return ffi.cast("void*", self._torch.cuda.current_stream(self.device).cuda_stream)

  def _p(self, t, dtype, numel=None, ctype="void*"):
    if t is None:
      return ffi.NULL
    torch = self._torch
    if not (t.is_cuda and t.device == self.device):
      raise ValueError("tensor must live on %s" % (self.device,))
    if t.dtype != dtype:
      raise ValueError("expected dtype %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous() and numel is not None:
      raise ValueError("tensor must be contiguous")
    if numel is not None and t.numel() < numel:
      raise ValueError("tensor too small: need %d elements" % numel)
    del torch
    return ffi.cast(ctype, t.data_ptr())

  def _obs_ptr(self, obs):
    if obs is None:
      return ffi.NULL, 0
    torch = self._torch
    if obs.dtype != torch.uint8 or obs.dim() != 2 or obs.stride(1) != 1:
      raise ValueError("obs must be a uint8 [num_games, >=obs_size] tensor with unit inner stride")
    if obs.shape[0] < self.num_games or obs.shape[1] < self.obs_size:
      raise ValueError("obs too small")
    if not (obs.is_cuda and obs.device == self.device):
      raise ValueError("obs must live on %s" % (self.device,))
    return ffi.cast("uint8_t*", obs.data_ptr()), int(obs.stride(0))

  # -- the batched verbs
  def reset(self, which=None):
    torch = self._torch
    _check(lib.BatchReset(self._c, self._p(which, torch.uint8, self.num_games, "uint8_t*"),
                          self._stream()), "BatchReset")

  def step(self, actions, rewards=None, dones=None, scores=None, auto_reset=True):
    torch = self._torch
    n = self.num_games
    _check(lib.BatchStep(self._c, self._p(actions, torch.int32, n, "int32_t*"),
                         self._p(rewards, torch.int32, n, "int32_t*"),
                         self._p(dones, torch.uint8, n, "uint8_t*"),
                         self._p(scores, torch.int32, n, "int32_t*"), int(bool(auto_reset)),
                         self._stream()), "BatchStep")

  def observe(self, obs=None, mask=None, cur_player=None, observer=-1):
    torch = self._torch
    n = self.num_games
    optr, pitch = self._obs_ptr(obs)
    _check(lib.BatchObserve(self._c, int(observer), optr, pitch,
                            self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
                            self._p(cur_player, torch.int32, n, "int32_t*"), self._stream()),
           "BatchObserve")

  def encode(self, obs, observer=-1):
    optr, pitch = self._obs_ptr(obs)
    _check(lib.BatchEncode(self._c, int(observer), optr, pitch, self._stream()), "BatchEncode")

  def legal_mask(self, mask):
    torch = self._torch
    _check(lib.BatchLegalMask(self._c, self._p(mask, torch.uint8, self.num_games * self.max_moves,
                                               "uint8_t*"), self._stream()), "BatchLegalMask")

  def cur_player(self, out):
    torch = self._torch
    _check(lib.BatchCurPlayer(self._c, self._p(out, torch.int32, self.num_games, "int32_t*"),
                              self._stream()), "BatchCurPlayer")

  def step_encode(self, actions, rewards=None, dones=None, obs=None, mask=None, cur_player=None,
                  scores=None, auto_reset=True):
    torch = self._torch
    n = self.num_games
    optr, pitch = self._obs_ptr(obs)
    _check(lib.BatchStepEncodeScores(
        self._c, self._p(actions, torch.int32, n, "int32_t*"),
        self._p(rewards, torch.int32, n, "int32_t*"), self._p(dones, torch.uint8, n, "uint8_t*"),
        self._p(scores, torch.int32, n, "int32_t*"), optr, pitch,
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), self._stream()),
           "BatchStepEncodeScores")

  # -- host-buffer verbs (CPU tensors, ideally pinned); synchronous
  def _h(self, t, dtype, numel, ctype):
    if t is None:
      return ffi.NULL
    if t.is_cuda or t.dtype != dtype or not t.is_contiguous() or t.numel() < numel:
      raise ValueError("expected a contiguous CPU tensor of dtype %s with >= %d elements" %
                       (dtype, numel))
    return ffi.cast(ctype, t.data_ptr())

  def _h_obs(self, obs):
    if obs is None:
      return ffi.NULL, 0
    torch = self._torch
    if (obs.is_cuda or obs.dtype != torch.uint8 or obs.dim() != 2 or obs.stride(1) != 1 or
        obs.shape[0] < self.num_games or obs.shape[1] < self.obs_size):
      raise ValueError("obs must be a CPU uint8 [num_games, >=obs_size] tensor")
    return ffi.cast("uint8_t*", obs.data_ptr()), int(obs.stride(0))

  def step_encode_host(self, actions, rewards=None, dones=None, obs=None, mask=None,
                       cur_player=None, scores=None, auto_reset=True):
    torch = self._torch
    n = self.num_games
    optr, pitch = self._h_obs(obs)
    _check(lib.BatchStepEncodeHost(
        self._c, self._h(actions, torch.int32, n, "int32_t*"),
        self._h(rewards, torch.int32, n, "int32_t*"), self._h(dones, torch.uint8, n, "uint8_t*"),
        self._h(scores, torch.int32, n, "int32_t*"), optr, pitch,
        self._h(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._h(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), self._stream()),
           "BatchStepEncodeHost")

  def reset_host(self, obs=None, mask=None, cur_player=None):
    torch = self._torch
    n = self.num_games
    optr, pitch = self._h_obs(obs)
    _check(lib.BatchResetHost(self._c, optr, pitch,
                              self._h(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
                              self._h(cur_player, torch.int32, n, "int32_t*"), self._stream()),
           "BatchResetHost")

  # -- state I/O
  def state_bytes(self):
    return int(lib.BatchStateBytes(self._c))

  def get_state(self):
    buf = bytearray(self.state_bytes())
    _check(lib.BatchGetState(self._c, ffi.from_buffer(buf), len(buf)), "BatchGetState")
    return bytes(buf)

  def set_state(self, blob):
    _check(lib.BatchSetState(self._c, ffi.from_buffer(blob), len(blob)), "BatchSetState")

  def unpack_state(self, dump=None):
    torch = self._torch
    if dump is None:
      dump = torch.empty((self.num_games, DUMP_LEN), dtype=torch.int32, device=self.device)
    _check(lib.BatchUnpackState(self._c, self._p(dump, torch.int32, self.num_games * DUMP_LEN,
                                                 "int32_t*"), self._stream()), "BatchUnpackState")
    return dump

  def pack_state(self, dump, extra=None):
    torch = self._torch
    _check(lib.BatchPackState(self._c, self._p(dump, torch.int32, self.num_games * DUMP_LEN,
                                               "int32_t*"),
                              self._p(extra, torch.int32, self.num_games * 2, "int32_t*"),
                              self._stream()), "BatchPackState")

  # -- statistics
  def set_stats_buffer(self, stats):
    torch = self._torch
    self._stats_keepalive = stats
    _check(lib.BatchSetStatsBuffer(self._c, self._p(stats, torch.int64, 32, "int64_t*")),
           "BatchSetStatsBuffer")

  def read_stats(self, clear=False):
    out = ffi.new("int64_t[32]")
    _check(lib.BatchReadStats(self._c, out, int(bool(clear)), self._stream()), "BatchReadStats")
    vals = [int(out[i]) for i in range(32)]
    return {"episodes": vals[0], "score_sum": vals[1], "length_sum": vals[2],
            "score_hist": vals[3:29]}

  # -- test hooks
  def set_debug(self, sampler_mode=0, generic=False):
    _check(lib.BatchSetDebug(self._c, int(sampler_mode), int(bool(generic))), "BatchSetDebug")

  def debug_deal(self, decks, sizes, lo, hi, mode):
    torch = self._torch
    n = int(decks.numel())
    out = torch.empty(n, dtype=torch.int32, device=self.device)
    amb = torch.empty(n, dtype=torch.int32, device=self.device)
    _check(lib.BatchDebugDeal(self._c, self._p(decks, torch.int64, n, "uint64_t*"),
                              self._p(sizes, torch.int32, n, "int32_t*"),
                              self._p(lo.view(torch.int32), torch.int32, n, "uint32_t*"),
                              self._p(hi.view(torch.int32), torch.int32, n, "uint32_t*"), n,
                              int(mode), self._p(out, torch.int32, n, "int32_t*"),
                              self._p(amb, torch.int32, n, "int32_t*"), self._stream()),
           "BatchDebugDeal")
    return out, amb



# ---------------------------------------------------------------------------
# Fused actor/learner kernels (free functions: they work on any tensors that
# live on the current CUDA device, not only on a HanabiBatch's outputs).
def _tp(t, dtype, ctype):
  import torch
  if t is None:
    return ffi.NULL
  if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
    raise ValueError("expected a contiguous CUDA tensor of dtype %s" % (dtype,))
  del torch
  return ffi.cast(ctype, t.data_ptr())


def _cur_stream(t):
  import torch
  return ffi.cast("void*", torch.cuda.current_stream(t.device).cuda_stream)


def select_action(values, mask, epsilon, seed, step, support=None, out=None):
  """Masked epsilon-greedy for a batch.  values: None (epsilon>=1), q[n,A] or C51
  logits[n,A,atoms] (then `support[atoms]` is required); mask uint8[n,A]."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  if out is None:
    out = torch.empty(n, dtype=torch.int32, device=mask.device)
  atoms = 0
  if values is not None and values.dim() == 3:
    atoms = int(values.shape[2])
  with torch.cuda.device(mask.device):
    _check(lib.BatchSelectAction(_tp(values, torch.float32, "float*"), atoms,
                                 _tp(support, torch.float32, "float*"),
                                 _tp(mask, torch.uint8, "uint8_t*"), n, A, float(epsilon),
                                 int(seed) & (2**64 - 1), int(step) & (2**64 - 1),
                                 _tp(out, torch.int32, "int32_t*"), _cur_stream(mask)),
           "BatchSelectAction")
  return out


def c51_q_values(logits, support, mask):
  import torch
  _require_lib()
  n, A, atoms = (int(x) for x in logits.shape)
  q = torch.empty((n, A), dtype=torch.float32, device=logits.device)
  greedy = torch.empty(n, dtype=torch.int32, device=logits.device)
  with torch.cuda.device(logits.device):
    _check(lib.BatchC51QValues(_tp(logits, torch.float32, "float*"),
                               _tp(support, torch.float32, "float*"),
                               _tp(mask, torch.uint8, "uint8_t*"), n, A, atoms,
                               _tp(q, torch.float32, "float*"), _tp(greedy, torch.int32, "int32_t*"),
                               _cur_stream(logits)), "BatchC51QValues")
  return q, greedy


def project_distribution(supports, weights, target_support, out=None):
  """Rainbow's C51 projection; same arguments as the reference's
  project_distribution (agents/rainbow_copy/rainbow_agent.py:252)."""
  import torch
  _require_lib()
  B, N = int(supports.shape[0]), int(supports.shape[1])
  if tuple(weights.shape) != (B, N) or tuple(target_support.shape) != (N,):
    raise ValueError("incompatible shapes")  # the reference raises ValueError too (:284-290)
  if out is None:
    out = torch.empty((B, N), dtype=torch.float32, device=supports.device)
  with torch.cuda.device(supports.device):
    _check(lib.BatchProjectDistribution(_tp(supports, torch.float32, "float*"),
                                        _tp(weights, torch.float32, "float*"),
                                        _tp(target_support, torch.float32, "float*"),
                                        _tp(out, torch.float32, "float*"), B, N,
                                        _cur_stream(supports)), "BatchProjectDistribution")
  return out


def c51_target_distribution(rewards, terminals, next_logits, next_legal_mask, support,
                            cumulative_gamma):
  import torch
  _require_lib()
  B, A, N = (int(x) for x in next_logits.shape)
  out = torch.empty((B, N), dtype=torch.float32, device=next_logits.device)
  arg = torch.empty(B, dtype=torch.int32, device=next_logits.device)
  with torch.cuda.device(next_logits.device):
    _check(lib.BatchC51TargetDistribution(
        _tp(rewards, torch.float32, "float*"), _tp(terminals, torch.uint8, "uint8_t*"),
        _tp(next_logits, torch.float32, "float*"), _tp(next_legal_mask, torch.uint8, "uint8_t*"),
        _tp(support, torch.float32, "float*"), float(cumulative_gamma),
        _tp(out, torch.float32, "float*"), _tp(arg, torch.int32, "int32_t*"), B, A, N,
        _cur_stream(next_logits)), "BatchC51TargetDistribution")
  return out, arg


try_cdef()
if cdef_loaded():
  try_load()
