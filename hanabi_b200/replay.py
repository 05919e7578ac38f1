"""Device-resident replay memory: the reference's OutOfGraphReplayMemory /
OutOfGraphPrioritizedReplayMemory (agents/rainbow_copy/replay_memory.py:67-347,
prioritized_replay_memory.py:41-170) for stack_size 1, with torch CUDA tensors as
buffers.  Same method names and return order as the reference classes; `add` takes
a batch of transitions (appended in order) instead of one."""
from hanabi_b200 import pyhanabi
from hanabi_b200.pyhanabi import ffi

DEFAULT_PRIORITY = 100.0


class DeviceReplayMemory(object):

  def __init__(self, num_actions, observation_size, replay_capacity, batch_size=32,
               update_horizon=1, gamma=1.0, prioritized=True, device=0, seed=0):
    import torch
    pyhanabi._require_lib()  # pylint: disable=protected-access
    self._torch = torch
    self._lib = pyhanabi.lib
    self._num_actions, self._observation_size = int(num_actions), int(observation_size)
    self._replay_capacity, self._batch_size = int(replay_capacity), int(batch_size)
    self._update_horizon, self._gamma = int(update_horizon), float(gamma)
    self.prioritized = bool(prioritized)
    self.device = torch.device("cuda", int(device))
    self._seed, self._draws = int(seed), 0
    self._c = ffi.new("pyhanabi_replay_t*")
    self._check(self._lib.NewReplay(self._c, self._num_actions, self._observation_size,
                                    self._replay_capacity, self._update_horizon, self._gamma,
                                    int(self.prioritized), int(device)), "NewReplay")

  def __del__(self):
    c = getattr(self, "_c", None)
    lib = getattr(self, "_lib", None)
    if c is not None and lib is not None:
      try:
        lib.DeleteReplay(c)
      except Exception:  # interpreter shutdown: cffi may already be torn down
        pass
      self._c = None

  def _check(self, rc, what):
    if rc != 0:
      raise RuntimeError("%s failed: %s" % (what, pyhanabi.encode_ffi_string(
          self._lib.BatchLastError())))

  def _p(self, t, dtype, ctype, numel):
    if not t.is_cuda or t.device != self.device or t.dtype != dtype or not t.is_contiguous() \
        or t.numel() < numel:
      raise ValueError("expected a contiguous %s tensor with >= %d elements on %s" %
                       (dtype, numel, self.device))
    return ffi.cast(ctype, t.data_ptr())

  def _stream(self):
    return ffi.cast("void*", self._torch.cuda.current_stream(self.device).cuda_stream)

  # -- bookkeeping, as in the reference
  @property
  def add_count(self):
    return int(self._lib.ReplayAddCount(self._c))

  def cursor(self):
    return self.add_count % self._replay_capacity

  def is_empty(self):
    return self.add_count == 0

  def is_full(self):
    return self.add_count >= self._replay_capacity

  # -- writing
  def add(self, observation, action, reward, terminal, legal_actions, priority=None):
    """Batch add: observation uint8 [B, obs], action int32 [B], reward float32 [B],
    terminal uint8 [B], legal_actions float32 [B, A]; appended in order."""
    torch = self._torch
    B = int(action.shape[0])
    self._check(self._lib.ReplayAdd(
        self._c, self._p(observation, torch.uint8, "uint8_t*", B * self._observation_size),
        self._p(action, torch.int32, "int32_t*", B), self._p(reward, torch.float32, "float*", B),
        self._p(terminal, torch.uint8, "uint8_t*", B),
        self._p(legal_actions, torch.float32, "float*", B * self._num_actions), B,
        -1.0 if priority is None else float(priority), self._stream()), "ReplayAdd")

  # -- reading
  def is_valid_transition(self, indices):
    torch = self._torch
    B = int(indices.numel())
    out = torch.empty(B, dtype=torch.uint8, device=self.device)
    self._check(self._lib.ReplayIsValidTransition(self._c, self._p(indices, torch.int32, "int32_t*", B),
                                                  B, self._p(out, torch.uint8, "uint8_t*", B),
                                                  self._stream()), "ReplayIsValidTransition")
    return out

  def tree_sample(self, query_values):
    torch = self._torch
    B = int(query_values.numel())
    out = torch.empty(B, dtype=torch.int32, device=self.device)
    self._check(self._lib.ReplayTreeSample(self._c, self._p(query_values, torch.float64, "double*", B),
                                           B, self._p(out, torch.int32, "int32_t*", B),
                                           self._stream()), "ReplayTreeSample")
    return out

  def sample_index_batch(self, batch_size, stratified=False, check=True):
    """Indices of valid transitions (replay_memory.py:260-288 / prioritized_replay_memory.py:
    109-132).  Like the reference, raises RuntimeError when no valid transition is found for
    some element (check=True reads a device counter back: one stream synchronisation; pass
    check=False inside captured / fully asynchronous loops and poll sample_failures())."""
    torch = self._torch
    out = torch.empty(batch_size, dtype=torch.int32, device=self.device)
    self._draws += 1
    self._check(self._lib.ReplaySampleIndexBatch(self._c, int(batch_size), int(bool(stratified)),
                                                 self._seed, self._draws,
                                                 self._p(out, torch.int32, "int32_t*", batch_size),
                                                 self._stream()), "ReplaySampleIndexBatch")
    if check and self.sample_failures(clear=True):
      raise RuntimeError("Max sample attempts: Tried 1000 times but only sampled fewer than "
                         "batch_size=%d valid indices" % batch_size)
    return out

  def sample_failures(self, clear=True):
    return int(self._lib.ReplaySampleFailures(self._c, int(bool(clear)), self._stream()))

  def sample_transition_batch(self, batch_size=None, indices=None):
    """Returns (state, action, reward, next_state, terminal, indices, next_legal_actions) like
    the reference (replay_memory.py:344-346), plus the sampling probabilities when
    prioritized (prioritized_replay_memory.py, WrappedPrioritizedReplayMemory)."""
    torch = self._torch
    if batch_size is None:
      batch_size = self._batch_size if indices is None else int(indices.numel())
    if indices is None:
      indices = self.sample_index_batch(batch_size)
    B, D, A = int(batch_size), self._observation_size, self._num_actions
    dev = self.device
    state = torch.empty((B, D), dtype=torch.uint8, device=dev)
    next_state = torch.empty((B, D), dtype=torch.uint8, device=dev)
    action = torch.empty(B, dtype=torch.int32, device=dev)
    reward = torch.empty(B, dtype=torch.float32, device=dev)
    terminal = torch.empty(B, dtype=torch.uint8, device=dev)
    next_legal = torch.empty((B, A), dtype=torch.float32, device=dev)
    self._check(self._lib.ReplaySampleTransitionBatch(
        self._c, self._p(indices, torch.int32, "int32_t*", B), B,
        self._p(state, torch.uint8, "uint8_t*", B * D), self._p(action, torch.int32, "int32_t*", B),
        self._p(reward, torch.float32, "float*", B), self._p(next_state, torch.uint8, "uint8_t*", B * D),
        self._p(terminal, torch.uint8, "uint8_t*", B), self._p(next_legal, torch.float32, "float*", B * A),
        self._stream()), "ReplaySampleTransitionBatch")
    return state, action, reward, next_state, terminal, indices, next_legal

  def set_priority(self, indices, priorities):
    torch = self._torch
    B = int(indices.numel())
    self._check(self._lib.ReplaySetPriority(
        self._c, self._p(indices, torch.int32, "int32_t*", B),
        self._p(priorities, torch.float32, "float*", B), B, float(priorities.max().item()),
        self._stream()), "ReplaySetPriority")

  def get_priority(self, indices):
    torch = self._torch
    B = int(indices.numel())
    out = torch.empty(B, dtype=torch.float32, device=self.device)
    self._check(self._lib.ReplayGetPriority(self._c, self._p(indices, torch.int32, "int32_t*", B), B,
                                            self._p(out, torch.float32, "float*", B), self._stream()),
                "ReplayGetPriority")
    return out
