"""Device-side bookkeeping of a self-play actor loop over a BatchedHanabiEnv: what
DQNAgent.begin_episode / step / end_episode do around the network
(agents/rainbow_copy/dqn_agent.py:286-385) and what run_one_episode does with the
rewards (agents/run_experiment.py:344-403), for N games at once on the GPU.

Per (game, seat) the recorder keeps the episode's transitions (o_t, l_t, a_t); every
reward is added to every seat's "reward since my last action" (run_experiment.py
:368-377); when a seat acts again that sum becomes the reward of its previous
transition (dqn_agent.py:311-324, 365-375) and is cleared; when the game ends each
seat's last transition gets the seat's pending sum and the terminal flag, and each
seat's episode is appended to the replay memory as one contiguous run
(_post_transitions, dqn_agent.py:359-385) -- which is what makes the replay's n-step
windows meaningful.  Plain torch indexing: this is glue, not a hot kernel.
"""


class SelfPlayRecorder(object):

  def __init__(self, num_games, num_players, observation_size, num_actions, device,
               replay=None, max_turns_per_seat=64):
    import torch
    self._torch = torch
    self.N, self.P, self.D, self.A, self.L = (num_games, num_players, observation_size,
                                              num_actions, max_turns_per_seat)
    self.replay = replay
    dev = device
    self.obs = torch.zeros((self.N, self.P, self.L, self.D), dtype=torch.uint8, device=dev)
    self.legal = torch.zeros((self.N, self.P, self.L, self.A), dtype=torch.uint8, device=dev)
    self.action = torch.zeros((self.N, self.P, self.L), dtype=torch.int32, device=dev)
    self.reward = torch.zeros((self.N, self.P, self.L), dtype=torch.float32, device=dev)
    self.count = torch.zeros((self.N, self.P), dtype=torch.int64, device=dev)
    self.since = torch.zeros((self.N, self.P), dtype=torch.float32, device=dev)
    self._rows = torch.arange(self.N, device=dev)
    self.dropped = 0  # transitions beyond max_turns_per_seat (never in a legal Hanabi game at 64)

  def before_step(self, current_player, observation, legal_mask, action):
    """The acting seat's view and choice: begin_episode / step -> _record_transition."""
    torch = self._torch
    rows, cur = self._rows, current_player.long()
    k = self.count[rows, cur]
    prev = k - 1
    has_prev = prev >= 0
    # the reward handed to step() belongs to the seat's previous transition
    self.reward[rows[has_prev], cur[has_prev], prev[has_prev]] = self.since[rows[has_prev], cur[has_prev]]
    self.since[rows, cur] = 0
    fits = k < self.L
    self.dropped += int((~fits).sum())
    r, c, kk = rows[fits], cur[fits], k[fits]
    self.obs[r, c, kk] = observation[fits]
    self.legal[r, c, kk] = legal_mask[fits]
    self.action[r, c, kk] = action[fits].to(torch.int32)
    self.count[r, c] = kk + 1

  def after_step(self, reward, done):
    """reward / done of the env step just taken.  Returns the posted transitions
    (obs, action, reward, terminal, legal float 0/-inf) or None."""
    torch = self._torch
    self.since += reward.to(torch.float32)[:, None]
    finished = torch.nonzero(done != 0)[:, 0]
    if finished.numel() == 0:
      return None
    F = finished.numel()
    lengths = self.count[finished]                                  # [F, P]
    flat_len = lengths.reshape(-1)
    total = int(flat_len.sum())
    if total == 0:
      self.since[finished] = 0
      return None
    # trajectory (f, p) occupies a contiguous run; runs ordered by game, then seat
    traj = torch.repeat_interleave(torch.arange(F * self.P, device=flat_len.device), flat_len)
    start = torch.cumsum(flat_len, 0) - flat_len
    step = torch.arange(total, device=flat_len.device) - start[traj]
    g = finished[traj // self.P]
    p = traj % self.P
    last = step == (flat_len[traj] - 1)
    obs = self.obs[g, p, step]
    action = self.action[g, p, step]
    reward_out = torch.where(last, self.since[g, p], self.reward[g, p, step])  # terminal_rewards
    terminal = last.to(torch.uint8)
    legal = torch.zeros((total, self.A), dtype=torch.float32, device=obs.device)
    legal.masked_fill_(self.legal[g, p, step] == 0, float("-inf"))  # format_legal_moves
    if self.replay is not None:
      self.replay.add(obs.contiguous(), action.contiguous(), reward_out.contiguous(),
                      terminal.contiguous(), legal.contiguous())
    self.count[finished] = 0
    self.since[finished] = 0
    return obs, action, reward_out, terminal, legal


class DeviceRecorder(object):
  """The same bookkeeping as SelfPlayRecorder in hand-written kernels (NewRecorder /
  RecorderBeforeStep / RecorderAfterStep, hb_replay.cu): episode buffers keep the
  observations bit-packed, finished episodes are written straight into the device
  replay memory's ring.  One stream synchronisation per step (the number of posted
  transitions), no torch indexing."""

  def __init__(self, num_games, num_players, observation_size, num_actions, replay,
               max_turns_per_seat=64, device=0):
    import torch
    from hanabi_b200 import pyhanabi
    pyhanabi._require_lib()  # pylint: disable=protected-access
    self._torch, self._ph = torch, pyhanabi
    self._lib, self._ffi = pyhanabi.lib, pyhanabi.ffi
    self.N, self.P, self.D, self.A = int(num_games), int(num_players), int(observation_size), int(num_actions)
    self.replay = replay
    self.device = torch.device("cuda", int(device))
    self._c = self._ffi.new("pyhanabi_recorder_t*")
    self._check(self._lib.NewRecorder(self._c, self.N, self.P, self.D, self.A,
                                      int(max_turns_per_seat), int(device)), "NewRecorder")
    self.packed_words = int(self._lib.RecorderPackedWords(self._c))
    self._posted = self._ffi.new("int64_t*")

  def __del__(self):
    c, lib = getattr(self, "_c", None), getattr(self, "_lib", None)
    if c is not None and lib is not None:
      try:
        lib.DeleteRecorder(c)
      except Exception:  # interpreter shutdown
        pass
      self._c = None

  def _check(self, rc, what):
    if rc != 0:
      raise RuntimeError("%s failed: %s" % (what, self._ph.encode_ffi_string(self._lib.BatchLastError())))

  def _p(self, t, dtype, ctype, numel):
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous() or t.numel() < numel:
      raise ValueError("expected a contiguous CUDA %s tensor with >= %d elements" % (dtype, numel))
    return self._ffi.cast(ctype, t.data_ptr())

  def _stream(self):
    return self._ffi.cast("void*", self._torch.cuda.current_stream(self.device).cuda_stream)

  def before_step(self, current_player, obs_bits, legal_mask, action):
    """obs_bits: int32 [N, packed_words] as step_encode_packed / observe_packed write them."""
    torch = self._torch
    self._check(self._lib.RecorderBeforeStep(
        self._c, self._p(current_player, torch.int32, "int32_t*", self.N),
        self._p(obs_bits, torch.int32, "uint32_t*", self.N * self.packed_words),
        self._p(legal_mask, torch.uint8, "uint8_t*", self.N * self.A),
        self._p(action, torch.int32, "int32_t*", self.N), self._stream()), "RecorderBeforeStep")

  def after_step(self, reward, done, priority=None, return_count=True):
    """reward int32 [N], done uint8 [N].  Finished episodes go into the replay memory; the
    whole update is driven from device memory (cursor and count never visit the host), so
    with return_count=False the call is asynchronous and can be captured in a CUDA graph.
    return_count=True synchronises once and returns the number of transitions appended."""
    torch = self._torch
    self._check(self._lib.RecorderAfterStep(
        self._c, self._p(reward, torch.int32, "int32_t*", self.N),
        self._p(done, torch.uint8, "uint8_t*", self.N), self.replay._c,  # pylint: disable=protected-access
        -1.0 if priority is None else float(priority),
        self._posted if return_count else self._ffi.NULL, self._stream()), "RecorderAfterStep")
    return int(self._posted[0]) if return_count else None

  @property
  def dropped(self):
    return int(self._lib.RecorderDropped(self._c, self._stream()))


def linearly_decaying_epsilon(decay_period, step, warmup_steps, epsilon):
  """Epsilon schedule of the DQN/Rainbow agents
  (agents/rainbow_copy/dqn_agent.py:44-59): 1.0 during warm-up, then linear decay to
  `epsilon` over `decay_period` steps."""
  steps_left = decay_period + warmup_steps - step
  bonus = (1.0 - epsilon) * steps_left / decay_period
  bonus = min(max(bonus, 0.0), 1.0 - epsilon)
  return epsilon + bonus
