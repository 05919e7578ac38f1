"""Batched counterpart of the reference's tf-agents environment wrapper
(training/tf_agents_lib/pyhanabi_env_wrapper.py:10-123) without a tf-agents
dependency: the same TimeStep structure and episode-boundary protocol, for N
games at once, as torch tensors on the GPU.

  TimeStep(step_type, reward, discount, observation)
    step_type    int32 [N]     0 FIRST, 1 MID, 2 LAST          (tf_agents StepType values)
    reward       float32 [N]   score delta of the move, 0 on FIRST (wrapper :50-51, :77)
    discount     float32 []    0.999                            (wrapper :7)
    observation  {"state": uint8 [N, obs_dim]  (dtype_vectorized = 'uint8', :8),
                  "mask":  float32 [N, A], 0 for legal moves, float32.min otherwise (:31-39),
                  "info":  int32 [N], the current score (:59, :83)}

Protocol per game, as in the reference wrapper (:61-67): the step after a LAST
time step ignores that game's action, starts a new episode and returns FIRST.
"""
import collections

TimeStep = collections.namedtuple("TimeStep", ["step_type", "reward", "discount", "observation"])
FIRST, MID, LAST = 0, 1, 2
DISCOUNT = 0.999


class BatchedPyhanabiEnvWrapper(object):
  """Wraps a hanabi_b200.rl_env.BatchedHanabiEnv (created with any auto_reset
  setting; the wrapper drives resets itself)."""

  batched = True

  def __init__(self, env):
    import torch
    self._torch = torch
    self._env = env
    self._batch = env.batch
    n = env.num_games
    dev = env.device
    self.batch_size = n
    self._ended = torch.zeros(n, dtype=torch.uint8, device=dev)
    self._eff_actions = torch.zeros(n, dtype=torch.int32, device=dev)
    self._reward = torch.zeros(n, dtype=torch.int32, device=dev)
    self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
    self._score = torch.zeros(n, dtype=torch.int32, device=dev)
    self._obs = torch.zeros((n, env.batch.obs_size), dtype=torch.uint8, device=dev)
    self._mask = torch.zeros((n, env.batch.max_moves), dtype=torch.uint8, device=dev)
    self._cur = torch.zeros(n, dtype=torch.int32, device=dev)
    self._discount = torch.tensor(DISCOUNT, dtype=torch.float32, device=dev)
    self._float_min = float(torch.finfo(torch.float32).min)

  def __getattr__(self, name):
    return getattr(self._env, name)  # forward everything else to the environment

  def _float_mask(self):
    torch = self._torch
    out = torch.zeros(self._mask.shape, dtype=torch.float32, device=self._mask.device)
    return out.masked_fill_(self._mask == 0, self._float_min)

  def _time_step(self, step_type, reward):
    obs = {"state": self._obs, "mask": self._float_mask(), "info": self._score.clone()}
    return TimeStep(step_type, reward, self._discount, obs)

  def reset(self):
    torch = self._torch
    self._batch.reset()
    self._batch.observe(self._obs, self._mask, self._cur)
    self._ended.zero_()
    self._score.zero_()
    n = self.batch_size
    return self._time_step(torch.full((n,), FIRST, dtype=torch.int32, device=self._obs.device),
                           torch.zeros(n, dtype=torch.float32, device=self._obs.device))

  def step(self, action):
    """action: integer tensor [N] of move uids."""
    torch = self._torch
    ended = self._ended != 0
    # games whose last step was LAST start a new episode and ignore their action
    self._batch.reset(self._ended)
    self._eff_actions.copy_(torch.where(ended, torch.full_like(action, -1), action).to(torch.int32))
    self._batch.step_encode(self._eff_actions, self._reward, self._done, self._obs, self._mask,
                            self._cur, scores=self._score, auto_reset=False)
    done = (self._done != 0) & ~ended
    step_type = torch.where(ended, torch.full_like(self._reward, FIRST),
                            torch.where(done, torch.full_like(self._reward, LAST),
                                        torch.full_like(self._reward, MID)))
    reward = torch.where(ended, torch.zeros_like(self._reward), self._reward).to(torch.float32)
    self._ended.copy_(done.to(torch.uint8))
    return self._time_step(step_type, reward)

  def observation_spec(self):
    return {"state": {"shape": tuple(self._env.vectorized_observation_shape()), "dtype": "uint8",
                      "minimum": 0, "maximum": 1, "name": "state"},
            "mask": {"shape": (self._env.num_moves(),), "dtype": "float32", "name": "mask"},
            "info": {"shape": (), "dtype": "int32", "name": "info"}}

  def action_spec(self):
    return {"shape": (), "dtype": "int64", "minimum": 0, "maximum": self._env.num_moves() - 1,
            "name": "action"}
