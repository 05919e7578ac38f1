"""RL environment interface of the batched engine.

`BatchedHanabiEnv` is the many-games sibling of the reference's
rl_env.HanabiEnv (hanabi_learning_environment/rl_env.py:67-366): the same
constructor config dict, the same verbs (reset / step /
vectorized_observation_shape / num_moves) and the same reward / done semantics
(rl_env.py:353-363), but every quantity is a torch tensor over `num_games`
games resident on the GPU, produced by one fused kernel per step.

`make()` mirrors rl_env.make (rl_env.py:501-585) and its four presets.
"""
from __future__ import absolute_import
from __future__ import division

from hanabi_b200 import pyhanabi

PRESETS = {  # rl_env.py:521-582
    "Hanabi-Full": dict(colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3,
                        observation_type=1),
    "Hanabi-Full-CardKnowledge": dict(colors=5, ranks=5, max_information_tokens=8,
                                      max_life_tokens=3, observation_type=1),
    "Hanabi-Full-Minimal": dict(colors=5, ranks=5, max_information_tokens=8, max_life_tokens=3,
                                observation_type=0),
    "Hanabi-Small": dict(colors=2, ranks=5, hand_size=2, max_information_tokens=3,
                         max_life_tokens=1, observation_type=1),
    "Hanabi-Very-Small": dict(colors=1, ranks=5, hand_size=2, max_information_tokens=3,
                              max_life_tokens=1, observation_type=1),
}


class BatchedHanabiEnv(object):
  """`num_games` independent Hanabi games stepped together on one GPU.

  ```python
  env = rl_env.make_batched("Hanabi-Full", num_players=2, num_games=65536)
  obs = env.reset()                  # dict of tensors
  while True:
    actions = policy(obs["vectorized"], obs["legal_moves_mask"])   # int32 [num_games]
    obs, reward, done, info = env.step(actions)
  ```

  Observations are those of the player to act (`obs["current_player"]`), which is
  what the reference's trainers consume (agents/run_experiment.py:282-308).  With
  auto_reset (default) a finished game is re-dealt inside the same step, so
  `obs` of a game with done=1 is the first observation of its next episode.
  """

  def __init__(self, config, num_games, device=0, seeds=None, auto_reset=True):
    assert isinstance(config, dict), "Expected config to be of type dict."
    import torch
    self._torch = torch
    config = dict(config)
    base_seed = int(config.pop("seed", 1))
    if base_seed == -1:  # the reference draws from std::random_device (hanabi_game.cc:48-50)
      import random
      base_seed = random.SystemRandom().randrange(1, 2**31 - 1 - num_games)
    if seeds is None:
      seeds = [base_seed + i for i in range(num_games)]
    self.config = config
    self.num_games = int(num_games)
    self.auto_reset = bool(auto_reset)
    self.batch = pyhanabi.HanabiBatch(config, num_games, seeds, device=device)
    self.device = self.batch.device
    self.players = self.batch.num_players
    n, D, A = self.num_games, self.batch.obs_size, self.batch.max_moves
    self._obs = torch.zeros((n, D), dtype=torch.uint8, device=self.device)
    self._mask = torch.zeros((n, A), dtype=torch.uint8, device=self.device)
    self._cur = torch.zeros(n, dtype=torch.int32, device=self.device)
    self._reward = torch.zeros(n, dtype=torch.int32, device=self.device)
    self._done = torch.zeros(n, dtype=torch.uint8, device=self.device)
    self._score = torch.zeros(n, dtype=torch.int32, device=self.device)

  def _observation(self):
    return {"current_player": self._cur, "vectorized": self._obs, "legal_moves_mask": self._mask}

  def reset(self, which=None):
    """Starts new games (all, or those flagged in the uint8 tensor `which`)."""
    self.batch.reset(which)
    self.batch.observe(self._obs, self._mask, self._cur)
    return self._observation()

  def vectorized_observation_shape(self):
    return [self.batch.obs_size]

  def num_moves(self):
    return self.batch.max_moves

  def step(self, actions):
    """actions: int32 CUDA tensor [num_games] of move uids (rl_env.py:346-348).

    Returns (observation dict, reward int32 [n], done uint8 [n], info) with
    reward = score delta (rl_env.py:363) and info["score"] = score after the move
    (the finished episode's score where done)."""
    torch = self._torch
    if actions.dtype != torch.int32:
      actions = actions.to(torch.int32)
    self.batch.step_encode(actions, self._reward, self._done, self._obs, self._mask, self._cur,
                           scores=self._score, auto_reset=self.auto_reset)
    return self._observation(), self._reward, self._done, {"score": self._score}

  def legal_moves_as_float(self, illegal_value=float("-inf")):
    """The 0 / -inf vector the Rainbow agents add to their q-values
    (agents/run_experiment.py:260-279); tf-agents uses float32.min instead
    (training/tf_agents_lib/pyhanabi_env_wrapper.py:31-39)."""
    torch = self._torch
    out = torch.zeros(self._mask.shape, dtype=torch.float32, device=self.device)
    return out.masked_fill_(self._mask == 0, illegal_value)

  def episode_statistics(self, clear=False):
    return self.batch.read_stats(clear=clear)


def _preset(environment_name, num_players):
  if environment_name not in PRESETS:
    raise ValueError("Unknown environment {}".format(environment_name))
  config = dict(PRESETS[environment_name])
  config["players"] = num_players
  return config


def make_batched(environment_name="Hanabi-Full", num_players=2, num_games=1024, device=0, seed=1,
                 auto_reset=True):
  config = _preset(environment_name, num_players)
  config["seed"] = seed
  return BatchedHanabiEnv(config, num_games, device=device, auto_reset=auto_reset)
