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
from hanabi_b200.pyhanabi import color_char_to_idx

MOVE_TYPES = [_.name for _ in pyhanabi.HanabiMoveType]

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


class Environment(object):
  """Abstract environment interface (reference rl_env.py:28-64)."""

  def reset(self, config):
    raise NotImplementedError("Not implemented in Abstract Base class")

  def step(self, action):
    raise NotImplementedError("Not implemented in Abstract Base class")


def observation_dict(state, game, encoder, observation):
  """One player's observation as the reference's dict (rl_env.py:382-438): 14 keys
  including `legal_moves_as_int`, `vectorized` and the live `pyhanabi` handle."""
  d = {}
  d["current_player"] = state.cur_player()
  d["current_player_offset"] = observation.cur_player_offset()
  d["life_tokens"] = observation.life_tokens()
  d["information_tokens"] = observation.information_tokens()
  d["num_players"] = observation.num_players()
  d["deck_size"] = observation.deck_size()
  d["fireworks"] = dict(zip(pyhanabi.COLOR_CHAR, state.fireworks()))
  d["legal_moves"] = []
  d["legal_moves_as_int"] = []
  for move in observation.legal_moves():
    d["legal_moves"].append(move.to_dict())
    d["legal_moves_as_int"].append(game.get_move_uid(move))
  d["observed_hands"] = [[card.to_dict() for card in hand]
                         for hand in observation.observed_hands()]
  d["discard_pile"] = [card.to_dict() for card in observation.discard_pile()]
  d["card_knowledge"] = []
  for player_hints in observation.card_knowledge():
    hints = []
    for hint in player_hints:
      color = hint.color()
      hints.append({"color": pyhanabi.color_idx_to_char(color) if color is not None else None,
                    "rank": hint.rank()})
    d["card_knowledge"].append(hints)
  d["vectorized"] = encoder.encode(observation)
  d["pyhanabi"] = observation
  return d


class HanabiEnv(Environment):
  """The reference's single-game environment, unchanged in behaviour
  (rl_env.py:67-498): same config dict, same per-player observation dicts (all 14
  keys including `legal_moves_as_int`, `vectorized` and the live `pyhanabi`
  observation handle), same reward (score delta, rl_env.py:363) and done flag.
  It runs on the host-side per-object API of libpyhanabi.so; use
  BatchedHanabiEnv for throughput."""

  def __init__(self, config):
    assert isinstance(config, dict), "Expected config to be of type dict."
    self.game = pyhanabi.HanabiGame(config)
    self.observation_encoder = pyhanabi.ObservationEncoder(
        self.game, pyhanabi.ObservationEncoderType.CANONICAL)
    self.players = self.game.num_players()

  def _deal(self):
    while self.state.cur_player() == pyhanabi.CHANCE_PLAYER_ID:
      self.state.deal_random_card()

  def reset(self):
    self.state = self.game.new_initial_state()
    self._deal()
    obs = self._make_observation_all_players()
    obs["current_player"] = self.state.cur_player()
    return obs

  def vectorized_observation_shape(self):
    return self.observation_encoder.shape()

  def num_moves(self):
    return self.game.max_moves()

  def step(self, action):
    if isinstance(action, dict):
      action = self._build_move(action)
    elif isinstance(action, int):
      action = self.game.get_move(action)
    else:  # numpy integers are rejected, as in the reference (rl_env.py:349-351)
      raise ValueError("Expected action as dict or int, got: {}".format(action))
    last_score = self.state.score()
    self.state.apply_move(action)
    self._deal()
    observation = self._make_observation_all_players()
    done = self.state.is_terminal()
    reward = self.state.score() - last_score
    return (observation, reward, done, {})

  def _make_observation_all_players(self):
    return {
        "player_observations": [
            self._extract_dict_from_backend(pid, self.state.observation(pid))
            for pid in range(self.players)],
        "current_player": self.state.cur_player(),
    }

  def _extract_dict_from_backend(self, player_id, observation):
    return observation_dict(self.state, self.game, self.observation_encoder, observation)

  def _build_move(self, action):
    assert isinstance(action, dict), "Expected dict, got: {}".format(action)
    assert "action_type" in action, ("Action should contain `action_type`. "
                                     "action: {}").format(action)
    action_type = action["action_type"]
    assert (action_type in MOVE_TYPES), (
        "action_type: {} should be one of: {}".format(action_type, MOVE_TYPES))
    if action_type == "PLAY":
      move = pyhanabi.HanabiMove.get_play_move(card_index=action["card_index"])
    elif action_type == "DISCARD":
      move = pyhanabi.HanabiMove.get_discard_move(card_index=action["card_index"])
    elif action_type == "REVEAL_RANK":
      move = pyhanabi.HanabiMove.get_reveal_rank_move(
          target_offset=action["target_offset"], rank=action["rank"])
    elif action_type == "REVEAL_COLOR":
      assert isinstance(action["color"], str)
      move = pyhanabi.HanabiMove.get_reveal_color_move(
          target_offset=action["target_offset"], color=color_char_to_idx(action["color"]))
    else:
      raise ValueError("Unknown action_type: {}".format(action_type))
    legal_moves = self.state.legal_moves()
    assert (str(move) in map(str, legal_moves)), (
        "Illegal action: {}. Move should be one of : {}".format(move, legal_moves))
    return move


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

  def game_state(self, game_index):
    """Game `game_index` as a single-game pyhanabi.HanabiState (host side, on demand)."""
    return self.batch.export_state(game_index)

  def game_observations(self, game_index):
    """What the reference's HanabiEnv returns for one game -- {'player_observations':
    [14-key dict per player], 'current_player'} (rl_env.py:368-438) -- built on demand
    from game `game_index` of the batch, for debugging, logging and the GUI layer.  After
    batch.enable_history() every key is exact, including the order of the discard pile and
    `pyhanabi.last_moves()`; without it the pile is sorted by card type and last_moves()
    holds the last player move only."""
    state = self.batch.export_state(game_index)
    game = self.batch.host_game
    if getattr(self, "_host_encoder", None) is None:
      self._host_encoder = pyhanabi.ObservationEncoder(game, pyhanabi.ObservationEncoderType.CANONICAL)
    return {
        "player_observations": [
            observation_dict(state, game, self._host_encoder, state.observation(pid))
            for pid in range(self.players)],
        "current_player": state.cur_player(),
    }


def _preset(environment_name, num_players):
  if environment_name not in PRESETS:
    raise ValueError("Unknown environment {}".format(environment_name))
  config = dict(PRESETS[environment_name])
  config["players"] = num_players
  return config


def make(environment_name="Hanabi-Full", num_players=2, pyhanabi_path=None):
  """rl_env.make (reference rl_env.py:501-585): the single-game environment for
  one of the preset names.  Like the reference it passes no seed, so games are
  seeded from std::random_device."""
  if pyhanabi_path is not None:
    prefixes = (pyhanabi_path,)
    assert pyhanabi.try_cdef(prefixes=prefixes), "cdef failed to load"
    assert pyhanabi.try_load(prefixes=prefixes), "library failed to load"
  return HanabiEnv(config=_preset(environment_name, num_players))


def make_batched(environment_name="Hanabi-Full", num_players=2, num_games=1024, device=0, seed=1,
                 auto_reset=True):
  config = _preset(environment_name, num_players)
  config["seed"] = seed
  return BatchedHanabiEnv(config, num_games, device=device, auto_reset=auto_reset)


class Agent(object):
  """Agent interface (reference rl_env.py:592-641)."""

  def __init__(self, config, *args, **kwargs):
    raise NotImplementedError("Not implemented in Abstract Base class")

  def reset(self, config):
    raise NotImplementedError("Not implemented in Abstract Base class")

  def act(self, observation):
    raise NotImplementedError("Not implemented in Abstract Base class")
