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


def read_cdef(cdef_file):
  """The declarations between the header's `extern "C" {` and
  `} /* extern "C" */` marker lines (same rule as the reference's try_cdef,
  pyhanabi.py:56-69)."""
  reading_cdef = False
  cdef_string = ""
  for line in open(cdef_file).readlines():
    line = line.rstrip()
    if re.match("extern *\"C\" *{", line):
      reading_cdef = True
      continue
    elif re.match("} */[*] *extern *\"C\" *[*]/", line):
      reading_cdef = False
      continue
    if reading_cdef:
      cdef_string = cdef_string + line + "\n"
  return cdef_string


def declared_functions(cdef_file):
  """Names of all functions the header declares inside its extern "C" block."""
  text = re.sub(r"/\*.*?\*/", " ", read_cdef(cdef_file), flags=re.S)
  return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)))


def try_cdef(header=PYHANABI_HEADER, prefixes=DEFAULT_CDEF_PREFIXES):
  """Parse the library header; same contract as the reference's try_cdef."""
  global cdef_loaded_flag
  if cdef_loaded_flag:
    return True
  for prefix in prefixes:
    try:
      cdef_file = header if prefix is None else prefix + "/" + header
      ffi.cdef(read_cdef(cdef_file))
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
  if library is None and os.environ.get("HANABI_B200_LIB"):
    library = os.environ["HANABI_B200_LIB"]  # experiments: an alternative build of the library
    prefixes = (None,)
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


# ---------------------------------------------------------------------------
# Per-object API: the same classes and methods as the reference binding
# (hanabi_learning_environment/pyhanabi.py:117-972), over the same C symbols.
import enum  # noqa: E402


def _take_string(c_str):
  text = encode_ffi_string(c_str)
  lib.DeleteString(c_str)
  return text


def color_idx_to_char(color_idx):
  """Colour index -> one of "RYGWB"; -1 -> None (reference pyhanabi.py:117-133)."""
  assert isinstance(color_idx, int)
  return None if color_idx == -1 else COLOR_CHAR[color_idx]


def color_char_to_idx(color_char):
  """One of "RYGWB" -> colour index (reference pyhanabi.py:136-153)."""
  assert isinstance(color_char, str)
  if color_char not in COLOR_CHAR:
    raise ValueError("Invalid color: {}. Should be one of {}.".format(color_char, COLOR_CHAR))
  return COLOR_CHAR.index(color_char)


class HanabiCard(object):
  """A card: colour index (RYGWB order) and 0-based rank; -1/-1 is "unknown"."""

  def __init__(self, color, rank):
    self._color, self._rank = color, rank

  def color(self):
    return self._color

  def rank(self):
    return self._rank

  def valid(self):
    return self._color >= 0 and self._rank >= 0

  def __str__(self):
    return COLOR_CHAR[self._color] + str(self._rank + 1) if self.valid() else "XX"

  __repr__ = __str__

  def __eq__(self, other):
    return self._color == other.color() and self._rank == other.rank()

  def to_dict(self):
    return {"color": color_idx_to_char(self.color()), "rank": self.rank()}


class HanabiCardKnowledge(object):
  """Hint knowledge about one card: the directly hinted colour/rank (or None)
  and which values no hint has excluded yet."""

  def __init__(self, knowledge):
    self._knowledge = knowledge

  def color(self):
    return lib.KnownColor(self._knowledge) if lib.ColorWasHinted(self._knowledge) else None

  def color_plausible(self, color_index):
    return lib.ColorIsPlausible(self._knowledge, color_index)

  def rank(self):
    return lib.KnownRank(self._knowledge) if lib.RankWasHinted(self._knowledge) else None

  def rank_plausible(self, rank_index):
    return lib.RankIsPlausible(self._knowledge, rank_index)

  def __str__(self):
    return _take_string(lib.CardKnowledgeToString(self._knowledge))

  __repr__ = __str__

  def to_dict(self):
    return {"color": color_idx_to_char(self.color()), "rank": self.rank()}


class HanabiMoveType(enum.IntEnum):
  INVALID = 0
  PLAY = 1
  DISCARD = 2
  REVEAL_COLOR = 3
  REVEAL_RANK = 4
  DEAL = 5


class HanabiMove(object):
  """Owns one C move object."""

  def __init__(self, move):
    assert move is not None
    self._move = move

  @property
  def c_move(self):
    return self._move

  def type(self):
    return HanabiMoveType(lib.MoveType(self._move))

  def card_index(self):
    return lib.CardIndex(self._move)

  def target_offset(self):
    return lib.TargetOffset(self._move)

  def color(self):
    return lib.MoveColor(self._move)

  def rank(self):
    return lib.MoveRank(self._move)

  @staticmethod
  def _make(fn, *args):
    c_move = ffi.new("pyhanabi_move_t*")
    assert fn(*(args + (c_move,)))
    return HanabiMove(c_move)

  @staticmethod
  def get_discard_move(card_index):
    return HanabiMove._make(lib.GetDiscardMove, card_index)

  @staticmethod
  def get_play_move(card_index):
    return HanabiMove._make(lib.GetPlayMove, card_index)

  @staticmethod
  def get_reveal_color_move(target_offset, color):
    return HanabiMove._make(lib.GetRevealColorMove, target_offset, color)

  @staticmethod
  def get_reveal_rank_move(target_offset, rank):
    return HanabiMove._make(lib.GetRevealRankMove, target_offset, rank)

  def __str__(self):
    return _take_string(lib.MoveToString(self._move))

  __repr__ = __str__

  def __del__(self):
    if self._move is not None and lib is not None:
      lib.DeleteMove(self._move)
      self._move = None

  def to_dict(self):
    kind = self.type()
    d = {"action_type": kind.name}
    if kind in (HanabiMoveType.PLAY, HanabiMoveType.DISCARD):
      d["card_index"] = self.card_index()
    elif kind == HanabiMoveType.REVEAL_COLOR:
      d["target_offset"] = self.target_offset()
      d["color"] = color_idx_to_char(self.color())
    elif kind == HanabiMoveType.REVEAL_RANK:
      d["target_offset"] = self.target_offset()
      d["rank"] = self.rank()
    elif kind == HanabiMoveType.DEAL:
      d["color"] = color_idx_to_char(self.color())
      d["rank"] = self.rank()
    else:
      raise ValueError("Unsupported move: {}".format(self))
    return d


def _bits(mask):
  return [i for i in range(8) if mask & (1 << i)]


class HanabiHistoryItem(object):
  """A move together with what it did (scored, token gained, cards touched)."""

  def __init__(self, item):
    self._item = item

  def move(self):
    c_move = ffi.new("pyhanabi_move_t*")
    lib.HistoryItemMove(self._item, c_move)
    return HanabiMove(c_move)

  def player(self):
    return lib.HistoryItemPlayer(self._item)

  def scored(self):
    return bool(lib.HistoryItemScored(self._item))

  def information_token(self):
    return bool(lib.HistoryItemInformationToken(self._item))

  def color(self):
    return lib.HistoryItemColor(self._item)

  def rank(self):
    return lib.HistoryItemRank(self._item)

  def card_info_revealed(self):
    return _bits(lib.HistoryItemRevealBitmask(self._item))

  def card_info_newly_revealed(self):
    return _bits(lib.HistoryItemNewlyRevealedBitmask(self._item))

  def deal_to_player(self):
    return lib.HistoryItemDealToPlayer(self._item)

  def __str__(self):
    return _take_string(lib.HistoryItemToString(self._item))

  __repr__ = __str__

  def __del__(self):
    if self._item is not None and lib is not None:
      lib.DeleteHistoryItem(self._item)
      self._item = None


class HanabiEndOfGameType(enum.IntEnum):
  NOT_FINISHED = 0
  OUT_OF_LIFE_TOKENS = 1
  OUT_OF_CARDS = 2
  COMPLETED_FIREWORKS = 3


def _card_from(fn, *args):
  c_card = ffi.new("pyhanabi_card_t*")
  fn(*(args + (c_card,)))
  return HanabiCard(c_card.color, c_card.rank)


def _moves_from_list(c_list):
  out = []
  for i in range(lib.NumMoves(c_list)):
    c_move = ffi.new("pyhanabi_move_t*")
    lib.GetMove(c_list, i, c_move)
    out.append(HanabiMove(c_move))
  lib.DeleteMoveList(c_list)
  return out


class HanabiState(object):
  """One game in progress; chance events are explicit (cur_player() == -1)."""

  def __init__(self, game, c_state=None):
    self._state = ffi.new("pyhanabi_state_t*")
    if c_state is None:
      self._game = game.c_game
      lib.NewState(self._game, self._state)
    else:
      # NOTE: like the reference, `game` is ignored and the copied state's game is used
      self._game = ffi.new("pyhanabi_game_t*")
      self._game.game = ffi.cast("void*", lib.StateParentGame(c_state))
      lib.CopyState(c_state, self._state)

  def copy(self):
    return HanabiState(None, self._state)

  def observation(self, player):
    return HanabiObservation(self._state, self._game, player)

  def apply_move(self, move):
    lib.StateApplyMove(self._state, move.c_move)

  def cur_player(self):
    return lib.StateCurPlayer(self._state)

  def deck_size(self):
    return lib.StateDeckSize(self._state)

  def discard_pile(self):
    return [_card_from(lib.StateGetDiscard, self._state, i)
            for i in range(lib.StateDiscardPileSize(self._state))]

  def fireworks(self):
    return [lib.StateFireworks(self._state, c) for c in range(lib.NumColors(self._game))]

  def deal_random_card(self):
    lib.StateDealRandomCard(self._state)

  def player_hands(self):
    hands = []
    for pid in range(self.num_players()):
      hands.append([_card_from(lib.StateGetHandCard, self._state, pid, i)
                    for i in range(lib.StateGetHandSize(self._state, pid))])
    return hands

  def information_tokens(self):
    return lib.StateInformationTokens(self._state)

  def end_of_game_status(self):
    return HanabiEndOfGameType(lib.StateEndOfGameStatus(self._state))

  def is_terminal(self):
    return lib.StateEndOfGameStatus(self._state) != HanabiEndOfGameType.NOT_FINISHED

  def legal_moves(self):
    return _moves_from_list(lib.StateLegalMoves(self._state))

  def move_is_legal(self, move):
    return lib.MoveIsLegal(self._state, move.c_move)

  def card_playable_on_fireworks(self, color, rank):
    return lib.CardPlayableOnFireworks(self._state, color, rank)

  def life_tokens(self):
    return lib.StateLifeTokens(self._state)

  def num_players(self):
    return lib.StateNumPlayers(self._state)

  def score(self):
    return lib.StateScore(self._state)

  def move_history(self):
    out = []
    for i in range(lib.StateLenMoveHistory(self._state)):
      c_item = ffi.new("pyhanabi_history_item_t*")
      lib.StateGetMoveHistory(self._state, i, c_item)
      out.append(HanabiHistoryItem(c_item))
    return out

  def __str__(self):
    return _take_string(lib.StateToString(self._state))

  __repr__ = __str__

  def __del__(self):
    if self._state is not None and lib is not None:
      lib.DeleteState(self._state)
      self._state = None
      self._game = None


class AgentObservationType(enum.IntEnum):
  """What an agent may see (hanabi_lib/hanabi_game.h:28-39)."""
  MINIMAL = 0
  CARD_KNOWLEDGE = 1
  SEER = 2


class HanabiGame(object):
  """Rule set plus the game's own RNG stream.  `params` keys: players, colors,
  ranks, hand_size, max_information_tokens, max_life_tokens, seed,
  random_start_player, observation_type (hanabi_lib/hanabi_game.cc:29-47)."""

  def __init__(self, params=None):
    _require_lib()
    self._game = ffi.new("pyhanabi_game_t*")
    if params is None:
      lib.NewDefaultGame(self._game)
    else:
      keep = []
      for key, value in params.items():
        keep.append(ffi.new("char[]", str(key).encode("ascii")))
        keep.append(ffi.new("char[]", str(value).encode("ascii")))
      c_array = ffi.new("char * [" + str(max(len(keep), 1)) + "]", keep if keep else [ffi.NULL])
      lib.NewGame(self._game, len(keep), c_array)

  def new_initial_state(self):
    return HanabiState(self)

  @property
  def c_game(self):
    return self._game

  def __del__(self):
    if self._game is not None and lib is not None:
      lib.DeleteGame(self._game)
      self._game = None

  def parameter_string(self):
    return _take_string(lib.GameParamString(self._game))

  def num_players(self):
    return lib.NumPlayers(self._game)

  def num_colors(self):
    return lib.NumColors(self._game)

  def num_ranks(self):
    return lib.NumRanks(self._game)

  def hand_size(self):
    return lib.HandSize(self._game)

  def max_information_tokens(self):
    return lib.MaxInformationTokens(self._game)

  def max_life_tokens(self):
    return lib.MaxLifeTokens(self._game)

  def observation_type(self):
    return AgentObservationType(lib.ObservationType(self._game))

  def max_moves(self):
    return lib.MaxMoves(self._game)

  def num_cards(self, color, rank):
    return lib.NumCards(self._game, color, rank)

  def get_move_uid(self, move):
    return lib.GetMoveUid(self._game, move.c_move)

  def get_move(self, move_uid):
    c_move = ffi.new("pyhanabi_move_t*")
    lib.GetMoveByUid(self._game, move_uid, c_move)
    return HanabiMove(c_move)


class HanabiObservation(object):
  """One player's view: own cards hidden, every player index relative to the
  observer (who is always index 0)."""

  def __init__(self, state, game, player):
    self._observation = ffi.new("pyhanabi_observation_t*")
    self._game = game
    lib.NewObservation(state, player, self._observation)

  def __str__(self):
    return _take_string(lib.ObsToString(self._observation))

  __repr__ = __str__

  def __del__(self):
    if self._observation is not None and lib is not None:
      lib.DeleteObservation(self._observation)
      self._observation = None
      self._game = None

  def observation(self):
    return self._observation

  def cur_player_offset(self):
    return lib.ObsCurPlayerOffset(self._observation)

  def num_players(self):
    return lib.ObsNumPlayers(self._observation)

  def observed_hands(self):
    return [[_card_from(lib.ObsGetHandCard, self._observation, pid, i)
             for i in range(lib.ObsGetHandSize(self._observation, pid))]
            for pid in range(self.num_players())]

  def card_knowledge(self):
    out = []
    for pid in range(self.num_players()):
      player = []
      for i in range(lib.ObsGetHandSize(self._observation, pid)):
        c_knowledge = ffi.new("pyhanabi_card_knowledge_t*")
        lib.ObsGetHandCardKnowledge(self._observation, pid, i, c_knowledge)
        player.append(HanabiCardKnowledge(c_knowledge))
      out.append(player)
    return out

  def discard_pile(self):
    return [_card_from(lib.ObsGetDiscard, self._observation, i)
            for i in range(lib.ObsDiscardPileSize(self._observation))]

  def fireworks(self):
    return [lib.ObsFireworks(self._observation, c) for c in range(lib.NumColors(self._game))]

  def deck_size(self):
    return lib.ObsDeckSize(self._observation)

  def last_moves(self):
    out = []
    for i in range(lib.ObsNumLastMoves(self._observation)):
      c_item = ffi.new("pyhanabi_history_item_t*")
      lib.ObsGetLastMove(self._observation, i, c_item)
      out.append(HanabiHistoryItem(c_item))
    return out

  def information_tokens(self):
    return lib.ObsInformationTokens(self._observation)

  def life_tokens(self):
    return lib.ObsLifeTokens(self._observation)

  def legal_moves(self):
    out = []
    for i in range(lib.ObsNumLegalMoves(self._observation)):
      c_move = ffi.new("pyhanabi_move_t*")
      lib.ObsGetLegalMove(self._observation, i, c_move)
      out.append(HanabiMove(c_move))
    return out

  def card_playable_on_fireworks(self, color, rank):
    return lib.ObsCardPlayableOnFireworks(self._observation, color, rank)


class ObservationEncoderType(enum.IntEnum):
  CANONICAL = 0


class ObservationEncoder(object):
  """Canonical bit-vector encoder.  encode() returns a list of ints like the
  reference (pyhanabi.py:963-972) but reads the bits as bytes instead of
  parsing the comma-separated string."""

  def __init__(self, game, enc_type=ObservationEncoderType.CANONICAL):
    self._game = game.c_game
    self._encoder = ffi.new("pyhanabi_observation_encoder_t*")
    lib.NewObservationEncoder(self._encoder, self._game, enc_type)
    self._shape = None

  def __del__(self):
    if self._encoder is not None and lib is not None:
      lib.DeleteObservationEncoder(self._encoder)
      self._encoder = None
      self._game = None

  def shape(self):
    if self._shape is None:
      self._shape = [int(x) for x in _take_string(lib.ObservationShape(self._encoder)).split(",")]
    return list(self._shape)

  def encode(self, observation):
    if hasattr(lib, "EncodeObservationBytes"):
      n = self.shape()[0]
      buf = ffi.new("uint8_t[]", n)
      lib.EncodeObservationBytes(self._encoder, observation.observation(), buf, n)
      return list(bytes(ffi.buffer(buf, n)))
    text = _take_string(lib.EncodeObservation(self._encoder, observation.observation()))
    return [int(x) for x in text.split(",")]


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
    self._params = {k: v for k, v in dict(params).items() if k != "seed"}
    self._host_game = None
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

  def step_encode_random(self, actions, seed, step, rewards=None, dones=None, obs=None, mask=None,
                         cur_player=None, scores=None, auto_reset=True, random_actions=None):
    """step_encode with the fused uniform-random policy (BatchStepEncodeRandom): applies
    `actions`, then writes a uniformly random legal move of every game's new state into
    `random_actions` (default: in place into `actions`)."""
    torch = self._torch
    n = self.num_games
    optr, pitch = self._obs_ptr(obs)
    out = actions if random_actions is None else random_actions
    _check(lib.BatchStepEncodeRandom(
        self._c, self._p(actions, torch.int32, n, "int32_t*"),
        self._p(rewards, torch.int32, n, "int32_t*"), self._p(dones, torch.uint8, n, "uint8_t*"),
        self._p(scores, torch.int32, n, "int32_t*"), optr, pitch,
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)),
        self._p(out, torch.int32, n, "int32_t*"), int(seed) & (2**64 - 1), int(step) & (2**64 - 1),
        self._stream()), "BatchStepEncodeRandom")

  def set_game_index_base(self, first_game_index):
    """Global index of this batch's game 0 (shards of a larger set of games): the fused
    random policy is keyed (seed, step, first_game_index + i)."""
    _check(lib.BatchSetGameIndexBase(self._c, int(first_game_index)), "BatchSetGameIndexBase")
    self.game_index_base = int(first_game_index)

  def random_legal_actions(self, actions, seed, step):
    """Uniformly random legal move of every game's current state (the draw the fused
    policy of step_encode_random makes), written into `actions`."""
    torch = self._torch
    _check(lib.BatchRandomLegalActions(self._c, int(seed) & (2**64 - 1), int(step) & (2**64 - 1),
                                       self._p(actions, torch.int32, self.num_games, "int32_t*"),
                                       self._stream()), "BatchRandomLegalActions")

  def bind_step_encode_random(self, actions, seed, rewards=None, dones=None, obs=None, mask=None,
                              cur_player=None, scores=None, auto_reset=True, stream=None):
    """Pre-bound in-place variant of step_encode_random for tight loops: run(step)."""
    torch = self._torch
    n = self.num_games
    optr, pitch = self._obs_ptr(obs)
    act = self._p(actions, torch.int32, n, "int32_t*")
    head = (self._c, act, self._p(rewards, torch.int32, n, "int32_t*"),
            self._p(dones, torch.uint8, n, "uint8_t*"), self._p(scores, torch.int32, n, "int32_t*"),
            optr, pitch, self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
            self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), act,
            int(seed) & (2**64 - 1))
    tail = (self._stream() if stream is None else ffi.cast("void*", stream.cuda_stream),)
    fn = lib.BatchStepEncodeRandom
    keep = (actions, rewards, dones, obs, mask, cur_player, scores)

    def run(step, _head=head, _tail=tail, _fn=fn, _keep=keep):
      if _fn(*(_head + (step,) + _tail)) != 0:
        raise RuntimeError("BatchStepEncodeRandom failed: %s" %
                           encode_ffi_string(lib.BatchLastError()))
    return run

  # -- bit-packed observations (int32 tensors [num_games, packed_words])
  @property
  def packed_words(self):
    return int(lib.BatchPackedObservationWords(self._c))

  def step_encode_packed(self, actions, rewards=None, dones=None, obs_bits=None, mask=None,
                         cur_player=None, scores=None, auto_reset=True):
    torch = self._torch
    n = self.num_games
    _check(lib.BatchStepEncodePacked(
        self._c, self._p(actions, torch.int32, n, "int32_t*"),
        self._p(rewards, torch.int32, n, "int32_t*"), self._p(dones, torch.uint8, n, "uint8_t*"),
        self._p(scores, torch.int32, n, "int32_t*"),
        self._p(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), self._stream()),
           "BatchStepEncodePacked")

  def observe_packed(self, obs_bits, mask=None, cur_player=None, observer=-1):
    torch = self._torch
    n = self.num_games
    _check(lib.BatchObservePacked(
        self._c, int(observer), self._p(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), self._stream()), "BatchObservePacked")

  # -- the observation as the network's input (float32 / bfloat16 / float16 rows, K padded)
  def _net(self, net_input):
    n = self.num_games
    if (net_input.dim() != 2 or net_input.stride(1) != 1 or int(net_input.shape[0]) != n or
        not net_input.is_cuda):
      raise ValueError("net_input must be a CUDA tensor [num_games, pitch] with unit column stride")
    return (ffi.cast("void*", net_input.data_ptr()), _gemm_dtype_code(net_input),
            int(net_input.stride(0)))

  def step_encode_net(self, actions, net_input, rewards=None, dones=None, obs_bits=None, mask=None,
                      cur_player=None, scores=None, auto_reset=True):
    """step_encode writing the observation straight into the first GEMM's input matrix
    `net_input` [num_games, pitch] (BatchStepEncodeNet); obs_bits optionally gets the
    bit-packed rows too."""
    torch = self._torch
    n = self.num_games
    ptr, code, pitch = self._net(net_input)
    _check(lib.BatchStepEncodeNet(
        self._c, self._p(actions, torch.int32, n, "int32_t*"),
        self._p(rewards, torch.int32, n, "int32_t*"), self._p(dones, torch.uint8, n, "uint8_t*"),
        self._p(scores, torch.int32, n, "int32_t*"), ptr, code, pitch,
        self._p(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), self._stream()),
           "BatchStepEncodeNet")

  def observe_net(self, net_input, obs_bits=None, mask=None, cur_player=None, observer=-1):
    torch = self._torch
    n = self.num_games
    ptr, code, pitch = self._net(net_input)
    _check(lib.BatchObserveNet(
        self._c, int(observer), ptr, code, pitch,
        self._p(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),
        self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._p(cur_player, torch.int32, n, "int32_t*"), self._stream()), "BatchObserveNet")

  def bind_step_encode(self, actions, rewards=None, dones=None, obs=None, mask=None,
                       cur_player=None, scores=None, auto_reset=True, stream=None):
    """Pre-validated, pre-cast variant of step_encode for tight loops: returns a
    zero-argument callable that launches the fused kernel on `stream` (default:
    the stream current at bind time) with the given tensors.  The tensors must
    stay alive and in place while the callable is used."""
    torch = self._torch
    n = self.num_games
    optr, pitch = self._obs_ptr(obs)
    args = (self._c, self._p(actions, torch.int32, n, "int32_t*"),
            self._p(rewards, torch.int32, n, "int32_t*"), self._p(dones, torch.uint8, n, "uint8_t*"),
            self._p(scores, torch.int32, n, "int32_t*"), optr, pitch,
            self._p(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
            self._p(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)),
            self._stream() if stream is None else ffi.cast("void*", stream.cuda_stream))
    fn = lib.BatchStepEncodeScores
    keep = (actions, rewards, dones, obs, mask, cur_player, scores)

    def run(_args=args, _fn=fn, _keep=keep):
      if _fn(*_args) != 0:
        raise RuntimeError("BatchStepEncodeScores failed: %s" %
                           encode_ffi_string(lib.BatchLastError()))
    return run

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

  def step_encode_host_packed(self, actions, rewards=None, dones=None, obs_bits=None, mask=None,
                              cur_player=None, scores=None, auto_reset=True):
    torch = self._torch
    n = self.num_games
    _check(lib.BatchStepEncodeHostPacked(
        self._c, self._h(actions, torch.int32, n, "int32_t*"),
        self._h(rewards, torch.int32, n, "int32_t*"), self._h(dones, torch.uint8, n, "uint8_t*"),
        self._h(scores, torch.int32, n, "int32_t*"),
        self._h(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),
        self._h(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
        self._h(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)), self._stream()),
           "BatchStepEncodeHostPacked")

  def bind_step_encode_host_range(self, actions, rewards=None, dones=None, obs=None, mask=None,
                                  cur_player=None, scores=None, auto_reset=True, obs_bits=None):
    """Pre-bound BatchStepEncodeHostRange: returns (launch(first, count), wait(first)).
    launch queues the step of games [first, first + count) and returns at once; wait
    blocks until that range's host buffers are filled (first < 0: everything)."""
    torch = self._torch
    n = self.num_games
    head = (self._h(actions, torch.int32, n, "int32_t*"), self._h(rewards, torch.int32, n, "int32_t*"),
            self._h(dones, torch.uint8, n, "uint8_t*"), self._h(scores, torch.int32, n, "int32_t*"))
    tail = (self._h(mask, torch.uint8, n * self.max_moves, "uint8_t*"),
            self._h(cur_player, torch.int32, n, "int32_t*"), int(bool(auto_reset)))
    if obs_bits is not None:  # bit-packed rows in the host buffer instead of bytes
      fn = lib.BatchStepEncodeHostRangePacked
      args = head + (self._h(obs_bits, torch.int32, n * self.packed_words, "uint32_t*"),) + tail
    else:
      fn = lib.BatchStepEncodeHostRange
      args = head + self._h_obs(obs) + tail
    keep = (actions, rewards, dones, obs, obs_bits, mask, cur_player, scores)
    c = self._c

    def launch(first, count, _args=args, _keep=keep, _fn=fn):
      if _fn(c, first, count, *_args) != 0:
        raise RuntimeError("BatchStepEncodeHostRange failed: %s" %
                           encode_ffi_string(lib.BatchLastError()))

    def wait(first=-1):
      if lib.BatchHostWait(c, first) != 0:
        raise RuntimeError("BatchHostWait failed: %s" % encode_ffi_string(lib.BatchLastError()))
    return launch, wait

  def set_host_link(self, packed):
    """True (default): the uint8 rows of the host entry points cross PCIe bit-packed and are
    expanded by the host worker pool; False: they cross as bytes (BatchSetHostLink)."""
    _check(lib.BatchSetHostLink(self._c, int(bool(packed))), "BatchSetHostLink")

  @property
  def host_link_packed(self):
    return bool(lib.BatchGetHostLink(self._c))

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

  # -- the reference's hand-written policies, batched
  def enable_history(self):
    """Keep the last non-deal moves per game (needed by agent_act kind 0); call before reset()."""
    _check(lib.BatchEnableHistory(self._c), "BatchEnableHistory")

  def agent_act(self, kind, seat_mask, actions, agent_state=None, seat_slot=None):
    """kind 0 = RuleBasedAgent, 1 = SimpleAgent; writes actions[i] where the player
    to act sits in seat_mask.  agent_state: int32 CUDA tensor [num_games, num_slots]."""
    torch = self._torch
    n = self.num_games
    slots = 0
    sptr = ffi.NULL
    if agent_state is not None:
      if agent_state.dtype != torch.int32 or agent_state.dim() != 2 or agent_state.shape[0] != n:
        raise ValueError("agent_state must be an int32 [num_games, num_slots] tensor")
      slots = int(agent_state.shape[1])
      sptr = self._p(agent_state, torch.int32, n * slots, "uint32_t*")
    cslot = ffi.NULL if seat_slot is None else ffi.new("int32_t[]", [int(x) for x in seat_slot])
    _check(lib.BatchAgentAct(self._c, int(kind), int(seat_mask), cslot, sptr, slots,
                             self._p(actions, torch.int32, n, "int32_t*"), self._stream()),
           "BatchAgentAct")

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

  def error_flags(self, clear=True):
    """uint8[num_games] CUDA tensor: 1 where an illegal action was submitted (the game
    was left unchanged) since the flags were last cleared (BatchErrorFlags)."""
    torch = self._torch
    flags = torch.zeros(self.num_games, dtype=torch.uint8, device=self.device)
    _check(lib.BatchErrorFlags(self._c, self._p(flags, torch.uint8, self.num_games, "uint8_t*"),
                               ffi.NULL, int(bool(clear)), self._stream()), "BatchErrorFlags")
    return flags

  def export_state(self, game_index):
    """Game `game_index` as a single-game HanabiState (BatchExportGame +
    StateFromBatchGame): observations, encoders, rule-based agents and the GUI layer
    of the per-object API work on it.  With enable_history() the discard pile keeps its
    order and the move history holds everything an observer can see (the newest P player
    moves and the deals after them); without it the pile comes back sorted by card type and
    the history holds the last player move only."""
    if getattr(self, "_host_game", None) is None:
      self._host_game = HanabiGame(dict(self._params))
    rec = ffi.new("int32_t[290]")
    _check(lib.BatchExportGame(self._c, int(game_index), rec, self._stream()), "BatchExportGame")
    state = HanabiState(self._host_game)
    lib.StateFromBatchGame(self._host_game.c_game, rec, state._state)
    return state

  @property
  def host_game(self):
    """A single-game HanabiGame with this batch's parameters (move uids, encoders)."""
    if getattr(self, "_host_game", None) is None:
      self._host_game = HanabiGame(dict(self._params))
    return self._host_game

  def set_l2_residency(self, max_bytes=0):
    """Keeps the packed state in L2 between steps (BatchSetL2Residency); returns the
    number of state bytes covered.  max_bytes < 0 switches it off."""
    covered = ffi.new("int64_t*")
    _check(lib.BatchSetL2Residency(self._c, int(max_bytes), covered), "BatchSetL2Residency")
    return int(covered[0])

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


def bind_select_action(values, mask, epsilon, seed, out, support=None, stream=None):
  """Pre-validated variant of select_action for tight loops: returns run(step)."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  atoms = int(values.shape[2]) if values is not None and values.dim() == 3 else 0
  s = _cur_stream(mask) if stream is None else ffi.cast("void*", stream.cuda_stream)
  head = (_tp(values, torch.float32, "float*"), atoms, _tp(support, torch.float32, "float*"),
          _tp(mask, torch.uint8, "uint8_t*"), n, A, float(epsilon), int(seed) & (2**64 - 1))
  tail = (_tp(out, torch.int32, "int32_t*"), s)
  fn = lib.BatchSelectAction
  keep = (values, mask, out, support)

  def run(step, _head=head, _tail=tail, _fn=fn, _keep=keep):
    if _fn(*(_head + (step,) + _tail)) != 0:
      raise RuntimeError("BatchSelectAction failed: %s" % encode_ffi_string(lib.BatchLastError()))
  return run


def _gemm_dtype_code(t):
  import torch
  codes = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
  if t.dtype not in codes:
    raise ValueError("expected float32, bfloat16 or float16, got %s" % (t.dtype,))
  return codes[t.dtype]


def bind_select_action_c51(logits, num_atoms, mask, epsilon, seed, out, support, stream=None):
  """Masked epsilon-greedy on the raw output of a C51 head's last GEMM
  (BatchSelectActionC51): logits is a 2-d CUDA tensor [n, >= A*num_atoms] in float32,
  bfloat16 or float16 whose rows may be padded (row pitch = logits.stride(0)).
  Returns run(step)."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  if logits.dim() != 2 or logits.stride(1) != 1 or int(logits.shape[0]) != n:
    raise ValueError("logits must be [n, pitch] with unit column stride")
  head = (ffi.cast("void*", logits.data_ptr()), _gemm_dtype_code(logits), int(logits.stride(0)),
          int(num_atoms), _tp(support, torch.float32, "float*"), _tp(mask, torch.uint8, "uint8_t*"),
          n, A, float(epsilon), int(seed) & (2**64 - 1))
  s = _cur_stream(mask) if stream is None else ffi.cast("void*", stream.cuda_stream)
  tail = (_tp(out, torch.int32, "int32_t*"), s)
  fn = lib.BatchSelectActionC51
  keep = (logits, mask, out, support)

  def run(step, _head=head, _tail=tail, _fn=fn, _keep=keep):
    if _fn(*(_head + (step,) + _tail)) != 0:
      raise RuntimeError("BatchSelectActionC51 failed: %s" % encode_ffi_string(lib.BatchLastError()))
  return run


def select_action_c51(logits, num_atoms, mask, epsilon, seed, step, support, out=None):
  import torch
  if out is None:
    out = torch.empty(int(mask.shape[0]), dtype=torch.int32, device=mask.device)
  with torch.cuda.device(mask.device):
    bind_select_action_c51(logits, num_atoms, mask, epsilon, seed, out, support)(int(step) & (2**64 - 1))
  return out


def sample_masked_categorical(logits, mask, seed, step, greedy=False, actions=None, log_probs=None):
  """Masked categorical policy head of the PPO / REINFORCE actors
  (BatchSampleMaskedCategorical): logits [n, >= A] float32 / bfloat16 / float16, mask uint8
  [n, A].  Returns (actions int32 [n], log_probs float32 [n])."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  if logits.dim() != 2 or logits.stride(1) != 1 or int(logits.shape[0]) != n:
    raise ValueError("logits must be [n, pitch] with unit column stride")
  if actions is None:
    actions = torch.empty(n, dtype=torch.int32, device=mask.device)
  if log_probs is None:
    log_probs = torch.empty(n, dtype=torch.float32, device=mask.device)
  with torch.cuda.device(mask.device):
    _check(lib.BatchSampleMaskedCategorical(
        ffi.cast("void*", logits.data_ptr()), _gemm_dtype_code(logits), int(logits.stride(0)),
        _tp(mask, torch.uint8, "uint8_t*"), n, A, int(bool(greedy)), int(seed) & (2**64 - 1),
        int(step) & (2**64 - 1), _tp(actions, torch.int32, "int32_t*"),
        _tp(log_probs, torch.float32, "float*"), _cur_stream(mask)), "BatchSampleMaskedCategorical")
  return actions, log_probs


def bind_expand_packed(obs_bits, obs_size, out, stream=None):
  """Bit-packed observation rows (int32 [n, packed_words]) -> `out` [n, pitch] of 0/1 in
  float32 / bfloat16 / float16, columns beyond obs_size zeroed (BatchExpandPacked).
  Returns run()."""
  import torch
  _require_lib()
  n, words = int(obs_bits.shape[0]), int(obs_bits.shape[1])
  if out.dim() != 2 or out.stride(1) != 1 or int(out.shape[0]) != n:
    raise ValueError("out must be [n, pitch] with unit column stride")
  args = (_tp(obs_bits, torch.int32, "uint32_t*"), words, int(obs_size), n,
          ffi.cast("void*", out.data_ptr()), int(out.stride(0)), _gemm_dtype_code(out),
          _cur_stream(out) if stream is None else ffi.cast("void*", stream.cuda_stream))
  fn = lib.BatchExpandPacked
  keep = (obs_bits, out)

  def run(_args=args, _fn=fn, _keep=keep):
    if _fn(*_args) != 0:
      raise RuntimeError("BatchExpandPacked failed: %s" % encode_ffi_string(lib.BatchLastError()))
  return run


def expand_packed(obs_bits, obs_size, out):
  import torch
  with torch.cuda.device(out.device):
    bind_expand_packed(obs_bits, obs_size, out)()
  return out


def host_random_legal_actions(mask, seed, step, out=None, threads=0, first_game=0):
  """Uniform random legal move per game on the HOST (CPU uint8 mask [n, A] ->
  CPU int32 actions [n]); bit-identical to select_action(None, mask, 1.0, seed, step).
  Row i is game first_game + i of the (seed, step, game) key."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  if mask.is_cuda or mask.dtype != torch.uint8 or not mask.is_contiguous():
    raise ValueError("mask must be a contiguous CPU uint8 tensor")
  if out is None:
    out = torch.empty(n, dtype=torch.int32)
  _check(lib.HostRandomLegalActionsAt(ffi.cast("uint8_t*", mask.data_ptr()), n, A,
                                      int(seed) & (2**64 - 1), int(step) & (2**64 - 1),
                                      int(first_game), ffi.cast("int32_t*", out.data_ptr()),
                                      int(threads)),
         "HostRandomLegalActions")
  return out


def bind_host_random_legal_actions(mask, seed, out, first_game=0, threads=0):
  """Pre-bound host_random_legal_actions for tight loops: run(step)."""
  import torch
  _require_lib()
  n, A = int(mask.shape[0]), int(mask.shape[1])
  if mask.is_cuda or mask.dtype != torch.uint8 or not mask.is_contiguous():
    raise ValueError("mask must be a contiguous CPU uint8 tensor")
  if out.is_cuda or out.dtype != torch.int32 or out.numel() < n:
    raise ValueError("out must be a CPU int32 tensor with one element per game")
  head = (ffi.cast("uint8_t*", mask.data_ptr()), n, A, int(seed) & (2**64 - 1))
  tail = (int(first_game), ffi.cast("int32_t*", out.data_ptr()), int(threads))
  fn = lib.HostRandomLegalActionsAt
  keep = (mask, out)

  def run(step, _head=head, _tail=tail, _fn=fn, _keep=keep):
    if _fn(*(_head + (step,) + _tail)) != 0:
      raise RuntimeError("HostRandomLegalActions failed: %s" % encode_ffi_string(lib.BatchLastError()))
  return run


def host_expand_packed(obs_bits, obs_size, out=None, threads=0):
  """Bit-packed observation rows (CPU int32 [n, packed_words]) -> CPU uint8 rows of 0/1
  ([n, >= obs_size]) on the host worker pool: the CPU twin of the step kernel's byte
  expansion (HostExpandPacked), what the packed host link runs behind BatchHostWait."""
  import torch
  _require_lib()
  if obs_bits.is_cuda or obs_bits.dtype != torch.int32 or obs_bits.dim() != 2 or not obs_bits.is_contiguous():
    raise ValueError("obs_bits must be a contiguous CPU int32 [n, packed_words] tensor")
  n, words = int(obs_bits.shape[0]), int(obs_bits.shape[1])
  if out is None:
    out = torch.empty((n, int(obs_size)), dtype=torch.uint8)
  if (out.is_cuda or out.dtype != torch.uint8 or out.dim() != 2 or out.stride(1) != 1 or
      out.shape[0] < n or out.shape[1] < obs_size):
    raise ValueError("out must be a CPU uint8 [n, >= obs_size] tensor")
  _check(lib.HostExpandPacked(ffi.cast("uint32_t*", obs_bits.data_ptr()), words, int(obs_size), n,
                              ffi.cast("uint8_t*", out.data_ptr()), int(out.stride(0)), int(threads)),
         "HostExpandPacked")
  return out


def host_policy_threads():
  """Size of the persistent host worker pool behind host_random_legal_actions."""
  _require_lib()
  return int(lib.HostPolicyThreads())


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
