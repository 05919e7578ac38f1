"""CPU tests of the per-object C-ABI kept from the reference (include/pyhanabi.h,
first block; hanabi_b200/csrc/hb_single.cc) and of the Python layers above it.

  * every symbol the public header declares is exported by libpyhanabi.so;
  * the library's answers, call by call, equal those of the reference's own
    libpyhanabi.so (oracle/_ref, built from the unmodified sources) on scripted
    games over several rule sets -- strings, getters, histories, encodings;
  * hanabi_b200.rl_env.HanabiEnv reproduces the committed traces recorded from
    the reference's rl_env.HanabiEnv (tests/golden/rl_env_traces.npz).
"""
import os
import re

import cffi
import numpy as np
import pytest

from oracle import oracle as O
from tests.common import CONFIGS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pyhanabi.h")
MINE = os.path.join(ROOT, "hanabi_b200", "libpyhanabi.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_cdef():
  from hanabi_b200 import pyhanabi
  return pyhanabi.read_cdef(HEADER)


def declared_symbols():
  from hanabi_b200 import pyhanabi
  return pyhanabi.declared_functions(HEADER)


@pytest.fixture(scope="module")
def ffi_and_libs(built_lib):
  ffi = cffi.FFI()
  ffi.cdef(header_cdef())
  mine = ffi.dlopen(built_lib)
  ref = ffi.dlopen(O.REF_PYHANABI_SO) if os.path.exists(O.REF_PYHANABI_SO) else None
  return ffi, mine, ref


def test_library_exports_every_declared_symbol(ffi_and_libs):
  ffi, mine, _ = ffi_and_libs
  names = declared_symbols()
  assert len(names) >= 96 + 20
  for name in names:
    assert getattr(mine, name) is not None, name
  # the reference's 96 functions are all there (counted from its own header when present)
  ref_header = "/root/reference/hanabi_learning_environment/pyhanabi.h"
  if os.path.exists(ref_header):
    from hanabi_b200 import pyhanabi
    ref_names = set(pyhanabi.declared_functions(ref_header))
    assert len(ref_names) == 96
    assert ref_names <= set(names)


class Walker(object):
  """Plays one scripted game through a libpyhanabi and records every answer."""

  def __init__(self, ffi, lib, params):
    self.ffi, self.lib = ffi, lib
    self.keep = []
    flat = []
    for k, v in params.items():
      flat += [ffi.new("char[]", str(k).encode()), ffi.new("char[]", str(v).encode())]
    self.keep.append(flat)
    self.game = ffi.new("pyhanabi_game_t*")
    lib.NewGame(self.game, len(flat), ffi.new("char*[]", flat))
    self.enc = ffi.new("pyhanabi_observation_encoder_t*")
    lib.NewObservationEncoder(self.enc, self.game, 0)

  def s(self, c_str):
    text = self.ffi.string(c_str).decode()
    self.lib.DeleteString(c_str)
    return text

  def move_record(self, mv):
    lib = self.lib
    return (self.s(lib.MoveToString(mv)), lib.MoveType(mv), lib.CardIndex(mv), lib.TargetOffset(mv),
            lib.MoveColor(mv), lib.MoveRank(mv), lib.GetMoveUid(self.game, mv))

  def item_record(self, it):
    lib, ffi = self.lib, self.ffi
    mv = ffi.new("pyhanabi_move_t*")
    lib.HistoryItemMove(it, mv)
    rec = (self.s(lib.HistoryItemToString(it)), self.move_record(mv)[:6], lib.HistoryItemPlayer(it),
           lib.HistoryItemScored(it), lib.HistoryItemInformationToken(it), lib.HistoryItemColor(it),
           lib.HistoryItemRank(it), lib.HistoryItemRevealBitmask(it),
           lib.HistoryItemNewlyRevealedBitmask(it), lib.HistoryItemDealToPlayer(it))
    lib.DeleteMove(mv)
    return rec

  def game_record(self):
    lib, g = self.lib, self.game
    params = sorted(p for p in self.s(lib.GameParamString(g)).split("\n") if p)
    rec = [params, lib.NumPlayers(g), lib.NumColors(g), lib.NumRanks(g), lib.HandSize(g),
           lib.MaxInformationTokens(g), lib.MaxLifeTokens(g), lib.ObservationType(g), lib.MaxMoves(g),
           self.s(lib.ObservationShape(self.enc))]
    rec.append([lib.NumCards(g, c, r) for c in range(-1, 6) for r in range(-1, 6)])
    for uid in range(lib.MaxMoves(g)):
      mv = self.ffi.new("pyhanabi_move_t*")
      lib.GetMoveByUid(g, uid, mv)
      rec.append(self.move_record(mv))
      lib.DeleteMove(mv)
    return rec

  def state_record(self, st):
    lib, ffi = self.lib, self.ffi
    P, C, R = lib.NumPlayers(self.game), lib.NumColors(self.game), lib.NumRanks(self.game)
    rec = [self.s(lib.StateToString(st)), lib.StateCurPlayer(st), lib.StateDeckSize(st),
           lib.StateInformationTokens(st), lib.StateLifeTokens(st), lib.StateScore(st),
           lib.StateEndOfGameStatus(st), lib.StateNumPlayers(st), lib.StateDiscardPileSize(st),
           [lib.StateFireworks(st, c) for c in range(C)],
           [lib.CardPlayableOnFireworks(st, c, r) for c in range(-1, C + 1) for r in range(R)]]
    card = ffi.new("pyhanabi_card_t*")
    for i in range(lib.StateDiscardPileSize(st)):
      lib.StateGetDiscard(st, i, card)
      rec.append((card.color, card.rank, lib.CardValid(card)))
    for p in range(P):
      for i in range(lib.StateGetHandSize(st, p)):
        lib.StateGetHandCard(st, p, i, card)
        rec.append((p, i, card.color, card.rank))
    return rec

  def observation_record(self, st, player):
    lib, ffi = self.lib, self.ffi
    P, C, R = lib.NumPlayers(self.game), lib.NumColors(self.game), lib.NumRanks(self.game)
    ob = ffi.new("pyhanabi_observation_t*")
    lib.NewObservation(st, player, ob)
    rec = [self.s(lib.ObsToString(ob)), lib.ObsCurPlayerOffset(ob), lib.ObsNumPlayers(ob),
           lib.ObsDeckSize(ob), lib.ObsInformationTokens(ob), lib.ObsLifeTokens(ob),
           [lib.ObsFireworks(ob, c) for c in range(C)],
           [lib.ObsCardPlayableOnFireworks(ob, c, r) for c in range(C) for r in range(R)],
           self.s(lib.EncodeObservation(self.enc, ob))]
    card = ffi.new("pyhanabi_card_t*")
    kn = ffi.new("pyhanabi_card_knowledge_t*")
    for p in range(P):
      for i in range(lib.ObsGetHandSize(ob, p)):
        lib.ObsGetHandCard(ob, p, i, card)
        lib.ObsGetHandCardKnowledge(ob, p, i, kn)
        rec.append((card.color, card.rank, lib.CardValid(card), self.s(lib.CardKnowledgeToString(kn)),
                    lib.ColorWasHinted(kn), lib.KnownColor(kn), lib.RankWasHinted(kn), lib.KnownRank(kn),
                    [lib.ColorIsPlausible(kn, c) for c in range(C)],
                    [lib.RankIsPlausible(kn, r) for r in range(R)]))
    for i in range(lib.ObsDiscardPileSize(ob)):
      lib.ObsGetDiscard(ob, i, card)
      rec.append((card.color, card.rank))
    for i in range(lib.ObsNumLastMoves(ob)):
      it = ffi.new("pyhanabi_history_item_t*")
      lib.ObsGetLastMove(ob, i, it)
      rec.append(self.item_record(it))
      lib.DeleteHistoryItem(it)
    for i in range(lib.ObsNumLegalMoves(ob)):
      mv = ffi.new("pyhanabi_move_t*")
      lib.ObsGetLegalMove(ob, i, mv)
      rec.append(self.move_record(mv))
      lib.DeleteMove(mv)
    lib.DeleteObservation(ob)
    return rec

  def play(self, steps, pick):
    lib, ffi = self.lib, self.ffi
    out = [self.game_record()]
    st = ffi.new("pyhanabi_state_t*")
    lib.NewState(self.game, st)
    t = 0
    while t < steps:
      while lib.StateCurPlayer(st) == -1:
        lib.StateDealRandomCard(st)
      out.append(self.state_record(st))
      for p in range(lib.NumPlayers(self.game)):
        out.append(self.observation_record(st, p))
      if t % 7 == 0:
        cp = ffi.new("pyhanabi_state_t*")
        lib.CopyState(st, cp)
        out.append(self.state_record(cp))
        lib.DeleteState(cp)
      if lib.StateEndOfGameStatus(st) != 0:
        hist = []
        for i in range(lib.StateLenMoveHistory(st)):
          it = ffi.new("pyhanabi_history_item_t*")
          lib.StateGetMoveHistory(st, i, it)
          hist.append(self.item_record(it))
          lib.DeleteHistoryItem(it)
        out.append(hist)
        lib.DeleteState(st)
        lib.NewState(self.game, st)
        continue
      moves = lib.StateLegalMoves(st)
      n = lib.NumMoves(moves)
      mv = ffi.new("pyhanabi_move_t*")
      lib.GetMove(moves, pick(t, n), mv)
      lib.DeleteMoveList(moves)
      out.append((self.move_record(mv), lib.MoveIsLegal(st, mv)))
      lib.StateApplyMove(st, mv)
      lib.DeleteMove(mv)
      t += 1
    lib.DeleteState(st)
    return out


@pytest.mark.skipif(not os.path.exists(O.REF_PYHANABI_SO), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(CONFIGS))
@pytest.mark.parametrize("policy", ["cycle", "hinty"])
def test_every_abi_answer_matches_the_reference_library(ffi_and_libs, name, policy):
  ffi, mine, ref = ffi_and_libs
  params = dict(CONFIGS[name])
  params["seed"] = 321
  if policy == "cycle":
    pick = lambda t, n: (t * 7 + 3) % n
  else:  # prefer the last legal moves (hints) so that games last and knowledge builds up
    pick = lambda t, n: n - 1 - (t % 3 == 0) * ((t // 3) % n) % n
  a = Walker(ffi, ref, params).play(90, pick)
  b = Walker(ffi, mine, params).play(90, pick)
  assert len(a) == len(b)
  for i, (x, y) in enumerate(zip(a, b)):
    assert x == y, "record %d differs:\n%r\n%r" % (i, x, y)


def test_move_constructors_and_strings(ffi_and_libs):
  ffi, mine, ref = ffi_and_libs
  for lib in [l for l in (mine, ref) if l is not None]:
    mv = ffi.new("pyhanabi_move_t*")
    assert lib.GetDiscardMove(2, mv)
    assert ffi.string(lib.MoveToString(mv)).decode() == "(Discard 2)"
    lib.DeleteMove(mv)
    assert lib.GetPlayMove(0, mv)
    assert ffi.string(lib.MoveToString(mv)).decode() == "(Play 0)"
    lib.DeleteMove(mv)
    assert lib.GetRevealColorMove(1, 3, mv)
    assert ffi.string(lib.MoveToString(mv)).decode() == "(Reveal player +1 color W)"
    lib.DeleteMove(mv)
    assert lib.GetRevealRankMove(2, 4, mv)
    assert ffi.string(lib.MoveToString(mv)).decode() == "(Reveal player +2 rank 5)"
    lib.DeleteMove(mv)


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_hanabi_env_reproduces_reference_rl_env_traces(built_lib, name):
  """hanabi_b200.rl_env.HanabiEnv vs traces recorded from the reference's
  rl_env.HanabiEnv (tests/golden/make_golden.py)."""
  from hanabi_b200 import rl_env
  traces = np.load(os.path.join(GOLDEN, "rl_env_traces.npz"))
  for seed in (1234, 7):
    key = "%s/%d/" % (name, seed)
    cfg = dict(CONFIGS[name])
    cfg["seed"] = seed
    env = rl_env.HanabiEnv(cfg)
    D, A = [int(x) for x in traces[key + "shape"]]
    assert env.vectorized_observation_shape() == [D] and env.num_moves() == A
    obs = env.reset()
    action, cur, vec, mask = (traces[key + k] for k in ("action", "cur", "obs", "mask"))
    reward, done = traces[key + "reward"], traces[key + "done"]
    for t in range(len(action)):
      if t > 0:
        obs, r, d, _ = env.step(int(action[t]))
        assert r == reward[t] and bool(d) == bool(done[t])
        if d:
          obs = env.reset()
      assert obs["current_player"] == cur[t]
      po = obs["player_observations"][obs["current_player"]]
      want = np.unpackbits(vec[t], bitorder="little")[:D].tolist()
      assert po["vectorized"] == want, (name, seed, t)
      assert po["legal_moves_as_int"] == np.flatnonzero(mask[t]).tolist()
      assert sorted(po.keys()) == sorted([
          "current_player", "current_player_offset", "life_tokens", "information_tokens",
          "num_players", "deck_size", "fireworks", "legal_moves", "legal_moves_as_int",
          "observed_hands", "discard_pile", "card_knowledge", "vectorized", "pyhanabi"])
      for other in obs["player_observations"]:
        if other is not po:
          assert other["legal_moves"] == [] and other["legal_moves_as_int"] == []


def test_hanabi_env_api_quirks(built_lib):
  from hanabi_b200 import rl_env
  env = rl_env.HanabiEnv({"players": 2, "seed": 3})
  obs = env.reset()
  with pytest.raises(ValueError):
    env.step(np.int64(5))  # numpy ints are rejected like in the reference (rl_env.py:349-351)
  po = obs["player_observations"][0]
  mv = po["legal_moves"][0]
  obs2, r, d, info = env.step(mv)  # dict actions
  assert isinstance(r, int) and d is False and info == {}
  with pytest.raises(AssertionError):
    env.step({"action_type": "DISCARD", "card_index": 0})  # info tokens are at max: illegal
  with pytest.raises(ValueError):
    rl_env.make("No-Such-Game")
  e = rl_env.make("Hanabi-Small", num_players=3)
  assert e.vectorized_observation_shape() == [191 + 35] or e.num_moves() == 18
  # the fork's extra key: the live observation object used by the rule-based agent
  assert po["pyhanabi"].card_playable_on_fireworks(0, 0)
