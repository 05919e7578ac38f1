"""oracle/agents_oracle.py -- TEST INFRASTRUCTURE ONLY.

Restatement of the reference's two hand-written policies over the per-player
observation dict of rl_env.HanabiEnv (the reference's or hanabi_b200's):
  RuleBasedAgent   agents/rule_based_agent.py:15-302
  SimpleAgent      agents/simple_agent.py:20-69
Both return move uids (RuleBasedAgent.act_train semantics: an action that is not
among the legal moves becomes the last legal move, :284-296).

PARITY PINNED: tests/test_agents.py replays tests/golden/agent_traces.npz, recorded
from the unmodified reference agents (tests/golden/make_agent_golden.py).
"""
COLORS = "RYGWB"


def _uid(obs, action):
  for idx, legal in enumerate(obs["legal_moves"]):
    if legal == action:
      return int(obs["legal_moves_as_int"][idx])
  return int(obs["legal_moves_as_int"][-1])


class RuleBasedAgent(object):
  def __init__(self, players):
    self.players = players
    num_cards = 5 if players < 4 else 4                      # :24-27
    self.table = [[False] * num_cards for _ in range(players)]  # rank_hinted_but_no_play

  def _pop(self, row, index):
    self.table[row].pop(index)
    self.table[row].append(False)

  def act_uid(self, obs):
    view = obs["pyhanabi"]
    # check_if_not_playable_hint (:44-73)
    for item in view.last_moves():
      move = item.move()
      kind = move.type().name
      if kind == "REVEAL_RANK" and move.target_offset() == 1 and 0 in item.card_info_revealed():
        target = (item.player() + 1) % (self.players - 1)
        for idx in item.card_info_revealed():
          self.table[target][idx] = True
        break
      elif kind in ("DISCARD", "PLAY"):
        self._pop(item.player(), move.card_index())
    # maybe_play_lowest_playable_card (:75-112)
    own = view.card_knowledge()[0]
    for index, k in enumerate(own):
      if k.color() is not None:
        self._pop(0, index)
        return _uid(obs, {"action_type": "PLAY", "card_index": index})
    for index, k in enumerate(own):
      if k.rank() is not None and not self.table[0][index]:
        self._pop(0, index)
        return _uid(obs, {"action_type": "PLAY", "card_index": index})
    # maybe_give_helpful_hint (:114-228)
    hint = None
    if obs["information_tokens"] != 0:
      fireworks = obs["fireworks"]
      playable = lambda c: c["rank"] is not None and c["rank"] == fireworks[c["color"]]
      best = 0
      for offset in range(1, obs["num_players"]):
        hand = obs["observed_hands"][offset]
        know = view.card_knowledge()[offset]
        colors, ranks = [], []
        for card in hand:
          if playable(card):
            if card["color"] not in colors:
              colors.append(card["color"])
            if card["rank"] not in ranks:
              ranks.append(card["rank"])
        for color in colors:
          content, bad = 0, False
          for card, k in zip(hand, know):
            if card["color"] != color:
              continue
            if playable(card) and k.color() is None:
              content += 1
            elif not playable(card):
              bad = True
              break
          if not bad and content > best:
            best = content
            hint = {"action_type": "REVEAL_COLOR", "color": color, "target_offset": offset}
        for rank in ranks:
          content, bad = 0, False
          for index, (card, k) in enumerate(zip(hand, know)):
            if card["rank"] != rank:
              continue
            if playable(card) and (k.rank() is None or self.table[offset][index]):
              content += 1
            elif not playable(card):
              bad = True
              break
          if not bad and content > best:
            best = content
            hint = {"action_type": "REVEAL_RANK", "rank": rank, "target_offset": offset}
    if hint is not None:
      return _uid(obs, hint)
    # discard the oldest card if allowed, else hint the right neighbour (:246-282)
    if any(m["action_type"] == "DISCARD" for m in obs["legal_moves"]):
      self._pop(0, 0)
      return _uid(obs, {"action_type": "DISCARD", "card_index": 0})
    return _uid(obs, {"action_type": "REVEAL_RANK", "rank": obs["observed_hands"][1][0]["rank"],
                      "target_offset": 1})


class SimpleAgent(object):
  def __init__(self, config):
    self.max_information_tokens = config["max_information_tokens"]

  def act_uid(self, obs):
    for index, hint in enumerate(obs["card_knowledge"][0]):
      if hint["color"] is not None or hint["rank"] is not None:
        return _uid(obs, {"action_type": "PLAY", "card_index": index})
    fireworks = obs["fireworks"]
    if obs["information_tokens"] > 0:
      for offset in range(1, obs["num_players"]):
        for card, hint in zip(obs["observed_hands"][offset], obs["card_knowledge"][offset]):
          if card["rank"] == fireworks[card["color"]] and hint["color"] is None:
            return _uid(obs, {"action_type": "REVEAL_RANK", "rank": card["rank"],
                              "target_offset": offset})
    if obs["information_tokens"] < self.max_information_tokens:
      return _uid(obs, {"action_type": "DISCARD", "card_index": 0})
    return _uid(obs, {"action_type": "PLAY", "card_index": 0})
