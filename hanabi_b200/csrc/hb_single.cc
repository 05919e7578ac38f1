// hb_single.cc -- the reference's per-object C-ABI (include/pyhanabi.h, first
// block; reference hanabi_learning_environment/pyhanabi.h:26-190), kept so that
// rl_env.HanabiEnv, the rule-based agents, the GUI translation layer and the
// example scripts keep working unchanged against this library.
//
// This is the host-side mirror of the reference interface, NOT the accelerated
// path and not a fallback for it: it serves the one-game-at-a-time object API
// (cards, knowledge, moves, history items, states, observations) that is
// inherently host-resident.  The batched hot path (Batch*) lives in
// hb_batch.cu and never calls into this file.
//
// Clean-room engine: one small value-type model of the game written against the
// behaviour documented in SURVEY.md appendices A-C, checked function by function
// against the reference's own libpyhanabi.so (tests/test_single_game_abi.py).
// Deals use the same standard-library facilities the reference uses
// (std::mt19937 + std::discrete_distribution over doubles count/deck_size,
// hanabi_lib/hanabi_game.cc:106-112, hanabi_lib/hanabi_state.cc:277-280), so the
// RNG streams agree by construction.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pyhanabi.h"
#include "hb_spec.h"

namespace hbs {

#define HBS_REQUIRE(expr)                                                                  \
  ((expr) ? (void)0                                                                        \
          : (std::fprintf(stderr, "Input requirements failed at %s:%d in %s: %s\n", __FILE__, \
                          __LINE__, __func__, #expr),                                      \
             std::abort()))

enum MoveKind { kInvalid = 0, kPlay = 1, kDiscard = 2, kRevealColor = 3, kRevealRank = 4, kDeal = 5 };
enum EndKind { kNotFinished = 0, kOutOfLifeTokens = 1, kOutOfCards = 2, kCompletedFireworks = 3 };
enum ObsKind { kMinimal = 0, kCardKnowledge = 1, kSeer = 2 };
constexpr int kChance = -1;

char ColorChar(int c) { return (c >= 0 && c <= 5) ? "RYGWB"[c] : 'X'; }  // util.cc:22-28 quirk: <= max
char RankChar(int r) { return (r >= 0 && r <= 5) ? "12345"[r] : 'X'; }

struct Card {
  int color = -1, rank = -1;
  bool valid() const { return color >= 0 && rank >= 0; }
  std::string str() const {
    return valid() ? std::string() + ColorChar(color) + RankChar(rank) : std::string("XX");
  }
};

// What the table knows about one card from hints (hanabi_hand.h:28-92).
struct Knowledge {
  int num_colors = 0, num_ranks = 0;
  int color = -1, rank = -1;       // directly hinted value or -1
  uint32_t color_ok = 0, rank_ok = 0;  // bit v: value v not excluded yet
  Knowledge() {}
  Knowledge(int C, int R) : num_colors(C), num_ranks(R), color_ok((1u << C) - 1), rank_ok((1u << R) - 1) {}
  void is_color(int c) { color = c; color_ok = 1u << c; }
  void not_color(int c) { color_ok &= ~(1u << c); }
  void is_rank(int r) { rank = r; rank_ok = 1u << r; }
  void not_rank(int r) { rank_ok &= ~(1u << r); }
  std::string str() const {  // hanabi_hand.cc:47-62
    std::string s;
    s += color >= 0 ? ColorChar(color) : 'X';
    s += rank >= 0 ? RankChar(rank) : 'X';
    s += '|';
    for (int c = 0; c < num_colors; ++c)
      if (color_ok >> c & 1) s += ColorChar(c);
    for (int r = 0; r < num_ranks; ++r)
      if (rank_ok >> r & 1) s += RankChar(r);
    return s;
  }
};

struct Move {
  int kind = kInvalid;
  int8_t card_index = -1, target_offset = -1, color = -1, rank = -1;
  Move() {}
  Move(int k, int ci, int to, int c, int r)
      : kind(k), card_index((int8_t)ci), target_offset((int8_t)to), color((int8_t)c), rank((int8_t)r) {}
  std::string str() const {  // hanabi_move.cc:45-69
    switch (kind) {
      case kPlay: return "(Play " + std::to_string(card_index) + ")";
      case kDiscard: return "(Discard " + std::to_string(card_index) + ")";
      case kRevealColor:
        return "(Reveal player +" + std::to_string(target_offset) + " color " + ColorChar(color) + ")";
      case kRevealRank:
        return "(Reveal player +" + std::to_string(target_offset) + " rank " + RankChar(rank) + ")";
      case kDeal:
        if (color >= 0) return std::string("(Deal ") + ColorChar(color) + RankChar(rank) + ")";
        return "(Deal XX)";
      default: return "(INVALID)";
    }
  }
};

struct HistoryItem {  // hanabi_history_item.h:33-56
  Move move;
  int8_t player = -1;
  bool scored = false, information_token = false;
  int8_t color = -1, rank = -1;
  uint8_t reveal_bitmask = 0, newly_revealed_bitmask = 0;
  int8_t deal_to_player = -1;
  std::string str() const {  // hanabi_history_item.cc:23-55
    std::string s = "<" + move.str();
    if (player >= 0) s += " by player " + std::to_string(player);
    if (scored) s += " scored";
    if (information_token) s += " info_token";
    if (color >= 0) { s += " "; s += ColorChar(color); s += RankChar(rank); }
    if (reveal_bitmask) {
      s += " reveal ";
      bool first = true;
      for (int i = 0; i < 8; ++i)
        if (reveal_bitmask & (1 << i)) {
          if (!first) s += ",";
          first = false;
          s += std::to_string(i);
        }
    }
    return s + ">";
  }
};

struct Hand {
  std::vector<Card> cards;
  std::vector<Knowledge> knowledge;
  std::string str() const {
    std::string s;
    for (size_t i = 0; i < cards.size(); ++i) s += cards[i].str() + " || " + knowledge[i].str() + '\n';
    return s;
  }
};

struct Game {
  std::unordered_map<std::string, std::string> params;
  int players, colors, ranks, hand_size, max_info, max_life, seed, obs_type;
  bool random_start;
  int cards_per_color = 0;
  mutable std::mt19937 rng;  // one stream per game, shared by all its states (hanabi_game.h:114)
  std::vector<Move> moves;

  static int IntParam(const std::unordered_map<std::string, std::string>& p, const char* k, int d) {
    auto it = p.find(k);
    return it == p.end() ? d : std::stoi(it->second);
  }
  static bool BoolParam(const std::unordered_map<std::string, std::string>& p, const char* k, bool d) {
    auto it = p.find(k);
    if (it == p.end()) return d;
    return it->second == "1" || it->second == "true" || it->second == "True";
  }
  explicit Game(const std::unordered_map<std::string, std::string>& p) : params(p) {
    players = IntParam(p, "players", 2);
    HBS_REQUIRE(players >= 2 && players <= 5);
    colors = IntParam(p, "colors", 5);
    HBS_REQUIRE(colors > 0 && colors <= 5);
    ranks = IntParam(p, "ranks", 5);
    HBS_REQUIRE(ranks > 0 && ranks <= 5);
    hand_size = IntParam(p, "hand_size", players < 4 ? 5 : 4);
    max_info = IntParam(p, "max_information_tokens", 8);
    max_life = IntParam(p, "max_life_tokens", 3);
    seed = IntParam(p, "seed", -1);
    random_start = BoolParam(p, "random_start_player", false);
    obs_type = IntParam(p, "observation_type", kCardKnowledge);
    while (seed == -1) seed = (int)std::random_device()();
    rng.seed(seed);
    for (int r = 0; r < ranks; ++r) cards_per_color += instances(0, r);
    HBS_REQUIRE(hand_size * players <= cards_per_color * colors);
    for (int uid = 0; uid < max_moves(); ++uid) moves.push_back(make_move(uid));
  }
  int instances(int c, int r) const {
    if (c < 0 || c >= colors || r < 0 || r >= ranks) return 0;
    return r == 0 ? 3 : (r == ranks - 1 ? 1 : 2);
  }
  int deck_size() const { return cards_per_color * colors; }
  int max_moves() const { return 2 * hand_size + (players - 1) * (colors + ranks); }
  Move make_move(int uid) const {  // uid layout hanabi_game.cc:154-183
    if (uid < 0 || uid >= max_moves()) return Move();
    if (uid < hand_size) return Move(kDiscard, uid, -1, -1, -1);
    uid -= hand_size;
    if (uid < hand_size) return Move(kPlay, uid, -1, -1, -1);
    uid -= hand_size;
    if (uid < (players - 1) * colors) return Move(kRevealColor, -1, 1 + uid / colors, uid % colors, -1);
    uid -= (players - 1) * colors;
    return Move(kRevealRank, -1, 1 + uid / ranks, -1, uid % ranks);
  }
  int move_uid(const Move& m) const {  // hanabi_game.cc:79-95
    switch (m.kind) {
      case kDiscard: return m.card_index;
      case kPlay: return hand_size + m.card_index;
      case kRevealColor: return 2 * hand_size + (m.target_offset - 1) * colors + m.color;
      case kRevealRank:
        return 2 * hand_size + (players - 1) * colors + (m.target_offset - 1) * ranks + m.rank;
      default: return -1;
    }
  }
  int sample_start_player() const {  // hanabi_game.cc:138-145
    if (!random_start) return 0;
    std::uniform_int_distribution<std::mt19937::result_type> dist(0, players - 1);
    return (int)dist(rng);
  }
  std::string param_string() const {  // GameParamString, pyhanabi.cc:528-537 (map order unspecified)
    std::string s;
    s += "players=" + std::to_string(players) + "\n";
    s += "colors=" + std::to_string(colors) + "\n";
    s += "ranks=" + std::to_string(ranks) + "\n";
    s += "hand_size=" + std::to_string(hand_size) + "\n";
    s += "max_information_tokens=" + std::to_string(max_info) + "\n";
    s += "max_life_tokens=" + std::to_string(max_life) + "\n";
    s += "seed=" + std::to_string(seed) + "\n";
    s += std::string("random_start_player=") + (random_start ? "true" : "false") + "\n";
    s += "observation_type=" + std::to_string(obs_type) + "\n";
    return s;
  }
};

struct State {
  const Game* game;
  std::vector<int> deck;  // remaining instances per card type (colour-major)
  int deck_total = 0;
  std::vector<Card> discard_pile;
  std::vector<Hand> hands;
  std::vector<HistoryItem> history;
  int cur_player = kChance, next_player = 0, info, life, turns_to_play;
  std::vector<int> fireworks;

  explicit State(const Game* g)
      : game(g), deck(g->colors * g->ranks, 0), hands(g->players), info(g->max_info),
        life(g->max_life), turns_to_play(g->players), fireworks(g->colors, 0) {
    for (int c = 0; c < g->colors; ++c)
      for (int r = 0; r < g->ranks; ++r) {
        deck[c * g->ranks + r] = g->instances(c, r);
        deck_total += g->instances(c, r);
      }
    next_player = g->sample_start_player();
  }
  int player_to_deal() const {
    for (int p = 0; p < game->players; ++p)
      if ((int)hands[p].cards.size() < game->hand_size) return p;
    return -1;
  }
  void advance() {  // hanabi_state.cc:104-111
    if (deck_total != 0 && player_to_deal() >= 0) {
      cur_player = kChance;
    } else {
      cur_player = next_player;
      next_player = (cur_player + 1) % game->players;
    }
  }
  const Hand& hand_at(int offset) const { return hands[(cur_player + offset) % game->players]; }
  Hand& hand_at(int offset) { return hands[(cur_player + offset) % game->players]; }

  bool legal(const Move& m) const {  // hanabi_state.cc:166-219
    switch (m.kind) {
      case kDeal:
        return cur_player == kChance && m.color >= 0 && m.rank >= 0 && m.color < game->colors &&
               m.rank < game->ranks && deck[m.color * game->ranks + m.rank] != 0;
      case kDiscard:
        return info < game->max_info && m.card_index < (int)hands[cur_player].cards.size();
      case kPlay:
        return m.card_index < (int)hands[cur_player].cards.size();
      case kRevealColor:
      case kRevealRank: {
        if (info <= 0 || m.target_offset < 1 || m.target_offset >= game->players) return false;
        for (const Card& c : hand_at(m.target_offset).cards)
          if (m.kind == kRevealColor ? c.color == m.color : c.rank == m.rank) return true;
        return false;
      }
      default:
        return false;
    }
  }
  bool gain_info() {
    if (info < game->max_info) { ++info; return true; }
    return false;
  }
  void remove_card(int p, int idx, bool to_discard) {
    if (to_discard) discard_pile.push_back(hands[p].cards[idx]);
    hands[p].cards.erase(hands[p].cards.begin() + idx);
    hands[p].knowledge.erase(hands[p].knowledge.begin() + idx);
  }
  void apply(const Move& m) {  // hanabi_state.cc:221-275
    HBS_REQUIRE(legal(m));
    if (deck_total == 0) --turns_to_play;
    HistoryItem h;
    h.move = m;
    h.player = (int8_t)cur_player;
    switch (m.kind) {
      case kDeal: {
        h.deal_to_player = (int8_t)player_to_deal();
        Knowledge k(game->colors, game->ranks);
        if (game->obs_type == kSeer) { k.is_color(m.color); k.is_rank(m.rank); }
        Card c; c.color = m.color; c.rank = m.rank;
        --deck[m.color * game->ranks + m.rank];
        --deck_total;
        hands[h.deal_to_player].cards.push_back(c);
        hands[h.deal_to_player].knowledge.push_back(k);
        break;
      }
      case kDiscard: {
        h.information_token = gain_info();
        const Card c = hands[cur_player].cards[m.card_index];
        h.color = (int8_t)c.color; h.rank = (int8_t)c.rank;
        remove_card(cur_player, m.card_index, true);
        break;
      }
      case kPlay: {
        const Card c = hands[cur_player].cards[m.card_index];
        h.color = (int8_t)c.color; h.rank = (int8_t)c.rank;
        if (c.rank == fireworks[c.color]) {
          ++fireworks[c.color];
          h.scored = true;
          if (fireworks[c.color] == game->ranks) h.information_token = gain_info();
        } else {
          --life;
        }
        remove_card(cur_player, m.card_index, !h.scored);
        break;
      }
      case kRevealColor:
      case kRevealRank: {
        --info;
        Hand& t = hand_at(m.target_offset);
        for (size_t i = 0; i < t.cards.size(); ++i) {
          Knowledge& k = t.knowledge[i];
          if (m.kind == kRevealColor) {
            if (t.cards[i].color == m.color) {
              h.reveal_bitmask |= (uint8_t)(1u << i);
              if (k.color < 0) h.newly_revealed_bitmask |= (uint8_t)(1u << i);
              k.is_color(m.color);
            } else {
              k.not_color(m.color);
            }
          } else {
            if (t.cards[i].rank == m.rank) {
              h.reveal_bitmask |= (uint8_t)(1u << i);
              if (k.rank < 0) h.newly_revealed_bitmask |= (uint8_t)(1u << i);
              k.is_rank(m.rank);
            } else {
              k.not_rank(m.rank);
            }
          }
        }
        break;
      }
      default:
        std::abort();
    }
    history.push_back(h);
    advance();
  }
  void deal_random() {  // hanabi_state.cc:282-286,313-325; hanabi_game.cc:106-112
    std::vector<Move> outcomes;
    std::vector<double> probs;
    for (int c = 0; c < game->colors; ++c)
      for (int r = 0; r < game->ranks; ++r) {
        Move m(kDeal, -1, -1, c, r);
        if (legal(m)) {
          outcomes.push_back(m);
          probs.push_back(static_cast<double>(deck[c * game->ranks + r]) / static_cast<double>(deck_total));
        }
      }
    HBS_REQUIRE(!outcomes.empty());
    std::discrete_distribution<std::mt19937::result_type> dist(probs.begin(), probs.end());
    apply(outcomes[dist(game->rng)]);
  }
  std::vector<Move> legal_moves(int player) const {  // hanabi_state.cc:288-304
    std::vector<Move> out;
    HBS_REQUIRE(player >= 0 && player < game->players);
    if (player != cur_player) return out;
    for (const Move& m : game->moves)
      if (legal(m)) out.push_back(m);
    return out;
  }
  int score() const {
    if (life <= 0) return 0;
    int s = 0;
    for (int f : fireworks) s += f;
    return s;
  }
  int end_status() const {
    if (life < 1) return kOutOfLifeTokens;
    if (score() >= game->colors * game->ranks) return kCompletedFireworks;
    if (turns_to_play <= 0) return kOutOfCards;
    return kNotFinished;
  }
  bool playable(int color, int rank) const {
    return color >= 0 && color < game->colors && rank == fireworks[color];
  }
};

std::string BoardString(int life, int info, const std::vector<int>& fireworks,
                        const std::vector<Hand>& hands, int cur, int deck_size,
                        const std::vector<Card>& discards) {  // hanabi_state.cc:332-357
  std::string s;
  s += "Life tokens: " + std::to_string(life) + "\n";
  s += "Info tokens: " + std::to_string(info) + "\n";
  s += "Fireworks: ";
  for (size_t i = 0; i < fireworks.size(); ++i) {
    s += ColorChar((int)i);
    s += std::to_string(fireworks[i]) + " ";
  }
  s += "\nHands:\n";
  for (size_t i = 0; i < hands.size(); ++i) {
    if (i > 0) s += "-----\n";
    if ((int)i == cur) s += "Cur player\n";
    s += hands[i].str();
  }
  s += "Deck size: " + std::to_string(deck_size) + "\n";
  s += "Discards:";
  for (const Card& c : discards) s += " " + c.str();
  return s;
}

// One player's view (hanabi_observation.cc:52-93).
struct Observation {
  const Game* game;
  int cur_player_offset;
  std::vector<Hand> hands;  // observer first, then clockwise
  std::vector<Card> discard_pile;
  std::vector<int> fireworks;
  int deck_size, info, life;
  std::vector<HistoryItem> last_moves;  // newest first
  std::vector<Move> legal_moves;

  Observation(const State& s, int observer) : game(s.game) {
    const int P = game->players;
    HBS_REQUIRE(observer >= 0 && observer < P);
    cur_player_offset = s.cur_player >= 0 ? (s.cur_player - observer + P) % P : s.cur_player;
    discard_pile = s.discard_pile;
    fireworks = s.fireworks;
    deck_size = s.deck_total;
    info = s.info;
    life = s.life;
    legal_moves = s.legal_moves(observer);
    const bool hide_knowledge = game->obs_type == kMinimal;
    const bool show_own = game->obs_type == kSeer;
    for (int q = 0; q < P; ++q) {
      Hand h = s.hands[(observer + q) % P];
      if (q == 0 && !show_own)
        for (Card& c : h.cards) c = Card();
      if (hide_knowledge)
        for (Knowledge& k : h.knowledge) k = Knowledge(game->colors, game->ranks);
      hands.push_back(h);
    }
    // history back to the observer's own previous action, newest first; the
    // initial chance moves are never included
    size_t first = 0;
    while (first < s.history.size() && s.history[first].player == kChance) ++first;
    for (size_t i = s.history.size(); i > first; --i) {
      HistoryItem it = s.history[i - 1];
      const bool own_move = it.player == observer;
      if (it.move.kind == kDeal) {
        it.deal_to_player = (int8_t)((it.deal_to_player - observer + P) % P);
        if (it.deal_to_player == 0 && !show_own) it.move = Move(kDeal, -1, -1, -1, -1);
      } else {
        it.player = (int8_t)((it.player - observer + P) % P);
      }
      last_moves.push_back(it);
      if (own_move) break;
    }
  }
  bool playable(int color, int rank) const {
    return color >= 0 && color < game->colors && rank == fireworks[color];
  }
};

// The canonical encoding (canonical_encoders.cc:62-451) over an Observation.
struct Encoder {
  const Game* game;
  int type;
  int shape() const {
    const Game& g = *game;
    const int bpc = g.colors * g.ranks;
    int n = (g.players - 1) * g.hand_size * bpc + g.players;
    n += g.deck_size() - g.players * g.hand_size + bpc + g.max_info + g.max_life;
    n += g.deck_size();
    n += g.players + 4 + g.players + g.colors + g.ranks + 2 * g.hand_size + bpc + 2;
    if (g.obs_type != kMinimal) n += g.players * g.hand_size * (bpc + g.colors + g.ranks);
    return n;
  }
  std::vector<int> encode(const Observation& o) const {
    const Game& g = *game;
    const int P = g.players, C = g.colors, R = g.ranks, H = g.hand_size, bpc = C * R;
    std::vector<int> e(shape(), 0);
    int off = 0;
    for (int q = 1; q < P; ++q) {
      for (size_t i = 0; i < o.hands[q].cards.size(); ++i)
        e[off + (int)i * bpc + o.hands[q].cards[i].color * R + o.hands[q].cards[i].rank] = 1;
      off += H * bpc;
    }
    for (int q = 0; q < P; ++q)
      if ((int)o.hands[q].cards.size() < H) e[off + q] = 1;
    off += P;
    for (int i = 0; i < o.deck_size; ++i) e[off + i] = 1;
    off += g.deck_size() - P * H;
    for (int c = 0; c < C; ++c) {
      if (o.fireworks[c] > 0) e[off + o.fireworks[c] - 1] = 1;
      off += R;
    }
    for (int i = 0; i < o.info; ++i) e[off + i] = 1;
    off += g.max_info;
    for (int i = 0; i < o.life; ++i) e[off + i] = 1;
    off += g.max_life;
    std::vector<int> counts(bpc, 0);
    for (const Card& c : o.discard_pile) ++counts[c.color * R + c.rank];
    for (int c = 0; c < C; ++c)
      for (int r = 0; r < R; ++r) {
        for (int i = 0; i < counts[c * R + r]; ++i) e[off + i] = 1;
        off += g.instances(c, r);
      }
    const HistoryItem* last = nullptr;
    for (const HistoryItem& it : o.last_moves)
      if (it.move.kind != kDeal) { last = &it; break; }
    if (last) {
      int p = off;
      const int k = last->move.kind;
      e[p + last->player] = 1;
      p += P;
      e[p + (k == kPlay ? 0 : k == kDiscard ? 1 : k == kRevealColor ? 2 : 3)] = 1;
      p += 4;
      const bool hint = k == kRevealColor || k == kRevealRank, card = k == kPlay || k == kDiscard;
      if (hint) e[p + (last->player + last->move.target_offset) % P] = 1;
      p += P;
      if (k == kRevealColor) e[p + last->move.color] = 1;
      p += C;
      if (k == kRevealRank) e[p + last->move.rank] = 1;
      p += R;
      if (hint)
        for (int i = 0; i < H; ++i)
          if (last->reveal_bitmask & (1 << i)) e[p + i] = 1;
      p += H;
      if (card) e[p + last->move.card_index] = 1;
      p += H;
      if (card) e[p + last->color * R + last->rank] = 1;
      p += bpc;
      if (k == kPlay) {
        if (last->scored) e[p] = 1;
        if (last->information_token) e[p + 1] = 1;
      }
    }
    off += P + 4 + P + C + R + 2 * H + bpc + 2;
    if (g.obs_type != kMinimal) {
      for (int q = 0; q < P; ++q) {
        for (size_t i = 0; i < o.hands[q].knowledge.size(); ++i) {
          const Knowledge& kn = o.hands[q].knowledge[i];
          const int p = off + (int)i * (bpc + C + R);
          for (int c = 0; c < C; ++c)
            if (kn.color_ok >> c & 1)
              for (int r = 0; r < R; ++r)
                if (kn.rank_ok >> r & 1) e[p + c * R + r] = 1;
          if (kn.color >= 0) e[p + bpc + kn.color] = 1;
          if (kn.rank >= 0) e[p + bpc + C + kn.rank] = 1;
        }
        off += H * (bpc + C + R);
      }
    }
    HBS_REQUIRE(off == (int)e.size());
    return e;
  }
};

char* Dup(const std::string& s) { return strdup(s.c_str()); }

}  // namespace hbs

using namespace hbs;  // NOLINT

#define AS(T, p) (reinterpret_cast<T*>(p))
#define CAS(T, p) (reinterpret_cast<const T*>(p))

extern "C" {

void DeleteString(char* str) { free(str); }

int CardValid(pyhanabi_card_t* card) {
  HBS_REQUIRE(card != nullptr);
  return card->color >= 0;
}

char* CardKnowledgeToString(pyhanabi_card_knowledge_t* k) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return Dup(CAS(Knowledge, k->knowledge)->str());
}
int ColorWasHinted(pyhanabi_card_knowledge_t* k) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return CAS(Knowledge, k->knowledge)->color >= 0;
}
int KnownColor(pyhanabi_card_knowledge_t* k) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return CAS(Knowledge, k->knowledge)->color;
}
int ColorIsPlausible(pyhanabi_card_knowledge_t* k, int color) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return (CAS(Knowledge, k->knowledge)->color_ok >> color) & 1;
}
int RankWasHinted(pyhanabi_card_knowledge_t* k) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return CAS(Knowledge, k->knowledge)->rank >= 0;
}
int KnownRank(pyhanabi_card_knowledge_t* k) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return CAS(Knowledge, k->knowledge)->rank;
}
int RankIsPlausible(pyhanabi_card_knowledge_t* k, int rank) {
  HBS_REQUIRE(k != nullptr && k->knowledge != nullptr);
  return (CAS(Knowledge, k->knowledge)->rank_ok >> rank) & 1;
}

void DeleteMoveList(void* movelist) { delete AS(std::vector<Move>, movelist); }
int NumMoves(void* movelist) {
  HBS_REQUIRE(movelist != nullptr);
  return (int)AS(std::vector<Move>, movelist)->size();
}
void GetMove(void* movelist, int index, pyhanabi_move_t* move) {
  HBS_REQUIRE(movelist != nullptr && move != nullptr);
  move->move = new Move(AS(std::vector<Move>, movelist)->at(index));
}
void DeleteMove(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr);
  delete AS(Move, move->move);
  move->move = nullptr;
}
char* MoveToString(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return Dup(CAS(Move, move->move)->str());
}
int MoveType(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return CAS(Move, move->move)->kind;
}
int CardIndex(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return CAS(Move, move->move)->card_index;
}
int TargetOffset(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return CAS(Move, move->move)->target_offset;
}
int MoveColor(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return CAS(Move, move->move)->color;
}
int MoveRank(pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr && move->move != nullptr);
  return CAS(Move, move->move)->rank;
}
bool GetDiscardMove(int card_index, pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr);
  move->move = new Move(kDiscard, card_index, -1, -1, -1);
  return move->move != nullptr;
}
bool GetPlayMove(int card_index, pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr);
  move->move = new Move(kPlay, card_index, -1, -1, -1);
  return move->move != nullptr;
}
bool GetRevealColorMove(int target_offset, int color, pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr);
  move->move = new Move(kRevealColor, -1, target_offset, color, -1);
  return move->move != nullptr;
}
bool GetRevealRankMove(int target_offset, int rank, pyhanabi_move_t* move) {
  HBS_REQUIRE(move != nullptr);
  move->move = new Move(kRevealRank, -1, target_offset, -1, rank);
  return move->move != nullptr;
}

void DeleteHistoryItem(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr);
  delete AS(HistoryItem, item->item);
  item->item = nullptr;
}
char* HistoryItemToString(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return Dup(CAS(HistoryItem, item->item)->str());
}
void HistoryItemMove(pyhanabi_history_item_t* item, pyhanabi_move_t* move) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr && move != nullptr);
  move->move = new Move(CAS(HistoryItem, item->item)->move);
}
int HistoryItemPlayer(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->player;
}
int HistoryItemScored(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->scored;
}
int HistoryItemInformationToken(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->information_token;
}
int HistoryItemColor(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->color;
}
int HistoryItemRank(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->rank;
}
int HistoryItemRevealBitmask(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->reveal_bitmask;
}
int HistoryItemNewlyRevealedBitmask(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->newly_revealed_bitmask;
}
int HistoryItemDealToPlayer(pyhanabi_history_item_t* item) {
  HBS_REQUIRE(item != nullptr && item->item != nullptr);
  return CAS(HistoryItem, item->item)->deal_to_player;
}

void NewState(pyhanabi_game_t* game, pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && game != nullptr && game->game != nullptr);
  state->state = new State(CAS(Game, game->game));
}
void CopyState(const pyhanabi_state_t* src, pyhanabi_state_t* dest) {
  HBS_REQUIRE(src != nullptr && src->state != nullptr && dest != nullptr);
  dest->state = new State(*CAS(State, src->state));
}
void DeleteState(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr);
  delete AS(State, state->state);
  state->state = nullptr;
}
// Rebuilds `state` (an existing state of `game`) from one game of a batch as
// BatchExportGame returns it, so that the reference's per-object API -- observations,
// encoders, the rule-based agents, the GUI layer -- can look at a batched game.
// With BatchEnableHistory the discard pile comes back in the order the cards were discarded
// and the move history holds what any observer can see: the newest P player moves with the
// deals that followed them (without it: the pile sorted by card type, the last player move
// only).  newly_revealed_bitmask is not recorded and returns as the reveal bitmask.
// Everything the canonical encoder and LegalMoves look at is exact either way.
void StateFromBatchGame(pyhanabi_game_t* game, const int32_t* rec, pyhanabi_state_t* state) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr && rec != nullptr && state != nullptr);
  const Game* g = CAS(Game, game->game);
  State* st = new State(g);
  const int C = g->colors, R = g->ranks, P = g->players;
  st->cur_player = rec[0];
  st->next_player = (rec[0] + 1) % P;
  st->info = rec[1];
  st->life = rec[2];
  st->deck_total = rec[3];
  for (int c = 0; c < C; ++c) st->fireworks[c] = rec[6 + c];
  for (int p = 0; p < P; ++p) {
    Hand& h = st->hands[p];
    h.cards.clear();
    h.knowledge.clear();
    for (int i = 0; i < rec[11 + p]; ++i) {
      const int id = rec[16 + p * 8 + i];
      Card card;
      card.color = id / R;
      card.rank = id % R;
      Knowledge k(C, R);
      k.color_ok = (uint32_t)rec[56 + p * 8 + i];
      k.rank_ok = (uint32_t)rec[96 + p * 8 + i];
      k.color = rec[136 + p * 8 + i];
      k.rank = rec[176 + p * 8 + i];
      h.cards.push_back(card);
      h.knowledge.push_back(k);
    }
  }
  st->discard_pile.clear();
  int total_discards = 0;
  for (int k = 0; k < C * R; ++k) {
    st->deck[k] = rec[216 + k];
    total_discards += rec[241 + k];
  }
  // the pile in the order the cards were discarded (rl_env.py:416-418), when the batch logged
  // it (BatchEnableHistory) and the log agrees with the per-type counts; else sorted by type
  bool ordered = rec[hb::kDumpLen + 8 + hb::kDiscardLogWords] != 0 && total_discards * 5 <= 32 * hb::kDiscardLogWords;
  if (ordered) {
    std::vector<int> seen(C * R, 0);
    std::vector<Card> pile;
    for (int i = 0; i < total_discards && ordered; ++i) {
      const int bit = 5 * i, w = bit >> 5, sh = bit & 31;
      uint32_t v = (uint32_t)rec[hb::kDumpLen + 8 + w] >> sh;
      if (sh > 27 && w + 1 < hb::kDiscardLogWords) v |= (uint32_t)rec[hb::kDumpLen + 9 + w] << (32 - sh);
      const int k = (int)(v & 31u);
      if (k >= C * R) { ordered = false; break; }
      seen[k] += 1;
      Card card;
      card.color = k / R;
      card.rank = k % R;
      pile.push_back(card);
    }
    for (int k = 0; k < C * R && ordered; ++k) ordered = seen[k] == rec[241 + k];
    if (ordered) st->discard_pile = pile;
  }
  if (!ordered) {
    for (int k = 0; k < C * R; ++k)
      for (int j = 0; j < rec[241 + k]; ++j) {
        Card card;
        card.color = k / R;
        card.rank = k % R;
        st->discard_pile.push_back(card);
      }
  }
  st->turns_to_play = rec[hb::kDumpLen];
  st->history.clear();
  // Nobody sees further back than his own previous move (HanabiObservation,
  // hanabi_observation.cc:77-92), i.e. at most the newest P player moves -- each by a different
  // player, none of whom has moved since.  That is why the deal that followed a play/discard
  // can be restored: the actor's hand is full again exactly when a card was dealt, and the
  // card is the last one in his hand.
  int kept = 0;
  for (int j = 0; j <= 4; ++j)
    if ((uint32_t)rec[hb::kDumpLen + 1 + j] & 1u) ++kept; else break;
  if (kept > P) kept = P;
  for (int j = kept - 1; j >= 0; --j) {  // oldest first; [1] is the newest
    const uint32_t la = (uint32_t)rec[hb::kDumpLen + 1 + j];
    HistoryItem h;
    const int type = (la >> hb::kLaTypeSh) & 3;
    h.player = (int8_t)((la >> hb::kLaActorSh) & 7);
    if (type == hb::kLaPlay || type == hb::kLaDiscard) {
      h.move = Move(type == hb::kLaPlay ? kPlay : kDiscard, (la >> hb::kLaIdxSh) & 7, -1, -1, -1);
      h.color = (int8_t)((la >> hb::kLaColSh) & 7);
      h.rank = (int8_t)((la >> hb::kLaRankSh) & 7);
      h.scored = (la >> hb::kLaScoredSh) & 1;
      h.information_token = (la >> hb::kLaInfoSh) & 1;
    } else {
      const int off = (la >> hb::kLaOffSh) & 7, val = (la >> hb::kLaValSh) & 7;
      h.move = type == hb::kLaRevealColor ? Move(kRevealColor, -1, off, val, -1)
                                          : Move(kRevealRank, -1, off, -1, val);
      h.reveal_bitmask = (uint8_t)hb::la_reveal(la);
      h.newly_revealed_bitmask = h.reveal_bitmask;
    }
    st->history.push_back(h);
    if ((type == hb::kLaPlay || type == hb::kLaDiscard) && h.player >= 0 && h.player < P) {
      const Hand& hand = st->hands[h.player];
      if ((int)hand.cards.size() == g->hand_size && !hand.cards.empty()) {
        HistoryItem d;
        const Card& c = hand.cards.back();
        d.move = Move(kDeal, -1, -1, c.color, c.rank);
        d.player = -1;  // the chance player
        d.deal_to_player = h.player;
        st->history.push_back(d);
      }
    }
  }
  delete AS(State, state->state);
  state->state = st;
}

const void* StateParentGame(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->game;
}
void StateApplyMove(pyhanabi_state_t* state, pyhanabi_move_t* move) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && move != nullptr && move->move != nullptr);
  AS(State, state->state)->apply(*CAS(Move, move->move));
}
int StateCurPlayer(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->cur_player;
}
void StateDealRandomCard(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  AS(State, state->state)->deal_random();
}
int StateDeckSize(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->deck_total;
}
int StateFireworks(pyhanabi_state_t* state, int color) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->fireworks.at(color);
}
int StateDiscardPileSize(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return (int)CAS(State, state->state)->discard_pile.size();
}
void StateGetDiscard(pyhanabi_state_t* state, int index, pyhanabi_card_t* card) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && card != nullptr);
  const Card& c = CAS(State, state->state)->discard_pile.at(index);
  card->color = c.color;
  card->rank = c.rank;
}
int StateGetHandSize(pyhanabi_state_t* state, int pid) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return (int)CAS(State, state->state)->hands.at(pid).cards.size();
}
void StateGetHandCard(pyhanabi_state_t* state, int pid, int index, pyhanabi_card_t* card) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && card != nullptr);
  const Card& c = CAS(State, state->state)->hands.at(pid).cards.at(index);
  card->color = c.color;
  card->rank = c.rank;
}
int StateEndOfGameStatus(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->end_status();
}
int StateInformationTokens(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->info;
}
void* StateLegalMoves(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  const State* s = CAS(State, state->state);
  return new std::vector<Move>(s->legal_moves(s->cur_player));
}
int StateLifeTokens(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->life;
}
int StateNumPlayers(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->game->players;
}
int StateScore(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->score();
}
char* StateToString(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  const State* s = CAS(State, state->state);
  return Dup(BoardString(s->life, s->info, s->fireworks, s->hands, s->cur_player, s->deck_total,
                         s->discard_pile));
}
bool MoveIsLegal(const pyhanabi_state_t* state, const pyhanabi_move_t* move) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && move != nullptr && move->move != nullptr);
  return CAS(State, state->state)->legal(*CAS(Move, move->move));
}
bool CardPlayableOnFireworks(const pyhanabi_state_t* state, int color, int rank) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return CAS(State, state->state)->playable(color, rank);
}
int StateLenMoveHistory(pyhanabi_state_t* state) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr);
  return (int)CAS(State, state->state)->history.size();
}
void StateGetMoveHistory(pyhanabi_state_t* state, int index, pyhanabi_history_item_t* item) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && item != nullptr);
  item->item = new HistoryItem(CAS(State, state->state)->history.at(index));
}

void DeleteGame(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr);
  delete AS(Game, game->game);
  game->game = nullptr;
}
void NewDefaultGame(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr);
  game->game = new Game(std::unordered_map<std::string, std::string>());
}
void NewGame(pyhanabi_game_t* game, int list_length, const char** param_list) {
  HBS_REQUIRE(game != nullptr && param_list != nullptr);
  std::unordered_map<std::string, std::string> params;
  for (int p = 0; p + 1 < list_length; p += 2) params[param_list[p]] = param_list[p + 1];
  game->game = new Game(params);
}
char* GameParamString(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return Dup(CAS(Game, game->game)->param_string());
}
int NumPlayers(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->players;
}
int NumColors(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->colors;
}
int NumRanks(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->ranks;
}
int HandSize(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->hand_size;
}
int MaxInformationTokens(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->max_info;
}
int MaxLifeTokens(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->max_life;
}
int ObservationType(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->obs_type;
}
int NumCards(pyhanabi_game_t* game, int color, int rank) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->instances(color, rank);
}
int GetMoveUid(pyhanabi_game_t* game, pyhanabi_move_t* move) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr && move != nullptr && move->move != nullptr);
  return CAS(Game, game->game)->move_uid(*CAS(Move, move->move));
}
void GetMoveByUid(pyhanabi_game_t* game, int move_uid, pyhanabi_move_t* move) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr && move != nullptr);
  move->move = new Move(CAS(Game, game->game)->moves.at(move_uid));
}
int MaxMoves(pyhanabi_game_t* game) {
  HBS_REQUIRE(game != nullptr && game->game != nullptr);
  return CAS(Game, game->game)->max_moves();
}

void NewObservation(pyhanabi_state_t* state, int player, pyhanabi_observation_t* observation) {
  HBS_REQUIRE(state != nullptr && state->state != nullptr && observation != nullptr);
  observation->observation = new Observation(*CAS(State, state->state), player);
}
void DeleteObservation(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr);
  delete AS(Observation, observation->observation);
  observation->observation = nullptr;
}
char* ObsToString(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  const Observation* o = CAS(Observation, observation->observation);
  return Dup(BoardString(o->life, o->info, o->fireworks, o->hands, o->cur_player_offset,
                         o->deck_size, o->discard_pile));
}
int ObsCurPlayerOffset(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->cur_player_offset;
}
int ObsNumPlayers(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->game->players;
}
int ObsGetHandSize(pyhanabi_observation_t* observation, int pid) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return (int)CAS(Observation, observation->observation)->hands.at(pid).cards.size();
}
void ObsGetHandCard(pyhanabi_observation_t* observation, int pid, int index, pyhanabi_card_t* card) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && card != nullptr);
  const Card& c = CAS(Observation, observation->observation)->hands.at(pid).cards.at(index);
  card->color = c.color;
  card->rank = c.rank;
}
void ObsGetHandCardKnowledge(pyhanabi_observation_t* observation, int pid, int index,
                             pyhanabi_card_knowledge_t* knowledge) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && knowledge != nullptr);
  // borrowed pointer into the observation, like the reference (pyhanabi.cc:678-690)
  knowledge->knowledge = &CAS(Observation, observation->observation)->hands.at(pid).knowledge.at(index);
}
int ObsDiscardPileSize(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return (int)CAS(Observation, observation->observation)->discard_pile.size();
}
void ObsGetDiscard(pyhanabi_observation_t* observation, int index, pyhanabi_card_t* card) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && card != nullptr);
  const Card& c = CAS(Observation, observation->observation)->discard_pile.at(index);
  card->color = c.color;
  card->rank = c.rank;
}
int ObsFireworks(pyhanabi_observation_t* observation, int color) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->fireworks.at(color);
}
int ObsDeckSize(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->deck_size;
}
int ObsNumLastMoves(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return (int)CAS(Observation, observation->observation)->last_moves.size();
}
void ObsGetLastMove(pyhanabi_observation_t* observation, int index, pyhanabi_history_item_t* item) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && item != nullptr);
  item->item = new HistoryItem(CAS(Observation, observation->observation)->last_moves.at(index));
}
int ObsInformationTokens(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->info;
}
int ObsLifeTokens(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->life;
}
int ObsNumLegalMoves(pyhanabi_observation_t* observation) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return (int)CAS(Observation, observation->observation)->legal_moves.size();
}
void ObsGetLegalMove(pyhanabi_observation_t* observation, int index, pyhanabi_move_t* move) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && move != nullptr);
  move->move = new Move(CAS(Observation, observation->observation)->legal_moves.at(index));
}
bool ObsCardPlayableOnFireworks(const pyhanabi_observation_t* observation, int color, int rank) {
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  return CAS(Observation, observation->observation)->playable(color, rank);
}

void NewObservationEncoder(pyhanabi_observation_encoder_t* encoder, pyhanabi_game_t* game, int type) {
  HBS_REQUIRE(encoder != nullptr && game != nullptr && game->game != nullptr);
  if (type != 0) {  // only the canonical encoder exists (observation_encoder.h)
    std::fprintf(stderr, "Encoder type not recognized.\n");
    encoder->encoder = nullptr;
    std::abort();
  }
  Encoder* e = new Encoder;
  e->game = CAS(Game, game->game);
  e->type = type;
  encoder->encoder = e;
}
void DeleteObservationEncoder(pyhanabi_observation_encoder_t* encoder) {
  HBS_REQUIRE(encoder != nullptr);
  delete AS(Encoder, encoder->encoder);
  encoder->encoder = nullptr;
}
char* ObservationShape(pyhanabi_observation_encoder_t* encoder) {
  HBS_REQUIRE(encoder != nullptr && encoder->encoder != nullptr);
  return Dup(std::to_string(CAS(Encoder, encoder->encoder)->shape()));
}
char* EncodeObservation(pyhanabi_observation_encoder_t* encoder, pyhanabi_observation_t* observation) {
  HBS_REQUIRE(encoder != nullptr && encoder->encoder != nullptr);
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr);
  const std::vector<int> bits =
      CAS(Encoder, encoder->encoder)->encode(*CAS(Observation, observation->observation));
  std::string s;
  s.reserve(bits.size() * 2);
  for (size_t i = 0; i < bits.size(); ++i) {
    if (i) s += ',';
    s += bits[i] ? '1' : '0';
  }
  return Dup(s);
}

// Binary twin of EncodeObservation: one byte per bit into a caller buffer of
// ObservationShape bytes (no ASCII round trip).  Returns the length written.
int EncodeObservationBytes(pyhanabi_observation_encoder_t* encoder,
                           pyhanabi_observation_t* observation, uint8_t* out, int capacity) {
  HBS_REQUIRE(encoder != nullptr && encoder->encoder != nullptr);
  HBS_REQUIRE(observation != nullptr && observation->observation != nullptr && out != nullptr);
  const std::vector<int> bits =
      CAS(Encoder, encoder->encoder)->encode(*CAS(Observation, observation->observation));
  HBS_REQUIRE((int)bits.size() <= capacity);
  for (size_t i = 0; i < bits.size(); ++i) out[i] = (uint8_t)bits[i];
  return (int)bits.size();
}

}  // extern "C"
