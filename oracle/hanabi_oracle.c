/*
 * oracle/hanabi_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).
 *
 * Plain-C restatement of the reference engine's batched hot path.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference/hanabi_learning_environment/).  The state here is deliberately
 * UNPACKED (one int per field) so that it shares no representation tricks with
 * the packed CUDA engine it is used to check.
 *
 * The deal sampler depends on a third-party library that is not in the
 * reference tree: libstdc++ (GCC 13.3) <random> -- std::mt19937,
 * std::discrete_distribution<unsigned long>, std::generate_canonical<double,53>,
 * std::uniform_int_distribution (bits/random.tcc:2655-2712, 3346-3385;
 * bits/uniform_int_dist.h:256-282).  Its published algorithm is restated in
 * orc_mt_*, orc_canonical, orc_discrete and orc_uniform_int below and pinned
 * against the real library through oracle/_ref (RefDiscreteSample,
 * RefMtWords, RefUniformInts).
 *
 * Compile with -ffp-contract=off: the fp64 sums must round after every
 * operation exactly as the x86-64 SSE2 code of the reference does.
 */
#include "hanabi_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { MT_N = 624, MT_M = 397 };
enum { MAXP = 5, MAXH = 8, MAXC = 5, MAXR = 5, MAXT = 25 };
/* Move types, hanabi_lib/hanabi_move.h:34. */
enum { kInvalid = 0, kPlay = 1, kDiscard = 2, kRevealColor = 3, kRevealRank = 4, kDeal = 5 };
/* hanabi_lib/hanabi_state.h:61-66 */
enum { kNotFinished = 0, kOutOfLifeTokens = 1, kOutOfCards = 2, kCompletedFireworks = 3 };
/* hanabi_lib/hanabi_game.h:39 */
enum { kMinimal = 0, kCardKnowledge = 1, kSeer = 2 };

typedef struct {
  uint32_t mt[MT_N];
  int mti;
} OrcMt;

typedef struct {
  int P, C, R, H, IT, LT, obs_type, random_start;
  int deck_max, obs_size, max_moves;
} OrcSpec;

typedef struct {
  int cur_player, next_player, info, life, turns_to_play, deck_size;
  int fireworks[MAXC];
  int hand_n[MAXP];
  int card[MAXP][MAXH];   /* color*R+rank */
  int cplaus[MAXP][MAXH]; /* bit c: color c still plausible */
  int rplaus[MAXP][MAXH];
  int chint[MAXP][MAXH];  /* hinted color or -1 */
  int rhint[MAXP][MAXH];
  int deck[MAXT];
  int discard[MAXT];
  /* the most recent non-deal history item, absolute player index */
  int la_valid, la_type, la_player, la_target_offset, la_color, la_rank;
  int la_card_index, la_card_color, la_card_rank, la_reveal, la_scored, la_info_token;
} OrcState;

typedef struct {
  OrcSpec spec;
  int n;
  OrcMt* rng;     /* one stream per game: hanabi_lib/hanabi_game.h:114 */
  OrcState* st;
  uint64_t* policy_rng;
} OrcBatch;

/* ---------------------------------------------------------------- mt19937 */
/* std::mt19937::seed(value): x[0]=value; x[i]=1812433253*(x[i-1]^(x[i-1]>>30))+i
 * (bits/random.tcc, mersenne_twister_engine::seed). */
static void orc_mt_seed(OrcMt* m, uint32_t seed) {
  m->mt[0] = seed;
  for (int i = 1; i < MT_N; ++i)
    m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
  m->mti = MT_N;
}

static uint32_t orc_mt_next(OrcMt* m) {
  if (m->mti >= MT_N) { /* _M_gen_rand: regenerate the whole block */
    for (int k = 0; k < MT_N; ++k) {
      uint32_t y = (m->mt[k] & 0x80000000u) | (m->mt[(k + 1) % MT_N] & 0x7fffffffu);
      m->mt[k] = m->mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    m->mti = 0;
  }
  uint32_t z = m->mt[m->mti++];
  z ^= (z >> 11);
  z ^= (z << 7) & 0x9d2c5680u;
  z ^= (z << 15) & 0xefc60000u;
  z ^= (z >> 18);
  return z;
}

/* generate_canonical<double,53> over a 32-bit URNG: two words, low word first
 * (bits/random.tcc:3346-3385).  Words are passed in so the same code serves
 * the replay test. */
static double orc_canonical(uint32_t lo, uint32_t hi) {
  double sum = 0.0, tmp = 1.0;
  sum += (double)lo * tmp;
  tmp *= 4294967296.0;
  sum += (double)hi * tmp;
  tmp *= 4294967296.0;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}

/* discrete_distribution::param_type::_M_initialize + operator()
 * (bits/random.tcc:2655-2712).  Returns the index; *used tells whether two
 * words were consumed (none when fewer than 2 weights). */
static int orc_discrete(const double* w, int n, uint32_t lo, uint32_t hi, int* used) {
  double p[MAXT], cp[MAXT];
  if (n < 2) {
    if (used) *used = 0;
    return 0;
  }
  if (used) *used = 1;
  double sum = 0.0;
  for (int k = 0; k < n; ++k) sum = sum + w[k]; /* std::accumulate(.., 0.0) */
  for (int k = 0; k < n; ++k) p[k] = w[k] / sum; /* __normalize */
  cp[0] = p[0];                                  /* std::partial_sum */
  for (int k = 1; k < n; ++k) cp[k] = cp[k - 1] + p[k];
  cp[n - 1] = 1.0;
  double u = orc_canonical(lo, hi);
  int k = 0; /* std::lower_bound: first k with !(cp[k] < u) */
  while (k < n && cp[k] < u) ++k;
  return k;
}

/* uniform_int_distribution<unsigned long>(0, hi) on a 32-bit URNG: Lemire's
 * method, bits/uniform_int_dist.h:256-282 (the __urngrange == UINT32_MAX branch). */
static uint32_t orc_uniform_int(OrcMt* m, uint32_t hi) {
  uint32_t range = hi + 1u;
  uint64_t product = (uint64_t)orc_mt_next(m) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)orc_mt_next(m) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32);
}

/* ------------------------------------------------------------------- spec */
/* HanabiGame::NumberCardInstances, hanabi_lib/hanabi_game.cc:126-136. */
static int orc_instances(const OrcSpec* s, int rank) {
  if (rank < 0 || rank >= s->R) return 0;
  if (rank == 0) return 3;
  if (rank == s->R - 1) return 1;
  return 2;
}

static const char* kv_find(int n_kv, const char** kv, const char* key) {
  const char* v = NULL;
  for (int k = 0; k + 1 < n_kv; k += 2)
    if (strcmp(kv[k], key) == 0) v = kv[k + 1];
  return v;
}
static int kv_int(int n_kv, const char** kv, const char* key, int dflt) {
  const char* v = kv_find(n_kv, kv, key); /* ParameterValue<int>: std::stoi, util.cc:37-47 */
  return v ? (int)strtol(v, NULL, 10) : dflt;
}
static int kv_bool(int n_kv, const char** kv, const char* key, int dflt) {
  const char* v = kv_find(n_kv, kv, key); /* ParameterValue<bool>, util.cc:74-86 */
  if (!v) return dflt;
  return (strcmp(v, "1") == 0 || strcmp(v, "true") == 0 || strcmp(v, "True") == 0) ? 1 : 0;
}

/* HanabiGame ctor, hanabi_lib/hanabi_game.cc:29-67; section lengths
 * hanabi_lib/canonical_encoders.cc:52-55,107-112,169,213-223,340-343,423-431;
 * MaxMoves hanabi_lib/hanabi_game.cc:69-72. */
static int orc_spec_init(OrcSpec* s, int n_kv, const char** kv) {
  s->P = kv_int(n_kv, kv, "players", 2);
  s->C = kv_int(n_kv, kv, "colors", MAXC);
  s->R = kv_int(n_kv, kv, "ranks", MAXR);
  s->H = kv_int(n_kv, kv, "hand_size", s->P < 4 ? 5 : 4);
  s->IT = kv_int(n_kv, kv, "max_information_tokens", 8);
  s->LT = kv_int(n_kv, kv, "max_life_tokens", 3);
  s->random_start = kv_bool(n_kv, kv, "random_start_player", 0);
  s->obs_type = kv_int(n_kv, kv, "observation_type", kCardKnowledge);
  if (s->P < 2 || s->P > MAXP || s->C < 1 || s->C > MAXC || s->R < 1 || s->R > MAXR) return -1;
  if (s->H < 1 || s->H > MAXH) return -1;
  int per_color = 0;
  for (int r = 0; r < s->R; ++r) per_color += orc_instances(s, r);
  s->deck_max = per_color * s->C;
  if (s->H * s->P > s->deck_max) return -1;
  int bits_per_card = s->C * s->R;
  int hands = (s->P - 1) * s->H * bits_per_card + s->P;
  int board = s->deck_max - s->P * s->H + s->C * s->R + s->IT + s->LT;
  int discards = s->deck_max;
  int last = s->P + 4 + s->P + s->C + s->R + s->H + s->H + bits_per_card + 2;
  int know = s->P * s->H * (bits_per_card + s->C + s->R);
  s->obs_size = hands + board + discards + last + (s->obs_type == kMinimal ? 0 : know);
  s->max_moves = 2 * s->H + (s->P - 1) * s->C + (s->P - 1) * s->R;
  return 0;
}

/* ------------------------------------------------------------------ moves */
typedef struct { int type, card_index, target_offset, color, rank; } OrcMove;

/* HanabiGame::ConstructMove, hanabi_lib/hanabi_game.cc:159-183. */
static OrcMove orc_move_from_uid(const OrcSpec* s, int uid) {
  OrcMove m = {kInvalid, -1, -1, -1, -1};
  if (uid < 0 || uid >= s->max_moves) return m;
  if (uid < s->H) { m.type = kDiscard; m.card_index = uid; return m; }
  uid -= s->H;
  if (uid < s->H) { m.type = kPlay; m.card_index = uid; return m; }
  uid -= s->H;
  if (uid < (s->P - 1) * s->C) {
    m.type = kRevealColor; m.target_offset = 1 + uid / s->C; m.color = uid % s->C; return m;
  }
  uid -= (s->P - 1) * s->C;
  m.type = kRevealRank; m.target_offset = 1 + uid / s->R; m.rank = uid % s->R;
  return m;
}

/* HanabiState::PlayerToDeal, hanabi_lib/hanabi_state.cc:157-164. */
static int orc_player_to_deal(const OrcSpec* s, const OrcState* st) {
  for (int i = 0; i < s->P; ++i)
    if (st->hand_n[i] < s->H) return i;
  return -1;
}

/* HanabiState::MoveIsLegal (player moves), hanabi_lib/hanabi_state.cc:166-219. */
static int orc_move_is_legal(const OrcSpec* s, const OrcState* st, OrcMove m) {
  int cur = st->cur_player;
  switch (m.type) {
    case kDiscard:
      if (st->info >= s->IT) return 0;
      if (m.card_index >= st->hand_n[cur]) return 0;
      return 1;
    case kPlay:
      if (m.card_index >= st->hand_n[cur]) return 0;
      return 1;
    case kRevealColor:
    case kRevealRank: {
      if (st->info <= 0) return 0; /* HintingIsLegal :146-155 */
      if (m.target_offset < 1 || m.target_offset >= s->P) return 0;
      int t = (cur + m.target_offset) % s->P;
      for (int i = 0; i < st->hand_n[t]; ++i) {
        int c = st->card[t][i] / s->R, r = st->card[t][i] % s->R;
        if (m.type == kRevealColor ? (c == m.color) : (r == m.rank)) return 1;
      }
      return 0;
    }
    default:
      return 0;
  }
}

/* HanabiState::AdvanceToNextPlayer, hanabi_lib/hanabi_state.cc:104-111. */
static void orc_advance(const OrcSpec* s, OrcState* st) {
  if (st->deck_size != 0 && orc_player_to_deal(s, st) >= 0) {
    st->cur_player = -1;
  } else {
    st->cur_player = st->next_player;
    st->next_player = (st->cur_player + 1) % s->P;
  }
}

/* HanabiHand::RemoveFromHand, hanabi_lib/hanabi_hand.cc:87-94. */
static void orc_remove_card(OrcState* st, int p, int idx, int to_discard) {
  if (to_discard) st->discard[st->card[p][idx]] += 1;
  for (int i = idx; i + 1 < st->hand_n[p]; ++i) {
    st->card[p][i] = st->card[p][i + 1];
    st->cplaus[p][i] = st->cplaus[p][i + 1];
    st->rplaus[p][i] = st->rplaus[p][i + 1];
    st->chint[p][i] = st->chint[p][i + 1];
    st->rhint[p][i] = st->rhint[p][i + 1];
  }
  st->hand_n[p] -= 1;
}

/* HanabiState::IncrementInformationTokens, hanabi_lib/hanabi_state.cc:113-120. */
static int orc_inc_info(const OrcSpec* s, OrcState* st) {
  if (st->info < s->IT) { st->info += 1; return 1; }
  return 0;
}

/* HanabiState::ApplyMove for player moves, hanabi_lib/hanabi_state.cc:221-275. */
static void orc_apply_move(const OrcSpec* s, OrcState* st, OrcMove m) {
  int cur = st->cur_player;
  if (st->deck_size == 0) st->turns_to_play -= 1; /* :223-225 */
  st->la_valid = 1; st->la_type = m.type; st->la_player = cur;
  st->la_target_offset = m.target_offset; st->la_color = m.color; st->la_rank = m.rank;
  st->la_card_index = m.card_index; st->la_card_color = -1; st->la_card_rank = -1;
  st->la_reveal = 0; st->la_scored = 0; st->la_info_token = 0;
  switch (m.type) {
    case kDiscard: { /* :242-247 */
      st->la_info_token = orc_inc_info(s, st);
      st->la_card_color = st->card[cur][m.card_index] / s->R;
      st->la_card_rank = st->card[cur][m.card_index] % s->R;
      orc_remove_card(st, cur, m.card_index, 1);
      break;
    }
    case kPlay: { /* :248-255 with AddToFireworks :132-144 */
      int c = st->card[cur][m.card_index] / s->R, r = st->card[cur][m.card_index] % s->R;
      st->la_card_color = c; st->la_card_rank = r;
      if (r == st->fireworks[c]) {
        st->fireworks[c] += 1;
        st->la_scored = 1;
        if (st->fireworks[c] == s->R) st->la_info_token = orc_inc_info(s, st);
      } else {
        st->life -= 1;
      }
      orc_remove_card(st, cur, m.card_index, st->la_scored ? 0 : 1);
      break;
    }
    case kRevealColor: { /* :256-262, HanabiHand::RevealColor hanabi_hand.cc:96-110 */
      st->info -= 1;
      int t = (cur + m.target_offset) % s->P;
      for (int i = 0; i < st->hand_n[t]; ++i) {
        if (st->card[t][i] / s->R == m.color) {
          st->la_reveal |= 1 << i;
          st->chint[t][i] = m.color;          /* ApplyIsValueHint :29-36 */
          st->cplaus[t][i] = 1 << m.color;
        } else {
          st->cplaus[t][i] &= ~(1 << m.color); /* ApplyIsNotValueHint :38-42 */
        }
      }
      break;
    }
    case kRevealRank: { /* :263-269, HanabiHand::RevealRank hanabi_hand.cc:112-126 */
      st->info -= 1;
      int t = (cur + m.target_offset) % s->P;
      for (int i = 0; i < st->hand_n[t]; ++i) {
        if (st->card[t][i] % s->R == m.rank) {
          st->la_reveal |= 1 << i;
          st->rhint[t][i] = m.rank;
          st->rplaus[t][i] = 1 << m.rank;
        } else {
          st->rplaus[t][i] &= ~(1 << m.rank);
        }
      }
      break;
    }
    default:
      abort();
  }
  orc_advance(s, st);
}

/* HanabiState::ApplyRandomChance -> ChanceOutcomes -> PickRandomChance ->
 * ApplyMove(kDeal): hanabi_lib/hanabi_state.cc:282-286, 313-325, 277-280,
 * hanabi_lib/hanabi_game.cc:106-112, hanabi_lib/hanabi_state.cc:229-241. */
static void orc_deal_random(const OrcSpec* s, OrcState* st, OrcMt* rng) {
  double w[MAXT];
  int idx[MAXT], n = 0;
  for (int k = 0; k < s->C * s->R; ++k)
    if (st->deck[k] > 0) {
      idx[n] = k;
      w[n] = (double)st->deck[k] / (double)st->deck_size;
      ++n;
    }
  if (n == 0) abort();
  uint32_t lo = 0, hi = 0;
  if (n >= 2) { lo = orc_mt_next(rng); hi = orc_mt_next(rng); }
  int type = idx[orc_discrete(w, n, lo, hi, NULL)];
  /* ApplyMove(kDeal): the deck is never empty here so turns_to_play is untouched. */
  int p = orc_player_to_deal(s, st);
  int i = st->hand_n[p];
  st->card[p][i] = type;
  st->cplaus[p][i] = (1 << s->C) - 1;
  st->rplaus[p][i] = (1 << s->R) - 1;
  st->chint[p][i] = -1;
  st->rhint[p][i] = -1;
  if (s->obs_type == kSeer) { /* :233-236 */
    st->chint[p][i] = type / s->R; st->cplaus[p][i] = 1 << (type / s->R);
    st->rhint[p][i] = type % s->R; st->rplaus[p][i] = 1 << (type % s->R);
  }
  st->hand_n[p] += 1;
  st->deck[type] -= 1;
  st->deck_size -= 1;
  orc_advance(s, st);
}

/* HanabiState::Score / EndOfGameStatus, hanabi_lib/hanabi_state.cc:359-377. */
static int orc_score(const OrcSpec* s, const OrcState* st) {
  if (st->life <= 0) return 0;
  int sum = 0;
  for (int c = 0; c < s->C; ++c) sum += st->fireworks[c];
  return sum;
}
static int orc_end_status(const OrcSpec* s, const OrcState* st) {
  if (st->life < 1) return kOutOfLifeTokens;
  if (orc_score(s, st) >= s->C * s->R) return kCompletedFireworks;
  if (st->turns_to_play <= 0) return kOutOfCards;
  return kNotFinished;
}

/* HanabiState ctor + rl_env.HanabiEnv.reset: hanabi_lib/hanabi_state.cc:90-102,
 * hanabi_lib/hanabi_game.cc:138-145, rl_env.py:210-217. */
static void orc_reset(const OrcSpec* s, OrcState* st, OrcMt* rng) {
  memset(st, 0, sizeof(*st));
  st->cur_player = -1;
  st->next_player = s->random_start ? (int)orc_uniform_int(rng, (uint32_t)(s->P - 1)) : 0;
  st->info = s->IT;
  st->life = s->LT;
  st->turns_to_play = s->P;
  for (int c = 0; c < s->C; ++c)
    for (int r = 0; r < s->R; ++r) {
      st->deck[c * s->R + r] = orc_instances(s, r);
      st->deck_size += orc_instances(s, r);
    }
  /* The constructor leaves cur_player at the chance player; rl_env deals until
   * a real player is to act.  With an empty deck AdvanceToNextPlayer is only
   * reached from ApplyMove, so a P*H == deck game still gets its P*H deals. */
  while (st->cur_player == -1) orc_deal_random(s, st, rng);
}

/* HanabiState::LegalMoves as a 0/1 mask over move uids, hanabi_lib/hanabi_state.cc:288-304. */
static void orc_legal_mask(const OrcSpec* s, const OrcState* st, uint8_t* mask) {
  for (int uid = 0; uid < s->max_moves; ++uid)
    mask[uid] = (uint8_t)orc_move_is_legal(s, st, orc_move_from_uid(s, uid));
}

/* CanonicalObservationEncoder::Encode for `observer`, with the observation
 * view of HanabiObservation folded in (hanabi_lib/canonical_encoders.cc:433-451,
 * hanabi_lib/hanabi_observation.cc:52-93). */
static void orc_encode(const OrcSpec* s, const OrcState* st, int observer, uint8_t* e) {
  const int P = s->P, C = s->C, R = s->R, H = s->H, bpc = s->C * s->R;
  memset(e, 0, (size_t)s->obs_size);
  int off = 0;
  /* EncodeHands :62-105: other players' cards, then "hand is short" bits. */
  for (int q = 1; q < P; ++q) {
    int p = (observer + q) % P;
    for (int i = 0; i < st->hand_n[p]; ++i) e[off + i * bpc + st->card[p][i]] = 1;
    off += H * bpc;
  }
  for (int q = 0; q < P; ++q)
    if (st->hand_n[(observer + q) % P] < H) e[off + q] = 1;
  off += P;
  /* EncodeBoard :123-167 */
  for (int i = 0; i < st->deck_size; ++i) e[off + i] = 1;
  off += s->deck_max - H * P;
  for (int c = 0; c < C; ++c) {
    if (st->fireworks[c] > 0) e[off + st->fireworks[c] - 1] = 1;
    off += R;
  }
  for (int i = 0; i < st->info; ++i) e[off + i] = 1;
  off += s->IT;
  for (int i = 0; i < st->life; ++i) e[off + i] = 1;
  off += s->LT;
  /* EncodeDiscards :188-211 */
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      for (int i = 0; i < st->discard[c * R + r]; ++i) e[off + i] = 1;
      off += orc_instances(s, r);
    }
  /* EncodeLastAction :236-338; player made observer-relative as in
   * ChangeHistoryItemToObserverRelative hanabi_observation.cc:34-49. */
  if (st->la_valid) {
    int rel = (st->la_player - observer + P) % P;
    int o = off;
    e[o + rel] = 1;
    o += P;
    switch (st->la_type) {
      case kPlay: e[o] = 1; break;
      case kDiscard: e[o + 1] = 1; break;
      case kRevealColor: e[o + 2] = 1; break;
      case kRevealRank: e[o + 3] = 1; break;
      default: abort();
    }
    o += 4;
    int is_hint = st->la_type == kRevealColor || st->la_type == kRevealRank;
    int is_card = st->la_type == kPlay || st->la_type == kDiscard;
    if (is_hint) e[o + (rel + st->la_target_offset) % P] = 1;
    o += P;
    if (st->la_type == kRevealColor) e[o + st->la_color] = 1;
    o += C;
    if (st->la_type == kRevealRank) e[o + st->la_rank] = 1;
    o += R;
    if (is_hint)
      for (int i = 0; i < H; ++i)
        if (st->la_reveal & (1 << i)) e[o + i] = 1;
    o += H;
    if (is_card) e[o + st->la_card_index] = 1;
    o += H;
    if (is_card) e[o + st->la_card_color * R + st->la_card_rank] = 1;
    o += bpc;
    if (st->la_type == kPlay) {
      if (st->la_scored) e[o] = 1;
      if (st->la_info_token) e[o + 1] = 1;
    }
  }
  off += P + 4 + P + C + R + H + H + bpc + 2;
  /* EncodeCardKnowledge :366-419 (absent for kMinimal :446-448). */
  if (s->obs_type != kMinimal) {
    for (int q = 0; q < P; ++q) {
      int p = (observer + q) % P;
      for (int i = 0; i < st->hand_n[p]; ++i) {
        int o = off + i * (bpc + C + R);
        for (int c = 0; c < C; ++c)
          if (st->cplaus[p][i] & (1 << c))
            for (int r = 0; r < R; ++r)
              if (st->rplaus[p][i] & (1 << r)) e[o + c * R + r] = 1;
        if (st->chint[p][i] >= 0) e[o + bpc + st->chint[p][i]] = 1;
        if (st->rhint[p][i] >= 0) e[o + bpc + C + st->rhint[p][i]] = 1;
      }
      off += H * (bpc + C + R);
    }
  }
  if (off != s->obs_size) abort();
}

/* rl_env.HanabiEnv.step, rl_env.py:343-366.  Returns 1 for an illegal action
 * (the reference would abort in REQUIRE, hanabi_state.cc:222). */
static int orc_step(const OrcSpec* s, OrcState* st, OrcMt* rng, int action, int auto_reset,
                    int32_t* reward, uint8_t* done, int32_t* score) {
  if (orc_end_status(s, st) != kNotFinished) {
    *reward = 0; *done = 1;
    if (score) *score = orc_score(s, st);
    return 0;
  }
  OrcMove m = orc_move_from_uid(s, action);
  if (m.type == kInvalid || !orc_move_is_legal(s, st, m)) {
    *reward = 0; *done = 0;
    if (score) *score = orc_score(s, st);
    return 1;
  }
  int last = orc_score(s, st);
  orc_apply_move(s, st, m);
  while (st->cur_player == -1) orc_deal_random(s, st, rng);
  *reward = orc_score(s, st) - last;
  *done = orc_end_status(s, st) != kNotFinished;
  if (score) *score = orc_score(s, st);
  if (*done && auto_reset) orc_reset(s, st, rng);
  return 0;
}

/* -------------------------------------------------------------- batch API */
void* OrcBatchNew(int n_kv, const char** kv, int num_games, const int32_t* seeds) {
  OrcBatch* b = (OrcBatch*)calloc(1, sizeof(OrcBatch));
  if (orc_spec_init(&b->spec, n_kv, kv) != 0) { free(b); return NULL; }
  b->n = num_games;
  b->rng = (OrcMt*)calloc((size_t)num_games, sizeof(OrcMt));
  b->st = (OrcState*)calloc((size_t)num_games, sizeof(OrcState));
  b->policy_rng = (uint64_t*)calloc((size_t)num_games, sizeof(uint64_t));
  for (int i = 0; i < num_games; ++i) {
    orc_mt_seed(&b->rng[i], (uint32_t)seeds[i]); /* rng_.seed(seed_), hanabi_game.cc:51 */
    b->policy_rng[i] = 0x1234567ull + (uint64_t)i * 7919ull;
  }
  return b;
}
void OrcBatchDelete(void* vb) {
  OrcBatch* b = (OrcBatch*)vb;
  if (!b) return;
  free(b->rng); free(b->st); free(b->policy_rng); free(b);
}
int OrcBatchObservationSize(void* vb) { return ((OrcBatch*)vb)->spec.obs_size; }
int OrcBatchMaxMoves(void* vb) { return ((OrcBatch*)vb)->spec.max_moves; }
int OrcBatchNumPlayers(void* vb) { return ((OrcBatch*)vb)->spec.P; }
int OrcBatchNumGames(void* vb) { return ((OrcBatch*)vb)->n; }

void OrcBatchReset(void* vb, const uint8_t* which) {
  OrcBatch* b = (OrcBatch*)vb;
  for (int i = 0; i < b->n; ++i)
    if (!which || which[i]) orc_reset(&b->spec, &b->st[i], &b->rng[i]);
}
void OrcBatchStep(void* vb, const int32_t* actions, int32_t* rewards, uint8_t* dones,
                  int32_t* scores, int auto_reset, uint8_t* illegal) {
  OrcBatch* b = (OrcBatch*)vb;
  for (int i = 0; i < b->n; ++i) {
    int rc = orc_step(&b->spec, &b->st[i], &b->rng[i], actions[i], auto_reset, &rewards[i],
                      &dones[i], scores ? &scores[i] : NULL);
    if (illegal) illegal[i] = (uint8_t)rc;
  }
}
void OrcBatchEncode(void* vb, int observer, uint8_t* obs, int64_t pitch) {
  OrcBatch* b = (OrcBatch*)vb;
  for (int i = 0; i < b->n; ++i)
    orc_encode(&b->spec, &b->st[i], observer < 0 ? b->st[i].cur_player : observer,
               obs + pitch * (int64_t)i);
}
void OrcBatchLegalMask(void* vb, uint8_t* mask) {
  OrcBatch* b = (OrcBatch*)vb;
  for (int i = 0; i < b->n; ++i)
    orc_legal_mask(&b->spec, &b->st[i], mask + (int64_t)b->spec.max_moves * i);
}
void OrcBatchCurPlayer(void* vb, int32_t* cur) {
  OrcBatch* b = (OrcBatch*)vb;
  for (int i = 0; i < b->n; ++i) cur[i] = b->st[i].cur_player;
}

void OrcBatchDumpState(void* vb, int32_t* out) {
  OrcBatch* b = (OrcBatch*)vb;
  const OrcSpec* s = &b->spec;
  for (int g = 0; g < b->n; ++g) {
    const OrcState* st = &b->st[g];
    int32_t* d = out + (int64_t)HB_DUMP_LEN * g;
    memset(d, 0, sizeof(int32_t) * HB_DUMP_LEN);
    d[HB_DUMP_CUR] = st->cur_player;
    d[HB_DUMP_INFO] = st->info;
    d[HB_DUMP_LIFE] = st->life;
    d[HB_DUMP_DECK] = st->deck_size;
    d[HB_DUMP_SCORE] = orc_score(s, st);
    d[HB_DUMP_END] = orc_end_status(s, st);
    for (int c = 0; c < s->C; ++c) d[HB_DUMP_FIREWORKS + c] = st->fireworks[c];
    for (int p = 0; p < MAXP; ++p)
      for (int k = 0; k < MAXH; ++k) {
        int have = p < s->P && k < st->hand_n[p];
        d[HB_DUMP_CARDS + p * 8 + k] = have ? st->card[p][k] : -1;
        d[HB_DUMP_CPLAUS + p * 8 + k] = have ? st->cplaus[p][k] : 0;
        d[HB_DUMP_RPLAUS + p * 8 + k] = have ? st->rplaus[p][k] : 0;
        d[HB_DUMP_CHINT + p * 8 + k] = have ? st->chint[p][k] : -1;
        d[HB_DUMP_RHINT + p * 8 + k] = have ? st->rhint[p][k] : -1;
      }
    for (int p = 0; p < s->P; ++p) d[HB_DUMP_HANDSIZE + p] = st->hand_n[p];
    for (int k = 0; k < s->C * s->R; ++k) {
      d[HB_DUMP_DECKCOUNT + k] = st->deck[k];
      d[HB_DUMP_DISCARDCOUNT + k] = st->discard[k];
    }
  }
}

void OrcBatchLoadState(void* vb, int game, const int32_t* d, int next_player, int turns_to_play) {
  OrcBatch* b = (OrcBatch*)vb;
  const OrcSpec* s = &b->spec;
  OrcState* st = &b->st[game];
  memset(st, 0, sizeof(*st));
  st->cur_player = d[HB_DUMP_CUR];
  st->next_player = next_player;
  st->info = d[HB_DUMP_INFO];
  st->life = d[HB_DUMP_LIFE];
  st->deck_size = d[HB_DUMP_DECK];
  st->turns_to_play = turns_to_play;
  for (int c = 0; c < s->C; ++c) st->fireworks[c] = d[HB_DUMP_FIREWORKS + c];
  for (int p = 0; p < s->P; ++p) {
    st->hand_n[p] = d[HB_DUMP_HANDSIZE + p];
    for (int k = 0; k < st->hand_n[p]; ++k) {
      st->card[p][k] = d[HB_DUMP_CARDS + p * 8 + k];
      st->cplaus[p][k] = d[HB_DUMP_CPLAUS + p * 8 + k];
      st->rplaus[p][k] = d[HB_DUMP_RPLAUS + p * 8 + k];
      st->chint[p][k] = d[HB_DUMP_CHINT + p * 8 + k];
      st->rhint[p][k] = d[HB_DUMP_RHINT + p * 8 + k];
    }
  }
  for (int k = 0; k < s->C * s->R; ++k) {
    st->deck[k] = d[HB_DUMP_DECKCOUNT + k];
    st->discard[k] = d[HB_DUMP_DISCARDCOUNT + k];
  }
}

static uint64_t orc_splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* Timed CPU port baseline (single thread): the same loop as RefBatchRollout. */
int64_t OrcBatchRollout(void* vb, int steps, uint8_t* obs, int64_t pitch, uint8_t* mask,
                        int64_t* episodes_out, int64_t* score_sum_out) {
  OrcBatch* b = (OrcBatch*)vb;
  const OrcSpec* s = &b->spec;
  int64_t episodes = 0, score_sum = 0;
  uint8_t m[64];
  for (int step = 0; step < steps; ++step)
    for (int i = 0; i < b->n; ++i) {
      OrcState* st = &b->st[i];
      orc_legal_mask(s, st, m);
      int n_legal = 0, legal[64];
      for (int a = 0; a < s->max_moves; ++a)
        if (m[a]) legal[n_legal++] = a;
      int action = legal[orc_splitmix(&b->policy_rng[i]) % (uint64_t)n_legal];
      int32_t reward, score; uint8_t done;
      orc_step(s, st, &b->rng[i], action, 1, &reward, &done, &score);
      if (done) { episodes += 1; score_sum += score; }
      orc_encode(s, st, st->cur_player, obs + pitch * (int64_t)i);
      orc_legal_mask(s, st, mask + (int64_t)s->max_moves * i);
    }
  if (episodes_out) *episodes_out = episodes;
  if (score_sum_out) *score_sum_out = score_sum;
  return (int64_t)steps * b->n;
}

/* ----------------------------------------------------- sampler test hooks */
int OrcDiscreteSample(const double* weights, int n, uint32_t lo, uint32_t hi) {
  return orc_discrete(weights, n, lo, hi, NULL);
}
int OrcDealFromCounts(const int32_t* counts, int n_types, uint32_t lo, uint32_t hi) {
  double w[MAXT];
  int idx[MAXT], n = 0, total = 0;
  for (int k = 0; k < n_types; ++k) total += counts[k];
  for (int k = 0; k < n_types; ++k)
    if (counts[k] > 0) { idx[n] = k; w[n] = (double)counts[k] / (double)total; ++n; }
  if (n == 0) return -1;
  return idx[orc_discrete(w, n, lo, hi, NULL)];
}
void OrcMtWords(int32_t seed, int n, uint32_t* out) {
  OrcMt m;
  orc_mt_seed(&m, (uint32_t)seed);
  for (int i = 0; i < n; ++i) out[i] = orc_mt_next(&m);
}
void OrcUniformInts(int32_t seed, int hi, int n, int32_t* out) {
  OrcMt m;
  orc_mt_seed(&m, (uint32_t)seed);
  for (int i = 0; i < n; ++i) out[i] = (int32_t)orc_uniform_int(&m, (uint32_t)hi);
}
