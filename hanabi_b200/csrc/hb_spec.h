// hb_spec.h -- game rule parameters shared by host and device code.
//
// Mirrors the parameter surface of the reference's HanabiGame constructor
// (hanabi_learning_environment/hanabi_lib/hanabi_game.cc:29-67): same keys,
// same defaults, same bounds.  Section lengths follow
// hanabi_lib/canonical_encoders.cc:52-55,107-112,169,213-223,340-343,423-431 and
// MaxMoves follows hanabi_lib/hanabi_game.cc:69-72.
#ifndef HB_SPEC_H_
#define HB_SPEC_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

namespace hb {

constexpr int kMaxPlayers = 5;
constexpr int kMaxColors = 5;
constexpr int kMaxRanks = 5;
constexpr int kMaxHand = 8;        // the reference's 8-bit hand masks (hanabi_state.cc:30,43)
constexpr int kMaxHandNarrow = 5;  // 32-bit hand words: 6-bit card slots, 5-bit knowledge slots x 5;
                                   // larger hands use the 64-bit words of the wide view (hb_engine.cuh)
constexpr int kMaxInfo = 31;     // 5-bit field
constexpr int kMaxLife = 15;     // 4-bit field
constexpr int kMtN = 624;
constexpr int kMtM = 397;
// Lazy slot regeneration (DESIGN.md "mt19937"): a deal inside an episode only READS its two
// RNG words; the slots are regenerated later, together and coalesced, by the job that
// re-deals the game (or by a flush job once `pend_cap - 1` pairs are pending).  The in-place
// scheme stays valid while fewer than 227 words are pending: at most 2 * 15 here.
constexpr int kMaxPend = 15;  // 4-bit field
#ifndef HB_LAZY_REGEN
#define HB_LAZY_REGEN 1
#endif
#ifndef HB_JOBS_MIN_CARDS
#define HB_JOBS_MIN_CARDS 12  // initial deals of this many cards or more use the job path (hb_engine.cuh)
#endif

// Observation types, hanabi_lib/hanabi_game.h:39.
enum ObsType { kMinimal = 0, kCardKnowledge = 1, kSeer = 2 };
// Last-action move kinds in the order the encoder one-hots them
// (hanabi_lib/canonical_encoders.cc:258-269), not the HanabiMove::Type enum order.
enum LaType { kLaPlay = 0, kLaDiscard = 1, kLaRevealColor = 2, kLaRevealRank = 3 };

// Packed record of a non-deal move (state column P+1 .z and the optional history ring):
// bit 0 valid, then type (LaType), acting player (absolute), target offset, hinted value,
// card index, colour and rank of the played/discarded card, reveal bitmask, scored,
// information token gained.
constexpr int kLaValid = 0, kLaTypeSh = 1, kLaActorSh = 3, kLaOffSh = 6, kLaValSh = 9,
              kLaIdxSh = 12, kLaColSh = 15, kLaRankSh = 18, kLaRevealSh = 21, kLaScoredSh = 26,
              kLaInfoSh = 27, kLaRevealHiSh = 28;
// The reveal bitmask of a hint: cards 0-4 in bits [21,26), cards 5-7 (hands of six to eight
// cards) in bits [28,31) -- records of hands up to five cards keep their 28-bit form.
HB_HD uint32_t la_pack_reveal(uint32_t mask) {
  return ((mask & 31u) << kLaRevealSh) | ((mask >> 5) << kLaRevealHiSh);
}
HB_HD uint32_t la_reveal(uint32_t la) {
  return ((la >> kLaRevealSh) & 31u) | (((la >> kLaRevealHiSh) & 7u) << 5);
}
// BatchExportGame: canonical dump (272 ints, oracle/oracle_api.h layout) followed by
// [272] turns_to_play [273] last-move record [274..277] the four records before it, newest
// first (0 unless BatchEnableHistory) [278] done flag [279] episode length
// [280..287] the discard pile in the order the cards were discarded, 5-bit card types
// (colour * R + rank) packed from bit 0 on [288] 1 when that log is valid (BatchEnableHistory
// was on for the whole episode), else 0 [289] spare
constexpr int kDumpLen = 272;
constexpr int kExportLen = 290;
constexpr int kDiscardLogWords = 8;  // 50 cards x 5 bits

struct Spec {
  int32_t P, C, R, H, IT, LT;
  int32_t obs_type, random_start;
  int32_t deck_max, per_color, bits_per_card;
  int32_t obs_size, max_moves;
  int32_t hands_len, board_len, discard_len, last_len, know_len;
  int32_t fast_reset;  // the unrolled re-deal applies (no random start, >= 2 card types left throughout)
  int32_t pend_cap;    // lazy regeneration: pending pairs a game may carry (0 = regenerate at once), PendCapFor()
  // Deck bitmap helpers (one bit per card instance, same layout as the discard
  // thermometers): first bit of every type block by block width.
  uint64_t full_deck, type_m23, type_m3, type_m1;
};

HB_HD int InstancesOfRank(int R, int rank) {  // hanabi_game.cc:126-136
  return rank == 0 ? 3 : (rank == R - 1 ? 1 : 2);
}
// Bit offset of rank r inside one colour's discard thermometer block.
HB_HD int DiscardRankOffset(int rank) { return rank == 0 ? 0 : 2 * rank + 1; }

// Pending pairs a game of this rule set may carry.  Games of small decks deal few cards per
// episode; a small cap keeps their re-deal jobs narrow (a job has P*H + pend_cap pair slots).
HB_HD constexpr int PendCapFor(int P, int H, int deck_max, int random_start) {
  if (!HB_LAZY_REGEN || P * H < HB_JOBS_MIN_CARDS || random_start || deck_max - P * H < 3 || P * H > 30) return 0;
  int n = (deck_max - P * H) / 3;
  n = n < 4 ? 4 : n;
  return n > kMaxPend ? kMaxPend : n;
}

// Fills the derived fields.  Returns 0 on success, else an error code whose
// text hb::SpecError() returns.
inline int FinishSpec(Spec* s) {
  if (s->P < 2 || s->P > kMaxPlayers) return 1;
  if (s->C < 1 || s->C > kMaxColors) return 2;
  if (s->R < 1 || s->R > kMaxRanks) return 3;
  if (s->H < 1 || s->H > kMaxHand) return 4;
  if (s->IT < 0 || s->IT > kMaxInfo) return 5;
  if (s->LT < 1 || s->LT > kMaxLife) return 6;
  if (s->obs_type < 0 || s->obs_type > 2) return 7;
  s->per_color = 0;
  for (int r = 0; r < s->R; ++r) s->per_color += InstancesOfRank(s->R, r);
  s->deck_max = s->per_color * s->C;
  if (s->H * s->P > s->deck_max) return 8;
  s->bits_per_card = s->C * s->R;
  s->hands_len = (s->P - 1) * s->H * s->bits_per_card + s->P;
  s->board_len = s->deck_max - s->P * s->H + s->C * s->R + s->IT + s->LT;
  s->discard_len = s->deck_max;
  s->last_len = s->P + 4 + s->P + s->C + s->R + s->H + s->H + s->bits_per_card + 2;
  s->know_len = s->P * s->H * (s->bits_per_card + s->C + s->R);
  s->obs_size = s->hands_len + s->board_len + s->discard_len + s->last_len +
                (s->obs_type == kMinimal ? 0 : s->know_len);
  s->max_moves = 2 * s->H + (s->P - 1) * s->C + (s->P - 1) * s->R;
  s->full_deck = s->type_m23 = s->type_m3 = s->type_m1 = 0;
  for (int c = 0; c < s->C; ++c)
    for (int r = 0; r < s->R; ++r) {
      const int n = InstancesOfRank(s->R, r), off = c * s->per_color + DiscardRankOffset(r);
      s->full_deck |= ((1ull << n) - 1ull) << off;
      if (n >= 2) s->type_m23 |= 1ull << off;
      if (n == 3) s->type_m3 |= 1ull << off;
      if (n == 1) s->type_m1 |= 1ull << off;
    }
  // A deck with more cards than the largest multiplicity (3) holds >= 2 types.
  s->fast_reset = (!s->random_start && s->deck_max - s->P * s->H >= 3 && s->P * s->H <= 30) ? 1 : 0;
  s->pend_cap = PendCapFor(s->P, s->H, s->deck_max, s->random_start);
  return 0;
}

inline const char* SpecError(int code) {
  switch (code) {
    case 0: return "ok";
    case 1: return "players must be in [2,5] (hanabi_game.cc:33)";
    case 2: return "colors must be in [1,5] (hanabi_game.cc:35)";
    case 3: return "ranks must be in [1,5] (hanabi_game.cc:37)";
    case 4: return "hand_size must be in [1,8] (hanabi_state.cc:30,43: 8-bit hand masks)";
    case 5: return "max_information_tokens must be in [0,31] for the packed batch layout";
    case 6: return "max_life_tokens must be in [1,15] for the packed batch layout";
    case 7: return "observation_type must be 0 (minimal), 1 (card knowledge) or 2 (seer)";
    case 8: return "hand_size * players exceeds the deck (hanabi_game.cc:58)";
    default: return "bad game parameters";
  }
}

}  // namespace hb
#endif  // HB_SPEC_H_
