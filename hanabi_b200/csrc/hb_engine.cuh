// hb_engine.cuh -- device-side batched Hanabi engine (sm_100a).
//
// One thread owns one game; a warp owns a tile of 32 games and provides the
// cooperative services (mt19937 block regeneration, bit->byte expansion of the
// observation tile with 128-bit stores).  The whole game lives in registers in
// packed form; see DESIGN.md "Data layout in HBM".
//
// Reference behaviour restated here (paths relative to
// /root/reference/hanabi_learning_environment/hanabi_lib/):
//   apply move / advance / score / terminal   hanabi_state.cc:104-144,221-275,359-377
//   legal moves                               hanabi_state.cc:146-219,288-304
//   hints on card knowledge                   hanabi_hand.cc:24-42,96-126
//   deals                                     hanabi_state.cc:229-241,277-286,313-325
//                                             hanabi_game.cc:106-112,138-145 (+ libstdc++ <random>)
//   observation view + canonical encoding     hanabi_observation.cc:34-93,
//                                             canonical_encoders.cc:62-451
#ifndef HB_ENGINE_CUH_
#define HB_ENGINE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hb_spec.h"

namespace hb {

#define HB_DEV __device__ __forceinline__
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ spec views
// Static view: every rule parameter that shapes the bit layout is a compile-time
// constant so the encoder unrolls into straight-line shifts on register words.
// Threads (= games) per CTA, measured per rule set with two launches in flight per GPU
// (768 resident threads per SM in every case): 256 where the re-deal is the parallel
// Lehmer decode (initial deals of 12+ cards), 192 for the short serial re-deals.
#ifndef HB_JOBS_MIN_CARDS
#define HB_JOBS_MIN_CARDS 12
#endif
#ifdef HB_BLOCK
constexpr int block_for_deal(int) { return HB_BLOCK; }
#else
constexpr int block_for_deal(int cards) { return cards >= 12 ? 256 : 192; }
#endif

template <int P_, int H_, int C_, int R_, int IT_, int LT_>
struct SView {
  static constexpr bool kStatic = true;
  static constexpr bool kWide = false;  // 32-bit hand words: up to five cards per hand
  typedef uint32_t Word;
  static constexpr int kP = P_, kH = H_, kC = C_, kR = R_;
  static constexpr int kBlock = block_for_deal(P_ * H_);
  static constexpr int kResident = 0;  // resident threads per SM the kernel is built for (0: the default)
  static constexpr int kStatePolicy = (P_ == 2 || H_ <= 2) ? 1 : 2;  // L2 policy of the packed state (st_state)
  static constexpr int kStreamOut = C_ * R_ * H_ >= 100 ? 1 : 0;  // streaming stores for long rows (store_out)
  // Re-deal machinery, measured per rule set (profiles/r02i_ab.log): initial deals of 12+ cards
  // go through the job path (one thread per RNG word pair over all jobs of the CTA, serial
  // chain per job, lazy slot regeneration); deals of up to 10 cards keep round 1's serial path.
  static constexpr bool kJobs = P_ * H_ >= HB_JOBS_MIN_CARDS;
  // pending-pair cap of this shape without a random start player (0: lazy regeneration off)
  static constexpr int kPendCap = kJobs ? PendCapFor(P_, H_, (R_ == 1 ? 3 : 2 * R_) * C_, 0) : 0;
  HB_DEV explicit SView(const Spec&) {}
  HB_DEV constexpr int P() const { return P_; }
  HB_DEV constexpr int H() const { return H_; }
  HB_DEV constexpr int C() const { return C_; }
  HB_DEV constexpr int R() const { return R_; }
  HB_DEV constexpr int IT() const { return IT_; }
  HB_DEV constexpr int LT() const { return LT_; }
  HB_DEV constexpr int per_color() const { return R_ == 1 ? 3 : 2 * R_; }
  HB_DEV constexpr int deck_max() const { return per_color() * C_; }
  HB_DEV constexpr int bpc() const { return C_ * R_; }
};
// Dynamic view: any legal parameter set, loop bounds are the layout maxima.
template <bool WIDE> struct WordOf { typedef uint32_t type; };
template <> struct WordOf<true> { typedef uint64_t type; };
template <bool WIDE>
struct DViewT {
  static constexpr bool kStatic = false;
  static constexpr bool kWide = WIDE;
  typedef typename WordOf<WIDE>::type Word;
  static constexpr int kP = kMaxPlayers, kH = WIDE ? kMaxHand : kMaxHandNarrow, kC = kMaxColors, kR = kMaxRanks;
  static constexpr int kBlock = WIDE ? 128 : block_for_deal(kMaxPlayers * kMaxHandNarrow);
  static constexpr int kResident = WIDE ? 384 : 0;  // the 64-bit hand words cost registers: fewer, fatter threads
  static constexpr int kStatePolicy = 0;
  static constexpr int kStreamOut = -1;  // decided by the rule set's row length at run time
  static constexpr bool kJobs = true;
  static constexpr int kPendCap = HB_LAZY_REGEN ? kMaxPend : 0;
  const Spec& s;
  HB_DEV explicit DViewT(const Spec& sp) : s(sp) {}
  HB_DEV int P() const { return s.P; }
  HB_DEV int H() const { return s.H; }
  HB_DEV int C() const { return s.C; }
  HB_DEV int R() const { return s.R; }
  HB_DEV int IT() const { return s.IT; }
  HB_DEV int LT() const { return s.LT; }
  HB_DEV int per_color() const { return s.per_color; }
  HB_DEV int deck_max() const { return s.deck_max; }
  HB_DEV int bpc() const { return s.bits_per_card; }
};
typedef DViewT<false> DView;
// Hands of six to eight cards (the reference allows hand_size * players <= deck,
// hanabi_game.cc:38,58; its 8-bit hand masks cap a hand at eight, hanabi_state.cc:30,43):
// the same engine on 64-bit hand words (6-bit card slots, 5-bit knowledge slots x 8) and a
// state layout of 2 columns per player (load_game).
typedef DViewT<true> WView;

// ------------------------------------------------------------------- the game
// Packed HBM layout: columns of uint4, column-major over games (state[col * N + game]) so a
// warp reads/writes 512 contiguous bytes per column.  General layout, P + 2 columns (the
// compact default, P + 1 columns, folds column P + 1 into the others: see load_game):
//   column p < P  : x = cards   (6-bit slots: colour | rank << 3, slot 0 = oldest card)
//                   y = colour-plausible masks (5-bit slots)
//                   z = rank-plausible masks   (5-bit slots)
//                   w = colour-hinted flags [0,8) | rank-hinted flags [8,16) | hand size [16,20)
//   column P      : x,y = deck bitmap (one bit per card instance still in the deck),
//                   z,w = discard thermometers; both in the bit layout of the encoder's
//                   discard section (colour-major, 3/2/.../1 bits per rank)
//   column P + 1  : x = fireworks 5x3 | info 5 | life 4 | cur 3 | turns_to_play 3 | done | error
//                   y = deck size 6 | mt19937 index 10 | episode length 12
//                   z = last non-deal action record, w = spare
template <class W>
struct GameT {
  W cards[kMaxPlayers];
  W cpl[kMaxPlayers];
  W rpl[kMaxPlayers];
  uint32_t hw[kMaxPlayers];
  uint64_t deck;  // bit set = that card instance is still in the deck
  uint64_t disc;  // discard thermometers
  uint32_t fw;    // 3 bits per colour
  int info, life, cur, turns, done, err;
  int deck_size, mti, eplen;
  int pend;  // RNG word pairs taken but not yet regenerated (they sit right below mti)
  uint32_t last;
  uint32_t spare;
};
typedef GameT<uint32_t> Game;
#define HB_GAME(V) GameT<typename V::Word>

// last-action record fields: kLa* in hb_spec.h

template <class W>
HB_DEV int hand_size(const GameT<W>& G, int p_static) { return (G.hw[p_static] >> 16) & 15; }
// one bit per 5-bit knowledge slot
template <class W> HB_DEV constexpr W slot_ones();
template <> HB_DEV constexpr uint32_t slot_ones<uint32_t>() { return 0x108421u; }
template <> HB_DEV constexpr uint64_t slot_ones<uint64_t>() { return 0x842108421ull; }

// The packed state is the one array the next step reads again.  Where it pays, its STORES carry
// the L2 evict-last policy, so that the streaming observation stores (evict-first) do not push it
// out of the 126 MB L2.  Measured per rule set on B200, 2^20 games (profiles/r02ae_exp_ab_evict.log,
// r02af_*, r02ag_*, r02aj_*, r02ak_*): Full 2p +2.2 % (also at 2^22 games, where the state is
// larger than the L2), Small 2p / 4p +2.1 % / +2.5 %, Full 4p -0.2...-1.5 %, Full 5p -1.3...-4 % --
// hence a property of the view (kStatePolicy).  The hint on the state's loads adds nothing (and
// costs 0.5-1 % next to the stores'); on the RNG words alone it gains the same 2.4 % on Full 2p
// and loses it together with the state (the L2 cannot keep both); on the action buffer it
// changes nothing.
// The rule sets whose state the L2 cannot keep anyway (Full 3p / 4p / 5p: 64-96 B per game) do
// best with STREAMING loads and stores of the state (ld.cs / st.cs: +0.5 % Full 4p, +0.9 % Full 5p,
// profiles/r02an_exp_ab_stcs.log).  kStatePolicy of the view: 0 plain, 1 keep (evict-last stores),
// 2 stream.  HB_STATE_POLICY >= 0 forces one policy everywhere.
#ifndef HB_STATE_POLICY
#define HB_STATE_POLICY -1
#endif
template <class V>
constexpr int state_policy_of() {
  return HB_STATE_POLICY < 0 ? V::kStatePolicy : HB_STATE_POLICY;
}
HB_DEV uint64_t state_policy() {
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int POLICY>
HB_DEV uint4 ld_state(const uint4* p) {
  if constexpr (POLICY == 2) return __ldcs(p);
  else return *p;
}
template <int POLICY>
HB_DEV void st_state(uint4* p, uint4 v) {
  if constexpr (POLICY == 1) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(state_policy())
                 : "memory");
  } else if constexpr (POLICY == 2) {
    __stcs(p, v);
  } else {
    *p = v;
  }
}

// Compact layout (HB_COMPACT_STATE, the default): P + 1 columns.  The fields of the general
// layout's column P + 1 travel in the unused high bits of players 0 and 1 and of the deck /
// discard column (cards 30 of 32 bits, plausible masks 25, flags + hand size 20, deck and
// discard bitmaps 50 of 64; the deck size is the bitmap's population count, the spare word
// is dropped): 16 bytes less per game and direction -- the step kernel's time follows its DRAM
// bytes.
//   column 0/1 : x = cards | bit 30: m[21] / L[21]      y = colour masks | m[0,7) / m[14,21) << 25
//                z = rank masks | m[7,14) / L[14,21) << 25   w = flags, hand size | X[0,12) / X[12,24) << 20
//   column P   : x = deck lo   y = deck hi (18 bits) | X[24,32) << 18 | L[22,28) << 26
//                z = discard lo   w = discard hi (18 bits) | L[0,14) << 18
// with X = the general layout's word x (fireworks ... done, error: the error flag is bit 25 of
// column P's y), m = mt19937 index | episode length (8 bits, saturating) << 10 | pending pairs << 18,
// L = last-action record.
#ifndef HB_COMPACT_STATE
#define HB_COMPACT_STATE 1
#endif
constexpr int state_columns(int players) { return HB_COMPACT_STATE ? players + 1 : players + 2; }
// Wide layout (WView, hands of 6-8 cards), 2 P + 2 columns:
//   column 2p     : x,y = cards (6-bit slots x 8)   z,w = colour-plausible masks (5-bit slots x 8)
//   column 2p + 1 : x,y = rank-plausible masks      z = flags | hand size (as above)   w = 0
//   column 2P     : deck bitmap, discard thermometers
//   column 2P + 1 : the general layout's column P + 1
constexpr int state_columns_wide(int players) { return 2 * players + 2; }

template <class W>
HB_DEV void unpack_misc(GameT<W>& G, uint32_t x) {
  G.fw = x & 0x7fffu;
  G.info = (x >> 15) & 31;
  G.life = (x >> 20) & 15;
  G.cur = (x >> 24) & 7;
  G.turns = (x >> 27) & 7;
  G.done = (x >> 30) & 1;
  G.err = (x >> 31) & 1;
}

template <class V>
HB_DEV void load_game(const uint4* __restrict__ st, int64_t N, int64_t g, V v, HB_GAME(V)& G) {
  if constexpr (V::kWide) {
#pragma unroll
    for (int p = 0; p < V::kP; ++p) {
      G.cards[p] = 0; G.cpl[p] = 0; G.rpl[p] = 0; G.hw[p] = 0;
      if (p < v.P()) {
        const uint4 a = ld_state<state_policy_of<V>()>(st + (int64_t)(2 * p) * N + g), b = ld_state<state_policy_of<V>()>(st + (int64_t)(2 * p + 1) * N + g);
        G.cards[p] = (uint64_t)a.x | ((uint64_t)a.y << 32);
        G.cpl[p] = (uint64_t)a.z | ((uint64_t)a.w << 32);
        G.rpl[p] = (uint64_t)b.x | ((uint64_t)b.y << 32);
        G.hw[p] = b.z;
      }
    }
    const uint4 a = ld_state<state_policy_of<V>()>(st + (int64_t)(2 * v.P()) * N + g);
    G.deck = (uint64_t)a.x | ((uint64_t)a.y << 32);
    G.disc = (uint64_t)a.z | ((uint64_t)a.w << 32);
    const uint4 b = ld_state<state_policy_of<V>()>(st + (int64_t)(2 * v.P() + 1) * N + g);
    unpack_misc(G, b.x);
    G.deck_size = b.y & 63;
    G.mti = (b.y >> 6) & 1023;
    G.eplen = (b.y >> 16) & 255;
    G.pend = (b.y >> 28) & 15;
    G.last = b.z;
    G.spare = b.w;
    return;
  } else {
  if (HB_COMPACT_STATE) {
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
#pragma unroll
    for (int p = 0; p < V::kP; ++p) {
      uint4 c = make_uint4(0, 0, 0, 0);
      if (p < v.P()) c = ld_state<state_policy_of<V>()>(st + (int64_t)p * N + g);
      if (p == 0) c0 = c;
      if (p == 1) c1 = c;
      G.cards[p] = c.x & 0x3fffffffu; G.cpl[p] = c.y & 0x1ffffffu; G.rpl[p] = c.z & 0x1ffffffu; G.hw[p] = c.w & 0xfffffu;
    }
    const uint4 c2 = ld_state<state_policy_of<V>()>(st + (int64_t)v.P() * N + g);
    G.deck = (uint64_t)c2.x | ((uint64_t)(c2.y & 0x3ffffu) << 32);
    G.disc = (uint64_t)c2.z | ((uint64_t)(c2.w & 0x3ffffu) << 32);
    unpack_misc(G, (c0.w >> 20) | ((c1.w >> 20) << 12) | (((c2.y >> 18) & 0xffu) << 24));
    const uint32_t m = (c0.y >> 25) | ((c0.z >> 25) << 7) | ((c1.y >> 25) << 14) | (((c0.x >> 30) & 1u) << 21);
    G.deck_size = __popcll(G.deck);
    G.mti = (int)(m & 1023u);
    G.eplen = (int)((m >> 10) & 255u);
    G.pend = V::kPendCap > 0 ? (int)(m >> 18) : 0;  // (views without lazy regeneration never have pending pairs)
    G.last = (c2.w >> 18) | ((c1.z >> 25) << 14) | (((c1.x >> 30) & 1u) << 21) | ((c2.y >> 26) << 22);
    G.spare = 0;
    return;
  }
#pragma unroll
  for (int p = 0; p < V::kP; ++p) {
    if (p < v.P()) {
      uint4 c = ld_state<state_policy_of<V>()>(st + (int64_t)p * N + g);
      G.cards[p] = c.x; G.cpl[p] = c.y; G.rpl[p] = c.z; G.hw[p] = c.w;
    } else {
      G.cards[p] = 0; G.cpl[p] = 0; G.rpl[p] = 0; G.hw[p] = 0;
    }
  }
  uint4 a = ld_state<state_policy_of<V>()>(st + (int64_t)v.P() * N + g);
  G.deck = (uint64_t)a.x | ((uint64_t)a.y << 32);
  G.disc = (uint64_t)a.z | ((uint64_t)a.w << 32);
  uint4 b = ld_state<state_policy_of<V>()>(st + (int64_t)(v.P() + 1) * N + g);
  G.fw = b.x & 0x7fffu;
  G.info = (b.x >> 15) & 31;
  G.life = (b.x >> 20) & 15;
  G.cur = (b.x >> 24) & 7;
  G.turns = (b.x >> 27) & 7;
  G.done = (b.x >> 30) & 1;
  G.err = (b.x >> 31) & 1;
  G.deck_size = b.y & 63;
  G.mti = (b.y >> 6) & 1023;
  G.eplen = (b.y >> 16) & 255;
  G.pend = (b.y >> 28) & 15;
  G.last = b.z;
  G.spare = b.w;
  }
}

template <class V>
HB_DEV void store_game(uint4* __restrict__ st, int64_t N, int64_t g, V v, const HB_GAME(V)& G) {
  if constexpr (V::kWide) {
#pragma unroll
    for (int p = 0; p < V::kP; ++p) {
      if (p < v.P()) {
        st_state<state_policy_of<V>()>(st + (int64_t)(2 * p) * N + g, make_uint4((uint32_t)G.cards[p], (uint32_t)(G.cards[p] >> 32),
                                                           (uint32_t)G.cpl[p], (uint32_t)(G.cpl[p] >> 32)));
        st_state<state_policy_of<V>()>(st + (int64_t)(2 * p + 1) * N + g, make_uint4((uint32_t)G.rpl[p], (uint32_t)(G.rpl[p] >> 32), G.hw[p], 0u));
      }
    }
    st_state<state_policy_of<V>()>(st + (int64_t)(2 * v.P()) * N + g, make_uint4((uint32_t)G.deck, (uint32_t)(G.deck >> 32),
                                                           (uint32_t)G.disc, (uint32_t)(G.disc >> 32)));
    const uint32_t x = G.fw | ((uint32_t)G.info << 15) | ((uint32_t)G.life << 20) |
                       ((uint32_t)G.cur << 24) | ((uint32_t)G.turns << 27) | ((uint32_t)G.done << 30) |
                       ((uint32_t)G.err << 31);
    const uint32_t y = (uint32_t)G.deck_size | ((uint32_t)G.mti << 6) |
                       ((uint32_t)(G.eplen > 255 ? 255 : G.eplen) << 16) | ((uint32_t)G.pend << 28);
    st_state<state_policy_of<V>()>(st + (int64_t)(2 * v.P() + 1) * N + g, make_uint4(x, y, G.last, G.spare));
    return;
  } else {
  if (HB_COMPACT_STATE) {
    const uint32_t X = G.fw | ((uint32_t)G.info << 15) | ((uint32_t)G.life << 20) |
                       ((uint32_t)G.cur << 24) | ((uint32_t)G.turns << 27) | ((uint32_t)G.done << 30) |
                       ((uint32_t)G.err << 31);
    const uint32_t m = (uint32_t)G.mti | ((uint32_t)(G.eplen > 255 ? 255 : G.eplen) << 10) |
                       (V::kPendCap > 0 ? (uint32_t)G.pend << 18 : 0u);
    const uint32_t L = G.last & 0xfffffffu;
#pragma unroll
    for (int p = 0; p < V::kP; ++p) {
      if (p < v.P()) {
        uint4 c = make_uint4(G.cards[p] & 0x3fffffffu, G.cpl[p] & 0x1ffffffu, G.rpl[p] & 0x1ffffffu, G.hw[p] & 0xfffffu);
        if (p == 0) {
          c.x |= ((m >> 21) & 1u) << 30; c.y |= (m & 0x7fu) << 25; c.z |= ((m >> 7) & 0x7fu) << 25; c.w |= (X & 0xfffu) << 20;
        }
        if (p == 1) {
          c.x |= ((L >> 21) & 1u) << 30; c.y |= ((m >> 14) & 0x7fu) << 25; c.z |= ((L >> 14) & 0x7fu) << 25; c.w |= ((X >> 12) & 0xfffu) << 20;
        }
        st_state<state_policy_of<V>()>(st + (int64_t)p * N + g, c);
      }
    }
    const uint32_t dh = (uint32_t)(G.deck >> 32) & 0x3ffffu, sh = (uint32_t)(G.disc >> 32) & 0x3ffffu;
    st_state<state_policy_of<V>()>(st + (int64_t)v.P() * N + g,
             make_uint4((uint32_t)G.deck, dh | (((X >> 24) & 0xffu) << 18) | (((L >> 22) & 0x3fu) << 26),
                        (uint32_t)G.disc, sh | ((L & 0x3fffu) << 18)));
    return;
  }
#pragma unroll
  for (int p = 0; p < V::kP; ++p) {
    if (p < v.P()) st_state<state_policy_of<V>()>(st + (int64_t)p * N + g, make_uint4(G.cards[p], G.cpl[p], G.rpl[p], G.hw[p]));
  }
  st_state<state_policy_of<V>()>(st + (int64_t)v.P() * N + g, make_uint4((uint32_t)G.deck, (uint32_t)(G.deck >> 32),
                                                   (uint32_t)G.disc, (uint32_t)(G.disc >> 32)));
  uint32_t x = G.fw | ((uint32_t)G.info << 15) | ((uint32_t)G.life << 20) |
               ((uint32_t)G.cur << 24) | ((uint32_t)G.turns << 27) | ((uint32_t)G.done << 30) |
               ((uint32_t)G.err << 31);
  uint32_t y = (uint32_t)G.deck_size | ((uint32_t)G.mti << 6) |
               ((uint32_t)(G.eplen > 255 ? 255 : G.eplen) << 16) | ((uint32_t)G.pend << 28);
  st_state<state_policy_of<V>()>(st + (int64_t)(v.P() + 1) * N + g, make_uint4(x, y, G.last, G.spare));
  }
}

// Register-array access with a runtime player index: a predicated select chain
// keeps the arrays in registers (no local memory).
template <class V, class T>
HB_DEV T sel(const T (&a)[kMaxPlayers], int p) {
  T r = a[0];
#pragma unroll
  for (int i = 1; i < V::kP; ++i) r = (p == i) ? a[i] : r;
  return r;
}
template <class V, class T>
HB_DEV void put(T (&a)[kMaxPlayers], int p, T val) {
#pragma unroll
  for (int i = 0; i < V::kP; ++i) a[i] = (p == i) ? val : a[i];
}

// -------------------------------------------------------------------- mt19937
HB_DEV uint32_t mt_temper(uint32_t z) {
  z ^= (z >> 11);
  z ^= (z << 7) & 0x9d2c5680u;
  z ^= (z << 15) & 0xefc60000u;
  z ^= (z >> 18);
  return z;
}

// mt19937 without the block "twist": every game keeps its 624-word block in HBM
// and an index k.  Invariant: slots < k already hold the NEXT block's words, slots
// >= k the current block's.  Taking word k immediately regenerates slot k for the
// next block, x'[k] = x[k+397] ^ ((x[k] & UPPER | x[k+1] & LOWER) >> 1) ^ mag,
// which reads exactly the values the sequential block regeneration of
// std::mt19937 would read at that point (x[k+1] still old; x[k+397] old for
// k < 227, already new for k >= 227; slot 623 sees the new slot 0).  The output
// stream is therefore identical, no lane ever stalls on a 624-word regeneration,
// and a draw costs five word loads and two word stores local to the thread.
HB_DEV int mt_wrap(int i) { return i >= kMtN ? i - kMtN : i; }

HB_DEV uint32_t mt_next_block_word(uint32_t cur, uint32_t next, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (next & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// Raw words needed to take two consecutive outputs at index `at`: x[at], x[at+1],
// x[at+2], x[at+397], x[at+398] (indices mod 624).  Loaded together with the game
// state so that the deal that follows most moves does not wait on a second
// dependent global load.  at < 0 = nothing prefetched.
struct Prefetch {
  uint32_t a, b, c, p, q;
  int at;
};

HB_DEV void mt_load5(const uint32_t* x, int k, Prefetch& f) {
  const int k1 = mt_wrap(k + 1), k2 = mt_wrap(k + 2), m0 = mt_wrap(k + kMtM), m1 = mt_wrap(k + kMtM + 1);
  f.a = __ldcg(x + k);
  f.b = __ldcg(x + k1);
  f.c = __ldcg(x + k2);
  f.p = __ldcg(x + m0);
  f.q = __ldcg(x + m1);
  f.at = k;
}

// Takes outputs k and k+1 of the game's stream (tempered) and regenerates both slots.
HB_DEV void mt_take2(uint32_t* x, int& k, uint32_t& w0, uint32_t& w1, Prefetch& pre) {
  // `pre` by reference (never by address) so that it stays in registers
  Prefetch f;
  if (pre.at == k) { f = pre; pre.at = -1; }
  else mt_load5(x, k, f);
  w0 = mt_temper(f.a);
  w1 = mt_temper(f.b);
  const int k1 = mt_wrap(k + 1);
  __stcg(x + k, mt_next_block_word(f.a, f.b, f.p));
  __stcg(x + k1, mt_next_block_word(f.b, f.c, f.q));
  k = mt_wrap(k + 2);
}

// L2 prefetch of the words the initial deal of a new game will take from index k
// on (2 * pairs outputs): x[k .. k + 2*pairs] and x[k+397 .. k+397 + 2*pairs - 1],
// indices mod 624.  Issued the moment a game is known to end, so that the re-deal's
// loads find the lines in L2.
HB_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
HB_DEV void mt_prefetch_deal(const uint32_t* x, int k, int pairs) {
  const int n = 2 * pairs;
  prefetch_l2(x + k);
  const int e = k + n;
  prefetch_l2(x + (e < kMtN ? e : kMtN - 1));
  if (e >= kMtN) prefetch_l2(x);
  const int m = mt_wrap(k + kMtM), me = m + n - 1;
  prefetch_l2(x + m);
  prefetch_l2(x + (me < kMtN ? me : kMtN - 1));
  if (me >= kMtN) prefetch_l2(x);
}

// L2 prefetch of `count` consecutive slots from `start` on (indices mod 624), one prefetch
// per 128 bytes.
HB_DEV void mt_prefetch_span(const uint32_t* x, int start, int count) {
  if (count <= 0) return;
  for (int o = 0; o < count; o += 32) prefetch_l2(x + mt_wrap(start + o));
  prefetch_l2(x + mt_wrap(start + count - 1));
}

// Takes one output (random start player draw).
HB_DEV uint32_t mt_take1(uint32_t* x, int& k) {
  const uint32_t a = __ldcg(x + k), b = __ldcg(x + mt_wrap(k + 1)), p = __ldcg(x + mt_wrap(k + kMtM));
  __stcg(x + k, mt_next_block_word(a, b, p));
  k = mt_wrap(k + 1);
  return mt_temper(a);
}

// ------------------------------------------------------- lazy regeneration
// A take that only reads: outputs k and k+1 (tempered), the two slots stay as they are and
// count as pending (G.pend) until a job or flush_pending regenerates them, in stream order.
HB_DEV void mt_load2(const uint32_t* x, int k, Prefetch& f) {
  f.a = __ldcg(x + k);
  f.b = __ldcg(x + mt_wrap(k + 1));
  f.at = k;
}
HB_DEV void mt_take2_lazy(const uint32_t* x, int& k, uint32_t& w0, uint32_t& w1, Prefetch& pre) {
  Prefetch f;
  if (pre.at == k) { f = pre; pre.at = -1; }
  else mt_load2(x, k, f);
  w0 = mt_temper(f.a);
  w1 = mt_temper(f.b);
  k = mt_wrap(k + 2);
}

// Regenerates the pending slots one after the other (rare paths only: state injection,
// job-queue overflow, draws that need the exact sampler); afterwards every slot below mti
// is regenerated.
HB_DEV void flush_pending(uint32_t* x, int mti, int& pend) {
  int at = mti - 2 * pend;
  at += at < 0 ? kMtN : 0;
  for (int i = 0; i < pend; ++i) {
    Prefetch f;
    mt_load5(x, at, f);
    __stcg(x + at, mt_next_block_word(f.a, f.b, f.p));
    __stcg(x + mt_wrap(at + 1), mt_next_block_word(f.b, f.c, f.q));
    at = mt_wrap(at + 2);
  }
  pend = 0;
}

// ------------------------------------------------------------- deal sampling
// The reference draws a card with libstdc++'s std::discrete_distribution over
// doubles count/deck_size (hanabi_state.cc:277-280,313-325; hanabi_game.cc:106-112;
// bits/random.tcc:2655-2712,3346-3385).  Mathematically that is "card number
// floor(u * deck_size) of the type-sorted deck".  The deck is kept as a bitmap
// with one bit per card instance in type order, so that card is simply the
// q-th set bit.  pick_fast evaluates q in exact 128-bit integer arithmetic and
// reports `ambiguous` when u*deck_size lies within 2^-32 of an integer -- the
// only place where fp64 rounding in the reference could decide differently (its
// accumulated error is < 2^-46).  Ambiguous draws are re-done by pick_exact, an
// operation-for-operation fp64 restatement of the library code.
constexpr uint64_t kGuard = 1ull << 32;

// Position of the q-th (0-based) set bit of x; q < popcount(x).
HB_DEV int select64(uint64_t x, int q) {
  uint32_t w = (uint32_t)x;
  int pos = 0;
  int c = __popc(w);
  if (q >= c) { q -= c; w = (uint32_t)(x >> 32); pos = 32; }
  c = __popc(w & 0xffffu);
  if (q >= c) { q -= c; w >>= 16; pos += 16; }
  c = __popc(w & 0xffu);
  if (q >= c) { q -= c; w >>= 8; pos += 8; }
  c = __popc(w & 0xfu);
  if (q >= c) { q -= c; w >>= 4; pos += 4; }
  c = __popc(w & 0x3u);
  if (q >= c) { q -= c; w >>= 2; pos += 2; }
  pos += (q >= (int)(w & 1u)) ? 1 : 0;
  return pos;
}

// Deck bit position -> (colour, rank).
template <class V>
HB_DEV void card_of_position(V v, int pos, int& color, int& rank) {
  color = pos / v.per_color();
  const int o = pos - color * v.per_color();
  rank = o < 3 ? 0 : (o - 1) >> 1;
}

// floor(u * deck_size) for the 64-bit uniform (hi:lo), and whether it is too
// close to an integer for the integer result to be trusted.
HB_DEV int draw_index(int deck_size, uint32_t lo, uint32_t hi, bool& ambiguous) {
  const uint64_t U = ((uint64_t)hi << 32) | lo;
  const uint64_t q = __umul64hi(U, (uint64_t)deck_size);
  const uint64_t F = U * (uint64_t)deck_size;
  ambiguous = (F + kGuard) < 2 * kGuard;
  return (int)q;
}

// Returns the deck bit position of the drawn card.
HB_DEV int pick_fast(uint64_t deck, int deck_size, uint32_t lo, uint32_t hi, bool& ambiguous) {
  return select64(deck, draw_index(deck_size, lo, hi, ambiguous));
}

template <class V>
__device__ __noinline__ int pick_exact(V v, uint64_t deck, int deck_size, uint32_t lo, uint32_t hi) {
  double w[kMaxColors * kMaxRanks];
  int first_bit[kMaxColors * kMaxRanks];
  int n = 0;
  for (int c = 0; c < v.C(); ++c)
    for (int r = 0; r < v.R(); ++r) {
      const int off = c * v.per_color() + DiscardRankOffset(r);
      const uint64_t block = (deck >> off) & ((1ull << InstancesOfRank(v.R(), r)) - 1ull);
      const int cnt = __popcll(block);
      if (cnt > 0) {
        first_bit[n] = off + __ffsll((long long)block) - 1;
        w[n] = __ddiv_rn((double)cnt, (double)deck_size);
        ++n;
      }
    }
  if (n < 2) return first_bit[0];
  double sum = 0.0;
  for (int k = 0; k < n; ++k) sum = __dadd_rn(sum, w[k]);
  double cp = 0.0;
  double u = __dadd_rn(__dmul_rn((double)lo, 1.0), __dmul_rn((double)hi, 4294967296.0));
  u = __ddiv_rn(u, 18446744073709551616.0);
  if (u >= 1.0) u = __longlong_as_double(0x3FEFFFFFFFFFFFFFll);  // nextafter(1, 0)
  int res = n - 1;
  bool found = false;
  for (int k = 0; k < n; ++k) {
    const double p = __ddiv_rn(w[k], sum);
    cp = (k == 0) ? p : __dadd_rn(cp, p);
    const double c = (k == n - 1) ? 1.0 : cp;
    if (!found && !(c < u)) { res = k; found = true; }
  }
  return first_bit[res];
}

// Number of distinct card types left in the deck (>= 2 means the reference
// consumes two RNG words for the draw, else none: random.tcc:2660-2664,2696-2697).
HB_DEV int deck_types(const Spec& spec, uint64_t deck) {
  const uint64_t any2 = deck | (deck >> 1);
  const uint64_t nz = (any2 & spec.type_m23) | ((deck >> 2) & spec.type_m3) | (deck & spec.type_m1);
  return __popcll(nz);
}

// Lemire's nearly-divisionless bounded integer as libstdc++ uses it for
// uniform_int_distribution on a 32-bit URNG (bits/uniform_int_dist.h:256-282).
// Returns true when the word is accepted.
HB_DEV bool lemire_accept(uint32_t word, uint32_t range, uint32_t& out) {
  const uint64_t product = (uint64_t)word * (uint64_t)range;
  const uint32_t low = (uint32_t)product;
  out = (uint32_t)(product >> 32);
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    if (low < threshold) return false;
  }
  return true;
}

// -------------------------------------------------------------------- rules
template <class V>
HB_DEV bool needs_deal(V v, const HB_GAME(V)& G) {
  bool short_hand = false;
#pragma unroll
  for (int p = 0; p < V::kP; ++p)
    if (p < v.P()) short_hand |= hand_size(G, p) < v.H();
  return G.deck_size > 0 && short_hand;
}

// HanabiState::ApplyMove(kDeal) of the card at deck position `pos`
// (hanabi_state.cc:229-241).
template <class V>
HB_DEV void deal_card(V v, HB_GAME(V)& G, int pos, int obs_type) {
  int color, rank;
  card_of_position(v, pos, color, rank);
  bool placed = false;
#pragma unroll
  for (int p = 0; p < V::kP; ++p) {
    if (p < v.P()) {
      const int n = hand_size(G, p);
      if (!placed && n < v.H()) {  // PlayerToDeal: lowest index with a short hand
        placed = true;
        typedef typename V::Word W;
        G.cards[p] |= (W)(uint32_t)(color | (rank << 3)) << (6 * n);
        uint32_t cm = (1u << v.C()) - 1u, rm = (1u << v.R()) - 1u, flags = 0;
        if (obs_type == kSeer) {  // hanabi_state.cc:233-236
          cm = 1u << color; rm = 1u << rank; flags = (1u << n) | (1u << (8 + n));
        }
        G.cpl[p] |= (W)cm << (5 * n);
        G.rpl[p] |= (W)rm << (5 * n);
        G.hw[p] = (G.hw[p] | flags) + (1u << 16);
      }
    }
  }
  G.deck &= ~(1ull << pos);
  G.deck_size -= 1;
}

template <class V>
HB_DEV int score_of(V v, const HB_GAME(V)& G) {  // hanabi_state.cc:359-364
  if (G.life <= 0) return 0;
  int s = 0;
#pragma unroll
  for (int c = 0; c < V::kC; ++c)
    if (c < v.C()) s += (G.fw >> (3 * c)) & 7;
  return s;
}

template <class V>
HB_DEV bool is_terminal(V v, const HB_GAME(V)& G) {  // hanabi_state.cc:366-377
  return G.life < 1 || score_of(v, G) >= v.C() * v.R() || G.turns <= 0;
}

// HanabiState::LegalMoves(cur) as a bitmask over move uids
// (hanabi_state.cc:166-219,288-304; uid layout hanabi_game.cc:79-95).
template <class V>
HB_DEV uint64_t legal_bits(V v, const HB_GAME(V)& G) {
  const int n = (int)((sel<V>(G.hw, G.cur) >> 16) & 15);
  const uint64_t hm = (1ull << n) - 1ull;
  uint64_t bits = (G.info < v.IT() ? hm : 0ull) | (hm << v.H());
  if (G.info > 0) {
#pragma unroll
    for (int p = 0; p < V::kP; ++p) {
      if (p < v.P() && p != G.cur) {
        int off = p - G.cur;
        off += off < 0 ? v.P() : 0;
        const int m = hand_size(G, p);
        uint32_t cm = 0, rm = 0;
#pragma unroll
        for (int i = 0; i < V::kH; ++i) {
          if (i < m) {
            const uint32_t f = (uint32_t)(G.cards[p] >> (6 * i));
            cm |= 1u << (f & 7);
            rm |= 1u << ((f >> 3) & 7);
          }
        }
        bits |= (uint64_t)cm << (2 * v.H() + (off - 1) * v.C());
        bits |= (uint64_t)rm << (2 * v.H() + (v.P() - 1) * v.C() + (off - 1) * v.R());
      }
    }
  }
  return bits;
}

// HanabiState::MoveIsLegal for one move uid of the player to act
// (hanabi_state.cc:166-219); cheaper than building the whole mask.
template <class V>
HB_DEV bool move_is_legal(V v, const HB_GAME(V)& G, int uid, int max_moves) {
  if (uid < 0 || uid >= max_moves) return false;
  const int P = v.P(), H = v.H(), C = v.C(), R = v.R();
  const int n = (int)((sel<V>(G.hw, G.cur) >> 16) & 15);
  if (uid < H) return G.info < v.IT() && uid < n;
  if (uid < 2 * H) return uid - H < n;
  if (G.info <= 0) return false;
  int u = uid - 2 * H;
  const bool is_rank = u >= (P - 1) * C;
  if (is_rank) u -= (P - 1) * C;
  const int span = is_rank ? R : C;
  // (two divisions by the view's constants instead of one by their runtime selection)
  const int offm1 = is_rank ? u / R : u / C, val = u - offm1 * span;
  int t = G.cur + offm1 + 1;
  t -= t >= P ? P : 0;
  const typename V::Word hand = sel<V>(G.cards, t);
  const int m = (int)((sel<V>(G.hw, t) >> 16) & 15);
  bool any = false;
#pragma unroll
  for (int i = 0; i < V::kH; ++i) {
    const uint32_t f = (uint32_t)(hand >> (6 * i));
    any |= (i < m) && ((int)(is_rank ? ((f >> 3) & 7) : (f & 7)) == val);
  }
  return any;
}

template <class T>
HB_DEV T drop_slot(T w, int idx, int width) {
  const T low = ((T)1 << (width * idx)) - (T)1;
  return (w & low) | ((w >> width) & ~low);
}

// HanabiState::ApplyMove for a player move (hanabi_state.cc:221-275), followed by
// the turn hand-over of AdvanceToNextPlayer (:104-111; pending deals are resolved
// by the caller's deal loop before anybody observes the state).
template <class V>
HB_DEV void apply_move(V v, HB_GAME(V)& G, int uid) {
  const int cur = G.cur;
  const int P = v.P(), H = v.H(), C = v.C(), R = v.R();
  if (G.deck_size == 0) G.turns -= 1;
  uint32_t la = 1u | ((uint32_t)cur << kLaActorSh);
  if (uid < 2 * H) {
    const bool is_play = uid >= H;
    const int idx = is_play ? uid - H : uid;
    const typename V::Word hand = sel<V>(G.cards, cur);
    const int color = (int)(hand >> (6 * idx)) & 7, rank = (int)(hand >> (6 * idx + 3)) & 7;
    bool to_discard = true, scored = false, token = false;
    if (is_play) {
      const int top = (G.fw >> (3 * color)) & 7;
      if (rank == top) {  // AddToFireworks :132-144
        G.fw += 1u << (3 * color);
        scored = true;
        to_discard = false;
        if (top + 1 == R && G.info < v.IT()) { G.info += 1; token = true; }
      } else {
        G.life -= 1;
      }
    } else {
      if (G.info < v.IT()) { G.info += 1; token = true; }  // IncrementInformationTokens
    }
    if (to_discard) {
      const int off = color * v.per_color() + DiscardRankOffset(rank);
      const int cnt = __popcll((G.disc >> off) & 7ull & ((1ull << InstancesOfRank(R, rank)) - 1ull));
      G.disc |= 1ull << (off + cnt);
    }
    // RemoveFromHand (hanabi_hand.cc:87-94): later cards shift down.
    put<V>(G.cards, cur, drop_slot(hand, idx, 6));
    put<V>(G.cpl, cur, drop_slot(sel<V>(G.cpl, cur), idx, 5));
    put<V>(G.rpl, cur, drop_slot(sel<V>(G.rpl, cur), idx, 5));
    const uint32_t hw = sel<V>(G.hw, cur);
    const uint32_t ch = drop_slot<uint32_t>(hw & 0xffu, idx, 1), rh = drop_slot<uint32_t>((hw >> 8) & 0xffu, idx, 1);
    put<V>(G.hw, cur, ch | (rh << 8) | ((hw & 0xf0000u) - (1u << 16)));
    la |= (uint32_t)(is_play ? kLaPlay : kLaDiscard) << kLaTypeSh;
    la |= (uint32_t)idx << kLaIdxSh | (uint32_t)color << kLaColSh | (uint32_t)rank << kLaRankSh;
    la |= (uint32_t)scored << kLaScoredSh | (uint32_t)token << kLaInfoSh;
  } else {
    int u = uid - 2 * H;
    const bool is_rank = u >= (P - 1) * C;
    if (is_rank) u -= (P - 1) * C;
    const int span = is_rank ? R : C;
    const int offm1 = is_rank ? u / R : u / C, val = u - offm1 * span;
    int t = cur + offm1 + 1;
    t -= t >= P ? P : 0;
    G.info -= 1;
    typedef typename V::Word W;
    const W hand = sel<V>(G.cards, t);
    const uint32_t hw = sel<V>(G.hw, t);
    const int n = (hw >> 16) & 15;
    uint32_t mm = 0;
    W M = 0;
#pragma unroll
    for (int i = 0; i < V::kH; ++i) {
      if (i < n) {
        const uint32_t f = (uint32_t)(hand >> (6 * i));
        const int cv = is_rank ? ((f >> 3) & 7) : (f & 7);
        if (cv == val) { mm |= 1u << i; M |= (W)31u << (5 * i); }
      }
    }
    // RevealColor/RevealRank (hanabi_hand.cc:96-126): matching cards become
    // exactly {val} and are flagged hinted; the others lose val.
    const W all = (W)(1u << val) * slot_ones<W>();
    if (is_rank) {
      put<V>(G.rpl, t, (sel<V>(G.rpl, t) & ~all & ~M) | (all & M));
      put<V>(G.hw, t, hw | (mm << 8));
    } else {
      put<V>(G.cpl, t, (sel<V>(G.cpl, t) & ~all & ~M) | (all & M));
      put<V>(G.hw, t, hw | mm);
    }
    la |= (uint32_t)(is_rank ? kLaRevealRank : kLaRevealColor) << kLaTypeSh;
    la |= (uint32_t)(offm1 + 1) << kLaOffSh | (uint32_t)val << kLaValSh |
          (V::kWide ? la_pack_reveal(mm) : mm << kLaRevealSh);
  }
  G.last = la;
  G.cur = cur + 1 == P ? 0 : cur + 1;
  G.eplen += 1;
}

// Fresh game before any deal (HanabiState ctor, hanabi_state.cc:90-102).
template <class V>
HB_DEV void init_fresh(V v, const Spec& spec, HB_GAME(V)& G, int start_player) {
#pragma unroll
  for (int p = 0; p < V::kP; ++p) { G.cards[p] = 0; G.cpl[p] = 0; G.rpl[p] = 0; G.hw[p] = 0; }
  G.deck = spec.full_deck;
  G.disc = 0;
  G.fw = 0;
  G.info = v.IT();
  G.life = v.LT();
  G.cur = start_player;
  G.turns = v.P();
  G.done = 0;
  G.deck_size = v.deck_max();
  G.eplen = 0;
  G.last = 0;
}

// Deals until the game is not waiting for a card any more.  With `restart` the
// deals of the old game are finished first (the reference deals before it looks
// at IsTerminal, rl_env.py:357-361), then a new game starts on the same RNG
// stream (one std::mt19937 per HanabiGame, hanabi_game.h:114) and gets its P*H
// cards.  Purely per-thread.
// `lazy_cap` > 0: takes only read their words (pending regeneration, see above) while the
// game carries fewer than lazy_cap pending pairs; a restart always regenerates at once
// (callers flush first).
template <class V>
HB_DEV void deal_loop(V v, const Spec& spec, HB_GAME(V)& G, bool restart, int64_t g, uint32_t* mt_all,
                      int sampler_mode, Prefetch& pre, int lazy_cap = 0) {
  uint32_t* x = mt_all + g * (int64_t)kMtN;
  for (;;) {
    bool need = needs_deal(v, G);
    if (!need && restart) {
      restart = false;
      init_fresh(v, spec, G, 0);
      if (spec.random_start) {
        // one bounded draw before the first deal (hanabi_game.cc:138-145),
        // re-drawn on the (2^-30 rare) Lemire rejection
        uint32_t val;
        while (!lemire_accept(mt_take1(x, G.mti), (uint32_t)v.P(), val)) {}
        G.cur = (int)val;
      }
      need = needs_deal(v, G);
    }
    if (!need) break;
    int pos;
    if (deck_types(spec, G.deck) >= 2) {
      uint32_t lo, hi;
      if (V::kPendCap > 0 && G.pend < lazy_cap) {
        mt_take2_lazy(x, G.mti, lo, hi, pre);
        G.pend += 1;
      } else {
        if (V::kPendCap > 0 && G.pend > 0) flush_pending(x, G.mti, G.pend);  // slots regenerate in stream order
        mt_take2(x, G.mti, lo, hi, pre);
      }
      bool amb = false;
      pos = pick_fast(G.deck, G.deck_size, lo, hi, amb);
      if (sampler_mode == 1 || (amb && sampler_mode != 2))
        pos = pick_exact(v, G.deck, G.deck_size, lo, hi);
    } else {
      pos = __ffsll((long long)G.deck) - 1;  // one type left: no RNG words are consumed
    }
    deal_card(v, G, pos, spec.obs_type);
  }
}

// ---- re-deal / flush jobs, compacted over the CTA ----------------------------
// A finished game that qualifies (spec.fast_reset: no random start player, every draw of the
// initial deal is a >= 2-type draw) and a running game whose pending pairs reach the cap
// enqueue a JOB in shared memory; after a CTA barrier all threads work the jobs off, one
// thread per RNG word pair (k_main, phase B).  The result of a re-deal job:
template <class W>
struct JobResultT {
  W cards[kMaxPlayers];         // the new hands
  uint32_t deck_lo, deck_hi;    // deck bitmap after the initial deal
};
typedef JobResultT<uint32_t> JobResult;
#define HB_JOBRESULT(V) JobResultT<typename V::Word>

// Draws the initial deal of one game.  Draw r (deck size deck_max - r) comes with
// its precomputed index qa[r] = floor(u_r * size) | ambiguous << 8 and its tempered
// RNG words lo[r], hi[r] (only read for the rare exact re-draw), all in shared
// memory.  The only serial dependency left is deck -> select -> deck.  Player 0's
// hand is filled first, then player 1's, ... (PlayerToDeal, hanabi_state.cc:157-164).
template <class V>
HB_DEV void reset_draw(V v, const Spec& spec, const uint32_t* qa, const uint32_t* lo,
                       const uint32_t* hi, int sampler_mode, HB_JOBRESULT(V)& out) {
  typedef typename V::Word W;
  uint64_t deck = spec.full_deck;
  int n = v.deck_max();
  const int H = v.H(), total = v.P() * H;
#pragma unroll
  for (int p = 0; p < kMaxPlayers; ++p) out.cards[p] = 0;
  int p = 0, i = 0;
#pragma unroll 1
  for (int r = 0; r < total; ++r) {
    const uint32_t w = qa[r];
    int pos = select64(deck, (int)(w & 0xffu));
    if (sampler_mode == 1 || ((w >> 8) && sampler_mode != 2))
      pos = pick_exact(v, deck, n, lo[r], hi[r]);
    int color, rank;
    card_of_position(v, pos, color, rank);
    const W card = (W)(uint32_t)(color | (rank << 3)) << (6 * i);
#pragma unroll
    for (int k = 0; k < V::kP; ++k) out.cards[k] |= (k == p) ? card : (W)0;
    deck &= ~(1ull << pos);
    --n;
    if (++i == H) { i = 0; ++p; }
  }
  out.deck_lo = (uint32_t)deck;
  out.deck_hi = (uint32_t)(deck >> 32);
}

// Installs a drawn initial deal as the lane's new game.
template <class V>
HB_DEV void reset_apply(V v, const Spec& spec, HB_GAME(V)& G, const HB_JOBRESULT(V)& res) {
  typedef typename V::Word W;
  init_fresh(v, spec, G, 0);
  const int H = v.H();
  const bool seer = spec.obs_type == kSeer;
  W call = 0, rall = 0;
#pragma unroll
  for (int i = 0; i < V::kH; ++i)
    if (i < H) { call |= (W)((1u << v.C()) - 1u) << (5 * i); rall |= (W)((1u << v.R()) - 1u) << (5 * i); }
  const uint32_t flags = (1u << H) - 1u;
#pragma unroll
  for (int p = 0; p < V::kP; ++p) {
    if (p < v.P()) {
      G.cards[p] = res.cards[p];
      G.cpl[p] = call;
      G.rpl[p] = rall;
      G.hw[p] = (uint32_t)H << 16;
      if (seer) {  // hanabi_state.cc:233-236: knowledge starts as the card itself
        W c1 = 0, r1 = 0;
#pragma unroll
        for (int i = 0; i < V::kH; ++i)
          if (i < H) {
            const uint32_t f = (uint32_t)(res.cards[p] >> (6 * i));
            c1 |= (W)(1u << (f & 7)) << (5 * i);
            r1 |= (W)(1u << ((f >> 3) & 7)) << (5 * i);
          }
        G.cpl[p] = c1;
        G.rpl[p] = r1;
        G.hw[p] |= flags | (flags << 8);
      }
    }
  }
  G.deck = (uint64_t)res.deck_lo | ((uint64_t)res.deck_hi << 32);
  G.deck_size = v.deck_max() - v.P() * H;
  G.mti = mt_wrap(G.mti + 2 * v.P() * H);
}

// ------------------------------------------------------------------ encoding
// A warp's 32 observations form ONE packed bit stream in shared memory: game
// `lane` owns bits [lane*dim, (lane+1)*dim), bit j of the stream is bit j & 31 of
// word j >> 5.  With dense output rows (pitch == dim) the stream maps 1:1 onto
// the tile's bytes, so 16 output bytes are exactly one aligned halfword of it.
//
// BitWriter appends fields to the lane's slice.  Fields are first gathered into
// 32-bit words aligned to the slice start (with a static view every width, hence
// every shift, is a compile-time constant); each completed word is then shifted
// by the lane's start offset s0 into stream position with one funnel shift.  A
// stream word shared by two neighbouring slices is written by the lower lane,
// which fetches the upper lane's contribution (`head`) with a shuffle.
struct BitWriter {
  uint32_t* out;  // stream word that holds the first bit of this lane's slice
  uint64_t acc;
  uint32_t prev, head, last_sh;
  int nbits, wi, s0;
  HB_DEV BitWriter(uint32_t* stream, int lane, int dim) : acc(0), prev(0), head(0), last_sh(0),
                                                           nbits(0), wi(0) {
    const int start = lane * dim;
    out = stream + (start >> 5);
    s0 = start & 31;
  }
  HB_DEV void emit(uint32_t word) {
    const uint32_t sh = __funnelshift_l(prev, word, s0);  // (word << s0) | (prev >> (32 - s0))
    if (wi == 0 && s0 != 0) head = sh;  // the lower neighbour owns this word
    else out[wi] = sh;
    last_sh = sh;
    prev = word;
    ++wi;
  }
  HB_DEV void put(uint32_t value, int width) {  // width in [0, 32], value < 2^width
    acc |= (uint64_t)value << nbits;
    nbits += width;
    if (nbits >= 32) {
      emit((uint32_t)acc);
      acc >>= 32;
      nbits -= 32;
    }
  }
  HB_DEV void put64(uint64_t value, int width) {  // width in [0, 64]
    const int w0 = width < 32 ? width : 32;
    put((uint32_t)(value & ((1ull << w0) - 1ull)), w0);
    put((uint32_t)(value >> 32), width - w0);
  }
  // Ends the slice (dim bits were put).  Must be reached by all 32 lanes.
  HB_DEV void finish(int lane, int dim) {
    const int rem = dim & 31;
    if (rem) emit((uint32_t)acc);               // last, partially filled aligned word
    const int tail_bits = rem ? rem : 32;       // valid bits of the last aligned word
    if (s0 + tail_bits > 32) emit(0u);          // its top bits spill into one more stream word
    const uint32_t next_head = __shfl_down_sync(kFull, head, 1);
    if (lane != 31 && next_head != 0u) out[wi - 1] = last_sh | next_head;
  }
};

// CanonicalObservationEncoder::Encode for `observer` with the HanabiObservation
// view folded in.  Writes the lane's obs_size bits into the warp's stream.
// Warp-convergent (every lane calls it).
// spread[cm] has bit c*R set for every colour c in the 5-bit mask cm, so that
// spread[cm] * rank_mask is the colour x rank outer product of the knowledge
// section in one multiply.  32 words of shared memory, one bank each: any access
// pattern is conflict-free.
HB_DEV uint32_t spread_entry(int cm, int C, int R) {
  uint32_t s = 0;
  for (int c = 0; c < C; ++c) s |= ((cm >> c) & 1u) << (c * R);
  return s;
}

template <class V>
HB_DEV void encode_obs(V v, const Spec& spec, const HB_GAME(V)& G, int observer, uint32_t* stream,
                       int lane, const uint32_t* spread, int slice_bits) {
  const int P = v.P(), H = v.H(), C = v.C(), R = v.R(), bpc = v.bpc();
  // slice_bits = distance between two lanes' slices in the stream: obs_size for
  // the dense byte expansion, a multiple of 128 for the bit-packed output rows
  BitWriter bw(stream, lane, slice_bits);
  // --- hands (canonical_encoders.cc:62-105)
  uint32_t short_bits = 0;
#pragma unroll
  for (int q = 0; q < V::kP; ++q) {
    if (q < P) {
      int p = observer + q;
      p -= p >= P ? P : 0;
      const typename V::Word hand = sel<V>(G.cards, p);
      const int n = (int)((sel<V>(G.hw, p) >> 16) & 15);
      short_bits |= (uint32_t)(n < H) << q;
      if (q >= 1) {
#pragma unroll
        for (int i = 0; i < V::kH; ++i) {
          if (i < H) {
            const uint32_t f = (uint32_t)(hand >> (6 * i));
            const int id = (int)(f & 7) * R + (int)((f >> 3) & 7);
            bw.put(i < n ? (1u << id) : 0u, bpc);
          }
        }
      }
    }
  }
  bw.put(short_bits, P);
  // --- board (canonical_encoders.cc:123-167)
  bw.put64((1ull << G.deck_size) - 1ull, v.deck_max() - P * H);
#pragma unroll
  for (int c = 0; c < V::kC; ++c) {
    if (c < C) {
      const int top = (G.fw >> (3 * c)) & 7;
      bw.put(top > 0 ? (1u << (top - 1)) : 0u, R);
    }
  }
  bw.put((1u << G.info) - 1u, v.IT());
  bw.put((1u << G.life) - 1u, v.LT());
  // --- discards (canonical_encoders.cc:188-211): stored in this very layout
  bw.put64(G.disc, v.deck_max());
  // --- last action (canonical_encoders.cc:236-338; observer-relative actor as
  //     hanabi_observation.cc:34-49)
  {
    const uint32_t la = G.last;
    const bool valid = la & 1u;
    const int type = (la >> kLaTypeSh) & 3;
    int rel = (int)((la >> kLaActorSh) & 7) - observer;
    rel += rel < 0 ? P : 0;
    const bool is_hint = valid && type >= kLaRevealColor;
    const bool is_card = valid && type <= kLaDiscard;
    bw.put(valid ? (1u << rel) : 0u, P);
    bw.put(valid ? (1u << type) : 0u, 4);
    int tgt = rel + (int)((la >> kLaOffSh) & 7);
    tgt -= tgt >= P ? P : 0;
    bw.put(is_hint ? (1u << tgt) : 0u, P);
    const int val = (la >> kLaValSh) & 7;
    bw.put((valid && type == kLaRevealColor) ? (1u << val) : 0u, C);
    bw.put((valid && type == kLaRevealRank) ? (1u << val) : 0u, R);
    bw.put(is_hint ? (V::kWide ? la_reveal(la) : ((la >> kLaRevealSh) & 31u)) : 0u, H);
    bw.put(is_card ? (1u << ((la >> kLaIdxSh) & 7)) : 0u, H);
    const int id = (int)((la >> kLaColSh) & 7) * R + (int)((la >> kLaRankSh) & 7);
    bw.put(is_card ? (1u << id) : 0u, bpc);
    bw.put((valid && type == kLaPlay) ? ((la >> kLaScoredSh) & 3u) : 0u, 2);
  }
  // --- card knowledge (canonical_encoders.cc:366-419), absent for kMinimal
  if (spec.obs_type != kMinimal) {
#pragma unroll
    for (int q = 0; q < V::kP; ++q) {
      if (q < P) {
        int p = observer + q;
        p -= p >= P ? P : 0;
        const typename V::Word cpl = sel<V>(G.cpl, p), rpl = sel<V>(G.rpl, p);
        const uint32_t hw = sel<V>(G.hw, p);
        const int n = (int)((hw >> 16) & 15);
#pragma unroll
        for (int i = 0; i < V::kH; ++i) {
          if (i < H) {
            const uint32_t cm = (uint32_t)(cpl >> (5 * i)) & 31u, rm = (uint32_t)(rpl >> (5 * i)) & 31u;
            const uint32_t plaus = spread[cm] * rm;  // rm < 2^R: the products do not overlap
            const bool have = i < n;
            bw.put(have ? plaus : 0u, bpc);
            bw.put((have && ((hw >> i) & 1u)) ? cm : 0u, C);
            bw.put((have && ((hw >> (8 + i)) & 1u)) ? rm : 0u, R);
          }
        }
      }
    }
  }
  bw.finish(lane, spec.obs_size);
  if (slice_bits != spec.obs_size) {  // packed rows: zero the padding words of this lane's row
    uint32_t* row = stream + lane * (slice_bits >> 5);
    for (int w = (spec.obs_size + 31) >> 5; w < (slice_bits >> 5); ++w) row[w] = 0u;
  }
}

// The observation / mask rows are written once and never read back by the engine.  Long rows
// (the full game's 658+ bytes: a warp's tile is 20+ KB) go out with streaming, evict-first stores
// so that they do not displace the game state in L2 -- without them Full 4p is 9 % slower.
// Short rows (the small games' 171 / 281 bytes: 5-9 KB per tile, every tile edge a 128-byte line
// shared with the neighbouring warp) are 2.7 % faster with plain L2 stores: an evict-first line
// leaves the L2 before the neighbour has completed it (profiles/r02al_exp_ab_out.log).
#ifndef HB_STREAM_STORES
#define HB_STREAM_STORES -1  // -1: by row length (view / rule set), 0: never, 1: always
#endif
template <bool STREAM = true>
HB_DEV void store_out(uint4* p, uint4 v) {
  if constexpr (HB_STREAM_STORES != 0 && (STREAM || HB_STREAM_STORES == 1)) __stcs(p, v);
  else __stcg(p, v);
}

// Bit-packed observation rows: copies the tile's stream (n_valid rows of
// `row_words` 32-bit words, bit j of a row = observation bit j) to global memory
// with 128-bit stores.
HB_DEV void store_packed_rows(const uint32_t* stream, int row_words, uint32_t* out, int n_valid,
                              int lane) {
  const int quads = (n_valid * row_words) >> 2;  // row_words is a multiple of 4
  const uint4* src = reinterpret_cast<const uint4*>(stream);
  uint4* dst = reinterpret_cast<uint4*>(out);
  for (int k = lane; k < quads; k += 32) store_out(dst + k, src[k]);
}

// Legal-move masks of the tile as a packed stream (A bits per game).  The slices
// are narrower than a word, so they are OR-ed into pre-zeroed words.
// Warp-convergent.
HB_DEV void pack_mask_stream(uint32_t* stream, int A, uint64_t bits, bool valid, int lane) {
  for (int w = lane; w < A + 1; w += 32) stream[w] = 0u;
  __syncwarp();
  if (valid) {
    const int start = lane * A, w0 = start >> 5, sh = start & 31;
    const uint64_t lo = bits << sh;
    const uint32_t hi = sh ? (uint32_t)(bits >> (64 - sh)) : 0u;
    atomicOr(&stream[w0], (uint32_t)lo);
    if ((uint32_t)(lo >> 32)) atomicOr(&stream[w0 + 1], (uint32_t)(lo >> 32));
    if (hi) atomicOr(&stream[w0 + 2], hi);
  }
}

// Network-input rows: the tile's packed rows (row_words 32-bit words per game, bit j =
// observation bit j, padding zero) expanded to 0/1 elements of a GEMM input type, K padded
// with zeros up to `pitch` elements.  lut16[b] holds the eight 16-bit elements of byte b
// (bfloat16 or float16 ones); float32 rows use the same table: a float32 1.0 is a bfloat16
// 1.0 in the upper half of the word.  One 128-bit store per 8 elements (two for float32).
// `magic` = ceil(2^32 / chunks) with chunks = pitch / 8 turns the flat chunk index into
// (game, chunk) with one multiply.
HB_DEV void expand_net_rows(const uint32_t* stream, int row_words, void* out, int64_t pitch, int chunks,
                            uint32_t magic, int n_valid, int lane, const uint4* lut16, bool f32) {
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(stream);
  const int row_bytes = row_words * 4;
  const int total = n_valid * chunks;
#pragma unroll 2
  for (int idx = lane; idx < total; idx += 32) {
    const int game = (int)__umulhi((uint32_t)idx, magic);
    const int j = idx - game * chunks;
    const uint32_t b = j < row_bytes ? bytes[game * row_bytes + j] : 0u;
    const uint4 v = lut16[b];
    if (!f32) {
      store_out(reinterpret_cast<uint4*>(static_cast<uint16_t*>(out) + (int64_t)game * pitch) + j, v);
    } else {
      uint4* dst = reinterpret_cast<uint4*>(static_cast<float*>(out) + (int64_t)game * pitch) + 2 * j;
      store_out(dst, make_uint4(v.x << 16, v.x & 0xffff0000u, v.y << 16, v.y & 0xffff0000u));
      store_out(dst + 1, make_uint4(v.z << 16, v.z & 0xffff0000u, v.w << 16, v.w & 0xffff0000u));
    }
  }
}

// 4 bits -> 4 bytes of 0/1 (little-endian byte i = bit i).
HB_DEV uint32_t nibble_to_bytes(uint32_t n) { return ((n & 15u) * 0x00204081u) & 0x01010101u; }

HB_DEV uint4 bits16_to_bytes(uint32_t bits) {
  uint4 val;
  val.x = nibble_to_bytes(bits);
  val.y = nibble_to_bytes(bits >> 4);
  val.z = nibble_to_bytes(bits >> 8);
  val.w = ((bits >> 12) * 0x00204081u) & 0x01010101u;
  return val;
}

// Warp-cooperative expansion of a tile's packed bit stream (`dim` bits per game,
// n_valid games) into bytes of 0/1 in global memory.  Dense rows (pitch == dim):
// the tile is one contiguous, 16-byte aligned byte range and every 128-bit store
// is fed by one halfword of the stream, a warp covering 512 contiguous bytes per
// instruction.  The stream must have one readable pad word after its last word.
// (A 256-entry shared-memory table instead of the multiplies was measured 3 % slower.)
#ifndef HB_EXPAND_LUT16
#define HB_EXPAND_LUT16 0  // 1: nibble -> four bytes through a 16-word shared-memory table (experiment)
#endif
template <bool STREAM = true>
HB_DEV void expand_stream(const uint32_t* stream, int dim, uint8_t* out, int64_t pitch,
                          int n_valid, int lane, const uint32_t* nib = nullptr) {
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (pitch == dim && aligned) {
    const int total = n_valid * dim;
    const int full_chunks = total >> 4;
    const uint16_t* half = reinterpret_cast<const uint16_t*>(stream);
    uint4* dst = reinterpret_cast<uint4*>(out);
    // unroll factor measured on B200 (Full 2p, 2^20 games): 1: 0.1892, 2: 0.1882, 3: 0.1903,
    // 4: 0.1906, 8: 0.1923 ms per step -- the smaller loop body wins (code size, registers)
#ifndef HB_EXPAND_UNROLL
#define HB_EXPAND_UNROLL 2
#endif
    constexpr int kUnroll = HB_EXPAND_UNROLL;
#if HB_EXPAND_LUT16
#pragma unroll kUnroll
    for (int k = lane; k < full_chunks; k += 32) {
      const uint32_t b = half[k];
      store_out<STREAM>(dst + k, make_uint4(nib[b & 15u], nib[(b >> 4) & 15u], nib[(b >> 8) & 15u], nib[b >> 12]));
    }
#else
#pragma unroll kUnroll
    for (int k = lane; k < full_chunks; k += 32) store_out<STREAM>(dst + k, bits16_to_bytes(half[k]));
#endif
    for (int b = (full_chunks << 4) + lane; b < total; b += 32)  // ragged tail of a partial tile
      out[b] = (uint8_t)((stream[b >> 5] >> (b & 31)) & 1u);
  } else if ((pitch & 15) == 0 && aligned) {
    const int chunks = (dim + 15) >> 4;
    for (int idx = lane; idx < n_valid * chunks; idx += 32) {
      const int g = idx / chunks, j = idx - g * chunks;
      const int o = j * 16, avail = dim - o;
      const int pos = g * dim + o;
      const uint32_t bits =
          __funnelshift_r(stream[pos >> 5], stream[(pos >> 5) + 1], pos & 31) & 0xffffu;
      uint8_t* d = out + (int64_t)g * pitch + o;
      if (avail >= 16) {
        store_out<STREAM>(reinterpret_cast<uint4*>(d), bits16_to_bytes(bits));
      } else {
        for (int i = 0; i < avail; ++i) d[i] = (uint8_t)((bits >> i) & 1u);
      }
    }
  } else {
    for (int g = 0; g < n_valid; ++g)
      for (int o = lane; o < dim; o += 32) {
        const int pos = g * dim + o;
        out[(int64_t)g * pitch + o] = (uint8_t)((stream[pos >> 5] >> (pos & 31)) & 1u);
      }
  }
}

}  // namespace hb
#endif  // HB_ENGINE_CUH_
