/*
 * oracle/oracle_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Common C interface shared by the two CPU checkers under oracle/:
 *   - libhanabi_oracle.so  (oracle/hanabi_oracle.c: plain-C restatement, prefix Orc*)
 *   - _ref/libhanabi_ref.so (oracle/ref_driver.cc compiled together with the
 *     UNMODIFIED reference sources under /root/reference, prefix Ref*)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load these libraries.  The product (hanabi_b200/) never does.
 *
 * Both libraries run N independent single-threaded games, one RNG stream per
 * game (the reference keeps one std::mt19937 per HanabiGame:
 * hanabi_lib/hanabi_game.h:114, seeded hanabi_lib/hanabi_game.cc:48-51).
 */
#ifndef HANABI_ORACLE_API_H_
#define HANABI_ORACLE_API_H_

#include <stdint.h>

/* Canonical unpacked per-game state dump: int32[HB_DUMP_LEN].
 * Layout (P<=5 players, <=8 cards per hand, <=25 card types):
 *   [0] cur_player  [1] information_tokens  [2] life_tokens  [3] deck_size
 *   [4] score       [5] end_of_game_status  [6..10] fireworks[color]
 *   [11..15] hand size of player p
 *   [16 + p*8 + s]        card id (color*R+rank) of player p slot s, -1 if none
 *   [56 + p*8 + s]        color-plausible bitmask
 *   [96 + p*8 + s]        rank-plausible bitmask
 *   [136 + p*8 + s]       hinted color or -1
 *   [176 + p*8 + s]       hinted rank or -1
 *   [216 + k]             deck count of card type k (k < 25)
 *   [241 + k]             discard count of card type k
 */
#define HB_DUMP_LEN 272
#define HB_DUMP_CUR 0
#define HB_DUMP_INFO 1
#define HB_DUMP_LIFE 2
#define HB_DUMP_DECK 3
#define HB_DUMP_SCORE 4
#define HB_DUMP_END 5
#define HB_DUMP_FIREWORKS 6
#define HB_DUMP_HANDSIZE 11
#define HB_DUMP_CARDS 16
#define HB_DUMP_CPLAUS 56
#define HB_DUMP_RPLAUS 96
#define HB_DUMP_CHINT 136
#define HB_DUMP_RHINT 176
#define HB_DUMP_DECKCOUNT 216
#define HB_DUMP_DISCARDCOUNT 241

#endif
