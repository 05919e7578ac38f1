// hb_agents.cuh -- the reference's hand-written policies as batch kernels, so that
// ad-hoc teams (Rainbow + rule-based seats, training/run_adhoc_experiment.py:180-263,
// 394-469) run without a host round trip per game:
//   kind 0  RuleBasedAgent   agents/rule_based_agent.py:15-302 (act / act_train)
//   kind 1  SimpleAgent      agents/simple_agent.py:20-69
// One thread per game; the policy acts for the player to move if that seat is in
// `seat_mask`, reading the packed game state directly (no observation object).
//
// RuleBasedAgent keeps `rank_hinted_but_no_play[players][num_cards]` per agent
// OBJECT (rule_based_agent.py:30-33; num_cards = 5 below four players, else 4); the
// reference never resets it between episodes, and create_adhoc_team gives several
// seats the same object (run_adhoc_experiment.py:251-258), so the state lives in a
// caller-owned slot array with a seat -> slot map.  Its quirks are reproduced on
// purpose: rows are indexed by observer-relative player, the hint target row is
// (player + 1) % (players - 1) (:63), the scan over `last_moves` stops at the
// first qualifying rank hint (:70), and an action that is not legal becomes the
// last legal move (act_train, :291-296).
#ifndef HB_AGENTS_CUH_
#define HB_AGENTS_CUH_

#include "hb_engine.cuh"

namespace hb {

struct AgentArgs {
  Spec spec;
  const uint4* state;
  const uint4* hist;
  int64_t N;
  uint32_t* agent_state;  // [N][num_slots], bit q*5+i = rank_hinted_but_no_play[q][i]
  int num_slots;
  int seat_slot[kMaxPlayers];
  uint32_t seat_mask;
  int kind;
  int max_information_tokens;  // SimpleAgent's config['max_information_tokens']
  int32_t* actions;
};

// list.pop(idx); list.append(False) on the nc-bit row q of the packed table.
HB_DEV uint32_t agent_pop(uint32_t table, int q, int idx, int nc) {
  const uint32_t row = (table >> (5 * q)) & 31u;
  const uint32_t low = (1u << idx) - 1u;
  const uint32_t nrow = ((row & low) | ((row >> 1) & ~low)) & ((1u << nc) - 1u);
  return (table & ~(31u << (5 * q))) | (nrow << (5 * q));
}

__global__ void k_agent_act(const __grid_constant__ AgentArgs a) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  const DView v(a.spec);
  Game G;
  load_game(a.state, a.N, g, v, G);
  if (G.done) return;
  const int cur = G.cur;
  if (!((a.seat_mask >> cur) & 1u)) return;
  const int P = a.spec.P, H = a.spec.H, C = a.spec.C, R = a.spec.R;
  const bool see_knowledge = a.spec.obs_type != kMinimal;  // kMinimal hides all knowledge
  const uint32_t own_hw = sel<DView>(G.hw, cur);
  const int own_n = (int)((own_hw >> 16) & 15);
  const uint32_t own_ch = see_knowledge ? (own_hw & 0xffu) : 0u;
  const uint32_t own_rh = see_knowledge ? ((own_hw >> 8) & 0xffu) : 0u;
  int uid = -1;

  if (a.kind == 1) {
    // ---- SimpleAgent.act (simple_agent.py:32-69)
    for (int i = 0; i < own_n && uid < 0; ++i)
      if (((own_ch | own_rh) >> i) & 1u) uid = H + i;  // play the first hinted card
    if (uid < 0 && G.info > 0) {
      for (int q = 1; q < P && uid < 0; ++q) {
        int p = cur + q;
        p -= p >= P ? P : 0;
        const uint32_t hand = sel<DView>(G.cards, p), hw = sel<DView>(G.hw, p);
        const int n = (int)((hw >> 16) & 15);
        const uint32_t ch = see_knowledge ? (hw & 0xffu) : 0u;
        for (int i = 0; i < n && uid < 0; ++i) {
          const uint32_t f = hand >> (6 * i);
          const int color = f & 7, rank = (f >> 3) & 7;
          if (rank == (int)((G.fw >> (3 * color)) & 7) && !((ch >> i) & 1u))
            uid = 2 * H + (P - 1) * C + (q - 1) * R + rank;  // REVEAL_RANK
        }
      }
    }
    if (uid < 0) uid = G.info < a.max_information_tokens ? 0 : H;  // DISCARD 0 / PLAY 0
    a.actions[g] = uid;
    return;
  }

  // ---- RuleBasedAgent.act_train
  const int nc = P < 4 ? 5 : 4;
  const int slot = a.seat_slot[cur];
  uint32_t T = a.agent_state[g * (int64_t)a.num_slots + slot];

  // check_if_not_playable_hint (:44-73): non-deal moves newest -> oldest, back to
  // and including the observer's own previous action
  {
    const uint4 h = a.hist ? a.hist[g] : make_uint4(0, 0, 0, 0);
    const uint32_t rec[5] = {G.last, h.x, h.y, h.z, h.w};
    for (int j = 0; j < P && j < 5; ++j) {
      const uint32_t la = rec[j];
      if (!(la & 1u)) break;
      const int type = (la >> kLaTypeSh) & 3;
      const int actor = (la >> kLaActorSh) & 7;
      int rel = actor - cur;
      rel += rel < 0 ? P : 0;
      const uint32_t reveal = (la >> kLaRevealSh) & 31u;
      if (type == kLaRevealRank && ((la >> kLaOffSh) & 7) == 1 && (reveal & 1u)) {
        const int tgt = (rel + 1) % (P - 1);
        T |= (reveal & ((1u << nc) - 1u)) << (5 * tgt);
        break;
      } else if (type == kLaPlay || type == kLaDiscard) {
        T = agent_pop(T, rel, (la >> kLaIdxSh) & 7, nc);
      }
      if (actor == cur) break;
    }
  }

  // maybe_play_lowest_playable_card (:75-112)
  for (int i = 0; i < own_n && uid < 0; ++i)
    if ((own_ch >> i) & 1u) { T = agent_pop(T, 0, i, nc); uid = H + i; }
  for (int i = 0; i < own_n && uid < 0; ++i)
    if (((own_rh >> i) & 1u) && !((T >> i) & 1u)) { T = agent_pop(T, 0, i, nc); uid = H + i; }

  // maybe_give_helpful_hint (:114-228)
  if (uid < 0 && G.info != 0) {
    int best = 0, best_uid = -1;
    for (int q = 1; q < P; ++q) {
      int p = cur + q;
      p -= p >= P ? P : 0;
      const uint32_t hand = sel<DView>(G.cards, p), hw = sel<DView>(G.hw, p);
      const int n = (int)((hw >> 16) & 15);
      const uint32_t ch = see_knowledge ? (hw & 0xffu) : 0u;
      const uint32_t rh = see_knowledge ? ((hw >> 8) & 0xffu) : 0u;
      uint32_t playable = 0;
      int colors[kMaxHand], ranks[kMaxHand];
      for (int i = 0; i < n; ++i) {
        const uint32_t f = hand >> (6 * i);
        colors[i] = f & 7;
        ranks[i] = (f >> 3) & 7;
        if (ranks[i] == (int)((G.fw >> (3 * colors[i])) & 7)) playable |= 1u << i;
      }
      // colour hints, colours in order of first appearance among playable cards
      uint32_t seen = 0;
      for (int i = 0; i < n; ++i) {
        if (!((playable >> i) & 1u) || ((seen >> colors[i]) & 1u)) continue;
        seen |= 1u << colors[i];
        int content = 0;
        bool bad = false;
        for (int k = 0; k < n && !bad; ++k) {
          if (colors[k] != colors[i]) continue;
          if (((playable >> k) & 1u) && !((ch >> k) & 1u)) ++content;
          else if (!((playable >> k) & 1u)) bad = true;
        }
        if (!bad && content > best) { best = content; best_uid = 2 * H + (q - 1) * C + colors[i]; }
      }
      // rank hints
      seen = 0;
      for (int i = 0; i < n; ++i) {
        if (!((playable >> i) & 1u) || ((seen >> ranks[i]) & 1u)) continue;
        seen |= 1u << ranks[i];
        int content = 0;
        bool bad = false;
        for (int k = 0; k < n && !bad; ++k) {
          if (ranks[k] != ranks[i]) continue;
          const bool marked = (T >> (5 * q + k)) & 1u;  // rank_hinted_but_no_play[player_offset][index]
          if (((playable >> k) & 1u) && (!((rh >> k) & 1u) || marked)) ++content;
          else if (!((playable >> k) & 1u)) bad = true;
        }
        if (!bad && content > best) {
          best = content;
          best_uid = 2 * H + (P - 1) * C + (q - 1) * R + ranks[i];
        }
      }
    }
    if (best > 0) uid = best_uid;
  }

  // discard the oldest card if allowed, else hint the right neighbour's oldest card's rank (:246-282)
  if (uid < 0) {
    if (G.info < a.spec.IT && own_n > 0) {
      T = agent_pop(T, 0, 0, nc);
      uid = 0;
    } else {
      int p = cur + 1;
      p -= p >= P ? P : 0;
      const int rank = (sel<DView>(G.cards, p) >> 3) & 7;
      uid = 2 * H + (P - 1) * C + rank;  // REVEAL_RANK, target_offset 1
    }
  }
  // act_train (:284-296): an action that is not among the legal moves becomes the last one
  const uint64_t legal = legal_bits(v, G);
  if (!((legal >> uid) & 1ull)) uid = 63 - __clzll((long long)legal);
  a.agent_state[g * (int64_t)a.num_slots + slot] = T;
  a.actions[g] = uid;
}

}  // namespace hb
#endif  // HB_AGENTS_CUH_
