// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).
//
// Thin batch driver around the UNMODIFIED reference engine.  It is compiled by
// oracle/Makefile together with the reference's own sources, taken where they
// lie under /root/reference/hanabi_learning_environment/hanabi_lib/*.cc, into
// oracle/_ref/libhanabi_ref.so.  No reference source is copied into this repo;
// this file only *calls* the reference's public C++ classes:
//   HanabiGame            hanabi_lib/hanabi_game.h
//   HanabiState           hanabi_lib/hanabi_state.h   (ApplyMove, ApplyRandomChance,
//                                                      LegalMoves, Score, EndOfGameStatus)
//   HanabiObservation     hanabi_lib/hanabi_observation.h
//   CanonicalObservationEncoder  hanabi_lib/canonical_encoders.h
//
// The per-game loop mirrors rl_env.HanabiEnv.reset/step
// (hanabi_learning_environment/rl_env.py:210-217, 343-366): apply the move,
// deal while the chance player is to act, reward = score - last_score,
// done = IsTerminal().
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "canonical_encoders.h"
#include "hanabi_game.h"
#include "hanabi_observation.h"
#include "hanabi_state.h"
#include "oracle_api.h"

namespace hle = hanabi_learning_env;

namespace {

struct RefGame {
  std::unique_ptr<hle::HanabiGame> game;
  std::unique_ptr<hle::HanabiState> state;
  std::unique_ptr<hle::CanonicalObservationEncoder> encoder;
  uint64_t policy_rng = 0;
};

struct RefBatch {
  std::vector<RefGame> games;
  int obs_size = 0;
  int max_moves = 0;
  int num_players = 0;
};

void DealAll(hle::HanabiState* s) {
  while (s->CurPlayer() == hle::kChancePlayerId) s->ApplyRandomChance();
}

void ResetGame(RefGame* g) {
  g->state.reset(new hle::HanabiState(g->game.get()));
  DealAll(g->state.get());
}

void EncodeInto(const RefGame& g, int observer, uint8_t* row) {
  hle::HanabiObservation obs(*g.state, observer);
  std::vector<int> enc = g.encoder->Encode(obs);
  for (size_t i = 0; i < enc.size(); ++i) row[i] = static_cast<uint8_t>(enc[i]);
}

void MaskInto(const RefGame& g, int max_moves, uint8_t* row) {
  std::memset(row, 0, max_moves);
  int cur = g.state->CurPlayer();
  if (cur < 0) return;
  for (const hle::HanabiMove& m : g.state->LegalMoves(cur)) {
    row[g.game->GetMoveUid(m)] = 1;
  }
}

// One env step for one game.  Returns 0 ok, 1 illegal action (state unchanged).
int StepGame(RefGame* g, int action, int auto_reset, int32_t* reward,
             uint8_t* done, int32_t* score) {
  hle::HanabiState* s = g->state.get();
  if (s->IsTerminal()) {  // stepping a finished game without auto-reset: no-op
    *reward = 0;
    *done = 1;
    if (score) *score = s->Score();
    return 0;
  }
  if (action < 0 || action >= g->game->MaxMoves() ||
      !s->MoveIsLegal(g->game->GetMove(action))) {
    *reward = 0;
    *done = 0;
    if (score) *score = s->Score();
    return 1;
  }
  int last_score = s->Score();
  s->ApplyMove(g->game->GetMove(action));
  DealAll(s);
  *reward = s->Score() - last_score;
  *done = s->IsTerminal() ? 1 : 0;
  if (score) *score = s->Score();
  if (*done && auto_reset) ResetGame(g);
  return 0;
}

inline uint64_t SplitMix(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace

extern "C" {

void* RefBatchNew(int n_kv, const char** kv, int num_games,
                  const int32_t* seeds) {
  auto* b = new RefBatch;
  b->games.resize(num_games);
  for (int i = 0; i < num_games; ++i) {
    std::unordered_map<std::string, std::string> params;
    for (int k = 0; k + 1 < n_kv; k += 2) params[kv[k]] = kv[k + 1];
    params["seed"] = std::to_string(seeds[i]);
    RefGame& g = b->games[i];
    g.game.reset(new hle::HanabiGame(params));
    g.encoder.reset(new hle::CanonicalObservationEncoder(g.game.get()));
    g.policy_rng = 0x1234567ull + static_cast<uint64_t>(i) * 7919ull;
  }
  const hle::HanabiGame& g0 = *b->games[0].game;
  b->obs_size = b->games[0].encoder->Shape()[0];
  b->max_moves = g0.MaxMoves();
  b->num_players = g0.NumPlayers();
  return b;
}

void RefBatchDelete(void* vb) { delete static_cast<RefBatch*>(vb); }
int RefBatchObservationSize(void* vb) { return static_cast<RefBatch*>(vb)->obs_size; }
int RefBatchMaxMoves(void* vb) { return static_cast<RefBatch*>(vb)->max_moves; }
int RefBatchNumPlayers(void* vb) { return static_cast<RefBatch*>(vb)->num_players; }
int RefBatchNumGames(void* vb) { return (int)static_cast<RefBatch*>(vb)->games.size(); }

void RefBatchReset(void* vb, const uint8_t* which) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i) {
    if (which == nullptr || which[i]) ResetGame(&b->games[i]);
  }
}

void RefBatchStep(void* vb, const int32_t* actions, int32_t* rewards,
                  uint8_t* dones, int32_t* scores, int auto_reset,
                  uint8_t* illegal) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i) {
    int rc = StepGame(&b->games[i], actions[i], auto_reset, &rewards[i],
                      &dones[i], scores ? &scores[i] : nullptr);
    if (illegal) illegal[i] = static_cast<uint8_t>(rc);
  }
}

// observer < 0: the current player of each game.
void RefBatchEncode(void* vb, int observer, uint8_t* obs, int64_t pitch) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i) {
    int o = observer < 0 ? b->games[i].state->CurPlayer() : observer;
    EncodeInto(b->games[i], o, obs + pitch * (int64_t)i);
  }
}

void RefBatchLegalMask(void* vb, uint8_t* mask) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i)
    MaskInto(b->games[i], b->max_moves, mask + (int64_t)b->max_moves * i);
}

void RefBatchCurPlayer(void* vb, int32_t* cur) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i)
    cur[i] = b->games[i].state->CurPlayer();
}

void RefBatchDumpState(void* vb, int32_t* out) {
  auto* b = static_cast<RefBatch*>(vb);
  for (size_t i = 0; i < b->games.size(); ++i) {
    int32_t* d = out + (int64_t)HB_DUMP_LEN * i;
    for (int k = 0; k < HB_DUMP_LEN; ++k) d[k] = 0;
    const hle::HanabiGame& game = *b->games[i].game;
    const hle::HanabiState& s = *b->games[i].state;
    const int C = game.NumColors(), R = game.NumRanks();
    d[HB_DUMP_CUR] = s.CurPlayer();
    d[HB_DUMP_INFO] = s.InformationTokens();
    d[HB_DUMP_LIFE] = s.LifeTokens();
    d[HB_DUMP_DECK] = s.Deck().Size();
    d[HB_DUMP_SCORE] = s.Score();
    d[HB_DUMP_END] = static_cast<int>(s.EndOfGameStatus());
    for (int c = 0; c < C; ++c) d[HB_DUMP_FIREWORKS + c] = s.Fireworks()[c];
    for (int p = 0; p < 5; ++p)
      for (int k = 0; k < 8; ++k) {
        d[HB_DUMP_CARDS + p * 8 + k] = -1;
        d[HB_DUMP_CHINT + p * 8 + k] = -1;
        d[HB_DUMP_RHINT + p * 8 + k] = -1;
      }
    for (int p = 0; p < game.NumPlayers(); ++p) {
      const hle::HanabiHand& hand = s.Hands()[p];
      d[HB_DUMP_HANDSIZE + p] = static_cast<int>(hand.Cards().size());
      for (size_t k = 0; k < hand.Cards().size() && k < 8; ++k) {
        const hle::HanabiCard& card = hand.Cards()[k];
        const auto& kn = hand.Knowledge()[k];
        d[HB_DUMP_CARDS + p * 8 + k] = card.Color() * R + card.Rank();
        int cm = 0, rm = 0;
        for (int c = 0; c < C; ++c) cm |= kn.ColorPlausible(c) ? (1 << c) : 0;
        for (int r = 0; r < R; ++r) rm |= kn.RankPlausible(r) ? (1 << r) : 0;
        d[HB_DUMP_CPLAUS + p * 8 + k] = cm;
        d[HB_DUMP_RPLAUS + p * 8 + k] = rm;
        d[HB_DUMP_CHINT + p * 8 + k] = kn.ColorHinted() ? kn.Color() : -1;
        d[HB_DUMP_RHINT + p * 8 + k] = kn.RankHinted() ? kn.Rank() : -1;
      }
    }
    for (int c = 0; c < C; ++c)
      for (int r = 0; r < R; ++r)
        d[HB_DUMP_DECKCOUNT + c * R + r] = s.Deck().CardCount(c, r);
    for (const hle::HanabiCard& card : s.DiscardPile())
      d[HB_DUMP_DISCARDCOUNT + card.Color() * R + card.Rank()] += 1;
  }
}

// Timed CPU baseline: `steps` env-steps for every game with a uniform random
// legal policy (private splitmix stream; not part of parity), auto-reset,
// current-player observation encoded to uint8 + legal mask written every step,
// `threads` std::threads over disjoint slices of games.  Modelled on the
// reference's own C++ rollout loop, hanabi_learning_environment/game_example.cc:34-87.
// obs: [num_games, pitch] uint8, mask: [num_games, max_moves] uint8.
// Returns the total number of env steps executed.
int64_t RefBatchRollout(void* vb, int steps, int threads, uint8_t* obs,
                        int64_t pitch, uint8_t* mask, int64_t* episodes_out,
                        int64_t* score_sum_out) {
  auto* b = static_cast<RefBatch*>(vb);
  const int n = static_cast<int>(b->games.size());
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  std::vector<int64_t> episodes(threads, 0), score_sum(threads, 0);
  auto worker = [&](int t) {
    int lo = (int)((int64_t)n * t / threads), hi = (int)((int64_t)n * (t + 1) / threads);
    for (int step = 0; step < steps; ++step) {
      for (int i = lo; i < hi; ++i) {
        RefGame& g = b->games[i];
        hle::HanabiState* s = g.state.get();
        auto legal = s->LegalMoves(s->CurPlayer());
        hle::HanabiMove mv = legal[SplitMix(&g.policy_rng) % legal.size()];
        s->ApplyMove(mv);
        DealAll(s);
        if (s->IsTerminal()) {
          episodes[t] += 1;
          score_sum[t] += s->Score();
          ResetGame(&g);
        }
        EncodeInto(g, g.state->CurPlayer(), obs + pitch * (int64_t)i);
        MaskInto(g, b->max_moves, mask + (int64_t)b->max_moves * i);
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  int64_t e = 0, sc = 0;
  for (int t = 0; t < threads; ++t) { e += episodes[t]; sc += score_sum[t]; }
  if (episodes_out) *episodes_out = e;
  if (score_sum_out) *score_sum_out = sc;
  return (int64_t)steps * n;
}

// A URNG that replays two chosen 32-bit words through libstdc++'s own
// std::discrete_distribution, i.e. the exact arithmetic of
// HanabiGame::PickRandomChance (hanabi_lib/hanabi_game.cc:106-112) for an
// arbitrary point of the mt19937 output space.  Used to pin the restated
// sampler (oracle + CUDA) at adversarial boundary values.
struct ReplayUrng {
  typedef uint32_t result_type;
  uint32_t w[2];
  int i = 0;
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return 0xFFFFFFFFu; }
  result_type operator()() { return w[(i++) & 1]; }
};

int RefDiscreteSample(const double* weights, int n, uint32_t lo, uint32_t hi) {
  std::discrete_distribution<std::mt19937::result_type> dist(weights, weights + n);
  ReplayUrng g;
  g.w[0] = lo;
  g.w[1] = hi;
  return static_cast<int>(dist(g));
}

// Same, but from a deck: weights = count/deck_size for the non-empty types in
// ascending type order, as HanabiState::ChanceOutcomes builds them
// (hanabi_lib/hanabi_state.cc:313-325, 277-280).  Returns the card type index.
int RefDealFromCounts(const int32_t* counts, int n_types, uint32_t lo, uint32_t hi) {
  std::vector<double> p;
  std::vector<int> idx;
  int total = 0;
  for (int k = 0; k < n_types; ++k) total += counts[k];
  for (int k = 0; k < n_types; ++k)
    if (counts[k] > 0) {
      idx.push_back(k);
      p.push_back(static_cast<double>(counts[k]) / static_cast<double>(total));
    }
  if (idx.empty()) return -1;
  return idx[RefDiscreteSample(p.data(), (int)p.size(), lo, hi)];
}

// First n raw outputs of std::mt19937 seeded like HanabiGame does.
void RefMtWords(int32_t seed, int n, uint32_t* out) {
  std::mt19937 rng;
  rng.seed(seed);
  for (int i = 0; i < n; ++i) out[i] = rng();
}

// uniform_int_distribution(0, hi) as GetSampledStartPlayer uses it
// (hanabi_lib/hanabi_game.cc:138-145), fed from a seeded mt19937; returns the
// first n draws so the restated sampler can be checked.
void RefUniformInts(int32_t seed, int hi, int n, int32_t* out) {
  std::mt19937 rng;
  rng.seed(seed);
  for (int i = 0; i < n; ++i) {
    std::uniform_int_distribution<std::mt19937::result_type> dist(0, hi);
    out[i] = static_cast<int32_t>(dist(rng));
  }
}

}  // extern "C"
