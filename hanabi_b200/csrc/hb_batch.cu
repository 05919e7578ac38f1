// hb_batch.cu -- kernels and the batched C-ABI (include/pyhanabi.h, "Batch*" block).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 (see hanabi_b200/build.py).
// There is no CPU fallback: every entry point launches CUDA work on the batch's
// device and reports failure through its int status + BatchLastError().
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pyhanabi.h"
#include "hb_agents.cuh"
#include "hb_engine.cuh"
#include "hb_policy_rng.h"
#include "hb_spec.h"

namespace hb {

enum Mode { kModeStep = 0, kModeReset = 1, kModeObserve = 2 };
#ifndef HB_ITEMS
#define HB_ITEMS 3
#endif
// Serial path (initial deals of up to 10 cards), phase B2: 1 = one thread per game draws the
// cards one after the other, 0 = one thread per (game, draw) decodes its card from the game's
// draw indices (two more CTA barriers), 2 = the former for deals of 10+ cards, the latter below
// (measured: Full 2p 1.6 % faster with the chain, Small 2p / 4p 1.2 % / 3 % faster with the decode).
#ifndef HB_SERIAL_B2
#define HB_SERIAL_B2 2
#endif
#ifndef HB_EARLY_OUT
#define HB_EARLY_OUT 0  // measured: -1.2 % on Full 2p, +1.1 % on Full 4p
#endif
#ifndef HB_RESET_PREFETCH
#define HB_RESET_PREFETCH 1
#endif
// Games per CTA come with the rule-set view (V::kBlock); 768 threads are resident per SM.
#ifndef HB_RESIDENT
#define HB_RESIDENT 768  // measured again on the final kernel: 640 (90 registers, no spills) 0.5-0.7 % slower
#endif
constexpr int kResidentThreads = HB_RESIDENT;
// serial path: three words per queued draw, bounded so that the table stays within 640 draws
constexpr int serial_queue_cap(int block, int cards) {
  return block / 2 < 640 / cards ? block / 2 : 640 / cards;
}
constexpr int kStreamPad = 2;    // readable pad words after each packed bit stream
constexpr int kNumStats = 32;    // [0] episodes [1] sum score [2] sum length [3..28] score histogram

struct KArgs {
  Spec spec;
  uint4* state;
  uint32_t* mt;
  int64_t N;
  int64_t g_begin, g_end;  // range of games this launch covers (g_begin % 32 == 0)
  const int32_t* actions;
  const uint8_t* which;
  int32_t* rewards;
  uint8_t* dones;
  int32_t* scores;
  int32_t* cur_player;
  uint8_t* obs;
  int64_t obs_pitch;
  uint32_t* obs_bits;  // alternative to obs: bit-packed rows of packed_words 32-bit words
  // alternative to obs, may accompany obs_bits: the network's input rows, 0/1 elements of
  // net_dtype (0 float32, 1 bfloat16, 2 float16), net_pitch elements per row (multiple of 8)
  void* net_in;
  int net_dtype, net_chunks, net_lut_offset_words;
  uint32_t net_magic;
  int64_t net_pitch;
  int packed_words;
  uint8_t* mask;
  // optional fused random policy: after the step, a uniformly random legal move of the
  // new state (the generator of BatchSelectAction keyed (seed, step, game)); may alias
  // `actions` (every thread reads its action before it writes the next one)
  int32_t* random_actions;
  uint64_t policy_seed, policy_step;
  uint64_t policy_key;     // hbp::policy_step_key(policy_seed, policy_step), computed once on the host
  int64_t game_base;  // global index of this batch's game 0 (BatchSetGameIndexBase): the policy key's game
  unsigned long long* stats;
  uint4* hist;  // optional: the 4 non-deal moves before `last`, newest first (BatchEnableHistory)
  uint32_t* disc_log;  // optional, with hist: the discard pile in order, kDiscardLogWords words per game
  int auto_reset;
  int sampler_mode;
  int observer;
  int warp_smem_words;  // shared-memory words per warp: obs stream + mask stream + pads
  int obs_stream_words; // words reserved for the observation stream inside that
};

// One CTA = kBlock games = kBlock/32 warp tiles.  Phases separated by CTA barriers keep the
// warps of an SM inside the same stretch of code (the kernel is ~100 KB of SASS,
// far beyond the instruction caches) and let the RNG work be done by compacted lanes:
//   A  load state (+ prefetch two RNG words) -> apply move -> deal (the deal only READS its
//      two words: lazy regeneration) -> terminal; finished games and games with many pending
//      pairs enqueue a job in shared memory
//   B  all threads work the jobs off, one thread per RNG word pair: regenerate the pending
//      slots and draw the initial deals of the new games
//   C  install new games, store state, encode observation + legal mask into the
//      warp's packed bit stream, expand to bytes with 128-bit stores
// NET: the instantiation that also writes network-input rows (BatchStepEncodeNet /
// BatchObserveNet); kept out of the plain kernels so that their code stays as small as it was.
template <class V, int MODE, bool NET = false>
__global__ void __launch_bounds__(V::kBlock, (V::kResident ? V::kResident : kResidentThreads) / V::kBlock)
k_main(const __grid_constant__ KArgs a) {
  constexpr int kBlock = V::kBlock, kWarps = kBlock / 32;
  constexpr bool kJobs = V::kJobs;  // which re-deal machinery this rule set uses (hb_engine.cuh)
  // finished games (and flushes) a CTA works off through shared memory per launch; the rest
  // take the per-thread path
  constexpr int kQueueCap = kJobs ? kBlock / 2 : hb::serial_queue_cap(kBlock, V::kP * V::kH);
  extern __shared__ uint32_t smem[];
  __shared__ int s_stats[kNumStats];
  __shared__ int s_nq;
  __shared__ int s_q_tid[kQueueCap];
  __shared__ uint32_t s_q_meta[kQueueCap];  // jobs: r0 | npend << 10 | nfresh << 14 | slow << 31; serial: mt index
  __shared__ HB_JOBRESULT(V) s_res[kQueueCap];
  constexpr int kItems = HB_ITEMS;  // job path: RNG word pairs per thread in flight in phase B
  __shared__ uint8_t s_qb[kJobs ? kItems * kBlock : 4];  // job path: draw indices of one pass
  __shared__ uint32_t s_spread[32];
#if HB_EXPAND_LUT16
  __shared__ uint32_t s_nib[16];
  if (threadIdx.x < 16) s_nib[threadIdx.x] = nibble_to_bytes(threadIdx.x);
#else
  const uint32_t* const s_nib = nullptr;
#endif
  // byte -> its eight 16-bit elements: 4 KB of dynamic shared memory that only the launches
  // writing network-input rows ask for
  uint4* s_net_lut = reinterpret_cast<uint4*>(smem + a.net_lut_offset_words);  // behind this view's areas
  if (NET) {
    const uint32_t one = a.net_dtype == 2 ? 0x3c00u : 0x3f80u;  // float16 / bfloat16 (and float32 >> 16)
    for (int b = threadIdx.x; b < 256; b += kBlock) {
      uint4 e;
      e.x = ((b & 1) ? one : 0u) | ((b & 2) ? one << 16 : 0u);
      e.y = ((b & 4) ? one : 0u) | ((b & 8) ? one << 16 : 0u);
      e.z = ((b & 16) ? one : 0u) | ((b & 32) ? one << 16 : 0u);
      e.w = ((b & 64) ? one : 0u) | ((b & 128) ? one << 16 : 0u);
      s_net_lut[b] = e;
    }
  }
  const V v(a.spec);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t cta0 = a.g_begin + (int64_t)blockIdx.x * kBlock;
  const int64_t g = cta0 + threadIdx.x;
  const int64_t tile0 = g - lane;  // first game of this warp's tile
  const bool valid = g < a.g_end;
  const bool tile_live = tile0 < a.g_end;
  // lazy regeneration: pending pairs a game of this batch may carry (0 = off; compiled out
  // where the shape cannot use it)
  const int pend_cap = V::kPendCap > 0 ? a.spec.pend_cap : 0;
  const bool lazy = pend_cap > 0;
  if (threadIdx.x < 32) s_spread[threadIdx.x] = spread_entry(threadIdx.x, a.spec.C, a.spec.R);
  if (MODE != kModeObserve) {
    if (threadIdx.x < kNumStats) s_stats[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_nq = 0;
  }
  __syncthreads();
  HB_GAME(V) G;
  Prefetch pre;
  pre.at = -1;
  bool moved_any = false, restarted = false;
  int job = 0;  // 0 none, 1 re-deal job queued, 2 flush job queued
  int slot = -1;
  int reward = 0, done_out = 0, score_out = 0;
  uint32_t prev_last = 0;

  // ---------------------------------------------------------------- phase A
  if (tile_live) {
    load_game(a.state, a.N, valid ? g : tile0, v, G);  // idle lanes mirror a live game; never stored
    bool restart = false, want_flush = false;
    if (MODE == kModeStep) {
      const int uid = valid ? a.actions[g] : -1;
      // a play/discard is followed by a deal: start fetching its RNG words now
      if (valid && !G.done && uid >= 0 && uid < 2 * v.H() && G.deck_size > 0) {
        if (lazy) mt_load2(a.mt + g * (int64_t)kMtN, G.mti, pre);
        else mt_load5(a.mt + g * (int64_t)kMtN, G.mti, pre);
      }
      bool moved = false, term = false;
      int ep_len = 0;
      if (valid) {
        if (G.done || uid == -1) {
          // finished game without auto-reset, or action -1 ("no move for this game"):
          // the step is a no-op for it
          done_out = G.done;
          score_out = score_of(v, G);
        } else {
          if (!move_is_legal(v, G, uid, a.spec.max_moves)) {
            G.err = 1;  // illegal action: game unchanged, flagged
            score_out = score_of(v, G);
          } else {
            const int before = score_of(v, G);
            prev_last = G.last;
            apply_move(v, G, uid);
            score_out = score_of(v, G);
            reward = score_out - before;  // rl_env.py:363
            term = is_terminal(v, G);
            done_out = term ? 1 : 0;
            ep_len = G.eplen;
            moved = true;
            moved_any = true;
          }
        }
      }
      // the deal that follows a play/discard -- also for a game that just ended:
      // the reference deals before it looks at IsTerminal (rl_env.py:357-361)
      if (moved) deal_loop(v, a.spec, G, false, g, a.mt, a.sampler_mode, pre, pend_cap);
      if (term && !a.auto_reset) G.done = 1;
      restart = term && a.auto_reset;
      want_flush = lazy && !restart && G.pend >= pend_cap - 1;
#if HB_RESET_PREFETCH
      if constexpr (!kJobs)
        if (restart && a.spec.fast_reset) mt_prefetch_deal(a.mt + g * (int64_t)kMtN, G.mti, v.P() * v.H());
#endif
#if HB_EARLY_OUT
      // the step's results are final here (a re-deal does not touch them): write them
      // now instead of carrying three registers through phases B and C
      if (valid) {
        if (a.rewards) a.rewards[g] = reward;
        if (a.dones) a.dones[g] = (uint8_t)done_out;
        if (a.scores) a.scores[g] = score_out;
      }
#endif
      // episode statistics: warp aggregate -> CTA shared counters
      const unsigned tm = __ballot_sync(kFull, term);
      if (tm) {
        const int ssum = __reduce_add_sync(kFull, term ? score_out : 0);
        const int lsum = __reduce_add_sync(kFull, term ? ep_len : 0);
        if (lane == 0) {
          atomicAdd(&s_stats[0], __popc(tm));
          atomicAdd(&s_stats[1], ssum);
          atomicAdd(&s_stats[2], lsum);
        }
        if (term) atomicAdd(&s_stats[3 + (score_out > 25 ? 25 : score_out)], 1);
      }
    } else if (MODE == kModeReset) {
      restart = valid && (a.which == nullptr || a.which[g] != 0);
      if (restart) { G.done = 0; G.err = 0; }
    }
    if (MODE != kModeObserve) {
      restarted = restart;
      if constexpr (kJobs) {
        // ---- enqueue the RNG work: re-deals of finished games, flushes of pending pairs
        const bool want_job = (restart && a.spec.fast_reset) || want_flush;
        const unsigned qm = __ballot_sync(kFull, want_job);
        if (qm) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_nq, __popc(qm));
          base = __shfl_sync(kFull, base, 0);
          slot = base + __popc(qm & ((1u << lane) - 1u));
          if (want_job && slot < kQueueCap) {
            job = restart ? 1 : 2;
            int r0 = G.mti - 2 * G.pend;
            r0 += r0 < 0 ? kMtN : 0;
            const int nfresh = restart ? v.P() * v.H() : 0;
            s_q_tid[slot] = threadIdx.x;
            s_q_meta[slot] = (uint32_t)r0 | ((uint32_t)G.pend << 10) | ((uint32_t)nfresh << 14) |
                             (a.sampler_mode == 1 ? 0x80000000u : 0u);
#if HB_RESET_PREFETCH
            // the job's lines: [r0, r0 + 2T] and 397 slots ahead
            mt_prefetch_span(a.mt + g * (int64_t)kMtN, r0, 2 * (G.pend + nfresh) + 1);
            mt_prefetch_span(a.mt + g * (int64_t)kMtN, mt_wrap(r0 + kMtM), 2 * (G.pend + nfresh));
#endif
          }
        }
        // random start player / tiny decks / job-queue overflow: the per-thread path,
        // straight on the game's mt19937 block
        if ((want_job || restart) && job == 0) {
          if (lazy) flush_pending(a.mt + g * (int64_t)kMtN, G.mti, G.pend);
          if (restart) {
            Prefetch none;
            none.at = -1;
            deal_loop(v, a.spec, G, true, g, a.mt, a.sampler_mode, none);
          }
        }
      } else {
        // ---- serial path: finished games enqueue themselves
        bool queued = restart && a.spec.fast_reset;
        const unsigned qm = __ballot_sync(kFull, queued);
        if (qm) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_nq, __popc(qm));
          base = __shfl_sync(kFull, base, 0);
          slot = base + __popc(qm & ((1u << lane) - 1u));
          if (queued && slot >= kQueueCap) queued = false;  // queue full: general path below
          if (queued) {
            job = 1;
            s_q_tid[slot] = threadIdx.x;
            s_q_meta[slot] = (uint32_t)G.mti;
          }
        }
        // random start player / tiny decks / queue overflow: general per-thread path
        if (restart && !queued) {
          Prefetch none;
          none.at = -1;
          deal_loop(v, a.spec, G, true, g, a.mt, a.sampler_mode, none);
        }
      }
    }
  }

  // ---------------------------------------------------------------- phase B
  if (MODE != kModeObserve) {
    __syncthreads();
    const int nq = s_nq < kQueueCap ? s_nq : kQueueCap;
    if constexpr (kJobs) {
    if (nq > 0) {  // CTA-uniform
      // B1, one thread per (job, RNG word pair), kItems pairs per thread in flight at once:
      //   pairs [0, npend)               taken earlier by in-game deals: regenerate
      //   pairs [npend, npend + nfresh)  the initial deal of a new game: draw + regenerate
      // Draw d of a fresh deck has the known size deck_max - d, so its index q_d =
      // floor(u_d * size) does not depend on the earlier draws and is computed here, in
      // parallel, for every draw of every job.  All loads come first, then a barrier, then
      // the regenerated slots are stored (a pair's x[k+2] is the next pair's slot, so stores
      // wait for every load).
      // B2, one thread per re-deal job: the short serial chain deck -> q-th set bit -> deck
      // over the job's indices in shared memory (jobs side by side in the lanes of a warp).
      const int PH = v.P() * v.H(), H = v.H();
      const int tmax = PH + pend_cap;                   // pair slots per job (pend <= pend_cap)
      const int chunk_jobs = (kItems * kBlock) / tmax;  // jobs per pass
      for (int e_base = 0; e_base < nq; e_base += chunk_jobs) {  // CTA-uniform
        const int n_jobs = min(chunk_jobs, nq - e_base);
        const int n_slots = n_jobs * tmax;
        uint32_t new0[kItems], new1[kItems];
        uint32_t* xs[kItems];
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          const int k = u * kBlock + threadIdx.x;
          xs[u] = nullptr;
          new0[u] = new1[u] = 0u;
          if (k < n_slots) {
            const int ej = k / tmax, it = k - ej * tmax;
            const int e = e_base + ej;
            const uint32_t meta = s_q_meta[e];
            const int r0 = meta & 1023, npend = (meta >> 10) & 15, nfresh = (meta >> 14) & 63;
            if (it < npend + nfresh) {
              uint32_t* x = a.mt + (cta0 + s_q_tid[e]) * (int64_t)kMtN;
              int at = r0 + 2 * it;
              at -= at >= kMtN ? kMtN : 0;
              Prefetch f;
              mt_load5(x, at, f);
              new0[u] = mt_next_block_word(f.a, f.b, f.p);
              new1[u] = mt_next_block_word(f.b, f.c, f.q);
              xs[u] = x + at;
              const int d = it - npend;
              if (d >= 0) {
                bool amb = false;
                const int q = draw_index(v.deck_max() - d, mt_temper(f.a), mt_temper(f.b), amb);
                s_qb[ej * tmax + d] = (uint8_t)q;
                // a job with a draw that needs the exact fp64 sampler is left untouched here and
                // redone by its own thread on the general path (phase C)
                if (amb && a.sampler_mode != 2) atomicOr(&s_q_meta[e], 0x80000000u);
              }
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kItems; ++u) {
          if (xs[u] != nullptr) {
            const int k = u * kBlock + threadIdx.x;
            const int e = e_base + k / tmax;
            if (!(s_q_meta[e] >> 31)) {
              // (without a random start player the take index is even: a pair never wraps)
              __stcg(xs[u], new0[u]);
              __stcg(xs[u] + 1, new1[u]);
            }
          }
        }
        if ((int)threadIdx.x < n_jobs) {
          const int e = e_base + threadIdx.x;
          const uint32_t meta = s_q_meta[e];
          const int nfresh = (meta >> 14) & 63;
          if (!(meta >> 31) && nfresh > 0) {
            uint64_t deck = a.spec.full_deck;
            const uint8_t* qb = s_qb + threadIdx.x * tmax;
            typename V::Word hand = 0;
            int pl = 0, slot_i = 0;
#pragma unroll 1
            for (int d = 0; d < nfresh; ++d) {
              const int pos = select64(deck, (int)qb[d]);
              deck &= ~(1ull << pos);
              int color, rank;
              card_of_position(v, pos, color, rank);
              hand |= (typename V::Word)(uint32_t)(color | (rank << 3)) << (6 * slot_i);
              if (++slot_i == H) {  // player 0's hand first, then player 1's, ... (PlayerToDeal)
                s_res[e].cards[pl++] = hand;
                hand = 0;
                slot_i = 0;
              }
            }
            s_res[e].deck_lo = (uint32_t)deck;
            s_res[e].deck_hi = (uint32_t)(deck >> 32);
          }
        }
        if (e_base + chunk_jobs < nq) __syncthreads();  // s_qb is reused by the next pass
      }
    }
    } else {
    if (nq > 0) {  // CTA-uniform
      // B1: one thread per (queued game, word pair) takes the pair from the game's
      // stream: all loads first, barrier, then the regenerated slots are stored
      // (a pair's x[k+2] is the next pair's slot, so stores wait for every load)
      uint32_t* s_words = smem + (size_t)kWarps * a.warp_smem_words;
      const int pairs = v.P() * v.H();
      const int cap = kQueueCap * pairs;
      for (int k0 = 0; k0 < nq * pairs; k0 += kBlock) {  // CTA-uniform trip count
        const int k = k0 + threadIdx.x;
        const bool act = k < nq * pairs;
        Prefetch f;
        uint32_t* x = a.mt;
        int at = 0;
        if (act) {
          const int e = k / pairs, r = k - e * pairs;
          x = a.mt + (cta0 + s_q_tid[e]) * (int64_t)kMtN;
          at = mt_wrap((int)s_q_meta[e] + 2 * r);
          mt_load5(x, at, f);
          const uint32_t lo = mt_temper(f.a), hi = mt_temper(f.b);
          bool amb = false;
          const int q = draw_index(v.deck_max() - r, lo, hi, amb);  // deck size of draw r is known
          s_words[k] = (uint32_t)q | ((uint32_t)amb << 8);
          s_words[cap + k] = lo;
          s_words[2 * cap + k] = hi;
        }
        __syncthreads();
        if (act) {
          __stcg(x + at, mt_next_block_word(f.a, f.b, f.p));
          __stcg(x + mt_wrap(at + 1), mt_next_block_word(f.b, f.c, f.q));
        }
      }
      __syncthreads();
      constexpr bool kChain = HB_SERIAL_B2 == 1 || (HB_SERIAL_B2 == 2 && V::kP * V::kH >= 10);
      if (kChain) {
      // B2: one thread per queued game runs the select chain on the precomputed indices
      for (int i = threadIdx.x; i < nq; i += kBlock)
        reset_draw(v, a.spec, s_words + i * pairs, s_words + cap + i * pairs,
                   s_words + 2 * cap + i * pairs, a.sampler_mode, s_res[i]);
      } else {
      // B2: the draw indices are all in shared memory, so the cards follow without a serial
      // chain (Lehmer decode as above, through shared memory instead of shuffles): one thread
      // per (game, draw) walks back through the game's earlier indices, then ORs its card into
      // the game's hands and clears its deck bit.  A game with a draw that needs the exact
      // fp64 sampler (or the exact-only test mode) is redone serially afterwards.
      for (int i = threadIdx.x; i < nq; i += kBlock) {
#pragma unroll
        for (int p = 0; p < kMaxPlayers; ++p) s_res[i].cards[p] = 0;
        s_res[i].deck_lo = (uint32_t)a.spec.full_deck;
        s_res[i].deck_hi = (uint32_t)(a.spec.full_deck >> 32);
      }
      __syncthreads();
      for (int k = threadIdx.x; k < nq * pairs; k += kBlock) {
        const int e = k / pairs, r = k - e * pairs;
        const uint32_t* qa = s_words + e * pairs;
        int idx = (int)(qa[r] & 0xffu);
        for (int j = r - 1; j >= 0; --j) idx += ((int)(qa[j] & 0xffu) <= idx) ? 1 : 0;
        int color, rank;
        card_of_position(v, idx, color, rank);
        const int pl = r / v.H(), slot_i = r - pl * v.H();
        atomicOr(&s_res[e].cards[pl], (uint32_t)(color | (rank << 3)) << (6 * slot_i));
        if (idx < 32) atomicAnd(&s_res[e].deck_lo, ~(1u << idx));
        else atomicAnd(&s_res[e].deck_hi, ~(1u << (idx - 32)));
      }
      __syncthreads();
      for (int i = threadIdx.x; i < nq; i += kBlock) {
        bool slow = a.sampler_mode == 1;
        if (a.sampler_mode != 2)
          for (int r = 0; r < pairs; ++r) slow |= (s_words[i * pairs + r] >> 8) != 0;
        if (slow)
          reset_draw(v, a.spec, s_words + i * pairs, s_words + cap + i * pairs,
                     s_words + 2 * cap + i * pairs, a.sampler_mode, s_res[i]);
      }
      }
    }
    }
    __syncthreads();
    if (MODE == kModeStep && a.stats && threadIdx.x < kNumStats && s_stats[threadIdx.x] != 0)
      atomicAdd(&a.stats[threadIdx.x], (unsigned long long)s_stats[threadIdx.x]);
  }
  if (!tile_live) return;

  // ---------------------------------------------------------------- phase C
  if constexpr (kJobs) {
    if (job != 0) {
      if (!(s_q_meta[slot] >> 31)) {
        G.pend = 0;  // the job regenerated the pending slots
        if (job == 1) reset_apply(v, a.spec, G, s_res[slot]);
      } else {
        // a draw of this job needs the exact sampler (or the exact-only test mode): nothing was
        // stored for it; redo it on the per-thread path
        flush_pending(a.mt + g * (int64_t)kMtN, G.mti, G.pend);
        if (job == 1) {
          Prefetch none;
          none.at = -1;
          deal_loop(v, a.spec, G, true, g, a.mt, a.sampler_mode, none);
        }
      }
    }
  } else {
    if (job != 0) reset_apply(v, a.spec, G, s_res[slot]);
  }
  if (valid) {
#if !HB_EARLY_OUT
    if (MODE == kModeStep) {
      if (a.rewards) a.rewards[g] = reward;
      if (a.dones) a.dones[g] = (uint8_t)done_out;
      if (a.scores) a.scores[g] = score_out;
    }
#endif
    if (a.cur_player) a.cur_player[g] = G.cur;
    if (MODE != kModeObserve) {
      store_game(a.state, a.N, g, v, G);
    }
    if (MODE != kModeObserve && a.hist != nullptr) {
      // move history for the rule-based agents: a new game starts with none
      uint32_t* log = a.disc_log + g * (int64_t)kDiscardLogWords;
      if (restarted) {
        a.hist[g] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int w = 0; w < kDiscardLogWords; ++w) log[w] = 0u;
      } else if (moved_any) {
        const uint4 h = a.hist[g];
        a.hist[g] = make_uint4(prev_last, h.x, h.y, h.z);
        // a discarded or misplayed card goes on the pile: append its type to the ordered log
        // (the pile itself is kept as order-free thermometers; rl_env.py:416-418 wants the order)
        const uint32_t la = G.last;
        const int type = (la >> kLaTypeSh) & 3;
        if (type == kLaDiscard || (type == kLaPlay && !((la >> kLaScoredSh) & 1u))) {
          const int idx = __popcll(G.disc) - 1;  // its position on the pile
          const uint32_t card = ((la >> kLaColSh) & 7u) * (uint32_t)v.R() + ((la >> kLaRankSh) & 7u);
          const int bit = 5 * idx, w = bit >> 5, sh = bit & 31;
          if (w < kDiscardLogWords) {
            log[w] |= card << sh;
            if (sh > 27 && w + 1 < kDiscardLogWords) log[w + 1] |= card >> (32 - sh);
          }
        }
      }
    }
  }
  const int n_valid = (int)((a.g_end - tile0) < 32 ? (a.g_end - tile0) : 32);
  uint32_t* obs_stream = smem + (size_t)warp * a.warp_smem_words;
  uint32_t* mask_stream = obs_stream + a.obs_stream_words + kStreamPad;
  if (a.obs || a.obs_bits || NET) {
    const int observer = a.observer < 0 ? G.cur : a.observer;
    encode_obs(v, a.spec, G, observer, obs_stream, lane, s_spread,
               (a.obs_bits || NET) ? a.packed_words * 32 : a.spec.obs_size);
  }
  if (a.mask || a.random_actions) {
    const uint64_t legal = legal_bits(v, G);
    if (a.random_actions && valid) {
      const int n_legal = __popcll(legal);
      const uint64_t bits = hbp::policy_bits_keyed(a.policy_key, (uint64_t)(g + a.game_base));
      a.random_actions[g] = n_legal > 0 ? select64(legal, hbp::random_legal_rank(bits, n_legal)) : -1;
    }
    if (a.mask) pack_mask_stream(mask_stream, a.spec.max_moves, legal, valid, lane);
  }
  __syncwarp();
  // streaming (evict-first) stores for long rows, plain L2 stores for short ones (store_out)
  const bool stream_out = V::kStreamOut < 0 ? a.spec.obs_size >= 512 : V::kStreamOut != 0;
  if (a.obs) {
    if (stream_out)
      expand_stream<true>(obs_stream, a.spec.obs_size, a.obs + tile0 * a.obs_pitch, a.obs_pitch, n_valid,
                          lane, s_nib);
    else
      expand_stream<false>(obs_stream, a.spec.obs_size, a.obs + tile0 * a.obs_pitch, a.obs_pitch, n_valid,
                           lane, s_nib);
  }
  if (a.obs_bits)
    store_packed_rows(obs_stream, a.packed_words, a.obs_bits + tile0 * a.packed_words, n_valid, lane);
  if (NET) {
    const int esz = a.net_dtype == 0 ? 4 : 2;
    expand_net_rows(obs_stream, a.packed_words, static_cast<uint8_t*>(a.net_in) + tile0 * a.net_pitch * esz,
                    a.net_pitch, a.net_chunks, a.net_magic, n_valid, lane, s_net_lut, a.net_dtype == 0);
  }
  if (a.mask) {
    if (stream_out)
      expand_stream<true>(mask_stream, a.spec.max_moves, a.mask + tile0 * a.spec.max_moves,
                          a.spec.max_moves, n_valid, lane, s_nib);
    else
      expand_stream<false>(mask_stream, a.spec.max_moves, a.mask + tile0 * a.spec.max_moves,
                           a.spec.max_moves, n_valid, lane, s_nib);
  }
}

// std::mt19937::seed(value) per game followed by the first block regeneration
// (libstdc++ does it lazily on the first draw), so that the block holds outputs
// 0..623 (untempered) and the index starts at 0.  Two kernels, both with coalesced
// traffic on the 2.5 KB blocks:
//   k_seed_init   init_genrand: a serial recurrence per game, one thread per game; a warp's
//                 32 games go through a transposed shared-memory tile, 32 words at a time, so
//                 that every store instruction writes 128 contiguous bytes of one game's block
//   k_seed_twist  the first block regeneration, one warp per game in shared memory: slot k
//                 needs the old slots k, k+1 and slot k+397 (old for k < 227, new beyond), so
//                 the warp walks k upwards 32 slots at a time
__global__ void __launch_bounds__(128) k_seed_init(uint32_t* mt, const int32_t* seeds, int64_t N) {
  __shared__ uint32_t tile[4][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t g0 = ((int64_t)blockIdx.x * 4 + warp) * 32;  // first game of this warp
  if (g0 >= N) return;
  const int64_t g = g0 + lane;
  const int n_games = (int)(N - g0 < 32 ? N - g0 : 32);
  uint32_t x = g < N ? (uint32_t)seeds[g] : 0u;
  for (int c = 0; c < kMtN; c += 32) {  // 19 chunks of 32 words and one of 16
    const int len = kMtN - c < 32 ? kMtN - c : 32;
    for (int j = 0; j < len; ++j) {
      const int i = c + j;
      if (i > 0) x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
      tile[warp][lane][j] = x;
    }
    __syncwarp();
    for (int r = 0; r < n_games; ++r)
      if (lane < len) mt[(g0 + r) * (int64_t)kMtN + c + lane] = tile[warp][r][lane];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) k_seed_twist(uint32_t* mt, int64_t N) {
  __shared__ uint32_t blk[4][kMtN + 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t g = (int64_t)blockIdx.x * 4 + warp;
  if (g >= N) return;
  uint32_t* my = mt + g * (int64_t)kMtN;
  uint32_t* s = blk[warp];
  for (int k = lane; k < kMtN; k += 32) s[k] = my[k];
  __syncwarp();
  for (int k0 = 0; k0 < kMtN; k0 += 32) {
    const int k = k0 + lane;
    uint32_t v = 0;
    if (k < kMtN) v = mt_next_block_word(s[k], s[mt_wrap(k + 1)], s[mt_wrap(k + kMtM)]);
    // slot 623 reads the NEW slot 0 as its successor (written in the first round); every other
    // successor and every slot k + 397 < 624 is still old, every wrapped k + 397 already new
    __syncwarp();
    if (k < kMtN) s[k] = v;
    __syncwarp();
  }
  for (int k = lane; k < kMtN; k += 32) my[k] = s[k];
}

__global__ void k_blank(uint4* state, int64_t N, int ncols) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  for (int c = 0; c < ncols; ++c) state[(int64_t)c * N + g] = make_uint4(0, 0, 0, 0);
  state[(int64_t)(ncols - 1) * N + g] = make_uint4(0, 0, 0, 0);
}

// Packed state -> canonical unpacked dump (oracle/oracle_api.h layout).
template <class V>
__device__ void unpack_one(const KArgs& a, int64_t g, int32_t* d, HB_GAME(V)& G) {
  const V v(a.spec);
  load_game(a.state, a.N, g, v, G);
  for (int i = 0; i < kDumpLen; ++i) d[i] = 0;
  const int C = a.spec.C, R = a.spec.R;
  d[0] = G.cur; d[1] = G.info; d[2] = G.life; d[3] = G.deck_size;
  d[4] = score_of(v, G);
  d[5] = G.life < 1 ? 1 : (score_of(v, G) >= C * R ? 3 : (G.turns <= 0 ? 2 : 0));
  for (int c = 0; c < C; ++c) d[6 + c] = (G.fw >> (3 * c)) & 7;
  for (int p = 0; p < kMaxPlayers; ++p) {
    const int n = p < a.spec.P ? (int)((G.hw[p] >> 16) & 15) : 0;
    if (p < a.spec.P) d[11 + p] = n;
    for (int i = 0; i < 8; ++i) {
      const bool have = i < n;
      // (narrow hand words hold five slots: shifts stay below the word size)
      const bool slot = have && i < V::kH;
      const uint32_t f = slot ? (uint32_t)(G.cards[p] >> (6 * (i < V::kH ? i : 0))) : 0;
      d[16 + p * 8 + i] = have ? (int)(f & 7) * R + (int)((f >> 3) & 7) : -1;
      const int cm = slot ? (int)(G.cpl[p] >> (5 * (i < V::kH ? i : 0))) & 31 : 0;
      const int rm = slot ? (int)(G.rpl[p] >> (5 * (i < V::kH ? i : 0))) & 31 : 0;
      d[56 + p * 8 + i] = cm;
      d[96 + p * 8 + i] = rm;
      d[136 + p * 8 + i] = (have && ((G.hw[p] >> i) & 1)) ? 31 - __clz(cm) : -1;
      d[176 + p * 8 + i] = (have && ((G.hw[p] >> (8 + i)) & 1)) ? 31 - __clz(rm) : -1;
    }
  }
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      const int k = c * R + r;
      const int off = c * a.spec.per_color + DiscardRankOffset(r);
      const uint64_t bm = (1ull << InstancesOfRank(R, r)) - 1ull;
      d[216 + k] = __popcll((G.deck >> off) & bm);
      d[241 + k] = __popcll((G.disc >> off) & bm);
    }
}

template <class V>
__global__ void k_unpack(KArgs a, int32_t* out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  HB_GAME(V) G;
  unpack_one<V>(a, g, out + g * (int64_t)kDumpLen, G);
}

// One game's dump plus what the dump does not carry (BatchExportGame).
template <class V>
__global__ void k_export(KArgs a, int64_t g, int32_t* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  HB_GAME(V) G;
  unpack_one<V>(a, g, out, G);
  out[kDumpLen + 0] = G.turns;
  out[kDumpLen + 1] = (int32_t)G.last;
  const uint4 h = a.hist ? a.hist[g] : make_uint4(0, 0, 0, 0);
  out[kDumpLen + 2] = (int32_t)h.x; out[kDumpLen + 3] = (int32_t)h.y;
  out[kDumpLen + 4] = (int32_t)h.z; out[kDumpLen + 5] = (int32_t)h.w;
  out[kDumpLen + 6] = G.done;
  out[kDumpLen + 7] = G.eplen;
  for (int w = 0; w < kDiscardLogWords; ++w)
    out[kDumpLen + 8 + w] = a.disc_log ? (int32_t)a.disc_log[g * (int64_t)kDiscardLogWords + w] : 0;
  out[kDumpLen + 8 + kDiscardLogWords] = a.disc_log ? 1 : 0;
  out[kDumpLen + 9 + kDiscardLogWords] = 0;
}

// Per-game error flags (an illegal action was submitted since the last clear).
// The flag is bit `bit` of word `word` of column `col` (general layout: column P + 1, x, bit 31;
// compact layout: column P, y, bit 25).
__global__ void k_errors(uint4* state, int64_t N, int col, int word, int bit, uint8_t* flags, int clear,
                         unsigned long long* count) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool err = false;
  if (g < N) {
    uint32_t* x = &state[(int64_t)col * N + g].x + word;
    const uint32_t v = *x;
    err = (v >> bit) & 1u;
    if (flags) flags[g] = err ? 1 : 0;
    if (clear && err) *x = v & ~(1u << bit);
  }
  const unsigned m = __ballot_sync(kFull, err);
  if (count && m && (threadIdx.x & 31) == 0) atomicAdd(count, (unsigned long long)__popc(m));
}

// Canonical dump -> packed state (keeps each game's RNG position).  The dump does
// not carry turns_to_play / last action; they arrive in `extra[g*2+{0,1}]`
// (turns_to_play, packed last-action record or 0).
template <class V>
__global__ void k_pack(KArgs a, const int32_t* in, const int32_t* extra) {
  typedef typename V::Word W;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.N) return;
  const V v(a.spec);
  HB_GAME(V) G;
  load_game(a.state, a.N, g, v, G);
  const int32_t* d = in + g * (int64_t)kDumpLen;
  const int C = a.spec.C, R = a.spec.R;
  G.cur = d[0] < 0 ? 0 : d[0]; G.info = d[1]; G.life = d[2]; G.deck_size = d[3];
  G.fw = 0;
  for (int c = 0; c < C; ++c) G.fw |= (uint32_t)d[6 + c] << (3 * c);
  for (int p = 0; p < kMaxPlayers; ++p) {
    G.cards[p] = G.cpl[p] = G.rpl[p] = G.hw[p] = 0;
    if (p >= a.spec.P) continue;
    const int n = d[11 + p];
    G.hw[p] = (uint32_t)n << 16;
    for (int i = 0; i < n; ++i) {
      const int id = d[16 + p * 8 + i];
      G.cards[p] |= (W)(uint32_t)((id / R) | ((id % R) << 3)) << (6 * i);
      G.cpl[p] |= (W)(uint32_t)d[56 + p * 8 + i] << (5 * i);
      G.rpl[p] |= (W)(uint32_t)d[96 + p * 8 + i] << (5 * i);
      if (d[136 + p * 8 + i] >= 0) G.hw[p] |= 1u << i;
      if (d[176 + p * 8 + i] >= 0) G.hw[p] |= 1u << (8 + i);
    }
  }
  G.deck = 0; G.disc = 0;
  for (int c = 0; c < C; ++c)
    for (int r = 0; r < R; ++r) {
      const int k = c * R + r;
      const int off = c * a.spec.per_color + DiscardRankOffset(r);
      G.deck |= ((1ull << d[216 + k]) - 1ull) << off;
      G.disc |= ((1ull << d[241 + k]) - 1ull) << off;
    }
  G.turns = extra ? extra[g * 2] : a.spec.P;
  G.last = extra ? (uint32_t)extra[g * 2 + 1] : 0u;
  G.done = 0; G.err = 0; G.eplen = 0;
  G.done = is_terminal(v, G) ? 1 : 0;
  store_game(a.state, a.N, g, v, G);
  if (a.hist) a.hist[g] = make_uint4(0, 0, 0, 0);
  // (an injected pile has no order: the log restarts empty and stays marked valid only for
  //  cards discarded from here on -- BatchExportGame reports it invalid when the counts disagree)
  if (a.disc_log)
    for (int w = 0; w < kDiscardLogWords; ++w) a.disc_log[g * (int64_t)kDiscardLogWords + w] = 0u;
}

// Test hook: the deal sampler alone on caller-provided decks (2-bit counts per
// card type) and RNG words.  Returns the card type colour * R + rank.
__global__ void k_deal_test(Spec spec, const uint64_t* decks, const int32_t* sizes,
                            const uint32_t* lo, const uint32_t* hi, int n, int mode,
                            int32_t* out_type, int32_t* out_amb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const DView v(spec);
  uint64_t deck = 0;
  for (int c = 0; c < spec.C; ++c)
    for (int r = 0; r < spec.R; ++r) {
      const int cnt = (int)((decks[i] >> (2 * (c * spec.R + r))) & 3);
      deck |= ((1ull << cnt) - 1ull) << (c * spec.per_color + DiscardRankOffset(r));
    }
  bool amb = false;
  int pos;
  if (deck_types(spec, deck) < 2) {
    pos = __ffsll((long long)deck) - 1;
  } else {
    pos = pick_fast(deck, sizes[i], lo[i], hi[i], amb);
    if (mode == 1 || (mode == 0 && amb)) pos = pick_exact(v, deck, sizes[i], lo[i], hi[i]);
  }
  int color, rank;
  card_of_position(v, pos, color, rank);
  out_type[i] = color * spec.R + rank;
  if (out_amb) out_amb[i] = amb ? 1 : 0;
}

}  // namespace hb

// ============================================================== host side
namespace {

thread_local std::string g_last_error;
}  // namespace

// Shared with hb_policy.cu / hb_host.cu: records the text BatchLastError() returns.
int HbFail(const std::string& msg) {
  g_last_error = msg;
  return 1;
}

namespace {
int Fail(const std::string& msg) { return HbFail(msg); }
int CudaFail(const char* what, cudaError_t e) {
  return Fail(std::string(what) + ": " + cudaGetErrorString(e));
}
#define HB_CUDA(expr)                                   \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return CudaFail(#expr, e__); \
  } while (0)

struct Batch {
  hb::Spec spec;
  int device = 0;
  int64_t N = 0;
  int ncols = 0;
  uint4* state = nullptr;
  uint32_t* mt = nullptr;
  unsigned long long* stats = nullptr;      // active stats buffer
  unsigned long long* own_stats = nullptr;  // library-owned fallback
  uint4* hist = nullptr;                    // move history ring, BatchEnableHistory
  uint32_t* disc_log = nullptr;             // ordered discard pile, allocated with hist
  int sampler_mode = 0;
  int force_generic = 0;
  int64_t game_base = 0;  // global index of game 0 (sharded batches), keys the fused policy
  int warp_smem_words = 0;
  int packed_words = 0, obs_stream_words = 0;
  size_t smem_pad = 0;  // experiments: extra dynamic shared memory per CTA (fewer resident CTAs)
  // L2 residency of the packed game state (BatchSetL2Residency): launches carry an
  // access-policy window over `state` so that it stays in L2 between steps
  size_t l2_window_bytes = 0;
  float l2_hit_ratio = 0.f;
  // pinned/pipelined host path
  cudaStream_t h_streams[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t h_events[3] = {nullptr, nullptr, nullptr};
  // BatchStepEncodeHostRange: completion events of the most recent ranges
  static constexpr int kRangeSlots = 64;
  cudaEvent_t range_events[kRangeSlots] = {};
  int64_t range_first[kRangeSlots] = {};
  int range_next = 0;
  // packed link (BatchSetHostLink): byte rows requested by the caller travel bit-packed and are
  // expanded by the host worker pool when the range is waited for
  int host_link = -1;  // -1: HANABI_B200_HOST_LINK or the default, 0: byte rows on the link, 1: packed rows
  uint8_t* h_stage = nullptr;  // pinned host staging of the packed rows, N rows + pad
  int64_t range_count[kRangeSlots] = {};
  uint8_t* range_obs[kRangeSlots] = {};    // non-null: the slot's rows still have to be expanded
  int64_t range_pitch[kRangeSlots] = {};
  // device staging for the host-buffer entry point, allocated on first use
  int32_t* d_actions = nullptr;
  int32_t* d_rewards = nullptr;
  int32_t* d_scores = nullptr;
  int32_t* d_cur = nullptr;
  uint8_t* d_dones = nullptr;
  uint8_t* d_obs = nullptr;
  uint8_t* d_mask = nullptr;
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

const char* FindParam(int n, const char** kv, const char* key) {
  const char* v = nullptr;
  for (int i = 0; i + 1 < n; i += 2)
    if (kv[i] && std::strcmp(kv[i], key) == 0) v = kv[i + 1];
  return v;
}
int IntParam(int n, const char** kv, const char* key, int dflt) {
  const char* v = FindParam(n, kv, key);  // ParameterValue<int>: std::stoi (hanabi_lib/util.cc:37-47)
  return v ? (int)std::strtol(v, nullptr, 10) : dflt;
}
bool BoolParam(int n, const char** kv, const char* key, bool dflt) {
  const char* v = FindParam(n, kv, key);  // ParameterValue<bool> (hanabi_lib/util.cc:74-86)
  if (!v) return dflt;
  return !std::strcmp(v, "1") || !std::strcmp(v, "true") || !std::strcmp(v, "True");
}

// HanabiGame ctor defaults, hanabi_lib/hanabi_game.cc:20-47.
int ParseSpec(int n, const char** kv, hb::Spec* s) {
  std::memset(s, 0, sizeof(*s));
  s->P = IntParam(n, kv, "players", 2);
  s->C = IntParam(n, kv, "colors", hb::kMaxColors);
  s->R = IntParam(n, kv, "ranks", hb::kMaxRanks);
  s->H = IntParam(n, kv, "hand_size", s->P < 4 ? 5 : 4);
  s->IT = IntParam(n, kv, "max_information_tokens", 8);
  s->LT = IntParam(n, kv, "max_life_tokens", 3);
  s->random_start = BoolParam(n, kv, "random_start_player", false) ? 1 : 0;
  s->obs_type = IntParam(n, kv, "observation_type", hb::kCardKnowledge);
  return hb::FinishSpec(s);
}

hb::KArgs BaseArgs(const Batch* b) {
  hb::KArgs a;
  std::memset(&a, 0, sizeof(a));
  a.spec = b->spec;
  a.state = b->state;
  a.mt = b->mt;
  a.N = b->N;
  a.g_begin = 0;
  a.g_end = b->N;
  a.stats = b->stats;
  a.hist = b->hist;
  a.disc_log = b->disc_log;
  a.sampler_mode = b->sampler_mode;
  a.game_base = b->game_base;
  a.observer = -1;
  a.warp_smem_words = b->warp_smem_words;
  a.packed_words = b->packed_words;
  a.obs_stream_words = b->obs_stream_words;
  return a;
}

// Dynamic shared memory of k_main<V>: one packed-stream area per warp, the (index, lo, hi)
// table of the serial re-deal path, and -- for the launches that write network-input rows
// -- the 4 KB element table behind them.
template <class V>
size_t ViewSmemWords(const Batch* b) {
  const size_t words = (size_t)(V::kBlock / 32) * b->warp_smem_words +
                       (V::kJobs ? 0 : (size_t)hb::serial_queue_cap(V::kBlock, V::kP * V::kH) * 3 * b->spec.P * b->spec.H);
  return (words + 3) & ~(size_t)3;
}

template <class V, int MODE, bool NET>
cudaError_t LaunchKernel(const Batch* b, const hb::KArgs& a_in, cudaStream_t stream) {
  hb::KArgs a = a_in;
  const unsigned grid = (unsigned)((a.g_end - a.g_begin + V::kBlock - 1) / V::kBlock);
  static size_t allowed[16] = {0};  // per device: dynamic shared memory already opted in
  const int dev = b->device & 15;
  const size_t base_words = ViewSmemWords<V>(b);
  a.net_lut_offset_words = (int)base_words;
  const size_t smem_bytes = base_words * sizeof(uint32_t) + (NET ? 256 * sizeof(uint4) : 0) + b->smem_pad;
  if (smem_bytes > allowed[dev]) {
    cudaError_t e = cudaFuncSetAttribute(hb::k_main<V, MODE, NET>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes);
    if (e != cudaSuccess) return e;
    allowed[dev] = smem_bytes;
  }
  if (b->l2_window_bytes == 0) {
    hb::k_main<V, MODE, NET><<<grid, V::kBlock, smem_bytes, stream>>>(a);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(V::kBlock);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  std::memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeAccessPolicyWindow;
  attr.val.accessPolicyWindow.base_ptr = b->state;
  attr.val.accessPolicyWindow.num_bytes = b->l2_window_bytes;
  attr.val.accessPolicyWindow.hitRatio = b->l2_hit_ratio;
  attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, hb::k_main<V, MODE, NET>, a);
}

template <class V, int MODE>
cudaError_t LaunchOne(const Batch* b, const hb::KArgs& a, cudaStream_t stream) {
  if (a.net_in != nullptr && MODE != hb::kModeReset) return LaunchKernel<V, MODE, MODE != hb::kModeReset>(b, a, stream);
  return LaunchKernel<V, MODE, false>(b, a, stream);
}

template <int MODE>
cudaError_t Launch(const Batch* b, const hb::KArgs& a, cudaStream_t stream) {
  const hb::Spec& s = b->spec;
  if (!b->force_generic) {
#define HB_CASE(P_, H_, C_, R_, IT_, LT_)                                                      \
  if (s.P == P_ && s.H == H_ && s.C == C_ && s.R == R_ && s.IT == IT_ && s.LT == LT_)          \
    return LaunchOne<hb::SView<P_, H_, C_, R_, IT_, LT_>, MODE>(b, a, stream);
    HB_CASE(2, 5, 5, 5, 8, 3)  // Hanabi-Full 2p
    HB_CASE(3, 5, 5, 5, 8, 3)  // Hanabi-Full 3p
    HB_CASE(4, 4, 5, 5, 8, 3)  // Hanabi-Full 4p
    HB_CASE(5, 4, 5, 5, 8, 3)  // Hanabi-Full 5p
    HB_CASE(2, 2, 2, 5, 3, 1)  // Hanabi-Small 2p
    HB_CASE(4, 2, 2, 5, 3, 1)  // Hanabi-Small 4p
    HB_CASE(2, 2, 1, 5, 3, 1)  // Hanabi-Very-Small 2p
#undef HB_CASE
  }
  if (s.H > hb::kMaxHandNarrow) return LaunchOne<hb::WView, MODE>(b, a, stream);  // hands of 6-8 cards
  return LaunchOne<hb::DView, MODE>(b, a, stream);
}

Batch* AsBatch(void* p) { return static_cast<Batch*>(p); }

}  // namespace

// ============================================================== C ABI
extern "C" {

const char* BatchLastError(void) { return g_last_error.c_str(); }

int NewBatch(pyhanabi_batch_t* out, int list_length, const char** param_list, int num_games,
             int device, const int32_t* seeds) {
  if (!out) return Fail("NewBatch: null handle");
  out->batch = nullptr;
  if (num_games <= 0) return Fail("NewBatch: num_games must be positive");
  if (!seeds) return Fail("NewBatch: seeds must point to num_games int32 values (host memory)");
  Batch* b = new Batch;
  int rc = ParseSpec(list_length, param_list, &b->spec);
  if (rc != 0) {
    delete b;
    return Fail(std::string("NewBatch: ") + hb::SpecError(rc));
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    delete b;
    return Fail(std::string("NewBatch: no CUDA device available (") + cudaGetErrorString(e) +
                "); the batched engine has no CPU fallback");
  }
  if (device < 0 || device >= count) {
    delete b;
    return Fail("NewBatch: device index out of range");
  }
  b->device = device;
  b->N = num_games;
  b->ncols = b->spec.H > hb::kMaxHandNarrow ? hb::state_columns_wide(b->spec.P) : hb::state_columns(b->spec.P);
  b->packed_words = 4 * ((b->spec.obs_size + 127) / 128);
  b->obs_stream_words = b->spec.obs_size > 32 * b->packed_words ? b->spec.obs_size : 32 * b->packed_words;
  b->obs_stream_words = (b->obs_stream_words + 3) & ~3;
  // every warp's region starts 16-byte aligned (the packed rows are read as uint4)
  b->warp_smem_words = (b->obs_stream_words + b->spec.max_moves + 2 * hb::kStreamPad + 3) & ~3;
  // (the CTA's total dynamic shared memory depends on the view's CTA size: ViewSmemWords)
  if (const char* pad = std::getenv("HANABI_B200_SMEM_PAD"))  // experiments: fewer resident CTAs
    b->smem_pad = (size_t)std::atoll(pad);
  DeviceGuard guard(device);
  if (!guard.ok) { delete b; return Fail("NewBatch: cudaSetDevice failed"); }
  int32_t* d_seeds = nullptr;
  auto cleanup = [&]() {
    cudaFree(b->state); cudaFree(b->mt); cudaFree(b->own_stats); cudaFree(d_seeds);
    delete b;
  };
#define HB_TRY(expr)                                                           \
  do {                                                                         \
    cudaError_t e__ = (expr);                                                  \
    if (e__ != cudaSuccess) { cleanup(); return CudaFail("NewBatch: " #expr, e__); } \
  } while (0)
  HB_TRY(cudaMalloc(&b->state, sizeof(uint4) * (size_t)b->ncols * (size_t)b->N));
  HB_TRY(cudaMalloc(&b->mt, sizeof(uint32_t) * (size_t)hb::kMtN * (size_t)b->N));
  HB_TRY(cudaMalloc(&b->own_stats, sizeof(unsigned long long) * hb::kNumStats));
  HB_TRY(cudaMemset(b->own_stats, 0, sizeof(unsigned long long) * hb::kNumStats));
  b->stats = b->own_stats;
  HB_TRY(cudaMalloc(&d_seeds, sizeof(int32_t) * (size_t)b->N));
  HB_TRY(cudaMemcpy(d_seeds, seeds, sizeof(int32_t) * (size_t)b->N, cudaMemcpyHostToDevice));
  const unsigned grid = (unsigned)((b->N + 127) / 128);
  hb::k_seed_init<<<grid, 128>>>(b->mt, d_seeds, b->N);
  hb::k_seed_twist<<<(unsigned)((b->N + 3) / 4), 128>>>(b->mt, b->N);
  hb::k_blank<<<grid, 128>>>(b->state, b->N, b->ncols);
  HB_TRY(cudaGetLastError());
  HB_TRY(cudaDeviceSynchronize());
  cudaFree(d_seeds);
  d_seeds = nullptr;
#undef HB_TRY
  out->batch = b;
  if (const char* env = std::getenv("HANABI_B200_L2_RESIDENT")) {
    const long long mb = std::atoll(env);
    if (mb != 0 && BatchSetL2Residency(out, mb > 0 ? mb << 20 : 0, nullptr) != 0) {
      DeleteBatch(out);
      return 1;
    }
  }
  return 0;
}

void DeleteBatch(pyhanabi_batch_t* h) {
  if (!h || !h->batch) return;
  Batch* b = AsBatch(h->batch);
  DeviceGuard guard(b->device);
  for (int i = 0; i < 3; ++i) {
    if (b->h_streams[i]) cudaStreamDestroy(b->h_streams[i]);
    if (b->h_events[i]) cudaEventDestroy(b->h_events[i]);
  }
  for (int i = 0; i < Batch::kRangeSlots; ++i)
    if (b->range_events[i]) cudaEventDestroy(b->range_events[i]);
  cudaFree(b->d_actions); cudaFree(b->d_rewards); cudaFree(b->d_scores); cudaFree(b->d_cur);
  cudaFree(b->d_dones); cudaFree(b->d_obs); cudaFree(b->d_mask);
  if (b->h_stage) cudaFreeHost(b->h_stage);
  cudaFree(b->state);
  cudaFree(b->mt);
  cudaFree(b->own_stats);
  cudaFree(b->hist);
  cudaFree(b->disc_log);
  delete b;
  h->batch = nullptr;
}

#define HB_BATCH(h)                                              \
  if (!(h) || !(h)->batch) return Fail("null batch handle");     \
  Batch* b = AsBatch((h)->batch);                                \
  DeviceGuard guard__(b->device);                                \
  if (!guard__.ok) return Fail("cudaSetDevice failed")

int BatchObservationSize(pyhanabi_batch_t* h) { return (h && h->batch) ? AsBatch(h->batch)->spec.obs_size : -1; }
int BatchMaxMoves(pyhanabi_batch_t* h) { return (h && h->batch) ? AsBatch(h->batch)->spec.max_moves : -1; }
int BatchNumGames(pyhanabi_batch_t* h) { return (h && h->batch) ? (int)AsBatch(h->batch)->N : -1; }
int BatchNumPlayers(pyhanabi_batch_t* h) { return (h && h->batch) ? AsBatch(h->batch)->spec.P : -1; }
int BatchDevice(pyhanabi_batch_t* h) { return (h && h->batch) ? AsBatch(h->batch)->device : -1; }

int BatchReset(pyhanabi_batch_t* h, const uint8_t* which, void* stream) {
  HB_BATCH(h);
  hb::KArgs a = BaseArgs(b);
  a.which = which;
  HB_CUDA(Launch<hb::kModeReset>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchStep(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards, uint8_t* dones,
              int32_t* scores, int auto_reset, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStep: actions is null");
  hb::KArgs a = BaseArgs(b);
  a.actions = actions; a.rewards = rewards; a.dones = dones; a.scores = scores;
  a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchEncode(pyhanabi_batch_t* h, int observer, uint8_t* obs, int64_t row_pitch, void* stream) {
  HB_BATCH(h);
  if (!obs) return Fail("BatchEncode: obs is null");
  if (row_pitch < b->spec.obs_size) return Fail("BatchEncode: row_pitch smaller than the observation");
  if (observer >= b->spec.P) return Fail("BatchEncode: observer out of range");
  hb::KArgs a = BaseArgs(b);
  a.obs = obs; a.obs_pitch = row_pitch; a.observer = observer;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchLegalMask(pyhanabi_batch_t* h, uint8_t* mask, void* stream) {
  HB_BATCH(h);
  if (!mask) return Fail("BatchLegalMask: mask is null");
  hb::KArgs a = BaseArgs(b);
  a.mask = mask;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchCurPlayer(pyhanabi_batch_t* h, int32_t* cur_player, void* stream) {
  HB_BATCH(h);
  if (!cur_player) return Fail("BatchCurPlayer: cur_player is null");
  hb::KArgs a = BaseArgs(b);
  a.cur_player = cur_player;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

// Observation + legal mask + current player of the state as it stands (no move).
int BatchObserve(pyhanabi_batch_t* h, int observer, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                 int32_t* cur_player, void* stream) {
  HB_BATCH(h);
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchObserve: row_pitch smaller than the observation");
  if (observer >= b->spec.P) return Fail("BatchObserve: observer out of range");
  hb::KArgs a = BaseArgs(b);
  a.obs = obs; a.obs_pitch = row_pitch; a.observer = observer; a.mask = mask; a.cur_player = cur_player;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchStepEncode(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards, uint8_t* dones,
                    uint8_t* obs, int64_t row_pitch, uint8_t* mask, int32_t* cur_player,
                    int auto_reset, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncode: actions is null");
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchStepEncode: row_pitch smaller than the observation");
  hb::KArgs a = BaseArgs(b);
  a.actions = actions; a.rewards = rewards; a.dones = dones;
  a.obs = obs; a.obs_pitch = row_pitch; a.mask = mask; a.cur_player = cur_player;
  a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

// Same as BatchStepEncode plus the post-move score of every game (the score of
// the finished episode when done[i] is set).
int BatchStepEncodeScores(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards,
                          uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                          uint8_t* mask, int32_t* cur_player, int auto_reset, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodeScores: actions is null");
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchStepEncodeScores: row_pitch smaller than the observation");
  hb::KArgs a = BaseArgs(b);
  a.actions = actions; a.rewards = rewards; a.dones = dones; a.scores = scores;
  a.obs = obs; a.obs_pitch = row_pitch; a.mask = mask; a.cur_player = cur_player;
  a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

// BatchStepEncodeScores plus a fused uniform-random policy: after the move,
// actions[i] is overwritten with a uniformly random legal move of game i's new
// state (-1 when there is none) -- exactly what BatchSelectAction(NULL, ..., epsilon
// >= 1, seed, step) would choose from the mask this call writes, without the second
// kernel and without reading the mask back.  `random_actions` may be `actions`
// itself (in-place rollout loop) or a separate buffer.
int BatchStepEncodeRandom(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards,
                          uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                          uint8_t* mask, int32_t* cur_player, int auto_reset,
                          int32_t* random_actions, uint64_t seed, uint64_t step, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodeRandom: actions is null");
  if (!random_actions) return Fail("BatchStepEncodeRandom: random_actions is null");
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchStepEncodeRandom: row_pitch smaller than the observation");
  hb::KArgs a = BaseArgs(b);
  a.actions = actions; a.rewards = rewards; a.dones = dones; a.scores = scores;
  a.obs = obs; a.obs_pitch = row_pitch; a.mask = mask; a.cur_player = cur_player;
  a.auto_reset = auto_reset;
  a.random_actions = random_actions; a.policy_seed = seed; a.policy_step = step;
  a.policy_key = hbp::policy_step_key(seed, step);
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

// A batch that is one shard of a larger set of games (one process per GPU, several
// shards per GPU) tells the engine the global index of its game 0: the fused random
// policy is keyed (seed, step, base + i), so a sharded rollout reproduces the
// unsharded one game for game (the games themselves never interact:
// hanabi_lib/hanabi_game.h:114, one RNG per game).
int BatchSetGameIndexBase(pyhanabi_batch_t* h, int64_t first_game_index) {
  HB_BATCH(h);
  if (first_game_index < 0) return Fail("BatchSetGameIndexBase: negative index");
  b->game_base = first_game_index;
  return 0;
}

// A uniformly random legal move of every game's CURRENT state (no move applied):
// the draw BatchStepEncodeRandom would leave behind, keyed (seed, step, base + i) --
// the first actions of a fused random rollout.
int BatchRandomLegalActions(pyhanabi_batch_t* h, uint64_t seed, uint64_t step, int32_t* actions,
                            void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchRandomLegalActions: actions is null");
  hb::KArgs a = BaseArgs(b);
  a.random_actions = actions; a.policy_seed = seed; a.policy_step = step;
  a.policy_key = hbp::policy_step_key(seed, step);
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

// The fused hot path with HOST buffers (pinned memory recommended): the games
// are cut into chunks that flow through three internal streams so that the
// action upload, the kernel and the result download of neighbouring chunks
// overlap.  Synchronous: the host buffers are complete on return.  `stream` is
// the stream whose earlier work on this batch must finish first (may be NULL).
// Staging buffers and internal streams of the host-buffer entry points (first use).
static int HostSetup(Batch* b) {
  if (b->h_streams[0]) return 0;
  const size_t N = (size_t)b->N;
  const int D = b->spec.obs_size, A = b->spec.max_moves;
  for (int i = 0; i < 3; ++i) {
    HB_CUDA(cudaStreamCreateWithFlags(&b->h_streams[i], cudaStreamNonBlocking));
    HB_CUDA(cudaEventCreateWithFlags(&b->h_events[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < Batch::kRangeSlots; ++i) {
    HB_CUDA(cudaEventCreateWithFlags(&b->range_events[i], cudaEventDisableTiming));
    b->range_first[i] = -1;
  }
  HB_CUDA(cudaMalloc(&b->d_actions, N * sizeof(int32_t)));
  HB_CUDA(cudaMalloc(&b->d_rewards, N * sizeof(int32_t)));
  HB_CUDA(cudaMalloc(&b->d_scores, N * sizeof(int32_t)));
  HB_CUDA(cudaMalloc(&b->d_cur, N * sizeof(int32_t)));
  HB_CUDA(cudaMalloc(&b->d_dones, N));
  HB_CUDA(cudaMalloc(&b->d_obs, N * (size_t)D));
  HB_CUDA(cudaMalloc(&b->d_mask, N * (size_t)A));
  return 0;
}

// Packed link: whether byte rows asked for by the caller cross PCIe bit-packed (and are expanded
// by the host worker pool, HostExpandPacked) or as the bytes the kernel's own expansion writes.
// Default: packed -- the link carries 5.5x fewer bytes per Full-2p game (131 B instead of 720 B),
// and the expansion runs on the pool while the next ranges are on the link.
static bool UsePackedLink(Batch* b) {
  if (b->host_link < 0) {
    const char* env = std::getenv("HANABI_B200_HOST_LINK");
    b->host_link = env ? (std::strcmp(env, "bytes") == 0 || std::strcmp(env, "0") == 0 ? 0 : 1) : 1;
  }
  return b->host_link == 1;
}
static int LinkSetup(Batch* b) {
  if (b->h_stage) return 0;
  const size_t row = (size_t)b->packed_words * sizeof(uint32_t);
  HB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->h_stage), (size_t)b->N * row + 64, cudaHostAllocDefault));
  return 0;
}
// Rows [g0, g1) of the pinned staging -> the caller's byte rows.
static int ExpandRange(Batch* b, size_t g0, size_t g1, uint8_t* obs, int64_t row_pitch) {
  const size_t row = (size_t)b->packed_words * sizeof(uint32_t);
  return HostExpandPacked(reinterpret_cast<const uint32_t*>(b->h_stage + g0 * row), b->packed_words,
                          b->spec.obs_size, (int64_t)(g1 - g0), obs + g0 * (size_t)row_pitch, row_pitch, 0);
}

// Enqueues games [g0, g1) on stream s: action upload, the fused kernel, result download.
// The host pointers are the bases of the whole arrays (indexed by game).
static int EnqueueHostRange(Batch* b, size_t g0, size_t g1, const int32_t* actions, int32_t* rewards,
                            uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                            uint32_t* obs_bits, uint8_t* mask, int32_t* cur_player, int auto_reset,
                            cudaStream_t s, bool packed_link = false) {
  const int D = b->spec.obs_size, A = b->spec.max_moves;
  const size_t n = g1 - g0;
  if (packed_link && obs) {
    // the caller's byte rows are produced on the host from the packed rows (ExpandRange)
    obs = nullptr;
    obs_bits = reinterpret_cast<uint32_t*>(b->h_stage);
  }
  HB_CUDA(cudaMemcpyAsync(b->d_actions + g0, actions + g0, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  hb::KArgs a = BaseArgs(b);
  a.g_begin = (int64_t)g0; a.g_end = (int64_t)g1;
  a.actions = b->d_actions; a.rewards = b->d_rewards; a.dones = b->d_dones; a.scores = b->d_scores;
  a.obs = obs ? b->d_obs : nullptr; a.obs_pitch = D; a.mask = mask ? b->d_mask : nullptr;
  a.obs_bits = (!obs && obs_bits) ? reinterpret_cast<uint32_t*>(b->d_obs) : nullptr;  // staging reused
  a.cur_player = b->d_cur; a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, s));
  if (a.obs_bits) {
    const size_t row = (size_t)b->packed_words * sizeof(uint32_t);
    HB_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(obs_bits) + g0 * row, b->d_obs + g0 * row, n * row,
                            cudaMemcpyDeviceToHost, s));
  }
  if (obs) {
    if (row_pitch == D)
      HB_CUDA(cudaMemcpyAsync(obs + g0 * (size_t)D, b->d_obs + g0 * (size_t)D, n * (size_t)D, cudaMemcpyDeviceToHost, s));
    else
      HB_CUDA(cudaMemcpy2DAsync(obs + g0 * (size_t)row_pitch, (size_t)row_pitch, b->d_obs + g0 * (size_t)D,
                                (size_t)D, (size_t)D, n, cudaMemcpyDeviceToHost, s));
  }
  if (mask) HB_CUDA(cudaMemcpyAsync(mask + g0 * (size_t)A, b->d_mask + g0 * (size_t)A, n * (size_t)A, cudaMemcpyDeviceToHost, s));
  if (rewards) HB_CUDA(cudaMemcpyAsync(rewards + g0, b->d_rewards + g0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (dones) HB_CUDA(cudaMemcpyAsync(dones + g0, b->d_dones + g0, n, cudaMemcpyDeviceToHost, s));
  if (scores) HB_CUDA(cudaMemcpyAsync(scores + g0, b->d_scores + g0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (cur_player) HB_CUDA(cudaMemcpyAsync(cur_player + g0, b->d_cur + g0, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  return 0;
}

static int StepHost(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards, uint8_t* dones,
                    int32_t* scores, uint8_t* obs, int64_t row_pitch, uint32_t* obs_bits,
                    uint8_t* mask, int32_t* cur_player, int auto_reset, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodeHost: actions is null");
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchStepEncodeHost: row_pitch smaller than the observation");
  const size_t N = (size_t)b->N;
  if (HostSetup(b)) return 1;
  HB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  const bool link = obs && UsePackedLink(b);
  if (link && LinkSetup(b)) return 1;
  int nchunks = (int)((N + 32767) / 32768);
  if (nchunks > 12) nchunks = 12;
  if (nchunks < 1) nchunks = 1;
  size_t cs = (N + nchunks - 1) / nchunks;
  cs = (cs + 127) / 128 * 128;
  int c = 0;
  cudaEvent_t* ev = b->range_events;  // (synchronous call: the ranged slots are free to borrow)
  if (BatchHostWait(h, -1)) return 1;  // ranged work still in flight (and its pending expansions) first
  for (size_t g0 = 0; g0 < N; g0 += cs, ++c) {
    const size_t g1 = g0 + cs < N ? g0 + cs : N;
    if (EnqueueHostRange(b, g0, g1, actions, rewards, dones, scores, obs, row_pitch, obs_bits, mask,
                         cur_player, auto_reset, b->h_streams[c % 3], link))
      return 1;
    if (link) HB_CUDA(cudaEventRecord(ev[c], b->h_streams[c % 3]));
  }
  if (link) {
    // the pool expands chunk c while the later chunks are still on the link
    c = 0;
    for (size_t g0 = 0; g0 < N; g0 += cs, ++c) {
      const size_t g1 = g0 + cs < N ? g0 + cs : N;
      HB_CUDA(cudaEventSynchronize(ev[c]));
      if (ExpandRange(b, g0, g1, obs, row_pitch)) return 1;
    }
  }
  for (int i = 0; i < 3; ++i) HB_CUDA(cudaStreamSynchronize(b->h_streams[i]));
  return 0;
}

// Asynchronous, ranged form for callers that pipeline their own host work: games
// [first_game, first_game + num_games) only (first_game a multiple of 128), pointers are
// the bases of the whole arrays.  Returns as soon as the work is queued; the range's host
// buffers are complete after BatchHostWait(batch, first_game).  While one range is in
// flight the caller can prepare the actions of the next (the policy of range c overlaps the
// transfers of range c-1), which keeps the PCIe link busy for the whole step.
static int StepHostRange(pyhanabi_batch_t* h, int64_t first_game, int64_t num_games,
                         const int32_t* actions, int32_t* rewards, uint8_t* dones,
                         int32_t* scores, uint8_t* obs, int64_t row_pitch, uint32_t* obs_bits,
                         uint8_t* mask, int32_t* cur_player, int auto_reset) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodeHostRange: actions is null");
  if (obs && row_pitch < b->spec.obs_size) return Fail("BatchStepEncodeHostRange: row_pitch smaller than the observation");
  if (first_game < 0 || num_games <= 0 || first_game + num_games > b->N || (first_game & 127))
    return Fail("BatchStepEncodeHostRange: range must lie in the batch and start at a multiple of 128");
  if (HostSetup(b)) return 1;
  const bool link = obs && UsePackedLink(b);
  if (link && LinkSetup(b)) return 1;
  // a range queued again before anybody waited for it: its earlier rows are superseded
  for (int i = 0; i < Batch::kRangeSlots; ++i)
    if (b->range_obs[i] && b->range_first[i] == first_game) b->range_obs[i] = nullptr;
  const int slot = b->range_next++ % Batch::kRangeSlots;
  if (b->range_obs[slot]) {
    // the ring wrapped around a range nobody waited for: deliver it before its slot is reused
    HB_CUDA(cudaEventSynchronize(b->range_events[slot]));
    if (ExpandRange(b, (size_t)b->range_first[slot], (size_t)(b->range_first[slot] + b->range_count[slot]),
                    b->range_obs[slot], b->range_pitch[slot]))
      return 1;
    b->range_obs[slot] = nullptr;
  }
  // a range always runs on the same internal stream (chosen by its first game), so consecutive
  // steps of one range are ordered by that stream even if the caller skips BatchHostWait;
  // ranges must not overlap other ranges that are still in flight
  cudaStream_t s = b->h_streams[(first_game >> 7) % 3];
  if (EnqueueHostRange(b, (size_t)first_game, (size_t)(first_game + num_games), actions, rewards, dones,
                       scores, obs, row_pitch, obs_bits, mask, cur_player, auto_reset, s, link))
    return 1;
  HB_CUDA(cudaEventRecord(b->range_events[slot], s));
  b->range_first[slot] = first_game;
  b->range_count[slot] = num_games;
  b->range_obs[slot] = link ? obs : nullptr;
  b->range_pitch[slot] = row_pitch;
  return 0;
}

int BatchStepEncodeHostRange(pyhanabi_batch_t* h, int64_t first_game, int64_t num_games,
                             const int32_t* actions, int32_t* rewards, uint8_t* dones,
                             int32_t* scores, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                             int32_t* cur_player, int auto_reset) {
  return StepHostRange(h, first_game, num_games, actions, rewards, dones, scores, obs, row_pitch, nullptr,
                       mask, cur_player, auto_reset);
}

// Same with bit-packed observation rows in the host buffer (BatchStepEncodeHostPacked, ranged).
int BatchStepEncodeHostRangePacked(pyhanabi_batch_t* h, int64_t first_game, int64_t num_games,
                                   const int32_t* actions, int32_t* rewards, uint8_t* dones,
                                   int32_t* scores, uint32_t* obs_bits, uint8_t* mask,
                                   int32_t* cur_player, int auto_reset) {
  return StepHostRange(h, first_game, num_games, actions, rewards, dones, scores, nullptr, 0, obs_bits,
                       mask, cur_player, auto_reset);
}

// Blocks until the most recent BatchStepEncodeHostRange call that started at first_game has
// delivered its host buffers; first_game < 0 waits for everything queued.
int BatchHostWait(pyhanabi_batch_t* h, int64_t first_game) {
  HB_BATCH(h);
  if (!b->h_streams[0]) return 0;
  // packed link: the slot's byte rows are expanded here, on the worker pool, once its packed
  // rows have arrived
  auto deliver = [&](int slot) -> int {
    HB_CUDA(cudaEventSynchronize(b->range_events[slot]));
    if (b->range_obs[slot]) {
      uint8_t* obs = b->range_obs[slot];
      b->range_obs[slot] = nullptr;
      if (ExpandRange(b, (size_t)b->range_first[slot], (size_t)(b->range_first[slot] + b->range_count[slot]),
                      obs, b->range_pitch[slot]))
        return 1;
    }
    return 0;
  };
  if (first_game < 0) {
    // oldest first, so that the pool works on rows that have arrived while the rest still travels
    for (int k = Batch::kRangeSlots; k >= 1; --k) {
      if (b->range_next - k < 0) continue;
      const int slot = (b->range_next - k) % Batch::kRangeSlots;
      if (b->range_obs[slot] && deliver(slot)) return 1;
    }
    for (int i = 0; i < 3; ++i) HB_CUDA(cudaStreamSynchronize(b->h_streams[i]));
    return 0;
  }
  for (int k = 1; k <= Batch::kRangeSlots; ++k) {  // newest first
    if (b->range_next - k < 0) break;
    const int slot = (b->range_next - k) % Batch::kRangeSlots;
    if (b->range_first[slot] == first_game) return deliver(slot);
  }
  return 0;  // nothing queued for that range
}

// 1: byte rows of the host entry points cross the link bit-packed and are expanded on the host
// (default), 0: they cross as bytes.  Not while ranges are in flight.
int BatchSetHostLink(pyhanabi_batch_t* h, int packed) {
  HB_BATCH(h);
  if (BatchHostWait(h, -1)) return 1;
  b->host_link = packed ? 1 : 0;
  return 0;
}
int BatchGetHostLink(pyhanabi_batch_t* h) {
  HB_BATCH(h);
  return UsePackedLink(b) ? 1 : 0;
}

int BatchStepEncodeHost(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards,
                        uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                        uint8_t* mask, int32_t* cur_player, int auto_reset, void* stream) {
  return StepHost(h, actions, rewards, dones, scores, obs, row_pitch, nullptr, mask, cur_player,
                  auto_reset, stream);
}

// Same with bit-packed observation rows in the HOST buffer (8x fewer PCIe bytes).
int BatchStepEncodeHostPacked(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards,
                              uint8_t* dones, int32_t* scores, uint32_t* obs_bits, uint8_t* mask,
                              int32_t* cur_player, int auto_reset, void* stream) {
  return StepHost(h, actions, rewards, dones, scores, nullptr, 0, obs_bits, mask, cur_player,
                  auto_reset, stream);
}

// rl_env.HanabiEnv.reset with HOST result buffers (observation, legal mask and
// current player of every game after the deal).  Synchronous.
int BatchResetHost(pyhanabi_batch_t* h, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                   int32_t* cur_player, void* stream) {
  HB_BATCH(h);
  const size_t N = (size_t)b->N;
  const int D = b->spec.obs_size, A = b->spec.max_moves;
  if (obs && row_pitch < D) return Fail("BatchResetHost: row_pitch smaller than the observation");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t *d_obs = nullptr, *d_mask = nullptr;
  int32_t* d_cur = nullptr;
  hb::KArgs a = BaseArgs(b);
  HB_CUDA(Launch<hb::kModeReset>(b, a, s));
  if (obs) HB_CUDA(cudaMallocAsync(&d_obs, N * (size_t)D, s));
  if (mask) HB_CUDA(cudaMallocAsync(&d_mask, N * (size_t)A, s));
  if (cur_player) HB_CUDA(cudaMallocAsync(&d_cur, N * sizeof(int32_t), s));
  a.obs = d_obs; a.obs_pitch = D; a.mask = d_mask; a.cur_player = d_cur;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, s));
  if (obs) HB_CUDA(cudaMemcpy2DAsync(obs, (size_t)row_pitch, d_obs, (size_t)D, (size_t)D, N, cudaMemcpyDeviceToHost, s));
  if (mask) HB_CUDA(cudaMemcpyAsync(mask, d_mask, N * (size_t)A, cudaMemcpyDeviceToHost, s));
  if (cur_player) HB_CUDA(cudaMemcpyAsync(cur_player, d_cur, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (d_obs) HB_CUDA(cudaFreeAsync(d_obs, s));
  if (d_mask) HB_CUDA(cudaFreeAsync(d_mask, s));
  if (d_cur) HB_CUDA(cudaFreeAsync(d_cur, s));
  HB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Bit-packed observations: row i = BatchPackedObservationWords() 32-bit words
// (a multiple of 4, i.e. 16-byte rows), bit j of the row = observation bit j, padding
// bits zero.  8x less HBM / PCIe traffic than one byte per bit, for consumers that
// unpack on the fly; np.unpackbits(rows.view(uint8), bitorder="little")[:, :obs_size]
// gives back the byte form.
int BatchPackedObservationWords(pyhanabi_batch_t* h) {
  return (h && h->batch) ? AsBatch(h->batch)->packed_words : -1;
}

int BatchStepEncodePacked(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards,
                          uint8_t* dones, int32_t* scores, uint32_t* obs_bits, uint8_t* mask,
                          int32_t* cur_player, int auto_reset, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodePacked: actions is null");
  if (obs_bits && (reinterpret_cast<uintptr_t>(obs_bits) & 15)) return Fail("BatchStepEncodePacked: obs_bits must be 16-byte aligned");
  hb::KArgs a = BaseArgs(b);
  a.actions = actions; a.rewards = rewards; a.dones = dones; a.scores = scores;
  a.obs_bits = obs_bits; a.mask = mask; a.cur_player = cur_player; a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

namespace {
int SetNetOutput(const Batch* b, hb::KArgs* a, void* net_input, int dtype, int64_t row_pitch, const char* who) {
  if (!net_input) return Fail(std::string(who) + ": net_input is null");
  if (dtype < 0 || dtype > 2) return Fail(std::string(who) + ": dtype must be 0 (float32), 1 (bfloat16) or 2 (float16)");
  if (row_pitch < b->spec.obs_size || (row_pitch & 7)) return Fail(std::string(who) + ": row_pitch must be >= the observation size and a multiple of 8");
  if (reinterpret_cast<uintptr_t>(net_input) & 15) return Fail(std::string(who) + ": net_input must be 16-byte aligned");
  a->net_in = net_input; a->net_dtype = dtype; a->net_pitch = row_pitch;
  a->net_chunks = (int)(row_pitch >> 3);
  a->net_magic = (uint32_t)((((uint64_t)1 << 32) + a->net_chunks - 1) / a->net_chunks);
  return 0;
}
}  // namespace

// The fused step writing the observation as the NETWORK'S INPUT: rows of 0/1 in float32 /
// bfloat16 / float16 (dtype 0 / 1 / 2), row_pitch elements per game (>= observation size,
// multiple of 8; the columns beyond the observation are zero, so K can be padded to the
// GEMM's alignment) -- BatchStepEncodePacked + BatchExpandPacked in one kernel.  obs_bits
// (nullable) receives the bit-packed rows as well (what the recorder keeps).
int BatchStepEncodeNet(pyhanabi_batch_t* h, const int32_t* actions, int32_t* rewards, uint8_t* dones,
                       int32_t* scores, void* net_input, int dtype, int64_t row_pitch,
                       uint32_t* obs_bits, uint8_t* mask, int32_t* cur_player, int auto_reset,
                       void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchStepEncodeNet: actions is null");
  if (obs_bits && (reinterpret_cast<uintptr_t>(obs_bits) & 15)) return Fail("BatchStepEncodeNet: obs_bits must be 16-byte aligned");
  hb::KArgs a = BaseArgs(b);
  if (SetNetOutput(b, &a, net_input, dtype, row_pitch, "BatchStepEncodeNet")) return 1;
  a.actions = actions; a.rewards = rewards; a.dones = dones; a.scores = scores;
  a.obs_bits = obs_bits; a.mask = mask; a.cur_player = cur_player; a.auto_reset = auto_reset;
  HB_CUDA(Launch<hb::kModeStep>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchObserveNet(pyhanabi_batch_t* h, int observer, void* net_input, int dtype, int64_t row_pitch,
                    uint32_t* obs_bits, uint8_t* mask, int32_t* cur_player, void* stream) {
  HB_BATCH(h);
  if (observer >= b->spec.P) return Fail("BatchObserveNet: observer out of range");
  if (obs_bits && (reinterpret_cast<uintptr_t>(obs_bits) & 15)) return Fail("BatchObserveNet: obs_bits must be 16-byte aligned");
  hb::KArgs a = BaseArgs(b);
  if (SetNetOutput(b, &a, net_input, dtype, row_pitch, "BatchObserveNet")) return 1;
  a.obs_bits = obs_bits; a.observer = observer; a.mask = mask; a.cur_player = cur_player;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

int BatchObservePacked(pyhanabi_batch_t* h, int observer, uint32_t* obs_bits, uint8_t* mask,
                       int32_t* cur_player, void* stream) {
  HB_BATCH(h);
  if (observer >= b->spec.P) return Fail("BatchObservePacked: observer out of range");
  if (obs_bits && (reinterpret_cast<uintptr_t>(obs_bits) & 15)) return Fail("BatchObservePacked: obs_bits must be 16-byte aligned");
  hb::KArgs a = BaseArgs(b);
  a.obs_bits = obs_bits; a.observer = observer; a.mask = mask; a.cur_player = cur_player;
  HB_CUDA(Launch<hb::kModeObserve>(b, a, (cudaStream_t)stream));
  return 0;
}

// Checkpoint blob: a 128-byte header (magic, layout version, the complete rule set, the
// number of games and state columns, whether the move-history ring is on), then the packed
// state columns, every game's mt19937 block, the history ring (if on)
// and the 32 statistics counters.
namespace {
constexpr int64_t kBlobMagic = 0x4842323030ll, kBlobVersion = 2;
constexpr int kBlobHeader = 16;  // int64 words
void FillBlobHeader(const Batch* b, int64_t* hd) {
  const hb::Spec& s = b->spec;
  const int64_t v[kBlobHeader] = {kBlobMagic, kBlobVersion, b->N, b->ncols, s.P, s.C, s.R, s.H, s.IT, s.LT,
                                  s.obs_type, s.random_start, s.pend_cap, b->hist ? 1 : 0, s.obs_size, s.max_moves};
  std::memcpy(hd, v, sizeof(v));
}
}  // namespace

int64_t BatchStateBytes(pyhanabi_batch_t* h) {
  if (!h || !h->batch) return -1;
  Batch* b = AsBatch(h->batch);
  return (int64_t)sizeof(int64_t) * kBlobHeader + (int64_t)sizeof(uint4) * b->ncols * b->N +
         (int64_t)sizeof(uint32_t) * hb::kMtN * b->N +
         (b->hist ? (int64_t)(sizeof(uint4) + sizeof(uint32_t) * hb::kDiscardLogWords) * b->N : 0) +
         (int64_t)sizeof(int64_t) * hb::kNumStats;
}

int BatchGetState(pyhanabi_batch_t* h, void* host_blob, int64_t nbytes) {
  HB_BATCH(h);
  if (nbytes < BatchStateBytes(h)) return Fail("BatchGetState: blob too small, see BatchStateBytes");
  HB_CUDA(cudaDeviceSynchronize());
  int64_t header[kBlobHeader];
  FillBlobHeader(b, header);
  std::memcpy(host_blob, header, sizeof(header));
  char* p = static_cast<char*>(host_blob) + sizeof(header);
  const size_t N = (size_t)b->N;
  const size_t sb = sizeof(uint4) * (size_t)b->ncols * N, qb = 0,
               mb = sizeof(uint32_t) * (size_t)hb::kMtN * N, hbytes = b->hist ? sizeof(uint4) * N : 0;
  HB_CUDA(cudaMemcpy(p, b->state, sb, cudaMemcpyDeviceToHost));
  HB_CUDA(cudaMemcpy(p + sb + qb, b->mt, mb, cudaMemcpyDeviceToHost));
  const size_t lbytes = b->hist ? sizeof(uint32_t) * hb::kDiscardLogWords * N : 0;
  if (hbytes) HB_CUDA(cudaMemcpy(p + sb + qb + mb, b->hist, hbytes, cudaMemcpyDeviceToHost));
  if (lbytes) HB_CUDA(cudaMemcpy(p + sb + qb + mb + hbytes, b->disc_log, lbytes, cudaMemcpyDeviceToHost));
  HB_CUDA(cudaMemcpy(p + sb + qb + mb + hbytes + lbytes, b->stats, sizeof(int64_t) * hb::kNumStats, cudaMemcpyDeviceToHost));
  return 0;
}

int BatchSetState(pyhanabi_batch_t* h, const void* host_blob, int64_t nbytes) {
  HB_BATCH(h);
  if (nbytes < BatchStateBytes(h)) return Fail("BatchSetState: blob too small, see BatchStateBytes");
  int64_t header[kBlobHeader], mine[kBlobHeader];
  std::memcpy(header, host_blob, sizeof(header));
  FillBlobHeader(b, mine);
  if (std::memcmp(header, mine, sizeof(header)) != 0)
    return Fail("BatchSetState: blob does not match this batch (layout version, games, rule set, "
                "observation type or history setting differ)");
  HB_CUDA(cudaDeviceSynchronize());
  const char* p = static_cast<const char*>(host_blob) + sizeof(header);
  const size_t N = (size_t)b->N;
  const size_t sb = sizeof(uint4) * (size_t)b->ncols * N, qb = 0,
               mb = sizeof(uint32_t) * (size_t)hb::kMtN * N, hbytes = b->hist ? sizeof(uint4) * N : 0;
  HB_CUDA(cudaMemcpy(b->state, p, sb, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(b->mt, p + sb + qb, mb, cudaMemcpyHostToDevice));
  const size_t lbytes = b->hist ? sizeof(uint32_t) * hb::kDiscardLogWords * N : 0;
  if (hbytes) HB_CUDA(cudaMemcpy(b->hist, p + sb + qb + mb, hbytes, cudaMemcpyHostToDevice));
  if (lbytes) HB_CUDA(cudaMemcpy(b->disc_log, p + sb + qb + mb + hbytes, lbytes, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(b->stats, p + sb + qb + mb + hbytes + lbytes, sizeof(int64_t) * hb::kNumStats, cudaMemcpyHostToDevice));
  return 0;
}

int BatchUnpackState(pyhanabi_batch_t* h, int32_t* dump, void* stream) {
  HB_BATCH(h);
  if (!dump) return Fail("BatchUnpackState: dump is null");
  hb::KArgs a = BaseArgs(b);
  const unsigned grid = (unsigned)((b->N + 127) / 128);
  if (b->spec.H > hb::kMaxHandNarrow) hb::k_unpack<hb::WView><<<grid, 128, 0, (cudaStream_t)stream>>>(a, dump);
  else hb::k_unpack<hb::DView><<<grid, 128, 0, (cudaStream_t)stream>>>(a, dump);
  HB_CUDA(cudaGetLastError());
  return 0;
}

// Illegal actions do not abort: the game is left unchanged and flagged.  Writes the
// flags (device uint8[num_games], nullable), adds their number to *count (device
// uint64, nullable; the caller zeroes it) and optionally clears them.
int BatchErrorFlags(pyhanabi_batch_t* h, uint8_t* flags, uint64_t* count, int clear, void* stream) {
  HB_BATCH(h);
  const unsigned grid = (unsigned)((b->N + 127) / 128);
  const bool wide = b->spec.H > hb::kMaxHandNarrow;  // wide layout: column 2 P + 1, x, bit 31
  const bool compact = !wide && hb::state_columns(b->spec.P) == b->spec.P + 1;
  hb::k_errors<<<grid, 128, 0, (cudaStream_t)stream>>>(b->state, b->N,
                                                       wide ? 2 * b->spec.P + 1 : compact ? b->spec.P : b->spec.P + 1,
                                                       compact ? 1 : 0, compact ? 25 : 31, flags, clear,
                                                       reinterpret_cast<unsigned long long*>(count));
  HB_CUDA(cudaGetLastError());
  return 0;
}

// One game of the batch as a host-side record (hb::kExportLen int32 values: the
// canonical dump, turns_to_play, the last non-deal move and the moves before it when
// the history ring is on).  StateFromBatchGame (hb_single.cc) turns it into a
// single-game HanabiState for the reference's per-object API.  Synchronous.
int BatchExportGame(pyhanabi_batch_t* h, int64_t game, int32_t* host_record, void* stream) {
  HB_BATCH(h);
  if (!host_record) return Fail("BatchExportGame: host_record is null");
  if (game < 0 || game >= b->N) return Fail("BatchExportGame: game index out of range");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* d = nullptr;
  HB_CUDA(cudaMallocAsync(&d, sizeof(int32_t) * hb::kExportLen, s));
  hb::KArgs a = BaseArgs(b);
  if (b->spec.H > hb::kMaxHandNarrow) hb::k_export<hb::WView><<<1, 32, 0, s>>>(a, game, d);
  else hb::k_export<hb::DView><<<1, 32, 0, s>>>(a, game, d);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(host_record, d, sizeof(int32_t) * hb::kExportLen, cudaMemcpyDeviceToHost, s);
  cudaFreeAsync(d, s);
  if (e != cudaSuccess) return CudaFail("BatchExportGame", e);
  HB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int BatchPackState(pyhanabi_batch_t* h, const int32_t* dump, const int32_t* extra, void* stream) {
  HB_BATCH(h);
  if (!dump) return Fail("BatchPackState: dump is null");
  hb::KArgs a = BaseArgs(b);
  const unsigned grid = (unsigned)((b->N + 127) / 128);
  if (b->spec.H > hb::kMaxHandNarrow) hb::k_pack<hb::WView><<<grid, 128, 0, (cudaStream_t)stream>>>(a, dump, extra);
  else hb::k_pack<hb::DView><<<grid, 128, 0, (cudaStream_t)stream>>>(a, dump, extra);
  HB_CUDA(cudaGetLastError());
  return 0;
}

// Keeps the last non-deal moves of every game (16 bytes per game, updated by every
// step) so that the rule-based policies can see what happened since their previous
// turn (HanabiObservation::LastMoves, hanabi_lib/hanabi_observation.cc:77-92).
// Off by default: the hot path does not pay for it unless asked.
int BatchEnableHistory(pyhanabi_batch_t* h) {
  HB_BATCH(h);
  if (b->hist) return 0;
  HB_CUDA(cudaMalloc(&b->hist, sizeof(uint4) * (size_t)b->N));
  HB_CUDA(cudaMemset(b->hist, 0, sizeof(uint4) * (size_t)b->N));
  HB_CUDA(cudaMalloc(&b->disc_log, sizeof(uint32_t) * hb::kDiscardLogWords * (size_t)b->N));
  HB_CUDA(cudaMemset(b->disc_log, 0, sizeof(uint32_t) * hb::kDiscardLogWords * (size_t)b->N));
  return 0;
}

// The reference's hand-written policies for a whole batch (hb_agents.cuh): writes
// actions[i] for every game whose player to act sits in seat_mask.  agent_kind 0 =
// RuleBasedAgent (needs BatchEnableHistory before the games are reset, and
// agent_state: device uint32[num_games][num_slots], zero-initialised by the caller,
// never reset between episodes -- like the Python object), 1 = SimpleAgent
// (stateless; agent_state may be NULL).  seat_slot: HOST int[players] mapping a seat
// to its state slot (NULL = slot 0 for every seat, i.e. one shared agent object).
int BatchAgentAct(pyhanabi_batch_t* h, int agent_kind, uint32_t seat_mask, const int32_t* seat_slot,
                  uint32_t* agent_state, int num_slots, int32_t* actions, void* stream) {
  HB_BATCH(h);
  if (!actions) return Fail("BatchAgentAct: actions is null");
  if (agent_kind != 0 && agent_kind != 1) return Fail("BatchAgentAct: agent_kind must be 0 (RuleBasedAgent) or 1 (SimpleAgent)");
  if (b->spec.H > hb::kMaxHandNarrow)
    return Fail("BatchAgentAct: the batched agents read hands of up to five cards (rule_based_agent.py:143 "
                "hard-codes five slots)");
  hb::AgentArgs a;
  std::memset(&a, 0, sizeof(a));
  if (agent_kind == 0) {
    if (!b->hist) return Fail("BatchAgentAct: RuleBasedAgent needs BatchEnableHistory");
    if (!agent_state || num_slots < 1) return Fail("BatchAgentAct: RuleBasedAgent needs agent_state");
    if (b->spec.H > (b->spec.P < 4 ? 5 : 4))
      return Fail("BatchAgentAct: RuleBasedAgent tracks 5 (4 from four players) cards per hand");
  }
  for (int p = 0; p < hb::kMaxPlayers; ++p) {
    a.seat_slot[p] = (seat_slot && p < b->spec.P) ? seat_slot[p] : 0;
    if (agent_kind == 0 && (a.seat_slot[p] < 0 || a.seat_slot[p] >= num_slots))
      return Fail("BatchAgentAct: seat_slot out of range");
  }
  a.spec = b->spec;
  a.state = b->state;
  a.hist = b->hist;
  a.N = b->N;
  a.agent_state = agent_state;
  a.num_slots = num_slots;
  a.seat_mask = seat_mask;
  a.kind = agent_kind;
  a.max_information_tokens = b->spec.IT;
  a.actions = actions;
  hb::k_agent_act<<<(unsigned)((b->N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
  HB_CUDA(cudaGetLastError());
  return 0;
}

// Episode statistics accumulate into a caller-owned device buffer of 32 int64
// (e.g. a torch tensor that torch.distributed all-reduces over NCCL); NULL
// switches back to the library-owned buffer.
int BatchSetStatsBuffer(pyhanabi_batch_t* h, int64_t* dev_stats) {
  HB_BATCH(h);
  b->stats = dev_stats ? reinterpret_cast<unsigned long long*>(dev_stats) : b->own_stats;
  return 0;
}

int BatchReadStats(pyhanabi_batch_t* h, int64_t* host_stats, int clear, void* stream) {
  HB_BATCH(h);
  cudaStream_t s = (cudaStream_t)stream;
  HB_CUDA(cudaMemcpyAsync(host_stats, b->stats, sizeof(int64_t) * hb::kNumStats, cudaMemcpyDeviceToHost, s));
  if (clear) HB_CUDA(cudaMemsetAsync(b->stats, 0, sizeof(int64_t) * hb::kNumStats, s));
  HB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Keeps the packed game state (16 bytes per column and game) resident in the L2 cache
// between steps: every launch on this batch carries an access-policy window over
// the state array (persisting on hit, streaming on miss), and the device's
// persisting-L2 carve-out is raised to `max_bytes` (0 = the device maximum,
// negative = switch the window off).  The observation / mask stores of the step
// stream past it.  Returns the number of state bytes the window covers through
// *covered (may be NULL).  Worth it when the state fits: 64 MB at 2^20 2p games.
int BatchSetL2Residency(pyhanabi_batch_t* h, int64_t max_bytes, int64_t* covered) {
  HB_BATCH(h);
  if (covered) *covered = 0;
  if (max_bytes < 0) {
    b->l2_window_bytes = 0;
    b->l2_hit_ratio = 0.f;
    return 0;
  }
  int max_persist = 0, max_window = 0;
  HB_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, b->device));
  HB_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, b->device));
  if (max_persist <= 0 || max_window <= 0) return Fail("BatchSetL2Residency: the device has no persisting L2");
  size_t carve = (size_t)max_persist;
  if (max_bytes > 0 && (size_t)max_bytes < carve) carve = (size_t)max_bytes;
  HB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
  const size_t state_bytes = sizeof(uint4) * (size_t)b->ncols * (size_t)b->N;
  size_t window = state_bytes < (size_t)max_window ? state_bytes : (size_t)max_window;
  b->l2_window_bytes = window;
  b->l2_hit_ratio = window <= carve ? 1.0f : (float)((double)carve / (double)window);
  if (covered) *covered = (int64_t)((double)window * b->l2_hit_ratio);
  return 0;
}

// 0 = integer fast path with exact fp64 fallback (default), 1 = always the exact
// fp64 restatement, 2 = fast path only (tests).  `generic` != 0 forces the
// runtime-parameter kernels even for the pre-instantiated rule sets (tests).
int BatchSetDebug(pyhanabi_batch_t* h, int sampler_mode, int generic) {
  HB_BATCH(h);
  b->sampler_mode = sampler_mode;
  b->force_generic = generic;
  return 0;
}

int BatchDebugDeal(pyhanabi_batch_t* h, const uint64_t* decks, const int32_t* sizes,
                   const uint32_t* lo, const uint32_t* hi, int n, int mode, int32_t* out_type,
                   int32_t* out_ambiguous, void* stream) {
  HB_BATCH(h);
  hb::k_deal_test<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(b->spec, decks, sizes, lo, hi, n,
                                                                    mode, out_type, out_ambiguous);
  HB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
