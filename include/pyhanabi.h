// include/pyhanabi.h -- C-ABI of the B200-native batched Hanabi engine.
//
// Drop-in replacement for the reference's
// hanabi_learning_environment/pyhanabi.h (the file its cffi binding text-parses,
// hanabi_learning_environment/pyhanabi.py:42-74): the same single
// `extern "C" { ... } /* extern "C" */` block of plain C declarations, loaded in
// cffi ABI mode, extended with the batched entry points ("Batch*") that replace
// the per-game hot path
//   StateApplyMove / StateDealRandomCard / StateLegalMoves / StateScore /
//   StateEndOfGameStatus / NewObservation / EncodeObservation
//   (reference pyhanabi.h:115,117,127,130,125,157,189; driven one game at a time
//    from rl_env.HanabiEnv.step, rl_env.py:343-366)
// for thousands of games per call.
//
// Conventions of the batched block
//   * int status: 0 = ok, non-zero = failure, text from BatchLastError()
//     (the reference aborts on bad input, hanabi_lib/util.h:53-61; a batch must
//     not take the process down for one bad game).
//   * Every pointer argument is DEVICE memory owned by the caller (e.g. a torch
//     tensor's data_ptr()) unless the comment says "host".  `stream` is a
//     cudaStream_t (torch.cuda.current_stream().cuda_stream) or NULL; calls are
//     asynchronous on it.
//   * Game i uses its own std::mt19937 stream seeded with seeds[i], exactly like
//     one reference HanabiGame per environment (hanabi_lib/hanabi_game.cc:48-51,
//     hanabi_lib/hanabi_game.h:114); episodes of a slot continue that stream.
//   * Move uids are the reference's (hanabi_lib/hanabi_game.cc:79-95,154-183).
//   * An illegal action leaves that game unchanged (reward 0, done 0) and sets
//     its error flag instead of aborting (hanabi_lib/hanabi_state.cc:222).
//   * There is no CPU fallback: without a CUDA device NewBatch fails.

#ifndef __PYHANABI_H__
#define __PYHANABI_H__

/* C++ header, like the reference's: everything between the two marker lines
 * below must stay plain C declarations (no preprocessor lines) because the cffi
 * binding feeds that text to ffi.cdef(). */
#include <stdbool.h>
#include <stdint.h>

extern "C" {

/* ===== per-object API kept from the reference ============================
 * Identical names, argument lists and ownership rules as the reference header
 * (hanabi_learning_environment/pyhanabi.h:26-190): every New* / Get*Move /
 * *GetMoveHistory / ObsGetLastMove / ObsGetLegalMove call allocates an object the
 * caller releases with the matching Delete*; strings are released with
 * DeleteString; a pyhanabi_card_knowledge_t borrows from its observation.
 * Input errors abort the process (hanabi_lib/util.h:53-61).  Host-side,
 * implemented in hanabi_b200/csrc/hb_single.cc. */

/* handles: one pointer each (reference pyhanabi.h:26-64) */
typedef struct PyHanabiCard { int color; int rank; } pyhanabi_card_t;
typedef struct PyHanabiCardKnowledge { const void* knowledge; } pyhanabi_card_knowledge_t;
typedef struct PyHanabiMove { void* move; } pyhanabi_move_t;
typedef struct PyHanabiHistoryItem { void* item; } pyhanabi_history_item_t;
typedef struct PyHanabiState { void* state; } pyhanabi_state_t;
typedef struct PyHanabiGame { void* game; } pyhanabi_game_t;
typedef struct PyHanabiObservation { void* observation; } pyhanabi_observation_t;
typedef struct PyHanabiObservationEncoder { void* encoder; } pyhanabi_observation_encoder_t;

/* strings and cards (pyhanabi.h:67,70) */
void DeleteString(char* str);
int CardValid(pyhanabi_card_t* card);

/* hint knowledge about one card (pyhanabi.h:73-79) */
char* CardKnowledgeToString(pyhanabi_card_knowledge_t* knowledge);
int ColorWasHinted(pyhanabi_card_knowledge_t* knowledge);
int KnownColor(pyhanabi_card_knowledge_t* knowledge);
int ColorIsPlausible(pyhanabi_card_knowledge_t* knowledge, int color);
int RankWasHinted(pyhanabi_card_knowledge_t* knowledge);
int KnownRank(pyhanabi_card_knowledge_t* knowledge);
int RankIsPlausible(pyhanabi_card_knowledge_t* knowledge, int rank);

/* moves and move lists (pyhanabi.h:82-95) */
void DeleteMoveList(void* movelist);
int NumMoves(void* movelist);
void GetMove(void* movelist, int index, pyhanabi_move_t* move);
void DeleteMove(pyhanabi_move_t* move);
char* MoveToString(pyhanabi_move_t* move);
int MoveType(pyhanabi_move_t* move);
int CardIndex(pyhanabi_move_t* move);
int TargetOffset(pyhanabi_move_t* move);
int MoveColor(pyhanabi_move_t* move);
int MoveRank(pyhanabi_move_t* move);
bool GetDiscardMove(int card_index, pyhanabi_move_t* move);
bool GetPlayMove(int card_index, pyhanabi_move_t* move);
bool GetRevealColorMove(int target_offset, int color, pyhanabi_move_t* move);
bool GetRevealRankMove(int target_offset, int rank, pyhanabi_move_t* move);

/* history items: a move plus its side effects (pyhanabi.h:98-108) */
void DeleteHistoryItem(pyhanabi_history_item_t* item);
char* HistoryItemToString(pyhanabi_history_item_t* item);
void HistoryItemMove(pyhanabi_history_item_t* item, pyhanabi_move_t* move);
int HistoryItemPlayer(pyhanabi_history_item_t* item);
int HistoryItemScored(pyhanabi_history_item_t* item);
int HistoryItemInformationToken(pyhanabi_history_item_t* item);
int HistoryItemColor(pyhanabi_history_item_t* item);
int HistoryItemRank(pyhanabi_history_item_t* item);
int HistoryItemRevealBitmask(pyhanabi_history_item_t* item);
int HistoryItemNewlyRevealedBitmask(pyhanabi_history_item_t* item);
int HistoryItemDealToPlayer(pyhanabi_history_item_t* item);

/* one game state (pyhanabi.h:111-139) */
void NewState(pyhanabi_game_t* game, pyhanabi_state_t* state);
void CopyState(const pyhanabi_state_t* src, pyhanabi_state_t* dest);
void DeleteState(pyhanabi_state_t* state);
const void* StateParentGame(pyhanabi_state_t* state);
void StateApplyMove(pyhanabi_state_t* state, pyhanabi_move_t* move);
int StateCurPlayer(pyhanabi_state_t* state);
void StateDealRandomCard(pyhanabi_state_t* state);
int StateDeckSize(pyhanabi_state_t* state);
int StateFireworks(pyhanabi_state_t* state, int color);
int StateDiscardPileSize(pyhanabi_state_t* state);
void StateGetDiscard(pyhanabi_state_t* state, int index, pyhanabi_card_t* card);
int StateGetHandSize(pyhanabi_state_t* state, int pid);
void StateGetHandCard(pyhanabi_state_t* state, int pid, int index, pyhanabi_card_t* card);
int StateEndOfGameStatus(pyhanabi_state_t* state);
int StateInformationTokens(pyhanabi_state_t* state);
void* StateLegalMoves(pyhanabi_state_t* state);
int StateLifeTokens(pyhanabi_state_t* state);
int StateNumPlayers(pyhanabi_state_t* state);
int StateScore(pyhanabi_state_t* state);
char* StateToString(pyhanabi_state_t* state);
bool MoveIsLegal(const pyhanabi_state_t* state, const pyhanabi_move_t* move);
bool CardPlayableOnFireworks(const pyhanabi_state_t* state, int color, int rank);
int StateLenMoveHistory(pyhanabi_state_t* state);
void StateGetMoveHistory(pyhanabi_state_t* state, int index, pyhanabi_history_item_t* item);

/* rule set + the game's own RNG stream (pyhanabi.h:142-154) */
void DeleteGame(pyhanabi_game_t* game);
void NewDefaultGame(pyhanabi_game_t* game);
void NewGame(pyhanabi_game_t* game, int list_length, const char** param_list);
char* GameParamString(pyhanabi_game_t* game);
int NumPlayers(pyhanabi_game_t* game);
int NumColors(pyhanabi_game_t* game);
int NumRanks(pyhanabi_game_t* game);
int HandSize(pyhanabi_game_t* game);
int MaxInformationTokens(pyhanabi_game_t* game);
int MaxLifeTokens(pyhanabi_game_t* game);
int ObservationType(pyhanabi_game_t* game);
int NumCards(pyhanabi_game_t* game, int color, int rank);
int GetMoveUid(pyhanabi_game_t* game, pyhanabi_move_t* move);
void GetMoveByUid(pyhanabi_game_t* game, int move_uid, pyhanabi_move_t* move);
int MaxMoves(pyhanabi_game_t* game);

/* one player's view of a state (pyhanabi.h:157-182) */
void NewObservation(pyhanabi_state_t* state, int player, pyhanabi_observation_t* observation);
void DeleteObservation(pyhanabi_observation_t* observation);
char* ObsToString(pyhanabi_observation_t* observation);
int ObsCurPlayerOffset(pyhanabi_observation_t* observation);
int ObsNumPlayers(pyhanabi_observation_t* observation);
int ObsGetHandSize(pyhanabi_observation_t* observation, int pid);
void ObsGetHandCard(pyhanabi_observation_t* observation, int pid, int index,
                    pyhanabi_card_t* card);
void ObsGetHandCardKnowledge(pyhanabi_observation_t* observation, int pid, int index,
                             pyhanabi_card_knowledge_t* knowledge);
int ObsDiscardPileSize(pyhanabi_observation_t* observation);
void ObsGetDiscard(pyhanabi_observation_t* observation, int index, pyhanabi_card_t* card);
int ObsFireworks(pyhanabi_observation_t* observation, int color);
int ObsDeckSize(pyhanabi_observation_t* observation);
int ObsNumLastMoves(pyhanabi_observation_t* observation);
void ObsGetLastMove(pyhanabi_observation_t* observation, int index,
                    pyhanabi_history_item_t* item);
int ObsInformationTokens(pyhanabi_observation_t* observation);
int ObsLifeTokens(pyhanabi_observation_t* observation);
int ObsNumLegalMoves(pyhanabi_observation_t* observation);
void ObsGetLegalMove(pyhanabi_observation_t* observation, int index, pyhanabi_move_t* move);
bool ObsCardPlayableOnFireworks(const pyhanabi_observation_t* observation, int color, int rank);

/* canonical observation encoder; EncodeObservation returns "0,1,0,..." as the
 * reference does (pyhanabi.h:185-190, pyhanabi.cc:839-858) */
void NewObservationEncoder(pyhanabi_observation_encoder_t* encoder, pyhanabi_game_t* game,
                           int type);
void DeleteObservationEncoder(pyhanabi_observation_encoder_t* encoder);
char* ObservationShape(pyhanabi_observation_encoder_t* encoder);
char* EncodeObservation(pyhanabi_observation_encoder_t* encoder,
                        pyhanabi_observation_t* observation);
/* new: the same bits as bytes into a caller buffer, no ASCII round trip */
int EncodeObservationBytes(pyhanabi_observation_encoder_t* encoder,
                           pyhanabi_observation_t* observation, uint8_t* out, int capacity);

/* ===== batched engine (new) ============================================== */

typedef struct PyHanabiBatch {
  /* Points to the device-resident batch of games. */
  void* batch;
} pyhanabi_batch_t;

/* Same key/value parameter list as NewGame (reference pyhanabi.h:142,
 * pyhanabi.cc:514-526; keys hanabi_lib/hanabi_game.cc:29-47) except "seed":
 * seeds is a HOST array of num_games int32. */
int NewBatch(pyhanabi_batch_t* batch, int list_length, const char** param_list,
             int num_games, int device, const int32_t* seeds);
void DeleteBatch(pyhanabi_batch_t* batch);
const char* BatchLastError(void);

int BatchObservationSize(pyhanabi_batch_t* batch); /* ObservationShape, pyhanabi.h:187 */
int BatchMaxMoves(pyhanabi_batch_t* batch);        /* MaxMoves, pyhanabi.h:154 */
int BatchNumGames(pyhanabi_batch_t* batch);
int BatchNumPlayers(pyhanabi_batch_t* batch);      /* NumPlayers, pyhanabi.h:144 */
int BatchDevice(pyhanabi_batch_t* batch);

/* rl_env.HanabiEnv.reset (rl_env.py:210-217) for the games with which[i] != 0
 * (NULL = all): new initial state, cards dealt until a player is to act. */
int BatchReset(pyhanabi_batch_t* batch, const uint8_t* which, void* stream);

/* rl_env.HanabiEnv.step (rl_env.py:343-366) for every game: apply move uid
 * actions[i], deal, reward = score delta, done = IsTerminal; with auto_reset a
 * finished game is re-dealt in the same call.  actions[i] == -1 means "no move
 * for game i in this call" (game untouched, reward 0).  A game that is finished
 * and not reset ignores its action (reward 0, done 1).  rewards/dones/scores may
 * be NULL. */
int BatchStep(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
              uint8_t* dones, int32_t* scores, int auto_reset, void* stream);

/* CanonicalObservationEncoder::Encode (hanabi_lib/canonical_encoders.cc:433-451)
 * of every game for `observer` (-1 = the player to act), one byte per bit into
 * obs[i * row_pitch + j].  row_pitch == BatchObservationSize gives the dense
 * learner input and the fastest (128-bit store) path. */
int BatchEncode(pyhanabi_batch_t* batch, int observer, uint8_t* obs,
                int64_t row_pitch, void* stream);

/* HanabiState::LegalMoves of the player to act (hanabi_lib/hanabi_state.cc:288-304)
 * as mask[i * BatchMaxMoves + uid] = 1 if legal else 0. */
int BatchLegalMask(pyhanabi_batch_t* batch, uint8_t* mask, void* stream);
int BatchCurPlayer(pyhanabi_batch_t* batch, int32_t* cur_player, void* stream);
/* BatchEncode + BatchLegalMask + BatchCurPlayer in one launch; any output may be NULL. */
int BatchObserve(pyhanabi_batch_t* batch, int observer, uint8_t* obs,
                 int64_t row_pitch, uint8_t* mask, int32_t* cur_player, void* stream);

/* The fused hot path: BatchStep followed by the next player's observation,
 * legal mask and index, one kernel.  Outputs may be NULL. */
int BatchStepEncode(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
                    uint8_t* dones, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                    int32_t* cur_player, int auto_reset, void* stream);
int BatchStepEncodeScores(pyhanabi_batch_t* batch, const int32_t* actions,
                          int32_t* rewards, uint8_t* dones, int32_t* scores, uint8_t* obs,
                          int64_t row_pitch, uint8_t* mask, int32_t* cur_player,
                          int auto_reset, void* stream);

/* BatchStepEncodeScores with a fused uniform-random policy (the reference's
   agents/random_agent.py for every game at once): after the move, random_actions[i]
   = a uniformly random legal move of game i's new state, -1 if none; the same draw
   as BatchSelectAction(NULL, 0, NULL, mask, ..., epsilon >= 1, seed, step).
   random_actions may alias actions (in-place rollout loop). */
int BatchStepEncodeRandom(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
                          uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                          uint8_t* mask, int32_t* cur_player, int auto_reset,
                          int32_t* random_actions, uint64_t seed, uint64_t step, void* stream);
/* A batch that is one shard of a larger set of games (one process per GPU, several
   shards per GPU) sets the global index of its game 0: the fused random policy is keyed
   (seed, step, first_game_index + i), so a sharded rollout reproduces the unsharded one
   game for game (one RNG per game, hanabi_lib/hanabi_game.h:114; games never interact). */
int BatchSetGameIndexBase(pyhanabi_batch_t* batch, int64_t first_game_index);
/* A uniformly random legal move of every game's CURRENT state (no move applied), the
   draw BatchStepEncodeRandom would leave behind for (seed, step): the first actions of
   a fused random rollout. */
int BatchRandomLegalActions(pyhanabi_batch_t* batch, uint64_t seed, uint64_t step,
                            int32_t* actions, void* stream);

/* Bit-packed observations (optional, reported separately from the byte form): row i
 * = BatchPackedObservationWords() uint32 words (16-byte multiple), bit j of the row
 * = observation bit j, padding zero.  obs_bits must be 16-byte aligned. */
int BatchPackedObservationWords(pyhanabi_batch_t* batch);
int BatchStepEncodePacked(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
                          uint8_t* dones, int32_t* scores, uint32_t* obs_bits, uint8_t* mask,
                          int32_t* cur_player, int auto_reset, void* stream);
int BatchObservePacked(pyhanabi_batch_t* batch, int observer, uint32_t* obs_bits,
                       uint8_t* mask, int32_t* cur_player, void* stream);

/* The observation as the NETWORK'S INPUT, written by the step kernel itself: rows of
 * 0/1 in float32 / bfloat16 / float16 (dtype 0 / 1 / 2), row_pitch elements per game
 * (>= observation size, multiple of 8, columns beyond the observation zero -- K padded
 * to the GEMM's alignment); obs_bits (nullable) also receives the bit-packed rows.
 * net_input must be 16-byte aligned. */
int BatchStepEncodeNet(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
                       uint8_t* dones, int32_t* scores, void* net_input, int dtype,
                       int64_t row_pitch, uint32_t* obs_bits, uint8_t* mask,
                       int32_t* cur_player, int auto_reset, void* stream);
int BatchObserveNet(pyhanabi_batch_t* batch, int observer, void* net_input, int dtype,
                    int64_t row_pitch, uint32_t* obs_bits, uint8_t* mask, int32_t* cur_player,
                    void* stream);

/* The same two verbs with HOST buffers (pinned memory recommended), for callers
 * that keep their data on the CPU like the reference's trainers do: uploads the
 * actions, runs the fused kernel and downloads the results, chunked over
 * internal streams so copies and compute overlap.  Synchronous.  Any output
 * may be NULL. */
int BatchStepEncodeHost(pyhanabi_batch_t* batch, const int32_t* actions, int32_t* rewards,
                        uint8_t* dones, int32_t* scores, uint8_t* obs, int64_t row_pitch,
                        uint8_t* mask, int32_t* cur_player, int auto_reset, void* stream);
int BatchStepEncodeHostPacked(pyhanabi_batch_t* batch, const int32_t* actions,
                              int32_t* rewards, uint8_t* dones, int32_t* scores,
                              uint32_t* obs_bits, uint8_t* mask, int32_t* cur_player,
                              int auto_reset, void* stream);
/* Asynchronous, ranged BatchStepEncodeHost for callers that pipeline their own host
   work: games [first_game, first_game + num_games) only (first_game % 128 == 0), the
   pointers are the bases of the whole arrays.  Returns once the work is queued; the
   range's host buffers are complete after BatchHostWait(batch, first_game)
   (first_game < 0: everything queued).  Ordering: a range runs on an internal stream
   chosen by its first game, so successive steps of the SAME range are ordered among
   themselves; work queued on other streams (BatchReset, BatchObserve on the caller's
   stream) must have completed before the first ranged call, and ranges in flight must
   not overlap. */
int BatchStepEncodeHostRange(pyhanabi_batch_t* batch, int64_t first_game, int64_t num_games,
                             const int32_t* actions, int32_t* rewards, uint8_t* dones,
                             int32_t* scores, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                             int32_t* cur_player, int auto_reset);
int BatchStepEncodeHostRangePacked(pyhanabi_batch_t* batch, int64_t first_game,
                                   int64_t num_games, const int32_t* actions, int32_t* rewards,
                                   uint8_t* dones, int32_t* scores, uint32_t* obs_bits,
                                   uint8_t* mask, int32_t* cur_player, int auto_reset);
int BatchHostWait(pyhanabi_batch_t* batch, int64_t first_game);
/* How the uint8 observation rows of the four host entry points above cross PCIe.
   1 (default; HANABI_B200_HOST_LINK=bytes selects 0): the kernel writes bit-packed rows,
   those are downloaded (96 B instead of 658 B per Hanabi-Full 2p game) into a pinned
   staging buffer of the batch, and the host worker pool expands them into the caller's
   rows -- inside BatchStepEncodeHost, or inside BatchHostWait for the ranged form, while
   later ranges are still on the link.  0: the kernel expands, the bytes cross the link.
   The caller's buffers hold the same bytes either way. */
int BatchSetHostLink(pyhanabi_batch_t* batch, int packed);
int BatchGetHostLink(pyhanabi_batch_t* batch);
/* Bit-packed rows (HOST; what BatchStepEncodeHostPacked delivers: packed_words 32-bit
   words per row, bit j of a row = observation bit j) -> HOST rows of obs_size bytes of
   0/1, row_pitch bytes apart, on the host worker pool (threads <= 0: all of it). */
int HostExpandPacked(const uint32_t* obs_bits, int packed_words, int obs_size, int64_t n,
                     uint8_t* out, int64_t row_pitch, int threads);
int BatchResetHost(pyhanabi_batch_t* batch, uint8_t* obs, int64_t row_pitch, uint8_t* mask,
                   int32_t* cur_player, void* stream);

/* Checkpoint / restore of the whole batch including every game's RNG stream
 * (HOST blobs of BatchStateBytes bytes). */
int64_t BatchStateBytes(pyhanabi_batch_t* batch);
int BatchGetState(pyhanabi_batch_t* batch, void* host_blob, int64_t nbytes);
int BatchSetState(pyhanabi_batch_t* batch, const void* host_blob, int64_t nbytes);

/* Inspection / injection of game states as int32[num_games][272] device arrays
 * (field layout in DESIGN.md; the accessors StateFireworks, StateGetHandCard,
 * ... of reference pyhanabi.h:118-131 in bulk).  extra = int32[num_games][2]
 * {turns_to_play, packed last action} or NULL. */
int BatchUnpackState(pyhanabi_batch_t* batch, int32_t* dump, void* stream);
int BatchPackState(pyhanabi_batch_t* batch, const int32_t* dump, const int32_t* extra,
                   void* stream);

/* Episode statistics {episodes, sum score, sum length, histogram[26]} as 32
 * int64, accumulated on the device by every step; dev_stats is a caller-owned
 * device buffer (to be all-reduced over NCCL) or NULL for the internal one. */
int BatchSetStatsBuffer(pyhanabi_batch_t* batch, int64_t* dev_stats);
int BatchReadStats(pyhanabi_batch_t* batch, int64_t* host_stats, int clear, void* stream);

/* ===== fused actor / learner kernels on the same batch ==================== */

/* Legal-mask-aware epsilon-greedy (DQNAgent._select_action,
 * agents/rainbow_copy/dqn_agent.py:387-420): with probability epsilon a uniform
 * legal move, else argmax_a(q[a] + legal[a]) with the first maximal index
 * (dqn_agent.py:200).  num_atoms == 0: q_or_logits is q[n][num_actions];
 * num_atoms > 0: logits[n][num_actions][num_atoms] of a C51 head and
 * q[a] = sum_i support[i] * softmax(logits[a,:])[i] (rainbow_agent.py:163-172).
 * mask[n][num_actions] is 1 for legal moves.  q_or_logits may be NULL when
 * epsilon >= 1 (uniform random legal policy).  Randomness is a counter-based
 * generator keyed (seed, step, game). */
int BatchSelectAction(const float* q_or_logits, int num_atoms, const float* support,
                      const uint8_t* mask, int n, int num_actions, float epsilon,
                      uint64_t seed, uint64_t step, int32_t* actions, void* stream);
/* BatchSelectAction for a C51 head whose logits come straight out of a GEMM:
 * dtype 0 = float32, 1 = bfloat16, 2 = float16; row_pitch = elements between two
 * games' rows (>= num_actions * num_atoms), so an N-padded GEMM output is used in
 * place. */
int BatchSelectActionC51(const void* logits, int dtype, int64_t row_pitch, int num_atoms,
                         const float* support, const uint8_t* mask, int n, int num_actions,
                         float epsilon, uint64_t seed, uint64_t step, int32_t* actions,
                         void* stream);
/* The PPO / REINFORCE actors' masked categorical policy head
 * (training/tf_agents_lib/masked_networks.py:26-50): softmax over the legal actions of
 * logits[n][row_pitch] (dtype as above), one draw per game from the counter-based
 * generator keyed (seed, step, game) -- or the mode when greedy != 0 -- and optionally
 * log pi(action | state) (float32[n], nullable). */
int BatchSampleMaskedCategorical(const void* logits, int dtype, int64_t row_pitch,
                                 const uint8_t* mask, int n, int num_actions, int greedy,
                                 uint64_t seed, uint64_t step, int32_t* actions,
                                 float* log_probs, void* stream);
/* Bit-packed observation rows -> the network's input matrix: out[i][j] = bit j of
 * row i as 0/1 in float32 / bfloat16 / float16 (dtype as above), columns
 * [obs_size, row_pitch) zero (K padded to the GEMM's alignment).  row_pitch % 8 == 0,
 * out 16-byte aligned. */
int BatchExpandPacked(const uint32_t* obs_bits, int packed_words, int obs_size, int n,
                      void* out, int64_t row_pitch, int dtype, void* stream);
/* Host-side twin of BatchSelectAction with epsilon >= 1 (the reference's
 * agents/random_agent.py for a whole batch): HOST mask in, HOST actions out,
 * identical to the device kernel for equal (seed, step, game). */
int HostRandomLegalActions(const uint8_t* mask, int n, int num_actions, uint64_t seed,
                           uint64_t step, int32_t* actions, int threads);
/* Same for rows that are games first_game_index .. first_game_index + n - 1 of the key
   (ranges of a batch, shards of a larger set).  Both run on a persistent worker pool of
   HostPolicyThreads() threads: cores / LOCAL_WORLD_SIZE, or HANABI_B200_HOST_THREADS. */
int HostRandomLegalActionsAt(const uint8_t* mask, int n, int num_actions, uint64_t seed,
                             uint64_t step, int64_t first_game_index, int32_t* actions,
                             int threads);
int HostPolicyThreads(void);
int BatchC51QValues(const float* logits, const float* support, const uint8_t* mask, int n,
                    int num_actions, int num_atoms, float* q_out, int32_t* greedy_actions,
                    void* stream);
/* project_distribution (agents/rainbow_copy/rainbow_agent.py:252-404):
 * supports[batch][num_dims], weights[batch][num_dims], target_support[num_dims]
 * -> out[batch][num_dims], fp32. */
int BatchProjectDistribution(const float* supports, const float* weights,
                             const float* target_support, float* out, int batch,
                             int num_dims, void* stream);
/* _build_target_distribution (rainbow_agent.py:181-213) fused: softmax of the
 * target net's next_logits[batch][num_actions][num_atoms], masked argmax,
 * supports = rewards + cumulative_gamma * (1 - terminals) * support, projection.
 * next_argmax may be NULL. */
int BatchC51TargetDistribution(const float* rewards, const uint8_t* terminals,
                               const float* next_logits, const uint8_t* next_legal_mask,
                               const float* support, float cumulative_gamma, float* out,
                               int32_t* next_argmax, int batch, int num_actions,
                               int num_atoms, void* stream);

/* ===== the reference's hand-written policies for a whole batch ============
 * BatchAgentAct writes actions[i] (device int32[num_games]) for every game whose
 * player to act sits in seat_mask (bit p = seat p); other games keep their entry.
 * agent_kind 0 = RuleBasedAgent.act_train (agents/rule_based_agent.py:230-296):
 *   needs BatchEnableHistory (called before the games are reset) and agent_state,
 *   a zero-initialised device uint32[num_games][num_slots] holding each agent
 *   object's rank_hinted_but_no_play table; seat_slot (HOST int[players], NULL = all
 *   seats share slot 0, as create_adhoc_team makes them share one object,
 *   training/run_adhoc_experiment.py:251-258) maps seats to slots.
 * agent_kind 1 = SimpleAgent.act (agents/simple_agent.py:32-69): stateless. */
int BatchEnableHistory(pyhanabi_batch_t* batch);
int BatchAgentAct(pyhanabi_batch_t* batch, int agent_kind, uint32_t seat_mask,
                  const int32_t* seat_slot, uint32_t* agent_state, int num_slots,
                  int32_t* actions, void* stream);

/* ===== device-resident replay memory for the Rainbow learner ==============
 * The reference's numpy ring buffer + sum tree (agents/rainbow_copy/replay_memory.py,
 * prioritized_replay_memory.py, third_party/dopamine/sum_tree.py) with stack_size 1,
 * kept in HBM.  All array arguments are DEVICE pointers.  Formats as in the
 * reference: uint8 observations, int32 actions, float32 rewards, uint8 terminals,
 * float32 legal-action vectors (0 / -inf).  New elements get priority 100
 * (DEFAULT_PRIORITY) when `priority` < 0. */
typedef struct PyHanabiReplay {
  /* Points to the device-resident replay memory. */
  void* replay;
} pyhanabi_replay_t;
int NewReplay(pyhanabi_replay_t* replay, int num_actions, int observation_size,
              int64_t capacity, int update_horizon, float gamma, int prioritized, int device);
void DeleteReplay(pyhanabi_replay_t* replay);
int64_t ReplayAddCount(pyhanabi_replay_t* replay);
int64_t ReplayCapacity(pyhanabi_replay_t* replay);
double ReplayMaxRecordedPriority(pyhanabi_replay_t* replay);
/* OutOfGraphReplayMemory.add for `count` transitions, appended in order at the cursor. */
int ReplayAdd(pyhanabi_replay_t* replay, const uint8_t* obs, const int32_t* actions,
              const float* rewards, const uint8_t* terminals, const float* legal_actions,
              int count, float priority, void* stream);
/* is_valid_transition (replay_memory.py:233-258) per index. */
int ReplayIsValidTransition(pyhanabi_replay_t* replay, const int32_t* indices, int count,
                            uint8_t* valid, void* stream);
/* SumTree.sample(query_value) per query in [0,1) (sum_tree.py:99-141). */
int ReplayTreeSample(pyhanabi_replay_t* replay, const double* query_values, int count,
                     int32_t* indices, void* stream);
/* sample_index_batch: valid indices, prioritized (optionally stratified) or uniform;
 * -1 where no valid transition was found in 1000 attempts. */
int ReplaySampleIndexBatch(pyhanabi_replay_t* replay, int count, int stratified, uint64_t seed,
                           uint64_t step, int32_t* indices, void* stream);
/* sample_transition_batch(indices=...) (replay_memory.py:290-347): n-step rewards
 * truncated at the first terminal, next state / next legal actions update_horizon
 * slots later. */
/* Elements of ReplaySampleIndexBatch calls that found no valid transition (index -1)
   since the last clear; the reference raises there (replay_memory.py:268).  Gathering
   such an index yields an all-zero terminal row, never an out-of-bounds read.
   Synchronises `stream`. */
int64_t ReplaySampleFailures(pyhanabi_replay_t* replay, int clear, void* stream);
int ReplaySampleTransitionBatch(pyhanabi_replay_t* replay, const int32_t* indices, int count,
                                uint8_t* states, int32_t* actions, float* rewards,
                                uint8_t* next_states, uint8_t* terminals,
                                float* next_legal_actions, void* stream);
int ReplaySetPriority(pyhanabi_replay_t* replay, const int32_t* indices,
                      const float* priorities, int count, float host_max_priority, void* stream);
int ReplayGetPriority(pyhanabi_replay_t* replay, const int32_t* indices, int count,
                      float* priorities, void* stream);

/* Illegal actions do not abort a batch: the game stays unchanged and is flagged.
   flags: device uint8[num_games] (nullable), count: device uint64 that the number of
   flagged games is added to (nullable), clear != 0 resets the flags. */
int BatchErrorFlags(pyhanabi_batch_t* batch, uint8_t* flags, uint64_t* count, int clear,
                    void* stream);

/* One game of a batch for the per-object API above (debugging, the GUI layer, the
   rule-based agents): BatchExportGame copies game `game` into host_record (290
   int32: the canonical state dump, turns_to_play, the last player move and -- with
   BatchEnableHistory -- the four before it and the discard pile in order);
   StateFromBatchGame rebuilds `state` (any existing state of a HanabiGame with the
   batch's parameters) from it.  Exact for everything LegalMoves, Score, the
   observations' hands/knowledge/tokens and the canonical encoding read.  With
   BatchEnableHistory also the discard pile's order (rl_env.py:416-418) and the moves an
   observer sees (HanabiObservation::LastMoves, hanabi_observation.cc:77-92: the newest
   P player moves and the deals after them); without it the pile comes back sorted by
   card type and only the last player move is kept. */
int BatchExportGame(pyhanabi_batch_t* batch, int64_t game, int32_t* host_record, void* stream);
void StateFromBatchGame(pyhanabi_game_t* game, const int32_t* host_record,
                        pyhanabi_state_t* state);

/* Keeps the packed game state resident in the L2 cache between steps (an
   access-policy window on every launch of this batch; the observation stores
   stream past it).  max_bytes: 0 = the device's maximum persisting carve-out,
   negative = off.  *covered (nullable) receives the state bytes held. */
int BatchSetL2Residency(pyhanabi_batch_t* batch, int64_t max_bytes, int64_t* covered);

/* ===== device-side self-play recorder (the transition bookkeeping of
 * DQNAgent.begin_episode / step / end_episode, dqn_agent.py:286-385, and the per-player
 * "reward since my last action" of run_one_episode, run_experiment.py:344-403) ===== */
typedef struct PyHanabiRecorder {
  void* recorder;
} pyhanabi_recorder_t;
int NewRecorder(pyhanabi_recorder_t* recorder, int num_games, int num_players,
                int observation_size, int num_actions, int max_turns_per_seat, int device);
void DeleteRecorder(pyhanabi_recorder_t* recorder);
int RecorderPackedWords(pyhanabi_recorder_t* recorder);
/* Before the environment step: the acting seat's observation (bit-packed rows of
 * BatchStepEncodePacked / BatchObservePacked), legal mask and chosen action of every
 * game.  Device buffers, asynchronous on stream. */
int RecorderBeforeStep(pyhanabi_recorder_t* recorder, const int32_t* cur_player,
                       const uint32_t* obs_bits, const uint8_t* mask, const int32_t* actions,
                       void* stream);
/* After it: rewards int32[num_games] and dones uint8[num_games] of the step.  The
 * episodes of finished games go into `replay`, each seat's episode as one contiguous
 * run (runs ordered by game, then seat), last transition terminal with the seat's
 * pending reward; *posted (host, nullable) = number of transitions appended. */
int RecorderAfterStep(pyhanabi_recorder_t* recorder, const int32_t* rewards,
                      const uint8_t* dones, pyhanabi_replay_t* replay, float priority,
                      int64_t* posted, void* stream);
int64_t RecorderDropped(pyhanabi_recorder_t* recorder, void* stream);

/* Test hooks. */
int BatchSetDebug(pyhanabi_batch_t* batch, int sampler_mode, int generic);
int BatchDebugDeal(pyhanabi_batch_t* batch, const uint64_t* decks, const int32_t* sizes,
                   const uint32_t* lo, const uint32_t* hi, int n, int mode,
                   int32_t* out_type, int32_t* out_ambiguous, void* stream);

} /* extern "C" */

#endif
