/*
 * oracle/hanabi_oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle/oracle_api.h).
 *
 * Plain-C restatement of the reference's Hanabi engine for the batched hot path
 * (apply move, deal, legal moves, score/terminal, canonical observation encoding).
 * PARITY PINNED: tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference built into oracle/_ref (same seeds, same
 * moves, every config of SURVEY.md section 8), against the reference's own
 * golden vector (gui/tests/tests_pyhanabi_to_gui.py:4-76) and against committed
 * fixtures in tests/golden/ generated from the reference.
 */
#ifndef HANABI_ORACLE_H_
#define HANABI_ORACLE_H_

#include <stdint.h>

#include "oracle_api.h"

#ifdef __cplusplus
extern "C" {
#endif

void* OrcBatchNew(int n_kv, const char** kv, int num_games, const int32_t* seeds);
void OrcBatchDelete(void* b);
int OrcBatchObservationSize(void* b);
int OrcBatchMaxMoves(void* b);
int OrcBatchNumPlayers(void* b);
int OrcBatchNumGames(void* b);
void OrcBatchReset(void* b, const uint8_t* which);
void OrcBatchStep(void* b, const int32_t* actions, int32_t* rewards, uint8_t* dones,
                  int32_t* scores, int auto_reset, uint8_t* illegal);
void OrcBatchEncode(void* b, int observer, uint8_t* obs, int64_t pitch);
void OrcBatchLegalMask(void* b, uint8_t* mask);
void OrcBatchCurPlayer(void* b, int32_t* cur);
void OrcBatchDumpState(void* b, int32_t* out);
/* Overwrite game i from a canonical dump (oracle_api.h) plus the hidden fields
 * the dump does not carry; used to inject hand-built states (golden vector). */
void OrcBatchLoadState(void* b, int game, const int32_t* dump, int next_player,
                       int turns_to_play);
int64_t OrcBatchRollout(void* b, int steps, uint8_t* obs, int64_t pitch, uint8_t* mask,
                        int64_t* episodes_out, int64_t* score_sum_out);

int OrcDiscreteSample(const double* weights, int n, uint32_t lo, uint32_t hi);
int OrcDealFromCounts(const int32_t* counts, int n_types, uint32_t lo, uint32_t hi);
void OrcMtWords(int32_t seed, int n, uint32_t* out);
void OrcUniformInts(int32_t seed, int hi, int n, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif
