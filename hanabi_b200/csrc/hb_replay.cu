// hb_replay.cu -- device-resident replay memory for the Rainbow learner
// (SURVEY.md section 8f, row f3): the reference's numpy ring buffer and sum tree
//   agents/rainbow_copy/replay_memory.py:67-347            OutOfGraphReplayMemory
//   agents/rainbow_copy/prioritized_replay_memory.py:41-170 OutOfGraphPrioritizedReplayMemory
//   agents/rainbow_copy/third_party/dopamine/sum_tree.py:30-205  SumTree
// with the Hanabi configuration stack_size = 1 (training/configs/hanabi_rainbow.gin:39),
// kept in HBM so that the batched actor can append and the learner can sample
// without a host round trip.  Same field formats as the reference: uint8
// observations, int32 actions, float32 rewards, uint8 terminals, float32 legal
// action vectors (0 / -inf), float64 priorities.
//
// Semantics kept from the reference, including the odd ones:
//  * transitions are appended in order at the cursor, next_state is "whatever is
//    stored update_horizon slots later" (replay_memory.py:326-331);
//  * is_valid_transition (:233-258) with stack_size 1: index in range, not the most
//    recently written slot, and -- only while the buffer is not full -- at least
//    update_horizon slots behind the cursor;
//  * the n-step reward stops at the first terminal in the window (:307-324);
//  * new elements enter the sum tree with priority 100 (DEFAULT_PRIORITY, :38).
// Difference on purpose: inner sum-tree nodes are recomputed as left + right after
// every batch of updates (deterministic, one rounding per node) instead of being
// nudged by += delta per update; leaves are identical, inner sums agree to ~1e-13.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <string>

#include "../../include/pyhanabi.h"

int HbFail(const std::string& msg);  // hb_batch.cu

namespace hbr {

constexpr int kMaxHorizon = 16;
constexpr int kMaxAttempts = 1000;  // MAX_SAMPLE_ATTEMPTS, replay_memory.py:36

struct Replay {
  int device = 0;
  int A = 0, D = 0, horizon = 1, prioritized = 0, depth = 0;
  int64_t cap = 0, add_count = 0;
  float gamma = 1.f;
  float disc[kMaxHorizon];
  uint8_t* obs = nullptr;
  int32_t* actions = nullptr;
  float* rewards = nullptr;
  uint8_t* terminals = nullptr;
  float* legal = nullptr;
  double* tree = nullptr;  // heap layout: level l (2^l nodes) starts at 2^l - 1; leaves at level `depth`
  double max_recorded_priority = 1.0;
  // Device copy of add_count ([0]) for appends that are driven entirely from the device
  // (the recorder's asynchronous posting); `dirty` = the host value is behind it.
  unsigned long long* d_counts = nullptr;
  unsigned long long* h_counts = nullptr;  // pinned
  bool dirty = false;
};

struct View {
  int A, D, horizon, depth;
  int64_t cap, add_count;
  float disc[kMaxHorizon];
  const uint8_t* obs;
  const int32_t* actions;
  const float* rewards;
  const uint8_t* terminals;
  const float* legal;
  const double* tree;
  unsigned long long* sample_failures;  // elements for which sample_index_batch found no valid transition
};

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// is_valid_transition, replay_memory.py:233-258 (stack_size == 1).
__device__ inline bool valid_transition(const View& v, int64_t index) {
  if (index < 0 || index >= v.cap) return false;
  const int64_t cursor = v.add_count % v.cap;
  if (v.add_count < v.cap) {
    if (index >= cursor - v.horizon) return false;
  }
  const int64_t newest = (cursor - 1 + v.cap) % v.cap;  // invalid_range(cursor, capacity, 1)
  return index != newest;
}

// SumTree.sample(query_value), sum_tree.py:99-141.
__device__ inline int64_t tree_sample(const View& v, double query) {
  query *= v.tree[0];
  int64_t node = 0;
  for (int l = 1; l <= v.depth; ++l) {
    const int64_t left = node * 2;
    const double left_sum = v.tree[((int64_t)1 << l) - 1 + left];
    if (query < left_sum) {
      node = left;
    } else {
      node = left + 1;
      query -= left_sum;
    }
  }
  return node;
}

__global__ void k_set_leaves_range(double* tree, int depth, int64_t cap, int64_t cursor, int count,
                                   double value) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  tree[((int64_t)1 << depth) - 1 + (cursor + i) % cap] = value;
}

// Batch of SumTree.set calls: for duplicate indices the last one in the batch wins,
// as it would sequentially (writer[] holds the winning position per leaf).
__global__ void k_claim(int32_t* writer, const int32_t* indices, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) atomicMax(&writer[indices[i]], i);
}
__global__ void k_set_leaves(double* tree, int depth, int32_t* writer, const int32_t* indices,
                             const float* values, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int32_t leaf = indices[i];
  if (writer[leaf] == i) tree[((int64_t)1 << depth) - 1 + leaf] = (double)values[i];
}
__global__ void k_unclaim(int32_t* writer, const int32_t* indices, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) writer[indices[i]] = -1;
}

// Recompute level `l` parents of the touched leaves: parent = left + right.
__global__ void k_rebuild_level_range(double* tree, int depth, int l, int64_t cap, int64_t cursor,
                                      int count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t leaf = (cursor + i) % cap;
  const int64_t node = leaf >> (depth - l);
  const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
  tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
}
// ---- the same updates with cursor and count read from device memory (no host round
// trip): the leaf range [cursor, cursor + total) mod capacity is one or two contiguous
// pieces; launches use a fixed grid and stride over whatever the device says.
struct Pieces {
  int64_t a[2], n[2];
};
__device__ __forceinline__ Pieces range_pieces(const unsigned long long* counts, const unsigned long long* total,
                                               int64_t cap) {
  Pieces p;
  const int64_t cursor = (int64_t)(counts[0] % (unsigned long long)cap);
  int64_t t = (int64_t)*total;
  if (t > cap) t = cap;
  p.a[0] = cursor; p.n[0] = t < cap - cursor ? t : cap - cursor;
  p.a[1] = 0; p.n[1] = t - p.n[0];
  return p;
}
__global__ void k_set_leaves_dev(double* tree, int depth, int64_t cap, const unsigned long long* counts,
                                 const unsigned long long* total, double value) {
  const Pieces p = range_pieces(counts, total, cap);
  const int64_t t = p.n[0] + p.n[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x)
    tree[((int64_t)1 << depth) - 1 + (p.a[0] + i) % cap] = value;
}
__global__ void k_rebuild_span_dev(double* tree, int depth, int l, int64_t cap, const unsigned long long* counts,
                                   const unsigned long long* total) {
  const Pieces p = range_pieces(counts, total, cap);
  const int sh = depth - l;
  const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
  for (int piece = 0; piece < 2; ++piece) {
    if (p.n[piece] <= 0) continue;
    const int64_t first = p.a[piece] >> sh, cnt = ((p.a[piece] + p.n[piece] - 1) >> sh) - first + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t node = first + i;
      tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
    }
  }
}
__global__ void k_rebuild_top_dev(double* tree, int depth, int l_start, int64_t cap, const unsigned long long* counts,
                                  const unsigned long long* total) {
  const Pieces p = range_pieces(counts, total, cap);
  if (p.n[0] + p.n[1] <= 0) return;
  for (int l = l_start; l >= 0; --l) {
    const int sh = depth - l;
    const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
    for (int piece = 0; piece < 2; ++piece) {
      if (p.n[piece] <= 0) continue;
      const int64_t first = p.a[piece] >> sh, cnt = ((p.a[piece] + p.n[piece] - 1) >> sh) - first + 1;
      for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int64_t node = first + i;
        tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
      }
    }
    __syncthreads();
  }
}
__global__ void k_commit_count(unsigned long long* counts, const unsigned long long* total, int64_t cap) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t = *total;
    if (t > (unsigned long long)cap) t = (unsigned long long)cap;
    counts[0] += t;
  }
}
__global__ void k_store_count(unsigned long long* counts, unsigned long long value) {
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = value;
}

// Contiguous leaf range: level l touches the contiguous nodes [first, first + count).
__global__ void k_rebuild_span(double* tree, int l, int64_t first, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t node = first + i;
  const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
  tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
}
// The upper levels of the same range (at most blockDim.x nodes each) in one launch:
// one block walks from level l_start up to the root.
__global__ void k_rebuild_top(double* tree, int depth, int l_start, int64_t first_leaf, int64_t last_leaf) {
  for (int l = l_start; l >= 0; --l) {
    const int sh = depth - l;
    const int64_t first = first_leaf >> sh, count = (last_leaf >> sh) - first + 1;
    const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
    for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
      const int64_t node = first + i;
      tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
    }
    __syncthreads();
  }
}
__global__ void k_rebuild_level(double* tree, int depth, int l, const int32_t* indices, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t node = (int64_t)indices[i] >> (depth - l);
  const double* child = tree + (((int64_t)1 << (l + 1)) - 1);
  tree[((int64_t)1 << l) - 1 + node] = child[2 * node] + child[2 * node + 1];
}

__global__ void k_get_priority(View v, const int32_t* indices, int count, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t idx = indices[i];  // -1 (failed draw) or any other out-of-range index reads nothing
  out[i] = (idx >= 0 && idx < v.cap) ? (float)v.tree[((int64_t)1 << v.depth) - 1 + idx] : 0.f;
}

__global__ void k_is_valid(View v, const int32_t* indices, int count, uint8_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = valid_transition(v, indices[i]) ? 1 : 0;
}

__global__ void k_tree_sample(View v, const double* queries, int count, int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (int32_t)tree_sample(v, queries[i]);
}

// sample_index_batch: prioritized (prioritized_replay_memory.py:109-132) or
// uniform (replay_memory.py:265-288), each element retried until valid.
__global__ void k_sample_indices(View v, int prioritized, int stratified, uint64_t seed, uint64_t step,
                                 int count, int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int64_t pick = -1;
  for (int attempt = 0; attempt < kMaxAttempts && pick < 0; ++attempt) {
    const uint64_t bits = mix64(mix64(seed ^ (step * 0xD1B54A32D192ED03ull)) ^
                                ((uint64_t)i * 0x8CB92BA72F3D8DD7ull) ^ ((uint64_t)attempt << 48));
    const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
    int64_t index;
    if (prioritized) {
      // stratified: element i draws from [i/count, (i+1)/count) (sum_tree.py:143-168)
      const double q = stratified ? ((double)i + u) / (double)count : u;
      index = tree_sample(v, q);
    } else if (v.add_count >= v.cap) {
      index = (int64_t)(u * (double)v.cap);                      // np.random.randint(0, capacity)
    } else {
      const int64_t hi = v.add_count % v.cap - 1;                // randint(stack_size - 1, cursor - 1)
      index = hi > 0 ? (int64_t)(u * (double)hi) : 0;
    }
    if (valid_transition(v, index)) pick = index;
  }
  out[i] = (int32_t)pick;  // -1: no valid transition found (the reference raises; see ReplaySampleFailures)
  if (pick < 0) atomicAdd(v.sample_failures, 1ull);
}

// sample_transition_batch(indices=...), replay_memory.py:290-347.  One warp per element.
__global__ void k_gather(View v, const int32_t* indices, int count, uint8_t* states, int32_t* actions,
                         float* rewards, uint8_t* next_states, uint8_t* terminals, float* next_legal) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= count) return;
  const int64_t idx = indices[b];
  if (idx < 0 || idx >= v.cap) {
    // a failed draw (-1) or a foreign index: an all-zero, terminal row instead of an
    // out-of-bounds read; the caller learns about failed draws from ReplaySampleFailures
    for (int k = lane; k < v.D; k += 32) { states[b * v.D + k] = 0; next_states[b * v.D + k] = 0; }
    for (int k = lane; k < v.A; k += 32) next_legal[b * v.A + k] = 0.f;
    if (lane == 0) { actions[b] = 0; rewards[b] = 0.f; terminals[b] = 1; }
    return;
  }
  const int64_t boot = (idx + v.horizon) % v.cap;
  for (int k = lane; k < v.D; k += 32) {
    states[b * v.D + k] = v.obs[idx * v.D + k];
    next_states[b * v.D + k] = v.obs[boot * v.D + k];
  }
  for (int k = lane; k < v.A; k += 32) next_legal[b * v.A + k] = v.legal[boot * v.A + k];
  if (lane == 0) {
    actions[b] = v.actions[idx];
    float acc = 0.f;
    uint8_t term = 0;
    for (int j = 0; j < v.horizon; ++j) {
      const int64_t t = (idx + j) % v.cap;
      acc = __fadd_rn(acc, __fmul_rn(v.disc[j], v.rewards[t]));
      if (v.terminals[t]) { term = 1; break; }  // truncated at the first terminal (inclusive)
    }
    rewards[b] = acc;
    terminals[b] = term;
  }
}


// ---- self-play recorder -------------------------------
// Per (game, seat) episode buffers on the device: what DQNAgent.begin_episode / step /
// end_episode keep around the network (agents/rainbow_copy/dqn_agent.py:286-385) and
// what run_one_episode does with the rewards (agents/run_experiment.py:344-403), for N
// games at once.  Observations are kept bit-packed (16-byte multiples) and expanded to
// the replay memory's uint8 rows when an episode is posted.
struct RecView {
  int N, P, L, W, A, D;
  uint32_t* obs;     // [N][P][L][W]
  uint64_t* legal;   // [N][P][L]
  int32_t* action;   // [N][P][L]
  float* reward;     // [N][P][L]
  int32_t* count;    // [N][P]
  float* since;      // [N][P]  reward since the seat's last action
  int32_t* run_len;  // [N][P]  snapshot of a finished game's episode lengths
  float* run_since;  // [N][P]
  int32_t* lens;     // [N][P] run lengths, scanned in place into the runs' offsets
  int32_t* block_sums;
  int32_t* owner;    // [owner_cap] run (game * P + seat) of every transition posted this step
  int64_t owner_cap;
  unsigned long long* total;    // transitions posted this step
  unsigned long long* dropped;  // transitions beyond L turns of a seat (never, for L = 64)
};

// begin_episode / step -> _record_transition for the acting seat of every game
// (dqn_agent.py:286-357): the reward handed to step() closes the seat's previous
// transition, then (o_t, l_t, a_t) opens a new one.
__global__ void k_rec_before(RecView v, const int32_t* __restrict__ cur, const uint32_t* __restrict__ bits,
                             const uint8_t* __restrict__ mask, const int32_t* __restrict__ actions) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= v.N) return;
  const int c = cur[g];
  if (c < 0 || c >= v.P) return;
  const int64_t gp = (int64_t)g * v.P + c;
  const int k = v.count[gp];
  if (k > 0 && k <= v.L) v.reward[gp * v.L + (k - 1)] = v.since[gp];
  v.since[gp] = 0.f;
  if (k >= v.L) {
    atomicAdd(v.dropped, 1ull);
    v.count[gp] = k + 1;
    return;
  }
  const uint4* src = reinterpret_cast<const uint4*>(bits + (int64_t)g * v.W);
  uint4* dst = reinterpret_cast<uint4*>(v.obs + (gp * v.L + k) * v.W);
  for (int j = 0; j < v.W / 4; ++j) dst[j] = src[j];
  uint64_t legal = 0;
  const uint8_t* m = mask + (int64_t)g * v.A;
  if ((v.A & 3) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3) == 0) {
    const uint32_t* m4 = reinterpret_cast<const uint32_t*>(m);
    for (int j = 0; j < (v.A >> 2); ++j) {
      uint32_t wv = m4[j];
      wv |= wv >> 4; wv |= wv >> 2; wv |= wv >> 1;  // byte != 0 -> bit 0 of the byte
      wv &= 0x01010101u;
      legal |= (uint64_t)((wv * 0x10204080u) >> 28) << (4 * j);  // bits 0, 8, 16, 24 -> 4 bits
    }
  } else {
    for (int a = 0; a < v.A; ++a) legal |= (uint64_t)(m[a] != 0) << a;
  }
  v.legal[gp * v.L + k] = legal;
  v.action[gp * v.L + k] = actions[g];
  v.count[gp] = k + 1;
}

// After the environment step: every seat's pending reward grows by the step's reward
// (run_experiment.py:374-377); a finished game snapshots its episodes for posting and
// starts empty.
__global__ void k_rec_after(RecView v, const int32_t* __restrict__ rewards, const uint8_t* __restrict__ dones) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= v.N) return;
  const float r = (float)rewards[g];
  const bool done = dones[g] != 0;
  for (int p = 0; p < v.P; ++p) {
    const int64_t gp = (int64_t)g * v.P + p;
    const float s = v.since[gp] + r;
    int k = 0;
    if (done) {
      k = v.count[gp] < v.L ? v.count[gp] : v.L;
      v.run_since[gp] = s;
      v.count[gp] = 0;
      v.since[gp] = 0.f;
    } else {
      v.since[gp] = s;
    }
    v.run_len[gp] = k;
    v.lens[gp] = k;
  }
}

// Exclusive prefix sum of lens[N] in three small passes (1024 elements per block).
__device__ __forceinline__ int block_exclusive_scan(int x, int* warp_sums, int& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    warp_sums[lane] = wi - w;          // exclusive offset of each warp
    if (lane == 31) warp_sums[32] = wi;  // block total
  }
  __syncthreads();
  block_total = warp_sums[32];
  return incl - x + warp_sums[warp];
}
__global__ void k_scan_blocks(int32_t* data, int n, int32_t* block_sums) {
  __shared__ int warp_sums[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = i < n ? data[i] : 0;
  int total;
  const int ex = block_exclusive_scan(x, warp_sums, total);
  if (i < n) data[i] = ex;
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void k_scan_sums(int32_t* block_sums, int nblocks, unsigned long long* total_out) {
  __shared__ int warp_sums[33];
  int carry = 0;
  for (int base = 0; base < nblocks; base += blockDim.x) {  // one block; loops when > 1024 block sums
    const int i = base + threadIdx.x;
    const int x = i < nblocks ? block_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(x, warp_sums, total);
    if (i < nblocks) block_sums[i] = ex + carry;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = (unsigned long long)carry;
}
__global__ void k_scan_add(int32_t* data, int n, const int32_t* block_sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) data[i] += block_sums[blockIdx.x];
}

// Which run every posted transition belongs to: the non-empty runs write their index into
// the slots [offset, offset + length) of the owner table.
__global__ void k_rec_owner(RecView v) {
  const int64_t gp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gp >= (int64_t)v.N * v.P) return;
  const int len = v.run_len[gp];
  if (len <= 0) return;
  const int64_t off = v.lens[gp];
  for (int k = 0; k < len; ++k)
    if (off + k < v.owner_cap) v.owner[off + k] = (int32_t)gp;
}

// _post_transitions (dqn_agent.py:359-385) for the finished games: every seat's episode
// becomes a contiguous run of replay slots -- runs ordered by game, then seat -- of
// (o_t, a_t, r_{t+1}, terminal, l_t); the last transition carries the seat's pending
// reward and the terminal flag.  One warp per posted transition (strided over a fixed
// grid; cursor and count come from device memory): it expands the packed observation row
// into the ring's uint8 row with 32-bit stores and writes the scalars.
__global__ void k_rec_post(RecView v, const unsigned long long* counts, int64_t cap, uint8_t* r_obs,
                           int32_t* r_actions, float* r_rewards, uint8_t* r_terminals, float* r_legal) {
  const int lane = threadIdx.x & 31;
  int64_t total = (int64_t)*v.total;
  if (total > cap) total = cap;  // more than the ring holds: the tail is dropped (never in practice)
  if (total > v.owner_cap) total = v.owner_cap;
  const int64_t cursor = (int64_t)(counts[0] % (unsigned long long)cap);
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += n_warps) {
    const int64_t gp = v.owner[w];
    const int k = (int)(w - v.lens[gp]);
    const int len = v.run_len[gp];
    const int64_t slot = (cursor + w) % cap;
    const uint32_t* row = v.obs + (gp * v.L + k) * v.W;
    uint8_t* dst = r_obs + slot * v.D;
    // bytes up to the first 4-byte boundary, then whole words, then the tail
    const int lead = (int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    const int head = lead < v.D ? lead : v.D;
    if (lane < head) dst[lane] = (uint8_t)((row[0] >> lane) & 1u);
    const int words = (v.D - head) >> 2;
    uint32_t* dst4 = reinterpret_cast<uint32_t*>(dst + head);
    for (int j = lane; j < words; j += 32) {
      const int bit = head + 4 * j;  // 4 consecutive bits, possibly across two row words
      const int wi = bit >> 5;
      const uint32_t lo = row[wi], hi = row[(wi + 1 < v.W && (bit & 31) > 28) ? wi + 1 : wi];
      const uint32_t four = __funnelshift_r(lo, hi, bit & 31) & 15u;
      dst4[j] = (four * 0x00204081u) & 0x01010101u;
    }
    for (int b = head + 4 * words + lane; b < v.D; b += 32) dst[b] = (uint8_t)((row[b >> 5] >> (b & 31)) & 1u);
    const uint64_t legal = v.legal[gp * v.L + k];
    for (int a = lane; a < v.A; a += 32)
      r_legal[slot * v.A + a] = ((legal >> a) & 1ull) ? 0.f : -CUDART_INF_F;  // format_legal_moves
    if (lane == 0) {
      const bool last = k == len - 1;
      r_actions[slot] = v.action[gp * v.L + k];
      r_rewards[slot] = last ? v.run_since[gp] : v.reward[gp * v.L + k];
      r_terminals[slot] = last ? 1 : 0;
    }
  }
}

}  // namespace hbr

namespace {
using hbr::Replay;

struct Guard {
  int prev = -1;
  explicit Guard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~Guard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

hbr::View MakeView(const Replay* r) {
  hbr::View v;
  v.A = r->A; v.D = r->D; v.horizon = r->horizon; v.depth = r->depth;
  v.cap = r->cap; v.add_count = r->add_count;
  std::memcpy(v.disc, r->disc, sizeof(v.disc));
  v.obs = r->obs; v.actions = r->actions; v.rewards = r->rewards; v.terminals = r->terminals;
  v.legal = r->legal; v.tree = r->tree;
  v.sample_failures = r->d_counts + 2;
  return v;
}

int Check(cudaError_t e, const char* what) {
  return e == cudaSuccess ? 0 : HbFail(std::string(what) + ": " + cudaGetErrorString(e));
}
#define HBR_CUDA(expr)                       \
  do {                                       \
    if (Check((expr), #expr)) return 1;      \
  } while (0)
// Brings the host's add_count up to date after device-driven appends.
int SyncAddCount(Replay* r) {
  if (!r->dirty) return 0;
  cudaError_t e = cudaDeviceSynchronize();  // the appends may sit on any (non-blocking) stream
  if (e == cudaSuccess)
    e = cudaMemcpy(r->h_counts, r->d_counts, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return Check(e, "replay add_count read-back");
  r->add_count = (int64_t)r->h_counts[0];
  r->dirty = false;
  return 0;
}
#define HBR_REPLAY(h)                                             \
  if (!(h) || !(h)->replay) return HbFail("null replay handle");  \
  Replay* r = static_cast<Replay*>((h)->replay);                  \
  Guard guard__(r->device);                                       \
  if (SyncAddCount(r)) return 1
}  // namespace

extern "C" {

int NewReplay(pyhanabi_replay_t* out, int num_actions, int observation_size, int64_t capacity,
              int update_horizon, float gamma, int prioritized, int device) {
  if (!out) return HbFail("NewReplay: null handle");
  out->replay = nullptr;
  if (num_actions <= 0 || observation_size <= 0 || capacity <= 0)
    return HbFail("NewReplay: sizes must be positive");
  if (update_horizon < 1 || update_horizon > hbr::kMaxHorizon)
    return HbFail("NewReplay: update_horizon must be in [1,16]");
  if (capacity > (int64_t)1 << 30) return HbFail("NewReplay: capacity above 2^30");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return HbFail("NewReplay: no such CUDA device (the replay memory has no CPU fallback)");
  Replay* r = new Replay;
  r->device = device; r->A = num_actions; r->D = observation_size; r->cap = capacity;
  r->horizon = update_horizon; r->gamma = gamma; r->prioritized = prioritized ? 1 : 0;
  for (int j = 0; j < hbr::kMaxHorizon; ++j)
    r->disc[j] = (float)std::pow((double)gamma, (double)j);  // math.pow -> float32, replay_memory.py:106-108
  r->depth = 0;
  while (((int64_t)1 << r->depth) < capacity) ++r->depth;  // ceil(log2(capacity)), sum_tree.py:75
  Guard guard(device);
  const size_t cap = (size_t)capacity;
  auto fail = [&](const char* what, cudaError_t e) {
    cudaFree(r->obs); cudaFree(r->actions); cudaFree(r->rewards); cudaFree(r->terminals);
    cudaFree(r->legal); cudaFree(r->tree); cudaFree(r->d_counts);
    if (r->h_counts) cudaFreeHost(r->h_counts);
    delete r;
    return Check(e, what);
  };
  cudaError_t e;
  if ((e = cudaMalloc(&r->obs, cap * (size_t)observation_size)) != cudaSuccess) return fail("cudaMalloc obs", e);
  if ((e = cudaMalloc(&r->actions, cap * sizeof(int32_t))) != cudaSuccess) return fail("cudaMalloc actions", e);
  if ((e = cudaMalloc(&r->rewards, cap * sizeof(float))) != cudaSuccess) return fail("cudaMalloc rewards", e);
  if ((e = cudaMalloc(&r->terminals, cap)) != cudaSuccess) return fail("cudaMalloc terminals", e);
  if ((e = cudaMalloc(&r->legal, cap * (size_t)num_actions * sizeof(float))) != cudaSuccess) return fail("cudaMalloc legal", e);
  const size_t nodes = ((size_t)2 << r->depth);  // 2^(depth+1) - 1 nodes, + 1 writer-free pad
  // the tree buffer also hosts the int32 "last writer" scratch behind the nodes
  if ((e = cudaMalloc(&r->tree, nodes * sizeof(double) + ((size_t)1 << r->depth) * sizeof(int32_t))) != cudaSuccess)
    return fail("cudaMalloc tree", e);
  if ((e = cudaMalloc(&r->d_counts, 4 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMalloc counters", e);
  if ((e = cudaMallocHost(&r->h_counts, 2 * sizeof(unsigned long long))) != cudaSuccess) return fail("cudaMallocHost counters", e);
  cudaMemset(r->d_counts, 0, 4 * sizeof(unsigned long long));
  cudaMemset(r->obs, 0, cap * (size_t)observation_size);
  cudaMemset(r->actions, 0, cap * sizeof(int32_t));
  cudaMemset(r->rewards, 0, cap * sizeof(float));
  cudaMemset(r->terminals, 0, cap);
  cudaMemset(r->legal, 0, cap * (size_t)num_actions * sizeof(float));
  cudaMemset(r->tree, 0, nodes * sizeof(double));
  cudaMemset(reinterpret_cast<char*>(r->tree) + nodes * sizeof(double), 0xff, ((size_t)1 << r->depth) * sizeof(int32_t));
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail("NewReplay init", e);
  out->replay = r;
  return 0;
}

void DeleteReplay(pyhanabi_replay_t* h) {
  if (!h || !h->replay) return;
  Replay* r = static_cast<Replay*>(h->replay);
  Guard guard(r->device);
  cudaFree(r->obs); cudaFree(r->actions); cudaFree(r->rewards); cudaFree(r->terminals);
  cudaFree(r->legal); cudaFree(r->tree); cudaFree(r->d_counts);
  if (r->h_counts) cudaFreeHost(r->h_counts);
  delete r;
  h->replay = nullptr;
}

int64_t ReplayAddCount(pyhanabi_replay_t* h) {
  if (!h || !h->replay) return -1;
  Replay* r = static_cast<Replay*>(h->replay);
  Guard guard(r->device);
  if (SyncAddCount(r)) return -1;
  return r->add_count;
}
int64_t ReplayCapacity(pyhanabi_replay_t* h) { return (h && h->replay) ? static_cast<Replay*>(h->replay)->cap : -1; }
double ReplayMaxRecordedPriority(pyhanabi_replay_t* h) {
  return (h && h->replay) ? static_cast<Replay*>(h->replay)->max_recorded_priority : -1.0;
}

// Parents of the leaves [cursor, cursor + count) mod capacity: the range is one or two
// contiguous leaf spans; per level the touched nodes are contiguous too, and once a
// level has at most 1024 of them one block finishes the walk to the root.
static int RebuildRange(Replay* r, int64_t cursor, int count, cudaStream_t s) {
  const int64_t first_len = (cursor + count <= r->cap) ? count : (r->cap - cursor);
  for (int piece = 0; piece < 2; ++piece) {
    const int64_t a = piece == 0 ? cursor : 0;
    const int64_t n = piece == 0 ? first_len : count - first_len;
    if (n <= 0) continue;
    const int64_t b = a + n - 1;
    for (int l = r->depth - 1; l >= 0; --l) {
      const int sh = r->depth - l;
      const int64_t first = a >> sh, cnt = (b >> sh) - first + 1;
      if (cnt > 1024) {
        hbr::k_rebuild_span<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(r->tree, l, first, cnt);
      } else {
        hbr::k_rebuild_top<<<1, 1024, 0, s>>>(r->tree, r->depth, l, a, b);
        break;
      }
    }
  }
  return Check(cudaGetLastError(), "sum tree rebuild");
}

int ReplayAdd(pyhanabi_replay_t* h, const uint8_t* obs, const int32_t* actions, const float* rewards,
              const uint8_t* terminals, const float* legal_actions, int count, float priority,
              void* stream) {
  HBR_REPLAY(h);
  if (!obs || !actions || !rewards || !terminals || !legal_actions) return HbFail("ReplayAdd: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  // A batch longer than the ring overwrites its own head, like sequential adds would:
  // feed it in stream-ordered pieces of at most `capacity` transitions.
  for (int done = 0; done < count;) {
    const int part = (int)((int64_t)(count - done) < r->cap ? (count - done) : r->cap);
    const int64_t cursor = r->add_count % r->cap;
    // consecutive slots are contiguous memory (two pieces when the ring wraps): the
    // append is a set of device-to-device copies at full copy bandwidth
    const int64_t first = (cursor + part <= r->cap) ? part : (r->cap - cursor);
    for (int piece = 0; piece < 2; ++piece) {
      const int64_t cnt = piece == 0 ? first : part - first;
      if (cnt <= 0) continue;
      const int64_t dst = piece == 0 ? cursor : 0, src = done + (piece == 0 ? 0 : first);
      HBR_CUDA(cudaMemcpyAsync(r->obs + dst * r->D, obs + src * r->D, (size_t)cnt * r->D, cudaMemcpyDeviceToDevice, s));
      HBR_CUDA(cudaMemcpyAsync(r->legal + dst * r->A, legal_actions + src * r->A, (size_t)cnt * r->A * sizeof(float), cudaMemcpyDeviceToDevice, s));
      HBR_CUDA(cudaMemcpyAsync(r->actions + dst, actions + src, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
      HBR_CUDA(cudaMemcpyAsync(r->rewards + dst, rewards + src, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToDevice, s));
      HBR_CUDA(cudaMemcpyAsync(r->terminals + dst, terminals + src, (size_t)cnt, cudaMemcpyDeviceToDevice, s));
    }
    r->add_count += part;
    hbr::k_store_count<<<1, 1, 0, s>>>(r->d_counts, (unsigned long long)r->add_count);
    if (r->prioritized) {
      const double p = priority < 0.f ? 100.0 : (double)priority;  // DEFAULT_PRIORITY
      if (p > r->max_recorded_priority) r->max_recorded_priority = p;
      hbr::k_set_leaves_range<<<(part + 255) / 256, 256, 0, s>>>(r->tree, r->depth, r->cap, cursor, part, p);
      HBR_CUDA(cudaGetLastError());
      if (RebuildRange(r, cursor, part, s)) return 1;
    }
    done += part;
  }
  return 0;
}

int ReplaySetPriority(pyhanabi_replay_t* h, const int32_t* indices, const float* priorities, int count,
                      float host_max_priority, void* stream) {
  HBR_REPLAY(h);
  if (!r->prioritized) return HbFail("ReplaySetPriority: the replay memory is not prioritized");
  if (count <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* writer = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(r->tree) + ((size_t)2 << r->depth) * sizeof(double));
  const int g = (count + 255) / 256;
  hbr::k_claim<<<g, 256, 0, s>>>(writer, indices, count);
  hbr::k_set_leaves<<<g, 256, 0, s>>>(r->tree, r->depth, writer, indices, priorities, count);
  hbr::k_unclaim<<<g, 256, 0, s>>>(writer, indices, count);
  for (int l = r->depth - 1; l >= 0; --l)
    hbr::k_rebuild_level<<<g, 256, 0, s>>>(r->tree, r->depth, l, indices, count);
  HBR_CUDA(cudaGetLastError());
  if ((double)host_max_priority > r->max_recorded_priority) r->max_recorded_priority = host_max_priority;
  return 0;
}

int ReplayGetPriority(pyhanabi_replay_t* h, const int32_t* indices, int count, float* priorities, void* stream) {
  HBR_REPLAY(h);
  if (!r->prioritized) return HbFail("ReplayGetPriority: the replay memory is not prioritized");
  hbr::k_get_priority<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(MakeView(r), indices, count, priorities);
  HBR_CUDA(cudaGetLastError());
  return 0;
}

int ReplayIsValidTransition(pyhanabi_replay_t* h, const int32_t* indices, int count, uint8_t* valid, void* stream) {
  HBR_REPLAY(h);
  hbr::k_is_valid<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(MakeView(r), indices, count, valid);
  HBR_CUDA(cudaGetLastError());
  return 0;
}

int ReplayTreeSample(pyhanabi_replay_t* h, const double* query_values, int count, int32_t* indices, void* stream) {
  HBR_REPLAY(h);
  if (!r->prioritized) return HbFail("ReplayTreeSample: the replay memory is not prioritized");
  hbr::k_tree_sample<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(MakeView(r), query_values, count, indices);
  HBR_CUDA(cudaGetLastError());
  return 0;
}

int ReplaySampleIndexBatch(pyhanabi_replay_t* h, int count, int stratified, uint64_t seed, uint64_t step,
                           int32_t* indices, void* stream) {
  HBR_REPLAY(h);
  // no transition can be valid before update_horizon + 1 adds (is_valid_transition,
  // replay_memory.py:233-258); the reference raises there (replay_memory.py:268)
  if (r->add_count <= r->horizon)
    return HbFail("ReplaySampleIndexBatch: the replay memory holds no valid transition yet "
                  "(needs more than update_horizon adds)");
  if (count <= 0 || !indices) return HbFail("ReplaySampleIndexBatch: bad argument");
  hbr::k_sample_indices<<<(count + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      MakeView(r), r->prioritized, stratified, seed, step, count, indices);
  HBR_CUDA(cudaGetLastError());
  return 0;
}

// Number of elements for which ReplaySampleIndexBatch found no valid transition in its
// 1000 attempts (they hold -1) since the counter was last cleared; the reference raises a
// RuntimeError in that case (replay_memory.py:268, prioritized_replay_memory.py:128).
// Synchronises `stream`.
int64_t ReplaySampleFailures(pyhanabi_replay_t* h, int clear, void* stream) {
  if (!h || !h->replay) return -1;
  Replay* r = static_cast<Replay*>(h->replay);
  Guard guard(r->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaMemcpyAsync(r->h_counts + 1, r->d_counts + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return -1;
  if (clear && cudaMemsetAsync(r->d_counts + 2, 0, sizeof(unsigned long long), s) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
  return (int64_t)r->h_counts[1];
}

int ReplaySampleTransitionBatch(pyhanabi_replay_t* h, const int32_t* indices, int count, uint8_t* states,
                                int32_t* actions, float* rewards, uint8_t* next_states,
                                uint8_t* terminals, float* next_legal_actions, void* stream) {
  HBR_REPLAY(h);
  if (!indices || !states || !actions || !rewards || !next_states || !terminals || !next_legal_actions)
    return HbFail("ReplaySampleTransitionBatch: null argument");
  const int64_t threads = (int64_t)count * 32;
  hbr::k_gather<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      MakeView(r), indices, count, states, actions, rewards, next_states, terminals, next_legal_actions);
  HBR_CUDA(cudaGetLastError());
  return 0;
}


// ---- self-play recorder -------------------------------------------------------------
struct Recorder {
  int device = 0;
  hbr::RecView v;
  int nblocks = 0;
  int sm_count = 148;
  unsigned long long* h_total = nullptr;  // pinned
};

int NewRecorder(pyhanabi_recorder_t* out, int num_games, int num_players, int observation_size,
                int num_actions, int max_turns_per_seat, int device) {
  if (!out) return HbFail("NewRecorder: null handle");
  out->recorder = nullptr;
  if (num_games <= 0 || num_players < 2 || num_players > 5 || observation_size <= 0 ||
      num_actions <= 0 || num_actions > 64 || max_turns_per_seat <= 0)
    return HbFail("NewRecorder: bad sizes (players 2..5, actions <= 64)");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return HbFail("NewRecorder: no such CUDA device (the recorder has no CPU fallback)");
  Recorder* r = new Recorder;
  r->device = device;
  Guard guard(device);
  cudaDeviceGetAttribute(&r->sm_count, cudaDevAttrMultiProcessorCount, device);
  hbr::RecView& v = r->v;
  std::memset(&v, 0, sizeof(v));
  v.N = num_games; v.P = num_players; v.L = max_turns_per_seat; v.A = num_actions; v.D = observation_size;
  v.W = 4 * ((observation_size + 127) / 128);
  const size_t NP = (size_t)v.N * v.P, NPL = NP * v.L;
  r->nblocks = (int)((NP + 1023) / 1024);
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
  };
  alloc((void**)&v.obs, NPL * v.W * sizeof(uint32_t));
  alloc((void**)&v.legal, NPL * sizeof(uint64_t));
  alloc((void**)&v.action, NPL * sizeof(int32_t));
  alloc((void**)&v.reward, NPL * sizeof(float));
  alloc((void**)&v.count, NP * sizeof(int32_t));
  alloc((void**)&v.since, NP * sizeof(float));
  alloc((void**)&v.run_len, NP * sizeof(int32_t));
  alloc((void**)&v.run_since, NP * sizeof(float));
  alloc((void**)&v.lens, NP * sizeof(int32_t));
  alloc((void**)&v.block_sums, (size_t)r->nblocks * sizeof(int32_t));
  v.owner_cap = (int64_t)(NPL < ((size_t)1 << 26) ? NPL : ((size_t)1 << 26));
  alloc((void**)&v.owner, (size_t)v.owner_cap * sizeof(int32_t));
  alloc((void**)&v.total, sizeof(unsigned long long));
  alloc((void**)&v.dropped, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&r->h_total, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  out->recorder = r;
  if (e != cudaSuccess) {
    DeleteRecorder(out);
    return Check(e, "NewRecorder");
  }
  return 0;
}

void DeleteRecorder(pyhanabi_recorder_t* h) {
  if (!h || !h->recorder) return;
  Recorder* r = static_cast<Recorder*>(h->recorder);
  Guard guard(r->device);
  hbr::RecView& v = r->v;
  cudaFree(v.obs); cudaFree(v.legal); cudaFree(v.action); cudaFree(v.reward); cudaFree(v.count);
  cudaFree(v.since); cudaFree(v.run_len); cudaFree(v.run_since); cudaFree(v.lens);
  cudaFree(v.block_sums); cudaFree(v.owner); cudaFree(v.total); cudaFree(v.dropped);
  if (r->h_total) cudaFreeHost(r->h_total);
  delete r;
  h->recorder = nullptr;
}

int RecorderPackedWords(pyhanabi_recorder_t* h) {
  return (h && h->recorder) ? static_cast<Recorder*>(h->recorder)->v.W : -1;
}

// The acting seat's observation (bit-packed rows as BatchStepEncodePacked /
// BatchObservePacked write them), legal mask and chosen action of every game, before the
// environment step.  Asynchronous on `stream`.
int RecorderBeforeStep(pyhanabi_recorder_t* h, const int32_t* cur_player, const uint32_t* obs_bits,
                       const uint8_t* mask, const int32_t* actions, void* stream) {
  if (!h || !h->recorder) return HbFail("null recorder handle");
  Recorder* r = static_cast<Recorder*>(h->recorder);
  if (!cur_player || !obs_bits || !mask || !actions) return HbFail("RecorderBeforeStep: null argument");
  Guard guard(r->device);
  hbr::k_rec_before<<<(r->v.N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(r->v, cur_player, obs_bits, mask, actions);
  return Check(cudaGetLastError(), "RecorderBeforeStep");
}

// reward / done of the environment step just taken.  The episodes of the finished games
// are appended to `replay` (each seat's episode one contiguous run, runs ordered by game
// then seat; new transitions enter the sum tree with `priority`, < 0 = the default 100).
// Synchronises once on `stream` to learn how many transitions were posted (*posted).
int RecorderAfterStep(pyhanabi_recorder_t* h, const int32_t* rewards, const uint8_t* dones,
                      pyhanabi_replay_t* replay, float priority, int64_t* posted, void* stream) {
  if (!h || !h->recorder) return HbFail("null recorder handle");
  Recorder* rec = static_cast<Recorder*>(h->recorder);
  if (!rewards || !dones) return HbFail("RecorderAfterStep: null argument");
  if (!replay || !replay->replay) return HbFail("RecorderAfterStep: null replay handle");
  Replay* r = static_cast<Replay*>(replay->replay);
  if (r->device != rec->device) return HbFail("RecorderAfterStep: recorder and replay memory sit on different devices");
  if (r->A != rec->v.A || r->D != rec->v.D) return HbFail("RecorderAfterStep: replay memory has other observation / action sizes");
  Guard guard(rec->device);
  cudaStream_t s = (cudaStream_t)stream;
  hbr::RecView& v = rec->v;
  hbr::k_rec_after<<<(v.N + 127) / 128, 128, 0, s>>>(v, rewards, dones);
  const int np = v.N * v.P;
  hbr::k_scan_blocks<<<rec->nblocks, 1024, 0, s>>>(v.lens, np, v.block_sums);
  hbr::k_scan_sums<<<1, 1024, 0, s>>>(v.block_sums, rec->nblocks, v.total);
  hbr::k_scan_add<<<rec->nblocks, 1024, 0, s>>>(v.lens, np, v.block_sums);
  HBR_CUDA(cudaGetLastError());
  // everything below reads the cursor and the count from device memory: fixed launch
  // geometry, no host round trip, capturable in a CUDA graph
  const int sms = rec->sm_count;
  hbr::k_rec_owner<<<(unsigned)((np + 255) / 256), 256, 0, s>>>(v);
  hbr::k_rec_post<<<sms * 16, 256, 0, s>>>(v, r->d_counts, r->cap, r->obs, r->actions, r->rewards,
                                           r->terminals, r->legal);
  if (r->prioritized) {
    const double p = priority < 0.f ? 100.0 : (double)priority;  // DEFAULT_PRIORITY
    if (p > r->max_recorded_priority) r->max_recorded_priority = p;
    hbr::k_set_leaves_dev<<<sms * 2, 256, 0, s>>>(r->tree, r->depth, r->cap, r->d_counts, v.total, p);
    // wide levels with strided grids, then one block walks the narrow ones to the root
    int l = r->depth - 1;
    for (; l >= 0 && (r->depth - l) <= 6; --l)
      hbr::k_rebuild_span_dev<<<sms, 256, 0, s>>>(r->tree, r->depth, l, r->cap, r->d_counts, v.total);
    if (l >= 0) hbr::k_rebuild_top_dev<<<1, 1024, 0, s>>>(r->tree, r->depth, l, r->cap, r->d_counts, v.total);
  }
  hbr::k_commit_count<<<1, 1, 0, s>>>(r->d_counts, v.total, r->cap);
  HBR_CUDA(cudaGetLastError());
  r->dirty = true;
  if (posted) {  // the caller wants the number now: one synchronisation
    HBR_CUDA(cudaMemcpyAsync(rec->h_total, v.total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    HBR_CUDA(cudaStreamSynchronize(s));
    *posted = (int64_t)rec->h_total[0];
  }
  return 0;
}

// Transitions that did not fit a seat's max_turns_per_seat slots (0 for the default 64).
int64_t RecorderDropped(pyhanabi_recorder_t* h, void* stream) {
  if (!h || !h->recorder) return -1;
  Recorder* r = static_cast<Recorder*>(h->recorder);
  Guard guard(r->device);
  if (cudaMemcpyAsync(r->h_total + 1, r->v.dropped, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                      (cudaStream_t)stream) != cudaSuccess ||
      cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
    return -1;
  return (int64_t)r->h_total[1];
}

}  // extern "C"
