// hb_policy.cu -- fused actor/learner-side kernels on the same batch:
//   BatchSelectAction          legal-mask-aware epsilon-greedy / argmax
//                              (agents/rainbow_copy/dqn_agent.py:387-420, :200;
//                               rainbow_agent.py:163-172 for the C51 q-values)
//   BatchProjectDistribution   Rainbow's C51 categorical projection
//                              (agents/rainbow_copy/rainbow_agent.py:252-404)
//   BatchC51TargetDistribution _build_target_distribution fused end to end
//                              (agents/rainbow_copy/rainbow_agent.py:181-213)
// fp32 throughout; no tensor cores (there is no dense contraction here).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/pyhanabi.h"
#include "hb_policy_rng.h"

namespace hbp {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// q-values given: one thread per game.
__global__ void k_select_q(const float* __restrict__ q, const uint8_t* __restrict__ mask, int n,
                           int A, float epsilon, uint64_t seed, uint64_t step,
                           int32_t* __restrict__ actions) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint8_t* m = mask + (int64_t)g * A;
  int n_legal = 0;
  for (int a = 0; a < A; ++a) n_legal += m[a] != 0;
  if (n_legal == 0) { actions[g] = -1; return; }
  const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
  const float u = (float)(bits >> 40) * (1.0f / 16777216.0f);  // 24-bit uniform in [0,1)
  int choice = -1;
  if (u <= epsilon && epsilon > 0.0f) {  // random.random() <= epsilon
    int k = (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
    for (int a = 0; a < A; ++a)
      if (m[a] != 0 && k-- == 0) { choice = a; break; }
  } else {
    // argmax_a(q[a] + legal[a]) with legal in {0, -inf}: first maximal index (tf.argmax).
    float best = -CUDART_INF_F;
    const float* row = q + (int64_t)g * A;
    for (int a = 0; a < A; ++a) {
      const float val = m[a] != 0 ? row[a] : -CUDART_INF_F;
      if (choice < 0 ? (m[a] != 0) : (val > best)) { best = val; choice = a; }
    }
  }
  actions[g] = choice;
}

// Position of the k-th (0-based) set bit of x; k < popcount(x).
__device__ __forceinline__ int kth_set_bit(uint64_t x, int k) {
  uint32_t w = (uint32_t)x;
  int pos = 0, c = __popc(w);
  if (k >= c) { k -= c; w = (uint32_t)(x >> 32); pos = 32; }
  c = __popc(w & 0xffffu);
  if (k >= c) { k -= c; w >>= 16; pos += 16; }
  c = __popc(w & 0xffu);
  if (k >= c) { k -= c; w >>= 8; pos += 8; }
  c = __popc(w & 0xfu);
  if (k >= c) { k -= c; w >>= 4; pos += 4; }
  c = __popc(w & 0x3u);
  if (k >= c) { k -= c; w >>= 2; pos += 2; }
  return pos + ((k >= (int)(w & 1u)) ? 1 : 0);
}

// Same selection for A <= 64, one warp per tile of 32 games: the tile's 32*A mask
// bytes are one contiguous range, fetched with coalesced 128-bit loads into shared
// memory; every lane then folds its row into a 64-bit legal set and picks the k-th
// member by popcount bisection instead of walking the row.
__global__ void k_select_q_tile(const float* __restrict__ q, const uint8_t* __restrict__ mask,
                                int n, int A, float epsilon, uint64_t seed, uint64_t step,
                                int32_t* __restrict__ actions) {
  extern __shared__ uint4 s_tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32;
  if (tile0 >= n) return;
  const int n_valid = (n - tile0) < 32 ? (int)(n - tile0) : 32;
  uint4* mine4 = s_tile + (size_t)warp * 2 * A;  // 32 * A bytes
  uint8_t* mine = reinterpret_cast<uint8_t*>(mine4);
  const uint8_t* src = mask + tile0 * A;
  const int total = n_valid * A;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int quads = total >> 4;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    for (int k = lane; k < quads; k += 32) mine4[k] = __ldcs(src4 + k);
    for (int b = (quads << 4) + lane; b < total; b += 32) mine[b] = src[b];
  } else {
    for (int b = lane; b < total; b += 32) mine[b] = src[b];
  }
  __syncwarp();
  if (lane >= n_valid) return;
  const int64_t g = tile0 + lane;
  const uint8_t* m = mine + lane * A;
  uint64_t legal = 0;
  for (int a = 0; a < A; ++a) legal |= (uint64_t)(m[a] != 0) << a;
  const int n_legal = __popcll(legal);
  if (n_legal == 0) { actions[g] = -1; return; }
  const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
  const float u = (float)(bits >> 40) * (1.0f / 16777216.0f);  // 24-bit uniform in [0,1)
  int choice = -1;
  if (u <= epsilon && epsilon > 0.0f) {  // random.random() <= epsilon
    const int k = (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
    choice = kth_set_bit(legal, k);
  } else {
    // argmax_a(q[a] + legal[a]) with legal in {0, -inf}: first maximal index (tf.argmax)
    float best = -CUDART_INF_F;
    const float* row = q + g * A;
    for (uint64_t rest = legal; rest; rest &= rest - 1) {
      const int a = __ffsll((long long)rest) - 1;
      const float val = row[a];
      if (choice < 0 || val > best) { best = val; choice = a; }
    }
  }
  actions[g] = choice;
}

// C51 logits given: one warp per game.  q[a] = sum_i support[i] * softmax(logits[a,:])[i].
__global__ void k_select_c51(const float* __restrict__ logits, const float* __restrict__ support,
                             const uint8_t* __restrict__ mask, int n, int A, int atoms,
                             float epsilon, uint64_t seed, uint64_t step,
                             int32_t* __restrict__ actions, float* __restrict__ q_out) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n) return;
  const uint8_t* m = mask + (int64_t)g * A;
  const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
  const float u = (float)(bits >> 40) * (1.0f / 16777216.0f);
  int n_legal = 0;
  for (int a = lane; a < A; a += 32) n_legal += m[a] != 0;
  n_legal = __reduce_add_sync(kFull, n_legal);
  if (n_legal == 0) { if (lane == 0) actions[g] = -1; return; }
  const bool explore = (u <= epsilon && epsilon > 0.0f);
  int choice = -1;
  if (explore && q_out == nullptr) {
    int k = (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
    for (int a = 0; a < A; ++a)
      if (m[a] != 0 && k-- == 0) { choice = a; break; }
  } else {
    float best = -CUDART_INF_F;
    for (int a = 0; a < A; ++a) {
      const bool legal = m[a] != 0;
      if (!legal && q_out == nullptr) continue;  // warp-uniform: skip the read entirely
      const float* row = logits + ((int64_t)g * A + a) * atoms;
      float mx = -CUDART_INF_F;
      for (int i = lane; i < atoms; i += 32) mx = fmaxf(mx, row[i]);
      mx = warp_max(mx);
      float se = 0.f, sw = 0.f;
      for (int i = lane; i < atoms; i += 32) {
        const float e = expf(row[i] - mx);
        se += e;
        sw += e * support[i];
      }
      se = warp_sum(se);
      sw = warp_sum(sw);
      const float qa = sw / se;
      if (q_out && lane == 0) q_out[(int64_t)g * A + a] = qa;
      if (legal && (choice < 0 || qa > best)) { best = qa; choice = a; }
    }
    if (explore) {
      int k = (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
      choice = -1;
      for (int a = 0; a < A; ++a)
        if (m[a] != 0 && k-- == 0) { choice = a; break; }
    }
  }
  if (lane == 0) actions[g] = choice;
}

// project_distribution, one warp per batch row, op for op in fp32:
//   clip(supports, vmin, vmax) -> |.-z_i| -> 1 - x/dz -> clip[0,1] -> * w -> sum_j (index order)
__device__ __forceinline__ void project_row(const float* s_sup, const float* s_w, const float* z,
                                            int N, float* out_row, int lane) {
  const float vmin = z[0], vmax = z[N - 1];
  const float dz = z[1] - z[0];
  for (int i = lane; i < N; i += 32) {
    const float zi = z[i];
    float acc = 0.f;
    for (int j = 0; j < N; ++j) {
      const float sj = fminf(fmaxf(s_sup[j], vmin), vmax);
      const float quotient = __fsub_rn(1.0f, __fdiv_rn(fabsf(__fsub_rn(sj, zi)), dz));
      const float cq = fminf(fmaxf(quotient, 0.0f), 1.0f);
      acc = __fadd_rn(acc, __fmul_rn(cq, s_w[j]));
    }
    out_row[i] = acc;
  }
}

__global__ void k_project(const float* __restrict__ supports, const float* __restrict__ weights,
                          const float* __restrict__ z, float* __restrict__ out, int B, int N) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float* s_sup = sm + (size_t)warp * 2 * N;
  float* s_w = s_sup + N;
  if (b >= B) return;
  for (int j = lane; j < N; j += 32) {
    s_sup[j] = supports[(int64_t)b * N + j];
    s_w[j] = weights[(int64_t)b * N + j];
  }
  __syncwarp();
  project_row(s_sup, s_w, z, N, out + (int64_t)b * N, lane);
}

// _build_target_distribution: softmax(next logits) -> q -> masked argmax ->
// gather that action's probabilities -> supports = r + gamma^n (1 - terminal) z
// -> projection.  One warp per batch row, nothing round-trips through HBM.
__global__ void k_c51_target(const float* __restrict__ rewards, const uint8_t* __restrict__ terminals,
                             const float* __restrict__ next_logits,
                             const uint8_t* __restrict__ next_mask, const float* __restrict__ z,
                             float cumulative_gamma, float* __restrict__ out,
                             int32_t* __restrict__ argmax_out, int B, int A, int N) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float* s_sup = sm + (size_t)warp * 2 * N;
  float* s_w = s_sup + N;
  if (b >= B) return;
  const uint8_t* m = next_mask + (int64_t)b * A;
  int choice = -1;
  float best = -CUDART_INF_F;
  for (int a = 0; a < A; ++a) {
    if (m[a] == 0) continue;
    const float* row = next_logits + ((int64_t)b * A + a) * N;
    float mx = -CUDART_INF_F;
    for (int i = lane; i < N; i += 32) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    float se = 0.f, sw = 0.f;
    for (int i = lane; i < N; i += 32) {
      const float e = expf(row[i] - mx);
      se += e;
      sw += e * z[i];
    }
    se = warp_sum(se);
    sw = warp_sum(sw);
    const float qa = sw / se;
    if (choice < 0 || qa > best) { best = qa; choice = a; }
  }
  if (choice < 0) choice = 0;  // no legal next action (terminal rows): any row, gamma term is 0
  if (argmax_out && lane == 0) argmax_out[b] = choice;
  {
    const float* row = next_logits + ((int64_t)b * A + choice) * N;
    float mx = -CUDART_INF_F;
    for (int i = lane; i < N; i += 32) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int i = lane; i < N; i += 32) se += expf(row[i] - mx);
    se = warp_sum(se);
    const float g = cumulative_gamma * (1.0f - (terminals[b] ? 1.0f : 0.0f));
    const float r = rewards[b];
    for (int i = lane; i < N; i += 32) {
      s_w[i] = expf(row[i] - mx) / se;
      s_sup[i] = r + g * z[i];
    }
  }
  __syncwarp();
  project_row(s_sup, s_w, z, N, out + (int64_t)b * N, lane);
}

}  // namespace hbp

int HbFail(const std::string& msg);  // hb_batch.cu
namespace {
int PFail(const std::string& s) { return HbFail(s); }
}  // namespace

extern "C" {

int BatchSelectAction(const float* q_or_logits, int num_atoms, const float* support,
                      const uint8_t* mask, int n, int num_actions, float epsilon, uint64_t seed,
                      uint64_t step, int32_t* actions, void* stream) {
  if (!mask || !actions) return PFail("BatchSelectAction: mask/actions is null");
  if (n <= 0 || num_actions <= 0) return PFail("BatchSelectAction: bad sizes");
  if (!q_or_logits && epsilon < 1.0f) return PFail("BatchSelectAction: values are required unless epsilon >= 1");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_atoms <= 0 || !q_or_logits) {
    const float eps = q_or_logits ? epsilon : 2.0f;
    if (num_actions <= 64) {
      const int warps = 4;
      const int64_t tiles = ((int64_t)n + 31) / 32;
      hbp::k_select_q_tile<<<(unsigned)((tiles + warps - 1) / warps), warps * 32,
                             (size_t)warps * 32 * num_actions, s>>>(q_or_logits, mask, n, num_actions,
                                                                    eps, seed, step, actions);
    } else {
      hbp::k_select_q<<<(n + 255) / 256, 256, 0, s>>>(q_or_logits, mask, n, num_actions, eps, seed,
                                                     step, actions);
    }
  } else {
    if (!support) return PFail("BatchSelectAction: support is null");
    const int64_t threads = (int64_t)n * 32;
    hbp::k_select_c51<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
        q_or_logits, support, mask, n, num_actions, num_atoms, epsilon, seed, step, actions, nullptr);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchSelectAction: ") + cudaGetErrorString(e));
}

// Host-side twin of BatchSelectAction for epsilon >= 1 (uniform random legal
// move; the reference's agents/random_agent.py): same counter-based generator,
// so for equal (seed, step, game) it returns exactly the device kernel's actions.
// mask and actions are HOST arrays; `threads` <= 0 uses the hardware concurrency.
int HostRandomLegalActions(const uint8_t* mask, int n, int num_actions, uint64_t seed,
                           uint64_t step, int32_t* actions, int threads) {
  if (!mask || !actions || n < 0 || num_actions <= 0 || num_actions > 64) return PFail("HostRandomLegalActions: bad argument");
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > 64) threads = 64;
  if ((int64_t)threads * 4096 > n) threads = n / 4096 > 0 ? n / 4096 : 1;
  auto work = [=](int t) {
    const int lo = (int)((int64_t)n * t / threads), hi = (int)((int64_t)n * (t + 1) / threads);
    for (int g = lo; g < hi; ++g) {
      const uint8_t* m = mask + (int64_t)g * num_actions;
      uint64_t legal = 0;
      for (int a = 0; a < num_actions; ++a) legal |= (uint64_t)(m[a] != 0) << a;
      const int n_legal = __builtin_popcountll(legal);
      int choice = -1;
      if (n_legal > 0) {
        const uint64_t bits = hbp::policy_bits(seed, step, (uint64_t)g);
        int k = (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
        while (k-- > 0) legal &= legal - 1;
        choice = __builtin_ctzll(legal);
      }
      actions[g] = choice;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return 0;
}

// C51 q-values q[n, A] = sum_i support_i softmax(logits[n, a, :])_i (all actions).
int BatchC51QValues(const float* logits, const float* support, const uint8_t* mask, int n,
                    int num_actions, int num_atoms, float* q_out, int32_t* greedy_actions,
                    void* stream) {
  if (!logits || !support || !mask || !q_out || !greedy_actions) return PFail("BatchC51QValues: null argument");
  const int64_t threads = (int64_t)n * 32;
  hbp::k_select_c51<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      logits, support, mask, n, num_actions, num_atoms, 0.0f, 0, 0, greedy_actions, q_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchC51QValues: ") + cudaGetErrorString(e));
}

int BatchProjectDistribution(const float* supports, const float* weights,
                             const float* target_support, float* out, int batch, int num_dims,
                             void* stream) {
  if (!supports || !weights || !target_support || !out) return PFail("BatchProjectDistribution: null argument");
  if (batch <= 0 || num_dims < 2) return PFail("BatchProjectDistribution: need batch > 0 and num_dims >= 2");
  const int warps = 4;
  const size_t smem = (size_t)warps * 2 * num_dims * sizeof(float);
  if (smem > 48 * 1024) return PFail("BatchProjectDistribution: num_dims too large");
  hbp::k_project<<<(batch + warps - 1) / warps, warps * 32, smem, (cudaStream_t)stream>>>(
      supports, weights, target_support, out, batch, num_dims);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchProjectDistribution: ") + cudaGetErrorString(e));
}

int BatchC51TargetDistribution(const float* rewards, const uint8_t* terminals,
                               const float* next_logits, const uint8_t* next_legal_mask,
                               const float* support, float cumulative_gamma, float* out,
                               int32_t* next_argmax, int batch, int num_actions, int num_atoms,
                               void* stream) {
  if (!rewards || !terminals || !next_logits || !next_legal_mask || !support || !out)
    return PFail("BatchC51TargetDistribution: null argument");
  if (batch <= 0 || num_atoms < 2) return PFail("BatchC51TargetDistribution: bad sizes");
  const int warps = 4;
  const size_t smem = (size_t)warps * 2 * num_atoms * sizeof(float);
  if (smem > 48 * 1024) return PFail("BatchC51TargetDistribution: num_atoms too large");
  hbp::k_c51_target<<<(batch + warps - 1) / warps, warps * 32, smem, (cudaStream_t)stream>>>(
      rewards, terminals, next_logits, next_legal_mask, support, cumulative_gamma, out, next_argmax,
      batch, num_actions, num_atoms);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchC51TargetDistribution: ") + cudaGetErrorString(e));
}

}  // extern "C"
