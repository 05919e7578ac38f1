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
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdint.h>

#include <climits>
#include <cstdlib>
#include <string>
#include <type_traits>
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

// Logit element types of the C51 head (BatchSelectActionC51 `dtype`).
template <class T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// C51 logits given: one warp per game.  q[a] = sum_i support[i] * softmax(logits[a,:])[i].
// Game g's logits start at logits + g * row_pitch (elements), action a at + a * atoms.
template <class T>
__global__ void k_select_c51(const T* __restrict__ logits, int64_t row_pitch,
                             const float* __restrict__ support,
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
    int k = random_legal_rank(bits, n_legal);
    for (int a = 0; a < A; ++a)
      if (m[a] != 0 && k-- == 0) { choice = a; break; }
  } else {
    float best = -CUDART_INF_F;
    const T* base = logits + (int64_t)g * row_pitch;
    const bool two = atoms <= 64;  // the usual 51 atoms: every logit is read exactly once
    const float s0 = lane < atoms ? support[lane] : 0.f;
    const float s1 = two && lane + 32 < atoms ? support[lane + 32] : 0.f;
    for (int a = 0; a < A; ++a) {
      const bool legal = m[a] != 0;
      if (!legal && q_out == nullptr) continue;  // warp-uniform: skip the read entirely
      const T* row = base + (int64_t)a * atoms;
      float se = 0.f, sw = 0.f;
      if (two) {
        const float v0 = lane < atoms ? to_f32<T>(row[lane]) : -CUDART_INF_F;
        const float v1 = lane + 32 < atoms ? to_f32<T>(row[lane + 32]) : -CUDART_INF_F;
        const float mx = warp_max(fmaxf(v0, v1));
        const float e0 = lane < atoms ? expf(v0 - mx) : 0.f;
        const float e1 = lane + 32 < atoms ? expf(v1 - mx) : 0.f;
        se = e0 + e1;
        sw = e0 * s0 + e1 * s1;
      } else {
        float mx = -CUDART_INF_F;
        for (int i = lane; i < atoms; i += 32) mx = fmaxf(mx, to_f32<T>(row[i]));
        mx = warp_max(mx);
        for (int i = lane; i < atoms; i += 32) {
          const float e = expf(to_f32<T>(row[i]) - mx);
          se += e;
          sw += e * support[i];
        }
      }
      se = warp_sum(se);
      sw = warp_sum(sw);
      const float qa = sw / se;
      if (q_out && lane == 0) q_out[(int64_t)g * A + a] = qa;
      if (legal && (choice < 0 || qa > best)) { best = qa; choice = a; }
    }
    if (explore) {
      int k = random_legal_rank(bits, n_legal);
      choice = -1;
      for (int a = 0; a < A; ++a)
        if (m[a] != 0 && k-- == 0) { choice = a; break; }
    }
  }
  if (lane == 0) actions[g] = choice;
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Two consecutive 16-bit logits from one 32-bit shared-memory word.
template <class T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// sum_i e_i and sum_i e_i * support_i over one action's atoms, e_i = 2^(v_i * log2e + nm),
// from the staged row.  s_sup[0..63] = support, s_sup[64..127] = support shifted by one.
// KA > 0: the number of atoms is the compile-time constant KA (the usual 51), so the loops
// unroll completely; KA == 0: taken from `atoms`.
template <class T, int KA>
__device__ __forceinline__ void action_sums(const T* row, int start, int atoms, float nm,
                                            const float* s_sup, float& se, float& sw) {
  if (KA > 0) atoms = KA;
  constexpr float kL2e = 1.4426950408889634f;
  se = 0.f; sw = 0.f;
  if (sizeof(T) == 4) {
    const float* r = reinterpret_cast<const float*>(row) + start;
#pragma unroll 4
    for (int i = 0; i < atoms; ++i) {
      const float e = ex2_fast(fmaf(r[i], kL2e, nm));
      se += e;
      sw = fmaf(e, s_sup[i], sw);
    }
  } else {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + (start >> 1);
    const int odd = start & 1;
    if (odd) {  // the action starts in the high half of its first word
      const float e = ex2_fast(fmaf(unpack2<typename std::conditional<sizeof(T) == 4, __half, T>::type>(w[0]).y, kL2e, nm));
      se = e;
      sw = e * s_sup[0];
    }
    w += odd;
    const float2* sp = reinterpret_cast<const float2*>(s_sup + (odd ? 64 : 0));
    // an odd atom count gives the same number of whole pairs for either alignment
    const int pairs = (KA > 0 && (KA & 1)) ? (KA >> 1) : ((atoms - odd) >> 1);
#pragma unroll
    for (int j = 0; j < pairs; ++j) {
      const float2 v = unpack2<typename std::conditional<sizeof(T) == 4, __half, T>::type>(w[j]);
      const float2 sv = sp[j];
      const float e0 = ex2_fast(fmaf(v.x, kL2e, nm)), e1 = ex2_fast(fmaf(v.y, kL2e, nm));
      se += e0 + e1;
      sw = fmaf(e0, sv.x, sw);
      sw = fmaf(e1, sv.y, sw);
    }
    if ((atoms - odd) & 1) {  // one more element in the low half of the next word
      const float e = ex2_fast(fmaf(unpack2<typename std::conditional<sizeof(T) == 4, __half, T>::type>(w[pairs]).x, kL2e, nm));
      se += e;
      sw = fmaf(e, s_sup[atoms - 1], sw);  // `atoms` = this call's length
    }
  }
}

// The same selection with the game's whole logits row staged in shared memory: the
// warp fetches the row with coalesced 128-bit loads (all in flight at once), then lane
// a owns action a -- softmax . support over its atoms without any shuffle -- and a
// final warp reduction takes the first maximal legal action.  exp(v - M) uses the row
// maximum M (>= every action's own maximum: no overflow); an action whose terms all
// underflow against M is redone against its own maximum.
// Needs 16-byte aligned rows (base and pitch) and atoms <= 63;
// smem: 128 floats + warps * row_smem_bytes.
// action_sums for a slice of an action's atoms: `len` atoms starting at row element `start`,
// the first of which has support index `sup_idx` (row alignment and support alignment are
// independent here: the pair loads pick the copy of the support whose pairs are 8-byte
// aligned for this slice).
template <class T, int KL>
__device__ __forceinline__ void slice_sums(const T* row, int start, int len, float nm,
                                           const float* s_sup, int sup_idx, float& se, float& sw) {
  if (KL > 0) len = KL;
  constexpr float kL2e = 1.4426950408889634f;
  se = 0.f; sw = 0.f;
  if (sizeof(T) == 4) {
    const float* r = reinterpret_cast<const float*>(row) + start;
#pragma unroll
    for (int i = 0; i < len; ++i) {
      const float e = ex2_fast(fmaf(r[i], kL2e, nm));
      se += e;
      sw = fmaf(e, s_sup[sup_idx + i], sw);
    }
  } else {
    typedef typename std::conditional<sizeof(T) == 4, __half, T>::type T16;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + (start >> 1);
    const int odd = start & 1;
    if (odd) {
      const float e = ex2_fast(fmaf(unpack2<T16>(w[0]).y, kL2e, nm));
      se = e;
      sw = e * s_sup[sup_idx];
    }
    w += odd;
    const int p0 = sup_idx + odd;  // support index of the first pair
    const float2* sp = reinterpret_cast<const float2*>(s_sup + ((p0 & 1) ? 64 + p0 - 1 : p0));
    const int pairs = (KL > 0 && (KL & 1)) ? (KL >> 1) : ((len - odd) >> 1);
#pragma unroll
    for (int j = 0; j < pairs; ++j) {
      const float2 v = unpack2<T16>(w[j]);
      const float2 sv = sp[j];
      const float e0 = ex2_fast(fmaf(v.x, kL2e, nm)), e1 = ex2_fast(fmaf(v.y, kL2e, nm));
      se += e0 + e1;
      sw = fmaf(e0, sv.x, sw);
      sw = fmaf(e1, sv.y, sw);
    }
    if ((len - odd) & 1) {
      const float e = ex2_fast(fmaf(unpack2<T16>(w[pairs]).x, kL2e, nm));
      se += e;
      sw = fmaf(e, s_sup[sup_idx + len - 1], sw);
    }
  }
}

// k_select_c51_rows with every action cut into three slices of atoms, so that 3 * A work
// units (60 for Hanabi-Full 2p) keep all 32 lanes busy for two short passes instead of A
// lanes for one long one; the partial sums meet in shared memory.  KC = slice length when
// it is a compile-time constant (17 for 51 atoms), else 0.
// smem: 128 floats + warps * row_smem_bytes + warps * 3 * A float2.
template <class T, int KC>
__global__ void k_select_c51_thirds(const T* __restrict__ logits, int64_t row_pitch,
                                    const float* __restrict__ support,
                                    const uint8_t* __restrict__ mask, int n, int A, int atoms,
                                    float epsilon, uint64_t seed, uint64_t step,
                                    int32_t* __restrict__ actions, float* __restrict__ q_out,
                                    int row_smem_bytes) {
  extern __shared__ uint4 s_rows[];
  float* s_sup = reinterpret_cast<float*>(s_rows);  // [0,64): support, [64,128): shifted by one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  if (threadIdx.x < 128) {
    const int i = threadIdx.x & 63, sh = threadIdx.x >> 6;
    s_sup[threadIdx.x] = i + sh < atoms ? support[i + sh] : 0.f;
  }
  __syncthreads();
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n) return;
  uint8_t* base = reinterpret_cast<uint8_t*>(s_rows) + 512;
  T* row = reinterpret_cast<T*>(base + (size_t)warp * row_smem_bytes);
  float2* part = reinterpret_cast<float2*>(base + (size_t)warps * row_smem_bytes) + (size_t)warp * 3 * A;
  const T* src = logits + (int64_t)g * row_pitch;
  const int len_row = A * atoms;
  constexpr int kPer = 16 / (int)sizeof(T);
  const int full = len_row / kPer;
  float mx = -CUDART_INF_F;
  for (int k = lane; k < full; k += 32) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + k);
    reinterpret_cast<uint4*>(row)[k] = v;
    if (sizeof(T) == 4) {
      mx = fmaxf(fmaxf(mx, __uint_as_float(v.x)), fmaxf(__uint_as_float(v.y), fmaxf(__uint_as_float(v.z), __uint_as_float(v.w))));
    } else {
      typedef typename std::conditional<sizeof(T) == 4, __half, T>::type T16;
      const float2 a = unpack2<T16>(v.x), b = unpack2<T16>(v.y), c = unpack2<T16>(v.z), d = unpack2<T16>(v.w);
      mx = fmaxf(fmaxf(fmaxf(mx, a.x), fmaxf(a.y, b.x)), fmaxf(fmaxf(b.y, c.x), fmaxf(c.y, fmaxf(d.x, d.y))));
    }
  }
  for (int i = full * kPer + lane; i < len_row; i += 32) {
    const T v = src[i];
    row[i] = v;
    mx = fmaxf(mx, to_f32<T>(v));
  }
  mx = warp_max(mx);
  __syncwarp();
  const uint8_t* m = mask + (int64_t)g * A;
  const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
  const float u01 = (float)(bits >> 40) * (1.0f / 16777216.0f);
  const bool explore = (u01 <= epsilon && epsilon > 0.0f);
  const float nm = -mx * 1.4426950408889634f;
  const int chunk = KC > 0 ? KC : (atoms + 2) / 3;
  const int units = 3 * A;
  for (int u0 = 0; u0 < units; u0 += 32) {
    const int u = u0 + lane;
    if (u < units) {
      const int a = u / 3, c = u - 3 * a;
      float se = 0.f, sw = 0.f;
      if (q_out != nullptr || m[a] != 0) {
        const int off = c * chunk;
        const int len = atoms - off < chunk ? atoms - off : chunk;
        if (KC > 0 && len == KC) slice_sums<T, KC>(row, a * atoms + off, len, nm, s_sup, off, se, sw);
        else if (len > 0) slice_sums<T, 0>(row, a * atoms + off, len, nm, s_sup, off, se, sw);
      }
      part[u] = make_float2(se, sw);
    }
  }
  __syncwarp();
  float best = -CUDART_INF_F;
  int choice = -1, n_legal = 0;
  for (int a = lane; a < A; a += 32) {
    const bool legal = m[a] != 0;
    n_legal += legal;
    if (!legal && q_out == nullptr) continue;
    const float2 p0 = part[3 * a], p1 = part[3 * a + 1], p2 = part[3 * a + 2];
    float se = p0.x + p1.x + p2.x, sw = p0.y + p1.y + p2.y;
    if (!(se > 0.f)) {  // everything underflowed against the row maximum
      float am = -CUDART_INF_F;
      for (int i = 0; i < atoms; ++i) am = fmaxf(am, to_f32<T>(row[a * atoms + i]));
      slice_sums<T, 0>(row, a * atoms, atoms, -am * 1.4426950408889634f, s_sup, 0, se, sw);
    }
    const float qa = sw / se;
    if (q_out) q_out[(int64_t)g * A + a] = qa;
    if (legal && (choice < 0 || qa > best)) { best = qa; choice = a; }
  }
  n_legal = __reduce_add_sync(kFull, n_legal);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oq = __shfl_xor_sync(kFull, best, o);
    const int oa = __shfl_xor_sync(kFull, choice, o);
    if (oa >= 0 && (choice < 0 || oq > best || (oq == best && oa < choice))) { best = oq; choice = oa; }
  }
  if (n_legal == 0) choice = -1;
  else if (explore) {
    uint64_t legal_set = 0;
    for (int a0 = 0; a0 < A; a0 += 32) {
      const int a = a0 + lane;
      legal_set |= (uint64_t)__ballot_sync(kFull, a < A && m[a] != 0) << a0;
    }
    choice = kth_set_bit(legal_set, random_legal_rank(bits, n_legal));
  }
  if (lane == 0) actions[g] = choice;
}

template <class T, int KA>
__global__ void k_select_c51_rows(const T* __restrict__ logits, int64_t row_pitch,
                                  const float* __restrict__ support,
                                  const uint8_t* __restrict__ mask, int n, int A, int atoms,
                                  float epsilon, uint64_t seed, uint64_t step,
                                  int32_t* __restrict__ actions, float* __restrict__ q_out,
                                  int row_smem_bytes) {
  extern __shared__ uint4 s_rows[];
  float* s_sup = reinterpret_cast<float*>(s_rows);  // [0,64): support, [64,128): shifted by one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 128) {
    const int i = threadIdx.x & 63, sh = threadIdx.x >> 6;
    s_sup[threadIdx.x] = i + sh < atoms ? support[i + sh] : 0.f;
  }
  __syncthreads();
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= n) return;
  T* row = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(s_rows) + 512 + (size_t)warp * row_smem_bytes);
  const T* src = logits + (int64_t)g * row_pitch;
  const int len = A * atoms;
  constexpr int kPer = 16 / (int)sizeof(T);
  const int full = len / kPer;
  float mx = -CUDART_INF_F;
  for (int k = lane; k < full; k += 32) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + k);
    reinterpret_cast<uint4*>(row)[k] = v;
    if (sizeof(T) == 4) {
      mx = fmaxf(fmaxf(mx, __uint_as_float(v.x)), fmaxf(__uint_as_float(v.y), fmaxf(__uint_as_float(v.z), __uint_as_float(v.w))));
    } else {
      typedef typename std::conditional<sizeof(T) == 4, __half, T>::type T16;
      const float2 a = unpack2<T16>(v.x), b = unpack2<T16>(v.y), c = unpack2<T16>(v.z), d = unpack2<T16>(v.w);
      mx = fmaxf(fmaxf(fmaxf(mx, a.x), fmaxf(a.y, b.x)), fmaxf(fmaxf(b.y, c.x), fmaxf(c.y, fmaxf(d.x, d.y))));
    }
  }
  for (int i = full * kPer + lane; i < len; i += 32) {
    const T v = src[i];
    row[i] = v;
    mx = fmaxf(mx, to_f32<T>(v));
  }
  mx = warp_max(mx);
  __syncwarp();
  const uint8_t* m = mask + (int64_t)g * A;
  const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
  const float u = (float)(bits >> 40) * (1.0f / 16777216.0f);
  const bool explore = (u <= epsilon && epsilon > 0.0f);
  const float nm = -mx * 1.4426950408889634f;
  float best = -CUDART_INF_F;
  int choice = -1, n_legal = 0;
  // (measured alternative: compacting the legal actions and cutting each into three
  // atom chunks over 30 lanes is slower, 287 vs 257 us at 2^18 games -- the second,
  // sparsely filled round costs more than the idle lanes here)
  for (int a = lane; a < A; a += 32) {
    const bool legal = m[a] != 0;
    n_legal += legal;
    if (!legal && q_out == nullptr) continue;
    float se, sw;
    action_sums<T, KA>(row, a * atoms, atoms, nm, s_sup, se, sw);
    if (!(se > 0.f)) {  // everything underflowed against the row maximum
      float am = -CUDART_INF_F;
      for (int i = 0; i < atoms; ++i) am = fmaxf(am, to_f32<T>(row[a * atoms + i]));
      action_sums<T, 0>(row, a * atoms, atoms, -am * 1.4426950408889634f, s_sup, se, sw);
    }
    const float qa = sw / se;
    if (q_out) q_out[(int64_t)g * A + a] = qa;
    if (legal && (choice < 0 || qa > best)) { best = qa; choice = a; }
  }
  n_legal = __reduce_add_sync(kFull, n_legal);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oq = __shfl_xor_sync(kFull, best, o);
    const int oa = __shfl_xor_sync(kFull, choice, o);
    if (oa >= 0 && (choice < 0 || oq > best || (oq == best && oa < choice))) { best = oq; choice = oa; }
  }
  if (n_legal == 0) choice = -1;
  else if (explore) {
    // uniform legal move: the k-th legal action in ascending order
    uint64_t legal_set = 0;
    for (int a0 = 0; a0 < A; a0 += 32) {
      const int a = a0 + lane;
      legal_set |= (uint64_t)__ballot_sync(kFull, a < A && m[a] != 0) << a0;
    }
    choice = kth_set_bit(legal_set, random_legal_rank(bits, n_legal));
  }
  if (lane == 0) actions[g] = choice;
}

// The PPO / REINFORCE actors' policy head (training/tf_agents_lib/masked_networks.py:26-50,
// 108-132): masked_logits = logits + mask with mask in {0, float32.min}, then a categorical
// draw.  One thread per game: softmax over the legal actions, inverse-CDF draw with the
// counter-based generator keyed (seed, step, game), log-probability of the drawn action
// (what the PPO loss needs from the collect policy).  greedy != 0 takes the mode instead
// (the evaluation policy).  Illegal actions have probability exactly 0, as adding
// float32.min to a logit makes its softmax term underflow to 0.
template <class T>
__global__ void k_sample_masked(const T* __restrict__ logits, int64_t row_pitch,
                                const uint8_t* __restrict__ mask, int n, int A, int greedy,
                                uint64_t seed, uint64_t step, int32_t* __restrict__ actions,
                                float* __restrict__ log_probs) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const T* row = logits + (int64_t)g * row_pitch;
  const uint8_t* m = mask + (int64_t)g * A;
  float mx = -CUDART_INF_F;
  int first = -1, best = -1;
  for (int a = 0; a < A; ++a)
    if (m[a] != 0) {
      const float v = to_f32<T>(row[a]);
      if (first < 0) first = a;
      if (best < 0 || v > mx) { mx = v; best = a; }
    }
  if (first < 0) {
    actions[g] = -1;
    if (log_probs) log_probs[g] = 0.f;
    return;
  }
  float sum = 0.f;
  for (int a = 0; a < A; ++a)
    if (m[a] != 0) sum += expf(to_f32<T>(row[a]) - mx);
  int choice = best;
  if (!greedy) {
    const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
    const float u = (float)(bits >> 40) * (1.0f / 16777216.0f) * sum;  // in [0, sum)
    float acc = 0.f;
    choice = -1;
    int last = first;
    for (int a = 0; a < A; ++a)
      if (m[a] != 0) {
        acc += expf(to_f32<T>(row[a]) - mx);
        last = a;
        if (u < acc) { choice = a; break; }
      }
    if (choice < 0) choice = last;  // rounding at the upper end
  }
  actions[g] = choice;
  if (log_probs) log_probs[g] = to_f32<T>(row[choice]) - mx - logf(sum);
}

// Bit-packed observation rows -> rows of 0/1 in a GEMM input type, K padded with
// zeros up to row_pitch.  One thread expands 8 bits into 8 elements.
template <class T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void store(float* p, uint32_t b) {
    float4 lo, hi;
    lo.x = (float)(b & 1u); lo.y = (float)((b >> 1) & 1u); lo.z = (float)((b >> 2) & 1u); lo.w = (float)((b >> 3) & 1u);
    hi.x = (float)((b >> 4) & 1u); hi.y = (float)((b >> 5) & 1u); hi.z = (float)((b >> 6) & 1u); hi.w = (float)((b >> 7) & 1u);
    reinterpret_cast<float4*>(p)[0] = lo;
    reinterpret_cast<float4*>(p)[1] = hi;
  }
};
template <uint32_t ONE> __device__ __forceinline__ void store8x16(void* p, uint32_t b) {
  // two 16-bit elements per word: bit i -> ONE in the low half, bit i+1 -> ONE in the high half
  uint4 v;
  v.x = ((b & 1u) * ONE) | (((b >> 1) & 1u) * (ONE << 16));
  v.y = (((b >> 2) & 1u) * ONE) | (((b >> 3) & 1u) * (ONE << 16));
  v.z = (((b >> 4) & 1u) * ONE) | (((b >> 5) & 1u) * (ONE << 16));
  v.w = (((b >> 6) & 1u) * ONE) | (((b >> 7) & 1u) * (ONE << 16));
  *reinterpret_cast<uint4*>(p) = v;
}
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void store(__nv_bfloat16* p, uint32_t b) { store8x16<0x3f80u>(p, b); }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void store(__half* p, uint32_t b) { store8x16<0x3c00u>(p, b); }
};

template <class T>
__global__ void k_expand_bits(const uint32_t* __restrict__ bits, int packed_words, int obs_size,
                              int64_t n, T* __restrict__ out, int64_t row_pitch) {
  // one thread per 32 bits of a row: one word load, four (16-bit types) or eight
  // (float32) independent 128-bit stores; a ragged last group of 8 when the pitch is
  // not a multiple of 32
  const int words = (int)((row_pitch + 31) >> 5);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * words) return;
  int64_t g;
  if (n * words < 0x7fffffffll) g = (int64_t)((uint32_t)idx / (uint32_t)words);  // 32-bit divide
  else g = idx / words;
  const int j = (int)(idx - g * words);
  uint32_t w = j < packed_words ? __ldcs(bits + g * packed_words + j) : 0u;
  const int rem = obs_size - 32 * j;
  if (rem < 32) w = rem > 0 ? (w & ((1u << rem) - 1u)) : 0u;
  T* dst = out + g * row_pitch + 32 * j;
  const int groups = (int)((row_pitch - 32 * j) >> 3) < 4 ? (int)((row_pitch - 32 * j) >> 3) : 4;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q < groups) Vec8<T>::store(dst + 8 * q, (w >> (8 * q)) & 0xffu);
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

// C51 head with 51 atoms and 16-bit logits, the shape the Rainbow actor runs
// (agents/rainbow_copy/rainbow_agent.py:36-68,163-172): a streaming kernel, HBM-bound by design
// (one 2 KB row in, 4 bytes out per game).
//   * persistent warps, one game per warp and iteration; rows travel global -> shared memory
//     with cp.async (16 bytes per lane and copy, four copies per row), kStages rows in flight
//     per warp, no registers and no store instructions spent on staging;
//   * an action's 51 atoms are three UNITS of 17; a lane owns two consecutive units =
//     34 elements = 17 consecutive 32-bit words of the row, which it picks out of the staged
//     row with 17 conflict-free LDS.32 (17 is odd); unit boundaries are therefore compile-time
//     positions inside the lane's registers and everything below unrolls;
//   * per unit: maximum, sum of 2^((x - max) * log2e), and the same sum weighted with the
//     support (the lane's 34 support values stay in registers across games);
//   * the three units of an action meet through shared memory (rescaled to their common
//     maximum, as in any split softmax), q = weighted / plain sum, then the masked arg-max
//     with the first maximal index (tf.argmax) by one warp reduction on an order-preserving
//     integer image of q; epsilon-greedy as in k_select_c51.
// Needs 3 * A <= 64, rows 16-byte aligned.  T = __nv_bfloat16 or __half.
#ifndef HB_U17_WARPS
#define HB_U17_WARPS 8
#endif
#ifndef HB_U17_STAGES
#define HB_U17_STAGES 4
#endif
// Element-wise maximum of two packed 16-bit pairs.
template <class T> __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <> __device__ __forceinline__ uint32_t max2<__half>(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}
constexpr int kU17RowQuads = 128 + 8;  // 2 KB row + pad (the last lane reads words 527..543)
constexpr size_t kU17SmemBytes = (size_t)HB_U17_WARPS * HB_U17_STAGES * kU17RowQuads * sizeof(uint4) +
                                 (size_t)HB_U17_WARPS * 64 * 3 * sizeof(float);

template <class T>
__global__ void __launch_bounds__(32 * HB_U17_WARPS) k_select_c51_u17(const T* __restrict__ logits, int64_t row_pitch,
                                                                     const float* __restrict__ support,
                                                                     const uint8_t* __restrict__ mask, int n, int A,
                                                                     float epsilon, uint64_t seed, uint64_t step,
                                                                     int32_t* __restrict__ actions) {
  constexpr int kAtoms = 51, kUnit = 17, kWarps = HB_U17_WARPS, kStages = HB_U17_STAGES;
  constexpr float kL2e = 1.4426950408889634f;
  extern __shared__ uint4 s_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* rows = s_dyn + (size_t)warp * kStages * kU17RowQuads;          // this warp's stages
  float* part = reinterpret_cast<float*>(s_dyn + (size_t)kWarps * kStages * kU17RowQuads) + warp * 64 * 3;
  const int units = 3 * A, len_row = A * kAtoms;
  const int quads = (len_row + 7) >> 3;        // 16-byte groups of a row that hold logits
  // the lane's two units and their support values (loop invariant)
  const int u0 = 2 * lane, u1 = 2 * lane + 1;
  float z0[kUnit], z1[kUnit];
#pragma unroll
  for (int j = 0; j < kUnit; ++j) {
    z0[j] = support[(u0 % 3) * kUnit + j];
    z1[j] = support[(u1 % 3) * kUnit + j];
  }
  // the pad words are read (by lanes whose units do not exist) but never used: keep them finite
  for (int st = 0; st < kStages; ++st)
    if (lane < 8) rows[st * kU17RowQuads + 128 + lane] = make_uint4(0, 0, 0, 0);
  const int wstride = gridDim.x * kWarps;
  const int g0 = blockIdx.x * kWarps + warp;
  auto fetch = [&](int g, int st) {
    if (g < n) {
      const uint4* src = reinterpret_cast<const uint4*>(logits + (int64_t)g * row_pitch);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < quads) cp_async16(rows + st * kU17RowQuads + lane + 32 * k, src + lane + 32 * k);
    }
    cp_async_commit();  // (an empty group keeps the group count in step)
  };
#pragma unroll
  for (int st = 0; st < kStages - 1; ++st) fetch(g0 + st * wstride, st);
  int stage = 0;
  // the legal-move byte of the lane's action, fetched one game ahead like the rows
  uint32_t mask_next = (g0 < n && lane < A) ? mask[(int64_t)g0 * A + lane] : 0u;
  for (int g = g0; g < n; g += wstride) {
    // the row kStages - 1 iterations ahead goes into the stage consumed in the previous iteration
    __syncwarp();
    fetch(g + (kStages - 1) * wstride, stage == 0 ? kStages - 1 : stage - 1);
    const bool legal = mask_next != 0u;
    mask_next = (g + wstride < n && lane < A) ? mask[(int64_t)(g + wstride) * A + lane] : 0u;
    cp_async_wait<kStages - 1>();  // this iteration's row has landed (this thread's copies) ...
    __syncwarp();                  // ... and everybody else's
    const uint32_t* words = reinterpret_cast<const uint32_t*>(rows + stage * kU17RowQuads);
    // the lane's 34 elements: words 17*lane .. 17*lane + 16; unit 0 = x[0..16], unit 1 = x[17..33]
    uint32_t wd[kUnit];
#pragma unroll
    for (int w = 0; w < kUnit; ++w) wd[w] = words[kUnit * lane + w];
    // unit maxima on the packed 16-bit pairs (one instruction per two elements): words 0..7 are
    // unit 0, words 9..16 unit 1, word 8 holds the last element of unit 0 and the first of unit 1
    uint32_t p0 = wd[0], p1 = wd[9];
#pragma unroll
    for (int w = 1; w < 8; ++w) { p0 = max2<T>(p0, wd[w]); p1 = max2<T>(p1, wd[9 + w]); }
    const float2 f0 = unpack2<T>(p0), f1 = unpack2<T>(p1), mid = unpack2<T>(wd[8]);
    const float m0 = fmaxf(fmaxf(f0.x, f0.y), mid.x), m1 = fmaxf(fmaxf(f1.x, f1.y), mid.y);
    float x[2 * kUnit];
#pragma unroll
    for (int w = 0; w < kUnit; ++w) {
      const float2 v = unpack2<T>(wd[w]);
      x[2 * w] = v.x;
      x[2 * w + 1] = v.y;
    }
    const float n0 = -m0 * kL2e, n1 = -m1 * kL2e;
    float s0 = 0.f, w0 = 0.f, s1 = 0.f, w1 = 0.f;
#pragma unroll
    for (int j = 0; j < kUnit; ++j) {
      const float e0 = ex2_fast(fmaf(x[j], kL2e, n0));
      const float e1 = ex2_fast(fmaf(x[kUnit + j], kL2e, n1));
      s0 += e0; w0 = fmaf(e0, z0[j], w0);
      s1 += e1; w1 = fmaf(e1, z1[j], w1);
    }
    if (u0 < units) { part[3 * u0] = m0; part[3 * u0 + 1] = s0; part[3 * u0 + 2] = w0; }
    if (u1 < units) { part[3 * u1] = m1; part[3 * u1 + 1] = s1; part[3 * u1 + 2] = w1; }
    __syncwarp();
    int key = INT_MIN;  // order-preserving integer image of q (illegal actions: below everything)
    if (legal) {
      const float* p = part + 9 * lane;  // units 3a, 3a+1, 3a+2 of action a = lane
      const float mm = fmaxf(p[0], fmaxf(p[3], p[6]));
      const float c0 = ex2_fast((p[0] - mm) * kL2e), c1 = ex2_fast((p[3] - mm) * kL2e),
                  c2 = ex2_fast((p[6] - mm) * kL2e);
      const float se = fmaf(p[1], c0, fmaf(p[4], c1, p[7] * c2));
      const float sw = fmaf(p[2], c0, fmaf(p[5], c1, p[8] * c2));
      const int qi = __float_as_int(sw / se);
      key = qi ^ ((qi >> 31) & 0x7fffffff);
      key = key == INT_MIN ? INT_MIN + 1 : key;
    }
    const unsigned legal_m = __ballot_sync(kFull, legal);
    const int best = __reduce_max_sync(kFull, key);
    int choice = __ffs((int)__ballot_sync(kFull, legal && key == best)) - 1;  // first maximal index
    if (legal_m == 0u) {
      choice = -1;
    } else if (epsilon > 0.0f) {
      const uint64_t bits = policy_bits(seed, step, (uint64_t)g);
      const float u01 = (float)(bits >> 40) * (1.0f / 16777216.0f);
      if (u01 <= epsilon) choice = kth_set_bit((uint64_t)legal_m, random_legal_rank(bits, __popc(legal_m)));
    }
    if (lane == 0) actions[g] = choice;
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
  cp_async_wait<0>();
}

}  // namespace hbp

int HbFail(const std::string& msg);  // hb_batch.cu
namespace {
int PFail(const std::string& s) { return HbFail(s); }
}  // namespace

namespace {
// Row-staged kernel when the rows are 16-byte aligned and fit the default dynamic
// shared memory (8 warps per CTA), else the streaming one.
template <class T>
cudaError_t LaunchSelectC51(const T* logits, int64_t row_pitch, const float* support, const uint8_t* mask,
                            int n, int A, int atoms, float epsilon, uint64_t seed, uint64_t step,
                            int32_t* actions, cudaStream_t s) {
  const unsigned grid = (unsigned)(((int64_t)n * 32 + 255) / 256);
  const size_t row_bytes = (((size_t)A * atoms * sizeof(T)) + 15) / 16 * 16;
  const bool aligned = (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && ((row_pitch * sizeof(T)) & 15) == 0;
  const size_t thirds_bytes = 512 + 8 * row_bytes + 8 * (size_t)3 * A * sizeof(float2);
  static const bool use_thirds = std::getenv("HANABI_B200_SELECT_THIRDS") == nullptr ||
                                 std::atoi(std::getenv("HANABI_B200_SELECT_THIRDS")) != 0;
  static const bool use_u17 = std::getenv("HANABI_B200_SELECT_U17") == nullptr ||
                              std::atoi(std::getenv("HANABI_B200_SELECT_U17")) != 0;
  if (use_u17 && sizeof(T) == 2 && aligned && atoms == 51 && 3 * A <= 64 && A <= 32) {
    // the Rainbow actor's shape: streaming kernel, persistent warps (a few CTAs per SM)
    typedef typename std::conditional<sizeof(T) == 2, T, __half>::type T16;
    static int resident_ctas = 0;  // CTAs of this kernel the device holds at once
    if (resident_ctas == 0) {
      int dev = 0, sms = 148, per_sm = 2;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncSetAttribute(hbp::k_select_c51_u17<T16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)hbp::kU17SmemBytes);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hbp::k_select_c51_u17<T16>, 32 * HB_U17_WARPS,
                                                    hbp::kU17SmemBytes);
      resident_ctas = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int64_t want = ((int64_t)n + HB_U17_WARPS - 1) / HB_U17_WARPS;
    const unsigned pgrid = (unsigned)(want < resident_ctas ? want : resident_ctas);
    hbp::k_select_c51_u17<T16><<<pgrid, 32 * HB_U17_WARPS, hbp::kU17SmemBytes, s>>>(reinterpret_cast<const T16*>(logits), row_pitch, support, mask,
                                                     n, A, epsilon, seed, step, actions);
  } else if (use_thirds && aligned && atoms <= 63 && thirds_bytes <= 48 * 1024) {
    if (atoms == 51)
      hbp::k_select_c51_thirds<T, 17><<<grid, 256, thirds_bytes, s>>>(logits, row_pitch, support, mask, n, A, atoms,
                                                                     epsilon, seed, step, actions, nullptr,
                                                                     (int)row_bytes);
    else
      hbp::k_select_c51_thirds<T, 0><<<grid, 256, thirds_bytes, s>>>(logits, row_pitch, support, mask, n, A, atoms,
                                                                    epsilon, seed, step, actions, nullptr,
                                                                    (int)row_bytes);
  } else if (aligned && atoms <= 63 && 512 + 8 * row_bytes <= 48 * 1024) {
    if (atoms == 51)
      hbp::k_select_c51_rows<T, 51><<<grid, 256, 512 + 8 * row_bytes, s>>>(logits, row_pitch, support, mask, n, A,
                                                                          atoms, epsilon, seed, step, actions,
                                                                          nullptr, (int)row_bytes);
    else
      hbp::k_select_c51_rows<T, 0><<<grid, 256, 512 + 8 * row_bytes, s>>>(logits, row_pitch, support, mask, n, A,
                                                                         atoms, epsilon, seed, step, actions,
                                                                         nullptr, (int)row_bytes);
  } else {
    hbp::k_select_c51<T><<<grid, 256, 0, s>>>(logits, row_pitch, support, mask, n, A, atoms, epsilon, seed,
                                              step, actions, nullptr);
  }
  return cudaGetLastError();
}
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
    hbp::k_select_c51<float><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
        q_or_logits, (int64_t)num_actions * num_atoms, support, mask, n, num_actions, num_atoms,
        epsilon, seed, step, actions, nullptr);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchSelectAction: ") + cudaGetErrorString(e));
}

// BatchSelectAction for a C51 head whose logits come straight out of a GEMM: element
// type float32 (dtype 0), bfloat16 (1) or float16 (2), and a row pitch (elements
// between two games' rows, >= num_actions * num_atoms) so that an N-padded GEMM output
// can be consumed in place.
int BatchSelectActionC51(const void* logits, int dtype, int64_t row_pitch, int num_atoms,
                         const float* support, const uint8_t* mask, int n, int num_actions,
                         float epsilon, uint64_t seed, uint64_t step, int32_t* actions,
                         void* stream) {
  if (!logits || !support || !mask || !actions) return PFail("BatchSelectActionC51: null argument");
  if (n <= 0 || num_actions <= 0 || num_atoms <= 0) return PFail("BatchSelectActionC51: bad sizes");
  if (row_pitch < (int64_t)num_actions * num_atoms) return PFail("BatchSelectActionC51: row_pitch too small");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  switch (dtype) {
    case 0: e = LaunchSelectC51(static_cast<const float*>(logits), row_pitch, support, mask, n, num_actions, num_atoms, epsilon, seed, step, actions, s); break;
    case 1: e = LaunchSelectC51(static_cast<const __nv_bfloat16*>(logits), row_pitch, support, mask, n, num_actions, num_atoms, epsilon, seed, step, actions, s); break;
    case 2: e = LaunchSelectC51(static_cast<const __half*>(logits), row_pitch, support, mask, n, num_actions, num_atoms, epsilon, seed, step, actions, s); break;
    default:
      return PFail("BatchSelectActionC51: dtype must be 0 (float32), 1 (bfloat16) or 2 (float16)");
  }
  return e == cudaSuccess ? 0 : PFail(std::string("BatchSelectActionC51: ") + cudaGetErrorString(e));
}

// Masked categorical policy head of the PPO / REINFORCE actors: see k_sample_masked.
// logits: float32 / bfloat16 / float16 (dtype 0 / 1 / 2) [n][row_pitch >= num_actions];
// log_probs (nullable): float32[n], log pi(action | state) under the masked softmax.
int BatchSampleMaskedCategorical(const void* logits, int dtype, int64_t row_pitch, const uint8_t* mask,
                                 int n, int num_actions, int greedy, uint64_t seed, uint64_t step,
                                 int32_t* actions, float* log_probs, void* stream) {
  if (!logits || !mask || !actions) return PFail("BatchSampleMaskedCategorical: null argument");
  if (n <= 0 || num_actions <= 0 || row_pitch < num_actions) return PFail("BatchSampleMaskedCategorical: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 127) / 128);
  switch (dtype) {
    case 0: hbp::k_sample_masked<float><<<grid, 128, 0, s>>>(static_cast<const float*>(logits), row_pitch, mask, n, num_actions, greedy, seed, step, actions, log_probs); break;
    case 1: hbp::k_sample_masked<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), row_pitch, mask, n, num_actions, greedy, seed, step, actions, log_probs); break;
    case 2: hbp::k_sample_masked<__half><<<grid, 128, 0, s>>>(static_cast<const __half*>(logits), row_pitch, mask, n, num_actions, greedy, seed, step, actions, log_probs); break;
    default: return PFail("BatchSampleMaskedCategorical: dtype must be 0 (float32), 1 (bfloat16) or 2 (float16)");
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchSampleMaskedCategorical: ") + cudaGetErrorString(e));
}

// Bit-packed observation rows (BatchStepEncodePacked / BatchObservePacked) -> the
// network's input matrix: out[i][j] = bit j of row i as 0/1 in float32 / bfloat16 /
// float16 (dtype as above), columns [obs_size, row_pitch) zero, so that K can be padded
// to the GEMM's alignment.  row_pitch must be a multiple of 8 and `out` 16-byte aligned.
int BatchExpandPacked(const uint32_t* obs_bits, int packed_words, int obs_size, int n, void* out,
                      int64_t row_pitch, int dtype, void* stream) {
  if (!obs_bits || !out) return PFail("BatchExpandPacked: null argument");
  if (n <= 0 || obs_size <= 0 || packed_words * 32 < obs_size) return PFail("BatchExpandPacked: bad sizes");
  if (row_pitch < obs_size || (row_pitch & 7)) return PFail("BatchExpandPacked: row_pitch must be >= obs_size and a multiple of 8");
  if (reinterpret_cast<uintptr_t>(out) & 15) return PFail("BatchExpandPacked: out must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t threads = (int64_t)n * ((row_pitch + 31) >> 5);
  const unsigned grid = (unsigned)((threads + 255) / 256);
  switch (dtype) {
    case 0: hbp::k_expand_bits<float><<<grid, 256, 0, s>>>(obs_bits, packed_words, obs_size, n, static_cast<float*>(out), row_pitch); break;
    case 1: hbp::k_expand_bits<__nv_bfloat16><<<grid, 256, 0, s>>>(obs_bits, packed_words, obs_size, n, static_cast<__nv_bfloat16*>(out), row_pitch); break;
    case 2: hbp::k_expand_bits<__half><<<grid, 256, 0, s>>>(obs_bits, packed_words, obs_size, n, static_cast<__half*>(out), row_pitch); break;
    default: return PFail("BatchExpandPacked: dtype must be 0 (float32), 1 (bfloat16) or 2 (float16)");
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : PFail(std::string("BatchExpandPacked: ") + cudaGetErrorString(e));
}

// C51 q-values q[n, A] = sum_i support_i softmax(logits[n, a, :])_i (all actions).
int BatchC51QValues(const float* logits, const float* support, const uint8_t* mask, int n,
                    int num_actions, int num_atoms, float* q_out, int32_t* greedy_actions,
                    void* stream) {
  if (!logits || !support || !mask || !q_out || !greedy_actions) return PFail("BatchC51QValues: null argument");
  const int64_t threads = (int64_t)n * 32;
  hbp::k_select_c51<float><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      logits, (int64_t)num_actions * num_atoms, support, mask, n, num_actions, num_atoms, 0.0f, 0, 0,
      greedy_actions, q_out);
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
