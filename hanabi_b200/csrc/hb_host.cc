// hb_host.cc -- host-side helpers of the batched C-ABI that run on CPU threads, on one
// persistent worker pool:
//   * the uniform-random-legal policy for HOST masks (the host twin of the device select
//     kernels; AVX2 + BMI2 where the CPU has them);
//   * HostExpandPacked, the host half of the packed link: bit-packed observation rows
//     (what crosses PCIe) -> the caller's uint8 rows of 0/1 (AVX2, non-temporal stores).
// Nothing here computes game logic: both are transport / consumer-side helpers of the
// host-buffer entry points.
//
// The pool exists because a host consumer of BatchStepEncodeHostRange calls the
// policy once per range of games, several thousand times a second: spawning
// std::threads per call (round 1) cost more than the policy itself and did not
// scale when eight ranks shared one host.  Workers are created once, sized
// cores / local world size (torchrun's LOCAL_WORLD_SIZE) unless
// HANABI_B200_HOST_THREADS says otherwise, spin briefly for the next job and
// then sleep on a condition variable.
#include <sched.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <stdint.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyhanabi.h"
#include "hb_policy_rng.h"

int HbFail(const std::string& msg);  // hb_batch.cu: text for BatchLastError()

namespace {

class HostPool {
 public:
  static HostPool& Get() {
    static HostPool* pool = new HostPool;  // leaked on purpose: workers may outlive static dtors
    return *pool;
  }

  int size() const { return num_workers_ + 1; }

  // Runs fn(part) for part in [0, parts) on the pool (the caller takes parts too).
  void Run(int parts, const std::function<void(int)>& fn) {
    if (parts <= 1 || num_workers_ == 0) {
      for (int p = 0; p < parts; ++p) fn(p);
      return;
    }
    std::lock_guard<std::mutex> serial(run_mutex_);  // one job at a time
    const uint64_t gen = (ticket_.load(std::memory_order_relaxed) >> 32) + 1;
    Job& j = jobs_[gen & 3];
    j.gen.store(0, std::memory_order_release);  // descriptor invalid while it is rewritten
    j.fn = &fn;
    j.parts = parts;
    j.left.store(parts, std::memory_order_relaxed);
    j.gen.store(gen, std::memory_order_release);
    ticket_.store(gen << 32, std::memory_order_release);  // part counter restarts at 0
    { std::lock_guard<std::mutex> lk(m_); }
    cv_.notify_all();
    Drain();
    int spins = 0;
    while (j.left.load(std::memory_order_acquire) != 0)
      if (++spins > 4000) std::this_thread::yield();
  }

 private:
  struct Job {
    std::atomic<uint64_t> gen{0};
    const std::function<void(int)>* fn = nullptr;
    int parts = 0;
    std::atomic<int> left{0};
  };

  HostPool() {
    int n = 0;
    if (const char* env = std::getenv("HANABI_B200_HOST_THREADS")) n = std::atoi(env);
    if (n <= 0) {
      const int hw = (int)std::thread::hardware_concurrency();
      n = hw;
      cpu_set_t set;
      CPU_ZERO(&set);
      if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = CPU_COUNT(&set);
      int local_world = 1;
      if (const char* env = std::getenv("LOCAL_WORLD_SIZE")) local_world = std::atoi(env);
      // a process already pinned to its share of the cores keeps all of them; an unpinned
      // one of several local ranks takes its fraction
      if (local_world > 1 && n >= hw) n /= local_world;
    }
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    num_workers_ = n - 1;
    for (int i = 0; i < num_workers_; ++i) std::thread([this] { Worker(); }).detach();
  }

  // Claims parts of the current job until none is left; returns the generation seen last.
  // A ticket carries the generation it was drawn from, so a part index can never be applied
  // to another job; the descriptor is read seqlock-style (fields first, generation after).
  uint64_t Drain() {
    for (;;) {
      const uint64_t t = ticket_.fetch_add(1, std::memory_order_acq_rel);
      const uint64_t gen = t >> 32;
      const uint32_t part = (uint32_t)t;
      Job& j = jobs_[gen & 3];
      const std::function<void(int)>* fn = j.fn;
      const int parts = j.parts;
      if (j.gen.load(std::memory_order_acquire) != gen || part >= (uint32_t)parts) return gen;
      (*fn)((int)part);
      j.left.fetch_sub(1, std::memory_order_release);
    }
  }

  void Worker() {
    uint64_t seen = 0;
    auto fresh = [&] { return (ticket_.load(std::memory_order_acquire) >> 32) != seen; };
    for (;;) {
      bool got = false;
      const auto t0 = std::chrono::steady_clock::now();
      while (std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200))
        if (fresh()) { got = true; break; }
      if (!got) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, fresh);
      }
      seen = Drain();
    }
  }

  int num_workers_ = 0;
  std::mutex m_, run_mutex_;
  std::condition_variable cv_;
  Job jobs_[4];
  std::atomic<uint64_t> ticket_{0};  // generation << 32 | next part of that generation
};

// ------------------------------------------------------- random legal policy rows
// actions[g] for rows [lo, hi): the legal bytes of a row gathered into a bit set, then the
// (rank)-th set bit for the draw of hb_policy_rng.h.
void PolicyRowsScalar(const uint8_t* mask, int lo, int hi, int num_actions, uint64_t seed,
                      uint64_t step, int64_t first_game_index, int32_t* actions) {
  for (int g = lo; g < hi; ++g) {
    const uint8_t* m = mask + (int64_t)g * num_actions;
    // eight mask bytes per step: any non-zero byte -> 1, then the eight low bits gathered
    // with one multiply
    uint64_t legal = 0;
    int a = 0;
    for (; a + 8 <= num_actions; a += 8) {
      uint64_t x;
      std::memcpy(&x, m + a, 8);
      x |= x >> 4; x |= x >> 2; x |= x >> 1;
      x &= 0x0101010101010101ull;
      legal |= ((x * 0x0102040810204080ull) >> 56) << a;
    }
    for (; a < num_actions; ++a) legal |= (uint64_t)(m[a] != 0) << a;
    const int n_legal = __builtin_popcountll(legal);
    int choice = -1;
    if (n_legal > 0) {
      const uint64_t bits = hbp::policy_bits(seed, step, (uint64_t)(first_game_index + g));
      int k = hbp::random_legal_rank(bits, n_legal);
      while (k-- > 0) legal &= legal - 1;
      choice = __builtin_ctzll(legal);
    }
    actions[g] = choice;
  }
}

#if defined(__x86_64__)
// Same rows with one or two 32-byte loads per row (compare + movemask) and the k-th set bit
// by pdep: no data-dependent loop, no branch on the mask's contents.
__attribute__((target("avx2,bmi2"))) void PolicyRowsAvx2(const uint8_t* mask, int lo, int hi, int n,
                                                       int num_actions, uint64_t seed, uint64_t step,
                                                       int64_t first_game_index, int32_t* actions) {
  // rows whose 64-byte window would run past the end of the array take the scalar path
  const int64_t total = (int64_t)n * num_actions;
  int safe_hi = hi;
  while (safe_hi > lo && (int64_t)(safe_hi - 1) * num_actions + 64 > total) --safe_hi;
  const __m256i zero = _mm256_setzero_si256();
  const uint64_t keep = num_actions >= 64 ? ~0ull : ((1ull << num_actions) - 1ull);
  for (int g = lo; g < safe_hi; ++g) {
    const uint8_t* m = mask + (int64_t)g * num_actions;
    const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(m));
    uint64_t legal = (uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(v0, zero));
    if (num_actions > 32) {
      const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(m + 32));
      legal |= (uint64_t)(uint32_t)~_mm256_movemask_epi8(_mm256_cmpeq_epi8(v1, zero)) << 32;
    }
    legal &= keep;
    const int n_legal = __builtin_popcountll(legal);
    const uint64_t bits = hbp::policy_bits(seed, step, (uint64_t)(first_game_index + g));
    const int k = hbp::random_legal_rank(bits, n_legal > 0 ? n_legal : 1);
    const uint64_t pick = _pdep_u64(1ull << k, legal);  // the k-th set bit of legal (0 when none)
    actions[g] = n_legal > 0 ? (int32_t)__builtin_ctzll(pick) : -1;
  }
  PolicyRowsScalar(mask, safe_hi, hi, num_actions, seed, step, first_game_index, actions);
}
bool HaveAvx2Bmi2() {
  static const bool have = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  return have;
}
#else
bool HaveAvx2Bmi2() { return false; }
void PolicyRowsAvx2(const uint8_t*, int, int, int, int, uint64_t, uint64_t, int64_t, int32_t*) {}
#endif

}  // namespace

extern "C" {

// Host-side twin of BatchSelectAction for epsilon >= 1 (uniform random legal move;
// the reference's agents/random_agent.py for a whole batch): same counter-based
// generator, so for equal (seed, step, game) it returns exactly the device kernels'
// actions.  mask and actions are HOST arrays of n rows; row i is game
// first_game_index + i of the key.  `threads` <= 0 uses the pool's size.
int HostRandomLegalActionsAt(const uint8_t* mask, int n, int num_actions, uint64_t seed,
                             uint64_t step, int64_t first_game_index, int32_t* actions,
                             int threads) {
  if (!mask || !actions || n < 0 || num_actions <= 0 || num_actions > 64)
    return HbFail("HostRandomLegalActions: bad argument");
  HostPool& pool = HostPool::Get();
  if (threads <= 0 || threads > pool.size()) threads = pool.size();
  int parts = threads;
  if ((int64_t)parts * 2048 > n) parts = n / 2048 > 0 ? n / 2048 : 1;
  const std::function<void(int)> work = [=](int t) {
    const int lo = (int)((int64_t)n * t / parts), hi = (int)((int64_t)n * (t + 1) / parts);
    if (HaveAvx2Bmi2()) PolicyRowsAvx2(mask, lo, hi, n, num_actions, seed, step, first_game_index, actions);
    else PolicyRowsScalar(mask, lo, hi, num_actions, seed, step, first_game_index, actions);
  };
  pool.Run(parts, work);
  return 0;
}

int HostRandomLegalActions(const uint8_t* mask, int n, int num_actions, uint64_t seed,
                           uint64_t step, int32_t* actions, int threads) {
  return HostRandomLegalActionsAt(mask, n, num_actions, seed, step, 0, actions, threads);
}

int HostPolicyThreads(void) { return HostPool::Get().size(); }

}  // extern "C"

// ------------------------------------------------------- packed rows -> bytes
// The host half of the "packed link" transfer (BatchSetHostLink): the step kernel's
// bit-packed observation rows cross PCIe (96 B instead of 658 B per Full-2p game) and
// the worker pool turns them into the caller's uint8 rows of 0/1 -- the same bytes the
// kernel's own expansion would have written, with 6.9 x fewer bytes on the link.
namespace {

// Eight bits -> eight bytes of 0/1 (byte i = bit i), portable form.
inline uint64_t Bits8ToBytes(uint32_t b8) {
  uint64_t x = ((uint64_t)(b8 & 0xffu) * 0x0101010101010101ull) & 0x8040201008040201ull;
  return ((x + 0x7f7f7f7f7f7f7f7full) >> 7) & 0x0101010101010101ull;
}

// 64 bits of row `row` from bit `bit` on (bits beyond the buffer read as zero).
inline uint64_t LoadBits(const uint8_t* rows, int64_t total_bytes, int64_t row_off, int bit) {
  const int64_t at = row_off + (bit >> 3);
  uint64_t w = 0;
  if (at + 8 <= total_bytes) std::memcpy(&w, rows + at, 8);
  else if (at < total_bytes) std::memcpy(&w, rows + at, (size_t)(total_bytes - at));
  return w >> (bit & 7);  // >= 57 valid bits
}

// Bits [pos, pos + 32) of the dense bit stream "row r = bits [r * dim, (r + 1) * dim)",
// read from rows of row_bytes bytes; a chunk may straddle two rows (dim >= 32).
inline uint32_t StreamBits32(const uint8_t* rows, int64_t total_bytes, int row_bytes, int dim,
                             int64_t r, int bit) {
  uint64_t w = LoadBits(rows, total_bytes, r * row_bytes, bit);
  const int avail = dim - bit;
  if (avail < 32) {
    const uint64_t next = LoadBits(rows, total_bytes, (r + 1) * row_bytes, 0);
    w = (w & ((1ull << avail) - 1ull)) | (next << avail);
  }
  return (uint32_t)w;
}

void ExpandDenseScalar(const uint8_t* rows, int64_t total_bytes, int row_bytes, int dim,
                       uint8_t* out, int64_t lo, int64_t hi) {
  // bytes [lo, hi) of the dense output; byte j = bit j % dim of row j / dim
  int64_t r = lo / dim;
  int bit = (int)(lo - r * dim);
  int64_t j = lo;
  while (j < hi) {
    int run = dim - bit;
    if ((int64_t)run > hi - j) run = (int)(hi - j);
    int done = 0;
    while (done + 8 <= run) {
      const uint64_t v = Bits8ToBytes((uint32_t)LoadBits(rows, total_bytes, r * row_bytes, bit + done));
      std::memcpy(out + j + done, &v, 8);
      done += 8;
    }
    if (done < run) {
      const uint64_t v = Bits8ToBytes((uint32_t)LoadBits(rows, total_bytes, r * row_bytes, bit + done));
      std::memcpy(out + j + done, &v, (size_t)(run - done));
    }
    j += run;
    bit += run;
    if (bit == dim) { bit = 0; ++r; }
  }
}

#if defined(__x86_64__)
// Aligned 32-byte chunks [lo, hi) (both multiples of 32, out + lo 32-byte aligned) with
// non-temporal stores: the rows are written once and read by somebody else, so they
// should not be read for ownership nor displace the caller's working set.
__attribute__((target("avx2"))) void ExpandDenseAvx2(const uint8_t* rows, int64_t total_bytes,
                                                    int row_bytes, int dim, uint8_t* out,
                                                    int64_t lo, int64_t hi) {
  const __m256i shuf = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1,
                                        2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3);
  const __m256i bitm = _mm256_set1_epi64x((long long)0x8040201008040201ull);
  const __m256i ones = _mm256_set1_epi8(1);
  int64_t r = lo / dim;
  int bit = (int)(lo - r * dim);
  for (int64_t j = lo; j < hi; j += 32) {
    const uint32_t bits = StreamBits32(rows, total_bytes, row_bytes, dim, r, bit);
    __m256i v = _mm256_shuffle_epi8(_mm256_set1_epi32((int)bits), shuf);
    v = _mm256_min_epu8(_mm256_and_si256(v, bitm), ones);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(out + j), v);
    bit += 32;
    if (bit >= dim) { bit -= dim; ++r; }
  }
  _mm_sfence();
}
bool HaveAvx2() {
  static const bool have = __builtin_cpu_supports("avx2");
  return have;
}
#else
bool HaveAvx2() { return false; }
void ExpandDenseAvx2(const uint8_t*, int64_t, int, int, uint8_t*, int64_t, int64_t) {}
#endif

}  // namespace

extern "C" {

// Bit-packed observation rows (HOST, packed_words 32-bit words per row, bit j of a row =
// observation bit j; what BatchStepEncodeHostPacked delivers) -> HOST rows of obs_size
// bytes of 0/1, row_pitch bytes apart: the CPU twin of the step kernel's byte expansion,
// on the worker pool.  `threads` <= 0 uses the pool's size.
int HostExpandPacked(const uint32_t* obs_bits, int packed_words, int obs_size, int64_t n,
                     uint8_t* out, int64_t row_pitch, int threads) {
  if (!obs_bits || !out || n < 0 || packed_words <= 0 || obs_size <= 0 ||
      obs_size > packed_words * 32 || row_pitch < obs_size)
    return HbFail("HostExpandPacked: bad argument");
  if (n == 0) return 0;
  HostPool& pool = HostPool::Get();
  if (threads <= 0 || threads > pool.size()) threads = pool.size();
  const uint8_t* rows = reinterpret_cast<const uint8_t*>(obs_bits);
  const int row_bytes = packed_words * 4;
  const int64_t total_bytes = n * (int64_t)row_bytes;
  if (row_pitch == obs_size && obs_size >= 64) {
    // dense rows: one contiguous byte range, cut into 32-byte aligned spans
    const int64_t total = n * (int64_t)obs_size;
    int parts = threads;
    if ((int64_t)parts * (1 << 16) > total) parts = (int)(total >> 16) > 0 ? (int)(total >> 16) : 1;
    const uintptr_t base = reinterpret_cast<uintptr_t>(out);
    const std::function<void(int)> work = [=](int t) {
      // span [lo, hi) with both ends rounded to 32-byte aligned ADDRESSES (except the outer ends)
      auto cut = [&](int k) -> int64_t {
        if (k <= 0) return 0;
        if (k >= parts) return total;
        const int64_t c = total * k / parts;
        int64_t a = c - (int64_t)((base + (uintptr_t)c) & 31u);
        return a < 0 ? 0 : a;
      };
      const int64_t lo = cut(t), hi = cut(t + 1);
      if (lo >= hi) return;
      int64_t a_lo = lo + (int64_t)((32u - ((base + (uintptr_t)lo) & 31u)) & 31u);
      if (a_lo > hi) a_lo = hi;
      const int64_t a_hi = a_lo + ((hi - a_lo) & ~(int64_t)31);
      if (HaveAvx2() && a_hi > a_lo) {
        ExpandDenseScalar(rows, total_bytes, row_bytes, obs_size, out, lo, a_lo);
        ExpandDenseAvx2(rows, total_bytes, row_bytes, obs_size, out, a_lo, a_hi);
        ExpandDenseScalar(rows, total_bytes, row_bytes, obs_size, out, a_hi, hi);
      } else {
        ExpandDenseScalar(rows, total_bytes, row_bytes, obs_size, out, lo, hi);
      }
    };
    pool.Run(parts, work);
    return 0;
  }
  // padded rows: row by row
  int parts = threads;
  if ((int64_t)parts * 256 > n) parts = (int)(n / 256) > 0 ? (int)(n / 256) : 1;
  const std::function<void(int)> work = [=](int t) {
    const int64_t lo = n * t / parts, hi = n * (t + 1) / parts;
    for (int64_t g = lo; g < hi; ++g) {
      uint8_t* dst = out + g * row_pitch;
      int done = 0;
      while (done + 8 <= obs_size) {
        const uint64_t v = Bits8ToBytes((uint32_t)LoadBits(rows, total_bytes, g * row_bytes, done));
        std::memcpy(dst + done, &v, 8);
        done += 8;
      }
      if (done < obs_size) {
        const uint64_t v = Bits8ToBytes((uint32_t)LoadBits(rows, total_bytes, g * row_bytes, done));
        std::memcpy(dst + done, &v, (size_t)(obs_size - done));
      }
    }
  };
  pool.Run(parts, work);
  return 0;
}

}  // extern "C"
