// hb_host.cc -- host-side helpers of the batched C-ABI that run on CPU threads:
// the uniform-random-legal policy for HOST masks (the host twin of the device
// select kernels) on a persistent worker pool.
//
// The pool exists because a host consumer of BatchStepEncodeHostRange calls the
// policy once per range of games, several thousand times a second: spawning
// std::threads per call (round 1) cost more than the policy itself and did not
// scale when eight ranks shared one host.  Workers are created once, sized
// cores / local world size (torchrun's LOCAL_WORLD_SIZE) unless
// HANABI_B200_HOST_THREADS says otherwise, spin briefly for the next job and
// then sleep on a condition variable.
#include <sched.h>
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
