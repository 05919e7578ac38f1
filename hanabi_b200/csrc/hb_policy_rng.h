// hb_policy_rng.h -- counter-based policy randomness shared by the select kernels
// (hb_policy.cu), their host twin and the step kernel's fused random-action output.
#ifndef HB_POLICY_RNG_H_
#define HB_POLICY_RNG_H_
#include <stdint.h>
#if defined(__CUDACC__)
#define HBP_HD __host__ __device__ __forceinline__
#else
#define HBP_HD inline
#endif
namespace hbp {

HBP_HD uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Keyed (seed, step, game): reproducible and independent of launch geometry.  The
// reference uses Python's `random` and numpy's global RNG (dqn_agent.py:409-412),
// which cannot be reproduced on a device; only the distribution is matched (uniform
// over legal moves with probability epsilon).
HBP_HD uint64_t policy_step_key(uint64_t seed, uint64_t step) {  // the part that is the same for every game
  return mix64(seed ^ (step * 0xD1B54A32D192ED03ull));
}
HBP_HD uint64_t policy_bits_keyed(uint64_t step_key, uint64_t game) {
  return mix64(step_key ^ (game * 0x8CB92BA72F3D8DD7ull));
}
HBP_HD uint64_t policy_bits(uint64_t seed, uint64_t step, uint64_t game) {
  return policy_bits_keyed(policy_step_key(seed, step), game);
}

// Index (0-based, among n_legal legal moves in ascending uid order) of the uniformly
// drawn move.
HBP_HD int random_legal_rank(uint64_t bits, int n_legal) {
  return (int)(((bits & 0xffffffffull) * (uint64_t)n_legal) >> 32);
}

}  // namespace hbp
#endif  // HB_POLICY_RNG_H_
