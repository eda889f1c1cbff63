// miso_b200/csrc/philox.cuh -- the random stream of the product ("miso-b200 stream" v1 / v2).
//
// The reference draws from whatever RNG sits behind its vtable
// (/root/reference/pysplicing/include/splicing_random.h:23-36,96-103) and never
// seeds it; this framework defines the stream instead:
//   Philox4x32, key = seed, counter = (block, tag, gene_id, chain)
//   stream v2 (default): 7 rounds -- the smallest round count of Philox4x32 that passes
//   BigCrush (Salmon et al., SC'11, table 2); the generator is 30 % of the chain kernel's
//   instructions, three rounds less are ~9 % of the step.  stream v1: the 10 rounds of
//   round 1 (misob200_stream_version(1), MISOB200_STREAM=1); golden vectors exist for both.
//   uniform n : word n&3 of block n>>2, tag 0   ->  (w + 0.5) * 2^-32
//   normal  n : block n, tag 1  ->  Box-Muller on two 53-bit uniforms
// A gene-chain consumes uniforms and normals in the order the reference's loop
// would, so the unmodified reference driven by the same
// stream through its vtable makes the same decisions.
#pragma once
#include <cstdint>

namespace misob200 {

#define MISOB200_PHILOX_M0 0xD2511F53u
#define MISOB200_PHILOX_M1 0xCD9E8D57u
#define MISOB200_PHILOX_W0 0x9E3779B9u
#define MISOB200_PHILOX_W1 0xBB67AE85u

// The ten round keys (k0 + r*W0, k1 + r*W1) are the same for every draw of a
// run (the key is the seed), so the host expands them once and they travel in
// the kernel parameter block: on the device they are constant-bank operands of
// the round's XOR, not instructions.
// The round count is a compile-time property of the key type, hence of every kernel: the two
// stream versions are separate instantiations and only the selected one is ever resident in
// the instruction caches (a run-time branch around three extra rounds at every call site cost
// more in instruction fetch than the rounds it skipped).
template <int R> struct PhiloxKeyT {
  static constexpr int kRounds = R;
  uint32_t k0[10], k1[10];
};
using PhiloxKey = PhiloxKeyT<7>;       // layout of every PhiloxKeyT<R>

inline PhiloxKey philox_expand_key(uint64_t seed) {
  PhiloxKey k;
  uint32_t a = (uint32_t) seed, b = (uint32_t) (seed >> 32);
  for (int r = 0; r < 10; r++) {
    k.k0[r] = a; k.k1[r] = b;
    a += MISOB200_PHILOX_W0; b += MISOB200_PHILOX_W1;
  }
  return k;
}

#ifdef __CUDACC__
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
  const unsigned long long p0 = (unsigned long long) MISOB200_PHILOX_M0 * c0;   // IMAD.WIDE.U32
  const unsigned long long p1 = (unsigned long long) MISOB200_PHILOX_M1 * c2;
  const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
  const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
  c1 = (uint32_t) p1; c3 = (uint32_t) p0;
  c0 = n0; c2 = n2;
}
template <class KEY>
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           const KEY &key, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < KEY::kRounds; r++) philox_round(c0, c1, c2, c3, key.k0[r], key.k1[r]);
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// (w + 0.5) * 2^-32, exact in fp64
__device__ __forceinline__ double uniform_from_word(uint32_t w) {
  return __fma_rn((double) w, 0x1p-32, 0x1p-33);   // one instruction; exact either way
}

template <class KEY>
__device__ __forceinline__ double stream_uniform(unsigned long long n, uint32_t gene, uint32_t chain,
                                                 const KEY &key) {
  uint32_t x[4];
  philox4x32((uint32_t) (n >> 2), 0u, gene, chain, key, x);
  const uint32_t sel = (uint32_t) n & 3u;
  const uint32_t w = sel == 0 ? x[0] : sel == 1 ? x[1] : sel == 2 ? x[2] : x[3];
  return uniform_from_word(w);
}

template <class KEY>
__device__ __noinline__ double stream_normal(uint32_t n, uint32_t gene, uint32_t chain,
                                                const KEY &key) {
  uint32_t x[4];
  philox4x32(n, 1u, gene, chain, key, x);
  const unsigned long long a = ((unsigned long long) x[0] << 21) | (x[1] >> 11);
  const unsigned long long b = ((unsigned long long) x[2] << 21) | (x[3] >> 11);
  const double u1 = (double) (a + 1ull) * 0x1p-53;
  const double u2 = (double) b * 0x1p-53;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
#endif

}  // namespace misob200
