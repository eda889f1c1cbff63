/*
 * oracle/philox_ref.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C statement of the random stream the whole repo agrees on
 * ("miso-b200 stream", v1 = Philox4x32-10, v2 = Philox4x32-7; phx_rounds below,
 * set through refh_set_stream / mo_set_stream; the product's default is v2).  The reference never seeds its generator
 * (the only RNG hook is the vtable
 * splicing_rng_type_t, /root/reference/pysplicing/include/splicing_random.h:23-36),
 * so "same seed" is *defined* here and injected into the unmodified reference
 * through that vtable (oracle/ref_harness.c) and restated in
 * oracle/miso_oracle.c; the CUDA product implements the same map in
 * miso_b200/csrc/philox.cuh.
 *
 *   generator : Philox4x32-R (Salmon, Moraes, Dror, Shaw, SC'11), R = phx_rounds:
 *               10 (stream v1) or 7 (stream v2, the fewest rounds that pass BigCrush)
 *   key       : (seed & 0xffffffff, seed >> 32)
 *   counter   : (block, tag, gene_id, chain_id)    tag 0 = uniforms, 1 = normals
 *   uniform n : word (n & 3) of block (n >> 2), tag 0   ->  (w + 0.5) * 2^-32
 *   normal  n : block n, tag 1, words x0..x3
 *                 a = (x0 << 21) | (x1 >> 11)     53 bits
 *                 b = (x2 << 21) | (x3 >> 11)     53 bits
 *                 z = sqrt(-2 ln((a + 1) 2^-53)) * cos(2 pi * b 2^-53)
 *   gamma(a=1, scale) : scale * -ln(next uniform)   (an Exp(1) draw by inversion)
 *                 -- the only shape the path asks for: the START_RANDOM
 *                 Dirichlet draw, src/miso.c:309-326 called with alpha = 1 at
 *                 :388-404; consumes one uniform of the tag-0 sequence.
 * The n-th call of get_real / get_norm inside one splicing_miso[_paired]
 * invocation returns uniform n / normal n (SURVEY.md appendix C gives the
 * order in which the reference consumes them).
 */
#ifndef MISO_ORACLE_PHILOX_REF_H
#define MISO_ORACLE_PHILOX_REF_H

#include <stdint.h>
#include <math.h>

#define PHX_M0 0xD2511F53u
#define PHX_M1 0xCD9E8D57u
#define PHX_W0 0x9E3779B9u
#define PHX_W1 0xBB67AE85u

static int phx_rounds = 7;	/* stream v2; 10 = stream v1 */

static inline void phx_4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                               uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  int r;
  for (r = 0; r < phx_rounds; r++) {
    uint64_t p0 = (uint64_t) PHX_M0 * c0;
    uint64_t p1 = (uint64_t) PHX_M1 * c2;
    uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t) p1;
    uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t) p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHX_W0; k1 += PHX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
  uint64_t seed;
  uint32_t gene, chain;
  uint64_t n_unif, n_norm;	/* number of draws handed out so far */
} phx_stream_t;

static inline double phx_uniform_at(const phx_stream_t *s, uint64_t n) {
  uint32_t ctr[4], key[2], w[4];
  ctr[0] = (uint32_t) (n >> 2); ctr[1] = 0u; ctr[2] = s->gene; ctr[3] = s->chain;
  key[0] = (uint32_t) s->seed; key[1] = (uint32_t) (s->seed >> 32);
  phx_4x32_10(ctr, key, w);
  return (double) w[n & 3] * 0x1p-32 + 0x1p-33;
}

static inline double phx_normal_at(const phx_stream_t *s, uint64_t n) {
  uint32_t ctr[4], key[2], w[4];
  uint64_t a, b;
  double u1, u2;
  ctr[0] = (uint32_t) n; ctr[1] = 1u; ctr[2] = s->gene; ctr[3] = s->chain;
  key[0] = (uint32_t) s->seed; key[1] = (uint32_t) (s->seed >> 32);
  phx_4x32_10(ctr, key, w);
  a = ((uint64_t) w[0] << 21) | (w[1] >> 11);
  b = ((uint64_t) w[2] << 21) | (w[3] >> 11);
  u1 = (double) (a + 1) * 0x1p-53;
  u2 = (double) b * 0x1p-53;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static inline double phx_next_uniform(phx_stream_t *s) {
  return phx_uniform_at(s, s->n_unif++);
}
static inline double phx_next_normal(phx_stream_t *s) {
  return phx_normal_at(s, s->n_norm++);
}
/* gamma(1, scale); other shapes are never requested on the sampler path */
static inline double phx_next_gamma1(phx_stream_t *s, double scale) {
  return scale * -log(phx_next_uniform(s));
}

#endif
