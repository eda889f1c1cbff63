// miso_b200/csrc/class_pass.cuh -- reassignment passes over CLASS tiles (format 1).
//
// The reference decides read r of a pass (src/miso.c:59-83, src/miso_paired.c:45-78) by
//     rnd = u_r * C_{K-1},   chosen = #{k < K-1 : rnd > C_k}   (>= for two compatible)
// with C_k the running sum of psi_j * p_{r,j}.  u_r = (w_r + 0.5) * 2^-32 comes from
// one 32-bit Philox word, rnd is monotone in w_r, so every test is an INTEGER
// threshold on the raw word:  test k true  <=>  w_r > t_k.  t_k depends on the read
// only through the direction of its weight vector -- its weight class
// (plan.cpp) -- so it is computed once per (class, iteration) in fp64 by one lane,
// and the per-read work of a pass shrinks to "load the class's K-1 thresholds,
// compare, count": no fp64, no probability lookups, one byte of tile per read.
//
// Exactness.  Let rho_k = C_k / C_{K-1} in exact arithmetic.  The reference's test is
// true iff w + 0.5 > 2^32 rho_k (1 + eta), where eta collects the roundings of its
// products, sums and of u * C (|eta| <= (2n+3) 2^-53 for n <= 8 compatible isoforms).
// thr_update computes tau = C_k * (2^32 / C_{K-1}) - 0.5 with its own roundings of the
// same size; both errors together stay below 1.8e-5 in units of w (tau <= 2^32, so
// absolute errors are ~2^32 * 17 * 2^-53 + two half-ulps of 2^32).  If tau lies more
// than kThrMargin = 2^-15 away from every integer, no integer w can fall between the
// reference's boundary and tau, hence  test  <=>  w > floor(tau)  for every read of
// the class -- for the `>` and the `>=` rule alike.  Otherwise (probability
// ~6e-5 per threshold) the pass is declined and the caller runs the literal fp64
// rule for this iteration.  Uniform-code classes use weight 1.0 in place of
// ptab[code]: the common factor cancels in rho_k and its rounding is inside eta.
// tests/test_thresholds.py checks the claim against the oracle's arithmetic at the
// boundary words.
#pragma once
#include "philox.cuh"
#include "plan.hpp"
#include "tile_mem.cuh"

namespace misob200 {

constexpr double kThrMargin = 0x1p-15;

// Philox blocks in flight per lane in the counting pass / the read-score pass
#ifndef MISOB200_UNROLL_COUNT
#define MISOB200_UNROLL_COUNT 2
#endif
#ifndef MISOB200_UNROLL_SCORE
#define MISOB200_UNROLL_SCORE 1
#endif
#ifndef MISOB200_UNROLL1_FROM_K
#define MISOB200_UNROLL1_FROM_K 5
#endif
constexpr int kUnrollCount = MISOB200_UNROLL_COUNT, kUnrollScore = MISOB200_UNROLL_SCORE;
// The counting loop is unrolled by two only for K <= 4.  From K = 5 on, one Philox block per trip:
// the unrolled body (~370 instructions at K = 8) plus the per-iteration scalar code no longer fit
// the instruction caches (6 KB L0, 32 KB L1.5) and the fetch stalls cost more than the second
// block's ILP brings: K = 8 181 -> 167 ms, K = 7 152 -> 145, K = 6 130 -> 122, K = 5 105 -> 104,
// K = 4 unchanged (profiles/README.md, r1_ab17).
constexpr int kUnroll1FromK = MISOB200_UNROLL1_FROM_K;

// Threshold rows live in two planes so that a row never spans more than 16 bytes: plane A
// holds t_0..t_3 of every class (stride TSA <= 16), plane B t_4..t_6 (stride TSB, K >= 6).
// A warp's LDS.128 is served a quarter-warp at a time; with one 32-byte row per class only
// four of the eight 16-byte bank groups were ever addressed (2-way conflicts at best, ncu
// showed the LSU pipe ~85 % busy at K = 8); with 16-byte rows class c sits in group c mod 8.
// plan.cpp numbers the classes by falling read count, so the frequent ones get distinct groups.
template <int K> struct Thr {
  static constexpr int NT = K - 1;                                          // thresholds per class
  static constexpr int TSA = NT <= 1 ? 4 : NT <= 2 ? 8 : 16;                // bytes per class, plane A
  static constexpr int TSB = NT <= 4 ? 0 : NT == 5 ? 4 : NT == 6 ? 8 : 16;  // plane B
  static __host__ __device__ constexpr int plane_a_bytes(int ncls) { return ((ncls + 1) * TSA + 15) & ~15; }
  static __host__ __device__ constexpr int bytes(int ncls) { return plane_a_bytes(ncls) + (((ncls + 1) * TSB + 15) & ~15); }
  static __device__ __forceinline__ void load(uint32_t a, uint32_t b, uint32_t (&t)[8]) {
    if (NT == 1) {
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t[0]) : "r"(a));
    } else if (NT == 2) {
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(t[0]), "=r"(t[1]) : "r"(a));
    } else {
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]) : "r"(a));
      if (NT == 5) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t[4]) : "r"(b));
      if (NT == 6) asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(t[4]), "=r"(t[5]) : "r"(b));
      if (NT == 7)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]) : "r"(b));
    }
  }
  static __device__ __forceinline__ void store(uint32_t a, uint32_t b, const uint32_t (&t)[8]) {
    if (NT == 1) {
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(t[0]) : "memory");
    } else if (NT == 2) {
      asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(t[0]), "r"(t[1]) : "memory");
    } else {
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]) : "memory");
      if (NT == 5) asm volatile("st.shared.u32 [%0], %1;" ::"r"(b), "r"(t[4]) : "memory");
      if (NT == 6) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(b), "r"(t[4]), "r"(t[5]) : "memory");
      if (NT == 7)
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(b), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]) : "memory");
    }
  }
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// Shared-memory addresses of a gene's class data (all per warp).
struct ClassRef {
  uint32_t rec_s;    // ncls x 8 u16 ptab indices (0 = incompatible)
  uint32_t meta_s;   // (ncls + 1) x u32: bits 0-7 first compatible isoform, bit 8 uniform-code class
  uint32_t thr_s;    // plane A: (ncls + 1) x Thr<K>::TSA bytes of ~t_0..3, row ncls = null class (all 0: no test ever true)
  uint32_t thrb_s;   // plane B: (ncls + 1) x Thr<K>::TSB bytes of ~t_4..6
  uint32_t l_s;      // 8 ints: L_k of the paired-end read score, lp = L_k - (code - 1) (miso_paired.c:409-410)
  int ncls;
};

// ---- thresholds of every class for the current psi ---------------------------
// Lane c handles class c (c += 32).  Returns true (warp-uniform) when some threshold is
// too close to an integer to be trusted -- the caller then runs the literal rule.
template <int K>
__device__ __forceinline__ bool thr_update(const ClassRef &cr, uint32_t ptab_s, const double (&psi)[K]) {
  constexpr int NT = Thr<K>::NT;
  const int lane = threadIdx.x & 31;
  bool bad = false;
  __syncwarp();          // the previous pass's threshold loads (other lanes) before these stores
  for (int c = lane; c < cr.ncls; c += 32) {
    uint4 rec;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w) : "r"(cr.rec_s + 16u * c));
    const uint32_t rw[4] = {rec.x, rec.y, rec.z, rec.w};
    const int first = (int) (lds_u32(cr.meta_s + 4u * c) & 0xffu);
    double S = 0.0, C[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const uint32_t idx = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xffffu);
      S = S + psi[k] * lds_f64(ptab_s + idx * 8u);      // CUMSUM, miso_paired.c:11-22
      C[k] = S;
    }
    const double inv = d_div(4294967296.0, S);
    uint32_t t[8];                 // stored complemented: w > t  <=>  w + ~t carries out of 32 bits
#pragma unroll
    for (int k = 0; k < 8; k++) t[k] = 0u;
#pragma unroll
    for (int k = 0; k < NT; k++) {
      const double tau = C[k] * inv - 0.5;
      const uint32_t tk = __double2uint_rz(tau);         // saturating
      const double fr = tau - (double) tk;
      const bool good = fr > kThrMargin && fr < 1.0 - kThrMargin;
      // isoforms before the first compatible one: C_k is an exact 0 < rnd, the test is
      // always true; those reads are counted by GeneDesc.g_always, the row says "never"
      if (k >= first) {
        t[k] = ~tk;
        bad = bad || !good;
      }
    }
    Thr<K>::store(cr.thr_s + (uint32_t) (Thr<K>::TSA * c), cr.thrb_s + (uint32_t) (Thr<K>::TSB * c), t);
  }
  bad = __any_sync(0xffffffffu, bad);
  __syncwarp();
  return bad;
}

// -log(lp) of miso_paired.c:409-411 from the plan-wide table (lp is a small positive
// integer: an isoform length minus a fragment length); anything else is computed.
__device__ __forceinline__ double neg_log_lp(int lp, const double *__restrict__ neglog, int n_neglog) {
  if (lp > 0 && lp < n_neglog) return __ldg(neglog + lp);
  return -d_log((double) lp);
}

// ---- one reassignment pass ---------------------------------------------------------
// Lane handles Philox block Q0+T (uniform indices 4(Q0+T)..+3), T = lane + 32*step,
// i.e. ranks 4T-o .. 4T-o+3 with o = n_u & 3 (the stream is sequential and the accept
// draw is conditional, miso.c:870, so a pass starts at an arbitrary phase).  The id
// row carries 3 null ids in front and null ids behind: phantom ranks count nothing.
//   MODE 0: counts only.
//   MODE 1, 2: + the read score of the chosen isoforms (paired-end, miso_paired.c:157-163);
//           runs before an iteration that records a sample (the MH ratio does not need
//           the read score, the recorded log score does).
template <int K, int MODE, bool SMEM, bool WIDE, class KEY>
__device__ __forceinline__ void class_pass_body(typename TileMem<SMEM>::addr_t rows, int ucode_off, const ClassRef &cr,
                                                uint32_t ptab_s, unsigned long long n_u, int R2, uint32_t gene,
                                                uint32_t chain, const KEY &key,
                                                const int *__restrict__ g_always,
                                                const double *__restrict__ neglog, int n_neglog,
                                                int (&cnt)[K], double &rp) {
  using TM = TileMem<SMEM>;
  constexpr int NT = Thr<K>::NT, TSA = Thr<K>::TSA, TSB = Thr<K>::TSB;
  const int lane = threadIdx.x & 31;
  const int o = (int) (n_u & 3ull);
  const uint32_t Q0 = (uint32_t) (n_u >> 2);
  const int nsteps = (R2 + o + 127) >> 7;
  const uint32_t sel = 0x3210u + 0x1111u * (uint32_t) (3 - o);
  uint32_t G[NT];
#pragma unroll
  for (int k = 0; k < NT; k++) G[k] = 0;
  typename TM::addr_t a = rows + 4 * lane;
  typename TM::addr_t ua = rows + ucode_off + (WIDE ? 8 : 4) * lane;
  const int hs = 3 - o;                               // 16-bit codes: halfwords into the 8-halfword window
  const bool hb = (hs >> 1) != 0;
  const uint32_t hsh = 16u * (uint32_t) (hs & 1);
  const uint32_t thr_s = cr.thr_s, thrb_s = cr.thrb_s;
  const uint32_t ncls = (uint32_t) cr.ncls;
  double rp_lane = 0.0;
  uint32_t tot = 0;                                   // MODE 1: sum of the G_k so far
#pragma unroll (MODE == 0 ? (K >= kUnroll1FromK ? 1 : kUnrollCount) : kUnrollScore)
  for (int s = 0; s < nsteps; s++) {
    uint32_t x[4];
    philox4x32(Q0 + (uint32_t) (lane + 32 * s), 0u, gene, chain, key, x);
    const uint32_t ids = __byte_perm(TM::ld(a), TM::ld(a + 4), sel);
    a += 128;
    uint32_t uc01 = 0, uc23 = 0;                      // MODE 1: the 4 reads' own codes
    if (MODE >= 1) {
      if (!WIDE) {
        uc01 = __byte_perm(TM::ld(ua), TM::ld(ua + 4), sel);
      } else {
        const uint4 w = TM::ld4(ua);
        const uint32_t wa = hb ? w.y : w.x, wb = hb ? w.z : w.y, wc = hb ? w.w : w.z;
        uc01 = __funnelshift_r(wa, wb, hsh);
        uc23 = __funnelshift_r(wb, wc, hsh);
      }
      ua += WIDE ? 256 : 128;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t id = __byte_perm(ids, 0u, 0x4440u | (uint32_t) i);
      uint32_t nt[8];
      Thr<K>::load(thr_s + id * (uint32_t) TSA, thrb_s + id * (uint32_t) TSB, nt);
#pragma unroll
      for (int k = 0; k < NT; k++)      // G_k += (w > t_k): carry out of w + ~t_k, added with carry
        asm("{\n\t.reg .u32 j;\n\tadd.cc.u32 j, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(G[k]) : "r"(x[i]), "r"(nt[k]));
      if (MODE >= 1) {
        uint32_t now = 0;
#pragma unroll
        for (int k = 0; k < NT; k++) now += G[k];
        const uint32_t meta = lds_u32(cr.meta_s + 4u * id);
        const uint32_t chosen = (meta & 0xffu) + (now - tot);       // first compatible + tests passed
        tot = now;
        uint32_t cc;
        if (!WIDE) cc = __byte_perm(uc01, 0u, 0x4440u | (uint32_t) i);
        else cc = __byte_perm(i < 2 ? uc01 : uc23, 0u, (i & 1) ? 0x4432u : 0x4410u);
        if (!(meta & 0x100u)) cc = lds_u16(cr.rec_s + 16u * id + 2u * chosen);   // not a uniform-code class
        const int lp = (int) lds_u32(cr.l_s + 4u * chosen) - ((int) cc - 1);
        // isoscores, miso_paired.c:409-411; MODE 2: the host checked that every lp this gene can
        // produce is inside the table (GeneDesc.lp_safe), no range test and no log() call in the loop
        const double nl = MODE == 2 ? __ldg(neglog + lp) : neg_log_lp(lp, neglog, n_neglog);
        const double sc = nl + lds_f64(ptab_s + cc * 8u);
        if (id != ncls) rp_lane += sc;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NT; k++) G[k] = __reduce_add_sync(0xffffffffu, G[k]) + (uint32_t) g_always[k];
  cnt[0] = R2 - (int) G[0];
#pragma unroll
  for (int k = 1; k < NT; k++) cnt[k] = (int) (G[k - 1] - G[k]);
  cnt[K - 1] = (int) G[NT - 1];
  if (MODE >= 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rp_lane += __shfl_xor_sync(0xffffffffu, rp_lane, d);
    rp = rp_lane;
  }
}

#ifdef MISOB200_PASS_NOINLINE
#define MISOB200_PASS_INLINE __noinline__
#else
#define MISOB200_PASS_INLINE __forceinline__
#endif
template <int K, bool SMEM, class KEY>
__device__ MISOB200_PASS_INLINE void class_pass(typename TileMem<SMEM>::addr_t rows, const ClassRef &cr,
                                           unsigned long long n_u, int R2, uint32_t gene, uint32_t chain,
                                           const KEY &key, const int *__restrict__ g_always, int (&cnt)[K]) {
  double unused;
  class_pass_body<K, 0, SMEM, false>(rows, 0, cr, 0u, n_u, R2, gene, chain, key, g_always, nullptr, 0, cnt, unused);
}

// MODE 1 out of line: one pass in `lag` runs it
template <int K, bool SMEM, bool WIDE, class KEY>
__device__ __noinline__ void class_pass_rp(typename TileMem<SMEM>::addr_t rows, int ucode_off, const ClassRef &cr,
                                           uint32_t ptab_s, unsigned long long n_u, int R2, uint32_t gene,
                                           uint32_t chain, const KEY &key, const int *__restrict__ g_always,
                                           const double *__restrict__ neglog, int n_neglog, bool lp_safe, int *cnt_k,
                                           double *rp) {
  int cnt[K];
  if (lp_safe)
    class_pass_body<K, 2, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, n_u, R2, gene, chain, key, g_always, neglog, n_neglog,
                                      cnt, *rp);
  else
    class_pass_body<K, 1, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, n_u, R2, gene, chain, key, g_always, neglog, n_neglog,
                                      cnt, *rp);
  const int lane = threadIdx.x & 31;
  int c = 0;
#pragma unroll
  for (int k = 0; k < K; k++) c = ((lane & 7) == k) ? cnt[k] : c;
  *cnt_k = c;
}

// ---- the literal rule of miso.c:59-83 on a class tile --------------------------------
// Codes are rebuilt from the class record (uniform-code classes: the read's own code
// wherever the record is non-zero).  Used for the final pass of chain 0 (emits the
// per-read assignment, miso.c:943-946) and for passes thr_update declined.
template <int K, bool SMEM, bool WIDE, class KEY>
__device__ __noinline__ void class_literal(typename TileMem<SMEM>::addr_t rows, int ucode_off, const ClassRef &cr,
                                           uint32_t ptab_s, double psi_k, unsigned long long n_u, int R2,
                                           uint32_t gene, uint32_t chain, const KEY &key, int paired,
                                           const int *__restrict__ L, const double *__restrict__ neglog,
                                           int n_neglog, int *cnt_k, double *rp, uint8_t *__restrict__ ass_out) {
  using TM = TileMem<SMEM>;
  const int lane = threadIdx.x & 31;
  const int o = (int) (n_u & 3ull);
  const uint32_t Q0 = (uint32_t) (n_u >> 2);
  const int nsteps = (R2 + o + 127) >> 7;
  double psi[K];
#pragma unroll
  for (int k = 0; k < K; k++) psi[k] = shfl_d(psi_k, k);
  int n[K];
#pragma unroll
  for (int k = 0; k < K; k++) n[k] = 0;
  double rp_lane = 0.0;
  for (int s = 0; s < nsteps; s++) {
    const int T = lane + 32 * s;
    uint32_t x[4];
    philox4x32(Q0 + (uint32_t) T, 0u, gene, chain, key, x);
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const int rank = 4 * T - o + i;
      if (rank < 0 || rank >= R2) continue;
      const int el = kTilePadFront + rank;
      const uint32_t id = (TM::ld(rows + (el & ~3)) >> (8 * (el & 3))) & 0xffu;
      const uint32_t meta = lds_u32(cr.meta_s + 4u * id);
      uint32_t uc = 0;
      if (meta & 0x100u) {
        const int byte = WIDE ? 2 * el : el;
        const uint32_t w = TM::ld(rows + ucode_off + (byte & ~3));
        uc = WIDE ? (w >> (8 * (byte & 2))) & 0xffffu : (w >> (8 * (byte & 3))) & 0xffu;
      }
      const uint32_t xi = i == 0 ? x[0] : i == 1 ? x[1] : i == 2 ? x[2] : x[3];
      double S = 0.0, C[K];
      uint32_t code[K];
      int nv = 0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t idx = lds_u16(cr.rec_s + 16u * id + 2u * k);
        code[k] = idx == 0u ? 0u : ((meta & 0x100u) ? uc : idx);
        nv += idx != 0u;
        S = S + psi[k] * lds_f64(ptab_s + code[k] * 8u);
        C[k] = S;
      }
      const bool two = nv == 2;
      const double rnd = uniform_from_word(xi) * S;
      int chosen = -1;
      uint32_t cc = 0;
#pragma unroll
      for (int k = K - 1; k >= 0; k--) {
        const bool valid = code[k] != 0u;
        const bool hit = two ? (rnd < C[k]) : (rnd <= C[k]);     // miso.c:71 / :78
        if (valid && (hit || chosen < 0)) { chosen = k; cc = code[k]; }
      }
#pragma unroll
      for (int k = 0; k < K; k++) n[k] += (chosen == k);
      if (chosen >= 0 && paired) {
        const int lp = __ldg(L + chosen) - ((int) cc - 1);
        rp_lane += neg_log_lp(lp, neglog, n_neglog) + lds_f64(ptab_s + cc * 8u);
      }
      if (ass_out) ass_out[rank] = (uint8_t) chosen;
    }
  }
  int mine = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int t = __reduce_add_sync(0xffffffffu, n[k]);
    if ((lane & 7) == k) mine = t;        // member k of every lane group
  }
  *cnt_k = mine;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) rp_lane += __shfl_xor_sync(0xffffffffu, rp_lane, d);
  *rp = rp_lane;
}

}  // namespace misob200
