// miso_b200/csrc/dense_pass.cuh -- reassignment passes over DENSE tiles (format 0):
// one code row per isoform + a flag row, the weights psi_k * ptab[code] evaluated
// per read in fp64 exactly as the reference does.  This is the general path
// (any number of weight classes); class_pass.cuh is the fast path.
#pragma once
#include "philox.cuh"
#include "plan.hpp"
#include "tile_mem.cuh"

#ifndef MISOB200_READ_UNROLL
#define MISOB200_READ_UNROLL 2   /* measured: 2 beats 4 (I-cache) and 1 (ILP), profiles/README.md */
#endif

namespace misob200 {

constexpr int kReadUnroll = MISOB200_READ_UNROLL;   // reads of a lane's step unrolled together

// ---- one reassignment pass ----------------------------------------------------
// src/miso.c:30-91 / src/miso_paired.c:24-86 for the R2 reads that draw.
// Lane handles Philox block Q0+T (uniform indices 4(Q0+T)..+3) for
// T = lane + 32*step, i.e. ranks 4T-o .. 4T-o+3 with o = n_u & 3: the stream is
// sequential (the accept draw is conditional, miso.c:870), so a pass starts at
// an arbitrary phase o.  Rows carry 3 zero bytes in front and zero padding
// behind; a zero code means "incompatible" (weight psi_k * ptab[0] = 0).
//
// Choice rule.  With C_k the running sum of psi_k * p_k over ALL isoforms
// (incompatible ones add an exact 0.0, so C_k is the reference's cumsum at the
// last compatible isoform <= k) and rnd = u * C_{K-1}:
//   >= 3 compatible: first w with !(rnd > cumsum[w])      (miso.c:78)
//      2 compatible: rnd <  cumsum[0] ? first : second    (miso.c:71)
// Both are "the number of k < K-1 whose test says go on": rnd > C_k, resp.
// rnd >= C_k, and rnd >= C  <=>  nextup(rnd) > C for rnd >= 0 -- one integer add
// on the bit pattern per read.  (Leading incompatible isoforms have C_k = 0 <
// rnd and are skipped, the others repeat their predecessor's verdict.)  So the
// pass only keeps G_k = #{reads with test_k true}; the per-isoform counts the MH
// ratio needs are n_0 = R2 - G_0, n_k = G_{k-1} - G_k, n_{K-1} = G_{K-2}.
// Phantom ranks (padding) have rnd = 0 and C_k = 0: no test is true.
// The argument needs 0 < rnd < C_{K-1}, true whenever C_{K-1} is a normal number
// (u is in [2^-33, 1 - 2^-33]).  Every drawing read has a compatible isoform, so
// C_{K-1} >= min_k psi_k * min nonzero ptab: the caller checks that bound once per
// pass (>= 1e-290, which also rules out a negative psi_{K-1} = 1 - sum) and
// otherwise runs reassign_literal below instead.
//   MODE 0: counts only.  MODE 1: + read score of the chosen isoform
//   (miso_paired.c:157-163), needed when the next iteration records.
template <int K, int MODE, bool SMEM, bool WIDE, class KEY>
__device__ __forceinline__ void reassign_pass(typename TileMem<SMEM>::addr_t rows, int row_bytes, int flag_off,
                                              uint32_t ptab_s, const double (&psi)[K],
                                              unsigned long long n_u, int R2, uint32_t gene,
                                              uint32_t chain, const KEY &key, int paired,
                                              const int *__restrict__ L, int (&cnt)[K], double &rp) {
  using TM = TileMem<SMEM>;
  const int lane = threadIdx.x & 31;
  const int o = (int) (n_u & 3ull);
  const uint32_t Q0 = (uint32_t) (n_u >> 2);
  const int nsteps = (R2 + o + 127) >> 7;
  const uint32_t sel = 0x3210u + 0x1111u * (uint32_t) (3 - o);
  int G[K];
#pragma unroll
  for (int k = 0; k < K; k++) G[k] = 0;
  double rp_lane = 0.0;
  // this lane's window of row 0: elements 4*lane .. 4*lane+7 (4 codes after the phase shift)
  typename TM::addr_t a = rows + (WIDE ? 8 : 4) * lane;
  typename TM::addr_t fa = rows + flag_off + 4 * lane;
  // 16-bit codes: the lane's 4 codes start (3 - o) halfwords into its 8-halfword window
  const int hs = 3 - o;
  const bool hb = (hs >> 1) != 0;
  const uint32_t hsh = 16u * (uint32_t) (hs & 1);

  for (int s = 0; s < nsteps; s++) {
    const int T = lane + 32 * s;
    uint32_t x[4];
    philox4x32(Q0 + (uint32_t) T, 0u, gene, chain, key, x);
    uint32_t cw[K + 1], cx[WIDE ? K : 1];
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (!WIDE) {
        const uint32_t w0 = TM::ld(a + k * row_bytes), w1 = TM::ld(a + k * row_bytes + 4);
        cw[k] = __byte_perm(w0, w1, sel);
      } else {
        const uint4 w = TM::ld4(a + k * row_bytes);
        const uint32_t wa = hb ? w.y : w.x, wb = hb ? w.z : w.y, wc = hb ? w.w : w.z;
        cw[k] = __funnelshift_r(wa, wb, hsh);      // codes of reads 0,1
        cx[k] = __funnelshift_r(wb, wc, hsh);      // codes of reads 2,3
      }
    }
    cw[K] = __byte_perm(TM::ld(fa), TM::ld(fa + 4), sel);
    a += WIDE ? 256 : 128;
    fa += 128;
#pragma unroll (kReadUnroll)
    for (int i = 0; i < 4; i++) {
      // flag byte: 1 = exactly two compatible isoforms (compare with nextup(rnd)), else 0
      const uint32_t two = __byte_perm(cw[K], 0u, 0x4440u | (uint32_t) i);
      double S = 0.0, C[K];
      uint32_t code[K];
#pragma unroll
      for (int k = 0; k < K; k++) {
        if (!WIDE) code[k] = __byte_perm(cw[k], 0u, 0x4440u | (uint32_t) i);
        else code[k] = __byte_perm(i < 2 ? cw[k] : cx[k], 0u, (i & 1) ? 0x4432u : 0x4410u);
        S = S + psi[k] * lds_f64(ptab_s + code[k] * 8u);     // CUMSUM, miso_paired.c:11-22
        C[k] = S;
      }
      const double rnd = uniform_from_word(MISOB200_READ_UNROLL == 4 ? x[i] : (i == 0 ? x[0] : i == 1 ? x[1] : i == 2 ? x[2] : x[3])) * S;          // miso.c:70,76
      const double rc = __longlong_as_double(__double_as_longlong(rnd) + (long long) two);
      int chosen = 0;
#pragma unroll
      for (int k = 0; k < K - 1; k++) {
        const bool go_on = rc > C[k];
        G[k] += go_on;
        if (MODE == 1) chosen += go_on;
      }
      if (MODE == 1) {
        const int rank = 4 * T - o + i;
        if (rank >= 0 && rank < R2 && paired) {
          uint32_t cc = code[0];
#pragma unroll
          for (int k = 1; k < K; k++)
            if (chosen == k) cc = code[k];
          const int Lc = __ldg(L + chosen);
          const double lp = (double) (Lc - ((int) cc - 1));
          rp_lane += -log(lp) + lds_f64(ptab_s + cc * 8u);   // isoscores, miso_paired.c:409-411
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K - 1; k++) G[k] = __reduce_add_sync(0xffffffffu, G[k]);
  cnt[0] = R2 - G[0];
#pragma unroll
  for (int k = 1; k < K - 1; k++) cnt[k] = G[k - 1] - G[k];
  cnt[K - 1] = G[K - 2];
  if (MODE == 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rp_lane += __shfl_xor_sync(0xffffffffu, rp_lane, d);
    rp = rp_lane;
  }
}

// The literal rule of miso.c:59-83, read by read, with the compatibility tests
// spelled out.  Used for the final pass of chain 0 (which has to emit the
// per-read assignment, miso.c:943-946) and for passes the fast rule declined.
// Not inlined and not unrolled over reads: it runs once or twice per chain.
// psi_k is lane k's psi; returns lane k's count in cnt_k and the read score.
template <int K, bool SMEM, bool WIDE, class KEY>
__device__ __noinline__ void reassign_literal(typename TileMem<SMEM>::addr_t rows, int row_bytes, int flag_off,
                                              uint32_t ptab_s, double psi_k, unsigned long long n_u,
                                              int R2, uint32_t gene, uint32_t chain, const KEY &key,
                                              int paired, const int *__restrict__ L, int *cnt_k,
                                              double *rp, uint8_t *__restrict__ ass_out) {
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
      auto code_at = [&](int k) -> uint32_t {
        const int byte = WIDE ? 2 * el : el;
        const uint32_t w = TM::ld(rows + k * row_bytes + (byte & ~3));
        return WIDE ? (w >> (8 * (byte & 2))) & 0xffffu : (w >> (8 * (byte & 3))) & 0xffu;
      };
      const bool two = ((TM::ld(rows + flag_off + (el & ~3)) >> (8 * (el & 3))) & 0xffu) == 1u;
      const uint32_t xi = i == 0 ? x[0] : i == 1 ? x[1] : i == 2 ? x[2] : x[3];
      double S = 0.0, C[K];
      uint32_t code[K];
#pragma unroll
      for (int k = 0; k < K; k++) {
        code[k] = code_at(k);
        S = S + psi[k] * lds_f64(ptab_s + code[k] * 8u);
        C[k] = S;
      }
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
        const double lp = (double) (__ldg(L + chosen) - ((int) cc - 1));
        rp_lane += -log(lp) + lds_f64(ptab_s + cc * 8u);
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
