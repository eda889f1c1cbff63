// miso_b200/csrc/chain_kernel.cuh -- the per-gene MCMC chain on one warp (sm_100a).
//
// One warp owns one gene-chain for its whole life: burn-in + sampling loop of
// /root/reference/pysplicing/src/miso.c:827-947 (single-end) and
// src/miso_paired.c:431-538 (paired-end), appendix A of SURVEY.md.
//
// What lives where
//   shared memory  : the gene's tile -- K code rows + 1 flag row, one byte per
//                    read that draws, in the reference's draw order -- brought
//                    in once per gene-chain by a TMA bulk copy
//                    (cp.async.bulk + mbarrier, SASS UBLKCP); the insert-length
//                    probability table ptab (shared by the CTA's warps).
//   registers      : psi (replicated for the read pass), everything else
//                    lane-distributed: lane k holds alpha_k, psi_k, the
//                    normalised log psi_k, prior and length terms of isoform k;
//                    Philox counters/keys; per-lane assignment counts.
//   warp shuffles  : the K-wide serial sums / max of the reference
//                    (score_iso, ldirichlet, logit_inv, mvplogisnorm) are done
//                    in the reference's order by broadcasting lane i's term.
//   tensor cores   : unused on purpose -- no dense contraction on this path.
//
// Decision parity: every compare the reference makes in fp64
// (rand*sumpsi vs cumsum, U vs acceptP) is made here with the same operands
// in the same operation order (this file is compiled with -fmad=false so a
// multiply feeding an add is not contracted).  The MH ratio uses per-isoform
// assignment counts (sum_k n_k * logpsi_k) instead of a serial sum over reads:
// same value up to fp64 rounding, see DESIGN.md "parity contract".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"
#include "plan.hpp"

#ifndef MISOB200_READ_UNROLL
#define MISOB200_READ_UNROLL 2   /* measured: 2 beats 4 (I-cache) and 1 (ILP), profiles/README.md */
#endif

namespace misob200 {

constexpr int kReadUnroll = MISOB200_READ_UNROLL;   // reads of a lane's step unrolled together

struct ChainParams {
  const GeneDesc *desc;
  const int *items;          // gene indices of this K bucket, longest first
  int n_genes;               // entries in items
  int n_chains;
  const uint8_t *tiles;
  const double *ptab;
  int n_ptab;
  double ptab_min;           // smallest non-zero entry of ptab
  double *samples;
  double *loglik;
  uint8_t *drawn;            // final chain-0 assignment of the reads that draw, draw order
  int *accrej;               // [gene][chain][2]
  unsigned *queue;           // work counter
  int n_iters, burn_in, lag, start;
  PhiloxKey key;
  int slot_bytes;            // shared-memory bytes per warp for a tile (0: stream tiles from L2)
};

__device__ __forceinline__ double shfl_d(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier ---------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- explicit address-space loads ------------------------------------------------
// The tile normally sits in shared memory (32-bit shared addresses, LDS); genes
// whose tile does not fit in a slot are streamed from global/L2 instead.
template <bool SMEM> struct TileMem;
template <> struct TileMem<true> {
  using addr_t = uint32_t;
  static __device__ __forceinline__ addr_t base(const void *p) { return smem_u32(p); }
  static __device__ __forceinline__ uint32_t ld(addr_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
  }
  // 16 bytes from an 8-byte aligned address (two 8-byte loads)
  static __device__ __forceinline__ uint4 ld4(addr_t a) {
    uint4 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+8];" : "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
  }
};
template <> struct TileMem<false> {
  using addr_t = const unsigned char *;
  static __device__ __forceinline__ addr_t base(const void *p) { return static_cast<addr_t>(p); }
  static __device__ __forceinline__ uint32_t ld(addr_t a) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(a));
    return v;
  }
  static __device__ __forceinline__ uint4 ld4(addr_t a) {
    uint4 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(a));
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2+8];" : "=r"(v.z), "=r"(v.w) : "l"(a));
    return v;
  }
};
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

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
template <int K, int MODE, bool SMEM, bool WIDE>
__device__ __forceinline__ void reassign_pass(typename TileMem<SMEM>::addr_t rows, int row_bytes, int flag_off,
                                              uint32_t ptab_s, const double (&psi)[K],
                                              unsigned long long n_u, int R2, uint32_t gene,
                                              uint32_t chain, const PhiloxKey &key, int paired,
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
    philox4x32_10(Q0 + (uint32_t) T, 0u, gene, chain, key, x);
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
template <int K, bool SMEM, bool WIDE>
__device__ __noinline__ void reassign_literal(typename TileMem<SMEM>::addr_t rows, int row_bytes, int flag_off,
                                              uint32_t ptab_s, double psi_k, unsigned long long n_u,
                                              int R2, uint32_t gene, uint32_t chain, const PhiloxKey &key,
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
    philox4x32_10(Q0 + (uint32_t) T, 0u, gene, chain, key, x);
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
    if (lane == k) mine = t;
  }
  *cnt_k = mine;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) rp_lane += __shfl_xor_sync(0xffffffffu, rp_lane, d);
  *rp = rp_lane;
}

// Everything derived from a candidate alpha (lane i < K-1 holds alpha_i).
struct Derived {
  double psi;      // lane k < K     : psi_k             (logit_inv, miso.c:219-241,467)
  double lp;       // lane k < K     : normalised log psi (score_iso, miso.c:136-149)
  double q;        // lane i < K-1   : log(psi_i / psi_rest) (mvplogisnorm, miso.c:113)
  double dir;      // uniform        : ldirichlet (miso.c:165-182)
  double prod;     // uniform        : 1 / prod(theta) / ltheta (miso.c:110)
};

template <int K>
__device__ __forceinline__ Derived derive(double alpha, double offset_k, double hyper_m1_k,
                                          double lg_sum, double lg_each) {
  constexpr int len = K - 1;
  const int lane = threadIdx.x & 31;
  Derived r;
  const double e = exp(alpha);
  double sumexp = 0.0;
#pragma unroll
  for (int i = 0; i < len; i++) sumexp = sumexp + shfl_d(e, i);
  sumexp = sumexp + 1.0;
  double psi = e / sumexp;
  double sumpsi = 0.0;
#pragma unroll
  for (int i = 0; i < len; i++) sumpsi = sumpsi + shfl_d(psi, i);
  if (lane == len) psi = 1 - sumpsi;
  r.psi = psi;

  const double lg = log(psi);
  const double t = lg + offset_k;
  double mx = shfl_d(t, 0);
#pragma unroll
  for (int i = 1; i < K; i++) {
    const double v = shfl_d(t, i);
    if (v > mx) mx = v;
  }
  const double ex = exp(t - mx);
  double sum = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) sum = sum + shfl_d(ex, i);
  sum = log(sum) + mx;
  r.lp = t - sum;

  const double term = hyper_m1_k * lg;
  double dir = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) dir = dir + shfl_d(term, i);
  dir = dir + lg_sum;
  dir = dir - lg_each;
  r.dir = dir;

  double lth = 1.0, prod = 1.0;
#pragma unroll
  for (int i = 0; i < len; i++) {
    const double at = shfl_d(psi, i);
    lth = lth - at;
    prod = prod * at;
  }
  r.prod = 1.0 / prod / lth;
  r.q = log(psi / lth);
  return r;
}

// sum_k n_k * v_k in isoform order, skipping isoforms nothing is assigned to
template <int K>
__device__ __forceinline__ double count_dot(int cnt_k, double v_k) {
  const double a = cnt_k ? (double) cnt_k * v_k : 0.0;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < K; i++) s = s + shfl_d(a, i);
  return s;
}

template <int K, bool SMEM, bool WIDE>
__device__ void run_chain(const ChainParams &P, const GeneDesc &d, int gene_index, int chain,
                          typename TileMem<SMEM>::addr_t rows, uint32_t ptab_s) {
  constexpr int len = K - 1;
  const int lane = threadIdx.x & 31;
  const int kk = lane < K ? lane : K - 1;
  const double offset_k = d.offset[kk];
  const double hyper_m1_k = d.hyper_m1[kk];
  const double rs_se_k = d.rs_se[kk];
  const int nfix_k = d.n_fixed[kk];
  const double lg_sum = d.lg_sum, lg_each = d.lg_each;
  const double sigma = d.sigma, sd = d.sd, covar = d.covar_const;
  const int R2 = d.R2, row_bytes = d.row_bytes, flag_off = d.flag_off, paired = d.paired;
  const uint32_t gid = d.gene_id;
  const PhiloxKey &key = P.key;
  const int *L = d.L;

  unsigned long long n_u = 0;
  // ---- start state, splicing_drift_proposal_init (miso.c:330-447) ----------
  double alpha;
  if (P.start == MISOB200_START_AUTO) {
    if (K == 2) { n_u = 1; alpha = 0.0; }     // one uniform drawn and discarded (miso.c:365)
    else alpha = 1.0 / (K - 1);
  } else {
    alpha = 0.0;
  }

  // normals are produced 32 at a time (lane l holds normal zbase + l)
  uint32_t zbase = 0xffffff00u;           // forces a refill on first use
  double zbuf = 0.0;
  auto next_normals = [&](uint32_t first) -> double {   // lane i < len gets normal first + i
    if (first - zbase + (uint32_t) len > 32u) {
      zbase = first;
      zbuf = stream_normal(first + (uint32_t) lane, gid, (uint32_t) chain, key);
    }
    return shfl_d(zbuf, (int) ((first - zbase + (uint32_t) lane) & 31u));
  };

  Derived cur;
  cur.psi = cur.lp = cur.q = cur.dir = cur.prod = 0.0;
  double psi_r[K];
  int cnt[K];
  int cnt_k;               // lane k: reads currently assigned to isoform k
  double rp_drawn = 0.0;
  int lagc = 0, n_rec = 0, acc = 0, rej = 0;
  const int S_total = (P.n_iters - P.burn_in) / P.lag;
  uint8_t *ass_out = (chain == 0) ? P.drawn + d.drawn_off : nullptr;

  auto do_pass = [&](int m_next) {
    const bool last = (m_next >= P.n_iters);
    const bool rec_next = d.rp_always || (paired && m_next >= P.burn_in && lagc == P.lag - 1);
    bool ok = false;
    if (!(last && ass_out)) {
      double pmin = 1.0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        psi_r[k] = shfl_d(cur.psi, k);
        pmin = psi_r[k] < pmin ? psi_r[k] : pmin;       // a NaN psi never lowers pmin ...
        ok = ok || !(psi_r[k] == psi_r[k]);             // ... so flag it here
      }
      ok = !ok && (pmin * P.ptab_min >= 1e-290);        // fast rule valid for every read of this pass
    }
    if (ok) {
      if (rec_next && !last)
        reassign_pass<K, 1, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, psi_r, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                  cnt, rp_drawn);
      else
        reassign_pass<K, 0, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, psi_r, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                  cnt, rp_drawn);
      int c = 0;
#pragma unroll
      for (int k = 0; k < K; k++) c = (lane == k) ? cnt[k] : c;
      cnt_k = c + nfix_k;
    }
    if (!ok) {   // final pass of chain 0, or a read whose weights underflow: literal rule
      int c = 0;
      reassign_literal<K, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, cur.psi, n_u, R2, gid, (uint32_t) chain, key, paired,
                                L, &c, &rp_drawn, last ? ass_out : nullptr);
      cnt_k = c + nfix_k;
    }
    n_u += (unsigned long long) R2;
    return rec_next || !ok;
  };

  bool have_rp = false;

  // m == -1 is the start-up proposal, adopted unconditionally (miso.c:834), followed
  // by the initial assignment (miso.c:840-843); m >= 0 are the iterations proper.
  for (int m = -1; m < P.n_iters; m++) {
    // ---- propose (miso.c:851) --------------------------------------------
    const double alphaN = alpha + sd * next_normals((uint32_t) (m + 1) * (uint32_t) len);
    const Derived nw = derive<K>(alphaN, offset_k, hyper_m1_k, lg_sum, lg_each);
    if (m < 0) {
      alpha = alphaN; cur = nw;
    } else {
    // ---- joint scores (miso.c:524-529) -----------------------------------
    double rp;
    if (!paired) rp = count_dot<K>(cnt_k, rs_se_k);          // sum_r isoscores[ass_r], miso.c:267-271
    else rp = have_rp ? d.rp_fixed + rp_drawn : 0.0;        // cancels in the ratio when not recorded
    const double ppJS = rp + count_dot<K>(cnt_k, nw.lp) + nw.dir;
    const double pcJS = rp + count_dot<K>(cnt_k, cur.lp) + cur.dir;

    // ---- proposal densities (miso.c:531-534, :97-122) --------------------
    const double t1 = cur.q - alphaN;                        // theta = psi,    mu = alphaNew
    const double t2 = nw.q - alpha;                          // theta = psiNew, mu = alpha
    const double e1 = (-0.5) * t1 * t1 / sigma, e2 = (-0.5) * t2 * t2 / sigma;
    double ep1 = 0.0, ep2 = 0.0;
#pragma unroll
    for (int i = 0; i < len; i++) { ep1 = ep1 + shfl_d(e1, i); ep2 = ep2 + shfl_d(e2, i); }
    const double xe = exp(lane == 0 ? ep1 : ep2);
    const double pdf = covar * (lane == 0 ? cur.prod : nw.prod) * xe;
    const double sc = log(pdf);
    const double ptoCS = shfl_d(sc, 0), ctoPS = shfl_d(sc, 1);

    const double acceptP = (m > 0) ? exp(ppJS + ptoCS - (pcJS + ctoPS)) : exp(ppJS - pcJS);

    // ---- accept (miso.c:869-880): the uniform is drawn only if acceptP < 1 --
    bool accept = acceptP >= 1;
    if (!accept) {
      const double u = stream_uniform(n_u, gid, (uint32_t) chain, key);
      n_u++;
      accept = u < acceptP;
    }
    double cJS = pcJS;
    if (accept) {
      alpha = alphaN; cur = nw; cJS = ppJS; acc++;
    } else {
      rej++;
    }

    // ---- record (miso.c:882-893) ------------------------------------------
    if (m >= P.burn_in) {
      if (lagc == P.lag - 1) {
        if (n_rec < S_total) {
          const long long col = (long long) n_rec * P.n_chains + chain;
          if (lane < K) P.samples[d.sample_off + col * K + lane] = cur.psi;
          if (lane == 0) P.loglik[d.loglik_off + col] = cJS;
        }
        n_rec++;
        lagc = 0;
      } else {
        lagc++;
      }
    }

    }
    // ---- reassign (miso.c:895-898; for m == -1 the initial assignment) ---------
    have_rp = do_pass(m + 1);
  }

  if (lane == 0) {
    int *ar = P.accrej + ((long long) gene_index * P.n_chains + chain) * 2;
    ar[0] = acc; ar[1] = rej;
  }
}

template <int K, int WARPS, bool SMEM, bool WIDE>
__global__ void __launch_bounds__(WARPS * 32, (K <= 6 ? 4 : 3)) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // layout: [ptab | per-warp {mbarrier(16 B), tile slot}]
  double *s_ptab = reinterpret_cast<double *>(smem);
  const int ptab_bytes = (P.n_ptab * 8 + 15) & ~15;
  unsigned char *wbase = smem + ptab_bytes + (size_t) warp * (16 + P.slot_bytes);
  uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *slot = wbase + 16;

  for (int i = threadIdx.x; i < P.n_ptab; i += WARPS * 32) s_ptab[i] = P.ptab[i];
  if (lane == 0 && SMEM) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();

  const int n_items = P.n_genes * P.n_chains;
  uint32_t phase = 0;
  while (true) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(P.queue, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((int) item >= n_items) break;
    const int gi = P.items[item / P.n_chains];
    const int chain = (int) (item % P.n_chains);
    const GeneDesc &d = P.desc[gi];
    const uint32_t tile_bytes = (uint32_t) d.tile_bytes;
    if (SMEM) {
      __syncwarp();
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, tile_bytes);
        tma_bulk_g2s(slot, P.tiles + d.tile_off, tile_bytes, bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      run_chain<K, true, WIDE>(P, d, gi, chain, TileMem<true>::base(slot), smem_u32(s_ptab));
    } else {
      run_chain<K, false, WIDE>(P, d, gi, chain, TileMem<false>::base(P.tiles + d.tile_off), smem_u32(s_ptab));
    }
  }
}

}  // namespace misob200
