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
#include "tile_mem.cuh"

namespace misob200 {

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
  // class format only
  const double *neglog;      // neglog[n] = -log(n), n < n_neglog (read scores, miso_paired.c:409-411)
  int n_neglog;
  int thr_bytes;             // shared-memory bytes per warp for the threshold rows (after the slot)
};

}  // namespace misob200

#include "dense_pass.cuh"
#include "class_pass.cuh"

namespace misob200 {

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

template <int K, bool SMEM, bool WIDE, int FMT>
__device__ void run_chain(const ChainParams &P, const GeneDesc &d, int gene_index, int chain,
                          typename TileMem<SMEM>::addr_t rows, uint32_t ptab_s, const ClassRef &cr) {
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
  const int ucode_off = d.ucode_off;
  int g_always[K];
#pragma unroll
  for (int k = 0; k < K; k++) g_always[k] = FMT == 1 ? d.g_always[k] : 0;
  int thr_state = 0;       // class format: 0 thresholds stale (psi changed), 1 valid, 2 declined for this psi
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
    if (FMT == 0) {
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
    } else {
      if (ok && thr_state == 0) thr_state = thr_update<K>(cr, ptab_s, psi_r) ? 2 : 1;
      ok = ok && thr_state == 1;
      int c = 0;
      if (ok && !(rec_next && !last)) {
        class_pass<K, SMEM>(rows, cr, n_u, R2, gid, (uint32_t) chain, key, g_always, cnt);
#pragma unroll
        for (int k = 0; k < K; k++) c = (lane == k) ? cnt[k] : c;
      } else if (ok) {
        class_pass_rp<K, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                     P.neglog, P.n_neglog, &c, &rp_drawn);
      } else {     // final pass of chain 0, thresholds declined, or weights that underflow
        class_literal<K, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, cur.psi, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                     P.neglog, P.n_neglog, &c, &rp_drawn, last ? ass_out : nullptr);
      }
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
      alpha = alphaN; cur = nw; thr_state = 0;
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
      alpha = alphaN; cur = nw; cJS = ppJS; acc++; thr_state = 0;
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

// FMT 0: dense tiles (dense_pass.cuh); FMT 1: class tiles (class_pass.cuh).
// Shared memory: [ptab | per warp {mbarrier (16 B), slot, threshold rows (FMT 1)}].  The slot
// holds the gene's whole tile (SMEM), or only its class records when the rows are streamed
// from global/L2 (FMT 1, !SMEM).
template <int K, int WARPS, bool SMEM, bool WIDE, int FMT>
__global__ void __launch_bounds__(WARPS * 32, (K <= 6 ? 4 : 3)) chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *s_ptab = reinterpret_cast<double *>(smem);
  const int ptab_bytes = (P.n_ptab * 8 + 15) & ~15;
  const int thr_bytes = FMT == 1 ? P.thr_bytes : 0;
  unsigned char *wbase = smem + ptab_bytes + (size_t) warp * (16 + P.slot_bytes + thr_bytes);
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
    ClassRef cr;
    cr.ncls = FMT == 1 ? d.ncls : 0;
    cr.thr_s = smem_u32(slot + P.slot_bytes);
    cr.rec_s = smem_u32(slot) + (SMEM ? (uint32_t) d.cls_off : 0u);
    cr.meta_s = cr.rec_s + 16u * (uint32_t) cr.ncls;
    __syncwarp();
    if (SMEM) {
      if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, tile_bytes);
        tma_bulk_g2s(slot, P.tiles + d.tile_off, tile_bytes, bar);
      }
    } else if (FMT == 1) {
      // rows stay in global memory; the class records are small and go to the slot
      const uint4 *src = reinterpret_cast<const uint4 *>(P.tiles + d.tile_off + d.cls_off);
      const int n16 = (int) (tile_bytes - (uint32_t) d.cls_off) >> 4;
      for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4 *>(slot)[i] = __ldg(src + i);
    }
    if (FMT == 1 && lane == 0) {      // null class of the padding: no test is ever true
      uint32_t never[8];
#pragma unroll
      for (int k = 0; k < 8; k++) never[k] = 0xffffffffu;
      Thr<K>::store(cr.thr_s + (uint32_t) (Thr<K>::TS * cr.ncls), never);
    }
    if (SMEM) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    __syncwarp();
    if (SMEM)
      run_chain<K, true, WIDE, FMT>(P, d, gi, chain, TileMem<true>::base(slot), smem_u32(s_ptab), cr);
    else
      run_chain<K, false, WIDE, FMT>(P, d, gi, chain, TileMem<false>::base(P.tiles + d.tile_off), smem_u32(s_ptab), cr);
  }
}

}  // namespace misob200
