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
//   warp shuffles  : the K-wide sums / max of the reference (score_iso,
//                    ldirichlet, logit_inv, mvplogisnorm) are reductions inside
//                    a group of eight lanes (group_reduce: gather or xor fold).
//   tensor cores   : unused on purpose -- no dense contraction on this path.
//
// Decision parity: every compare the reference makes in fp64
// (rand*sumpsi vs cumsum, U vs acceptP) is made here on operands that equal the
// reference's up to fp64 rounding (this file is compiled with -fmad=false so a
// multiply feeding an add is not contracted; the thresholds of the read pass carry
// their own exactness argument, class_pass.cuh).  Three things are not the reference's
// operation sequence, each the same value up to rounding: the MH ratio uses
// per-isoform assignment counts (sum_k n_k * logpsi_k) instead of a serial sum over
// reads, the K-wide sums are not added left to right, and the proposal densities
// are evaluated in log space (proposal_scores).  See DESIGN.md "parity contract".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"
#include "plan.hpp"
#include "tile_mem.cuh"

namespace misob200 {

// A gene-chain is 5000 dependent iterations -- tens of milliseconds on one warp -- and the
// chains of a K bucket all cost about the same, so handing out whole chains makes the
// bucket run in lock-stepped waves with a nearly empty last one (bench: 3.02 waves at
// K = 8 took the time of ~3.6).  Chains are therefore cut into SEGMENTS of seg_len steps and
// handed from warp to warp through a ready queue per bucket: a warp that finishes a segment
// stores the chain's ChainState, appends the work unit (one chain here, four consecutive
// chains in quad_kernel.cuh) to the ring and takes the oldest ready unit.  Pop number p is
// unit p itself for p < n_units (first segments), else the (p - n_units)-th push; it can only
// wait when nothing at all is ready, and then some running warp is about to push.
// Everything else a chain carries across iterations is a deterministic function of the
// record: the current point's psi, log psi ... are re-derived from alpha, thresholds
// recomputed, normals are indexed by the iteration number -- the results do not depend on
// seg_len (tests/test_gpu_parity.py::test_results_do_not_depend_on_segment_length).
struct ChainState {
  double alpha[kMaxIso - 1];
  double rp_drawn;
  unsigned long long n_u;    // uniforms consumed
  int cnt[kMaxIso];          // reads assigned per isoform (drawn + fixed)
  int lagc, n_rec, acc, rej;
  int have_rp, pad_;
};
static_assert(sizeof(ChainState) % 16 == 0, "ChainState is 16-byte aligned");

constexpr unsigned kRingEmpty = 0xffffffffu;
#ifndef MISOB200_POLL_NS
#define MISOB200_POLL_NS 500
#endif

// Pop: returns the unit of pop number p.  The WHOLE warp polls (one relaxed load of the same word per
// trip, served by L2 so that the SM's L1 is not invalidated on every poll; one acquire fence at the end):
// a loop that only lane 0 runs leaves the warp split into {lane 0} and {lanes 1..31} -- the compiler
// cannot mark a region with a sleep loop reconvergent -- and a split warp issues every instruction of
// the following segment twice and takes the divergent fall-back of every shuffle
// (profiles/r3_k8_split_warps.txt: 22 % of the popped segments ran that way, the "slow resumed
// segments" of round 1).  After the fence every lane reads the hand-over record (ChainState, progress)
// with ld.global.cg (__ldcg): L2 is the point of coherence, so a line of the same chain left in this
// SM's L1 by an earlier segment can never be what a lane sees.
#ifdef MISOB200_SEG_DEBUG
__device__ unsigned long long g_seg_dbg[8];    // polls, wait cycles, pops, ring pops
#endif
__device__ __forceinline__ unsigned ring_pop(const unsigned *ring, unsigned p, unsigned n_units) {
#ifdef MISOB200_SEG_DEBUG
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_seg_dbg[2], 1ull);
#endif
  if (p < n_units) return p;
  const unsigned *slot = ring + (p - n_units);
  unsigned v;
#ifdef MISOB200_SEG_DEBUG
  const long long t0 = clock64();
  unsigned long long polls = 0;
#endif
  while (true) {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
    v = __shfl_sync(0xffffffffu, v, 0);          // one value for the warp: the loop stays uniform
    if (v != kRingEmpty) break;
    __nanosleep(MISOB200_POLL_NS);
#ifdef MISOB200_SEG_DEBUG
    polls++;
#endif
  }
#ifdef MISOB200_SEG_DEBUG
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&g_seg_dbg[0], polls);
    atomicAdd(&g_seg_dbg[1], (unsigned long long) (clock64() - t0));
    atomicAdd(&g_seg_dbg[3], 1ull);
  }
#endif
  __threadfence();
  __syncwarp();
  return v;
}
// Push: the caller's ChainState stores are ordered before the slot by the fence.
__device__ __forceinline__ void ring_push(unsigned *ring, unsigned *tail, unsigned unit) {
  __threadfence();
  const unsigned q = atomicAdd(tail, 1u);
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ring + q), "r"(unit) : "memory");
}

template <int ROUNDS> struct ChainParamsT {
  const GeneDesc *desc;
  const int *items;          // gene indices of this K bucket, longest first
  int n_genes;               // entries in items
  int n_chains;
  const uint8_t *tiles;
  const double *ptab;
  int n_ptab;
  double ptab_min;           // smallest non-zero entry of ptab
  double tame_slg;           // proposal densities in log space while sum_k log psi_k > tame_slg (kTameSlg; see proposal_scores)
  double *samples;
  double *loglik;
  // optional second destination of every recorded sample: the caller's PINNED host buffers, written
  // straight over PCIe as the chains produce them (8 K bytes per gene-chain every `lag`
  // iterations -- ~1.6 GB/s at cfg-3), so a run ends without a device->host copy of the posteriors
  double *samples_host;
  double *loglik_host;
  uint8_t *drawn;            // final chain-0 assignment of the reads that draw, draw order
  int *accrej;               // [gene][chain][2]
  unsigned *queue;           // work counter
  int n_iters, burn_in, lag, start;
  PhiloxKeyT<ROUNDS> key;     // stream v2: 7 rounds, v1: 10 (philox.cuh)
  int slot_bytes;            // shared-memory bytes per warp for a tile (0: stream tiles from L2)
  // class format only
  const double *neglog;      // neglog[n] = -log(n), n < n_neglog (read scores, miso_paired.c:409-411)
  int n_neglog;
  int thr_bytes;             // shared-memory bytes per warp for the threshold rows (after the slot)
  // chain segments (see ChainState)
  int seg_len;               // steps per segment (a step is the start-up, m = -1, or an iteration)
  int n_seg;                 // ceil((n_iters + 1) / seg_len)
  ChainState *state;         // [gene][chain]
  int *progress;             // [gene][chain]: segments completed (touched by the chain's current holder only)
  unsigned *ring;            // this bucket's ready queue: n_units * (n_seg - 1) slots, all kRingEmpty at launch
  unsigned *ring_tail;       // pushes so far
};
using ChainParams = ChainParamsT<7>;      // what run.cu fills: every ChainParamsT<R> has this layout

}  // namespace misob200

#include "dense_pass.cuh"
#include "class_pass.cuh"

namespace misob200 {

// ---- lane groups ------------------------------------------------------------------
// The scalar part of an iteration is a long chain of dependent fp64 operations
// (exp, log, divisions) on K values: lane-distributed, it keeps K <= 8 of the 32 lanes
// busy.  The warp therefore works on kSpec = 4 PROPOSALS at once: lanes 8g .. 8g+7
// (group g) evaluate everything that depends only on the proposal of iteration m0+g
// -- psi, the normalised log psi, the Dirichlet term, both proposal densities --
// under the assumption that iterations m0 .. m0+g-1 reject (the proposal of an
// iteration is alpha + sd * z, and alpha only changes on an accept).  Iterations then
// consume the groups in order; the first accept throws the rest of the batch away.
// With the 20-40 % acceptance of K >= 3 chains a batch serves ~2.5 iterations for the
// instruction count of one.  Nothing about the arithmetic changes: every group runs
// the same operations on its own operands, whichever group ends up being consumed.
constexpr int kSpec = 4;

// Everything derived from a candidate alpha (member i < K-1 of a group holds alpha_i).
struct Derived {
  double psi;      // member k < K    : psi_k             (logit_inv, miso.c:219-241,467)
  double lp;       // member k < K    : normalised log psi (score_iso, miso.c:136-149)
  double q;        // member i < K-1  : log(psi_i / psi_rest) (mvplogisnorm, miso.c:113)
  double dir;      // group-uniform   : ldirichlet (miso.c:165-182)
  double slg;      // group-uniform   : sum_k log psi_k = -log(1 / prod(theta) / ltheta) (miso.c:104-110), see derive
};

// ---- N-wide reductions inside a lane group ---------------------------------------------
// Two shapes, both leave the same bits in all EIGHT lanes of the group (idle members included: what
// is decided from these values -- accept, stream position, pass type -- has to stay group-uniform):
//   gather : every lane fetches the N members' terms (2 N shuffles, all independent) and adds them
//            pairwise: latency of ONE shuffle + log2(N) additions, 3 N instructions;
//   fold   : xor butterfly over the eight lanes (members >= N pass the identity): 9 instructions
//            whatever N is, but three DEPENDENT shuffle + add rounds.
// The chain is latency-bound (16 warps per SM), so the short dependency chain wins while N is small
// and the instruction count wins at N = 7, 8 (profiles/r3_ab2_reductions.log); kGatherMax is the
// largest N that gathers.  The reference adds the terms one after the other; the sums here differ
// from its by fp64 rounding only -- like the exp / log they are built from, which CUDA and glibc already
// round differently (parity contract).
#ifndef MISOB200_GATHER_MAX
#define MISOB200_GATHER_MAX 4
#endif
constexpr int kGatherMax = MISOB200_GATHER_MAX;
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
struct OpAdd { static __device__ __forceinline__ double id() { return 0.0; }
               static __device__ __forceinline__ double f(double a, double b) { return a + b; } };
struct OpMax { static __device__ __forceinline__ double id() { return -INFINITY; }       // NaN members are ignored
               static __device__ __forceinline__ double f(double a, double b) { return fmax(a, b); } };
template <int N, class OP> __device__ __forceinline__ double group_reduce(double v, int gb, int mi) {
  if (N <= kGatherMax) {
    double a[N];
#pragma unroll
    for (int i = 0; i < N; i++) a[i] = shfl_d(v, gb + i);
#pragma unroll
    for (int w = 1; w < N; w <<= 1)
#pragma unroll
      for (int i = 0; i + w < N; i += 2 * w) a[i] = OP::f(a[i], a[i + w]);
    return a[0];
  }
  v = mi < N ? v : OP::id();
  v = OP::f(v, shfl_xor_d(v, 1));
  v = OP::f(v, shfl_xor_d(v, 2));
  v = OP::f(v, shfl_xor_d(v, 4));
  return v;
}
template <int N> __device__ __forceinline__ double group_add(double v, int gb, int mi) { return group_reduce<N, OpAdd>(v, gb, mi); }
template <int N> __device__ __forceinline__ double group_max(double v, int gb, int mi) { return group_reduce<N, OpMax>(v, gb, mi); }

// mvplogisnorm (miso.c:97-122) evaluates log(covar * (1 / prod(theta) / ltheta) * exp(expPart)) with
// theta = psi_0..K-2, ltheta = 1 - sum(theta) = psi_{K-1} and tmp_i = log(theta_i / ltheta) - mu_i.
// In log space that is  log(covar) - sum_k log psi_k + expPart  and  tmp_i = log psi_i - log psi_{K-1} - mu_i:
// the K logs are already there (score_iso needs them), so the proposal densities cost no exp, no log and
// no division of their own.  The two forms agree to fp64 rounding as long as nothing under- or overflows
// on the literal route; kTameSlg / kTameExp fence that region off (every log psi_k in (-600, 0], so
// prod(theta) >= e^-600; expPart > -700, so exp() stays normal) and anything outside -- a psi of 0, NaN,
// a proposal hundreds of sigmas away -- is evaluated literally, operation by operation, by the *_literal
// bodies below (cold: a chain does not live there, the prior and the likelihood are -inf or NaN too).
// The bound on sum_k log psi_k travels in ChainParams.tame_slg: MISOB200_LITERAL_SCORES=1 sets it to +1
// (never tame), which sends every proposal down the literal route -- how the tests reach that code.
constexpr double kTameSlg = -600.0, kTameExp = -700.0;

// gb = first lane of this lane's group, mi = member index (lane - gb)
template <int K>
__device__ __forceinline__ Derived derive(double alpha, double offset_k, double hyper_m1_k,
                                          double lg_sum, double lg_each, int gb, int mi) {
  constexpr int len = K - 1;
  Derived r;
  // (idle members, mi >= K, carry copies of member 0 / K-1 operands: finite values that keep exp, log and
  // the division on their fast paths; they are masked out of every reduction)
  const double e = d_exp(alpha);
  const double sumexp = group_add<len>(e, gb, mi) + 1.0;
  double psi = d_div(e, sumexp);
  const double sumpsi = group_add<len>(psi, gb, mi);
  if (mi == len) psi = 1 - sumpsi;
  r.psi = psi;

  const double lg = d_log(psi);
  const double t = lg + offset_k;
  const double mx = group_max<K>(t, gb, mi);
  const double ex = d_exp(t - mx);
  const double sum = d_log(group_add<K>(ex, gb, mi)) + mx;
  r.lp = t - sum;

  r.dir = (group_add<K>(hyper_m1_k * lg, gb, mi) + lg_sum) - lg_each;
  r.slg = group_add<K>(lg, gb, mi);
  r.q = lg - shfl_d(lg, gb + len);
  return r;
}

// The literal route of mvplogisnorm for one (theta, mu) pair of a group: member i < K-1 holds theta_i
// and mu_i; returns the score (group-uniform).  Operation order as miso.c:104-119.
template <int K>
__device__ __noinline__ double proposal_score_literal(double theta, double mu, double sigma, double covar, int gb) {
  constexpr int len = K - 1;
  double lth = 1.0, prod = 1.0;
#pragma unroll 1
  for (int i = 0; i < len; i++) {
    const double at = shfl_d(theta, gb + i);
    lth = lth - at;
    prod = prod * at;
  }
  prod = d_div(d_div(1.0, prod), lth);
  const double tmp = d_log(d_div(theta, lth)) - mu;
  const double e = d_div((-0.5) * tmp * tmp, sigma);
  double ep = 0.0;
#pragma unroll 1
  for (int i = 0; i < len; i++) ep = ep + shfl_d(e, gb + i);
  return d_log(covar * prod * d_exp(ep));
}

// ptoCS (theta = psi, mu = alphaNew) and ctoPS (theta = psiNew, mu = alpha) of miso.c:531-534 for
// every group at once.  lcovar = log(covar).
template <int K>
__device__ __forceinline__ void proposal_scores(const Derived &cur, double alpha, const Derived &nw, double alphaN,
                                                double sigma, double covar, double lcovar, double tame_slg,
                                                int gb, int mi, double &ptoCS, double &ctoPS) {
  constexpr int len = K - 1;
  const double t1 = cur.q - alphaN;
  const double t2 = nw.q - alpha;
  const double e1 = d_div((-0.5) * t1 * t1, sigma), e2 = d_div((-0.5) * t2 * t2, sigma);
  const double ep1 = group_add<len>(e1, gb, mi), ep2 = group_add<len>(e2, gb, mi);
  ptoCS = (lcovar - cur.slg) + ep1;
  ctoPS = (lcovar - nw.slg) + ep2;
  const bool tame = cur.slg > tame_slg && nw.slg > tame_slg && ep1 > kTameExp && ep2 > kTameExp;
  if (__any_sync(0xffffffffu, !tame)) {
    const double lit1 = proposal_score_literal<K>(cur.psi, alphaN, sigma, covar, gb);
    const double lit2 = proposal_score_literal<K>(nw.psi, alpha, sigma, covar, gb);
    if (!tame) { ptoCS = lit1; ctoPS = lit2; }
  }
}

// sum_k n_k * v_k over the lane's group, skipping isoforms nothing is assigned to
// (the reference adds one term per read, miso.c:267-271 / :136-149: same value up to rounding)
template <int K>
__device__ __forceinline__ double count_dot(int cnt_k, double v_k, int gb, int mi) {
  return group_add<K>(cnt_k ? (double) cnt_k * v_k : 0.0, gb, mi);
}

template <int K, bool SMEM, bool WIDE, int FMT, int ROUNDS>
__device__ void run_chain(const ChainParamsT<ROUNDS> &P, const GeneDesc &d, int gene_index, int chain,
                          typename TileMem<SMEM>::addr_t rows, uint32_t ptab_s, const ClassRef &cr,
                          int m_begin, int m_end) {
  constexpr int len = K - 1;
  const int lane = threadIdx.x & 31;
  const int gb = lane & 24, mi = lane & 7, grp = lane >> 3;
  const int kk = mi < K ? mi : K - 1;
  const double offset_k = d.offset[kk];
  const double hyper_m1_k = d.hyper_m1[kk];
  const double rs_se_k = d.rs_se[kk];
  const int nfix_k = d.n_fixed[kk];
  const double lg_sum = d.lg_sum, lg_each = d.lg_each;
  const double sigma = d.sigma, sd = d.sd, covar = d.covar_const;
  const double lcovar = d_log(covar);
  const int R2 = d.R2, row_bytes = d.row_bytes, flag_off = d.flag_off, paired = d.paired;
  const int ucode_off = d.ucode_off;
  const bool lp_safe = d.lp_safe != 0 && d.lp_max < P.n_neglog;
  int g_always[K];
#pragma unroll
  for (int k = 0; k < K; k++) g_always[k] = FMT == 1 ? d.g_always[k] : 0;
  int thr_state = 0;       // class format: 0 thresholds stale (psi changed), 1 valid, 2 declined for this psi / fast rule not applicable
  const uint32_t gid = d.gene_id;
  const PhiloxKeyT<ROUNDS> &key = P.key;
  const int *L = d.L;

  ChainState *const st = P.state + ((long long) gene_index * P.n_chains + chain);
  const bool fresh = m_begin < 0;
  unsigned long long n_u = 0;
  // ---- start state, splicing_drift_proposal_init (miso.c:330-447) ----------
  double alpha;            // replicated in every group: member i < K-1 holds alpha_i
  if (!fresh) {
    alpha = __ldcg(&st->alpha[mi < len ? mi : 0]);      // hand-over record: read from L2 (see ring_pop)
    n_u = __ldcg(&st->n_u);
  } else if (P.start == MISOB200_START_AUTO) {
    if (K == 2) { n_u = 1; alpha = 0.0; }     // one uniform drawn and discarded (miso.c:365)
    else alpha = 1.0 / (K - 1);
  } else if (P.start == MISOB200_START_RANDOM) {
    // psi ~ Dirichlet(1,...,1): K gamma(1,1) = -log(uniform) draws, normalised
    // (splicing_rng_get_dirichlet, miso.c:309-326), alpha = logit(psi) (miso.c:202-217).
    // sigma stays SIGMA: the reference leaves it unassigned on this branch.
    const double g = 1.0 * -d_log(stream_uniform((unsigned long long) kk, gid, (uint32_t) chain, key));
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) sum = sum + shfl_d(g, gb + i);
    const double lpsi = d_log(d_div(g, sum));
    alpha = lpsi - shfl_d(lpsi, gb + K - 1);
    n_u = K;
  } else {
    alpha = 0.0;
  }

  // normals are produced 32 at a time into a 64-entry window (lane l holds normals
  // zbase + l and zbase + 32 + l); iteration m consumes normals (m+1)*len .. +len-1
  // (miso.c:851 -> :192-196), a batch looks kSpec iterations ahead
  uint32_t zbase = ((uint32_t) (m_begin + 1) * (uint32_t) len) & ~31u;
  double zbuf0 = stream_normal(zbase + (uint32_t) lane, gid, (uint32_t) chain, key);
  double zbuf1 = stream_normal(zbase + 32u + (uint32_t) lane, gid, (uint32_t) chain, key);

  Derived cur;             // replicated in every group
  cur.psi = cur.lp = cur.q = cur.dir = cur.slg = 0.0;
  double psi_r[K];
  int cnt[K];
  int cnt_k = 0;           // member k: reads currently assigned to isoform k (every group)
  double rp_drawn = 0.0;
  int lagc = 0, n_rec = 0, acc = 0, rej = 0;
  bool have_rp = false;
  if (!fresh) {            // resume: the current point is a function of alpha
    cur = derive<K>(alpha, offset_k, hyper_m1_k, lg_sum, lg_each, gb, mi);
    cnt_k = __ldcg(&st->cnt[kk]);
    rp_drawn = __ldcg(&st->rp_drawn);
    lagc = __ldcg(&st->lagc); n_rec = __ldcg(&st->n_rec); acc = __ldcg(&st->acc); rej = __ldcg(&st->rej);
    have_rp = __ldcg(&st->have_rp) != 0;
  }
  const int S_total = (P.n_iters - P.burn_in) / P.lag;
  uint8_t *ass_out = (chain == 0) ? P.drawn + d.drawn_off : nullptr;

  auto do_pass = [&](int m_next) {
    const bool last = (m_next >= P.n_iters);
    const bool rec_next = d.rp_always || (paired && m_next >= P.burn_in && lagc == P.lag - 1);
    // The fast rules hold for every read of a pass iff no psi is NaN and the smallest weight stays
    // normal; that and the thresholds depend on psi only: re-examined after an accept (thr_state 0),
    // not on every pass (class format; dense tiles pass psi itself to the pass).
    bool ok = false;
    if (!(last && ass_out) && (FMT == 0 || thr_state == 0)) {
      double pmin = 1.0;
#pragma unroll
      for (int k = 0; k < K; k++) {
        psi_r[k] = shfl_d(cur.psi, k);
        pmin = psi_r[k] < pmin ? psi_r[k] : pmin;       // a NaN psi never lowers pmin ...
        ok = ok || !(psi_r[k] == psi_r[k]);             // ... so flag it here
      }
      ok = !ok && (pmin * P.ptab_min >= 1e-290);        // fast rule valid for every read of this pass
      if (FMT == 1) thr_state = (ok && !thr_update<K>(cr, ptab_s, psi_r)) ? 1 : 2;
    }
    if (FMT == 1) ok = !(last && ass_out) && thr_state == 1;
    int c = 0;             // member k: drawing reads assigned to isoform k
    if (FMT == 0) {
      if (ok) {
        if (rec_next && !last)
          reassign_pass<K, 1, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, psi_r, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                    cnt, rp_drawn);
        else
          reassign_pass<K, 0, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, psi_r, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                    cnt, rp_drawn);
#pragma unroll
        for (int k = 0; k < K; k++) c = (mi == k) ? cnt[k] : c;
      } else {     // final pass of chain 0, or a read whose weights underflow: literal rule
        reassign_literal<K, SMEM, WIDE>(rows, row_bytes, flag_off, ptab_s, cur.psi, n_u, R2, gid, (uint32_t) chain, key, paired,
                                  L, &c, &rp_drawn, last ? ass_out : nullptr);
      }
    } else {
      if (ok && !(rec_next && !last && paired)) {
        class_pass<K, SMEM>(rows, cr, n_u, R2, gid, (uint32_t) chain, key, g_always, cnt);
#pragma unroll
        for (int k = 0; k < K; k++) c = (mi == k) ? cnt[k] : c;
      } else if (ok) {
        // (only paired-end passes get here with a read score to compute; see rec_next)
        class_pass_rp<K, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, n_u, R2, gid, (uint32_t) chain, key, g_always,
                                     P.neglog, P.n_neglog, lp_safe, &c, &rp_drawn);
      } else {     // final pass of chain 0, thresholds declined, or weights that underflow
        class_literal<K, SMEM, WIDE>(rows, ucode_off, cr, ptab_s, cur.psi, n_u, R2, gid, (uint32_t) chain, key, paired, L,
                                     P.neglog, P.n_neglog, &c, &rp_drawn, last ? ass_out : nullptr);
      }
    }
    cnt_k = c + nfix_k;
    n_u += (unsigned long long) R2;
    return rec_next || !ok;
  };

  // the batch: group g holds the proposal of iteration m0 + g, valid while no iteration
  // from m0 on has accepted
  int m0 = 0;
  bool batch_ok = false;
  double alphaB = 0.0, scP = 0.0, scC = 0.0;
  Derived nwB;
  nwB.psi = nwB.lp = nwB.q = nwB.dir = nwB.slg = 0.0;

  auto make_batch = [&](int m) {
    m0 = m;
    batch_ok = true;
    // ---- propose (miso.c:851): alphaNew = alpha + sd * N(0,1), group g for iteration m + g
    const uint32_t first = (uint32_t) (m + 1) * (uint32_t) len;
    while (first - zbase >= 32u) {         // slide the window (first grows by len < 32 per iteration)
      zbase += 32u;
      zbuf0 = zbuf1;
      zbuf1 = stream_normal(zbase + 32u + (uint32_t) lane, gid, (uint32_t) chain, key);
    }
    const uint32_t zi = first - zbase + (uint32_t) (grp * len + (mi < len ? mi : 0));     // < 32 + kSpec * len <= 60
    const double za = shfl_d(zbuf0, (int) (zi & 31u)), zb = shfl_d(zbuf1, (int) (zi & 31u));
    const double z = zi < 32u ? za : zb;
    alphaB = alpha + sd * z;
    nwB = derive<K>(alphaB, offset_k, hyper_m1_k, lg_sum, lg_each, gb, mi);
    // ---- proposal densities (miso.c:531-534, :97-122), see proposal_scores
    proposal_scores<K>(cur, alpha, nwB, alphaB, sigma, covar, lcovar, P.tame_slg, gb, mi, scP, scC);     // ptoCS, ctoPS
  };

  // m == -1 is the start-up proposal, adopted unconditionally (miso.c:834), followed
  // by the initial assignment (miso.c:840-843); m >= 0 are the iterations proper.
  for (int m = m_begin; m < m_end; m++) {
    if (!batch_ok || m - m0 >= kSpec) make_batch(m);
    const int src = 8 * (m - m0) + mi;          // this member in the group of iteration m
    if (m < 0) {
      alpha = shfl_d(alphaB, src);
      cur.psi = shfl_d(nwB.psi, src); cur.lp = shfl_d(nwB.lp, src); cur.q = shfl_d(nwB.q, src);
      cur.dir = shfl_d(nwB.dir, src); cur.slg = shfl_d(nwB.slg, src);
      batch_ok = false; thr_state = 0;
    } else {
    // ---- joint scores (miso.c:524-529); every group scores its own proposal --------
    double rp;
    if (!paired) rp = count_dot<K>(cnt_k, rs_se_k, gb, mi);      // sum_r isoscores[ass_r], miso.c:267-271
    else rp = have_rp ? d.rp_fixed + rp_drawn : 0.0;        // cancels in the ratio when not recorded
    const double ppJS_g = rp + count_dot<K>(cnt_k, nwB.lp, gb, mi) + nwB.dir;
    const double pcJS = rp + count_dot<K>(cnt_k, cur.lp, gb, mi) + cur.dir;
    const double acceptP_g = d_exp((m > 0) ? ppJS_g + scP - (pcJS + scC) : ppJS_g - pcJS);
    const double acceptP = shfl_d(acceptP_g, src & 24);

    // ---- accept (miso.c:869-880): the uniform is drawn only if acceptP < 1 --
    bool accept = acceptP >= 1;
    if (!accept) {
      const double u = stream_uniform(n_u, gid, (uint32_t) chain, key);
      n_u++;
      accept = u < acceptP;
    }
    double cJS = pcJS;
    if (accept) {
      cJS = shfl_d(ppJS_g, src & 24);
      alpha = shfl_d(alphaB, src);
      cur.psi = shfl_d(nwB.psi, src); cur.lp = shfl_d(nwB.lp, src); cur.q = shfl_d(nwB.q, src);
      cur.dir = shfl_d(nwB.dir, src); cur.slg = shfl_d(nwB.slg, src);
      acc++; thr_state = 0; batch_ok = false;
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
          if (P.samples_host) {
            if (lane < K) P.samples_host[d.sample_off + col * K + lane] = cur.psi;
            if (lane == 0) P.loglik_host[d.loglik_off + col] = cJS;
          }
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

  if (m_end >= P.n_iters) {
    if (lane == 0) {
      int *ar = P.accrej + ((long long) gene_index * P.n_chains + chain) * 2;
      ar[0] = acc; ar[1] = rej;
    }
  } else {                 // hand the chain over to whoever holds the next segment's ticket
    if (lane < len) st->alpha[lane] = alpha;
    if (lane < K) st->cnt[lane] = cnt_k;
    if (lane == 0) {
      st->rp_drawn = rp_drawn; st->n_u = n_u;
      st->lagc = lagc; st->n_rec = n_rec; st->acc = acc; st->rej = rej;
      st->have_rp = have_rp ? 1 : 0;
    }
  }
}

// FMT 0: dense tiles (dense_pass.cuh); FMT 1: class tiles (class_pass.cuh).
// Shared memory: [ptab | per warp {mbarrier (16 B), slot, threshold rows (FMT 1)}].  The slot
// holds the gene's whole tile (SMEM), or only its class records when the rows are streamed
// from global/L2 (FMT 1, !SMEM).
// Class-format kernels run as ONE 16-warp CTA per SM (128 registers x 512 threads = the whole
// register file): an SM then executes a single kernel's code at a time.  The per-iteration
// instruction footprint of one K is 30-48 KB; with CTAs of several K buckets resident on the same
// SM the instruction caches thrash and everything runs ~5x slower (profiles/r2_ab1_sched.log),
// so concurrency between buckets is arranged SM by SM (run.cu, "balanced").  The warps of a CTA
// stay independent (one __syncthreads at start-up).  Dense-format kernels (rare fallback) keep
// 4-warp CTAs, 3-4 per SM.
#ifndef MISOB200_CLASS_WARPS
#define MISOB200_CLASS_WARPS 16      /* A/B builds: 20 (96 registers) with MISOB200_PASS_NOINLINE */
#endif
constexpr int kClassWarps = MISOB200_CLASS_WARPS;
template <int K, int WARPS, bool SMEM, bool WIDE, int FMT, int ROUNDS>
__global__ void __launch_bounds__(WARPS * 32, (FMT == 1 ? 1 : (K <= 6 ? 4 : 3))) chain_kernel(const __grid_constant__ ChainParamsT<ROUNDS> P) {
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

  const unsigned n_items = (unsigned) (P.n_genes * P.n_chains);        // work unit = one gene-chain
  const unsigned n_pops = n_items * (unsigned) P.n_seg;
  uint32_t phase = 0;
  while (true) {
    unsigned p = 0;
    if (lane == 0) p = atomicAdd(P.queue, 1u);
    p = __shfl_sync(0xffffffffu, p, 0);
    const unsigned item = p < n_pops ? ring_pop(P.ring, p, n_items) : kRingEmpty;      // warp-uniform
    if (item == kRingEmpty) break;
    const int gi = P.items[item / (unsigned) P.n_chains];
    const int chain = (int) (item % (unsigned) P.n_chains);
    const GeneDesc &d = P.desc[gi];
    int *const progress = P.progress + ((long long) gi * P.n_chains + chain);
    const int seg = __ldcg(progress);
    const uint32_t tile_bytes = (uint32_t) d.tile_bytes;
    ClassRef cr;
    cr.ncls = FMT == 1 ? d.ncls : 0;
    cr.thr_s = smem_u32(slot + P.slot_bytes) + 32u;
    cr.thrb_s = cr.thr_s + (uint32_t) Thr<K>::plane_a_bytes(cr.ncls);
    cr.l_s = smem_u32(slot + P.slot_bytes);
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
      const int n16 = (d.core_bytes - d.cls_off) >> 4;
      for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4 *>(slot)[i] = __ldg(src + i);
    }
    if (FMT == 1 && lane < kMaxIso) reinterpret_cast<int *>(slot + P.slot_bytes)[lane] = d.L[lane];
    if (FMT == 1 && lane == 0) {      // null class of the padding: no test is ever true
      uint32_t never[8];
#pragma unroll
      for (int k = 0; k < 8; k++) never[k] = 0u;        // rows hold ~t_k
      Thr<K>::store(cr.thr_s + (uint32_t) (Thr<K>::TSA * cr.ncls), cr.thrb_s + (uint32_t) (Thr<K>::TSB * cr.ncls), never);
    }
    if (SMEM) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    __syncwarp();
    const int m_begin = seg * P.seg_len - 1;
    const int m_end = min(m_begin + P.seg_len, P.n_iters);
    if (SMEM)
      run_chain<K, true, WIDE, FMT>(P, d, gi, chain, TileMem<true>::base(slot), smem_u32(s_ptab), cr, m_begin, m_end);
    else
      run_chain<K, false, WIDE, FMT>(P, d, gi, chain, TileMem<false>::base(P.tiles + d.tile_off), smem_u32(s_ptab), cr,
                                     m_begin, m_end);
    if (m_end < P.n_iters) {
      __syncwarp();
      if (lane == 0) { *progress = seg + 1; ring_push(P.ring, P.ring_tail, item); }
    }
  }
}

}  // namespace misob200
