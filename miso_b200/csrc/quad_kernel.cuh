// miso_b200/csrc/quad_kernel.cuh -- FOUR gene-chains per warp, eight lanes each (sm_100a).
//
// chain_kernel.cuh gives a whole warp to one gene-chain.  ncu on that kernel
// (profiles/r1_v5_*) showed where its time goes at K = 5: the counting pass over the
// reads -- the only part that uses 32 lanes -- is 46 % of the instructions but 30 % of
// the time; the per-iteration scalar part (proposal, logit_inv, score_iso, ldirichlet,
// mvplogisnorm, MH ratio: SURVEY.md section 8a rows a-4 ... a-10) is a long dependent
// chain of fp64 exp/log/div on K <= 8 values that keeps 8 lanes busy and takes the
// rest, and for the K = 2 events (few reads that draw) it is everything.
//
// Here a warp carries four gene-chains of the same K bucket in lock step -- lanes
// 8g .. 8g+7 ("group g") own chain g: member k of a group holds alpha_k, psi_k, the
// normalised log psi_k ... of isoform k exactly as in chain_kernel.cuh, so every
// instruction of the scalar part now serves four chains, and all control flow that
// depends on the iteration number (record, read-score pass, last pass) stays
// warp-uniform because the four chains run the same iteration.  What depends on the
// chain -- accept or reject, the conditional accept draw (miso.c:870) -- is selects,
// not branches.
//
// The reads: a group walks its own gene's class-id row, lane m taking Philox blocks
// m, m+8, m+16 ... of the gene-chain's uniform stream (4 consecutive reads each),
// thresholds and counting as in class_pass.cuh.  The trip count is the largest of the
// four (the work list is sorted by reads that draw, neighbours are alike); a group that
// runs out re-reads the null ids of its own padding.
//
// Shared memory holds, per warp, four "core" tiles (class-id row + class records,
// GeneDesc.core_bytes, one TMA bulk copy each) and four threshold areas.  The
// uniform-code row (read-score passes: one pass in `lag`) and the insert-length table
// are read through L1 instead, so that 16 warps x 4 tiles still fit an SM.
//
// Same arithmetic, same streams, same decisions as chain_kernel.cuh: both are checked
// against the oracle by the same tests (tests/test_gpu_parity.py runs every layout).
#pragma once
#include "chain_kernel.cuh"

namespace misob200 {

constexpr int kQuad = 4;      // gene-chains per warp
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t group_sum(uint32_t v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 4);
  return v;
}
__device__ __forceinline__ double group_sum(double v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 4);
  return v;
}
__device__ __forceinline__ uint32_t ldg_u32(const unsigned char *p) {
  return __ldg(reinterpret_cast<const uint32_t *>(p));
}

// ---- thresholds (class_pass.cuh thr_update) for the groups that need them --------
// The (group, class) pairs of the groups whose psi changed are dealt to the 32 lanes;
// a lane fetches "its" group's psi and shared-memory addresses by shuffle.  Returns
// the groups (bit g) with a threshold too close to an integer to be trusted.
template <int K>
__device__ __forceinline__ uint32_t quad_thr_update(bool need, const ClassRef &cr, const double *__restrict__ ptab,
                                                    double psi_k) {
  constexpr int NT = Thr<K>::NT;
  const int lane = threadIdx.x & 31;
  const int n_mine = need ? cr.ncls : 0;
  const int o1 = __shfl_sync(kFull, n_mine, 0);
  const int o2 = o1 + __shfl_sync(kFull, n_mine, 8);
  const int o3 = o2 + __shfl_sync(kFull, n_mine, 16);
  const int total = o3 + __shfl_sync(kFull, n_mine, 24);
  uint32_t declined = 0;
  __syncwarp();          // the previous pass's threshold loads (other lanes) before these stores
  for (int base = 0; base < total; base += 32) {
    const int p = base + lane;
    const bool on = p < total;
    const int g = p >= o3 ? 3 : p >= o2 ? 2 : p >= o1 ? 1 : 0;
    const int c = on ? p - (g == 3 ? o3 : g == 2 ? o2 : g == 1 ? o1 : 0) : 0;
    const int src = 8 * g;
    double psi[K];
#pragma unroll
    for (int k = 0; k < K; k++) psi[k] = shfl_d(psi_k, src + k);
    const uint32_t rec_s = __shfl_sync(kFull, cr.rec_s, src), meta_s = __shfl_sync(kFull, cr.meta_s, src);
    const uint32_t thr_s = __shfl_sync(kFull, cr.thr_s, src), thrb_s = __shfl_sync(kFull, cr.thrb_s, src);
    bool bad = false;
    if (on) {
      uint4 rec;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(rec.x), "=r"(rec.y), "=r"(rec.z), "=r"(rec.w) : "r"(rec_s + 16u * c));
      const uint32_t rw[4] = {rec.x, rec.y, rec.z, rec.w};
      const int first = (int) (lds_u32(meta_s + 4u * c) & 0xffu);
      double S = 0.0, C[K];
#pragma unroll
      for (int k = 0; k < K; k++) {
        const uint32_t idx = (k & 1) ? (rw[k >> 1] >> 16) : (rw[k >> 1] & 0xffffu);
        S = S + psi[k] * __ldg(ptab + idx);             // CUMSUM, miso_paired.c:11-22
        C[k] = S;
      }
      const double inv = d_div(4294967296.0, S);
      uint32_t t[8];
#pragma unroll
      for (int k = 0; k < 8; k++) t[k] = 0u;
#pragma unroll
      for (int k = 0; k < NT; k++) {
        const double tau = C[k] * inv - 0.5;
        const uint32_t tk = __double2uint_rz(tau);
        const double fr = tau - (double) tk;
        const bool good = fr > kThrMargin && fr < 1.0 - kThrMargin;
        if (k >= first) {
          t[k] = ~tk;
          bad = bad || !good;
        }
      }
      Thr<K>::store(thr_s + (uint32_t) (Thr<K>::TSA * c), thrb_s + (uint32_t) (Thr<K>::TSB * c), t);
    }
#pragma unroll
    for (int q = 0; q < kQuad; q++)
      if (__any_sync(kFull, bad && g == q)) declined |= 1u << q;
  }
  __syncwarp();
  return declined;
}

// ---- one reassignment pass of the four chains --------------------------------------
// class_pass_body with lane m of a group on Philox blocks Q0 + m + 8 s.  `a0` is the
// lane's first id word, `a_end` the last (all-null) word of the group's id row: a lane
// past its gene's reads keeps re-reading that word.  A group that sits this pass out
// (literal rule instead) passes a0 = a_end.
//   MODE 0: counts.  MODE 1: + read score of the chosen isoforms (miso_paired.c:157-163),
//   uniform codes from global memory.
template <int K, int MODE, bool WIDE, class KEY>
__device__ __forceinline__ void quad_pass(uint32_t a0, uint32_t a_end, int my_steps,
                                          const unsigned char *__restrict__ ucode, int row_last,
                                          const ClassRef &cr, const double *__restrict__ ptab,
                                          unsigned long long n_u, int R2, uint32_t gene, uint32_t chain,
                                          const KEY &key, const int (&g_always)[K],
                                          const double *__restrict__ neglog, int n_neglog, int (&cnt)[K],
                                          double &rp) {
  constexpr int NT = Thr<K>::NT, TSA = Thr<K>::TSA, TSB = Thr<K>::TSB;
  const int mi = threadIdx.x & 7;
  const int o = (int) (n_u & 3ull);
  const uint32_t Q0 = (uint32_t) (n_u >> 2);
  const int nsteps = __reduce_max_sync(kFull, my_steps);
  const uint32_t sel = 0x3210u + 0x1111u * (uint32_t) (3 - o);
  uint32_t G[NT];
#pragma unroll
  for (int k = 0; k < NT; k++) G[k] = 0;
  uint32_t a = a0;
  int uw = mi;                                          // MODE 1: index of the lane's 4-read window
  const int hs = 3 - o;
  const bool hb = (hs >> 1) != 0;
  const uint32_t hsh = 16u * (uint32_t) (hs & 1);
  const uint32_t thr_s = cr.thr_s, thrb_s = cr.thrb_s;
  const uint32_t ncls = (uint32_t) cr.ncls;
  double rp_lane = 0.0;
  uint32_t tot = 0;
#pragma unroll (MODE == 0 ? kUnrollCount : kUnrollScore)
  for (int s = 0; s < nsteps; s++) {
    uint32_t x[4];
    philox4x32(Q0 + (uint32_t) (mi + 8 * s), 0u, gene, chain, key, x);
    const uint32_t aa = a < a_end ? a : a_end;
    const uint32_t ids = __byte_perm(TileMem<true>::ld(aa), TileMem<true>::ld(aa + 4), sel);
    a += 32;
    uint32_t uc01 = 0, uc23 = 0;
    if (MODE == 1) {
      // clamp like the id row (the ids there are null, the codes are never used); the 16-bit
      // row is read 16 bytes at a time and ends 16 bytes after 2 * padded
      const int lim = WIDE ? row_last - 2 : row_last;
      const int w = uw < lim ? uw : lim;
      if (!WIDE) {
        uc01 = __byte_perm(ldg_u32(ucode + 4 * w), ldg_u32(ucode + 4 * w + 4), sel);
      } else {
        const uint32_t w0 = ldg_u32(ucode + 8 * w), w1 = ldg_u32(ucode + 8 * w + 4);
        const uint32_t w2 = ldg_u32(ucode + 8 * w + 8), w3 = ldg_u32(ucode + 8 * w + 12);
        const uint32_t wa = hb ? w1 : w0, wb = hb ? w2 : w1, wc = hb ? w3 : w2;
        uc01 = __funnelshift_r(wa, wb, hsh);
        uc23 = __funnelshift_r(wb, wc, hsh);
      }
      uw += 8;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t id = __byte_perm(ids, 0u, 0x4440u | (uint32_t) i);
      uint32_t nt[8];
      Thr<K>::load(thr_s + id * (uint32_t) TSA, thrb_s + id * (uint32_t) TSB, nt);
#pragma unroll
      for (int k = 0; k < NT; k++)
        asm("{\n\t.reg .u32 j;\n\tadd.cc.u32 j, %1, %2;\n\taddc.u32 %0, %0, 0;\n\t}" : "+r"(G[k]) : "r"(x[i]), "r"(nt[k]));
      if (MODE == 1) {
        uint32_t now = 0;
#pragma unroll
        for (int k = 0; k < NT; k++) now += G[k];
        const uint32_t meta = lds_u32(cr.meta_s + 4u * id);
        const uint32_t chosen = (meta & 0xffu) + (now - tot);
        tot = now;
        uint32_t cc;
        if (!WIDE) cc = __byte_perm(uc01, 0u, 0x4440u | (uint32_t) i);
        else cc = __byte_perm(i < 2 ? uc01 : uc23, 0u, (i & 1) ? 0x4432u : 0x4410u);
        if (!(meta & 0x100u)) cc = lds_u16(cr.rec_s + 16u * id + 2u * chosen);
        const int lp = (int) lds_u32(cr.l_s + 4u * chosen) - ((int) cc - 1);
        const double sc = neg_log_lp(lp, neglog, n_neglog) + __ldg(ptab + cc);   // isoscores, miso_paired.c:409-411
        if (id != ncls) rp_lane += sc;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NT; k++) G[k] = group_sum(G[k]) + (uint32_t) g_always[k];
  cnt[0] = R2 - (int) G[0];
#pragma unroll
  for (int k = 1; k < NT; k++) cnt[k] = (int) (G[k - 1] - G[k]);
  cnt[K - 1] = (int) G[NT - 1];
  if (MODE == 1) rp = group_sum(rp_lane);
}

template <int K, bool WIDE, class KEY>
__device__ __noinline__ void quad_pass_rp(uint32_t a0, uint32_t a_end, int my_steps, const unsigned char *__restrict__ ucode,
                                          int row_last, const ClassRef &cr, const double *__restrict__ ptab,
                                          unsigned long long n_u, int R2, uint32_t gene, uint32_t chain,
                                          const KEY &key, const int (&g_always)[K],
                                          const double *__restrict__ neglog, int n_neglog, int *cnt_k, double *rp) {
  int cnt[K];
  quad_pass<K, 1, WIDE>(a0, a_end, my_steps, ucode, row_last, cr, ptab, n_u, R2, gene, chain, key, g_always, neglog,
                        n_neglog, cnt, *rp);
  int c = 0;
#pragma unroll
  for (int k = 0; k < K; k++) c = ((threadIdx.x & 7) == k) ? cnt[k] : c;
  *cnt_k = c;
}

// ---- the literal rule of miso.c:59-83 for ONE group (class_literal, eight lanes) -----
// Runs under divergence: only the lanes of `gmask` are here.
template <int K, bool WIDE, class KEY>
__device__ __noinline__ void quad_literal(unsigned gmask, uint32_t rows, const unsigned char *__restrict__ ucode,
                                          const ClassRef &cr, const double *__restrict__ ptab, double psi_k,
                                          unsigned long long n_u, int R2, uint32_t gene, uint32_t chain,
                                          const KEY &key, int paired, const double *__restrict__ neglog,
                                          int n_neglog, int *cnt_k, double *rp, uint8_t *__restrict__ ass_out) {
  const int lane = threadIdx.x & 31, mi = lane & 7, gb = lane & 24;
  const int o = (int) (n_u & 3ull);
  const uint32_t Q0 = (uint32_t) (n_u >> 2);
  const int nsteps = (R2 + o + 31) >> 5;
  double psi[K];
#pragma unroll
  for (int k = 0; k < K; k++) psi[k] = __shfl_sync(gmask, psi_k, gb + k);
  int n[K];
#pragma unroll
  for (int k = 0; k < K; k++) n[k] = 0;
  double rp_lane = 0.0;
  for (int s = 0; s < nsteps; s++) {
    const int T = mi + 8 * s;
    uint32_t x[4];
    philox4x32(Q0 + (uint32_t) T, 0u, gene, chain, key, x);
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const int rank = 4 * T - o + i;
      if (rank < 0 || rank >= R2) continue;
      const int el = kTilePadFront + rank;
      const uint32_t id = (TileMem<true>::ld(rows + (el & ~3)) >> (8 * (el & 3))) & 0xffu;
      const uint32_t meta = lds_u32(cr.meta_s + 4u * id);
      uint32_t uc = 0;
      if (meta & 0x100u) {
        const int byte = WIDE ? 2 * el : el;
        const uint32_t w = ldg_u32(ucode + (byte & ~3));
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
        S = S + psi[k] * __ldg(ptab + code[k]);
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
        const int lp = (int) lds_u32(cr.l_s + 4u * (uint32_t) chosen) - ((int) cc - 1);
        rp_lane += neg_log_lp(lp, neglog, n_neglog) + __ldg(ptab + cc);
      }
      if (ass_out) ass_out[rank] = (uint8_t) chosen;
    }
  }
  int mine = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    int t = n[k];
    t += __shfl_xor_sync(gmask, t, 1);
    t += __shfl_xor_sync(gmask, t, 2);
    t += __shfl_xor_sync(gmask, t, 4);
    if (mi == k) mine = t;
  }
  *cnt_k = mine;
  rp_lane += __shfl_xor_sync(gmask, rp_lane, 1);
  rp_lane += __shfl_xor_sync(gmask, rp_lane, 2);
  rp_lane += __shfl_xor_sync(gmask, rp_lane, 4);
  *rp = rp_lane;
}

// ---- the kernel ---------------------------------------------------------------------
// Shared memory per warp: [mbarrier 16 B | 4 x slot (core tile) | 4 x {L_k 32 B, threshold planes}].
// slot_bytes is 32 mod 128, so the four groups' id words of one step sit in different banks.
template <int K, int WARPS, bool WIDE, int ROUNDS>
__global__ void __launch_bounds__(WARPS * 32, 1) quad_kernel(const __grid_constant__ ChainParamsT<ROUNDS> P) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int len = K - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gb = lane & 24, mi = lane & 7, grp = lane >> 3;
  const int kk = mi < K ? mi : K - 1;
  unsigned char *wbase = smem + (size_t) warp * (16 + kQuad * (P.slot_bytes + P.thr_bytes));
  uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
  unsigned char *slot = wbase + 16 + grp * P.slot_bytes;
  unsigned char *thr = wbase + 16 + kQuad * P.slot_bytes + grp * P.thr_bytes;
  const double *__restrict__ ptab = P.ptab;
  const PhiloxKeyT<ROUNDS> &key = P.key;

  if (lane == 0) { mbar_init(bar, kQuad); fence_mbar_init(); }
  __syncwarp();

  const int n_items = P.n_genes * P.n_chains;
  const unsigned n_units = (unsigned) (n_items + kQuad - 1) / kQuad;    // work unit = four consecutive gene-chains
  const unsigned n_pops = n_units * (unsigned) P.n_seg;
  const int S_total = (P.n_iters - P.burn_in) / P.lag;
  uint32_t phase = 0;
  while (true) {
    unsigned p = 0;
    if (lane == 0) p = atomicAdd(P.queue, 1u);
    p = __shfl_sync(kFull, p, 0);
    const unsigned unit = p < n_pops ? ring_pop(P.ring, p, n_units) : kRingEmpty;      // the whole warp polls (ring_pop)
    if (unit == kRingEmpty) break;
#ifdef MISOB200_SEG_DEBUG
    const long long dbg_t0 = clock64();
#endif
    const int item0 = (int) unit * kQuad;
    // a group without an item of its own shadows the last one and writes nothing
    const bool live = item0 + grp < n_items;
    const int item = live ? item0 + grp : n_items - 1;
    const int gi = P.items[item / P.n_chains];
    const int chain = item % P.n_chains;
    const GeneDesc &d = P.desc[gi];
    ChainState *const st = P.state + ((long long) gi * P.n_chains + chain);
    int *const progress = P.progress + ((long long) gi * P.n_chains + chain);
    const int seg = __ldcg(progress);                 // the same for the four chains of a unit
    const int m_begin = seg * P.seg_len - 1;
    const int m_end = min(m_begin + P.seg_len, P.n_iters);
    const bool fresh = seg == 0;

    ClassRef cr;
    cr.ncls = d.ncls;
    cr.l_s = smem_u32(thr);
    cr.thr_s = cr.l_s + 32u;
    cr.thrb_s = cr.thr_s + (uint32_t) Thr<K>::plane_a_bytes(cr.ncls);
    cr.rec_s = smem_u32(slot) + (uint32_t) d.cls_off;
    cr.meta_s = cr.rec_s + 16u * (uint32_t) cr.ncls;
    __syncwarp();
    if (mi == 0) {
      fence_proxy_async();
      mbar_expect_tx(bar, (uint32_t) d.core_bytes);
      tma_bulk_g2s(slot, P.tiles + d.tile_off, (uint32_t) d.core_bytes, bar);
    }
    reinterpret_cast<int *>(thr)[mi] = d.L[mi];
    if (mi == 0) {       // null class of the padding: no test is ever true
      uint32_t never[8];
#pragma unroll
      for (int k = 0; k < 8; k++) never[k] = 0u;
      Thr<K>::store(cr.thr_s + (uint32_t) (Thr<K>::TSA * cr.ncls), cr.thrb_s + (uint32_t) (Thr<K>::TSB * cr.ncls), never);
    }

    // ---- per-chain constants ---------------------------------------------------------
    const double offset_k = d.offset[kk], hyper_m1_k = d.hyper_m1[kk], rs_se_k = d.rs_se[kk];
    const int nfix_k = d.n_fixed[kk];
    const double lg_sum = d.lg_sum, lg_each = d.lg_each;
    const double sigma = d.sigma, sd = d.sd, covar = d.covar_const, rp_fixed = d.rp_fixed;
    const double lcovar = d_log(covar);
    const int R2 = d.R2, paired = d.paired, rp_always = d.rp_always;
    const uint32_t gid = d.gene_id;
    int g_always[K];
#pragma unroll
    for (int k = 0; k < K; k++) g_always[k] = d.g_always[k];
    const uint32_t rows = smem_u32(slot);
    const uint32_t a_end = rows + (uint32_t) d.row_bytes - 8u;
    const int row_last = (d.row_bytes >> 2) - 2;          // word index of a_end
    const unsigned char *ucode = P.tiles + d.tile_off + d.ucode_off;
    uint8_t *ass_out = (chain == 0 && live) ? P.drawn + d.drawn_off : nullptr;
    const unsigned gmask = 0xffu << gb;

    mbar_wait(bar, phase);
    phase ^= 1u;
    __syncwarp();

    // ---- start state, splicing_drift_proposal_init (miso.c:330-447) ----------
    unsigned long long n_u = 0;
    double alpha;
    if (!fresh) {
      alpha = __ldcg(&st->alpha[mi < len ? mi : 0]);      // hand-over record: read from L2 (see ring_pop)
      n_u = __ldcg(&st->n_u);
    } else if (P.start == MISOB200_START_AUTO) {
      if (K == 2) { n_u = 1; alpha = 0.0; }     // one uniform drawn and discarded (miso.c:365)
      else alpha = 1.0 / (K - 1);
    } else if (P.start == MISOB200_START_RANDOM) {
      // Dirichlet(1,...,1) by K gamma(1,1) = -log(uniform) draws (miso.c:309-326, :388-404)
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

    Derived cur;
    cur.psi = cur.lp = cur.q = cur.dir = cur.slg = 0.0;
    int cnt_k = 0;
    double rp_drawn = 0.0;
    int lagc = 0, n_rec = 0, acc = 0, rej = 0;
    int thr_state = 0;       // 0 thresholds stale (psi changed), 1 valid, 2 declined for this psi
    bool have_rp = false;
    if (!fresh) {            // resume: the current point is a function of alpha (ChainState)
      cur = derive<K>(alpha, offset_k, hyper_m1_k, lg_sum, lg_each, gb, mi);
      cnt_k = __ldcg(&st->cnt[kk]);
      rp_drawn = __ldcg(&st->rp_drawn);
      lagc = __ldcg(&st->lagc); n_rec = __ldcg(&st->n_rec); acc = __ldcg(&st->acc); rej = __ldcg(&st->rej);
      have_rp = __ldcg(&st->have_rp) != 0;
    }

    // normals are produced eight at a time per chain into a 16-entry window (member l of a group
    // holds normals zbase + l and zbase + 8 + l of its chain): one Box-Muller evaluation per 8 / (K-1)
    // iterations instead of one per iteration.  The four chains run the same iteration, so the
    // window position is warp-uniform.
    uint32_t zbase = ((uint32_t) (m_begin + 1) * (uint32_t) len) & ~7u;
    double zbuf0 = stream_normal(zbase + (uint32_t) mi, gid, (uint32_t) chain, key);
    double zbuf1 = stream_normal(zbase + 8u + (uint32_t) mi, gid, (uint32_t) chain, key);

    // m == -1 is the start-up proposal, adopted unconditionally (miso.c:834), followed by the
    // initial assignment (miso.c:840-843); m >= 0 are the iterations proper.
    for (int m = m_begin; m < m_end; m++) {
      // ---- propose (miso.c:851): alphaNew = alpha + sd * N(0,1); normals (m+1)(K-1) .. +K-2
      const uint32_t first = (uint32_t) (m + 1) * (uint32_t) len;
      if (first - zbase >= 8u) {             // slide the window (first grows by len < 8 per iteration; warp-uniform)
        zbase += 8u;
        zbuf0 = zbuf1;
        zbuf1 = stream_normal(zbase + 8u + (uint32_t) mi, gid, (uint32_t) chain, key);
      }
      const uint32_t zi = first - zbase + (uint32_t) (mi < len ? mi : 0);        // < 8 + len <= 14
      const double za = shfl_d(zbuf0, gb + (int) (zi & 7u)), zb = shfl_d(zbuf1, gb + (int) (zi & 7u));
      const double z = zi < 8u ? za : zb;
      const double alphaN = alpha + sd * z;
      const Derived nw = derive<K>(alphaN, offset_k, hyper_m1_k, lg_sum, lg_each, gb, mi);
      bool accept = true;
      double cJS = 0.0;
      if (m >= 0) {
        // ---- proposal densities (miso.c:531-534, :97-122), see proposal_scores (chain_kernel.cuh)
        double scP, scC;                                   // ptoCS, ctoPS
        proposal_scores<K>(cur, alpha, nw, alphaN, sigma, covar, lcovar, P.tame_slg, gb, mi, scP, scC);
        // ---- joint scores (miso.c:524-529) ------------------------------------------------
        double rp;
        if (!paired) rp = count_dot<K>(cnt_k, rs_se_k, gb, mi);       // sum_r isoscores[ass_r], miso.c:267-271
        else rp = have_rp ? rp_fixed + rp_drawn : 0.0;           // cancels in the ratio when not recorded
        const double ppJS = rp + count_dot<K>(cnt_k, nw.lp, gb, mi) + nw.dir;
        const double pcJS = rp + count_dot<K>(cnt_k, cur.lp, gb, mi) + cur.dir;
        const double acceptP = d_exp((m > 0) ? ppJS + scP - (pcJS + scC) : ppJS - pcJS);
        // ---- accept (miso.c:869-880): the uniform is consumed only if !(acceptP >= 1) ------
        const bool sure = acceptP >= 1;
        const double u = stream_uniform(n_u, gid, (uint32_t) chain, key);
        accept = sure || u < acceptP;
        n_u += sure ? 0ull : 1ull;
        cJS = accept ? ppJS : pcJS;
        acc += accept ? 1 : 0;
        rej += accept ? 0 : 1;
      }
      if (accept) {
        alpha = alphaN;
        cur = nw;
        thr_state = 0;
      }

      // ---- record (miso.c:882-893) ----------------------------------------------------
      if (m >= P.burn_in) {
        if (lagc == P.lag - 1) {
          if (n_rec < S_total && live) {
            const long long col = (long long) n_rec * P.n_chains + chain;
            if (mi < K) P.samples[d.sample_off + col * K + mi] = cur.psi;
            if (mi == 0) P.loglik[d.loglik_off + col] = cJS;
            if (P.samples_host) {
              if (mi < K) P.samples_host[d.sample_off + col * K + mi] = cur.psi;
              if (mi == 0) P.loglik_host[d.loglik_off + col] = cJS;
            }
          }
          n_rec++;
          lagc = 0;
        } else {
          lagc++;
        }
      }

      // ---- reassign (miso.c:895-898; for m == -1 the initial assignment) ---------------
      {
        const int m_next = m + 1;
        const bool last = m_next >= P.n_iters;
        const bool rec_mine = rp_always || (paired && m_next >= P.burn_in && lagc == P.lag - 1);
        const bool rec_next = __any_sync(kFull, rec_mine);
        double psi_r[K];
        bool nan = false;
        double pmin = 1.0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          psi_r[k] = shfl_d(cur.psi, gb + k);
          pmin = psi_r[k] < pmin ? psi_r[k] : pmin;       // a NaN psi never lowers pmin ...
          nan = nan || !(psi_r[k] == psi_r[k]);           // ... so flag it here
        }
        // fast rule valid for every read of this pass (dense_pass.cuh) and not the pass
        // that has to emit chain 0's per-read assignment
        bool fast = !nan && (pmin * P.ptab_min >= 1e-290) && !(last && ass_out);
        const bool need = fast && thr_state == 0;
        if (__any_sync(kFull, need)) {
          const uint32_t declined = quad_thr_update<K>(need, cr, ptab, cur.psi);
          if (need) thr_state = ((declined >> grp) & 1u) ? 2 : 1;
        }
        fast = fast && thr_state == 1;
        const int o = (int) (n_u & 3ull);
        const int my_steps = fast ? (((R2 + o + 3) >> 2) + 7) >> 3 : 0;
        const uint32_t a0 = fast ? rows + 4u * (uint32_t) mi : a_end;
        int c = 0;
        if (!(rec_next && !last && paired)) {
          int cnt[K];
          double unused;
          quad_pass<K, 0, false>(a0, a_end, my_steps, ucode, row_last, cr, ptab, n_u, R2, gid, (uint32_t) chain, key,
                                 g_always, P.neglog, P.n_neglog, cnt, unused);
#pragma unroll
          for (int k = 0; k < K; k++) c = (mi == k) ? cnt[k] : c;
        } else {
          quad_pass_rp<K, WIDE>(a0, a_end, my_steps, ucode, row_last, cr, ptab, n_u, R2, gid, (uint32_t) chain, key,
                                g_always, P.neglog, P.n_neglog, &c, &rp_drawn);
        }
        if (!fast) {     // final pass of chain 0, thresholds declined, or weights that underflow
          quad_literal<K, WIDE>(gmask, rows, ucode, cr, ptab, cur.psi, n_u, R2, gid, (uint32_t) chain, key, paired,
                                P.neglog, P.n_neglog, &c, &rp_drawn, last ? ass_out : nullptr);
        }
        __syncwarp();
        cnt_k = c + nfix_k;
        n_u += (unsigned long long) R2;
        have_rp = (rec_next && !last && paired) || !fast;
      }
    }

#ifdef MISOB200_SEG_DEBUG
    if (lane == 0) atomicAdd(&g_seg_dbg[4 + (seg < 3 ? seg : 3)], (unsigned long long) (clock64() - dbg_t0));
#endif
    if (m_end >= P.n_iters) {
      if (mi == 0 && live) {
        int *ar = P.accrej + ((long long) gi * P.n_chains + chain) * 2;
        ar[0] = acc; ar[1] = rej;
      }
    } else if (live) {       // hand the chain over to whoever holds the next segment's ticket
      if (mi < len) st->alpha[mi] = alpha;
      if (mi < K) st->cnt[mi] = cnt_k;
      if (mi == 0) {
        st->rp_drawn = rp_drawn; st->n_u = n_u;
        st->lagc = lagc; st->n_rec = n_rec; st->acc = acc; st->rej = rej;
        st->have_rp = have_rp ? 1 : 0;
      }
      if (mi == 0) *progress = seg + 1;
    }
    if (m_end < P.n_iters) {
      __syncwarp();
      if (lane == 0) ring_push(P.ring, P.ring_tail, unit);
    }
  }
}

}  // namespace misob200
