// miso_b200/csrc/plan.cpp -- setup stage on the host (integer work), see plan.hpp.
//
// Design notes (this is not a translation of the reference's setup code):
//   * the compatibility matrix is never materialised as doubles; a read x
//     isoform cell is one small integer code: 0 incompatible, SE 1, PE
//     fragment_length - fragment_start + 1.  The probability the reference
//     stores in its match matrix (src/solve.c:196-197) is ptab[code].
//   * reads that can never draw (0 or 1 compatible isoform, src/miso.c:65-68)
//     are folded into per-isoform constants; only the R2 reads that draw are
//     shipped, as (K+1) byte rows in draw order (row K = flags).
//   * the draw order must reproduce the reference's unstable sort exactly
//     (which read receives the n-th uniform of a pass depends on it), so the
//     index sort below follows the same published algorithm, Bentley &
//     McIlroy's "Engineering a Sort Function" (the reference runs it from
//     src/qsort.c through include/matrix.pmt:579-589).
#include "plan.hpp"
#include "match_core.hpp"
#include "bm_sort.hpp"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>

namespace misob200 {

namespace {

struct ColCmp {           // include/matrix.pmt:546-561 on ptab[code]
  const int32_t *codes; const double *ptab; int K;
  int operator()(int32_t a, int32_t b) const {
    const int32_t *x = codes + (size_t) a * K, *y = codes + (size_t) b * K;
    for (int i = 0; i < K; i++) {
      const double p = ptab[x[i]], q = ptab[y[i]];
      if (p < q) return -1;
      if (p > q) return 1;
    }
    return 0;
  }
};

static std::atomic<long long> g_prof[8];
struct ProfT {
  int slot; std::chrono::steady_clock::time_point t0;
  explicit ProfT(int s) : slot(s), t0(std::chrono::steady_clock::now()) {}
  void next(int s) { auto t = std::chrono::steady_clock::now(); g_prof[slot] += std::chrono::duration_cast<std::chrono::nanoseconds>(t - t0).count(); slot = s; t0 = t; }
  ~ProfT() { next(slot); }
};
struct GeneOut {
  GeneHost h;
  GeneDesc d;
  std::vector<uint8_t> tile;
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

void build_gene(const Plan &plan, const misob200_reads_t &in, int g, GeneOut &out, const DeviceCodes *pre) {
  GeneHost &h = out.h;
  GeneDesc &d = out.d;
  std::memset(&d, 0, sizeof(d));
  const int iso0 = in.iso_off[g], K = in.iso_off[g + 1] - iso0;
  const long long r0 = in.read_off[g], nr = in.read_off[g + 1] - r0;
  const int paired = in.paired ? 1 : 0;
  const int R = (int) (paired ? nr / 2 : nr);
  h.K = K; h.R = R;
  d.K = K; d.paired = paired;
  d.gene_id = in.gene_id ? in.gene_id[g] : (unsigned) g;

  int overhang = in.overhang == 0 ? 1 : in.overhang;
  if (K < 2) { h.status = MISOB200_EINVAL; return; }
  if (K > kMaxIso) { h.status = MISOB200_UNIMPLEMENTED; return; }
  // src/miso.c:690-694
  if (overhang < 1 || overhang >= in.read_len / 2) { h.status = MISOB200_EINVAL; return; }
  const int n_codes = (int) plan.ptab.size();
  if (n_codes > kMaxCodes) { h.status = MISOB200_UNIMPLEMENTED; return; }   // insert model wider than ptab in shared memory

  IsoView gv{K, in.exon_off + iso0, in.exon_start, in.exon_end};
  int isolen[kMaxIso], nex[kMaxIso];
  for (int k = 0; k < K; k++) {
    isolen[k] = 0; nex[k] = gv.exon_off[k + 1] - gv.exon_off[k];
    for (int e = gv.exon_off[k]; e < gv.exon_off[k + 1]; e++)
      isolen[k] += gv.ex_end[e] - gv.ex_start[e] + 1;
  }

  ProfT prof(0);
  // ---- compatibility codes, K x R column-major --------------------------
  // match_core.hpp; either computed here, or taken from the device kernel (match.cu), which
  // compiles the same functions
  std::vector<int32_t> codes((size_t) K * (R > 0 ? R : 1), 0);
  if (pre) {
    if (pre->status[g]) { h.status = pre->status[g]; return; }
    if (pre->codes8) {
      const uint8_t *src = pre->codes8 + pre->code_off[g];
      for (size_t i = 0; i < (size_t) K * R; i++) codes[i] = src[i];
    } else {
      const uint16_t *src = pre->codes16 + pre->code_off[g];
      for (size_t i = 0; i < (size_t) K * R; i++) codes[i] = src[i];
    }
  } else {
    const MatchParams mp{in.read_len, overhang, paired, plan.frag_start, plan.frag_len_n};
    for (int r = 0; r < R; r++)
      if (match_read(gv, mp, in.position + r0, in.cigar_off + r0, in.cigar, r, codes.data() + (size_t) r * K)) {
        h.status = MISOB200_EINVAL;
        return;
      }
  }
  // odd trailing mate of a paired batch is ignored, as noreads/2 does (solve.c:187)

  prof.next(1);
  // ---- draw order ---------------------------------------------------------
  std::vector<int32_t> order(R);
  const int32_t *dev_order = (pre && pre->order && R > 0) ? pre->order + pre->pair_off[g] : nullptr;
  if (dev_order && dev_order[0] >= 0) {      // sorted by order_kernel (match.cu): the same code, bm_sort.hpp
    std::memcpy(order.data(), dev_order, (size_t) R * sizeof(int32_t));
  } else {
  for (int r = 0; r < R; r++) order[r] = r;
  if ((int) plan.code_rank.size() == n_codes && !std::getenv("MISOB200_SORT_DOUBLES")) {
    // one integer key per read: the ranks of its K probabilities, isoform 0 most significant.  The
    // comparison only sees the order of the keys, so the narrowest field that holds every rank does
    // (8 bits for the usual <= 256 distinct fragment-length probabilities: a 64-bit key for any K)
    int max_rank = 0;                      // (dense ranks: the largest is the last distinct value's)
    for (uint16_t v : plan.code_rank) max_rank = std::max<int>(max_rank, v);
    if (max_rank < 256) {
      std::vector<uint64_t> key(R > 0 ? R : 1);
      for (int r = 0; r < R; r++) {
        uint64_t v = 0;
        for (int k = 0; k < K; k++) v = (v << 8) | plan.code_rank[codes[(size_t) r * K + k]];
        key[r] = v;
      }
      BMSort<KeyCmp<uint64_t>> sorter{KeyCmp<uint64_t>{key.data()}};
      sorter.sort(order.data(), R);
    } else {
      std::vector<unsigned __int128> key(R > 0 ? R : 1);
      for (int r = 0; r < R; r++) {
        unsigned __int128 v = 0;
        for (int k = 0; k < K; k++) v = (v << 16) | plan.code_rank[codes[(size_t) r * K + k]];
        key[r] = v;
      }
      BMSort<KeyCmp<unsigned __int128>> sorter{KeyCmp<unsigned __int128>{key.data()}};
      sorter.sort(order.data(), R);
    }
  } else {
    BMSort<ColCmp> sorter{ColCmp{codes.data(), plan.ptab.data(), K}};
    sorter.sort(order.data(), R);
  }
  }

  prof.next(2);
  // ---- read classes: histogram over zero/non-zero patterns ---------------
  // Both reference tabulations list distinct patterns in ascending
  // lexicographic order with isoform 0 most significant
  // (miso_paired.c:576-619 for SE where codes are 0/1, :628-681 for PE).
  {
    std::vector<int> hist(1 << K, 0);
    for (int r = 0; r < R; r++) {
      int m = 0;
      for (int k = 0; k < K; k++) m = (m << 1) | (codes[(size_t) r * K + k] != 0);
      hist[m]++;
    }
    for (int m = 0; m < (1 << K); m++) {
      if (!hist[m]) continue;
      for (int k = 0; k < K; k++) h.class_templates.push_back((m >> (K - 1 - k)) & 1);
      h.class_counts.push_back(hist[m]);
    }
    h.ncls = (int) h.class_counts.size();
  }

  // ---- per-gene constants ---------------------------------------------------
  d.sigma = 0.2 / K / K;                                    // SIGMA, miso.c:328
  d.sd = (K - 1 == 1) ? d.sigma : std::sqrt(d.sigma);       // miso.c:188
  d.covar_const = std::pow(2 * M_PI * d.sigma, -0.5 * (K - 1));   // miso.c:101
  {
    double asum = 0.0, lsum = 0.0;
    for (int k = 0; k < K; k++) {
      const double a = in.hyper ? in.hyper[iso0 + k] : 1.0;
      asum += a; lsum += lgamma(a);
      d.hyper_m1[k] = a - 1.0;
    }
    d.lg_sum = lgamma(asum); d.lg_each = lsum;              // miso.c:177-178
  }
  for (int k = 0; k < K; k++) {
    if (!paired) {                                           // miso.c:777-784
      const int l = isolen[k] - in.read_len + 1 - 2 * (nex[k] - 1) * (overhang - 1);
      const int eff = l > 0 ? l : 0;
      d.rs_se[k] = -std::log((double) l);
      d.offset[k] = std::log((double) eff);
      d.L[k] = l;
    } else {                                                 // miso_paired.c:403-419
      d.L[k] = isolen[k] - plan.frag_start + 1 - 2 * (nex[k] - 1) * (overhang - 1);
      double acc = 0.0;
      for (int j = 0; j < plan.frag_len_n; j++) {
        const double lp = d.L[k] - j;
        if (lp > 0) acc += lp;
      }
      d.offset[k] = std::log(acc);
    }
  }
  if (paired) {
    // the read-score passes look -log(lp) up in a table, lp = L_k - (code - 1), code in [1, n_codes):
    // if every such lp is >= 1 the kernel can skip the range test (class_pass.cuh MODE 2)
    int lo = d.L[0], hi = d.L[0];
    for (int k = 0; k < K; k++) { lo = std::min(lo, d.L[k]); hi = std::max(hi, d.L[k]); }
    d.lp_min = lo - (n_codes - 2);
    d.lp_max = hi;
    d.lp_safe = (d.lp_min >= 1 && !std::getenv("MISOB200_NO_LP_SAFE")) ? 1 : 0;      // (switch: A/B runs)
  }
  auto read_score = [&](int k, int code) -> double {
    if (!paired) return d.rs_se[k];
    const double lp = d.L[k] - (code - 1);
    return -std::log(lp) + plan.ptab[code];                  // miso_paired.c:411
  };

  prof.next(3);
  // ---- split reads into fixed and drawn; pack the tile ------------------------
  h.fixed_ass.assign(R, -1);
  h.rank_read.clear();
  std::vector<int> drawn;
  drawn.reserve(R);
  for (int i = 0; i < R; i++) {
    const int r = order[i];
    const int32_t *col = codes.data() + (size_t) r * K;
    int nv = 0, last = -1;
    for (int k = 0; k < K; k++) if (col[k]) { nv++; last = k; }
    if (nv == 0) continue;
    if (nv == 1) {
      h.fixed_ass[r] = (int8_t) last;
      d.n_fixed[last]++;
      const double s = read_score(last, col[last]);
      d.rp_fixed += s;
      if (!std::isfinite(s)) d.rp_always = 1;
      continue;
    }
    h.fixed_ass[r] = -2;
    drawn.push_back(r);
    // a score is finite unless its length term is log of a non-positive number
    // (ptab entries are finite): no need to evaluate the logarithm to know
    for (int k = 0; k < K; k++)
      if (col[k] && (paired ? d.L[k] - (col[k] - 1) <= 0 : d.L[k] <= 0)) d.rp_always = 1;
  }
  const int R2 = (int) drawn.size();
  h.R2 = R2; d.R2 = R2;
  h.rank_read.assign(drawn.begin(), drawn.end());
  // ---- weight classes of the drawing reads ------------------------------------
  // Two reads whose weight vectors (psi_k * p_k)_k are proportional make the same
  // choice for the same uniform whatever psi is.  Reads whose compatible isoforms all
  // see the same fragment length (the common case) are proportional to the 0/1
  // pattern itself: their class key is the pattern, with the ptab index of 1.0
  // (n_codes, appended on upload) in place of the code.  Other reads key on their
  // full code vector.  class_kernel.cuh turns each class into K-1 integer thresholds
  // on the raw Philox word once per iteration.
  prof.next(4);
  const int one_idx = n_codes;
  std::vector<uint8_t> cls_id(R2);
  std::vector<uint16_t> ucode(R2, 0);
  std::vector<std::array<uint16_t, kMaxIso>> cls_keys;
  std::vector<int> cls_size;
  bool class_ok = plan.force_format != 0;
  if (class_ok) {
    std::unordered_map<std::string, int> seen;
    std::string key(2 * kMaxIso, '\0');
    std::array<uint16_t, kMaxIso> prev{};
    int prev_id = -1;
    for (int i = 0; i < R2 && class_ok; i++) {
      const int32_t *col = codes.data() + (size_t) drawn[i] * K;
      int32_t common = 0;
      bool uniform = true;
      for (int k = 0; k < K; k++) {
        if (!col[k]) continue;
        if (!common) common = col[k];
        else if (col[k] != common) uniform = false;
      }
      std::array<uint16_t, kMaxIso> v{};
      for (int k = 0; k < K; k++) v[k] = (uint16_t) (col[k] ? (uniform ? one_idx : col[k]) : 0);
      ucode[i] = uniform ? (uint16_t) common : 0;
      int id;
      if (prev_id >= 0 && v == prev) {      // the draw order is sorted by column: runs of one class
        id = prev_id;
      } else {
        std::memcpy(&key[0], v.data(), 2 * kMaxIso);
        auto it = seen.find(key);
        if (it == seen.end()) {
          id = (int) cls_keys.size();
          if (id >= kMaxClasses) { class_ok = false; break; }
          seen.emplace(key, id);
          cls_keys.push_back(v);
          cls_size.push_back(0);
        } else {
          id = it->second;
        }
        prev = v; prev_id = id;
      }
      cls_id[i] = (uint8_t) id;
      cls_size[id]++;
    }
  }

  if (class_ok) {
    // number the classes by falling read count: class c's threshold row lives in the 16-byte
    // bank group c mod 8 of shared memory (class_pass.cuh), so the classes most lanes ask for
    // in the same LDS should sit in different groups
    const int n = (int) cls_keys.size();
    std::vector<int> by(n), to(n);
    for (int c = 0; c < n; c++) by[c] = c;
    std::stable_sort(by.begin(), by.end(), [&](int a, int b) { return cls_size[a] > cls_size[b]; });
    std::vector<std::array<uint16_t, kMaxIso>> keys2(n);
    std::vector<int> size2(n);
    for (int c = 0; c < n; c++) { to[by[c]] = c; keys2[c] = cls_keys[by[c]]; size2[c] = cls_size[by[c]]; }
    cls_keys.swap(keys2);
    cls_size.swap(size2);
    for (int i = 0; i < R2; i++) cls_id[i] = (uint8_t) to[cls_id[i]];
  }

  prof.next(5);
  const int cb = plan.wide ? 2 : 1;
  const int padded = round_up(R2 + kTilePadFront, 128);
  if (class_ok) {
    // class tile: id row (bytes; the padding carries the null class id ncls), class records:
    // ncls x 8 ptab indices (u16), then ncls meta words (bits 0-7 first compatible isoform,
    // bit 8 uniform-code class), one more for the null class -- these two are the "core" every
    // pass reads -- then the uniform-code row (bytes, or 16-bit when the insert model has more
    // than 255 fragment lengths).
    const int ncls = (int) cls_keys.size();
    d.format = 1;
    d.ncls = ncls;
    d.row_bytes = padded + 16;
    d.cls_off = d.row_bytes;
    d.core_bytes = d.cls_off + ncls * 16 + round_up((ncls + 1) * 4, 16);
    d.ucode_off = d.core_bytes;      // last: only the read-score passes and the literal rule read it
    d.flag_off = 0;
    d.tile_bytes = d.ucode_off + padded * cb + 16;
    out.tile.assign((size_t) d.tile_bytes, 0);
    std::memset(out.tile.data(), ncls, (size_t) d.row_bytes);
    uint8_t *uc = out.tile.data() + d.ucode_off;
    for (int i = 0; i < R2; i++) {
      out.tile[kTilePadFront + i] = cls_id[i];
      if (plan.wide) {
        uc[2 * (kTilePadFront + i)] = (uint8_t) (ucode[i] & 0xff);
        uc[2 * (kTilePadFront + i) + 1] = (uint8_t) (ucode[i] >> 8);
      } else {
        uc[kTilePadFront + i] = (uint8_t) ucode[i];
      }
    }
    uint16_t *rec = reinterpret_cast<uint16_t *>(out.tile.data() + d.cls_off);
    uint32_t *meta = reinterpret_cast<uint32_t *>(out.tile.data() + d.cls_off + ncls * 16);
    for (int c = 0; c < ncls; c++) {
      int first = -1;
      bool uniform = false;
      for (int k = 0; k < kMaxIso; k++) {
        rec[c * 8 + k] = cls_keys[c][k];
        if (cls_keys[c][k] && first < 0) first = k;
        if (cls_keys[c][k] == one_idx) uniform = true;
      }
      meta[c] = (uint32_t) first | (uniform ? 0x100u : 0u);
      for (int k = 0; k < first; k++) d.g_always[k] += cls_size[c];
    }
    meta[ncls] = 0x100u;      // null class of the padding
  } else {
    // dense tile: per isoform a code row -- 3 pad elements, R2 codes, zero fill to a whole
    // number of 128-read warp steps plus 16 spare bytes (a lane reads a little past its 4
    // reads); elements are bytes, or 16-bit when the insert model has more than 255
    // fragment lengths.  The flag row (1 = exactly two compatible isoforms) is always bytes.
    const int row_bytes = padded * cb + 16;
    const int flag_row = padded + 16;
    d.format = 0;
    d.row_bytes = row_bytes;
    d.flag_off = row_bytes * K;
    d.tile_bytes = row_bytes * K + flag_row;
    out.tile.assign((size_t) d.tile_bytes, 0);
    for (int i = 0; i < R2; i++) {
      const int32_t *col = codes.data() + (size_t) drawn[i] * K;
      int nv = 0;
      for (int k = 0; k < K; k++) {
        uint8_t *row = out.tile.data() + (size_t) k * row_bytes;
        if (plan.wide) {
          row[2 * (kTilePadFront + i)] = (uint8_t) (col[k] & 0xff);
          row[2 * (kTilePadFront + i) + 1] = (uint8_t) (col[k] >> 8);
        } else {
          row[kTilePadFront + i] = (uint8_t) col[k];
        }
        nv += col[k] != 0;
      }
      out.tile[(size_t) d.flag_off + kTilePadFront + i] = (nv == 2) ? 1 : 0;
    }
  }
  if (plan.keep_match) { h.codes = std::move(codes); h.order = std::move(order); }
}

// discretised normal insert-length table, normalised
// (src/simulator.c:198-219, src/util.c:17-32, src/miso_paired.c:303-307)
void fragment_table(Plan &plan) {
  const double sd = std::sqrt(plan.frag_var);
  int fs = (int) (plan.frag_mean - sd * plan.num_devs);
  int fe = (int) (plan.frag_mean + sd * plan.num_devs);
  if (fs < plan.read_len) fs = plan.read_len;
  if (fe < fs) fe = fs;
  const int n = fe - fs + 1;
  std::vector<double> p(n);
  double sum = 0.0;
  for (int i = fs, j = 0; i <= fe; i++, j++) {
    const double x = (i - plan.frag_mean) / sd;
    p[j] = 0.398942280401432677939946059934 * std::exp(-0.5 * x * x) / sd;
  }
  for (int j = 0; j < n; j++) sum += p[j];
  const double by = 1.0 / sum;
  plan.ptab.assign(n + 1, 0.0);
  for (int j = 0; j < n; j++) plan.ptab[j + 1] = p[j] * by;
  plan.frag_start = fs; plan.frag_len_n = n;
}

}  // namespace

// Validation + the plan's library (fragment table, code ranks): what an append fixes before any read is looked at.
static int prepare_library(Plan &plan, const misob200_reads_t &in) {
  if (in.n_genes < 0 || !in.iso_off || !in.exon_off || !in.read_off) {
    set_error("plan_append: null or negative input"); return MISOB200_EINVAL;
  }
  const int paired = in.paired ? 1 : 0;
  // validate first; the plan's library fields are only committed by a batch that passes
  if (in.read_len < 0) { set_error("plan_append: negative read length"); return MISOB200_EINVAL; }
  if (paired && !(in.frag_var > 0)) { set_error("plan_append: frag_var must be positive"); return MISOB200_EINVAL; }
  if (paired && !(in.num_devs >= 0)) { set_error("plan_append: num_devs must not be negative"); return MISOB200_EINVAL; }
  if (plan.paired < 0) {
    plan.read_len = in.read_len; plan.overhang = in.overhang;
    plan.frag_mean = in.frag_mean; plan.frag_var = in.frag_var; plan.num_devs = in.num_devs;
    if (paired) {
      fragment_table(plan);
      plan.wide = plan.ptab.size() > 256;
    } else {
      plan.ptab = {0.0, 1.0}; plan.frag_start = 0; plan.frag_len_n = 1;
    }
    plan.paired = paired;
  } else if (plan.paired != paired || plan.read_len != in.read_len || plan.overhang != in.overhang ||
             (paired && (plan.frag_mean != in.frag_mean || plan.frag_var != in.frag_var ||
                         plan.num_devs != in.num_devs))) {
    set_error("plan_append: a plan holds one library (read length, overhang, insert model)");
    return MISOB200_EINVAL;
  }
  if (plan.code_rank.size() != plan.ptab.size()) {
    // dense rank of every code's probability (KeyCmp)
    const int n = (int) plan.ptab.size();
    std::vector<int> by(n);
    for (int i = 0; i < n; i++) by[i] = i;
    std::stable_sort(by.begin(), by.end(), [&](int a, int b) { return plan.ptab[a] < plan.ptab[b]; });
    plan.code_rank.assign(n, 0);
    int rank = 0;
    for (int i = 0; i < n; i++) {
      if (i && plan.ptab[by[i]] != plan.ptab[by[i - 1]]) rank++;
      plan.code_rank[by[i]] = (uint16_t) rank;
    }
    if (n > 65535) plan.code_rank.clear();      // (never: kMaxCodes)
  }
  return 0;
}

// the host half of an append: per-gene classes, constants and tiles on the worker threads, then the merge
static int build_and_merge(Plan &plan, const misob200_reads_t &in, int n_threads, const DeviceCodes *pre) {
  const int G = in.n_genes;
  std::vector<GeneOut> outs(G);
  std::atomic<int> next(0);
  int nt = n_threads > 0 ? n_threads : host_threads();
  if (nt < 1) nt = 1;
  if (nt > G) nt = G > 0 ? G : 1;
  auto work = [&]() {
    for (int g; (g = next.fetch_add(1)) < G;) build_gene(plan, in, g, outs[g], pre);
  };
  if (nt == 1) work();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  auto t_merge = std::chrono::steady_clock::now();
  // merge: offsets in gene order (serial, a few words per gene), then the tile bytes copied into the
  // arena by the worker threads -- one growth of the arena per append instead of one per gene
  size_t tile_at = plan.tiles.size();
  plan.desc.reserve(plan.desc.size() + (size_t) G);
  plan.host.reserve(plan.host.size() + (size_t) G);
  for (int g = 0; g < G; g++) {
    GeneOut &o = outs[g];
    o.h.read_base = plan.n_reads;
    o.d.status = o.h.status;
    o.d.tile_off = tile_at;
    o.d.drawn_off = plan.n_drawn;
    tile_at += o.tile.size();
    plan.n_reads += o.h.R;
    plan.n_drawn += (o.h.R2 + 15) / 16 * 16;
    plan.desc.push_back(o.d);
  }
  plan.tiles.resize(tile_at);
  const size_t first_desc = plan.desc.size() - (size_t) G;
  std::atomic<int> next_copy(0);
  auto copy = [&]() {
    for (int g0; (g0 = next_copy.fetch_add(256)) < G;)
      for (int g = g0; g < std::min(G, g0 + 256); g++) {
        GeneOut &o = outs[g];
        if (!o.tile.empty()) std::memcpy(plan.tiles.data() + plan.desc[first_desc + g].tile_off, o.tile.data(), o.tile.size());
        std::vector<uint8_t>().swap(o.tile);
      }
  };
  if (nt == 1 || G < 1024) copy();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(copy);
    for (auto &t : pool) t.join();
  }
  for (int g = 0; g < G; g++) plan.host.push_back(std::move(outs[g].h));
  if (std::getenv("MISOB200_PLAN_PROFILE")) {
    const double merge = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_merge).count();
    fprintf(stderr, "[plan profile] thread-seconds: match %.3f sort %.3f classes %.3f consts+split %.3f weight classes %.3f tile %.3f; merge (serial) %.3f s\n",
            g_prof[0] / 1e9, g_prof[1] / 1e9, g_prof[2] / 1e9, g_prof[3] / 1e9, g_prof[4] / 1e9, g_prof[5] / 1e9, merge);
    for (auto &x : g_prof) x = 0;
  }
  return 0;
}

struct PendingAppend {
  DeviceCodes codes;
  const misob200_reads_t *reads = nullptr;
  misob200_reads_t copy{};          // the struct itself (pointers + scalars); the arrays stay the caller's
};

int plan_append_device_begin(Plan &plan, const misob200_reads_t &in, int device, PendingAppend **pending) {
  if (!pending) return MISOB200_EINVAL;
  *pending = nullptr;
  if (int rc = prepare_library(plan, in)) return rc;
  PendingAppend *p = new PendingAppend();
  p->copy = in;
  const MatchParams mp{in.read_len, in.overhang == 0 ? 1 : in.overhang, in.paired ? 1 : 0, plan.frag_start, plan.frag_len_n};
  static const std::vector<uint16_t> no_rank;
  const int rc = stage_submit(in, mp, device, (int) plan.ptab.size(),
                              std::getenv("MISOB200_HOST_SORT") ? no_rank : plan.code_rank, p->codes);
  if (rc) { delete p; return rc; }
  *pending = p;
  return 0;
}

int plan_append_device_finish(Plan &plan, PendingAppend *p, int n_threads) {
  if (!p) { set_error("plan_append_device_finish: nothing pending"); return MISOB200_EINVAL; }
  int rc = stage_wait(p->codes);
  if (!rc) rc = build_and_merge(plan, p->copy, n_threads, &p->codes);
  stage_release(p->codes.stage);
  delete p;
  return rc;
}

int plan_append(Plan &plan, const misob200_reads_t &in, int n_threads, int match_device) {
  if (match_device >= 0) {      // optional: read <-> isoform compatibility and draw order on the GPU
    PendingAppend *p = nullptr;
    if (int rc = plan_append_device_begin(plan, in, match_device, &p)) return rc;
    return plan_append_device_finish(plan, p, n_threads);
  }
  if (int rc = prepare_library(plan, in)) return rc;
  return build_and_merge(plan, in, n_threads, nullptr);
}

}  // namespace misob200
