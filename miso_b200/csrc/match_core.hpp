// miso_b200/csrc/match_core.hpp -- read <-> isoform compatibility, one statement for the
// host plan stage (plan.cpp) and the device kernel (match.cu).
//
// What it computes is the reference's splicing_matchIso / splicing_matchIso_paired
// (/root/reference/pysplicing/src/solve.c:8-108, :141-218) with splicing_parse_cigar
// (:220-306) and the isoform coordinate of splicing_genomic_to_iso (src/gff.c:1041-1084),
// as one small integer code per (read, isoform): 0 incompatible, single-end 1, paired-end
// fragment_length - fragment_start + 1.  Integer work only; the two users compile the same
// functions, so the GPU and the host plan stage cannot disagree.
#pragma once
#include <cstdint>

#include "../../include/miso_b200.h"

#ifdef __CUDACC__
#define MISOB200_HD __host__ __device__ __forceinline__
#else
#define MISOB200_HD inline
#endif

namespace misob200 {

struct IsoView {          // one gene's isoforms
  int K;
  const int32_t *exon_off;   // K+1 entries, absolute
  const int32_t *ex_start, *ex_end;
};

// ---- CIGAR ------------------------------------------------------------
// Semantics of splicing_parse_cigar (src/solve.c:220-306): M = X S H D are
// match-like and clipped to read_len in total, N is an intron (negative),
// I is skipped, S/H only at the ends, anything else is an error.
constexpr int kMaxCigarOps = 64;
struct Cigar {
  int n = 0, len = 0;
  int op[kMaxCigarOps];
};

// strtol(s, &end, 10) as the reference calls it: optional white space, optional sign, digits;
// no digits -> 0 and end = s.  (Saturates at +-2^62 instead of LONG_MAX: lengths that large
// are clipped to read_len or make every isoform incompatible either way.)
MISOB200_HD long long parse_long(const char *s, const char **end) {
  const char *p = s;
  while (*p == ' ' || (*p >= '\t' && *p <= '\r')) p++;
  bool neg = false;
  if (*p == '+' || *p == '-') { neg = *p == '-'; p++; }
  if (*p < '0' || *p > '9') { *end = s; return 0; }
  long long v = 0;
  while (*p >= '0' && *p <= '9') {
    if (v < (1LL << 58)) v = v * 10 + (*p - '0');
    p++;
  }
  *end = p;
  return neg ? -v : v;
}

MISOB200_HD int parse_cigar(const char *s, int read_len, Cigar &out) {
  int mode = 0;
  out.n = 0; out.len = 0;
  while (*s) {
    const char *end;
    long long l = parse_long(s, &end);
    const char c = *end;
    const bool clip = (c == 'S' || c == 'H');
    if (mode == 0 && !clip) mode = 1;
    else if (mode == 1 && clip) mode = 2;
    else if (mode == 2 && !clip) return MISOB200_EINVAL;
    if (c == 'M' || c == '=' || c == 'X' || clip || c == 'D') {
      if (read_len > 0 && out.len + l > read_len) l = read_len - out.len;
      if (out.n >= kMaxCigarOps) return MISOB200_EINVAL;
      out.op[out.n++] = (int) l;
      out.len += (int) l;
    } else if (c == 'N') {
      if (out.n >= kMaxCigarOps) return MISOB200_EINVAL;
      out.op[out.n++] = (int) -l;
    } else if (c == 'I') {
      // not on the genome: nothing to do
    } else {
      return MISOB200_EINVAL;   // also hit by a trailing number without a letter
    }
    s = end + 1;
  }
  return 0;
}

// 1 if the read's blocks tile isoform k's exons from pos (src/solve.c:65-95)
MISOB200_HD int compatible(const IsoView &g, int k, int pos, const Cigar &cg) {
  int ex = g.exon_off[k];
  const int ex_hi = g.exon_off[k + 1];
  while (ex < ex_hi && (pos < g.ex_start[ex] || g.ex_end[ex] < pos)) ex++;
  if (ex >= ex_hi) return 0;
  for (int c = 0; c < cg.n; c++) {
    const int o = cg.op[c];
    if (o > 0) {
      if (pos + o - 1 > g.ex_end[ex]) return 0;
      pos += o;
    } else {
      if (pos != g.ex_end[ex] + 1) return 0;
      pos -= o;
      ex++;
      if (ex >= ex_hi || pos != g.ex_start[ex]) return 0;
    }
  }
  return 1;
}

// position on the spliced isoform, 1-based, or -1 (src/gff.c:855-900,1041-1084)
MISOB200_HD int iso_coordinate(const IsoView &g, int k, int pos) {
  int before = 0;
  for (int ex = g.exon_off[k]; ex < g.exon_off[k + 1]; ex++) {
    if (g.ex_end[ex] < pos) { before += g.ex_end[ex] - g.ex_start[ex] + 1; continue; }
    if (g.ex_start[ex] <= pos) return pos - g.ex_start[ex] + 1 + before;
    return -1;
  }
  return -1;
}

struct MatchParams {      // what a read's code depends on besides the gene and the read
  int read_len, overhang, paired;
  int frag_start, frag_len_n;
};

// usable alignment: long enough, first and last block at least `overhang` (solve.c:55-61)
MISOB200_HD bool cigar_usable(const Cigar &cg, int read_len, int overhang) {
  return !(cg.n == 0 || cg.len < read_len || cg.op[0] < overhang || cg.op[cg.n - 1] < overhang);
}

// Codes of read r of a gene (single-end: read r; paired-end: mates 2r, 2r+1) against its K
// isoforms.  `position` / `cigar_off` point at the gene's first read.  col[k] receives the
// code.  Returns 0 or MISOB200_EINVAL (unparsable CIGAR: the reference aborts the whole call,
// solve.c:295-298; here the gene gets that status).
template <class Code, class Off>
MISOB200_HD int match_read(const IsoView &gv, const MatchParams &mp, const int32_t *position,
                           const Off *cigar_off, const char *cigar, int r, Code *col) {
  const int K = gv.K;
  for (int k = 0; k < K; k++) col[k] = 0;
  Cigar cg;
  if (!mp.paired) {
    if (parse_cigar(cigar + cigar_off[r], mp.read_len, cg)) return MISOB200_EINVAL;
    if (!cigar_usable(cg, mp.read_len, mp.overhang)) return 0;
    for (int k = 0; k < K; k++) col[k] = (Code) compatible(gv, k, position[r], cg);
    return 0;
  }
  Cigar cg2;
  const int r1 = 2 * r, r2 = r1 + 1;
  if (parse_cigar(cigar + cigar_off[r1], mp.read_len, cg) || parse_cigar(cigar + cigar_off[r2], mp.read_len, cg2))
    return MISOB200_EINVAL;
  if (!cigar_usable(cg, mp.read_len, mp.overhang) || !cigar_usable(cg2, mp.read_len, mp.overhang)) return 0;
  const int p1 = position[r1], p2 = position[r2];
  for (int k = 0; k < K; k++) {
    if (!compatible(gv, k, p1, cg) || !compatible(gv, k, p2, cg2)) continue;
    // src/solve.c:190-198
    const int frag = iso_coordinate(gv, k, p2) - iso_coordinate(gv, k, p1) + mp.read_len;
    if (frag < mp.frag_start || frag >= mp.frag_len_n + mp.frag_start) continue;
    col[k] = (Code) (frag - mp.frag_start + 1);
  }
  return 0;
}

}  // namespace misob200
