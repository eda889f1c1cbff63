// workloads/synth.cpp -- synthetic workloads for tests and bench.py (libmiso_synth.so).
//
// Bench / test infrastructure, NOT part of the product library: both arms of bench.py
// (ours and --impl reference) draw their inputs from here, and the reference arm maps
// nothing of libmiso_b200.so.
//
// Generates the inputs BASELINE.json's configs name:
//   kind 0  cfg-2: skipped-exon events, K = 2, single-end reads
//   kind 1  cfg-3: K ~ U{2..8} isoforms (isoform 0 = all K+1 exons of 200 nt,
//           isoform k skips exon k), paired-end reads with a discretised normal
//           insert-length model
// The read model is the one the reference's simulators implement
// (/root/reference/pysplicing/src/simulator.c:68-196, :221-442): isoform chosen
// in proportion to psi times its number of start positions, start uniform on
// the isoform, CIGAR = the exon blocks the read covers.  This is an
// independent implementation with its own RNG (splitmix64 keyed by
// (seed, gene id)); it does not have to match the reference draw for draw --
// parity is checked downstream, on whatever reads come out.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/miso_synth.h"

namespace misob200 {

namespace {

thread_local std::string g_synth_error;
void set_error(const std::string &m) { g_synth_error = m; }

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double unif() { return ((next() >> 11) + 0.5) * 0x1p-53; }
  int below(int n) { return (int) (unif() * n); }     // 0 .. n-1
};

struct GeneBuf {
  std::vector<int32_t> exon_off, ex_start, ex_end, pos;
  std::vector<int32_t> cig_len;     // per read
  std::string cig;
  std::vector<double> psi;
  int K = 0;
};

// genomic start + CIGAR of a read occupying isoform coordinates [at, at+len-1]
void place_read(const std::vector<int> &es, const std::vector<int> &ee, int at, int len, GeneBuf &gb) {
  size_t e = 0;
  int before = 0;
  while (before + (ee[e] - es[e] + 1) < at) { before += ee[e] - es[e] + 1; e++; }
  int g = es[e] + (at - before - 1);
  gb.pos.push_back(g);
  char tmp[32];
  const size_t c0 = gb.cig.size();
  int left = len;
  while (true) {
    const int room = ee[e] - g + 1;
    if (left <= room) {
      snprintf(tmp, sizeof tmp, "%dM", left); gb.cig += tmp;
      break;
    }
    snprintf(tmp, sizeof tmp, "%dM%dN", room, es[e + 1] - ee[e] - 1); gb.cig += tmp;
    left -= room; e++; g = es[e];
  }
  gb.cig.push_back('\0');
  gb.cig_len.push_back((int32_t) (gb.cig.size() - c0));
}

}  // namespace

struct Workload {
  int kind = 0, paired = 0, read_len = 0;
  double frag_mean = 0, frag_var = 0, num_devs = 0;
  std::vector<int32_t> iso_off, exon_off, ex_start, ex_end, position;
  std::vector<int64_t> read_off, cigar_off;
  std::vector<char> cigar;
  std::vector<uint32_t> gene_id;
  std::vector<double> psi;
  std::vector<int64_t> psi_off;
};

int workload_create(int kind, int n_genes, int reads_per_gene, int read_len, double frag_mean,
                    double frag_var, double num_devs, uint64_t seed, uint32_t first_gene_id,
                    const uint32_t *gene_ids, uint32_t sample, int n_threads, Workload **out) {
  if (kind != 0 && kind != 1) { set_error("workload kind must be 0 (SE K=2) or 1 (PE mixed K)"); return MISOB200_EINVAL; }
  if (n_genes < 0 || reads_per_gene < 0 || read_len < 4) { set_error("workload: bad sizes"); return MISOB200_EINVAL; }
  Workload *w = new Workload();
  w->kind = kind; w->paired = kind == 1; w->read_len = read_len;
  w->frag_mean = frag_mean; w->frag_var = frag_var; w->num_devs = num_devs;

  // insert-length table (same construction as the sampler's, simulator.c:198-219)
  std::vector<double> fp;
  int fs = 0;
  if (kind == 1) {
    if (!(frag_var > 0)) { delete w; set_error("workload: frag_var must be positive"); return MISOB200_EINVAL; }
    const double sd = std::sqrt(frag_var);
    fs = (int) (frag_mean - sd * num_devs);
    int fe = (int) (frag_mean + sd * num_devs);
    if (fs < read_len) fs = read_len;
    if (fe < fs) fe = fs;
    fp.resize(fe - fs + 1);
    double sum = 0;
    for (size_t j = 0; j < fp.size(); j++) {
      const double x = ((fs + (int) j) - frag_mean) / sd;
      fp[j] = std::exp(-0.5 * x * x); sum += fp[j];
    }
    for (auto &p : fp) p /= sum;
  }

  std::vector<GeneBuf> bufs(n_genes);
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int g; (g = next.fetch_add(1)) < n_genes;) {
      GeneBuf &gb = bufs[g];
      const uint32_t gid = gene_ids ? gene_ids[g] : first_gene_id + (uint32_t) g;
      Rng rng(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (gid + 1));
      std::vector<std::vector<int>> iso;    // exon indices per isoform
      std::vector<int> xs, xe;              // exon table
      if (kind == 0) {
        const int mid = 50 + rng.below(251);
        xs = {1, 401, 801}; xe = {200, 400 + mid, 1000};
        iso = {{0, 1, 2}, {0, 2}};
      } else {
        const int K = 2 + rng.below(7);
        for (int i = 0; i <= K; i++) { xs.push_back(1 + 400 * i); xe.push_back(200 + 400 * i); }
        std::vector<int> all;
        for (int i = 0; i <= K; i++) all.push_back(i);
        iso.push_back(all);
        for (int k = 1; k < K; k++) {
          std::vector<int> v;
          for (int i = 0; i <= K; i++) if (i != k) v.push_back(i);
          iso.push_back(v);
        }
      }
      const int K = (int) iso.size();
      gb.K = K;
      // a second sample of the same events (cfg-5): same gene structures, its own psi and reads
      if (sample) rng = Rng((seed + 0x51ED270B7F4A7C15ull * sample) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * (gid + 1));
      // psi ~ Dirichlet(1,...,1)
      gb.psi.resize(K);
      double ps = 0;
      for (int k = 0; k < K; k++) { gb.psi[k] = -std::log(rng.unif()); ps += gb.psi[k]; }
      for (int k = 0; k < K; k++) gb.psi[k] /= ps;
      std::vector<std::vector<int>> ies(K), iee(K);
      std::vector<int> isolen(K, 0);
      gb.exon_off.push_back(0);
      for (int k = 0; k < K; k++) {
        for (int e : iso[k]) {
          ies[k].push_back(xs[e]); iee[k].push_back(xe[e]);
          gb.ex_start.push_back(xs[e]); gb.ex_end.push_back(xe[e]);
          isolen[k] += xe[e] - xs[e] + 1;
        }
        gb.exon_off.push_back((int32_t) gb.ex_start.size());
      }
      // isoform weights
      std::vector<double> wgt(K), cum(K);
      double tot = 0;
      for (int k = 0; k < K; k++) {
        double places = 0;
        if (kind == 0) places = std::max(isolen[k] - read_len + 1, 0);
        else for (size_t j = 0; j < fp.size(); j++) places += fp[j] * std::max(isolen[k] - (fs + (int) j) + 1, 0);
        wgt[k] = gb.psi[k] * places; tot += wgt[k]; cum[k] = tot;
      }
      for (int r = 0; r < reads_per_gene; r++) {
        const double u = rng.unif() * tot;
        int k = 0;
        while (k < K - 1 && u > cum[k]) k++;
        if (kind == 0) {
          const int at = 1 + rng.below(isolen[k] - read_len + 1);
          place_read(ies[k], iee[k], at, read_len, gb);
        } else {
          // fragment length in proportion to P(l) * number of placements
          double z = 0;
          for (size_t j = 0; j < fp.size(); j++) z += fp[j] * std::max(isolen[k] - (fs + (int) j) + 1, 0);
          double v = rng.unif() * z, c = 0;
          size_t j = 0;
          for (; j < fp.size(); j++) {
            c += fp[j] * std::max(isolen[k] - (fs + (int) j) + 1, 0);
            if (v <= c) break;
          }
          if (j >= fp.size()) j = fp.size() - 1;
          int fl = fs + (int) j;
          if (fl > isolen[k]) fl = isolen[k];                  // rounding fell off the end of the table
          const int at = 1 + rng.below(isolen[k] - fl + 1);
          place_read(ies[k], iee[k], at, read_len, gb);
          place_read(ies[k], iee[k], at + fl - read_len, read_len, gb);
        }
      }
    }
  };
  int nt = n_threads > 0 ? n_threads : (int) std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > n_genes) nt = n_genes > 0 ? n_genes : 1;
  if (nt == 1) work();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }

  w->iso_off.push_back(0); w->exon_off.push_back(0); w->read_off.push_back(0);
  w->cigar_off.push_back(0); w->psi_off.push_back(0);
  for (int g = 0; g < n_genes; g++) {
    GeneBuf &gb = bufs[g];
    const int32_t ebase = (int32_t) w->ex_start.size();
    for (int k = 0; k < gb.K; k++) w->exon_off.push_back(ebase + gb.exon_off[k + 1]);
    w->iso_off.push_back(w->iso_off.back() + gb.K);
    w->ex_start.insert(w->ex_start.end(), gb.ex_start.begin(), gb.ex_start.end());
    w->ex_end.insert(w->ex_end.end(), gb.ex_end.begin(), gb.ex_end.end());
    w->position.insert(w->position.end(), gb.pos.begin(), gb.pos.end());
    for (int32_t l : gb.cig_len) w->cigar_off.push_back(w->cigar_off.back() + l);
    w->cigar.insert(w->cigar.end(), gb.cig.begin(), gb.cig.end());
    w->read_off.push_back((int64_t) w->position.size());
    w->gene_id.push_back(gene_ids ? gene_ids[g] : first_gene_id + (uint32_t) g);
    w->psi.insert(w->psi.end(), gb.psi.begin(), gb.psi.end());
    w->psi_off.push_back((int64_t) w->psi.size());
    std::vector<int32_t>().swap(gb.exon_off);
    std::vector<int32_t>().swap(gb.pos);
    std::string().swap(gb.cig);
  }
  *out = w;
  return 0;
}

}  // namespace misob200

using misob200::Workload;
struct misob200_workload { Workload *w; };

extern "C" {

const char *misob200_workload_last_error(void) { return misob200::g_synth_error.c_str(); }

int misob200_workload_create(int kind, int32_t n_genes, int32_t reads_per_gene, int32_t read_len,
                             double frag_mean, double frag_var, double num_devs, uint64_t seed,
                             uint32_t first_gene_id, int n_threads, misob200_workload_t **out) {
  if (!out) return MISOB200_EINVAL;
  Workload *w = nullptr;
  int rc = misob200::workload_create(kind, n_genes, reads_per_gene, read_len, frag_mean, frag_var, num_devs,
                                     seed, first_gene_id, nullptr, 0, n_threads, &w);
  if (rc) return rc;
  *out = new misob200_workload{w};
  return 0;
}

int misob200_workload_create_ids(int kind, int32_t n_genes, const uint32_t *gene_ids, int32_t reads_per_gene,
                                 int32_t read_len, double frag_mean, double frag_var, double num_devs,
                                 uint64_t seed, uint32_t sample, int n_threads, misob200_workload_t **out) {
  if (!out || (n_genes > 0 && !gene_ids)) return MISOB200_EINVAL;
  Workload *w = nullptr;
  int rc = misob200::workload_create(kind, n_genes, reads_per_gene, read_len, frag_mean, frag_var, num_devs,
                                     seed, 0, gene_ids, sample, n_threads, &w);
  if (rc) return rc;
  *out = new misob200_workload{w};
  return 0;
}

/* isoform count of every gene of a workload (cheap: no reads are generated when the
   workload was created with reads_per_gene = 0) */
int misob200_workload_n_iso(const misob200_workload_t *wl, int32_t *n_iso) {
  if (!wl || !n_iso) return MISOB200_EINVAL;
  const Workload &w = *wl->w;
  for (size_t g = 0; g + 1 < w.iso_off.size(); g++) n_iso[g] = w.iso_off[g + 1] - w.iso_off[g];
  return 0;
}

int misob200_workload_view(const misob200_workload_t *wl, misob200_reads_t *v) {
  if (!wl || !v) return MISOB200_EINVAL;
  const Workload &w = *wl->w;
  std::memset(v, 0, sizeof(*v));
  v->n_genes = (int32_t) w.gene_id.size();
  v->iso_off = w.iso_off.data(); v->exon_off = w.exon_off.data();
  v->exon_start = w.ex_start.data(); v->exon_end = w.ex_end.data();
  v->read_off = w.read_off.data(); v->position = w.position.data();
  v->cigar_off = w.cigar_off.data(); v->cigar = w.cigar.data();
  v->hyper = nullptr; v->gene_id = w.gene_id.data();
  v->read_len = w.read_len; v->overhang = 1; v->paired = w.paired;
  v->frag_mean = w.frag_mean; v->frag_var = w.frag_var; v->num_devs = w.num_devs;
  return 0;
}

int misob200_workload_truth(const misob200_workload_t *wl, int32_t gene, double *psi) {
  if (!wl || !psi) return MISOB200_EINVAL;
  const Workload &w = *wl->w;
  if (gene < 0 || (size_t) gene >= w.gene_id.size()) return MISOB200_EINVAL;
  for (int64_t i = w.psi_off[gene]; i < w.psi_off[gene + 1]; i++) psi[i - w.psi_off[gene]] = w.psi[i];
  return 0;
}

int misob200_workload_destroy(misob200_workload_t *wl) {
  if (wl) { delete wl->w; delete wl; }
  return 0;
}

}  // extern "C"
