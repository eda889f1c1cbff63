// miso_b200/csrc/writer.cpp -- batched `.miso` writer: the posterior files of a whole plan,
// formatted and written by a pool of host threads straight from the (pinned) output buffers.
//
// The reference writes one file per gene from Python (misopy/miso_sampler.py:456-465):
//     header line, "sampled_psi\tlog_score\n", then per recorded sample
//     "%s\t%.2f\n" % (",".join(["%.4f" % psi for psi in psi_sample]), log_score)
// 50k events x 450 samples x (K + 1) numbers is ~1.3e8 conversions per step, which would take
// longer than the sampling itself.  Here a number is formatted with exact integer arithmetic:
// v * 10^d is split into the rounded product and its exact residual with one FMA, so the
// result is rounded to nearest, ties to even, on the exact binary value -- what "%.4f" of
// CPython / glibc prints -- without a bignum conversion.  tests/test_formats.py compares it
// with Python's "%" on random and on exact-tie values.
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "plan.hpp"

namespace misob200 {

namespace {

// Appends "%.<D>f" of v (D = 2 or 4) to p, returns the new end.  |v| < 2^50 / 10^D.
template <int D>
inline char *put_fixed(char *p, double v) {
  constexpr double scale = D == 2 ? 100.0 : 10000.0;
  if (!(std::fabs(v) < 1e11)) {                  // inf, nan, huge: the C library agrees with Python here
    if (std::isnan(v)) { std::memcpy(p, "nan", 3); return p + 3; }
    if (std::isinf(v)) { if (v < 0) *p++ = '-'; std::memcpy(p, "inf", 3); return p + 3; }
    return p + std::snprintf(p, 64, D == 2 ? "%.2f" : "%.4f", v);
  }
  const bool neg = std::signbit(v);
  const double a = std::fabs(v);
  const double t = a * scale;                    // rounded product
  const double err = std::fma(a, scale, -t);     // exact residual: a * scale == t + err
  double f = std::floor(t);
  const double frac = t - f;                     // exact
  bool up;
  if (frac > 0.5) up = true;
  else if (frac < 0.5) up = false;
  else up = err > 0 || (err == 0 && std::fmod(f, 2.0) != 0.0);     // exact tie: to even
  if (frac == 0.0 && err < 0) up = false;        // just below the integer f: rounds to f either way
  unsigned long long q = (unsigned long long) f + (up ? 1ull : 0ull);
  if (neg) *p++ = '-';                           // ("-0.00" for a negative that rounds to zero, as printf)
  constexpr unsigned long long mod = D == 2 ? 100ull : 10000ull;
  unsigned long long ip = q / mod, fp = q % mod;
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char) ('0' + ip % 10); ip /= 10; } while (ip);
  while (n) *p++ = tmp[--n];
  *p++ = '.';
  for (int i = D - 1; i >= 0; i--) { p[i] = (char) ('0' + fp % 10); fp /= 10; }
  return p + D;
}

bool write_all(int fd, const char *buf, size_t n) {
  while (n) {
    const ssize_t w = ::write(fd, buf, n);
    if (w < 0) { if (errno == EINTR) continue; return false; }
    buf += w; n -= (size_t) w;
  }
  return true;
}

}  // namespace

// exported for the format test
int format_fixed(double v, int decimals, char *out) {
  char *e = decimals == 2 ? put_fixed<2>(out, v) : put_fixed<4>(out, v);
  *e = '\0';
  return (int) (e - out);
}

int write_miso_files(int n_files, const char *const *paths, const char *const *headers, const double *samples,
                     const int64_t *sample_off, const double *loglik, const int64_t *loglik_off,
                     const int32_t *n_iso, int n_rows, int n_threads, int64_t *bytes_written) {
  if (n_files < 0 || n_rows < 0 || (n_files && (!paths || !headers || !samples || !sample_off || !loglik || !loglik_off || !n_iso))) {
    set_error("write_miso_files: null or negative argument");
    return MISOB200_EINVAL;
  }
  int nt = n_threads > 0 ? n_threads : host_threads();
  if (nt > n_files) nt = n_files > 0 ? n_files : 1;
  std::atomic<int> next(0), failed(-1);
  std::atomic<long long> total(0);
  auto work = [&]() {
    std::string buf;
    long long mine = 0;
    for (int i; (i = next.fetch_add(1)) < n_files;) {
      const int K = n_iso[i];
      const size_t hl = std::strlen(headers[i]);
      static const char cols[] = "sampled_psi\tlog_score\n";
      buf.resize(hl + sizeof(cols) + (size_t) n_rows * ((size_t) K * 24 + 32));
      char *p = &buf[0];
      std::memcpy(p, headers[i], hl); p += hl;
      std::memcpy(p, cols, sizeof(cols) - 1); p += sizeof(cols) - 1;
      const double *s = samples + sample_off[i];
      const double *l = loglik + loglik_off[i];
      for (int r = 0; r < n_rows; r++) {
        for (int k = 0; k < K; k++) {
          if (k) *p++ = ',';
          p = put_fixed<4>(p, s[(size_t) r * K + k]);
        }
        *p++ = '\t';
        p = put_fixed<2>(p, l[r]);
        *p++ = '\n';
      }
      const size_t n = (size_t) (p - buf.data());
      const int fd = ::open(paths[i], O_WRONLY | O_CREAT | O_TRUNC, 0644);
      if (fd < 0 || !write_all(fd, buf.data(), n)) {
        int none = -1;
        failed.compare_exchange_strong(none, i);
        if (fd >= 0) ::close(fd);
        continue;
      }
      ::close(fd);
      mine += (long long) n;
    }
    total += mine;
  };
  if (nt <= 1) work();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  if (bytes_written) *bytes_written = total.load();
  if (failed.load() >= 0) {
    set_error(std::string("write_miso_files: cannot write ") + paths[failed.load()] + ": " + std::strerror(errno));
    return MISOB200_FAILURE;
  }
  return 0;
}


// The whole plan: header = prefix[g] + the run-dependent fields of misopy/miso_sampler.py:444-454
// ("iters=..\tburn_in=..\tlag=..\tpercent_accept=%.2f\tproposal_type=drift\tcounts=..\tassigned_counts=..")
// + suffix[g]; prefix ("#isoforms=[..]\texon_lens=..\t") and suffix ("\tchrom=..\n") depend on the
// annotation only and are prepared once by the caller.  Genes with status != 0, without a path, or
// whose reads are all incompatible (miso_sampler.py:352-354) are skipped; *n_written counts the rest.
int plan_write_miso(const Plan &plan, const misob200_params_t &p, const char *const *paths, const char *const *prefix,
                    const char *const *suffix, const double *samples, const double *loglik, const int32_t *assignment,
                    const int32_t *rundata, int n_threads, int64_t *n_written, int64_t *bytes_written) {
  const int G = (int) plan.desc.size();
  if (p.lag < 1 || p.n_chains < 1 || (G && (!paths || !prefix || !suffix || !samples || !loglik || !assignment || !rundata))) {
    set_error("plan_write_miso: null argument or invalid parameters");
    return MISOB200_EINVAL;
  }
  if (G && plan.lay_genes != plan.desc.size()) { set_error("plan_write_miso: the plan has not run"); return MISOB200_EINVAL; }
  const int n_rows = p.n_chains * ((p.n_iters - p.burn_in) / p.lag);
  int nt = n_threads > 0 ? n_threads : host_threads();
  if (nt > G) nt = G > 0 ? G : 1;
  std::atomic<int> next(0), failed(-1);
  std::atomic<long long> total(0), files(0);
  auto work = [&]() {
    std::string buf;
    long long mine = 0, nf = 0;
    char num[64];
    for (int g; (g = next.fetch_add(1)) < G;) {
      const GeneHost &h = plan.host[g];
      const GeneDesc &d = plan.desc[g];
      if (!paths[g] || h.status != 0) continue;
      const int K = h.K;
      const int32_t *a = assignment + h.read_base;
      int amax = -1;
      long long cnt[kMaxIso] = {0};
      for (int r = 0; r < h.R; r++) {
        const int v = a[r];
        if (v > amax) amax = v;
        if (v >= 0 && v < kMaxIso) cnt[v]++;
      }
      if (amax < 0) continue;                      // no read compatible with an isoform: no file
      buf.assign(prefix[g]);
      const int32_t *rd = rundata + (size_t) g * 9;
      const double acc = rd[5], rej = rd[6];
      buf += "iters=" + std::to_string(p.n_iters) + "\tburn_in=" + std::to_string(p.burn_in) + "\tlag=" + std::to_string(p.lag);
      format_fixed(acc / (acc + rej) * 100, 2, num);
      buf += "\tpercent_accept="; buf += num; buf += "\tproposal_type=drift\tcounts=";
      for (int c = 0; c < h.ncls; c++) {
        if (c) buf += ',';
        buf += '(';
        for (int k = 0; k < K; k++) {
          if (k) buf += ',';
          buf += std::to_string((long long) h.class_templates[(size_t) c * K + k]);
        }
        buf += "):" + std::to_string((long long) h.class_counts[c]);
      }
      buf += "\tassigned_counts=";
      for (int k = 0; k <= amax; k++) {
        if (k) buf += ',';
        buf += std::to_string(k) + ":" + std::to_string(cnt[k]);
      }
      buf += suffix[g];
      static const char cols[] = "sampled_psi\tlog_score\n";
      buf += cols;
      const size_t head = buf.size();
      buf.resize(head + (size_t) n_rows * ((size_t) K * 24 + 32));
      char *q = &buf[head];
      const double *s = samples + d.sample_off;
      const double *l = loglik + d.loglik_off;
      for (int r = 0; r < n_rows; r++) {
        for (int k = 0; k < K; k++) {
          if (k) *q++ = ',';
          q = put_fixed<4>(q, s[(size_t) r * K + k]);
        }
        *q++ = '\t';
        q = put_fixed<2>(q, l[r]);
        *q++ = '\n';
      }
      const size_t n = (size_t) (q - buf.data());
      const int fd = ::open(paths[g], O_WRONLY | O_CREAT | O_TRUNC, 0644);
      if (fd < 0 || !write_all(fd, buf.data(), n)) {
        int none = -1;
        failed.compare_exchange_strong(none, g);
        if (fd >= 0) ::close(fd);
        continue;
      }
      ::close(fd);
      mine += (long long) n; nf++;
    }
    total += mine; files += nf;
  };
  if (nt <= 1) work();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto &t : pool) t.join();
  }
  if (bytes_written) *bytes_written = total.load();
  if (n_written) *n_written = files.load();
  if (failed.load() >= 0) {
    set_error(std::string("plan_write_miso: cannot write ") + paths[failed.load()] + ": " + std::strerror(errno));
    return MISOB200_FAILURE;
  }
  return 0;
}

}  // namespace misob200

extern "C" {

int misob200_plan_write_miso(const misob200_plan_t *plan, const misob200_params_t *params, const char *const *paths,
                             const char *const *prefix, const char *const *suffix, const double *samples,
                             const double *loglik, const int32_t *assignment, const int32_t *rundata, int n_threads,
                             int64_t *n_written, int64_t *bytes_written) {
  if (!plan || !params) { misob200::set_error("plan_write_miso: null argument"); return MISOB200_EINVAL; }
  return misob200::plan_write_miso(plan->p, *params, paths, prefix, suffix, samples, loglik, assignment, rundata,
                                   n_threads, n_written, bytes_written);
}

int misob200_write_miso_files(int32_t n_files, const char *const *paths, const char *const *headers,
                              const double *samples, const int64_t *sample_off, const double *loglik,
                              const int64_t *loglik_off, const int32_t *n_iso, int32_t n_rows, int n_threads,
                              int64_t *bytes_written) {
  return misob200::write_miso_files(n_files, paths, headers, samples, sample_off, loglik, loglik_off, n_iso, n_rows,
                                    n_threads, bytes_written);
}

int misob200_format_fixed(double v, int decimals, char *out32) {
  if (!out32 || (decimals != 2 && decimals != 4)) return -1;
  return misob200::format_fixed(v, decimals, out32);
}

}  // extern "C"
