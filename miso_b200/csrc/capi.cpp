// miso_b200/csrc/capi.cpp -- extern "C" entry points declared in include/miso_b200.h.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "plan.hpp"

namespace misob200 {
const char *last_error();
void plan_layout(Plan &plan, const misob200_params_t &p, long long *n_samples, long long *n_loglik,
                 long long (*range)[4] = nullptr);
int device_init(int device);
int upload(Plan &plan, const misob200_params_t &p);
int run_resident(Plan &plan, double *kernel_ms, int *launches, double *h_samples = nullptr, double *h_loglik = nullptr,
                 int32_t *h_assignment = nullptr);
int download(Plan &plan, double *samples, double *loglik, int32_t *assignment, int32_t *rundata,
             int32_t *status);
int release_device(Plan &plan);
int release_pool();
int stream_version(int set);
int summarize(Plan &plan, double *summary);
int run_timing(Plan &plan, double *timing_ms);
int device_count(int *n);
int bucket_timing(Plan &plan, double *ms);
long long input_bytes(const Plan &plan);
long long output_bytes(Plan &plan);
int compare(Plan &pa, Plan &pb, double *out);
}  // namespace misob200

using namespace misob200;


extern "C" {

int misob200_version(void) { return 100; }
const char *misob200_last_error(void) { return last_error(); }

int misob200_init(int device) { return device_init(device); }
int misob200_shutdown(void) { stage_pool_release(); return release_pool(); }
int misob200_device_count(int *count) { return device_count(count); }

int misob200_plan_create(misob200_plan_t **plan) {
  if (!plan) return MISOB200_EINVAL;
  *plan = new misob200_plan();
  return 0;
}
int misob200_plan_destroy(misob200_plan_t *plan) {
  if (!plan) return 0;
  release_device(plan->p);
  delete plan;
  return 0;
}
int misob200_plan_keep_match(misob200_plan_t *plan, int on) {
  if (!plan) return MISOB200_EINVAL;
  plan->p.keep_match = on != 0;
  return 0;
}
int misob200_plan_tile_format(misob200_plan_t *plan, int format) {
  if (!plan || format < -1 || format > 1) return MISOB200_EINVAL;
  plan->p.force_format = format;
  return 0;
}
int misob200_plan_gene_tile(const misob200_plan_t *plan, int32_t gene, int32_t *format, int32_t *n_weight_classes,
                            int32_t *tile_bytes) {
  if (!plan || gene < 0 || (size_t) gene >= plan->p.desc.size()) return MISOB200_EINVAL;
  const GeneDesc &d = plan->p.desc[gene];
  if (format) *format = d.format;
  if (n_weight_classes) *n_weight_classes = d.ncls;
  if (tile_bytes) *tile_bytes = d.tile_bytes;
  return 0;
}
int misob200_plan_append(misob200_plan_t *plan, const misob200_reads_t *reads, int n_threads) {
  if (!plan || !reads) { set_error("plan_append: null argument"); return MISOB200_EINVAL; }
  if (plan->p.dev) release_device(plan->p);
  return plan_append(plan->p, *reads, n_threads);
}
int misob200_plan_append_device(misob200_plan_t *plan, const misob200_reads_t *reads, int n_threads, int device) {
  if (!plan || !reads) { set_error("plan_append_device: null argument"); return MISOB200_EINVAL; }
  if (device < 0) { set_error("plan_append_device: device ordinal out of range"); return MISOB200_EINVAL; }
  if (plan->p.dev) release_device(plan->p);
  return plan_append(plan->p, *reads, n_threads, device);
}
int misob200_plan_append_device_begin(misob200_plan_t *plan, const misob200_reads_t *reads, int device, void **pending) {
  if (!plan || !reads || !pending) { set_error("plan_append_device_begin: null argument"); return MISOB200_EINVAL; }
  if (device < 0) { set_error("plan_append_device_begin: device ordinal out of range"); return MISOB200_EINVAL; }
  if (plan->p.dev) release_device(plan->p);
  PendingAppend *p = nullptr;
  const int rc = plan_append_device_begin(plan->p, *reads, device, &p);
  *pending = p;
  return rc;
}
int misob200_plan_append_device_finish(misob200_plan_t *plan, void *pending, int n_threads) {
  if (!plan || !pending) { set_error("plan_append_device_finish: null argument"); return MISOB200_EINVAL; }
  return plan_append_device_finish(plan->p, static_cast<PendingAppend *>(pending), n_threads);
}
int misob200_last_match_stats(double *kernel_ms, double *h2d_ms, double *d2h_ms, int64_t *bytes_in, int64_t *bytes_out) {
  long long bi = 0, bo = 0;
  last_match_stats(kernel_ms, h2d_ms, d2h_ms, &bi, &bo);
  if (bytes_in) *bytes_in = bi;
  if (bytes_out) *bytes_out = bo;
  return 0;
}
int misob200_plan_size(const misob200_plan_t *plan, int32_t *n_genes, int64_t *n_reads, int64_t *tile_bytes) {
  if (!plan) return MISOB200_EINVAL;
  if (n_genes) *n_genes = (int32_t) plan->p.desc.size();
  if (n_reads) *n_reads = plan->p.n_reads;
  if (tile_bytes) *tile_bytes = (int64_t) plan->p.tiles.size();
  return 0;
}
static int check_gene(const misob200_plan_t *plan, int32_t gene) {
  if (!plan || gene < 0 || (size_t) gene >= plan->p.host.size()) {
    set_error("gene index out of range");
    return MISOB200_EINVAL;
  }
  return 0;
}
int misob200_plan_gene_info(const misob200_plan_t *plan, int32_t gene, int32_t *n_iso, int32_t *n_reads,
                            int32_t *n_drawn, int32_t *n_classes, int32_t *status) {
  if (int rc = check_gene(plan, gene)) return rc;
  const GeneHost &h = plan->p.host[gene];
  if (n_iso) *n_iso = h.K;
  if (n_reads) *n_reads = h.R;
  if (n_drawn) *n_drawn = h.R2;
  if (n_classes) *n_classes = h.ncls;
  if (status) *status = h.status;
  return 0;
}
/* all genes at once: info5 = n_genes x {K, R, R2, classes, status} */
int misob200_plan_info_all(const misob200_plan_t *plan, int32_t *info5) {
  if (!plan || !info5) return MISOB200_EINVAL;
  const auto &host = plan->p.host;
  for (size_t g = 0; g < host.size(); g++) {
    int32_t *o = info5 + 5 * g;
    o[0] = host[g].K; o[1] = host[g].R; o[2] = host[g].R2; o[3] = host[g].ncls; o[4] = host[g].status;
  }
  return 0;
}
int misob200_plan_gene_classes(const misob200_plan_t *plan, int32_t gene, double *class_templates,
                               double *class_counts) {
  if (int rc = check_gene(plan, gene)) return rc;
  const GeneHost &h = plan->p.host[gene];
  if (class_templates && !h.class_templates.empty())
    std::memcpy(class_templates, h.class_templates.data(), h.class_templates.size() * sizeof(double));
  if (class_counts && !h.class_counts.empty())
    std::memcpy(class_counts, h.class_counts.data(), h.class_counts.size() * sizeof(double));
  return 0;
}
int misob200_plan_gene_match(const misob200_plan_t *plan, int32_t gene, int32_t *codes, int32_t *order) {
  if (int rc = check_gene(plan, gene)) return rc;
  const GeneHost &h = plan->p.host[gene];
  if (h.status == 0 && h.R > 0 && h.codes.empty()) {
    set_error("plan was built without keep_match");
    return MISOB200_EINVAL;
  }
  if (codes && !h.codes.empty()) std::memcpy(codes, h.codes.data(), (size_t) h.K * h.R * sizeof(int32_t));
  if (order && !h.order.empty()) std::memcpy(order, h.order.data(), (size_t) h.R * sizeof(int32_t));
  return 0;
}
int misob200_plan_fragment_table(const misob200_plan_t *plan, int32_t cap, double *prob, int32_t *frag_start,
                                 int32_t *n_len) {
  if (!plan) return MISOB200_EINVAL;
  const Plan &p = plan->p;
  if (n_len) *n_len = p.frag_len_n;
  if (frag_start) *frag_start = p.frag_start;
  if (prob) {
    if (cap < p.frag_len_n) { set_error("fragment table: buffer too small"); return MISOB200_EINVAL; }
    for (int j = 0; j < p.frag_len_n; j++) prob[j] = p.ptab[j + 1];
  }
  return 0;
}
int misob200_plan_offsets(const misob200_plan_t *plan, const misob200_params_t *params, int32_t gene,
                          int64_t *sample_off, int64_t *loglik_off, int64_t *assign_off) {
  if (int rc = check_gene(plan, gene)) return rc;
  if (!params || params->lag < 1 || params->n_iters < 0 || params->burn_in < 0 || params->n_chains < 1 ||
      params->burn_in > params->n_iters) {
    set_error("invalid sampler parameters (iterations/burn-in/lag/chains)");
    return MISOB200_EINVAL;
  }
  // the layout follows the run order of the buckets (run.cu plan_layout); offsets are derived
  // fields of the descriptors, filled on demand
  Plan &p = const_cast<Plan &>(plan->p);
  long long ns = 0, nl = 0;
  plan_layout(p, *params, &ns, &nl);
  const long long so = p.desc[gene].sample_off, lo = p.desc[gene].loglik_off;
  if (sample_off) *sample_off = so;
  if (loglik_off) *loglik_off = lo;
  if (assign_off) *assign_off = p.host[gene].read_base;
  return 0;
}
int misob200_plan_offsets_all(const misob200_plan_t *plan, const misob200_params_t *params, int64_t *sample_off,
                              int64_t *loglik_off, int64_t *assign_off) {
  if (!plan) return MISOB200_EINVAL;
  if (plan->p.host.empty()) return 0;
  if (int rc = misob200_plan_offsets(plan, params, 0, nullptr, nullptr, nullptr)) return rc;     // validates, fills the layout
  const Plan &p = plan->p;
  for (size_t g = 0; g < p.desc.size(); g++) {
    if (sample_off) sample_off[g] = p.desc[g].sample_off;
    if (loglik_off) loglik_off[g] = p.desc[g].loglik_off;
    if (assign_off) assign_off[g] = p.host[g].read_base;
  }
  return 0;
}
int misob200_plan_output_sizes(const misob200_plan_t *plan, const misob200_params_t *params,
                               int64_t *n_samples_f64, int64_t *n_loglik_f64, int64_t *n_assign_i32) {
  if (!plan || !params || params->lag < 1) return MISOB200_EINVAL;
  const Plan &p = plan->p;
  if (params->n_iters < 0 || params->burn_in < 0 || params->n_chains < 1 || params->burn_in > params->n_iters) {
    set_error("invalid sampler parameters (iterations/burn-in/lag/chains)");
    return MISOB200_EINVAL;
  }
  const long long cols = (long long) params->n_chains * ((params->n_iters - params->burn_in) / params->lag);
  long long so = 0, lo = 0;
  for (size_t g = 0; g < p.desc.size(); g++) {
    so += (long long) p.desc[g].K * cols;
    lo += cols;
  }
  if (n_samples_f64) *n_samples_f64 = so;
  if (n_loglik_f64) *n_loglik_f64 = lo;
  if (n_assign_i32) *n_assign_i32 = p.n_reads;
  return 0;
}

int misob200_upload(misob200_plan_t *plan, const misob200_params_t *params) {
  if (!plan || !params) { set_error("upload: null argument"); return MISOB200_EINVAL; }
  return upload(plan->p, *params);
}
int misob200_run_resident(misob200_plan_t *plan, double *kernel_ms, int32_t *launches) {
  if (!plan) return MISOB200_EINVAL;
  return run_resident(plan->p, kernel_ms, launches);
}
int misob200_download(misob200_plan_t *plan, double *samples, double *loglik, int32_t *assignment,
                      int32_t *rundata, int32_t *status) {
  if (!plan) return MISOB200_EINVAL;
  return download(plan->p, samples, loglik, assignment, rundata, status);
}
int misob200_release_device(misob200_plan_t *plan) {
  if (!plan) return MISOB200_EINVAL;
  return release_device(plan->p);
}
int misob200_run(misob200_plan_t *plan, const misob200_params_t *params, double *samples, double *loglik,
                 int32_t *assignment, int32_t *rundata, int32_t *status, double *timing_ms,
                 int32_t *launches) {
  if (!plan || !params) { set_error("run: null argument"); return MISOB200_EINVAL; }
  const bool dbg = std::getenv("MISOB200_RUN_DEBUG") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto t0 = now();
  if (int rc = upload(plan->p, *params)) return rc;
  auto t1 = now();
  // pinned output buffers are written by the kernels themselves; download() then only fetches
  // the assignments and counters
  if (int rc = run_resident(plan->p, nullptr, launches, samples, loglik, assignment)) return rc;
  auto t2 = now();
  if (int rc = download(plan->p, nullptr, nullptr, assignment, rundata, status)) return rc;
  run_timing(plan->p, timing_ms);
  if (dbg) {
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[run] %zu genes: upload %.1f ms, run_resident %.1f ms (kernels %.1f), download + epilogue %.1f ms\n",
            plan->p.desc.size(), ms(t0, t1), ms(t1, t2), timing_ms ? timing_ms[1] : 0.0, ms(t2, now()));
  }
  return 0;
}
int misob200_bucket_timing(misob200_plan_t *plan, double *ms9) {
  if (!plan || !ms9) return MISOB200_EINVAL;
  return bucket_timing(plan->p, ms9);
}
int misob200_transfer_bytes(misob200_plan_t *plan, int64_t *h2d, int64_t *d2h) {
  if (!plan) return MISOB200_EINVAL;
  if (h2d) *h2d = input_bytes(plan->p);
  if (d2h) *d2h = output_bytes(plan->p);
  return 0;
}
int misob200_compare(misob200_plan_t *plan_a, misob200_plan_t *plan_b, double *out) {
  if (!plan_a || !plan_b || !out) return MISOB200_EINVAL;
  return compare(plan_a->p, plan_b->p, out);
}
int misob200_host_threads(void) { return host_threads(); }
int misob200_stream_version(int set) { return stream_version(set); }
int misob200_summarize(misob200_plan_t *plan, double *summary) {
  if (!plan || !summary) return MISOB200_EINVAL;
  return summarize(plan->p, summary);
}

}  // extern "C"
