// miso_b200/csrc/run.cu -- device side of the C ABI: upload, chain kernels, download, summaries.
//
// One launch per isoform count K (the chain kernel is specialised on K so psi
// and the cumulative sums stay in registers); genes of a bucket are dealt to
// persistent warps through an atomic counter, longest first.  Buckets run on
// separate streams so the tail of one overlaps the head of the next.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>

#include <cuda_runtime.h>

#include "chain_kernel.cuh"
#include "quad_kernel.cuh"
#include "plan.hpp"

namespace misob200 {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *last_error() { return g_error.c_str(); }

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
      return MISOB200_ECUDA;                                                              \
    }                                                                                     \
  } while (0)

constexpr int kBuckets = 2 * (kMaxIso + 1);

// Stream version: 2 (Philox4x32-7, default) or 1 (Philox4x32-10, round 1's stream).
static int g_stream_version = 0;
int stream_version(int set) {
  if (set == 1 || set == 2) g_stream_version = set;
  if (!g_stream_version) {
    const char *e = std::getenv("MISOB200_STREAM");
    g_stream_version = (e && std::atoi(e) == 1) ? 1 : 2;
  }
  return g_stream_version;
}
static int stream_rounds() { return stream_version(0) == 1 ? 10 : 7; }

int host_threads() {
  static int cached = 0;
  if (cached) return cached;
  int n = 0;
  if (const char *e = std::getenv("MISOB200_HOST_THREADS")) n = std::atoi(e);
  if (n < 1) {
    n = (int) std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n > 0 ? n : 1 << 20, CPU_COUNT(&set));
    if (FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {      // "<quota> <period>" or "max <period>"
      char q[64];
      long long period = 0;
      if (std::fscanf(f, "%63s %lld", q, &period) == 2 && std::strcmp(q, "max") != 0 && period > 0)
        n = std::min<long long>(n, std::max<long long>(1, (std::atoll(q) + period / 2) / period));
      std::fclose(f);
    }
    int local_world = 1;
    if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) local_world = std::max(1, std::atoi(e));
    n = std::max(1, n / local_world);
  }
  cached = std::max(1, std::min(n, 32));
  return cached;
}

struct DevState {
  int device = 0;
  misob200_params_t params{};
  cudaStream_t stream = nullptr;            // copies + summary
  // buckets: b = fmt * (kMaxIso + 1) + K, fmt 0 dense tiles, 1 class tiles
  cudaStream_t kstream[kBuckets] = {};      // one per bucket
  cudaEvent_t ev[6] = {};
  cudaStream_t hstream[kBuckets] = {};      // low priority: a bucket's helper grid (run_resident, "balanced")
  cudaEvent_t kdone[kBuckets] = {};
  cudaEvent_t kmain[kBuckets] = {};
  cudaEvent_t kbeg[kBuckets] = {}, kend[kBuckets] = {};
  uint8_t *d_tiles = nullptr;
  GeneDesc *d_desc = nullptr;
  double *d_ptab = nullptr, *d_neglog = nullptr;
  int n_neglog = 0;
  double *d_samples = nullptr, *d_loglik = nullptr, *d_summary = nullptr;
  double *zc_samples = nullptr, *zc_loglik = nullptr;   // device views of the caller's pinned output buffers (this run)
  double *d_compare = nullptr;      // two-sample records (compare_device), grow-only
  size_t compare_cap = 0;           // ... in events
  uint8_t *d_drawn = nullptr;
  int *d_accrej = nullptr;
  unsigned *d_queue = nullptr;
  ChainState *d_state = nullptr;    // chain segments: hand-over records, progress counters, ready queues
  int *d_progress = nullptr;
  unsigned *d_ring = nullptr, *d_ring_tail = nullptr;
  size_t ring_cap = 0;              // entries
  int *d_items = nullptr;
  std::vector<int> items[kBuckets];
  int item_off[kBuckets + 1] = {};
  long long n_samples = 0, n_loglik = 0;
  long long range[kBuckets + 1][4] = {};   // output ranges per bucket (plan_layout)
  cudaStream_t cstream = nullptr;          // device->host copies of finished buckets
  cudaEvent_t cdone = nullptr;
  bool uploaded = false, have_run = false;
  bool summary_valid = false;        // d_summary holds the records of the last run
  const int32_t *prefilled = nullptr;      // assignment buffer run_resident pre-filled during the last run
  // pinned staging of the plan's tile arena and descriptors (grow-only, pooled with the state): the
  // plan's own arrays are pageable std::vectors, and page-locking them per plan (cudaHostRegister)
  // cost 0.4 s per fresh plan while other threads of the process were allocating
  uint8_t *h_stage = nullptr;
  size_t cap_h_stage = 0;
  bool tiles_staged = false;
  long long staged_layout = -1;      // Plan.lay_generation of the descriptors in h_stage
  std::vector<double> h_ptab;       // plan.ptab + the 1.0 of the uniform-code classes
  std::vector<double> h_neglog;     // -log(n), read scores of the class format
  uint8_t *h_drawn = nullptr;       // pinned staging of the drawn assignments / accept counters (grow-only)
  int *h_accrej = nullptr;
  size_t cap_h_drawn = 0, cap_h_accrej = 0;
  // capacities (bytes) of the grow-only device buffers: a released state goes back to a small pool
  // and serves the next plan without a cudaMalloc (drop-in calls run one gene per plan)
  size_t cap_tiles = 0, cap_desc = 0, cap_ptab = 0, cap_neglog = 0, cap_samples = 0, cap_loglik = 0, cap_summary = 0,
         cap_drawn = 0, cap_accrej = 0, cap_state = 0, cap_progress = 0, cap_items = 0;
  int sm_count = 0;
  size_t device_bytes() const {
    return cap_tiles + cap_desc + cap_ptab + cap_neglog + cap_samples + cap_loglik + cap_summary + cap_drawn +
           cap_accrej + cap_state + cap_progress + cap_items + ring_cap * sizeof(unsigned) + compare_cap * 256;
  }
};

// Released device states, kept for the next plan on the same GPU: streams, events, device buffers
// and pinned staging survive, so back-to-back pysplicing.MISO calls (one gene per plan, the way
// misopy/run_miso.py drives the sampler) and the batches of a pipeline (miso_b200/pipeline.py)
// pay no allocation after the first.
static std::mutex g_pool_mu;
static std::vector<DevState *> g_pool;
constexpr size_t kPoolMaxStates = 4, kPoolMaxBytes = 4ull << 30;      // (device bytes held by the pool, all states together)

template <class T>
static int ensure_dev(T *&p, size_t &cap, size_t need_bytes) {
  need_bytes = std::max<size_t>(need_bytes, 16);
  if (need_bytes <= cap) return 0;
  cudaFree(p); p = nullptr; cap = 0;
  CK(cudaMalloc(&p, need_bytes));
  cap = need_bytes;
  return 0;
}
template <class T>
static int ensure_pinned(T *&p, size_t &cap, size_t need_bytes) {
  need_bytes = std::max<size_t>(need_bytes, 16);
  if (need_bytes <= cap) return 0;
  if (p) cudaFreeHost(p);
  p = nullptr; cap = 0;
  CK(cudaHostAlloc(&p, need_bytes, cudaHostAllocDefault));
  cap = need_bytes;
  return 0;
}

static int S_of(const misob200_params_t &p) { return p.lag > 0 ? (p.n_iters - p.burn_in) / p.lag : 0; }
// Columns of a gene's sample block: n_chains * S.  (The reference sizes its matrix with noSamples =
// noChains * (noIterations - noBurnIn) / noLag (miso.c:661), which exceeds n_chains * S when the lag does
// not divide the sampling span and leaves that many all-zero columns behind the recorded ones
// (miso.c:822); here every column is a recorded sample and rundata.noSamples says n_chains * S --
// documented deviation, include/miso_b200.h.)
static long long cols_of(const misob200_params_t &p) { return (long long) p.n_chains * S_of(p); }

static int check_params(const misob200_params_t &p) {
  if (p.n_iters < 0 || p.burn_in < 0 || p.lag < 1 || p.n_chains < 1 || p.burn_in > p.n_iters) {
    set_error("invalid sampler parameters (iterations/burn-in/lag/chains)");
    return MISOB200_EINVAL;
  }
  if (p.start != MISOB200_START_AUTO && p.start != MISOB200_START_UNIFORM && p.start != MISOB200_START_RANDOM) {
    set_error("only MISO_START_AUTO, MISO_START_UNIFORM and MISO_START_RANDOM are implemented "
              "(GIVEN has no start_psi at the pysplicing boundary, LINEAR needs the NNLS solver)");
    return MISOB200_UNIMPLEMENTED;
  }
  if (p.stop != MISOB200_STOP_FIXEDNO) {
    set_error("only MISO_STOP_FIXEDNO is implemented");
    return MISOB200_UNIMPLEMENTED;
  }
  if (p.algo != MISOB200_ALGO_REASSIGN) {
    set_error("only MISO_ALGO_REASSIGN is implemented (the one misopy uses, miso_sampler.py:322)");
    return MISOB200_UNIMPLEMENTED;
  }
  return 0;
}

// Work lists: per (tile format, K) bucket, genes by falling number of drawing reads.
static int bucket_of(const GeneDesc &d) {
  return (d.status == 0 && d.K >= 2 && d.K <= kMaxIso) ? (d.format ? 1 : 0) * (kMaxIso + 1) + d.K : -1;
}
void plan_buckets(const Plan &plan, std::vector<int> (&items)[2 * (kMaxIso + 1)]) {
  for (auto &v : items) v.clear();
  for (size_t g = 0; g < plan.desc.size(); g++) {
    const int b = bucket_of(plan.desc[g]);
    if (b >= 0) items[b].push_back((int) g);
  }
  for (auto &v : items)
    std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return plan.desc[a].R2 > plan.desc[b].R2; });
}

// Output layout.  The posterior samples of a gene are one block [s*C+c][K] (the reference's
// column-major K x (C*S), miso.c:884-888); the blocks are laid out in the order the buckets
// run -- dense before class tiles, K = 8 down to 2, work-list order inside a bucket, genes
// that do not run last -- so that a finished bucket's outputs are one contiguous range and its
// device->host copy overlaps the buckets still running (run_resident).  Callers locate a gene
// through misob200_plan_offsets.  `range` (optional): per bucket [begin, end) in f64 elements
// of samples, then of loglik; entry 2*(kMaxIso+1) is the tail of genes that do not run.
void plan_layout(Plan &plan, const misob200_params_t &p, long long *n_samples, long long *n_loglik,
                 long long (*range_out)[4]) {
  const long long S = cols_of(p);      // (cache key: columns per gene block)
  long long (*range)[4] = range_out;
  if (plan.lay_chains == p.n_chains && plan.lay_S == S && plan.lay_genes == plan.desc.size()) {
    if (range) std::memcpy(range, plan.lay_range, sizeof(plan.lay_range));
    *n_samples = plan.lay_n_samples; *n_loglik = plan.lay_n_loglik;
    return;
  }
  range = plan.lay_range;
  std::vector<int> items[2 * (kMaxIso + 1)];
  plan_buckets(plan, items);
  long long so = 0, lo = 0;
  auto place = [&](int g) {
    plan.desc[g].sample_off = so;
    plan.desc[g].loglik_off = lo;
    so += (long long) plan.desc[g].K * S;
    lo += S;
  };
  for (int fmt = 0; fmt < 2; fmt++)
    for (int k = kMaxIso; k >= 0; k--) {
      const int b = fmt * (kMaxIso + 1) + k;
      const long long s0 = so, l0 = lo;
      for (int g : items[b]) place(g);
      range[b][0] = s0; range[b][1] = so; range[b][2] = l0; range[b][3] = lo;
    }
  const long long s0 = so, l0 = lo;
  for (size_t g = 0; g < plan.desc.size(); g++)
    if (bucket_of(plan.desc[g]) < 0) place((int) g);
  range[2 * (kMaxIso + 1)][0] = s0; range[2 * (kMaxIso + 1)][1] = so; range[2 * (kMaxIso + 1)][2] = l0; range[2 * (kMaxIso + 1)][3] = lo;
  plan.lay_chains = p.n_chains; plan.lay_S = S; plan.lay_genes = plan.desc.size();
  plan.lay_generation++;
  plan.lay_n_samples = so; plan.lay_n_loglik = lo;
  if (range_out) std::memcpy(range_out, plan.lay_range, sizeof(plan.lay_range));
  *n_samples = so; *n_loglik = lo;
}

static void free_dev(DevState *st) {
  if (!st) return;
  cudaSetDevice(st->device);
  cudaFree(st->d_tiles); cudaFree(st->d_desc); cudaFree(st->d_ptab); cudaFree(st->d_neglog); cudaFree(st->d_samples);
  cudaFree(st->d_loglik); cudaFree(st->d_summary); cudaFree(st->d_compare); cudaFree(st->d_drawn); cudaFree(st->d_accrej);
  cudaFree(st->d_queue); cudaFree(st->d_items); cudaFree(st->d_state); cudaFree(st->d_progress);
  cudaFree(st->d_ring); cudaFree(st->d_ring_tail);
  for (auto &e : st->ev) if (e) cudaEventDestroy(e);
  for (auto &e : st->kdone) if (e) cudaEventDestroy(e);
  for (auto &e : st->kmain) if (e) cudaEventDestroy(e);
  for (auto &s : st->hstream) if (s) cudaStreamDestroy(s);
  for (auto &e : st->kbeg) if (e) cudaEventDestroy(e);
  for (auto &e : st->kend) if (e) cudaEventDestroy(e);
  if (st->h_drawn) cudaFreeHost(st->h_drawn);
  if (st->h_accrej) cudaFreeHost(st->h_accrej);
  if (st->h_stage) cudaFreeHost(st->h_stage);
  for (auto &s : st->kstream) if (s) cudaStreamDestroy(s);
  if (st->stream) cudaStreamDestroy(st->stream);
  if (st->cstream) cudaStreamDestroy(st->cstream);
  if (st->cdone) cudaEventDestroy(st->cdone);
  delete st;
}

int release_device(Plan &plan) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (st) {
    st->tiles_staged = false;
  st->staged_layout = -1;
    st->staged_layout = -1;
    st->uploaded = st->have_run = false;
    bool pooled = false;
    {
      std::lock_guard<std::mutex> lock(g_pool_mu);
      size_t held = 0;
      for (DevState *q : g_pool) held += q->device_bytes();
      if (g_pool.size() < kPoolMaxStates && held + st->device_bytes() <= kPoolMaxBytes) { g_pool.push_back(st); pooled = true; }
    }
    if (!pooled) free_dev(st);
  }
  plan.dev = nullptr;
  return 0;
}

int release_pool() {
  std::vector<DevState *> all;
  {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    all.swap(g_pool);
  }
  for (DevState *st : all) free_dev(st);
  return 0;
}

int device_count(int *n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
  if (n) *n = c;
  return 0;
}

int device_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) +
              " -- miso_b200 has no CPU path by design");
    return MISOB200_ECUDA;
  }
  if (device < 0 || device >= n) { set_error("device ordinal out of range"); return MISOB200_EINVAL; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("miso_b200 is built for sm_100a only; device " + std::to_string(device) + " is sm_" +
              std::to_string(prop.major) + std::to_string(prop.minor));
    return MISOB200_ECUDA;
  }
  CK(cudaSetDevice(device));
  return 0;
}

static int copy_inputs(Plan &plan, DevState *st) {
  const size_t G = plan.desc.size(), tile_bytes = plan.tiles.size();
  // tiles (once per plan) and descriptors (the output layout may have changed) into pinned staging
  const size_t tile_pad = (tile_bytes + 255) & ~(size_t) 255, desc_bytes = G * sizeof(GeneDesc);
  if (int r = ensure_pinned(st->h_stage, st->cap_h_stage, tile_pad + desc_bytes)) return r;
  if (!st->tiles_staged) {
    const int nt = tile_bytes >= (8u << 20) ? std::max(1, std::min(host_threads(), 8)) : 1;
    if (nt == 1) std::memcpy(st->h_stage, plan.tiles.data(), tile_bytes);
    else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; t++) {
        const size_t a = tile_bytes * t / nt, b = tile_bytes * (t + 1) / nt;
        pool.emplace_back([=, &plan] { std::memcpy(st->h_stage + a, plan.tiles.data() + a, b - a); });
      }
      for (auto &t : pool) t.join();
    }
    st->tiles_staged = true;
  }
  if (G && st->staged_layout != plan.lay_generation) {
    std::memcpy(st->h_stage + tile_pad, plan.desc.data(), desc_bytes);
    st->staged_layout = plan.lay_generation;
  }
  CK(cudaEventRecord(st->ev[0], st->stream));
  if (tile_bytes) CK(cudaMemcpyAsync(st->d_tiles, st->h_stage, tile_bytes, cudaMemcpyHostToDevice, st->stream));
  if (G) CK(cudaMemcpyAsync(st->d_desc, st->h_stage + tile_pad, desc_bytes, cudaMemcpyHostToDevice, st->stream));
  CK(cudaMemcpyAsync(st->d_ptab, st->h_ptab.data(), st->h_ptab.size() * sizeof(double), cudaMemcpyHostToDevice, st->stream));
  CK(cudaMemcpyAsync(st->d_neglog, st->h_neglog.data(), st->h_neglog.size() * sizeof(double), cudaMemcpyHostToDevice, st->stream));
  for (int b = 0; b < kBuckets; b++)
    if (!st->items[b].empty())
      CK(cudaMemcpyAsync(st->d_items + st->item_off[b], st->items[b].data(), st->items[b].size() * sizeof(int),
                         cudaMemcpyHostToDevice, st->stream));
  CK(cudaEventRecord(st->ev[1], st->stream));
  CK(cudaStreamSynchronize(st->stream));
  st->uploaded = true;
  st->have_run = false;
  return 0;
}

long long input_bytes(const Plan &plan) {
  long long b = (long long) plan.tiles.size() + (long long) plan.desc.size() * (sizeof(GeneDesc) + sizeof(int)) +
                (long long) (plan.ptab.size() + 1) * sizeof(double);
  return b;
}

int upload(Plan &plan, const misob200_params_t &p) {
  int rc = check_params(p);
  if (rc) return rc;
  if (plan.dev) {
    // same plan, same parameters: keep the device buffers and the pinned
    // registration, only redo the host->device copies
    DevState *old = static_cast<DevState *>(plan.dev);
    if (std::memcmp(&old->params, &p, sizeof(p)) == 0) {
      CK(cudaSetDevice(old->device));
      plan_layout(plan, p, &old->n_samples, &old->n_loglik, old->range);      // (offsets may have been queried for other parameters)
      return copy_inputs(plan, old);
    }
  }
  rc = device_init(p.device);
  if (rc) return rc;
  release_device(plan);
  DevState *st = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (size_t i = 0; i < g_pool.size(); i++)
      if (g_pool[i]->device == p.device) { st = g_pool[i]; g_pool.erase(g_pool.begin() + i); break; }
  }
  CK(cudaSetDevice(p.device));
  if (!st) {
    st = new DevState();
    st->device = p.device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p.device));
    st->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&st->cstream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&st->cdone, cudaEventDisableTiming));
    for (auto &e : st->ev) CK(cudaEventCreate(&e));
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int b = 0; b < kBuckets; b++) {
      if (b % (kMaxIso + 1) < 2) continue;
      CK(cudaStreamCreateWithPriority(&st->kstream[b], cudaStreamNonBlocking, prio_hi));
      CK(cudaStreamCreateWithPriority(&st->hstream[b], cudaStreamNonBlocking, prio_lo));
      CK(cudaEventCreateWithFlags(&st->kdone[b], cudaEventDisableTiming));
      CK(cudaEventCreate(&st->kmain[b]));      // (timed: the calibration reads kbeg -> kmain)
      CK(cudaEventCreate(&st->kbeg[b]));
      CK(cudaEventCreate(&st->kend[b]));
    }
  }
  plan.dev = st;
  st->params = p;

  plan_layout(plan, p, &st->n_samples, &st->n_loglik, st->range);
  const size_t G = plan.desc.size();

  // work lists: per (tile format, K), genes ordered by decreasing number of drawing reads
  int total = 0;
  plan_buckets(plan, st->items);
  for (int b = 0; b < kBuckets; b++) {
    st->item_off[b] = total;
    total += (int) st->items[b].size();
  }
  st->item_off[kBuckets] = total;

  // probability table + the weight 1.0 of the uniform-code classes (index n_codes)
  st->h_ptab = plan.ptab;
  st->h_ptab.push_back(1.0);
  // -log(lp) for the paired-end read scores, lp = L_k - (code - 1) (miso_paired.c:409-411)
  {
    int max_l = 0;
    for (size_t g = 0; g < G; g++)
      if (plan.desc[g].paired && plan.desc[g].format == 1)
        for (int k = 0; k < plan.desc[g].K; k++) max_l = std::max(max_l, plan.desc[g].L[k]);
    st->n_neglog = std::min(max_l + 1, 1 << 21);
    // two spare entries: the padding reads of a tile (null class, code 0) look up lp = L_k + 1 in the unchecked
    // read-score pass (class_pass.cuh MODE 2) before the result is discarded -- found by compute-sanitizer
    st->h_neglog.assign((size_t) st->n_neglog + 2, 0.0);
    for (int n = 1; n < st->n_neglog; n++) st->h_neglog[n] = -std::log((double) n);
  }

  const size_t tile_bytes = plan.tiles.size();
  const size_t GC = std::max<size_t>(G, 1) * p.n_chains;
  if (int r = ensure_dev(st->d_tiles, st->cap_tiles, tile_bytes)) return r;
  if (int r = ensure_dev(st->d_desc, st->cap_desc, G * sizeof(GeneDesc))) return r;
  if (int r = ensure_dev(st->d_ptab, st->cap_ptab, st->h_ptab.size() * sizeof(double))) return r;
  if (int r = ensure_dev(st->d_neglog, st->cap_neglog, st->h_neglog.size() * sizeof(double))) return r;
  if (int r = ensure_dev(st->d_samples, st->cap_samples, (size_t) st->n_samples * sizeof(double))) return r;
  if (int r = ensure_dev(st->d_loglik, st->cap_loglik, (size_t) st->n_loglik * sizeof(double))) return r;
  if (int r = ensure_dev(st->d_summary, st->cap_summary, std::max<size_t>(G, 1) * MISOB200_SUMMARY_F64 * sizeof(double))) return r;
  if (int r = ensure_dev(st->d_drawn, st->cap_drawn, (size_t) plan.n_drawn)) return r;
  if (int r = ensure_dev(st->d_accrej, st->cap_accrej, GC * 2 * sizeof(int))) return r;
  if (int r = ensure_dev(st->d_state, st->cap_state, GC * sizeof(ChainState))) return r;
  if (int r = ensure_dev(st->d_progress, st->cap_progress, GC * sizeof(int))) return r;
  if (int r = ensure_dev(st->d_items, st->cap_items, (size_t) std::max(total, 1) * sizeof(int))) return r;
  if (!st->d_queue) CK(cudaMalloc(&st->d_queue, kBuckets * sizeof(unsigned)));
  if (!st->d_ring_tail) CK(cudaMalloc(&st->d_ring_tail, kBuckets * sizeof(unsigned)));

  st->tiles_staged = false;
  st->staged_layout = -1;
  if (int r = ensure_pinned(st->h_drawn, st->cap_h_drawn, (size_t) plan.n_drawn)) return r;
  if (int r = ensure_pinned(st->h_accrej, st->cap_h_accrej, GC * 2 * sizeof(int))) return r;
  return copy_inputs(plan, st);
}

// Steps per chain segment (ChainState in chain_kernel.cuh).  256 keeps the per-segment overhead
// (tile reload, re-derived current point, thresholds) near 1 % and a bucket's tail to the time of
// 256 iterations; MISOB200_SEG_ITERS overrides (tests use odd small values).
static int segment_length(const DevState *st) {
  const char *e = std::getenv("MISOB200_SEG_ITERS");
  int len = e ? std::atoi(e) : 256;
  const int steps = st->params.n_iters + 1;
  if (len < 1 || len > steps) len = steps;
  return len;
}
// `n_units` work units on `n_warps` resident warps.  With at most one unit per warp there is
// nothing to balance, every chain is in flight from start to end; chains are only cut when the
// bucket needs more than one wave.  (Measured: in the one-wave regime hand-overs made the quad
// kernel's resumed segments run up to 2x slower, profiles/README.md "segments".)
static void set_segments(ChainParams &P, DevState *st, int b, long long n_units, long long n_warps) {
  const int steps = st->params.n_iters + 1;
  P.seg_len = segment_length(st);
  if (n_units <= n_warps && !std::getenv("MISOB200_SEG_ALWAYS")) P.seg_len = steps;
  P.n_seg = (steps + P.seg_len - 1) / P.seg_len;
  P.state = st->d_state;
  P.progress = st->d_progress;
  // bucket b's ring: room for one push per (gene-chain, segment)
  P.ring = st->d_ring + (size_t) st->item_off[b] * st->params.n_chains * P.n_seg;
  P.ring_tail = st->d_ring_tail + b;
}

// One bucket, ready to launch: kernel, parameters (all but the segment fields), occupancy.
struct Launch {
  const void *kern = nullptr;
  ChainParams P;
  size_t smem = 0;
  int per_sm = 0, warps = 4, b = 0;
  bool quad = false;
  long long n_units = 0;       // work units (gene-chains, or groups of four)
  long long need_blocks = 0;   // CTAs that give every unit a warp
  double work_ms = 0;          // estimated time on the whole machine (bucket_work_ms)
  long long blocks = 0;        // CTAs of the main grid
};

// Estimated whole-machine time of a bucket, ms per 5000 iterations and chain: the measured
// bucket times of the cfg-3 benchmark (profiles/r2_ab6_stream_v2_ct.log, ~7.1k genes per isoform
// count: 26 (four chains per warp) / 64 / 82 / 91 / 108 / 132 / 153 ms) scaled by the drawing reads of each gene
// -- the scalar part of an iteration costs about as much as the counting pass over 1000 reads
// (profiles/r1_v8_single_K5_lines.txt).  Only the RATIOS between buckets matter: they set
// each bucket's share of the SMs; helper grids absorb the error.
// Self-calibration of that table: after every balanced step of some size the main grids' run times
// say which buckets were given too few SMs (they end last); the per-(layout, K) correction factors
// below follow them with a damped update, so the second step of a plan -- and the first step of the
// next plan of the process, e.g. the next batch of a pipeline -- starts from measured ratios.
static std::mutex g_cal_mu;
static double g_cal[2][kMaxIso + 1] = {{1, 1, 1, 1, 1, 1, 1, 1, 1}, {1, 1, 1, 1, 1, 1, 1, 1, 1}};

static double bucket_work_ms(const Plan &plan, const std::vector<int> &v, int K, bool quad, bool dense) {
  static const double ms_per_gene[kMaxIso + 1] = {0, 0, 25.9 / 7012, 63.5 / 7226, 81.8 / 7122, 91.3 / 7083,
                                                  108.4 / 7168, 131.9 / 7166, 152.5 / 7223};
  static const double r2_ref[kMaxIso + 1] = {0, 0, 14, 973, 1415, 1557, 1620, 1672, 1708};
  double c = ms_per_gene[K];
  if (K == 2 && !quad) c *= 2.1;          // one chain per warp at K = 2: 54 vs 26 ms (r2_ab5)
  if (K > 2 && quad) c *= (K == 3 ? 0.75 : K == 4 ? 0.85 : K == 5 ? 0.93 : 1.05);      // r2_ab8_quad_fair.log
  if (dense) c *= 2.0;
  else {
    std::lock_guard<std::mutex> lock(g_cal_mu);
    c *= g_cal[quad ? 1 : 0][K];
  }
  double w = 0;
  for (int g : v) w += c * (1000.0 + plan.desc[g].R2) / (1000.0 + r2_ref[K]);
  return w;
}

template <int K, int FMT>
static int prepare_bucket(Plan &plan, DevState *st, Launch *out) {
  const int b = FMT * (kMaxIso + 1) + K;
  const auto &v = st->items[b];
  if (v.empty()) return 0;
  constexpr int WARPS = FMT == 1 ? kClassWarps : 4;      // class format: one 16-warp CTA per SM (chain_kernel.cuh)
  // per-warp shared memory: the largest tile of the bucket (+ threshold rows, class format)
  int slot = 0, cls = 0, thr = 0;
  for (int g : v) {
    const GeneDesc &d = plan.desc[g];
    slot = std::max(slot, d.tile_bytes);
    if (FMT == 1) {
      cls = std::max(cls, d.core_bytes - d.cls_off);
      thr = std::max(thr, 32 + Thr<K>::bytes(d.ncls));     // L_k, then the two planes of threshold rows
    }
  }
  const int n_ptab = (int) st->h_ptab.size();
  const int ptab_bytes = (n_ptab * 8 + 15) & ~15;
  size_t smem = (size_t) ptab_bytes + WARPS * (16 + (size_t) slot + thr);
  const size_t smem_cap = 227 * 1024;
  bool in_smem = true;
  if (smem > smem_cap) {      // stream the rows through L2
    in_smem = false;
    slot = FMT == 1 ? cls : 0;
    smem = (size_t) ptab_bytes + WARPS * (16 + (size_t) slot + thr);
  }
  // (the two stream versions are separate kernels with identical parameter layout, philox.cuh)
  const bool v1 = stream_rounds() == 10;
  typedef void (*kern_t)(ChainParams);
#define MISOB200_PICK(SM, WD) (v1 ? (kern_t) (void *) chain_kernel<K, WARPS, SM, WD, FMT, 10> : (kern_t) (void *) chain_kernel<K, WARPS, SM, WD, FMT, 7>)
  kern_t kern = plan.wide ? (in_smem ? MISOB200_PICK(true, true) : MISOB200_PICK(false, true))
                          : (in_smem ? MISOB200_PICK(true, false) : MISOB200_PICK(false, false));
#undef MISOB200_PICK
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  {
    // the seven K buckets run as concurrent kernels: give them all the same L1/shared split
    const char *cv = std::getenv("MISOB200_CARVEOUT");
    const int carve = cv ? std::atoi(cv) : -2;
    if (carve >= -1) CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
  }
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
  if (per_sm < 1) { set_error("chain kernel does not fit on an SM"); return MISOB200_ECUDA; }
  const long long n_items = (long long) v.size() * st->params.n_chains;

  ChainParams P;
  P.desc = st->d_desc;
  P.items = st->d_items + st->item_off[b];
  P.n_genes = (int) v.size();
  P.n_chains = st->params.n_chains;
  P.tiles = st->d_tiles;
  P.ptab = st->d_ptab;
  P.n_ptab = n_ptab;
  P.ptab_min = 1.0;
  for (double v : plan.ptab) if (v > 0 && v < P.ptab_min) P.ptab_min = v;
  P.tame_slg = std::getenv("MISOB200_LITERAL_SCORES") ? 1.0 : kTameSlg;      // (switch: tests of the literal route)
  P.samples = st->d_samples;
  P.loglik = st->d_loglik;
  P.samples_host = st->zc_samples;
  P.loglik_host = st->zc_loglik;
  P.drawn = st->d_drawn;
  P.accrej = st->d_accrej;
  P.queue = st->d_queue + b;
  P.n_iters = st->params.n_iters; P.burn_in = st->params.burn_in; P.lag = st->params.lag;
  P.start = st->params.start;
  P.key = philox_expand_key(st->params.seed);
  P.slot_bytes = slot;
  P.neglog = st->d_neglog;
  P.n_neglog = st->n_neglog;
  P.thr_bytes = thr;
  out->kern = (const void *) kern;
  out->P = P;
  out->smem = smem; out->per_sm = per_sm; out->warps = WARPS; out->b = b; out->quad = false;
  out->n_units = n_items;
  out->need_blocks = (n_items + WARPS - 1) / WARPS;
  out->work_ms = bucket_work_ms(plan, v, K, false, FMT == 0) * st->params.n_chains * (st->params.n_iters / 5000.0);
  return 0;
}

// Four gene-chains per warp (quad_kernel.cuh): class-format buckets whose core tiles fit.
// Returns 1 when it launched, 0 when the caller should use chain_kernel instead.
template <int K>
static int prepare_quad(Plan &plan, DevState *st, Launch *out, int *rc) {
  const int b = (kMaxIso + 1) + K;
  const auto &v = st->items[b];
  constexpr int WARPS = kClassWarps;
  int core = 0, thr = 0;
  for (int g : v) {
    const GeneDesc &d = plan.desc[g];
    core = std::max(core, d.core_bytes);
    thr = std::max(thr, 32 + Thr<K>::bytes(d.ncls));
  }
  // The quad layout pays when the per-iteration scalar part dominates, i.e. for genes with few
  // reads that draw; with many reads the counting pass dominates and costs the same either
  // way, while four tiles per warp lower the occupancy (measured, profiles/README.md).
  {
    long long r2 = 0;
    for (int g : v) r2 += plan.desc[g].R2;
    const char *lim = std::getenv("MISOB200_QUAD_MAX_READS");
    const char *cpw = std::getenv("MISOB200_CHAINS_PER_WARP");      // "4": tests force the layout
    // With the machine kept full either way (28 000 events of one isoform count, profiles/r2_ab8_quad_fair.log)
    // four chains per warp win up to K = 5 at the 1000-1600 drawing reads of cfg-3 -- K = 3: 193 vs 257 ms,
    // K = 5: 351 vs 379 ms -- and lose from K = 7 (525 vs 500 ms): the scalar part they save shrinks relative
    // to the counting pass as K and the read count grow.  (Round 1 measured the layouts bucket by bucket,
    // where a bucket of four-chain units is a single wave and pays its full latency; under the balanced
    // policy every bucket has several waves on its own SM share.)
    const long long max_mean = lim ? std::atoll(lim) : (cpw && std::atoi(cpw) == 4) ? (1LL << 40) : (K == 2 ? 2500 : K <= 5 ? 2000 : 850);
    if (!v.empty() && r2 > max_mean * (long long) v.size()) return 0;
  }
  const int slot = ((core + 127) & ~127) + 32;     // 32 mod 128: the four groups' id words fall in different banks
  const size_t smem = (size_t) WARPS * (16 + (size_t) kQuad * (slot + thr));
  if (smem > 227 * 1024) return 0;
  const bool v1 = stream_rounds() == 10;
  typedef void (*kern_t)(ChainParams);
  kern_t kern = plan.wide ? (v1 ? (kern_t) (void *) quad_kernel<K, WARPS, true, 10> : (kern_t) (void *) quad_kernel<K, WARPS, true, 7>)
                          : (v1 ? (kern_t) (void *) quad_kernel<K, WARPS, false, 10> : (kern_t) (void *) quad_kernel<K, WARPS, false, 7>);
  *rc = MISOB200_ECUDA;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) {
    set_error(std::string("quad kernel: ") + cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
  if (per_sm < 1) { cudaGetLastError(); *rc = 0; return 0; }
  const long long n_items = (long long) v.size() * st->params.n_chains;

  ChainParams P;
  P.desc = st->d_desc;
  P.items = st->d_items + st->item_off[b];
  P.n_genes = (int) v.size();
  P.n_chains = st->params.n_chains;
  P.tiles = st->d_tiles;
  P.ptab = st->d_ptab;
  P.n_ptab = (int) st->h_ptab.size();
  P.ptab_min = 1.0;
  for (double v : plan.ptab) if (v > 0 && v < P.ptab_min) P.ptab_min = v;
  P.tame_slg = std::getenv("MISOB200_LITERAL_SCORES") ? 1.0 : kTameSlg;      // (switch: tests of the literal route)
  P.samples = st->d_samples;
  P.loglik = st->d_loglik;
  P.samples_host = st->zc_samples;
  P.loglik_host = st->zc_loglik;
  P.drawn = st->d_drawn;
  P.accrej = st->d_accrej;
  P.queue = st->d_queue + b;
  P.n_iters = st->params.n_iters; P.burn_in = st->params.burn_in; P.lag = st->params.lag;
  P.start = st->params.start;
  P.key = philox_expand_key(st->params.seed);
  P.slot_bytes = slot;
  P.neglog = st->d_neglog;
  P.n_neglog = st->n_neglog;
  P.thr_bytes = thr;
  out->kern = (const void *) kern;
  out->P = P;
  out->smem = smem; out->per_sm = per_sm; out->warps = WARPS; out->b = b; out->quad = true;
  out->n_units = (n_items + kQuad - 1) / kQuad;
  out->need_blocks = (out->n_units + WARPS - 1) / WARPS;
  out->work_ms = bucket_work_ms(plan, v, K, true, false) * st->params.n_chains * (st->params.n_iters / 5000.0);
  *rc = 0;
  return 1;
}

static int prepare_quad_k(Plan &plan, DevState *st, int k, Launch *nl, int *rc) {
  switch (k) {
    case 2: return prepare_quad<2>(plan, st, nl, rc);
    case 3: return prepare_quad<3>(plan, st, nl, rc);
    case 4: return prepare_quad<4>(plan, st, nl, rc);
    case 5: return prepare_quad<5>(plan, st, nl, rc);
    case 6: return prepare_quad<6>(plan, st, nl, rc);
    case 7: return prepare_quad<7>(plan, st, nl, rc);
    case 8: return prepare_quad<8>(plan, st, nl, rc);
  }
  return 0;
}

template <int FMT>
static int prepare_k(Plan &plan, DevState *st, int k, Launch *nl) {
  switch (k) {
    case 2: return prepare_bucket<2, FMT>(plan, st, nl);
    case 3: return prepare_bucket<3, FMT>(plan, st, nl);
    case 4: return prepare_bucket<4, FMT>(plan, st, nl);
    case 5: return prepare_bucket<5, FMT>(plan, st, nl);
    case 6: return prepare_bucket<6, FMT>(plan, st, nl);
    case 7: return prepare_bucket<7, FMT>(plan, st, nl);
    case 8: return prepare_bucket<8, FMT>(plan, st, nl);
  }
  return 0;
}

// cluster > 1: the grid is launched as thread-block clusters of that many CTAs, which the hardware
// places on neighbouring SMs (a pair shares a TPC) -- the kernels do not use the cluster, it only
// keeps the SMs of one bucket together (run_resident, "balanced").
static int fire(Launch &L, long long blocks, cudaStream_t stream, int cluster = 1) {
  void *args[] = {&L.P};
  if (cluster > 1 && blocks % cluster == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) blocks); cfg.blockDim = dim3(L.warps * 32);
    cfg.dynamicSmemBytes = L.smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned) cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, L.kern, args));
    return 0;
  }
  CK(cudaLaunchKernel(L.kern, dim3((unsigned) blocks), dim3(L.warps * 32), args, L.smem, stream));
  return 0;
}

// The part of the output epilogue that does not depend on the run -- every read's assignment
// starts as -1 / its only compatible isoform (miso.c:943-946 for the reads that never draw) -- is
// written while the chain kernels run and the host has nothing else to do; download() then only
// scatters the drawn reads.
static void prefill_assignment(const Plan &plan, int32_t *assignment) {
  const size_t G = plan.host.size();
  auto fill = [&](size_t g0, size_t g1) {
    for (size_t g = g0; g < g1; g++) {
      const GeneHost &h = plan.host[g];
      int32_t *a = assignment + h.read_base;
      if (h.status == 0) for (int r = 0; r < h.R; r++) a[r] = h.fixed_ass[r];
      else for (int r = 0; r < h.R; r++) a[r] = -1;
    }
  };
  unsigned nt = G < 256 ? 1u : (unsigned) host_threads();
  if (nt == 1) { fill(0, G); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; t++) pool.emplace_back(fill, G * t / nt, G * (t + 1) / nt);
  for (auto &t : pool) t.join();
}

int run_resident(Plan &plan, double *kernel_ms, int *launches, double *h_samples, double *h_loglik, int32_t *h_assignment) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st || !st->uploaded) { set_error("run_resident: plan is not on the device (call misob200_upload)"); return MISOB200_EINVAL; }
  CK(cudaSetDevice(st->device));
  int nl = 0;
  CK(cudaMemsetAsync(st->d_queue, 0, kBuckets * sizeof(unsigned), st->stream));
  CK(cudaMemsetAsync(st->d_progress, 0, std::max<size_t>(plan.desc.size(), 1) * st->params.n_chains * sizeof(int), st->stream));
  {
    // ready queues of the chain segments: all slots empty, no pushes yet
    const int seg_len = segment_length(st);
    const size_t n_seg = (size_t) (st->params.n_iters + seg_len) / seg_len;
    const size_t need = std::max<size_t>((size_t) st->item_off[kBuckets] * st->params.n_chains * n_seg, 1);
    if (need > st->ring_cap) {
      CK(cudaStreamSynchronize(st->stream));
      cudaFree(st->d_ring);
      st->d_ring = nullptr;
      CK(cudaMalloc(&st->d_ring, need * sizeof(unsigned)));
      st->ring_cap = need;
    }
    CK(cudaMemsetAsync(st->d_ring, 0xff, need * sizeof(unsigned), st->stream));
    CK(cudaMemsetAsync(st->d_ring_tail, 0, kBuckets * sizeof(unsigned), st->stream));
  }
  // recorded samples that a short chain never writes stay zero, like the
  // reference's zero-initialised sample matrix
  CK(cudaMemsetAsync(st->d_samples, 0, std::max<long long>(st->n_samples, 1) * sizeof(double), st->stream));
  CK(cudaMemsetAsync(st->d_loglik, 0, std::max<long long>(st->n_loglik, 1) * sizeof(double), st->stream));
  CK(cudaEventRecord(st->ev[2], st->stream));
  // Pinned output buffers are written by the kernels themselves (ChainParams.samples_host):
  // nothing is left to copy when the chains end.  Pageable buffers get the device->host copies.
  st->zc_samples = st->zc_loglik = nullptr;
  bool zc = false;
  if (h_samples && h_loglik && !std::getenv("MISOB200_NO_ZEROCOPY")) {
    cudaPointerAttributes pa, pl;
    if (cudaPointerGetAttributes(&pa, h_samples) == cudaSuccess && cudaPointerGetAttributes(&pl, h_loglik) == cudaSuccess &&
        pa.type == cudaMemoryTypeHost && pl.type == cudaMemoryTypeHost && pa.devicePointer && pl.devicePointer) {
      zc = true;
      st->zc_samples = static_cast<double *>(pa.devicePointer);
      st->zc_loglik = static_cast<double *>(pl.devicePointer);
    } else {
      cudaGetLastError();
    }
  }
  int rc = 0;
  // development only (tools/ab_bench.py): time one isoform-count bucket by itself; the
  // other genes' outputs are then left unset
  const char *only = std::getenv("MISOB200_ONLY_K");
  const int only_k = only ? std::atoi(only) : 0;
  const bool serial = std::getenv("MISOB200_SERIAL") != nullptr;                 // development knobs
  const bool concurrent_all = std::getenv("MISOB200_CONCURRENT") != nullptr;
  // chains per warp: four (quad_kernel.cuh) once there are enough gene-chains to fill the
  // machine that way, else one (chain_kernel.cuh: a chain alone on a warp finishes sooner).
  // MISOB200_CHAINS_PER_WARP = 1 | 4 overrides (tests run both).
  bool quad;
  {
    long long n_class_items = 0;
    for (int k = 2; k <= kMaxIso; k++) n_class_items += (long long) st->items[(kMaxIso + 1) + k].size() * st->params.n_chains;
    quad = n_class_items >= 2LL * st->sm_count * 16;
    const char *cpw = std::getenv("MISOB200_CHAINS_PER_WARP");
    if (cpw) quad = std::atoi(cpw) == 4;
  }
  // A four-chain unit takes ~4x as long as one chain, so a bucket of them has a quarter of the
  // waves: four chains per warp only where the bucket still gets >= 2.5 waves on its SM share
  // (waves = step time / unit latency; with fewer, the bucket's own latency sets the step: a rank's
  // 6 250-event shard ran 137 instead of 87 ms with K = 4, 5 four to a warp).  K = 2 units are short.
  bool quad_ok[kMaxIso + 1];
  {
    const bool forced = std::getenv("MISOB200_CHAINS_PER_WARP") || std::getenv("MISOB200_QUAD_MAX_READS");
    const double resident_warps = 16.0 * st->sm_count;
    double total = 0, work[kMaxIso + 1] = {0};
    for (int fmt = 0; fmt < 2; fmt++)
      for (int k = 2; k <= kMaxIso; k++) {
        const auto &v = st->items[fmt * (kMaxIso + 1) + k];
        if (v.empty()) continue;
        const double w = bucket_work_ms(plan, v, k, k == 2 && fmt == 1, fmt == 0);
        total += w;
        if (fmt == 1) work[k] = w;
      }
    for (int k = 2; k <= kMaxIso; k++) {
      const size_t n = st->items[(kMaxIso + 1) + k].size();
      const double unit_ms = n ? 4.0 * work[k] * resident_warps / (double) n : 0.0;      // latency of a four-chain unit
      quad_ok[k] = forced || k == 2 || (unit_ms > 0 && total / unit_ms >= 2.5);
    }
  }
  // dense buckets first (slowest per read), big K before small K (longest chains)
  std::vector<Launch> Ls;
  for (int fmt = 0; fmt < 2 && !rc; fmt++)
    for (int k = kMaxIso; k >= 2 && !rc; k--) {
      const int b = fmt * (kMaxIso + 1) + k;
      if (st->items[b].empty()) continue;
      if (only_k && k != only_k) continue;
      Launch L;
      if (!(fmt && quad && quad_ok[k] && prepare_quad_k(plan, st, k, &L, &rc))) rc = fmt ? prepare_k<1>(plan, st, k, &L) : prepare_k<0>(plan, st, k, &L);
      if (rc) return rc;
      Ls.push_back(L);
    }
  // Launch policy for the class-format buckets.  "balanced" (default): every bucket gets a share of
  // the SMs in proportion to its estimated work and all buckets run side by side from start to
  // end, each on its own SMs (one 16-warp CTA per SM, chain_kernel.cuh) and cut into segments
  // where it has more units than warps -- no bucket waits for another, one tail per step instead
  // of one per bucket, and a small batch (a rank's shard of a workload dealt to several GPUs)
  // still fills the machine.  Behind the main grids every bucket gets a low-priority HELPER grid
  // on the same work queue: the block scheduler places its CTAs on the SMs that fall free when
  // some bucket ends before the others (the estimate is only that), and they exit at once
  // when their bucket has nothing left.
  // "serial" (MISOB200_SCHED=serial, round 1): buckets that fill the machine one after the other.
  // Dense-format buckets (rare fallback, 4-warp CTAs) always run first, one after the other.
  const char *sched = std::getenv("MISOB200_SCHED");
  std::vector<Launch *> cls;
  for (auto &L : Ls) if (L.b > kMaxIso) cls.push_back(&L);
  const bool balanced = !(sched && std::strcmp(sched, "serial") == 0) && !serial && !concurrent_all && cls.size() > 1;
  // main grids go out as clusters of two CTAs = whole TPCs: the two SMs of a TPC share an
  // instruction cache level, and pairs running different buckets cost ~3 % of the step
  // (profiles/r2_ab3_sched_cluster.log: 770 -> 748 ms; clusters of four were slower)
  const char *cl_env = std::getenv("MISOB200_CLUSTER");
  const int cluster = cl_env ? std::max(1, std::atoi(cl_env)) : 2;
  if (balanced) {
    // shares are counted in clusters (TPCs): `slots` of them on the machine
    const long long slots = st->sm_count / cluster;
    auto need = [&](const Launch *L) { return (L->need_blocks + cluster - 1) / cluster; };
    double total = 0;
    for (auto *L : cls) total += L->work_ms;
    // water-filling: a bucket never gets more CTAs than it has units for; what it leaves goes to the rest
    std::vector<char> capped(cls.size(), 0);
    long long free_slots = slots;
    double open = total;
    for (bool again = true; again;) {
      again = false;
      for (size_t i = 0; i < cls.size(); i++) {
        if (capped[i]) continue;
        const double want = open > 0 ? cls[i]->work_ms / open * free_slots : 0;
        if ((double) need(cls[i]) <= want) {
          cls[i]->blocks = need(cls[i]); capped[i] = 1;
          free_slots -= cls[i]->blocks; open -= cls[i]->work_ms; again = true;
        }
      }
    }
    // largest-remainder rounding of the open buckets' shares, at least one cluster each
    long long given = 0;
    std::vector<std::pair<double, size_t>> rem;
    for (size_t i = 0; i < cls.size(); i++)
      if (!capped[i]) {
        const double want = cls[i]->work_ms / open * free_slots;
        cls[i]->blocks = std::max<long long>(1, (long long) want);
        given += cls[i]->blocks;
        rem.push_back({want - (double) (long long) want, i});
      }
    std::sort(rem.begin(), rem.end(), [](const std::pair<double, size_t> &a, const std::pair<double, size_t> &c) { return a.first > c.first; });
    for (size_t r = 0; r < rem.size() && given < free_slots; r++, given++) cls[rem[r].second]->blocks++;
    for (auto *L : cls) L->blocks *= cluster;
  }
  if (std::getenv("MISOB200_SCHED_DEBUG"))
    for (auto &L : Ls)
      fprintf(stderr, "[sched] bucket %2d (K = %d%s): units %lld, work %.2f ms (calibration %.3f), need %lld CTAs, share %lld CTAs of %d warps\n", L.b,
              L.b % (kMaxIso + 1), L.quad ? ", four chains per warp" : "", L.n_units, L.work_ms, g_cal[L.quad ? 1 : 0][L.b % (kMaxIso + 1)],
              L.need_blocks, L.blocks, L.warps);
  int prev = -1;
  for (auto &L : Ls) {
    const int b = L.b;
    const bool bal = balanced && b > kMaxIso;
    CK(cudaStreamWaitEvent(st->kstream[b], st->ev[2], 0));
    if (!bal) {
      L.blocks = std::min<long long>(L.need_blocks, (long long) L.per_sm * st->sm_count);
      // A bucket that fills most of the machine runs alone and the next one starts after it
      // (measured, profiles/r1_ab10/ab11).  Small buckets overlap with each other.
      if (prev >= 0 && !concurrent_all) CK(cudaStreamWaitEvent(st->kstream[b], st->kdone[prev], 0));
    } else if (prev >= 0) {
      CK(cudaStreamWaitEvent(st->kstream[b], st->kdone[prev], 0));      // after the dense buckets
    }
    set_segments(L.P, st, b, L.n_units, L.blocks * L.warps);
    CK(cudaEventRecord(st->kbeg[b], st->kstream[b]));
    if (int r = fire(L, L.blocks, st->kstream[b], bal ? cluster : 1)) return r;
    nl++;
    CK(cudaEventRecord(st->kmain[b], st->kstream[b]));
    if (!bal) {
      CK(cudaEventRecord(st->kend[b], st->kstream[b]));
      CK(cudaEventRecord(st->kdone[b], st->kstream[b]));
      const double fill = (double) (L.blocks * L.warps) / ((double) L.per_sm * L.warps * st->sm_count);
      if (serial || b <= kMaxIso || L.P.n_seg > 1 || fill >= 0.6) prev = b;
    }
  }
  if (balanced) {
    // helper grids, biggest bucket first (largest absolute error of the estimate)
    std::vector<Launch *> order(cls);
    std::stable_sort(order.begin(), order.end(), [](const Launch *a, const Launch *c) { return a->work_ms > c->work_ms; });
    for (Launch *L : order) {
      const int b = L->b;
      CK(cudaStreamWaitEvent(st->hstream[b], st->ev[2], 0));
      if (prev >= 0) CK(cudaStreamWaitEvent(st->hstream[b], st->kdone[prev], 0));
      if (L->n_units > L->blocks * L->warps) {      // (a bucket with a warp per unit has nothing to hand out)
        const long long hb = std::min<long long>(L->need_blocks - L->blocks, (long long) L->per_sm * st->sm_count);
        if (int r = fire(*L, hb, st->hstream[b])) return r;
        nl++;
      }
      CK(cudaStreamWaitEvent(st->hstream[b], st->kmain[b], 0));
      CK(cudaEventRecord(st->kend[b], st->hstream[b]));
      CK(cudaEventRecord(st->kdone[b], st->hstream[b]));
    }
  }
  for (auto &L : Ls) {
    const int b = L.b;
    // a finished bucket's posteriors are one contiguous range (plan_layout): copy them out
    // while the other buckets run
    if ((h_samples || h_loglik) && !zc) {
      CK(cudaStreamWaitEvent(st->cstream, st->kdone[b], 0));
      const long long *r = st->range[b];
      if (h_samples && r[1] > r[0])
        CK(cudaMemcpyAsync(h_samples + r[0], st->d_samples + r[0], (r[1] - r[0]) * sizeof(double), cudaMemcpyDeviceToHost, st->cstream));
      if (h_loglik && r[3] > r[2])
        CK(cudaMemcpyAsync(h_loglik + r[2], st->d_loglik + r[2], (r[3] - r[2]) * sizeof(double), cudaMemcpyDeviceToHost, st->cstream));
    }
    CK(cudaStreamWaitEvent(st->stream, st->kdone[b], 0));
  }
  CK(cudaEventRecord(st->ev[3], st->stream));
  st->prefilled = nullptr;
  if (h_assignment) { prefill_assignment(plan, h_assignment); st->prefilled = h_assignment; }
  if (h_samples || h_loglik) {
    // genes that did not run (status != 0): their zeroed blocks, the tail of the layout; and
    // buckets skipped by a development switch
    for (int b = 0; b <= kBuckets; b++) {
      const bool ran = b < kBuckets && !st->items[b].empty() && !(only_k && b % (kMaxIso + 1) != only_k);
      if (ran) continue;
      const long long *r = st->range[b];
      if (zc) {
        if (r[1] > r[0]) std::memset(h_samples + r[0], 0, (r[1] - r[0]) * sizeof(double));
        if (r[3] > r[2]) std::memset(h_loglik + r[2], 0, (r[3] - r[2]) * sizeof(double));
        continue;
      }
      CK(cudaStreamWaitEvent(st->cstream, st->ev[2], 0));
      if (h_samples && r[1] > r[0])
        CK(cudaMemcpyAsync(h_samples + r[0], st->d_samples + r[0], (r[1] - r[0]) * sizeof(double), cudaMemcpyDeviceToHost, st->cstream));
      if (h_loglik && r[3] > r[2])
        CK(cudaMemcpyAsync(h_loglik + r[2], st->d_loglik + r[2], (r[3] - r[2]) * sizeof(double), cudaMemcpyDeviceToHost, st->cstream));
    }
    CK(cudaStreamSynchronize(st->cstream));
  }
  CK(cudaStreamSynchronize(st->stream));
  CK(cudaGetLastError());
#ifdef MISOB200_SEG_DEBUG
  {
    unsigned long long h[8] = {0}, z[8] = {0};
    cudaMemcpyFromSymbol(h, g_seg_dbg, sizeof(h));
    cudaMemcpyToSymbol(g_seg_dbg, z, sizeof(z));
    fprintf(stderr, "[seg debug] polls %llu wait Mcycles %.1f pops %llu ring pops %llu; Mcycles in segment 0/1/2/3+: %.0f %.0f %.0f %.0f\n",
            h[0], h[1] / 1e6, h[2], h[3], h[4] / 1e6, h[5] / 1e6, h[6] / 1e6, h[7] / 1e6);
  }
#endif
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, st->ev[2], st->ev[3]));
  if (balanced && ms >= 20.0f && !std::getenv("MISOB200_NO_CALIBRATION")) {
    // measured work of a bucket ~ run time of its main grid x its SMs; compare the buckets' shares of
    // that with their shares of the estimate (only buckets with several waves: the others are latency)
    double sum_m = 0, sum_w = 0;
    std::vector<double> m(cls.size(), 0.0);
    for (size_t i = 0; i < cls.size(); i++) {
      float t = 0;
      if (cls[i]->n_units < 2 * cls[i]->blocks * cls[i]->warps) { m[i] = -1; continue; }
      if (cudaEventElapsedTime(&t, st->kbeg[cls[i]->b], st->kmain[cls[i]->b]) != cudaSuccess) { cudaGetLastError(); m[i] = -1; continue; }
      m[i] = (double) t * (double) cls[i]->blocks;
      sum_m += m[i]; sum_w += cls[i]->work_ms;
    }
    if (sum_m > 0 && sum_w > 0) {
      std::lock_guard<std::mutex> lock(g_cal_mu);
      for (size_t i = 0; i < cls.size(); i++) {
        if (m[i] <= 0) continue;
        const double r = (m[i] / sum_m) / (cls[i]->work_ms / sum_w);
        double &c = g_cal[cls[i]->quad ? 1 : 0][cls[i]->b % (kMaxIso + 1)];
        c = std::min(4.0, std::max(0.25, c * std::pow(std::min(2.0, std::max(0.5, r)), 0.7)));
      }
    }
  }
  if (kernel_ms) *kernel_ms = ms;
  if (launches) *launches = nl;
  st->have_run = true;
  st->summary_valid = false;
  return 0;
}

int summarize_device(Plan &plan, const double **d_summary, cudaStream_t *stream);

int download(Plan &plan, double *samples, double *loglik, int32_t *assignment, int32_t *rundata,
             int32_t *status) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st || !st->have_run) { set_error("download: nothing has run"); return MISOB200_EINVAL; }
  CK(cudaSetDevice(st->device));
  const size_t G = plan.desc.size();
  const misob200_params_t &p = st->params;
  CK(cudaEventRecord(st->ev[4], st->stream));
  if (samples && st->n_samples)
    CK(cudaMemcpyAsync(samples, st->d_samples, st->n_samples * sizeof(double), cudaMemcpyDeviceToHost, st->stream));
  if (loglik && st->n_loglik)
    CK(cudaMemcpyAsync(loglik, st->d_loglik, st->n_loglik * sizeof(double), cudaMemcpyDeviceToHost, st->stream));
  if (assignment && plan.n_drawn)
    CK(cudaMemcpyAsync(st->h_drawn, st->d_drawn, plan.n_drawn, cudaMemcpyDeviceToHost, st->stream));
  if (G)
    CK(cudaMemcpyAsync(st->h_accrej, st->d_accrej, G * p.n_chains * 2 * sizeof(int), cudaMemcpyDeviceToHost, st->stream));
  CK(cudaEventRecord(st->ev[5], st->stream));
  CK(cudaStreamSynchronize(st->stream));

  // the summary records (mean, 95% CI, assigned counts) while the host scatters the assignments: the
  // kernel needs only what is already on the device
  if (!std::getenv("MISOB200_NO_EAGER_SUMMARY")) {
    if (summarize_device(plan, nullptr, nullptr) != 0) cudaGetLastError();      // (not fatal here: misob200_summarize reports it)
  }
  // host epilogue: scatter chain-0 assignments back to input read order
  // (miso.c:943-946) and fill rundata (include/splicing.h:143-146)
  const bool prefilled = assignment && st->prefilled == assignment;      // (run_resident wrote the run-independent part)
  st->prefilled = nullptr;
  auto epilogue = [&](size_t g0, size_t g1) {
    for (size_t g = g0; g < g1; g++) {
      const GeneHost &h = plan.host[g];
      const GeneDesc &d = plan.desc[g];
      if (status) status[g] = h.status;
      if (rundata) {
        int acc = 0, rej = 0;
        if (h.status == 0)
          for (int c = 0; c < p.n_chains; c++) {
            acc += st->h_accrej[(g * p.n_chains + c) * 2];
            rej += st->h_accrej[(g * p.n_chains + c) * 2 + 1];
          }
        int32_t *rd = rundata + g * 9;
        rd[0] = h.K; rd[1] = p.n_iters; rd[2] = 0; rd[3] = p.burn_in; rd[4] = p.lag;
        rd[5] = acc; rd[6] = rej; rd[7] = p.n_chains;
        rd[8] = (int) cols_of(p);
      }
      if (assignment) {
        int32_t *a = assignment + h.read_base;
        if (!prefilled) for (int r = 0; r < h.R; r++) a[r] = h.status == 0 ? h.fixed_ass[r] : -1;
        if (h.status == 0) {
          const uint8_t *dr = st->h_drawn + d.drawn_off;
          for (int i = 0; i < h.R2; i++) a[h.rank_read[i]] = dr[i];
        }
      }
    }
  };
  unsigned nt = (unsigned) host_threads();
  if (G < 256) nt = 1;
  if (nt == 1) epilogue(0, G);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++) pool.emplace_back(epilogue, G * t / nt, G * (t + 1) / nt);
    for (auto &t : pool) t.join();
  }
  return 0;
}

// ---- posterior summaries --------------------------------------------------
// One CTA per gene: mean over the C*S recorded samples and the 95% credible
// interval by order statistics -- /root/reference/misopy/credible_intervals.py:31-55:
// alpha = 1 - 0.95, indices int(round((alpha/2) n)) - 1 and
// int(round((1 - alpha/2) n)) - 1 of the sorted samples, where `round` is
// numpy's (the module does `from numpy import *`), i.e. half-to-even on the
// fp64 product; the host computes the two indices with the same expression --
// plus the per-isoform assigned-read counts of chain 0.
__global__ void summary_kernel(const GeneDesc *desc, int n_genes, int n_chains, int n, int lo, int hi, int n_pad,
                               const double *samples, const uint8_t *drawn, const int *accrej,
                               double *summary) {
  extern __shared__ double vals[];
  const int g = blockIdx.x;
  if (g >= n_genes) return;
  const GeneDesc &d = desc[g];
  double *out = summary + (size_t) g * MISOB200_SUMMARY_F64;
  const int K = d.K;
  int *iout = reinterpret_cast<int *>(out + 24);
  if (d.status != 0) {
    for (int i = threadIdx.x; i < MISOB200_SUMMARY_F64; i += blockDim.x) out[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) { iout[8] = K; iout[11] = d.status; }
    return;
  }
  __shared__ int s_cnt[kMaxIso];
  __shared__ double s_red[32];
  if (threadIdx.x < kMaxIso) s_cnt[threadIdx.x] = threadIdx.x < K ? d.n_fixed[threadIdx.x] : 0;
  __syncthreads();
  for (int i = threadIdx.x; i < d.R2; i += blockDim.x) {
    const int a = drawn[d.drawn_off + i];
    if (a < K) atomicAdd(&s_cnt[a], 1);
  }
  for (int k = 0; k < kMaxIso; k++) {
    double mean = 0.0, vlo = 0.0, vhi = 0.0;
    if (k < K && n > 0) {
      double part = 0.0;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = samples[d.sample_off + (long long) i * K + k];
        vals[i] = v;
        part += v;
      }
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) t += s_red[w];
        s_red[0] = t / n;
      }
      __syncthreads();
      mean = s_red[0];
      // order statistics: bitonic sort of the n values (padded with +inf to a power of two)
      // in shared memory, then elements lo and hi (Python indexing: -1 wraps)
      for (int i = n + threadIdx.x; i < n_pad; i += blockDim.x) vals[i] = INFINITY;
      __syncthreads();
      for (int kk = 2; kk <= n_pad; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
            const int p = i ^ j;
            if (p > i) {
              const double a = vals[i], b = vals[p];
              if ((a > b) == ((i & kk) == 0)) { vals[i] = b; vals[p] = a; }
            }
          }
          __syncthreads();
        }
      if (threadIdx.x == 0) { s_red[1] = vals[lo < 0 ? n + lo : lo]; s_red[2] = vals[hi < 0 ? n + hi : hi]; }
      __syncthreads();
      vlo = s_red[1]; vhi = s_red[2];
      __syncthreads();
    }
    if (threadIdx.x == 0) { out[k] = mean; out[8 + k] = vlo; out[16 + k] = vhi; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0, rej = 0;
    for (int c = 0; c < n_chains; c++) {
      acc += accrej[((size_t) g * n_chains + c) * 2];
      rej += accrej[((size_t) g * n_chains + c) * 2 + 1];
    }
    for (int k = 0; k < kMaxIso; k++) iout[k] = s_cnt[k];
    iout[8] = K; iout[9] = acc; iout[10] = rej; iout[11] = 0;
    for (int k = 12; k < 16; k++) iout[k] = 0;
  }
}

// Launches the summary kernel on the plan's stream; the records stay on the device
// (d_summary, n_genes x 32 f64).  comm.cu all-gathers them from there.
int summarize_device(Plan &plan, const double **d_summary, cudaStream_t *stream) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st || !st->have_run) { set_error("summarize: nothing has run"); return MISOB200_EINVAL; }
  CK(cudaSetDevice(st->device));
  if (d_summary) *d_summary = st->d_summary;
  if (stream) *stream = st->stream;
  const int G = (int) plan.desc.size();
  if (G == 0 || st->summary_valid) return 0;      // (download() already launched it behind the chains, see there)
  const int n = (int) cols_of(st->params);      // every column of the block, like the reference's Python (samples_utils.py)
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  const size_t smem = (size_t) n_pad * sizeof(double);
  if (smem > 200 * 1024) { set_error("summarize: too many samples per gene for the on-chip selection"); return MISOB200_UNIMPLEMENTED; }
  CK(cudaFuncSetAttribute(summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  const double alpha = 1 - 0.95;
  const int lo = (int) nearbyint((alpha / 2) * n) - 1, hi = (int) nearbyint((1 - alpha / 2) * n) - 1;
  summary_kernel<<<G, 128, smem, st->stream>>>(st->d_desc, G, st->params.n_chains, n, lo, hi, n_pad, st->d_samples,
                                              st->d_drawn, st->d_accrej, st->d_summary);
  CK(cudaGetLastError());
  st->summary_valid = true;
  return 0;
}

int summarize(Plan &plan, double *summary) {
  const double *d = nullptr;
  cudaStream_t s = nullptr;
  if (int rc = summarize_device(plan, &d, &s)) return rc;
  const size_t G = plan.desc.size();
  if (G == 0) return 0;
  CK(cudaMemcpyAsync(summary, d, G * MISOB200_SUMMARY_F64 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

// ---- two-sample comparison (misopy/hypothesis_test.py:89-179, :348-380) --------
// One CTA per event.  For isoform k: delta_i = psi1[i,k] - psi2[i,k] over the n
// paired posterior samples; if mean|delta| <= 0.009 or all deltas are equal the
// posterior is taken as peaked on the null (BF = 0); otherwise a Gaussian KDE with
// bandwidth factor 0.3 (scipy gaussian_kde subclass: covariance = var(delta, ddof=1)
// * 0.3^2) is evaluated at 0 and BF = 1 / KDE(0), capped at 1e12.
// out record (32 f64): bf[8], mean1-mean2 [8], mean|delta| [8], KDE(0) [8].
__global__ void compare_kernel(const GeneDesc *da, const GeneDesc *db, int n_genes, int n,
                               const double *sa, const double *sb, double *out) {
  const int g = blockIdx.x;
  if (g >= n_genes) return;
  const GeneDesc &a = da[g], &b = db[g];
  double *o = out + (size_t) g * 32;
  __shared__ double red[4][32];
  const int K = a.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  auto block_sum = [&](double v, int slot) -> double {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane == 0) red[slot][warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; w++) t += red[slot][w];
    return t;
  };
  for (int k = 0; k < kMaxIso; k++) {
    double bf = 0.0, dm = 0.0, mabs = 0.0, kde = 0.0;
    if (k < K && K == b.K && a.status == 0 && b.status == 0 && n > 0) {
      double s = 0.0, sab = 0.0, same = 1.0;
      const double d0 = sa[a.sample_off + k] - sb[b.sample_off + k];
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = sa[a.sample_off + (long long) i * K + k] - sb[b.sample_off + (long long) i * K + k];
        s += d; sab += fabs(d);
        if (d - d0 != 0.0) same = 0.0;
      }
      const double mean = block_sum(s, 0) / n;
      mabs = block_sum(sab, 1) / n;
      const bool all_same = block_sum(1.0 - same, 2) == 0.0;
      dm = mean;
      if (mabs <= 0.009 || all_same) {
        bf = 0.0; kde = INFINITY;               // NullPeakedDensity, hypothesis_test.py:15-26
      } else {
        double ss = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const double d = sa[a.sample_off + (long long) i * K + k] - sb[b.sample_off + (long long) i * K + k];
          ss += (d - mean) * (d - mean);
        }
        const double var = block_sum(ss, 3) / (n - 1);
        const double cov = var * 0.3 * 0.3;
        double e = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const double d = sa[a.sample_off + (long long) i * K + k] - sb[b.sample_off + (long long) i * K + k];
          e += exp(-0.5 * d * d / cov);
        }
        kde = block_sum(e, 0) / (n * sqrt(2.0 * 3.14159265358979323846 * cov));
        bf = kde == 0.0 ? 1e12 : 1.0 / kde;
        if (bf > 1e12) bf = 1e12;
      }
    }
    if (threadIdx.x == 0) { o[k] = bf; o[8 + k] = dm; o[16 + k] = mabs; o[24 + k] = kde; }
    __syncthreads();
  }
}

// Launches compare_kernel on sample A's stream; the records stay on the device in a buffer
// owned by A's device state (grow-only, no allocation per call).
int compare_device(Plan &pa, Plan &pb, const double **d_out, cudaStream_t *stream) {
  DevState *a = static_cast<DevState *>(pa.dev), *b = static_cast<DevState *>(pb.dev);
  if (!a || !b || !a->have_run || !b->have_run) { set_error("compare: both plans must have run and stay resident"); return MISOB200_EINVAL; }
  if (a->device != b->device) { set_error("compare: the two samples of an event must live on the same GPU"); return MISOB200_EINVAL; }
  if (pa.desc.size() != pb.desc.size() || a->params.n_chains != b->params.n_chains || cols_of(a->params) != cols_of(b->params)) {
    set_error("compare: plans differ in events or in the number of recorded samples"); return MISOB200_EINVAL;
  }
  for (size_t g = 0; g < pa.desc.size(); g++)
    if (pa.desc[g].K != pb.desc[g].K) {
      set_error("compare: event " + std::to_string(g) + " has a different number of isoforms in the two samples");
      return MISOB200_EINVAL;
    }
  CK(cudaSetDevice(a->device));
  const int G = (int) pa.desc.size();
  if ((size_t) G > a->compare_cap) {
    cudaFree(a->d_compare); a->d_compare = nullptr; a->compare_cap = 0;
    CK(cudaMalloc(&a->d_compare, (size_t) G * MISOB200_COMPARE_F64 * sizeof(double)));
    a->compare_cap = (size_t) G;
  }
  if (d_out) *d_out = a->d_compare;
  if (stream) *stream = a->stream;
  if (G == 0) return 0;
  const int n = (int) cols_of(a->params);
  // sample B's kernels ran on B's streams: order after them
  CK(cudaEventRecord(b->cdone, b->stream));
  CK(cudaStreamWaitEvent(a->stream, b->cdone, 0));
  compare_kernel<<<G, 128, 0, a->stream>>>(a->d_desc, b->d_desc, G, n, a->d_samples, b->d_samples, a->d_compare);
  CK(cudaGetLastError());
  return 0;
}

int compare(Plan &pa, Plan &pb, double *out) {
  const double *d = nullptr;
  cudaStream_t s = nullptr;
  if (int rc = compare_device(pa, pb, &d, &s)) return rc;
  const size_t G = pa.desc.size();
  if (G == 0) return 0;
  CK(cudaMemcpyAsync(out, d, G * MISOB200_COMPARE_F64 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

// per-bucket kernel durations of the last resident run, ms[k] for K = 0..8 (0 when unused)
int bucket_timing(Plan &plan, double *ms) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st || !st->have_run) { set_error("bucket_timing: nothing has run"); return MISOB200_EINVAL; }
  for (int k = 0; k <= kMaxIso; k++) {
    ms[k] = 0.0;
    for (int fmt = 0; fmt < 2 && k >= 2; fmt++) {
      const int b = fmt * (kMaxIso + 1) + k;
      if (st->items[b].empty()) continue;
      float t = 0;
      if (cudaEventElapsedTime(&t, st->kbeg[b], st->kend[b]) == cudaSuccess) ms[k] = std::max(ms[k], (double) t);
      else cudaGetLastError();
    }
  }
  return 0;
}

long long output_bytes(Plan &plan) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st) return 0;
  return (st->n_samples + st->n_loglik) * 8 + plan.n_drawn + (long long) plan.desc.size() * st->params.n_chains * 8;
}

int run_timing(Plan &plan, double *timing_ms) {
  DevState *st = static_cast<DevState *>(plan.dev);
  if (!st || !timing_ms) return 0;
  float h2d = 0, k = 0, d2h = 0, tot = 0;
  cudaEventElapsedTime(&h2d, st->ev[0], st->ev[1]);
  cudaEventElapsedTime(&k, st->ev[2], st->ev[3]);
  cudaEventElapsedTime(&d2h, st->ev[4], st->ev[5]);
  cudaEventElapsedTime(&tot, st->ev[0], st->ev[5]);
  timing_ms[0] = h2d; timing_ms[1] = k; timing_ms[2] = d2h; timing_ms[3] = tot;
  return 0;
}

}  // namespace misob200
