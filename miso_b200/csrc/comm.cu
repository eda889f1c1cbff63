// miso_b200/csrc/comm.cu -- the one collective of the path: an all-gather of the
// per-gene posterior summary records over NCCL (NVLink 5 / NVSwitch).
//
// Genes are independent (the reference exploits this with OS processes,
// /root/reference/misopy/miso.py:163-188), so sampling needs no exchange; only
// the fixed-size summaries are gathered, once, at the end.  NCCL is loaded with
// dlopen so that a single-GPU process never needs it.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "plan.hpp"

namespace misob200 {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclMax = 2 };

struct Nccl {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} N;

ncclComm_t g_comm = nullptr;
cudaStream_t g_stream = nullptr;
int g_ranks = 0;
double *g_scalar = nullptr;
// persistent, grow-only exchange buffers: one padded record block in, n_ranks blocks out,
// and a pinned host mirror of the gathered table (nothing is allocated per step)
double *g_in = nullptr, *g_out = nullptr, *g_host = nullptr;
size_t g_in_cap = 0, g_out_cap = 0, g_host_cap = 0;
cudaEvent_t g_ready = nullptr;

int load_nccl() {
  if (N.h) return 0;
  const char *names[] = {getenv("MISOB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n) continue;
    N.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (N.h) break;
  }
  if (!N.h) { set_error(std::string("cannot load libnccl: ") + dlerror()); return MISOB200_ENCCL; }
#define SYM(f) *(void **) (&N.f) = dlsym(N.h, "nccl" #f); if (!N.f) { set_error("libnccl lacks nccl" #f); return MISOB200_ENCCL; }
  SYM(GetUniqueId) SYM(CommInitRank) SYM(AllGather) SYM(AllReduce) SYM(CommDestroy) SYM(GetErrorString)
#undef SYM
  return 0;
}

#define NK(call)                                                                   \
  do {                                                                             \
    ncclResult_t r_ = (call);                                                      \
    if (r_ != 0) { set_error(std::string(#call) + ": " + N.GetErrorString(r_)); return MISOB200_ENCCL; } \
  } while (0)
#define CK(call)                                                                   \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return MISOB200_ECUDA; } \
  } while (0)

int reserve(size_t n_in, size_t n_out) {
  if (n_in > g_in_cap) {
    cudaFree(g_in); g_in = nullptr; g_in_cap = 0;
    CK(cudaMalloc(&g_in, n_in * sizeof(double)));
    g_in_cap = n_in;
  }
  if (n_out > g_out_cap) {
    cudaFree(g_out); g_out = nullptr; g_out_cap = 0;
    CK(cudaMalloc(&g_out, n_out * sizeof(double)));
    g_out_cap = n_out;
  }
  if (n_out > g_host_cap) {
    if (g_host) cudaFreeHost(g_host);
    g_host = nullptr; g_host_cap = 0;
    CK(cudaHostAlloc(&g_host, n_out * sizeof(double), cudaHostAllocDefault));
    g_host_cap = n_out;
  }
  return 0;
}

// records already on this rank's GPU (n of them, rec_f64 doubles each, produced on `producer`)
// -> every rank's records, rank-major, padded to n_pad records per rank (pad rows: status = -1
// for the summary layout, zeros otherwise), into `all` (host).
int allgather_device(const double *d_src, long long n, long long n_pad, int rec_f64, bool summary_layout,
                     cudaStream_t producer, double *all) {
  if (!g_comm) { set_error("comm: communicator not initialised (misob200_comm_init)"); return MISOB200_ENCCL; }
  if (n < 0 || n_pad < n) { set_error("comm: padded record count smaller than the local one"); return MISOB200_EINVAL; }
  const size_t per = (size_t) std::max<long long>(n_pad, 1) * rec_f64;
  if (int rc = reserve(per, per * g_ranks)) return rc;
  CK(cudaEventRecord(g_ready, producer));
  CK(cudaStreamWaitEvent(g_stream, g_ready, 0));
  const double *send = d_src;
  if (n != n_pad || !d_src) {       // unequal shards: pad on the device
    CK(cudaMemsetAsync(g_in, 0, per * sizeof(double), g_stream));
    if (n) CK(cudaMemcpyAsync(g_in, d_src, (size_t) n * rec_f64 * sizeof(double), cudaMemcpyDeviceToDevice, g_stream));
    if (summary_layout && n_pad > n) {
      // int32 field 11 of the record's integer tail (f64 index 24 + 5, high half) = status -1
      std::vector<double> tail((size_t) (n_pad - n) * rec_f64, 0.0);
      for (long long r = 0; r < n_pad - n; r++) reinterpret_cast<int *>(&tail[(size_t) r * rec_f64 + 24])[11] = -1;
      CK(cudaMemcpyAsync(g_in + (size_t) n * rec_f64, tail.data(), tail.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
      CK(cudaStreamSynchronize(g_stream));     // `tail` is a stack-lifetime pageable source
    }
    send = g_in;
  }
  NK(N.AllGather(send, g_out, per, ncclFloat64, g_comm, g_stream));
  // straight into the caller's buffer when it is page-locked, else through the pinned mirror
  bool direct = false;
  if (all) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, all) == cudaSuccess && pa.type == cudaMemoryTypeHost) direct = true;
    else cudaGetLastError();
  }
  CK(cudaMemcpyAsync(direct ? all : g_host, g_out, per * g_ranks * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  if (all && !direct) std::memcpy(all, g_host, per * g_ranks * sizeof(double));
  return 0;
}
}  // namespace

// run.cu
int summarize_device(Plan &plan, const double **d_summary, cudaStream_t *stream);
int compare_device(Plan &pa, Plan &pb, const double **d_out, cudaStream_t *stream);

}  // namespace misob200

using namespace misob200;

extern "C" {

int misob200_comm_unique_id(char *id128) {
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  NK(N.GetUniqueId(&id));
  std::memcpy(id128, id.internal, 128);
  return 0;
}

int misob200_comm_init(const char *id128, int n_ranks, int rank) {
  if (int rc = load_nccl()) return rc;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  NK(N.CommInitRank(&g_comm, n_ranks, id, rank));
  CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&g_scalar, 2 * sizeof(double)));
  CK(cudaEventCreateWithFlags(&g_ready, cudaEventDisableTiming));
  g_ranks = n_ranks;
  return 0;
}

// The path's one collective, device to device: summary kernel on this rank's resident
// posteriors -> ncclAllGather straight out of the plan's summary buffer (NVLink) -> one
// device->host copy of the gathered table into pinned memory.  all: n_ranks * n_pad * 32 f64.
int misob200_comm_allgather_summaries(misob200_plan_t *plan, int64_t n_pad, double *all) {
  if (!plan) { set_error("comm_allgather_summaries: null plan"); return MISOB200_EINVAL; }
  const double *d = nullptr;
  cudaStream_t s = nullptr;
  if (int rc = summarize_device(plan->p, &d, &s)) return rc;
  return allgather_device(d, (long long) plan->p.desc.size(), n_pad, MISOB200_SUMMARY_F64, true, s, all);
}

// cfg-5: Bayes-factor records of this rank's events (both samples resident here) gathered
// the same way.  all: n_ranks * n_pad * 32 f64.
int misob200_comm_allgather_compare(misob200_plan_t *plan_a, misob200_plan_t *plan_b, int64_t n_pad, double *all) {
  if (!plan_a || !plan_b) { set_error("comm_allgather_compare: null plan"); return MISOB200_EINVAL; }
  const double *d = nullptr;
  cudaStream_t s = nullptr;
  if (int rc = compare_device(plan_a->p, plan_b->p, &d, &s)) return rc;
  return allgather_device(d, (long long) plan_a->p.desc.size(), n_pad, MISOB200_COMPARE_F64, false, s, all);
}

// mine: n_f64_per_rank doubles in HOST memory; all: n_ranks * n_f64_per_rank doubles,
// rank-major.  (Generic host-buffer form; the data path uses the device-to-device calls above.)
int misob200_comm_allgather(const double *mine, int64_t n, double *all) {
  if (!g_comm) { set_error("comm_allgather: communicator not initialised"); return MISOB200_ENCCL; }
  const size_t per = (size_t) std::max<int64_t>(n, 1);
  if (int rc = reserve(per, per * g_ranks)) return rc;
  CK(cudaMemcpyAsync(g_in, mine, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  NK(N.AllGather(g_in, g_out, (size_t) n, ncclFloat64, g_comm, g_stream));
  CK(cudaMemcpyAsync(all, g_out, n * g_ranks * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

// barrier + max over ranks in one all-reduce
int misob200_comm_barrier_max(double *value) {
  if (!g_comm) { set_error("comm_barrier_max: communicator not initialised"); return MISOB200_ENCCL; }
  CK(cudaMemcpyAsync(g_scalar, value, sizeof(double), cudaMemcpyHostToDevice, g_stream));
  NK(N.AllReduce(g_scalar, g_scalar + 1, 1, ncclFloat64, ncclMax, g_comm, g_stream));
  CK(cudaMemcpyAsync(value, g_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int misob200_comm_destroy(void) {
  if (g_comm) { N.CommDestroy(g_comm); g_comm = nullptr; }
  if (g_scalar) { cudaFree(g_scalar); g_scalar = nullptr; }
  cudaFree(g_in); cudaFree(g_out); g_in = g_out = nullptr; g_in_cap = g_out_cap = 0;
  if (g_host) { cudaFreeHost(g_host); g_host = nullptr; g_host_cap = 0; }
  if (g_ready) { cudaEventDestroy(g_ready); g_ready = nullptr; }
  if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
  return 0;
}

void *misob200_host_alloc(int64_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, (size_t) bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
int misob200_host_free(void *p) {
  if (p) cudaFreeHost(p);
  return 0;
}

}  // extern "C"
