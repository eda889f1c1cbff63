// miso_b200/csrc/comm.cu -- the one collective of the path: an all-gather of the
// per-gene posterior summary records over NCCL (NVLink 5 / NVSwitch).
//
// Genes are independent (the reference exploits this with OS processes,
// /root/reference/misopy/miso.py:163-188), so sampling needs no exchange; only
// the fixed-size summaries are gathered, once, at the end.  NCCL is loaded with
// dlopen so that a single-GPU process never needs it.
#include <dlfcn.h>

#include <cstring>
#include <string>

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
}  // namespace

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
  g_ranks = n_ranks;
  return 0;
}

// mine: n_f64_per_rank doubles in HOST memory (the summaries misob200_summarize
// returned); all: n_ranks * n_f64_per_rank doubles, rank-major.
int misob200_comm_allgather(const double *mine, int64_t n, double *all) {
  if (!g_comm) { set_error("comm_allgather: communicator not initialised"); return MISOB200_ENCCL; }
  double *d_in = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_in, std::max<int64_t>(n, 1) * sizeof(double)));
  CK(cudaMalloc(&d_out, std::max<int64_t>(n, 1) * g_ranks * sizeof(double)));
  CK(cudaMemcpyAsync(d_in, mine, n * sizeof(double), cudaMemcpyHostToDevice, g_stream));
  NK(N.AllGather(d_in, d_out, (size_t) n, ncclFloat64, g_comm, g_stream));
  CK(cudaMemcpyAsync(all, d_out, n * g_ranks * sizeof(double), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  cudaFree(d_in); cudaFree(d_out);
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
