// miso_b200/csrc/match.cu -- the O(reads) half of the setup stage on the GPU:
// read <-> isoform compatibility and the draw order.
//
// The setup stage of the reference spends most of its time deciding, for every read and every
// isoform, whether the read's alignment blocks tile the isoform's exons
// (/root/reference/pysplicing/src/solve.c:8-108 splicing_matchIso, :141-218 _paired, :220-306
// splicing_parse_cigar; src/gff.c:1041-1084 for the fragment length of a pair), and then sorting
// the reads by their probability column (splicing_order_matches, src/miso.c:988-993).
//
//   match_kernel   one thread per read (pair); a CTA walks the reads of one gene at a time so that
//                  positions, offsets and strings are read as contiguous runs and the gene's small
//                  exon table stays in L1; CTAs take genes from an atomic counter.  The arithmetic
//                  is match_core.hpp, the same functions the host plan stage compiles.  Integer
//                  and string work, bound by instruction issue (divergent per-thread parsing), not
//                  by the ~34 bytes per pair it reads.
//   order_kernel   one warp per gene: the lanes pack each read's column into an integer sort key
//                  (dense ranks of the code probabilities, plan.cpp) in shared memory, then lane 0
//                  runs the reference's Bentley-McIlroy quicksort on the index array -- the very
//                  code of bm_sort.hpp the host compiles, because the tie order of this unstable
//                  sort is observable.  ~10^5 dependent shared-memory steps per gene,
//                  thousands of genes in flight.
//
// Codes and order come back to pinned staging buffers; read classes and tile packing stay on the
// host (plan.cpp).  All device and pinned buffers are grow-only and reused from call to call.
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "bm_sort.hpp"
#include "match_core.hpp"
#include "plan.hpp"

namespace misob200 {

namespace {

struct MatchArgs {
  int n_genes;
  const int32_t *iso_off, *exon_off, *exon_start, *exon_end;
  const int64_t *read_off;
  const int32_t *position;
  const uint32_t *cigar_off;      // relative to the batch's first CIGAR byte
  const char *cigar;
  const long long *code_off;
  MatchParams mp;
  void *codes;                    // u8 (narrow insert models) or u16, [read][isoform]
  int *status;
  unsigned *next_gene;
};

constexpr int kMatchThreads = 256;

template <class Code>
__global__ void __launch_bounds__(kMatchThreads) match_kernel(const MatchArgs a) {
  __shared__ int s_gene;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_gene = (int) atomicAdd(a.next_gene, 1u);
    __syncthreads();
    const int g = s_gene;
    if (g >= a.n_genes) return;
    const int iso0 = a.iso_off[g], K = a.iso_off[g + 1] - iso0;
    if (K < 1 || K > kMaxIso) continue;          // the host reports these genes (plan.cpp)
    const long long r0 = a.read_off[g], nr = a.read_off[g + 1] - r0;
    const int R = (int) (a.mp.paired ? nr / 2 : nr);
    const IsoView gv{K, a.exon_off + iso0, a.exon_start, a.exon_end};
    Code *out = static_cast<Code *>(a.codes) + a.code_off[g];
    bool bad = false;
    for (int r = threadIdx.x; r < R; r += kMatchThreads) {
      Code col[kMaxIso];
      if (match_read(gv, a.mp, a.position + r0, a.cigar_off + r0, a.cigar, r, col)) bad = true;
      Code *dst = out + (size_t) r * K;
#pragma unroll
      for (int k = 0; k < kMaxIso; k++)
        if (k < K) dst[k] = col[k];
    }
    if (bad) atomicOr(a.status + g, MISOB200_EINVAL);
  }
}

struct OrderArgs {
  int n_genes, paired, r_cap;
  const int32_t *iso_off;
  const int64_t *read_off;
  const long long *code_off, *pair_off;
  const void *codes;
  const uint16_t *rank;           // dense rank of ptab[code]
  int32_t *order;                 // [pair_off[g] + i] = read drawn i-th; order[pair_off[g]] = -1: not sorted here
  unsigned *next_gene;
};

template <class KeyT, int BITS, class Code>
__global__ void __launch_bounds__(32) order_kernel(const OrderArgs a) {
  extern __shared__ __align__(16) unsigned char sm[];
  KeyT *key = reinterpret_cast<KeyT *>(sm);
  int32_t *ord = reinterpret_cast<int32_t *>(sm + (size_t) a.r_cap * sizeof(KeyT));
  const int lane = threadIdx.x;
  while (true) {
    int g = 0;
    if (lane == 0) g = (int) atomicAdd(a.next_gene, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= a.n_genes) return;
    const int K = a.iso_off[g + 1] - a.iso_off[g];
    if (K < 1 || K > kMaxIso) continue;
    const long long nr = a.read_off[g + 1] - a.read_off[g];
    const int R = (int) (a.paired ? nr / 2 : nr);
    int32_t *out = a.order + a.pair_off[g];
    if (R > a.r_cap) {                       // does not fit the shared-memory arrays: the host sorts this gene
      if (lane == 0 && R > 0) out[0] = -1;
      continue;
    }
    const Code *codes = static_cast<const Code *>(a.codes) + a.code_off[g];
    for (int r = lane; r < R; r += 32) {
      KeyT v = 0;
      for (int k = 0; k < K; k++) v = (v << BITS) | (KeyT) a.rank[codes[(size_t) r * K + k]];
      key[r] = v;
      ord[r] = r;
    }
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
      BMSort<KeyCmp<KeyT>> sorter{KeyCmp<KeyT>{key}};
      ok = sorter.sort(ord, R);
    }
    ok = __shfl_sync(0xffffffffu, (int) ok, 0) != 0;
    __syncwarp();
    for (int r = lane; r < R; r += 32) out[r] = ord[r];
    if (!ok && lane == 0 && R > 0) out[0] = -1;
    __syncwarp();
  }
}

// ---- persistent buffers ------------------------------------------------------------------
// kStages independent sets of device + pinned buffers, so that the batches of a pipeline can be in
// different phases at once (miso_b200/pipeline.py): one being copied in, one in the kernels or
// copying out, one being finished by the host threads.  A stage is acquired by stage_submit and
// released by stage_release; submit blocks while all are taken.
struct Stage {
  bool busy = false;
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[5] = {};
  void *d[12] = {};
  size_t cap[12] = {};
  void *h[3] = {};                // pinned: 0 all inputs (CIGAR offsets as u32), 1 codes (out), 2 order + status (out)
  size_t hcap[3] = {};
};
constexpr int kStages = 3;
std::mutex g_mu;
std::condition_variable g_cv;
Stage g_stage[kStages];

thread_local double t_kernel_ms = 0, t_h2d_ms = 0, t_d2h_ms = 0;
thread_local long long t_bytes_in = 0, t_bytes_out = 0;

#define MCK(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      set_error(std::string("match_on_device: ") + cudaGetErrorString(e_));                \
      return MISOB200_ECUDA;                                                               \
    }                                                                                      \
  } while (0)

int pool_device(Stage &p, int device) {
  MCK(cudaSetDevice(device));
  if (p.device == device) return 0;
  if (p.device >= 0) {            // a different GPU than last time: start over
    cudaSetDevice(p.device);
    for (auto &x : p.d) { cudaFree(x); x = nullptr; }
    for (auto &c : p.cap) c = 0;
    if (p.stream) cudaStreamDestroy(p.stream);
    for (auto &e : p.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    p.stream = nullptr;
    MCK(cudaSetDevice(device));
  }
  p.device = device;
  MCK(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
  for (auto &e : p.ev) MCK(cudaEventCreate(&e));
  return 0;
}
int pool_dev(Stage &p, int i, size_t bytes) {
  bytes = std::max<size_t>(bytes, 16);
  if (bytes <= p.cap[i]) return 0;
  cudaFree(p.d[i]); p.d[i] = nullptr; p.cap[i] = 0;
  MCK(cudaMalloc(&p.d[i], bytes + bytes / 8));
  p.cap[i] = bytes + bytes / 8;
  return 0;
}
int pool_host(Stage &p, int i, size_t bytes) {
  bytes = std::max<size_t>(bytes, 16);
  if (bytes <= p.hcap[i]) return 0;
  if (p.h[i]) cudaFreeHost(p.h[i]);
  p.h[i] = nullptr; p.hcap[i] = 0;
  MCK(cudaHostAlloc(&p.h[i], bytes + bytes / 8, cudaHostAllocDefault));
  p.hcap[i] = bytes + bytes / 8;
  return 0;
}

}  // namespace

void last_match_stats(double *kernel_ms, double *h2d_ms, double *d2h_ms, long long *bytes_in, long long *bytes_out) {
  if (kernel_ms) *kernel_ms = t_kernel_ms;
  if (h2d_ms) *h2d_ms = t_h2d_ms;
  if (d2h_ms) *d2h_ms = t_d2h_ms;
  if (bytes_in) *bytes_in = t_bytes_in;
  if (bytes_out) *bytes_out = t_bytes_out;
}

// frees the buffers of every idle stage (misob200_shutdown)
void stage_pool_release() {
  std::lock_guard<std::mutex> lock(g_mu);
  for (Stage &p : g_stage) {
    if (p.busy || p.device < 0) continue;
    cudaSetDevice(p.device);
    for (auto &x : p.d) { cudaFree(x); x = nullptr; }
    for (auto &c : p.cap) c = 0;
    for (auto &x : p.h) { if (x) cudaFreeHost(x); x = nullptr; }
    for (auto &c : p.hcap) c = 0;
    if (p.stream) { cudaStreamDestroy(p.stream); p.stream = nullptr; }
    for (auto &e : p.ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    p.device = -1;
  }
}

void stage_release(int stage) {
  if (stage < 0 || stage >= kStages) return;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    g_stage[stage].busy = false;
  }
  g_cv.notify_all();
}

// Phase 1: copies in, kernels and copies out are ENQUEUED on the stage's stream; the call returns
// when the pageable inputs have been handed to the driver.  code_rank: dense rank of ptab[code]
// (empty: no device sort, the host orders the reads).  out.stage identifies the stage for
// stage_wait / stage_release.
int stage_submit(const misob200_reads_t &in, const MatchParams &mp, int device, int n_codes,
                 const std::vector<uint16_t> &code_rank, DeviceCodes &out) {
  out.stage = -1;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error("match_on_device: no CUDA device -- miso_b200 has no CPU path for device matching");
    return MISOB200_ECUDA;
  }
  if (device < 0 || device >= n_dev) { set_error("match_on_device: device ordinal out of range"); return MISOB200_EINVAL; }
  const int G = in.n_genes;
  out.code_off.assign((size_t) G + 1, 0);
  out.pair_off.assign((size_t) G + 1, 0);
  out.status.assign((size_t) std::max(G, 1), 0);
  out.codes16 = nullptr; out.codes8 = nullptr; out.order = nullptr;
  if (G == 0) return 0;
  const int n_iso = in.iso_off[G];
  const int n_exon = in.exon_off[n_iso];
  const long long n_reads = in.read_off[G];
  const long long cig0 = n_reads > 0 ? in.cigar_off[0] : 0;
  const long long n_cig = n_reads > 0 ? in.cigar_off[n_reads] - cig0 : 0;
  if (n_cig >= (1LL << 32)) { set_error("match_on_device: more than 4 GiB of CIGAR text in one batch; append in smaller batches"); return MISOB200_EINVAL; }
  int r_max = 0;
  for (int g = 0; g < G; g++) {
    const long long K = in.iso_off[g + 1] - in.iso_off[g];
    const long long nr = in.read_off[g + 1] - in.read_off[g];
    const long long R = in.paired ? nr / 2 : nr;
    const bool ok = K >= 1 && K <= kMaxIso;
    out.code_off[g + 1] = out.code_off[g] + (ok ? K * R : 0);
    out.pair_off[g + 1] = out.pair_off[g] + (ok ? R : 0);
    if (ok) r_max = (int) std::max<long long>(r_max, R);
  }
  const long long n_codes_total = out.code_off[G], n_pairs = out.pair_off[G];
  const bool narrow = n_codes <= 256;
  const size_t code_bytes = (size_t) n_codes_total * (narrow ? 1 : 2);
  const bool do_sort = (int) code_rank.size() == n_codes && n_codes > 0;
  int max_rank = 0;
  for (uint16_t r : code_rank) max_rank = std::max<int>(max_rank, r);
  const bool key64 = max_rank < 256;

  int stage = -1;
  {
    std::unique_lock<std::mutex> lock(g_mu);
    g_cv.wait(lock, [&] { for (int i = 0; i < kStages; i++) if (!g_stage[i].busy) { stage = i; return true; } return false; });
    g_stage[stage].busy = true;
  }
  out.stage = stage;
  Stage &P = g_stage[stage];
  struct Guard { int s; bool armed; ~Guard() { if (armed) stage_release(s); } } guard{stage, true};
  if (int rc = pool_device(P, device)) return rc;
  // device buffers: 0-3 gene structure, 4 read_off, 5 position, 6 u32 CIGAR offsets, 7 CIGAR text, 8 code_off,
  // 9 codes, 10 per-gene status + two work counters, 11 pair_off + order + rank table
  const size_t sz[12] = {
      (size_t) (G + 1) * 4, (size_t) (n_iso + 1) * 4, (size_t) std::max(n_exon, 1) * 4, (size_t) std::max(n_exon, 1) * 4,
      (size_t) (G + 1) * 8, (size_t) std::max<long long>(n_reads, 1) * 4, (size_t) (n_reads + 1) * 4,
      (size_t) std::max<long long>(n_cig, 1), (size_t) (G + 1) * 8, code_bytes, (size_t) G * 4 + 8,
      (size_t) (G + 1) * 8 + (size_t) std::max<long long>(n_pairs, 1) * 4 + (size_t) std::max(n_codes, 1) * 2 + 64};
  for (int i = 0; i < 12; i++) if (int rc = pool_dev(P, i, sz[i])) return rc;
  // every input goes through ONE pinned staging buffer (offsets 256-byte aligned): the copies to the
  // device are then truly asynchronous -- a pageable source makes cudaMemcpyAsync hold the calling
  // thread (and, for hundreds of MB, other threads' CUDA calls) until the data has been staged
  size_t in_off[11], in_total = 0;
  const size_t in_sz[11] = {sz[0], sz[1], sz[2], sz[3], sz[4], sz[5], sz[6], sz[7], sz[8], (size_t) (G + 1) * 8,
                            (size_t) std::max(n_codes, 1) * 2};
  for (int i = 0; i < 11; i++) { in_off[i] = in_total; in_total += (in_sz[i] + 255) & ~(size_t) 255; }
  if (int rc = pool_host(P, 0, in_total)) return rc;
  if (int rc = pool_host(P, 1, code_bytes)) return rc;
  if (int rc = pool_host(P, 2, (size_t) std::max<long long>(n_pairs, 1) * 4 + (size_t) G * 4)) return rc;      // order, then status
  unsigned char *hin = static_cast<unsigned char *>(P.h[0]);
  {
    const void *small_src[11] = {in.iso_off, in.exon_off, in.exon_start, in.exon_end, in.read_off, nullptr, nullptr, nullptr,
                                 out.code_off.data(), out.pair_off.data(), do_sort ? code_rank.data() : nullptr};
    const size_t small_n[11] = {(size_t) (G + 1) * 4, (size_t) (n_iso + 1) * 4, (size_t) n_exon * 4, (size_t) n_exon * 4, (size_t) (G + 1) * 8,
                                0, 0, 0, (size_t) (G + 1) * 8, (size_t) (G + 1) * 8, do_sort ? (size_t) n_codes * 2 : 0};
    for (int i = 0; i < 11; i++) if (small_src[i] && small_n[i]) std::memcpy(hin + in_off[i], small_src[i], small_n[i]);
    // the three big ones on the worker threads: positions, CIGAR text, and the CIGAR offsets as 32-bit
    // offsets relative to the batch's text (half the bytes of the 64-bit ABI array)
    uint32_t *o32 = reinterpret_cast<uint32_t *>(hin + in_off[6]);
    const int nt = n_reads < (1 << 20) ? 1 : std::max(1, std::min(host_threads(), 8));
    auto part = [&](int t) {
      const long long a = (n_reads + 1) * t / nt, b = (n_reads + 1) * (t + 1) / nt;
      for (long long i = a; i < b; i++) o32[i] = (uint32_t) (in.cigar_off[i] - cig0);
      const long long pa = n_reads * t / nt, pb = n_reads * (t + 1) / nt;
      if (pb > pa) std::memcpy(hin + in_off[5] + (size_t) pa * 4, in.position + pa, (size_t) (pb - pa) * 4);
      const long long ca = n_cig * t / nt, cb = n_cig * (t + 1) / nt;
      if (cb > ca) std::memcpy(hin + in_off[7] + (size_t) ca, in.cigar + cig0 + ca, (size_t) (cb - ca));
    };
    if (nt == 1) part(0);
    else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; t++) pool.emplace_back(part, t);
      for (auto &t : pool) t.join();
    }
  }
  int *d_status = static_cast<int *>(P.d[10]);
  unsigned *d_next = reinterpret_cast<unsigned *>(d_status + G);
  MCK(cudaMemsetAsync(d_status, 0, (size_t) G * 4 + 8, P.stream));
  MCK(cudaEventRecord(P.ev[0], P.stream));
  long long bytes_in = 0;
  for (int i = 0; i < 9; i++) {
    MCK(cudaMemcpyAsync(P.d[i], hin + in_off[i], sz[i], cudaMemcpyHostToDevice, P.stream));
    bytes_in += (long long) sz[i];
  }
  unsigned char *d11 = static_cast<unsigned char *>(P.d[11]);
  long long *d_pair_off = reinterpret_cast<long long *>(d11);
  int32_t *d_order = reinterpret_cast<int32_t *>(d11 + (size_t) (G + 1) * 8);
  uint16_t *d_rank = reinterpret_cast<uint16_t *>(d11 + (size_t) (G + 1) * 8 + (size_t) std::max<long long>(n_pairs, 1) * 4);
  MCK(cudaMemcpyAsync(d_pair_off, hin + in_off[9], (size_t) (G + 1) * 8, cudaMemcpyHostToDevice, P.stream));
  if (do_sort) MCK(cudaMemcpyAsync(d_rank, hin + in_off[10], (size_t) n_codes * 2, cudaMemcpyHostToDevice, P.stream));
  MCK(cudaEventRecord(P.ev[1], P.stream));

  MatchArgs a;
  a.n_genes = G;
  a.iso_off = (const int32_t *) P.d[0]; a.exon_off = (const int32_t *) P.d[1];
  a.exon_start = (const int32_t *) P.d[2]; a.exon_end = (const int32_t *) P.d[3];
  a.read_off = (const int64_t *) P.d[4]; a.position = (const int32_t *) P.d[5];
  a.cigar_off = (const uint32_t *) P.d[6]; a.cigar = (const char *) P.d[7];
  a.code_off = (const long long *) P.d[8];
  a.mp = mp;
  a.codes = P.d[9];
  a.status = d_status;
  a.next_gene = d_next;
  cudaDeviceProp prop;
  MCK(cudaGetDeviceProperties(&prop, device));
  const int blocks = std::min(G, prop.multiProcessorCount * 8);      // 8 x 256 threads per SM: 2048 resident threads
  if (narrow) match_kernel<uint8_t><<<blocks, kMatchThreads, 0, P.stream>>>(a);
  else match_kernel<uint16_t><<<blocks, kMatchThreads, 0, P.stream>>>(a);
  MCK(cudaGetLastError());
  MCK(cudaEventRecord(P.ev[2], P.stream));

  bool sorted = false;
  if (do_sort && n_pairs > 0) {
    // shared memory per warp-CTA: r_cap keys + r_cap indices; at most 96 KB, so that several genes share an SM
    const size_t per_read = (key64 ? 8 : 16) + 4;
    const int r_cap = (int) std::min<long long>(r_max, (96 * 1024) / (long long) per_read);
    const size_t smem = (size_t) std::max(r_cap, 1) * per_read;
    OrderArgs o;
    o.n_genes = G; o.paired = mp.paired; o.r_cap = r_cap;
    o.iso_off = a.iso_off; o.read_off = a.read_off; o.code_off = a.code_off; o.pair_off = d_pair_off;
    o.codes = P.d[9]; o.rank = d_rank; o.order = d_order; o.next_gene = d_next + 1;
    const void *kern = key64 ? (narrow ? (const void *) order_kernel<unsigned long long, 8, uint8_t>
                                       : (const void *) order_kernel<unsigned long long, 8, uint16_t>)
                             : (narrow ? (const void *) order_kernel<unsigned __int128, 16, uint8_t>
                                       : (const void *) order_kernel<unsigned __int128, 16, uint16_t>);
    MCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0;
    MCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    const int oblocks = std::min(G, std::max(1, per_sm) * prop.multiProcessorCount);
    void *args[] = {&o};
    MCK(cudaLaunchKernel(kern, dim3(oblocks), dim3(32), args, smem, P.stream));
    sorted = true;
  }
  MCK(cudaEventRecord(P.ev[3], P.stream));
  MCK(cudaMemcpyAsync(P.h[1], P.d[9], std::max<size_t>(code_bytes, 1), cudaMemcpyDeviceToHost, P.stream));
  if (sorted) MCK(cudaMemcpyAsync(P.h[2], d_order, (size_t) n_pairs * 4, cudaMemcpyDeviceToHost, P.stream));
  // (status goes to pinned memory too: a pageable destination would make this call wait for the kernels)
  out.n_pairs = n_pairs;
  MCK(cudaMemcpyAsync(static_cast<char *>(P.h[2]) + (size_t) std::max<long long>(n_pairs, 1) * 4, d_status, (size_t) G * 4,
                      cudaMemcpyDeviceToHost, P.stream));
  MCK(cudaEventRecord(P.ev[4], P.stream));
  if (narrow) out.codes8 = static_cast<const uint8_t *>(P.h[1]); else out.codes16 = static_cast<const uint16_t *>(P.h[1]);
  out.order = sorted ? static_cast<const int32_t *>(P.h[2]) : nullptr;
  out.bytes_in = bytes_in;
  out.bytes_out = (long long) code_bytes + (sorted ? n_pairs * 4 : 0);
  guard.armed = false;
  return 0;
}

// Phase 2: wait for the stage's copies and kernels; codes / order / status are then readable.
int stage_wait(DeviceCodes &out) {
  if (out.stage < 0 || out.stage >= kStages) { set_error("match stage: no batch in flight"); return MISOB200_EINVAL; }
  Stage &P = g_stage[out.stage];
  MCK(cudaSetDevice(P.device));
  MCK(cudaStreamSynchronize(P.stream));
  std::memcpy(out.status.data(), static_cast<char *>(P.h[2]) + (size_t) std::max<long long>(out.n_pairs, 1) * 4,
              out.status.size() * sizeof(int));
  float ms = 0;
  cudaEventElapsedTime(&ms, P.ev[0], P.ev[1]); out.h2d_ms = ms;
  cudaEventElapsedTime(&ms, P.ev[1], P.ev[2]); out.kernel_ms = ms;
  cudaEventElapsedTime(&ms, P.ev[2], P.ev[3]); out.sort_ms = ms;
  cudaEventElapsedTime(&ms, P.ev[3], P.ev[4]); out.d2h_ms = ms;
  t_kernel_ms = out.kernel_ms + out.sort_ms; t_h2d_ms = out.h2d_ms; t_d2h_ms = out.d2h_ms;
  t_bytes_in = out.bytes_in; t_bytes_out = out.bytes_out;
  return 0;
}

}  // namespace misob200
