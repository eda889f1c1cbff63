// miso_b200/csrc/match.cu -- read <-> isoform compatibility on the GPU.
//
// The setup stage of the reference spends most of its time deciding, for every read and every
// isoform, whether the read's alignment blocks tile the isoform's exons
// (/root/reference/pysplicing/src/solve.c:8-108 splicing_matchIso, :141-218 _paired, :220-306
// splicing_parse_cigar; src/gff.c:1041-1084 for the fragment length of a pair).  It is integer
// and string work, embarrassingly parallel over reads, and bound by the bytes it has to touch:
// per read a 4-byte position, an 8-byte offset and a CIGAR string of a few characters in, 2K
// bytes of codes out.  One thread per read (pair); a CTA walks the reads of one gene at a time
// so that positions, offsets and strings are read as contiguous runs and the gene's small exon
// table stays in L1; CTAs take genes from an atomic counter.  The arithmetic is
// match_core.hpp, the same functions the host plan stage compiles.
//
// The sort into draw order and the class/tile packing stay on the host (plan.cpp): the draw
// order must reproduce the reference's unstable qsort exactly.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "match_core.hpp"
#include "plan.hpp"

namespace misob200 {

namespace {

struct MatchArgs {
  int n_genes;
  const int32_t *iso_off, *exon_off, *exon_start, *exon_end;
  const int64_t *read_off;
  const int32_t *position;
  const int64_t *cigar_off;
  const char *cigar;
  const long long *code_off;
  MatchParams mp;
  uint16_t *codes;
  int *status;
  unsigned *next_gene;
};

constexpr int kMatchThreads = 256;

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
    uint16_t *out = a.codes + a.code_off[g];
    bool bad = false;
    for (int r = threadIdx.x; r < R; r += kMatchThreads) {
      uint16_t col[kMaxIso];
      if (match_read(gv, a.mp, a.position + r0, a.cigar_off + r0, a.cigar, r, col)) bad = true;
      // K <= 8 contiguous 16-bit codes per read: one 16-byte store when K = 8
      uint16_t *dst = out + (size_t) r * K;
#pragma unroll
      for (int k = 0; k < kMaxIso; k++)
        if (k < K) dst[k] = col[k];
    }
    if (bad) atomicOr(a.status + g, MISOB200_EINVAL);
  }
}

#define MCK(call)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      set_error(std::string("match_on_device: ") + cudaGetErrorString(e_));                \
      cleanup();                                                                           \
      return MISOB200_ECUDA;                                                               \
    }                                                                                      \
  } while (0)

thread_local double t_kernel_ms = 0, t_h2d_ms = 0, t_d2h_ms = 0;
thread_local long long t_bytes_in = 0, t_bytes_out = 0;

}  // namespace

void last_match_stats(double *kernel_ms, double *h2d_ms, double *d2h_ms, long long *bytes_in, long long *bytes_out) {
  if (kernel_ms) *kernel_ms = t_kernel_ms;
  if (h2d_ms) *h2d_ms = t_h2d_ms;
  if (d2h_ms) *d2h_ms = t_d2h_ms;
  if (bytes_in) *bytes_in = t_bytes_in;
  if (bytes_out) *bytes_out = t_bytes_out;
}

int match_on_device(const misob200_reads_t &in, const MatchParams &mp, int device, DeviceCodes &out) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error("match_on_device: no CUDA device -- miso_b200 has no CPU path for device matching");
    return MISOB200_ECUDA;
  }
  if (device < 0 || device >= n_dev) { set_error("match_on_device: device ordinal out of range"); return MISOB200_EINVAL; }
  const int G = in.n_genes;
  out.code_off.assign((size_t) G + 1, 0);
  out.status.assign((size_t) std::max(G, 1), 0);
  if (G == 0) { out.codes_store.clear(); out.codes = out.codes_store.data(); return 0; }
  const int n_iso = in.iso_off[G];
  const int n_exon = in.exon_off[n_iso];
  const long long n_reads = in.read_off[G];
  const long long n_cig = n_reads > 0 ? in.cigar_off[n_reads] : 0;
  for (int g = 0; g < G; g++) {
    const long long K = in.iso_off[g + 1] - in.iso_off[g];
    const long long nr = in.read_off[g + 1] - in.read_off[g];
    const long long R = in.paired ? nr / 2 : nr;
    out.code_off[g + 1] = out.code_off[g] + (K >= 1 && K <= kMaxIso ? K * R : 0);
  }
  const long long n_codes = out.code_off[G];
  out.codes_store.assign((size_t) std::max<long long>(n_codes, 1), 0);
  out.codes = out.codes_store.data();

  // device buffers: 0-8 inputs (order of `src`), 9 codes, 10 per-gene status + the work counter
  void *d[11] = {nullptr};
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  auto cleanup = [&]() {
    for (auto &p : d) if (p) cudaFree(p);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  };
  MCK(cudaSetDevice(device));
  MCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (auto &e : ev) MCK(cudaEventCreate(&e));
  const size_t sz[11] = {
      (size_t) (G + 1) * 4, (size_t) (n_iso + 1) * 4, (size_t) std::max(n_exon, 1) * 4, (size_t) std::max(n_exon, 1) * 4,
      (size_t) (G + 1) * 8, (size_t) std::max<long long>(n_reads, 1) * 4, (size_t) (n_reads + 1) * 8,
      (size_t) std::max<long long>(n_cig, 1), (size_t) (G + 1) * 8, (size_t) std::max<long long>(n_codes, 1) * 2,
      (size_t) G * 4 + 4};
  const void *src[9] = {in.iso_off, in.exon_off, in.exon_start, in.exon_end, in.read_off, in.position, in.cigar_off,
                        in.cigar, out.code_off.data()};
  for (int i = 0; i < 11; i++) MCK(cudaMalloc(&d[i], sz[i]));
  int *d_status = static_cast<int *>(d[10]);
  unsigned *d_next = reinterpret_cast<unsigned *>(d_status + G);
  MCK(cudaMemsetAsync(d_status, 0, (size_t) G * 4 + 4, stream));
  MCK(cudaEventRecord(ev[0], stream));
  long long bytes_in = 0;
  for (int i = 0; i < 9; i++) {
    if (!src[i]) continue;
    MCK(cudaMemcpyAsync(d[i], src[i], sz[i], cudaMemcpyHostToDevice, stream));
    bytes_in += (long long) sz[i];
  }
  MCK(cudaEventRecord(ev[1], stream));

  MatchArgs a;
  a.n_genes = G;
  a.iso_off = (const int32_t *) d[0]; a.exon_off = (const int32_t *) d[1];
  a.exon_start = (const int32_t *) d[2]; a.exon_end = (const int32_t *) d[3];
  a.read_off = (const int64_t *) d[4]; a.position = (const int32_t *) d[5];
  a.cigar_off = (const int64_t *) d[6]; a.cigar = (const char *) d[7];
  a.code_off = (const long long *) d[8];
  a.mp = mp;
  a.codes = (uint16_t *) d[9];
  a.status = d_status;
  a.next_gene = d_next;
  cudaDeviceProp prop;
  MCK(cudaGetDeviceProperties(&prop, device));
  const int blocks = std::min(G, prop.multiProcessorCount * 8);      // 8 x 256 threads per SM: 2048 resident threads
  match_kernel<<<blocks, kMatchThreads, 0, stream>>>(a);
  MCK(cudaGetLastError());
  MCK(cudaEventRecord(ev[2], stream));
  MCK(cudaMemcpyAsync(out.codes_store.data(), d[9], (size_t) std::max<long long>(n_codes, 1) * 2, cudaMemcpyDeviceToHost, stream));
  MCK(cudaMemcpyAsync(out.status.data(), d_status, (size_t) G * 4, cudaMemcpyDeviceToHost, stream));
  MCK(cudaEventRecord(ev[3], stream));
  MCK(cudaStreamSynchronize(stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); out.h2d_ms = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); out.kernel_ms = ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); out.d2h_ms = ms;
  out.bytes_in = bytes_in;
  out.bytes_out = n_codes * 2;
  t_kernel_ms = out.kernel_ms; t_h2d_ms = out.h2d_ms; t_d2h_ms = out.d2h_ms;
  t_bytes_in = out.bytes_in; t_bytes_out = out.bytes_out;
  cleanup();
  return 0;
}

}  // namespace misob200
