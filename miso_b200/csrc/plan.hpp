// miso_b200/csrc/plan.hpp -- host-side plan: per-gene packed inputs of the chain kernel.
//
// A "plan" is everything the setup stage of the reference computes once per
// gene before its MCMC loop (SURVEY.md section 8a, rows a-12 ... a-18), laid
// out for the GPU:
//   * compatibility codes   (splicing_matchIso / _paired, src/solve.c:8-218)
//   * draw order            (splicing_order_matches, src/miso.c:988-993)
//   * read classes          (splicing_i_miso_classes, src/miso_paired.c:576-681)
//   * effective lengths, score tables (src/miso.c:773-784, miso_paired.c:396-419)
// Paths are relative to /root/reference/pysplicing/.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/miso_b200.h"

namespace misob200 {

constexpr int kMaxIso = MISOB200_MAX_ISO;
constexpr int kMaxCodes = 4096;    // ptab entries that fit the shared-memory table (32 KB)
constexpr int kTilePadFront = 3;   // bytes in front of rank 0 (stream misalignment, see chain_kernel.cu)
constexpr int kMaxClasses = 254;   // class ids are bytes; id ncls is the null class of the padding

// Device-visible per-gene descriptor (one per gene, 16-byte aligned, POD).
struct GeneDesc {
  unsigned long long tile_off;   // byte offset of this gene's tile in the tile arena
  long long sample_off;          // element offset into samples (f64)
  long long loglik_off;          // element offset into loglik (f64)
  long long drawn_off;           // element offset into the drawn-assignment arena (u8)
  int K, R2, row_bytes, paired;   // row_bytes: one code row (u8 codes, or u16 when the plan is "wide")
  int n_fixed[kMaxIso];          // reads with exactly one compatible isoform
  int L[kMaxIso];                // PE: lp(j) = L[k] - j  (miso_paired.c:409-410)
  double offset[kMaxIso];        // SE log(effisolen_k) (miso.c:136) / PE assscores_k (miso_paired.c:418)
  double hyper_m1[kMaxIso];      // alpha_k - 1 (miso.c:174)
  double rs_se[kMaxIso];         // SE isoscores_k = -log(l_k) (miso.c:783)
  double lg_sum, lg_each;        // lgamma(sum alpha), sum lgamma(alpha_k) (miso.c:177-178)
  double sigma, sd, covar_const; // miso.c:328, :188, :101
  double rp_fixed;               // read score of the single-isoform reads
  unsigned gene_id;
  int rp_always;                 // some read score is not finite: keep readProb in every MH ratio
  int status;
  int flag_off;                  // byte offset of the flag row (always u8) inside the tile
  int tile_bytes;                // whole tile, multiple of 16
  int core_bytes;                // class format: id row + class records (what quad_kernel.cuh keeps in shared memory)
  int lp_safe;                   // paired-end: every lp = L_k - (code - 1) a drawing read can produce lies in [lp_min, lp_max], lp_min >= 1
  // ---- class format (format == 1, class_kernel.cuh) -------------------------
  int format;                    // 0: dense code rows + flag row; 1: class ids + uniform codes + class records
  int ncls;                      // weight classes among the drawing reads (<= kMaxClasses)
  int cls_off;                   // byte offset of the class records (ncls x 8 u16 ptab indices, then ncls + 1 u32 meta)
  int ucode_off;                 // byte offset of the uniform-code row (u8, or u16 when the plan is "wide")
  int lp_min, lp_max, pad2_[2];  // see lp_safe
  int g_always[kMaxIso];         // drawing reads whose first compatible isoform comes after k (their test k
                                 // is true whatever the uniform: the cumulative sum is still an exact 0)
};
static_assert(sizeof(GeneDesc) % 16 == 0, "GeneDesc is copied as a 16-byte aligned POD");

struct GeneHost {
  int K = 0, R = 0, R2 = 0, ncls = 0, status = 0;
  long long read_base = 0;               // first read/pair of this gene in the plan-wide assignment output
  std::vector<double> class_templates;   // ncls x K, row per class
  std::vector<double> class_counts;
  std::vector<int32_t> rank_read;        // rank -> read index (draw order of the reads that draw)
  std::vector<int8_t> fixed_ass;         // per read: -1 incompatible, k single isoform, -2 drawn
  std::vector<int32_t> codes, order;     // kept only when keep_match is on (parity tests)
};

struct Plan {
  bool keep_match = false;
  int paired = -1;                       // fixed by the first append
  int read_len = 0, overhang = 1;
  double frag_mean = 0, frag_var = 0, num_devs = 0;
  int frag_start = 0, frag_len_n = 0;
  std::vector<double> ptab;              // ptab[0] = 0, ptab[j+1] = fragment prob j (SE: {0,1})
  std::vector<uint16_t> code_rank;       // dense rank of ptab[code] (plan.cpp KeyCmp: the draw-order sort key)
  bool wide = false;                     // more than 255 fragment lengths: 16-bit codes
  int force_format = -1;                 // tests: 0 = dense tiles only, 1 = class tiles whenever possible (default)
  std::vector<GeneDesc> desc;
  std::vector<GeneHost> host;
  std::vector<uint8_t> tiles;            // tile arena, each tile 16-byte aligned
  long long n_reads = 0;
  long long n_drawn = 0;
  // output layout cache (run.cu plan_layout): valid for (lay_chains, lay_S) and lay_genes genes
  int lay_chains = -1;
  long long lay_S = -1, lay_n_samples = 0, lay_n_loglik = 0;
  size_t lay_genes = 0;
  long long lay_generation = 0;          // bumped whenever the descriptors' output offsets are rewritten
  long long lay_range[2 * (kMaxIso + 1) + 1][4] = {};
  // device side (owned by run.cu)
  void *dev = nullptr;
};

// Compatibility codes and draw order computed by match.cu for a whole batch.  The arrays live in
// the pinned staging buffers of one of match.cu's stages and stay valid until stage_release.
struct MatchParams;
struct DeviceCodes {
  std::vector<long long> code_off;       // per gene: offset of its R x K block (read-major, K contiguous)
  std::vector<long long> pair_off;       // per gene: offset of its R order entries
  std::vector<int> status;               // per gene: 0 or MISOB200_EINVAL (unparsable CIGAR)
  const uint16_t *codes16 = nullptr;     // one of the two: 8-bit codes when the plan has at most 256 codes
  const uint8_t *codes8 = nullptr;
  const int32_t *order = nullptr;        // draw order (order[pair_off[g]] == -1: this gene is sorted on the host); may be null
  int stage = -1;                        // match.cu stage holding the buffers
  long long n_pairs = 0;
  double kernel_ms = 0, sort_ms = 0, h2d_ms = 0, d2h_ms = 0;
  long long bytes_in = 0, bytes_out = 0;
};
int stage_submit(const misob200_reads_t &reads, const MatchParams &mp, int device, int n_codes,
                 const std::vector<uint16_t> &code_rank, DeviceCodes &out);
int stage_wait(DeviceCodes &out);
void stage_release(int stage);
void stage_pool_release();
// timing of the last device matching of this thread's plan_append (ms; bytes)
void last_match_stats(double *kernel_ms, double *h2d_ms, double *d2h_ms, long long *bytes_in, long long *bytes_out);

void set_error(const std::string &msg);
// match_device < 0: compatibility on the host; >= 0: on that GPU
int plan_append(Plan &plan, const misob200_reads_t &reads, int n_threads, int match_device = -1);
// the same in two calls, so that a pipeline can have several batches in flight: _begin validates,
// fixes the plan's library and ENQUEUES the device work of the batch (copies in, match_kernel,
// order_kernel, copies out); _finish waits for it and runs the host half (classes, tiles).
// `reads` must stay alive and unchanged between the two calls.
struct PendingAppend;
int plan_append_device_begin(Plan &plan, const misob200_reads_t &reads, int device, PendingAppend **pending);
int plan_append_device_finish(Plan &plan, PendingAppend *pending, int n_threads);

// Host worker threads for the plan stage and the output epilogues: MISOB200_HOST_THREADS, else
// the cores this process may use (affinity mask, cgroup cpu.max quota) divided by the ranks
// sharing the node (LOCAL_WORLD_SIZE, set by torchrun), at most 32.
int host_threads();

}  // namespace misob200

struct misob200_plan { misob200::Plan p; };
