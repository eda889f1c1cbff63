/*
 * include/miso_b200.h -- C ABI of libmiso_b200.so
 *
 * B200-native drop-in for the reference's per-gene MCMC PSI sampler, i.e. the
 * C entry points that pysplicing's extension module binds
 * (/root/reference/pysplicing/src/pysplicing.c:41-131 `MISO`, :152-244
 * `MISOPaired`, :246-278 `createGene`) and that are declared in
 * /root/reference/pysplicing/include/splicing.h:203-238
 * (`splicing_miso`, `splicing_miso_paired`), re-cut for a GPU: many genes per
 * call, flat pointer + size arguments, caller-owned output buffers.
 *
 * Plain C only: no C++ and no framework types cross this boundary.  Every
 * function returns 0 on success or one of the MISOB200_E* codes (the values
 * follow /root/reference/pysplicing/include/splicing_error.h:316-366 where a
 * counterpart exists); misob200_last_error() gives the message.
 *
 * Pipeline
 *   misob200_plan_create        empty plan (opaque, host side)
 *   misob200_plan_append        reads + gene structures  ->  compatibility
 *                               codes, draw order, read classes, packed tiles
 *                               (replaces splicing_matchIso[_paired],
 *                               splicing_order_matches, splicing_i_miso_classes
 *                               and the effective-length block:
 *                               src/solve.c:8-218, src/miso.c:758-784,
 *                               src/miso_paired.c:378-419)
 *   misob200_run                H2D tiles, sm_100a chain kernel (one warp per
 *                               gene-chain, whole burn-in + sampling loop on
 *                               chip), D2H posteriors (replaces the loop of
 *                               src/miso.c:827-947 / src/miso_paired.c:431-538)
 *   misob200_summarize          posterior mean + 95% CI per gene on the device
 *                               (misopy/credible_intervals.py:31-55)
 */
#ifndef MISO_B200_H
#define MISO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (splicing_error.h values where they exist) ---------- */
#define MISOB200_SUCCESS        0
#define MISOB200_FAILURE        1
#define MISOB200_ENOMEM         2
#define MISOB200_EINVAL         4
#define MISOB200_UNIMPLEMENTED 12
#define MISOB200_ECUDA        100	/* CUDA runtime / no device */
#define MISOB200_ENCCL        101

/* ---- enums: pysplicing/pysplicing/__init__.py:2-13 ------------------- */
#define MISOB200_START_AUTO      0
#define MISOB200_START_UNIFORM   1
#define MISOB200_START_RANDOM    2	/* Dirichlet(1) start, src/miso.c:388-404 */
#define MISOB200_STOP_FIXEDNO    0
#define MISOB200_ALGO_REASSIGN   0

#define MISOB200_MAX_ISO         8	/* isoforms per gene handled on chip */

/* ---- input: a batch of genes with their reads ------------------------- */
/*
 * Gene structure is what createGene receives (pysplicing.c:246-278) after
 * splicing_create_gene / splicing_gff_exon_start_end have expanded it
 * (src/simulator.c:9-66, src/gff.c:728-779): for every isoform the list of
 * its exons as 1-based inclusive [start,end], in isoform order.
 * Reads are what MISO / MISOPaired receive: 1-based start positions and
 * CIGAR strings (pysplicing.c:62,173); for paired-end the two mates of a
 * pair are consecutive (src/solve.c:139).
 */
typedef struct misob200_reads {
  int32_t n_genes;
  const int32_t *iso_off;	/* [n_genes+1]  first isoform of each gene      */
  const int32_t *exon_off;	/* [n_iso+1]    first exon of each isoform       */
  const int32_t *exon_start;	/* [n_exon]                                      */
  const int32_t *exon_end;	/* [n_exon]                                      */
  const int64_t *read_off;	/* [n_genes+1]  first read (SE) / mate (PE)      */
  const int32_t *position;	/* [n_reads]    1-based                          */
  const int64_t *cigar_off;	/* [n_reads+1]  byte offsets into cigar          */
  const char *cigar;		/* NUL-terminated strings, back to back          */
  const double *hyper;		/* [n_iso] Dirichlet prior, or NULL for all 1.0  */
  const uint32_t *gene_id;	/* [n_genes] RNG stream ids, or NULL for 0..n-1  */
  int32_t read_len;
  int32_t overhang;		/* 0 means 1, as src/miso.c:690                  */
  int32_t paired;		/* 0: MISO, 1: MISOPaired                        */
  double frag_mean, frag_var, num_devs;	/* paired only (pysplicing.c:173) */
} misob200_reads_t;

typedef struct misob200_plan misob200_plan_t;	/* opaque */

/* ---- run parameters: positional tail of MISO(...) --------------------- */
typedef struct misob200_params {
  int32_t n_iters;		/* noIterations, default 5000 */
  int32_t burn_in;		/* noBurnIn,     default 500  */
  int32_t lag;			/* noLag,        default 10   */
  int32_t n_chains;		/* no_chains,    default 6    */
  int32_t start;		/* MISOB200_START_*           */
  int32_t stop;			/* MISOB200_STOP_FIXEDNO only */
  int32_t algo;			/* MISOB200_ALGO_REASSIGN only */
  int32_t device;		/* CUDA ordinal               */
  uint64_t seed;		/* stream = Philox4x32 keyed (seed; gene_id, chain) */
} misob200_params_t;

/* per-gene layout of the outputs of misob200_run (all caller-owned):
 *   samples   : for gene g, K_g x (n_chains*S) doubles, column-major, column
 *               s*n_chains + c = sample s of chain c (src/miso.c:884-888),
 *               S = (n_iters-burn_in)/lag; the gene blocks lie back to back
 *               in the device's run order, locate gene g with
 *               misob200_plan_offsets(plan, params, g, ...)
 *   loglik    : n_chains*S doubles per gene, same column order
 *   assignment: one int32 per read (SE) / pair (PE) in input order, chain 0,
 *               -1 = incompatible (src/miso.c:943-946)
 *   rundata   : 9 int32 per gene, the fields of splicing_miso_rundata_t
 *               (include/splicing.h:143-146) in declaration order.  noSamples
 *               (field 8) = n_chains*S, the columns actually recorded and the
 *               size of the gene's block; the reference reports
 *               n_chains*(n_iters-burn_in)/lag (src/miso.c:661), which differs
 *               when lag does not divide n_iters-burn_in -- its matrix then
 *               ends in that many all-zero columns, which are not reproduced
 *   status    : 1 int32 per gene, 0 ok, else MISOB200_E*
 */

int misob200_version(void);
const char *misob200_last_error(void);

/* device bring-up; fails loudly (MISOB200_ECUDA) if no sm_100 device */
int misob200_init(int device);
int misob200_shutdown(void);	/* frees the pooled device states and setup stages */
int misob200_device_count(int *count);

int misob200_plan_create(misob200_plan_t **plan);
int misob200_plan_destroy(misob200_plan_t *plan);
int misob200_plan_append(misob200_plan_t *plan, const misob200_reads_t *reads,
			 int n_threads);
/* the same, with the read <-> isoform compatibility (splicing_matchIso[_paired]
   + splicing_parse_cigar, src/solve.c:8-306) and the draw order
   (splicing_order_matches, src/miso.c:988-993) computed by kernels on `device`
   instead of the host threads; identical plan, MISOB200_ECUDA without a GPU */
int misob200_plan_append_device(misob200_plan_t *plan,
				const misob200_reads_t *reads, int n_threads,
				int device);
/* the same in two calls, for pipelines with several batches in flight
   (miso_b200/pipeline.py): _begin validates, fixes the plan's library and
   ENQUEUES the batch's device work (copies in, match_kernel, order_kernel =
   the draw-order sort of splicing_order_matches, src/miso.c:988-993, copies
   out) on one of three internal stages; _finish waits for it and runs the host
   half (read classes, tile packing).  `reads` and its arrays must stay alive
   and unchanged in between.  Every _begin that returned 0 must be _finish-ed. */
int misob200_plan_append_device_begin(misob200_plan_t *plan,
				      const misob200_reads_t *reads,
				      int device, void **pending);
int misob200_plan_append_device_finish(misob200_plan_t *plan, void *pending,
				       int n_threads);
/* timing / traffic of the calling thread's last device matching:
   kernel, H2D, D2H in ms; input and output bytes */
int misob200_last_match_stats(double *kernel_ms, double *h2d_ms, double *d2h_ms,
			      int64_t *bytes_in, int64_t *bytes_out);

/* keep the code matrix and draw order of later appends for
   misob200_plan_gene_match (parity tests; costs 4 bytes per read x isoform) */
int misob200_plan_keep_match(misob200_plan_t *plan, int on);

/* tile format of later appends: -1 (default) class tiles whenever a gene has at
   most 254 weight classes among its drawing reads, dense tiles otherwise;
   0 dense tiles only (the general fp64 pass; parity tests run both) */
int misob200_plan_tile_format(misob200_plan_t *plan, int format);
/* per gene: tile format chosen (0 dense, 1 class), number of weight classes
   (class format), tile bytes */
int misob200_plan_gene_tile(const misob200_plan_t *plan, int32_t gene,
			    int32_t *format, int32_t *n_weight_classes,
			    int32_t *tile_bytes);

int misob200_plan_size(const misob200_plan_t *plan, int32_t *n_genes,
		       int64_t *n_reads, int64_t *tile_bytes);
/* per gene: K, number of reads/pairs R, reads that draw each pass R2,
   number of read classes, per-gene status from the setup stage */
int misob200_plan_gene_info(const misob200_plan_t *plan, int32_t gene,
			    int32_t *n_iso, int32_t *n_reads, int32_t *n_drawn,
			    int32_t *n_classes, int32_t *status);
/* the same for every gene of the plan: info5 = n_genes x {n_iso, n_reads,
   n_drawn, n_classes, status} */
int misob200_plan_info_all(const misob200_plan_t *plan, int32_t *info5);
/* class_templates: n_classes x K doubles, row per class (this is the
   transposed form pysplicing returns, pysplicing.c:120-121); counts likewise.
   SE: exact columns (miso_paired.c:576-619); PE: zero/non-zero patterns
   (miso_paired.c:628-681) */
int misob200_plan_gene_classes(const misob200_plan_t *plan, int32_t gene,
			       double *class_templates, double *class_counts);
/* the setup-stage products, for parity tests: code matrix K x R column-major
   (0 = incompatible, SE 1, PE fragment - fragment_start + 1) and draw order */
int misob200_plan_gene_match(const misob200_plan_t *plan, int32_t gene,
			     int32_t *codes, int32_t *order);
int misob200_plan_fragment_table(const misob200_plan_t *plan, int32_t cap,
				 double *prob, int32_t *frag_start,
				 int32_t *n_len);
/* offsets (in elements) of gene g inside samples / loglik / assignment.
   The per-gene blocks of samples and loglik are NOT in gene order: they follow
   the order in which the device runs the genes (isoform-count buckets), so
   that finished buckets copy out while others still run; always go through
   this call.  assignment is in input read order. */
int misob200_plan_offsets(const misob200_plan_t *plan,
			  const misob200_params_t *params, int32_t gene,
			  int64_t *sample_off, int64_t *loglik_off,
			  int64_t *assign_off);
int misob200_plan_offsets_all(const misob200_plan_t *plan,
			      const misob200_params_t *params,
			      int64_t *sample_off, int64_t *loglik_off,
			      int64_t *assign_off);
int misob200_plan_output_sizes(const misob200_plan_t *plan,
			       const misob200_params_t *params,
			       int64_t *n_samples_f64, int64_t *n_loglik_f64,
			       int64_t *n_assign_i32);

/* timing_ms (may be NULL): [0] H2D, [1] kernels, [2] D2H, [3] total, from
   CUDA events on the run stream; launches = kernels launched (may be NULL) */
int misob200_run(misob200_plan_t *plan, const misob200_params_t *params,
		 double *samples, double *loglik, int32_t *assignment,
		 int32_t *rundata, int32_t *status, double *timing_ms,
		 int32_t *launches);

/* Device-resident variant for measurement: upload once, run many times,
   download once.  misob200_run == upload + run_resident + download. */
int misob200_upload(misob200_plan_t *plan, const misob200_params_t *params);
int misob200_run_resident(misob200_plan_t *plan, double *kernel_ms,
			  int32_t *launches);
int misob200_download(misob200_plan_t *plan, double *samples, double *loglik,
		      int32_t *assignment, int32_t *rundata, int32_t *status);
int misob200_release_device(misob200_plan_t *plan);

/* measurement helpers: kernel time of each K bucket of the last resident run
   (ms9[K], K = 2..8, CUDA events on that bucket's stream) and the bytes one
   misob200_run moves over PCIe in each direction */
int misob200_bucket_timing(misob200_plan_t *plan, double *ms9);
int misob200_transfer_bytes(misob200_plan_t *plan, int64_t *h2d, int64_t *d2h);

/* Posterior summaries on the device (needs a resident run): per gene a
   256-byte record {mean[8], ci_low[8], ci_high[8] (f64); assigned_counts[8],
   n_iso, accepted, rejected, status (i32); pad}.  summary: n_genes*32 f64. */
#define MISOB200_SUMMARY_F64 32
int misob200_summarize(misob200_plan_t *plan, double *summary);

/* Two-sample comparison on the device (misopy/hypothesis_test.py:89-179,348-380;
   compare_miso --compare-samples): both plans hold the same events in the same
   order, have run with the same iterations/burn-in/lag/chains and are resident
   on the same GPU.  out: n_genes records of 32 f64 = bayes_factor[8],
   mean1-mean2 [8], mean|delta| [8], KDE(0) [8]. */
#define MISOB200_COMPARE_F64 32
int misob200_compare(misob200_plan_t *plan_a, misob200_plan_t *plan_b, double *out);

/* ---- multi-GPU: genes are sharded by the caller, one process per GPU;
   the only exchange is one all-gather of the summary records ------------- */
int misob200_comm_unique_id(char *id128);		/* rank 0 */
int misob200_comm_init(const char *id128, int n_ranks, int rank);
/* The path's one collective, device to device: summary kernel on this rank's
   resident posteriors -> ncclAllGather of the 256-byte records over NVLink ->
   one device->host copy of the gathered table.  Every rank passes the same
   n_pad >= its own gene count (all-gather needs equal counts; pad records
   carry status = -1).  all: n_ranks * n_pad * 32 f64, rank-major. */
int misob200_comm_allgather_summaries(misob200_plan_t *plan, int64_t n_pad,
				      double *all);
/* the same for the two-sample records of misob200_compare (cfg-5) */
int misob200_comm_allgather_compare(misob200_plan_t *plan_a,
				    misob200_plan_t *plan_b, int64_t n_pad,
				    double *all);
/* generic host-buffer form (tests, small payloads) */
int misob200_comm_allgather(const double *mine, int64_t n_f64_per_rank,
			    double *all);
int misob200_comm_barrier_max(double *value);	/* in: local, out: max */
int misob200_comm_destroy(void);

/* Batched `.miso` writer (misopy/miso_sampler.py:456-465, one file per event):
   file i = headers[i] (the "#isoforms=...\n" line, built by the caller) +
   "sampled_psi\tlog_score\n" + n_rows lines "%.4f,...,%.4f\t%.2f\n" from the
   event's block of the output buffers of misob200_run (offsets in f64 elements,
   rows of n_iso[i] psi values).  Formatted and written by n_threads host
   threads (0: misob200_host_threads()); the directories must exist. */
int misob200_write_miso_files(int32_t n_files, const char *const *paths,
			      const char *const *headers,
			      const double *samples, const int64_t *sample_off,
			      const double *loglik, const int64_t *loglik_off,
			      const int32_t *n_iso, int32_t n_rows,
			      int n_threads, int64_t *bytes_written);
/* The whole plan at once, headers included: file g = prefix[g] + the
   run-dependent header fields of misopy/miso_sampler.py:444-454 (iters,
   burn_in, lag, percent_accept, proposal_type, counts = read classes of the
   setup stage, assigned_counts = chain 0's assignments per isoform) +
   suffix[g] + the body as above, from the buffers misob200_run filled.
   prefix ("#isoforms=[..]\texon_lens=..\t") and suffix ("\tchrom=..\tstrand=..
   \tmRNA_starts=..\tmRNA_ends=..\n") depend on the annotation only.  Skipped:
   genes with paths[g] == NULL, status != 0, or no compatible read
   (miso_sampler.py:352-354). */
int misob200_plan_write_miso(const misob200_plan_t *plan,
			     const misob200_params_t *params,
			     const char *const *paths,
			     const char *const *prefix,
			     const char *const *suffix, const double *samples,
			     const double *loglik, const int32_t *assignment,
			     const int32_t *rundata, int n_threads,
			     int64_t *n_written, int64_t *bytes_written);
/* the writer's number formatting, "%.2f" / "%.4f" (decimals = 2 | 4): exact
   round-half-even on the binary value, as CPython's "%" operator; returns the
   length written to out32 (NUL-terminated, at most 31 characters) */
int misob200_format_fixed(double v, int decimals, char *out32);

/* The random stream ("same seed" is defined by this framework, the reference
   never seeds): Philox4x32 keyed (seed; gene_id, chain), see
   miso_b200/csrc/philox.cuh and oracle/philox_ref.h.  Version 2 (default) runs
   the 7 rounds that pass BigCrush, version 1 the 10 rounds of the first
   release.  set = 1 | 2 selects (process-wide, before misob200_run), 0 only
   queries; the environment variable MISOB200_STREAM=1 selects v1 at start-up.
   Returns the version in force. */
int misob200_stream_version(int set);

/* host worker threads the library uses for the plan stage and the output
   epilogue: MISOB200_HOST_THREADS, else usable cores (affinity, cgroup quota)
   divided by LOCAL_WORLD_SIZE */
int misob200_host_threads(void);

/* page-locked host buffers for the outputs of misob200_run / _download
   (optional: any host pointer works, pinned ones copy at link speed) */
void *misob200_host_alloc(int64_t bytes);
int misob200_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
