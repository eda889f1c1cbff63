/*
 * include/miso_synth.h -- C ABI of workloads/libmiso_synth.so
 *
 * Synthetic inputs for tests and bench.py (BASELINE.json configs 2, 3, 4, 5).
 * Bench / test infrastructure: NOT part of the product library.  The read
 * model is the one the reference's simulators implement
 * (/root/reference/pysplicing/src/simulator.c:68-196, :221-442); the output
 * is a misob200_reads_t view, i.e. exactly what MISO / MISOPaired receive
 * (/root/reference/pysplicing/src/pysplicing.c:62,173).
 */
#ifndef MISO_SYNTH_H
#define MISO_SYNTH_H

#include "miso_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic workloads (BASELINE.json configs 2 and 3) -------------- */
typedef struct misob200_workload misob200_workload_t;	/* opaque */
/* kind 0: K=2 skipped-exon SE events (cfg-2); kind 1: K in [2,8] paired-end
   events with a N(frag_mean, frag_var) insert model (cfg-3).  Genes get ids
   first_gene_id .. first_gene_id+n_genes-1 and are generated from (seed, id)
   only, so any shard of a workload can be rebuilt independently. */
int misob200_workload_create(int kind, int32_t n_genes, int32_t reads_per_gene,
			     int32_t read_len, double frag_mean,
			     double frag_var, double num_devs, uint64_t seed,
			     uint32_t first_gene_id, int n_threads,
			     misob200_workload_t **out);
/* the same for an explicit list of gene ids (a rank's shard of a workload).
   sample > 0: another sample of the same events -- identical gene structures,
   its own psi and reads (the two conditions of cfg-5) */
int misob200_workload_create_ids(int kind, int32_t n_genes,
				 const uint32_t *gene_ids,
				 int32_t reads_per_gene, int32_t read_len,
				 double frag_mean, double frag_var,
				 double num_devs, uint64_t seed, uint32_t sample,
				 int n_threads, misob200_workload_t **out);
/* isoform count per gene (a workload created with reads_per_gene = 0 holds the
   gene structures only: cheap way to cost a workload before dealing it) */
int misob200_workload_n_iso(const misob200_workload_t *w, int32_t *n_iso);
const char *misob200_workload_last_error(void);
int misob200_workload_view(const misob200_workload_t *w,
			   misob200_reads_t *view);
int misob200_workload_truth(const misob200_workload_t *w, int32_t gene,
			    double *psi);
int misob200_workload_destroy(misob200_workload_t *w);

#ifdef __cplusplus
}
#endif
#endif
