/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin plain-C driver around the UNMODIFIED reference C core
 * (/root/reference/pysplicing/src/{miso,miso_paired,solve,gff,simulator,
 * random,vector,matrix,strvector,memory,error,util,qsort,qsort_r}.c, compiled
 * where they lie by oracle/Makefile into oracle/_ref/libsplicing_ref.so).
 * Nothing of the reference is copied here: this file only
 *   (a) installs a deterministic random stream through the reference's own
 *       RNG vtable (splicing_rng_type_t, include/splicing_random.h:23-36,
 *       splicing_rng_set_default, src/random.c:467-469) -- the stream is the
 *       one specified in oracle/philox_ref.h;
 *   (b) exposes splicing_miso / splicing_miso_paired and the setup functions
 *       they call with flat pointer arguments so tests can drive them through
 *       ctypes;
 * The four off-path symbols the linker still wants (NNLS / assignment
 * matrices, reached only by START_LINEAR / ALGO_CLASSES) are stubbed in
 * oracle/ref_stubs.c so that the vendored f2c LAPACK need not be built.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>

#include "splicing.h"
#include "splicing_error.h"
#include "splicing_random.h"

#include "philox_ref.h"

/* ------------------------------------------------------------------ */
/* RNG vtable                                                          */

static phx_stream_t refh_stream;

static double refh_get_real(void *state) {
  return phx_next_uniform((phx_stream_t *) state);
}
static double refh_get_norm(void *state) {
  return phx_next_normal((phx_stream_t *) state);
}

/* gamma(1, scale) only: the START_RANDOM Dirichlet draw (src/miso.c:318 with
   alpha = 1, :394).  Any other shape is not part of the stream definition. */
static double refh_get_gamma(void *state, double a, double scale) {
  if (a != 1.0) {
    fprintf(stderr, "ref_harness: gamma shape %g is not defined by the stream\n", a);
    abort();
  }
  return phx_next_gamma1((phx_stream_t *) state, scale);
}

static splicing_rng_type_t refh_rngtype = {
  /* name= */      "miso-b200 stream v1 (Philox4x32-10)",
  /* min= */       0,
  /* max= */       0xffffffffUL,
  /* init= */      0,
  /* destroy= */   0,
  /* seed= */      0,
  /* get= */       0,
  /* get_real= */  refh_get_real,
  /* get_norm= */  refh_get_norm,
  /* get_geom= */  0,
  /* get_binom= */ 0,
  /* get_gamma= */ refh_get_gamma
};

static splicing_rng_t refh_saved_default;
static int refh_have_saved = 0;

/* stream version: 1 = Philox4x32-10, 2 = Philox4x32-7 (philox_ref.h) */
int refh_set_stream(int version) {
  if (version == 1) phx_rounds = 10;
  if (version == 2) phx_rounds = 7;
  return phx_rounds == 10 ? 1 : 2;
}

/* rng_mode 0: Philox stream keyed (seed, gene, chain).
   rng_mode 1: the reference's own MT19937 default, seeded with `seed'
               (its fastest configuration; used for CPU-baseline timing). */
static void refh_install_rng(int rng_mode, uint64_t seed, uint32_t gene,
			     uint32_t chain) {
  if (!refh_have_saved) {
    refh_saved_default = splicing_rng_default;
    refh_have_saved = 1;
  }
  if (rng_mode == 0) {
    splicing_rng_t r;
    refh_stream.seed = seed; refh_stream.gene = gene;
    refh_stream.chain = chain;
    refh_stream.n_unif = refh_stream.n_norm = 0;
    r.type = &refh_rngtype; r.state = &refh_stream; r.def = 0;
    splicing_rng_set_default(&r);
  } else {
    splicing_rng_set_default(&refh_saved_default);
    splicing_rng_default.def = 2;
    splicing_rng_seed(&splicing_rng_default, (unsigned long) seed);
  }
}

void refh_rng_counts(uint64_t *n_unif, uint64_t *n_norm) {
  *n_unif = refh_stream.n_unif; *n_norm = refh_stream.n_norm;
}

/* START_RANDOM leaves `sigma' of splicing_miso[_paired] unassigned
   (src/miso.c:655 vs :388-404; src/miso_paired.c:264): the unmodified
   reference then reads whatever its stack slot holds.  To compare it with the
   port on that branch the harness fills the stack region the callee's frame
   is about to occupy with the bit pattern of SIGMA = 0.2/K/K right before the
   call.  This relies on the compiler giving the address-taken local its own
   slot and is checked, not assumed: tests/test_oracle_vs_reference.py only
   accepts the comparison when the outputs agree bit for bit.            */
static void __attribute__((noinline)) refh_paint_stack(double v) {
  volatile double pad[4096];
  int i;
  for (i = 0; i < 4096; i++) { pad[i] = v; }
}

/* splicing_miso prints "no chains: N" on every call (src/miso.c:837).  Keep
   test logs and timings clean by pointing fd 1 at /dev/null meanwhile. */
static int refh_quiet_fd = -1;
static void refh_quiet_begin(void) {
  int nul;
  fflush(stdout);
  refh_quiet_fd = dup(1);
  nul = open("/dev/null", O_WRONLY);
  if (nul >= 0) { dup2(nul, 1); close(nul); }
}
static void refh_quiet_end(void) {
  fflush(stdout);
  if (refh_quiet_fd >= 0) { dup2(refh_quiet_fd, 1); close(refh_quiet_fd); }
  refh_quiet_fd = -1;
}

/* ------------------------------------------------------------------ */
/* gene construction                                                   */

static int refh_make_gene(splicing_gff_t *gff, int nexons, const int *exons,
			  int isolen, const int *isoforms) {
  splicing_vector_int_t ex, iso;
  int i, ret;
  splicing_vector_int_init(&ex, 2 * nexons);
  splicing_vector_int_init(&iso, isolen);
  for (i = 0; i < 2 * nexons; i++) { VECTOR(ex)[i] = exons[i]; }
  for (i = 0; i < isolen; i++) { VECTOR(iso)[i] = isoforms[i]; }
  splicing_gff_init(gff, 0);
  ret = splicing_create_gene(&ex, &iso, "insilicogene", "seq1",
			     "protein_coding", SPLICING_STRAND_UNKNOWN, gff);
  splicing_vector_int_destroy(&iso);
  splicing_vector_int_destroy(&ex);
  return ret;
}

static void refh_copy_matrix(const splicing_matrix_t *m, double *out) {
  if (out) {
    memcpy(out, &MATRIX(*m, 0, 0),
	   sizeof(double) * splicing_matrix_nrow(m) * splicing_matrix_ncol(m));
  }
}

/* Errors come back as return codes instead of abort(): install the
   reference's own "free the finally-stack and return" handler
   (include/splicing_error.h:208-215). */
int refh_init(void) {
  splicing_set_error_handler(splicing_error_handler_ignore);
  return 0;
}

/* ------------------------------------------------------------------ */
/* setup-stage probes                                                   */

int refh_gene_info(int nexons, const int *exons, int isolen,
		   const int *isoforms, int *noiso, int *isolength,
		   int *noexons) {
  splicing_gff_t gff;
  splicing_vector_int_t il, ne;
  size_t n, i;
  int ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_vector_int_init(&il, 0);
  splicing_vector_int_init(&ne, 0);
  splicing_gff_noiso_one(&gff, 0, &n);
  splicing_gff_isolength_one(&gff, 0, &il);
  splicing_gff_noexons_one(&gff, 0, &ne);
  *noiso = (int) n;
  for (i = 0; i < n; i++) {
    isolength[i] = VECTOR(il)[i]; noexons[i] = VECTOR(ne)[i];
  }
  splicing_vector_int_destroy(&ne);
  splicing_vector_int_destroy(&il);
  splicing_gff_destroy(&gff);
  return 0;
}

/* match (K x R, column-major double), order (R), class templates (K x ncls)
   + counts.  src/solve.c:8-108, src/miso.c:988-993, src/miso_paired.c:576-619 */
int refh_match_se(int nexons, const int *exons, int isolen,
		  const int *isoforms, int nreads, const int *pos,
		  const char **cigars, int readLength, int overhang,
		  double *match, int *order, int *ncls,
		  double *class_templates, double *class_counts) {
  splicing_gff_t gff;
  splicing_vector_int_t position, ord;
  splicing_matrix_t m, ct;
  splicing_vector_t cc;
  int i, ret;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_vector_int_init(&position, nreads);
  for (i = 0; i < nreads; i++) { VECTOR(position)[i] = pos[i]; }
  splicing_matrix_init(&m, 0, 0);
  splicing_matrix_init(&ct, 0, 0);
  splicing_vector_init(&cc, 0);
  splicing_vector_int_init(&ord, 0);
  ret = splicing_matchIso(&gff, 0, &position, cigars, overhang, readLength,
			  &m);
  if (!ret) { ret = splicing_order_matches(&m, &ord); }
  if (!ret) { ret = splicing_i_miso_classes(&m, &ord, &ct, &cc, 0, 0); }
  if (!ret) {
    refh_copy_matrix(&m, match);
    if (order) {
      for (i = 0; i < nreads; i++) { order[i] = VECTOR(ord)[i]; }
    }
    *ncls = splicing_matrix_ncol(&ct);
    refh_copy_matrix(&ct, class_templates);
    if (class_counts) {
      memcpy(class_counts, VECTOR(cc), sizeof(double) * (*ncls));
    }
  }
  splicing_vector_int_destroy(&ord);
  splicing_vector_destroy(&cc);
  splicing_matrix_destroy(&ct);
  splicing_matrix_destroy(&m);
  splicing_vector_int_destroy(&position);
  splicing_gff_destroy(&gff);
  return ret;
}

/* Insert-length table as splicing_miso_paired builds it when called with
   fragmentProb == NULL (src/miso_paired.c:299-308; src/simulator.c:198-219). */
int refh_fragment_table(double mean, double var, double numDevs,
			int readLength, int cap, double *prob, int *start,
			int *il) {
  splicing_vector_t fp;
  int ret, n;
  splicing_vector_init(&fp, 0);
  ret = splicing_normal_fragment(mean, var, numDevs, readLength, &fp, start);
  if (!ret) {
    splicing_vector_scale(&fp, 1.0 / splicing_vector_sum(&fp));
    n = splicing_vector_size(&fp);
    *il = n;
    if (n > cap) { ret = -1; } else {
      memcpy(prob, VECTOR(fp), sizeof(double) * n);
    }
  }
  splicing_vector_destroy(&fp);
  return ret;
}

/* nreads = number of single reads (2 per pair).  src/solve.c:141-218 */
int refh_match_pe(int nexons, const int *exons, int isolen,
		  const int *isoforms, int nreads, const int *pos,
		  const char **cigars, int readLength, int overhang,
		  double mean, double var, double numDevs,
		  double *match, int *fraglen, int *order, int *ncls,
		  double *bin_class_templates, double *bin_class_counts) {
  splicing_gff_t gff;
  splicing_vector_int_t position, ord;
  splicing_matrix_t m, ct;
  splicing_matrix_int_t fl;
  splicing_vector_t cc;
  int i, ret, npairs = nreads / 2;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_vector_int_init(&position, nreads);
  for (i = 0; i < nreads; i++) { VECTOR(position)[i] = pos[i]; }
  splicing_matrix_init(&m, 0, 0);
  splicing_matrix_int_init(&fl, 0, 0);
  splicing_matrix_init(&ct, 0, 0);
  splicing_vector_init(&cc, 0);
  splicing_vector_int_init(&ord, 0);
  ret = splicing_matchIso_paired(&gff, 0, &position, cigars, readLength,
				 overhang, /*fragmentProb=*/ 0,
				 /*fragmentStart=*/ 0, mean, var, numDevs,
				 &m, &fl);
  if (!ret) { ret = splicing_order_matches(&m, &ord); }
  if (!ret) { ret = splicing_i_miso_classes(&m, &ord, 0, 0, &ct, &cc); }
  if (!ret) {
    refh_copy_matrix(&m, match);
    if (fraglen) {
      memcpy(fraglen, &MATRIX(fl, 0, 0),
	     sizeof(int) * splicing_matrix_int_nrow(&fl) *
	     splicing_matrix_int_ncol(&fl));
    }
    if (order) {
      for (i = 0; i < npairs; i++) { order[i] = VECTOR(ord)[i]; }
    }
    *ncls = splicing_matrix_ncol(&ct);
    refh_copy_matrix(&ct, bin_class_templates);
    if (bin_class_counts) {
      memcpy(bin_class_counts, VECTOR(cc), sizeof(double) * (*ncls));
    }
  }
  splicing_vector_int_destroy(&ord);
  splicing_vector_destroy(&cc);
  splicing_matrix_destroy(&ct);
  splicing_matrix_int_destroy(&fl);
  splicing_matrix_destroy(&m);
  splicing_vector_int_destroy(&position);
  splicing_gff_destroy(&gff);
  return ret;
}

/* ------------------------------------------------------------------ */
/* the sampler itself                                                   */

/* samples: K x noSamples column-major, noSamples = C*(I-B)/L (src/miso.c:661)
   rundata: the 9 ints of splicing_miso_rundata_t in declaration order.   */
int refh_miso_se(int nexons, const int *exons, int isolen,
		 const int *isoforms, int nreads, const int *pos,
		 const char **cigars, int readLength, int overhang,
		 int noChains, int noIterations, int noBurnIn, int noLag,
		 const double *hyper, int start, int stop,
		 int rng_mode, uint64_t seed, uint32_t gene_id,
		 uint32_t chain_id,
		 double *samples, double *logLik, int *ncls,
		 double *class_templates, double *class_counts,
		 int *assignment, int *rundata) {
  splicing_gff_t gff;
  splicing_vector_int_t position, ass;
  splicing_vector_t hyp, ll, cc;
  splicing_matrix_t smp, ct;
  splicing_miso_rundata_t rd;
  size_t noiso;
  int i, ret;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_gff_noiso_one(&gff, 0, &noiso);
  splicing_vector_int_init(&position, nreads);
  for (i = 0; i < nreads; i++) { VECTOR(position)[i] = pos[i]; }
  splicing_vector_init(&hyp, noiso);
  for (i = 0; i < (int) noiso; i++) { VECTOR(hyp)[i] = hyper[i]; }
  splicing_matrix_init(&smp, 0, 0);
  splicing_vector_init(&ll, 0);
  splicing_matrix_init(&ct, 0, 0);
  splicing_vector_init(&cc, 0);
  splicing_vector_int_init(&ass, 0);
  memset(&rd, 0, sizeof(rd));

  refh_install_rng(rng_mode, seed, gene_id, chain_id);
  refh_quiet_begin();
  if (start == 2) { refh_paint_stack(0.2 / (double) noiso / (double) noiso); }
  ret = splicing_miso(&gff, 0, &position, cigars, readLength, overhang,
		      noChains, noIterations, /*maxIterations=*/ 100000,
		      noBurnIn, noLag, &hyp, SPLICING_ALGO_REASSIGN,
		      (splicing_miso_start_t) start,
		      (splicing_miso_stop_t) stop, /*start_psi=*/ 0,
		      &smp, &ll, /*match_matrix=*/ 0, &ct, &cc, &ass, &rd);
  refh_quiet_end();

  if (!ret) {
    refh_copy_matrix(&smp, samples);
    if (logLik) {
      memcpy(logLik, VECTOR(ll), sizeof(double) * splicing_vector_size(&ll));
    }
    *ncls = splicing_matrix_ncol(&ct);
    refh_copy_matrix(&ct, class_templates);
    if (class_counts) {
      memcpy(class_counts, VECTOR(cc), sizeof(double) * (*ncls));
    }
    if (assignment) {
      memcpy(assignment, VECTOR(ass), sizeof(int) * nreads);
    }
    memcpy(rundata, &rd, sizeof(rd));
  }
  splicing_vector_int_destroy(&ass);
  splicing_vector_destroy(&cc);
  splicing_matrix_destroy(&ct);
  splicing_vector_destroy(&ll);
  splicing_matrix_destroy(&smp);
  splicing_vector_destroy(&hyp);
  splicing_vector_int_destroy(&position);
  splicing_gff_destroy(&gff);
  return ret;
}

int refh_miso_pe(int nexons, const int *exons, int isolen,
		 const int *isoforms, int nreads, const int *pos,
		 const char **cigars, int readLength, int overhang,
		 double mean, double var, double numDevs,
		 int noChains, int noIterations, int noBurnIn, int noLag,
		 const double *hyper, int start, int stop,
		 int rng_mode, uint64_t seed, uint32_t gene_id,
		 uint32_t chain_id,
		 double *samples, double *logLik, int *ncls,
		 double *bin_class_templates, double *bin_class_counts,
		 int *assignment, int *rundata) {
  splicing_gff_t gff;
  splicing_vector_int_t position, ass;
  splicing_vector_t hyp, ll, cc;
  splicing_matrix_t smp, ct;
  splicing_miso_rundata_t rd;
  size_t noiso;
  int i, ret;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_gff_noiso_one(&gff, 0, &noiso);
  splicing_vector_int_init(&position, nreads);
  for (i = 0; i < nreads; i++) { VECTOR(position)[i] = pos[i]; }
  splicing_vector_init(&hyp, noiso);
  for (i = 0; i < (int) noiso; i++) { VECTOR(hyp)[i] = hyper[i]; }
  splicing_matrix_init(&smp, 0, 0);
  splicing_vector_init(&ll, 0);
  splicing_matrix_init(&ct, 0, 0);
  splicing_vector_init(&cc, 0);
  splicing_vector_int_init(&ass, 0);
  memset(&rd, 0, sizeof(rd));

  refh_install_rng(rng_mode, seed, gene_id, chain_id);
  if (start == 2) { refh_paint_stack(0.2 / (double) noiso / (double) noiso); }
  ret = splicing_miso_paired(&gff, 0, &position, cigars, readLength,
			     overhang, noChains, noIterations,
			     /*maxIterations=*/ 100000, noBurnIn, noLag,
			     &hyp, (splicing_miso_start_t) start,
			     (splicing_miso_stop_t) stop, /*start_psi=*/ 0,
			     /*fragmentProb=*/ 0, /*fragmentStart=*/ 0,
			     mean, var, numDevs, &smp, &ll,
			     /*match_matrix=*/ 0, /*class_templates=*/ 0,
			     /*class_counts=*/ 0, &ct, &cc, &ass, &rd);

  if (!ret) {
    refh_copy_matrix(&smp, samples);
    if (logLik) {
      memcpy(logLik, VECTOR(ll), sizeof(double) * splicing_vector_size(&ll));
    }
    *ncls = splicing_matrix_ncol(&ct);
    refh_copy_matrix(&ct, bin_class_templates);
    if (bin_class_counts) {
      memcpy(bin_class_counts, VECTOR(cc), sizeof(double) * (*ncls));
    }
    if (assignment) {
      memcpy(assignment, VECTOR(ass), sizeof(int) * (nreads / 2));
    }
    memcpy(rundata, &rd, sizeof(rd));
  }
  splicing_vector_int_destroy(&ass);
  splicing_vector_destroy(&cc);
  splicing_matrix_destroy(&ct);
  splicing_vector_destroy(&ll);
  splicing_matrix_destroy(&smp);
  splicing_vector_destroy(&hyp);
  splicing_vector_int_destroy(&position);
  splicing_gff_destroy(&gff);
  return ret;
}

/* ------------------------------------------------------------------ */
/* synthetic reads from the reference's own simulators                  */
/* (src/simulator.c:68-196 and :221-442).  cigar_buf receives the        */
/* NUL-terminated strings back to back, cigar_off[i] their offsets.      */

static int refh_pack_cigars(const splicing_strvector_t *cig, int n,
			    char *cigar_buf, int cap, int *cigar_off) {
  int i, at = 0;
  for (i = 0; i < n; i++) {
    const char *s = splicing_strvector_get(cig, i);
    int l = (int) strlen(s) + 1;
    if (at + l > cap) { return -1; }
    memcpy(cigar_buf + at, s, l);
    cigar_off[i] = at;
    at += l;
  }
  cigar_off[n] = at;
  return 0;
}

int refh_simulate_se(int nexons, const int *exons, int isolen,
		     const int *isoforms, const double *expression,
		     int noreads, int readLength, int rng_mode, uint64_t seed,
		     int *isoform_out, int *pos, char *cigar_buf, int cap,
		     int *cigar_off) {
  splicing_gff_t gff;
  splicing_vector_t expr;
  splicing_vector_int_t iso, position;
  splicing_strvector_t cig;
  size_t noiso;
  int i, ret;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_gff_noiso_one(&gff, 0, &noiso);
  splicing_vector_init(&expr, noiso);
  for (i = 0; i < (int) noiso; i++) { VECTOR(expr)[i] = expression[i]; }
  splicing_vector_int_init(&iso, 0);
  splicing_vector_int_init(&position, 0);
  splicing_strvector_init(&cig, 0);
  refh_install_rng(rng_mode, seed, 0xfffffffeu, 0);
  ret = splicing_simulate_reads(&gff, 0, &expr, noreads, readLength, &iso,
				&position, &cig, 0);
  if (!ret) {
    for (i = 0; i < noreads; i++) {
      if (isoform_out) { isoform_out[i] = VECTOR(iso)[i]; }
      pos[i] = VECTOR(position)[i];
    }
    ret = refh_pack_cigars(&cig, noreads, cigar_buf, cap, cigar_off);
  }
  splicing_strvector_destroy(&cig);
  splicing_vector_int_destroy(&position);
  splicing_vector_int_destroy(&iso);
  splicing_vector_destroy(&expr);
  splicing_gff_destroy(&gff);
  return ret;
}

/* noreads = number of PAIRS; pos / cigars get 2*noreads entries. */
int refh_simulate_pe(int nexons, const int *exons, int isolen,
		     const int *isoforms, const double *expression,
		     int noreads, int readLength, double mean, double var,
		     double numDevs, int rng_mode, uint64_t seed,
		     int *isoform_out, int *pos, char *cigar_buf, int cap,
		     int *cigar_off) {
  splicing_gff_t gff;
  splicing_vector_t expr;
  splicing_vector_int_t iso, position;
  splicing_strvector_t cig;
  size_t noiso;
  int i, ret, n;
  ret = refh_make_gene(&gff, nexons, exons, isolen, isoforms);
  if (ret) { return ret; }
  splicing_gff_noiso_one(&gff, 0, &noiso);
  splicing_vector_init(&expr, noiso);
  for (i = 0; i < (int) noiso; i++) { VECTOR(expr)[i] = expression[i]; }
  splicing_vector_int_init(&iso, 0);
  splicing_vector_int_init(&position, 0);
  splicing_strvector_init(&cig, 0);
  refh_install_rng(rng_mode, seed, 0xfffffffeu, 0);
  ret = splicing_simulate_paired_reads(&gff, 0, &expr, noreads, readLength,
				       /*fragmentProb=*/ 0,
				       /*fragmentStart=*/ 0, mean, var,
				       numDevs, &iso, &position, &cig, 0);
  if (!ret) {
    n = splicing_vector_int_size(&position);
    for (i = 0; i < n; i++) { pos[i] = VECTOR(position)[i]; }
    if (isoform_out) {
      int ni = splicing_vector_int_size(&iso);
      for (i = 0; i < ni; i++) { isoform_out[i] = VECTOR(iso)[i]; }
    }
    ret = refh_pack_cigars(&cig, n, cigar_buf, cap, cigar_off);
  }
  splicing_strvector_destroy(&cig);
  splicing_vector_int_destroy(&position);
  splicing_vector_int_destroy(&iso);
  splicing_vector_destroy(&expr);
  splicing_gff_destroy(&gff);
  return ret;
}
