/*
 * oracle/ref_stubs.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Link-time stubs for the four reference symbols that the default sampler
 * path (START_AUTO, ALGO_REASSIGN; misopy/miso_sampler.py:210,322) never
 * reaches but miso.c / solve.c reference: NNLS + BLAS (START_LINEAR,
 * src/miso.c:410-443, src/solve.c:308-536) and the theoretical assignment
 * matrices (ALGO_CLASSES, src/miso.c:790-803).  Stubbing them keeps the 233
 * vendored f2c/LAPACK files out of the oracle build (SURVEY.md section 2 marks
 * them out of scope).  Deliberately no reference header is included here.
 */
#include <stdio.h>
#include <stdlib.h>

static int refh_offpath(const char *what) {
  fprintf(stderr, "oracle/_ref: off-path symbol %s reached -- the oracle "
          "build leaves NNLS / assignment matrices out on purpose\n", what);
  abort();
  return 1;
}
int splicing_nnls(void) { return refh_offpath("splicing_nnls"); }
int splicing_dgemv(void) { return refh_offpath("splicing_dgemv"); }
int splicing_assignment_matrix(void) {
  return refh_offpath("splicing_assignment_matrix");
}
int splicing_paired_assignment_matrix(void) {
  return refh_offpath("splicing_paired_assignment_matrix");
}
