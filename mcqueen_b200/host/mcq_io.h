/* mcq_io.h - input and output of the dmodel driver: FASTA, HLA csv, state JSON.
 * Formats follow the reference (load_data.c:66-157, write_json.c:9-67). */
#ifndef MCQ_IO_H_
#define MCQ_IO_H_

#include <stdio.h>
#include "mcq_windows.h"

void mcq_die(const char *fmt, ...) __attribute__((format(printf, 1, 2), noreturn));

/* all sequences of a FASTA file as NUL-terminated strings; returns their number */
int mcq_load_fasta(const char *path, char ***seqs);
/* a 0/1 csv with a header line and one leading label column; rows[q][h] in '0'/'1' plus NUL.
 * Returns the number of type columns. */
int mcq_load_hla_csv(const char *path, char ***rows, int num_rows);
int mcq_mkpath(const char *path);

void mcq_write_state_json(FILE *fh, int state, double llk, double mu, int num_sites, const double *omega,
                          const double *R, const McqHlaWindows *hw, int num_hla, int num_query,
                          int closest_n, const int *closest_refs);
#endif
