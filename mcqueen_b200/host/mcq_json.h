/* mcq_json.h - reading a saved chain state (the format of write_json.c:9-67 / mcq_write_state_json),
 * what the reference does in load_from_json.c:47-176 for `dmodel -r`. */
#ifndef MCQ_JSON_H_
#define MCQ_JSON_H_

#include "mcq_windows.h"

/* Loads the LAST state of the "chain" array: llk, mu, omega[], R[], the window sets of every HLA type
 * (start + coefficient; lengths are recomputed, as the reference does) and the closest-L lists.
 * Dies with a message on any mismatch with the expected sizes. */
void mcq_load_state_json(const char *path, double *llk, double *mu, double *R, double *omega,
                         McqHlaWindows *hw, int num_sites, int num_hla, int num_query, int closest_n,
                         int *closest_refs);
/* dmodel -p (load_from_json.c:248-295): {"kappa": k, "phi": f, "codons": [{"codon_prior": p} x 64]} */
void mcq_load_fixed_parameters_json(const char *path, double *kappa, double *phi, double *codon_prior);
/* dmodel -m (load_from_json.c:178-246): {"true_sequences": [{"number": q, "closest_n": [refs the query was copied
 * from]}, ...]}; which_seqs is [num_query][num_sites], num_seqs [num_query] */
void mcq_load_true_mosaics_json(const char *path, int num_query, int num_sites, int *which_seqs, int *num_seqs);
#endif
