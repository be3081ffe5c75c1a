/* mcq_prep.h - everything dmodel derives from the loaded sequences before the first likelihood
 * evaluation (dmodel.c:489-849): consensus, per-query consensus + gap fill, encodings, per-site
 * codon lists, rate_out_of_codons. */
#ifndef MCQ_PREP_H_
#define MCQ_PREP_H_

typedef struct {
  int S, N, Q, L, H, nbases;
  char *ref_triplets;      /* [N][S] base-5 (amino_acids.h:28-30), gaps kept */
  char *ref_codons;        /* [N][S] base-4, gaps mapped to the consensus */
  char *query_codons;      /* [Q][S] */
  char *cons_query_nucls;  /* [Q][3S] codes 0..3 */
  char *hla;               /* [Q][H] '0'/'1' */
  char *consensus_codons;  /* [S] */
  int *n_site_codons;      /* [S] */
  char *site_codons;       /* [S][64] */
  int alpha_S[64], alpha_V[64], beta_S[64], beta_V[64];
} McqData;

int mcq_base_code(int ch);   /* amino_acids.c:10-17 */
/* majority base per column over `n` rows (rows = seqs[idx[i]] when idx != NULL), ties T<C<A<G
 * (amino_acids.c:94-141); writes ASCII into out[len] */
void mcq_consensus(char *const *seqs, const int *idx, int n, int len, char *out);
/* builds everything in `d` from the sequences; fills gaps of queries/references in place the way
 * dmodel.c:664-701 does.  closest is [Q][L]. */
void mcq_prepare(McqData *d, int device, char **ref_seqs, int N, char **query_seqs, int Q, char **cons_seqs,
                 int ncons, char **hla_rows, int H, const int *closest, int L);
void mcq_data_free(McqData *d);

#endif
