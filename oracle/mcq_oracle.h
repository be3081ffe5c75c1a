/* TEST INFRASTRUCTURE ONLY - the CPU oracle for the McQueen hot path.
 *
 * A plain-C restatement (written from scratch, not copied) of the reference's
 * algorithm for the per-query copying-model HMM likelihood path.  Each function
 * cites the reference file:line it follows.  Array shapes are the REFERENCE's
 * shapes (state-major F[L][S], codon-major c2c[64][S]) so that outputs can be
 * compared element by element with the reference's own objects (oracle/_ref).
 *
 * Pinning: the reference ships no tests / golden vectors for this path
 * (SURVEY.md §4, §8c).  This oracle is pinned against outputs of the reference
 * itself, compiled unmodified into oracle/_ref/libmcq_ref.so, on seeded synthetic
 * inputs (tests/test_oracle_vs_ref.py, run wherever oracle/_ref exists) and
 * against the committed fixtures under tests/golden/ generated from that library
 * (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (libmcq_b200.so) never does.
 */
#ifndef MCQ_ORACLE_H_
#define MCQ_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NCODON 64

/* emission-update selectors, same values as mcmc_moves.h:6-9 */
#define ORC_ESC  0
#define ORC_REV  1
#define ORC_BOTH 2
#define ORC_ALL_HLAS (-1)

/* amino_acids.c:10-17 : T,C,A,G (either case) -> 0..3, anything else -> 4 */
int orc_base_code(int ch);

/* amino_acids.c:50-66 */
int orc_triplet_to_codon(int triplet, const signed char cons3[3]);

/* constants.c:405-690 substitution-class tables, derived from the genetic code
 * (amino_acids.c:20-28): kind 0=NS_TS 1=NS_TV 2=S_TS 3=S_TV 4=two_step */
int orc_subst_class(int kind, int c1, int c2);
/* kind 0=alpha_S 1=alpha_V 2=beta_S 3=beta_V (row sums of the above) */
int orc_subst_count(int kind, int c1);

/* forward_backward.c:11-76 and :80-171 in one routine.
 * start==0 reseeds column 0 (and zeroes rnorm_in[0], rnorm_out[0], as :108-109).
 * F_in/F_out are [L][S]; may alias. */
void orc_forward(int S, int L, int N, const double *c2c /*[64][S]*/,
                 const signed char *ref_triplets /*[N][S]*/, const double *R,
                 const double *F_in, int *rnorm_in, int *rnorm_out, double *F_out,
                 int start, int end, const int *closest /*[L]*/,
                 const signed char *cons_nucls /*[3S]*/);

/* forward_backward.c:173-244 */
void orc_backward(int S, int L, int N, const double *c2c, const signed char *ref_triplets,
                  const double *R, double *B, int *B_rnorm, int update,
                  const int *closest, const signed char *cons_nucls);

/* mcmc_moves.c:220-223 / dmodel.c:899-901: log(sum_r F[r][s]*B[r][s]) - (Fn+Bn)*log(BIG);
 * B==NULL means the forward-only form (B=1, Bn=0). */
double orc_site_llk(int S, int L, const double *F, const int *Fn,
                    const double *B, const int *Bn, int site);

/* mcmc_moves.c:33-170 */
void orc_emission(int S, int N, int Q, int H,
                  const int *n_site_codons /*[S]*/, const signed char *site_codons /*[S][64]*/,
                  double *c2c /*[Q][64][S]*/, const double *rate_out /*[64][S]*/,
                  int start, int end, int which_hla, double mu,
                  const double *esc /*[H][S]*/, const double *rev /*[H][S] (row 0 used)*/,
                  const char *hla /*[Q][H] of '0'/'1'*/, const signed char *cons_codons /*[S]*/,
                  const double *omega, const signed char *query_codons /*[Q][S]*/,
                  double *A /*[Q][64][S]*/, int mode,
                  double kappa, double phi, const double *prior /*[64]*/);

/* mcmc_moves.c:1082-1111 */
void orc_mu_rescale(int S, int N, int Q, const int *n_site_codons, const signed char *site_codons,
                    const double *c2c, const double *A, double *c2c_out, double *A_out,
                    const signed char *query_codons, double mu, double mu_new,
                    double phi, const double *prior);

/* closest_seqs.c:17-48 (including the always-true ref-side test at :32) */
void orc_hamming(const char *const *refs, const char *query, int N, int nbases, int *out);

/* closest_seqs.c:66-99; draws from libc random() exactly as util.c:24-30 does */
void orc_closest(int N, const int *hamming, int n, int *out);

#ifdef __cplusplus
}
#endif
#endif
