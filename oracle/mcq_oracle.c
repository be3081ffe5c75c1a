/* TEST INFRASTRUCTURE ONLY - see mcq_oracle.h for what this is and who may use it.
 *
 * Arithmetic notes that make this restatement bit-faithful to the reference as
 * compiled by its own Makefile (-O3, x86-64 baseline, hence no FMA contraction):
 * built with -ffp-contract=off; every sum accumulates in ascending state index;
 * "x * R / N" is evaluated left to right with N converted from int.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "mcq_oracle.h"

static const double ORC_BIG = 1000000.0;   /* constants.c:13 */
static const double ORC_BIGI = 100.0;      /* constants.c:14 */

/* ------------------------------------------------------------------ encodings */

int orc_base_code(int ch)
{
  switch(ch) {
    case 'T': case 't': return 0;
    case 'C': case 'c': return 1;
    case 'A': case 'a': return 2;
    case 'G': case 'g': return 3;
    default: return 4;
  }
}

/* A triplet is a base-5 number d0*25+d1*5+d2 (amino_acids.h:28-30); digit 4 is
 * "unknown" and is replaced by the query's local consensus code at the same
 * codon position (amino_acids.c:50-66). */
int orc_triplet_to_codon(int triplet, const signed char cons3[3])
{
  int d0 = (triplet / 25) % 5, d1 = (triplet / 5) % 5, d2 = triplet % 5;
  if(d0 == 4) d0 = cons3[0];
  if(d1 == 4) d1 = cons3[1];
  if(d2 == 4) d2 = cons3[2];
  return d0 * 16 + d1 * 4 + d2;
}

/* ------------------------------------------------- substitution-class tables */

/* Standard genetic code in TCAG order (amino_acids.c:20-28). */
static const char orc_aa[65] =
  "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

int orc_subst_class(int kind, int c1, int c2)
{
  int ndiff = 0, a = 0, b = 0, p;
  for(p = 0; p < 3; p++) {
    int x = (c1 >> (4 - 2*p)) & 3, y = (c2 >> (4 - 2*p)) & 3;
    if(x != y) { ndiff++; a = x; b = y; }
  }
  if(kind == 4) return ndiff >= 2;
  if(ndiff != 1) return 0;
  int transition = ((a ^ b) == 1);          /* T<->C (0,1) or A<->G (2,3) */
  int synonymous = (orc_aa[c1] == orc_aa[c2]);
  switch(kind) {
    case 0: return !synonymous && transition;
    case 1: return !synonymous && !transition;
    case 2: return synonymous && transition;
    case 3: return synonymous && !transition;
  }
  return 0;
}

int orc_subst_count(int kind, int c1)
{
  int c2, n = 0;
  for(c2 = 0; c2 < ORC_NCODON; c2++) n += orc_subst_class(kind, c1, c2);
  return n;
}

/* ------------------------------------------------------------ forward / backward */

static inline double emis(const double *c2c, int S, const signed char *ref_triplets,
                          const int *closest, const signed char *cons_nucls, int r, int s)
{
  int codon = orc_triplet_to_codon(ref_triplets[(size_t)closest[r] * S + s], cons_nucls + 3*s);
  return c2c[(size_t)codon * S + s];
}

void orc_forward(int S, int L, int N, const double *c2c,
                 const signed char *ref_triplets, const double *R,
                 const double *F_in, int *rnorm_in, int *rnorm_out, double *F_out,
                 int start, int end, const int *closest, const signed char *cons_nucls)
{
  const double *prev = F_in;       /* matrix the previous column is read from */
  int s = start, r, count;

  if(s == 0) {
    /* column 0 is the bare emission, never rescaled (forward_backward.c:27-39,102-111) */
    for(r = 0; r < L; r++)
      F_out[(size_t)r * S] = emis(c2c, S, ref_triplets, closest, cons_nucls, r, 0);
    rnorm_out[0] = 0;
    rnorm_in[0] = 0;
    prev = F_out;
    s = 1;
  }
  count = rnorm_in[s - 1];

  for(; s <= end; s++) {
    double jump = 0, colsum = 0;
    for(r = 0; r < L; r++) jump += prev[(size_t)r * S + s - 1];
    jump = jump * R[s] / N;                       /* forward_backward.c:138 */
    for(r = 0; r < L; r++) {
      double v = jump;
      v += (1 - R[s]) * prev[(size_t)r * S + s - 1];
      v *= emis(c2c, S, ref_triplets, closest, cons_nucls, r, s);
      F_out[(size_t)r * S + s] = v;
      colsum += v;
    }
    rnorm_out[s] = count;
    if(colsum < ORC_BIGI) {                       /* forward_backward.c:163-167 */
      rnorm_out[s]++;
      for(r = 0; r < L; r++) F_out[(size_t)r * S + s] *= ORC_BIG;
    }
    prev = F_out;
    count = rnorm_out[s];
  }
}

void orc_backward(int S, int L, int N, const double *c2c, const signed char *ref_triplets,
                  const double *R, double *B, int *B_rnorm, int update,
                  const int *closest, const signed char *cons_nucls)
{
  int s = update, r;
  if(s == S - 1) {                                /* forward_backward.c:197-203 */
    for(r = 0; r < L; r++) B[(size_t)r * S + S - 1] = 1;
    B_rnorm[S - 1] = 0;
    s--;
  }
  for(; s >= 0; s--) {
    double jump = 0, colsum = 0;
    for(r = 0; r < L; r++)
      jump += B[(size_t)r * S + s + 1] * emis(c2c, S, ref_triplets, closest, cons_nucls, r, s + 1);
    jump = jump * R[s + 1] / N;
    for(r = 0; r < L; r++) {
      double v = jump;
      v += (1 - R[s + 1]) * B[(size_t)r * S + s + 1] *
           emis(c2c, S, ref_triplets, closest, cons_nucls, r, s + 1);
      B[(size_t)r * S + s] = v;
      colsum += v;
    }
    B_rnorm[s] = B_rnorm[s + 1];
    if(colsum < ORC_BIGI) {
      B_rnorm[s]++;
      for(r = 0; r < L; r++) B[(size_t)r * S + s] *= ORC_BIG;
    }
  }
}

double orc_site_llk(int S, int L, const double *F, const int *Fn,
                    const double *B, const int *Bn, int site)
{
  double dot = 0;
  int r;
  if(B != NULL) {
    for(r = 0; r < L; r++) dot += F[(size_t)r * S + site] * B[(size_t)r * S + site];
    return log(dot) - (Fn[site] + Bn[site]) * log(ORC_BIG);
  }
  for(r = 0; r < L; r++) dot += F[(size_t)r * S + site];
  return log(dot) - Fn[site] * log(ORC_BIG);
}

/* -------------------------------------------------------------------- emissions */

/* single-step rate c1->c2 given omega: omega*(kappa*NS_TS+NS_TV) + kappa*S_TS + S_TV,
 * evaluated in the reference's association order (mcmc_moves.c:85-88,111-112). */
static inline double step_rate(double omega, double kappa, int c1, int c2)
{
  return omega * (kappa * orc_subst_class(0, c1, c2) + orc_subst_class(1, c1, c2)) +
         kappa * orc_subst_class(2, c1, c2) + orc_subst_class(3, c1, c2);
}

void orc_emission(int S, int N, int Q, int H,
                  const int *n_site_codons, const signed char *site_codons,
                  double *c2c, const double *rate_out,
                  int start, int end, int which_hla, double mu,
                  const double *esc, const double *rev,
                  const char *hla, const signed char *cons_codons,
                  const double *omega, const signed char *query_codons,
                  double *A, int mode, double kappa, double phi, const double *prior)
{
  int q, s, h, k;
  for(q = 0; q < Q; q++) {
    const char *types = hla + (size_t)q * H;
    /* an escape-only update of one HLA leaves non-carriers untouched (mcmc_moves.c:61-62) */
    if(mode == ORC_ESC && which_hla != ORC_ALL_HLAS && types[which_hla] == '0') continue;
    double *cq = c2c + (size_t)q * ORC_NCODON * S, *Aq = A + (size_t)q * ORC_NCODON * S;

    for(s = start; s <= end; s++) {
      int cons = cons_codons[s], to = query_codons[(size_t)q * S + s];
      double esc_prod = 1, rev_coef = rev[s];     /* rev row 0 is shared (mcmc_moves.c:70) */
      double back = (1 - phi) * prior[to];        /* multi-step background term */
      for(h = 0; h < H; h++)
        if(types[h] == '1') esc_prod *= esc[(size_t)h * S + s];

      if(mode == ORC_BOTH || mode == ORC_REV) {
        /* rows for every non-consensus codon the panel shows at this site (:78-124) */
        for(k = 0; k < n_site_codons[s]; k++) {
          int c1 = site_codons[(size_t)s * ORC_NCODON + k];
          if(c1 == cons) continue;
          double rate = rate_out[(size_t)c1 * S + s] +
                        (rev_coef - 1) * step_rate(omega[s], kappa, c1, cons);
          double a = 1 - N / (N + rate * 4 * mu);
          double gamma = phi / rate;
          double m = step_rate(omega[s], kappa, c1, to);
          double v;
          if(to == cons) v = a * (gamma * rev_coef * m + back);   /* non-cons -> cons */
          else {
            v = a * (gamma * m + back);                           /* non-cons -> non-cons */
            if(c1 == to) v += (1 - a);
          }
          Aq[(size_t)c1 * S + s] = a;
          cq[(size_t)c1 * S + s] = v;
        }
      }
      if(mode == ORC_BOTH || mode == ORC_ESC) {
        /* the consensus row (:125-167) */
        int nts = orc_subst_count(0, cons), ntv = orc_subst_count(1, cons);
        double rate = rate_out[(size_t)cons * S + s] +
                      (esc_prod - 1) * omega[s] * (kappa * nts + ntv);
        double a = 1 - N / (N + rate * 4 * mu);
        double v;
        if(cons == 15 || cons == 35) {            /* TGG, ATG: no synonymous neighbours */
          double gamma = phi / (kappa * nts + ntv);
          v = a * (gamma * (kappa * orc_subst_class(0, cons, to) + orc_subst_class(1, cons, to)) + back);
        } else {
          double gamma = phi / rate;
          v = a * (gamma * (esc_prod * omega[s] *
                            (kappa * orc_subst_class(0, cons, to) + orc_subst_class(1, cons, to)) +
                            kappa * orc_subst_class(2, cons, to) + orc_subst_class(3, cons, to)) + back);
        }
        if(to == cons) v = (1 - a) * (1 - back) + back;
        Aq[(size_t)cons * S + s] = a;
        cq[(size_t)cons * S + s] = v;
      }
    }
  }
}

void orc_mu_rescale(int S, int N, int Q, const int *n_site_codons, const signed char *site_codons,
                    const double *c2c, const double *A, double *c2c_out, double *A_out,
                    const signed char *query_codons, double mu, double mu_new,
                    double phi, const double *prior)
{
  int q, s, k;
  for(q = 0; q < Q; q++)
    for(s = 0; s < S; s++) {
      int to = query_codons[(size_t)q * S + s];
      for(k = 0; k < n_site_codons[s]; k++) {
        int c1 = site_codons[(size_t)s * ORC_NCODON + k];
        size_t i = ((size_t)q * ORC_NCODON + c1) * S + s;
        double a = A[i], rate = 0, a_new, v;
        if(a != 0) rate = (N / (1 - a) - N) / (4 * mu);          /* mcmc_moves.c:1091-1095 */
        a_new = 1 - N / (N + rate * 4 * mu_new);
        if(to != c1) v = (a == 0) ? 0 : c2c[i] * a_new / a;
        else v = 1 - a_new + (1 - phi) * prior[c1] * a_new;
        A_out[i] = a_new;
        c2c_out[i] = v;
      }
    }
}

/* ------------------------------------------------------------------- closest-L */

void orc_hamming(const char *const *refs, const char *query, int N, int nbases, int *out)
{
  int r, i;
  for(r = 0; r < N; r++) {
    int mism = 0;
    for(i = 0; i < nbases; i++) {
      int qc = orc_base_code((unsigned char)query[i]);
      if(qc == 4) continue;                       /* unknown query base: skipped */
      /* the reference's ref-side "unknown" test is dnachar_to_code[ref != 4], i.e. table
       * entry 0 or 1, both 4, so it always passes (closest_seqs.c:32): an unknown ref base
       * opposite a known query base is a mismatch and the gap term (:46) is always 0 */
      mism += (qc != orc_base_code((unsigned char)refs[r][i]));
    }
    out[r] = 2 * mism;
  }
}

typedef struct { int idx, dist; } orc_pair;

/* stable merge sort by dist: equal distances keep ascending index, which is what glibc's
 * qsort (a merge sort when it can allocate) gives the reference */
static void sort_pairs(orc_pair *a, orc_pair *tmp, int n)
{
  if(n < 2) return;
  int h = n / 2, i = 0, j = h, k = 0;
  sort_pairs(a, tmp, h);
  sort_pairs(a + h, tmp, n - h);
  while(i < h && j < n) tmp[k++] = (a[j].dist < a[i].dist) ? a[j++] : a[i++];
  while(i < h) tmp[k++] = a[i++];
  while(j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(orc_pair));
}

void orc_closest(int N, const int *hamming, int n, int *out)
{
  orc_pair *d = malloc((size_t)N * sizeof(orc_pair)), *tmp = malloc((size_t)N * sizeof(orc_pair));
  int i;
  for(i = 0; i < N; i++) { d[i].idx = i; d[i].dist = hamming[i]; }
  sort_pairs(d, tmp, N);

  /* tie window around the cut, keyed on the element one PAST the cut (closest_seqs.c:84-88) */
  int lo = n - 1, hi = n - 1, edge = d[n].dist;
  while(lo > 0 && d[lo - 1].dist == edge) lo--;
  while(hi + 1 < N && d[hi + 1].dist == edge) hi++;

  /* forward "Knuth" shuffle with j = (int)(limit * random()/RAND_MAX) (closest_seqs.c:57-64,
   * util.c:24-30) */
  for(i = 0; i <= hi - lo; i++) {
    double u = (i + 1) * ((double)random() / (double)RAND_MAX);
    int j = (int)u;
    orc_pair t = d[lo + i]; d[lo + i] = d[lo + j]; d[lo + j] = t;
  }
  for(i = 0; i < n; i++) out[i] = d[i].idx;
  free(d); free(tmp);
}
