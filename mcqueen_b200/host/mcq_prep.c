#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mcq_b200.h"
#include "mcq_io.h"
#include "mcq_prep.h"

int mcq_base_code(int ch)
{
  switch(ch) {
    case 'T': case 't': return 0;
    case 'C': case 'c': return 1;
    case 'A': case 'a': return 2;
    case 'G': case 'g': return 3;
    default: return 4;
  }
}

void mcq_consensus(char *const *seqs, const int *idx, int n, int len, char *out)
{
  static const char letters[4] = {'T', 'C', 'A', 'G'};
  for(int i = 0; i < len; i++) {
    int cnt[5] = {0, 0, 0, 0, 0};
    for(int r = 0; r < n; r++) cnt[mcq_base_code((unsigned char)seqs[idx ? idx[r] : r][i])]++;
    int best = 0;
    for(int b = 1; b < 4; b++) if(cnt[b] > cnt[best]) best = b;
    out[i] = letters[best];
  }
  out[len] = '\0';
}

void mcq_prepare(McqData *d, int device, char **ref_seqs, int N, char **query_seqs, int Q, char **cons_seqs,
                 int ncons, char **hla_rows, int H, const int *closest, int L)
{
  const int nb = (int)strlen(ref_seqs[0]);
  const int S = nb / 3;
  d->S = S; d->N = N; d->Q = Q; d->L = L; d->H = H; d->nbases = nb;
  d->ref_triplets = malloc((size_t)N * S);
  d->ref_codons = malloc((size_t)N * S);
  d->query_codons = malloc((size_t)Q * S);
  d->cons_query_nucls = malloc((size_t)Q * nb);
  d->hla = malloc((size_t)Q * H);
  d->consensus_codons = malloc(S);
  d->n_site_codons = malloc(sizeof(int) * S);
  d->site_codons = calloc((size_t)S * 64, 1);

  char *consensus = malloc(nb + 1), *local = malloc(nb + 1);
  mcq_consensus(cons_seqs, NULL, ncons, nb, consensus);

  /* per query: consensus of its closest references (on the device: Q*L*nb character reads); unknown
   * query bases take it (dmodel.c:664-681) */
  {
    static const char letters[4] = {'T', 'C', 'A', 'G'};
    char *rflat = malloc((size_t)N * nb);
    for(int r = 0; r < N; r++) memcpy(rflat + (size_t)r * nb, ref_seqs[r], nb);
    if(mcq_consensus_batch(device, rflat, nb, N, closest, Q, L, nb, d->cons_query_nucls) != 0)
      mcq_die("mcq_consensus_batch: %s", mcq_last_error());
    free(rflat);
    for(int q = 0; q < Q; q++) {
      for(int i = 0; i < nb; i++)
        if(mcq_base_code((unsigned char)query_seqs[q][i]) == 4)
          query_seqs[q][i] = letters[(int)d->cons_query_nucls[(size_t)q * nb + i]];
      memcpy(d->hla + (size_t)q * H, hla_rows[q], H);
    }
  }
  /* triplets keep unknown bases (dmodel.c:684-686); codons are taken after mapping them to the
   * global consensus (dmodel.c:693-701) */
  for(int r = 0; r < N; r++)
    for(int s = 0; s < S; s++) {
      const unsigned char *t = (const unsigned char *)ref_seqs[r] + 3 * s;
      d->ref_triplets[(size_t)r * S + s] = (char)(mcq_base_code(t[0]) * 25 + mcq_base_code(t[1]) * 5 + mcq_base_code(t[2]));
    }
  for(int r = 0; r < N; r++)
    for(int i = 0; i < nb; i++)
      if(mcq_base_code((unsigned char)ref_seqs[r][i]) == 4) ref_seqs[r][i] = consensus[i];
  for(int r = 0; r < N; r++)
    for(int s = 0; s < S; s++) {
      const unsigned char *t = (const unsigned char *)ref_seqs[r] + 3 * s;
      d->ref_codons[(size_t)r * S + s] = (char)(mcq_base_code(t[0]) * 16 + mcq_base_code(t[1]) * 4 + mcq_base_code(t[2]));
    }
  for(int q = 0; q < Q; q++)
    for(int s = 0; s < S; s++) {
      const unsigned char *t = (const unsigned char *)query_seqs[q] + 3 * s;
      d->query_codons[(size_t)q * S + s] = (char)(mcq_base_code(t[0]) * 16 + mcq_base_code(t[1]) * 4 + mcq_base_code(t[2]));
    }
  for(int s = 0; s < S; s++) {
    const unsigned char *t = (const unsigned char *)consensus + 3 * s;
    d->consensus_codons[s] = (char)(mcq_base_code(t[0]) * 16 + mcq_base_code(t[1]) * 4 + mcq_base_code(t[2]));
  }
  /* codons the panel shows at each site, ascending (amino_acids.c:155-176) */
  for(int s = 0; s < S; s++) {
    char seen[64] = {0};
    for(int r = 0; r < N; r++) seen[(int)d->ref_codons[(size_t)r * S + s]] = 1;
    int k = 0;
    for(int c = 0; c < 64; c++) if(seen[c]) d->site_codons[(size_t)s * 64 + k++] = (char)c;
    d->n_site_codons[s] = k;
  }
  /* alpha/beta: numbers of (non-)synonymous transitions/transversions out of each codon */
  for(int c1 = 0; c1 < 64; c1++) {
    d->alpha_S[c1] = d->alpha_V[c1] = d->beta_S[c1] = d->beta_V[c1] = 0;
    for(int c2 = 0; c2 < 64; c2++)
      switch(mcq_codon_class(c1, c2)) {
        case 1: d->alpha_S[c1]++; break;
        case 2: d->alpha_V[c1]++; break;
        case 3: d->beta_S[c1]++; break;
        case 4: d->beta_V[c1]++; break;
        default: break;
      }
  }
  printf("Consensus:\n%s\n", consensus);
  free(consensus); free(local);
}

void mcq_data_free(McqData *d)
{
  free(d->ref_triplets); free(d->ref_codons); free(d->query_codons); free(d->cons_query_nucls);
  free(d->hla); free(d->consensus_codons); free(d->n_site_codons); free(d->site_codons);
}
