/* TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * CPU baseline harness: times the REFERENCE's own, unmodified functions (linked from the objects
 * oracle/Makefile compiles out of /root/reference/src) on a problem written by bench.py, over
 * disjoint shards of queries from host threads.  Queries are independent, so sharding them in
 * the harness changes no reference code.  One "evaluation" per query is what
 * dmodel does at start-up (dmodel.c:873-902): create_codon_to_codon_HLA for the query's rows,
 * initialise_F_numerical_recipes, and the log-likelihood reduction; optionally update_B too.
 *
 * usage: ref_bench <problem.bin> <threads> <num_queries_to_time> <with_backward 0|1> <repeats>
 * prints one JSON line.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* the reference's symbols (forward_backward.h, mcmc_moves.h, constants.h); VLA parameters decay to
 * these flat pointers */
extern void initialise_F_numerical_recipes(int, int, int, double *, char *, char *, double *, double *, int *,
                                           int, int, int *, int, char *);
extern void update_B_numerical_recipes(int, int, int, double *, char *, char *, double *, double *, int *, int,
                                       int *, int, char *);
extern void create_codon_to_codon_HLA(int, int, int, int, int *, char *, double *, double *, int, int, int,
                                      double, double *, double *, char *const *, const char *, const double *,
                                      const int *, const int *, char *, char *, double *, int);
extern double kappa, phi, prior_C1[64];
extern const double BIG;

typedef struct {
  int S, N, Q, L, H;
  char *ref_triplets, *ref_codons, *query_codons, *cons_nucls, *hla, *cons_codons, *site_codons, *toc;
  int *closest, *n_site, *ns_ts, *ns_tv;
  double *R, *omega, *esc, *rev, *rate_out, mu;
} Problem;

static void *rd(FILE *f, size_t n) {
  void *p = malloc(n ? n : 1);
  if (fread(p, 1, n, f) != n) { fprintf(stderr, "ref_bench: short read\n"); exit(1); }
  return p;
}

static void load(const char *path, Problem *p) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(1); }
  int hdr[5];
  double sc[3];
  if (fread(hdr, sizeof(int), 5, f) != 5 || fread(sc, sizeof(double), 3, f) != 3) exit(1);
  p->S = hdr[0]; p->N = hdr[1]; p->Q = hdr[2]; p->L = hdr[3]; p->H = hdr[4];
  p->mu = sc[0]; kappa = sc[1]; phi = sc[2];
  size_t S = p->S, N = p->N, Q = p->Q, L = p->L, H = p->H;
  double *prior = rd(f, 64 * sizeof(double));
  memcpy(prior_C1, prior, 64 * sizeof(double));
  free(prior);
  p->ref_triplets = rd(f, N * S); p->ref_codons = rd(f, N * S); p->query_codons = rd(f, Q * S);
  p->cons_nucls = rd(f, Q * 3 * S); p->hla = rd(f, Q * (H + 1)); p->cons_codons = rd(f, S);
  p->site_codons = rd(f, S * 64); p->toc = rd(f, S * 64 * 64);
  p->closest = rd(f, Q * L * sizeof(int)); p->n_site = rd(f, S * sizeof(int));
  p->ns_ts = rd(f, S * sizeof(int)); p->ns_tv = rd(f, S * sizeof(int));
  p->R = rd(f, S * sizeof(double)); p->omega = rd(f, S * sizeof(double));
  p->esc = rd(f, H * S * sizeof(double)); p->rev = rd(f, H * S * sizeof(double));
  p->rate_out = rd(f, 64 * S * sizeof(double));
  fclose(f);
}

typedef struct {
  const Problem *p;
  int q0, q1, with_b, repeats;
  double llk;
} Job;

static void *work(void *arg) {
  Job *j = arg;
  const Problem *p = j->p;
  const size_t S = p->S, L = p->L, H = p->H;
  double *c2c = calloc(64 * S, sizeof(double)), *A = calloc(64 * S, sizeof(double));
  double *F = malloc(L * S * sizeof(double)), *B = malloc(L * S * sizeof(double));
  int *Fn = malloc(S * sizeof(int)), *Bn = malloc(S * sizeof(int));
  double acc = 0;
  for (int rep = 0; rep < j->repeats; rep++) {
    acc = 0;
    for (int q = j->q0; q < j->q1; q++) {
      char *hrow = p->hla + (size_t)q * (H + 1);
      char *const rows[1] = {hrow};
      create_codon_to_codon_HLA(p->S, p->N, 1, p->H, p->n_site, p->site_codons, c2c, p->rate_out, 0, p->S - 1,
                                -1, p->mu, p->esc, p->rev, rows, p->cons_codons, p->omega, p->ns_ts, p->ns_tv,
                                p->toc, p->query_codons + (size_t)q * S, A, 2);
      initialise_F_numerical_recipes(p->S, p->L, p->N, c2c, p->ref_codons, p->ref_triplets, p->R, F, Fn, 0,
                                     p->S - 1, p->closest + (size_t)q * L, 3 * p->S, p->cons_nucls + (size_t)q * 3 * S);
      if (j->with_b)
        update_B_numerical_recipes(p->S, p->L, p->N, c2c, p->ref_codons, p->ref_triplets, p->R, B, Bn, p->S - 1,
                                   p->closest + (size_t)q * L, 3 * p->S, p->cons_nucls + (size_t)q * 3 * S);
      double s = 0;
      for (size_t r = 0; r < L; r++) s += F[r * S + S - 1];
      acc += log(s) - Fn[S - 1] * log(BIG);
    }
  }
  j->llk = acc;
  free(c2c); free(A); free(F); free(B); free(Fn); free(Bn);
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: ref_bench problem.bin threads nq with_backward repeats\n"); return 2; }
  Problem p;
  load(argv[1], &p);
  int T = atoi(argv[2]), nq = atoi(argv[3]), with_b = atoi(argv[4]), reps = atoi(argv[5]);
  if (nq > p.Q || nq <= 0) nq = p.Q;
  if (T < 1) T = 1;
  if (T > nq) T = nq;
  pthread_t *th = malloc(T * sizeof(pthread_t));
  Job *jobs = malloc(T * sizeof(Job));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < T; t++) {
    jobs[t].p = &p; jobs[t].with_b = with_b; jobs[t].repeats = reps;
    jobs[t].q0 = (int)((long long)nq * t / T); jobs[t].q1 = (int)((long long)nq * (t + 1) / T);
    pthread_create(&th[t], NULL, work, &jobs[t]);
  }
  double llk = 0;
  for (int t = 0; t < T; t++) { pthread_join(th[t], NULL); llk += jobs[t].llk; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  double cells = (double)nq * p.S * p.L * reps;
  printf("{\"cells\": %.0f, \"seconds\": %.6f, \"cells_per_s\": %.6e, \"threads\": %d, \"queries\": %d, "
         "\"repeats\": %d, \"with_backward\": %d, \"llk\": %.10f}\n",
         cells, sec, cells / sec, T, nq, reps, with_b, llk);
  return 0;
}
