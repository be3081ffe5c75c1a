/* TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * CPU baseline harness: times the REFERENCE's own, unmodified functions (linked from the objects
 * oracle/Makefile compiles out of /root/reference/src) on a problem written by bench.py, over
 * disjoint shards of queries from host threads.  Queries are independent, so sharding them in
 * the harness changes no reference code.  One "evaluation" per query is what
 * dmodel does at start-up (dmodel.c:873-902): create_codon_to_codon_HLA for the query's rows,
 * initialise_F_numerical_recipes, and the log-likelihood reduction; optionally update_B too.
 *
 * usage: ref_bench <problem.bin> <threads> <num_queries_to_time> <mode> <repeats>
 *   mode 0  full evaluation: emission + initialise_F + llk          (the headline region)
 *        1  the same followed by update_B (update = S-1)
 *        2  update_B alone (emission and forward done outside the timed region)
 *        3  window proposal: update_F over sites S/3 .. S/3+S/6-1 into a tmp matrix + sum_r tmp_F*B
 *        4  one-column proposal: the same at site S/2 (mcmc_moves.c:213-224)
 *        5  closest-L: hamming_distance_considering_unknown + closest_sequences per query (needs the
 *           sequences section of the problem file)
 * Modes 2-5 time only their own region, per thread (clock around the region); the reported rate is the work
 * of all threads divided by the slowest thread's accumulated region time.
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
extern void update_F_numerical_recipes(int, int, int, double *, char *, char *, double *, double *, int *, int *,
                                       double *, int, int, int *, int, char *);
extern void hamming_distance_considering_unknown(char **, char *, int, int, int *);
extern void closest_sequences(int, int *, int, int *);
extern double kappa, phi, prior_C1[64];
extern const double BIG;

typedef struct {
  int S, N, Q, L, H;
  char *ref_triplets, *ref_codons, *query_codons, *cons_nucls, *hla, *cons_codons, *site_codons, *toc;
  int *closest, *n_site, *ns_ts, *ns_tv;
  double *R, *omega, *esc, *rev, *rate_out, mu;
  int nb;                      /* bases per sequence; 0 = the file carries no sequences */
  char *ref_seqs, *query_seqs; /* [N][nb+1], [Q][nb+1] NUL-terminated ASCII */
  char **ref_rows;
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
  p->nb = 0;
  int nb;
  if (fread(&nb, sizeof(int), 1, f) == 1 && nb > 0) {      /* optional: the sequences, for mode 5 */
    p->nb = nb;
    p->ref_seqs = rd(f, N * (size_t)(nb + 1)); p->query_seqs = rd(f, Q * (size_t)(nb + 1));
    p->ref_rows = malloc(N * sizeof(char *));
    for (size_t r = 0; r < N; r++) p->ref_rows[r] = p->ref_seqs + r * (nb + 1);
  }
  fclose(f);
}

typedef struct {
  const Problem *p;
  int q0, q1, mode, repeats;
  double llk, region_s, units;
} Job;

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *work(void *arg) {
  Job *j = arg;
  const Problem *p = j->p;
  const size_t S = p->S, L = p->L, H = p->H;
  double *c2c = calloc(64 * S, sizeof(double)), *A = calloc(64 * S, sizeof(double));
  double *F = malloc(L * S * sizeof(double)), *B = malloc(L * S * sizeof(double));
  double *T = j->mode >= 3 ? malloc(L * S * sizeof(double)) : NULL;
  int *Fn = malloc(S * sizeof(int)), *Bn = malloc(S * sizeof(int)), *Tn = malloc(S * sizeof(int));
  int *ham = j->mode == 5 ? malloc((size_t)p->N * sizeof(int)) : NULL, *cl = malloc((L + 1) * sizeof(int));
  const int w0 = j->mode == 4 ? p->S / 2 : p->S / 3, w1 = j->mode == 4 ? p->S / 2 : p->S / 3 + p->S / 6 - 1;
  double acc = 0;
  j->region_s = 0; j->units = 0;
  for (int rep = 0; rep < j->repeats; rep++) {
    acc = 0;
    for (int q = j->q0; q < j->q1; q++) {
      if (j->mode == 5) {
        double t0 = now_s();
        hamming_distance_considering_unknown(p->ref_rows, p->query_seqs + (size_t)q * (p->nb + 1), p->N, p->nb, ham);
        closest_sequences(p->N, ham, p->L, cl);
        j->region_s += now_s() - t0;
        j->units += (double)p->N * p->nb;
        acc += cl[0];
        continue;
      }
      char *hrow = p->hla + (size_t)q * (H + 1);
      char *const rows[1] = {hrow};
      int *clq = p->closest + (size_t)q * L;
      char *cons = p->cons_nucls + (size_t)q * 3 * S;
      create_codon_to_codon_HLA(p->S, p->N, 1, p->H, p->n_site, p->site_codons, c2c, p->rate_out, 0, p->S - 1,
                                -1, p->mu, p->esc, p->rev, rows, p->cons_codons, p->omega, p->ns_ts, p->ns_tv,
                                p->toc, p->query_codons + (size_t)q * S, A, 2);
      initialise_F_numerical_recipes(p->S, p->L, p->N, c2c, p->ref_codons, p->ref_triplets, p->R, F, Fn, 0,
                                     p->S - 1, clq, 3 * p->S, cons);
      if (j->mode >= 1) {
        double t0 = now_s();
        update_B_numerical_recipes(p->S, p->L, p->N, c2c, p->ref_codons, p->ref_triplets, p->R, B, Bn, p->S - 1,
                                   clq, 3 * p->S, cons);
        if (j->mode == 2) { j->region_s += now_s() - t0; j->units += (double)S * L; }
      }
      double s = 0;
      if (j->mode >= 3) {
        double t0 = now_s();
        update_F_numerical_recipes(p->S, p->L, p->N, c2c, p->ref_codons, p->ref_triplets, p->R, F, Fn, Tn, T, w0, w1,
                                   clq, 3 * p->S, cons);
        for (size_t r = 0; r < L; r++) s += T[r * S + w1] * B[r * S + w1];      /* mcmc_moves.c:220-223 */
        acc += log(s) - (Tn[w1] + Bn[w1]) * log(BIG);
        j->region_s += now_s() - t0;
        j->units += (double)(w1 - w0 + 1) * L;
        continue;
      }
      for (size_t r = 0; r < L; r++) s += F[r * S + S - 1];
      acc += log(s) - Fn[S - 1] * log(BIG);
    }
  }
  j->llk = acc;
  free(c2c); free(A); free(F); free(B); free(T); free(Fn); free(Bn); free(Tn); free(ham); free(cl);
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: ref_bench problem.bin threads nq mode repeats\n"); return 2; }
  Problem p;
  load(argv[1], &p);
  int T = atoi(argv[2]), nq = atoi(argv[3]), mode = atoi(argv[4]), reps = atoi(argv[5]);
  if (nq > p.Q || nq <= 0) nq = p.Q;
  if (T < 1) T = 1;
  if (T > nq) T = nq;
  if (mode == 5 && p.nb == 0) { fprintf(stderr, "ref_bench: mode 5 needs the sequences in the problem file\n"); return 2; }
  pthread_t *th = malloc(T * sizeof(pthread_t));
  Job *jobs = malloc(T * sizeof(Job));
  double t0 = now_s();
  for (int t = 0; t < T; t++) {
    jobs[t].p = &p; jobs[t].mode = mode; jobs[t].repeats = reps;
    jobs[t].q0 = (int)((long long)nq * t / T); jobs[t].q1 = (int)((long long)nq * (t + 1) / T);
    pthread_create(&th[t], NULL, work, &jobs[t]);
  }
  double llk = 0, region = 0, units = 0;
  for (int t = 0; t < T; t++) {
    pthread_join(th[t], NULL);
    llk += jobs[t].llk; units += jobs[t].units;
    if (jobs[t].region_s > region) region = jobs[t].region_s;
  }
  double sec = now_s() - t0;
  double cells = (double)nq * p.S * p.L * reps;
  if (mode >= 2) { cells = units; }
  double timed = mode >= 2 ? region : sec;
  printf("{\"cells\": %.0f, \"seconds\": %.6f, \"cells_per_s\": %.6e, \"threads\": %d, \"queries\": %d, "
         "\"repeats\": %d, \"mode\": %d, \"wall_seconds\": %.6f, \"llk\": %.10f}\n",
         cells, timed, cells / timed, T, nq, reps, mode, sec, llk);
  return 0;
}
