/* mcq_dmodel.c - the dmodel driver on the B200 library.
 *
 * Same command line, same model, same Metropolis-Hastings moves, same output files as the
 * reference's dmodel (src/tools/dmodel.c, src/mcmc_moves.c); every loop over queries is one call
 * into the batched context of libmcq_b200 (include/mcq_b200.h), where F, B and the emission rows
 * stay resident on the GPU.  The host keeps what is O(sites) or O(windows): the parameter vectors,
 * the selection windows, the priors, the random streams and the accept/reject decision.
 *
 * Random draws are made in the reference's order (closest-L tie shuffles, window initialisation,
 * then per step: move choice, proposal, accept test), so a chain started from the same seeds
 * makes the same decisions (tests/test_mcmc_trace.py).
 *
 * Multi-GPU: one process per GPU (RANK / WORLD_SIZE / LOCAL_RANK from the environment, as set by
 * torchrun or mpirun-style launchers); every rank runs the same host chain on its shard of the
 * queries and the log-likelihood is all-reduced inside the library, so all ranks take the same
 * decisions without any further exchange.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <limits.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/mcq_b200.h"
#include "mcq_io.h"
#include "mcq_json.h"
#include "mcq_model.h"
#include "mcq_prep.h"
#include "mcq_rng.h"
#include "mcq_windows.h"

#define CHECK(call)                                                         \
  do {                                                                      \
    if((call) != 0) mcq_die("%s: %s", #call, mcq_last_error());             \
  } while(0)

/* ------------------------------------------------------------------------------------------
 * chain state kept on the host
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int S, N, Q, L, H;
  double *omega, *R, mu;
  double *esc;               /* [H][S]  HLA_select_esc */
  double *rev;               /* [S]     HLA_select_rev[0] */
  McqHlaWindows *hw;
  double llk;
  double lambda_rec;
  int omega_prior, coeff_prior_esc, coeff_prior_rev;
  bool sample_prior;
  double gamma_const_esc, gamma_const_rev;
  mcq_ctx *ctx;
  /* accept / reject counters, as mcmc_moves.c:13-19 */
  int rec_acc, rec_rej, sel_acc, sel_rej, merge_acc, merge_rej, split_acc, split_rej;
  int grow_acc, grow_rej, coeff_acc, coeff_rej, mut_acc, mut_rej;
  FILE *trace;
} Chain;

/* the accept test of every move: exp(min(0, ratio)) * RAND_MAX > rand() (mcmc_moves.c:249-251) */
static bool accept_test(Chain *c, double ratio)
{
  double alpha = exp(ratio <= 0 ? ratio : 0);
  int draw = rand();
  if(c->trace) {
    fprintf(c->trace, "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", c->rec_acc, c->rec_rej, c->sel_acc,
            c->sel_rej, c->merge_acc, c->merge_rej, c->split_acc, c->split_rej, c->grow_acc, c->grow_rej,
            c->coeff_acc, c->coeff_rej, c->mut_acc, c->mut_rej, draw);
    fflush(c->trace);
  }
  return alpha * RAND_MAX > draw;
}

static double propose_llk(Chain *c, const mcq_proposal *p)
{
  double v = 0;
  CHECK(mcq_ctx_propose(c->ctx, p, &v));
  return v;
}

/* ---- recombination (mcmc_moves.c:172-289) ------------------------------------------------ */
static void move_recombination(Chain *c, double temperature)
{
  int site = mcq_rand_lim(c->S - 1) + 1;
  double old_rate = -c->N * log(1 - c->R[site]);
  double new_rate = old_rate * exp(mcq_rand_lim(4) - 2);
  double new_R = 1 - exp(-new_rate / c->N);
  double llk_ratio = 0, prop = 0;
  if(!c->sample_prior) {
    mcq_proposal p = {MCQ_PROP_REC, site, site, 0, site + 1, new_R, 0};
    prop = propose_llk(c, &p);
    llk_ratio = (1 / temperature) * (prop - c->llk);
  }
  if(isnan(llk_ratio)) mcq_die("recombination_move llk_ratio: NaN");
  double ratio = llk_ratio - c->lambda_rec * (new_rate - old_rate) + log(new_rate / old_rate);
  if(accept_test(c, ratio)) {
    c->rec_acc++;
    c->R[site] = new_R;
    if(!c->sample_prior) { c->llk = prop; CHECK(mcq_ctx_accept(c->ctx)); }
  } else {
    c->rec_rej++;
    if(!c->sample_prior) CHECK(mcq_ctx_reject(c->ctx));
  }
}

/* ---- selection / omega (mcmc_moves.c:291-492) ------------------------------------------- */
static void move_selection(Chain *c, double temperature)
{
  int site = mcq_rand_lim(c->S);
  double old = c->omega[site];
  double new_omega = old * exp(mcq_rand_lim(2) - 1);
  double llk_ratio = 0, prop = 0;
  if(!c->sample_prior) {
    mcq_proposal p = {MCQ_PROP_SEL, site, site, 0, site + 1, new_omega, 0};
    prop = propose_llk(c, &p);
    llk_ratio = (1 / temperature) * (prop - c->llk);
    if(isnan(llk_ratio)) mcq_die("selection_move llk_ratio: NaN");
  }
  double prior_diff;
  switch(c->omega_prior) {
    case MCQ_PRIOR_EXPONENTIAL: prior_diff = -MCQ_LAMBDA_SEL * (new_omega - old); break;
    case MCQ_PRIOR_GAMMA:
      prior_diff = (MCQ_OMEGA_GAMMA_K - 1) * log(new_omega / old) + (old - new_omega) / MCQ_OMEGA_GAMMA_THETA;
      break;
    case MCQ_PRIOR_LOG_NORMAL:
      prior_diff = log(old / new_omega) - (pow(log(new_omega / MCQ_OMEGA_MU_SEL), 2) - pow(log(old / MCQ_OMEGA_MU_SEL), 2)) /
                                              (2 * pow(MCQ_OMEGA_SIGMA_SEL, 2));
      break;
    case MCQ_PRIOR_FLAT: prior_diff = 0; break;
    default: mcq_die("No prior detected!");
  }
  double ratio = llk_ratio + prior_diff + log(new_omega / old);
  if(accept_test(c, ratio)) {
    c->sel_acc++;
    c->omega[site] = new_omega;
    if(!c->sample_prior) { c->llk = prop; CHECK(mcq_ctx_accept(c->ctx)); }
  } else {
    c->sel_rej++;
    if(!c->sample_prior) CHECK(mcq_ctx_reject(c->ctx));
  }
}

/* ---- priors of the window moves (mcmc_moves.c:494-616) ---------------------------------- */
static double prior_diff_merge(int prior, double add, double o0, double o1, double nc, double gconst)
{
  switch(prior) {
    case MCQ_PRIOR_LOG_NORMAL:
      return log(((1 - add) * o0 * o1 * sqrt(2 * M_PI) * MCQ_SIGMA_SEL_PRIOR) / (nc * add)) +
             (-pow(log(nc / MCQ_MU_SEL_PRIOR), 2) + pow(log(o0 / MCQ_MU_SEL_PRIOR), 2) +
              pow(log(o1 / MCQ_MU_SEL_PRIOR), 2)) / (2 * pow(MCQ_SIGMA_SEL_PRIOR, 2));
    case MCQ_PRIOR_NORMAL:
      return log(MCQ_SIGMA_SEL_NORM_PRIOR * sqrt(2 * M_PI)) +
             (pow(o0 - MCQ_MU_SEL_NORM_PRIOR, 2) + pow(o1 - MCQ_MU_SEL_NORM_PRIOR, 2) -
              pow(nc - MCQ_MU_SEL_NORM_PRIOR, 2)) / (2 * pow(MCQ_SIGMA_SEL_NORM_PRIOR, 2)) +
             log((1 - add) / add);
    case MCQ_PRIOR_FLAT: return log((1 - add) / add);
    case MCQ_PRIOR_GAMMA:
      return (MCQ_GAMMA_K_SHAPE - 1) * (log(nc) - log(o0 * o1)) + log((1 - add) / add) +
             (o0 + o1 - nc) / MCQ_GAMMA_THETA_SCALE + gconst;
    default: mcq_die("Unknown prior!");
  }
}

static double prior_diff_split(int prior, double add, double o0, double n0, double n1, double gconst)
{
  switch(prior) {
    case MCQ_PRIOR_LOG_NORMAL:
      return log((add * o0) / ((1 - add) * n0 * n1 * sqrt(2 * M_PI) * MCQ_SIGMA_SEL_PRIOR)) +
             (-pow(log(n0 / MCQ_MU_SEL_PRIOR), 2) - pow(log(n1 / MCQ_MU_SEL_PRIOR), 2) +
              pow(log(o0 / MCQ_MU_SEL_PRIOR), 2)) / (2 * pow(MCQ_SIGMA_SEL_PRIOR, 2));
    case MCQ_PRIOR_NORMAL:
      return -log(MCQ_SIGMA_SEL_NORM_PRIOR * sqrt(2 * M_PI)) +
             (pow(o0 - MCQ_MU_SEL_NORM_PRIOR, 2) - pow(n0 - MCQ_MU_SEL_NORM_PRIOR, 2) +
              pow(n1 - MCQ_MU_SEL_NORM_PRIOR, 2)) / (2 * pow(MCQ_SIGMA_SEL_NORM_PRIOR, 2)) +
             log(add / (1 - add));
    case MCQ_PRIOR_FLAT: return log(add / (1 - add));
    case MCQ_PRIOR_GAMMA:
      return (MCQ_GAMMA_K_SHAPE - 1) * (log(n0 * n1) - log(o0)) + log(add / (1 - add)) -
             (n0 + n1 - o0) / MCQ_GAMMA_THETA_SCALE - gconst;
    default: mcq_die("Unknown prior!");
  }
}

static double prior_diff_coeff(int prior, double o0, double nc)
{
  switch(prior) {
    case MCQ_PRIOR_LOG_NORMAL:
      return log(o0 / nc) - (pow(log(nc / MCQ_MU_SEL_PRIOR), 2) - pow(log(o0 / MCQ_MU_SEL_PRIOR), 2)) /
                                (2 * pow(MCQ_SIGMA_SEL_PRIOR, 2));
    case MCQ_PRIOR_NORMAL:
      return (pow(o0 - MCQ_MU_SEL_NORM_PRIOR, 2) - pow(nc - MCQ_MU_SEL_NORM_PRIOR, 2)) /
             (2 * pow(MCQ_SIGMA_SEL_NORM_PRIOR, 2));
    case MCQ_PRIOR_FLAT: return 0;
    case MCQ_PRIOR_GAMMA: return (MCQ_GAMMA_K_SHAPE - 1) * log(nc / o0) + (o0 - nc) / MCQ_GAMMA_THETA_SCALE;
    default: mcq_die("Unknown prior!");
  }
}

/* ---- window moves: merge/split, grow, coefficient; escape or reversion (mcmc_moves.c:618-1036) */
enum { ACT_MERGE_OR_SPLIT = 0, ACT_MERGE = 1, ACT_SPLIT = 2, ACT_GROW = 3, ACT_COEFF = 4 };

static void move_window(Chain *c, double temperature, bool is_rev, int action)
{
  const int S = c->S;
  int h = mcq_rand_lim(c->H);                     /* drawn for reversion moves too (mcmc_moves.c:657) */
  McqWindowSet *ws = is_rev ? c->hw[0].rev : c->hw[h].esc;
  double *per_site = is_rev ? c->rev : c->esc + (size_t)h * S;
  const double add = is_rev ? MCQ_PRIOR_ADD_WINDOW_REV : MCQ_PRIOR_ADD_WINDOW_ESC;
  const int prior = is_rev ? c->coeff_prior_rev : c->coeff_prior_esc;
  const double gconst = is_rev ? c->gamma_const_rev : c->gamma_const_esc;

  double prob_split = 1, prob_merge = 1;
  if(action == ACT_MERGE_OR_SPLIT) {
    int joins = ws->count - 1;
    double a = (double)(S - ws->count) / (double)ws->count * add / (1 - add);
    prob_split = a <= 1 ? a : 1;
    double b = (double)joins / (double)(S - joins) * (1 - add) / add;
    prob_merge = b <= 1 ? b : 1;
    double u = mcq_rand_lim(1);
    action = (u < prob_split / (prob_split + prob_merge)) ? ACT_SPLIT : ACT_MERGE;
  }

  int w, old_start = 0;
  double old0 = 0, old1 = 0;
  switch(action) {
    case ACT_MERGE:
      w = mcq_windows_merge(ws, &old_start, &old0, &old1);
      if(w < 0) mcq_die("This should not happen");
      break;
    case ACT_SPLIT:
      w = mcq_windows_split(ws, S, &old0);
      if(w < 0) mcq_die("This should not happen");
      break;
    case ACT_GROW:
      w = mcq_windows_grow(ws, MCQ_WINDOW_GEOM_PARAM, &old_start);
      if(w == -1) { c->grow_rej++; return; }      /* no accept draw in this case (mcmc_moves.c:725-728) */
      break;
    case ACT_COEFF: w = mcq_windows_coeff(ws, &old0); break;
    default: mcq_die("No action!");
  }

  /* the sites whose coefficient changes, as one or two constant runs (mcmc_moves.c:736-770) */
  int start, end, split;
  double v0, v1 = 0;
  if(action != ACT_GROW) {
    start = ws->w[w].start;
    end = start + ws->w[w].length - 1;
    split = end + 1;
    v0 = ws->w[w].coeff;
    if(action == ACT_SPLIT) { end += ws->w[w+1].length; v1 = ws->w[w+1].coeff; }
  } else {
    int new_start = ws->w[w+1].start;
    if(old_start < new_start) { start = old_start; end = new_start - 1; v0 = ws->w[w].coeff; }
    else { start = new_start; end = old_start - 1; v0 = ws->w[w+1].coeff; }
    split = end + 1;
  }
  for(int s = start; s <= end; s++) per_site[s] = s < split ? v0 : v1;

  double llk_ratio = 0, prop = 0;
  if(!c->sample_prior) {
    mcq_proposal p = {is_rev ? MCQ_PROP_REV : MCQ_PROP_ESC, start, end, h, split, v0, v1};
    prop = propose_llk(c, &p);
    llk_ratio = (1 / temperature) * (prop - c->llk);
    if(isnan(llk_ratio)) mcq_die("window_move llk_ratio: NaN");
  }

  double ratio;
  switch(action) {
    case ACT_MERGE: {
      double nc = ws->w[w].coeff;
      double odds = add / (1 - add);
      double birth = ((S - (double)ws->count) / (double)ws->count) * odds;
      double death = 1 / birth <= 1 ? 1 / birth : 1;
      birth = birth <= 1 ? birth : 1;
      ratio = llk_ratio + prior_diff_merge(prior, add, old0, old1, nc, gconst) +
              log((birth / death) * (ws->count * nc) / ((S - ws->count) * pow(old0 + old1, 2)));
      break;
    }
    case ACT_SPLIT: {
      double n0 = ws->w[w].coeff, n1 = ws->w[w+1].coeff, sum = n0 + n1;
      double odds = (1 - add) / add;
      double death = (double)(ws->count - 1) / (double)(S - (ws->count - 1)) * odds;
      double birth = 1 / death <= 1 ? 1 / death : 1;
      death = death <= 1 ? death : 1;
      ratio = llk_ratio + prior_diff_split(prior, add, old0, n0, n1, gconst) +
              log((death / birth) * ((S - (ws->count - 1)) * pow(sum, 2)) / ((ws->count - 1) * old0));
      break;
    }
    case ACT_GROW: ratio = llk_ratio; break;
    default: {
      double nc = ws->w[w].coeff;
      ratio = llk_ratio + prior_diff_coeff(prior, old0, nc) + log(nc / old0);
    }
  }

  if(accept_test(c, ratio)) {
    switch(action) {
      case ACT_MERGE: c->merge_acc++; break;
      case ACT_SPLIT: c->split_acc++; break;
      case ACT_GROW:
        c->grow_acc++;
        mcq_windows_grow_accept(ws, ws->w[w+1].start, old_start, S);
        break;
      default: c->coeff_acc++;
    }
    if(!c->sample_prior) { c->llk = prop; CHECK(mcq_ctx_accept(c->ctx)); }
  } else {
    switch(action) {
      case ACT_MERGE: c->merge_rej++; mcq_windows_merge_undo(ws, w, old_start, old0, old1); break;
      case ACT_SPLIT: c->split_rej++; mcq_windows_split_undo(ws, w, old0); break;
      case ACT_GROW: c->grow_rej++; mcq_windows_grow_undo(ws, w, old_start); break;
      default: c->coeff_rej++; mcq_windows_coeff_undo(ws, w, old0);
    }
    if(!c->sample_prior) CHECK(mcq_ctx_reject(c->ctx));
    /* the per-site coefficients go back to what the restored windows say (mcmc_moves.c:1018-1034) */
    for(int k = w; k <= w + ((action == ACT_MERGE || action == ACT_GROW) ? 1 : 0); k++)
      for(int s = ws->w[k].start; s < ws->w[k].start + ws->w[k].length; s++) per_site[s] = ws->w[k].coeff;
  }
}

/* ---- mutation rate (mcmc_moves.c:1038-1188) ---------------------------------------------- */
static void move_mutation(Chain *c, double temperature)
{
  double new_mu = c->mu * exp((mcq_rand_lim(1) - 0.5) / 4);
  double llk_ratio = 0, prop = 0;
  if(!c->sample_prior) {
    mcq_proposal p = {MCQ_PROP_MU, 0, c->S - 1, 0, c->S, new_mu, 0};
    prop = propose_llk(c, &p);
    llk_ratio = (1 / temperature) * (prop - c->llk);
    if(isnan(llk_ratio)) mcq_die("mutation_move llk_ratio: NaN");
  }
  double ratio = llk_ratio - MCQ_MU_PRIOR * (new_mu - c->mu) + log(new_mu / c->mu);
  if(accept_test(c, ratio)) {
    c->mut_acc++;
    c->mu = new_mu;
    if(!c->sample_prior) { CHECK(mcq_ctx_accept(c->ctx)); c->llk = prop; }
  } else {
    c->mut_rej++;
    if(!c->sample_prior) CHECK(mcq_ctx_reject(c->ctx));
  }
}

/* ------------------------------------------------------------------------------------------
 * command line (the reference's option table, dmodel.c:108-137, plus --seed/--gsl_seed/--device/
 * --sparse_esc/--trace which the reference lacks)
 * ---------------------------------------------------------------------------------------- */
static int chain_length = 50000, sample_rate = 100, closest_n = 100, runtime_hours = 0;
static const char *ref_path, *query_path, *out_dir, *hla_path, *cons_path;
static bool separate_reference_set = true, no_hla = false, no_rev = false, no_rec = false, no_anneal = false;
static int opt_sparse = -1;   /* -1 auto, 0 dense, 1 sparse */
static int opt_checkpoint = -1;   /* -1 auto, else the stride (1 = whole matrices) */
static bool opt_sample_prior = false, runtime_set = false, states_set = false, opt_nccl = false,
            opt_eager = false, opt_profile = false;
static int opt_omega_prior = MCQ_PRIOR_EXPONENTIAL, opt_esc_prior = MCQ_PRIOR_LOG_NORMAL;
static int opt_rev_prior = MCQ_PRIOR_LOG_NORMAL, opt_device = -1;
static long opt_seed = -1, opt_gsl_seed = -1;
static const char *trace_path = NULL, *resume_path = NULL, *params_path = NULL, *mosaic_path = NULL;

static void usage(void)
{
  fprintf(stderr,
          "usage: mcq_dmodel [options] <ref_seqs.fasta> <query.fasta> <outdir>\n"
          "  -H <HLA.csv>  -n <chain length>  -s <sample every>  -t <hours>  -c <consensus.fasta>  -l <closest L>\n"
          "  -r <state.json> resume from a saved state  -p <parameters.json> kappa, phi, codon priors\n"
          "  -m <mosaic.json> closest-L lists from the true mosaics of a simulation\n"
          "  -q self-reference mode  -i no HLA inference  -j no reversion inference  -o no recombination\n"
          "  -v/-e/-w omega prior gamma/log-normal/uniform   -g/-k/-f escape prior gamma/normal/flat\n"
          "  -a/-d/-u reversion prior gamma/normal/flat   -b sample the prior   -z no simulated annealing\n"
          "  --seed N --gsl_seed N (default: clock)  --device D  --sparse_esc | --dense_esc  --trace FILE  --nccl\n"
          "  --eager_refill  refill F and B at every accept (the reference's schedule) instead of when next read\n"
          "  --profile  wall time per kind of move at the end of the run\n"
          "  --checkpoint K  keep every K-th column of F and B (1: all of them, as the reference; default: by shard size)\n");
  exit(EXIT_FAILURE);
}

static void parse(int argc, char **argv)
{
  static struct option lo[] = {
      {"HLA_inference", required_argument, NULL, 'H'}, {"chain_length", required_argument, NULL, 'n'},
      {"sample_every", required_argument, NULL, 's'}, {"runtime_hours", required_argument, NULL, 't'},
      {"consensus_fasta", required_argument, NULL, 'c'}, {"closest_n", required_argument, NULL, 'l'},
      {"resume_json", required_argument, NULL, 'r'}, {"parameters_json", required_argument, NULL, 'p'},
      {"true_mosaic_fasta", required_argument, NULL, 'm'},
      {"separate_reference_fasta", no_argument, NULL, 'q'}, {"no_HLA_inference", no_argument, NULL, 'i'},
      {"no_rev_inference", no_argument, NULL, 'j'}, {"no_recombination_inference", no_argument, NULL, 'o'},
      {"omega_gamma", no_argument, NULL, 'v'}, {"omega_log_normal", no_argument, NULL, 'e'},
      {"omega_uniform", no_argument, NULL, 'w'}, {"HLA_coeff_esc_gamma", no_argument, NULL, 'g'},
      {"HLA_coeff_esc_normal", no_argument, NULL, 'k'}, {"HLA_coeff_esc_uniform", no_argument, NULL, 'f'},
      {"HLA_coeff_rev_gamma", no_argument, NULL, 'a'}, {"HLA_coeff_rev_normal", no_argument, NULL, 'd'},
      {"HLA_coeff_rev_uniform", no_argument, NULL, 'u'}, {"sample_prior", no_argument, NULL, 'b'},
      {"no_simulated_annealing", no_argument, NULL, 'z'}, {"seed", required_argument, NULL, 1001},
      {"gsl_seed", required_argument, NULL, 1002}, {"device", required_argument, NULL, 1003},
      {"sparse_esc", no_argument, NULL, 1004}, {"trace", required_argument, NULL, 1005},
      {"nccl", no_argument, NULL, 1006}, {"eager_refill", no_argument, NULL, 1007}, {"dense_esc", no_argument, NULL, 1008},
      {"profile", no_argument, NULL, 1009}, {"checkpoint", required_argument, NULL, 1010},
      {NULL, 0, NULL, 0}};
  int o;
  while((o = getopt_long(argc, argv, "H:n:s:t:c:r:p:m:l:qijovewgkfadubz", lo, NULL)) != -1) {
    switch(o) {
      case 'H': hla_path = optarg; break;
      case 'n': chain_length = atoi(optarg); states_set = true; break;
      case 's': sample_rate = atoi(optarg); break;
      case 't': runtime_hours = atoi(optarg); runtime_set = true; break;
      case 'c': cons_path = optarg; break;
      case 'r': resume_path = optarg; break;
      case 'p': params_path = optarg; break;
      case 'm': mosaic_path = optarg; break;
      case 'l': closest_n = atoi(optarg); break;
      case 'q': separate_reference_set = false; break;
      case 'i': no_hla = true; break;
      case 'j': no_rev = true; break;
      case 'o': no_rec = true; break;
      case 'v': opt_omega_prior = MCQ_PRIOR_GAMMA; break;
      case 'e': opt_omega_prior = MCQ_PRIOR_LOG_NORMAL; break;
      case 'w': opt_omega_prior = MCQ_PRIOR_FLAT; break;
      case 'g': opt_esc_prior = MCQ_PRIOR_GAMMA; break;
      case 'k': opt_esc_prior = MCQ_PRIOR_NORMAL; break;
      case 'f': opt_esc_prior = MCQ_PRIOR_FLAT; break;
      case 'a': opt_rev_prior = MCQ_PRIOR_GAMMA; break;
      case 'd': opt_rev_prior = MCQ_PRIOR_NORMAL; break;
      case 'u': opt_rev_prior = MCQ_PRIOR_FLAT; break;
      case 'b': opt_sample_prior = true; break;
      case 'z': no_anneal = true; break;
      case 1001: opt_seed = atol(optarg); break;
      case 1002: opt_gsl_seed = atol(optarg); break;
      case 1003: opt_device = atoi(optarg); break;
      case 1004: opt_sparse = 1; break;
      case 1008: opt_sparse = 0; break;
      case 1005: trace_path = optarg; break;
      case 1006: opt_nccl = true; break;
      case 1007: opt_eager = true; break;
      case 1009: opt_profile = true; break;
      case 1010: opt_checkpoint = atoi(optarg); break;
      default: usage();
    }
  }
  if(argc - optind != 3) usage();
  if(runtime_set && states_set) mcq_die("Set both a time limit and number of states");
  ref_path = argv[optind]; query_path = argv[optind + 1]; out_dir = argv[optind + 2];
  if(cons_path == NULL) cons_path = ref_path;
}

static int env_int(const char *name, int dflt)
{
  const char *v = getenv(name);
  return v ? atoi(v) : dflt;
}

/* Hand-over of small blobs between the ranks of one box through files under /tmp: rank r writes
 * <tag>_<r>, everyone polls for the files it needs.  Used for the 64-byte CUDA IPC handles of the peer
 * mailboxes (default) or the 128-byte NCCL id (--nccl). */
static void blob_path(char *path, size_t n, const char *tag, int r)
{
  /* the launcher's pid is common to the ranks of one launch and differs between launches, so files a
   * crashed earlier run left behind are never picked up */
  const char *run = getenv("MCQ_RUN_TAG");      /* ranks started by different parents name their launch themselves */
  char ppid[32];
  snprintf(ppid, sizeof(ppid), "%ld", (long)getppid());
  snprintf(path, n, "/tmp/mcq_%s_%s_%s_%s_%d", tag, getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0",
           getenv("TORCHELASTIC_RUN_ID") ? getenv("TORCHELASTIC_RUN_ID") : "run", run ? run : ppid, r);
}

static void blob_put(const char *tag, int r, const unsigned char *data, size_t len)
{
  char path[PATH_MAX], tmp[PATH_MAX + 8];
  blob_path(path, sizeof(path), tag, r);
  snprintf(tmp, sizeof(tmp), "%s.tmp", path);
  FILE *f = fopen(tmp, "wb");
  if(!f || fwrite(data, 1, len, f) != len) mcq_die("cannot write %s", tmp);
  fclose(f);
  rename(tmp, path);
}

static void blob_get(const char *tag, int r, unsigned char *data, size_t len)
{
  char path[PATH_MAX];
  blob_path(path, sizeof(path), tag, r);
  for(int tries = 0; tries < 12000; tries++) {
    FILE *f = fopen(path, "rb");
    if(f) {
      size_t n = fread(data, 1, len, f);
      fclose(f);
      if(n == len) return;
    }
    usleep(10000);
  }
  mcq_die("timed out waiting for %s", path);
}

/* a blob of any length: rank r writes {length, bytes}, the reader allocates */
static void blob_put_var(const char *tag, int r, const void *data, size_t len)
{
  unsigned char *buf = malloc(sizeof(size_t) + len);
  memcpy(buf, &len, sizeof(size_t));
  memcpy(buf + sizeof(size_t), data, len);
  blob_put(tag, r, buf, sizeof(size_t) + len);
  free(buf);
}

static void *blob_get_var(const char *tag, int r, size_t *len)
{
  char path[PATH_MAX];
  blob_path(path, sizeof(path), tag, r);
  for(int tries = 0; tries < 12000; tries++) {
    FILE *f = fopen(path, "rb");           /* (the file appears complete: written aside and renamed) */
    if(f) {
      void *data = NULL;
      if(fread(len, sizeof(size_t), 1, f) == 1 && (data = malloc(*len ? *len : 1)) != NULL &&
         fread(data, 1, *len, f) == *len) {
        fclose(f);
        return data;
      }
      free(data);
      fclose(f);
    }
    usleep(10000);
  }
  mcq_die("timed out waiting for %s", path);
}

int main(int argc, char **argv)
{
  parse(argc, argv);
  const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1);
  const int device = opt_device >= 0 ? opt_device : env_int("LOCAL_RANK", 0);

  /* the two random streams (dmodel.c:391-392); --seed / MCQ_SEED make a run reproducible */
  if(opt_seed < 0 && getenv("MCQ_SEED")) opt_seed = atol(getenv("MCQ_SEED"));
  if(opt_gsl_seed < 0 && getenv("MCQ_GSL_SEED")) opt_gsl_seed = atol(getenv("MCQ_GSL_SEED"));
  {
    /* clock-derived seeds differ between processes: with several ranks rank 0 draws them and hands them
     * over before anybody makes a draw, so that every rank runs the same chain */
    unsigned long seeds[2];
    seeds[0] = opt_seed >= 0 ? (unsigned long)opt_seed : (unsigned long)(unsigned)time(NULL);
    seeds[1] = opt_gsl_seed >= 0 ? (unsigned long)opt_gsl_seed : (unsigned long)((double)time(NULL) * getpid());
    if(world > 1) {
      if(rank == 0) blob_put("seed", 0, (const unsigned char *)seeds, sizeof(seeds));
      else blob_get("seed", 0, (unsigned char *)seeds, sizeof(seeds));
    }
    mcq_seed_libc((unsigned)seeds[0]);
    mcq_mt_seed(seeds[1]);
    if(rank == 0) printf("seeds: %lu %lu\n", seeds[0], seeds[1]);
  }

  if(!mcq_mkpath(out_dir)) mcq_die("Cannot create output dir: %s", out_dir);
  char stamp[64], path[PATH_MAX];
  {
    time_t now = time(NULL);
    struct tm *tm = localtime(&now);
    snprintf(stamp, sizeof(stamp), "%d_%d_%d_%d_%d_%d", tm->tm_year + 1900, tm->tm_mon + 1, tm->tm_mday,
             tm->tm_hour, tm->tm_min, tm->tm_sec);
    if(world > 1) snprintf(stamp + strlen(stamp), sizeof(stamp) - strlen(stamp), "_rank%d", rank);
  }
  FILE *log_file, *window_file = NULL, *priors_file, *summary_file, *json_file;
#define OPEN(var, suffix)                                                         \
  do {                                                                            \
    snprintf(path, sizeof(path), "%s/%s_%s", out_dir, stamp, suffix);             \
    var = fopen(path, "w");                                                       \
    if(!var) mcq_die("Cannot open output file: %s", path);                        \
  } while(0)
  OPEN(log_file, "log.txt");
  if(hla_path != NULL || !no_rev) OPEN(window_file, "window.txt");
  OPEN(priors_file, "priors.txt");
  OPEN(summary_file, "summary.txt");
  OPEN(json_file, "final_state.json");
  FILE *trace = NULL;
  if(trace_path == NULL) trace_path = getenv("MCQ_TRACE");
  if(trace_path != NULL && rank == 0) trace = fopen(trace_path, "w");

  printf("Loading sequences...\n");
  char **ref_seqs, **query_seqs, **cons_seqs;
  int N = mcq_load_fasta(ref_path, &ref_seqs);
  int Qall = mcq_load_fasta(query_path, &query_seqs);
  int ncons = mcq_load_fasta(cons_path, &cons_seqs);
  if(N <= 0 || Qall <= 0 || ncons <= 0) mcq_die("empty sequence file");
  printf("Loaded %i reference sequences.\nLoaded %i query sequences.\nLoaded %i sequences to determine consensus.\n",
         N, Qall, ncons);
  size_t nb = strlen(ref_seqs[0]);
  if(nb % 3 != 0) mcq_die("Sequences contain partial codons [%zu mod 3 != 0].", nb);
  for(int r = 1; r < N; r++) if(strlen(ref_seqs[r]) != nb) mcq_die("Reference sequences aren't all the same length.");
  for(int q = 0; q < Qall; q++) if(strlen(query_seqs[q]) != nb) mcq_die("Query sequences aren't all the same length.");
  for(int i = 0; i < ncons; i++)
    if(strlen(cons_seqs[i]) != nb) mcq_die("Sequences from which to derive the consensus sequence aren't all the same length.");
  const int S = (int)nb / 3;

  char **hla_rows;
  int H = 0;
  if(hla_path != NULL) {
    printf("Loading HLA types...\n");
    H = mcq_load_hla_csv(hla_path, &hla_rows, Qall);
    if(H <= 0) mcq_die("no HLA types in %s", hla_path);
    printf("Loaded %i HLA types.\n", H);
  } else {
    /* no HLA file: one all-'0' pseudo type so that the reversion coefficients still apply (dmodel.c:449-458,581-584) */
    hla_rows = malloc(sizeof(char *) * Qall);
    for(int q = 0; q < Qall; q++) { hla_rows[q] = malloc(2); hla_rows[q][0] = '0'; hla_rows[q][1] = '\0'; }
    H = 1;
  }
  /* the fixed viral parameters; dmodel -p replaces them (dmodel.c:461-463) */
  double kappa = MCQ_KAPPA, phi = MCQ_PHI, prior[64];
  memcpy(prior, mcq_prior_c1, sizeof(prior));
  if(params_path != NULL) mcq_load_fixed_parameters_json(params_path, &kappa, &phi, prior);
  printf("kappa:%f\nphi:%f\n", kappa, phi);
  double lambda_rec = 1 / (-log(1 - MCQ_P_REC) * N);
  printf("lambda_rec:%f\n", lambda_rec);
  printf("Loaded %i sites.\n", S);

  /* this rank's shard of the queries */
  const int q0 = (int)((long long)Qall * rank / world), q1 = (int)((long long)Qall * (rank + 1) / world);
  const int Q = q1 - q0;

  /* closest-L.  The distances and the selection groups (the references closer than the boundary distance, those
   * at it) are computed for this rank's queries only; the tie shuffles consume the shared random() stream in
   * query order (closest_seqs.c:57-64), so the groups of all ranks are exchanged and every rank replays all of
   * them: the lists are those of a single process. */
  const bool skip_first = !separate_reference_set;
  if((skip_first ? closest_n + 1 : closest_n) > N) mcq_die("closest_n exceeds the number of references");
  int *closest_all = malloc((size_t)Qall * closest_n * sizeof(int));
  {
    char *rflat = malloc((size_t)N * nb), *qflat = malloc((size_t)(Q ? Q : 1) * nb);
    for(int r = 0; r < N; r++) memcpy(rflat + (size_t)r * nb, ref_seqs[r], nb);
    for(int q = 0; q < Q; q++) memcpy(qflat + (size_t)q * nb, query_seqs[q0 + q], nb);
    int *groups = NULL;
    size_t count = 0;
    if(Q == 0) mcq_die("more ranks than query sequences");
    CHECK(mcq_closest_groups(device, rflat, nb, N, qflat, nb, Q, (int)nb, closest_n, skip_first, &groups, &count));
    free(rflat); free(qflat);
    if(world > 1) blob_put_var("closest", rank, groups, count * sizeof(int));
    for(int r = 0; r < world; r++) {
      const int r0 = (int)((long long)Qall * r / world);
      int *g = groups, nq = 0;
      size_t cnt = count;
      if(r != rank) {
        size_t len = 0;
        g = blob_get_var("closest", r, &len);
        cnt = len / sizeof(int);
      }
      CHECK(mcq_closest_replay(g, cnt, closest_n, skip_first, closest_all + (size_t)r0 * closest_n, &nq));
      if(nq != (int)((long long)Qall * (r + 1) / world) - r0) mcq_die("rank %d selected for %d queries", r, nq);
      if(r != rank) free(g);
    }
    mcq_free(groups);
  }
  printf("Example - closest references to query sequence 1:\n");
  for(int r = 0; r < closest_n; r++) printf("%i ", closest_all[r]);
  printf("\n");

  if(mosaic_path != NULL) {
    /* dmodel -m (dmodel.c:505-558): every query's list becomes the closest relatives of the references it was
     * truly copied from, floor(L/k) of each and one more for the last L mod k of them.  Draws (the permutation,
     * then the tie shuffles of every selection) are made in the reference's order, one selection at a time
     * through the library's one-query symbols. */
    int *which = malloc(sizeof(int) * (size_t)Qall * S), *num = malloc(sizeof(int) * Qall);
    int *tmp_hamming = malloc(sizeof(int) * N), *sel = malloc(sizeof(int) * (closest_n + 1));
    mcq_load_true_mosaics_json(mosaic_path, Qall, S, which, num);
    for(int q = 0; q < Qall; q++) {
      int *w = which + (size_t)q * S, k = num[q];
      if(k <= 0) mcq_die("mosaic of query %i is empty", q);
      for(int i = 0; i < k; i++) {                /* knuth_perm, util.c:103-112 */
        int j = (int)mcq_rand_lim(i + 1), t = w[i];
        w[i] = w[j]; w[j] = t;
      }
      int each = closest_n / k, remain = closest_n % k, m_total = 0;
      for(int r = 0; r < k; r++) {
        int n = r < k - remain ? each : each + 1;
        if(w[r] < 0 || w[r] >= N) mcq_die("mosaic of query %i names reference %i", q, w[r]);
        hamming_distance_considering_unknown(ref_seqs, ref_seqs[w[r]], N, (int)nb, tmp_hamming);
        closest_sequences(N, tmp_hamming, n, sel);
        for(int m = 0; m < n; m++) closest_all[(size_t)q * closest_n + m_total++] = sel[m];
      }
    }
    free(which); free(num); free(tmp_hamming); free(sel);
  }

  /* parameters and windows (dmodel.c:564-642) */
  Chain c;
  memset(&c, 0, sizeof(c));
  c.S = S; c.N = N; c.L = closest_n; c.H = H; c.trace = trace;
  c.omega = malloc(sizeof(double) * S); c.R = malloc(sizeof(double) * S);
  for(int s = 0; s < S; s++) { c.omega[s] = 1.0; c.R[s] = no_rec ? 0 : MCQ_INIT_R; }
  c.mu = MCQ_INIT_MU;
  c.lambda_rec = lambda_rec;
  c.omega_prior = opt_omega_prior; c.coeff_prior_esc = opt_esc_prior; c.coeff_prior_rev = opt_rev_prior;
  c.sample_prior = opt_sample_prior;
  c.gamma_const_esc = MCQ_GAMMA_K_SHAPE * log(MCQ_GAMMA_THETA_SCALE) + lgamma(MCQ_GAMMA_K_SHAPE);
  c.gamma_const_rev = c.gamma_const_esc;
  c.hw = malloc(sizeof(McqHlaWindows) * H);
  mcq_windows_alloc(S, H, MCQ_INIT_WINDOWS, MCQ_INIT_WINDOWS, c.hw);
  double resume_llk = 0;
  if(resume_path != NULL) {
    /* dmodel -r (dmodel.c:589-604): parameters, windows and the closest-L lists come from the saved
     * state; the tie shuffles above have still consumed their random() draws, as in the reference */
    mcq_load_state_json(resume_path, &resume_llk, &c.mu, c.R, c.omega, c.hw, S, H, Qall, closest_n, closest_all);
  } else {
    for(int h = 0; h < H; h++) {
      mcq_windows_randomise(c.hw[h].esc, S, MCQ_INIT_WINDOWS, log(MCQ_MU_SEL_PRIOR), MCQ_SIGMA_SEL_PRIOR);
      mcq_windows_randomise(c.hw[h].rev, S, MCQ_INIT_WINDOWS, log(MCQ_MU_SEL_PRIOR), MCQ_SIGMA_SEL_PRIOR);
    }
    if(hla_path == NULL || no_hla)
      for(int h = 0; h < H; h++) for(int i = 0; i < MCQ_INIT_WINDOWS; i++) c.hw[h].esc->w[i].coeff = 1;
    if(no_rev) for(int i = 0; i < MCQ_INIT_WINDOWS; i++) c.hw[0].rev->w[i].coeff = 1;
    for(int h = 1; h < H; h++) for(int i = 0; i < MCQ_INIT_WINDOWS; i++) c.hw[h].rev->w[i].coeff = 1;
  }

  c.Q = Q;
  McqData d;
  mcq_prepare(&d, device, ref_seqs, N, query_seqs + q0, Q, cons_seqs, ncons, hla_rows + q0, H,
              closest_all + (size_t)q0 * closest_n, closest_n);

  printf("Reference codons: ");
  for(int s = 0; s < S; s++) {
    if(s > 0) printf(" | ");
    for(int i = 0; i < d.n_site_codons[s]; i++) printf(i ? ",%i" : "%i", (int)d.site_codons[(size_t)s * 64 + i]);
  }
  printf("\n");
  fprintf(summary_file, "chain length: %i\nsample every: %i\nreference sequence file: %s\nquery sequence file: %s%s%s\n"
                        "consensus sequence file: %s\n", chain_length, sample_rate, ref_path, query_path,
          hla_path ? "\nHLA csv file: " : "", hla_path ? hla_path : "", cons_path);
  fclose(summary_file);

  /* per-site coefficient arrays from the windows (dmodel.c:852-862) */
  c.esc = malloc(sizeof(double) * H * S);
  c.rev = malloc(sizeof(double) * S);
  double *rev_all = malloc(sizeof(double) * S);
  for(int h = 0; h < H; h++) {
    mcq_windows_to_sites(c.hw[h].esc, c.esc + (size_t)h * S);
    mcq_windows_to_sites(c.hw[h].rev, h == 0 ? c.rev : rev_all);
  }
  free(rev_all);

  double acc = 0;
  for(int i = 0; i < 63; i++) acc += prior[i];
  prior[63] = 1 - acc;                                     /* dmodel.c:867-869 */
  double *rate_out = malloc(sizeof(double) * 64 * S);     /* dmodel.c:836-849 */
  for(int c1 = 0; c1 < 64; c1++) {
    double v1 = kappa * d.beta_S[c1] + d.beta_V[c1], v2 = kappa * d.alpha_S[c1] + d.alpha_V[c1];
    for(int s = 0; s < S; s++) rate_out[(size_t)c1 * S + s] = v1 + c.omega[s] * v2;
  }

  /* the device context */
  mcq_dims dims = {S, closest_n, N, Q, H, device};
  CHECK(mcq_ctx_create(&c.ctx, &dims));
  /* Only every fourth column of F and B is kept, the ones in between are recomputed when a proposal needs them:
   * a sweep moves 3 bytes per cell instead of 9 and the matrices take a quarter of the memory.  Measured on B200:
   * 1.6 x the steps per second at 10 000 queries x L = 500 on one GPU, 1.15 x at 1 250 queries, 1.07 x at the
   * 1 000 queries x L = 100 of config 2; a stride of 8 adds another 4 % at most.  --checkpoint 1 keeps whole
   * matrices, as the reference does. */
  if(opt_checkpoint < 0) opt_checkpoint = S >= 64 ? 4 : 1;
  if(opt_checkpoint > 1) CHECK(mcq_ctx_set_option(c.ctx, MCQ_OPT_CHECKPOINT_STRIDE, opt_checkpoint));
  printf("forward / backward matrices: %s\n", opt_checkpoint > 1 ? "every k-th column kept (--checkpoint 1 for all of them)" : "whole");
  CHECK(mcq_ctx_upload_static(c.ctx, d.ref_triplets, closest_all + (size_t)q0 * closest_n, d.cons_query_nucls,
                              d.query_codons, d.hla, d.consensus_codons, d.n_site_codons, d.site_codons));
  mcq_params prm = {c.mu, kappa, phi, prior, c.R, c.omega, rate_out, c.esc, c.rev};
  CHECK(mcq_ctx_set_params(c.ctx, &prm));
  /* An escape-window proposal changes the emissions of the carriers of one HLA type only.  Re-evaluating just
   * those queries pays when the evaluation is bound by bytes (large shards of wide queries) and costs when it
   * is bound by the chain of dependent sites (small shards: the catch-ups of differently stale queries no
   * longer coincide); by default the shard's size decides (the same on every rank). */
  if(opt_sparse < 0) opt_sparse = (double)Qall / world * ((closest_n + 15) / 16 * 16) >= 262144.0 ? 1 : 0;
  if(opt_sparse) CHECK(mcq_ctx_set_option(c.ctx, MCQ_OPT_SPARSE_ESC, 1));
  printf("escape-window proposals: %s\n", opt_sparse ? "carriers only (--dense_esc for all queries)" : "all queries (--sparse_esc for carriers only)");
  if(opt_eager) CHECK(mcq_ctx_set_option(c.ctx, MCQ_OPT_LAZY_REFILL, 0));
  if(world > 1 && !opt_nccl) {
    /* shard totals are exchanged by NVLink peer stores inside the reduction kernel */
    unsigned char *handles = malloc((size_t)64 * world);
    CHECK(mcq_ctx_peer_handle(c.ctx, handles + (size_t)64 * rank));
    blob_put("peer", rank, handles + (size_t)64 * rank, 64);
    for(int r = 0; r < world; r++) if(r != rank) blob_get("peer", r, handles + (size_t)64 * r, 64);
    CHECK(mcq_ctx_attach_peers(c.ctx, world, rank, handles));
    free(handles);
  } else if(world > 1) {
    unsigned char id[128];
    if(rank == 0) { CHECK(mcq_comm_unique_id(id)); blob_put("nccl", 0, id, 128); }
    else blob_get("nccl", 0, id, 128);
    CHECK(mcq_ctx_attach_comm(c.ctx, world, rank, id));
  }
  if(!c.sample_prior) {
    CHECK(mcq_ctx_init_state(c.ctx, &c.llk));
    printf("Total log-likelihood using initialise F: %f.\n", c.llk);
    CHECK(mcq_ctx_tune(c.ctx, NULL));   /* residency of the sweep kernels for this shard size */
    if(world > 1) {
      /* the first exchange has completed on every rank, so every rank has read every hand-over file */
      char mine[PATH_MAX];
      blob_path(mine, sizeof(mine), opt_nccl ? "nccl" : "peer", opt_nccl ? 0 : rank);
      if(!opt_nccl || rank == 0) unlink(mine);
    }
  }
  if(world > 1) {
    /* the seed hand-over file: every rank has read it once the first exchange is through */
    double one = 1;
    CHECK(mcq_ctx_exchange_sum(c.ctx, &one));
    if((int)one != world) mcq_die("rank exchange returned %f for %d ranks", one, world);
    if(rank == 0) { char f[PATH_MAX]; blob_path(f, sizeof(f), "seed", 0); unlink(f); }
    { char f[PATH_MAX]; blob_path(f, sizeof(f), "closest", rank); unlink(f); }
  }
  if(resume_path != NULL) printf("Likelihood from loaded .json file: %f.\n", resume_llk);

  /* move weights (dmodel.c:915-962) */
  const bool esc_on = (hla_path != NULL) && !no_hla, rev_on = !no_rev;
  const int k = 5;
  int rec = no_rec ? 0 : 20;
  int sel = 20 + rec;
  int esc_ms = (esc_on ? H * k : 0) + sel;
  int rev_ms = (rev_on ? k : 0) + esc_ms;
  int esc_gr = (esc_on ? H * k : 0) + rev_ms;
  int rev_gr = (rev_on ? k : 0) + esc_gr;
  int esc_co = (esc_on ? H * k : 0) + rev_gr;
  int rev_co = (rev_on ? k : 0) + esc_co;
  int total = 1 + rev_co;

  fprintf(log_file, "state posterior mu ");
  for(int s = 0; s < S; s++) fprintf(log_file, "omega[%i] ", s);
  for(int s = 1; s < S; s++) fprintf(log_file, "R[%i] ", s);
  fprintf(log_file, "\n");
  if(hla_path != NULL) fprintf(window_file, "number of HLAs: %i\n", H);
  if(hla_path == NULL && !no_rev) fprintf(window_file, "reversion inference only\n");
  fprintf(priors_file, "Closest sequences: %i\nMutation\nmu_prior: %f\nRecombination\nlambda_rec: %f\n", closest_n,
          MCQ_MU_PRIOR, lambda_rec);
  fclose(priors_file);

  const double T0 = no_anneal ? 1 : 500;
  struct timespec ts0, ts1;
  clock_gettime(CLOCK_MONOTONIC, &ts0);
  ts1 = ts0;
  time_t t_start = time(NULL);
  int hours_passed = 1, sample_i = 1, chain_i, hours = 0;
  double prof_s[9] = {0}, state_files_s = 0;
  long prof_n[9] = {0};
  for(chain_i = 0; chain_i < chain_length; chain_i++, sample_i++) {
    int pick = mcq_rand_lim(total);
    int move = pick < rec ? 0 : pick < sel ? 1 : pick < esc_ms ? 2 : pick < rev_ms ? 3 : pick < esc_gr ? 4 :
               pick < rev_gr ? 5 : pick < esc_co ? 6 : pick < rev_co ? 7 : 8;
    double temperature = 1 + (T0 - 1) * pow(0.999, chain_i);

    if(sample_i == sample_rate) {
      sample_i = 0;
      fprintf(log_file, "%i %f %f ", chain_i, c.llk, c.mu);
      for(int s = 0; s < S; s++) fprintf(log_file, "%f ", c.omega[s]);
      for(int s = 1; s < S; s++) fprintf(log_file, "%f ", c.R[s]);
      fprintf(log_file, "\n");
      if(hla_path != NULL) {
        fprintf(window_file, "%i\n", chain_i);
        mcq_windows_print(c.hw[0].rev, window_file);
        for(int h = 0; h < H; h++) mcq_windows_print(c.hw[h].esc, window_file);
      } else if(!no_rev) {
        fprintf(window_file, "%i\n", chain_i);
        mcq_windows_print(c.hw[0].rev, window_file);
      }
    }

    struct timespec tm0, tm1;
    if(opt_profile) clock_gettime(CLOCK_MONOTONIC, &tm0);
    switch(move) {
      case 0: move_recombination(&c, temperature); break;
      case 1: move_selection(&c, temperature); break;
      case 8: move_mutation(&c, temperature); break;
      default:
        move_window(&c, temperature, (move & 1) != 0, move <= 3 ? ACT_MERGE_OR_SPLIT : move <= 5 ? ACT_GROW : ACT_COEFF);
    }
    if(opt_profile) {
      clock_gettime(CLOCK_MONOTONIC, &tm1);
      prof_s[move] += (tm1.tv_sec - tm0.tv_sec) + 1e-9 * (tm1.tv_nsec - tm0.tv_nsec);
      prof_n[move]++;
    }

    /* end of chain / wall-clock limit / hourly interim state (dmodel.c:1203-1235) */
    /* with several ranks the clock of rank 0 decides for everyone (checked every 256 steps and on the
     * last step of the chain): the ranks stop at the same step and write the same state */
    if(world == 1) hours = (int)(difftime(time(NULL), t_start) / 3600);
    else if((chain_i & 255) == 255 || (runtime_set && chain_i == chain_length - 1)) {
      double h0 = rank == 0 ? floor(difftime(time(NULL), t_start) / 3600) : 0;
      CHECK(mcq_ctx_exchange_sum(c.ctx, &h0));
      hours = (int)h0;
    }
    bool last = false;
    if(runtime_set) {
      if(hours >= runtime_hours) last = true;
      else if(chain_i == chain_length - 1) chain_length *= 2;
    } else if(chain_i == chain_length - 1) last = true;
    if(last) {
      /* the chain ends here: the device is drained and the clock stopped before the state (with the closest-L
       * lists of every query: megabytes of text at large shapes) is written out */
      if(!c.sample_prior) CHECK(mcq_ctx_synchronize(c.ctx));
      clock_gettime(CLOCK_MONOTONIC, &ts1);
      mcq_write_state_json(json_file, chain_i, c.llk, c.mu, S, c.omega, c.R, c.hw, H, Qall, closest_n, closest_all);
      break;
    }
    if(hours >= hours_passed) {
      snprintf(path, sizeof(path), "%s/%s_interim_final_state.json", out_dir, stamp);
      FILE *f = fopen(path, "w");
      if(f) {
        struct timespec tj0, tj1;
        clock_gettime(CLOCK_MONOTONIC, &tj0);
        mcq_write_state_json(f, chain_i, c.llk, c.mu, S, c.omega, c.R, c.hw, H, Qall, closest_n, closest_all);
        fclose(f);
        clock_gettime(CLOCK_MONOTONIC, &tj1);
        state_files_s += (tj1.tv_sec - tj0.tv_sec) + 1e-9 * (tj1.tv_nsec - tj0.tv_nsec);
      }
      hours_passed++;
    }
  }
  printf("Proposed %i moves.\n", chain_i);
  printf("rec_accept: %i rec_reject: %i\nsel_accept: %i sel_reject: %i\nmerge_accept: %i merge_reject: %i\n"
         "split_accept: %i split_reject: %i\ngrow_accept: %i grow_reject: %i\n"
         "window_coeff_accept: %i window_coeff_reject: %i\nmut_accept: %i mut_reject: %i\n",
         c.rec_acc, c.rec_rej, c.sel_acc, c.sel_rej, c.merge_acc, c.merge_rej, c.split_acc, c.split_rej,
         c.grow_acc, c.grow_rej, c.coeff_acc, c.coeff_rej, c.mut_acc, c.mut_rej);
  {
    /* steps per second of the chain itself: hourly interim state files are not counted, the final one is
     * written after the clock has stopped */
    double sec = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec) - state_files_s;
    int done = chain_i < chain_length ? chain_i + 1 : chain_i;
    printf("MCMC loop: %d steps in %.6f s = %.3f steps/s\n", done, sec, sec > 0 ? done / sec : 0.0);
  }
  if(opt_profile) {
    static const char *names[9] = {"recombination", "selection", "esc merge/split", "rev merge/split", "esc grow",
                                   "rev grow", "esc coeff", "rev coeff", "mutation"};
    for(int m = 0; m < 9; m++)
      if(prof_n[m]) printf("profile: %-16s %7ld moves %10.3f ms total %9.1f us each\n", names[m], prof_n[m],
                           prof_s[m] * 1e3, prof_s[m] / prof_n[m] * 1e6);
  }
  printf("final log-likelihood: %.10f\n", c.llk);
  printf("time taken: %.2f\n", difftime(time(NULL), t_start));

  mcq_ctx_destroy(c.ctx);
  fclose(log_file);
  if(window_file) fclose(window_file);
  fclose(json_file);
  if(trace) fclose(trace);
  mcq_data_free(&d);
  return EXIT_SUCCESS;
}
