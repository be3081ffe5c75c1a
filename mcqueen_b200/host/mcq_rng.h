/* mcq_rng.h - the two random streams of dmodel, host side.
 *
 * (1) libc random(): every rand_lim() draw and, through rand(), the four accept tests
 *     (util.c:24-30, mcmc_moves.c:251,437,941,1153).  One shared glibc state; the library's
 *     closest-L tie shuffles draw from it too (closest_seqs.c:57-64), so the driver never
 *     re-seeds it after start-up.
 * (2) a Mersenne twister for the two distributions selection_window.c takes from GSL
 *     (gsl_ran_geometric :231, gsl_ran_lognormal :317).  GSL is not a dependency: MT19937 and
 *     the two samplers are implemented here from their published definitions.
 */
#ifndef MCQ_RNG_H_
#define MCQ_RNG_H_

/* limit * random()/RAND_MAX, as util.c:24-30 (can return `limit` itself with probability 2^-31) */
double mcq_rand_lim(int limit);

void mcq_seed_libc(unsigned seed);
void mcq_mt_seed(unsigned long seed);
/* trials up to and including the first success, success probability p */
unsigned mcq_ran_geometric(double p);
/* exp(N(zeta, sigma^2)), polar Box-Muller */
double mcq_ran_lognormal(double zeta, double sigma);

#endif
