/* TEST INFRASTRUCTURE ONLY.  GSL 1.16 stand-in (see shim/gsl/gsl_rng.h).
 *
 * Seeding: the reference seeds GSL with time*getpid (selection_window.c:23-27)
 * and cannot be given a seed.  For reproducible fixtures, when the environment
 * variable MCQ_GSL_SEED is set, gsl_rng_set() uses that value instead of the
 * one it is passed.
 */
#include <stdlib.h>
#include <math.h>
#include "gsl/gsl_rng.h"
#include "gsl/gsl_randist.h"

#define MT_N 624
#define MT_M 397

static const gsl_rng_type mt19937_type = { "mt19937" };
const gsl_rng_type *gsl_rng_default = &mt19937_type;

const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }

void gsl_rng_set(gsl_rng *r, unsigned long seed)
{
  const char *env = getenv("MCQ_GSL_SEED");
  if(env != NULL) seed = strtoul(env, NULL, 10);
  if(seed == 0) seed = 4357;
  r->mt[0] = seed & 0xffffffffUL;
  for(int i = 1; i < MT_N; i++) {
    unsigned long prev = r->mt[i-1];
    r->mt[i] = (1812433253UL * (prev ^ (prev >> 30)) + (unsigned long)i) & 0xffffffffUL;
  }
  r->mti = MT_N;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
  (void)T;
  gsl_rng *r = malloc(sizeof(gsl_rng));
  if(r != NULL) gsl_rng_set(r, 0);
  return r;
}

void gsl_rng_free(gsl_rng *r) { free(r); }

unsigned long gsl_rng_get(gsl_rng *r)
{
  unsigned long *mt = r->mt;
  if(r->mti >= MT_N) {
    int k;
    for(k = 0; k < MT_N; k++) {
      unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k+1) % MT_N] & 0x7fffffffUL);
      mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    r->mti = 0;
  }
  unsigned long y = mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680UL;
  y ^= (y << 15) & 0xefc60000UL;
  y ^= (y >> 18);
  return y & 0xffffffffUL;
}

double gsl_rng_uniform(gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

double gsl_rng_uniform_pos(gsl_rng *r)
{
  double x;
  do { x = gsl_rng_uniform(r); } while(x == 0);
  return x;
}

/* Number of Bernoulli(p) trials up to and including the first success. */
unsigned int gsl_ran_geometric(gsl_rng *r, double p)
{
  double u = gsl_rng_uniform_pos(r);
  if(p == 1) return 1;
  return (unsigned int)(log(u) / log(1 - p) + 1);
}

static double ran_gaussian_polar(gsl_rng *r, double sigma)
{
  double x, y, r2;
  do {
    x = -1 + 2 * gsl_rng_uniform_pos(r);
    y = -1 + 2 * gsl_rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while(r2 > 1.0 || r2 == 0);
  return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

double gsl_ran_lognormal(gsl_rng *r, double zeta, double sigma)
{
  return exp(ran_gaussian_polar(r, sigma) + zeta);
}

double gsl_ran_exponential(gsl_rng *r, double mu)
{
  return -mu * log(gsl_rng_uniform_pos(r));
}
