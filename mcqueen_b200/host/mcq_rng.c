#include <math.h>
#include <stdlib.h>

#include "mcq_rng.h"

double mcq_rand_lim(int limit) { return limit * ((double)random() / (double)RAND_MAX); }

void mcq_seed_libc(unsigned seed) { srandom(seed); }

/* MT19937 (Matsumoto & Nishimura 1998, 2002 initialisation) */
#define MT_N 624
#define MT_M 397
static unsigned long mt[MT_N];
static int mti = MT_N + 1;

void mcq_mt_seed(unsigned long seed)
{
  if(seed == 0) seed = 4357;   /* the twister must not start from an all-zero state */
  mt[0] = seed & 0xffffffffUL;
  for(int i = 1; i < MT_N; i++)
    mt[i] = (1812433253UL * (mt[i-1] ^ (mt[i-1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  mti = MT_N;
}

static unsigned long mt_next(void)
{
  if(mti > MT_N) mcq_mt_seed(0);
  if(mti >= MT_N) {
    for(int k = 0; k < MT_N; k++) {
      unsigned long y = (mt[k] & 0x80000000UL) | (mt[(k+1) % MT_N] & 0x7fffffffUL);
      mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    mti = 0;
  }
  unsigned long y = mt[mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680UL;
  y ^= (y << 15) & 0xefc60000UL;
  y ^= (y >> 18);
  return y & 0xffffffffUL;
}

static double uniform_pos(void)
{
  double x;
  do { x = mt_next() / 4294967296.0; } while(x == 0);
  return x;
}

unsigned mcq_ran_geometric(double p)
{
  double u = uniform_pos();
  if(p == 1) return 1;
  return (unsigned)(log(u) / log(1 - p) + 1);
}

double mcq_ran_lognormal(double zeta, double sigma)
{
  double x, y, r2;
  do {
    x = -1 + 2 * uniform_pos();
    y = -1 + 2 * uniform_pos();
    r2 = x * x + y * y;
  } while(r2 > 1.0 || r2 == 0);
  return exp(sigma * y * sqrt(-2.0 * log(r2) / r2) + zeta);
}
