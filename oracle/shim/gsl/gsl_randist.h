/* TEST INFRASTRUCTURE ONLY: stand-in for GSL 1.16 <gsl/gsl_randist.h>
 * (see gsl_rng.h in this directory; parity unpinned at the GSL boundary). */
#ifndef MCQ_SHIM_GSL_RANDIST_H_
#define MCQ_SHIM_GSL_RANDIST_H_
#include "gsl_rng.h"

unsigned int gsl_ran_geometric(gsl_rng *r, double p);
double gsl_ran_lognormal(gsl_rng *r, double zeta, double sigma);
double gsl_ran_exponential(gsl_rng *r, double mu);

#endif
