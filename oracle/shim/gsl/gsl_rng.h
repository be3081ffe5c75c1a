/* TEST INFRASTRUCTURE ONLY (oracle build): minimal stand-in for the parts of
 * GSL 1.16 <gsl/gsl_rng.h> that the reference's selection_window.c needs
 * (SURVEY.md §8c).  GSL is not vendored in /root/reference and not installed
 * here; the algorithms below are the published ones (MT19937, Matsumoto &
 * Nishimura 1998; seeding as in the 2002 reference implementation, which GSL's
 * "mt19937" generator follows).  PARITY UNPINNED at the GSL boundary: no GSL
 * build exists in this container to check the draws against.  GSL draws never
 * enter the likelihood kernels; the oracle and the GPU host driver share this
 * same shim so both see identical draws.
 */
#ifndef MCQ_SHIM_GSL_RNG_H_
#define MCQ_SHIM_GSL_RNG_H_

typedef struct { const char *name; } gsl_rng_type;
typedef struct { unsigned long mt[624]; int mti; } gsl_rng;

extern const gsl_rng_type *gsl_rng_default;

const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
void gsl_rng_free(gsl_rng *r);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);

#endif
