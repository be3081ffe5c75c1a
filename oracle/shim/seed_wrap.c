/* TEST INFRASTRUCTURE ONLY.  Link-time seeding override for the reference
 * dmodel (-Wl,--wrap=init_rand,--wrap=rand): the reference seeds libc with
 * time(NULL) (util.c:19-22) and cannot be given a seed.  MCQ_SEED (default 1)
 * replaces it.  The rand() wrapper records, at each of the four accept tests
 * (mcmc_moves.c:251,437,941,1153 - the only callers of rand()), the accept /
 * reject counters of mcmc_moves.c:13-19 so the per-step decision sequence can
 * be reconstructed without touching reference code.
 */
#include <stdio.h>
#include <stdlib.h>

extern int rec_accept, rec_reject, sel_accept, sel_reject, merge_accept, merge_reject;
extern int split_accept, split_reject, grow_accept, grow_reject;
extern int wcoeff_accept, wcoeff_reject, mut_accept, mut_reject;

int __real_rand(void);

void __wrap_init_rand(void)
{
  const char *env = getenv("MCQ_SEED");
  srandom(env != NULL ? (unsigned)strtoul(env, NULL, 10) : 1u);
}

static FILE *trace_fh = NULL;
static int trace_init = 0;

int __wrap_rand(void)
{
  if(!trace_init) {
    const char *path = getenv("MCQ_TRACE");
    if(path != NULL) trace_fh = fopen(path, "w");
    trace_init = 1;
  }
  int v = __real_rand();
  if(trace_fh != NULL) {
    /* counters BEFORE this decision + the draw; decisions follow by differencing */
    fprintf(trace_fh, "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n",
            rec_accept, rec_reject, sel_accept, sel_reject, merge_accept, merge_reject,
            split_accept, split_reject, grow_accept, grow_reject,
            wcoeff_accept, wcoeff_reject, mut_accept, mut_reject, v);
    fflush(trace_fh);
  }
  return v;
}
