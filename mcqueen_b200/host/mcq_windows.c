#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mcq_rng.h"
#include "mcq_windows.h"

void mcq_windows_alloc(int num_sites, int num_hla, int init_esc, int init_rev, McqHlaWindows *hw)
{
  McqWindowSet *sets = malloc((size_t)num_hla * 2 * sizeof(McqWindowSet));
  McqWindow *all = malloc((size_t)num_sites * num_hla * 2 * sizeof(McqWindow));
  for(int i = 0; i < num_hla * 2; i++) {
    sets[i].count = (i % 2 == 0) ? init_esc : init_rev;
    sets[i].w = all + (size_t)num_sites * i;
    for(int j = 0; j < num_sites; j++) { sets[i].w[j].start = j; sets[i].w[j].length = 0; sets[i].w[j].coeff = 0; }
  }
  for(int h = 0; h < num_hla; h++) { hw[h].esc = &sets[2*h]; hw[h].rev = &sets[2*h+1]; }
}

void mcq_windows_free(McqHlaWindows *hw)
{
  free(hw[0].esc->w);
  free(hw[0].esc);
}

static int by_start(const void *a, const void *b)
{
  return ((const McqWindow *)a)->start - ((const McqWindow *)b)->start;
}

static void set_lengths(McqWindowSet *ws, int num_sites)
{
  for(int i = 0; i + 1 < ws->count; i++) ws->w[i].length = ws->w[i+1].start - ws->w[i].start;
  ws->w[ws->count-1].length = num_sites - ws->w[ws->count-1].start;
}

void mcq_windows_randomise(McqWindowSet *ws, int num_sites, int count, double log_mu, double sigma)
{
  /* window 0 always starts at site 0; the other starts are drawn without replacement from the pool
   * (selection_window.c:289-303): one rand_lim per window */
  for(int i = 1; i < count; i++) {
    int pick = i + mcq_rand_lim(num_sites - i);
    McqWindow t = ws->w[i]; ws->w[i] = ws->w[pick]; ws->w[pick] = t;
  }
  ws->count = count;
  qsort(ws->w, count, sizeof(McqWindow), by_start);
  set_lengths(ws, num_sites);
  for(int i = 0; i < count; i++) ws->w[i].coeff = mcq_ran_lognormal(log_mu, sigma);
}

void mcq_windows_set(McqWindowSet *ws, int num_sites, int count, const int *start, const double *coeff)
{
  for(int i = 1; i < count; i++) {
    McqWindow t = ws->w[i]; ws->w[i] = ws->w[start[i]]; ws->w[start[i]] = t;
  }
  ws->count = count;
  qsort(ws->w, count, sizeof(McqWindow), by_start);
  set_lengths(ws, num_sites);
  for(int i = 0; i < count; i++) ws->w[i].coeff = coeff[i];
}

void mcq_windows_to_sites(const McqWindowSet *ws, double *per_site)
{
  for(int i = 0; i < ws->count; i++)
    for(int s = ws->w[i].start; s < ws->w[i].start + ws->w[i].length; s++) per_site[s] = ws->w[i].coeff;
}

/* the live window containing `site` (selection_window.c:96-113) */
static int find_window(const McqWindowSet *ws, int site)
{
  int lo = 0, hi = ws->count - 1;
  for(;;) {
    int mid = (lo + hi) / 2;
    if(ws->w[mid].start > site) hi = mid - 1;
    else if(ws->w[mid].start + ws->w[mid].length <= site) lo = mid + 1;
    else return mid;
  }
}

int mcq_windows_split(McqWindowSet *ws, int num_sites, double *old_coeff)
{
  int n = ws->count;
  if(n == num_sites) return -1;
  /* a random unused start site: swap it to the head of the pool (selection_window.c:86-94) */
  int pick = n + mcq_rand_lim(num_sites - n);
  McqWindow t = ws->w[n]; ws->w[n] = ws->w[pick]; ws->w[pick] = t;
  McqWindow fresh = ws->w[n];
  int i = find_window(ws, fresh.start);
  *old_coeff = ws->w[i].coeff;
  int old_len = ws->w[i].length;
  int len0 = fresh.start - ws->w[i].start, len1 = old_len - len0;
  ws->w[i].length = len0;
  fresh.length = len1;
  /* the two new coefficients keep the length-weighted geometric mean (selection_window.c:139-150) */
  double u = mcq_rand_lim(1);
  double ratio = (1 - u) / u;
  double c0 = ws->w[i].coeff * pow(ratio, -(double)len1 / old_len);
  double c1 = ws->w[i].coeff * pow(ratio, (double)len0 / old_len);
  ws->w[i].coeff = c0;
  fresh.coeff = c1;
  memmove(ws->w + i + 2, ws->w + i + 1, (size_t)(n - i - 1) * sizeof(McqWindow));
  ws->w[i+1] = fresh;
  ws->count++;
  return i;
}

void mcq_windows_split_undo(McqWindowSet *ws, int w, double old_coeff)
{
  ws->w[w].length += ws->w[w+1].length;
  ws->w[w].coeff = old_coeff;
  int freed = ws->w[w+1].start;
  memmove(ws->w + w + 1, ws->w + w + 2, (size_t)(ws->count - w - 2) * sizeof(McqWindow));
  ws->count--;
  ws->w[ws->count].start = freed;
}

int mcq_windows_merge(McqWindowSet *ws, int *old_start, double *old0, double *old1)
{
  if(ws->count == 1) return -1;
  /* window i is absorbed by its left neighbour (selection_window.c:173-203) */
  int i = 1 + mcq_rand_lim(ws->count - 1);
  *old_start = ws->w[i].start;
  *old0 = ws->w[i-1].coeff;
  *old1 = ws->w[i].coeff;
  int l0 = ws->w[i-1].length, l1 = ws->w[i].length;
  ws->w[i-1].coeff *= pow(*old1 / *old0, (double)l1 / (l0 + l1));
  ws->w[i-1].length += l1;
  memmove(ws->w + i, ws->w + i + 1, (size_t)(ws->count - i - 1) * sizeof(McqWindow));
  ws->count--;
  ws->w[ws->count].start = *old_start;
  return i - 1;
}

void mcq_windows_merge_undo(McqWindowSet *ws, int w, int old_start, double old0, double old1)
{
  int end = ws->w[w].start + ws->w[w].length;
  ws->w[w].length = old_start - ws->w[w].start;
  ws->w[w].coeff = old0;
  memmove(ws->w + w + 2, ws->w + w + 1, (size_t)(ws->count - w - 1) * sizeof(McqWindow));
  ws->count++;
  ws->w[w+1].start = old_start;
  ws->w[w+1].coeff = old1;
  ws->w[w+1].length = end - old_start;
}

int mcq_windows_grow(McqWindowSet *ws, double geom_p, int *old_start)
{
  if(ws->count == 1) return -1;
  /* the boundary between windows i-1 and i moves by a signed geometric step
   * (selection_window.c:225-248); all three draws happen before the feasibility test */
  int i = 1 + mcq_rand_lim(ws->count - 1);
  int dir = (int)mcq_rand_lim(2) * 2 - 1;
  int step = dir * (int)mcq_ran_geometric(geom_p);
  if(step >= ws->w[i].length || -step >= ws->w[i-1].length) return -1;
  *old_start = ws->w[i].start;
  ws->w[i-1].length += step;
  ws->w[i].start += step;
  ws->w[i].length -= step;
  return i - 1;
}

void mcq_windows_grow_accept(McqWindowSet *ws, int new_start, int old_start, int num_sites)
{
  /* the pool entry of the site that became a start now stands for the site that stopped being one */
  for(int i = ws->count; i < num_sites; i++)
    if(ws->w[i].start == new_start) { ws->w[i].start = old_start; break; }
}

void mcq_windows_grow_undo(McqWindowSet *ws, int w, int old_start)
{
  ws->w[w].length = old_start - ws->w[w].start;
  int end = ws->w[w+1].start + ws->w[w+1].length;
  ws->w[w+1].length = end - old_start;
  ws->w[w+1].start = old_start;
}

int mcq_windows_coeff(McqWindowSet *ws, double *old_coeff)
{
  int w = mcq_rand_lim(ws->count);
  *old_coeff = ws->w[w].coeff;
  ws->w[w].coeff = *old_coeff * exp(mcq_rand_lim(2) - 1);
  return w;
}

void mcq_windows_coeff_undo(McqWindowSet *ws, int w, double old_coeff) { ws->w[w].coeff = old_coeff; }

void mcq_windows_print(const McqWindowSet *ws, FILE *fh)
{
  fprintf(fh, "%i\n", ws->count);
  for(int i = 0; i < ws->count; i++)
    fprintf(fh, "%i,%i,%f%s", ws->w[i].start, ws->w[i].length, ws->w[i].coeff, i + 1 < ws->count ? " | " : "\n");
  fflush(fh);
}
