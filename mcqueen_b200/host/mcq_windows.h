/* mcq_windows.h - selection-window state of the MCMC (host side, O(windows) work).
 *
 * Same model as the reference's selection_window.c: per HLA type a partition of the sites into
 * windows with one escape coefficient each, plus one shared partition for reversion.  A window set
 * is an array of num_sites entries: the first `count` are the live windows in site order, the rest
 * is the pool of still-unused start sites (only `.start` is meaningful there).  The proposal
 * functions draw from the same streams, in the same order, as selection_window.c:117-263, so that a
 * chain run through this driver makes the reference's decisions.
 */
#ifndef MCQ_WINDOWS_H_
#define MCQ_WINDOWS_H_

#include <stdio.h>

typedef struct { int start, length; double coeff; } McqWindow;
typedef struct { McqWindow *w; int count; } McqWindowSet;
typedef struct { McqWindowSet *esc, *rev; } McqHlaWindows;

void mcq_windows_alloc(int num_sites, int num_hla, int init_esc, int init_rev, McqHlaWindows *hw);
void mcq_windows_free(McqHlaWindows *hw);
/* selec_windows_randomise, selection_window.c:289-320 */
void mcq_windows_randomise(McqWindowSet *ws, int num_sites, int count, double log_mu, double sigma);
/* selec_windows_set (used when a state is resumed), selection_window.c:322-350 */
void mcq_windows_set(McqWindowSet *ws, int num_sites, int count, const int *start, const double *coeff);
/* update_HLA_selection, selection_window.c:375-385 */
void mcq_windows_to_sites(const McqWindowSet *ws, double *per_site);

/* proposals; each returns the index of the left-hand window (or -1: proposal impossible) */
int  mcq_windows_split(McqWindowSet *ws, int num_sites, double *old_coeff);
void mcq_windows_split_undo(McqWindowSet *ws, int w, double old_coeff);
int  mcq_windows_merge(McqWindowSet *ws, int *old_start, double *old0, double *old1);
void mcq_windows_merge_undo(McqWindowSet *ws, int w, int old_start, double old0, double old1);
int  mcq_windows_grow(McqWindowSet *ws, double geom_p, int *old_start);
void mcq_windows_grow_accept(McqWindowSet *ws, int new_start, int old_start, int num_sites);
void mcq_windows_grow_undo(McqWindowSet *ws, int w, int old_start);
int  mcq_windows_coeff(McqWindowSet *ws, double *old_coeff);
void mcq_windows_coeff_undo(McqWindowSet *ws, int w, double old_coeff);

/* selec_window_print_to_file, selection_window.c:361-373 */
void mcq_windows_print(const McqWindowSet *ws, FILE *fh);

#endif
