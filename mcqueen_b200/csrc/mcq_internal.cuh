// mcq_internal.cuh - device-side data layout and kernel launch interfaces of libmcq_b200.
//
// Layout in HBM (one context = one shard of Q queries on one GPU):
//   G      u8  [Q][S][Lg]   slot of the codon that copy state r shows at site s
//                           (0 = the site's consensus codon, 1..K_s = its other reference codons in
//                           ascending order, 65 = a codon no reference shows / padding); Lg = 64*NJ,
//                           bytes permuted so that a lane's 2*NJ states are contiguous
//   F,B,T  f64 [Q][S][Ls]   forward, backward, proposal ("tmp") forward; SITE-major, state fastest
//   Fn,Bn,Tn i32 [Q][S]     rescale counters (F_rnorm etc. of the reference)
//   llk_cur, llk_prop f64 [Q]
// Ls = L rounded up to 16 (128-byte rows).  A warp owns one query; lane l owns states
// 64*j + 2*l + {0,1}, j < NJ = ceil(Ls/64).
//
// Emission rows.  The reference stores codon_to_codon_HLA[q][c1][s] per query (mcmc_moves.c:33-170).
// For a non-consensus c1 the value depends on the query only through its codon `to` at the site
// (the reversion coefficient is shared by all queries, mcmc_moves.c:70), so those rows are kept once
// per (site, to):   TAB f64 [sum_s 64*K_s]: site s owns 64 rows (one per `to`) of K_s entries at
// toff[s]; ATAB f64 [sum_s K_s] holds the matching "A" values (independent of `to`).  Only the
// consensus row depends on the query's HLA profile: EC, AC f64 [Q][S].  A warp stages, per site,
// row (s, to(q,s)) of TAB into shared-memory slots 1..K_s and EC[q][s] into slot 0.
#ifndef MCQ_INTERNAL_CUH_
#define MCQ_INTERNAL_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define MCQ_BIG 1000000.0   // constants.c:13
#define MCQ_BIGI 100.0      // constants.c:14
#define MCQ_NOSLOT 65       // slot value of "a codon no reference shows" / padding: staged emission 0

struct McqModel {            // device-resident parameter block (all device pointers)
  double *R;                 // [S]
  double *omega;             // [S]
  double *rate_out;          // [S][64]  (site-major transpose of rate_out_of_codons[64][S])
  double *esc;               // [H][S]
  double *rev;               // [S]
  double *prior;             // [64]
  double mu, kappa, phi;
};

struct McqStatic {           // device-resident static tables
  int Q, S, L, Ls, N, H;
  int Lg;                    // bytes per G row: 64 * NJ with NJ = 1,2,4,8,16 covering Ls
  const uint8_t *G;          // [Q][S][Lg], states permuted lane-major (see k_gather)
  const int *aoff;           // [S+1] first non-consensus slot of site s in ATAB / ncodon; K_s = aoff[s+1]-aoff[s]
  const int *toff;           // [S+1] offset of site s's block of rows in TAB
  const int2 *desc;          // [S] {toff[s], K_s}
  const uint32_t *rowdesc;   // [Q][S] K_s << 25 | offset of row (s, to(q,s)) in TAB: what a warp needs to stage
                             // the emission row of a site, in one word (built by mcq_launch_rowdesc)
  const uint8_t *ncodon;     // [aoff[S]] codon of each non-consensus slot
  const int *nsite;          // [aoff[S]] site of each non-consensus slot
  const uint8_t *cons_codon; // [S]
  const uint8_t *cons_in_list;  // [S] 1 when some reference shows the consensus codon at the site
  const uint8_t *query_codon;// [Q][S]; nullptr = every row is row 0 (the drop-in symbols)
  const uint8_t *hla;        // [Q][H] 0/1
};

struct McqEmis {             // one consistent set of emission rows
  const double *tab;         // TAB
  const double *ec;          // EC [Q][S]; nullptr = slot 0 is never referenced
};

// ---- proposal overrides applied on top of McqModel inside a kernel (no H2D copy per step)
struct McqOverride {
  int kind;                  // -1 none, else MCQ_PROP_*
  int start, end, hla, split;
  double v0, v1;
};

// F and B are kept valid LAZILY: per query, F[q][s] is current for s <= fv[q] and B[q][s] for
// s >= bv[q].  An accepted proposal on sites c..d sets fv = bv = d for the queries it touched; the next
// evaluation that needs F[c'-1] or B[d'] first catches the missing columns up (k_forward runs the sites
// fv+1 .. start-1 with the current parameters before the proposal's own range; k_backward runs bv-1
// down to the requested column).  The reference refills both matrices completely after every accept
// (mcmc_moves.c:261-279); the values are the same, they are just computed when first needed.
struct FwdArgs {
  McqStatic st;
  McqEmis em;                // emission rows read for sites start..end
  McqEmis em_pre;            // current emission rows, for the catch-up sites before `start`
  int *fv;                   // [Q] last valid column of Fin per query; nullptr = column start-1 is valid
  int *mark_fv;              // [Q] or nullptr: receives `end` for every query the launch ran (a full sweep)
  const double *R;
  int r_site; double r_value;     // R override at one site (-1 = none)
  // Two column buffers: Fin = buffer 0, Fout = buffer 1.  sel[q][s] (0/1, or nullptr = always 0) names the
  // buffer that holds the CURRENT column s of query q.  Column start-1 is read from the current buffer;
  // catch-up sites and in_place runs write the current buffer, anything else (a proposal) writes the other
  // one - so an accept only flips selectors (k_commit_cols) instead of copying columns.
  const double *Fin; const int *Fn_in;
  double *Fout; int *Fn_out;
  const uint8_t *sel;
  int in_place;              // 1: sites start..end go where the current column is (refills, the full sweep)
  int no_store;              // 1: likelihood only, no column is written (start must be 0)
  int start, end;
  int llk_mode;              // 0 none, 1 dot with B at `end`, 2 plain column sum at `end`
  const double *B; const int *Bn;
  double *llk_out;           // [Q]
  const double *llk_keep;    // [Q] value copied to llk_out for inactive queries (may be null)
  const uint8_t *active;     // [Q][H] carrier table or null
  int active_hla;            // column of `active` to test
  double logBIG;
  int pipe;                  // 1: few queries in flight - launch the variant with the shorter per-site chain (same bits)
  // checkpointed matrices (ck > 1, see k_forward_ck): only the columns of sites s with s % ck == ck-1 exist, as
  // column s / ck of nc per query (Fin, Fout, sel, Fn_in, Fn_out are indexed that way); Tcol / Tncol ([Q][Ls],
  // [Q]) receive the column of site `end` and its counter when not null
  int ck, nc;
  double *Tcol; int *Tncol;
};

struct BwdArgs {
  McqStatic st;
  McqEmis em;
  const double *R;
  double *B; int *Bn;
  int update;                // first column to compute when bv == nullptr (S-1 = fresh start)
  int bottom;                // last column to compute (0 = down to the first site)
  int *bv;                   // [Q] first valid column of B per query (S = none); nullptr = use `update`
  const uint8_t *active; int active_hla;
  int pipe;                  // as FwdArgs::pipe
  // checkpointed matrices (ck > 1, see k_backward_ck): only the columns of sites s with s % ck == 0 exist, as
  // column s / ck of nc per query; Bcol / Bncol ([Q][Ls], [Q]) receive the column of site `bottom` when not null
  int ck, nc;
  double *Bcol; int *Bncol;
};

struct DotArgs {             // per-query log-likelihood from a proposal column and the backward column
  McqStatic st;
  const double *T; const int *Tn;   // buffer 1 / proposal counters
  const double *F;                  // buffer 0
  const uint8_t *sel;               // [Q][S] which buffer holds the current column (nullptr: buffer 0)
  const double *B; const int *Bn;
  int site;
  double *llk_out;
  const double *llk_keep;
  const uint8_t *active; int active_hla;
  double logBIG;
  int wpq;                          // consumer warps per query of the sweeps (0: one warp): fixes the summation order
  int cols;                         // 1: T / B are per-query columns [Q][Ls] with counters Tn / Bn [Q] (checkpointed matrices)
};

struct EmisArgs {
  McqStatic st;
  McqModel m;
  McqOverride ov;
  int start, end, which_hla;
  int a0, a1;                // non-consensus slot bounds of the range: aoff[start], aoff[end+1]
  const double *tab_src, *atab_src, *ec_src, *ac_src;
  double *tab_dst, *atab_dst, *ec_dst, *ac_dst;
};

// launchers (mcq_kernels.cu); all asynchronous on `stream`, return cudaGetLastError()
cudaError_t mcq_launch_tables_init();
// pitch of the re-pitched reference slot rows (trip_pad, N rows): whole 128-byte lines, so that a gather tile (64
// sites) never reads past a row
static inline size_t mcq_gather_pitch(int S) { return ((size_t)S + 127) / 128 * 128; }
// the query-independent tables of the gather: trip_tab [ceil32(S)][128] (triplet -> slot per site) and trip_pad
// [N][mcq_gather_pitch(S)], the SLOT of every reference codon (both scratch)
cudaError_t mcq_launch_gather_prepare(const int8_t *ref_triplets, int8_t *trip_pad, const uint8_t *slot_of /*[S][64]*/,
                                      uint8_t *trip_tab, int S, int N, cudaStream_t stream);
// G rows of Q queries (pointers address the first of them); trip = ref_triplets on the device (for the triplets with
// unknown bases); hla_chars [Q][H] '0'/'1' -> 0/1 in place, or null; err: device flag, 2 = closest out of range
cudaError_t mcq_launch_gather(const int8_t *trip_pad, const int8_t *trip, const int *closest, const int8_t *cons_nucls,
                              const uint8_t *slot_of, uint8_t *G, uint8_t *hla_chars, int Q,
                              int S, int L, int Lg, int N, int H, int *err, cudaStream_t stream);
// rowdesc[q][s] for queries q0 .. q0+nq-1 from desc and query_codon (st.rowdesc itself is not read)
cudaError_t mcq_launch_rowdesc(const McqStatic &st, uint32_t *out, int q0, int nq, cudaStream_t stream);
cudaError_t mcq_launch_forward(const FwdArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_backward(const BwdArgs &a, cudaStream_t stream);
// the sweeps over checkpointed matrices (FwdArgs::ck / BwdArgs::ck > 1)
cudaError_t mcq_launch_forward_ck(const FwdArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_backward_ck(const BwdArgs &a, cudaStream_t stream);
// a one-column proposal with nothing owed (start == end > 0, all queries): same values as mcq_launch_forward's
cudaError_t mcq_launch_column(const FwdArgs &a, cudaStream_t stream);
// time the launch at every residency (blocks per SM) the kernel allows and keep the fastest for launches of
// this kernel with this many queries; the launch itself must be repeatable (same inputs, same outputs)
cudaError_t mcq_tune_forward(const FwdArgs &a, cudaStream_t stream, int *best_blocks, float *best_ms);
cudaError_t mcq_tune_backward(const BwdArgs &a, cudaStream_t stream, int *best_blocks, float *best_ms);
// the same sweeps with W consumer warps (+ one producer warp) per query (mcq_sweeps.cu); mcq_mw_supported says
// which (NJ = Lg/64, W) pairs are built
bool mcq_mw_supported(int NJ, int W);
cudaError_t mcq_launch_forward_mw(const FwdArgs &a, int W, cudaStream_t stream);
cudaError_t mcq_launch_backward_mw(const BwdArgs &a, int W, cudaStream_t stream);
cudaError_t mcq_launch_dot(const DotArgs &a, cudaStream_t stream);
// fv[q] := f, bv[q] := b for the queries selected by (active, hla) (all when active == nullptr); a
// negative value leaves that array untouched
cudaError_t mcq_launch_set_valid(int *fv, int *bv, int Q, int H, int f, int b, const uint8_t *active, int hla,
                                 cudaStream_t stream);
// the reversion part (rows of non-consensus codons) / the escape part (the consensus row) of
// create_codon_to_codon_HLA
cudaError_t mcq_launch_emission_table(const EmisArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_emission_cons(const EmisArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_emission_both(const EmisArgs &a, cudaStream_t stream);   // the two above in one launch
cudaError_t mcq_launch_mu_rescale(const McqStatic &st, const McqModel &m, double mu_new, int nslots,
                                  const double *tab, const double *atab, const double *ec, const double *ac,
                                  double *tab_o, double *atab_o, double *ec_o, double *ac_o, cudaStream_t stream);
// everything an accepted window / column proposal makes current, in one launch (k_commit_all)
struct CommitArgs {
  McqStatic st; McqModel m; McqOverride ov;
  const double *tab_t, *atab_t, *ec_t, *ac_t;
  double *tab, *atab, *ec, *ac;
  size_t t0, t1, a0, a1;          // ranges of the shared table / A values to copy (empty when t0 == t1)
  int copy_cons;                  // 1: per-query consensus values of start..end
  uint8_t *sel; const int *Tn; int *Fn;  // indexed [Q][nc]: checkpoint columns c0..c1 of every query flip
  int nc, c0, c1;
  const uint8_t *active; int active_hla;
  int *fv, *bv; int mark;         // mark >= 0: fv = bv = mark for the active queries
};

cudaError_t mcq_launch_commit_all(const CommitArgs &a, cudaStream_t stream);
// host-mapped (pinned) result slot: the reduction kernel writes the total, then the evaluation number
struct McqHostSlot {
  double value;
  unsigned long long seq;
  int err;
};
// hs: device view of the slot (nullptr = none), seq: the number to publish
// err: device error flag copied into the slot with the result (nullptr: none)
cudaError_t mcq_launch_sum(const double *v, int n, double *out, McqHostSlot *hs, unsigned long long seq,
                           const int *err, cudaStream_t stream);
// barrier over the attached ranks (monotone `count`); err: device flag, 3 = a peer never arrived
cudaError_t mcq_launch_peer_barrier(const void *box, unsigned long long count, int *err, cudaStream_t stream);
// `box` points at a host copy of the PeerBox struct of mcq_kernels.cu (16 value + 16 seq + 16 bar pointers, nranks, rank)
cudaError_t mcq_launch_sum_exchange(const double *v, int n, double *out, const void *box, unsigned long long step,
                                    int *err, McqHostSlot *hs, unsigned long long seq, cudaStream_t stream);

#endif
