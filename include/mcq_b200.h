/* mcq_b200.h - C ABI of libmcq_b200.so, the B200 (sm_100a) implementation of McQueen's
 * per-query copying-model HMM likelihood path.
 *
 * Two faces, both plain C (pointers and sizes only):
 *
 *  (i)  DROP-IN SYMBOLS - the five external functions of the reference's
 *       forward_backward.c and closest_seqs.c with their exact signatures
 *       (forward_backward.h:6-42, closest_seqs.h:4-9; C99 variably-modified
 *       parameters decay to the flat pointers written here).  Linking the reference's
 *       dmodel/dparameters against libmcq_b200.so instead of those two objects needs no
 *       source change.  Host pointers in, host pointers out, one query per call.
 *
 *  (ii) BATCHED CONTEXT API (mcq_ctx_*) - all queries of a shard resident on one GPU;
 *       what the `for(q = 0; q < num_query; q++)` loops of mcmc_moves.c and dmodel.c
 *       collapse onto (INTEGRATION.md shows the patch).  One process per GPU.
 *
 * There is no CPU fallback: every entry point runs CUDA kernels and fails (non-zero
 * return + mcq_last_error(), or exit(EXIT_FAILURE) for the void drop-in symbols, the
 * reference's die() convention, util.c:54-73) when no device is usable.
 */
#ifndef MCQ_B200_H_
#define MCQ_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCQ_NUM_CODONS 64

/* ======================================================================================
 * (i) drop-in symbols
 * ==================================================================================== */

/* replaces forward_backward.c:11-76 (decl forward_backward.h:6-16).
 * codon_to_codon_HLA[64][num_sites], ref_codons/ref_triplets[total_num_refs][num_sites],
 * F[closest_n][num_sites] (state-major, as the reference), F_rnorm[num_sites]. */
void initialise_F_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                    double *codon_to_codon_HLA, char *ref_codons,
                                    char *ref_triplets, double *R, double *F, int *F_rnorm,
                                    int start_site, int end_site, int *closest_refs,
                                    int num_bases, char *cons_query_nucls);

/* replaces forward_backward.c:80-171 (decl forward_backward.h:18-30). F_in/F_out may alias;
 * start_site==0 reseeds column 0 and zeroes F_rnorm_in[0] and F_rnorm_out[0]. */
void update_F_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                double *codon_to_codon_HLA, char *ref_codons,
                                char *ref_triplets, double *R, double *F_in, int *F_rnorm_in,
                                int *F_rnorm_out, double *F_out, int start_site, int end_site,
                                int *closest_refs, int num_bases, char *cons_query_nucls);

/* replaces forward_backward.c:173-244 (decl forward_backward.h:32-42). */
void update_B_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                double *codon_to_codon_HLA, char *ref_codons,
                                char *ref_triplets, double *R, double *B, int *B_rnorm,
                                int update, int *closest_refs, int num_bases,
                                char *cons_query_nucls);

/* replaces closest_seqs.c:17-48 (decl closest_seqs.h:4-6), including its treatment of
 * unknown reference bases (closest_seqs.c:32). */
void hamming_distance_considering_unknown(char **ref_seqs, char *query_seqs, int num_refs,
                                          int num_bases, int *hamming);

/* replaces closest_seqs.c:66-99 (decl closest_seqs.h:8-9).  Consumes libc random() draws
 * exactly as the reference does through rand_lim (util.c:24-30). Requires closest_n < num_refs
 * (the reference reads element closest_n of the sorted list, closest_seqs.c:85). */
void closest_sequences(int num_refs, int *hamming, int closest_n, int *closest_refs);

/* ======================================================================================
 * (ii) batched context API
 * ==================================================================================== */

typedef struct mcq_ctx mcq_ctx;

/* Text of the last error raised on the calling thread ("" if none). */
const char *mcq_last_error(void);
/* Number of CUDA devices visible; <0 and an error text if the runtime cannot start. */
int mcq_device_count(void);
/* Library/ABI version, bumped when a signature changes. */
int mcq_abi_version(void);

typedef struct {
  int num_sites;        /* S, codons */
  int closest_n;        /* L, copy states per query (<= 1024) */
  int total_num_refs;   /* N, size of the whole panel: enters as R[s]/N (forward_backward.c:138) */
  int num_query;        /* Q of THIS shard */
  int num_hla_types;    /* H (>= 1; dmodel.c:583 forces 1 when no HLA file is given) */
  int device;           /* CUDA device ordinal */
} mcq_dims;

int  mcq_ctx_create(mcq_ctx **ctx, const mcq_dims *dims);
void mcq_ctx_destroy(mcq_ctx *ctx);

/* Static per-run data in the reference's layouts (all host pointers, copied):
 *   ref_triplets[N][S] base-5 codes (dmodel.c:684-686), closest_refs[Q][L] (dmodel.c:492),
 *   cons_query_nucls[Q][3S] codes 0..3 (dmodel.c:664-681), query_codons[Q][S] (dmodel.c:699),
 *   hla_types[Q][H] bytes '0'/'1' (rows of load_data.c:125-157 without the NUL),
 *   consensus_codons[S] (dmodel.c:757), num_ref_site_codons[S] + ref_site_codons[S][64]
 *   (amino_acids.c:155-176).
 * Builds the device-side gathered codon table (triplet_to_codon, amino_acids.c:50-66, is
 * evaluated once here instead of once per cell per sweep).  May be called again with new data of
 * the same shape and the same per-site codon lists (re-uploads and re-gathers, no allocation). */
int mcq_ctx_upload_static(mcq_ctx *ctx, const char *ref_triplets, const int *closest_refs,
                          const char *cons_query_nucls, const char *query_codons,
                          const char *hla_types, const char *consensus_codons,
                          const int *num_ref_site_codons, const char *ref_site_codons);

/* Model parameters (host pointers, copied).  Any pointer may be NULL = keep current.
 * rate_out_of_codons is [64][S] as dmodel.c:836-849; HLA_select_esc is [H][S];
 * HLA_select_rev is the shared row 0, [S] (mcmc_moves.c:70). */
typedef struct {
  double mu, kappa, phi;
  const double *prior_C2;            /* [64] */
  const double *R;                   /* [S] */
  const double *omega;               /* [S] */
  const double *rate_out_of_codons;  /* [64][S] */
  const double *HLA_select_esc;      /* [H][S] */
  const double *HLA_select_rev;      /* [S] */
} mcq_params;

int mcq_ctx_set_params(mcq_ctx *ctx, const mcq_params *p);

/* create_codon_to_codon_HLA (mcmc_moves.c:33-170) for every query of the shard, sites
 * start..end, from the context's current parameters, in place.
 * do_esc_or_rev: 0 ESC, 1 REV, 2 BOTH; which_HLA: -1 = all (mcmc_moves.h:6-9). */
int mcq_ctx_emission(mcq_ctx *ctx, int start_site, int end_site, int which_HLA, int do_esc_or_rev);

/* dmodel.c:873-902: emission table for all sites, forward fill of F, backward fill of B,
 * *llk = sum_q [ log(sum_r F[q][r][S-1]) - F_rnorm[q][S-1]*log(BIG) ].  Makes it the current state. */
int mcq_ctx_init_state(mcq_ctx *ctx, double *llk);

/* Forward fill only (initialise_F over all queries + the llk reduction). with_emission!=0
 * rebuilds the emission table first. This is the "full likelihood evaluation" of the headline
 * metric. */
int mcq_ctx_full_llk(mcq_ctx *ctx, int with_emission, double *llk);
/* The same value as mcq_ctx_full_llk without writing F: the forward recursion keeps each query's
 * running column in registers and only the per-query log-likelihoods leave the SMs (what a caller
 * needs when it evaluates parameters it will not propose moves from; mutation_move's whole-genome
 * proposal, mcmc_moves.c:1124-1134, is evaluated this way).  with_emission != 0 rebuilds the
 * emission table first and marks F and B as owed (see MCQ_OPT_LAZY_REFILL). */
int mcq_ctx_llk_only(mcq_ctx *ctx, int with_emission, double *llk);
/* One-off tuning of the sweep kernels for this shard (call after mcq_ctx_init_state): the full forward
 * sweep (stored and likelihood-only) and the full backward sweep are timed at every number of resident
 * blocks per SM the kernels allow, and the fastest is used for launches of that kernel with this many
 * queries from then on (process-wide).  Every block of a sweep lives for the whole sweep of its query,
 * so the residency decides how the last wave fills and how many write streams the HBM sees; the best
 * value depends on the shard size.  The state is left as it was.  blocks_per_sm (may be NULL) receives
 * the three choices.  Takes a few dozen sweeps. */
int mcq_ctx_tune(mcq_ctx *ctx, int blocks_per_sm[3]);
/* update_B_numerical_recipes(update = S-1) for every query. */
int mcq_ctx_backward_fill(mcq_ctx *ctx);

/* ---- MCMC proposals (the bodies of the for(q) loops in mcmc_moves.c) ------------------ */

enum {
  MCQ_PROP_REC = 0,  /* R[start] := value0; evaluates column start (mcmc_moves.c:213-224)          */
  MCQ_PROP_SEL = 1,  /* omega[start] := value0 (+ rate_out_of_codons, mcmc_moves.c:345-352),
                        emission BOTH/ALL at that site, evaluates column start (:366-395)           */
  MCQ_PROP_ESC = 2,  /* HLA_select_esc[hla][start..split-1] := value0, [split..end] := value1;
                        emission ESC for carriers of hla; evaluates start..end (:789-819)           */
  MCQ_PROP_REV = 3,  /* same on HLA_select_rev[0], emission REV for every query                     */
  MCQ_PROP_MU  = 4   /* mu := value0: rescale of every entry (:1082-1111) + full forward (:1124-1134) */
};

typedef struct {
  int kind;
  int start, end;   /* inclusive site range whose F columns are re-evaluated (REC/SEL: start==end) */
  int hla;          /* ESC only */
  int split;        /* first site taking value1; end+1 when the range takes one value */
  double value0, value1;
} mcq_proposal;

/* Evaluate a proposal into the tmp buffers; *sum_llk = total over the shard's queries (after
 * the all-reduce when a communicator is attached). Nothing current is modified. */
int mcq_ctx_propose(mcq_ctx *ctx, const mcq_proposal *prop, double *sum_llk);
/* Make the last proposal current: commit parameters, emission rows and F columns, refill F
 * after the range and B from the range down to 0 (mcmc_moves.c:261-279, 445-463, 955-981,
 * 1158-1181). */
int mcq_ctx_accept(mcq_ctx *ctx);
/* Drop the last proposal (nothing to undo on the device; kept for symmetry). */
int mcq_ctx_reject(mcq_ctx *ctx);
/* Complete every refill still owed (MCQ_OPT_LAZY_REFILL): afterwards F and B hold what the
 * reference's would after its last accept.  mcq_ctx_get_matrix does this itself. */
int mcq_ctx_flush(mcq_ctx *ctx);

/* Options. */
enum {
  MCQ_OPT_SPARSE_ESC = 1,  /* 1: an ESC proposal re-evaluates only carriers of that HLA and
                              reuses the cached per-query llk of the others (the reference
                              re-evaluates all, mcmc_moves.c:803; results agree to ~1e-13 rel) */
  MCQ_OPT_LAZY_REFILL = 2, /* default 1: after an accept the refills of F behind the range and of B
                              below it (mcmc_moves.c:266-279) are owed per query and completed by the
                              next evaluation that needs those columns, with the parameters then
                              current - the same values, computed when first read.  0: refill at
                              once, the reference's schedule. */
  MCQ_OPT_SHARED_PANEL = 3, /* default 1, effective once mcq_ctx_attach_peers has run: in
                              mcq_ctx_upload_static every rank DMAs 1/nranks of ref_triplets from the
                              host and reads the rest from its peers over NVLink.  All ranks must then
                              call mcq_ctx_upload_static together and with the same panel.  0: every
                              rank uploads the whole panel itself. */
  MCQ_OPT_ASYNC_UPLOAD = 4, /* default 0.  1: mcq_ctx_upload_static returns as soon as the caller's buffers
                              have been DMA'd (they may be reused); the device-side set-up runs on, chunk
                              by chunk, and the next mcq_ctx_full_llk evaluates each chunk of queries as
                              soon as it is ready, under the set-up of the next one.  A closest_refs entry
                              out of range is then reported by that evaluation, not by the upload.
                              2: the upload does not wait for the DMAs either (the cudaMemcpyAsync contract:
                              the caller keeps its page-locked buffers unchanged until the next
                              mcq_ctx_full_llk, evaluation or mcq_ctx_synchronize has returned). */
  MCQ_OPT_EXCHANGE = 6,     /* which exchange carries the shard totals when both are attached: 0 (default) the
                              peer mailboxes, 1 the NCCL all-reduce */
  MCQ_OPT_WARPS_PER_QUERY = 7, /* 0: the sweeps run one warp per query (bound by the HBM when the GPU holds many
                              queries).  W = 1, 2, 4 or 8 (NJ/2 or NJ for NJ = ceil(closest_n/64) groups of 64 copy
                              states): a block of W consumer warps + one warp that stages the rows owns a query,
                              which shortens the dependent instruction stream per site - for GPUs that hold few
                              queries.  Set it before mcq_ctx_init_state: the grouping of the per-site sums, and
                              with it the last bits of every value, follows the choice. */
  MCQ_OPT_PIPELINE_MAX_QUERIES = 8, /* a sweep launch that runs at most this many queries (the shard, or the carriers
                              of one HLA type) uses the software-pipelined kernels: with few warps per SM a sweep
                              costs the length of one warp's dependent chain per site, and there the look-ups of
                              the next site travel under the butterfly of the current one.  Same bits either way.
                              Default: see mcq_ctx.cu (measured cross-over); 0 = never. */
  MCQ_OPT_CHECKPOINT_STRIDE = 9, /* default 1: F, tmp_F and B are kept whole, as the reference keeps them.  k = 2, 4, 8,
                              ...: only every k-th column of each is kept; a column in between is recomputed from the
                              nearest one (at most k-1 sites of the recursion, 1 byte per cell) when a proposal needs
                              it.  A sweep then moves 1 + 8/k bytes per cell instead of 9 and the matrices take 1/k of
                              the memory; the values are bit for bit those of k = 1.  Set it before
                              mcq_ctx_upload_static (the matrices are allocated there). */
  MCQ_OPT_UPLOAD_CHUNK_QUERIES = 5 /* default 1250: mcq_ctx_upload_static transfers and sets up the per-query
                              data in up to 8 chunks of at least this many queries (one chunk = no
                              pipelining; smaller chunks starve the kernels that follow them). */
};
int mcq_ctx_set_option(mcq_ctx *ctx, int option, int value);

/* ---- inspection (tests, drop-in glue); outputs in the reference's layouts -------------- */
/* F, B, tmp_F: the whole matrix of a query, as the reference holds it (with MCQ_OPT_CHECKPOINT_STRIDE > 1 computed
 * afresh for the call).  F_KEPT / B_KEPT: only the columns the context really keeps, the others 0 (counters -1). */
enum { MCQ_BUF_F = 0, MCQ_BUF_B = 1, MCQ_BUF_TMP_F = 2, MCQ_BUF_F_KEPT = 3, MCQ_BUF_B_KEPT = 4 };
int mcq_ctx_get_matrix(mcq_ctx *ctx, int which, int q, double *out /*[L][S]*/, int *rnorm /*[S]*/);
enum { MCQ_TAB_C2C = 0, MCQ_TAB_A = 1, MCQ_TAB_TMP_C2C = 2, MCQ_TAB_TMP_A = 3 };
/* rows never written by the reference (codons absent from the site) come back as 0 */
int mcq_ctx_get_table(mcq_ctx *ctx, int which, int q, double *out /*[64][S]*/);
int mcq_ctx_get_query_llk(mcq_ctx *ctx, int proposed, double *out /*[Q]*/);
int mcq_ctx_get_params(mcq_ctx *ctx, double *mu, double *R, double *omega, double *rate_out,
                       double *esc, double *rev);
/* Bytes of device memory held by the context. */
size_t mcq_ctx_device_bytes(const mcq_ctx *ctx);
/* Device time (ms, CUDA events on the context's stream) and launch count of the kernels
 * issued by the previous API call on this context. */
int mcq_ctx_last_timing(const mcq_ctx *ctx, float *ms, int *launches);
/* When enabled every API call brackets its kernels with CUDA events (costs a sync).
 * on == 2 additionally brackets every kernel launch, accumulated per kernel class.
 * on == 3 brackets only the forward-kernel launches and adds no host wait to the calls: the pairs
 * are read out by the next mcq_ctx_kernel_time (what bench.py uses inside its timed region). */
int mcq_ctx_enable_timing(mcq_ctx *ctx, int on);
enum { MCQ_K_GATHER = 0, MCQ_K_EMISSION = 1, MCQ_K_FORWARD = 2, MCQ_K_BACKWARD = 3, MCQ_K_COMMIT = 4,
       MCQ_K_SUM = 5, MCQ_K_CLASSES = 6 };
/* Accumulated device time and launch count of one kernel class since the last reset
 * (needs mcq_ctx_enable_timing(ctx, 2)); reset != 0 clears all classes after reading. */
int mcq_ctx_kernel_time(mcq_ctx *ctx, int kclass, float *ms_total, int *launches, int reset);
/* Stopwatch on the context's stream (CUDA events): start records, stop records + waits. */
int mcq_ctx_timer_start(mcq_ctx *ctx);
int mcq_ctx_timer_stop(mcq_ctx *ctx, float *ms);
/* Wait for everything queued on the context's stream. */
int mcq_ctx_synchronize(mcq_ctx *ctx);
/* Page-lock / unlock a caller-owned host buffer so that uploads from it run at full PCIe rate. */
int mcq_host_register(void *ptr, size_t bytes);
int mcq_host_unregister(void *ptr);

/* ---- multi-GPU: queries are sharded over ranks, one scalar is exchanged ---------------- */
/* (a) NVLink peer exchange (preferred on one NVSwitch box): each rank exports a 64-byte CUDA IPC
 * handle of its mailbox (and of its copy of the reference panel, MCQ_OPT_SHARED_PANEL), the launcher
 * gathers them, every rank attaches all of them; the shard totals
 * are then exchanged by peer stores inside the reduction kernel and added in rank order.
 * `handles` is [nranks][64], entry `rank` may be anything. At most 16 ranks. */
int mcq_ctx_peer_handle(mcq_ctx *ctx, unsigned char handle[64]);
int mcq_ctx_attach_peers(mcq_ctx *ctx, int nranks, int rank, const unsigned char *handles);
/* (b) NCCL all-reduce of the one double.
 * 128-byte NCCL unique id, created on rank 0 and passed to the others by the launcher. */
int mcq_comm_unique_id(unsigned char id[128]);
int mcq_ctx_attach_comm(mcq_ctx *ctx, int nranks, int rank, const unsigned char id[128]);

/* Sum of one double per rank over the attached ranks, added in rank order (identical bits on every
 * rank); identity without peers.  The same exchange the likelihood totals use, for what the host chain
 * has to decide collectively (mcq_dmodel's wall-clock limit: rank 0's clock decides for everyone). */
int mcq_ctx_exchange_sum(mcq_ctx *ctx, double *value);

/* Substitution class of the codon pair (c1,c2), codons numbered base-4 over T,C,A,G: 0 none or
 * multi-step, 1 non-synonymous transition, 2 non-synonymous transversion, 3 synonymous transition,
 * 4 synonymous transversion - the content of NS_TS / NS_TV / S_TS / S_TV (constants.c:424-688). */
int mcq_codon_class(int c1, int c2);

/* ---- batched closest-L (closest_seqs.c over all queries; fill_closest_n, dmodel.c:356-387) -- */
/* ref_seqs: num_refs rows of num_bases ASCII bytes (row stride ref_stride), queries likewise.
 * skip_first = the -q self-reference mode (find closest_n+1, drop the first, dmodel.c:362,380-386).
 * Boundary ties are shuffled on the host with libc random() in query order, as the reference. */
int mcq_closest_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                      const char *query_seqs, size_t query_stride, int num_query,
                      int num_bases, int closest_n, int skip_first, int *closest_refs /*[Q][L]*/);
/* The same in two steps, for queries sharded over ranks.  mcq_closest_groups does the device part for the given
 * queries - per query the references closer than the distance T of sorted element closest_n, sorted by (distance,
 * index), and those at distance T in index order - and returns them as one flat int block (malloc'ed; mcq_free).
 * mcq_closest_replay does the sequential part for one block: the reference's tie shuffle (closest_seqs.c:57-64)
 * with libc random(), query after query.  Every rank computes the block of its shard, the blocks are exchanged, and
 * every rank replays all of them in rank order: the lists are then those of a single-process run. */
int mcq_closest_groups(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                       const char *query_seqs, size_t query_stride, int num_query, int num_bases,
                       int closest_n, int skip_first, int **groups, size_t *count);
int mcq_closest_replay(const int *groups, size_t count, int closest_n, int skip_first,
                       int *closest_refs /*[Q of the block][L]*/, int *num_query /* may be NULL */);
void mcq_free(void *p);
/* Per-query consensus of the closest references (generate_consensus_subset over every query,
 * dmodel.c:664-681 / amino_acids.c:118-141): cons_codes[Q][num_bases] receives base codes 0..3
 * (T,C,A,G), i.e. cons_query_nucls.  Reference sequences must still carry their unknown bases. */
int mcq_consensus_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                        const int *closest_refs, int num_query, int closest_n, int num_bases,
                        char *cons_codes);
/* Device time (CUDA events) of the distance kernel and of the two selection kernels in the last
 * mcq_closest_batch / mcq_hamming_batch call of the calling thread. */
int mcq_closest_last_timing(float *hamming_ms, float *select_ms);
/* Device time of the consensus kernel in the last mcq_consensus_batch call of the calling thread. */
int mcq_consensus_last_timing(float *kernel_ms);
/* Hamming distances only (values as closest_seqs.c:17-48), hamming[Q][N]. */
int mcq_hamming_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                      const char *query_seqs, size_t query_stride, int num_query,
                      int num_bases, int *hamming);

#ifdef __cplusplus
}
#endif
#endif /* MCQ_B200_H_ */
