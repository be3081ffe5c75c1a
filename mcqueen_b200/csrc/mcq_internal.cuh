// mcq_internal.cuh - device-side data layout and kernel launch interfaces of libmcq_b200.
//
// Layout in HBM (one context = one shard of Q queries on one GPU):
//   G      u8  [Q][S][Ls]   slot of the codon that copy state r shows at site s (0xFF = none/pad)
//   F,B,T  f64 [Q][S][Ls]   forward, backward, proposal("tmp") forward; SITE-major, state fastest
//   Fn,Bn,Tn i32 [Q][S]     rescale counters (F_rnorm etc. of the reference)
//   E,A,Et,At f64 [Q][ES]   emission / "A" rows, compact: site s owns slots eoff[s]..eoff[s]+K_s-1,
//                           slot k <-> codon slot_codon[eoff[s]+k]; ES = sum_s K_s (padded even)
//   llk_cur, llk_prop f64 [Q]
// Ls = L rounded up to 16 (128-byte rows).  A warp owns one query; lane l owns states
// 64*j + 2*l + {0,1}, j < NJ = ceil(Ls/64).
#ifndef MCQ_INTERNAL_CUH_
#define MCQ_INTERNAL_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define MCQ_BIG 1000000.0   // constants.c:13
#define MCQ_BIGI 100.0      // constants.c:14
#define MCQ_NOSLOT 0xFF

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
  int Q, S, L, Ls, N, H, ES;
  const uint8_t *G;          // [Q][S][Ls]
  const int *eoff;           // [S+1]
  const uint8_t *slot_codon; // [ES]  codon of each slot
  const uint8_t *slot_flags; // [ES]  bit0: codon is in ref_site_codons[s]; bit1: codon is the consensus
  const int *slot_site;      // [ES]  site each slot belongs to
  const uint8_t *cons_codon; // [S]
  const uint8_t *query_codon;// [Q][S]
  const uint8_t *hla;        // [Q][H] 0/1
};

// ---- proposal overrides applied on top of McqModel inside a kernel (no H2D copy per step)
struct McqOverride {
  int kind;                  // -1 none, else MCQ_PROP_*
  int start, end, hla, split;
  double v0, v1;
};

struct FwdArgs {
  McqStatic st;
  const double *E;           // emission rows read for sites start..end
  const double *R;
  int r_site; double r_value;     // R override at one site (-1 = none)
  const double *Fin; const int *Fn_in;
  double *Fout; int *Fn_out;
  int start, end;
  int llk_mode;              // 0 none, 1 dot with B at `end`, 2 plain column sum at `end`
  const double *B; const int *Bn;
  double *llk_out;           // [Q]
  const double *llk_keep;    // [Q] value copied to llk_out for inactive queries (may be null)
  const uint8_t *active;     // [Q][H] carrier table or null
  int active_hla;            // column of `active` to test
  double logBIG;
};

struct BwdArgs {
  McqStatic st;
  const double *E;
  const double *R;
  double *B; int *Bn;
  int update;                // start column (S-1 = fresh)
  const uint8_t *active; int active_hla;
};

struct EmisArgs {
  McqStatic st;
  McqModel m;
  McqOverride ov;
  int start, end, which_hla, mode;
  int k0, k1;                // slot bounds of the site range: eoff[start], eoff[end+1]
  const double *Esrc, *Asrc;
  double *Edst, *Adst;
};

// launchers (mcq_kernels.cu); all asynchronous on `stream`, return cudaGetLastError()
cudaError_t mcq_launch_tables_init();
cudaError_t mcq_launch_gather(const int8_t *ref_triplets, const int *closest, const int8_t *cons_nucls,
                              const uint8_t *slot_of /*[S][64]*/, uint8_t *G, int Q, int S, int L, int Ls,
                              cudaStream_t stream);
cudaError_t mcq_launch_forward(const FwdArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_backward(const BwdArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_emission(const EmisArgs &a, cudaStream_t stream);
cudaError_t mcq_launch_mu_rescale(const McqStatic &st, const McqModel &m, double mu_new,
                                  const double *E, const double *A, double *Eo, double *Ao, int nslots,
                                  cudaStream_t stream);
cudaError_t mcq_launch_commit_cols(const McqStatic &st, const double *T, const int *Tn, double *F, int *Fn,
                                   int start, int end, const uint8_t *active, int active_hla,
                                   cudaStream_t stream);
cudaError_t mcq_launch_commit_rows(const McqStatic &st, const double *Et, const double *At, double *E, double *A,
                                   int start, int end, cudaStream_t stream);
cudaError_t mcq_launch_commit_params(const McqStatic &st, McqModel m, McqOverride ov, cudaStream_t stream);
cudaError_t mcq_launch_sum(const double *v, int n, double *out, cudaStream_t stream);

#endif
