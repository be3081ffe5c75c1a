// mcq_dropin.cu - the three HMM symbols of the reference's forward_backward.c with their exact
// signatures (forward_backward.h:6-42), host pointers in and out, one query per call, running the
// same CUDA kernels as the batched context (Q = 1).  This is the link-compatible face of the
// library: the reference's dmodel builds against it unchanged.  It moves a query's matrices over
// PCIe on every call, so it is a correctness and migration path; throughput comes from the
// batched mcq_ctx_* API.
//
// Layout bridge: the reference's F[L][S] (state-major) <-> the device's [S][Ls] (site-major);
// codon_to_codon_HLA[64][S] <-> one dense emission row per site, E[s][64] (slot == codon + 1; the
// table degenerates to a single row per site because the caller's table is already per query).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "mcq_internal.cuh"
#include "../../include/mcq_b200.h"

namespace {

[[noreturn]] void die_cuda(const char *what, cudaError_t e, int line) {
  fflush(stdout);
  fprintf(stderr, "[mcq_dropin.cu:%d] Error: %s: %s\n", line, what, cudaGetErrorString(e));
  exit(EXIT_FAILURE);   // the reference's die() convention (util.c:54-73)
}
#define DK(x)                                           \
  do {                                                  \
    cudaError_t e_ = (x);                               \
    if (e_ != cudaSuccess) die_cuda(#x, e_, __LINE__);  \
  } while (0)

struct Scratch {
  int S = 0, L = 0, Ls = 0, Lg = 0;
  bool tables = false;
  cudaStream_t stream = nullptr;
  // device
  int8_t *trip = nullptr, *trip_pad = nullptr, *cons = nullptr;
  int *closest = nullptr, *aoff = nullptr, *Fn = nullptr, *Tn = nullptr, *d_err = nullptr;
  int2 *desc = nullptr;
  uint32_t *rowdesc = nullptr;
  uint8_t *slot_of = nullptr, *trip_tab = nullptr, *G = nullptr;
  double *E = nullptr, *R = nullptr, *F = nullptr, *T = nullptr;
  // host staging
  std::vector<int8_t> h_trip;
  std::vector<double> h_E, h_cols;
  std::vector<int> h_n;

  void release() {
    void *p[] = {trip, trip_pad, cons, closest, aoff, desc, rowdesc, Fn, Tn, slot_of, trip_tab, G, E, R, F, T};
    for (void *x : p)
      if (x) cudaFree(x);
    trip = trip_pad = cons = nullptr; closest = aoff = Fn = Tn = nullptr; desc = nullptr; rowdesc = nullptr; slot_of = trip_tab = G = nullptr;
    E = R = F = T = nullptr;
  }

  void ensure(int S_, int L_) {
    if (!tables) {
      int ndev = 0;
      cudaError_t e = cudaGetDeviceCount(&ndev);
      if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "libmcq_b200: no CUDA device available (there is no CPU fallback)\n");
        exit(EXIT_FAILURE);
      }
      DK(mcq_launch_tables_init());
      DK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      tables = true;
    }
    if (S_ == S && L_ == L) return;
    release();
    S = S_; L = L_; Ls = (L + 15) / 16 * 16;
    Lg = 64;
    while (Lg < Ls) Lg *= 2;
    DK(cudaMalloc((void **)&trip, (size_t)L * S));
    DK(cudaMalloc((void **)&trip_pad, (size_t)L * mcq_gather_pitch(S)));
    DK(cudaMalloc((void **)&cons, (size_t)3 * S));
    DK(cudaMalloc((void **)&closest, (size_t)L * sizeof(int)));
    DK(cudaMalloc((void **)&aoff, (size_t)(S + 1) * sizeof(int)));
    if (!d_err) DK(cudaMalloc((void **)&d_err, sizeof(int)));
    DK(cudaMalloc((void **)&Fn, (size_t)S * sizeof(int)));
    DK(cudaMalloc((void **)&Tn, (size_t)S * sizeof(int)));
    DK(cudaMalloc((void **)&slot_of, (size_t)S * 64));
    DK(cudaMalloc((void **)&trip_tab, (size_t)(S + 31) / 32 * 32 * 128));
    DK(cudaMalloc((void **)&G, (size_t)S * Lg));
    DK(cudaMalloc((void **)&E, (size_t)S * 64 * sizeof(double)));
    DK(cudaMalloc((void **)&R, (size_t)S * sizeof(double)));
    DK(cudaMalloc((void **)&F, (size_t)S * Ls * sizeof(double)));
    DK(cudaMalloc((void **)&T, (size_t)S * Ls * sizeof(double)));
    // identity slot scheme: slots 1..64 of every site are codons 0..63, slot 0 is unused
    std::vector<int> eo(S + 1);
    std::vector<uint8_t> so((size_t)S * 64);
    std::vector<int> cl(L);
    for (int s = 0; s <= S; s++) eo[s] = 64 * s;
    for (size_t i = 0; i < so.size(); i++) so[i] = (uint8_t)((i & 63) + 1);
    for (int r = 0; r < L; r++) cl[r] = r;
    DK(cudaMemcpy(aoff, eo.data(), eo.size() * sizeof(int), cudaMemcpyHostToDevice));
    {
      std::vector<int2> dh(S);
      for (int s = 0; s < S; s++) dh[s] = make_int2(64 * s, 64);
      DK(cudaMalloc((void **)&desc, (size_t)S * sizeof(int2)));
      DK(cudaMemcpy(desc, dh.data(), dh.size() * sizeof(int2), cudaMemcpyHostToDevice));
      std::vector<uint32_t> rh(S);
      for (int s = 0; s < S; s++) rh[s] = (64u << 25) | (uint32_t)(64 * s);
      DK(cudaMalloc((void **)&rowdesc, (size_t)S * sizeof(uint32_t)));
      DK(cudaMemcpy(rowdesc, rh.data(), rh.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    DK(cudaMemcpy(slot_of, so.data(), so.size(), cudaMemcpyHostToDevice));
    DK(cudaMemcpy(closest, cl.data(), cl.size() * sizeof(int), cudaMemcpyHostToDevice));
    DK(cudaMemset(F, 0, (size_t)S * Ls * sizeof(double)));
    DK(cudaMemset(T, 0, (size_t)S * Ls * sizeof(double)));
  }

  McqStatic statics(int N) const {
    McqStatic st{};
    st.Q = 1; st.S = S; st.L = L; st.Ls = Ls; st.Lg = Lg; st.N = N; st.H = 1;
    st.G = G; st.aoff = aoff; st.toff = aoff; st.desc = desc; st.rowdesc = rowdesc;   // one row of 64 entries per site
    st.query_codon = nullptr;                   // -> always row 0 of the site's block
    return st;
  }

  // inputs every call needs: the L reference rows the query copies from, its consensus codes, the
  // emission table and R
  void upload_common(const double *c2c, const char *ref_triplets, const double *Rh, const int *closest_refs,
                     const char *cons_nucls) {
    h_trip.resize((size_t)L * S);
    for (int r = 0; r < L; r++)
      memcpy(&h_trip[(size_t)r * S], ref_triplets + (size_t)closest_refs[r] * S, S);
    h_E.resize((size_t)S * 64);
    for (int c = 0; c < 64; c++)
      for (int s = 0; s < S; s++) h_E[(size_t)s * 64 + c] = c2c[(size_t)c * S + s];
    DK(cudaMemcpyAsync(trip, h_trip.data(), h_trip.size(), cudaMemcpyHostToDevice, stream));
    DK(cudaMemcpyAsync(cons, cons_nucls, (size_t)3 * S, cudaMemcpyHostToDevice, stream));
    DK(cudaMemcpyAsync(E, h_E.data(), h_E.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
    DK(cudaMemcpyAsync(R, Rh, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, stream));
    DK(mcq_launch_gather_prepare(trip, trip_pad, slot_of, trip_tab, S, L, stream));
    DK(mcq_launch_gather(trip_pad, trip, closest, cons, slot_of, G, nullptr, 1, S, L, Lg, L, 1, d_err, stream));
  }

  // one column host[L][S] (stride S) -> device column
  void put_column(double *dev, const double *host, int s) {
    h_cols.assign(Ls, 0.0);
    for (int r = 0; r < L; r++) h_cols[r] = host[(size_t)r * S + s];
    DK(cudaMemcpyAsync(dev + (size_t)s * Ls, h_cols.data(), Ls * sizeof(double), cudaMemcpyHostToDevice, stream));
    DK(cudaStreamSynchronize(stream));   // h_cols is reused
  }

  // device columns s0..s1 -> host[L][S]
  void get_columns(const double *dev, double *host, int s0, int s1) {
    const int W = s1 - s0 + 1;
    h_cols.resize((size_t)W * Ls);
    DK(cudaMemcpyAsync(h_cols.data(), dev + (size_t)s0 * Ls, h_cols.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
    DK(cudaStreamSynchronize(stream));
    for (int r = 0; r < L; r++)
      for (int s = s0; s <= s1; s++) host[(size_t)r * S + s] = h_cols[(size_t)(s - s0) * Ls + r];
  }

  void forward(int N, int start, int end, const double *Fin, const int *Fn_in, double *Fout, int *Fn_out) {
    FwdArgs a{};
    a.st = statics(N);
    a.em.tab = E; a.em.ec = nullptr; a.R = R; a.r_site = -1; a.r_value = 0.0;
    a.Fin = Fin; a.Fn_in = Fn_in; a.Fout = Fout; a.Fn_out = Fn_out;
    a.start = start; a.end = end; a.llk_mode = 0;
    a.active = nullptr; a.logBIG = log(MCQ_BIG);
    DK(mcq_launch_forward(a, stream));
  }
};

Scratch g_s;

void check_dims(int S, int L, int N) {
  if (S <= 0 || L <= 0 || N <= 0 || L > 1024) {
    fprintf(stderr, "libmcq_b200: unsupported dimensions num_sites=%d closest_n=%d total_num_refs=%d\n", S, L, N);
    exit(EXIT_FAILURE);
  }
}

}  // namespace

extern "C" void update_F_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                           double *codon_to_codon_HLA, char *ref_codons,
                                           char *ref_triplets, double *R, double *F_in, int *F_rnorm_in,
                                           int *F_rnorm_out, double *F_out, int start_site, int end_site,
                                           int *closest_refs, int num_bases, char *cons_query_nucls) {
  (void)ref_codons; (void)num_bases;   // USE_TRIPLET is hard-defined in the reference (forward_backward.h:4)
  const int S = num_sites, L = closest_n;
  check_dims(S, L, total_num_refs);
  Scratch &k = g_s;
  k.ensure(S, L);
  k.upload_common(codon_to_codon_HLA, ref_triplets, R, closest_refs, cons_query_nucls);
  if (start_site > 0) {
    k.put_column(k.F, F_in, start_site - 1);
    DK(cudaMemcpyAsync(k.Fn + start_site - 1, F_rnorm_in + start_site - 1, sizeof(int), cudaMemcpyHostToDevice, k.stream));
  }
  if (start_site > end_site && start_site > 0) return;   // empty loop in the reference
  k.forward(total_num_refs, start_site, end_site, k.F, k.Fn, k.T, k.Tn);
  const int first = start_site;                          // column 0 is written when start_site == 0
  const int last = end_site < first ? first : end_site;
  if (start_site == 0) F_rnorm_in[0] = 0;                // forward_backward.c:109
  k.get_columns(k.T, F_out, first, start_site == 0 && end_site < 0 ? 0 : last);
  DK(cudaMemcpy(F_rnorm_out + first, k.Tn + first, (size_t)(last - first + 1) * sizeof(int), cudaMemcpyDeviceToHost));
}

extern "C" void initialise_F_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                               double *codon_to_codon_HLA, char *ref_codons,
                                               char *ref_triplets, double *R, double *F, int *F_rnorm,
                                               int start_site, int end_site, int *closest_refs,
                                               int num_bases, char *cons_query_nucls) {
  // forward_backward.c:11-76: column 0 is always seeded from site 0; the loop runs start_site+1..end_site.
  // Every caller passes (0, S-1) (dmodel.c:887-891).
  update_F_numerical_recipes(num_sites, closest_n, total_num_refs, codon_to_codon_HLA, ref_codons, ref_triplets,
                             R, F, F_rnorm, F_rnorm, F, 0, start_site == 0 ? end_site : 0, closest_refs,
                             num_bases, cons_query_nucls);
  if (start_site != 0 && start_site + 1 <= end_site)
    update_F_numerical_recipes(num_sites, closest_n, total_num_refs, codon_to_codon_HLA, ref_codons, ref_triplets,
                               R, F, F_rnorm, F_rnorm, F, start_site + 1, end_site, closest_refs,
                               num_bases, cons_query_nucls);
}

extern "C" void update_B_numerical_recipes(int num_sites, int closest_n, int total_num_refs,
                                           double *codon_to_codon_HLA, char *ref_codons,
                                           char *ref_triplets, double *R, double *B, int *B_rnorm,
                                           int update, int *closest_refs, int num_bases,
                                           char *cons_query_nucls) {
  (void)ref_codons; (void)num_bases;
  const int S = num_sites, L = closest_n;
  check_dims(S, L, total_num_refs);
  Scratch &k = g_s;
  k.ensure(S, L);
  k.upload_common(codon_to_codon_HLA, ref_triplets, R, closest_refs, cons_query_nucls);
  if (update < S - 1) {
    k.put_column(k.F, B, update + 1);
    DK(cudaMemcpyAsync(k.Fn + update + 1, B_rnorm + update + 1, sizeof(int), cudaMemcpyHostToDevice, k.stream));
  }
  BwdArgs a{};
  a.st = k.statics(total_num_refs);
  a.em.tab = k.E; a.em.ec = nullptr; a.R = k.R; a.B = k.F; a.Bn = k.Fn; a.update = update; a.active = nullptr;
  DK(mcq_launch_backward(a, k.stream));
  const int top = update;   // update == S-1 rewrites the last column too
  k.get_columns(k.F, B, 0, top);
  DK(cudaMemcpy(B_rnorm, k.Fn, (size_t)(top + 1) * sizeof(int), cudaMemcpyDeviceToHost));
}
