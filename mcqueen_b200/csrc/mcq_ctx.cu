// mcq_ctx.cu - the batched context API of include/mcq_b200.h: device memory, static-table
// construction, parameter upload, proposal / accept sequencing, inspection, NCCL all-reduce.
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mcq_internal.cuh"
#include "../../include/mcq_b200.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

int mcq_fail(const char *file, int line, const char *what) {
  char buf[512];
  snprintf(buf, sizeof(buf), "[%s:%d] %s", file, line, what);
  g_err = buf;
  return 1;
}
#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) return mcq_fail(__FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)
#define REQUIRE(c, msg)                                         \
  do {                                                          \
    if (!(c)) return mcq_fail(__FILE__, __LINE__, msg);         \
  } while (0)

extern "C" const char *mcq_last_error(void) { return g_err.c_str(); }
extern "C" int mcq_abi_version(void) { return 1; }
extern "C" int mcq_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    mcq_fail(__FILE__, __LINE__, cudaGetErrorString(e));
    return -1;
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU users need no NCCL at all
// ------------------------------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return mcq_fail(__FILE__, __LINE__, "cannot dlopen libnccl.so.2");
  g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return mcq_fail(__FILE__, __LINE__, "libnccl lacks expected symbols");
  return 0;
}
static const int kNcclDouble = 8;  // ncclFloat64
static const int kNcclSum = 0;     // ncclSum

// ------------------------------------------------------------------------------------------
// the context
// ------------------------------------------------------------------------------------------
static const int kPipeStreams = 3;    // chunks of a pipelined evaluation in flight at once
static const int kUploadChunks = 8;   // query chunks of mcq_ctx_upload_static (transfer under gather)

struct mcq_ctx {
  mcq_dims d;
  int Ls = 0, Lg = 0, NA = 0;          // NA = number of non-consensus slots (sum_s K_s)
  size_t TL = 0;               // length of the shared emission table (64 * NA)
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;            // the backward refill of an accept runs beside the forward refill
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_chunk[1 + kUploadChunks] = {nullptr};   // upload_static: panel landed, chunk k landed
  // MCQ_OPT_ASYNC_UPLOAD: mcq_ctx_upload_static returns once the DMAs are done; the chunks' gathers are
  // still running on the main stream, ev_ready[k] marks chunk k complete, and the next mcq_ctx_full_llk
  // evaluates chunk by chunk on stream3 behind them (pipe_n chunks, queries pipe_q[k] .. pipe_q[k+1]-1)
  cudaStream_t stream_p = nullptr;  // parameters travelling beside an asynchronous upload
  cudaStream_t stream3 = nullptr;   // the pipelined evaluation: stream3 and pipe_st[] in turn
  cudaStream_t pipe_st[kPipeStreams - 1] = {nullptr};
  cudaEvent_t ev_ps[kPipeStreams - 1] = {nullptr};
  cudaEvent_t ev_ready[kUploadChunks] = {nullptr}, ev_params = nullptr, ev_static = nullptr, ev_pipe = nullptr;
  int pipe_n = 0, pipe_q[kUploadChunks + 1] = {0};
  bool params_on_copy_stream = false;
  int opt_async_upload = 0;
  int opt_chunk_queries = 1250;   // MCQ_OPT_UPLOAD_CHUNK_QUERIES
  size_t bytes = 0;
  bool have_static = false, have_params = false, have_state = false;
  bool have_fwd = false;     // a full forward sweep has run since the last upload: with a backward fill it makes a state

  // static (device)
  uint8_t *G = nullptr, *ncodon = nullptr, *cons_in_list = nullptr, *cons_codon = nullptr;
  uint8_t *query_codon = nullptr, *hla = nullptr;
  int *aoff = nullptr, *toff = nullptr, *nsite = nullptr;
  int2 *desc = nullptr;
  uint32_t *rowdesc = nullptr;
  // static (host mirrors)
  std::vector<int> aoff_h, toff_h;
  std::vector<uint8_t> ncodon_h, cons_in_list_h, cons_codon_h, slot_of_h;

  McqModel m{};
  // F and T are the two forward column buffers (buffer 0 / buffer 1); sel[q][s] names the one that holds
  // the current column, the other one receives proposals (an accept flips selectors, nothing is copied)
  double *F = nullptr, *B = nullptr, *T = nullptr;
  int *Fn = nullptr, *Bn = nullptr, *Tn = nullptr;
  uint8_t *sel = nullptr;
  // emission rows: shared table + per-query consensus values, current and proposal ("tmp") sets
  double *tab = nullptr, *atab = nullptr, *ec = nullptr, *ac = nullptr;
  double *tab_t = nullptr, *atab_t = nullptr, *ec_t = nullptr, *ac_t = nullptr;
  double *llk_cur = nullptr, *llk_prop = nullptr, *d_sum = nullptr, *d_xchg = nullptr;
  int *fv = nullptr, *bv = nullptr;   // per query: F valid up to fv, B valid from bv (lazy refills)
  double *h_sum = nullptr;   // pinned
  double *pblock = nullptr, *h_pblock = nullptr;   // parameter block (device) and its pinned staging copy
  size_t pblock_n = 0;
  McqHostSlot *h_slot = nullptr, *d_slot = nullptr;   // host-mapped result slot and its device view
  unsigned long long eval_seq = 0;

  McqOverride pending{};
  bool has_pending = false;
  int opt_sparse_esc = 0;
  int opt_lazy = 1;          // MCQ_OPT_LAZY_REFILL
  // host mirror of the per-query validity marks fv / bv (every update is issued from the host, so the
  // mirror is exact) and of the 0/1 carrier table: a catch-up is launched only when one is really owed
  std::vector<int> fv_h, bv_h;
  std::vector<uint8_t> hla_h;

  int timing = 0;            // 0 off, 1 per API call, 2 per kernel launch as well, 3 forward launches only (read later)
  std::vector<cudaEvent_t> fev;      // level 3: pairs bracketing forward launches, resolved by mcq_ctx_kernel_time
  size_t fev_used = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evt0 = nullptr, evt1 = nullptr;
  float last_ms = 0.f;
  int last_launches = 0, launches = 0;
  std::vector<cudaEvent_t> kev;      // pairs of events bracketing single launches
  std::vector<int> kev_class;
  float k_ms[MCQ_K_CLASSES] = {0};
  int k_n[MCQ_K_CLASSES] = {0};

  // staging buffers of mcq_ctx_upload_static (kept so that a re-upload allocates nothing)
  int8_t *st_trip = nullptr, *st_trip_pad = nullptr, *st_cons = nullptr;
  int *st_closest = nullptr;
  uint8_t *st_slot_of = nullptr, *st_trip_tab = nullptr;

  ncclComm_t comm = nullptr;
  int nranks = 1;
  // NVLink peer mailboxes (mcq_ctx_attach_peers)
  struct {
    double *value[16];
    unsigned long long *seq[16];
    unsigned long long *bar[16];
    int nranks, rank;
  } box{};
  void *mailbox = nullptr;          // this rank's own mailbox allocation (+ the shared panel behind it)
  bool trip_in_mailbox = false;     // st_trip lives behind the mailbox: peers read their slices of the panel from it
  const int8_t *peer_trip[16] = {nullptr};
  unsigned long long bar_seq = 0;
  int opt_shared_panel = 1;         // MCQ_OPT_SHARED_PANEL
  int opt_exchange = 0;             // MCQ_OPT_EXCHANGE
  int pipe_max = 0;                 // launches that run at most this many queries use the pipelined sweep kernels
  std::vector<int> carriers;        // queries carrying each HLA type (host count)
  int ck = 1, nc = 0;               // MCQ_OPT_CHECKPOINT_STRIDE: every ck-th column of F / T / B is kept; nc columns per query
  double *Tcol = nullptr, *Bcol = nullptr;   // ck > 1: the proposal's last column / the backward column at it, per query
  int *Tncol = nullptr, *Bncol = nullptr;
  const double *pend_tab = nullptr, *pend_ec = nullptr;   // the rows the pending proposal was evaluated with
  int pend_rsite = -1; double pend_rvalue = 0.0;
  int wpq = 0;                      // MCQ_OPT_WARPS_PER_QUERY: 0 = the one-warp sweeps, else consumer warps per query
  void *peer_map[16] = {nullptr};   // IPC mappings of the peers' mailboxes
  bool have_peers = false;
  unsigned long long step = 0;
  int *d_err = nullptr, *h_err = nullptr;
  double logBIG = 0.0;
};

template <typename T>
static int dev_alloc(mcq_ctx *c, T **p, size_t n) {
  if (n == 0) n = 1;
  CK(cudaMalloc((void **)p, n * sizeof(T)));
  c->bytes += n * sizeof(T);
  return 0;
}
#define ALLOC(ptr, n)                          \
  do {                                         \
    if (dev_alloc(c, &(ptr), (n))) return 1;   \
  } while (0)

static McqStatic make_static(const mcq_ctx *c) {
  McqStatic st;
  st.Q = c->d.num_query; st.S = c->d.num_sites; st.L = c->d.closest_n; st.Ls = c->Ls; st.Lg = c->Lg;
  st.N = c->d.total_num_refs; st.H = c->d.num_hla_types;
  st.G = c->G; st.aoff = c->aoff; st.toff = c->toff; st.ncodon = c->ncodon; st.nsite = c->nsite;
  st.cons_in_list = c->cons_in_list; st.desc = c->desc; st.rowdesc = c->rowdesc;
  st.cons_codon = c->cons_codon; st.query_codon = c->query_codon; st.hla = c->hla;
  return st;
}

struct CallScope {   // brackets one API call for mcq_ctx_last_timing
  mcq_ctx *c;
  explicit CallScope(mcq_ctx *ctx) : c(ctx) {
    c->launches = 0;
    c->kev_class.clear();
    if (c->timing == 1 || c->timing == 2) cudaEventRecord(c->ev0, c->stream);
  }
  ~CallScope() {
    c->last_launches = c->launches;
    if (c->timing == 1 || c->timing == 2) {
      cudaEventRecord(c->ev1, c->stream);
      cudaEventSynchronize(c->ev1);
      cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1);
      for (size_t i = 0; i < c->kev_class.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->kev[2 * i], c->kev[2 * i + 1]);
        c->k_ms[c->kev_class[i]] += ms;
      }
    }
  }
};

struct KScope {      // brackets one kernel launch (timing level 2; level 3: forward launches only) and counts it
  mcq_ctx *c;
  int idx = -1;
  cudaEvent_t late = nullptr;
  KScope(mcq_ctx *ctx, int kclass) : c(ctx) {
    c->launches++;
    c->k_n[kclass]++;
    if (c->timing == 3 && kclass == MCQ_K_FORWARD) {
      // no host wait in the call: the pair is read out when the caller asks for the class total
      while (c->fev.size() < c->fev_used + 2) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->fev.push_back(e);
      }
      cudaEventRecord(c->fev[c->fev_used], c->stream);
      late = c->fev[c->fev_used + 1];
      c->fev_used += 2;
    }
    if (c->timing == 2) {
      idx = (int)c->kev_class.size();
      while (c->kev.size() < (size_t)2 * (idx + 1)) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->kev.push_back(e);
      }
      c->kev_class.push_back(kclass);
      cudaEventRecord(c->kev[2 * idx], c->stream);
    }
  }
  ~KScope() {
    if (idx >= 0) cudaEventRecord(c->kev[2 * idx + 1], c->stream);
    if (late) cudaEventRecord(late, c->stream);
  }
};

extern "C" int mcq_ctx_create(mcq_ctx **out, const mcq_dims *dims) {
  REQUIRE(out && dims, "null argument");
  REQUIRE(dims->num_sites > 0 && dims->closest_n > 0 && dims->total_num_refs > 0 && dims->num_query > 0 &&
              dims->num_hla_types > 0, "dimensions must be positive");
  REQUIRE(dims->closest_n <= 1024, "closest_n > 1024 is not supported");
  REQUIRE((double)dims->total_num_refs * dims->num_sites < 4.0e9, "total_num_refs * num_sites must stay below 4e9");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(ndev > 0, "no CUDA device: libmcq_b200 has no CPU fallback");
  REQUIRE(dims->device >= 0 && dims->device < ndev, "device ordinal out of range");
  CK(cudaSetDevice(dims->device));
  static bool tables_ready[64] = {false};
  if (!tables_ready[dims->device]) {
    CK(mcq_launch_tables_init());
    tables_ready[dims->device] = true;
  }
  mcq_ctx *c = new mcq_ctx();
  c->d = *dims;
  c->Ls = (dims->closest_n + 15) / 16 * 16;
  c->Lg = 64;
  while (c->Lg < c->Ls) c->Lg *= 2;   // 64 * NJ, NJ in {1,2,4,8,16}
  c->logBIG = log(MCQ_BIG);   // the host libm value the reference itself uses (mcmc_moves.c:223)
  if (getenv("MCQ_PIPE_MAX")) c->pipe_max = atoi(getenv("MCQ_PIPE_MAX"));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (cudaEvent_t &e : c->ev_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (cudaEvent_t &e : c->ev_ready) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_params, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_static, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_pipe, cudaEventDisableTiming));
  {
    // the pipelined evaluation outranks the set-up kernels still queued on the main stream: its long-lived
    // one-warp blocks are dispatched first and the set-up fills the rest of the SMs
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithFlags(&c->stream_p, cudaStreamNonBlocking));
    for (int i = 0; i < kPipeStreams - 1; i++) {
      CK(cudaStreamCreateWithPriority(&c->pipe_st[i], cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&c->ev_ps[i], cudaEventDisableTiming));
    }
  }
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaEventCreate(&c->evt0));
  CK(cudaEventCreate(&c->evt1));
  const size_t S = dims->num_sites, Q = dims->num_query, H = dims->num_hla_types;
  // the parameter arrays are one block, mirrored by a pinned staging block: a full mcq_ctx_set_params is
  // one DMA.  Order: prior[64] R[S] omega[S] rev[S] esc[H][S] rate_out[S][64]
  c->pblock_n = 64 + 3 * S + H * S + 64 * S;
  ALLOC(c->pblock, c->pblock_n);
  CK(cudaMallocHost((void **)&c->h_pblock, c->pblock_n * sizeof(double)));
  c->m.prior = c->pblock; c->m.R = c->m.prior + 64; c->m.omega = c->m.R + S; c->m.rev = c->m.omega + S;
  c->m.esc = c->m.rev + S; c->m.rate_out = c->m.esc + H * S;
  ALLOC(c->llk_cur, Q); ALLOC(c->llk_prop, Q); ALLOC(c->d_sum, 1); ALLOC(c->d_xchg, 1);
  ALLOC(c->fv, Q); ALLOC(c->bv, Q);
  CK(cudaMallocHost((void **)&c->h_sum, sizeof(double)));
  CK(cudaMallocHost((void **)&c->h_err, sizeof(int)));
  *c->h_err = 0;
  CK(cudaHostAlloc((void **)&c->h_slot, sizeof(McqHostSlot), cudaHostAllocMapped));
  memset(c->h_slot, 0, sizeof(McqHostSlot));
  CK(cudaHostGetDevicePointer((void **)&c->d_slot, c->h_slot, 0));
  ALLOC(c->d_err, 1);
  CK(cudaMemset(c->d_err, 0, sizeof(int)));
  c->pending.kind = -1;
  *out = c;
  return 0;
}

extern "C" void mcq_ctx_destroy(mcq_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->d.device);
  cudaStreamSynchronize(c->stream);
  void *ptrs[] = {c->desc, c->rowdesc, c->nsite, c->G, c->ncodon, c->cons_in_list, c->cons_codon, c->query_codon, c->hla, c->aoff, c->toff,
                  c->pblock,
                  c->F, c->B, c->T, c->Fn, c->Bn, c->Tn, c->tab, c->atab, c->ec, c->ac, c->tab_t, c->atab_t, c->ec_t, c->ac_t,
                  c->llk_cur, c->llk_prop, c->d_sum, c->d_xchg, c->fv, c->bv, c->sel, c->Tcol, c->Bcol, c->Tncol, c->Bncol, c->st_trip, c->st_trip_pad, c->st_cons, c->st_closest, c->st_slot_of, c->st_trip_tab};
  for (void *p : ptrs)
    if (p && !(p == (void *)c->st_trip && c->trip_in_mailbox)) cudaFree(p);
  if (c->h_sum) cudaFreeHost(c->h_sum);
  if (c->h_pblock) cudaFreeHost(c->h_pblock);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->h_slot) cudaFreeHost(c->h_slot);
  if (c->d_err) cudaFree(c->d_err);
  for (int r = 0; r < 16; r++)
    if (c->peer_map[r]) cudaIpcCloseMemHandle(c->peer_map[r]);
  if (c->mailbox) cudaFree(c->mailbox);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->evt0);
  cudaEventDestroy(c->evt1);
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->fev) cudaEventDestroy(e);
  cudaStreamSynchronize(c->stream2);
  for (cudaEvent_t e : c->ev_chunk) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_ready) if (e) cudaEventDestroy(e);
  if (c->ev_params) cudaEventDestroy(c->ev_params);
  if (c->ev_static) cudaEventDestroy(c->ev_static);
  if (c->ev_pipe) cudaEventDestroy(c->ev_pipe);
  if (c->stream3) { cudaStreamSynchronize(c->stream3); cudaStreamDestroy(c->stream3); }
  if (c->stream_p) { cudaStreamSynchronize(c->stream_p); cudaStreamDestroy(c->stream_p); }
  for (int i = 0; i < kPipeStreams - 1; i++) {
    if (c->pipe_st[i]) { cudaStreamSynchronize(c->pipe_st[i]); cudaStreamDestroy(c->pipe_st[i]); }
    if (c->ev_ps[i]) cudaEventDestroy(c->ev_ps[i]);
  }
  cudaEventDestroy(c->ev_fork);
  cudaEventDestroy(c->ev_join);
  cudaStreamDestroy(c->stream2);
  cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" size_t mcq_ctx_device_bytes(const mcq_ctx *c) { return c ? c->bytes : 0; }

extern "C" int mcq_ctx_enable_timing(mcq_ctx *c, int on) {
  REQUIRE(c, "null context");
  c->timing = on;
  return 0;
}
extern "C" int mcq_ctx_kernel_time(mcq_ctx *c, int kclass, float *ms_total, int *launches, int reset) {
  REQUIRE(c, "null context");
  REQUIRE(kclass >= 0 && kclass < MCQ_K_CLASSES, "kernel class");
  if (c->fev_used) {           // level 3: the deferred forward pairs
    CK(cudaSetDevice(c->d.device));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < c->fev_used; i += 2) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, c->fev[i], c->fev[i + 1]));
      c->k_ms[MCQ_K_FORWARD] += ms;
    }
    c->fev_used = 0;
  }
  if (ms_total) *ms_total = c->k_ms[kclass];
  if (launches) *launches = c->k_n[kclass];
  if (reset)
    for (int i = 0; i < MCQ_K_CLASSES; i++) { c->k_ms[i] = 0.f; c->k_n[i] = 0; }
  return 0;
}
extern "C" int mcq_ctx_timer_start(mcq_ctx *c) {
  REQUIRE(c, "null context");
  CK(cudaSetDevice(c->d.device));
  CK(cudaEventRecord(c->evt0, c->stream));
  return 0;
}
extern "C" int mcq_ctx_timer_stop(mcq_ctx *c, float *ms) {
  REQUIRE(c && ms, "null argument");
  CK(cudaSetDevice(c->d.device));
  CK(cudaEventRecord(c->evt1, c->stream));
  CK(cudaEventSynchronize(c->evt1));
  CK(cudaEventElapsedTime(ms, c->evt0, c->evt1));
  return 0;
}
extern "C" int mcq_ctx_synchronize(mcq_ctx *c) {
  REQUIRE(c, "null context");
  CK(cudaSetDevice(c->d.device));
  CK(cudaStreamSynchronize(c->stream2));      // the DMAs of an asynchronous upload
  CK(cudaStreamSynchronize(c->stream_p));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int mcq_host_register(void *ptr, size_t bytes) {
  REQUIRE(ptr && bytes, "null argument");
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return 0;
}
extern "C" int mcq_host_unregister(void *ptr) {
  REQUIRE(ptr, "null argument");
  CK(cudaHostUnregister(ptr));
  return 0;
}
extern "C" int mcq_ctx_last_timing(const mcq_ctx *c, float *ms, int *launches) {
  REQUIRE(c, "null context");
  if (ms) *ms = c->last_ms;
  if (launches) *launches = c->last_launches;
  return 0;
}
extern "C" int mcq_ctx_set_option(mcq_ctx *c, int option, int value) {
  REQUIRE(c, "null context");
  if (option == MCQ_OPT_SPARSE_ESC) { c->opt_sparse_esc = value; return 0; }
  if (option == MCQ_OPT_SHARED_PANEL) { c->opt_shared_panel = value; return 0; }
  if (option == MCQ_OPT_ASYNC_UPLOAD) { c->opt_async_upload = value; return 0; }
  if (option == MCQ_OPT_EXCHANGE) { c->opt_exchange = value; return 0; }
  if (option == MCQ_OPT_PIPELINE_MAX_QUERIES) { c->pipe_max = value; return 0; }
  if (option == MCQ_OPT_CHECKPOINT_STRIDE) {
    REQUIRE(value >= 1 && value <= 64 && (value & (value - 1)) == 0, "the checkpoint stride is a power of two, at most 64");
    REQUIRE(!c->have_static, "set MCQ_OPT_CHECKPOINT_STRIDE before mcq_ctx_upload_static (the matrices are allocated there)");
    REQUIRE(value == 1 || c->wpq == 0, "checkpointed matrices go with the one-warp sweeps");
    c->ck = value;
    return 0;
  }
  if (option == MCQ_OPT_WARPS_PER_QUERY) {
    REQUIRE(value == 0 || mcq_mw_supported(c->Lg / 64, value), "this number of warps per query is not built for this closest_n");
    REQUIRE(value == 0 || c->ck == 1, "checkpointed matrices go with the one-warp sweeps");
    // the grouping of the per-site column sum changes with it: F, B and the cached per-query values of one state
    // must all come from one grouping
    REQUIRE(!c->have_state || value == c->wpq, "set MCQ_OPT_WARPS_PER_QUERY before mcq_ctx_init_state");
    c->wpq = value;
    return 0;
  }
  if (option == MCQ_OPT_UPLOAD_CHUNK_QUERIES) {
    REQUIRE(value >= 1, "chunk size must be positive");
    c->opt_chunk_queries = value;
    return 0;
  }
  if (option == MCQ_OPT_LAZY_REFILL) {
    if (!value && c->have_state && mcq_ctx_flush(c)) return 1;   // leave lazy mode with F and B complete
    c->opt_lazy = value;
    return 0;
  }
  return mcq_fail(__FILE__, __LINE__, "unknown option");
}

// ------------------------------------------------------------------------------------------
// static data
// ------------------------------------------------------------------------------------------
extern "C" int mcq_ctx_upload_static(mcq_ctx *c, const char *ref_triplets, const int *closest_refs,
                                     const char *cons_query_nucls, const char *query_codons,
                                     const char *hla_types, const char *consensus_codons,
                                     const int *num_ref_site_codons, const char *ref_site_codons) {
  REQUIRE(c, "null context");
  REQUIRE(ref_triplets && closest_refs && cons_query_nucls && query_codons && hla_types &&
              consensus_codons && num_ref_site_codons && ref_site_codons, "null argument");
  CK(cudaSetDevice(c->d.device));
  const int S = c->d.num_sites, Q = c->d.num_query, L = c->d.closest_n, N = c->d.total_num_refs,
            H = c->d.num_hla_types;
  CallScope scope(c);

  // slots per site: 0 = the consensus codon, 1..K_s = the other codons the panel shows at the
  // site in ascending order (amino_acids.c:155-176); everything else has no slot
  std::vector<int> aoff_h(S + 1, 0), toff_h(S + 1, 0);
  std::vector<uint8_t> ncodon_h, cons_in_list_h(S, 0), cons_codon_h(S, 0);
  std::vector<uint8_t> slot_of((size_t)S * 64, MCQ_NOSLOT);
  for (int s = 0; s < S; s++) {
    const int n = num_ref_site_codons[s];
    REQUIRE(n >= 0 && n <= 64, "num_ref_site_codons out of range");
    const int cons = (unsigned char)consensus_codons[s];
    REQUIRE(cons < 64, "consensus codon out of range");
    cons_codon_h[s] = (uint8_t)cons;
    slot_of[(size_t)s * 64 + cons] = 0;
    int k = 0;
    for (int i = 0; i < n; i++) {
      const int cd = (unsigned char)ref_site_codons[(size_t)s * 64 + i];
      REQUIRE(cd < 64, "reference codon out of range");
      if (cd == cons) {
        REQUIRE(!cons_in_list_h[s], "duplicate codon in ref_site_codons");
        cons_in_list_h[s] = 1;
        continue;
      }
      REQUIRE(slot_of[(size_t)s * 64 + cd] == MCQ_NOSLOT, "duplicate codon in ref_site_codons");
      slot_of[(size_t)s * 64 + cd] = (uint8_t)(1 + k++);
      ncodon_h.push_back((uint8_t)cd);
    }
    aoff_h[s + 1] = aoff_h[s] + k;
    toff_h[s + 1] = toff_h[s] + 64 * k;
  }

  if (!c->have_static) {
    c->aoff_h = aoff_h; c->toff_h = toff_h; c->ncodon_h = ncodon_h; c->cons_in_list_h = cons_in_list_h;
    c->cons_codon_h = cons_codon_h; c->slot_of_h = slot_of;
    c->NA = aoff_h[S];
    c->TL = (size_t)toff_h[S];
    REQUIRE(c->TL <= 0x1FFFFFFu, "emission table too large for the packed row descriptors (sum of 64*K_s must stay below 2^25)");
    ALLOC(c->aoff, (size_t)S + 1);
    ALLOC(c->toff, (size_t)S + 1);
    ALLOC(c->ncodon, (size_t)c->NA);
    ALLOC(c->nsite, (size_t)c->NA);
    ALLOC(c->cons_in_list, (size_t)S);
    ALLOC(c->cons_codon, (size_t)S);
    ALLOC(c->query_codon, (size_t)Q * S);
    ALLOC(c->rowdesc, (size_t)Q * S);
    ALLOC(c->hla, (size_t)Q * H);
    ALLOC(c->G, (size_t)Q * S * c->Lg);
    if (!c->st_trip) ALLOC(c->st_trip, (size_t)N * S);
    ALLOC(c->st_trip_pad, (size_t)N * mcq_gather_pitch(S));
    ALLOC(c->st_cons, (size_t)Q * 3 * S);
    ALLOC(c->st_closest, (size_t)Q * L);
    ALLOC(c->st_slot_of, (size_t)S * 64);
    ALLOC(c->st_trip_tab, (size_t)(S + 31) / 32 * 32 * 128);
    // the forward / proposal / backward matrices: every ck-th column (all of them with ck = 1)
    c->nc = (S + c->ck - 1) / c->ck;
    const size_t ncol = (size_t)Q * c->nc, cells = ncol * c->Ls;
    ALLOC(c->F, cells); ALLOC(c->B, cells); ALLOC(c->T, cells);
    ALLOC(c->Fn, ncol); ALLOC(c->Bn, ncol); ALLOC(c->Tn, ncol);
    ALLOC(c->sel, ncol);
    CK(cudaMemsetAsync(c->sel, 0, ncol, c->stream));
    if (c->ck > 1) {
      ALLOC(c->Tcol, (size_t)Q * c->Ls); ALLOC(c->Bcol, (size_t)Q * c->Ls);
      ALLOC(c->Tncol, (size_t)Q); ALLOC(c->Bncol, (size_t)Q);
    }
    ALLOC(c->tab, c->TL); ALLOC(c->tab_t, c->TL); ALLOC(c->atab, (size_t)c->NA); ALLOC(c->atab_t, (size_t)c->NA);
    const size_t qs = (size_t)Q * S;
    ALLOC(c->ec, qs); ALLOC(c->ac, qs); ALLOC(c->ec_t, qs); ALLOC(c->ac_t, qs);
    double *zero[] = {c->tab, c->tab_t, c->atab, c->atab_t, c->ec, c->ac, c->ec_t, c->ac_t};
    const size_t zn[] = {c->TL, c->TL, (size_t)c->NA, (size_t)c->NA, qs, qs, qs, qs};
    for (int i = 0; i < 8; i++) CK(cudaMemsetAsync(zero[i], 0, (zn[i] ? zn[i] : 1) * sizeof(double), c->stream));
    CK(cudaMemsetAsync(c->Fn, 0, ncol * sizeof(int), c->stream));
    std::vector<int> site_of((size_t)c->NA);
    for (int s = 0; s < S; s++)
      for (int k = aoff_h[s]; k < aoff_h[s + 1]; k++) site_of[k] = s;
    CK(cudaMemcpy(c->nsite, site_of.data(), site_of.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->aoff, aoff_h.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->toff, toff_h.data(), (S + 1) * sizeof(int), cudaMemcpyHostToDevice));
    {
      std::vector<int2> desc_h(S);
      for (int s = 0; s < S; s++) desc_h[s] = make_int2(toff_h[s], aoff_h[s + 1] - aoff_h[s]);
      ALLOC(c->desc, (size_t)S);
      CK(cudaMemcpy(c->desc, desc_h.data(), (size_t)S * sizeof(int2), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(c->ncodon, ncodon_h.data(), ncodon_h.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->cons_in_list, cons_in_list_h.data(), S, cudaMemcpyHostToDevice));
  } else {
    REQUIRE(aoff_h == c->aoff_h && ncodon_h == c->ncodon_h && cons_in_list_h == c->cons_in_list_h &&
                cons_codon_h == c->cons_codon_h,
            "re-upload changes the per-site codon lists: create a new context");
  }
  // The inputs are DMA'd from the caller's buffers on the copy stream, the panel first, then the per-query
  // arrays in chunks of queries; each chunk is gathered on the main stream as soon as it has landed, so the
  // gather of chunk k runs under the transfer of chunk k+1.  The '0'/'1' characters of hla_types are turned
  // into 0/1 and closest_refs is range-checked on the device.
  CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
  CK(cudaEventRecord(c->ev_fork, c->stream));                 // the copy stream starts behind earlier work
  CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  // every DMA out of a caller's buffer runs on the copy stream: MCQ_OPT_ASYNC_UPLOAD = 1 returns once that
  // stream has drained, and the caller may then reuse all of its buffers
  CK(cudaMemcpyAsync(c->cons_codon, consensus_codons, S, cudaMemcpyHostToDevice, c->stream2));
  CK(cudaMemcpyAsync(c->st_slot_of, slot_of.data(), (size_t)S * 64, cudaMemcpyHostToDevice, c->stream2));
  CK(cudaEventRecord(c->ev_static, c->stream2));              // a pipelined evaluation starts behind the small tables
  const bool shared_panel = c->have_peers && c->trip_in_mailbox && c->opt_shared_panel && c->box.nranks > 1;
  if (shared_panel) {
    // every rank DMAs one slice of the panel rows from the host; the others arrive from the peers' copies
    // over NVLink once all slices have landed (a barrier over the ranks: all of them are in this call)
    const int R = c->box.nranks, me = c->box.rank;
    const size_t r0 = (size_t)N * me / R, r1 = (size_t)N * (me + 1) / R;
    CK(cudaMemcpyAsync(c->st_trip + r0 * S, ref_triplets + r0 * S, (r1 - r0) * S, cudaMemcpyHostToDevice, c->stream2));
    CK(mcq_launch_peer_barrier(&c->box, ++c->bar_seq, c->d_err, c->stream2));
    for (int k = 1; k < R; k++) {
      const int p = (me + k) % R;            // staggered, so that no peer serves everybody at once
      const size_t p0 = (size_t)N * p / R, p1 = (size_t)N * (p + 1) / R;
      CK(cudaMemcpyAsync(c->st_trip + p0 * S, c->peer_trip[p] + p0 * S, (p1 - p0) * S, cudaMemcpyDeviceToDevice, c->stream2));
    }
    c->launches++;
  } else {
    CK(cudaMemcpyAsync(c->st_trip, ref_triplets, (size_t)N * S, cudaMemcpyHostToDevice, c->stream2));
  }
  CK(cudaEventRecord(c->ev_chunk[0], c->stream2));
  CK(cudaStreamWaitEvent(c->stream, c->ev_chunk[0], 0));
  {
    KScope ks(c, MCQ_K_GATHER);
    CK(mcq_launch_gather_prepare(c->st_trip, c->st_trip_pad, c->st_slot_of, c->st_trip_tab, S, N, c->stream));
  }
  // new data: nothing of F or B is valid until the next evaluation (marked before any chunk is released to a
  // pipelined evaluation, which sets the F marks of the queries it sweeps)
  c->fv_h.assign(Q, -1);
  c->bv_h.assign(Q, S);
  CK(mcq_launch_set_valid(c->fv, c->bv, Q, H, -1, S, nullptr, 0, c->stream));
  c->launches++;
  // chunks of at least opt_chunk_queries queries: smaller ones leave the kernels behind them short of warps
  const int nchunk = std::max(1, std::min(kUploadChunks, Q / std::max(1, c->opt_chunk_queries)));
  const McqStatic st_all = make_static(c);
  for (int k = 0; k < nchunk; k++) {
    const size_t q0 = (size_t)Q * k / nchunk, q1 = (size_t)Q * (k + 1) / nchunk, nq = q1 - q0;
    CK(cudaMemcpyAsync(c->st_closest + q0 * L, closest_refs + q0 * L, nq * L * sizeof(int), cudaMemcpyHostToDevice, c->stream2));
    CK(cudaMemcpyAsync(c->st_cons + q0 * 3 * S, cons_query_nucls + q0 * 3 * S, nq * 3 * S, cudaMemcpyHostToDevice, c->stream2));
    CK(cudaMemcpyAsync(c->query_codon + q0 * S, query_codons + q0 * S, nq * S, cudaMemcpyHostToDevice, c->stream2));
    CK(cudaMemcpyAsync(c->hla + q0 * H, hla_types + q0 * H, nq * H, cudaMemcpyHostToDevice, c->stream2));
    CK(cudaEventRecord(c->ev_chunk[1 + k], c->stream2));
    CK(cudaStreamWaitEvent(c->stream, c->ev_chunk[1 + k], 0));
    {
      KScope ks(c, MCQ_K_GATHER);
      CK(mcq_launch_gather(c->st_trip_pad, c->st_trip, c->st_closest + q0 * L, c->st_cons + q0 * 3 * S, c->st_slot_of,
                           c->G + q0 * S * c->Lg, c->hla + q0 * H, (int)nq, S, L, c->Lg, N, H, c->d_err, c->stream));
    }
    CK(mcq_launch_rowdesc(st_all, c->rowdesc, (int)q0, (int)nq, c->stream));
    c->launches++;
    CK(cudaEventRecord(c->ev_ready[k], c->stream));
    c->pipe_q[k] = (int)q0; c->pipe_q[k + 1] = (int)q1;
  }
  c->hla_h.resize((size_t)Q * H);
  for (size_t i = 0; i < c->hla_h.size(); i++) c->hla_h[i] = hla_types[i] == '1';
  c->carriers.assign(H, 0);
  for (int q = 0; q < Q; q++)
    for (int h = 0; h < H; h++) c->carriers[h] += c->hla_h[(size_t)q * H + h];
  c->have_state = false;
  c->have_fwd = false;
  c->has_pending = false;
  if (shared_panel) {
    // nobody may overwrite its slice (the next upload) before every peer has read it
    CK(mcq_launch_peer_barrier(&c->box, ++c->bar_seq, c->d_err, c->stream2));
    CK(cudaEventRecord(c->ev_join, c->stream2));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->launches++;
  }
  if (c->opt_async_upload) {
    // the caller's buffers are free once the DMAs are done (mode 1 waits for that, mode 2 leaves it to the
    // next evaluation, which cannot finish before them); the gathers run on, the range check of closest_refs
    // is reported by the next evaluation (its reduction publishes the device flag)
    if (c->opt_async_upload == 1) CK(cudaStreamSynchronize(c->stream2));
    c->pipe_n = nchunk;
    c->have_static = true;
    return 0;
  }
  c->pipe_n = 0;
  CK(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));   // host staging vectors go out of scope
  REQUIRE(*c->h_err != 3, "panel exchange timed out: a rank did not reach the same mcq_ctx_upload_static");
  REQUIRE(*c->h_err == 0, "closest_refs entry out of range");
  c->have_static = true;
  c->have_state = false;
  c->have_fwd = false;
  c->has_pending = false;
  return 0;
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
extern "C" int mcq_ctx_set_params(mcq_ctx *c, const mcq_params *p) {
  REQUIRE(c && p, "null argument");
  CK(cudaSetDevice(c->d.device));
  const size_t S = c->d.num_sites, H = c->d.num_hla_types;
  c->m.mu = p->mu; c->m.kappa = p->kappa; c->m.phi = p->phi;
  // With an asynchronous upload in flight the main stream is busy gathering and the copy stream busy with the
  // upload's DMAs: the parameters then travel on a stream of their own and the pipelined evaluation waits for
  // ev_params.
  cudaStream_t pst = c->pipe_n > 0 ? c->stream_p : c->stream;
  c->params_on_copy_stream = c->pipe_n > 0;
  if (c->pipe_n > 0) CK(cudaEventSynchronize(c->ev_params));   // the staging block may still feed the last such copy
  else CK(cudaStreamSynchronize(c->stream));                    // ... or an earlier copy on the main stream
  double *h = c->h_pblock;
  double *h_prior = h, *h_R = h_prior + 64, *h_omega = h_R + S, *h_rev = h_omega + S, *h_esc = h_rev + S,
         *h_rate = h_esc + H * S;
  if (p->prior_C2) memcpy(h_prior, p->prior_C2, 64 * sizeof(double));
  if (p->R) memcpy(h_R, p->R, S * sizeof(double));
  if (p->omega) memcpy(h_omega, p->omega, S * sizeof(double));
  if (p->HLA_select_rev) memcpy(h_rev, p->HLA_select_rev, S * sizeof(double));
  if (p->HLA_select_esc) memcpy(h_esc, p->HLA_select_esc, H * S * sizeof(double));
  if (p->rate_out_of_codons)      // [64][S] -> site-major [S][64]
    for (size_t cd = 0; cd < 64; cd++) {
      const double *src = p->rate_out_of_codons + cd * S;
      for (size_t s = 0; s < S; s++) h_rate[s * 64 + cd] = src[s];
    }
  if (p->prior_C2 && p->R && p->omega && p->HLA_select_rev && p->HLA_select_esc && p->rate_out_of_codons) {
    CK(cudaMemcpyAsync(c->pblock, h, c->pblock_n * sizeof(double), cudaMemcpyHostToDevice, pst));
  } else {                        // only what was given
    auto put = [&](double *dst, const double *src, size_t n) {
      return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, pst);
    };
    if (p->prior_C2) CK(put(c->m.prior, h_prior, 64));
    if (p->R) CK(put(c->m.R, h_R, S));
    if (p->omega) CK(put(c->m.omega, h_omega, S));
    if (p->HLA_select_rev) CK(put(c->m.rev, h_rev, S));
    if (p->HLA_select_esc) CK(put(c->m.esc, h_esc, H * S));
    if (p->rate_out_of_codons) CK(put(c->m.rate_out, h_rate, 64 * S));
  }
  // no host wait: the copy is ordered before every later kernel on the context's stream
  if (c->params_on_copy_stream) {
    CK(cudaEventRecord(c->ev_params, c->stream_p));
    CK(cudaStreamWaitEvent(c->stream, c->ev_params, 0));      // everything later on the main stream sees them too
  }
  c->have_params = true;
  c->has_pending = false;
  return 0;
}

extern "C" int mcq_ctx_get_params(mcq_ctx *c, double *mu, double *R, double *omega, double *rate_out,
                                  double *esc, double *rev) {
  REQUIRE(c, "null context");
  CK(cudaSetDevice(c->d.device));
  const size_t S = c->d.num_sites, H = c->d.num_hla_types;
  CK(cudaStreamSynchronize(c->stream));
  if (mu) *mu = c->m.mu;
  if (R) CK(cudaMemcpy(R, c->m.R, S * sizeof(double), cudaMemcpyDeviceToHost));
  if (omega) CK(cudaMemcpy(omega, c->m.omega, S * sizeof(double), cudaMemcpyDeviceToHost));
  if (rate_out) {
    std::vector<double> tr(S * 64);
    CK(cudaMemcpy(tr.data(), c->m.rate_out, S * 64 * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t cd = 0; cd < 64; cd++)
      for (size_t s = 0; s < S; s++) rate_out[cd * S + s] = tr[s * 64 + cd];
  }
  if (esc) CK(cudaMemcpy(esc, c->m.esc, H * S * sizeof(double), cudaMemcpyDeviceToHost));
  if (rev) CK(cudaMemcpy(rev, c->m.rev, S * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------
// kernels sequencing helpers
// ------------------------------------------------------------------------------------------
struct EmisSet { double *tab, *atab, *ec, *ac; };
static EmisSet cur_set(mcq_ctx *c) { return {c->tab, c->atab, c->ec, c->ac}; }
static EmisSet tmp_set(mcq_ctx *c) { return {c->tab_t, c->atab_t, c->ec_t, c->ac_t}; }

// create_codon_to_codon_HLA over sites start..end: mode 0 ESC (consensus values), 1 REV (shared table), 2 BOTH
static int run_emission(mcq_ctx *c, int start, int end, int which_hla, int mode, const McqOverride &ov,
                        const EmisSet &src, const EmisSet &dst) {
  EmisArgs a;
  a.st = make_static(c); a.m = c->m; a.ov = ov;
  a.start = start; a.end = end;
  a.which_hla = (mode == 0) ? which_hla : -1;   // only an ESC-only update skips non-carriers (mcmc_moves.c:61-62)
  a.a0 = c->aoff_h[start]; a.a1 = c->aoff_h[end + 1];
  a.tab_src = src.tab; a.atab_src = src.atab; a.ec_src = src.ec; a.ac_src = src.ac;
  a.tab_dst = dst.tab; a.atab_dst = dst.atab; a.ec_dst = dst.ec; a.ac_dst = dst.ac;
  if (mode == 2) {
    KScope ks(c, MCQ_K_EMISSION);
    CK(mcq_launch_emission_both(a, c->stream));
    return 0;
  }
  if (mode != 0) {
    KScope ks(c, MCQ_K_EMISSION);
    CK(mcq_launch_emission_table(a, c->stream));
  }
  if (mode != 1) {
    KScope ks(c, MCQ_K_EMISSION);
    CK(mcq_launch_emission_cons(a, c->stream));
  }
  return 0;
}

// where the columns of a forward run go
enum FwdMode {
  FWD_IN_PLACE,   // into the buffer holding the current column (the full sweep, refills): counters into Fn
  FWD_PROPOSAL,   // into the other buffer (tmp_F of the reference): counters into Tn
  FWD_NO_STORE    // nowhere: likelihood only
};

static int run_forward(mcq_ctx *c, const double *tab, const double *ec, int start, int end, FwdMode mode, int llk_mode,
                       double *llk_out, int r_site, double r_value, bool sparse, int hla, bool lazy = false,
                       bool mark_end = false, int *tuned_blocks = nullptr) {
  FwdArgs a{};
  a.st = make_static(c);
  a.em.tab = tab; a.em.ec = ec; a.R = c->m.R; a.r_site = r_site; a.r_value = r_value;
  a.em_pre.tab = c->tab; a.em_pre.ec = c->ec;
  a.fv = lazy ? c->fv : nullptr;   // catch up F before `start` where it is stale
  a.mark_fv = mark_end ? c->fv : nullptr;
  if (mark_end) for (int &v : c->fv_h) v = end;      // (never combined with `sparse`)
  a.Fin = c->F; a.Fn_in = c->Fn; a.Fout = c->T; a.sel = c->sel;
  a.Fn_out = (mode == FWD_PROPOSAL) ? c->Tn : c->Fn;
  a.in_place = mode == FWD_IN_PLACE; a.no_store = mode == FWD_NO_STORE;
  a.start = start; a.end = end; a.llk_mode = llk_mode;
  a.B = c->B; a.Bn = c->Bn; a.llk_out = llk_out; a.llk_keep = c->llk_cur;
  a.active = sparse ? c->hla : nullptr; a.active_hla = hla;
  a.logBIG = c->logBIG;
  a.pipe = (sparse ? c->carriers[hla] : c->d.num_query) <= c->pipe_max;
  KScope ks(c, MCQ_K_FORWARD);
  if (c->ck > 1 && mode != FWD_NO_STORE) {
    // checkpointed matrices: the proposal's last column is handed to k_dot (llk_mode 1 becomes Tcol + 0)
    a.ck = c->ck; a.nc = c->nc;
    a.fv = (lazy || mode == FWD_PROPOSAL) && c->opt_lazy ? c->fv : a.fv;
    if (llk_mode == 1) { a.Tcol = c->Tcol; a.Tncol = c->Tncol; a.llk_mode = 0; }
    if (tuned_blocks != nullptr) *tuned_blocks = 0;
    CK(mcq_launch_forward_ck(a, c->stream));
    return 0;
  }
  if (c->wpq == 0 && mode == FWD_PROPOSAL && start == end && start > 0 && !lazy && !sparse && llk_mode == 1 &&
      tuned_blocks == nullptr) {
    CK(mcq_launch_column(a, c->stream));       // a one-column proposal, nothing owed: the direct kernel
    return 0;
  }
  if (c->wpq > 0) {
    if (tuned_blocks != nullptr) *tuned_blocks = 0;      // (the residency sweep is for the one-warp kernels)
    CK(mcq_launch_forward_mw(a, c->wpq, c->stream));
  } else if (tuned_blocks != nullptr) CK(mcq_tune_forward(a, c->stream, tuned_blocks, nullptr));
  else CK(mcq_launch_forward(a, c->stream));
  return 0;
}

// lazy: every query continues from its own bv mark down to column `bottom`; otherwise the reference's
// update_B(update) down to column 0
static int run_backward(mcq_ctx *c, int update, bool sparse, int hla, cudaStream_t stream = nullptr,
                        bool lazy = false, int bottom = 0, int *tuned_blocks = nullptr, bool deliver = false) {
  if (stream == nullptr) stream = c->stream;
  BwdArgs a{};
  a.st = make_static(c);
  a.em.tab = c->tab; a.em.ec = c->ec; a.R = c->m.R; a.B = c->B; a.Bn = c->Bn; a.update = update;
  a.bottom = bottom; a.bv = lazy ? c->bv : nullptr;
  a.active = sparse ? c->hla : nullptr; a.active_hla = hla;
  a.pipe = (sparse ? c->carriers[hla] : c->d.num_query) <= c->pipe_max;
  if (c->ck > 1) {
    a.ck = c->ck; a.nc = c->nc;
    if (deliver) { a.Bcol = c->Bcol; a.Bncol = c->Bncol; }
    if (tuned_blocks != nullptr) *tuned_blocks = 0;
    if (stream != c->stream) { c->launches++; c->k_n[MCQ_K_BACKWARD]++; CK(mcq_launch_backward_ck(a, stream)); return 0; }
    KScope ks(c, MCQ_K_BACKWARD);
    CK(mcq_launch_backward_ck(a, c->stream));
    return 0;
  }
  if (stream != c->stream) {          // beside the main stream: counted, not bracketed
    c->launches++;
    c->k_n[MCQ_K_BACKWARD]++;
    if (c->wpq > 0) CK(mcq_launch_backward_mw(a, c->wpq, stream));
    else CK(mcq_launch_backward(a, stream));
    return 0;
  }
  KScope ks(c, MCQ_K_BACKWARD);
  if (c->wpq > 0) {
    if (tuned_blocks != nullptr) *tuned_blocks = 0;
    CK(mcq_launch_backward_mw(a, c->wpq, c->stream));
  } else if (tuned_blocks != nullptr) CK(mcq_tune_backward(a, c->stream, tuned_blocks, nullptr));
  else CK(mcq_launch_backward(a, c->stream));
  return 0;
}

// Wait for the reduction kernel to publish evaluation `want` in the host-mapped slot.  The stream is
// queried now and then so that a failed launch cannot leave the host spinning.
static int wait_slot(mcq_ctx *c, unsigned long long want, cudaStream_t st) {
  volatile unsigned long long *seq = &c->h_slot->seq;
  for (unsigned spins = 1; *seq != want; spins++) {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
    if ((spins & 0x3FFF) == 0) {
      const cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess) {
        if (*seq == want) break;
        return mcq_fail(__FILE__, __LINE__, "the reduction kernel finished without publishing its result");
      }
      if (e != cudaErrorNotReady) return mcq_fail(__FILE__, __LINE__, cudaGetErrorString(e));
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  return 0;
}

// total over the shard (+ all ranks), brought to the host
// what the device error flag said when the reduction published its result
static int slot_error(mcq_ctx *c) {
  switch (c->h_slot->err) {
    case 0: return 0;
    case 1: return mcq_fail(__FILE__, __LINE__, "peer exchange timed out: a rank did not reach the same evaluation");
    case 2: return mcq_fail(__FILE__, __LINE__, "closest_refs entry out of range (found by the last mcq_ctx_upload_static)");
    case 3: return mcq_fail(__FILE__, __LINE__, "panel exchange timed out: a rank did not reach the same mcq_ctx_upload_static");
    default: return mcq_fail(__FILE__, __LINE__, "device error flag set");
  }
}

static int reduce_llk(mcq_ctx *c, const double *per_query, double *out, cudaStream_t st = nullptr, int n = -1) {
  if (st == nullptr) st = c->stream;
  if (n < 0) n = c->d.num_query;
  const unsigned long long want = ++c->eval_seq;
  if (c->have_peers && !(c->opt_exchange == 1 && c->comm)) {
    // shard total + NVLink mailbox exchange in one kernel
    {
      KScope ks(c, MCQ_K_SUM);
      c->step++;
      CK(mcq_launch_sum_exchange(per_query, n, c->d_sum, &c->box, c->step, c->d_err, c->d_slot, want, st));
    }
    if (wait_slot(c, want, st)) return 1;
    if (slot_error(c)) return 1;
    if (out) *out = c->h_slot->value;
    return 0;
  }
  {
    KScope ks(c, MCQ_K_SUM);
    CK(mcq_launch_sum(per_query, n, c->d_sum, c->comm ? nullptr : c->d_slot, want, c->d_err, st));
  }
  if (c->comm) {
    int r = g_nccl.AllReduce(c->d_sum, c->d_sum, 1, kNcclDouble, kNcclSum, c->comm, st);
    if (r != 0) return mcq_fail(__FILE__, __LINE__, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclAllReduce failed");
    // the reduced value reaches the host through the mapped slot, as on the other paths (no copy + stream wait)
    CK(mcq_launch_sum(c->d_sum, 1, c->d_sum, c->d_slot, want, c->d_err, st));
    c->launches++;
    if (wait_slot(c, want, st)) return 1;
    if (slot_error(c)) return 1;
    if (out) *out = c->h_slot->value;
    return 0;
  }
  if (wait_slot(c, want, st)) return 1;
  if (slot_error(c)) return 1;
  if (out) *out = c->h_slot->value;
  return 0;
}

// One double per rank, summed over the attached ranks in rank order (the same bits everywhere): the
// exchange of the likelihood totals, offered to the host chain for what it must decide collectively.
extern "C" int mcq_ctx_exchange_sum(mcq_ctx *c, double *value) {
  REQUIRE(c && value, "null argument");
  CK(cudaSetDevice(c->d.device));
  *c->h_sum = *value;
  CK(cudaMemcpyAsync(c->d_xchg, c->h_sum, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->launches = 0;
  return reduce_llk(c, c->d_xchg, value, nullptr, 1);
}

#define READY(c)                                                                   \
  do {                                                                             \
    REQUIRE(c, "null context");                                                    \
    REQUIRE((c)->have_static, "mcq_ctx_upload_static has not been called");        \
    REQUIRE((c)->have_params, "mcq_ctx_set_params has not been called");           \
    CK(cudaSetDevice((c)->d.device));                                              \
  } while (0)

// the queries an operation touches: all of them, or the carriers of one HLA type
template <typename Fn>
static void for_active(mcq_ctx *c, bool sparse, int hla, Fn fn) {
  const int Q = c->d.num_query, H = c->d.num_hla_types;
  if (!sparse) { for (int q = 0; q < Q; q++) fn(q); return; }
  const uint8_t *col = c->hla_h.data() + hla;
  for (int q = 0; q < Q; q++) if (col[(size_t)q * H]) fn(q);
}

// validity marks of F and B (lazy refills): on the device per query, mirrored on the host
static int set_marks(mcq_ctx *c, int f, int b, bool sparse, int hla) {
  CK(mcq_launch_set_valid(c->fv, c->bv, c->d.num_query, c->d.num_hla_types, f, b, sparse ? c->hla : nullptr, hla, c->stream));
  c->launches++;
  for_active(c, sparse, hla, [&](int q) {
    if (f >= -1) c->fv_h[q] = f;
    if (b >= 0) c->bv_h[q] = b;
  });
  return 0;
}

// Complete F (to the last site) and B (down to site 0) for every query whose refill is still owed.
static int flush_lazy(mcq_ctx *c) {
  const int S = c->d.num_sites;
  bool owe_f = false, owe_b = false;
  const int last_cp = (S / c->ck) * c->ck - 1;       // the last forward checkpoint (S - 1 with ck = 1)
  for_active(c, false, 0, [&](int q) { owe_f |= ((c->fv_h[q] + 1) / c->ck) * c->ck - 1 < last_cp; owe_b |= c->bv_h[q] > 0; });
  if (owe_f) {
    if (run_forward(c, c->tab, c->ec, S, S - 1, FWD_IN_PLACE, 0, nullptr, -1, 0.0, false, 0, true)) return 1;
    for (int &v : c->fv_h) v = S - 1;
  }
  if (owe_b) {
    if (run_backward(c, S - 1, false, 0, nullptr, true, 0)) return 1;
    for (int &v : c->bv_h) v = 0;
  }
  return 0;
}

extern "C" int mcq_ctx_flush(mcq_ctx *c) {
  READY(c);
  REQUIRE(c->have_state, "mcq_ctx_init_state has not been called");
  CallScope scope(c);
  if (flush_lazy(c)) return 1;
  if (c->timing == 0 || c->timing == 3) CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mcq_ctx_emission(mcq_ctx *c, int start, int end, int which_hla, int mode) {
  READY(c);
  REQUIRE(start >= 0 && end < c->d.num_sites && start <= end, "site range");
  REQUIRE(mode >= 0 && mode <= 2 && which_hla >= -1 && which_hla < c->d.num_hla_types, "mode / HLA");
  CallScope scope(c);
  McqOverride none{}; none.kind = -1;
  if (run_emission(c, start, end, which_hla, mode, none, cur_set(c), cur_set(c))) return 1;
  if (c->timing == 0 || c->timing == 3) CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// The evaluation behind an asynchronous upload: chunk k of the queries is evaluated on stream3 as soon as its
// gather has finished on the main stream (ev_ready[k]), so the sweep of chunk k runs under the gather of chunk
// k+1 and the transfers behind it.  Same kernels on views of the per-query arrays that start at the chunk.
static int full_llk_pipelined(mcq_ctx *c, int with_emission, double *llk) {
  const int S = c->d.num_sites, H = c->d.num_hla_types, Ls = c->Ls;
  // kPipeStreams streams in turn: a chunk of queries alone leaves the SMs short of warps, several in flight do not
  cudaStream_t st = c->stream3;
  CK(cudaStreamWaitEvent(st, c->ev_static, 0));          // the small static tables of this upload
  if (c->params_on_copy_stream) CK(cudaStreamWaitEvent(st, c->ev_params, 0));
  McqOverride none{}; none.kind = -1;
  const McqStatic all = make_static(c);
  if (with_emission) {       // the shared rows once, the per-query consensus values chunk by chunk below
    EmisArgs e;
    e.st = all; e.m = c->m; e.ov = none; e.start = 0; e.end = S - 1; e.which_hla = -1;
    e.a0 = c->aoff_h[0]; e.a1 = c->aoff_h[S];
    e.tab_src = e.tab_dst = c->tab; e.atab_src = e.atab_dst = c->atab; e.ec_src = e.ec_dst = c->ec; e.ac_src = e.ac_dst = c->ac;
    c->launches++; c->k_n[MCQ_K_EMISSION]++;
    CK(mcq_launch_emission_table(e, st));
  }
  CK(cudaEventRecord(c->ev_ps[0], c->stream3));          // the other streams start behind the shared rows
  for (int i = 0; i < kPipeStreams - 1; i++) CK(cudaStreamWaitEvent(c->pipe_st[i], c->ev_ps[0], 0));
  for (int k = 0; k < c->pipe_n; k++) {
    const size_t q0 = (size_t)c->pipe_q[k], nq = (size_t)c->pipe_q[k + 1] - q0;
    st = (k % kPipeStreams) ? c->pipe_st[k % kPipeStreams - 1] : c->stream3;
    CK(cudaStreamWaitEvent(st, c->ev_ready[k], 0));
    McqStatic v = all;       // the chunk's view
    v.Q = (int)nq;
    v.G = all.G + q0 * S * c->Lg; v.rowdesc = all.rowdesc + q0 * S;
    v.query_codon = all.query_codon + q0 * S; v.hla = all.hla + q0 * H;
    if (with_emission) {
      EmisArgs e;
      e.st = v; e.m = c->m; e.ov = none; e.start = 0; e.end = S - 1; e.which_hla = -1;
      e.a0 = c->aoff_h[0]; e.a1 = c->aoff_h[S];
      e.tab_src = e.tab_dst = c->tab; e.atab_src = e.atab_dst = c->atab;
      e.ec_src = e.ec_dst = c->ec + q0 * S; e.ac_src = e.ac_dst = c->ac + q0 * S;
      c->launches++; c->k_n[MCQ_K_EMISSION]++;
      CK(mcq_launch_emission_cons(e, st));
    }
    FwdArgs a{};
    a.st = v;
    a.em.tab = c->tab; a.em.ec = c->ec + q0 * S; a.em_pre = a.em; a.R = c->m.R; a.r_site = -1; a.r_value = 0.0;
    a.fv = nullptr; a.mark_fv = c->fv + q0;
    const size_t nc = (size_t)c->nc;
    a.Fin = c->F + q0 * nc * Ls; a.Fout = c->T + q0 * nc * Ls; a.sel = c->sel + q0 * nc;
    a.Fn_in = c->Fn + q0 * nc; a.Fn_out = c->Fn + q0 * nc;
    a.ck = c->ck; a.nc = c->nc;
    a.in_place = 1; a.no_store = 0; a.start = 0; a.end = S - 1; a.llk_mode = 2;
    a.B = c->B + q0 * nc * Ls; a.Bn = c->Bn + q0 * nc; a.llk_out = c->llk_cur + q0; a.llk_keep = c->llk_cur + q0;
    a.active = nullptr; a.active_hla = 0; a.logBIG = c->logBIG;
    a.pipe = 0;
    c->launches++; c->k_n[MCQ_K_FORWARD]++;
    if (c->ck > 1) CK(mcq_launch_forward_ck(a, st));
    else if (c->wpq > 0) CK(mcq_launch_forward_mw(a, c->wpq, st));
    else CK(mcq_launch_forward(a, st));
  }
  for (int &m : c->fv_h) m = S - 1;
  c->pipe_n = 0;
  c->params_on_copy_stream = false;
  st = c->stream3;
  for (int i = 0; i < kPipeStreams - 1; i++) {
    CK(cudaEventRecord(c->ev_ps[i], c->pipe_st[i]));
    CK(cudaStreamWaitEvent(st, c->ev_ps[i], 0));
  }
  if (reduce_llk(c, c->llk_cur, llk, st)) return 1;      // (waits for the result: everything above is done)
  CK(cudaEventRecord(c->ev_pipe, st));
  CK(cudaStreamWaitEvent(c->stream, c->ev_pipe, 0));
  c->has_pending = false;
  c->have_fwd = true;
  return 0;
}

extern "C" int mcq_ctx_full_llk(mcq_ctx *c, int with_emission, double *llk) {
  READY(c);
  CallScope scope(c);
  if (c->pipe_n > 0 && (c->timing == 0 || c->timing == 3)) return full_llk_pipelined(c, with_emission, llk);
  c->pipe_n = 0;
  const int S = c->d.num_sites;
  McqOverride none{}; none.kind = -1;
  if (with_emission && run_emission(c, 0, S - 1, -1, 2, none, cur_set(c), cur_set(c))) return 1;
  if (run_forward(c, c->tab, c->ec, 0, S - 1, FWD_IN_PLACE, 2, c->llk_cur, -1, 0.0, false, 0, false, true)) return 1;
  if (reduce_llk(c, c->llk_cur, llk)) return 1;
  c->has_pending = false;
  c->have_fwd = true;
  return 0;
}

extern "C" int mcq_ctx_llk_only(mcq_ctx *c, int with_emission, double *llk) {
  READY(c);
  c->pipe_n = 0;             // (everything here runs on the main stream, behind an asynchronous upload's set-up)
  CallScope scope(c);
  const int S = c->d.num_sites;
  McqOverride none{}; none.kind = -1;
  if (with_emission) {
    if (run_emission(c, 0, S - 1, -1, 2, none, cur_set(c), cur_set(c))) return 1;
    if (set_marks(c, -1, S, false, 0)) return 1;     // F and B no longer match the rows: all of both owed
  }
  if (run_forward(c, c->tab, c->ec, 0, S - 1, FWD_NO_STORE, 2, c->llk_cur, -1, 0.0, false, 0)) return 1;
  if (reduce_llk(c, c->llk_cur, llk)) return 1;
  c->has_pending = false;
  return 0;
}

extern "C" int mcq_ctx_tune(mcq_ctx *c, int blocks_per_sm[3]) {
  READY(c);
  REQUIRE(c->have_state, "mcq_ctx_init_state has not been called");
  c->pipe_n = 0;
  CallScope scope(c);
  if (flush_lazy(c)) return 1;            // the sweeps below recompute what is current: same inputs, same outputs
  const int S = c->d.num_sites;
  int k[3] = {0, 0, 0};
  if (run_forward(c, c->tab, c->ec, 0, S - 1, FWD_IN_PLACE, 2, c->llk_cur, -1, 0.0, false, 0, false, true, &k[0])) return 1;
  if (run_forward(c, c->tab, c->ec, 0, S - 1, FWD_NO_STORE, 2, c->llk_prop, -1, 0.0, false, 0, false, false, &k[1])) return 1;
  if (run_backward(c, S - 1, false, 0, nullptr, false, 0, &k[2])) return 1;
  CK(cudaStreamSynchronize(c->stream));
  c->has_pending = false;
  if (blocks_per_sm) for (int i = 0; i < 3; i++) blocks_per_sm[i] = k[i];
  return 0;
}

extern "C" int mcq_ctx_backward_fill(mcq_ctx *c) {
  READY(c);
  CallScope scope(c);
  if (run_backward(c, c->d.num_sites - 1, false, 0)) return 1;
  if (set_marks(c, -2, 0, false, 0)) return 1;
  if (c->have_fwd) c->have_state = true;      // mcq_ctx_full_llk + mcq_ctx_backward_fill = mcq_ctx_init_state
  if (c->timing == 0 || c->timing == 3) CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int mcq_ctx_init_state(mcq_ctx *c, double *llk) {
  READY(c);
  c->pipe_n = 0;
  CallScope scope(c);
  const int S = c->d.num_sites;
  McqOverride none{}; none.kind = -1;
  if (run_emission(c, 0, S - 1, -1, 2, none, cur_set(c), cur_set(c))) return 1;
  if (run_forward(c, c->tab, c->ec, 0, S - 1, FWD_IN_PLACE, 2, c->llk_cur, -1, 0.0, false, 0)) return 1;
  if (run_backward(c, S - 1, false, 0)) return 1;
  if (set_marks(c, S - 1, 0, false, 0)) return 1;
  if (reduce_llk(c, c->llk_cur, llk)) return 1;
  c->have_state = true;
  c->has_pending = false;
  return 0;
}

// ------------------------------------------------------------------------------------------
// proposals
// ------------------------------------------------------------------------------------------
// Columns ov.start..ov.end of a proposal into T and its per-query log-likelihood (the dot with B at
// ov.end, mcmc_moves.c:220-223) into llk_prop.  With lazy refills the columns of F before the range and
// of B down to ov.end are completed first where an earlier accept left them owed: F inside the forward
// kernel itself, B by a backward launch beside it on the second stream, joined by k_dot.
static int eval_range(mcq_ctx *c, const double *tab, const double *ec, const McqOverride &ov, int r_site,
                      double r_value, bool sparse) {
  c->pend_tab = tab; c->pend_ec = ec; c->pend_rsite = r_site; c->pend_rvalue = r_value;
  if (c->ck > 1) {
    // Checkpointed matrices: neither F[start-1] nor B[end] need exist.  The forward kernel starts from the last
    // current checkpoint before the range (completing owed ones on the way) and hands out the proposal's last
    // column; the backward kernel, beside it on the second stream, does the same from above for B[end]; k_dot
    // joins the two columns.
    const int cp_start = (ov.start / c->ck) * c->ck - 1;
    if (c->opt_lazy)
      for_active(c, sparse, ov.hla, [&](int q) {
        if (((c->fv_h[q] + 1) / c->ck) * c->ck - 1 < cp_start) c->fv_h[q] = cp_start;
        if (c->bv_h[q] > ov.end) c->bv_h[q] = ov.end;
      });
    const bool dual = c->timing != 2;
    if (dual) {
      CK(cudaEventRecord(c->ev_fork, c->stream));
      CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    }
    if (run_backward(c, -1, sparse, ov.hla, dual ? c->stream2 : nullptr, c->opt_lazy != 0, ov.end, nullptr, true)) return 1;
    if (dual) CK(cudaEventRecord(c->ev_join, c->stream2));
    if (run_forward(c, tab, ec, ov.start, ov.end, FWD_PROPOSAL, 1, nullptr, r_site, r_value, sparse, ov.hla, false))
      return 1;
    if (dual) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    DotArgs d{};
    d.st = make_static(c);
    d.cols = 1;
    d.T = c->Tcol; d.Tn = c->Tncol; d.B = c->Bcol; d.Bn = c->Bncol; d.site = ov.end;
    d.llk_out = c->llk_prop; d.llk_keep = c->llk_cur;
    d.active = sparse ? c->hla : nullptr; d.active_hla = ov.hla;
    d.logBIG = c->logBIG;
    d.wpq = 0;
    KScope ks(c, MCQ_K_SUM);
    CK(mcq_launch_dot(d, c->stream));
    return 0;
  }
  bool need_f = false, need_b = false;
  if (c->opt_lazy)
    for_active(c, sparse, ov.hla, [&](int q) { need_f |= c->fv_h[q] < ov.start - 1; need_b |= c->bv_h[q] > ov.end; });
  // what the launches below leave behind: F current up to start-1, B current down to end, where they were not
  if (need_f || need_b)
    for_active(c, sparse, ov.hla, [&](int q) {
      if (c->fv_h[q] < ov.start - 1) c->fv_h[q] = ov.start - 1;
      if (c->bv_h[q] > ov.end) c->bv_h[q] = ov.end;
    });
  if (!need_b)
    return run_forward(c, tab, ec, ov.start, ov.end, FWD_PROPOSAL, 1, c->llk_prop, r_site, r_value, sparse, ov.hla,
                       need_f);
  const bool dual = c->timing != 2;
  if (dual) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    if (run_backward(c, 0, sparse, ov.hla, c->stream2, true, ov.end)) return 1;
    CK(cudaEventRecord(c->ev_join, c->stream2));
  } else if (run_backward(c, 0, sparse, ov.hla, nullptr, true, ov.end)) return 1;
  if (run_forward(c, tab, ec, ov.start, ov.end, FWD_PROPOSAL, 0, nullptr, r_site, r_value, sparse, ov.hla, need_f))
    return 1;
  if (dual) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  DotArgs d{};
  d.st = make_static(c);
  d.T = c->T; d.Tn = c->Tn; d.F = c->F; d.sel = c->sel; d.B = c->B; d.Bn = c->Bn; d.site = ov.end;
  d.llk_out = c->llk_prop; d.llk_keep = c->llk_cur;
  d.active = sparse ? c->hla : nullptr; d.active_hla = ov.hla;
  d.logBIG = c->logBIG;
  d.wpq = c->wpq;
  {
    KScope ks(c, MCQ_K_SUM);
    CK(mcq_launch_dot(d, c->stream));
  }
  return 0;
}

extern "C" int mcq_ctx_propose(mcq_ctx *c, const mcq_proposal *p, double *sum_llk) {
  READY(c);
  REQUIRE(p, "null proposal");
  REQUIRE(c->have_state, "mcq_ctx_init_state has not been called");
  const int S = c->d.num_sites;
  McqOverride ov;
  ov.kind = p->kind; ov.start = p->start; ov.end = p->end; ov.hla = p->hla; ov.split = p->split;
  ov.v0 = p->value0; ov.v1 = p->value1;
  if (p->kind == MCQ_PROP_MU) { ov.start = 0; ov.end = S - 1; }
  REQUIRE(ov.start >= 0 && ov.end < S && ov.start <= ov.end, "proposal site range");
  CallScope scope(c);
  switch (p->kind) {
    case MCQ_PROP_REC:
      REQUIRE(ov.start == ov.end, "REC proposal covers one site");
      if (eval_range(c, c->tab, c->ec, ov, ov.start, ov.v0, false)) return 1;
      break;
    case MCQ_PROP_SEL:
      REQUIRE(ov.start == ov.end, "SEL proposal covers one site");
      if (run_emission(c, ov.start, ov.end, -1, 2, ov, cur_set(c), tmp_set(c))) return 1;
      if (eval_range(c, c->tab_t, c->ec_t, ov, -1, 0.0, false)) return 1;
      break;
    case MCQ_PROP_ESC:
      REQUIRE(ov.hla >= 0 && ov.hla < c->d.num_hla_types, "ESC proposal HLA index");
      if (run_emission(c, ov.start, ov.end, ov.hla, 0, ov, cur_set(c), tmp_set(c))) return 1;
      if (eval_range(c, c->tab, c->ec_t, ov, -1, 0.0, c->opt_sparse_esc != 0)) return 1;
      break;
    case MCQ_PROP_REV:
      if (run_emission(c, ov.start, ov.end, -1, 1, ov, cur_set(c), tmp_set(c))) return 1;
      if (eval_range(c, c->tab_t, c->ec, ov, -1, 0.0, false)) return 1;
      break;
    case MCQ_PROP_MU:
      {
        KScope ks(c, MCQ_K_EMISSION);
        CK(mcq_launch_mu_rescale(make_static(c), c->m, ov.v0, c->NA, c->tab, c->atab, c->ec, c->ac,
                                 c->tab_t, c->atab_t, c->ec_t, c->ac_t, c->stream));
        c->launches++;   // two kernels
      }
      // the whole-genome sweep of mcmc_moves.c:1124-1134 without writing tmp_F: only an accept needs it
      if (run_forward(c, c->tab_t, c->ec_t, 0, S - 1, FWD_NO_STORE, 2, c->llk_prop, -1, 0.0, false, 0)) return 1;
      break;
    default:
      return mcq_fail(__FILE__, __LINE__, "unknown proposal kind");
  }
  if (reduce_llk(c, c->llk_prop, sum_llk)) return 1;
  c->pending = ov;
  c->has_pending = true;
  return 0;
}

extern "C" int mcq_ctx_reject(mcq_ctx *c) {
  REQUIRE(c, "null context");
  c->has_pending = false;
  return 0;
}

template <typename T>
static void swap_ptr(T *&a, T *&b) { T *t = a; a = b; b = t; }

extern "C" int mcq_ctx_accept(mcq_ctx *c) {
  READY(c);
  REQUIRE(c->has_pending, "no proposal to accept");
  const int S = c->d.num_sites;
  const McqOverride ov = c->pending;
  CallScope scope(c);
  const McqStatic st = make_static(c);
  if (ov.kind == MCQ_PROP_MU) {
    // the proposal was evaluated without writing tmp_F: the same sweep again, this time stored - straight
    // into the current columns, with the proposal's rows, which then become the current ones (the
    // reference swaps the roles of the tmp and current arrays, mcmc_moves.c:1166-1169)
    if (run_forward(c, c->tab_t, c->ec_t, 0, S - 1, FWD_IN_PLACE, 2, c->llk_prop, -1, 0.0, false, 0)) return 1;
    swap_ptr(c->tab, c->tab_t); swap_ptr(c->atab, c->atab_t);
    swap_ptr(c->ec, c->ec_t); swap_ptr(c->ac, c->ac_t);
    c->m.mu = ov.v0;
    if (c->opt_lazy) {
      if (set_marks(c, S - 1, S, false, 0)) return 1;    // F complete, all of B owed
    } else if (run_backward(c, S - 1, false, 0)) return 1;
  } else {
    const bool sparse = (ov.kind == MCQ_PROP_ESC) && c->opt_sparse_esc != 0;
    {
      CommitArgs k{};
      k.st = st; k.m = c->m; k.ov = ov;
      k.tab_t = c->tab_t; k.atab_t = c->atab_t; k.ec_t = c->ec_t; k.ac_t = c->ac_t;
      k.tab = c->tab; k.atab = c->atab; k.ec = c->ec; k.ac = c->ac;
      if (ov.kind == MCQ_PROP_SEL || ov.kind == MCQ_PROP_REV) {   // shared rows of the range
        k.t0 = (size_t)c->toff_h[ov.start]; k.t1 = (size_t)c->toff_h[ov.end + 1];
        k.a0 = (size_t)c->aoff_h[ov.start]; k.a1 = (size_t)c->aoff_h[ov.end + 1];
      }
      k.copy_cons = (ov.kind == MCQ_PROP_SEL || ov.kind == MCQ_PROP_ESC) ? 1 : 0;   // per-query consensus values
      k.sel = c->sel; k.Tn = c->Tn; k.Fn = c->Fn;
      // the checkpoint columns inside the range (every site with ck = 1)
      k.nc = c->nc; k.c0 = ov.start / c->ck; k.c1 = (ov.end + 1) / c->ck - 1;
      k.active = sparse ? c->hla : nullptr; k.active_hla = ov.hla;
      // with lazy refills the refills of mcmc_moves.c:266-279 are owed, not run: the next evaluation completes
      // what it needs (marks on the device here, mirrored on the host below)
      k.fv = c->fv; k.bv = c->bv; k.mark = c->opt_lazy ? ov.end : -1;
      KScope ks(c, MCQ_K_COMMIT);
      CK(mcq_launch_commit_all(k, c->stream));
    }
    if (c->opt_lazy) {
      for_active(c, sparse, ov.hla, [&](int q) { c->fv_h[q] = ov.end; c->bv_h[q] = ov.end; });
    } else {
      // F after the range and B from the range down to 0 do not depend on each other: two streams
      const bool dual = c->timing != 2 && ov.end + 1 < S;
      if (dual) {
        CK(cudaEventRecord(c->ev_fork, c->stream));
        CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        if (run_backward(c, ov.end, sparse, ov.hla, c->stream2)) return 1;
        CK(cudaEventRecord(c->ev_join, c->stream2));
      }
      if (ov.end + 1 < S)
        if (run_forward(c, c->tab, c->ec, ov.end + 1, S - 1, FWD_IN_PLACE, 0, nullptr, -1, 0.0, sparse, ov.hla)) return 1;
      if (dual) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      else if (run_backward(c, ov.end, sparse, ov.hla)) return 1;
    }
  }
  swap_ptr(c->llk_cur, c->llk_prop);
  c->has_pending = false;
  // no host wait here: everything later is ordered behind this on the context's stream
  return 0;
}

// ------------------------------------------------------------------------------------------
// inspection
// ------------------------------------------------------------------------------------------
// Checkpointed matrices hold every ck-th column only: for inspection the whole matrix of one query is computed
// afresh, into scratch of the reference's size, by the sweeps of the un-checkpointed layout on a one-query view -
// the same cell updates, so the same bits the kept columns have.
static int recompute_matrix(mcq_ctx *c, int which, int q, double *out, int *rnorm) {
  const int S = c->d.num_sites, L = c->d.closest_n, Ls = c->Ls, H = c->d.num_hla_types;
  struct Scratch {
    void *p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
  } f0, f1, n0, n1, z;
  const size_t cells = (size_t)S * Ls;
  CK(cudaMalloc(&f0.p, cells * sizeof(double))); CK(cudaMalloc(&f1.p, cells * sizeof(double)));
  CK(cudaMalloc(&n0.p, (size_t)S * sizeof(int))); CK(cudaMalloc(&n1.p, (size_t)S * sizeof(int)));
  CK(cudaMalloc(&z.p, (size_t)S));
  CK(cudaMemsetAsync(z.p, 0, (size_t)S, c->stream));
  CK(cudaMemsetAsync(f1.p, 0, cells * sizeof(double), c->stream));
  CK(cudaMemsetAsync(n1.p, 0, (size_t)S * sizeof(int), c->stream));
  McqStatic v = make_static(c);
  v.Q = 1;
  v.G += (size_t)q * S * c->Lg; v.rowdesc += (size_t)q * S; v.query_codon += (size_t)q * S; v.hla += (size_t)q * H;
  const double *src = (const double *)f0.p;
  const int *nsrc = (const int *)n0.p;
  if (which == MCQ_BUF_B) {
    BwdArgs b{};
    b.st = v; b.em.tab = c->tab; b.em.ec = c->ec + (size_t)q * S; b.R = c->m.R;
    b.B = (double *)f0.p; b.Bn = (int *)n0.p; b.update = S - 1; b.bottom = 0;
    CK(mcq_launch_backward(b, c->stream));
  } else {
    FwdArgs a{};
    a.st = v; a.em.tab = c->tab; a.em.ec = c->ec + (size_t)q * S; a.em_pre = a.em; a.R = c->m.R; a.r_site = -1;
    a.Fin = (double *)f0.p; a.Fn_in = (int *)n0.p; a.Fout = (double *)f1.p; a.Fn_out = (int *)n0.p;
    a.sel = (const uint8_t *)z.p; a.in_place = 1; a.start = 0; a.end = S - 1; a.logBIG = c->logBIG;
    CK(mcq_launch_forward(a, c->stream));
    if (which == MCQ_BUF_TMP_F && c->has_pending && c->pending.kind != MCQ_PROP_MU) {
      a.em.tab = c->pend_tab; a.em.ec = c->pend_ec + (size_t)q * S;
      a.r_site = c->pend_rsite; a.r_value = c->pend_rvalue;
      a.in_place = 0; a.Fn_out = (int *)n1.p; a.start = c->pending.start; a.end = c->pending.end;
      CK(mcq_launch_forward(a, c->stream));
    }
    if (which == MCQ_BUF_TMP_F) { src = (const double *)f1.p; nsrc = (const int *)n1.p; }
  }
  CK(cudaStreamSynchronize(c->stream));
  if (out) {
    std::vector<double> slab(cells);
    CK(cudaMemcpy(slab.data(), src, cells * sizeof(double), cudaMemcpyDeviceToHost));
    for (int r = 0; r < L; r++)
      for (int s2 = 0; s2 < S; s2++) out[(size_t)r * S + s2] = slab[(size_t)s2 * Ls + r];
  }
  if (rnorm) CK(cudaMemcpy(rnorm, nsrc, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int mcq_ctx_get_matrix(mcq_ctx *c, int which, int q, double *out, int *rnorm) {
  REQUIRE(c && c->have_static, "context not ready");
  REQUIRE(q >= 0 && q < c->d.num_query, "query index");
  CK(cudaSetDevice(c->d.device));
  if (which == MCQ_BUF_F_KEPT || which == MCQ_BUF_B_KEPT) {
    // the columns that are actually kept (all of them with a checkpoint stride of 1), owed refills completed:
    // the other columns come back as 0 and their counters as -1
    const int S = c->d.num_sites, L = c->d.closest_n, Ls = c->Ls, ck = c->ck, nc = c->nc;
    const bool fwd = which == MCQ_BUF_F_KEPT;
    if (c->have_state && c->have_params && flush_lazy(c)) return 1;
    CK(cudaStreamSynchronize(c->stream));
    std::vector<double> k0((size_t)nc * Ls), k1;
    std::vector<uint8_t> sel(nc, 0);
    std::vector<int> cnt(nc);
    CK(cudaMemcpy(k0.data(), (fwd ? c->F : c->B) + (size_t)q * nc * Ls, k0.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt.data(), (fwd ? c->Fn : c->Bn) + (size_t)q * nc, nc * sizeof(int), cudaMemcpyDeviceToHost));
    if (fwd) {
      k1.resize(k0.size());
      CK(cudaMemcpy(k1.data(), c->T + (size_t)q * nc * Ls, k1.size() * sizeof(double), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(sel.data(), c->sel + (size_t)q * nc, nc, cudaMemcpyDeviceToHost));
    }
    if (out) memset(out, 0, sizeof(double) * (size_t)L * S);
    for (int s2 = 0; s2 < S; s2++) {
      const bool kept = fwd ? (s2 % ck == ck - 1) : (s2 % ck == 0);
      if (rnorm) rnorm[s2] = kept ? cnt[s2 / ck] : -1;
      if (!kept || !out) continue;
      const std::vector<double> &k = sel[s2 / ck] ? k1 : k0;
      for (int r = 0; r < L; r++) out[(size_t)r * S + s2] = k[(size_t)(s2 / ck) * Ls + r];
    }
    return 0;
  }
  if (c->ck > 1) {
    REQUIRE(c->have_params, "mcq_ctx_set_params has not been called");
    CK(cudaStreamSynchronize(c->stream));
    return recompute_matrix(c, which, q, out, rnorm);
  }
  const int S = c->d.num_sites, L = c->d.closest_n, Ls = c->Ls;
  const int *nsrc = which == MCQ_BUF_F ? c->Fn : which == MCQ_BUF_B ? c->Bn : c->Tn;
  if (c->have_state && c->have_params && which != MCQ_BUF_TMP_F && flush_lazy(c)) return 1;
  CK(cudaStreamSynchronize(c->stream));
  if (out) {
    std::vector<double> slab((size_t)S * Ls), slab1;
    std::vector<uint8_t> sel(S, 0);
    if (which == MCQ_BUF_B) CK(cudaMemcpy(slab.data(), c->B + (size_t)q * S * Ls, slab.size() * sizeof(double), cudaMemcpyDeviceToHost));
    else {   // column by column from the buffer the selector names (the other one for tmp_F)
      slab1.resize(slab.size());
      CK(cudaMemcpy(slab.data(), c->F + (size_t)q * S * Ls, slab.size() * sizeof(double), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(slab1.data(), c->T + (size_t)q * S * Ls, slab.size() * sizeof(double), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(sel.data(), c->sel + (size_t)q * S, S, cudaMemcpyDeviceToHost));
      if (which == MCQ_BUF_TMP_F) for (auto &b : sel) b ^= 1;
    }
    for (int r = 0; r < L; r++)
      for (int s = 0; s < S; s++) out[(size_t)r * S + s] = (sel[s] ? slab1 : slab)[(size_t)s * Ls + r];
  }
  if (rnorm) CK(cudaMemcpy(rnorm, nsrc + (size_t)q * S, S * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int mcq_ctx_get_table(mcq_ctx *c, int which, int q, double *out) {
  REQUIRE(c && c->have_static && out, "context not ready");
  REQUIRE(q >= 0 && q < c->d.num_query, "query index");
  CK(cudaSetDevice(c->d.device));
  const int S = c->d.num_sites;
  const bool tmp = (which == MCQ_TAB_TMP_C2C || which == MCQ_TAB_TMP_A);
  const bool want_a = (which == MCQ_TAB_A || which == MCQ_TAB_TMP_A);
  CK(cudaStreamSynchronize(c->stream));
  std::vector<double> cons(S), shared(want_a ? (size_t)c->NA : c->TL);
  std::vector<uint8_t> qc(S);
  const double *cons_src = want_a ? (tmp ? c->ac_t : c->ac) : (tmp ? c->ec_t : c->ec);
  const double *sh_src = want_a ? (tmp ? c->atab_t : c->atab) : (tmp ? c->tab_t : c->tab);
  CK(cudaMemcpy(cons.data(), cons_src + (size_t)q * S, S * sizeof(double), cudaMemcpyDeviceToHost));
  if (!shared.empty()) CK(cudaMemcpy(shared.data(), sh_src, shared.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(qc.data(), c->query_codon + (size_t)q * S, S, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(double) * 64 * S);
  for (int s = 0; s < S; s++) {
    out[(size_t)c->cons_codon_h[s] * S + s] = cons[s];
    const int a0 = c->aoff_h[s], kr = c->aoff_h[s + 1] - a0;
    for (int k = 0; k < kr; k++)
      out[(size_t)c->ncodon_h[a0 + k] * S + s] =
          want_a ? shared[a0 + k] : shared[(size_t)c->toff_h[s] + (size_t)qc[s] * kr + k];
  }
  return 0;
}

extern "C" int mcq_ctx_get_query_llk(mcq_ctx *c, int proposed, double *out) {
  REQUIRE(c && out, "null argument");
  CK(cudaSetDevice(c->d.device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out, proposed ? c->llk_prop : c->llk_cur, c->d.num_query * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------------
// ---- NVLink peer mailboxes ----------------------------------------------------------------
// layout of the exported allocation: double value[2][16]; u64 seq[2][16]; u64 bar[16]; then, from
// kPanelOffset, this rank's copy of the reference panel (st_trip): with it behind the same IPC handle the
// ranks of a box DMA one slice of the panel each from the host and read the other slices from their peers
// over NVLink (mcq_ctx_upload_static), instead of every rank pulling the whole panel over PCIe.
static const size_t kMailboxBytes = 2 * 16 * (sizeof(double) + sizeof(unsigned long long)) + 16 * sizeof(unsigned long long);
static const size_t kPanelOffset = 1024;
static_assert(kMailboxBytes <= kPanelOffset, "mailbox header");

extern "C" int mcq_ctx_peer_handle(mcq_ctx *c, unsigned char handle[64]) {
  REQUIRE(c && handle, "null argument");
  CK(cudaSetDevice(c->d.device));
  if (!c->mailbox) {
    const size_t panel = (size_t)c->d.total_num_refs * c->d.num_sites;
    CK(cudaMalloc(&c->mailbox, kPanelOffset + panel));
    c->bytes += kPanelOffset + panel;
    CK(cudaMemset(c->mailbox, 0, kPanelOffset));
    CK(cudaDeviceSynchronize());
    // the panel staging buffer moves behind the mailbox (it is refilled by every mcq_ctx_upload_static)
    if (c->st_trip) {
      CK(cudaFree(c->st_trip));
      c->bytes -= panel;
    }
    c->st_trip = (int8_t *)c->mailbox + kPanelOffset;
    c->trip_in_mailbox = true;
  }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->mailbox));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle, &h, 64);
  return 0;
}

extern "C" int mcq_ctx_attach_peers(mcq_ctx *c, int nranks, int rank, const unsigned char *handles) {
  REQUIRE(c && handles, "null argument");
  REQUIRE(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks, "rank / nranks (at most 16 ranks on one box)");
  REQUIRE(c->mailbox, "call mcq_ctx_peer_handle first");
  CK(cudaSetDevice(c->d.device));
  for (int r = 0; r < nranks; r++) {
    char *base;
    if (r == rank) base = (char *)c->mailbox;
    else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)r * 64, 64);
      CK(cudaIpcOpenMemHandle(&c->peer_map[r], h, cudaIpcMemLazyEnablePeerAccess));
      base = (char *)c->peer_map[r];
    }
    // every rank indexes a mailbox as [slot][R] with R = nranks
    c->box.value[r] = (double *)base;
    c->box.seq[r] = (unsigned long long *)(base + 2 * 16 * sizeof(double));
    c->box.bar[r] = (unsigned long long *)(base + 2 * 16 * (sizeof(double) + sizeof(unsigned long long)));
    c->peer_trip[r] = (const int8_t *)base + kPanelOffset;
  }
  c->box.nranks = nranks;
  c->box.rank = rank;
  c->nranks = nranks;
  c->have_peers = true;
  c->step = 0;
  c->bar_seq = 0;
  return 0;
}

extern "C" int mcq_comm_unique_id(unsigned char id[128]) {
  if (nccl_load()) return 1;
  ncclUniqueId u;
  int r = g_nccl.GetUniqueId(&u);
  if (r != 0) return mcq_fail(__FILE__, __LINE__, "ncclGetUniqueId failed");
  memcpy(id, u.internal, 128);
  return 0;
}

extern "C" int mcq_ctx_attach_comm(mcq_ctx *c, int nranks, int rank, const unsigned char id[128]) {
  REQUIRE(c && id, "null argument");
  REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks");
  if (nccl_load()) return 1;
  CK(cudaSetDevice(c->d.device));
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  int r = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
  if (r != 0) return mcq_fail(__FILE__, __LINE__, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclCommInitRank failed");
  c->nranks = nranks;
  return 0;
}
