// mcq_closest.cu - Hamming top-L reference selection (closest_seqs.c) on the GPU.
//
//   K1a k_pack      ASCII -> three bit planes per sequence (code low bit, code high bit, "unknown")
//   K1b k_hamming   bit-parallel mismatch counts for a tile of queries x references
//   K2a k_select_count  per query: the distance T of sorted element `n` (radix select) and the
//                       sizes of the "< T" and "== T" groups
//   K2b k_select_fill   per query: the "< T" group sorted by (distance, index) and the "== T"
//                       group in ascending index order
//   host            replays the reference's boundary-tie shuffle with libc random() in query order
//
// Distance semantics follow closest_seqs.c:17-48 as compiled: a position
// counts when the QUERY base is one of TCAG and the reference code differs from it - an unknown
// reference base (code 4) therefore counts as a mismatch - and the value is 2 x that count; the
// gap term of :46 is identically 0.  The kernels keep the count (0..num_bases) as u16.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "mcq_internal.cuh"
#include "../../include/mcq_b200.h"

int mcq_fail(const char *file, int line, const char *what);
#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) return mcq_fail(__FILE__, __LINE__, cudaGetErrorString(e_));   \
  } while (0)
#define REQUIRE(c, msg)                                   \
  do {                                                    \
    if (!(c)) return mcq_fail(__FILE__, __LINE__, msg);   \
  } while (0)

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------
// K1a: pack.  One warp per sequence; 32 characters per step through a ballot.
// planes: seq-major [n][W][3] when seq_major, else word-major [W][3][npad] (coalesced over
// references in k_hamming).  plane 0 = code bit 0, 1 = code bit 1, 2 = unknown (code 4).
// amino_acids.c:10-17: T,t->0  C,c->1  A,a->2  G,g->3  else 4.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int base_code(unsigned char ch) {
  switch (ch | 0x20) {   // fold case for letters
    case 't': return 0;
    case 'c': return 1;
    case 'a': return 2;
    case 'g': return 3;
    default: return 4;
  }
}

__global__ void __launch_bounds__(256) k_pack(const unsigned char *__restrict__ seqs, size_t stride, int n,
                                              int nbases, int W, u64 *__restrict__ planes, int seq_major,
                                              size_t npad) {
  const int seq = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (seq >= n) return;
  const unsigned char *row = seqs + (size_t)seq * stride;
  for (int w = 0; w < W; w++) {
    unsigned m[3][2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int i = w * 64 + h * 32 + lane;
      // only 'A'..'Z','a'..'z' can be bases; (ch|0x20) would also fold e.g. 'T'^0x20 punctuation: guard
      int code = 4;
      if (i < nbases) {
        const unsigned char ch = row[i];
        if ((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z')) code = base_code(ch);
      }
      // beyond the end of the sequence: query side "unknown" -> never counted; reference side is
      // irrelevant there because the query mask is 0
      m[0][h] = __ballot_sync(0xffffffffu, code & 1 && code != 4);
      m[1][h] = __ballot_sync(0xffffffffu, code & 2 && code != 4);
      m[2][h] = __ballot_sync(0xffffffffu, code == 4);
    }
    if (lane < 3) {
      const u64 v = (u64)m[lane][0] | ((u64)m[lane][1] << 32);
      if (seq_major) planes[((size_t)seq * W + w) * 3 + lane] = v;
      else planes[((size_t)w * 3 + lane) * npad + seq] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1b: distances.  Block = 128 threads; thread t owns RT references (coalesced word-major
// loads), the block's TQ queries sit in shared memory (broadcast reads), counts in registers.
// mismatch word = q_valid & (r_unknown | (q_lo ^ r_lo) | (q_hi ^ r_hi)); q_valid = ~q_unknown.
// ------------------------------------------------------------------------------------------
template <int TQ, int RT>
__global__ void __launch_bounds__(128) k_hamming(const u64 *__restrict__ refp, size_t npad, int N,
                                                 const u64 *__restrict__ qryp, int Q, int W,
                                                 uint16_t *__restrict__ D) {
  extern __shared__ u64 sq[];   // [TQ][W][3], plane 2 stored as q_valid
  const int q0 = blockIdx.y * TQ;
  const int nq = min(TQ, Q - q0);
  for (int i = threadIdx.x; i < TQ * W * 3; i += blockDim.x) {
    const int q = i / (W * 3);
    u64 v = 0;
    if (q < nq) {
      v = qryp[(size_t)(q0 + q) * W * 3 + (i % (W * 3))];
      if (i % 3 == 2) v = ~v;
    }
    sq[i] = v;
  }
  __syncthreads();
  const int rbase = (blockIdx.x * 128 + threadIdx.x);   // references rbase + k*128*gridDim... (RT consecutive tiles)
  int ridx[RT];
#pragma unroll
  for (int k = 0; k < RT; k++) ridx[k] = (blockIdx.x * RT + k) * 128 + threadIdx.x;
  (void)rbase;
  int acc[TQ][RT];
#pragma unroll
  for (int q = 0; q < TQ; q++)
#pragma unroll
    for (int k = 0; k < RT; k++) acc[q][k] = 0;

  for (int w = 0; w < W; w++) {
    u64 rl[RT], rh[RT], ru[RT];
#pragma unroll
    for (int k = 0; k < RT; k++) {
      const size_t o = (size_t)w * 3 * npad + ridx[k];   // ridx < npad always (npad covers the grid)
      rl[k] = refp[o]; rh[k] = refp[o + npad]; ru[k] = refp[o + 2 * npad];
    }
#pragma unroll
    for (int q = 0; q < TQ; q++) {
      const u64 ql = sq[(q * W + w) * 3 + 0], qh = sq[(q * W + w) * 3 + 1], qv = sq[(q * W + w) * 3 + 2];
#pragma unroll
      for (int k = 0; k < RT; k++) {
        const u64 x = (((ql ^ rl[k]) | (qh ^ rh[k])) | ru[k]) & qv;
        acc[q][k] += __popcll(x);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < TQ; q++)
    if (q < nq)
#pragma unroll
      for (int k = 0; k < RT; k++)
        if (ridx[k] < N) D[(size_t)(q0 + q) * N + ridx[k]] = (uint16_t)acc[q][k];
}

// ------------------------------------------------------------------------------------------
// K2: selection.  Values are compared as unsigned keys; int32 inputs (the drop-in) are biased.
// ------------------------------------------------------------------------------------------
template <typename V> struct KeyOf;
template <> struct KeyOf<uint16_t> {
  static const int kBytes = 2;
  __device__ static unsigned key(uint16_t v) { return v; }
};
template <> struct KeyOf<int> {
  static const int kBytes = 4;
  __device__ static unsigned key(int v) { return (unsigned)v ^ 0x80000000u; }
};

// per query: T-key = key of the element of rank n (0-based) in ascending order, n_lt, n_eq
template <typename V>
__global__ void __launch_bounds__(512) k_select_count(const V *__restrict__ D, int N, int n,
                                                      unsigned *__restrict__ Tkey, int *__restrict__ n_lt,
                                                      int *__restrict__ n_eq) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank, s_less;
  const V *row = D + (size_t)blockIdx.x * N;
  unsigned prefix = 0, mask = 0, rank = (unsigned)n, less_total = 0;
  for (int b = KeyOf<V>::kBytes - 1; b >= 0; b--) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const unsigned k = KeyOf<V>::key(row[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> (8 * b)) & 255], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned cum = 0, d = 0;
      for (d = 0; d < 256; d++) {
        if (cum + hist[d] > rank) break;
        cum += hist[d];
      }
      s_prefix = prefix | (d << (8 * b));
      s_rank = rank - cum;
      s_less = less_total + cum;
    }
    __syncthreads();
    prefix = s_prefix; rank = s_rank; less_total = s_less;
    mask |= 255u << (8 * b);
    __syncthreads();
  }
  // prefix is now the full key of the rank-n element; less_total = #keys < it
  if (threadIdx.x == 0) { Tkey[blockIdx.x] = prefix; n_lt[blockIdx.x] = (int)less_total; }
  // count equals
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned loc = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) loc += (KeyOf<V>::key(row[i]) == prefix);
  atomicAdd(&hist[0], loc);
  __syncthreads();
  if (threadIdx.x == 0) n_eq[blockIdx.x] = (int)hist[0];
}

// per query: "< T" group sorted by (key, index) -> lt_idx/lt_key [Q][cap]; "== T" group in index
// order -> eq_pool[eq_off[q] ...]
template <typename V>
__global__ void __launch_bounds__(512) k_select_fill(const V *__restrict__ D, int N, int cap,
                                                     const unsigned *__restrict__ Tkey,
                                                     const long long *__restrict__ eq_off,
                                                     int *__restrict__ lt_idx, unsigned *__restrict__ lt_key,
                                                     int *__restrict__ eq_pool) {
  extern __shared__ u64 keys[];          // cap2 = next pow2 >= cap
  __shared__ unsigned warp_tot[16];
  __shared__ unsigned s_nlt, s_base;
  const int q = blockIdx.x;
  const V *row = D + (size_t)q * N;
  const unsigned T = Tkey[q];
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  for (int i = threadIdx.x; i < cap2; i += blockDim.x) keys[i] = ~0ull;
  if (threadIdx.x == 0) { s_nlt = 0; s_base = 0; }
  __syncthreads();
  int *eq_out = eq_pool + eq_off[q];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i0 = 0; i0 < N; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    unsigned k = 0xffffffffu;
    bool is_eq = false;
    if (i < N) {
      k = KeyOf<V>::key(row[i]);
      is_eq = (k == T);
      if (k < T) {
        const unsigned slot = atomicAdd(&s_nlt, 1u);
        if (slot < (unsigned)cap2) keys[slot] = ((u64)k << 32) | (unsigned)i;
      }
    }
    // ordered compaction of the equal group
    const unsigned bal = __ballot_sync(0xffffffffu, is_eq);
    const unsigned before = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned woff = 0, tot = 0;
    for (int w = 0; w < nw; w++) {
      if (w < wid) woff += warp_tot[w];
      tot += warp_tot[w];
    }
    const unsigned base = s_base;
    if (is_eq) eq_out[base + woff + before] = i;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + tot;
    __syncthreads();
  }
  // bitonic sort of the "< T" keys (padding = all ones sorts last)
  for (int size = 2; size <= cap2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < cap2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = ((i & size) == 0);
          const u64 a = keys[i], b = keys[j];
          if ((a > b) == up) { keys[i] = b; keys[j] = a; }
        }
      }
      __syncthreads();
    }
  const int nlt = min((int)s_nlt, cap);
  for (int i = threadIdx.x; i < nlt; i += blockDim.x) {
    lt_idx[(size_t)q * cap + i] = (int)(keys[i] & 0xffffffffu);
    lt_key[(size_t)q * cap + i] = (unsigned)(keys[i] >> 32);
  }
}

// ------------------------------------------------------------------------------------------
// The batched path (16-bit mismatch counts, at most num_bases): one read of the row builds the full
// histogram of the distances in shared memory (warp-aggregated atomics: the counts cluster on few
// values), from which T, n_lt and n_eq follow; the fill kernel gives every warp a contiguous span of the
// row, counts first, then writes the "< T" pairs and the "== T" indices at warp-local offsets (ballots
// keep the index order), two block barriers in all.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_select_hist(const uint16_t *__restrict__ D, int N, int n, int nbins,
                                                     unsigned *__restrict__ Tkey, int *__restrict__ n_lt,
                                                     int *__restrict__ n_eq) {
  extern __shared__ unsigned hist[];          // nbins counters
  __shared__ unsigned warp_sum[16];
  const uint16_t *row = D + (size_t)blockIdx.x * N;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < N; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool valid = i < N;
    const unsigned k = valid ? row[i] : 0x10000u + lane;      // a key nobody shares
    const unsigned peers = __match_any_sync(0xffffffffu, k);
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[k], (unsigned)__popc(peers));
  }
  __syncthreads();
  // the bin that holds rank n: every thread sums a contiguous chunk of bins, block scan of the partials
  const int chunk = (nbins + (int)blockDim.x - 1) / (int)blockDim.x;
  const int b0 = min(nbins, (int)threadIdx.x * chunk), b1 = min(nbins, b0 + chunk);
  unsigned local = 0;
  for (int b = b0; b < b1; b++) local += hist[b];
  unsigned incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  unsigned base = 0;
  for (int w = 0; w < wid; w++) base += warp_sum[w];
  const unsigned excl = base + incl - local;
  if ((unsigned)n >= excl && (unsigned)n < excl + local) {
    unsigned cum = excl;
    for (int b = b0; b < b1; b++) {
      const unsigned h = hist[b];
      if ((unsigned)n < cum + h) {
        Tkey[blockIdx.x] = (unsigned)b; n_lt[blockIdx.x] = (int)cum; n_eq[blockIdx.x] = (int)h;
        break;
      }
      cum += h;
    }
  }
}

__global__ void __launch_bounds__(512) k_select_fill_spans(const uint16_t *__restrict__ D, int N, int cap,
                                                           const unsigned *__restrict__ Tkey,
                                                           const long long *__restrict__ eq_off,
                                                           int *__restrict__ lt_idx, unsigned *__restrict__ lt_key,
                                                           int *__restrict__ eq_pool) {
  extern __shared__ u64 keys[];          // cap2 = next pow2 >= cap
  __shared__ unsigned cnt_lt[16], cnt_eq[16];
  const int q = blockIdx.x;
  const uint16_t *row = D + (size_t)q * N;
  const unsigned T = Tkey[q];
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  for (int i = threadIdx.x; i < cap2; i += blockDim.x) keys[i] = ~0ull;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int span = ((N + nw - 1) / nw + 31) / 32 * 32;          // a multiple of the warp width
  const int s0 = min(N, wid * span), s1 = min(N, s0 + span);
  unsigned c_lt = 0, c_eq = 0;
  for (int i = s0 + lane; i < s1; i += 32) {
    const unsigned k = row[i];
    c_lt += (k < T); c_eq += (k == T);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_lt += __shfl_xor_sync(0xffffffffu, c_lt, o);
    c_eq += __shfl_xor_sync(0xffffffffu, c_eq, o);
  }
  if (lane == 0) { cnt_lt[wid] = c_lt; cnt_eq[wid] = c_eq; }
  __syncthreads();
  unsigned lt_at = 0, eq_at = 0;
  for (int w = 0; w < wid; w++) { lt_at += cnt_lt[w]; eq_at += cnt_eq[w]; }
  int *eq_out = eq_pool + eq_off[q];
  const unsigned below = (1u << lane) - 1;
  for (int i0 = s0; i0 < s1; i0 += 32) {
    const int i = i0 + lane;
    const unsigned k = i < s1 ? row[i] : 0xffffffffu;
    const unsigned b_lt = __ballot_sync(0xffffffffu, k < T), b_eq = __ballot_sync(0xffffffffu, k == T);
    if (k < T) {
      const unsigned slot = lt_at + __popc(b_lt & below);
      if (slot < (unsigned)cap2) keys[slot] = ((u64)k << 32) | (unsigned)i;
    }
    if (k == T) eq_out[eq_at + __popc(b_eq & below)] = i;
    lt_at += __popc(b_lt); eq_at += __popc(b_eq);
  }
  __syncthreads();
  // bitonic sort of the "< T" keys (padding = all ones sorts last)
  for (int size = 2; size <= cap2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < cap2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = ((i & size) == 0);
          const u64 a = keys[i], b = keys[j];
          if ((a > b) == up) { keys[i] = b; keys[j] = a; }
        }
      }
      __syncthreads();
    }
  unsigned nlt = 0;
  for (int w = 0; w < nw; w++) nlt += cnt_lt[w];
  const int take = min((int)nlt, cap);
  for (int i = threadIdx.x; i < take; i += blockDim.x) {
    lt_idx[(size_t)q * cap + i] = (int)(keys[i] & 0xffffffffu);
    lt_key[(size_t)q * cap + i] = (unsigned)(keys[i] >> 32);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {

// device times of the last batched call on this thread (mcq_closest_last_timing)
thread_local float g_ms_hamming = 0.f, g_ms_select = 0.f;

struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  float *out;
  EventTimer(cudaStream_t s, float *o) : st(s), out(o) {
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
  }
  ~EventTimer() {
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *out += ms;
    cudaEventDestroy(a); cudaEventDestroy(b);
  }
};

struct DevBuf {
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <typename T> T *as() { return (T *)p; }
};

int pack_to_device(const char *seqs, size_t stride, int n, int nbases, int W, bool seq_major, size_t npad,
                   DevBuf &planes, cudaStream_t st) {
  DevBuf raw;
  CK(cudaMalloc(&raw.p, (size_t)n * stride));
  CK(cudaMemcpyAsync(raw.p, seqs, (size_t)n * stride, cudaMemcpyHostToDevice, st));
  const size_t words = seq_major ? (size_t)n * W * 3 : (size_t)W * 3 * npad;
  CK(cudaMalloc(&planes.p, words * sizeof(u64)));
  // word-major padding columns: unknown everywhere (never written to D anyway)
  CK(cudaMemsetAsync(planes.p, 0, words * sizeof(u64), st));
  k_pack<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, st>>>((const unsigned char *)raw.p, stride, n, nbases,
                                                                    W, planes.as<u64>(), seq_major ? 1 : 0, npad);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  return 0;
}

const int kTQ = 32, kRT = 2;

template <int TQ>
int hamming_launch(DevBuf &refp, size_t npad, int N, DevBuf &qryp, int Q, int W, uint16_t *D, cudaStream_t st) {
  const size_t smem = (size_t)TQ * W * 3 * sizeof(u64);
  CK(cudaFuncSetAttribute(k_hamming<TQ, kRT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(npad / (128 * kRT)), (unsigned)((Q + TQ - 1) / TQ));
  k_hamming<TQ, kRT><<<grid, 128, smem, st>>>(refp.as<u64>(), npad, N, qryp.as<u64>(), Q, W, D);
  CK(cudaGetLastError());
  return 0;
}

int hamming_device(const DevBuf &refp_, size_t npad, int N, const DevBuf &qryp_, int Q, int W, uint16_t *D,
                   cudaStream_t st) {
  DevBuf &refp = const_cast<DevBuf &>(refp_), &qryp = const_cast<DevBuf &>(qryp_);
  // 32 queries per block while their bit planes fit in shared memory (up to ~17 000 bases), else 8
  if ((size_t)kTQ * W * 3 * sizeof(u64) <= 200 * 1024) return hamming_launch<kTQ>(refp, npad, N, qryp, Q, W, D, st);
  return hamming_launch<8>(refp, npad, N, qryp, Q, W, D, st);
}

// closest_seqs.c:57-64 + util.c:24-30: forward shuffle with j = (int)((i+1) * random()/RAND_MAX), applied to
// the reference's window of the sorted list: element n-1 extended over its neighbours equal to the distance
// T of element n (closest_seqs.c:85-88).  With the "< T" group (n_lt <= n elements, sorted on the device)
// and the "== T" group (in index order) that window starts at lo = min(n_lt, n-1) and runs to the end of the
// "== T" group; everything before lo is output as sorted.
void tie_shuffle_and_take(const int *lt_sorted, int n_lt, const int *eq_group, long long n_eq, int n,
                          std::vector<int> &win, int *out) {
  const int lo = n_lt < n - 1 ? n_lt : n - 1;
  memcpy(out, lt_sorted, (size_t)lo * sizeof(int));
  win.clear();
  for (int i = lo; i < n_lt; i++) win.push_back(lt_sorted[i]);          // at most element n-1
  win.insert(win.end(), eq_group, eq_group + n_eq);
  const int M = (int)win.size();
  for (int i = 0; i < M; i++) {
    const double u = (i + 1) * ((double)random() / (double)RAND_MAX);
    int j = (int)u;
    if (j > i) j = i;        // random() == RAND_MAX: the reference swaps with element i+1, past the window (UB)
    std::swap(win[i], win[j]);
  }
  memcpy(out + lo, win.data(), (size_t)(n - lo) * sizeof(int));
}

// device selection over rows of D (V = uint16_t counts or int32 values), host shuffle, output
template <typename V>
int select_rows(const V *D, int Q, int N, int n, int skip_first, int L, int *closest_refs, cudaStream_t st,
                int max_key = -1) {
  DevBuf dT, dlt, deq, doff, dlti, dltk, dpool;
  CK(cudaMalloc(&dT.p, (size_t)Q * sizeof(unsigned)));
  CK(cudaMalloc(&dlt.p, (size_t)Q * sizeof(int)));
  CK(cudaMalloc(&deq.p, (size_t)Q * sizeof(int)));
  {
    EventTimer t(st, &g_ms_select);
    bool done = false;
    if constexpr (sizeof(V) == 2) {
      // the full histogram of the counts fits in shared memory: one read of the row
      const size_t hbytes = (size_t)(max_key + 1) * sizeof(unsigned);
      if (max_key >= 0 && hbytes <= 200 * 1024) {
        CK(cudaFuncSetAttribute(k_select_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hbytes));
        k_select_hist<<<Q, 512, hbytes, st>>>(D, N, n, max_key + 1, dT.as<unsigned>(), dlt.as<int>(), deq.as<int>());
        done = true;
      }
    }
    if (!done) k_select_count<V><<<Q, 512, 0, st>>>(D, N, n, dT.as<unsigned>(), dlt.as<int>(), deq.as<int>());
    CK(cudaGetLastError());
  }
  std::vector<int> n_lt(Q), n_eq(Q);
  std::vector<unsigned> Tk(Q);
  CK(cudaMemcpyAsync(n_lt.data(), dlt.p, (size_t)Q * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(n_eq.data(), deq.p, (size_t)Q * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(Tk.data(), dT.p, (size_t)Q * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<long long> off(Q + 1, 0);
  for (int q = 0; q < Q; q++) off[q + 1] = off[q] + n_eq[q];
  const int cap = n;   // "< T" holds at most n elements
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  CK(cudaMalloc(&doff.p, (size_t)(Q + 1) * sizeof(long long)));
  CK(cudaMalloc(&dlti.p, (size_t)Q * cap * sizeof(int)));
  CK(cudaMalloc(&dltk.p, (size_t)Q * cap * sizeof(unsigned)));
  CK(cudaMalloc(&dpool.p, (size_t)std::max<long long>(off[Q], 1) * sizeof(int)));
  CK(cudaMemcpyAsync(doff.p, off.data(), (size_t)(Q + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)cap2 * sizeof(u64);
  {
    EventTimer t(st, &g_ms_select);
    if constexpr (sizeof(V) == 2) {
      CK(cudaFuncSetAttribute(k_select_fill_spans, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_select_fill_spans<<<Q, 512, smem, st>>>(D, N, cap, dT.as<unsigned>(), doff.as<long long>(), dlti.as<int>(),
                                                dltk.as<unsigned>(), dpool.as<int>());
    } else {
      CK(cudaFuncSetAttribute(k_select_fill<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_select_fill<V><<<Q, 512, smem, st>>>(D, N, cap, dT.as<unsigned>(), doff.as<long long>(), dlti.as<int>(),
                                             dltk.as<unsigned>(), dpool.as<int>());
    }
    CK(cudaGetLastError());
  }
  // (the keys of the "< T" group stay on the device: the host only needs the sorted indices)
  std::unique_ptr<int[]> lti(new int[(size_t)Q * cap]), pool(new int[(size_t)std::max<long long>(off[Q], 1)]);
  CK(cudaMemcpyAsync(lti.get(), dlti.p, (size_t)Q * cap * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pool.get(), dpool.p, (size_t)off[Q] * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  // host: the reference's sequential tie shuffle, in query order (one shared libc random() stream)
  std::vector<int> win, tmp(n);
  for (int q = 0; q < Q; q++) {
    tie_shuffle_and_take(lti.get() + (size_t)q * cap, n_lt[q], pool.get() + off[q], off[q + 1] - off[q], n, win, tmp.data());
    if (skip_first) memcpy(closest_refs + (size_t)q * L, tmp.data() + 1, (size_t)L * sizeof(int));   // dmodel.c:380-386
    else memcpy(closest_refs + (size_t)q * L, tmp.data(), (size_t)L * sizeof(int));
  }
  return 0;
}

struct StreamGuard {
  cudaStream_t s = nullptr;
  ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};

int distances(int device, const char *ref_seqs, size_t ref_stride, int N, const char *query_seqs,
              size_t query_stride, int Q, int nbases, DevBuf &D, StreamGuard &sg) {
  REQUIRE(ref_seqs && query_seqs, "null argument");
  REQUIRE(N > 0 && Q > 0 && nbases > 0, "dimensions must be positive");
  REQUIRE(nbases <= 65535, "num_bases > 65535 is not supported");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(ndev > 0, "no CUDA device: libmcq_b200 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device ordinal out of range");
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
  const int W = (nbases + 63) / 64;
  const size_t tile = 128 * kRT;
  const size_t npad = ((size_t)N + tile - 1) / tile * tile;
  DevBuf refp, qryp;
  if (pack_to_device(ref_seqs, ref_stride, N, nbases, W, false, npad, refp, sg.s)) return 1;
  if (pack_to_device(query_seqs, query_stride, Q, nbases, W, true, 0, qryp, sg.s)) return 1;
  CK(cudaMalloc(&D.p, (size_t)Q * N * sizeof(uint16_t)));
  g_ms_hamming = 0.f; g_ms_select = 0.f;
  {
    EventTimer t(sg.s, &g_ms_hamming);
    if (hamming_device(refp, npad, N, qryp, Q, W, D.as<uint16_t>(), sg.s)) return 1;
  }
  CK(cudaStreamSynchronize(sg.s));
  return 0;
}

}  // namespace

extern "C" int mcq_closest_last_timing(float *hamming_ms, float *select_ms) {
  if (hamming_ms) *hamming_ms = g_ms_hamming;
  if (select_ms) *select_ms = g_ms_select;
  return 0;
}

extern "C" int mcq_hamming_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                 const char *query_seqs, size_t query_stride, int num_query,
                                 int num_bases, int *hamming) {
  REQUIRE(hamming, "null argument");
  DevBuf D;
  StreamGuard sg;
  if (distances(device, ref_seqs, ref_stride, num_refs, query_seqs, query_stride, num_query, num_bases, D, sg))
    return 1;
  std::vector<uint16_t> h((size_t)num_query * num_refs);
  CK(cudaMemcpy(h.data(), D.p, h.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < h.size(); i++) hamming[i] = 2 * (int)h[i];   // closest_seqs.c:45
  return 0;
}

extern "C" int mcq_closest_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                 const char *query_seqs, size_t query_stride, int num_query,
                                 int num_bases, int closest_n, int skip_first, int *closest_refs) {
  REQUIRE(closest_refs, "null argument");
  const int n = skip_first ? closest_n + 1 : closest_n;
  REQUIRE(closest_n > 0, "closest_n must be positive");
  REQUIRE(n < num_refs, "closest_n must be smaller than num_refs (closest_seqs.c:85 reads element n)");
  DevBuf D;
  StreamGuard sg;
  if (distances(device, ref_seqs, ref_stride, num_refs, query_seqs, query_stride, num_query, num_bases, D, sg))
    return 1;
  return select_rows<uint16_t>(D.as<uint16_t>(), num_query, num_refs, n, skip_first, closest_n, closest_refs, sg.s,
                               num_bases);
}

// ------------------------------------------------------------------------------------------
// per-query consensus of the closest references (generate_consensus_subset, amino_acids.c:118-141,
// as used by dmodel.c:664-681): majority base per position over the query's L references, ties to
// the first of T,C,A,G, 'T' when no reference has a base there.  One thread per (query, position),
// references walked in list order (reads coalesced along the position).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_consensus(const unsigned char *__restrict__ refs, size_t stride,
                                                   const int *__restrict__ closest, int L, int nbases,
                                                   signed char *__restrict__ out) {
  const int q = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbases) return;
  const int *cl = closest + (size_t)q * L;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int r = 0; r < L; r++) {
    const unsigned char ch = refs[(size_t)cl[r] * stride + i];
    int code = 4;
    if ((ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z')) code = base_code(ch);
    c0 += (code == 0); c1 += (code == 1); c2 += (code == 2); c3 += (code == 3);
  }
  int best = 0, cnt = c0;
  if (c1 > cnt) { best = 1; cnt = c1; }
  if (c2 > cnt) { best = 2; cnt = c2; }
  if (c3 > cnt) { best = 3; cnt = c3; }
  out[(size_t)q * nbases + i] = (signed char)best;
}

extern "C" int mcq_consensus_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                   const int *closest_refs, int num_query, int closest_n, int num_bases,
                                   char *cons_codes) {
  REQUIRE(ref_seqs && closest_refs && cons_codes, "null argument");
  REQUIRE(num_refs > 0 && num_query > 0 && closest_n > 0 && num_bases > 0, "dimensions must be positive");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(ndev > 0, "no CUDA device: libmcq_b200 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device ordinal out of range");
  for (size_t i = 0; i < (size_t)num_query * closest_n; i++)
    REQUIRE(closest_refs[i] >= 0 && closest_refs[i] < num_refs, "closest_refs entry out of range");
  CK(cudaSetDevice(device));
  DevBuf d_refs, d_cl, d_out;
  CK(cudaMalloc(&d_refs.p, (size_t)num_refs * ref_stride));
  CK(cudaMalloc(&d_cl.p, (size_t)num_query * closest_n * sizeof(int)));
  CK(cudaMalloc(&d_out.p, (size_t)num_query * num_bases));
  CK(cudaMemcpy(d_refs.p, ref_seqs, (size_t)num_refs * ref_stride, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cl.p, closest_refs, (size_t)num_query * closest_n * sizeof(int), cudaMemcpyHostToDevice));
  for (int q0 = 0; q0 < num_query; q0 += 65535) {
    const int nq = std::min(65535, num_query - q0);
    dim3 grid((unsigned)((num_bases + 255) / 256), (unsigned)nq);
    k_consensus<<<grid, 256>>>(d_refs.as<unsigned char>(), ref_stride, d_cl.as<int>() + (size_t)q0 * closest_n, closest_n,
                               num_bases, d_out.as<signed char>() + (size_t)q0 * num_bases);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(cons_codes, d_out.p, (size_t)num_query * num_bases, cudaMemcpyDeviceToHost));
  return 0;
}

// ------------------------------------------------------------------------------------------
// drop-in symbols (closest_seqs.h:4-9)
// ------------------------------------------------------------------------------------------
namespace {
[[noreturn]] void die_msg(const char *what) {
  fflush(stdout);
  fprintf(stderr, "[libmcq_b200] Error: %s: %s\n", what, mcq_last_error());
  exit(EXIT_FAILURE);
}

// packed reference panel cached across calls: dmodel calls once per query with the same panel
// (dmodel.c:368-377).  The cache is validated by a 64-bit hash of the panel's bytes on every call.
struct RefCache {
  DevBuf planes;
  size_t npad = 0;
  int N = 0, nbases = 0, W = 0;
  u64 hash = 0;
  bool valid = false;
  cudaStream_t st = nullptr;
} g_refs;

u64 hash_rows(char **rows, int N, int nb) {
  u64 h = 1469598103934665603ull;
  for (int r = 0; r < N; r++) {
    const unsigned char *p = (const unsigned char *)rows[r];
    int i = 0;
    for (; i + 8 <= nb; i += 8) {
      u64 v;
      memcpy(&v, p + i, 8);
      h = (h ^ v) * 1099511628211ull;
      h ^= h >> 29;
    }
    for (; i < nb; i++) h = (h ^ p[i]) * 1099511628211ull;
  }
  return h;
}
}  // namespace

extern "C" void hamming_distance_considering_unknown(char **ref_seqs, char *query_seqs, int num_refs,
                                                     int num_bases, int *hamming) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "libmcq_b200: no CUDA device available (there is no CPU fallback)\n");
    exit(EXIT_FAILURE);
  }
  const int W = (num_bases + 63) / 64;
  const size_t tile = 128 * kRT;
  const u64 h = hash_rows(ref_seqs, num_refs, num_bases);
  if (!g_refs.valid || g_refs.N != num_refs || g_refs.nbases != num_bases || g_refs.hash != h) {
    if (!g_refs.st && cudaStreamCreateWithFlags(&g_refs.st, cudaStreamNonBlocking) != cudaSuccess) die_msg("stream");
    std::vector<char> flat((size_t)num_refs * num_bases);
    for (int r = 0; r < num_refs; r++) memcpy(&flat[(size_t)r * num_bases], ref_seqs[r], num_bases);
    if (g_refs.planes.p) { cudaFree(g_refs.planes.p); g_refs.planes.p = nullptr; }
    g_refs.npad = ((size_t)num_refs + tile - 1) / tile * tile;
    if (pack_to_device(flat.data(), num_bases, num_refs, num_bases, W, false, g_refs.npad, g_refs.planes, g_refs.st))
      die_msg("packing the reference panel");
    g_refs.N = num_refs; g_refs.nbases = num_bases; g_refs.W = W; g_refs.hash = h; g_refs.valid = true;
  }
  DevBuf qryp, D;
  if (pack_to_device(query_seqs, num_bases, 1, num_bases, W, true, 0, qryp, g_refs.st)) die_msg("packing the query");
  if (cudaMalloc(&D.p, (size_t)num_refs * sizeof(uint16_t)) != cudaSuccess) die_msg("cudaMalloc");
  if (hamming_device(g_refs.planes, g_refs.npad, num_refs, qryp, 1, W, D.as<uint16_t>(), g_refs.st)) die_msg("k_hamming");
  std::vector<uint16_t> hcount(num_refs);
  if (cudaMemcpyAsync(hcount.data(), D.p, (size_t)num_refs * sizeof(uint16_t), cudaMemcpyDeviceToHost, g_refs.st) != cudaSuccess ||
      cudaStreamSynchronize(g_refs.st) != cudaSuccess)
    die_msg("reading distances back");
  for (int r = 0; r < num_refs; r++) hamming[r] = 2 * (int)hcount[r];
}

extern "C" void closest_sequences(int num_refs, int *hamming, int closest_n, int *closest_refs) {
  if (num_refs <= 0 || closest_n <= 0 || closest_n >= num_refs) {
    // closest_seqs.c:69-70 asserts the first two; :85 reads element closest_n
    fprintf(stderr, "libmcq_b200: closest_sequences needs 0 < closest_n < num_refs (got %d, %d)\n", closest_n, num_refs);
    exit(EXIT_FAILURE);
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "libmcq_b200: no CUDA device available (there is no CPU fallback)\n");
    exit(EXIT_FAILURE);
  }
  StreamGuard sg;
  DevBuf D;
  if (cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&D.p, (size_t)num_refs * sizeof(int)) != cudaSuccess ||
      cudaMemcpyAsync(D.p, hamming, (size_t)num_refs * sizeof(int), cudaMemcpyHostToDevice, sg.s) != cudaSuccess)
    die_msg("uploading distances");
  if (select_rows<int>(D.as<int>(), 1, num_refs, closest_n, 0, closest_n, closest_refs, sg.s)) die_msg("selection");
}
