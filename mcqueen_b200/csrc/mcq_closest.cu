// mcq_closest.cu - Hamming top-L reference selection (closest_seqs.c) on the GPU.
//
//   K1a k_pack      ASCII -> three bit planes per sequence (code low bit, code high bit, "unknown")
//   K1b k_hamming   bit-parallel mismatch counts for a tile of queries x references
//   K2  k_select_row    per query, one block: histogram of the row of counts -> the distance T of sorted element
//                       `n`, the "< T" group sorted by (distance, index), the "== T" group in index order
//       k_select_count / k_select_fill: the same in two kernels by radix select, for the int32 distances of the
//                       one-query drop-in symbol and for sequences too long for a shared-memory histogram
//   host            replays the reference's boundary-tie shuffle with libc random() in query order
// Queries are processed in chunks of at most 256 MB of distances; nothing of size Q x N is allocated.
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
                                                 uint16_t *__restrict__ D, size_t ldD) {
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
        if ((size_t)ridx[k] < ldD)      // columns N .. ldD-1 pad the row: a count no reference can have
          D[(size_t)(q0 + q) * ldD + ridx[k]] = ridx[k] < N ? (uint16_t)acc[q][k] : (uint16_t)0xffff;
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
// K2 (batched path): the whole selection of one query in one block.  The row of 16-bit mismatch counts (pitch a
// multiple of 8 entries, so every thread reads 16 bytes at a time) belongs to a chunk of queries whose
// distances k_hamming has just written and that still sits in L2.
//   1. histogram of the counts in shared memory (warp-aggregated atomics: the counts cluster on few values);
//   2. the bin that holds sorted element n: T, #(< T), #(== T); the "== T" group gets its place in the chunk's
//      pool from a global cursor;
//   3. every warp counts the "< T" / "== T" entries of its contiguous span of the row;
//   4. and writes them - the "< T" pairs into shared memory, the "== T" indices into the pool IN INDEX ORDER (a
//      warp scan over the lanes' 8-entry masks; spans without either kind cost one vote);
//   5. the "< T" group is sorted by (distance, index) and written out.
// ------------------------------------------------------------------------------------------
struct SelectOut {
  unsigned *Tkey; int *n_lt; int *n_eq; long long *eq_off;   // [Qc]
  int *lt_idx;                                              // [Qc][cap]
  int *eq_pool; long long pool_cap; unsigned long long *cursor; int *overflow;
  u64 *cand; int cand_cap;                                  // [Qc][cand_cap] scratch: (count << 32 | index) of the entries <= U
  int sort_cap;                                             // entries of the shared-memory sort buffer (a power of two)
};

__global__ void __launch_bounds__(512) k_select_row(const uint16_t *__restrict__ D, size_t ldD, int N, int n, int nbins,
                                                    int cap, int cap2, SelectOut o) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64 *keys = reinterpret_cast<u64 *>(smem_raw);                       // sort_cap (>= cap2)
  unsigned *hist = reinterpret_cast<unsigned *>(keys + o.sort_cap);    // nbins
  __shared__ unsigned warp_sum[16], cnt_lt[16], cnt_eq[16];
  __shared__ unsigned s_T, s_nlt, s_neq;
  __shared__ long long s_off;
  const int q = blockIdx.x;
  const uint4 *__restrict__ row = reinterpret_cast<const uint4 *>(D + (size_t)q * ldD);
  const int nvec = (N + 7) >> 3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
  for (int i = threadIdx.x; i < cap2; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  // ---- 1. histogram.  First of a head of the row only (warp-aggregated: MATCH is the expensive instruction
  // here): the element of rank n among the head is an upper bound U of the row's element of rank n, and only
  // the few per cent of the remaining entries that are <= U are counted at all.
  const int nhead = min(nvec, max(16 * n, 4096) >> 3);
  __shared__ unsigned s_cand;
  if (threadIdx.x == 0) s_cand = 0;
  for (int v0 = wid * 32; v0 < nhead; v0 += blockDim.x) {
    const int v = v0 + lane;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (v < nhead) x = row[v];
    const unsigned w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const bool valid = v < nhead && v * 8 + e < N;
      const unsigned k = valid ? ((w[e >> 1] >> (16 * (e & 1))) & 0xffffu) : 0x10000u + lane;   // a key nobody shares
      const unsigned peers = __match_any_sync(0xffffffffu, k);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&hist[k], (unsigned)__popc(peers));
    }
  }
  __syncthreads();
  // the bin that holds rank `rank`: every thread sums a contiguous chunk of bins, block scan of the partials
  auto find_bin = [&](unsigned rank, bool final) {
    const int chunk = (nbins + (int)blockDim.x - 1) / (int)blockDim.x;
    const int b0 = min(nbins, (int)threadIdx.x * chunk), b1 = min(nbins, b0 + chunk);
    unsigned local = 0;
    for (int b2 = b0; b2 < b1; b2++) local += hist[b2];
    unsigned incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    unsigned base = 0;
    for (int w2 = 0; w2 < wid; w2++) base += warp_sum[w2];
    const unsigned excl = base + incl - local;
    if (rank >= excl && rank < excl + local) {
      unsigned cum = excl;
      for (int b2 = b0; b2 < b1; b2++) {
        const unsigned h = hist[b2];
        if (rank < cum + h) {
          s_T = (unsigned)b2; s_nlt = cum; s_neq = h;
          if (final) {        // the "== T" group's place in the chunk's pool
            const unsigned long long at = atomicAdd(o.cursor, (unsigned long long)h);
            s_off = (long long)at;
            if ((long long)(at + h) > o.pool_cap) *o.overflow = 1;
          }
          break;
        }
        cum += h;
      }
    }
    __syncthreads();
  };
  // (the head holds more than n entries unless it is the whole row)
  const int head_count = min(N, nhead * 8);
  find_bin((unsigned)min(n, head_count - 1), false);
  const unsigned U = s_T;
  __syncthreads();
  // the one pass over the whole row: entries <= U are kept, with their index, as candidates.  Every warp owns a
  // contiguous span of the row and a region of the candidate scratch of its own, which it fills in index order
  // (a warp scan over the lanes' counts; no atomics in this pass).  The pad columns of the row hold 0xffff.  A
  // region that overflows sends the block down the slow path below.
  u64 *__restrict__ cand = o.cand + (size_t)q * o.cand_cap;
  const unsigned below = (1u << lane) - 1;
  const int region = o.cand_cap / nw;
  u64 *__restrict__ mycand = cand + (size_t)wid * region;
  const int cspan = (nvec + nw - 1) / nw;
  const int c0 = min(nvec, wid * cspan), c1 = min(nvec, c0 + cspan);
  unsigned mine_n = 0;                       // candidates of this warp so far (the same in every lane)
  uint4 nx = make_uint4(~0u, ~0u, ~0u, ~0u);
  if (c0 + lane < c1) nx = row[c0 + lane];
  for (int v0 = c0; v0 < c1; v0 += 32) {
    const int v = v0 + lane;
    const uint4 x = nx;                       // (the next 512 bytes are requested before these are looked at)
    nx = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (v + 32 < c1) nx = row[v + 32];
    const unsigned w[4] = {x.x, x.y, x.z, x.w};
    unsigned mk = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) mk |= (unsigned)(((w[e >> 1] >> (16 * (e & 1))) & 0xffffu) <= U) << e;
    if (!__any_sync(0xffffffffu, mk != 0)) continue;
    const unsigned cnt = (unsigned)__popc(mk);
    unsigned incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    unsigned slot = mine_n + incl - cnt;
    while (mk != 0) {
      const int e = __ffs(mk) - 1;
      mk &= mk - 1;
      if (slot < (unsigned)region) mycand[slot] = ((u64)((w[e >> 1] >> (16 * (e & 1))) & 0xffffu) << 32) | (unsigned)(v * 8 + e);
      slot++;
    }
    mine_n += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0 && mine_n > (unsigned)region) atomicExch(&s_cand, 0xffffffffu);
  __syncthreads();
  const bool fast = s_cand != 0xffffffffu;
  if (fast) {
    // the candidates beyond the head join the histogram (warp-aggregated; the head's entries are in it already)
    for (unsigned i0 = 0; i0 < mine_n; i0 += 32) {
      const unsigned i = i0 + lane;
      const u64 c = i < mine_n ? mycand[i] : 0;
      const bool cnt = i < mine_n && (int)((unsigned)c >> 3) >= nhead;
      const unsigned k = cnt ? (unsigned)(c >> 32) : 0x10000u + lane;
      const unsigned peers = __match_any_sync(0xffffffffu, k);
      if (cnt && lane == __ffs(peers) - 1) atomicAdd(&hist[k], (unsigned)__popc(peers));
    }
  } else {
    // a region overflowed: count the whole row beyond the head the plain way
    for (int v = nhead + (int)threadIdx.x; v < nvec; v += blockDim.x) {
      const uint4 x = row[v];
      const unsigned w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const unsigned k = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
        if (k <= U && k != 0xffffu) atomicAdd(&hist[k], 1u);
      }
    }
  }
  __syncthreads();
  // ---- 2. the bin of rank n in the whole row (bins above U are incomplete and not needed: T <= U)
  find_bin((unsigned)n, true);
  const unsigned T = s_T;
  const long long off = s_off;
  const bool room = off + (long long)s_neq <= o.pool_cap;
  if (threadIdx.x == 0) { o.Tkey[q] = T; o.n_lt[q] = (int)s_nlt; o.n_eq[q] = (int)s_neq; o.eq_off[q] = off; }
  // ---- the usual case: every entry <= T is among the candidates, whose regions, taken in warp order, are in index
  // order.  The "== T" entries go to the pool as they come, the "< T" entries (at most n) are sorted by (distance,
  // index).
  if (fast) {
    unsigned c_lt = 0, c_eq = 0;
    for (unsigned i = lane; i < mine_n; i += 32) {
      const unsigned k = (unsigned)(mycand[i] >> 32);
      c_lt += k < T; c_eq += k == T;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      c_lt += __shfl_xor_sync(0xffffffffu, c_lt, d);
      c_eq += __shfl_xor_sync(0xffffffffu, c_eq, d);
    }
    if (lane == 0) { cnt_lt[wid] = c_lt; cnt_eq[wid] = c_eq; }
    __syncthreads();
    unsigned lt_at = 0, eq_at = 0;
    for (int w2 = 0; w2 < wid; w2++) { lt_at += cnt_lt[w2]; eq_at += cnt_eq[w2]; }
    for (unsigned i0 = 0; i0 < mine_n; i0 += 32) {
      const unsigned i = i0 + lane;
      const u64 c = i < mine_n ? mycand[i] : ~0ull;
      const unsigned k = (unsigned)(c >> 32);
      const unsigned b_lt = __ballot_sync(0xffffffffu, k < T), b_eq = __ballot_sync(0xffffffffu, k == T);
      if (k < T) keys[lt_at + __popc(b_lt & below)] = c;
      if (k == T && room) o.eq_pool[off + eq_at + __popc(b_eq & below)] = (int)(unsigned)c;
      lt_at += __popc(b_lt); eq_at += __popc(b_eq);
    }
    __syncthreads();
    for (int size = 2; size <= cap2; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < cap2; i += blockDim.x) {
          const int j = i ^ stride;
          if (j > i) {
            const bool up = ((i & size) == 0);
            const u64 x = keys[i], y = keys[j];
            if ((x > y) == up) { keys[i] = y; keys[j] = x; }
          }
        }
        __syncthreads();
      }
    const int take = min((int)s_nlt, cap);
    for (int i = threadIdx.x; i < take; i += blockDim.x) o.lt_idx[(size_t)q * cap + i] = (int)(keys[i] & 0xffffffffu);
    return;
  }
  // ---- otherwise (huge tie groups, degenerate panels): two more passes over the row
  // ---- 3. per-warp counts over contiguous spans of 16-byte vectors
  const int span = (nvec + nw - 1) / nw;
  const int s0 = min(nvec, wid * span), s1 = min(nvec, s0 + span);
  unsigned c_lt = 0, c_eq = 0;
  for (int v = s0 + lane; v < s1; v += 32) {
    const uint4 x = row[v];
    const unsigned w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const unsigned k = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
      const bool valid = v * 8 + e < N;
      c_lt += valid && k < T; c_eq += valid && k == T;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    c_lt += __shfl_xor_sync(0xffffffffu, c_lt, d);
    c_eq += __shfl_xor_sync(0xffffffffu, c_eq, d);
  }
  if (lane == 0) { cnt_lt[wid] = c_lt; cnt_eq[wid] = c_eq; }
  __syncthreads();
  unsigned lt_at = 0, eq_at = 0;
  for (int w2 = 0; w2 < wid; w2++) { lt_at += cnt_lt[w2]; eq_at += cnt_eq[w2]; }
  // ---- 4. write, in index order within and across warps
  int *eq_out = o.eq_pool + off;
  for (int v0 = s0; v0 < s1; v0 += 32) {
    const int v = v0 + lane;
    unsigned m_lt = 0, m_eq = 0;
    unsigned w[4] = {0, 0, 0, 0};
    if (v < s1) {
      const uint4 x = row[v];
      w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const unsigned k = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
        const bool valid = v * 8 + e < N;
        m_lt |= (unsigned)(valid && k < T) << e; m_eq |= (unsigned)(valid && k == T) << e;
      }
    }
    if (!__any_sync(0xffffffffu, (m_lt | m_eq) != 0)) continue;
    const unsigned mine = (unsigned)__popc(m_lt) | ((unsigned)__popc(m_eq) << 16);
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const unsigned excl = incl - mine, total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned a = lt_at + (excl & 0xffffu), b = eq_at + (excl >> 16);
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const unsigned k = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
      if ((m_lt >> e) & 1) { if (a < (unsigned)cap2) keys[a] = ((u64)k << 32) | (unsigned)(v * 8 + e); a++; }
      if ((m_eq >> e) & 1) { if (room) eq_out[b] = v * 8 + e; b++; }
    }
    lt_at += total & 0xffffu; eq_at += total >> 16;
  }
  __syncthreads();
  // ---- 5. bitonic sort of the "< T" keys (padding = all ones sorts last)
  for (int size = 2; size <= cap2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < cap2; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = ((i & size) == 0);
          const u64 x = keys[i], y = keys[j];
          if ((x > y) == up) { keys[i] = y; keys[j] = x; }
        }
      }
      __syncthreads();
    }
  const int take = min((int)s_nlt, cap);
  for (int i = threadIdx.x; i < take; i += blockDim.x) o.lt_idx[(size_t)q * cap + i] = (int)(keys[i] & 0xffffffffu);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
namespace {

// device times of the last batched call on this thread (mcq_closest_last_timing)
thread_local float g_ms_hamming = 0.f, g_ms_select = 0.f, g_ms_consensus = 0.f;

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
int hamming_launch(const u64 *refp, size_t npad, int N, const u64 *qryp, int Q, int W, uint16_t *D, size_t ldD,
                   cudaStream_t st) {
  const size_t smem = (size_t)TQ * W * 3 * sizeof(u64);
  CK(cudaFuncSetAttribute(k_hamming<TQ, kRT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(npad / (128 * kRT)), (unsigned)((Q + TQ - 1) / TQ));
  k_hamming<TQ, kRT><<<grid, 128, smem, st>>>(refp, npad, N, qryp, Q, W, D, ldD);
  CK(cudaGetLastError());
  return 0;
}

// distances of Q queries (seq-major planes at qryp) to the panel, rows of pitch ldD
int hamming_device(const u64 *refp, size_t npad, int N, const u64 *qryp, int Q, int W, uint16_t *D, size_t ldD,
                   cudaStream_t st) {
  // 32 queries per block while their bit planes fit in shared memory (up to ~17 000 bases), else 8
  if ((size_t)kTQ * W * 3 * sizeof(u64) <= 200 * 1024) return hamming_launch<kTQ>(refp, npad, N, qryp, Q, W, D, ldD, st);
  return hamming_launch<8>(refp, npad, N, qryp, Q, W, D, ldD, st);
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

// ---- selection groups ---------------------------------------------------------------------------------
// What the device hands to the sequential part: per query the "< T" group sorted by (distance, index) and the
// "== T" group in index order, as one flat int array
//     { kGroupsMagic, Q, n,   then per query:  n_lt, n_eq, lt[0..n_lt), eq[0..n_eq) }
// Blocks of this form travel between ranks (queries are sharded, the tie shuffles are not: every rank replays
// all blocks in query order, mcq_closest_replay).
const int kGroupsMagic = 0x4D435147;   // "MCQG"

// one chunk's results -> records appended to `flat`
void append_records(std::vector<int> &flat, int nq, int cap, const int *n_lt, const int *n_eq, const long long *eq_off,
                    const int *lt_idx, const int *pool) {
  for (int q = 0; q < nq; q++) {
    flat.push_back(n_lt[q]);
    flat.push_back(n_eq[q]);
    flat.insert(flat.end(), lt_idx + (size_t)q * cap, lt_idx + (size_t)q * cap + n_lt[q]);
    flat.insert(flat.end(), pool + eq_off[q], pool + eq_off[q] + n_eq[q]);
  }
}

// the radix-select kernels over rows of pitch N (int32 values of the one-query symbol; 16-bit counts of sequences
// too long for a shared-memory histogram): records of Q queries appended to `flat`
template <typename V>
int select_rows_radix(const V *D, int Q, int N, int n, cudaStream_t st, std::vector<int> &flat) {
  DevBuf dT, dlt, deq, doff, dlti, dltk, dpool;
  CK(cudaMalloc(&dT.p, (size_t)Q * sizeof(unsigned)));
  CK(cudaMalloc(&dlt.p, (size_t)Q * sizeof(int)));
  CK(cudaMalloc(&deq.p, (size_t)Q * sizeof(int)));
  {
    EventTimer t(st, &g_ms_select);
    k_select_count<V><<<Q, 512, 0, st>>>(D, N, n, dT.as<unsigned>(), dlt.as<int>(), deq.as<int>());
    CK(cudaGetLastError());
  }
  std::vector<int> n_lt(Q), n_eq(Q);
  CK(cudaMemcpyAsync(n_lt.data(), dlt.p, (size_t)Q * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(n_eq.data(), deq.p, (size_t)Q * sizeof(int), cudaMemcpyDeviceToHost, st));
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
    CK(cudaFuncSetAttribute(k_select_fill<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_select_fill<V><<<Q, 512, smem, st>>>(D, N, cap, dT.as<unsigned>(), doff.as<long long>(), dlti.as<int>(),
                                           dltk.as<unsigned>(), dpool.as<int>());
    CK(cudaGetLastError());
  }
  std::unique_ptr<int[]> lti(new int[(size_t)Q * cap]), pool(new int[(size_t)std::max<long long>(off[Q], 1)]);
  CK(cudaMemcpyAsync(lti.get(), dlti.p, (size_t)Q * cap * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pool.get(), dpool.p, (size_t)off[Q] * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  append_records(flat, Q, cap, n_lt.data(), n_eq.data(), off.data(), lti.get(), pool.get());
  return 0;
}

// the sequential part: the reference's tie shuffle with libc random(), query after query
int replay_block(const int *flat, size_t count, int closest_n, int skip_first, int *closest_refs, size_t *used) {
  REQUIRE(count >= 3 && flat[0] == kGroupsMagic, "not a block of selection groups");
  const int Q = flat[1], n = flat[2];
  REQUIRE(Q >= 0 && n == (skip_first ? closest_n + 1 : closest_n), "selection groups were made for another closest_n");
  size_t at = 3;
  std::vector<int> win, tmp((size_t)n);
  for (int q = 0; q < Q; q++) {
    REQUIRE(at + 2 <= count, "truncated block of selection groups");
    const int n_lt = flat[at], n_eq = flat[at + 1];
    at += 2;
    REQUIRE(n_lt >= 0 && n_lt <= n && n_eq >= 1 && at + (size_t)n_lt + (size_t)n_eq <= count && n_lt + (long long)n_eq > n,
            "inconsistent selection groups");
    tie_shuffle_and_take(flat + at, n_lt, flat + at + n_lt, n_eq, n, win, tmp.data());
    at += (size_t)n_lt + (size_t)n_eq;
    memcpy(closest_refs + (size_t)q * closest_n, tmp.data() + (skip_first ? 1 : 0), (size_t)closest_n * sizeof(int));   // dmodel.c:380-386
  }
  if (used) *used = at;
  return 0;
}

struct StreamGuard {
  cudaStream_t s = nullptr;
  ~StreamGuard() { if (s) cudaStreamDestroy(s); }
};

struct HostPin {             // pinned host staging (asynchronous read-back of a chunk under the next one's kernels)
  void *p = nullptr;
  size_t bytes = 0;
  ~HostPin() { if (p) cudaFreeHost(p); }
  int need(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr; bytes = 0;
    CK(cudaMallocHost(&p, n));
    bytes = n;
    return 0;
  }
  template <typename T> T *as() { return (T *)p; }
};

int check_device(int device) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(ndev > 0, "no CUDA device: libmcq_b200 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device ordinal out of range");
  CK(cudaSetDevice(device));
  return 0;
}

// Distances and selection groups of Q queries, chunk by chunk.  The distances of a chunk (at most kChunkBytes)
// are written by k_hamming and read once by k_select_row; chunks
// alternate between two streams and two sets of buffers, the host collects chunk k while the device works on
// chunk k+1.  Nothing of size Q x N is ever allocated.
const size_t kChunkBytes = 256u << 20;

struct ChunkBufs {
  DevBuf D, T, nlt, neq, off, lti, pool, cursor, overflow, cand;
  HostPin h_nlt, h_neq, h_off, h_lti, h_pool, h_misc;
  long long pool_cap = 0;
  cudaStream_t st = nullptr;
  int q0 = 0, nq = 0;
};

int closest_groups(int device, const char *ref_seqs, size_t ref_stride, int N, const char *query_seqs,
                   size_t query_stride, int Q, int nbases, int n, std::vector<int> &flat) {
  REQUIRE(ref_seqs && query_seqs, "null argument");
  REQUIRE(N > 0 && Q > 0 && nbases > 0, "dimensions must be positive");
  REQUIRE(nbases <= 65535, "num_bases > 65535 is not supported");
  REQUIRE(n > 0 && n < N, "closest_n must be smaller than num_refs (closest_seqs.c:85 reads element n)");
  if (check_device(device)) return 1;
  StreamGuard sg[2];
  for (auto &g : sg) CK(cudaStreamCreateWithFlags(&g.s, cudaStreamNonBlocking));
  const int W = (nbases + 63) / 64;
  const size_t tile = 128 * kRT;
  const size_t npad = ((size_t)N + tile - 1) / tile * tile;
  DevBuf refp, qryp;
  if (pack_to_device(ref_seqs, ref_stride, N, nbases, W, false, npad, refp, sg[0].s)) return 1;
  if (pack_to_device(query_seqs, query_stride, Q, nbases, W, true, 0, qryp, sg[0].s)) return 1;
  g_ms_hamming = 0.f; g_ms_select = 0.f;
  flat.clear();
  flat.push_back(kGroupsMagic); flat.push_back(Q); flat.push_back(n);

  const int cap = n;
  int cap2 = 1;
  while (cap2 < cap) cap2 <<= 1;
  const int nbins = nbases + 1;
  const int sort_cap = cap2;
  const size_t smem = (size_t)sort_cap * sizeof(u64) + (size_t)nbins * sizeof(unsigned);
  const bool one_block = smem <= 200 * 1024;
  int cand_cap = 4096;       // about twice the expected number of entries below the head's bound (N / 16)
  while (cand_cap < N / 8 && cand_cap < 65536) cand_cap <<= 1;
  const size_t ldD = one_block ? ((size_t)N + 7) / 8 * 8 : (size_t)N;
  int qc = (int)std::min<size_t>((size_t)Q, std::max<size_t>(kTQ, kChunkBytes / (ldD * sizeof(uint16_t)) / kTQ * kTQ));
  if (one_block) CK(cudaFuncSetAttribute(k_select_row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  ChunkBufs cb[2];
  cudaEvent_t ev[2][4];
  for (int b = 0; b < 2; b++) {
    cb[b].st = sg[b].s;
    for (auto &e : ev[b]) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&cb[b].D.p, (size_t)qc * ldD * sizeof(uint16_t)));
    if (!one_block) continue;
    CK(cudaMalloc(&cb[b].T.p, (size_t)qc * sizeof(unsigned)));
    CK(cudaMalloc(&cb[b].nlt.p, (size_t)qc * sizeof(int)));
    CK(cudaMalloc(&cb[b].neq.p, (size_t)qc * sizeof(int)));
    CK(cudaMalloc(&cb[b].off.p, (size_t)qc * sizeof(long long)));
    CK(cudaMalloc(&cb[b].lti.p, (size_t)qc * cap * sizeof(int)));
    CK(cudaMalloc(&cb[b].cand.p, (size_t)qc * cand_cap * sizeof(u64)));
    CK(cudaMalloc(&cb[b].cursor.p, sizeof(unsigned long long)));
    CK(cudaMalloc(&cb[b].overflow.p, sizeof(int)));
    cb[b].pool_cap = std::max<long long>(1 << 20, (long long)qc * 4096);
    CK(cudaMalloc(&cb[b].pool.p, (size_t)cb[b].pool_cap * sizeof(int)));
    if (cb[b].h_nlt.need((size_t)qc * sizeof(int)) || cb[b].h_neq.need((size_t)qc * sizeof(int)) ||
        cb[b].h_off.need((size_t)qc * sizeof(long long)) || cb[b].h_lti.need((size_t)qc * cap * sizeof(int)) ||
        cb[b].h_misc.need(16))
      return 1;
  }
  float ms_h[2] = {0.f, 0.f}, ms_s[2] = {0.f, 0.f};

  auto issue = [&](int b, int q0, int nq) -> int {         // hamming + selection of one chunk, results towards the host
    ChunkBufs &c = cb[b];
    c.q0 = q0; c.nq = nq;
    CK(cudaEventRecord(ev[b][0], c.st));
    if (hamming_device(refp.as<u64>(), npad, N, qryp.as<u64>() + (size_t)q0 * W * 3, nq, W, c.D.as<uint16_t>(), ldD, c.st))
      return 1;
    CK(cudaEventRecord(ev[b][1], c.st));
    if (!one_block) return 0;
    CK(cudaMemsetAsync(c.cursor.p, 0, sizeof(unsigned long long), c.st));
    CK(cudaMemsetAsync(c.overflow.p, 0, sizeof(int), c.st));
    SelectOut o{c.T.as<unsigned>(), c.nlt.as<int>(), c.neq.as<int>(), c.off.as<long long>(), c.lti.as<int>(),
                c.pool.as<int>(), c.pool_cap, c.cursor.as<unsigned long long>(), c.overflow.as<int>(),
                c.cand.as<u64>(), cand_cap, sort_cap};
    CK(cudaEventRecord(ev[b][2], c.st));
    k_select_row<<<nq, 512, smem, c.st>>>(c.D.as<uint16_t>(), ldD, N, n, nbins, cap, cap2, o);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[b][3], c.st));
    CK(cudaMemcpyAsync(c.h_nlt.p, c.nlt.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, c.st));
    CK(cudaMemcpyAsync(c.h_neq.p, c.neq.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, c.st));
    CK(cudaMemcpyAsync(c.h_off.p, c.off.p, (size_t)nq * sizeof(long long), cudaMemcpyDeviceToHost, c.st));
    CK(cudaMemcpyAsync(c.h_lti.p, c.lti.p, (size_t)nq * cap * sizeof(int), cudaMemcpyDeviceToHost, c.st));
    CK(cudaMemcpyAsync(c.h_misc.as<char>(), c.cursor.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st));
    CK(cudaMemcpyAsync(c.h_misc.as<char>() + 8, c.overflow.p, sizeof(int), cudaMemcpyDeviceToHost, c.st));
    return 0;
  };
  auto collect = [&](int b) -> int {                      // wait for a chunk and append its records
    ChunkBufs &c = cb[b];
    CK(cudaStreamSynchronize(c.st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev[b][0], ev[b][1]));
    ms_h[b] += ms;
    if (!one_block) {
      // sequences too long for the shared-memory histogram: the radix-select kernels on this chunk's rows
      float keep = g_ms_select;
      g_ms_select = 0.f;
      if (select_rows_radix<uint16_t>(c.D.as<uint16_t>(), c.nq, N, n, c.st, flat)) return 1;
      ms_s[b] += g_ms_select;
      g_ms_select = keep;
      return 0;
    }
    CK(cudaEventElapsedTime(&ms, ev[b][2], ev[b][3]));
    ms_s[b] += ms;
    const unsigned long long used = *c.h_misc.as<unsigned long long>();
    const int overflow = *(int *)(c.h_misc.as<char>() + 8);
    if (overflow) {
      // more boundary ties than the pool holds (degenerate panels): a pool of the exact size, the chunk once more
      c.pool_cap = (long long)used;
      if (c.pool.p) { cudaFree(c.pool.p); c.pool.p = nullptr; }
      CK(cudaMalloc(&c.pool.p, (size_t)c.pool_cap * sizeof(int)));
      if (issue(b, c.q0, c.nq)) return 1;
      return 2;
    }
    if (c.h_pool.need(std::max<size_t>((size_t)used, 1) * sizeof(int))) return 1;
    CK(cudaMemcpyAsync(c.h_pool.p, c.pool.p, (size_t)used * sizeof(int), cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    append_records(flat, c.nq, cap, c.h_nlt.as<int>(), c.h_neq.as<int>(), c.h_off.as<long long>(), c.h_lti.as<int>(),
                   c.h_pool.as<int>());
    return 0;
  };
  auto collect_done = [&](int b) -> int {
    for (;;) {
      const int r = collect(b);
      if (r != 2) return r;
    }
  };

  const bool serial = getenv("MCQ_CLOSEST_SERIAL") != nullptr;      // (measurement aid: no overlap between chunks)
  int pending = -1, k = 0;
  for (int q0 = 0; q0 < Q; q0 += qc, k++) {
    const int b = k & 1;
    if (issue(b, q0, std::min(qc, Q - q0))) return 1;
    if (serial) { if (collect_done(b)) return 1; continue; }
    if (pending >= 0 && collect_done(pending)) return 1;      // the previous chunk, while this one runs
    pending = b;
  }
  if (pending >= 0 && collect_done(pending)) return 1;
  g_ms_hamming = ms_h[0] + ms_h[1];
  g_ms_select = ms_s[0] + ms_s[1];
  for (int b = 0; b < 2; b++)
    for (auto &e : ev[b]) cudaEventDestroy(e);
  return 0;
}

}  // namespace

extern "C" int mcq_closest_last_timing(float *hamming_ms, float *select_ms) {
  if (hamming_ms) *hamming_ms = g_ms_hamming;
  if (select_ms) *select_ms = g_ms_select;
  return 0;
}

extern "C" int mcq_consensus_last_timing(float *kernel_ms) {
  if (kernel_ms) *kernel_ms = g_ms_consensus;
  return 0;
}

extern "C" int mcq_hamming_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                 const char *query_seqs, size_t query_stride, int num_query,
                                 int num_bases, int *hamming) {
  REQUIRE(hamming && ref_seqs && query_seqs, "null argument");
  REQUIRE(num_refs > 0 && num_query > 0 && num_bases > 0, "dimensions must be positive");
  REQUIRE(num_bases <= 65535, "num_bases > 65535 is not supported");
  if (check_device(device)) return 1;
  StreamGuard sg;
  CK(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
  const int W = (num_bases + 63) / 64;
  const size_t tile = 128 * kRT;
  const size_t npad = ((size_t)num_refs + tile - 1) / tile * tile;
  DevBuf refp, qryp, D;
  if (pack_to_device(ref_seqs, ref_stride, num_refs, num_bases, W, false, npad, refp, sg.s)) return 1;
  if (pack_to_device(query_seqs, query_stride, num_query, num_bases, W, true, 0, qryp, sg.s)) return 1;
  // chunks of queries, so that the 16-bit counts of a chunk never exceed kChunkBytes
  const int qc = (int)std::min<size_t>((size_t)num_query, std::max<size_t>(kTQ, kChunkBytes / ((size_t)num_refs * 2) / kTQ * kTQ));
  CK(cudaMalloc(&D.p, (size_t)qc * num_refs * sizeof(uint16_t)));
  std::vector<uint16_t> h((size_t)qc * num_refs);
  g_ms_hamming = 0.f; g_ms_select = 0.f;
  for (int q0 = 0; q0 < num_query; q0 += qc) {
    const int nq = std::min(qc, num_query - q0);
    {
      EventTimer t(sg.s, &g_ms_hamming);
      if (hamming_device(refp.as<u64>(), npad, num_refs, qryp.as<u64>() + (size_t)q0 * W * 3, nq, W, D.as<uint16_t>(),
                         (size_t)num_refs, sg.s))
        return 1;
    }
    CK(cudaMemcpyAsync(h.data(), D.p, (size_t)nq * num_refs * sizeof(uint16_t), cudaMemcpyDeviceToHost, sg.s));
    CK(cudaStreamSynchronize(sg.s));
    int *out = hamming + (size_t)q0 * num_refs;
    for (size_t i = 0; i < (size_t)nq * num_refs; i++) out[i] = 2 * (int)h[i];   // closest_seqs.c:45
  }
  return 0;
}

// Selection groups of the given queries as one flat block (see kGroupsMagic), malloc'ed: free with mcq_free.
extern "C" int mcq_closest_groups(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                  const char *query_seqs, size_t query_stride, int num_query, int num_bases,
                                  int closest_n, int skip_first, int **groups, size_t *count) {
  REQUIRE(groups && count, "null argument");
  REQUIRE(closest_n > 0, "closest_n must be positive");
  std::vector<int> flat;
  if (closest_groups(device, ref_seqs, ref_stride, num_refs, query_seqs, query_stride, num_query, num_bases,
                     skip_first ? closest_n + 1 : closest_n, flat))
    return 1;
  *groups = (int *)malloc(flat.size() * sizeof(int));
  REQUIRE(*groups, "out of host memory");
  memcpy(*groups, flat.data(), flat.size() * sizeof(int));
  *count = flat.size();
  return 0;
}

extern "C" void mcq_free(void *p) { free(p); }

extern "C" int mcq_closest_replay(const int *groups, size_t count, int closest_n, int skip_first, int *closest_refs,
                                  int *num_query) {
  REQUIRE(groups && closest_refs, "null argument");
  size_t used = 0;
  if (replay_block(groups, count, closest_n, skip_first, closest_refs, &used)) return 1;
  REQUIRE(used == count, "trailing data after the block of selection groups");
  if (num_query) *num_query = groups[1];
  return 0;
}

extern "C" int mcq_closest_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                 const char *query_seqs, size_t query_stride, int num_query,
                                 int num_bases, int closest_n, int skip_first, int *closest_refs) {
  REQUIRE(closest_refs, "null argument");
  REQUIRE(closest_n > 0, "closest_n must be positive");
  std::vector<int> flat;
  if (closest_groups(device, ref_seqs, ref_stride, num_refs, query_seqs, query_stride, num_query, num_bases,
                     skip_first ? closest_n + 1 : closest_n, flat))
    return 1;
  return replay_block(flat.data(), flat.size(), closest_n, skip_first, closest_refs, nullptr);
}

// ------------------------------------------------------------------------------------------
// per-query consensus of the closest references (generate_consensus_subset, amino_acids.c:118-141,
// as used by dmodel.c:664-681): majority base per position over the query's L references, ties to
// the first of T,C,A,G, 'T' when no reference has a base there.
//
// Bit-sliced: the panel is packed as for the distances (three bit planes per 64 positions, sequence-major, so
// the 3*W words of a reference are one contiguous run).  A lane owns one 64-position word of one query and walks
// the query's references: per reference it forms the four indicator words "this position shows T / C / A / G"
// and adds each into a vertical counter (kBits bit-slice words per base: 64 counters advance with ~3 logic
// operations per slice), then compares the four counters slice by slice.  About one logic operation per
// (reference, position) instead of a byte load, a decode and four compares.
// ------------------------------------------------------------------------------------------
template <int kBits>
struct SlicedCounter {
  u64 b[kBits];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < kBits; i++) b[i] = 0;
  }
  __device__ __forceinline__ void add(u64 carry) {      // +1 where the bit of `carry` is set
#pragma unroll
    for (int i = 0; i < kBits; i++) {
      const u64 t = b[i] & carry;
      b[i] ^= carry;
      carry = t;
    }
  }
};

// mask of the positions where a > b
template <int kBits>
__device__ __forceinline__ u64 sliced_gt(const SlicedCounter<kBits> &a, const SlicedCounter<kBits> &b) {
  u64 gt = 0, eq = ~0ull;
#pragma unroll
  for (int i = kBits - 1; i >= 0; i--) {
    gt |= eq & a.b[i] & ~b.b[i];
    eq &= ~(a.b[i] ^ b.b[i]);
  }
  return gt;
}

template <int kBits>
__device__ __forceinline__ void sliced_take(SlicedCounter<kBits> &dst, const SlicedCounter<kBits> &src, u64 where) {
#pragma unroll
  for (int i = 0; i < kBits; i++) dst.b[i] = (dst.b[i] & ~where) | (src.b[i] & where);
}

// planes: [N][W][3] (k_pack, seq_major).  One warp per (query, group of 32 words); lane = word.
template <int kBits>
__global__ void __launch_bounds__(128) k_consensus(const u64 *__restrict__ planes, int W, const int *__restrict__ closest,
                                                   int L, int nbases, int nq, signed char *__restrict__ out) {
  const int wgroups = (W + 31) / 32;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int q = (int)(warp / wgroups), w = (int)(warp % wgroups) * 32 + lane;
  if (q >= nq) return;
  const int *cl = closest + (size_t)q * L;
  SlicedCounter<kBits> cT, cC, cA, cG;
  cT.clear(); cC.clear(); cA.clear(); cG.clear();
  const bool on = w < W;
  // the reference indices of the list, 32 at a time through the lanes; the planes of the next reference are
  // requested before the current one is counted
  for (int r0 = 0; r0 < L; r0 += 32) {
    const int mine = (r0 + lane < L) ? cl[r0 + lane] : 0;
    const int n = min(32, L - r0);
    u64 lo = 0, hi = 0, un = ~0ull;
    {
      const int ref = __shfl_sync(0xffffffffu, mine, 0);
      if (on) { const u64 *p = planes + ((size_t)ref * W + w) * 3; lo = p[0]; hi = p[1]; un = p[2]; }
    }
    for (int k = 0; k < n; k++) {
      u64 nlo = 0, nhi = 0, nun = ~0ull;
      if (k + 1 < n) {
        const int ref = __shfl_sync(0xffffffffu, mine, k + 1);
        if (on) { const u64 *p = planes + ((size_t)ref * W + w) * 3; nlo = p[0]; nhi = p[1]; nun = p[2]; }
      }
      const u64 base = ~un;                       // k_pack clears the code bits of unknown positions
      cT.add(base & ~lo & ~hi);
      cC.add(lo & ~hi);
      cA.add(hi & ~lo);
      cG.add(lo & hi);
      lo = nlo; hi = nhi; un = nun;
    }
  }
  if (!on) return;
  // strict '>' scan in T, C, A, G order (amino_acids.c:132-136)
  u64 code_lo = 0, code_hi = 0;
  SlicedCounter<kBits> best = cT;
  u64 g = sliced_gt(cC, best);
  code_lo = g;
  sliced_take(best, cC, g);
  g = sliced_gt(cA, best);
  code_lo &= ~g; code_hi = g;
  sliced_take(best, cA, g);
  g = sliced_gt(cG, best);
  code_lo |= g; code_hi |= g;
  signed char *o = out + (size_t)q * nbases + (size_t)w * 64;
  const int nb = min(64, nbases - w * 64);
  for (int i = 0; i < nb; i += 4) {              // four codes per 32-bit store (whole words inside the row)
    unsigned v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int bit = i + k;
      v |= (unsigned)(((code_lo >> bit) & 1) | (((code_hi >> bit) & 1) << 1)) << (8 * k);
    }
    if (i + 4 <= nb && (((size_t)q * nbases + (size_t)w * 64 + i) & 3) == 0) *reinterpret_cast<unsigned *>(o + i) = v;
    else
      for (int k = 0; k < 4 && i + k < nb; k++) o[i + k] = (signed char)((v >> (8 * k)) & 3);
  }
}

extern "C" int mcq_consensus_batch(int device, const char *ref_seqs, size_t ref_stride, int num_refs,
                                   const int *closest_refs, int num_query, int closest_n, int num_bases,
                                   char *cons_codes) {
  REQUIRE(ref_seqs && closest_refs && cons_codes, "null argument");
  REQUIRE(num_refs > 0 && num_query > 0 && closest_n > 0 && num_bases > 0, "dimensions must be positive");
  REQUIRE(closest_n < (1 << 16), "closest_n too large");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(ndev > 0, "no CUDA device: libmcq_b200 has no CPU fallback");
  REQUIRE(device >= 0 && device < ndev, "device ordinal out of range");
  for (size_t i = 0; i < (size_t)num_query * closest_n; i++)
    REQUIRE(closest_refs[i] >= 0 && closest_refs[i] < num_refs, "closest_refs entry out of range");
  CK(cudaSetDevice(device));
  StreamGuard sg;
  CK(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
  const int W = (num_bases + 63) / 64;
  DevBuf planes, d_cl, d_out;
  if (pack_to_device(ref_seqs, ref_stride, num_refs, num_bases, W, true, 0, planes, sg.s)) return 1;
  CK(cudaMalloc(&d_cl.p, (size_t)num_query * closest_n * sizeof(int)));
  CK(cudaMalloc(&d_out.p, (size_t)num_query * num_bases));
  CK(cudaMemcpyAsync(d_cl.p, closest_refs, (size_t)num_query * closest_n * sizeof(int), cudaMemcpyHostToDevice, sg.s));
  const long long warps = (long long)num_query * ((W + 31) / 32);
  const unsigned blocks = (unsigned)((warps * 32 + 127) / 128);
  g_ms_consensus = 0.f;
  {
    EventTimer t(sg.s, &g_ms_consensus);
    // counters of 8, 11 or 16 bits: the smallest that holds closest_n
    if (closest_n < 256)
      k_consensus<8><<<blocks, 128, 0, sg.s>>>(planes.as<u64>(), W, d_cl.as<int>(), closest_n, num_bases, num_query, d_out.as<signed char>());
    else if (closest_n < 2048)
      k_consensus<11><<<blocks, 128, 0, sg.s>>>(planes.as<u64>(), W, d_cl.as<int>(), closest_n, num_bases, num_query, d_out.as<signed char>());
    else
      k_consensus<16><<<blocks, 128, 0, sg.s>>>(planes.as<u64>(), W, d_cl.as<int>(), closest_n, num_bases, num_query, d_out.as<signed char>());
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(cons_codes, d_out.p, (size_t)num_query * num_bases, cudaMemcpyDeviceToHost, sg.s));
  CK(cudaStreamSynchronize(sg.s));
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
  if (hamming_device(g_refs.planes.as<u64>(), g_refs.npad, num_refs, qryp.as<u64>(), 1, W, D.as<uint16_t>(), (size_t)num_refs,
                     g_refs.st))
    die_msg("k_hamming");
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
  std::vector<int> flat{kGroupsMagic, 1, closest_n};
  if (select_rows_radix<int>(D.as<int>(), 1, num_refs, closest_n, sg.s, flat) ||
      replay_block(flat.data(), flat.size(), closest_n, 0, closest_refs, nullptr))
    die_msg("selection");
}
