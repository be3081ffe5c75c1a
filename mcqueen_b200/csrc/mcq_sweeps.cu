// mcq_sweeps.cu - the forward / backward recursions with SEVERAL WARPS PER QUERY (sm_100a).
//
// The one-warp-per-query sweeps of mcq_kernels.cu run at the HBM rate when a GPU holds many queries.  With few
// queries per GPU - a shard of an 8-GPU run, the carriers of one HLA type, the config-2 shape - the time of a
// sweep is the length of one warp's dependent instruction stream per site (~170 warp instructions at L = 500)
// times the number of sites.  Here a block of W + 1 warps owns one query: W consumer warps hold NJ / W of the
// NJ 64-state groups each, a producer warp stages the rows of the sites ahead into the shared-memory ring and
// does nothing else.  Per site a consumer does its share of the cell updates, one butterfly, and leaves its
// partial column sum in shared memory; ONE block barrier per site publishes the partial sums (added in warp
// order: the same bits in every warp, so every warp takes the same rescale decision, forward_backward.c:163-167)
// and hands the freed ring slot back to the producer.  The staging instructions, a quarter of the one-warp
// stream, leave the critical path altogether.
//
// Arithmetic per cell is that of k_forward / k_backward; the column sum is grouped by warp, so totals differ
// from the one-warp kernels in the last bits (a context uses one grouping for all its launches).
#include "mcq_internal.cuh"
#include "mcq_sweep_common.cuh"
#include "../../include/mcq_b200.h"

// Ring of staged sites.  A consumer warp is through with a site in ~0.2 us, a row needs 0.6 - 0.9 us from HBM to
// shared memory: seven sites are in flight beyond the one being computed (the one-warp kernels, slower per site,
// get by with three).
#define MCQ_RINGW 8
static_assert((MCQ_RINGW & (MCQ_RINGW - 1)) == 0 && MCQ_RINGW >= 4, "ring slots are addressed with a mask");

// Stage the rows of one site into ring slot `slot` (run time: the producer is off the critical path).  Same
// contract as stage_k.
template <int NJ>
__device__ __forceinline__ void stage_rt(unsigned g_sh, unsigned e_sh, int slot, bool valid, unsigned rd,
                                         const double *__restrict__ tabl, const double *__restrict__ ecp, bool has_ec,
                                         const double *__restrict__ Rp, const uint8_t *__restrict__ gsrc, int lane) {
#pragma unroll
  for (int i = 0; i < (NJ + 7) / 8; i++)
    cp_async16_if(g_sh + slot * (64 * NJ) + 512 * i, gsrc + 512 * i, valid && (lane + 32 * i) * 16 < 64 * NJ);
  const int kr = (int)(rd >> MCQ_RD_SHIFT);
  const double *row = tabl + (rd & MCQ_RD_MASK);
  const unsigned ed = e_sh + (slot * MCQ_ES + 1) * 8;
  cp_async8_if(ed, row, lane < kr);
  cp_async8_if(ed + 256, row + 32, lane + 32 < kr);
  const bool l0 = lane == 0;
  cp_async8_if(l0 ? ed - 8 : ed + (MCQ_RSLOT - 2) * 8, l0 ? ecp : Rp, valid && (l0 ? has_ec : lane == 1));
  cp_async_commit();
}

// The descriptor words of the sites the producer stages: lane i holds the word of site base + i (one coalesced
// load per 32 sites, the next 32 already in flight), a shuffle hands out the word of site t.  Sites are asked
// for in ascending (forward) or descending (backward) order.
template <bool DESC>
struct RowDescWindow {
  const uint32_t *rdq;
  int lo, hi, base, lane;      // valid sites lo..hi; base = first (forward) / last (backward) site of the window
  unsigned cur, nxt;
  __device__ __forceinline__ unsigned fetch(int b) const {
    const int t = DESC ? b - lane : b + lane;
    return (t >= lo && t <= hi) ? rdq[t] : 0u;
  }
  __device__ __forceinline__ void init(const uint32_t *r, int lo_, int hi_, int first, int lane_) {
    rdq = r; lo = lo_; hi = hi_; base = first; lane = lane_;
    cur = fetch(base);
    nxt = fetch(DESC ? base - 32 : base + 32);
  }
  __device__ __forceinline__ unsigned get(int t) {
    while (DESC ? t <= base - 32 : t >= base + 32) {
      cur = nxt;
      base = DESC ? base - 32 : base + 32;
      nxt = fetch(DESC ? base - 32 : base + 32);
    }
    return __shfl_sync(0xffffffffu, cur, DESC ? base - t : t - base);
  }
};

// the 2*NJW slot bytes a lane of one consumer warp owns, from ring slot K (gl = its bytes in slot 0)
template <int NJW, int NJ, int K>
__device__ __forceinline__ void load_slots_mw(const uint8_t *gl, unsigned (&w)[(2 * NJW + 3) / 4]) {
  const uint8_t *p = gl + K * (64 * NJ);
  if constexpr (NJW == 1) w[0] = *reinterpret_cast<const uint16_t *>(p);
  else if constexpr (NJW == 2) w[0] = *reinterpret_cast<const uint32_t *>(p);
  else if constexpr (NJW == 4) { const uint2 v = *reinterpret_cast<const uint2 *>(p); w[0] = v.x; w[1] = v.y; }
  else {
#pragma unroll
    for (int i = 0; i < NJW / 8; i++) {
      const uint4 v = reinterpret_cast<const uint4 *>(p)[i];
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  }
}

// sum of the W partial values the consumer warps left in shared memory, in warp order
template <int W>
__device__ __forceinline__ double sum_parts(const double *p) {
  double t = p[0];
#pragma unroll
  for (int w = 1; w < W; w++) t += p[w];
  return t;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int NJW>
struct FwdWarpMw {
  double p0[NJW], p1[NJW];
  double tot;
  int cnt;
  unsigned sel[MCQ_RINGW];
  double *Fs;
  int *Fnp;
  const uint8_t *selp;
};

template <int NJ, int W, bool FULL, bool PRE, bool STORE, int K>
__device__ __forceinline__ void fwd_site_mw(const FwdArgs &a, FwdWarpMw<NJ / W> &w, int s, const uint8_t *gl,
                                            const double *ering, double (*part)[W], int lane, int wid,
                                            const bool (&inrow)[NJ / W], double invN, ptrdiff_t delta, bool has_sel,
                                            int *Fnc) {
  constexpr int NJW = NJ / W;
  unsigned g[(2 * NJW + 3) / 4];
  load_slots_mw<NJW, NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  const double Rs = (s == a.r_site) ? a.r_value : es[MCQ_RSLOT];
  const double jump = w.tot * (Rs * invN);     // (the arithmetic of fwd_site, mcq_kernels.cu)
  const double stay = 1 - Rs;
  double la = 0.0, lb = 0.0;
#pragma unroll
  for (int j = 0; j < NJW; j++) {
    const double v0 = __fma_rn(stay, w.p0[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2);
    const double v1 = __fma_rn(stay, w.p1[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    w.p0[j] = v0; w.p1[j] = v1;
    la += v0; lb += v1;
  }
  const double mine = warp_sum(la + lb);
  if (lane == 0) part[K & 1][wid] = mine;
  __syncthreads();               // partial sums visible; the rows of site s are free for the producer
  w.tot = sum_parts<W>(part[K & 1]);
  if (w.tot < MCQ_BIGI) {                       // forward_backward.c:163-167
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJW; j++) { w.p0[j] *= MCQ_BIG; w.p1[j] *= MCQ_BIG; }
    w.tot *= MCQ_BIG;
  }
  if (STORE) {
    const bool pre = PRE && s < a.start;
    const unsigned cur1 = w.sel[K];
    if (has_sel && s + MCQ_RINGW <= a.end) w.sel[K] = w.selp[K];
    const bool into1 = (cur1 != 0) != (!a.in_place && !pre);
    const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;
    double *__restrict__ Fs = (into1 ? w.Fs + delta : w.Fs) + K * row;
#pragma unroll
    for (int j = 0; j < NJW; j++)
      if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Fs + 64 * j), make_double2(w.p0[j], w.p1[j]));
    if (lane == 0 && wid == 0) {
      if (pre) Fnc[s] = w.cnt;
      else w.Fnp[K] = w.cnt;
    }
    if (K == MCQ_RINGW - 1) { w.Fs += MCQ_RINGW * row; w.Fnp += MCQ_RINGW; w.selp += MCQ_RINGW; }
  }
}

template <int NJ, int W, bool FULL, bool PRE, bool STORE>
__global__ void __launch_bounds__(32 * (W + 1)) k_forward_mw(const FwdArgs a) {
  constexpr int NJW = NJ / W;
  static_assert(NJ % W == 0 && NJW >= 1, "the 64-state groups are split evenly over the consumer warps");
  __shared__ __align__(16) uint8_t s_g[MCQ_RINGW][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_RINGW][MCQ_ES];
  __shared__ __align__(16) double s_part[2][W];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) {
    if (threadIdx.x == 0 && a.llk_out != nullptr) a.llk_out[q] = a.llk_keep[q];
    return;
  }
  const int S = a.st.S, Ls = a.st.Ls;
  const size_t qoff = (size_t)q * S * Ls;
  int s0 = a.start;
  if (PRE) {
    const int v = a.fv[q];
    if (v < a.start - 1) s0 = v + 1;
  }
  const bool caught_up = PRE && s0 < a.start;
  const int sA = (s0 == 0) ? 1 : s0;
  const int nsites = a.end >= sA ? a.end - sA + 1 : 0;
  uint8_t *gring = &s_g[0][0];
  double *ering = &s_e[0][0];
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const double *ecq_pre = (PRE && a.em_pre.ec) ? a.em_pre.ec + (size_t)q * S : nullptr;

  if (wid == W) {
    // ---- producer: rows of the sites ahead, three in flight beyond the one being computed
    if (lane < MCQ_RINGW) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;   // the "no such codon" emission
    const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
    const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
    const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
    const double *tabl = a.em.tab + lane, *tabl_pre = PRE ? a.em_pre.tab + lane : nullptr;
    RowDescWindow<false> rdw;
    rdw.init(a.st.rowdesc + (size_t)q * S, s0 == 0 ? 0 : sA, a.end, s0 == 0 ? 0 : sA, lane);
    auto stage_site = [&](int t, int slot, bool on) {
      const bool valid = on && t <= a.end;
      const bool pre = PRE && t < a.start;
      const double *eq = pre ? ecq_pre : ecq;
      const unsigned rd = rdw.get(t <= a.end ? t : a.end);
      stage_rt<NJ>(g_sh, e_sh, slot, valid, valid ? rd : 0u, pre ? tabl_pre : tabl, eq + t, eq != nullptr, a.R + t,
                   Gl + (size_t)t * (64 * NJ), lane);
    };
    constexpr int RW = MCQ_RINGW;
    stage_site(0, RW - 1, s0 == 0);           // the seed column's rows (an empty group otherwise)
#pragma unroll
    for (int k = 0; k < RW - 1; k++) stage_site(sA + k, k, true);
    cp_async_wait<RW - 2>();                  // site 0 and site sA have landed
    __syncthreads();                          // (1) consumers start
    __syncthreads();                          // (2) the first column's sum: the seed's slot is free
    stage_site(sA + RW - 1, RW - 1, true);
    for (int s = sA; s <= a.end; s++) {
      cp_async_wait<RW - 2>();                // rows of site s+1
      __syncthreads();                        // consumers are through with the rows of site s
      stage_site(s + RW, (s - sA) & (RW - 1), true);
    }
    cp_async_wait<0>();
    if (a.llk_mode == 1) __syncthreads();
    return;
  }

  // ---- consumers: warp `wid` owns the 64-state groups wid*NJW .. wid*NJW + NJW - 1
  const int jb = wid * NJW;
  const uint8_t *gl = gring + lane * (2 * NJ) + 2 * jb;
  double *X0 = const_cast<double *>(a.Fin) + qoff + 64 * jb + 2 * lane;
  const ptrdiff_t delta = STORE ? (a.Fout - a.Fin) : 0;
  const bool has_sel = a.sel != nullptr;
  const uint8_t *selq = has_sel ? a.sel + (size_t)q * S : nullptr;
  int *Fno = STORE ? a.Fn_out + (size_t)q * S : nullptr;
  int *Fnc = const_cast<int *>(a.Fn_in) + (size_t)q * S;
  bool inrow[NJW];
#pragma unroll
  for (int j = 0; j < NJW; j++) inrow[j] = FULL || (64 * (jb + j) + 2 * lane < Ls);

  FwdWarpMw<NJW> w;
#pragma unroll
  for (int k = 0; k < MCQ_RINGW; k++) w.sel[k] = (has_sel && sA + k <= a.end) ? selq[sA + k] : 0u;
  w.Fs = X0 + (size_t)sA * Ls;
  w.Fnp = STORE ? Fno + sA : nullptr;
  w.selp = selq + sA + MCQ_RINGW;
  const double invN = 1.0 / (double)a.st.N;
  double loc = 0.0;
  if (s0 == 0) {
    // column 0 is the bare emission and is never rescaled (forward_backward.c:27-39,102-111)
    const bool into1 = ((has_sel ? selq[0] : 0) != 0) != (!a.in_place && !(PRE && 0 < a.start));
    double *F0 = into1 ? X0 + delta : X0;
    __syncthreads();                            // (1) rows of site 0 (the last slot) and site sA (slot 0)
    unsigned g[(2 * NJW + 3) / 4];
    load_slots_mw<NJW, NJ, MCQ_RINGW - 1>(gl, g);
    const double *es = ering + (MCQ_RINGW - 1) * MCQ_ES;
#pragma unroll
    for (int j = 0; j < NJW; j++) {
      w.p0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.p1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
      if (STORE && inrow[j]) __stcs(reinterpret_cast<double2 *>(F0 + 64 * j), make_double2(w.p0[j], w.p1[j]));
      loc += w.p0[j]; loc += w.p1[j];
    }
    w.cnt = 0;
    if (STORE && threadIdx.x == 0) {
      Fno[0] = 0;
      if (a.Fn_in != nullptr) Fnc[0] = 0;
    }
  } else {
    const double *__restrict__ Fi = X0 + (size_t)(s0 - 1) * Ls + ((has_sel && selq[s0 - 1]) ? (a.Fout - a.Fin) : 0);
#pragma unroll
    for (int j = 0; j < NJW; j++) {
      w.p0[j] = 0.0; w.p1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Fi + 64 * j);
        w.p0[j] = v.x; w.p1[j] = v.y;
      }
      loc += w.p0[j]; loc += w.p1[j];
    }
    w.cnt = Fnc[s0 - 1];
    __syncthreads();                            // (1)
  }
  {
    const double mine = warp_sum(loc);
    if (lane == 0) s_part[1][wid] = mine;
    __syncthreads();                            // (2)
    w.tot = sum_parts<W>(s_part[1]);
  }

#define MCQ_FWD_SITE(K, SITE) \
  fwd_site_mw<NJ, W, FULL, PRE, STORE, K>(a, w, SITE, gl, ering, s_part, lane, wid, inrow, invN, delta, has_sel, Fnc)
  for (int s = sA; s <= a.end; s += 8) {
    MCQ_FWD_SITE(0, s);
    if (s + 1 > a.end) break;
    MCQ_FWD_SITE(1, s + 1);
    if (s + 2 > a.end) break;
    MCQ_FWD_SITE(2, s + 2);
    if (s + 3 > a.end) break;
    MCQ_FWD_SITE(3, s + 3);
    if (s + 4 > a.end) break;
    MCQ_FWD_SITE(4, s + 4);
    if (s + 5 > a.end) break;
    MCQ_FWD_SITE(5, s + 5);
    if (s + 6 > a.end) break;
    MCQ_FWD_SITE(6, s + 6);
    if (s + 7 > a.end) break;
    MCQ_FWD_SITE(7, s + 7);
  }
#undef MCQ_FWD_SITE
  if (threadIdx.x == 0) {
    if (caught_up) a.fv[q] = a.start - 1;
    if (a.mark_fv != nullptr) a.mark_fv[q] = a.end;
  }
  if (a.llk_mode != 0) {
    double dot;
    int nrm = w.cnt;
    if (a.llk_mode == 1) {                      // mcmc_moves.c:220-223
      const double *__restrict__ Bc = a.B + qoff + (size_t)a.end * Ls + 64 * jb + 2 * lane;
      double l2 = 0.0;
#pragma unroll
      for (int j = 0; j < NJW; j++)
        if (inrow[j]) {
          const double2 b = *reinterpret_cast<const double2 *>(Bc + 64 * j);
          l2 += w.p0[j] * b.x; l2 += w.p1[j] * b.y;
        }
      const double mine = warp_sum(l2);
      const int par = nsites & 1;               // the buffer the last barrier but one protected
      if (lane == 0) s_part[par][wid] = mine;
      __syncthreads();
      dot = sum_parts<W>(s_part[par]);
      nrm += a.Bn[(size_t)q * S + a.end];
    } else {                                    // dmodel.c:899-901, mcmc_moves.c:1129-1134
      dot = w.tot;
    }
    if (threadIdx.x == 0) a.llk_out[q] = log(dot) - nrm * a.logBIG;
  }
}

// ------------------------------------------------------------------------------------------
// backward: b = B[.][s+1] (as stored), e = the emission at s+1, sx = sum_r b*e; the column sum of the new
// column and the next sx travel through the same barrier.  Rows are consumed in the order tA = s+1, s, ...,
// bottom; ring slot of site t = (tA - t) & (MCQ_RINGW - 1).
// ------------------------------------------------------------------------------------------
template <int NJW>
struct BwdWarpMw {
  double b0[NJW], b1[NJW], e0[NJW], e1[NJW];
  double sx, Rnext;
  int cnt;
  double *Bs;
  int *Bnp;
};

template <int NJ, int W, bool FULL, int K>
__device__ __forceinline__ void bwd_site_mw(const BwdArgs &a, BwdWarpMw<NJ / W> &w, const uint8_t *gl,
                                            const double *ering, double (*part)[2][W], int lane, int wid,
                                            const bool (&inrow)[NJ / W], const double (&m0)[NJ / W],
                                            const double (&m1)[NJ / W], double invN) {
  constexpr int NJW = NJ / W;
  unsigned g[(2 * NJW + 3) / 4];
  load_slots_mw<NJW, NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  const double Rn = w.Rnext;
  w.Rnext = es[MCQ_RSLOT];
  const double jump = w.sx * (Rn * invN);       // forward_backward.c:219 (the arithmetic of bwd_site)
  const double stay = 1 - Rn;
  double locv = 0.0, locy = 0.0, locv2 = 0.0, locy2 = 0.0;
#pragma unroll
  for (int j = 0; j < NJW; j++) {
    const double n0 = emis_at(es, g[j >> 1], (j & 1) * 2), n1 = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    const double v0 = __fma_rn(stay * w.b0[j], w.e0[j], jump) * m0[j];
    const double v1 = __fma_rn(stay * w.b1[j], w.e1[j], jump) * m1[j];
    w.b0[j] = v0; w.b1[j] = v1;
    locv += v0; locv2 += v1;
    locy = __fma_rn(v0, n0, locy); locy2 = __fma_rn(v1, n1, locy2);
    w.e0[j] = n0; w.e1[j] = n1;
  }
  locv += locv2; locy += locy2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    locv += __shfl_xor_sync(0xffffffffu, locv, o);
    locy += __shfl_xor_sync(0xffffffffu, locy, o);
  }
  if (lane == 0) { part[K & 1][0][wid] = locv; part[K & 1][1][wid] = locy; }
  __syncthreads();
  const double tot = sum_parts<W>(part[K & 1][0]);
  w.sx = sum_parts<W>(part[K & 1][1]);
  if (tot < MCQ_BIGI) {                         // forward_backward.c:236-242
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJW; j++) { w.b0[j] *= MCQ_BIG; w.b1[j] *= MCQ_BIG; }
    w.sx *= MCQ_BIG;
  }
  // (the sites run K = 1, 2, ..., 7, 0, 1, ...: the output pointers address the site of K = 1)
  constexpr int J = (K + MCQ_RINGW - 1) & (MCQ_RINGW - 1);
  const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;
  double *__restrict__ Bs = w.Bs - J * row;
#pragma unroll
  for (int j = 0; j < NJW; j++)
    if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Bs + 64 * j), make_double2(w.b0[j], w.b1[j]));
  if (lane == 0 && wid == 0) *(w.Bnp - J) = w.cnt;
  if (K == 0) { w.Bs -= MCQ_RINGW * row; w.Bnp -= MCQ_RINGW; }
}

template <int NJ, int W, bool FULL>
__global__ void __launch_bounds__(32 * (W + 1)) k_backward_mw(const BwdArgs a) {
  constexpr int NJW = NJ / W;
  static_assert(NJ % W == 0 && NJW >= 1, "the 64-state groups are split evenly over the consumer warps");
  __shared__ __align__(16) uint8_t s_g[MCQ_RINGW][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_RINGW][MCQ_ES];
  __shared__ __align__(16) double s_part[2][2][W];
  const int q = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) return;
  const int S = a.st.S, L = a.st.L, Ls = a.st.Ls;
  const size_t qoff = (size_t)q * S * Ls;
  int s = a.update;
  const int bottom = a.bottom;
  if (a.bv != nullptr) {
    const int v = a.bv[q];
    if (v <= bottom) return;                    // already valid down to the requested column
    s = v - 1;
  }
  const bool fresh = s == S - 1;                // forward_backward.c:197-203: column S-1 is all ones
  if (fresh) s--;
  // (every thread of the block has taken the same path up to here)
  if (s < bottom) {
    if (fresh && wid < W) {
      double *__restrict__ Bq = a.B + qoff + (size_t)(S - 1) * Ls + 64 * wid * NJW + 2 * lane;
#pragma unroll
      for (int j = 0; j < NJW; j++) {
        const int r = 64 * (wid * NJW + j) + 2 * lane;
        if (FULL || r < Ls)
          __stcs(reinterpret_cast<double2 *>(Bq + 64 * j), make_double2(r < L ? 1.0 : 0.0, r + 1 < L ? 1.0 : 0.0));
      }
      if (threadIdx.x == 0) a.Bn[(size_t)q * S + S - 1] = 0;
    }
    if (a.bv != nullptr && threadIdx.x == 0) a.bv[q] = bottom;
    return;
  }
  uint8_t *gring = &s_g[0][0];
  double *ering = &s_e[0][0];
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const int tA = s + 1;                         // the first site whose rows are consumed

  if (wid == W) {
    // ---- producer
    if (lane < MCQ_RINGW) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;
    const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
    const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
    const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
    const double *tabl = a.em.tab + lane;
    RowDescWindow<true> rdw;
    rdw.init(a.st.rowdesc + (size_t)q * S, bottom, tA, tA, lane);
    auto stage_site = [&](int t, int slot) {
      const bool valid = t >= bottom;
      const int tc = valid ? t : bottom;
      const unsigned rd = rdw.get(tc);
      stage_rt<NJ>(g_sh, e_sh, slot, valid, valid ? rd : 0u, tabl, ecq + tc, ecq != nullptr, a.R + tc,
                   Gl + (ptrdiff_t)tc * (64 * NJ), lane);
    };
    constexpr int RW = MCQ_RINGW;
#pragma unroll
    for (int k = 0; k < RW; k++) stage_site(tA - k, k);
    cp_async_wait<RW - 1>();                    // rows of site tA
    __syncthreads();                            // (1) consumers read the emission at tA
    // the consumers' first barrier (sx) also frees slot 0; from then on one barrier per column
    for (int t = tA; t > bottom; t--) {         // t = the site whose rows the consumers are reading now
      cp_async_wait<RW - 2>();                  // rows of site t-1
      __syncthreads();
      stage_site(t - RW, (tA - t) & (RW - 1));
    }
    // the consumers' last column (site `bottom`) ends with a barrier as well
    cp_async_wait<0>();
    __syncthreads();
    return;
  }

  // ---- consumers
  const int jb = wid * NJW;
  const uint8_t *gl = gring + lane * (2 * NJ) + 2 * jb;
  double *__restrict__ Bq = a.B + qoff + 64 * jb + 2 * lane;
  int *__restrict__ Bnq = a.Bn + (size_t)q * S;
  bool inrow[NJW];
  double m0[NJW], m1[NJW];   // 1 for real states, 0 for padding (r >= L): padding must not pick up `jump`
#pragma unroll
  for (int j = 0; j < NJW; j++) {
    const int r = 64 * (jb + j) + 2 * lane;
    inrow[j] = FULL || (r < Ls);
    m0[j] = (r < L) ? 1.0 : 0.0;
    m1[j] = (r + 1 < L) ? 1.0 : 0.0;
  }
  BwdWarpMw<NJW> w;
  if (fresh) {
#pragma unroll
    for (int j = 0; j < NJW; j++) {
      w.b0[j] = m0[j]; w.b1[j] = m1[j];
      if (inrow[j]) __stcs(reinterpret_cast<double2 *>(Bq + (size_t)(S - 1) * Ls + 64 * j), make_double2(w.b0[j], w.b1[j]));
    }
    w.cnt = 0;
    if (threadIdx.x == 0) Bnq[S - 1] = 0;
  } else {
#pragma unroll
    for (int j = 0; j < NJW; j++) {
      w.b0[j] = 0.0; w.b1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Bq + (size_t)(s + 1) * Ls + 64 * j);
        w.b0[j] = v.x; w.b1[j] = v.y;
      }
    }
    w.cnt = Bnq[s + 1];
  }
  w.Bs = Bq + (size_t)s * Ls;
  w.Bnp = Bnq + s;
  const double invN = 1.0 / (double)a.st.N;
  __syncthreads();                              // (1) rows of site tA (slot 0)
  {
    unsigned g[(2 * NJW + 3) / 4];
    load_slots_mw<NJW, NJ, 0>(gl, g);
    const double *es = ering;
    double loc = 0.0;
#pragma unroll
    for (int j = 0; j < NJW; j++) {
      w.e0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.e1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
      loc += w.b0[j] * w.e0[j]; loc += w.b1[j] * w.e1[j];
    }
    w.Rnext = es[MCQ_RSLOT];                    // R[s+1]
    const double mine = warp_sum(loc);
    if (lane == 0) s_part[0][1][wid] = mine;    // the parity of slot 0
    __syncthreads();
    w.sx = sum_parts<W>(s_part[0][1]);
  }
#define MCQ_BWD_SITE(K) bwd_site_mw<NJ, W, FULL, K>(a, w, gl, ering, s_part, lane, wid, inrow, m0, m1, invN)
  for (; s >= bottom; s -= 8) {
    MCQ_BWD_SITE(1);
    if (s - 1 < bottom) break;
    MCQ_BWD_SITE(2);
    if (s - 2 < bottom) break;
    MCQ_BWD_SITE(3);
    if (s - 3 < bottom) break;
    MCQ_BWD_SITE(4);
    if (s - 4 < bottom) break;
    MCQ_BWD_SITE(5);
    if (s - 5 < bottom) break;
    MCQ_BWD_SITE(6);
    if (s - 6 < bottom) break;
    MCQ_BWD_SITE(7);
    if (s - 7 < bottom) break;
    MCQ_BWD_SITE(0);
  }
#undef MCQ_BWD_SITE
  if (a.bv != nullptr && threadIdx.x == 0) a.bv[q] = bottom;
}

// ------------------------------------------------------------------------------------------
// launchers: a.wpq consumer warps per query
// ------------------------------------------------------------------------------------------
template <int NJ, int W, bool STORE>
static cudaError_t launch_fwd(const FwdArgs &a, cudaStream_t stream) {
  const bool full = a.st.Ls == 64 * NJ;
  void (*kernel)(const FwdArgs);
  if (a.fv != nullptr) kernel = full ? k_forward_mw<NJ, W, true, true, STORE> : k_forward_mw<NJ, W, false, true, STORE>;
  else kernel = full ? k_forward_mw<NJ, W, true, false, STORE> : k_forward_mw<NJ, W, false, false, STORE>;
  kernel<<<(unsigned)a.st.Q, 32 * (W + 1), 0, stream>>>(a);
  return cudaGetLastError();
}

template <int NJ, int W>
static cudaError_t launch_fwd_store(const FwdArgs &a, cudaStream_t stream) {
  if (a.no_store) {
    if (a.fv != nullptr || a.start != 0) return cudaErrorInvalidValue;
    return launch_fwd<NJ, W, false>(a, stream);
  }
  return launch_fwd<NJ, W, true>(a, stream);
}

template <int NJ, int W>
static cudaError_t launch_bwd(const BwdArgs &a, cudaStream_t stream) {
  void (*kernel)(const BwdArgs) = (a.st.Ls == 64 * NJ) ? k_backward_mw<NJ, W, true> : k_backward_mw<NJ, W, false>;
  kernel<<<(unsigned)a.st.Q, 32 * (W + 1), 0, stream>>>(a);
  return cudaGetLastError();
}

// the (NJ, W) pairs built: W = NJ/2 and W = NJ (two or one 64-state group per consumer warp), at most 8
#define MCQ_MW_CASES(X)            \
  X(2, 1) X(2, 2)                  \
  X(4, 1) X(4, 2) X(4, 4)          \
  X(8, 1) X(8, 2) X(8, 4) X(8, 8)  \
  X(16, 2) X(16, 4) X(16, 8)

bool mcq_mw_supported(int NJ, int W) {
#define X(nj, w) if (NJ == nj && W == w) return true;
  MCQ_MW_CASES(X)
#undef X
  return false;
}

cudaError_t mcq_launch_forward_mw(const FwdArgs &a, int W, cudaStream_t stream) {
  const int NJ = a.st.Lg / 64;
#define X(nj, w) if (NJ == nj && W == w) return launch_fwd_store<nj, w>(a, stream);
  MCQ_MW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t mcq_launch_backward_mw(const BwdArgs &a, int W, cudaStream_t stream) {
  const int NJ = a.st.Lg / 64;
#define X(nj, w) if (NJ == nj && W == w) return launch_bwd<nj, w>(a, stream);
  MCQ_MW_CASES(X)
#undef X
  return cudaErrorInvalidValue;
}
