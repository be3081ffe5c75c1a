// mcq_kernels.cu - the CUDA kernels of the McQueen likelihood path (sm_100a).
//
//   K0  k_gather        triplet_to_codon (amino_acids.c:50-66) evaluated once into G[q][s][r]
//   K3a k_emission_table  create_codon_to_codon_HLA, reversion part (mcmc_moves.c:78-124), shared rows
//   K3c k_emission_cons   create_codon_to_codon_HLA, escape part (mcmc_moves.c:125-167), per query
//   K3b k_mu_rescale_*    the mu-rescale of mutation_move (mcmc_moves.c:1082-1111)
//   K4  k_forward<NJ>   initialise_F / update_F (forward_backward.c:11-171) + K6 llk epilogue
//   K5  k_backward<NJ>  update_B (forward_backward.c:173-244)
//   K7  k_commit_*      accept-commit copies (mcmc_moves.c:263-265,447-449,959-965)
//   K6b k_sum           cross-query total (mcmc_moves.c:222 etc.), fixed-shape reduction
//
// All arithmetic is fp64 and the library is compiled with -fmad=false: the reference is built
// without FMA contraction (its Makefile:22 gives -O3 only), and with contraction off the emission
// rows are bit-identical to the reference's and a cell update differs from it only through the
// order of the per-site sum over copy states (a warp butterfly here, ascending r there).
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include "mcq_internal.cuh"
#include "mcq_sweep_common.cuh"
#include "../../include/mcq_b200.h"

// ------------------------------------------------------------------------------------------
// substitution classes, derived from the genetic code on the host (mcq_tables.cpp) and copied
// here once: 0 = none / multi-step, 1 = NS_TS, 2 = NS_TV, 3 = S_TS, 4 = S_TV
// (constants.c:424-688 stores these as four 64x64 0/1 int matrices).
// ------------------------------------------------------------------------------------------
__device__ uint8_t d_cls[64 * 64];
__device__ uint8_t d_cnt[4][64];     // alpha_S, alpha_V, beta_S, beta_V (constants.c:405-422)

extern "C" void mcq_host_tables(uint8_t cls[64 * 64], uint8_t cnt[4][64]);

cudaError_t mcq_launch_tables_init() {
  static uint8_t cls[64 * 64];
  static uint8_t cnt[4][64];
  mcq_host_tables(cls, cnt);
  cudaError_t e = cudaMemcpyToSymbol(d_cls, cls, sizeof(cls));
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(d_cnt, cnt, sizeof(cnt));
}

// omega*(kappa*NS_TS + NS_TV) + kappa*S_TS + S_TV for one codon pair.  Exactly one class bit is
// set per pair, so the reference's expression (mcmc_moves.c:85-88) reduces, bit for bit, to one
// of {omega*kappa, omega, kappa, 1, 0}.
__device__ __forceinline__ double step_rate(double omega, double kappa, int c1, int c2) {
  switch (d_cls[c1 * 64 + c2]) {
    case 1: return omega * kappa;
    case 2: return omega;
    case 3: return kappa;
    case 4: return 1.0;
    default: return 0.0;
  }
}
// kappa*NS_TS + NS_TV for one pair
__device__ __forceinline__ double ns_rate(double kappa, int c1, int c2) {
  uint8_t c = d_cls[c1 * 64 + c2];
  return c == 1 ? kappa : (c == 2 ? 1.0 : 0.0);
}
__device__ __forceinline__ double s_rate(double kappa, int c1, int c2) {
  uint8_t c = d_cls[c1 * 64 + c2];
  return c == 3 ? kappa : (c == 4 ? 1.0 : 0.0);
}

// ------------------------------------------------------------------------------------------
// K0: gathered codon-slot table.  One block per (query, 32 sites): reads run along a reference's
// row (32 consecutive sites = one sector), the tile is transposed and PERMUTED in shared memory,
// and every site row of G leaves as one contiguous run of 16-byte stores.
//
// Physical order inside a G row (Lg = 64*NJ bytes): the 2*NJ states a lane owns in k_forward /
// k_backward - 64j+2l, 64j+2l+1 for j < NJ - are contiguous at byte offset l*2*NJ, so a lane
// fetches all of them with one vector load from the staged row.
// ------------------------------------------------------------------------------------------
// state index stored at physical byte `pos` of a G row (the inverse of the lane-major permutation)
__device__ __forceinline__ int g_state_of(int pos, int lg2) {   // lg2 = log2(2*NJ), bytes per lane
  return 64 * ((pos & ((1 << lg2) - 1)) >> 1) + 2 * (pos >> lg2) + (pos & 1);
}

#define MCQ_NEEDS_CONS 0xFE   // table entry of a triplet with an unknown base: resolved per query

// triplet -> slot per site for triplets without unknown bases (query independent); 128 entries per
// site, 125..127 = "no state" -> MCQ_NOSLOT.  Stored [site][128]: k_gather keeps the 32 sites of its
// block in shared memory; the lanes of a warp look up the same site with different triplets, which land
// in different banks (125 triplets = 32 words) or broadcast.
__global__ void __launch_bounds__(128) k_trip_table(const uint8_t *__restrict__ slot_of, uint8_t *__restrict__ trip_tab,
                                                    int S) {
  const int s = blockIdx.x, t = threadIdx.x;
  uint8_t v = MCQ_NOSLOT;
  if (s < S && t < 125) {
    const int d0 = t / 25, d1 = (t / 5) % 5, d2 = t % 5;
    v = (d0 == 4 || d1 == 4 || d2 == 4) ? MCQ_NEEDS_CONS : slot_of[(size_t)s * 64 + d0 * 16 + d1 * 4 + d2];
  }
  trip_tab[(size_t)s * 128 + t] = v;
}

// ref_triplets[N][S] -> the SLOT of every reference codon, rows of pitch Sp (a multiple of 32, padded with "no
// slot"): the query-independent half of triplet_to_codon is done once per panel cell here instead of once per
// (query, copy state, site) in the gather, and every run of 32 sites of a reference is one aligned 32-byte sector.
// Triplets with an unknown base keep MCQ_NEEDS_CONS: they are resolved per query.
__global__ void __launch_bounds__(256) k_slot_rows(const int8_t *__restrict__ src, const uint8_t *__restrict__ trip_tab,
                                                   uint32_t *__restrict__ dst, int N, int S, int Sp) {
  const size_t wpr = (size_t)(Sp >> 2), n = (size_t)N * wpr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / wpr;
    const int c0 = (int)(i - row * wpr) * 4;
    const int8_t *p = src + row * S + c0;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const unsigned v = c0 + b < S ? trip_tab[(size_t)(c0 + b) * 128 + (uint8_t)p[b]] : (unsigned)MCQ_NOSLOT;
      w |= v << (8 * b);
    }
    dst[i] = w;
  }
}

// One block per query (and share of its 64-site tiles).  The slot rows of the query's references are brought
// into shared memory first, 64 bytes of each per tile (cp.async, 16 bytes per lane: a warp instruction covers
// 64-byte runs of 8 references), the next tile travelling while the current one is turned.  Then a lane owns
// SIXTEEN consecutive physical positions of the G rows (= sixteen copy states) and sixteen sites: a 16 x 16 byte
// transpose in registers, after which the lane stores 16 bytes per site and the LG/16 lanes of a row together one
// whole G row per store instruction (512 contiguous bytes at closest_n = 500).  Shared layout: the row of position
// pos sits at slot (pos % 16) * (LG/16) + pos / 16, its four 16-byte chunks XOR-swizzled with bits 1-2 of pos / 16:
// the 16-byte reads of a quarter warp then fall into eight different bank groups.
#define MCQ_GT 64
template <int LG>                  // LG = padded states per G row (64 ... 1024): every offset below is an immediate
__global__ void __launch_bounds__(LG / 4 < 32 ? 32 : LG / 4) k_gather(const int8_t *__restrict__ slot_rows, int Sp,
                                                const int8_t *__restrict__ trip,
                                                const int *__restrict__ closest,
                                                const int8_t *__restrict__ cons_nucls,
                                                const uint8_t *__restrict__ slot_of,
                                                uint8_t *__restrict__ G, int S, int L, int N,
                                                int *__restrict__ err) {
  constexpr int GT = MCQ_GT, CPR = GT / 16;                         // sites per tile, 16-byte chunks per reference
  constexpr int LPR = LG / 16;                                      // lanes per G row
  constexpr int NT = LG / 4 < 32 ? 32 : LG / 4;                     // threads per block
  constexpr int LG2 = LG == 64 ? 1 : LG == 128 ? 2 : LG == 256 ? 3 : LG == 512 ? 4 : 5;   // log2(2 * NJ)
  extern __shared__ __align__(16) uint8_t g_rows[];                 // two buffers [LG][GT] of slot bytes, then int refs[LG]
  int *s_ref = reinterpret_cast<int *>(g_rows + 2 * (size_t)LG * GT);
  const int q = blockIdx.y, ntiles = (S + GT - 1) / GT;
  // the references of the LG positions; padding states (r >= L) read row 0 and are overwritten below
#pragma unroll
  for (int p = threadIdx.x; p < LG; p += NT) {
    const int r = g_state_of(p, LG2);
    int ref = 0;
    if (r < L) {
      ref = closest[(size_t)q * L + r];
      if ((unsigned)ref >= (unsigned)N) { ref = 0; *err = 2; }   // reported by mcq_ctx_upload_static
    }
    s_ref[p] = ref;
  }
  __syncthreads();
  // chunk c = threadIdx.x + i * NT: NT is a multiple of CPR, so a thread keeps its 16-byte part of the row
  // segment and walks the positions in steps of NT / CPR
  const int part = threadIdx.x % CPR, p0 = threadIdx.x / CPR;
  const unsigned sm0 = (unsigned)__cvta_generic_to_shared(g_rows);
  auto fetch = [&](int tile, int buf) {             // one commit group per call (empty past the last tile)
    if (tile < ntiles) {
      const int8_t *src = slot_rows + tile * GT + part * 16;
      const unsigned sm = sm0 + buf * (LG * GT);
#pragma unroll 8
      for (int i = 0; i < LG * CPR / NT; i++) {
        const int p = p0 + i * (NT / CPR);
        const unsigned slot = (unsigned)(p & 15) * LPR + (unsigned)(p >> 4);
        cp_async16_if(sm + slot * GT + (((unsigned)part ^ ((unsigned)(p >> 5) & 3u)) << 4),
                      src + (size_t)s_ref[p] * (size_t)Sp, true);
      }
    }
    cp_async_commit();
  };
  const int l = threadIdx.x % LPR, sg = threadIdx.x / LPR;   // position group (16 l ...), group of 16 sites in the tile
  const int pos = 16 * l;
  const bool mine = sg < CPR;
  unsigned pad[4] = {0, 0, 0, 0};                   // byte k = 0xFF when position pos + 4 m + k is padding
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (g_state_of(pos + k, LG2) >= L) pad[k >> 2] |= 0xFFu << (8 * (k & 3));
  const bool any_pad = (pad[0] | pad[1] | pad[2] | pad[3]) != 0;
  const unsigned chunk = (((unsigned)sg ^ ((unsigned)(l >> 1) & 3u)) & (CPR - 1)) << 4;
  // the block's tiles, the rows of the next one travelling while this one is turned
  fetch(blockIdx.x, 0);
  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    fetch(tile + gridDim.x, (it + 1) & 1);
    cp_async_wait<1>();
    __syncthreads();
    const int s0 = tile * GT + 16 * sg;               // this lane's sixteen sites
    const int nrow = min(16, S - s0);                 // ... inside the genome (may be <= 0)
    if (mine && nrow > 0) {
      const uint8_t *rows = g_rows + (it & 1) * (LG * GT) + l * GT + chunk;
      unsigned w[16][4];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const uint4 a = *reinterpret_cast<const uint4 *>(rows + k * (LPR * GT));
        w[k][0] = a.x; w[k][1] = a.y; w[k][2] = a.z; w[k][3] = a.w;
      }
      uint8_t *dst = G + ((size_t)q * S + s0) * LG + pos;
      unsigned worst = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        // byte (i & 3) of word (i >> 2) of each of the sixteen rows -> four words
        const unsigned sel2 = 0x0040u + (i & 3) * 0x0011u;            // {b of x, b of y, 0, 0}
        unsigned o[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
          const unsigned p01 = __byte_perm(w[4 * m][i >> 2], w[4 * m + 1][i >> 2], sel2);
          const unsigned p23 = __byte_perm(w[4 * m + 2][i >> 2], w[4 * m + 3][i >> 2], sel2);
          o[m] = __byte_perm(p01, p23, 0x5410);
          if (any_pad) o[m] = (o[m] & ~pad[m]) | (pad[m] & (MCQ_NOSLOT * 0x01010101u));
          worst |= o[m];
        }
        if (i < nrow) __stcs(reinterpret_cast<uint4 *>(dst + i * LG), make_uint4(o[0], o[1], o[2], o[3]));   // streaming
      }
      // every real slot value is below 0x80, MCQ_NEEDS_CONS is not: bit 7 of the OR says whether any occurred
      if (worst & 0x80808080u) {
        // triplet_to_codon (amino_acids.c:50-66): unknown bases take the query's local consensus
#pragma unroll 1
        for (int i = 0; i < nrow; i++) {
          const int s = s0 + i;
#pragma unroll 1
          for (int k = 0; k < 16; k++) {
            if (g_state_of(pos + k, LG2) >= L) continue;
            const unsigned tt = (uint8_t)trip[(size_t)s_ref[pos + k] * S + s];
            int d0 = tt / 25, d1 = (tt / 5) % 5, d2 = tt % 5;
            if (tt >= 125 || (d0 != 4 && d1 != 4 && d2 != 4)) continue;
            const int8_t *cn = cons_nucls + (size_t)q * 3 * S + 3 * s;
            if (d0 == 4) d0 = cn[0];
            if (d1 == 4) d1 = cn[1];
            if (d2 == 4) d2 = cn[2];
            dst[(size_t)i * LG + k] = slot_of[(size_t)s * 64 + (d0 * 16 + d1 * 4 + d2)];
          }
        }
      }
    }
    __syncthreads();                                  // the buffer is refilled two tiles on
  }
  cp_async_wait<0>();
}

// '0'/'1' characters (load_data.c:125-157) -> 0/1, in place
__global__ void __launch_bounds__(256) k_hla_bits(uint8_t *hla, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) hla[i] = (hla[i] == '1');
}

// query-independent part: the per-site triplet -> slot table and the re-pitched reference rows (pitch
// mcq_gather_pitch(S): whole 128-byte lines)
cudaError_t mcq_launch_gather_prepare(const int8_t *ref_triplets, int8_t *trip_pad, const uint8_t *slot_of,
                                      uint8_t *trip_tab, int S, int N, cudaStream_t stream) {
  const int sblocks = (S + 31) / 32, Sp = (int)mcq_gather_pitch(S);
  k_trip_table<<<sblocks * 32, 128, 0, stream>>>(slot_of, trip_tab, S);
  const size_t words = (size_t)N * (Sp / 4);
  unsigned blocks = (unsigned)((words + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_slot_rows<<<blocks, 256, 0, stream>>>(ref_triplets, trip_tab, reinterpret_cast<uint32_t *>(trip_pad), N, S, Sp);
  return cudaGetLastError();
}

template <int LG>
static cudaError_t launch_gather_lg(const int8_t *trip_pad, int Sp, const int8_t *trip, const int *closest,
                                    const int8_t *cons_nucls, const uint8_t *slot_of, uint8_t *G, int Q, int S, int L,
                                    int N, int *err, cudaStream_t stream) {
  constexpr int threads = LG / 4 < 32 ? 32 : LG / 4;
  constexpr size_t smem = 2 * (size_t)LG * MCQ_GT + (size_t)LG * sizeof(int);
  // (per device; the largest case, closest_n = 1024, takes 128 KB + 4 KB)
  cudaError_t e = cudaFuncSetAttribute(k_gather<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int ntiles = (S + MCQ_GT - 1) / MCQ_GT;
  for (int q0 = 0; q0 < Q; q0 += 65535) {   // gridDim.y is limited to 65535
    int nq = Q - q0 < 65535 ? Q - q0 : 65535;
    // about eight tiles per block, fetched one ahead; more blocks per query when there are few queries
    int bx = (ntiles + 7) / 8;
    while (bx < ntiles && (long long)bx * nq < 148 * 6) bx *= 2;
    dim3 grid(bx < ntiles ? bx : ntiles, nq);
    k_gather<LG><<<grid, threads, smem, stream>>>(trip_pad, Sp, trip, closest + (size_t)q0 * L,
                                                  cons_nucls + (size_t)q0 * 3 * S, slot_of,
                                                  G + (size_t)q0 * S * LG, S, L, N, err);
  }
  return cudaGetLastError();
}

// G rows (and the 0/1 HLA flags) of Q queries; the pointers address the first of them
cudaError_t mcq_launch_gather(const int8_t *trip_pad, const int8_t *trip, const int *closest, const int8_t *cons_nucls,
                              const uint8_t *slot_of, uint8_t *G, uint8_t *hla_chars, int Q,
                              int S, int L, int Lg, int N, int H, int *err, cudaStream_t stream) {
  const int Sp = (int)mcq_gather_pitch(S);
  if (hla_chars != nullptr) {
    const size_t n = (size_t)Q * H;
    k_hla_bits<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(hla_chars, n);
  }
  switch (Lg) {
    case 64: return launch_gather_lg<64>(trip_pad, Sp, trip, closest, cons_nucls, slot_of, G, Q, S, L, N, err, stream);
    case 128: return launch_gather_lg<128>(trip_pad, Sp, trip, closest, cons_nucls, slot_of, G, Q, S, L, N, err, stream);
    case 256: return launch_gather_lg<256>(trip_pad, Sp, trip, closest, cons_nucls, slot_of, G, Q, S, L, N, err, stream);
    case 512: return launch_gather_lg<512>(trip_pad, Sp, trip, closest, cons_nucls, slot_of, G, Q, S, L, N, err, stream);
    case 1024: return launch_gather_lg<1024>(trip_pad, Sp, trip, closest, cons_nucls, slot_of, G, Q, S, L, N, err, stream);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------
// parameter views with the proposal override folded in
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double eff_omega(const McqModel &m, const McqOverride &ov, int s) {
  return (ov.kind == MCQ_PROP_SEL && s == ov.start) ? ov.v0 : m.omega[s];
}
__device__ __forceinline__ double eff_esc(const McqModel &m, const McqOverride &ov, int h, int s, int S) {
  if (ov.kind == MCQ_PROP_ESC && h == ov.hla && s >= ov.start && s <= ov.end)
    return s < ov.split ? ov.v0 : ov.v1;
  return m.esc[(size_t)h * S + s];
}
__device__ __forceinline__ double eff_rev(const McqModel &m, const McqOverride &ov, int s) {
  if (ov.kind == MCQ_PROP_REV && s >= ov.start && s <= ov.end) return s < ov.split ? ov.v0 : ov.v1;
  return m.rev[s];
}
// rate_out_of_codons[c][s]; a SEL proposal recomputes it for the codons the panel shows at the
// site (mcmc_moves.c:345-352), in the association order of dmodel.c:842-848
__device__ __forceinline__ double eff_rate_out(const McqModel &m, const McqOverride &ov, int s, int c,
                                               bool in_list) {
  if (ov.kind == MCQ_PROP_SEL && s == ov.start && in_list)
    return m.kappa * d_cnt[2][c] + d_cnt[3][c] + ov.v0 * (m.kappa * d_cnt[0][c] + d_cnt[1][c]);
  return m.rate_out[(size_t)s * 64 + c];
}

// ------------------------------------------------------------------------------------------
// K3a: the reversion part of create_codon_to_codon_HLA (mcmc_moves.c:78-124): rows of the
// non-consensus codons.  Nothing in it depends on the query except its codon `to`, so it is
// evaluated once per (slot, to): one thread each, 64 x (non-consensus slots of the range).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void emission_table_entry(const EmisArgs &a, unsigned idx) {
  if (idx >= (unsigned)(a.a1 - a.a0) * 64u) return;
  const int j = a.a0 + (int)(idx >> 6), to = (int)(idx & 63);
  const int s = a.st.nsite[j], c1 = a.st.ncodon[j];
  const McqModel &m = a.m;
  const int cons = a.st.cons_codon[s];
  const double omega = eff_omega(m, a.ov, s);
  const double kappa = m.kappa, phi = m.phi, mu = m.mu;
  const double N = (double)a.st.N;
  const double back = (1 - phi) * m.prior[to];
  const double rev = eff_rev(m, a.ov, s);
  const double rate = eff_rate_out(m, a.ov, s, c1, true) + (rev - 1) * step_rate(omega, kappa, c1, cons);
  const double A = 1 - N / (N + rate * 4 * mu);
  const double gamma = phi / rate;
  const double mm = step_rate(omega, kappa, c1, to);
  double v;
  if (to == cons) v = A * (gamma * rev * mm + back);
  else {
    v = A * (gamma * mm + back);
    if (c1 == to) v += (1 - A);
  }
  const int as = a.st.aoff[s], kr = a.st.aoff[s + 1] - as;
  a.tab_dst[(size_t)a.st.toff[s] + (size_t)to * kr + (j - as)] = v;
  if (to == 0) a.atab_dst[j] = A;
}

__global__ void __launch_bounds__(256) k_emission_table(const EmisArgs a) {
  emission_table_entry(a, blockIdx.x * blockDim.x + threadIdx.x);
}

cudaError_t mcq_launch_emission_table(const EmisArgs &a, cudaStream_t stream) {
  const unsigned n = (unsigned)(a.a1 - a.a0) * 64u;
  if (n == 0) return cudaSuccess;
  k_emission_table<<<(n + 255) / 256, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K3c: the escape part (mcmc_moves.c:125-167): the consensus row, one thread per (query, site).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void emission_cons_entry(const EmisArgs &a, unsigned idx) {
  const unsigned W = (unsigned)(a.end - a.start + 1);
  if (idx >= (unsigned)a.st.Q * W) return;
  const int q = (int)(idx / W), s = a.start + (int)(idx % W);
  const int S = a.st.S, H = a.st.H;
  const McqModel &m = a.m;
  const uint8_t *__restrict__ hla = a.st.hla + (size_t)q * H;
  const size_t o = (size_t)q * S + s;
  // an ESC update of one HLA leaves non-carriers untouched (mcmc_moves.c:61-62)
  if (a.which_hla >= 0 && hla[a.which_hla] == 0) {
    if (a.ec_dst != a.ec_src) { a.ec_dst[o] = a.ec_src[o]; a.ac_dst[o] = a.ac_src[o]; }
    return;
  }
  const int cons = a.st.cons_codon[s], to = a.st.query_codon[o];
  const double omega = eff_omega(m, a.ov, s);
  const double kappa = m.kappa, phi = m.phi, mu = m.mu;
  const double N = (double)a.st.N;
  const double back = (1 - phi) * m.prior[to];
  double esc_prod = 1.0;
  for (int h = 0; h < H; h++)
    if (hla[h]) esc_prod *= eff_esc(m, a.ov, h, s, S);
  const double nsum = kappa * d_cnt[0][cons] + d_cnt[1][cons];
  const double rate = eff_rate_out(m, a.ov, s, cons, a.st.cons_in_list[s] != 0) + (esc_prod - 1) * omega * nsum;
  const double A = 1 - N / (N + rate * 4 * mu);
  double v;
  if (cons == 15 || cons == 35) {
    const double gamma = phi / nsum;
    v = A * (gamma * ns_rate(kappa, cons, to) + back);
  } else {
    const double gamma = phi / rate;
    v = A * (gamma * (esc_prod * omega * ns_rate(kappa, cons, to) + s_rate(kappa, cons, to)) + back);
  }
  if (to == cons) v = (1 - A) * (1 - back) + back;
  a.ec_dst[o] = v;
  a.ac_dst[o] = A;
}

__global__ void __launch_bounds__(256) k_emission_cons(const EmisArgs a) {
  emission_cons_entry(a, blockIdx.x * blockDim.x + threadIdx.x);
}

// both parts in one launch (DO_HLA_BOTH): the first nb_table blocks take the shared rows, the rest the
// per-query consensus values
__global__ void __launch_bounds__(256) k_emission_both(const EmisArgs a, unsigned nb_table) {
  if (blockIdx.x < nb_table) emission_table_entry(a, blockIdx.x * blockDim.x + threadIdx.x);
  else emission_cons_entry(a, (blockIdx.x - nb_table) * blockDim.x + threadIdx.x);
}

cudaError_t mcq_launch_emission_both(const EmisArgs &a, cudaStream_t stream) {
  const unsigned n1 = (unsigned)(a.a1 - a.a0) * 64u, n2 = (unsigned)a.st.Q * (unsigned)(a.end - a.start + 1);
  const unsigned nb1 = (n1 + 255) / 256, nb2 = (n2 + 255) / 256;
  k_emission_both<<<nb1 + nb2, 256, 0, stream>>>(a, nb1);
  return cudaGetLastError();
}

cudaError_t mcq_launch_emission_cons(const EmisArgs &a, cudaStream_t stream) {
  const unsigned n = (unsigned)a.st.Q * (unsigned)(a.end - a.start + 1);
  k_emission_cons<<<(n + 255) / 256, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K3b: mu rescale (mcmc_moves.c:1082-1111), entry by entry, on the shared table and on the
// per-query consensus values.  The reference walks ref_site_codons only; the consensus value of a
// site whose consensus codon no reference shows is rescaled as well (the reference leaves that tmp
// entry unwritten - uninitialised memory, SURVEY.md §8c hazard 1).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rescale_entry(double a, double e, bool same, double N, double mu, double mu_new,
                                              double phi, double prior_c1, double &an, double &v) {
  double rate = 0.0;
  if (a != 0) rate = (N / (1 - a) - N) / (4 * mu);
  an = 1 - N / (N + rate * 4 * mu_new);
  if (!same) v = (a == 0) ? 0.0 : e * an / a;
  else v = 1 - an + (1 - phi) * prior_c1 * an;
}

__global__ void __launch_bounds__(256) k_mu_rescale_table(const McqStatic st, const McqModel m, double mu_new,
                                                          int nslots, const double *__restrict__ tab,
                                                          const double *__restrict__ atab,
                                                          double *__restrict__ tab_o, double *__restrict__ atab_o) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)nslots * 64u) return;
  const int j = (int)(idx >> 6), to = (int)(idx & 63);
  const int s = st.nsite[j], c1 = st.ncodon[j];
  const int as = st.aoff[s], kr = st.aoff[s + 1] - as;
  const size_t o = (size_t)st.toff[s] + (size_t)to * kr + (j - as);
  double an, v;
  rescale_entry(atab[j], tab[o], to == c1, (double)st.N, m.mu, mu_new, m.phi, m.prior[c1], an, v);
  tab_o[o] = v;
  if (to == 0) atab_o[j] = an;
}

__global__ void __launch_bounds__(256) k_mu_rescale_cons(const McqStatic st, const McqModel m, double mu_new,
                                                         const double *__restrict__ ec, const double *__restrict__ ac,
                                                         double *__restrict__ ec_o, double *__restrict__ ac_o) {
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (unsigned)st.Q * (unsigned)st.S) return;
  const int s = (int)(idx % (unsigned)st.S);
  const int c1 = st.cons_codon[s], to = st.query_codon[idx];
  double an, v;
  rescale_entry(ac[idx], ec[idx], to == c1, (double)st.N, m.mu, mu_new, m.phi, m.prior[c1], an, v);
  ec_o[idx] = v;
  ac_o[idx] = an;
}

cudaError_t mcq_launch_mu_rescale(const McqStatic &st, const McqModel &m, double mu_new, int nslots,
                                  const double *tab, const double *atab, const double *ec, const double *ac,
                                  double *tab_o, double *atab_o, double *ec_o, double *ac_o, cudaStream_t stream) {
  const unsigned n1 = (unsigned)nslots * 64u, n2 = (unsigned)st.Q * (unsigned)st.S;
  if (n1) k_mu_rescale_table<<<(n1 + 255) / 256, 256, 0, stream>>>(st, m, mu_new, nslots, tab, atab, tab_o, atab_o);
  k_mu_rescale_cons<<<(n2 + 255) / 256, 256, 0, stream>>>(st, m, mu_new, ec, ac, ec_o, ac_o);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_rowdesc(const McqStatic st, uint32_t *__restrict__ out, int q0, int nq) {
  const size_t i0 = (size_t)q0 * st.S, n = i0 + (size_t)nq * st.S;
  for (size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(i % (size_t)st.S);
    const int2 d = st.desc[s];     // {offset of the site's block in the table, entries per row}
    const unsigned to = st.query_codon ? st.query_codon[i] : 0u;
    out[i] = ((unsigned)d.y << MCQ_RD_SHIFT) | ((unsigned)d.x + to * (unsigned)d.y);
  }
}

cudaError_t mcq_launch_rowdesc(const McqStatic &st, uint32_t *out, int q0, int nq, cudaStream_t stream) {
  const size_t n = (size_t)nq * st.S;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_rowdesc<<<blocks, 256, 0, stream>>>(st, out, q0, nq);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K4: forward recursion.  One warp per query; lane l holds states 64j+2l, 64j+2l+1 (j < NJ) of
// the running column in registers.  Per site: the site's G row and emission row arrive in shared
// memory through cp.async, staged MCQ_PF sites ahead so that no global-load latency sits on the
// recurrence; one butterfly all-reduce (the column sum, which both decides the rescale and is the
// next site's recombination sum); one 16-byte streaming store of F per lane per j (512 contiguous
// bytes per warp instruction).
//   FULL : Ls == 64*NJ, no row-bound predicates.
//   PRE  : owed columns before `start` are completed first (lazy refills).
//   STORE: false = the columns are not written at all (likelihood only: the recursion then runs at
//          the instruction-issue / FP64 rate instead of the HBM store rate).
// The site loop is unrolled by the ring depth, so that the ring slot of every site is a compile-time
// constant: slot of site t = (t - sA) & 3, sA the first site of the loop.  The emission of a state is
// one byte extract (PRMT), one scaled add (LEA) and one LDS.
// ------------------------------------------------------------------------------------------
static_assert(MCQ_RING == 4, "the unrolled site loops assume a ring of four slots");


template <int NJ>
struct FwdWarp {              // what one warp carries from site to site
  double p0[NJ], p1[NJ];      // the running column
  double tot;
  int cnt, t;                 // rescale counter; next site to stage
  unsigned rd[MCQ_RING];      // descriptor words of sites t .. t+3 (entry = ring slot of the site)
  unsigned sel[MCQ_RING];     // column selectors of sites s .. s+3 (entry = ring slot of the site)
  // running pointers (advanced once per site, no per-site index arithmetic)
  const uint8_t *gsrc;        // G row of site t (this lane's 16 bytes)
  const uint32_t *rdp;        // descriptor word of site t + MCQ_RING
  const uint32_t *rdsafe;     // a descriptor word that may always be read (the last site's)
  const double *Rp;           // R of site t
  double *Fs;                 // column of the site being computed in buffer 0 (this lane's states)
  int *Fnp;                   // its rescale counter (output array)
  const uint8_t *selp;        // column selector of site s + MCQ_RING
  // the pipelined variant only: emissions and recombination probability of the site about to be computed,
  // fetched while the previous site's butterfly was in flight
  double e0[NJ], e1[NJ], Rs;
};

// one site of the recursion, its rows in ring slot K
template <int NJ, bool FULL, bool PRE, bool STORE, int K>
__device__ __forceinline__ void fwd_site(const FwdArgs &a, FwdWarp<NJ> &w, int s, const uint8_t *gl, const double *ering,
                                         unsigned g_sh, unsigned e_sh, const double *tabl,
                                         const double *tabl_pre, const double *ecq, const double *ecq_pre,
                                         int lane, const bool (&inrow)[NJ], double invN, ptrdiff_t delta,
                                         bool has_sel, int *Fnc) {
  cp_async_wait<MCQ_PF - 1>();   // rows of site s have landed (this lane's copies)
  __syncwarp();                  // ... and every other lane's; the slot of site s-1 is free again
  {
    constexpr int KT = (K + MCQ_PF) & 3;       // ring slot of site t = s + MCQ_PF
    // the running pointers address the group's first site: constant offsets inside the group of four,
    // one bump per group
    const int t = w.t + K;
    const unsigned cur = w.rd[KT];
    // consumed four sites from now.  The likelihood-only sweep loads unconditionally (beyond the range: the word of the
    // last site, ignored by stage_k): a select on the loaded VALUE parks the warp on the load's latency right here.
    // The stored sweep keeps the select: measured on one box at 1 250 queries, 0.533 ms against 0.556 ms with the
    // unconditional load (its stores leave the load/store unit little room; 10 000 queries: no difference)
    if constexpr (STORE) w.rd[KT] = (t + MCQ_RING <= a.end) ? w.rdp[K] : 0u;
    else w.rd[KT] = *((t + MCQ_RING <= a.end) ? w.rdp + K : w.rdsafe);
    const bool pre = PRE && t < a.start;
    const double *eq = pre ? ecq_pre : ecq;
    stage_k<NJ, KT, STORE>(g_sh, e_sh, t <= a.end, cur, pre ? tabl_pre : tabl, eq + t, eq != nullptr, w.Rp + K,
                    w.gsrc + K * (64 * NJ), lane);
    if (K == 3) { w.t += 4; w.rdp += 4; w.Rp += 4; w.gsrc += 4 * (64 * NJ); }
  }
  unsigned g[(2 * NJ + 3) / 4];
  load_slots<NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  const double Rs = (s == a.r_site) ? a.r_value : es[MCQ_RSLOT];
  // forward_backward.c:138 divides by N and :141-142 adds then multiplies; here the reciprocal and
  // a fused multiply-add keep the per-site dependency chain short (each one rounding away from
  // the reference's sequence, far inside the tolerance - the order of the sum differs anyway)
  const double jump = w.tot * (Rs * invN);
  const double stay = 1 - Rs;
  double la = 0.0, lb = 0.0;     // two partial sums: a shorter dependent chain than one
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const double v0 = __fma_rn(stay, w.p0[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2);
    const double v1 = __fma_rn(stay, w.p1[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    w.p0[j] = v0; w.p1[j] = v1;
    if (j == 0) { la = v0; lb = v1; } else { la += v0; lb += v1; }   // (0 + v = v: no add for the first)
  }
  w.tot = warp_sum(la + lb);
  if (w.tot < MCQ_BIGI) {                       // forward_backward.c:163-167
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJ; j++) { w.p0[j] *= MCQ_BIG; w.p1[j] *= MCQ_BIG; }
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));   // of the column as stored (mcq_sweep_common.cuh)
  }
  if (STORE) {
    // the column goes into the buffer that holds the current one (a catch-up site, an in-place refill)
    // or into the other one (a proposal): the same offset, `delta` doubles away
    const bool pre = PRE && s < a.start;
    const unsigned cur1 = w.sel[K];
    if (has_sel && s + MCQ_RING <= a.end) w.sel[K] = w.selp[K];   // consumed four sites from now
    const bool into1 = (cur1 != 0) != (!a.in_place && !pre);
    const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;   // doubles per column
    double *__restrict__ Fs = (into1 ? w.Fs + delta : w.Fs) + K * row;
#pragma unroll
    for (int j = 0; j < NJ; j++)
      if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Fs + 64 * j), make_double2(w.p0[j], w.p1[j]));
    if (lane == 0) {
      if (pre) Fnc[s] = w.cnt;
      else w.Fnp[K] = w.cnt;
    }
    if (K == 3) { w.Fs += 4 * row; w.Fnp += 4; w.selp += 4; }
  }
}

// ---- the pipelined variant (PIPE).  With few warps on an SM the time of a sweep is the length of ONE warp's
// dependent chain per site: the branch that closes a site (the rescale decision needs the butterfly's result)
// keeps everything of the next site - wait for its rows, slot bytes, emission look-ups: two shared-memory
// round trips - from issuing until the column sum is known.  Here a site is cut in two: its FRONT (wait for the
// rows, stage the site three ahead, fetch this lane's emissions and R into registers) runs between the cell
// updates of the previous site and that site's butterfly, so the look-ups travel while the shuffles do.  Same
// operations on the same values in the same order per cell: the results are bit-identical to fwd_site's.
template <int NJ, bool PRE, int K>
__device__ __forceinline__ void fwd_front(const FwdArgs &a, FwdWarp<NJ> &w, int s, const uint8_t *gl, const double *ering,
                                          unsigned g_sh, unsigned e_sh, const double *tabl, const double *tabl_pre,
                                          const double *ecq, const double *ecq_pre, int lane) {
  cp_async_wait<MCQ_PF - 1>();   // rows of site s have landed; the values of site s-1 are in registers already
  __syncwarp();
  {
    constexpr int KT = (K + MCQ_PF) & 3;
    const int t = w.t + K;
    const unsigned cur = w.rd[KT];
    w.rd[KT] = *((t + MCQ_RING <= a.end) ? w.rdp + K : w.rdsafe);
    const bool pre = PRE && t < a.start;
    const double *eq = pre ? ecq_pre : ecq;
    stage_k<NJ, KT>(g_sh, e_sh, t <= a.end, cur, pre ? tabl_pre : tabl, eq + t, eq != nullptr, w.Rp + K,
                    w.gsrc + K * (64 * NJ), lane);
    if (K == 3) { w.t += 4; w.rdp += 4; w.Rp += 4; w.gsrc += 4 * (64 * NJ); }
  }
  unsigned g[(2 * NJ + 3) / 4];
  load_slots<NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  w.Rs = (s == a.r_site) ? a.r_value : es[MCQ_RSLOT];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    w.e0[j] = emis_at(es, g[j >> 1], (j & 1) * 2);
    w.e1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
  }
}

// the cell updates of a site from the registers fwd_front filled; returns this lane's share of the column sum
template <int NJ>
__device__ __forceinline__ double fwd_cells(FwdWarp<NJ> &w, double invN) {
  const double jump = w.tot * (w.Rs * invN);
  const double stay = 1 - w.Rs;
  double la = 0.0, lb = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const double v0 = __fma_rn(stay, w.p0[j], jump) * w.e0[j];
    const double v1 = __fma_rn(stay, w.p1[j], jump) * w.e1[j];
    w.p0[j] = v0; w.p1[j] = v1;
    if (j == 0) { la = v0; lb = v1; } else { la += v0; lb += v1; }   // (0 + v = v: no add for the first)
  }
  return la + lb;
}

// butterfly, rescale, store
template <int NJ, bool FULL, bool PRE, bool STORE, int K>
__device__ __forceinline__ void fwd_finish(const FwdArgs &a, FwdWarp<NJ> &w, int s, double mine, int lane,
                                           const bool (&inrow)[NJ], ptrdiff_t delta, bool has_sel, int *Fnc) {
  w.tot = warp_sum(mine);
  if (w.tot < MCQ_BIGI) {                       // forward_backward.c:163-167
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJ; j++) { w.p0[j] *= MCQ_BIG; w.p1[j] *= MCQ_BIG; }
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));   // of the column as stored (mcq_sweep_common.cuh)
  }
  if (STORE) {
    const bool pre = PRE && s < a.start;
    const unsigned cur1 = w.sel[K];
    if (has_sel && s + MCQ_RING <= a.end) w.sel[K] = w.selp[K];
    const bool into1 = (cur1 != 0) != (!a.in_place && !pre);
    const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;
    double *__restrict__ Fs = (into1 ? w.Fs + delta : w.Fs) + K * row;
#pragma unroll
    for (int j = 0; j < NJ; j++)
      if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Fs + 64 * j), make_double2(w.p0[j], w.p1[j]));
    if (lane == 0) {
      if (pre) Fnc[s] = w.cnt;
      else w.Fnp[K] = w.cnt;
    }
    if (K == 3) { w.Fs += 4 * row; w.Fnp += 4; w.selp += 4; }
  }
}

// (register caps: the stored sweep at most 120 = 17 one-warp blocks per SM, which NJ = 8 misses by two registers
// otherwise; the likelihood-only sweep is issue-bound and wants the warps of 96; NJ = 16 and the catch-up
// variant are exempt)
template <int NJ, bool FULL, bool PRE, bool STORE, bool PIPE>
__global__ void __launch_bounds__(32 * MCQ_WPB) __maxnreg__((NJ <= 8 && !PRE && !PIPE) ? (STORE ? 120 : 96) : 255) k_forward(const FwdArgs a) {
  __shared__ __align__(16) uint8_t s_g[MCQ_WPB][MCQ_RING][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_WPB][MCQ_RING][MCQ_ES];
  const int q = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31, wib = MCQ_WPB == 1 ? 0 : (int)(threadIdx.x >> 5);   // (one warp per block: the rings sit at constant addresses)
  if (q >= a.st.Q) return;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) {
    if (lane == 0 && a.llk_out != nullptr) a.llk_out[q] = a.llk_keep[q];
    return;
  }
  const int S = a.st.S, Ls = a.st.Ls;
  const size_t qoff = (size_t)q * S * Ls;
  const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
  const uint32_t *__restrict__ rdq = a.st.rowdesc + (size_t)q * S;
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const double *ecq_pre = (PRE && a.em_pre.ec) ? a.em_pre.ec + (size_t)q * S : nullptr;
  const double *tabl = a.em.tab + lane, *tabl_pre = PRE ? a.em_pre.tab + lane : nullptr;
  // sites before `start` that are not valid yet are caught up first, with the current rows, into Fin
  // (PRE: the variant launched when the host knows of owed columns; the other one carries none of this)
  int s0 = a.start;
  if (PRE) {
    const int v = a.fv[q];
    if (v < a.start - 1) s0 = v + 1;
  }
  const bool caught_up = PRE && s0 < a.start;
  // buffer 0 = Fin, buffer 1 = Fout; sel[q][s] says which one holds the current column s (none: buffer 0)
  double *X0 = const_cast<double *>(a.Fin) + qoff + 2 * lane;
  const ptrdiff_t delta = STORE ? (a.Fout - a.Fin) : 0;
  const bool has_sel = a.sel != nullptr;
  const uint8_t *selq = has_sel ? a.sel + (size_t)q * S : nullptr;
  int *Fno = STORE ? a.Fn_out + (size_t)q * S : nullptr;
  int *Fnc = const_cast<int *>(a.Fn_in) + (size_t)q * S;
  uint8_t *gring = &s_g[wib][0][0];
  double *ering = &s_e[wib][0][0];
  if (lane < MCQ_RING) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;   // the "no such codon" emission
  const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
  const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
  const uint8_t *gl = gring + lane * (2 * NJ);

  bool inrow[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) inrow[j] = FULL || (64 * j + 2 * lane < Ls);

  // prologue.  The loop starts at site sA with ring slot 0; when the range starts at site 0 the rows of
  // site 0 (the seed column) go to slot 3, which the loop first reuses for site sA + 3.
  const int sA = (s0 == 0) ? 1 : s0;
  FwdWarp<NJ> w;
  {
    const bool pre = PRE && 0 < a.start;
    const double *eq = pre ? ecq_pre : ecq;
    stage_k<NJ, 3>(g_sh, e_sh, s0 == 0, s0 == 0 ? rdq[0] : 0u, pre ? tabl_pre : tabl, eq, eq != nullptr, a.R, Gl, lane);
  }
#define MCQ_FWD_PROLOGUE(K)                                                                                  \
  {                                                                                                          \
    const int tt = sA + K;                                                                                   \
    const bool pre = PRE && tt < a.start;                                                                    \
    const double *eq = pre ? ecq_pre : ecq;                                                                  \
    stage_k<NJ, K>(g_sh, e_sh, tt <= a.end, tt <= a.end ? rdq[tt] : 0u, pre ? tabl_pre : tabl, eq + tt,      \
                   eq != nullptr, a.R + tt, Gl + (size_t)tt * (64 * NJ), lane);                              \
  }
  MCQ_FWD_PROLOGUE(0) MCQ_FWD_PROLOGUE(1) MCQ_FWD_PROLOGUE(2)
#undef MCQ_FWD_PROLOGUE
  w.t = sA + MCQ_PF;
#pragma unroll
  for (int k = 0; k < MCQ_RING; k++) {
    const int tt = w.t + k;                      // ring slot (tt - sA) & 3 = (k + 3) & 3
    w.rd[(k + MCQ_PF) & 3] = (tt <= a.end) ? rdq[tt] : 0u;
    w.sel[k] = (has_sel && sA + k <= a.end) ? selq[sA + k] : 0u;
  }
  w.gsrc = Gl + (size_t)w.t * (64 * NJ);
  w.rdp = rdq + w.t + MCQ_RING;
  w.rdsafe = rdq + a.end;
  w.Rp = a.R + w.t;
  w.Fs = X0 + (size_t)sA * Ls;
  w.Fnp = STORE ? Fno + sA : nullptr;
  w.selp = selq + sA + MCQ_RING;
  const double invN = 1.0 / (double)a.st.N;

  if (s0 == 0) {
    // column 0 is the bare emission and is never rescaled (forward_backward.c:27-39,102-111)
    cp_async_wait<MCQ_PF>();
    __syncwarp();
    unsigned g[(2 * NJ + 3) / 4];
    load_slots<NJ, 3>(gl, g);
    const double *es = ering + 3 * MCQ_ES;
    // column 0 belongs to the catch-up when start > 0
    const bool into1 = ((has_sel ? selq[0] : 0) != 0) != (!a.in_place && !(PRE && 0 < a.start));
    double *F0 = into1 ? X0 + delta : X0;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.p0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.p1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
      if (STORE && inrow[j]) __stcs(reinterpret_cast<double2 *>(F0 + 64 * j), make_double2(w.p0[j], w.p1[j]));
    }
    w.cnt = 0;
    if (STORE && lane == 0) {
      // update_F with start 0 zeroes both counters of column 0 (forward_backward.c:108-109)
      Fno[0] = 0;
      if (a.Fn_in != nullptr) Fnc[0] = 0;
    }
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));
  } else {
    const double *__restrict__ Fi = X0 + (size_t)(s0 - 1) * Ls + ((has_sel && selq[s0 - 1]) ? (a.Fout - a.Fin) : 0);
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.p0[j] = 0.0; w.p1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Fi + 64 * j);
        w.p0[j] = v.x; w.p1[j] = v.y;
      }
    }
    w.cnt = Fnc[s0 - 1];
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));
  }

  if constexpr (!PIPE) {
#define MCQ_FWD_SITE(K, SITE)                                                                             \
  fwd_site<NJ, FULL, PRE, STORE, K>(a, w, SITE, gl, ering, g_sh, e_sh, tabl, tabl_pre, ecq, ecq_pre, lane, \
                                    inrow, invN, delta, has_sel, Fnc)
    for (int s = sA; s <= a.end; s += 4) {
      MCQ_FWD_SITE(0, s);
      if (s + 1 > a.end) break;
      MCQ_FWD_SITE(1, s + 1);
      if (s + 2 > a.end) break;
      MCQ_FWD_SITE(2, s + 2);
      if (s + 3 > a.end) break;
      MCQ_FWD_SITE(3, s + 3);
    }
#undef MCQ_FWD_SITE
  } else {
    // cells of site s | front of site s+1 | butterfly, rescale, store of site s
#define MCQ_FWD_FRONT(K, SITE) \
  fwd_front<NJ, PRE, K>(a, w, SITE, gl, ering, g_sh, e_sh, tabl, tabl_pre, ecq, ecq_pre, lane)
#define MCQ_FWD_SITE(K, KN, SITE)                                                       \
  {                                                                                     \
    const double mine = fwd_cells<NJ>(w, invN);                                         \
    if ((SITE) + 1 <= a.end) MCQ_FWD_FRONT(KN, (SITE) + 1);                             \
    fwd_finish<NJ, FULL, PRE, STORE, K>(a, w, SITE, mine, lane, inrow, delta, has_sel, Fnc); \
  }
    if (sA <= a.end) MCQ_FWD_FRONT(0, sA);
    for (int s = sA; s <= a.end; s += 4) {
      MCQ_FWD_SITE(0, 1, s);
      if (s + 1 > a.end) break;
      MCQ_FWD_SITE(1, 2, s + 1);
      if (s + 2 > a.end) break;
      MCQ_FWD_SITE(2, 3, s + 2);
      if (s + 3 > a.end) break;
      MCQ_FWD_SITE(3, 0, s + 3);
    }
#undef MCQ_FWD_SITE
#undef MCQ_FWD_FRONT
  }
  cp_async_wait<0>();
  if (caught_up && lane == 0) a.fv[q] = a.start - 1;
  if (a.mark_fv != nullptr && lane == 0) a.mark_fv[q] = a.end;

  // K6 epilogue: the per-query log-likelihood at the last column of the range
  if (a.llk_mode != 0) {
    double dot;
    int nrm = w.cnt;
    if (a.llk_mode == 1) {                      // mcmc_moves.c:220-223
      const double *__restrict__ Bc = a.B + qoff + (size_t)a.end * Ls + 2 * lane;
      double loc = 0.0;
#pragma unroll
      for (int j = 0; j < NJ; j++)
        if (inrow[j]) {
          const double2 b = *reinterpret_cast<const double2 *>(Bc + 64 * j);
          loc += w.p0[j] * b.x; loc += w.p1[j] * b.y;
        }
      dot = warp_sum(loc);
      nrm += a.Bn[(size_t)q * S + a.end];
    } else {                                    // dmodel.c:899-901, mcmc_moves.c:1129-1134
      dot = w.tot;
    }
    if (lane == 0) a.llk_out[q] = log(dot) - nrm * a.logBIG;
  }
}

// ------------------------------------------------------------------------------------------
// Resident blocks per SM of the sweep kernels.  Every block lives for the whole sweep of its query, so how
// many queries are in flight decides both how the last wave of a launch fills and how many write streams
// the HBM sees at once; the best number depends on the kernel and on the number of queries (measured at
// C4: 17 per SM for 10 000 queries, 12 for 5 000, 10 for 2 500).  mcq_tune_* time the candidates once and
// record the winner per (kernel, grid size); launches look it up and lower the residency with unused
// dynamic shared memory.  Untuned launches run at the hardware maximum.
// ------------------------------------------------------------------------------------------
namespace {
struct Residency {
  std::mutex mu;
  std::unordered_map<const void *, int> maxb;                       // hardware maximum per kernel
  std::unordered_map<const void *, std::unordered_map<int, int>> pad;   // kernel -> blocks -> bytes
  std::unordered_map<const void *, std::unordered_map<unsigned, int>> tuned;   // kernel -> grid -> blocks
} g_res;

int max_blocks(const void *func) {           // (g_res.mu held)
  auto it = g_res.maxb.find(func);
  if (it != g_res.maxb.end()) return it->second;
  int b = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, func, 32 * MCQ_WPB, 0) != cudaSuccess) b = 0;
  g_res.maxb[func] = b;
  return b;
}

// smallest dynamic shared memory request that leaves at most `blocks` blocks resident (g_res.mu held)
int pad_for(const void *func, int blocks) {
  if (blocks <= 0 || blocks >= max_blocks(func)) return 0;
  auto &m = g_res.pad[func];
  auto it = m.find(blocks);
  if (it != m.end()) return it->second;
  int lo = 0, hi = 46 * 1024;
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, func, 32 * MCQ_WPB, (size_t)mid) != cudaSuccess) return 0;
    if (b <= blocks) hi = mid;
    else lo = mid + 1;
  }
  m[blocks] = lo;
  return lo;
}

int residency_smem(const void *func, unsigned grid) {
  std::lock_guard<std::mutex> lock(g_res.mu);
  auto k = g_res.tuned.find(func);
  if (k == g_res.tuned.end()) return 0;
  auto g = k->second.find(grid);
  return g == k->second.end() ? 0 : pad_for(func, g->second);
}

// times `launch` (twice per candidate, the second run counts) for every residency from the maximum down to
// 9 blocks per SM and records the fastest for this grid size
template <typename Launch>
cudaError_t tune_residency(const void *func, unsigned grid, cudaStream_t stream, Launch launch, int *best_blocks,
                           float *best_ms) {
  int maxb;
  { std::lock_guard<std::mutex> lock(g_res.mu); maxb = max_blocks(func); }
  cudaEvent_t e0, e1;
  cudaError_t err = cudaEventCreate(&e0);
  if (err != cudaSuccess) return err;
  err = cudaEventCreate(&e1);
  if (err != cudaSuccess) return err;
  int best = maxb;
  float best_t = 1e30f;
  for (int k = maxb; k >= 9 && k >= maxb - 8; k--) {
    int smem;
    { std::lock_guard<std::mutex> lock(g_res.mu); smem = pad_for(func, k); }
    float t = 0.f;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0, stream);
      launch(smem);
      cudaEventRecord(e1, stream);
      err = cudaEventSynchronize(e1);
      if (err != cudaSuccess) return err;
      cudaEventElapsedTime(&t, e0, e1);
    }
    if (t < best_t) { best_t = t; best = k; }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  { std::lock_guard<std::mutex> lock(g_res.mu); g_res.tuned[func][grid] = best; }
  if (best_blocks) *best_blocks = best;
  if (best_ms) *best_ms = best_t;
  return cudaGetLastError();
}
}  // namespace

template <int NJ, bool STORE>
static cudaError_t launch_forward_nj(const FwdArgs &a, cudaStream_t stream, bool tune, int *best_blocks, float *best_ms) {
  const unsigned blocks = (unsigned)((a.st.Q + MCQ_WPB - 1) / MCQ_WPB);
  const bool full = a.st.Ls == 64 * NJ;
  void (*kernel)(const FwdArgs);
  if (a.pipe && !tune) {     // few queries in flight: the variant with the shorter per-site chain (same bits)
    if (a.fv != nullptr) kernel = full ? k_forward<NJ, true, true, STORE, true> : k_forward<NJ, false, true, STORE, true>;
    else kernel = full ? k_forward<NJ, true, false, STORE, true> : k_forward<NJ, false, false, STORE, true>;
    kernel<<<blocks, 32 * MCQ_WPB, 0, stream>>>(a);
    return cudaGetLastError();
  }
  if (a.fv != nullptr) kernel = full ? k_forward<NJ, true, true, STORE, false> : k_forward<NJ, false, true, STORE, false>;
  else kernel = full ? k_forward<NJ, true, false, STORE, false> : k_forward<NJ, false, false, STORE, false>;
  if (tune)
    return tune_residency((const void *)kernel, blocks, stream,
                          [&](int smem) { kernel<<<blocks, 32 * MCQ_WPB, smem, stream>>>(a); }, best_blocks, best_ms);
  kernel<<<blocks, 32 * MCQ_WPB, residency_smem((const void *)kernel, blocks), stream>>>(a);
  return cudaGetLastError();
}

static cudaError_t dispatch_forward(const FwdArgs &a, cudaStream_t stream, bool tune, int *best_blocks, float *best_ms) {
  if (a.no_store) {                 // likelihood only: no column is written
    if (a.fv != nullptr || a.start != 0) return cudaErrorInvalidValue;
    switch (a.st.Lg / 64) {
      case 1: return launch_forward_nj<1, false>(a, stream, tune, best_blocks, best_ms);
      case 2: return launch_forward_nj<2, false>(a, stream, tune, best_blocks, best_ms);
      case 4: return launch_forward_nj<4, false>(a, stream, tune, best_blocks, best_ms);
      case 8: return launch_forward_nj<8, false>(a, stream, tune, best_blocks, best_ms);
      case 16: return launch_forward_nj<16, false>(a, stream, tune, best_blocks, best_ms);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (a.st.Lg / 64) {
    case 1: return launch_forward_nj<1, true>(a, stream, tune, best_blocks, best_ms);
    case 2: return launch_forward_nj<2, true>(a, stream, tune, best_blocks, best_ms);
    case 4: return launch_forward_nj<4, true>(a, stream, tune, best_blocks, best_ms);
    case 8: return launch_forward_nj<8, true>(a, stream, tune, best_blocks, best_ms);
    case 16: return launch_forward_nj<16, true>(a, stream, tune, best_blocks, best_ms);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t mcq_tune_forward(const FwdArgs &a, cudaStream_t stream, int *best_blocks, float *best_ms) {
  return dispatch_forward(a, stream, true, best_blocks, best_ms);
}

cudaError_t mcq_launch_forward(const FwdArgs &a, cudaStream_t stream) {
  return dispatch_forward(a, stream, false, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------
// K5: backward recursion, same mapping and staging, descending sites.  b = B[.][s+1] (as stored),
// e = the emission at s+1, sx = sum_r b*e.  The column sum of the new column and the next sx are
// reduced together in one butterfly.  Ring slot of site t = (tA - t) & 3, tA the first site whose rows
// are consumed (the column above the first one computed).
// ------------------------------------------------------------------------------------------
template <int NJ>
struct BwdWarp {
  double b0[NJ], b1[NJ], e0[NJ], e1[NJ];
  double sx, Rnext;
  int cnt, t;
  unsigned rd[MCQ_RING];      // descriptor words of sites t .. t-3 (entry = ring slot of the site)
  const uint8_t *gsrc;        // G row of site t (this lane's 16 bytes)
  const uint32_t *rdp;        // descriptor word of site t - MCQ_RING
  const uint32_t *rdsafe;     // a descriptor word that may always be read (the bottom site's)
  const double *ecp;          // consensus-row value of site t (unused when the rows carry none)
  const double *Rp;           // R of site t
  double *Bs;                 // output column of the site being computed
  int *Bnp;
  double n0[NJ], n1[NJ], Rs;  // the pipelined variant only: emissions and R of the site about to be computed
};

// stage site w.t into ring slot (K + MCQ_PF) & 3 and fetch the descriptor word needed four sites on
template <int NJ, int K, bool WHOLE = false>
__device__ __forceinline__ void bwd_stage(const BwdArgs &a, BwdWarp<NJ> &w, unsigned g_sh, unsigned e_sh,
                                          const double *tabl, int bottom, int lane) {
  // (the staging calls run K = 0, 1, 2, 3, 0, ...: the running pointers address the site of K = 0, constant
  // offsets inside the group of four, one bump per group)
  constexpr int KT = (K + MCQ_PF) & 3;
  const int t = w.t - K;
  const unsigned cur = w.rd[KT];
  // (see fwd_site: the sweep that stores whole columns keeps the select on the value - 0.510 ms against 0.586 ms at
  // 1 250 queries -, the checkpointed and pipelined ones load unconditionally)
  if constexpr (WHOLE) w.rd[KT] = (t - MCQ_RING >= bottom) ? *(w.rdp - K) : 0u;
  else w.rd[KT] = *((t - MCQ_RING >= bottom) ? w.rdp - K : w.rdsafe);
  stage_k<NJ, KT, WHOLE>(g_sh, e_sh, t >= bottom, cur, tabl, w.ecp - K, a.em.ec != nullptr, w.Rp - K,
                  w.gsrc - K * (64 * NJ), lane);
  if (K == 3) { w.t -= 4; w.rdp -= 4; w.ecp -= 4; w.Rp -= 4; w.gsrc -= 4 * (64 * NJ); }
}

// one site of the recursion: computes column s from column s+1; the rows of site s are in ring slot K
template <int NJ, bool FULL, int K>
__device__ __forceinline__ void bwd_site(const BwdArgs &a, BwdWarp<NJ> &w, const uint8_t *gl, const double *ering,
                                         unsigned g_sh, unsigned e_sh, const double *tabl, int bottom,
                                         int lane, const bool (&inrow)[NJ], const double (&m0)[NJ],
                                         const double (&m1)[NJ], double invN) {
  cp_async_wait<MCQ_PF - 1>();   // rows of site s
  __syncwarp();
  bwd_stage<NJ, K, true>(a, w, g_sh, e_sh, tabl, bottom, lane);
  unsigned g[(2 * NJ + 3) / 4];
  load_slots<NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  const double Rn = w.Rnext;
  w.Rnext = es[MCQ_RSLOT];                      // R[(s-1)+1] for the next site
  const double jump = w.sx * (Rn * invN);       // forward_backward.c:219 (reciprocal: see k_forward)
  const double stay = 1 - Rn;
  double locv = 0.0, locy = 0.0, locv2 = 0.0, locy2 = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    // emission at s, for the next step's products
    const double n0 = emis_at(es, g[j >> 1], (j & 1) * 2), n1 = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    const double v0 = __fma_rn(stay * w.b0[j], w.e0[j], jump) * m0[j];
    const double v1 = __fma_rn(stay * w.b1[j], w.e1[j], jump) * m1[j];
    w.b0[j] = v0; w.b1[j] = v1;
    if (j == 0) { locv = v0; locv2 = v1; } else { locv += v0; locv2 += v1; }
    locy = __fma_rn(v0, n0, locy); locy2 = __fma_rn(v1, n1, locy2);
    w.e0[j] = n0; w.e1[j] = n1;
  }
  locv += locv2; locy += locy2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    locv += __shfl_xor_sync(0xffffffffu, locv, o);
    locy += __shfl_xor_sync(0xffffffffu, locy, o);
  }
  w.sx = locy;
  if (locv < MCQ_BIGI) {                        // forward_backward.c:236-242
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJ; j++) { w.b0[j] *= MCQ_BIG; w.b1[j] *= MCQ_BIG; }
    w.sx = warp_sum(lane_dot<NJ>(w.b0, w.b1, w.e0, w.e1));   // of the column as stored (mcq_sweep_common.cuh)
  }
  // (the sites run K = 1, 2, 3, 0, 1, ...: the output pointers address the site of K = 1)
  constexpr int J = (K + 3) & 3;
  const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;   // doubles per column
  double *__restrict__ Bs = w.Bs - J * row;
#pragma unroll
  for (int j = 0; j < NJ; j++)
    if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Bs + 64 * j), make_double2(w.b0[j], w.b1[j]));
  if (lane == 0) *(w.Bnp - J) = w.cnt;
  if (K == 0) { w.Bs -= 4 * row; w.Bnp -= 4; }
}

// ---- the pipelined variant (see fwd_front): the front of column s-1 runs between the cell updates of column s
// and its butterfly.  Bit-identical to bwd_site.
template <int NJ, int K>
__device__ __forceinline__ void bwd_front(const BwdArgs &a, BwdWarp<NJ> &w, const uint8_t *gl, const double *ering,
                                          unsigned g_sh, unsigned e_sh, const double *tabl, int bottom, int lane) {
  cp_async_wait<MCQ_PF - 1>();
  __syncwarp();
  bwd_stage<NJ, K>(a, w, g_sh, e_sh, tabl, bottom, lane);
  unsigned g[(2 * NJ + 3) / 4];
  load_slots<NJ, K>(gl, g);
  const double *es = ering + K * MCQ_ES;
  w.Rs = es[MCQ_RSLOT];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    w.n0[j] = emis_at(es, g[j >> 1], (j & 1) * 2);
    w.n1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
  }
}

// cell updates of one column from the registers bwd_front filled; lane sums in locv / locy
template <int NJ>
__device__ __forceinline__ void bwd_cells(BwdWarp<NJ> &w, const double (&m0)[NJ], const double (&m1)[NJ], double invN,
                                          double &locv, double &locy) {
  const double Rn = w.Rnext;
  w.Rnext = w.Rs;
  const double jump = w.sx * (Rn * invN);
  const double stay = 1 - Rn;
  double v_a = 0.0, y_a = 0.0, v_b = 0.0, y_b = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const double v0 = __fma_rn(stay * w.b0[j], w.e0[j], jump) * m0[j];
    const double v1 = __fma_rn(stay * w.b1[j], w.e1[j], jump) * m1[j];
    w.b0[j] = v0; w.b1[j] = v1;
    if (j == 0) { v_a = v0; v_b = v1; } else { v_a += v0; v_b += v1; }
    y_a = __fma_rn(v0, w.n0[j], y_a); y_b = __fma_rn(v1, w.n1[j], y_b);
    w.e0[j] = w.n0[j]; w.e1[j] = w.n1[j];
  }
  locv = v_a + v_b; locy = y_a + y_b;
}

template <int NJ, bool FULL, int K>
__device__ __forceinline__ void bwd_finish(const BwdArgs &a, BwdWarp<NJ> &w, double locv, double locy, int lane,
                                           const bool (&inrow)[NJ]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    locv += __shfl_xor_sync(0xffffffffu, locv, o);
    locy += __shfl_xor_sync(0xffffffffu, locy, o);
  }
  w.sx = locy;
  if (locv < MCQ_BIGI) {                        // forward_backward.c:236-242
    ++w.cnt;
#pragma unroll
    for (int j = 0; j < NJ; j++) { w.b0[j] *= MCQ_BIG; w.b1[j] *= MCQ_BIG; }
    w.sx = warp_sum(lane_dot<NJ>(w.b0, w.b1, w.e0, w.e1));   // of the column as stored (mcq_sweep_common.cuh)
  }
  constexpr int J = (K + 3) & 3;
  const size_t row = FULL ? (size_t)(64 * NJ) : (size_t)a.st.Ls;
  double *__restrict__ Bs = w.Bs - J * row;
#pragma unroll
  for (int j = 0; j < NJ; j++)
    if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Bs + 64 * j), make_double2(w.b0[j], w.b1[j]));
  if (lane == 0) *(w.Bnp - J) = w.cnt;
  if (K == 0) { w.Bs -= 4 * row; w.Bnp -= 4; }
}

template <int NJ, bool FULL, bool PIPE>
__global__ void __launch_bounds__(32 * MCQ_WPB) k_backward(const BwdArgs a) {
  __shared__ __align__(16) uint8_t s_g[MCQ_WPB][MCQ_RING][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_WPB][MCQ_RING][MCQ_ES];
  const int q = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31, wib = MCQ_WPB == 1 ? 0 : (int)(threadIdx.x >> 5);   // (one warp per block: the rings sit at constant addresses)
  if (q >= a.st.Q) return;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) return;
  const int S = a.st.S, L = a.st.L, Ls = a.st.Ls;
  const size_t qoff = (size_t)q * S * Ls;
  const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
  const uint32_t *__restrict__ rdq = a.st.rowdesc + (size_t)q * S;
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const double *tabl = a.em.tab + lane;
  double *__restrict__ Bq = a.B + qoff + 2 * lane;
  int *__restrict__ Bnq = a.Bn + (size_t)q * S;
  uint8_t *gring = &s_g[wib][0][0];
  double *ering = &s_e[wib][0][0];
  if (lane < MCQ_RING) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;
  const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
  const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
  const uint8_t *gl = gring + lane * (2 * NJ);

  bool inrow[NJ];
  double m0[NJ], m1[NJ];   // 1 for real states, 0 for padding (r >= L): padding must not pick up `jump`
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const int r = 64 * j + 2 * lane;
    inrow[j] = FULL || (r < Ls);
    m0[j] = (r < L) ? 1.0 : 0.0;
    m1[j] = (r + 1 < L) ? 1.0 : 0.0;
  }

  BwdWarp<NJ> w;
  int s = a.update;
  const int bottom = a.bottom;
  if (a.bv != nullptr) {
    const int v = a.bv[q];
    if (v <= bottom) return;                    // already valid down to the requested column
    s = v - 1;
  }
  if (s == S - 1) {                             // forward_backward.c:197-203
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.b0[j] = m0[j]; w.b1[j] = m1[j];
      if (inrow[j]) __stcs(reinterpret_cast<double2 *>(Bq + (size_t)s * Ls + 64 * j), make_double2(w.b0[j], w.b1[j]));
    }
    w.cnt = 0;
    if (lane == 0) Bnq[s] = 0;
    s--;
  } else {
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.b0[j] = 0.0; w.b1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Bq + (size_t)(s + 1) * Ls + 64 * j);
        w.b0[j] = v.x; w.b1[j] = v.y;
      }
    }
    w.cnt = Bnq[s + 1];
  }
  if (s < bottom) {
    if (a.bv != nullptr && lane == 0) a.bv[q] = bottom;
    return;
  }

  // rows are consumed in the order tA = s+1, s, s-1, ..., bottom; slot of site t = (tA - t) & 3
  const int tA = s + 1;
#define MCQ_BWD_PROLOGUE(K)                                                                                  \
  {                                                                                                          \
    const int tt = tA - K;                                                                                   \
    stage_k<NJ, K>(g_sh, e_sh, tt >= bottom, tt >= bottom ? rdq[tt] : 0u, tabl, ecq + tt, ecq != nullptr,    \
                   a.R + tt, Gl + (ptrdiff_t)tt * (64 * NJ), lane);                                          \
  }
  MCQ_BWD_PROLOGUE(0) MCQ_BWD_PROLOGUE(1) MCQ_BWD_PROLOGUE(2)
#undef MCQ_BWD_PROLOGUE
  w.t = tA - MCQ_PF;
#pragma unroll
  for (int k = 0; k < MCQ_RING; k++) {
    const int tt = w.t - k;                      // ring slot (tA - tt) & 3 = (k + 3) & 3
    w.rd[(k + MCQ_PF) & 3] = (tt >= bottom) ? rdq[tt] : 0u;
  }
  w.gsrc = Gl + (ptrdiff_t)w.t * (64 * NJ);
  w.rdp = rdq + (w.t - MCQ_RING);
  w.rdsafe = rdq + bottom;
  w.ecp = ecq + w.t;
  w.Rp = a.R + w.t;
  w.Bs = Bq + (size_t)s * Ls;
  w.Bnp = Bnq + s;
  const double invN = 1.0 / (double)a.st.N;

  // emission at s+1 and sx
  {
    cp_async_wait<MCQ_PF - 1>();
    __syncwarp();
    bwd_stage<NJ, 0>(a, w, g_sh, e_sh, tabl, bottom, lane);
    unsigned g[(2 * NJ + 3) / 4];
    load_slots<NJ, 0>(gl, g);
    const double *es = ering;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.e0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.e1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    }
    w.sx = warp_sum(lane_dot<NJ>(w.b0, w.b1, w.e0, w.e1));
    w.Rnext = es[MCQ_RSLOT];                    // R[s+1]
  }

  if constexpr (!PIPE) {
#define MCQ_BWD_SITE(K) \
  bwd_site<NJ, FULL, K>(a, w, gl, ering, g_sh, e_sh, tabl, bottom, lane, inrow, m0, m1, invN)
    for (; s >= bottom; s -= 4) {
      MCQ_BWD_SITE(1);
      if (s - 1 < bottom) break;
      MCQ_BWD_SITE(2);
      if (s - 2 < bottom) break;
      MCQ_BWD_SITE(3);
      if (s - 3 < bottom) break;
      MCQ_BWD_SITE(0);
    }
#undef MCQ_BWD_SITE
  } else {
    // cells of column s | front of column s-1 | butterfly, rescale, store of column s
#define MCQ_BWD_FRONT(K) bwd_front<NJ, K>(a, w, gl, ering, g_sh, e_sh, tabl, bottom, lane)
#define MCQ_BWD_SITE(K, KN, COL)                                       \
  {                                                                    \
    double locv, locy;                                                 \
    bwd_cells<NJ>(w, m0, m1, invN, locv, locy);                        \
    if ((COL) - 1 >= bottom) MCQ_BWD_FRONT(KN);                        \
    bwd_finish<NJ, FULL, K>(a, w, locv, locy, lane, inrow);            \
  }
    MCQ_BWD_FRONT(1);
    for (; s >= bottom; s -= 4) {
      MCQ_BWD_SITE(1, 2, s);
      if (s - 1 < bottom) break;
      MCQ_BWD_SITE(2, 3, s - 1);
      if (s - 2 < bottom) break;
      MCQ_BWD_SITE(3, 0, s - 2);
      if (s - 3 < bottom) break;
      MCQ_BWD_SITE(0, 1, s - 3);
    }
#undef MCQ_BWD_SITE
#undef MCQ_BWD_FRONT
  }
  cp_async_wait<0>();
  if (a.bv != nullptr && lane == 0) a.bv[q] = bottom;
}

template <int NJ>
static cudaError_t launch_backward_nj(const BwdArgs &a, cudaStream_t stream, bool tune, int *best_blocks, float *best_ms) {
  const unsigned blocks = (unsigned)((a.st.Q + MCQ_WPB - 1) / MCQ_WPB);
  if (a.pipe && !tune) {
    void (*kp)(const BwdArgs) = (a.st.Ls == 64 * NJ) ? k_backward<NJ, true, true> : k_backward<NJ, false, true>;
    kp<<<blocks, 32 * MCQ_WPB, 0, stream>>>(a);
    return cudaGetLastError();
  }
  void (*kernel)(const BwdArgs) = (a.st.Ls == 64 * NJ) ? k_backward<NJ, true, false> : k_backward<NJ, false, false>;
  if (tune)
    return tune_residency((const void *)kernel, blocks, stream,
                          [&](int smem) { kernel<<<blocks, 32 * MCQ_WPB, smem, stream>>>(a); }, best_blocks, best_ms);
  kernel<<<blocks, 32 * MCQ_WPB, residency_smem((const void *)kernel, blocks), stream>>>(a);
  return cudaGetLastError();
}

static cudaError_t dispatch_backward(const BwdArgs &a, cudaStream_t stream, bool tune, int *best_blocks, float *best_ms) {
  switch (a.st.Lg / 64) {
    case 1: return launch_backward_nj<1>(a, stream, tune, best_blocks, best_ms);
    case 2: return launch_backward_nj<2>(a, stream, tune, best_blocks, best_ms);
    case 4: return launch_backward_nj<4>(a, stream, tune, best_blocks, best_ms);
    case 8: return launch_backward_nj<8>(a, stream, tune, best_blocks, best_ms);
    case 16: return launch_backward_nj<16>(a, stream, tune, best_blocks, best_ms);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t mcq_launch_backward(const BwdArgs &a, cudaStream_t stream) {
  return dispatch_backward(a, stream, false, nullptr, nullptr);
}

cudaError_t mcq_tune_backward(const BwdArgs &a, cudaStream_t stream, int *best_blocks, float *best_ms) {
  return dispatch_backward(a, stream, true, best_blocks, best_ms);
}

// ------------------------------------------------------------------------------------------
// K4k / K5k: the sweeps over CHECKPOINTED matrices.  Only every ck-th column of F (sites s with s % ck == ck-1)
// and of B (sites with s % ck == 0) is kept in HBM; a column in between is recomputed from the nearest
// checkpoint - at most ck-1 sites of the 1 B/cell recursion - by whoever needs it.  A sweep then moves
// 1 + 8/ck bytes per cell instead of 9, and the matrices take 1/ck of the memory.  Same cell updates in the same
// order as k_forward / k_backward (the site halves fwd_front / fwd_cells, bwd_front / bwd_cells are theirs), so
// every column, wherever it is computed, has the bits it has there.
//   forward : starts at the last current checkpoint at or before start-1 (validity mark fv: checkpoints of sites
//             <= fv are current), runs the sites up to start-1 with the current rows (completing owed
//             checkpoints on the way), then start..end with the rows of the launch; checkpoints of start..end go
//             where the selector says (in place, or into the other buffer for a proposal); the column of site
//             `end` is handed out in Tcol.
//   backward: starts at the first current checkpoint at or above `bottom` (bv: checkpoints of sites >= bv are
//             current; or from the seed column), runs down to `bottom`, completing owed checkpoints, and hands
//             out the column of `bottom` in Bcol.
// ------------------------------------------------------------------------------------------
template <int NJ, bool FULL>
__global__ void __launch_bounds__(32) k_forward_ck(const FwdArgs a) {
  __shared__ __align__(16) uint8_t s_g[MCQ_RING][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_RING][MCQ_ES];
  const int q = blockIdx.x, lane = threadIdx.x;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) {
    if (lane == 0 && a.llk_out != nullptr) a.llk_out[q] = a.llk_keep[q];
    return;
  }
  const int S = a.st.S, Ls = a.st.Ls, ck = a.ck, km = ck - 1, NC = a.nc, cksh = __ffs(ck) - 1;
  const size_t qoff = (size_t)q * NC * Ls;
  const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
  const uint32_t *__restrict__ rdq = a.st.rowdesc + (size_t)q * S;
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const double *ecq_pre = a.em_pre.ec ? a.em_pre.ec + (size_t)q * S : nullptr;
  const double *tabl = a.em.tab + lane, *tabl_pre = a.em_pre.tab + lane;
  // the last current checkpoint at or before start - 1 (-1: none, the sweep starts with the seed column)
  const int cp_start = (a.start & ~km) - 1;
  int p = cp_start, known = S - 1;
  if (a.fv != nullptr) {
    known = a.fv[q];
    p = min(p, ((known + 1) & ~km) - 1);
  }
  const int s0 = p + 1;
  double *X0 = const_cast<double *>(a.Fin) + qoff + 2 * lane;
  const ptrdiff_t delta = a.Fout - a.Fin;
  const uint8_t *selq = a.sel + (size_t)q * NC;
  int *Fno = a.Fn_out + (size_t)q * NC;
  int *Fnc = const_cast<int *>(a.Fn_in) + (size_t)q * NC;
  uint8_t *gring = &s_g[0][0];
  double *ering = &s_e[0][0];
  if (lane < MCQ_RING) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;
  const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
  const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
  const uint8_t *gl = gring + lane * (2 * NJ);
  bool inrow[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) inrow[j] = FULL || (64 * j + 2 * lane < Ls);

  const int sA = (s0 == 0) ? 1 : s0;
  FwdWarp<NJ> w;
  {
    const bool pre = 0 < a.start;
    const double *eq = pre ? ecq_pre : ecq;
    stage_k<NJ, 3>(g_sh, e_sh, s0 == 0, s0 == 0 ? rdq[0] : 0u, pre ? tabl_pre : tabl, eq, eq != nullptr, a.R, Gl, lane);
  }
#define MCQ_FWD_PROLOGUE(K)                                                                                  \
  {                                                                                                          \
    const int tt = sA + K;                                                                                   \
    const bool pre = tt < a.start;                                                                           \
    const double *eq = pre ? ecq_pre : ecq;                                                                  \
    stage_k<NJ, K>(g_sh, e_sh, tt <= a.end, tt <= a.end ? rdq[tt] : 0u, pre ? tabl_pre : tabl, eq + tt,      \
                   eq != nullptr, a.R + tt, Gl + (size_t)tt * (64 * NJ), lane);                              \
  }
  MCQ_FWD_PROLOGUE(0) MCQ_FWD_PROLOGUE(1) MCQ_FWD_PROLOGUE(2)
#undef MCQ_FWD_PROLOGUE
  w.t = sA + MCQ_PF;
#pragma unroll
  for (int k = 0; k < MCQ_RING; k++) {
    const int tt = w.t + k;
    w.rd[(k + MCQ_PF) & 3] = (tt <= a.end) ? rdq[tt] : 0u;
  }
  w.gsrc = Gl + (size_t)w.t * (64 * NJ);
  w.rdp = rdq + w.t + MCQ_RING;
  w.rdsafe = rdq + a.end;
  w.Rp = a.R + w.t;
  const double invN = 1.0 / (double)a.st.N;

  if (s0 == 0) {
    // column 0 is the bare emission and is never rescaled (forward_backward.c:27-39,102-111); with ck > 1 site 0
    // is not a checkpoint
    cp_async_wait<MCQ_PF>();
    __syncwarp();
    unsigned g[(2 * NJ + 3) / 4];
    load_slots<NJ, 3>(gl, g);
    const double *es = ering + 3 * MCQ_ES;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.p0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.p1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    }
    w.cnt = 0;
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));
  } else {
    const int c = p >> cksh;
    const double *__restrict__ Fi = X0 + (size_t)c * Ls + (selq[c] ? delta : 0);
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.p0[j] = 0.0; w.p1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Fi + 64 * j);
        w.p0[j] = v.x; w.p1[j] = v.y;
      }
    }
    w.cnt = Fnc[c];
    w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));
  }

#define MCQ_FWD_SITE(K, SITE)                                                                                       \
  {                                                                                                                 \
    const int s_ = (SITE);                                                                                          \
    fwd_front<NJ, true, K>(a, w, s_, gl, ering, g_sh, e_sh, tabl, tabl_pre, ecq, ecq_pre, lane);                    \
    w.tot = warp_sum(fwd_cells<NJ>(w, invN));                                                                       \
    if (w.tot < MCQ_BIGI) {                       /* forward_backward.c:163-167 */                                  \
      ++w.cnt;                                                                                                      \
      _Pragma("unroll") for (int j = 0; j < NJ; j++) { w.p0[j] *= MCQ_BIG; w.p1[j] *= MCQ_BIG; }                    \
      w.tot = warp_sum(lane_sum<NJ>(w.p0, w.p1));                                                                   \
    }                                                                                                               \
    if ((s_ & km) == km && !a.no_store) {           /* a checkpoint */                                              \
      const int c = s_ >> cksh;                     /* ck is a power of two: no division inside the sweep */         \
      const bool pre = s_ < a.start;                                                                                \
      const bool into1 = (selq[c] != 0) != (!a.in_place && !pre);                                                   \
      double *__restrict__ Fs = X0 + (size_t)c * Ls + (into1 ? delta : 0);                                          \
      _Pragma("unroll") for (int j = 0; j < NJ; j++)                                                                \
        if (FULL || inrow[j]) __stcs(reinterpret_cast<double2 *>(Fs + 64 * j), make_double2(w.p0[j], w.p1[j]));     \
      if (lane == 0) (pre ? Fnc : Fno)[c] = w.cnt;                                                                  \
    }                                                                                                               \
  }
  for (int s = sA; s <= a.end; s += 4) {
    MCQ_FWD_SITE(0, s);
    if (s + 1 > a.end) break;
    MCQ_FWD_SITE(1, s + 1);
    if (s + 2 > a.end) break;
    MCQ_FWD_SITE(2, s + 2);
    if (s + 3 > a.end) break;
    MCQ_FWD_SITE(3, s + 3);
  }
#undef MCQ_FWD_SITE
  cp_async_wait<0>();
  if (lane == 0) {
    // the checkpoints up to start - 1 are current now; a full in-place sweep makes all of them current
    if (a.fv != nullptr && !a.no_store && known < cp_start) a.fv[q] = cp_start;
    if (a.mark_fv != nullptr) a.mark_fv[q] = a.end;
  }
  if (a.Tcol != nullptr) {
    double *__restrict__ Tc = a.Tcol + (size_t)q * Ls + 2 * lane;
#pragma unroll
    for (int j = 0; j < NJ; j++)
      if (FULL || inrow[j]) *reinterpret_cast<double2 *>(Tc + 64 * j) = make_double2(w.p0[j], w.p1[j]);
    if (lane == 0) a.Tncol[q] = w.cnt;
  }
  if (a.llk_mode == 2 && lane == 0) a.llk_out[q] = log(w.tot) - w.cnt * a.logBIG;   // dmodel.c:899-901
}

template <int NJ, bool FULL>
__global__ void __launch_bounds__(32) k_backward_ck(const BwdArgs a) {
  __shared__ __align__(16) uint8_t s_g[MCQ_RING][64 * NJ];
  __shared__ __align__(16) double s_e[MCQ_RING][MCQ_ES];
  const int q = blockIdx.x, lane = threadIdx.x;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) return;
  const int S = a.st.S, L = a.st.L, Ls = a.st.Ls, ck = a.ck, km = ck - 1, NC = a.nc, cksh = __ffs(ck) - 1;
  const int bottom = a.bottom;
  // checkpoints of sites >= vb are current
  const int vb = a.bv != nullptr ? a.bv[q] : a.update + 1;
  const bool deliver = a.Bcol != nullptr;
  if (!deliver && vb <= bottom) return;
  const size_t qoff = (size_t)q * NC * Ls;
  const uint8_t *__restrict__ Gl = a.st.G + (size_t)q * S * (64 * NJ) + lane * 16;
  const uint32_t *__restrict__ rdq = a.st.rowdesc + (size_t)q * S;
  const double *ecq = a.em.ec ? a.em.ec + (size_t)q * S : nullptr;
  const double *tabl = a.em.tab + lane;
  double *__restrict__ Bq = a.B + qoff + 2 * lane;
  int *__restrict__ Bnq = a.Bn + (size_t)q * NC;
  double *__restrict__ Bc = deliver ? a.Bcol + (size_t)q * Ls + 2 * lane : nullptr;
  uint8_t *gring = &s_g[0][0];
  double *ering = &s_e[0][0];
  if (lane < MCQ_RING) ering[lane * MCQ_ES + MCQ_NOSLOT] = 0.0;
  const unsigned g_sh = (unsigned)__cvta_generic_to_shared(gring) + lane * 16;
  const unsigned e_sh = (unsigned)__cvta_generic_to_shared(ering) + lane * 8;
  const uint8_t *gl = gring + lane * (2 * NJ);
  bool inrow[NJ];
  double m0[NJ], m1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const int r = 64 * j + 2 * lane;
    inrow[j] = FULL || (r < Ls);
    m0[j] = (r < L) ? 1.0 : 0.0;
    m1[j] = (r + 1 < L) ? 1.0 : 0.0;
  }
  BwdWarp<NJ> w;
  // the first current checkpoint at or above `bottom`, or the seed column (forward_backward.c:197-203)
  const int lo = max(vb, bottom);
  const int p = (lo + km) & ~km;
  int s;                                        // the first column to compute
  if (p <= S - 1) {
    const int c = p >> cksh;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.b0[j] = 0.0; w.b1[j] = 0.0;
      if (inrow[j]) {
        const double2 v = *reinterpret_cast<const double2 *>(Bq + (size_t)c * Ls + 64 * j);
        w.b0[j] = v.x; w.b1[j] = v.y;
      }
    }
    w.cnt = Bnq[c];
    s = p - 1;
  } else {
#pragma unroll
    for (int j = 0; j < NJ; j++) { w.b0[j] = m0[j]; w.b1[j] = m1[j]; }
    w.cnt = 0;
    if (((S - 1) & km) == 0) {
      const int c = (S - 1) >> cksh;
#pragma unroll
      for (int j = 0; j < NJ; j++)
        if (inrow[j]) __stcs(reinterpret_cast<double2 *>(Bq + (size_t)c * Ls + 64 * j), make_double2(w.b0[j], w.b1[j]));
      if (lane == 0) Bnq[c] = 0;
    }
    s = S - 2;
  }
  auto hand_out = [&]() {
    if (!deliver) return;
#pragma unroll
    for (int j = 0; j < NJ; j++)
      if (FULL || inrow[j]) *reinterpret_cast<double2 *>(Bc + 64 * j) = make_double2(w.b0[j], w.b1[j]);
    if (lane == 0) a.Bncol[q] = w.cnt;
  };
  if (s < bottom) {                             // the column asked for is the one just loaded / seeded
    hand_out();
    if (a.bv != nullptr && lane == 0 && bottom < vb) a.bv[q] = bottom;
    return;
  }
  const int tA = s + 1;
#define MCQ_BWD_PROLOGUE(K)                                                                                  \
  {                                                                                                          \
    const int tt = tA - K;                                                                                   \
    stage_k<NJ, K>(g_sh, e_sh, tt >= bottom, tt >= bottom ? rdq[tt] : 0u, tabl, ecq + tt, ecq != nullptr,    \
                   a.R + tt, Gl + (ptrdiff_t)tt * (64 * NJ), lane);                                          \
  }
  MCQ_BWD_PROLOGUE(0) MCQ_BWD_PROLOGUE(1) MCQ_BWD_PROLOGUE(2)
#undef MCQ_BWD_PROLOGUE
  w.t = tA - MCQ_PF;
#pragma unroll
  for (int k = 0; k < MCQ_RING; k++) {
    const int tt = w.t - k;
    w.rd[(k + MCQ_PF) & 3] = (tt >= bottom) ? rdq[tt] : 0u;
  }
  w.gsrc = Gl + (ptrdiff_t)w.t * (64 * NJ);
  w.rdp = rdq + (w.t - MCQ_RING);
  w.rdsafe = rdq + bottom;
  w.ecp = ecq + w.t;
  w.Rp = a.R + w.t;
  const double invN = 1.0 / (double)a.st.N;
  {
    cp_async_wait<MCQ_PF - 1>();
    __syncwarp();
    bwd_stage<NJ, 0>(a, w, g_sh, e_sh, tabl, bottom, lane);
    unsigned g[(2 * NJ + 3) / 4];
    load_slots<NJ, 0>(gl, g);
    const double *es = ering;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      w.e0[j] = emis_at(es, g[j >> 1], (j & 1) * 2); w.e1[j] = emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    }
    w.sx = warp_sum(lane_dot<NJ>(w.b0, w.b1, w.e0, w.e1));
    w.Rnext = es[MCQ_RSLOT];
  }
#define MCQ_BWD_SITE(K, COL)                                                                                        \
  {                                                                                                                 \
    const int s_ = (COL);                                                                                           \
    bwd_front<NJ, K>(a, w, gl, ering, g_sh, e_sh, tabl, bottom, lane);                                              \
    double locv, locy;                                                                                              \
    bwd_cells<NJ>(w, m0, m1, invN, locv, locy);                                                                     \
    _Pragma("unroll") for (int o = 16; o > 0; o >>= 1) {                                                            \
      locv += __shfl_xor_sync(0xffffffffu, locv, o);                                                                \
      locy += __shfl_xor_sync(0xffffffffu, locy, o);                                                                \
    }                                                                                                               \
    w.sx = locy;                                                                                                    \
    if (locv < MCQ_BIGI) {                        /* forward_backward.c:236-242 */                                  \
      ++w.cnt;                                                                                                      \
      _Pragma("unroll") for (int j = 0; j < NJ; j++) { w.b0[j] *= MCQ_BIG; w.b1[j] *= MCQ_BIG; }                    \
      w.sx = warp_sum(lane_dot<NJ>(w.b0, w.b1, w.e0, w.e1));                                                        \
    }                                                                                                               \
    if ((s_ & km) == 0) {                           /* a checkpoint */                                              \
      const int c = s_ >> cksh;                     /* ck is a power of two: no division inside the sweep */         \
      _Pragma("unroll") for (int j = 0; j < NJ; j++)                                                                \
        if (FULL || inrow[j])                                                                                       \
          __stcs(reinterpret_cast<double2 *>(Bq + (size_t)c * Ls + 64 * j), make_double2(w.b0[j], w.b1[j]));        \
      if (lane == 0) Bnq[c] = w.cnt;                                                                                \
    }                                                                                                               \
  }
  for (; s >= bottom; s -= 4) {
    MCQ_BWD_SITE(1, s);
    if (s - 1 < bottom) break;
    MCQ_BWD_SITE(2, s - 1);
    if (s - 2 < bottom) break;
    MCQ_BWD_SITE(3, s - 2);
    if (s - 3 < bottom) break;
    MCQ_BWD_SITE(0, s - 3);
  }
#undef MCQ_BWD_SITE
  cp_async_wait<0>();
  hand_out();
  if (a.bv != nullptr && lane == 0 && bottom < vb) a.bv[q] = bottom;
}

template <int NJ>
static cudaError_t launch_forward_ck_nj(const FwdArgs &a, cudaStream_t stream) {
  if (a.st.Ls == 64 * NJ) k_forward_ck<NJ, true><<<(unsigned)a.st.Q, 32, 0, stream>>>(a);
  else k_forward_ck<NJ, false><<<(unsigned)a.st.Q, 32, 0, stream>>>(a);
  return cudaGetLastError();
}
template <int NJ>
static cudaError_t launch_backward_ck_nj(const BwdArgs &a, cudaStream_t stream) {
  if (a.st.Ls == 64 * NJ) k_backward_ck<NJ, true><<<(unsigned)a.st.Q, 32, 0, stream>>>(a);
  else k_backward_ck<NJ, false><<<(unsigned)a.st.Q, 32, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t mcq_launch_forward_ck(const FwdArgs &a, cudaStream_t stream) {
  if (a.ck < 2 || (a.ck & (a.ck - 1)) != 0 || a.sel == nullptr || a.em_pre.tab == nullptr) return cudaErrorInvalidValue;
  switch (a.st.Lg / 64) {
    case 1: return launch_forward_ck_nj<1>(a, stream);
    case 2: return launch_forward_ck_nj<2>(a, stream);
    case 4: return launch_forward_ck_nj<4>(a, stream);
    case 8: return launch_forward_ck_nj<8>(a, stream);
    case 16: return launch_forward_ck_nj<16>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t mcq_launch_backward_ck(const BwdArgs &a, cudaStream_t stream) {
  if (a.ck < 2 || (a.ck & (a.ck - 1)) != 0) return cudaErrorInvalidValue;
  switch (a.st.Lg / 64) {
    case 1: return launch_backward_ck_nj<1>(a, stream);
    case 2: return launch_backward_ck_nj<2>(a, stream);
    case 4: return launch_backward_ck_nj<4>(a, stream);
    case 8: return launch_backward_ck_nj<8>(a, stream);
    case 16: return launch_backward_ck_nj<16>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------
// K6: per-query log-likelihood of a proposal from its last forward column and the backward column at
// the same site (mcmc_moves.c:220-223): one warp per query.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dot(const DotArgs a) {
  const int q = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= a.st.Q) return;
  if (a.active != nullptr && a.active[(size_t)q * a.st.H + a.active_hla] == 0) {
    if (lane == 0) a.llk_out[q] = a.llk_keep[q];
    return;
  }
  const int S = a.st.S, Ls = a.st.Ls;
  const size_t col = a.cols ? (size_t)q * Ls : ((size_t)q * S + a.site) * Ls;
  const size_t cnt_at = a.cols ? (size_t)q : (size_t)q * S + a.site;
  // the proposal's column is in the buffer that does NOT hold the current one
  const double *Tq = (!a.cols && a.sel != nullptr && a.sel[(size_t)q * S + a.site] != 0) ? a.F : a.T;
  const double2 *__restrict__ t = reinterpret_cast<const double2 *>(Tq + col);
  const double2 *__restrict__ b = reinterpret_cast<const double2 *>(a.B + col);
  double dot;
  if (a.wpq <= 0) {
    double loc = 0.0;        // the same lane mapping and order as the epilogue of k_forward
    for (int i = lane; i < Ls / 2; i += 32) {
      const double2 x = t[i], y = b[i];
      loc += x.x * y.x; loc += x.y * y.y;
    }
    dot = warp_sum(loc);
  } else {
    // ... and of k_forward_mw: the 64-state groups of one consumer warp first, then the warps in order
    const int njw = (a.st.Lg / 64) / a.wpq;
    dot = 0.0;
    for (int w = 0; w < a.wpq; w++) {
      double loc = 0.0;
      for (int j = w * njw; j < (w + 1) * njw; j++) {
        const int i = lane + 32 * j;
        if (i < Ls / 2) {
          const double2 x = t[i], y = b[i];
          loc += x.x * y.x; loc += x.y * y.y;
        }
      }
      const double part = warp_sum(loc);
      dot = (w == 0) ? part : dot + part;
    }
  }
  if (lane == 0)
    a.llk_out[q] = log(dot) - (a.Tn[cnt_at] + a.Bn[cnt_at]) * a.logBIG;
}

cudaError_t mcq_launch_dot(const DotArgs &a, cudaStream_t stream) {
  k_dot<<<(unsigned)(((size_t)a.st.Q * 32 + 127) / 128), 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K4c: a ONE-COLUMN proposal (recombination_move, selection_move: mcmc_moves.c:213-224, 366-395) when nothing is
// owed before or behind the column.  The general sweep kernel spends its time here on set-up: descriptor, ring
// prologue, column load, row wait, each a dependent trip to memory.  Here every load a warp needs - the current
// column s-1, the backward column s, the slot bytes and the emission row of site s - is issued at once; one
// warp per query, four queries per block.  Same operations in the same order as k_forward + its epilogue: the
// per-query values are bit-identical.
// ------------------------------------------------------------------------------------------
template <int NJ, bool FULL>
__global__ void __launch_bounds__(128) k_column(const FwdArgs a) {
  __shared__ __align__(16) double s_e[4][MCQ_NOSLOT + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * 4 + wid;
  if (q >= a.st.Q) return;
  const int S = a.st.S, Ls = a.st.Ls, s = a.start;
  const size_t qs = (size_t)q * S, qoff = qs * Ls;
  const unsigned sel0 = a.sel ? a.sel[qs + s - 1] : 0u, selS = a.sel ? a.sel[qs + s] : 0u;
  const unsigned rd = a.st.rowdesc[qs + s];
  int cnt = a.Fn_in[qs + s - 1];
  const int bn = a.Bn[qs + s];
  const double Rs = (s == a.r_site) ? a.r_value : a.R[s];
  bool inrow[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) inrow[j] = FULL || (64 * j + 2 * lane < Ls);
  unsigned g[(2 * NJ + 3) / 4];
  load_slots<NJ, 0>(a.st.G + (qs + s) * (64 * NJ) + lane * (2 * NJ), g);
  const double *__restrict__ Fi = (sel0 ? a.Fout : a.Fin) + qoff + (size_t)(s - 1) * Ls + 2 * lane;
  const double *__restrict__ Bc = a.B + qoff + (size_t)s * Ls + 2 * lane;
  double p0[NJ], p1[NJ], b0[NJ], b1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    p0[j] = p1[j] = b0[j] = b1[j] = 0.0;
    if (inrow[j]) {
      const double2 v = *reinterpret_cast<const double2 *>(Fi + 64 * j);
      const double2 b = *reinterpret_cast<const double2 *>(Bc + 64 * j);
      p0[j] = v.x; p1[j] = v.y; b0[j] = b.x; b1[j] = b.y;
    }
  }
  double *es = s_e[wid];
  {
    const int kr = (int)(rd >> MCQ_RD_SHIFT);
    const double *row = a.em.tab + (rd & MCQ_RD_MASK);
    if (lane < kr) es[1 + lane] = row[lane];
    if (lane + 32 < kr) es[33 + lane] = row[lane + 32];
    if (lane == 0) {
      es[0] = a.em.ec ? a.em.ec[qs + s] : 0.0;
      es[MCQ_NOSLOT] = 0.0;
    }
  }
  __syncwarp();
  double tot = warp_sum(lane_sum<NJ>(p0, p1));
  const double jump = tot * (Rs * (1.0 / (double)a.st.N));
  const double stay = 1 - Rs;
  double la = 0.0, lb = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const double v0 = __fma_rn(stay, p0[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2);
    const double v1 = __fma_rn(stay, p1[j], jump) * emis_at(es, g[j >> 1], (j & 1) * 2 + 1);
    p0[j] = v0; p1[j] = v1;
    if (j == 0) { la = v0; lb = v1; } else { la += v0; lb += v1; }   // (0 + v = v: no add for the first)
  }
  tot = warp_sum(la + lb);
  if (tot < MCQ_BIGI) {                         // forward_backward.c:163-167
    ++cnt;
#pragma unroll
    for (int j = 0; j < NJ; j++) { p0[j] *= MCQ_BIG; p1[j] *= MCQ_BIG; }
  }
  // the proposal's column goes into the buffer that does not hold the current column s
  double *__restrict__ Fo = (selS ? const_cast<double *>(a.Fin) : a.Fout) + qoff + (size_t)s * Ls + 2 * lane;
  double dl = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++)
    if (inrow[j]) {
      __stcs(reinterpret_cast<double2 *>(Fo + 64 * j), make_double2(p0[j], p1[j]));
      dl += p0[j] * b0[j]; dl += p1[j] * b1[j];
    }
  const double dot = warp_sum(dl);
  if (lane == 0) {
    a.Fn_out[qs + s] = cnt;
    a.llk_out[q] = log(dot) - (cnt + bn) * a.logBIG;     // mcmc_moves.c:220-223
  }
}

template <int NJ>
static cudaError_t launch_column_nj(const FwdArgs &a, cudaStream_t stream) {
  const unsigned blocks = (unsigned)((a.st.Q + 3) / 4);
  if (a.st.Ls == 64 * NJ) k_column<NJ, true><<<blocks, 128, 0, stream>>>(a);
  else k_column<NJ, false><<<blocks, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// a.start == a.end > 0, every query active, columns start-1 of F and start of B current, llk_mode 1
cudaError_t mcq_launch_column(const FwdArgs &a, cudaStream_t stream) {
  if (a.start != a.end || a.start <= 0 || a.active != nullptr || a.fv != nullptr || a.llk_mode != 1) return cudaErrorInvalidValue;
  switch (a.st.Lg / 64) {
    case 1: return launch_column_nj<1>(a, stream);
    case 2: return launch_column_nj<2>(a, stream);
    case 4: return launch_column_nj<4>(a, stream);
    case 8: return launch_column_nj<8>(a, stream);
    case 16: return launch_column_nj<16>(a, stream);
    default: return cudaErrorInvalidValue;
  }
}

__global__ void __launch_bounds__(256) k_set_valid(int *fv, int *bv, int Q, int H, int f, int b,
                                                   const uint8_t *__restrict__ active, int hla) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (active != nullptr && active[(size_t)q * H + hla] == 0) return;
  if (f >= -1 && fv != nullptr) fv[q] = f;
  if (b >= 0 && bv != nullptr) bv[q] = b;
}

cudaError_t mcq_launch_set_valid(int *fv, int *bv, int Q, int H, int f, int b, const uint8_t *active, int hla,
                                 cudaStream_t stream) {
  k_set_valid<<<(Q + 255) / 256, 256, 0, stream>>>(fv, bv, Q, H, f, b, active, hla);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K7: accept-commit copies
// ------------------------------------------------------------------------------------------
// One launch makes an accepted window / column proposal current (mcmc_moves.c:263-265,447-449,959-965):
//   the parameter change itself (R / omega + rate_out_of_codons, mcmc_moves.c:345-352 / esc / rev);
//   the shared emission rows and the per-query consensus values of the range (tmp -> current);
//   the columns start..end: their selectors flip (the proposal was written into the buffer that did not hold
//   the current column, so nothing is copied) and the rescale counters follow;
//   with lazy refills, the validity marks of the queries touched.
// The parts touch different arrays: no ordering between them inside the kernel.
__global__ void __launch_bounds__(256) k_commit_all(const CommitArgs a) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const McqStatic &st = a.st;
  const McqOverride &ov = a.ov;
  if (blockIdx.x == 0) {
    const int t = threadIdx.x;
    switch (ov.kind) {
      case MCQ_PROP_REC:
        if (t == 0) a.m.R[ov.start] = ov.v0;
        break;
      case MCQ_PROP_SEL: {
        const int s = ov.start;
        if (t == 0) a.m.omega[s] = ov.v0;
        for (int k = st.aoff[s] + t - 1; k < st.aoff[s + 1]; k += blockDim.x) {
          int c;
          if (k < st.aoff[s]) {
            if (!st.cons_in_list[s]) continue;
            c = st.cons_codon[s];
          } else c = st.ncodon[k];
          a.m.rate_out[(size_t)s * 64 + c] = a.m.kappa * d_cnt[2][c] + d_cnt[3][c] + ov.v0 * (a.m.kappa * d_cnt[0][c] + d_cnt[1][c]);
        }
        break;
      }
      case MCQ_PROP_ESC:
        for (int s = ov.start + t; s <= ov.end; s += blockDim.x)
          a.m.esc[(size_t)ov.hla * st.S + s] = s < ov.split ? ov.v0 : ov.v1;
        break;
      case MCQ_PROP_REV:
        for (int s = ov.start + t; s <= ov.end; s += blockDim.x) a.m.rev[s] = s < ov.split ? ov.v0 : ov.v1;
        break;
      default: break;
    }
  }
  for (size_t i = a.t0 + tid; i < a.t1; i += nth) a.tab[i] = a.tab_t[i];
  for (size_t i = a.a0 + tid; i < a.a1; i += nth) a.atab[i] = a.atab_t[i];
  if (a.copy_cons) {
    const unsigned W = (unsigned)(ov.end - ov.start + 1), n = (unsigned)st.Q * W;
    for (unsigned i = tid; i < n; i += nth) {
      const size_t o = (size_t)(i / W) * st.S + ov.start + (i % W);
      a.ec[o] = a.ec_t[o]; a.ac[o] = a.ac_t[o];
    }
  }
  if (a.c1 >= a.c0) {
    const unsigned W = (unsigned)(a.c1 - a.c0 + 1), n = (unsigned)st.Q * W;
    for (unsigned i = tid; i < n; i += nth) {
      const unsigned q = i / W;
      if (a.active != nullptr && a.active[(size_t)q * st.H + a.active_hla] == 0) continue;
      const size_t o = (size_t)q * a.nc + a.c0 + (i % W);
      a.sel[o] ^= 1;
      a.Fn[o] = a.Tn[o];
    }
  }
  if (a.mark >= 0)
    for (unsigned q = tid; q < (unsigned)st.Q; q += nth) {
      if (a.active != nullptr && a.active[(size_t)q * st.H + a.active_hla] == 0) continue;
      a.fv[q] = a.mark; a.bv[q] = a.mark;
    }
}

cudaError_t mcq_launch_commit_all(const CommitArgs &a, cudaStream_t stream) {
  const unsigned n = (unsigned)a.st.Q * (unsigned)(a.ov.end - a.ov.start + 1);
  size_t most = std::max<size_t>(std::max<size_t>(n, a.t1 - a.t0), (size_t)a.st.Q);
  unsigned blocks = (unsigned)((most + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks == 0) blocks = 1;
  k_commit_all<<<blocks, 256, 0, stream>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K6b: total over queries.  One block; thread t adds elements t, t+1024, ... in ascending order,
// then a fixed tree - the result does not depend on scheduling.
// ------------------------------------------------------------------------------------------
// the total also goes straight to a host-mapped slot, followed by the evaluation number: the host spins on
// the number instead of paying for a copy and a stream synchronisation
__device__ __forceinline__ void publish_to_host(McqHostSlot *hs, double total, unsigned long long seq,
                                                const int *err = nullptr) {
  if (hs == nullptr) return;
  // the device error flag of the work before this reduction (range check of an asynchronous upload, ...)
  if (err != nullptr) *reinterpret_cast<volatile int *>(&hs->err) = *reinterpret_cast<const volatile int *>(err);
  *reinterpret_cast<volatile double *>(&hs->value) = total;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long *>(&hs->seq) = seq;
}

__global__ void __launch_bounds__(1024) k_sum(const double *__restrict__ v, int n, double *__restrict__ out,
                                              McqHostSlot *hs, unsigned long long seq, const int *err) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *out = sh[0];
    publish_to_host(hs, sh[0], seq, err);
  }
}

// ------------------------------------------------------------------------------------------
// K6c: the shard total followed, in the same kernel, by the cross-GPU exchange over NVLink peer
// memory.  Every rank owns a mailbox {double value[2][R]; u64 seq[2][R]} that all peers have mapped
// (CUDA IPC).  Lane r of warp 0 stores this rank's total and then the step number into peer r's
// mailbox (slot = step parity, entry = this rank), waits for peer r's entry in its own mailbox, and
// lane 0 adds the R totals in rank order - the same bits on every rank.  One step of latency over
// NVSwitch instead of a collective launch.  A peer that never arrives trips a clock-based timeout.
// ------------------------------------------------------------------------------------------
struct PeerBox {
  double *value[16];               // value[r] = peer r's mailbox values  [2][R]
  unsigned long long *seq[16];     // seq[r]   = peer r's mailbox flags   [2][R]
  unsigned long long *bar[16];     // bar[r]   = peer r's barrier counters [R]
  int nranks, rank;
};

// Barrier over the ranks of one box through the mailboxes: lane r stores the (monotone) barrier number
// into entry `me` of peer r's counters and waits until entry r of its own counters has reached it.
// Everything earlier on the launching stream (copies included) is complete and visible before the store.
__global__ void __launch_bounds__(32) k_peer_barrier(PeerBox box, unsigned long long count, int *__restrict__ err) {
  const int R = box.nranks, me = box.rank, lane = threadIdx.x;
  bool ok = true;
  if (lane < R) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(box.bar[lane] + me) = count;
    volatile unsigned long long *mine = box.bar[me] + lane;
    const long long t0 = clock64();
    while (*mine < count) {
      if (clock64() - t0 > 20000000000LL) { ok = false; break; }   // ~10 s
    }
    __threadfence_system();
  }
  if (!__all_sync(0xffffffffu, ok) && lane == 0) *err = 3;
}

cudaError_t mcq_launch_peer_barrier(const void *box, unsigned long long count, int *err, cudaStream_t stream) {
  k_peer_barrier<<<1, 32, 0, stream>>>(*static_cast<const PeerBox *>(box), count, err);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(1024) k_sum_exchange(const double *__restrict__ v, int n, double *__restrict__ out,
                                                       PeerBox box, unsigned long long step, int *__restrict__ err,
                                                       McqHostSlot *hs, unsigned long long seq) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x >= 32) return;
  const int R = box.nranks, me = box.rank, lane = threadIdx.x;
  const int slot = (int)(step & 1);
  double got = 0.0;
  bool ok = true;
  if (lane < R) {
    // publish: value first, then the step number (system-scope release)
    volatile double *pv = box.value[lane] + slot * R + me;
    volatile unsigned long long *ps = box.seq[lane] + slot * R + me;
    *pv = sh[0];
    __threadfence_system();
    *ps = step;
    // collect peer `lane`'s entry from this rank's own mailbox
    volatile unsigned long long *ms = box.seq[me] + slot * R + lane;
    volatile double *mv = box.value[me] + slot * R + lane;
    const long long t0 = clock64();
    while (*ms != step) {
      if (clock64() - t0 > 20000000000LL) { ok = false; break; }   // ~10 s
    }
    __threadfence_system();
    got = *mv;
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) {
      *err = 1;
      if (hs != nullptr) { *reinterpret_cast<volatile int *>(&hs->err) = 1; __threadfence_system(); }
      publish_to_host(hs, 0.0, seq);
    }
    return;
  }
  double total = 0.0;
  for (int r = 0; r < R; r++) total += __shfl_sync(0xffffffffu, got, r);   // rank order on every rank
  if (lane == 0) {
    *out = total;
    publish_to_host(hs, total, seq, err);
  }
}

cudaError_t mcq_launch_sum_exchange(const double *v, int n, double *out, const void *box, unsigned long long step,
                                    int *err, McqHostSlot *hs, unsigned long long seq, cudaStream_t stream) {
  k_sum_exchange<<<1, 1024, 0, stream>>>(v, n, out, *static_cast<const PeerBox *>(box), step, err, hs, seq);
  return cudaGetLastError();
}

cudaError_t mcq_launch_sum(const double *v, int n, double *out, McqHostSlot *hs, unsigned long long seq,
                           const int *err, cudaStream_t stream) {
  k_sum<<<1, 1024, 0, stream>>>(v, n, out, hs, seq, err);
  return cudaGetLastError();
}
