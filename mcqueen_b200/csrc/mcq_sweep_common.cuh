// mcq_sweep_common.cuh - device helpers shared by the sweep kernels (mcq_kernels.cu: one warp per query;
// mcq_sweeps.cu: several warps per query): the warp butterfly, the cp.async staging of a site's rows into the
// shared-memory ring, and the emission look-up.
#ifndef MCQ_SWEEP_COMMON_CUH_
#define MCQ_SWEEP_COMMON_CUH_

#include "mcq_internal.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// The sum a sweep carries from one column to the next (the `ref_sum` of forward_backward.c:135-138, :218-227) is
// DEFINED on the column as it is stored - the reference re-forms it from F[.][s-1] / B[.][s+1] at every site - and
// formed in ONE order: each lane adds its even-state and its odd-state cells in two chains (fused multiply-adds for
// the backward products), adds the two, and the warp butterfly adds the lanes.  A sweep that continues from the
// registers, one that restarts from a stored column or a checkpoint, and the site after a rescaling (which re-forms
// the sum from the rescaled cells instead of scaling the sum) therefore carry the same bits: every refill schedule
// (eager, lazy, checkpointed at any stride) produces identical columns and totals.
// ------------------------------------------------------------------------------------------
template <int NJ>
__device__ __forceinline__ double lane_sum(const double (&p0)[NJ], const double (&p1)[NJ]) {
  double la = p0[0], lb = p1[0];
#pragma unroll
  for (int j = 1; j < NJ; j++) { la += p0[j]; lb += p1[j]; }
  return la + lb;
}
template <int NJ>
__device__ __forceinline__ double lane_dot(const double (&b0)[NJ], const double (&b1)[NJ], const double (&e0)[NJ],
                                           const double (&e1)[NJ]) {
  double ya = 0.0, yb = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; j++) { ya = __fma_rn(b0[j], e0[j], ya); yb = __fma_rn(b1[j], e1[j], yb); }
  return ya + yb;
}

// ------------------------------------------------------------------------------------------
// asynchronous global->shared copies (LDGSTS): a warp stages the codon-slot row G[q][s][.] and the
// emission row of the sites ahead of the one it is computing
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#define MCQ_PF 3                 // sites staged ahead of the one being computed
#define MCQ_RING (MCQ_PF + 1)    // ring slots per warp
#ifndef MCQ_WPB
#define MCQ_WPB 1                // warps (= queries) per block: small blocks balance best over the SMs
#endif
#define MCQ_ES 67                // doubles per staged row: emission slots 0..64, a constant 0 at MCQ_NOSLOT, R[s]
#define MCQ_RSLOT 66             // where the site's recombination probability travels

// What a warp needs to stage the emission row of (query, site) is one word of the static table rowdesc:
// bits 0..24 the offset of row (s, to(q,s)) in TAB, bits 25..31 the entries per row K_s (0 = nothing to
// stage, also used for sites outside the range).  The words are fetched MCQ_RING sites ahead of their use.
#define MCQ_RD_SHIFT 25
#define MCQ_RD_MASK 0x1FFFFFFu

// predicated forms: straight-line code, no branch around the copy
__device__ __forceinline__ void cp_async16_if(unsigned sa, const void *gmem, bool p) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}\n"
               ::"r"(sa), "l"(gmem), "r"((int)p) : "memory");
}
__device__ __forceinline__ void cp_async8_if(unsigned sa, const void *gmem, bool p) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}\n"
               ::"r"(sa), "l"(gmem), "r"((int)p) : "memory");
}

// Stage the rows of one site into ring slot KS (a constant): one commit group per call, empty when the
// site lies outside the range (its descriptor word is then 0 as well).  g_sh / e_sh: shared-window addresses
// of the rings, already offset by this lane's 16 / 8 bytes; gsrc = the site's G row, tabl = the emission
// table, both offset by this lane's share; ecp = the query's consensus-row value of the site (lane 0 copies
// it), Rp = the site's recombination probability (lane 1).
template <int NJ, int KS, bool RDZ = false>     // RDZ: the caller passes rd = 0 for a site that is not valid
__device__ __forceinline__ void stage_k(unsigned g_sh, unsigned e_sh, bool valid, unsigned rd,
                                        const double *__restrict__ tabl, const double *__restrict__ ecp, bool has_ec,
                                        const double *__restrict__ Rp, const uint8_t *__restrict__ gsrc, int lane) {
#pragma unroll
  for (int i = 0; i < (NJ + 7) / 8; i++)
    cp_async16_if(g_sh + KS * (64 * NJ) + 512 * i, gsrc + 512 * i, valid && (lane + 32 * i) * 16 < 64 * NJ);
  const int kr = (RDZ || valid) ? (int)(rd >> MCQ_RD_SHIFT) : 0;   // at most 64 entries: two predicated copies, no loop
  const double *row = tabl + (rd & MCQ_RD_MASK);
  const unsigned ed = e_sh + (KS * MCQ_ES + 1) * 8;
  cp_async8_if(ed, row, lane < kr);
  cp_async8_if(ed + 256, row + 32, lane + 32 < kr);
  // lane 0: the consensus-row value into slot 0; lane 1: R[s] into MCQ_RSLOT
  const bool l0 = lane == 0;
  cp_async8_if(l0 ? ed - 8 : ed + (MCQ_RSLOT - 2) * 8, l0 ? ecp : Rp, valid && (l0 ? has_ec : lane == 1));
  cp_async_commit();
}

// emission of the state whose slot is byte `b` of word w, from the staged row at `row`
__device__ __forceinline__ double emis_at(const double *row, unsigned w, int b) {
  const unsigned slot = __byte_perm(w, 0, 0x4440 + b);
  return *reinterpret_cast<const double *>(reinterpret_cast<const char *>(row) + (slot << 3));
}

// the 2*NJ slot bytes a lane owns at one site, from ring slot K of the staged G rows (gl = this lane's
// bytes in slot 0)
template <int NJ, int K>
__device__ __forceinline__ void load_slots(const uint8_t *gl, unsigned (&w)[(2 * NJ + 3) / 4]) {
  const uint8_t *p = gl + K * (64 * NJ);
  if constexpr (NJ == 1) w[0] = *reinterpret_cast<const uint16_t *>(p);
  else if constexpr (NJ == 2) w[0] = *reinterpret_cast<const uint32_t *>(p);
  else if constexpr (NJ == 4) { const uint2 v = *reinterpret_cast<const uint2 *>(p); w[0] = v.x; w[1] = v.y; }
  else {
#pragma unroll
    for (int i = 0; i < NJ / 8; i++) {
      const uint4 v = reinterpret_cast<const uint4 *>(p)[i];
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  }
}

#endif
