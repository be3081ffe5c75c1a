"""Host-side set-up between sequence loading and the first likelihood evaluation.

Restates, in numpy, what the reference's ``dmodel`` does in ``main`` between
``generate_consensus`` and ``create_codon_to_codon_HLA`` (dmodel.c:489-862): global
consensus, per-query consensus of the closest references + gap fill, triplet/codon
encodings, per-site lists of reference codons, ``rate_out_of_codons`` and the initial
parameter vectors.  Output arrays use the reference's layouts so they can be handed
unchanged to the reference's own functions (oracle/_ref), to the C oracle and to the
C-ABI of libmcq_b200.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import genetics as gen


def majority_codes(codes):
    """Per-column majority base over rows of codes 0..4, ties to the lower code and 'T'
    when no base is present (amino_acids.c:94-141: strict '>' scan in T,C,A,G order)."""
    counts = np.stack([(codes == b).sum(axis=0) for b in range(4)], axis=0)
    return counts.argmax(axis=0).astype(np.int8)  # argmax returns the first maximum


@dataclass
class Problem:
    """Everything the hot path reads, in the reference's layouts."""
    S: int
    N: int
    Q: int
    L: int
    H: int
    ref_triplets: np.ndarray      # (N,S) i8 base-5, gaps kept      dmodel.c:684-686
    ref_codons: np.ndarray        # (N,S) i8 base-4, gaps -> consensus dmodel.c:693-697
    query_codons: np.ndarray      # (Q,S) i8                         dmodel.c:699-701
    cons_query_nucls: np.ndarray  # (Q,3S) i8 codes 0..3             dmodel.c:664-681
    closest_refs: np.ndarray      # (Q,L) i32
    hla: np.ndarray               # (Q,H) u8 '0'/'1'
    consensus_codons: np.ndarray  # (S,) i8                          dmodel.c:757-762
    n_site_codons: np.ndarray     # (S,) i32                         amino_acids.c:155-176
    site_codons: np.ndarray       # (S,64) i8
    ns_ts_sum: np.ndarray         # (S,) i32                         dmodel.c:791-802
    ns_tv_sum: np.ndarray
    type_of_change: np.ndarray    # (S,64,64) i8                     dmodel.c:815-834
    prior: np.ndarray             # (64,) f64 normalised             dmodel.c:867-869
    kappa: float = gen.KAPPA
    phi: float = gen.PHI
    # mutable parameters (initial values dmodel.c:564-577)
    mu: float = gen.INIT_MU
    omega: np.ndarray = field(default=None)
    R: np.ndarray = field(default=None)
    esc: np.ndarray = field(default=None)   # (H,S)
    rev: np.ndarray = field(default=None)   # (H,S), row 0 used

    def rate_out_of_codons(self, omega=None):
        """(64,S): kappa*beta_S + beta_V + omega*(kappa*alpha_S + alpha_V)  (dmodel.c:840-849)."""
        om = self.omega if omega is None else omega
        v1 = self.kappa * gen.BETA_S + gen.BETA_V
        v2 = self.kappa * gen.ALPHA_S + gen.ALPHA_V
        return np.ascontiguousarray(v1[:, None] + om[None, :] * v2[:, None])


def build_problem(ref_seqs, query_seqs, hla, closest_refs, kappa=gen.KAPPA, phi=gen.PHI,
                  prior=None, cons_q=None) -> Problem:
    """``cons_q`` (Q,3S) may carry precomputed per-query consensus codes (same rule, computed
    elsewhere, e.g. on the GPU for large shapes)."""
    ref_seqs = np.asarray(ref_seqs, np.uint8)
    query_seqs = np.asarray(query_seqs, np.uint8)
    N, nb = ref_seqs.shape
    Q = query_seqs.shape[0]
    assert nb % 3 == 0, "Sequences contain partial codons"
    S = nb // 3
    closest_refs = np.ascontiguousarray(closest_refs, np.int32)
    L = closest_refs.shape[1]
    hla = np.ascontiguousarray(hla, np.uint8)
    H = hla.shape[1]

    ref_codes = gen.base_codes(ref_seqs)        # 0..4
    qry_codes = gen.base_codes(query_seqs)
    consensus = majority_codes(ref_codes)       # consensus of the references (default -c)

    if cons_q is None:
        cons_q = np.empty((Q, nb), np.int8)
        for q in range(Q):
            cons_q[q] = majority_codes(ref_codes[closest_refs[q]])
    cons_q = np.ascontiguousarray(cons_q, np.int8)
    qry_filled = np.where(qry_codes == 4, cons_q, qry_codes)

    ref_triplets = gen.to_triplets(ref_codes)
    ref_filled = np.where(ref_codes == 4, consensus[None, :], ref_codes)
    ref_codons = gen.to_codons(ref_filled)
    query_codons = gen.to_codons(qry_filled)
    consensus_codons = gen.to_codons(consensus[None, :])[0]

    n_site = np.zeros(S, np.int32)
    site_codons = np.zeros((S, 64), np.int8)
    for s in range(S):
        cs = np.nonzero(np.bincount(ref_codons[:, s].astype(np.int64), minlength=64))[0]
        n_site[s] = cs.size
        site_codons[s, :cs.size] = cs

    cc = consensus_codons.astype(np.int64)
    ns_ts_sum = gen.NS_TS[cc].sum(1).astype(np.int32)
    ns_tv_sum = gen.NS_TV[cc].sum(1).astype(np.int32)
    toc = np.full((S, 64, 64), 4, np.int8)          # NONCON_TO_NONCON
    toc[np.arange(S), :, cc] = 3                    # NONCON_TO_CON
    toc[np.arange(S), cc, :] = 2                    # CON_TO_NONCON
    toc[np.arange(S), cc, cc] = 1                   # CON_TO_CON

    p = Problem(S=S, N=N, Q=Q, L=L, H=H,
                ref_triplets=np.ascontiguousarray(ref_triplets),
                ref_codons=np.ascontiguousarray(ref_codons),
                query_codons=np.ascontiguousarray(query_codons),
                cons_query_nucls=np.ascontiguousarray(cons_q),
                closest_refs=closest_refs, hla=hla,
                consensus_codons=np.ascontiguousarray(consensus_codons),
                n_site_codons=n_site, site_codons=site_codons,
                ns_ts_sum=ns_ts_sum, ns_tv_sum=ns_tv_sum, type_of_change=toc,
                prior=gen.normalised_prior(prior), kappa=kappa, phi=phi)
    p.omega = np.ones(S)
    p.R = np.full(S, gen.INIT_R)
    p.esc = np.ones((H, S))
    p.rev = np.ones((H, S))
    return p


def random_parameters(p: Problem, seed=1, n_windows=6):
    """Fill omega/R/esc/rev/mu with a plausible mid-chain state: log-normal omega, R around
    its initial value, piecewise-constant (windowed) escape and reversion coefficients."""
    rng = np.random.default_rng(seed)
    p.omega = np.exp(rng.normal(0.0, 0.5, p.S))
    p.R = np.clip(gen.INIT_R * np.exp(rng.normal(0.0, 1.0, p.S)), 1e-6, 0.5)
    p.mu = float(gen.INIT_MU * np.exp(rng.normal(0, 0.2)))

    def windows():
        k = min(n_windows, p.S)
        starts = np.sort(np.concatenate([[0], rng.choice(np.arange(1, p.S), k - 1, replace=False)]))
        coeff = np.exp(rng.normal(0.0, 0.7, k))
        out = np.empty(p.S)
        for i, st in enumerate(starts):
            en = starts[i + 1] if i + 1 < k else p.S
            out[st:en] = coeff[i]
        return out

    p.esc = np.stack([windows() for _ in range(p.H)])
    p.rev = np.ones((p.H, p.S))
    p.rev[0] = windows()
    return p
