"""Seeded synthetic data of the shapes BASELINE.json names.

The reference's ``test_files/`` are absent and its simulator ``dsim`` cannot be linked
here, so every data set is generated: a root codon sequence drawn from the codon usage
``prior_C1`` (constants.c:93-109) without stop codons, a two-level reference panel
(clade founders a few per cent from the root, members a further few per cent from
their founder, every change a one-step codon substitution), and queries that are
mosaics of references (switching with probability ``rec`` per codon as in
simulate_sequences_sim.c:238-260) with one-step substitutions on top, boosted at
HLA-associated sites for carriers.  Output is ASCII ``TCAG`` (plus optional ``N``/``-``)
exactly as the reference's FASTA loader would hand it on (load_data.c:66-94).
"""
from __future__ import annotations

import numpy as np

from . import genetics as gen

__all__ = ["make_dataset", "CONFIGS", "dsim_rows", "dsim_queries", "dsim_simulate_in_reference_order",
           "write_simulation_parameters_json", "load_simulation_parameters_json"]

# the named shapes of BASELINE.json / SURVEY.md §8
CONFIGS = {
    "toy": dict(S=100, N=200, Q=50, L=20, H=4),
    "C2": dict(S=500, N=2000, Q=1000, L=100, H=20),
    "C3": dict(S=500, N=100000, Q=10000, L=500, H=20),
    "C4": dict(S=500, N=100000, Q=10000, L=500, H=20),
    "C5": dict(S=3000, N=50000, Q=50000, L=200, H=20),
    # one GPU's share of C5 when it is sharded 8 ways (self-reference mode: queries are references)
    "C5s": dict(S=3000, N=50000, Q=6250, L=200, H=20),
}
SELF_REFERENCE = {"C5", "C5s"}   # dmodel -q: the queries are the first Q references, L+1 found, self dropped

_ASCII = np.frombuffer(b"TCAG", dtype=np.uint8)


def _mutate(rng, codons, rate, neigh, ncount):
    """One-step substitution of a fraction ``rate`` of the codons (in place)."""
    mask = rng.random(codons.shape) < rate
    idx = np.nonzero(mask)
    if idx[0].size:
        c = codons[idx]
        pick = (rng.random(c.shape) * ncount[c]).astype(np.int64)
        codons[idx] = neigh[c, pick]
    return codons


def dsim_rows(frm, cons, omega, esc_prod, rev, mu, N, kappa=gen.KAPPA, phi=gen.PHI, prior=None):
    """The simulator's emission distribution (simulate_codon_to_codon_HLA,
    simulate_sequences_sim.c:14-137): P(new codon = to | copied codon = frm) for every `to`.
    All arguments broadcast against each other (frm/cons integer codons); returns (..., 64).
    These are the rows of the likelihood's emission table read the other way round
    (mcmc_moves.c:78-167), except that the simulator has no special case for TGG/ATG consensus."""
    prior = gen.normalised_prior() if prior is None else np.asarray(prior, float)
    frm, cons = np.asarray(frm, np.int64), np.asarray(cons, np.int64)
    omega, esc_prod, rev = (np.asarray(x, float) for x in (omega, esc_prod, rev))
    frm, cons, omega, esc_prod, rev = np.broadcast_arrays(frm, cons, omega, esc_prod, rev)
    ns = kappa * gen.NS_TS + gen.NS_TV            # (64,64) non-synonymous single steps
    sy = kappa * gen.S_TS + gen.S_TV              # synonymous single steps
    om = omega[..., None]
    step = om * ns[frm] + sy[frm]                 # (...,64): m(frm, to)
    rate0 = kappa * gen.BETA_S[frm] + gen.BETA_V[frm] + omega * (kappa * gen.ALPHA_S[frm] + gen.ALPHA_V[frm])
    back = (1 - phi) * prior                      # (64,)
    to = np.arange(64)
    is_cons_to = to == cons[..., None]
    # copied codon is not the consensus: reversion boosts the step onto the consensus
    m_fc = omega * ns[frm, cons] + sy[frm, cons]
    rate_r = rate0 + (rev - 1) * m_fc
    A_r = 1 - N / (N + rate_r * 4 * mu)
    g_r = (phi / rate_r)[..., None]
    rows_r = A_r[..., None] * (np.where(is_cons_to, g_r * rev[..., None] * step, g_r * step) + back)
    rows_r = rows_r + (to == frm[..., None]) * (1 - A_r[..., None])
    # copied codon is the consensus: escape boosts every non-synonymous step out of it
    nsum = kappa * gen.ALPHA_S[cons] + gen.ALPHA_V[cons]
    rate_e = rate0 + (esc_prod - 1) * omega * nsum
    A_e = 1 - N / (N + rate_e * 4 * mu)
    g_e = (phi / rate_e)[..., None]
    rows_e = A_e[..., None] * (g_e * (esc_prod[..., None] * om * ns[frm] + sy[frm]) + back)
    stay = (1 - A_e) * (1 - back[cons]) + back[cons]
    rows_e = np.where(is_cons_to, stay[..., None], rows_e)
    return np.where((frm == cons)[..., None], rows_e, rows_r)


def dsim_queries(rng, ref_codons, consensus_codons, hla, omega, R, esc, rev, mu, chunk=256):
    """Queries drawn from the model itself, as the reference's dsim does
    (simulate_sequences_sim.c:154-300): a mosaic of references switching with probability R[s], each codon
    drawn from the emission distribution of the copied codon given the query's HLA profile.
    Returns (queries (Q,S) int8, copy_from (Q,S) int32)."""
    N, S = ref_codons.shape
    Q, H = hla.shape
    carrier = (hla == ord("1"))
    copy_from = np.empty((Q, S), np.int32)
    copy_from[:, 0] = rng.integers(0, N, Q)
    switch = rng.random((Q, S)) < R[None, :]
    fresh = rng.integers(0, N, (Q, S))
    for s in range(1, S):
        copy_from[:, s] = np.where(switch[:, s], fresh[:, s], copy_from[:, s - 1])
    out = np.empty((Q, S), np.int8)
    sites = np.arange(S)
    for q0 in range(0, Q, chunk):
        q1 = min(Q, q0 + chunk)
        frm = ref_codons[copy_from[q0:q1], sites[None, :]].astype(np.int64)
        # product of the escape coefficients of the HLA types the query carries
        esc_prod = np.exp(carrier[q0:q1].astype(float) @ np.log(esc))
        rows = dsim_rows(frm, consensus_codons[None, :], omega[None, :], esc_prod, rev[None, :], mu, N)
        cdf = np.cumsum(rows, axis=-1)
        u = rng.random((q1 - q0, S, 1)) * cdf[..., -1:]
        out[q0:q1] = np.minimum((u > cdf).sum(-1), 63).astype(np.int8)
    return out, copy_from


def dsim_simulate_in_reference_order(ref_codons, consensus_codons, omega, R, esc, rev, mu, hla_prevalences, ploidy,
                                     n_HLA, proportion_no_hla, num_sims, uniform=None, kappa=gen.KAPPA, phi=gen.PHI,
                                     prior=None):
    """simulate_sequences (simulate_sequences_sim.c:154-300) restated draw for draw: the HLA profile of every
    simulated query (one discrete_sampling_dist draw per gene copy, util.c:519-533), the queries without HLA
    information, the copied reference per codon (a switch with probability R[s]), and one codon per site drawn from
    the cumulative emission distribution (bfind_pos, :139-152).  ``uniform()`` returns random()/RAND_MAX of the
    libc stream the reference draws from (default: this process's libc), so that the same seed gives the reference's
    own output.  Returns (hla (num_sims,H) u8 '0'/'1', copy_from (num_sims,S) i32, sequences (num_sims,S) i32)."""
    if uniform is None:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.random.restype = ctypes.c_long
        rand_max = 2147483647.0

        def uniform():
            return libc.random() / rand_max
    N, S = ref_codons.shape
    H = int(sum(n_HLA))
    prev = np.asarray(hla_prevalences, float)
    hla = np.full((num_sims, H), ord("0"), np.uint8)
    for i in range(num_sims):
        base = 0
        for ng in n_HLA:
            for _ in range(ploidy):
                u = 1 * uniform()
                cur, b = prev[base], 0
                for k in range(ng):
                    if u < cur:
                        break
                    b += 1
                    cur += prev[base + k + 1] if base + k + 1 < prev.size else 0.0
                hla[i, base + b] = ord("1")
            base += ng
    for i in range(num_sims):
        if 1 * uniform() < proportion_no_hla:
            hla[i, :] = ord("0")
    copy_from = np.zeros((num_sims, S), np.int32)
    for i in range(num_sims):
        copy_from[i, 0] = int(N * uniform())
        for s in range(1, S):
            if 1 * uniform() < R[s]:
                copy_from[i, s] = int(N * uniform())
            else:
                copy_from[i, s] = copy_from[i, s - 1]
    frm = ref_codons[copy_from, np.arange(S)[None, :]].astype(np.int64)
    esc_prod = np.ones((num_sims, S))
    for h in range(H):          # product in HLA order, as the reference multiplies
        esc_prod = np.where((hla[:, h] == ord("1"))[:, None], esc_prod * esc[h][None, :], esc_prod)
    rows = dsim_rows(frm, np.asarray(consensus_codons)[None, :], omega[None, :], esc_prod, rev[None, :], mu, N,
                     kappa=kappa, phi=phi, prior=prior)
    cdf = np.cumsum(rows, axis=-1)
    seqs = np.zeros((num_sims, S), np.int32)
    for i in range(num_sims):
        for s in range(S):
            where = 1 * uniform()
            j = 0
            for c in range(64):
                if where > cdf[i, s, c]:
                    j += 1
                else:
                    break
            seqs[i, s] = j
    return hla, copy_from, seqs


def write_simulation_parameters_json(path, kappa, mu, omega, R, reversion, hla_profiles, hla_prevalences, n_HLA, ploidy):
    """The parameter file dsim reads (load_lengths_for_simulation_from_json / load_parameters_for_simulation_from_json,
    load_from_json.c:297-443): scalars, the HLA prevalences and gene sizes, and per site omega, R, the reversion
    coefficient and one escape coefficient per HLA type."""
    import json
    omega, R, reversion = (np.asarray(x, float) for x in (omega, R, reversion))
    prof = np.asarray(hla_profiles, float)            # (total_n_HLA, S)
    doc = {"kappa": float(kappa), "mu": float(mu), "codon_sequence_length": int(omega.size),
           "total_n_HLA": int(prof.shape[0]), "ploidy": int(ploidy), "n_genes": len(n_HLA),
           "HLA_prevalences": [float(x) for x in hla_prevalences], "n_HLA": [int(x) for x in n_HLA],
           "sites": [{"omega": float(omega[s]), "R": float(R[s]), "reversion": float(reversion[s]),
                      "HLA": [float(x) for x in prof[:, s]]} for s in range(omega.size)]}
    with open(path, "w") as f:
        json.dump(doc, f)


def load_simulation_parameters_json(path):
    """Reads the file above, with the reference's consistency checks (gene sizes add up to total_n_HLA, one entry
    per site)."""
    import gzip
    import json
    with open(path, "rb") as f:
        raw = f.read()
    doc = json.loads(gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw)
    S, H = int(doc["codon_sequence_length"]), int(doc["total_n_HLA"])
    n_HLA = [int(x) for x in doc["n_HLA"]]
    if sum(n_HLA) != H or len(n_HLA) != int(doc["n_genes"]):
        raise ValueError(f"Mismatch in the number of HLA types: [{sum(n_HLA)} vs {H}].")
    if len(doc["sites"]) != S:
        raise ValueError(f"Mismatch in sequence length: [{len(doc['sites'])} vs {S}].")
    sites = doc["sites"]
    return dict(kappa=float(doc["kappa"]), mu=float(doc["mu"]), ploidy=int(doc["ploidy"]), n_HLA=n_HLA,
                hla_prevalences=np.array(doc["HLA_prevalences"], float),
                omega=np.array([s["omega"] for s in sites], float), R=np.array([s["R"] for s in sites], float),
                reversion=np.array([s["reversion"] for s in sites], float),
                hla_profiles=np.array([s["HLA"] for s in sites], float).T.copy())


def make_dataset(seed=20261019, S=100, N=200, Q=50, H=4, n_clades=None,
                 clade_div=0.04, member_div=0.02, query_div=0.03, rec=0.01,
                 ref_unknown_frac=0.0, query_gap=False, hla_site_frac=0.05, hla_boost=3.0,
                 self_reference=False, query_model="simple"):
    """Returns dict(ref_seqs (N,3S) u8 ASCII, query_seqs (Q,3S) u8 ASCII, hla (Q,H) u8 '0'/'1',
    parents (Q,S) i32 = the true copied reference per codon)."""
    rng = np.random.default_rng(seed)
    prior = gen.PRIOR_C1.copy()
    prior[list(gen.STOP_CODONS)] = 0
    prior /= prior.sum()
    root = rng.choice(64, size=S, p=prior).astype(np.int8)

    # one-step neighbours that are not stop codons
    neigh = np.zeros((64, 9), np.int8)
    ncount = np.zeros(64, np.int64)
    for c in range(64):
        ns = [d for d in gen.one_step_neighbours(c) if d not in gen.STOP_CODONS]
        neigh[c, :len(ns)] = ns
        ncount[c] = len(ns)

    if n_clades is None:
        n_clades = max(4, int(round(N ** 0.5)))
    founders = np.repeat(root[None, :], n_clades, axis=0)
    _mutate(rng, founders, clade_div, neigh, ncount)
    clade_of = rng.integers(0, n_clades, size=N)
    refs = founders[clade_of].copy()
    _mutate(rng, refs, member_div, neigh, ncount)

    # HLA profiles: 2 genes x ploidy 2 -> at most 4 ones; 20 % of queries carry none
    hla = np.full((Q, H), ord("0"), np.uint8)
    for q in range(Q):
        if rng.random() < 0.2:
            continue
        k = min(H, 4)
        hla[q, rng.choice(H, size=rng.integers(1, k + 1), replace=False)] = ord("1")
    hla_sites = rng.random((H, S)) < hla_site_frac

    # queries: mosaic of references
    switch = rng.random((Q, S)) < rec
    switch[:, 0] = True
    donor = rng.integers(0, N, size=(Q, S))
    seg = np.cumsum(switch, axis=1) - 1
    first = np.where(switch, np.arange(S)[None, :], 0)
    first = np.maximum.accumulate(first, axis=1)
    parents = donor[np.arange(Q)[:, None], first].astype(np.int32)
    del seg
    queries = refs[parents, np.arange(S)[None, :]].copy()
    carrier_boost = ((hla == ord("1")).astype(np.float64) @ hla_sites.astype(np.float64)) > 0
    rate = np.where(carrier_boost, np.minimum(1.0, query_div * hla_boost), query_div)
    _mutate(rng, queries, rate, neigh, ncount)

    if query_model == "dsim":
        # the reference's own generative recipe; parameters as SURVEY.md §8d: omega ~ LogN(0, 0.5),
        # R = 0.01, reversion 1, escape 1 except a few (HLA, site) pairs at hla_boost; mu scaled so that a
        # codon mutates with a few per cent probability whatever the panel size
        codes = gen.base_codes(gen.codons_to_ascii(refs))
        counts = np.stack([(codes == b).sum(0) for b in range(4)])
        cons = gen.to_codons(counts.argmax(0).astype(np.int8)[None, :])[0].astype(np.int64)
        omega = np.exp(rng.normal(0.0, 0.5, S))
        esc = np.where(hla_sites, hla_boost, 1.0)
        mu_gen = query_div * N / (4 * 30.0)
        queries, parents = dsim_queries(rng, refs, cons, hla, omega, np.full(S, rec), esc, np.ones(S), mu_gen)
    if self_reference:
        queries = refs[:Q].copy()
    ref_seqs = gen.codons_to_ascii(refs)
    query_seqs = gen.codons_to_ascii(queries)
    if ref_unknown_frac > 0:
        rows = np.nonzero(rng.random(N) < ref_unknown_frac)[0]
        ref_seqs[rows, rng.integers(0, 3 * S, size=rows.size)] = ord("N")
    if query_gap:
        query_seqs[np.arange(Q), rng.integers(0, 3 * S, size=Q)] = ord("-")
    return dict(ref_seqs=ref_seqs, query_seqs=query_seqs, hla=hla, parents=parents)
