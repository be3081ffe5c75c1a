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

__all__ = ["make_dataset", "CONFIGS"]

# the named shapes of BASELINE.json / SURVEY.md §8
CONFIGS = {
    "toy": dict(S=100, N=200, Q=50, L=20, H=4),
    "C2": dict(S=500, N=2000, Q=1000, L=100, H=20),
    "C3": dict(S=500, N=100000, Q=10000, L=500, H=20),
    "C4": dict(S=500, N=100000, Q=10000, L=500, H=20),
    "C5": dict(S=3000, N=50000, Q=50000, L=200, H=20),
}

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


def make_dataset(seed=20261019, S=100, N=200, Q=50, H=4, n_clades=None,
                 clade_div=0.04, member_div=0.02, query_div=0.03, rec=0.01,
                 ref_unknown_frac=0.0, query_gap=False, hla_site_frac=0.05, hla_boost=3.0):
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

    ref_seqs = gen.codons_to_ascii(refs)
    query_seqs = gen.codons_to_ascii(queries)
    if ref_unknown_frac > 0:
        rows = np.nonzero(rng.random(N) < ref_unknown_frac)[0]
        ref_seqs[rows, rng.integers(0, 3 * S, size=rows.size)] = ord("N")
    if query_gap:
        query_seqs[np.arange(Q), rng.integers(0, 3 * S, size=Q)] = ord("-")
    return dict(ref_seqs=ref_seqs, query_seqs=query_seqs, hla=hla, parents=parents)
