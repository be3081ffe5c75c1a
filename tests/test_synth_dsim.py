"""The dsim-style generator (mcqueen_b200/synth.py) against the likelihood's emission table: the
simulator's P(to | copied codon) is the emission table read the other way round (CPU only)."""
import numpy as np

import cpu_libs
from mcqueen_b200 import genetics as gen, synth
from problems import make_problem


def test_rows_are_distributions_and_match_the_emission_table():
    p = make_problem("toy")
    orc = cpu_libs.Checker("oracle")
    S, N = p.S, p.N
    cons = p.consensus_codons.astype(np.int64)
    carrier = (p.hla == ord("1")).astype(float)
    esc_prod = np.exp(carrier @ np.log(p.esc))                       # (Q,S)
    for q in (0, 7):
        frm_all = np.arange(64)[:, None] * np.ones((1, S), np.int64)
        rows = synth.dsim_rows(frm_all, cons[None, :], p.omega[None, :], esc_prod[q][None, :], p.rev[0][None, :],
                               p.mu, N, p.kappa, p.phi, p.prior)     # (64 from, S, 64 to)
        assert np.allclose(rows.sum(-1), 1.0, atol=1e-4)             # the reference asserts the same band
        assert rows.min() >= 0
        # emission table for a pseudo query whose codon at every site is `to`
        for to in (int(p.query_codons[q, 3]), 15, 35, 40):
            one = _with_query(p, q, to)
            c2c = np.zeros((1, 64, S)); A = np.zeros_like(c2c)
            orc.emission(one, c2c, A, 0, S - 1, -1, 2)
            for s in range(0, S, 9):
                special = cons[s] in (15, 35)                        # TGG/ATG: the likelihood has a special case
                for k in range(p.n_site_codons[s]):
                    c1 = int(p.site_codons[s, k])
                    if special and c1 == cons[s]:
                        continue
                    assert abs(rows[c1, s, to] - c2c[0, c1, s]) <= 1e-12 * max(c2c[0, c1, s], 1e-300), (q, to, s, c1)


def _with_query(p, q, to):
    import copy
    one = copy.copy(p)
    one.Q = 1
    one.query_codons = np.full((1, p.S), to, np.int8)
    one.hla = np.ascontiguousarray(p.hla[q:q + 1])
    one.cons_query_nucls = np.ascontiguousarray(p.cons_query_nucls[q:q + 1])
    one.closest_refs = np.ascontiguousarray(p.closest_refs[q:q + 1])
    return one


def test_dsim_dataset_is_seeded_and_model_like():
    a = synth.make_dataset(seed=3, S=60, N=300, Q=40, H=4, query_model="dsim")
    b = synth.make_dataset(seed=3, S=60, N=300, Q=40, H=4, query_model="dsim")
    assert np.array_equal(a["query_seqs"], b["query_seqs"]) and np.array_equal(a["parents"], b["parents"])
    refc = gen.to_codons(gen.base_codes(a["ref_seqs"]))
    qc = gen.to_codons(gen.base_codes(a["query_seqs"]))
    copied = refc[a["parents"], np.arange(60)[None, :]]
    diff = (qc != copied).mean()
    assert 0.0 < diff < 0.15                                          # a few per cent of codons mutate
    assert (np.diff(a["parents"], axis=1) != 0).mean() < 0.05         # switches at about R = 0.01
