"""Full-size runs (BASELINE.json configs C3 and C4) checked through size-independent properties and
through the oracle on a sample of queries - the oracle cannot sweep 2.5e9 cells in a test."""
import os
import sys

import numpy as np
import pytest

import cpu_libs

pytestmark = pytest.mark.gpu
ROOT = cpu_libs.ROOT
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def api():
    from mcqueen_b200 import api as m
    return m


@pytest.fixture(scope="module")
def c4(api):
    import bench
    p = bench.build_shard("C4", 0, 10000, 0)
    ctx = api.Context.from_problem(p)
    yield p, ctx
    ctx.close()


def test_c4_full_likelihood_properties(api, c4):
    p, ctx = c4
    llk = ctx.init_state()
    assert np.isfinite(llk)
    per_q = ctx.query_llk()
    # idempotence and determinism: the same evaluation twice gives the same bits
    assert ctx.full_llk(True) == llk and ctx.full_llk(False) == llk
    assert np.array_equal(ctx.query_llk(), per_q)
    # the F.B identity at full size: every site gives the query's log-likelihood
    for q in (0, 4999, 9999):
        F, Fn = ctx.matrix(api.BUF_F, q)
        B, Bn = ctx.matrix(api.BUF_B, q)
        per_site = np.log((F * B).sum(0)) - (Fn + Bn) * np.log(1e6)
        assert np.max(np.abs(per_site - per_q[q])) < 1e-9 * abs(per_q[q])
    # a sample of queries against the oracle (full L = 500, S = 500)
    chk = cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")
    sample = [0, 1234, 9999]
    c2c = np.zeros((1, 64, p.S)); A = np.zeros_like(c2c)
    for q in sample:
        one = _single_query_view(p, q)
        chk.emission(one, c2c, A, 0, p.S - 1, -1, 2)
        F = np.zeros((p.L, p.S)); Fn = np.zeros(p.S, np.int32)
        chk.init_forward(one, 0, c2c[0], p.R, F, Fn)
        want = chk.site_llk(one, F, Fn, None, None, p.S - 1)
        assert abs(per_q[q] - want) <= 1e-12 * abs(want)
        g, gn = ctx.matrix(api.BUF_F, q)
        assert np.array_equal(gn, Fn) and np.allclose(g, F, rtol=1e-11, atol=0)
        assert np.array_equal(ctx.table(api.TAB_C2C, q), c2c[0])


def _single_query_view(p, q):
    import copy
    one = copy.copy(p)
    one.Q = 1
    one.query_codons = np.ascontiguousarray(p.query_codons[q:q + 1])
    one.cons_query_nucls = np.ascontiguousarray(p.cons_query_nucls[q:q + 1])
    one.closest_refs = np.ascontiguousarray(p.closest_refs[q:q + 1])
    one.hla = np.ascontiguousarray(p.hla[q:q + 1])
    return one


def test_c4_proposals_are_reversible_and_consistent(api, c4):
    p, ctx = c4
    llk = ctx.init_state()
    per_q = ctx.query_llk()
    S = p.S
    # proposals that change nothing return the current likelihood (evaluated at another site: last bits differ)
    same = [
        ctx.propose(api.PROP_REC, 250, value0=float(p.R[250])),
        ctx.propose(api.PROP_SEL, 17, value0=float(p.omega[17])),
        ctx.propose(api.PROP_MU, 0, value0=float(p.mu)),
    ]
    for v in same:
        assert abs(v - llk) <= 1e-11 * abs(llk)
    ctx.reject()
    assert np.array_equal(ctx.query_llk(), per_q)             # a rejected proposal leaves no trace
    # a real window proposal: reject -> state intact; accept -> llk becomes the proposed value and the F.B
    # identity still holds at sites on both sides of the window
    a, b = 200, 280
    v1 = ctx.propose(api.PROP_ESC, a, b, hla=3, value0=1.7, value1=0.6, split=240)
    ctx.reject()
    v2 = ctx.propose(api.PROP_ESC, a, b, hla=3, value0=1.7, value1=0.6, split=240)
    assert v1 == v2 and v1 != llk
    ctx.accept()
    new_q = ctx.query_llk()
    assert abs(float(np.sum(new_q)) - v2) <= 1e-10 * abs(v2)
    for q in (5, 7777):
        F, Fn = ctx.matrix(api.BUF_F, q)
        B, Bn = ctx.matrix(api.BUF_B, q)
        per_site = np.log((F * B).sum(0)) - (Fn + Bn) * np.log(1e6)
        assert np.max(np.abs(per_site - new_q[q])) < 1e-9 * abs(new_q[q])
    # sparse evaluation of the same kind of move agrees with the dense one
    ctx.set_option(api.OPT_SPARSE_ESC, 1)
    s1 = ctx.propose(api.PROP_ESC, 50, 90, hla=1, value0=2.2)
    ctx.reject()
    ctx.set_option(api.OPT_SPARSE_ESC, 0)
    d1 = ctx.propose(api.PROP_ESC, 50, 90, hla=1, value0=2.2)
    ctx.reject()
    assert abs(s1 - d1) <= 1e-11 * abs(d1)


def test_c4_lazy_refills_and_likelihood_only_sweep(api, c4):
    """At the full C4 size: the sweep that writes no column gives the bits of the stored one; after a run of
    accepted moves with owed refills (selector flips, catch-ups inside later proposals) the F.B identity holds
    at every site, the state equals a fresh evaluation of the accepted parameters, and replaying the same
    moves with the reference's refill schedule gives the same proposal totals bit for bit."""
    p, ctx = c4
    ctx.set_option(api.OPT_SPARSE_ESC, 0)
    ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                   rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
    moves = [dict(kind=api.PROP_REV, start=300, end=380, value0=1.3),
             dict(kind=api.PROP_REC, start=120, value0=0.03),
             dict(kind=api.PROP_ESC, start=40, end=110, hla=2, value0=0.7, value1=1.6, split=80),
             dict(kind=api.PROP_SEL, start=450, value0=1.8),
             dict(kind=api.PROP_ESC, start=200, end=499, hla=5, value0=1.2)]
    totals = {}
    for lazy in (1, 0):
        ctx.set_option(api.OPT_LAZY_REFILL, lazy)
        ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, R=p.R, omega=p.omega,
                       rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
        llk = ctx.init_state()
        assert ctx.llk_only(False) == llk and ctx.llk_only(True) == llk
        out = []
        for mv in moves:
            out.append(ctx.propose(**mv))
            ctx.accept()
        totals[lazy] = out
        per_q = ctx.query_llk()
        assert abs(float(np.sum(per_q)) - out[-1]) <= 1e-10 * abs(out[-1])
        for q in (3, 6021):
            F, Fn = ctx.matrix(api.BUF_F, q)
            B, Bn = ctx.matrix(api.BUF_B, q)
            per_site = np.log((F * B).sum(0)) - (Fn + Bn) * np.log(1e6)
            assert np.max(np.abs(per_site - per_q[q])) < 1e-9 * abs(per_q[q])
        # the accepted parameters, evaluated from scratch
        fresh = ctx.llk_only(False)
        assert abs(fresh - out[-1]) <= 1e-11 * abs(out[-1])
    assert totals[0] == totals[1]
    ctx.set_option(api.OPT_LAZY_REFILL, 1)


def test_c3_closest_L_full_panel(api):
    """C3: 100 000 references x 1 500 nt, L = 500.  Bit-exact against the oracle on the first queries
    (the tie-shuffle stream of query k depends only on queries 0..k), structural properties on all."""
    from mcqueen_b200 import synth
    cfg = dict(synth.CONFIGS["C3"])
    L = cfg.pop("L")
    d = synth.make_dataset(seed=7, ref_unknown_frac=0.02, query_gap=True, **cfg)
    refs, qs = d["ref_seqs"], d["query_seqs"]
    chk = cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")
    K = 6
    cpu_libs.srandom(2026)
    want = np.stack([chk.closest(chk.hamming(refs, qs[q]), L) for q in range(K)])
    cpu_libs.srandom(2026)
    got = api.closest_batch(refs, np.ascontiguousarray(qs[:K]), L)
    assert np.array_equal(got, want)
    cpu_libs.srandom(2026)
    allq = api.closest_batch(refs, qs, L)
    assert np.array_equal(allq[:K], want)                    # same stream prefix in the full batch
    assert allq.min() >= 0 and allq.max() < refs.shape[0]
    srt = np.sort(allq, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])                 # no reference twice in a list
    ham = api.hamming_batch(refs, np.ascontiguousarray(qs[100:104]))
    for i, q in enumerate(range(100, 104)):
        dsel = ham[i][allq[q]]
        T = np.sort(ham[i])[L]                                # distance of sorted element L (one past the cut)
        assert np.all(dsel <= T)                              # nothing beyond the boundary distance is taken
        inside = dsel[dsel < T]
        assert np.all(np.diff(inside[:-1]) >= 0)             # ascending distance up to the tie window
        # every reference strictly closer than the boundary distance is in the list, except possibly one
        # (the element before the cut joins the shuffled tie window, closest_seqs.c:84-88)
        assert (ham[i] < T).sum() - (dsel < T).sum() in (0, 1)
