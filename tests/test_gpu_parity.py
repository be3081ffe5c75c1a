"""Parity of the CUDA path (through the C ABI) with the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): closest-L index lists bit-exact; log-likelihood within
1e-9 relative.  What is asserted here is tighter: emission rows bit-exact (integer-free fp64 with
FMA contraction off reproduces the reference's operation order), F/B cells within 1e-11 relative
(they differ from the reference only through the order of the per-site sum over copy states),
rescale counters exact, log-likelihoods within 1e-12 relative.
"""
import numpy as np
import pytest

import cpu_libs
from cpu_chain import CpuState, REC, SEL, ESC, REV, MU
from problems import make_problem

pytestmark = pytest.mark.gpu

RTOL_CELL = 1e-11
RTOL_LLK = 1e-12
RTOL_CONTRACT = 1e-9


def checker():
    # the compiled reference when it travelled with the snapshot, else the C restatement
    return cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def assert_state_matches(ctx, cpu, queries):
    from mcqueen_b200 import api
    for q in queries:
        for which, M, Mn in ((api.BUF_F, cpu.F, cpu.Fn), (api.BUF_B, cpu.B, cpu.Bn)):
            g, gn = ctx.matrix(which, q)
            assert np.array_equal(gn, Mn[q]), (which, q)
            assert rel(g, M[q]) < RTOL_CELL, (which, q, rel(g, M[q]))
        assert np.array_equal(ctx.table(api.TAB_C2C, q), cpu.c2c[q])
        assert np.array_equal(ctx.table(api.TAB_A, q), cpu.A[q])


@pytest.fixture(scope="module")
def api():
    from mcqueen_b200 import api as m
    return m


@pytest.mark.parametrize("gaps", [False, True])
def test_init_state(api, gaps):
    p = make_problem("toy", gaps=gaps)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    assert rel(ctx.query_llk(), cpu.llk_q) < RTOL_LLK
    assert_state_matches(ctx, cpu, range(p.Q))
    ctx.close()


@pytest.mark.parametrize("mode,which", [(0, 1), (0, -1), (1, 0), (2, -1), (2, 3)])
def test_emission_modes_bit_exact(api, mode, which):
    p = make_problem("toy")
    chk = checker()
    c2c = np.zeros((p.Q, 64, p.S)); A = np.zeros_like(c2c)
    chk.emission(p, c2c, A, 0, p.S - 1, -1, 2)
    ctx = api.Context.from_problem(p)
    ctx.emission(0, p.S - 1)
    esc = p.esc.copy(); rev = p.rev.copy()
    esc[:, 30:55] *= 1.7; rev[0, 30:55] *= 0.6
    chk.emission(p, c2c, A, 30, 54, which, mode, esc=esc, rev=rev)
    ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, esc=esc, rev=rev[0])
    ctx.emission(30, 54, which, mode)
    for q in range(p.Q):
        assert np.array_equal(ctx.table(api.TAB_C2C, q), c2c[q]), q
        assert np.array_equal(ctx.table(api.TAB_A, q), A[q]), q
    ctx.close()


def _moves(p, rng):
    S, H = p.S, p.H
    out = []
    s = int(rng.integers(1, S))
    out.append(dict(kind=REC, start=s, value0=float(p.R[s] * np.exp(rng.uniform(-2, 2)))))
    s = int(rng.integers(0, S))
    out.append(dict(kind=SEL, start=s, value0=float(p.omega[s] * np.exp(rng.uniform(-1, 1)))))
    a = int(rng.integers(0, S - 30)); b = a + int(rng.integers(1, 30))
    out.append(dict(kind=ESC, start=a, end=b, hla=int(rng.integers(0, H)), split=(a + b) // 2 + 1,
                    value0=float(np.exp(rng.normal())), value1=float(np.exp(rng.normal()))))
    a = int(rng.integers(0, S - 30)); b = a + int(rng.integers(1, 30))
    out.append(dict(kind=REV, start=a, end=b, value0=float(np.exp(rng.normal()))))
    out.append(dict(kind=MU, start=0, value0=float(p.mu * np.exp(rng.uniform(-0.125, 0.125)))))
    out.append(dict(kind=SEL, start=0, value0=0.7))
    out.append(dict(kind=SEL, start=S - 1, value0=1.9))
    out.append(dict(kind=REC, start=S - 1, value0=0.02))
    out.append(dict(kind=ESC, start=0, end=S - 1, hla=0, value0=1.3))
    return out


@pytest.mark.parametrize("sparse,lazy", [(0, 1), (1, 1), (0, 0), (1, 0)])
def test_proposals_and_accepts_track_the_oracle(api, sparse, lazy):
    """A scripted chain: every proposal's total matches, rejected ones leave no trace, accepted ones
    leave the same F/B/emission state as the reference's accept path - whether the refills after an
    accept run at once (lazy=0, the reference's schedule) or when first needed (lazy=1, the default)."""
    p = make_problem("toy", gaps=True)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    ctx.set_option(api.OPT_SPARSE_ESC, sparse)
    ctx.set_option(api.OPT_LAZY_REFILL, lazy)
    ctx.init_state()
    rng = np.random.default_rng(3)
    for rnd in range(3):
        for i, mv in enumerate(_moves(p, rng)):
            want = cpu.propose(**mv)
            got = ctx.propose(**mv)
            # sparse mode reuses cached per-query values for non-carriers: same to ~1e-13
            assert abs(got - want) <= (1e-11 if sparse else RTOL_LLK) * abs(want), (rnd, i, mv, got, want)
            if (i + rnd) % 2 == 0:
                cpu.accept(); ctx.accept()
            else:
                cpu.reject(); ctx.reject()
        assert_state_matches(ctx, cpu, [0, 7, 23, p.Q - 1])
        prm = ctx.params()
        assert np.array_equal(prm["R"], cpu.R) and np.array_equal(prm["omega"], cpu.omega)
        assert np.array_equal(prm["esc"], cpu.esc) and np.array_equal(prm["rev"], cpu.rev[0])
        assert np.array_equal(prm["rate_out"], cpu.rate_out) and prm["mu"] == cpu.mu
    ctx.close()


def test_proposal_columns_and_selector_flip(api):
    """A proposal's columns land in the buffer that does not hold the current ones (tmp_F of the
    reference, mcmc_moves.c:805) and an accept makes them current by flipping selectors: tmp_F after the
    proposal, F after the accept and after a second, overlapping accept all match the CPU chain."""
    p = make_problem("toy", gaps=True)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    ctx.init_state()
    for mv in (dict(kind=REV, start=10, end=30, value0=1.3), dict(kind=ESC, start=20, end=45, hla=2, value0=0.6),
               dict(kind=REC, start=25, value0=0.04), dict(kind=REV, start=0, end=p.S - 1, value0=0.9)):
        want, got = cpu.propose(**mv), ctx.propose(**mv)
        assert abs(got - want) <= RTOL_LLK * abs(want)
        end = mv.get("end", mv["start"])
        for q in (0, 11, p.Q - 1):
            T, Tn = ctx.matrix(api.BUF_TMP_F, q)
            sl = slice(mv["start"], end + 1)
            assert np.array_equal(Tn[sl], cpu.tFn[q][sl]) and rel(T[:, sl], cpu.tF[q][:, sl]) < RTOL_CELL
        cpu.accept(); ctx.accept()
        assert_state_matches(ctx, cpu, [0, 11, p.Q - 1])
    ctx.close()


@pytest.mark.parametrize("over", [dict(), dict(L=100, N=150), dict(L=500, N=600, S=60, Q=9), dict(L=1000, N=1100, S=40, Q=6)],
                         ids=["L20", "L100", "L500", "L1000"])
@pytest.mark.parametrize("sparse", [0, 1])
def test_lazy_refills_are_bit_identical_to_eager(api, sparse, over):
    """Owed refills (MCQ_OPT_LAZY_REFILL) compute the columns the reference would have computed at the
    accept, with the same kernels' arithmetic: every proposal total, every per-query value and the final
    F/B matrices are bit-identical to the eager schedule over a random chain with many accepts - at every number
    of states per lane (a sweep that restarts from a stored column carries the same column sum as one that
    continues from its registers: mcq_sweep_common.cuh, lane_sum / lane_dot)."""
    p = make_problem("toy", gaps=True, **over)
    ctxs = []
    for lazy in (0, 1):
        c = api.Context.from_problem(p)
        c.set_option(api.OPT_SPARSE_ESC, sparse)
        c.set_option(api.OPT_LAZY_REFILL, lazy)
        c.init_state()
        ctxs.append(c)
    eager, lz = ctxs
    rng = np.random.default_rng(11)
    for step in range(120):
        mv = _moves(p, rng)[int(rng.integers(0, 9))]
        a, b = eager.propose(**mv), lz.propose(**mv)
        assert a == b, (step, mv, a, b)
        assert np.array_equal(eager.query_llk(True), lz.query_llk(True)), (step, mv)
        if rng.random() < 0.6:
            eager.accept(); lz.accept()
        else:
            eager.reject(); lz.reject()
        if step % 40 == 39:
            for q in (0, 5, p.Q - 1):
                for which in (api.BUF_F, api.BUF_B):
                    (m0, n0), (m1, n1) = eager.matrix(which, q), lz.matrix(which, q)
                    assert np.array_equal(n0, n1) and np.array_equal(m0, m1), (step, which, q)
    lz.set_option(api.OPT_LAZY_REFILL, 0)      # leaving lazy mode completes what is owed
    mv = dict(kind=REC, start=3, value0=0.03)
    assert eager.propose(**mv) == lz.propose(**mv)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("over", [{}, dict(L=37, S=33, Q=9), dict(L=200, N=260, S=40, Q=6), dict(L=1000, N=1100, S=21, Q=4),
                                  dict(L=20, N=200, S=1, Q=5), dict(L=500, N=600, S=7, Q=3)])
def test_llk_only_equals_the_stored_sweep(api, over):
    """mcq_ctx_llk_only runs the same recursion without writing F: bit-identical per-query values, F and
    B untouched when the rows did not change, owed (and then recomputed correctly) when they did."""
    p = make_problem("toy", **over)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    per_query = ctx.query_llk().copy()
    assert ctx.llk_only(with_emission=False) == llk
    assert np.array_equal(ctx.query_llk(), per_query)
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    # new parameters: llk_only(with_emission) evaluates them; F and B follow when next read
    p2 = make_problem("toy", param_seed=5, **over)
    cpu2 = CpuState(checker(), p2)
    ctx.set_params(mu=p2.mu, kappa=p2.kappa, phi=p2.phi, R=p2.R, omega=p2.omega,
                   rate_out=p2.rate_out_of_codons(p2.omega), esc=p2.esc, rev=p2.rev[0])
    got = ctx.llk_only()
    assert abs(got - cpu2.llk) <= RTOL_LLK * abs(cpu2.llk)
    assert_state_matches(ctx, cpu2, [0, p.Q - 1])
    if p.S > 12:
        mv = dict(kind=ESC, start=3, end=11, hla=1, value0=1.7)
        want, got = cpu2.propose(**mv), ctx.propose(**mv)
        assert abs(got - want) <= RTOL_LLK * abs(want)
    ctx.close()


def test_dropin_symbols_match(api):
    p = make_problem("toy", gaps=True)
    chk = checker()
    base = cpu_libs.full_state(chk, p)
    S, L, N = p.S, p.L, p.N
    for q in (0, 11, p.Q - 1):
        cl = np.ascontiguousarray(p.closest_refs[q]); cons = np.ascontiguousarray(p.cons_query_nucls[q])
        F = np.zeros((L, S)); Fn = np.full(S, -1, np.int32)
        api.initialise_F_numerical_recipes(S, L, N, base["c2c"][q], p.ref_codons, p.ref_triplets, p.R, F, Fn,
                                           0, S - 1, cl, 3 * S, cons)
        assert np.array_equal(Fn, base["Fn"][q]) and rel(F, base["F"][q]) < RTOL_CELL
        B = np.zeros((L, S)); Bn = np.full(S, -1, np.int32)
        api.update_B_numerical_recipes(S, L, N, base["c2c"][q], p.ref_codons, p.ref_triplets, p.R, B, Bn,
                                       S - 1, cl, 3 * S, cons)
        assert np.array_equal(Bn, base["Bn"][q]) and rel(B, base["B"][q]) < RTOL_CELL
        # window update into a tmp matrix, then in-place suffix + prefix refills (the accept path)
        R2 = p.R.copy(); R2[20:41] *= 1.5
        tF = np.zeros((L, S)); tFn = np.zeros(S, np.int32)
        wF = np.zeros((L, S)); wFn = np.zeros(S, np.int32)
        api.update_F_numerical_recipes(S, L, N, base["c2c"][q], p.ref_codons, p.ref_triplets, R2, F, Fn, tFn, tF,
                                       20, 40, cl, 3 * S, cons)
        chk.forward(p, q, base["c2c"][q], R2, base["F"][q], base["Fn"][q], wFn, wF, 20, 40)
        assert np.array_equal(tFn[20:41], wFn[20:41]) and rel(tF[:, 20:41], wF[:, 20:41]) < RTOL_CELL
        F2, Fn2 = base["F"][q].copy(), base["Fn"][q].copy()
        G2, Gn2 = base["F"][q].copy(), base["Fn"][q].copy()
        api.update_F_numerical_recipes(S, L, N, base["c2c"][q], p.ref_codons, p.ref_triplets, R2, F2, Fn2, Fn2, F2,
                                       41, S - 1, cl, 3 * S, cons)
        chk.forward(p, q, base["c2c"][q], R2, G2, Gn2, Gn2, G2, 41, S - 1)
        assert np.array_equal(Fn2, Gn2) and rel(F2, G2) < RTOL_CELL
        B2, Bn2 = base["B"][q].copy(), base["Bn"][q].copy()
        H2, Hn2 = base["B"][q].copy(), base["Bn"][q].copy()
        api.update_B_numerical_recipes(S, L, N, base["c2c"][q], p.ref_codons, p.ref_triplets, R2, B2, Bn2, 40,
                                       cl, 3 * S, cons)
        chk.backward(p, q, base["c2c"][q], R2, H2, Hn2, 40)
        assert np.array_equal(Bn2, Hn2) and rel(B2, H2) < RTOL_CELL


@pytest.mark.parametrize("name,over", [("toy", {}), ("toy", dict(L=100, N=150)), ("toy", dict(L=37, S=33, Q=9)),
                                       ("toy", dict(L=200, N=260, S=40, Q=6)),      # NJ = 4, ragged row
                                       ("toy", dict(L=256, N=300, S=25, Q=5)),      # NJ = 4, full row
                                       ("toy", dict(L=1000, N=1100, S=21, Q=4)),    # NJ = 16, ragged
                                       ("toy", dict(L=1024, N=1100, S=12, Q=3)),    # NJ = 16, the maximum
                                       ("toy", dict(L=1, N=50, S=30, Q=7)),         # a single copy state
                                       ("toy", dict(L=20, N=200, S=1, Q=5))])       # a single site
def test_shapes_and_likelihood_identity(api, name, over):
    """Ragged shapes (L not a multiple of 16 or 64, tiny S) and the F.B identity: for every site s,
    log sum_r F[r][s] B[r][s] - (Fn+Bn) log BIG is the query's log-likelihood."""
    p = make_problem(name, **over)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    lq = ctx.query_llk()
    for q in (0, p.Q - 1):
        F, Fn = ctx.matrix(api.BUF_F, q)
        B, Bn = ctx.matrix(api.BUF_B, q)
        per_site = np.log((F * B).sum(0)) - (Fn + Bn) * np.log(1e6)
        assert np.max(np.abs(per_site - lq[q])) < 1e-9 * abs(lq[q])
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    if p.S > 2:      # one window proposal + accept through the same shapes
        mv = dict(kind=REV, start=1, end=min(p.S - 1, 9), value0=1.4)
        want, got = cpu.propose(**mv), ctx.propose(**mv)
        assert abs(got - want) <= RTOL_LLK * abs(want)
        cpu.accept(); ctx.accept()
        assert_state_matches(ctx, cpu, [0, p.Q - 1])
    ctx.close()


def test_contract_tolerance_on_c2_slice(api):
    """A C2-shaped slice (S=500, L=100, N=2000, H=20; 24 queries so the oracle finishes in seconds):
    total log-likelihood within the contract's 1e-9 relative (asserted at 1e-12)."""
    p = make_problem("C2", Q=24)
    chk = checker()
    st = cpu_libs.full_state(chk, p)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    want = CpuState._seq_sum(st["llk"])
    assert abs(llk - want) <= RTOL_LLK * abs(want) < RTOL_CONTRACT * abs(want)
    assert rel(ctx.query_llk(), st["llk"]) < RTOL_LLK
    for q in (0, 23):
        g, gn = ctx.matrix(api.BUF_B, q)
        assert np.array_equal(gn, st["Bn"][q]) and rel(g, st["B"][q]) < RTOL_CELL
    ctx.close()


@pytest.mark.parametrize("sparse", [0, 1])
def test_c5_shape_state_and_proposals(api, sparse):
    """The shape of BASELINE config 5 (3000 codons, L=200 -> ragged rows of 208 in the four-word kernels; 8 queries
    so that the CPU side finishes in seconds): initial state, every proposal kind, accepts and the owed refills
    they leave behind, against the CPU chain."""
    p = make_problem("toy", S=3000, N=600, Q=8, L=200, H=8, gaps=True)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    ctx.set_option(api.OPT_SPARSE_ESC, sparse)
    llk = ctx.init_state()
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    assert rel(ctx.query_llk(), cpu.llk_q) < RTOL_LLK
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    rng = np.random.default_rng(5)
    moves = _moves(p, rng) + [dict(kind=REV, start=2900, end=2999, value0=0.8),
                              dict(kind=ESC, start=1400, end=1650, hla=3, split=1500, value0=1.4, value1=0.7),
                              dict(kind=REC, start=1, value0=0.03)]
    for i, mv in enumerate(moves):
        want, got = cpu.propose(**mv), ctx.propose(**mv)
        assert abs(got - want) <= (1e-11 if sparse else RTOL_LLK) * abs(want), (i, mv, got, want)
        if i % 3 != 2:
            cpu.accept(); ctx.accept()
        else:
            cpu.reject(); ctx.reject()
    assert_state_matches(ctx, cpu, [0, 3, p.Q - 1])
    ctx.close()


def test_reupload_invalidates_the_state(api):
    """New static data makes F and B meaningless: proposals are refused until a state exists again
    (mcq_ctx_init_state, or mcq_ctx_full_llk + mcq_ctx_backward_fill), and that state matches the oracle."""
    import copy
    p = make_problem("toy")
    p2 = copy.copy(p)             # the same panel, the queries in reverse order
    for name in ("closest_refs", "cons_query_nucls", "query_codons", "hla"):
        setattr(p2, name, np.ascontiguousarray(getattr(p, name)[::-1]))
    cpu2 = CpuState(checker(), p2)
    ctx = api.Context.from_problem(p)
    ctx.init_state()
    ctx.propose(REC, 5, value0=0.02); ctx.accept()
    ctx.upload_static(p2.ref_triplets, p2.closest_refs, p2.cons_query_nucls, p2.query_codons, p2.hla,
                      p2.consensus_codons, p2.n_site_codons, p2.site_codons)
    ctx.set_params(mu=p2.mu, kappa=p2.kappa, phi=p2.phi, prior=p2.prior, R=p2.R, omega=p2.omega,
                   rate_out=p2.rate_out_of_codons(), esc=p2.esc, rev=p2.rev[0])
    with pytest.raises(api._lib.McqError, match="init_state"):
        ctx.propose(REC, 5, value0=0.02)
    llk = ctx.full_llk(True)
    assert abs(llk - cpu2.llk) <= RTOL_LLK * abs(cpu2.llk)
    ctx.backward_fill()
    mv = dict(kind=REV, start=10, end=30, value0=1.3)
    want, got = cpu2.propose(**mv), ctx.propose(**mv)
    assert abs(got - want) <= RTOL_LLK * abs(want)
    cpu2.accept(); ctx.accept()
    assert_state_matches(ctx, cpu2, [0, p.Q - 1])
    ctx.close()


@pytest.mark.parametrize("name,over", [("toy", {}), ("toy", dict(L=200, N=260, S=40, Q=70)), ("C2", dict(Q=64))])
def test_async_upload_pipelined_evaluation(api, name, over):
    """MCQ_OPT_ASYNC_UPLOAD: the upload returns after the DMAs, the next full evaluation runs chunk by chunk
    behind the device-side set-up.  Same bits as the synchronous path, repeatedly (the bench's e2e loop), a state
    can be built on it, and a bad closest_refs entry is reported by the evaluation."""
    p = make_problem(name, **over)
    ctx = api.Context.from_problem(p)
    want = ctx.full_llk(True)
    per_q = ctx.query_llk().copy()
    up = (p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
          p.n_site_codons, p.site_codons)
    prm = dict(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
               rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 1)
    ctx.set_option(api.OPT_UPLOAD_CHUNK_QUERIES, 5)      # several chunks even at test sizes
    for _ in range(3):
        ctx.upload_static(*up)
        ctx.set_params(**prm)
        assert ctx.full_llk(True) == want
        assert np.array_equal(ctx.query_llk(), per_q)
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 2)              # not even the DMAs are waited for (page-locked buffers)
    pinned = [np.ascontiguousarray(a) for a in up[:5]]
    for a in pinned:
        api.host_register(a)
    for _ in range(3):
        ctx.upload_static(*pinned, *up[5:])
        ctx.set_params(**prm)
        assert ctx.full_llk(True) == want
        assert np.array_equal(ctx.query_llk(), per_q)
    ctx.upload_static(*pinned, *up[5:])
    ctx.set_params(**prm)
    assert abs(ctx.init_state() - want) <= 1e-15 * abs(want)      # another consumer of a queued upload
    ctx.synchronize()
    for a in pinned:
        api.host_unregister(a)
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 1)
    ctx.set_option(api.OPT_UPLOAD_CHUNK_QUERIES, 1250)   # the default: one chunk at this size, no pipelining
    ctx.upload_static(*up)
    ctx.set_params(**prm)
    assert ctx.full_llk(True) == want
    assert np.array_equal(ctx.query_llk(), per_q)
    ctx.set_option(api.OPT_UPLOAD_CHUNK_QUERIES, 5)
    ctx.backward_fill()                       # full_llk + backward_fill = a state
    cpu = CpuState(checker(), p)
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    mv = dict(kind=REV, start=3, end=min(p.S - 1, 20), value0=1.2)
    w, g = cpu.propose(**mv), ctx.propose(**mv)
    assert abs(g - w) <= RTOL_LLK * abs(w)
    # another evaluation path behind an asynchronous upload: everything queues on the main stream
    ctx.upload_static(*up)
    ctx.set_params(**prm)
    assert abs(ctx.init_state() - want) <= 1e-15 * abs(want)
    bad = p.closest_refs.copy()
    bad[p.Q // 2, 1] = p.N
    ctx.upload_static(up[0], bad, *up[2:])
    ctx.set_params(**prm)
    with pytest.raises(api._lib.McqError, match="closest_refs"):
        ctx.full_llk(True)
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 0)
    ctx.upload_static(*up)
    ctx.set_params(**prm)
    assert ctx.full_llk(True) == want
    ctx.close()


def test_tune_leaves_state_and_results_unchanged(api):
    """mcq_ctx_tune only changes how many blocks of the sweep kernels are resident: same bits before and after,
    also for owed refills pending at the time of the call."""
    p = make_problem("toy", L=200, N=260, S=40, Q=70)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    mv = dict(kind=REV, start=5, end=20, value0=1.3)
    cpu.propose(**mv); cpu.accept()
    v = ctx.propose(**mv); ctx.accept()          # leaves refills owed
    k = ctx.tune()
    assert len(k) == 3 and all(x >= 1 for x in k)
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    assert ctx.full_llk(False) == ctx.llk_only(False)
    assert abs(ctx.full_llk(False) - v) <= 1e-12 * abs(v)
    mv2 = dict(kind=ESC, start=10, end=30, hla=1, value0=0.8)
    w, g = cpu.propose(**mv2), ctx.propose(**mv2)
    assert abs(g - w) <= RTOL_LLK * abs(w)
    ctx.close()


def test_errors_are_loud(api):
    with pytest.raises(api._lib.McqError):
        api.Context(10, 2000, 100, 5, 1)          # closest_n > 1024
    ctx = api.Context(10, 4, 100, 5, 1)
    with pytest.raises(api._lib.McqError):
        ctx.init_state()                           # nothing uploaded
    ctx.close()
    p = make_problem("toy")
    bad = p.closest_refs.copy()
    bad[3, 5] = p.N                                # one past the panel
    ctx = api.Context(p.S, p.L, p.N, p.Q, p.H)
    with pytest.raises(api._lib.McqError, match="closest_refs"):
        ctx.upload_static(p.ref_triplets, bad, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
                          p.n_site_codons, p.site_codons)
    ctx.close()
