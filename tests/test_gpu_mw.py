"""The sweeps with several warps per query (MCQ_OPT_WARPS_PER_QUERY, mcqueen_b200/csrc/mcq_sweeps.cu) against the
CPU chain: same tolerances as the one-warp kernels (tests/test_gpu_parity.py) - F/B cells 1e-11, rescale counters
exact, log-likelihoods 1e-12 - over every (groups of 64 states, warps) pair that is built, ragged and full rows,
lazy and eager refills, carriers-only escape proposals, and the likelihood-only sweep."""
import numpy as np
import pytest

import cpu_libs
from cpu_chain import CpuState, REC, SEL, ESC, REV, MU
from problems import make_problem
from test_gpu_parity import RTOL_LLK, _moves, assert_state_matches, checker, rel

pytestmark = pytest.mark.gpu

SHAPES = [  # (make_problem overrides, warps per query)
    (dict(L=100, N=150), 1), (dict(L=100, N=150), 2),                       # NJ = 2, ragged rows (Ls = 112)
    (dict(L=128, N=150, S=40, Q=9), 2),                                     # NJ = 2, full rows
    (dict(L=200, N=260, S=40, Q=6), 2), (dict(L=200, N=260, S=40, Q=6), 4),  # NJ = 4, ragged
    (dict(L=256, N=300, S=25, Q=5), 4),                                     # NJ = 4, full
    (dict(L=500, N=600, S=50, Q=7), 4), (dict(L=500, N=600, S=50, Q=7), 8),  # NJ = 8 (the closest_n of config 4)
    (dict(L=300, N=400, S=30, Q=5), 8),                                     # NJ = 8, three empty 64-state groups
    (dict(L=1000, N=1100, S=21, Q=4), 8), (dict(L=1024, N=1100, S=12, Q=3), 8),   # NJ = 16
    (dict(L=100, N=150, S=1, Q=5), 2), (dict(L=100, N=150, S=2, Q=5), 2), (dict(L=100, N=150, S=5, Q=5), 1),
]


@pytest.fixture(scope="module")
def api():
    from mcqueen_b200 import api as m
    return m


@pytest.mark.parametrize("over,W", SHAPES)
def test_mw_state_proposals_and_accepts(api, over, W):
    p = make_problem("toy", gaps=True, **over)
    cpu = CpuState(checker(), p)
    ctx = api.Context.from_problem(p)
    ctx.set_option(api.OPT_WARPS_PER_QUERY, W)
    llk = ctx.init_state()
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    assert rel(ctx.query_llk(), cpu.llk_q) < RTOL_LLK
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    assert ctx.llk_only(False) == llk                      # the same recursion without the stores
    if p.S < 40:
        ctx.close()
        return
    rng = np.random.default_rng(3)
    for rnd in range(2):
        for i, mv in enumerate(_moves(p, rng)):
            want, got = cpu.propose(**mv), ctx.propose(**mv)
            assert abs(got - want) <= RTOL_LLK * abs(want), (rnd, i, mv, got, want)
            if (i + rnd) % 2 == 0:
                cpu.accept(); ctx.accept()
            else:
                cpu.reject(); ctx.reject()
        assert_state_matches(ctx, cpu, [0, p.Q // 2, p.Q - 1])
    ctx.close()


@pytest.mark.parametrize("over,W", [(dict(L=100, N=150), 2), (dict(L=500, N=600, S=60, Q=9), 4)])
@pytest.mark.parametrize("sparse", [0, 1])
def test_mw_lazy_refills_are_bit_identical_to_eager(api, over, W, sparse):
    p = make_problem("toy", gaps=True, **over)
    ctxs = []
    for lazy in (0, 1):
        c = api.Context.from_problem(p)
        c.set_option(api.OPT_WARPS_PER_QUERY, W)
        c.set_option(api.OPT_SPARSE_ESC, sparse)
        c.set_option(api.OPT_LAZY_REFILL, lazy)
        c.init_state()
        ctxs.append(c)
    eager, lz = ctxs
    rng = np.random.default_rng(11)
    for step in range(80):
        mv = _moves(p, rng)[int(rng.integers(0, 9))]
        a, b = eager.propose(**mv), lz.propose(**mv)
        assert a == b, (step, mv, a, b)
        assert np.array_equal(eager.query_llk(True), lz.query_llk(True)), (step, mv)
        if rng.random() < 0.6:
            eager.accept(); lz.accept()
        else:
            eager.reject(); lz.reject()
    for q in (0, p.Q - 1):
        for which in (api.BUF_F, api.BUF_B):
            (m0, n0), (m1, n1) = eager.matrix(which, q), lz.matrix(which, q)
            assert np.array_equal(n0, n1) and np.array_equal(m0, m1), (which, q)
    for c in ctxs:
        c.close()


def test_mw_option_is_checked(api):
    p = make_problem("toy")                                  # L = 20: one group of 64 states, nothing to split
    ctx = api.Context.from_problem(p)
    with pytest.raises(api._lib.McqError, match="warps per query"):
        ctx.set_option(api.OPT_WARPS_PER_QUERY, 2)
    ctx.close()
    p = make_problem("toy", L=100, N=150)
    ctx = api.Context.from_problem(p)
    ctx.init_state()
    with pytest.raises(api._lib.McqError, match="before mcq_ctx_init_state"):
        ctx.set_option(api.OPT_WARPS_PER_QUERY, 2)
    ctx.close()


def test_mw_async_upload_pipelined_evaluation(api):
    p = make_problem("toy", L=200, N=260, S=40, Q=70)
    ctx = api.Context.from_problem(p)
    ctx.set_option(api.OPT_WARPS_PER_QUERY, 4)
    want = ctx.full_llk(True)
    per_q = ctx.query_llk().copy()
    up = (p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
          p.n_site_codons, p.site_codons)
    prm = dict(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
               rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 1)
    ctx.set_option(api.OPT_UPLOAD_CHUNK_QUERIES, 5)
    for _ in range(3):
        ctx.upload_static(*up)
        ctx.set_params(**prm)
        assert ctx.full_llk(True) == want
        assert np.array_equal(ctx.query_llk(), per_q)
    ctx.close()


@pytest.mark.parametrize("over", [dict(), dict(L=100, N=150), dict(L=200, N=260, S=40, Q=6), dict(L=500, N=600, S=50, Q=7),
                                  dict(L=1000, N=1100, S=21, Q=4), dict(L=37, S=33, Q=9), dict(L=20, N=200, S=2, Q=5)])
def test_pipelined_sweeps_are_bit_identical(api, over):
    """MCQ_OPT_PIPELINE_MAX_QUERIES only reorders independent work inside a site: every total, every per-query
    value and the F / B matrices equal the plain kernels' bit for bit over a chain with accepts and owed refills."""
    p = make_problem("toy", gaps=True, **over)
    ctxs = []
    for pipe in (0, 1 << 30):
        c = api.Context.from_problem(p)
        c.set_option(api.OPT_PIPELINE_MAX_QUERIES, pipe)
        ctxs.append(c)
    a, b = ctxs
    assert a.init_state() == b.init_state()
    assert a.llk_only(False) == b.llk_only(False)
    rng = np.random.default_rng(7)
    if p.S >= 33:
        for step in range(40):
            mv = _moves(p, rng)[int(rng.integers(0, 9))]
            x, y = a.propose(**mv), b.propose(**mv)
            assert x == y, (step, mv, x, y)
            assert np.array_equal(a.query_llk(True), b.query_llk(True)), (step, mv)
            if rng.random() < 0.6:
                a.accept(); b.accept()
            else:
                a.reject(); b.reject()
    for q in (0, p.Q - 1):
        for which in (api.BUF_F, api.BUF_B):
            (m0, n0), (m1, n1) = a.matrix(which, q), b.matrix(which, q)
            assert np.array_equal(n0, n1) and np.array_equal(m0, m1), (which, q)
    for c in ctxs:
        c.close()
