"""MCQ_OPT_CHECKPOINT_STRIDE: only every k-th column of F, tmp_F and B is kept, the columns in between are
recomputed from the nearest kept one when a proposal needs them.  Everything a caller can observe must be what the
whole matrices give: every proposal total and per-query value, the kept columns against the matching columns of the
whole matrices, the state after accepts against the CPU chain - for ragged and full rows, lazy and eager refills,
carriers-only escape proposals, strides 2 to 16.

Tolerance: none.  The cell updates are the same operations in the same order, and the column sum a sweep carries is
defined on the stored column and formed in one order (mcq_sweep_common.cuh, lane_sum / lane_dot), so a sweep that
restarts from a kept column continues with the bits the whole-matrix sweep had in its registers: totals, per-query
values and kept columns are asserted equal (RTOL_CK = 0)."""
import numpy as np
import pytest

import cpu_libs
from cpu_chain import CpuState, REC, SEL, ESC, REV, MU
from problems import make_problem
from test_gpu_parity import RTOL_LLK, _moves, assert_state_matches, checker, rel

RTOL_CK = 0.0

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from mcqueen_b200 import api as m
    return m


def make_ctx(api, p, stride, sparse=0, lazy=1):
    ctx = api.Context(p.S, p.L, p.N, p.Q, p.H)
    ctx.set_option(api.OPT_CHECKPOINT_STRIDE, stride)
    ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
                      p.n_site_codons, p.site_codons)
    ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                   rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
    ctx.set_option(api.OPT_SPARSE_ESC, sparse)
    ctx.set_option(api.OPT_LAZY_REFILL, lazy)
    return ctx


SHAPES = [dict(), dict(L=100, N=150), dict(L=200, N=260, S=40, Q=6), dict(L=500, N=600, S=50, Q=7),
          dict(L=1000, N=1100, S=21, Q=4), dict(L=37, S=33, Q=9), dict(L=20, N=200, S=5, Q=5), dict(L=64, N=100, S=43, Q=4)]


@pytest.mark.parametrize("over", SHAPES)
@pytest.mark.parametrize("stride", [2, 4, 16])
def test_checkpointed_state_matches_the_oracle_and_the_whole_matrices(api, over, stride):
    p = make_problem("toy", gaps=True, **over)
    cpu = CpuState(checker(), p)
    ctx, whole = make_ctx(api, p, stride), make_ctx(api, p, 1)
    llk = ctx.init_state()
    assert llk == whole.init_state()                     # (one uninterrupted sweep either way: the same bits)
    assert abs(llk - cpu.llk) <= RTOL_LLK * abs(cpu.llk)
    assert np.array_equal(ctx.query_llk(), whole.query_llk())
    assert_state_matches(ctx, cpu, [0, p.Q - 1])
    assert ctx.device_bytes() < whole.device_bytes()
    rng = np.random.default_rng(3)
    for rnd in range(2 if p.S >= 33 else 0):
        for i, mv in enumerate(_moves(p, rng)):
            want, got, ref = cpu.propose(**mv), ctx.propose(**mv), whole.propose(**mv)
            assert abs(got - ref) <= RTOL_CK * abs(ref), (rnd, i, mv, got, ref)
            assert abs(got - want) <= RTOL_LLK * abs(want), (rnd, i, mv, got, want)
            assert rel(ctx.query_llk(True), whole.query_llk(True)) <= RTOL_CK
            if (i + rnd) % 2 == 0:
                cpu.accept(); ctx.accept(); whole.accept()
            else:
                cpu.reject(); ctx.reject(); whole.reject()
        assert_state_matches(ctx, cpu, [0, p.Q // 2, p.Q - 1])
    # the kept columns are the matching columns of the whole matrices
    for q in (0, p.Q - 1):
        for kept, full in ((api.BUF_F_KEPT, api.BUF_F), (api.BUF_B_KEPT, api.BUF_B)):
            (mk, nk), (mf, nf) = ctx.matrix(kept, q), whole.matrix(full, q)
            cols = np.nonzero(nk >= 0)[0]
            assert cols.size == (p.S // stride if kept == api.BUF_F_KEPT else (p.S + stride - 1) // stride)
            assert np.array_equal(nk[cols], nf[cols]) and rel(mk[:, cols], mf[:, cols]) <= RTOL_CK, (kept, q)
    ctx.close(); whole.close()


@pytest.mark.parametrize("sparse,lazy", [(0, 0), (1, 1), (1, 0)])
def test_checkpointed_chain_tracks_the_whole_matrices(api, sparse, lazy):
    p = make_problem("toy", gaps=True, L=100, N=150)
    ctx, whole = make_ctx(api, p, 4, sparse, lazy), make_ctx(api, p, 1, sparse, lazy)
    assert ctx.init_state() == whole.init_state()
    rng = np.random.default_rng(11)
    for step in range(100):
        mv = _moves(p, rng)[int(rng.integers(0, 9))]
        a, b = ctx.propose(**mv), whole.propose(**mv)
        assert abs(a - b) <= RTOL_CK * abs(b), (step, mv, a, b)
        assert rel(ctx.query_llk(True), whole.query_llk(True)) <= RTOL_CK, (step, mv)
        if rng.random() < 0.6:
            ctx.accept(); whole.accept()
        else:
            ctx.reject(); whole.reject()
    for q in (0, p.Q - 1):
        for which in (api.BUF_F, api.BUF_B):
            (m0, n0), (m1, n1) = ctx.matrix(which, q), whole.matrix(which, q)
            assert np.array_equal(n0, n1) and rel(m0, m1) <= RTOL_CK, (which, q)
    assert ctx.full_llk(True) == whole.full_llk(True)
    ctx.close(); whole.close()


def test_checkpointed_tmp_F_and_option_checks(api):
    p = make_problem("toy", gaps=True)
    cpu = CpuState(checker(), p)
    ctx = make_ctx(api, p, 4)
    ctx.init_state()
    mv = dict(kind=REV, start=10, end=30, value0=1.3)
    want, got = cpu.propose(**mv), ctx.propose(**mv)
    assert abs(got - want) <= RTOL_LLK * abs(want)
    for q in (0, 11):
        T, Tn = ctx.matrix(api.BUF_TMP_F, q)
        assert np.array_equal(Tn[10:31], cpu.tFn[q][10:31]) and rel(T[:, 10:31], cpu.tF[q][:, 10:31]) < 1e-11
    with pytest.raises(api._lib.McqError, match="before mcq_ctx_upload_static"):
        ctx.set_option(api.OPT_CHECKPOINT_STRIDE, 2)
    ctx.close()
    c2 = api.Context(p.S, p.L, p.N, p.Q, p.H)
    with pytest.raises(api._lib.McqError, match="power of two"):
        c2.set_option(api.OPT_CHECKPOINT_STRIDE, 3)
    c2.close()
