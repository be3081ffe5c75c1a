"""CPU: the C oracle against the committed golden vectors (outputs of the reference itself, made by
tests/golden/make_golden.py).  GPU: the CUDA path against the same vectors."""
import os

import numpy as np
import pytest

import cpu_libs
from problems import make_problem

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "toy_state.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_matches_golden_state(gold):
    p = make_problem("toy", gaps=True)
    assert np.array_equal(p.closest_refs, gold["problem_closest"])     # oracle closest-L == reference's
    st = cpu_libs.full_state(cpu_libs.Checker("oracle"), p)
    qs, sites = gold["queries"], gold["sites"]
    assert np.array_equal(st["llk"], gold["llk"])
    assert np.array_equal(st["Fn"], gold["Fn"]) and np.array_equal(st["Bn"], gold["Bn"])
    assert np.array_equal(st["F"][qs][:, :, sites], gold["F_cols"])
    assert np.array_equal(st["B"][qs][:, :, sites], gold["B_cols"])
    assert np.array_equal(st["c2c"][qs], gold["c2c"]) and np.array_equal(st["A"][qs], gold["A"])


def test_oracle_matches_golden_closest(gold):
    p = make_problem("toy", gaps=True)
    orc = cpu_libs.Checker("oracle")
    d = p.data
    for q in range(p.Q):
        assert np.array_equal(orc.hamming(d["ref_seqs"], d["query_seqs"][q]), gold["hamming"][q])
    for n in (1, 20, 21):
        cpu_libs.srandom(int(gold["closest_seed"]))
        got = np.stack([orc.closest(gold["hamming"][q], n) for q in range(p.Q)])
        assert np.array_equal(got, gold[f"closest_{n}"])


@pytest.mark.gpu
def test_gpu_matches_golden(gold):
    from mcqueen_b200 import api
    p = make_problem("toy", gaps=True)
    d = p.data
    assert np.array_equal(api.hamming_batch(d["ref_seqs"], d["query_seqs"]), gold["hamming"])
    for n, skip in ((1, False), (20, False), (21, False), (20, True)):
        cpu_libs.srandom(int(gold["closest_seed"]))
        got = api.closest_batch(d["ref_seqs"], d["query_seqs"], n, skip_first=skip)
        want = gold["closest_21"][:, 1:] if skip else gold[f"closest_{n}"]
        assert np.array_equal(got, want), (n, skip)
    ctx = api.Context.from_problem(p)
    llk = ctx.init_state()
    want = 0.0
    for v in gold["llk"]:
        want += float(v)
    assert abs(llk - want) <= 1e-12 * abs(want)
    assert np.max(np.abs(ctx.query_llk() - gold["llk"]) / np.abs(gold["llk"])) < 1e-12
    for i, q in enumerate(gold["queries"]):
        F, Fn = ctx.matrix(api.BUF_F, int(q))
        B, Bn = ctx.matrix(api.BUF_B, int(q))
        assert np.array_equal(Fn, gold["Fn"][q]) and np.array_equal(Bn, gold["Bn"][q])
        assert np.allclose(F[:, gold["sites"]], gold["F_cols"][i], rtol=1e-11, atol=0)
        assert np.allclose(B[:, gold["sites"]], gold["B_cols"][i], rtol=1e-11, atol=0)
        assert np.array_equal(ctx.table(api.TAB_C2C, int(q)), gold["c2c"][i])     # emission rows: bit-exact
        assert np.array_equal(ctx.table(api.TAB_A, int(q)), gold["A"][i])
    ctx.close()
