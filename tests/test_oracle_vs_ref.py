"""Pins the C oracle (oracle/mcq_oracle.c) against the reference's own code compiled
unmodified (oracle/_ref/libmcq_ref.so).  CPU only; skipped where oracle/_ref is absent."""
import numpy as np
import pytest

import cpu_libs
from problems import make_problem

pytestmark = pytest.mark.skipif(not cpu_libs.have_ref(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    return cpu_libs.Checker("oracle"), cpu_libs.Checker("ref")


@pytest.mark.parametrize("gaps", [False, True])
def test_full_state_bit_exact(libs, gaps):
    orc, ref = libs
    p = make_problem("toy", gaps=gaps)
    so, sr = cpu_libs.full_state(orc, p), cpu_libs.full_state(ref, p)
    for k in so:
        assert np.array_equal(so[k], sr[k]), k


@pytest.mark.parametrize("mode,which", [(0, 1), (0, -1), (1, 0), (2, -1), (2, 2)])
def test_partial_emission_updates(libs, mode, which):
    """window-restricted ESC / REV / BOTH updates on top of an existing table (mcmc_moves.c:789)."""
    orc, ref = libs
    p = make_problem("toy")
    rng = np.random.default_rng(5)
    outs = []
    for chk in (orc, ref):
        c2c = np.zeros((p.Q, 64, p.S)); A = np.zeros_like(c2c)
        chk.emission(p, c2c, A, 0, p.S - 1, -1, 2)
        esc = p.esc.copy(); rev = p.rev.copy()
        esc[:, 30:55] *= 1.7; rev[0, 30:55] *= 0.6
        chk.emission(p, c2c, A, 30, 54, which, mode, esc=esc, rev=rev)
        outs.append((c2c, A))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("start,end", [(1, 1), (40, 40), (99, 99), (17, 63), (0, 99), (60, 99)])
def test_window_forward_and_backward(libs, start, end):
    orc, ref = libs
    p = make_problem("toy", gaps=True)
    base = cpu_libs.full_state(ref, p)
    R2 = p.R.copy(); R2[start:end + 1] *= 1.9
    for q in (0, 7, 49):
        res = []
        for chk in (orc, ref):
            F = base["F"][q].copy(); Fn = base["Fn"][q].copy()
            tF = np.full_like(F, np.nan); tFn = np.full_like(Fn, -7)
            chk.forward(p, q, base["c2c"][q], R2, F, Fn, tFn, tF, start, end)
            B = base["B"][q].copy(); Bn = base["Bn"][q].copy()
            chk.backward(p, q, base["c2c"][q], R2, B, Bn, end)
            res.append((tF[:, start:end + 1].copy(), tFn[start:end + 1].copy(), B, Bn, Fn))
        for a, b in zip(*res):
            assert np.array_equal(a, b)


def test_hamming_and_closest_bit_exact(libs):
    orc, ref = libs
    p = make_problem("toy", gaps=True)
    d = p.data
    for q in range(0, p.Q, 7):
        ho, hr = orc.hamming(d["ref_seqs"], d["query_seqs"][q]), ref.hamming(d["ref_seqs"], d["query_seqs"][q])
        assert np.array_equal(ho, hr)
        for n in (1, 5, 20, 21, 199):
            cpu_libs.srandom(11 + q)
            co = orc.closest(ho, n); ro = cpu_libs.libc_random()
            cpu_libs.srandom(11 + q)
            cr = ref.closest(hr, n); rr = cpu_libs.libc_random()
            assert np.array_equal(co, cr)
            assert ro == rr          # same number of random() draws consumed


def test_closest_survey_quirks(libs):
    """the element before the cut joins the tie window of the element after it."""
    orc, ref = libs
    ham = np.array([0, 2, 4, 6, 6, 6], np.int32)
    seen = set()
    for seed in range(40):
        cpu_libs.srandom(seed); a = orc.closest(ham, 3)
        cpu_libs.srandom(seed); b = ref.closest(ham, 3)
        assert np.array_equal(a, b)
        seen.add(int(a[2]))
    assert list(a[:2]) == [0, 1] and seen == {2, 3, 4, 5}


def test_hamming_survey_example(libs):
    orc, ref = libs
    refs = np.frombuffer(b"ACGTACGNNNNNacgt", np.uint8).reshape(4, 4).copy()
    q = np.frombuffer(b"AC-T", np.uint8).copy()
    assert list(orc.hamming(refs, q)) == [0, 2, 6, 0] == list(ref.hamming(refs, q))


def test_mu_rescale_matches_mutation_move(libs):
    """orc_mu_rescale against the rescale loop inside the reference's mutation_move
    (mcmc_moves.c:1082-1111), driven through the whole move with a seeded proposal."""
    import ctypes as C
    import math
    orc, ref = libs
    p = make_problem("toy")
    st = cpu_libs.full_state(ref, p)
    Q, S, L = p.Q, p.S, p.L
    lib = ref.lib
    C.c_double.in_dll(lib, "phi").value = p.phi
    prior_c1 = (C.c_double * 64).in_dll(lib, "prior_C1")
    for i in range(64):
        prior_c1[i] = p.prior[i]
    c2c, A = st["c2c"].copy(), st["A"].copy()
    tc2c, tA = np.zeros_like(c2c), np.zeros_like(A)
    F, tF = st["F"].copy(), np.zeros_like(st["F"])
    Fn, tFn = st["Fn"].copy(), np.zeros_like(st["Fn"])
    B, Bn = st["B"].copy(), st["Bn"].copy()
    mu = C.c_double(p.mu)
    llk = C.c_double(float(st["llk"].sum()))
    R = p.R.copy()

    def pp(a):
        return C.byref(C.c_void_p(a.ctypes.data))

    seed = 77
    cpu_libs.srandom(seed)
    u = cpu_libs.libc_random() / 2147483647.0
    mu_new = p.mu * math.exp((1 * u - 0.5) / 4)
    cpu_libs.srandom(seed)
    holders = [C.c_void_p(a.ctypes.data) for a in (F, Fn, c2c, A, tA, tc2c, tF, tFn)]
    lib.mutation_move(C.c_double(1.0), S, L, p.N, Q, cpu_libs.ip(p.n_site_codons), cpu_libs.cp(p.site_codons),
                      C.byref(mu), cpu_libs.dp(R), C.byref(holders[0]), cpu_libs.dp(B), C.byref(holders[1]),
                      cpu_libs.ip(Bn), C.byref(holders[2]), cpu_libs.cp(p.ref_codons), cpu_libs.cp(p.ref_triplets),
                      cpu_libs.cp(p.query_codons), C.byref(holders[3]), C.byref(holders[4]), C.byref(holders[5]),
                      C.byref(holders[6]), C.byref(llk), cpu_libs.ip(p.closest_refs), C.byref(holders[7]),
                      3 * S, cpu_libs.cp(p.cons_query_nucls))
    # the rescaled tables were written into the arrays that were "tmp" on entry (swapped or not)
    want_c2c, want_A = np.zeros_like(c2c), np.zeros_like(A)
    orc.lib.orc_mu_rescale(S, p.N, Q, cpu_libs.ip(p.n_site_codons), cpu_libs.cp(p.site_codons),
                           cpu_libs.dp(st["c2c"]), cpu_libs.dp(st["A"]), cpu_libs.dp(want_c2c), cpu_libs.dp(want_A),
                           cpu_libs.cp(p.query_codons), C.c_double(p.mu), C.c_double(mu_new), C.c_double(p.phi),
                           cpu_libs.dp(p.prior))
    assert np.array_equal(tc2c, want_c2c) and np.array_equal(tA, want_A)


def test_host_majority_rule_matches_generate_consensus_subset():
    """prep.majority_codes (the host mirror used to build problems) against the reference's
    generate_consensus_subset (amino_acids.c:118-141): ties, unknown bases, empty columns."""
    from mcqueen_b200 import genetics as gen, prep, synth
    d = synth.make_dataset(seed=3, S=30, N=120, Q=4, H=2, ref_unknown_frac=0.2)
    refs = d["ref_seqs"].copy()
    refs[:, 5] = ord("N"); refs[:60, 8] = ord("G"); refs[60:, 8] = ord("C"); refs[3, :9] = np.frombuffer(b"acgtnACGT", np.uint8)
    rng = np.random.default_rng(0)
    for L in (1, 2, 60, 119):
        closest = np.stack([rng.choice(120, L, replace=False) for _ in range(6)]).astype(np.int32)
        want = cpu_libs.reference_consensus_subset(refs, closest)
        codes = gen.base_codes(refs)
        got = np.stack([prep.majority_codes(codes[closest[q]]) for q in range(6)])
        assert np.array_equal(got, want), L
