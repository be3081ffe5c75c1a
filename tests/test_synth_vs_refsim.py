"""The data generator against the reference's own simulator (simulate_sequences_sim.c, compiled unmodified into
oracle/_ref/libmcq_refsim.so): the emission distribution synth.dsim_rows computes is the reference's
simulate_codon_to_codon_HLA, and synth.dsim_simulate_in_reference_order reproduces simulate_sequences draw for
draw from the same libc seed - HLA profiles, copied references and simulated codons identical.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import cpu_libs
from mcqueen_b200 import genetics as gen, prep, synth

REFSIM = os.path.join(cpu_libs.ORACLE_DIR, "_ref", "libmcq_refsim.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REFSIM), reason="oracle/_ref/libmcq_refsim.so not built")


def _problem(S=40, N=60, seed=3):
    d = synth.make_dataset(seed=seed, S=S, N=N, Q=4, H=6)
    closest = np.tile(np.arange(8, dtype=np.int32), (4, 1))
    p = prep.build_problem(d["ref_seqs"], d["query_seqs"], d["hla"], closest)
    prep.random_parameters(p, seed=seed)
    return p


def _lib(p):
    lib = C.CDLL(REFSIM)
    C.c_double.in_dll(lib, "kappa").value = p.kappa
    C.c_double.in_dll(lib, "phi").value = p.phi
    prior = (C.c_double * 64).in_dll(lib, "prior_C1")
    for i in range(64):
        prior[i] = p.prior[i]
    return lib


def test_dsim_rows_match_simulate_codon_to_codon_HLA():
    p = _problem()
    lib = _lib(p)
    S, N, H, nsim = p.S, p.N, p.H, 9
    rng = np.random.default_rng(1)
    hla = np.where(rng.random((nsim, H)) < 0.3, ord("1"), ord("0")).astype(np.uint8)
    copied = p.ref_codons[rng.integers(0, N, (nsim, S)), np.arange(S)[None, :]].astype(np.int32)
    copied[0] = p.consensus_codons            # a query that copies the consensus everywhere
    out = np.zeros((nsim, S, 64))
    A = np.zeros((nsim, 64, S))
    rate_out = p.rate_out_of_codons()         # [64][S]
    lib.simulate_codon_to_codon_HLA(S, N, nsim, H, cpu_libs.dp(out), cpu_libs.dp(rate_out), 0, S - 1, C.c_double(p.mu),
                                    cpu_libs.dp(np.ascontiguousarray(p.esc)), cpu_libs.dp(np.ascontiguousarray(p.rev[0])),
                                    cpu_libs.cp(hla), cpu_libs.cp(p.consensus_codons), cpu_libs.dp(p.omega),
                                    cpu_libs.ip(p.ns_ts_sum), cpu_libs.ip(p.ns_tv_sum), cpu_libs.cp(p.type_of_change),
                                    cpu_libs.ip(copied), cpu_libs.dp(A))
    esc_prod = np.exp((hla == ord("1")).astype(float) @ np.log(p.esc))
    rows = synth.dsim_rows(copied.astype(np.int64), p.consensus_codons[None, :], p.omega[None, :], esc_prod,
                           p.rev[0][None, :], p.mu, N, kappa=p.kappa, phi=p.phi, prior=p.prior)
    assert np.allclose(rows, out, rtol=1e-12, atol=1e-300)
    assert np.allclose(rows.sum(-1), 1.0, atol=1e-9)


def test_simulation_reproduces_the_reference_draw_for_draw(tmp_path):
    p = _problem(S=30, N=50, seed=5)
    lib = _lib(p)
    S, N, nsim = p.S, p.N, 12
    ploidy, n_HLA = 2, np.array([2, 4], np.int32)          # 6 HLA types in two genes
    H = int(n_HLA.sum())
    prev = np.array([0.6, 0.4, 0.1, 0.2, 0.3, 0.4])
    R = np.full(S, 0.15)
    rng = np.random.default_rng(2)
    esc = np.exp(rng.normal(0, 0.5, (H, S)))
    rev = np.exp(rng.normal(0, 0.5, S))
    rate_out = p.rate_out_of_codons()
    libc = C.CDLL("libc.so.6")
    libc.fopen.restype = C.c_void_p
    files = {}
    for name in ("hla_file", "mosaic_json"):
        files[name] = str(tmp_path / name)
        C.c_void_p.in_dll(lib, name).value = libc.fopen(files[name].encode(), b"w")
    sim_c2c = np.zeros((nsim, S, 64)); A = np.zeros((nsim, 64, S))
    ccf = np.zeros((nsim, S), np.int32); seqs = np.zeros((nsim, S), np.int32); copy_from = np.zeros((nsim, S), np.int32)
    cpu_libs.srandom(77)
    lib.simulate_sequences(S, N, nsim, H, cpu_libs.dp(prev), ploidy, 2, cpu_libs.ip(n_HLA), cpu_libs.cp(p.ref_codons),
                           C.c_double(p.mu), cpu_libs.dp(esc), cpu_libs.dp(rev), cpu_libs.cp(p.consensus_codons),
                           cpu_libs.dp(p.omega), cpu_libs.dp(R), cpu_libs.dp(rate_out), cpu_libs.ip(p.ns_ts_sum),
                           cpu_libs.ip(p.ns_tv_sum), cpu_libs.cp(p.type_of_change), cpu_libs.dp(sim_c2c),
                           C.c_double(0.25), cpu_libs.ip(ccf), cpu_libs.ip(seqs), cpu_libs.dp(A), cpu_libs.ip(copy_from))
    after_ref = cpu_libs.libc_random()
    for name in files:
        libc.fclose(C.c_void_p(C.c_void_p.in_dll(lib, name).value))
    cpu_libs.srandom(77)
    hla, cf, sq = synth.dsim_simulate_in_reference_order(
        p.ref_codons, p.consensus_codons, p.omega, R, esc, rev, p.mu, prev, ploidy, n_HLA, 0.25, nsim,
        uniform=lambda: cpu_libs.libc_random() / 2147483647.0, kappa=p.kappa, phi=p.phi, prior=p.prior)
    assert cpu_libs.libc_random() == after_ref                  # the same number of draws
    assert np.array_equal(cf, copy_from) and np.array_equal(sq, seqs)
    rows = [l.strip().split(",")[1:] for l in open(files["hla_file"]).read().strip().splitlines()[1:]]
    assert np.array_equal(np.array(rows), np.vectorize(chr)(hla))
    assert (hla == ord("1")).sum(1).max() <= ploidy * len(n_HLA)


def test_simulation_parameter_json_is_read_by_the_reference(tmp_path):
    """The dsim parameter file written by synth is loaded by the reference's own readers
    (load_from_json.c:297-443) into the same values, and read back by synth's loader."""
    lib = C.CDLL(REFSIM)
    S, n_HLA, ploidy = 17, [2, 3], 2
    H = sum(n_HLA)
    rng = np.random.default_rng(4)
    omega, R, rev = np.exp(rng.normal(0, 0.5, S)), rng.uniform(0, 0.1, S), np.exp(rng.normal(0, 0.3, S))
    prof, prev = np.exp(rng.normal(0, 0.5, (H, S))), rng.uniform(0.05, 0.5, H)
    path = str(tmp_path / "sim_params.json")
    synth.write_simulation_parameters_json(path, 6.5, 18.0, omega, R, rev, prof, prev, n_HLA, ploidy)
    kappa, mu = C.c_double(), C.c_double()
    ints = [C.c_int() for _ in range(4)]
    lib.load_lengths_for_simulation_from_json(path.encode(), C.byref(kappa), C.byref(mu), *[C.byref(i) for i in ints])
    assert (kappa.value, mu.value) == (6.5, 18.0)
    assert [i.value for i in ints] == [S, H, ploidy, len(n_HLA)]
    o, r, v = np.zeros(S), np.zeros(S), np.zeros(S)
    nh, pv, pf = np.zeros(len(n_HLA), np.int32), np.zeros(H), np.zeros((H, S))
    lib.load_parameters_for_simulation_from_json(path.encode(), S, cpu_libs.dp(o), cpu_libs.dp(r), cpu_libs.dp(v), H,
                                                 len(n_HLA), cpu_libs.ip(nh), cpu_libs.dp(pv), cpu_libs.dp(pf))
    # (the reference's JSON library converts decimals with pow(10, -digits): the last bit may differ from the
    # correctly rounded value json.dump wrote)
    close = lambda a, b: np.allclose(a, b, rtol=1e-15, atol=0)
    assert close(o, omega) and close(r, R) and close(v, rev)
    assert nh.tolist() == n_HLA and close(pv, prev) and close(pf, prof)
    back = synth.load_simulation_parameters_json(path)
    assert np.array_equal(back["omega"], omega) and np.array_equal(back["hla_profiles"], prof)
    assert back["n_HLA"] == n_HLA and back["ploidy"] == ploidy and back["kappa"] == 6.5
