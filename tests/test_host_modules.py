"""CPU checks of the host-side C modules of the driver (mcqueen_b200/host): the window bookkeeping and the
random streams against the reference's own selection_window.c (compiled unmodified into
oracle/_ref/libmcq_ref.so, GSL replaced by the shim), and the state JSON against itself.

What the GPU trace test proves end to end - the same decisions, draw for draw, as the reference dmodel -
is pinned here for the piece that does not need a GPU: for the same seeds the same script of window
proposals (split / merge / grow / coefficient, with undo and accept) leaves the same window sets and
consumes the same number of libc random() draws (selection_window.c:117-263)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cpu_libs

ROOT = cpu_libs.ROOT
HOST = os.path.join(ROOT, "mcqueen_b200", "host")
GEOM_P = 0.5            # constants.c:75
LOG_MU, SIGMA = 1.0, 2.0   # constants.c:48-52 (as passed by dmodel.c to selec_windows_randomise)


class Window(C.Structure):
    _fields_ = [("start", C.c_int), ("length", C.c_int), ("coeff", C.c_double)]


class WindowSet(C.Structure):
    _fields_ = [("w", C.POINTER(Window)), ("count", C.c_int)]


class HlaWindows(C.Structure):
    _fields_ = [("esc", C.POINTER(WindowSet)), ("rev", C.POINTER(WindowSet))]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostlib") / "libmcqhost.so")
    srcs = [os.path.join(HOST, f) for f in ("mcq_windows.c", "mcq_rng.c", "mcq_io.c", "mcq_json.c")]
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-fPIC", "-shared", "-ffp-contract=off", "-I",
                           os.path.join(ROOT, "include"), "-o", out] + srcs + ["-lz", "-lm"])
    lib = C.CDLL(out)
    lib.mcq_ran_lognormal.restype = C.c_double
    lib.mcq_ran_lognormal.argtypes = [C.c_double, C.c_double]
    lib.mcq_ran_geometric.restype = C.c_uint
    lib.mcq_ran_geometric.argtypes = [C.c_double]
    lib.mcq_rand_lim.restype = C.c_double
    lib.mcq_mt_seed.argtypes = [C.c_ulong]
    lib.mcq_windows_randomise.argtypes = [C.POINTER(WindowSet), C.c_int, C.c_int, C.c_double, C.c_double]
    lib.mcq_windows_grow.argtypes = [C.POINTER(WindowSet), C.c_double, C.POINTER(C.c_int)]
    return lib


@pytest.fixture(scope="module")
def ref():
    if not cpu_libs.have_ref():
        pytest.skip("oracle/_ref/libmcq_ref.so not built (reference sources absent)")
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmcq_ref.so"))
    lib.gsl_ran_lognormal.restype = C.c_double
    lib.gsl_ran_lognormal.argtypes = [C.c_void_p, C.c_double, C.c_double]
    lib.gsl_ran_geometric.restype = C.c_uint
    lib.gsl_ran_geometric.argtypes = [C.c_void_p, C.c_double]
    lib.gsl_rng_alloc.restype = C.c_void_p
    lib.gsl_rng_set.argtypes = [C.c_void_p, C.c_ulong]
    lib.selec_windows_randomise.argtypes = [C.POINTER(WindowSet), C.c_int, C.c_int, C.c_double, C.c_double]
    return lib


def snapshot(ws, S):
    """Live windows plus the pool of unused start sites behind them."""
    return ([(ws.w[i].start, ws.w[i].length, ws.w[i].coeff) for i in range(ws.count)],
            sorted(ws.w[i].start for i in range(ws.count, S)))


def test_mersenne_twister_and_samplers_match_the_shim(host, ref):
    """mcq_rng.c and oracle/shim/gsl_shim.c are two independent writings of MT19937 and of the two GSL samplers
    selection_window.c uses: same seed, same draws (geometric exactly, log-normal to the last bit)."""
    os.environ.pop("MCQ_GSL_SEED", None)
    r = ref.gsl_rng_alloc(None)
    for seed in (1, 5489, 20261019):
        ref.gsl_rng_set(r, seed)
        host.mcq_mt_seed(seed)
        for i in range(2000):
            if i % 3 == 0:
                assert host.mcq_ran_geometric(GEOM_P) == ref.gsl_ran_geometric(r, GEOM_P)
            else:
                assert host.mcq_ran_lognormal(LOG_MU, SIGMA) == ref.gsl_ran_lognormal(r, LOG_MU, SIGMA)


def run_script(lib, names, ws_ptr, S, script_seed, steps):
    """The same script of proposals through either implementation; returns the trace of snapshots."""
    rng = np.random.default_rng(script_seed)
    ws = ws_ptr.contents
    trace = []
    old_c, old0, old1 = C.c_double(), C.c_double(), C.c_double()
    old_s = C.c_int()
    for _ in range(steps):
        action, keep = int(rng.integers(0, 4)), bool(rng.random() < 0.6)
        if action == 0:
            w = names["split"](ws_ptr, S, C.byref(old_c))
            if w >= 0 and not keep:
                names["split_undo"](ws_ptr, w, old_c)
        elif action == 1:
            if ws.count < 2:
                continue
            w = names["merge"](ws_ptr, C.byref(old_s), C.byref(old0), C.byref(old1))
            if w >= 0 and not keep:
                names["merge_undo"](ws_ptr, w, old_s, old0, old1)
        elif action == 2:
            w = names["grow"](ws_ptr, C.byref(old_s))
            if w >= 0:
                if keep:
                    names["grow_accept"](ws_ptr, ws.w[w + 1].start, old_s, S)
                else:
                    names["grow_undo"](ws_ptr, w, old_s)
        else:
            w = names["coeff"](ws_ptr, C.byref(old_c))
            if w >= 0 and not keep:
                names["coeff_undo"](ws_ptr, w, old_c)
        trace.append((action, keep, w, snapshot(ws, S), cpu_libs._libc.random()))   # the libc stream position
    return trace


@pytest.mark.parametrize("S,init,seed", [(60, 6, 3), (200, 6, 11), (25, 3, 5)])
def test_window_proposals_follow_the_reference(host, ref, S, init, seed):
    os.environ["MCQ_GSL_SEED"] = str(seed)       # the shim's override: setup_gsl() cannot be given a seed
    ref.setup_gsl()
    # --- the reference
    rh = HlaWindows()
    ref.selec_window_hla_alloc(S, 1, C.byref(rh))
    cpu_libs.srandom(seed)
    ref.selec_windows_randomise(rh.esc, S, init, LOG_MU, SIGMA)
    ref_names = dict(split=ref.selec_window_split_rnd, split_undo=ref.selec_window_split_undo,
                     merge=ref.selec_window_merge_rnd, merge_undo=ref.selec_window_merge_undo,
                     grow=ref.selec_window_grow_rnd, grow_accept=ref.selec_window_grow_accept,
                     grow_undo=ref.selec_window_grow_undo, coeff=ref.selec_window_change_coeff_rnd,
                     coeff_undo=ref.selec_window_change_coeff_undo)
    for f in ("split_undo", "coeff_undo"):
        ref_names[f].argtypes = [C.POINTER(WindowSet), C.c_int, C.c_double]
    ref_names["merge_undo"].argtypes = [C.POINTER(WindowSet), C.c_int, C.c_int, C.c_double, C.c_double]
    want0 = snapshot(rh.esc.contents, S)
    want = run_script(ref, ref_names, rh.esc, S, seed, 400)
    # --- the driver's module, same seeds
    hh = HlaWindows()
    host.mcq_windows_alloc(S, 1, init, init, C.byref(hh))
    host.mcq_mt_seed(seed)
    cpu_libs.srandom(seed)
    host.mcq_windows_randomise(hh.esc, S, init, LOG_MU, SIGMA)
    host_names = dict(split=host.mcq_windows_split, split_undo=host.mcq_windows_split_undo,
                      merge=host.mcq_windows_merge, merge_undo=host.mcq_windows_merge_undo,
                      grow=lambda ws, old: host.mcq_windows_grow(ws, GEOM_P, old),
                      grow_accept=host.mcq_windows_grow_accept, grow_undo=host.mcq_windows_grow_undo,
                      coeff=host.mcq_windows_coeff, coeff_undo=host.mcq_windows_coeff_undo)
    for f in ("split_undo", "coeff_undo"):
        host_names[f].argtypes = [C.POINTER(WindowSet), C.c_int, C.c_double]
    host_names["merge_undo"].argtypes = [C.POINTER(WindowSet), C.c_int, C.c_int, C.c_double, C.c_double]
    assert snapshot(hh.esc.contents, S) == want0
    got = run_script(host, host_names, hh.esc, S, seed, 400)
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, (i, a[:3], b[:3])
    # partition invariant at the end
    wins = snapshot(hh.esc.contents, S)[0]
    assert wins[0][0] == 0 and sum(w[1] for w in wins) == S
    assert all(wins[i][0] + wins[i][1] == wins[i + 1][0] for i in range(len(wins) - 1))
    os.environ.pop("MCQ_GSL_SEED", None)


def test_state_json_round_trip(host, tmp_path):
    S, H, Q, L = 40, 3, 5, 4
    hw = (HlaWindows * H)()
    host.mcq_mt_seed(9)
    cpu_libs.srandom(9)
    for h in range(H):
        host.mcq_windows_alloc(S, 1, 4, 4, C.byref(hw[h]))
        host.mcq_windows_randomise(hw[h].esc, S, 4, LOG_MU, SIGMA)
        host.mcq_windows_randomise(hw[h].rev, S, 3, LOG_MU, SIGMA)
    rng = np.random.default_rng(0)
    omega, R = rng.random(S) + 0.5, rng.random(S) * 0.02
    closest = rng.integers(0, 100, (Q, L)).astype(np.int32)
    path = str(tmp_path / "state.json")
    libc = C.CDLL("libc.so.6")
    libc.fopen.restype = C.c_void_p
    fh = libc.fopen(path.encode(), b"w")
    host.mcq_write_state_json.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    host.mcq_write_state_json(fh, 7, -1234.5, 19.25, S, omega.ctypes.data, R.ctypes.data, hw, H, Q, L, closest.ctypes.data)
    libc.fclose.argtypes = [C.c_void_p]
    libc.fclose(fh)
    back = (HlaWindows * H)()
    for h in range(H):
        host.mcq_windows_alloc(S, 1, 1, 1, C.byref(back[h]))
    llk, mu = C.c_double(), C.c_double()
    omega2, R2 = np.zeros(S), np.zeros(S)
    closest2 = np.zeros((Q, L), np.int32)
    host.mcq_load_state_json.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    host.mcq_load_state_json(path.encode(), C.byref(llk), C.byref(mu), R2.ctypes.data, omega2.ctypes.data, back, S, H, Q, L,
                             closest2.ctypes.data)
    assert mu.value == pytest.approx(19.25, rel=1e-6) and llk.value == pytest.approx(-1234.5, rel=1e-6)
    assert np.allclose(omega2, omega, rtol=1e-5) and np.allclose(R2[1:], R[1:], rtol=1e-5, atol=1e-9)
    assert np.array_equal(closest2, closest)
    for h in range(H):
        for a, b in ((hw[h].esc, back[h].esc), (hw[h].rev if h == 0 else None, back[h].rev if h == 0 else None)):
            if a is None:
                continue
            wa, wb = snapshot(a.contents, S)[0], snapshot(b.contents, S)[0]
            assert [(s, n) for s, n, _ in wa] == [(s, n) for s, n, _ in wb]
            assert np.allclose([c for _, _, c in wa], [c for _, _, c in wb], rtol=1e-5)


def test_fasta_and_hla_csv_readers(host, tmp_path):
    """load_data.c:66-157 restated: multi-line records, CR/LF, blank lines; the 0/1 table with a header line and one
    leading label column."""
    fa = tmp_path / "a.fasta"
    fa.write_text(">r1 some text\nTTTCAG\nGG-\n\n>r2\r\nacgtnn\r\n>r3\n>r4\nA\n")
    seqs = C.POINTER(C.c_char_p)()
    host.mcq_load_fasta.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_char_p))]
    n = host.mcq_load_fasta(str(fa).encode(), C.byref(seqs))
    assert n == 4
    assert [seqs[i] for i in range(n)] == [b"TTTCAGGG-", b"acgtnn", b"", b"A"]
    csv = tmp_path / "h.csv"
    csv.write_text('"","A01","A02","B07"\nq1,0,1,0\nq2, 1 ,1,1\nq3,0,0,0\nq4,1,0,1\n')
    rows = C.POINTER(C.c_char_p)()
    host.mcq_load_hla_csv.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_char_p)), C.c_int]
    types = host.mcq_load_hla_csv(str(csv).encode(), C.byref(rows), 3)
    assert types == 3
    assert [rows[i] for i in range(3)] == [b"010", b"111", b"000"]
    d = tmp_path / "x" / "y" / "z"
    host.mcq_mkpath.argtypes = [C.c_char_p]
    assert host.mcq_mkpath(str(d).encode()) == 1 and d.is_dir()


def test_parameter_json_reads_the_doubles_the_reference_reads(host, tmp_path):
    """dmodel -p: kappa, phi and the codon priors come out of the file as exactly the doubles the reference's JSON
    library produces (digits times pow(10, exponent), cJSON.c:96-115 - not the correctly rounded strtod value), for
    plain decimals, long fractions and exponents alike (load_from_json.c:248-295)."""
    refsim = os.path.join(ROOT, "oracle", "_ref", "libmcq_refsim.so")
    if not os.path.exists(refsim):
        pytest.skip("oracle/_ref/libmcq_refsim.so not built")
    ref = C.CDLL(refsim)
    rng = np.random.default_rng(9)
    texts = ["0.1", "7.595744", "1e-05", "3.08971e-06", "0.9482171", "12345.678901234567", "2.5E+3", "0.000117409", "-0.5"]
    texts += [repr(float(x)) for x in rng.random(64 - len(texts)) * 10.0 ** rng.integers(-8, 3, 64 - len(texts))]
    path = str(tmp_path / "p.json")
    with open(path, "w") as f:
        f.write('{"kappa": 7.595744, "phi": 0.9482171,\n "codons": [' +
                ", ".join('{"codon": %d, "codon_prior": %s}' % (i, t) for i, t in enumerate(texts)) + "]}\n")
    ours, theirs = np.zeros(64), np.zeros(64)
    k = [C.c_double(), C.c_double()]
    p = [C.c_double(), C.c_double()]
    host.mcq_load_fixed_parameters_json(path.encode(), C.byref(k[0]), C.byref(p[0]), ours.ctypes.data_as(C.c_void_p))
    ref.load_fixed_parameters_from_json(path.encode(), C.byref(k[1]), C.byref(p[1]), theirs.ctypes.data_as(C.c_void_p))
    assert k[0].value == k[1].value and p[0].value == p[1].value
    assert np.array_equal(ours, theirs)
    assert any(float(t) != v for t, v in zip(texts, theirs))      # the point: these are not all strtod's values
