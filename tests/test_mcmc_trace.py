"""The north-star chain test: the C driver on the GPU library (mcqueen_b200/host/mcq_dmodel) and the
UNMODIFIED reference dmodel (oracle/_ref/dmodel_ref, seeded through its link-time wrappers) run the
same fixed-seed chain on the same files; every accept/reject decision, every rand() draw and the
logged state must agree for the first 1000 steps, the log-likelihood to 1e-9 relative."""
import os
import re
import subprocess

import numpy as np
import pytest

import cpu_libs
from chain_files import make_chain_inputs, read_trace

pytestmark = pytest.mark.gpu

ROOT = cpu_libs.ROOT
DRIVER = os.path.join(ROOT, "mcqueen_b200", "host", "mcq_dmodel")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def newest(dirname, suffix):
    fs = sorted(f for f in os.listdir(dirname) if f.endswith(suffix))
    assert fs, (dirname, suffix)
    return os.path.join(dirname, fs[-1])


def log_rows(path):
    rows = [l.split() for l in open(path).read().strip().splitlines()[1:]]
    return np.array(rows, dtype=float)


@pytest.mark.parametrize("extra,gaps,model", [([], False, "simple"), (["--sparse_esc"], False, "simple"),
                                              (["--checkpoint", "1"], True, "dsim"),
                                              ([], True, "simple"), ([], False, "dsim")])
def test_trace_matches_reference_dmodel(tmp_path, extra, gaps, model):
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(DRIVER)])
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy", gaps=gaps, query_model=model)
    args = ["-H", paths["hla"], "-l", str(L), "-n", "1000", "-s", "50", paths["ref"], paths["query"]]
    ours_trace = str(tmp_path / "ours.trace")
    out = run([DRIVER, "--seed", "1", "--gsl_seed", "1", "--trace", ours_trace] + extra + args + [str(tmp_path / "ours")])
    ours = read_trace(ours_trace)
    m = re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", out)
    ours_init = float(m.group(1))

    if cpu_libs.have_ref() and os.path.exists(cpu_libs.REF_DMODEL):
        ref_trace = str(tmp_path / "ref.trace")
        rout = run([cpu_libs.REF_DMODEL] + args + [str(tmp_path / "ref")],
                   env={"MCQ_SEED": "1", "MCQ_GSL_SEED": "1", "MCQ_TRACE": ref_trace})
        ref = read_trace(ref_trace)
        ref_init = float(re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", rout).group(1))
        ref_log = log_rows(newest(str(tmp_path / "ref"), "_log.txt"))
    else:
        if gaps or model != "simple":
            pytest.skip("golden trace covers the gap-free chain on the simple generator only")
        g = np.load(os.path.join(GOLDEN, "toy_chain.npz"))
        ref, ref_init, ref_log = g["trace"], float(g["init_llk"]), g["log"]

    assert abs(ours_init - ref_init) <= 1e-9 * abs(ref_init)
    # identical decisions: the counters before every accept test and the rand() draw itself
    assert ours.shape == ref.shape, (ours.shape, ref.shape)
    assert np.array_equal(ours, ref)
    ours_log = log_rows(newest(str(tmp_path / "ours"), "_log.txt"))
    assert ours_log.shape == ref_log.shape
    # state, mu, omega[], R[] are printed with %f: identical text; the posterior within 1e-9 relative
    assert np.array_equal(ours_log[:, 0], ref_log[:, 0])
    assert np.allclose(ours_log[:, 1], ref_log[:, 1], rtol=1e-9, atol=0)
    assert np.array_equal(ours_log[:, 2:], ref_log[:, 2:])


def test_reference_dmodel_linked_against_the_dropin_symbols(tmp_path):
    """The UNMODIFIED reference dmodel with forward_backward.o / closest_seqs.o replaced by
    libmcq_b200.so (oracle/_ref/dmodel_dropin): same decisions as the all-CPU reference."""
    dropin = os.path.join(ROOT, "oracle", "_ref", "dmodel_dropin")
    if not (os.path.exists(dropin) and os.path.exists(cpu_libs.REF_DMODEL)):
        pytest.skip("oracle/_ref binaries not present")
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy", Q=12, S=60)
    args = ["-H", paths["hla"], "-l", str(L), "-n", "300", "-s", "50", paths["ref"], paths["query"]]
    traces = []
    for exe, name in ((cpu_libs.REF_DMODEL, "ref"), (dropin, "dropin")):
        t = str(tmp_path / f"{name}.trace")
        run([exe] + args + [str(tmp_path / name)], env={"MCQ_SEED": "3", "MCQ_GSL_SEED": "3", "MCQ_TRACE": t})
        traces.append(read_trace(t))
    assert traces[0].shape == traces[1].shape and np.array_equal(traces[0], traces[1])


def test_resume_from_saved_state_matches_reference(tmp_path):
    """dmodel -r: a state written by this driver is read back by it AND by the reference dmodel
    (write_json.c / load_from_json.c format); both resumed chains make the same decisions."""
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(DRIVER)])
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy")
    base = ["-H", paths["hla"], "-l", str(L), "-s", "50", paths["ref"], paths["query"]]
    out = run([DRIVER, "--seed", "5", "--gsl_seed", "5", "-n", "200"] + base + [str(tmp_path / "first")])
    state = newest(str(tmp_path / "first"), "_final_state.json")
    saved = float(re.search(r'"llkhood": (-?[0-9]+\.[0-9]+)', open(state).read()).group(1))
    final = float(re.search(r"final log-likelihood: (-?[0-9]+\.[0-9]+)", out).group(1))
    assert abs(saved - final) <= 1e-9 * abs(final)

    t_ours = str(tmp_path / "ours.trace")
    o = run([DRIVER, "--seed", "9", "--gsl_seed", "9", "--trace", t_ours, "-r", state, "-n", "150"] + base +
            [str(tmp_path / "ours")])
    ours_init = float(re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", o).group(1))
    # omega / R / coefficients are stored with 10 decimals: the recomputed likelihood is close to the saved one
    assert abs(ours_init - saved) <= 1e-6 * abs(saved)
    if not (cpu_libs.have_ref() and os.path.exists(cpu_libs.REF_DMODEL)):
        return
    t_ref = str(tmp_path / "ref.trace")
    r = run([cpu_libs.REF_DMODEL, "-r", state, "-n", "150"] + base + [str(tmp_path / "ref")],
            env={"MCQ_SEED": "9", "MCQ_GSL_SEED": "9", "MCQ_TRACE": t_ref})
    ref_init = float(re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", r).group(1))
    assert abs(ours_init - ref_init) <= 1e-9 * abs(ref_init)
    a, b = read_trace(t_ours), read_trace(t_ref)
    assert a.shape == b.shape and np.array_equal(a, b)


def _trace_pair(tmp_path, extra_args, paths, L, steps="300", seed="4"):
    args = extra_args + ["-H", paths["hla"], "-l", str(L), "-n", steps, "-s", "50", paths["ref"], paths["query"]]
    t_ours, t_ref = str(tmp_path / "ours.trace"), str(tmp_path / "ref.trace")
    o = run([DRIVER, "--seed", seed, "--gsl_seed", seed, "--trace", t_ours] + args + [str(tmp_path / "ours")])
    r = run([cpu_libs.REF_DMODEL] + args + [str(tmp_path / "ref")],
            env={"MCQ_SEED": seed, "MCQ_GSL_SEED": seed, "MCQ_TRACE": t_ref})
    pat = r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)"
    return (read_trace(t_ours), float(re.search(pat, o).group(1)), o), (read_trace(t_ref), float(re.search(pat, r).group(1)), r)


def test_parameters_json_matches_reference(tmp_path):
    """dmodel -p (load_from_json.c:248-295): kappa, phi and the codon priors from a file - same initial
    likelihood and the same decisions as the reference given the same file."""
    import json
    if not (cpu_libs.have_ref() and os.path.exists(cpu_libs.REF_DMODEL)):
        pytest.skip("oracle/_ref/dmodel_ref not present")
    from mcqueen_b200 import genetics as gen
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy")
    rng = np.random.default_rng(2)
    prior = gen.PRIOR_C1 * np.exp(rng.normal(0, 0.3, 64))
    prior /= prior.sum()
    pj = str(tmp_path / "params.json")
    json.dump({"kappa": 5.25, "phi": 0.9, "codons": [{"codon": i, "codon_prior": float(x)} for i, x in enumerate(prior)]},
              open(pj, "w"))
    (a, ia, out), (b, ib, _) = _trace_pair(tmp_path, ["-p", pj], paths, L)
    assert "kappa:5.25" in out and "phi:0.9" in out
    assert abs(ia - ib) <= 1e-9 * abs(ib)
    assert a.shape == b.shape and np.array_equal(a, b)
    # ... and the file changed something: the default parameters give another likelihood
    (_, i0, _), _ = _trace_pair(tmp_path, [], paths, L, steps="1")
    assert abs(i0 - ia) > 1e-6 * abs(ia)


def test_true_mosaic_mode_matches_reference(tmp_path):
    """dmodel -m (dmodel.c:505-558): the closest-L lists are rebuilt from the references each query was truly
    copied from; the permutation and every tie shuffle draw from the shared stream in the reference's order."""
    import json
    if not (cpu_libs.have_ref() and os.path.exists(cpu_libs.REF_DMODEL)):
        pytest.skip("oracle/_ref/dmodel_ref not present")
    paths, L, d = make_chain_inputs(str(tmp_path / "in"), "toy", Q=20)
    sims = []
    for q in range(20):
        seen = []
        for r in d["parents"][q]:
            if int(r) not in seen:
                seen.append(int(r))
        sims.append({"number": q, "closest_n": seen[:L]})
    mj = str(tmp_path / "mosaic.json")
    json.dump({"true_sequences": sims}, open(mj, "w"))
    (a, ia, out), (b, ib, rout) = _trace_pair(tmp_path, ["-m", mj], paths, L)
    assert abs(ia - ib) <= 1e-9 * abs(ib)
    assert a.shape == b.shape and np.array_equal(a, b)
    # the saved closest-L lists are the mosaic ones, identical in both final states
    lists = []
    for name in ("ours", "ref"):
        st = json.load(open(newest(str(tmp_path / name), "_final_state.json")))
        lists.append([qq["closest_n"] for qq in st["query"]])
    assert lists[0] == lists[1]
    for q in range(20):
        assert set(sims[q]["closest_n"]) <= set(lists[0][q]) or len(sims[q]["closest_n"]) > L
