"""The 1000-step trace identity at the shapes the benchmarks run (tests/test_mcmc_trace.py covers the toy shape,
i.e. only the one-word kernels): a slice of config 2 (L=100: NJ=2, ragged rows), L=500 (NJ=8, full rows) and a
config-5-shaped slice (3000 codons, L=200: NJ=4, self-reference mode).  The C driver on the GPU library must make
every accept/reject decision (mcmc_moves.c:249-251, 437, 941, 1153) and every rand() draw of the UNMODIFIED
reference dmodel; the reference's traces were recorded in the build container by
tests/golden/make_trace_golden.py (the reference runs live instead when oracle/_ref/dmodel_ref is present and
the fixture is not).  The driver keeps every fourth column of F and B by default (--checkpoint); whole matrices and
other strides are run as well."""
import os
import re
import subprocess

import numpy as np
import pytest

import cpu_libs
from chain_files import make_chain_inputs, read_trace
from trace_shapes import SEED, SHAPES, chain_args

pytestmark = pytest.mark.gpu

ROOT = cpu_libs.ROOT
DRIVER = os.path.join(ROOT, "mcqueen_b200", "host", "mcq_dmodel")
GOLDEN = os.path.join(ROOT, "tests", "golden")
INIT_RE = r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)"


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def log_rows(dirname):
    fs = sorted(f for f in os.listdir(dirname) if f.endswith("_log.txt"))
    rows = [l.split() for l in open(os.path.join(dirname, fs[-1])).read().strip().splitlines()[1:]]
    return np.array(rows, dtype=float)


def reference_chain(name, tmp_path, paths, L):
    over, steps, flags = SHAPES[name]
    fixture = os.path.join(GOLDEN, f"trace_{name}.npz")
    if os.path.exists(fixture):
        g = np.load(fixture)
        assert int(g["seed"]) == SEED and int(g["steps"]) == steps
        return g["trace"], float(g["init_llk"]), g["log"]
    if not os.path.exists(cpu_libs.REF_DMODEL):
        pytest.skip("neither the recorded trace nor oracle/_ref/dmodel_ref is present")
    t = str(tmp_path / "ref.trace")
    out = run([cpu_libs.REF_DMODEL] + chain_args(paths, L, steps, flags) + [str(tmp_path / "ref")],
              env={"MCQ_SEED": str(SEED), "MCQ_GSL_SEED": str(SEED), "MCQ_TRACE": t})
    return read_trace(t), float(re.search(INIT_RE, out).group(1)), log_rows(str(tmp_path / "ref"))


@pytest.mark.parametrize("extra", [[], ["--sparse_esc"], ["--eager_refill"], ["--checkpoint", "1"], ["--checkpoint", "1", "--sparse_esc"],
                          ["--checkpoint", "8", "--sparse_esc"], ["--checkpoint", "2", "--eager_refill"]],
                         ids=["default", "sparse_esc", "eager", "whole", "whole_sparse", "checkpoint8_sparse", "checkpoint2_eager"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_trace_matches_reference_at_benchmark_shapes(tmp_path, name, extra):
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(DRIVER)])
    over, steps, flags = SHAPES[name]
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy", query_model="dsim", **over)
    ref, ref_init, ref_log = reference_chain(name, tmp_path, paths, L)
    t = str(tmp_path / "ours.trace")
    out = run([DRIVER, "--seed", str(SEED), "--gsl_seed", str(SEED), "--trace", t] + extra +
              chain_args(paths, L, steps, flags) + [str(tmp_path / "ours")])
    ours = read_trace(t)
    ours_init = float(re.search(INIT_RE, out).group(1))
    assert abs(ours_init - ref_init) <= 1e-9 * abs(ref_init)
    assert ours.shape == ref.shape, (ours.shape, ref.shape)
    bad = np.nonzero((ours != ref).any(axis=1))[0]
    assert bad.size == 0, f"first differing accept test: {bad[0]}: ours {ours[bad[0]]} reference {ref[bad[0]]}"
    ours_log = log_rows(str(tmp_path / "ours"))
    assert ours_log.shape == ref_log.shape
    assert np.array_equal(ours_log[:, 0], ref_log[:, 0])
    assert np.allclose(ours_log[:, 1], ref_log[:, 1], rtol=1e-9, atol=0)
    assert np.array_equal(ours_log[:, 2:], ref_log[:, 2:])
