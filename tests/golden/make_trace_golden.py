#!/usr/bin/env python
"""Records the accept/reject traces of the UNMODIFIED reference dmodel (oracle/_ref/dmodel_ref, built by
oracle/Makefile from /root/reference) for the chain shapes of tests/trace_shapes.py.  Run in the build
container:

    python tests/golden/make_trace_golden.py [shape ...]

Output per shape (small, committed): tests/golden/trace_<shape>.npz with the accept-test trace (14 counters +
the rand() draw per test), the initial log-likelihood and the sampled log rows.  The GPU box has no
/root/reference and the reference needs ~1 min of one core per shape: the GPU tests compare with these."""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_libs  # noqa: E402
from chain_files import make_chain_inputs, read_trace  # noqa: E402
from trace_shapes import SEED, SHAPES, chain_args  # noqa: E402


def record(name):
    over, steps, flags = SHAPES[name]
    with tempfile.TemporaryDirectory() as tmp:
        paths, L, _ = make_chain_inputs(os.path.join(tmp, "in"), "toy", query_model="dsim", **over)
        trace = os.path.join(tmp, "ref.trace")
        env = dict(os.environ, MCQ_SEED=str(SEED), MCQ_GSL_SEED=str(SEED), MCQ_TRACE=trace)
        t0 = time.time()
        r = subprocess.run([cpu_libs.REF_DMODEL] + chain_args(paths, L, steps, flags) + [os.path.join(tmp, "ref")],
                           capture_output=True, text=True, env=env, check=True)
        dt = time.time() - t0
        init = float(re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", r.stdout).group(1))
        logs = sorted(f for f in os.listdir(os.path.join(tmp, "ref")) if f.endswith("_log.txt"))
        rows = [l.split() for l in open(os.path.join(tmp, "ref", logs[-1])).read().strip().splitlines()[1:]]
        np.savez_compressed(os.path.join(HERE, f"trace_{name}.npz"), trace=read_trace(trace), init_llk=init,
                            log=np.array(rows, dtype=float), seed=SEED, steps=steps)
        print(f"{name}: {steps} steps, reference took {dt:.1f} s, initial llk {init}")


if __name__ == "__main__":
    assert os.path.exists(cpu_libs.REF_DMODEL), "build oracle/_ref first (make -C oracle ref)"
    for n in (sys.argv[1:] or list(SHAPES)):
        record(n)
