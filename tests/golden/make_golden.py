#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE ITSELF (oracle/_ref, i.e. the
unmodified translation units of /root/reference compiled by oracle/Makefile).  Run in the build
container (where /root/reference exists):

    python tests/golden/make_golden.py

Outputs (small, committed):
  toy_state.npz  - seeded toy problem: per-query log-likelihoods, rescale counters, sampled F/B columns,
                   sampled emission rows, Hamming distances and closest-L lists (with the seed used)
  toy_chain.npz  - the 1000-step fixed-seed dmodel chain: accept-test trace (counters + rand() draws),
                   initial log-likelihood and the sampled log rows
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_libs  # noqa: E402
from chain_files import make_chain_inputs, read_trace  # noqa: E402
from problems import make_problem  # noqa: E402


def main():
    assert cpu_libs.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    ref = cpu_libs.Checker("ref")

    # ---- state fixture
    p = make_problem("toy", gaps=True)
    st = cpu_libs.full_state(ref, p)
    d = p.data
    qs = [0, 7, 23, 49]
    sites = [0, 1, 50, 98, 99]
    ham = np.stack([ref.hamming(d["ref_seqs"], d["query_seqs"][q]) for q in range(p.Q)])
    closest = {}
    for n in (1, 20, 21):
        cpu_libs.srandom(4242)
        closest[n] = np.stack([ref.closest(ham[q], n) for q in range(p.Q)])
    np.savez_compressed(
        os.path.join(HERE, "toy_state.npz"),
        llk=st["llk"], Fn=st["Fn"], Bn=st["Bn"], queries=np.array(qs), sites=np.array(sites),
        F_cols=st["F"][qs][:, :, sites], B_cols=st["B"][qs][:, :, sites],
        c2c=st["c2c"][qs], A=st["A"][qs], hamming=ham.astype(np.int32),
        closest_1=closest[1], closest_20=closest[20], closest_21=closest[21], closest_seed=4242,
        problem_closest=p.closest_refs)

    # ---- chain fixture
    tmp = tempfile.mkdtemp()
    paths, L, _ = make_chain_inputs(os.path.join(tmp, "in"), "toy")
    env = dict(os.environ, MCQ_SEED="1", MCQ_GSL_SEED="1", MCQ_TRACE=os.path.join(tmp, "ref.trace"))
    out = subprocess.run([cpu_libs.REF_DMODEL, "-H", paths["hla"], "-l", str(L), "-n", "1000", "-s", "50",
                          paths["ref"], paths["query"], os.path.join(tmp, "ref")],
                         capture_output=True, text=True, env=env, check=True).stdout
    init = float(re.search(r"Total log-likelihood using initialise F: (-?[0-9]+\.[0-9]+)", out).group(1))
    logf = [f for f in os.listdir(os.path.join(tmp, "ref")) if f.endswith("_log.txt")][0]
    rows = [l.split() for l in open(os.path.join(tmp, "ref", logf)).read().strip().splitlines()[1:]]
    counts = [int(x) for x in re.findall(r"(?:accept|reject): (\d+)", out)]
    np.savez_compressed(os.path.join(HERE, "toy_chain.npz"), trace=read_trace(env["MCQ_TRACE"]), init_llk=init,
                        log=np.array(rows, dtype=float), final_counts=np.array(counts))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
