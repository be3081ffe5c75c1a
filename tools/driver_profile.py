"""Runs the C driver with --profile on a slice of a workload: python tools/driver_profile.py C4 1250 600 [flags...]"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chain_files import make_chain_inputs  # noqa: E402

name, nq, steps = sys.argv[1], int(sys.argv[2]), sys.argv[3]
tmp = tempfile.mkdtemp(prefix="mcq_prof_")
paths, L, _ = make_chain_inputs(os.path.join(tmp, "in"), name, seed=20261019, query_model="dsim", Q=nq)
cmd = [os.path.join(ROOT, "mcqueen_b200", "host", "mcq_dmodel"), "--seed", "1", "--gsl_seed", "1", "-n", steps, "--profile"] + \
    sys.argv[4:] + ["-H", paths["hla"], "-l", str(L), "-s", "1000000", paths["ref"], paths["query"], os.path.join(tmp, "out")]
r = subprocess.run(cmd, capture_output=True, text=True)
print("\n".join(l for l in r.stdout.splitlines() if l.startswith(("profile", "MCMC", "escape", "final", "rec_", "sel_", "merge", "split", "grow", "window", "mut"))))
print(r.stderr[-500:])
