"""TEST helper: the chain shapes of tests/test_mcmc_trace_shapes.py, shared with the script that records
the reference's traces for them (tests/golden/make_trace_golden.py)."""
from __future__ import annotations

# name -> (synth.make_dataset overrides on top of CONFIGS["toy"], chain steps, extra dmodel flags)
# c2slice : a slice of BASELINE config 2 - L=100 -> Ls=112, Lg=128: the two-word (NJ=2), ragged-row kernels
# l500    : the closest-L of config 4 - L=500 -> Ls=512=Lg: NJ=8, full rows
# c5slice : the shape of config 5 - 3000 codons, L=200 -> Ls=208, Lg=256 (NJ=4), self-reference mode (-q)
SHAPES = {
    "c2slice": (dict(S=500, N=2000, Q=64, L=100, H=20), 1000, []),
    "l500": (dict(S=500, N=2000, Q=24, L=500, H=8), 300, []),
    "c5slice": (dict(S=3000, N=600, Q=16, L=200, H=8, self_reference=True), 300, ["-q"]),
}
SEED = 11


def chain_args(paths, L, steps, flags):
    return flags + ["-H", paths["hla"], "-l", str(L), "-n", str(steps), "-s", "50", paths["ref"], paths["query"]]
