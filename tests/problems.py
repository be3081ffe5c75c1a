"""Seeded problems shared by the tests (TEST code: may use the CPU oracle)."""
from __future__ import annotations

import functools

import numpy as np

import cpu_libs
from mcqueen_b200 import prep, synth


@functools.lru_cache(maxsize=None)
def make_problem(name="toy", seed=20261019, gaps=False, param_seed=1, **over):
    cfg = dict(synth.CONFIGS[name])
    cfg.update(over)
    L = cfg.pop("L")
    data = synth.make_dataset(seed=seed, ref_unknown_frac=0.05 if gaps else 0.0, query_gap=gaps, **cfg)
    orc = cpu_libs.Checker("oracle")
    cpu_libs.srandom(seed & 0x7fffffff)
    closest = np.zeros((cfg["Q"], L), np.int32)
    for q in range(cfg["Q"]):
        ham = orc.hamming(data["ref_seqs"], data["query_seqs"][q])
        closest[q] = orc.closest(ham, L)
    p = prep.build_problem(data["ref_seqs"], data["query_seqs"], data["hla"], closest)
    prep.random_parameters(p, seed=param_seed)
    p.data = data
    return p
