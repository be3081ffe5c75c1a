"""TEST helper: writes a seeded toy data set as the files dmodel reads (FASTA + HLA csv)."""
from __future__ import annotations

import os

import numpy as np

from mcqueen_b200 import synth


def write_fasta(path, seqs, prefix):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">{prefix}{i}\n{bytes(s).decode()}\n")


def write_hla_csv(path, hla):
    H = hla.shape[1]
    with open(path, "w") as f:
        f.write("id," + ",".join(f"HLA{h}" for h in range(H)) + "\n")
        for q, row in enumerate(hla):
            f.write(f"q{q}," + ",".join(chr(c) for c in row) + "\n")


def make_chain_inputs(dirname, name="toy", seed=20261019, gaps=False, query_model="simple", **over):
    cfg = dict(synth.CONFIGS[name])
    cfg.update(over)
    L = cfg.pop("L")
    d = synth.make_dataset(seed=seed, ref_unknown_frac=0.03 if gaps else 0.0, query_gap=gaps,
                           query_model=query_model, **cfg)
    os.makedirs(dirname, exist_ok=True)
    paths = dict(ref=os.path.join(dirname, "ref.fasta"), query=os.path.join(dirname, "query.fasta"),
                 hla=os.path.join(dirname, "query.HLA.csv"))
    write_fasta(paths["ref"], d["ref_seqs"], "r")
    write_fasta(paths["query"], d["query_seqs"], "q")
    write_hla_csv(paths["hla"], d["hla"])
    return paths, L, d


def read_trace(path):
    """rows of the accept-test trace: 14 counters before the decision + the rand() draw"""
    return np.loadtxt(path, dtype=np.int64, ndmin=2)


def decisions(trace, final_counts=None):
    """per accept test: (move class 0..6, accepted 0/1), by differencing the counters"""
    out = []
    rows = list(trace[:, :14])
    if final_counts is not None:
        rows.append(np.asarray(final_counts))
    for a, b in zip(rows[:-1], rows[1:]):
        diff = b - a
        idx = np.nonzero(diff)[0]
        # a failed grow proposal bumps grow_reject without an accept test (mcmc_moves.c:725-728)
        idx = [i for i in idx]
        out.append(tuple(int(i) for i in idx))
    return out
