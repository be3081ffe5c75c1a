"""Profiling targets of round 2 (run under ncu, one kernel each): python tools/prof_r2.py <what>
  closest   : closest-L + consensus on a 960-query slice of C3 (k_pack, k_hamming, k_select_row, k_consensus)
  ckpt      : C4 on checkpointed matrices (stride 4): full evaluation, backward fill, a window proposal
  column    : C4, whole matrices: one-column proposals (k_column) and an accept (k_commit_all)
  upload    : C4 upload (k_slot_rows, k_gather)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402

what = sys.argv[1]
if what == "closest":
    data = bench.make_data("C3")
    refs, qs = data["ref_seqs"], np.ascontiguousarray(data["query_seqs"][:960])
    os.environ["MCQ_CLOSEST_SERIAL"] = "1"
    for _ in range(2):
        cl = api.closest_batch(refs, qs, 500)
        api.consensus_batch(refs, cl)
    print(api.closest_last_timing(), api.consensus_last_timing())
else:
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    p = bench.build_shard("C4", 0, nq, 0)
    S = p.S
    if what == "ckpt":
        ctx = api.Context(p.S, p.L, p.N, p.Q, p.H)
        ctx.set_option(api.OPT_CHECKPOINT_STRIDE, 4)
        ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
                          p.n_site_codons, p.site_codons)
        ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                       rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
        ctx.init_state()
        for _ in range(3):
            ctx.full_llk(True)
        ctx.backward_fill()
        for _ in range(3):
            ctx.propose(api.PROP_REV, S // 3, S // 3 + S // 6 - 1, value0=1.1)
    elif what == "column":
        ctx = api.Context.from_problem(p)
        ctx.init_state()
        for i in range(4):
            ctx.propose(api.PROP_REC, S // 2 + i, value0=0.02)
            ctx.reject()
        ctx.propose(api.PROP_REV, S // 3, S // 3 + S // 6 - 1, value0=1.1)
        ctx.accept()
    elif what == "upload":
        ctx = api.Context.from_problem(p)
        for _ in range(2):
            ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla,
                              p.consensus_codons, p.n_site_codons, p.site_codons)
    ctx.synchronize()
    ctx.close()
print("done")
