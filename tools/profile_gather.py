#!/usr/bin/env python
"""Tuning aid: repeats mcq_ctx_upload_static (H2D + k_gather) on a C4-shaped shard."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
p = bench.build_shard("C4", 0, Q, 0)
ctx = api.Context.from_problem(p, device=0)
ctx.enable_timing(2)
for a in (p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla):
    api.host_register(a)
for i in range(4):
    ctx.kernel_time(api.K_GATHER, reset=True)
    ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla,
                      p.consensus_codons, p.n_site_codons, p.site_codons)
    print("upload_static total ms", ctx.last_timing()[0], "gather ms", ctx.kernel_time(api.K_GATHER)[0], flush=True)
