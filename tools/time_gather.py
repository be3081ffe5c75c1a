"""Device time of the gather (k_slot_rows + k_gather chunks) of one C4 upload: python tools/time_gather.py [queries]"""
import os
import sys

ROOT = os.getcwd()
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
p = bench.build_shard("C4", 0, nq, 0)
ctx = api.Context.from_problem(p, device=0)
want = ctx.full_llk(True)
ctx.enable_timing(2)
for rep in range(3):
    ctx.kernel_time(api.K_GATHER, reset=True)
    for _ in range(3):
        ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla,
                          p.consensus_codons, p.n_site_codons, p.site_codons)
    ms, n = ctx.kernel_time(api.K_GATHER)
    print(f"{ROOT[-8:]} Q={nq} gather kernels per upload {ms / 3:.4f} ms ({n // 3} launches)")
ctx.enable_timing(0)
ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
               rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
got = ctx.full_llk(True)
print("llk after re-upload equal:", got == want, repr(got))
ctx.close()
