"""A/B of one build against another on the same box: the stored forward sweep, the likelihood-only sweep and the
backward fill at a given shard size.  Run from the root of the tree under test:
    python tools/ab_stored.py C4 1250 [blocks_per_sm]
(the same file is dropped into a checkout of another commit for the comparison)."""
import os
import sys
import time

ROOT = os.getcwd()
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402

name, nq = sys.argv[1], int(sys.argv[2])
p = bench.build_shard(name, 0, nq, 0)
ctx = api.Context.from_problem(p, device=0)
ctx.init_state()
tuned = ctx.tune() if len(sys.argv) <= 3 else None
for rep in range(3):
    ctx.enable_timing(3)
    ctx.kernel_time(api.K_FORWARD, reset=True)
    for _ in range(30):
        ctx.full_llk(True)
    ms, n = ctx.kernel_time(api.K_FORWARD)
    ctx.enable_timing(0)
    t0 = time.perf_counter()
    for _ in range(30):
        ctx.full_llk(True)
    wall = (time.perf_counter() - t0) / 30 * 1e3
    t0 = time.perf_counter()
    for _ in range(30):
        ctx.llk_only(False)
    wall_l = (time.perf_counter() - t0) / 30 * 1e3
    print(f"{ROOT[-8:]} Q={nq} stored forward kernel {ms / n:.4f} ms ({n} launches)  eval wall {wall:.4f} ms  llk-only wall {wall_l:.4f} ms  tuned {tuned}")
if hasattr(ctx, "backward_fill"):
    for rep in range(2):
        ctx.backward_fill()
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.backward_fill()
        print(f"{ROOT[-8:]} Q={nq} backward fill wall {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms")
if hasattr(api, "OPT_CHECKPOINT_STRIDE"):
    ck = api.Context(p.S, p.L, p.N, p.Q, p.H, device=0)
    ck.set_option(api.OPT_CHECKPOINT_STRIDE, 4)
    ck.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
                     p.n_site_codons, p.site_codons)
    ck.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                  rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
    print(f"{ROOT[-8:]} Q={nq} checkpoint 4: init_state llk {ck.init_state()!r}  full_llk {ck.full_llk(True)!r}  llk_only {ck.llk_only(True)!r}")
    for rep in range(2):
        t0 = time.perf_counter()
        for _ in range(30):
            ck.full_llk(True)
        a = (time.perf_counter() - t0) / 30 * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            ck.backward_fill()
        print(f"{ROOT[-8:]} Q={nq} checkpoint 4: eval wall {a:.4f} ms  backward fill wall {(time.perf_counter() - t0) / 20 * 1e3:.4f} ms")
    ck.close()
from mcqueen_b200.api import PROP_REC, PROP_SEL  # noqa: E402
for kind, val in ((PROP_REC, 0.013), (PROP_SEL, 0.7)):
    for _ in range(20):
        ctx.propose(kind=kind, start=250, value0=val); ctx.reject()
    t0 = time.perf_counter()
    for i in range(300):
        ctx.propose(kind=kind, start=100 + i, value0=val); ctx.reject()
    print(f"{ROOT[-8:]} Q={nq} single-site proposal kind {kind}: {(time.perf_counter() - t0) / 300 * 1e6:.2f} us (propose + reject, host wall)")
ctx.close()
