"""Times the sweeps (stored forward, likelihood-only, backward, an 80-site window proposal, a one-site proposal)
at a given shard size for every number of warps per query that is built.  Run on a GPU box:
    python tools/sweep_mw.py C4 1250 [2500 ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402


def main():
    name = sys.argv[1]
    sizes = [int(x) for x in sys.argv[2:]] or [1250]
    cfg = bench.workload_dims(name)
    S, L = cfg["S"], cfg["L"]
    p = bench.build_shard(name, 0, max(sizes), 0)
    NJ = 1
    while 64 * NJ < (L + 15) // 16 * 16:
        NJ *= 2
    for nq in sizes:
        ws = [0, -1] + ([w for w in (1, 2, 4, 8) if w <= NJ and NJ >= 2 and NJ // w <= 8] if os.environ.get("MW") else [])
        for W in ws:              # -1: one warp per query, pipelined
            ctx = api.Context.from_problem(p, device=0, q0=0, q1=nq)
            ctx.set_option(api.OPT_WARPS_PER_QUERY, max(W, 0))
            ctx.set_option(api.OPT_PIPELINE_MAX_QUERIES, 1 << 30 if W == -1 else 0)
            llk = ctx.init_state()
            if W == 0:
                ctx.tune()
            ctx.enable_timing(1)

            def timed(fn, reps=5):
                best = 1e30
                for _ in range(reps):
                    fn()
                    best = min(best, ctx.last_timing()[0])
                return best

            cells = nq * S * L
            t_full = timed(lambda: ctx.full_llk(False))
            t_only = timed(lambda: ctx.llk_only(False))
            ctx.init_state()
            t_bwd = timed(ctx.backward_fill)
            w0, w1 = S // 3, S // 3 + S // 6 - 1
            t_win = timed(lambda: ctx.propose(api.PROP_REV, w0, w1, value0=1.1))
            t_one = timed(lambda: ctx.propose(api.PROP_REC, S // 2, value0=0.02), reps=8)
            print(f"{name} Q={nq} W={W}: llk={llk:.6f} stored {t_full:.3f} ms ({cells / t_full * 1e3:.3e} cells/s)  "
                  f"llk-only {t_only:.3f} ms ({cells / t_only * 1e3:.3e})  backward {t_bwd:.3f} ms  "
                  f"window({w1 - w0 + 1}) {t_win * 1e3:.1f} us  one-site {t_one * 1e3:.1f} us", flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
