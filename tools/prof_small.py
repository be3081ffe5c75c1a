"""Profiling target: the likelihood-only forward sweep on a shard that leaves about one warp per SM (the regime
where a sweep costs the dependent chain of one warp per site).  python tools/prof_small.py C4 148 [pipe]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402

name, nq = sys.argv[1], int(sys.argv[2])
p = bench.build_shard(name, 0, nq, 0)
ctx = api.Context.from_problem(p, device=0)
ctx.set_option(api.OPT_PIPELINE_MAX_QUERIES, (1 << 30) if len(sys.argv) > 3 else 0)
ctx.init_state()
for _ in range(4):
    v = ctx.llk_only(False)
print(v)
ctx.close()
