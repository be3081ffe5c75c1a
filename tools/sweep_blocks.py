import os, sys
sys.path.insert(0, "/root/repo")
import bench
from mcqueen_b200 import api
p = bench.build_shard("C4", 0, 10000, 0)
for Q in (10000, 5000, 2500):
    ctx = api.Context.from_problem(p, device=0, q0=0, q1=Q)
    ctx.init_state(); ctx.enable_timing(2)
    for _ in range(3): ctx.full_llk(False)
    ctx.kernel_time(api.K_FORWARD, reset=True)
    for _ in range(12): ctx.full_llk(False)
    f_ms, f_n = ctx.kernel_time(api.K_FORWARD, reset=True)
    for _ in range(6): ctx.backward_fill()
    b_ms, b_n = ctx.kernel_time(api.K_BACKWARD, reset=True)
    print(f"blocks={os.environ.get('MCQ_FWD_BLOCKS')} Q={Q} forward {f_ms/f_n:.4f} ms backward {b_ms/b_n:.4f} ms", flush=True)
    ctx.close()
