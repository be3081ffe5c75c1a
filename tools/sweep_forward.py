#!/usr/bin/env python
"""Tuning aid: forward / backward fill time versus number of queries on one GPU (kernel timing via
the library's CUDA events).  Not part of the product or the tests."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    qs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [500, 1000, 2000, 3552, 5000, 7104, 10000]
    p = bench.build_shard(name, 0, max(qs), 0)
    S, L = p.S, p.L
    for Q in qs:
        ctx = api.Context.from_problem(p, device=0, q0=0, q1=Q)
        ctx.init_state()
        ctx.enable_timing(2)
        for _ in range(3):
            ctx.full_llk(False)
        ctx.kernel_time(api.K_FORWARD, reset=True)
        for _ in range(10):
            ctx.full_llk(False)
        f_ms, f_n = ctx.kernel_time(api.K_FORWARD, reset=True)
        for _ in range(5):
            ctx.backward_fill()
        b_ms, b_n = ctx.kernel_time(api.K_BACKWARD, reset=True)
        ctx.enable_timing(0)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            ctx.full_llk(True)
        step_ms = (time.perf_counter() - t0) / 50 * 1e3
        ctx.enable_timing(2)
        ctx.kernel_time(api.K_EMISSION, reset=True)
        ctx.full_llk(True)
        e_ms = ctx.kernel_time(api.K_EMISSION)[0]
        print(f"   step (emission+forward+sum+D2H, wall) {step_ms:.3f} ms; emission kernels {e_ms:.3f} ms")
        fb = (9 + 520 / L) * Q * S * L
        print(f"Q={Q:6d} forward {f_ms / f_n:7.3f} ms  {fb / (f_ms / f_n) / 1e6:7.0f} GB/s   "
              f"backward {b_ms / b_n:7.3f} ms {fb / (b_ms / b_n) / 1e6:7.0f} GB/s", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
