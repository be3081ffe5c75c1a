#!/usr/bin/env python
"""Tuning aid: wall and device time of the MCMC building blocks at a given workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mcqueen_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
p = bench.build_shard(name, 0, bench.workload_dims(name)["Q"], 0)
ctx = api.Context.from_problem(p)
ctx.init_state()
S = p.S
rng = np.random.default_rng(0)

def wall(fn, n=200):
    ctx.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize(); return (time.perf_counter() - t) / n * 1e6

def rec():
    ctx.propose(api.PROP_REC, int(rng.integers(1, S)), value0=0.011); ctx.reject()
def sel():
    ctx.propose(api.PROP_SEL, int(rng.integers(0, S)), value0=1.1); ctx.reject()
def win(accept):
    a = int(rng.integers(0, S - 90)); b = a + 80
    ctx.propose(api.PROP_ESC, a, b, hla=int(rng.integers(0, p.H)), value0=1.2)
    ctx.accept() if accept else ctx.reject()
def rec_acc():
    ctx.propose(api.PROP_REC, int(rng.integers(1, S)), value0=0.011); ctx.accept()
def full():
    ctx.full_llk(False)
print(f"{name}: Q={p.Q} S={S} L={p.L}")
print(f"REC propose+reject      {wall(rec):8.1f} us")
print(f"SEL propose+reject      {wall(sel):8.1f} us")
print(f"ESC(80 sites) + reject  {wall(lambda: win(False)):8.1f} us")
print(f"ESC(80 sites) + accept  {wall(lambda: win(True)):8.1f} us")
print(f"REC propose + accept    {wall(rec_acc):8.1f} us")
print(f"full forward sweep      {wall(full, 50):8.1f} us")
ctx.enable_timing(1)
ctx.propose(api.PROP_REC, 100, value0=0.011); print("REC propose device ms", ctx.last_timing()); ctx.accept(); print("REC accept device ms", ctx.last_timing())
