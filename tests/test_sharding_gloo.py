"""N > 1 host logic on CPU (gloo, world_size 2): contiguous query shards, per-rank partial
log-likelihoods, one all-reduce - the same decomposition bench.py and mcq_dmodel use with NCCL.
The per-rank compute is the CPU oracle here (TEST code); the product path needs a GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import cpu_libs
    from problems import make_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem("toy")
    q0, q1 = p.Q * rank // world, p.Q * (rank + 1) // world      # the sharding rule of bench.py / mcq_dmodel.c
    orc = cpu_libs.Checker("oracle")
    c2c = np.zeros((p.Q, 64, p.S)); A = np.zeros_like(c2c)
    orc.emission(p, c2c, A, 0, p.S - 1, -1, 2)
    part = 0.0
    for q in range(q0, q1):
        F = np.zeros((p.L, p.S)); Fn = np.zeros(p.S, np.int32)
        orc.init_forward(p, q, c2c[q], p.R, F, Fn)
        part += orc.site_llk(p, F, Fn, None, None, p.S - 1)
    t = torch.tensor([part], dtype=torch.float64)
    dist.all_reduce(t)                                           # the path's only exchange: one double
    cover = torch.zeros(p.Q, dtype=torch.int32)
    cover[q0:q1] = 1
    dist.all_reduce(cover)
    if rank == 0:
        np.save(out, np.array([t.item(), float(cover.min()), float(cover.max())]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_sum_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_libs
    from problems import make_problem
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, 29631, out), nprocs=2, join=True)
    total, cmin, cmax = np.load(out)
    assert cmin == 1 and cmax == 1                               # every query on exactly one rank
    p = make_problem("toy")
    st = cpu_libs.full_state(cpu_libs.Checker("oracle"), p)
    want = float(st["llk"].sum())
    assert abs(total - want) <= 1e-12 * abs(want)
