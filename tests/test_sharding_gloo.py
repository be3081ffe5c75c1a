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


def _closest_worker(rank, world, port, out):
    """closest-L sharded as bench.py / mcq_dmodel.c shard it: every rank makes the selection groups of its queries,
    the blocks are exchanged (all_gather_object over gloo), every rank replays ALL blocks in rank order from the
    shared seed (the product's host function mcq_closest_replay)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import cpu_libs
    from mcqueen_b200 import api, synth
    from test_closest_replay_cpu import groups_block
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synth.make_dataset(seed=9, S=30, N=90, Q=17, H=4, ref_unknown_frac=0.1, query_gap=True)
    refs, qs = d["ref_seqs"], d["query_seqs"]
    Q, n = qs.shape[0], 12
    q0, q1 = Q * rank // world, Q * (rank + 1) // world
    chk = cpu_libs.Checker("oracle")
    mine = groups_block([chk.hamming(refs, qs[q]) for q in range(q0, q1)], n)    # (the device part, on the CPU here)
    blocks = [None] * world
    dist.all_gather_object(blocks, mine)
    cpu_libs.srandom(21)
    lists = np.concatenate([api.closest_replay(b, n) for b in blocks])
    draws_after = cpu_libs.libc_random()
    np.save(out + f".{rank}.npy", np.concatenate([lists.ravel(), [draws_after]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_closest_replay_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_libs
    from mcqueen_b200 import build, synth
    build.build()
    out = str(tmp_path / "closest")
    mp.spawn(_closest_worker, args=(2, 29633, out), nprocs=2, join=True)
    r0, r1 = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(r0, r1)                                # both ranks hold the same lists and the same stream
    d = synth.make_dataset(seed=9, S=30, N=90, Q=17, H=4, ref_unknown_frac=0.1, query_gap=True)
    chk = cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")
    cpu_libs.srandom(21)
    want = np.stack([chk.closest(chk.hamming(d["ref_seqs"], q), 12) for q in d["query_seqs"]])
    after = cpu_libs.libc_random()
    assert np.array_equal(r0[:-1].reshape(want.shape), want) and int(r0[-1]) == after
