"""Two ranks on two GPUs (run with `gpurun --gpus 2`; skipped on a single-GPU box): the sharded
evaluation with the NVLink peer exchange and with the NCCL all-reduce both equal the single-GPU total,
and proposals / accepts stay in lock step."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mcqueen_b200 import api
    from problems import make_problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem("toy")
    q0, q1 = p.Q * rank // world, p.Q * (rank + 1) // world
    ctx = api.Context.from_problem(p, device=rank, q0=q0, q1=q1)
    if kind == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_handle())
        ctx.attach_peers(world, rank, handles)
    else:
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.attach_comm(world, rank, ids[0])
    # a second upload, now that the ranks know each other: with peers attached every rank DMAs one slice
    # of the reference panel and reads the other slices from its peers (MCQ_OPT_SHARED_PANEL)
    ctx.set_option(api.OPT_UPLOAD_CHUNK_QUERIES, 4)      # several chunks even at test sizes
    for _ in range(3):
        ctx.upload_static(p.ref_triplets, p.closest_refs[q0:q1], p.cons_query_nucls[q0:q1], p.query_codons[q0:q1],
                          p.hla[q0:q1], p.consensus_codons, p.n_site_codons, p.site_codons)
    vals = [ctx.init_state()]
    vals.append(ctx.propose(api.PROP_ESC, 10, 40, hla=1, value0=1.5, value1=0.8, split=25))
    ctx.accept()
    vals.append(ctx.propose(api.PROP_REC, 50, value0=0.03))
    ctx.reject()
    vals.append(ctx.propose(api.PROP_MU, 0, value0=p.mu * 1.05))
    ctx.accept()
    for _ in range(200):                       # many exchanges back to back (slot parity, step counter)
        v = ctx.full_llk(False)
    vals.append(v)
    gathered = [None] * world
    dist.all_gather_object(gathered, vals)
    if rank == 0:
        np.save(out, np.array(gathered))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["peer", "nccl"])
def test_two_gpu_exchange(tmp_path, kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from mcqueen_b200 import api
    from problems import make_problem
    out = str(tmp_path / "vals.npy")
    mp.spawn(_worker, args=(2, 29641 if kind == "peer" else 29642, kind, out), nprocs=2, join=True)
    got = np.load(out)
    assert np.array_equal(got[0], got[1])      # every rank sees the same bits
    p = make_problem("toy")
    ctx = api.Context.from_problem(p, device=0)
    want = [ctx.init_state()]
    want.append(ctx.propose(api.PROP_ESC, 10, 40, hla=1, value0=1.5, value1=0.8, split=25))
    ctx.accept()
    want.append(ctx.propose(api.PROP_REC, 50, value0=0.03))
    ctx.reject()
    want.append(ctx.propose(api.PROP_MU, 0, value0=p.mu * 1.05))
    ctx.accept()
    want.append(ctx.full_llk(False))
    ctx.close()
    assert np.allclose(got[0], want, rtol=1e-12, atol=0)
