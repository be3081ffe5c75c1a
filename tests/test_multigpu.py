"""Two ranks: the sharded evaluation with the peer-mailbox exchange and with the NCCL all-reduce both equal the
single-GPU total, and proposals / accepts stay in lock step.  With two GPUs (`gpurun --gpus 2`) each rank has its
own; on a single-GPU box the peer-mailbox variant still runs, both ranks time-sharing device 0 (the mailboxes are
CUDA IPC mappings either way), and only the NCCL variant - which refuses two ranks on one device - is skipped.
Also the driver: two ranks of mcq_dmodel started WITHOUT seeds run one and the same chain."""
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
    import torch
    ctx = api.Context.from_problem(p, device=rank % torch.cuda.device_count(), q0=q0, q1=q1)
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
    if torch.cuda.device_count() < 2 and kind == "nccl":
        pytest.skip("NCCL needs one GPU per rank")
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


def test_two_rank_driver_without_seeds_runs_one_chain(tmp_path):
    """mcq_dmodel under two ranks with clock-derived seeds: rank 0 hands its seeds over, so both ranks print the
    same counters and the same final log-likelihood - the ones a single rank gets from those seeds."""
    import re
    import subprocess
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from chain_files import make_chain_inputs
    driver = os.path.join(ROOT, "mcqueen_b200", "host", "mcq_dmodel")
    if not os.path.exists(driver):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(driver)])
    paths, L, _ = make_chain_inputs(str(tmp_path / "in"), "toy")
    args = ["-H", paths["hla"], "-l", str(L), "-n", "400", "-s", "50", paths["ref"], paths["query"]]
    ndev = torch.cuda.device_count()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r % ndev), MASTER_PORT="29655")
        env.pop("MCQ_SEED", None); env.pop("MCQ_GSL_SEED", None)
        procs.append(subprocess.Popen([driver] + args + [str(tmp_path / f"out{r}")], env=env, text=True,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs[0][-1500:] + outs[1][-1500:]

    def tail(o):
        return o[o.index("Proposed"):o.index("MCMC loop")], float(re.search(r"final log-likelihood: (\S+)", o).group(1))

    assert tail(outs[0]) == tail(outs[1])
    a, b = re.search(r"seeds: (\d+) (\d+)", outs[0]).groups()
    one = subprocess.run([driver, "--seed", a, "--gsl_seed", b] + args + [str(tmp_path / "single")],
                         capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-1500:] + one.stderr[-1500:]
    c1, l1 = tail(one.stdout)
    c2, l2 = tail(outs[0])
    assert c1 == c2 and abs(l1 - l2) <= 1e-11 * abs(l1)
