#!/usr/bin/env python
"""Benchmark of the McQueen likelihood path on B200 (contract: see DESIGN.md "Measurement").

One *step* = one full likelihood evaluation of the shard's queries: emission table for all sites
(K3), forward fill materialising F (K4), per-query log-likelihood (K6) and the cross-query /
cross-GPU sum - what dmodel does at start-up and what mutation_move re-does per proposal
(dmodel.c:873-902, mcmc_moves.c:1124-1134).  Metric: site x state cells per second = Q*S*L / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|toy] [--impl reference]

N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...`;
queries are sharded over ranks (strong scaling: BASELINE.json names one fixed workload "sharded
1/2/4/8"), the only exchange is one double per rank per evaluation: NVLink stores into the peers'
mailboxes inside the reduction kernel (MCQ_BENCH_COMM=peer, the default) or one ncclAllReduce
(MCQ_BENCH_COMM=nccl); the one that is not the headline is timed beside it (`exchange_comparison`).
The problem is the same at every N (closest-L lists of the whole workload, one tie-shuffle stream):
`llk_matches_single_gpu` checks the total against the recorded one-GPU value and a sample of
queries bit for bit against a one-GPU re-evaluation.  The dmodel MCMC steps/s are measured at
every N as well (one driver process per rank).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mcqueen_b200 import genetics as gen, prep, synth  # noqa: E402

METRIC = "query-HMM evals/sec (site x state cells/s)"
UNIT = "cells/s"
SEED = 20261019


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def workload_dims(name):
    return dict(synth.CONFIGS[name])


def make_data(name):
    cfg = workload_dims(name)
    cfg.pop("L")
    return synth.make_dataset(seed=SEED, self_reference=name in synth.SELF_REFERENCE, **cfg)


def build_shard(name, q0, q1, device_index, dist=None, world=1):
    """Problem restricted to queries q0..q1-1 (references are shared by every shard).

    The closest-L lists are those of the WHOLE workload: the reference shuffles boundary ties with one libc
    random() stream consumed in global query order (closest_seqs.c:57-64, called per query from
    dmodel.c:356-387).  The device part (distances, selection groups) is done for this rank's queries only; the
    groups of all ranks are exchanged and every rank replays the tie shuffles of all queries from the same seed - as
    mcq_dmodel does - and keeps its shard.  The problem is then the same at every number of GPUs."""
    from mcqueen_b200 import api
    cfg = workload_dims(name)
    L = cfg["L"]
    t0 = time.time()
    data = make_data(name)
    refs, qs_all = data["ref_seqs"], data["query_seqs"]
    qs, hla = np.ascontiguousarray(qs_all[q0:q1]), data["hla"][q0:q1]
    skip = name in synth.SELF_REFERENCE
    import ctypes
    # CUDA context creation and module load happen on the first call: keep them out of the closest-L wall time
    api.closest_batch(refs[:max(2 * L, 64)], qs[:2], min(L, 8), device=device_index)
    t1 = time.time()
    groups = api.closest_groups(refs, qs, L, skip_first=skip, device=device_index)
    blocks = [groups]
    if world > 1:
        blocks = [None] * world
        dist.all_gather_object(blocks, groups)
    ctypes.CDLL("libc.so.6").srandom(SEED & 0x7fffffff)
    closest_all = np.concatenate([api.closest_replay(b, L, skip_first=skip) for b in blocks])
    t2 = time.time()
    ham_ovl, sel_ovl = api.closest_last_timing()
    # the kernel times proper: the same call once more with the chunks one after the other (in the pass above they
    # alternate between two streams, so a chunk's events also span its wait for the other stream's kernels)
    os.environ["MCQ_CLOSEST_SERIAL"] = "1"
    api.closest_groups(refs, qs, L, skip_first=skip, device=device_index)
    del os.environ["MCQ_CLOSEST_SERIAL"]
    ham_ms, sel_ms = api.closest_last_timing()
    closest = np.ascontiguousarray(closest_all[q0:q1])
    cons_q = api.consensus_batch(refs, closest, device=device_index)
    p = prep.build_problem(refs, qs, hla, closest, cons_q=cons_q)
    prep.random_parameters(p, seed=1)
    t3 = time.time()
    log(f"[bench] data {t1 - t0:.1f}s  closest-L(GPU: {qs.shape[0]} queries here, tie shuffles of all {qs_all.shape[0]}) {t2 - t1:.2f}s  prep {t3 - t2:.1f}s  "
        f"Q={p.Q} S={p.S} L={p.L} N={p.N} H={p.H}")
    p.closest_seconds = t2 - t1
    p.full = dict(refs=refs, queries=qs_all, hla=data["hla"], closest=closest_all)
    nb = refs.shape[1]
    nq = qs.shape[0]
    p.closest_stats = {
        "shape": f"{refs.shape[0]} refs x {nq} queries on this GPU (of {qs_all.shape[0]}) x {nb} nt, L={L} (BASELINE config 3 when the workload is C4)",
        "hamming_kernel_ms": ham_ms, "select_kernel_ms": sel_ms,
        "kernel_times": "CUDA events per chunk of queries, summed, from a pass that runs the chunks one after the other; "
                        "the wall time is of the pass that alternates chunks between two streams",
        "event_sums_of_the_two_stream_pass_ms": {"hamming": ham_ovl, "select": sel_ovl},
        "wall_seconds_incl_transfers_exchange_and_tie_shuffles": t2 - t1,
        "nt_compares_per_s": refs.shape[0] * nq * nb / (ham_ms * 1e-3) if ham_ms > 0 else None,
        # one 64-bit POPC word per 64 nt per pair: B200 issues 16 POPC.32 lanes/clk/SM
        "popc_word_ops_per_s": refs.shape[0] * nq * ((nb + 63) // 64) / (ham_ms * 1e-3) if ham_ms > 0 else None,
        "queries_per_s": nq / (t2 - t1)}
    return p


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def mark(self):
        """Wall-clock mark (start / end of the timed region) in nvidia-smi's timestamp format."""
        import datetime
        return datetime.datetime.now()

    def summary(self, t0=None, t1=None):
        """Median SM clock and throttle reasons over the samples taken between t0 and t1 (all
        samples when the window caught none)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.path is None or not os.path.exists(self.path):
            return out
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        out["window"] = "timed region" if inside else "whole run (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for ts, a, b, flags in (inside or rows):
            sm.append(a); mx.append(b)
            for n, v in zip(names, flags):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own code (oracle/_ref/ref_bench), host threads over query shards
# ------------------------------------------------------------------------------------------------
def write_ref_problem(p, path, nq, seqs=None):
    """The problem file oracle/ref_bench.c reads; seqs = (ref_seqs, query_seqs) ASCII adds the section its
    closest-L mode needs."""
    nq = min(nq, p.Q)
    with open(path, "wb") as f:
        np.array([p.S, p.N, nq, p.L, p.H], np.int32).tofile(f)
        np.array([p.mu, p.kappa, p.phi], np.float64).tofile(f)
        p.prior.astype(np.float64).tofile(f)
        p.ref_triplets.tofile(f); p.ref_codons.tofile(f); p.query_codons[:nq].tofile(f)
        p.cons_query_nucls[:nq].tofile(f)
        hl = np.zeros((nq, p.H + 1), np.uint8); hl[:, :p.H] = p.hla[:nq]; hl.tofile(f)
        p.consensus_codons.tofile(f); p.site_codons.tofile(f); p.type_of_change.tofile(f)
        p.closest_refs[:nq].astype(np.int32).tofile(f); p.n_site_codons.astype(np.int32).tofile(f)
        p.ns_ts_sum.astype(np.int32).tofile(f); p.ns_tv_sum.astype(np.int32).tofile(f)
        p.R.tofile(f); p.omega.tofile(f); np.ascontiguousarray(p.esc).tofile(f)
        np.ascontiguousarray(p.rev).tofile(f); p.rate_out_of_codons().tofile(f)
        if seqs is not None:
            refs, qs = seqs
            nb = refs.shape[1]
            np.array([nb], np.int32).tofile(f)
            for a in (refs, qs[:nq]):
                z = np.zeros((a.shape[0], nb + 1), np.uint8); z[:, :nb] = a; z.tofile(f)
    return nq


REF_MODES = {0: "create_codon_to_codon_HLA + initialise_F_numerical_recipes + llk per query",
             2: "update_B_numerical_recipes (update = S-1) alone",
             3: "window proposal: update_F_numerical_recipes over S/6 sites + sum_r tmp_F*B",
             4: "one-column proposal: update_F_numerical_recipes at one site + sum_r tmp_F*B",
             5: "closest-L: hamming_distance_considering_unknown + closest_sequences per query"}
CPU_SAMPLE_QUERIES = 200      # time >= 200 queries of the large shapes, scale linearly in Q


def run_ref_bench(p, threads, seconds, path=None, total_queries=None, mode=0, seqs=None):
    """Times the reference's own objects on a bounded sample; returns (units/s, description dict).  Units are
    cells (nt-compares in mode 5)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    if not os.path.exists(exe):
        return None, {"error": "oracle/_ref/ref_bench missing"}
    own = path is None
    if own:
        fd, path = tempfile.mkstemp(suffix=".bin")
        os.close(fd)
        write_ref_problem(p, path, max(CPU_SAMPLE_QUERIES, threads * 4), seqs)
    nq = min(p.Q, max(CPU_SAMPLE_QUERIES, threads * 4))
    if mode == 5:
        nq = min(nq, max(threads, 16))      # 0.36 s per query per core: one query per thread is a sample already

    def once(reps):
        out = subprocess.run([exe, path, str(threads), str(nq), str(mode), str(reps)], capture_output=True, text=True,
                             check=True).stdout
        return json.loads(out.strip().splitlines()[-1])

    probe = once(1)
    reps = max(1, int(seconds / max(probe["wall_seconds"], 1e-3)))
    res = once(reps) if reps > 1 else probe
    if own:
        os.unlink(path)
    desc = dict(kind="reference", cores=threads, value=res["cells_per_s"], unit=UNIT if mode != 5 else "nt-compares/s",
                sample=f"{nq} of {total_queries or p.Q} queries x {reps} repeats ({res['wall_seconds']:.1f} s): "
                       f"{REF_MODES[mode]}, reference objects built -O3 (oracle/Makefile), {threads} host "
                       f"thread{'s' if threads > 1 else ''} over query shards" +
                       ("; every thread re-runs its ~%d queries, so its working set (one F matrix) stays in cache: "
                        "generous to the CPU" % max(1, nq // threads) if reps > 1 and mode == 0 else ""))
    return res["cells_per_s"], desc


def cpu_baseline_suite(p, seconds, seqs, total_queries):
    """the headline region on all host cores and on one core, plus backward / window /
    column / closest-L, every one the reference's own objects on a sample of the same workload."""
    ncpu = os.cpu_count() or 1
    fd, path = tempfile.mkstemp(suffix=".bin")
    os.close(fd)
    try:
        write_ref_problem(p, path, max(CPU_SAMPLE_QUERIES, ncpu * 4), seqs)
        v, desc = run_ref_bench(p, ncpu, seconds, path, total_queries)
        if v is None:
            return {"value": None, "unit": UNIT, "cores": ncpu, "kind": "reference", "sample": desc["error"]}
        other = {}
        v1, d1 = run_ref_bench(p, 1, min(4.0, seconds), path, total_queries)
        other["one_core"] = {"value": v1, "unit": UNIT, "cores": 1, "sample": d1["sample"]}
        for key, mode in (("backward_fill", 2), ("window_proposal", 3), ("single_site_proposal", 4), ("closest_L", 5)):
            if mode == 5 and seqs is None:
                continue
            vm, dm = run_ref_bench(p, ncpu, min(3.0, seconds), path, total_queries, mode=mode)
            other[key] = {"value": vm, "unit": dm["unit"], "cores": ncpu, "sample": dm["sample"]}
        desc["other_regions"] = other
        return desc
    finally:
        os.unlink(path)


# ------------------------------------------------------------------------------------------------
# dmodel MCMC steps/s (second half of BASELINE.json's metric): the C driver on the GPU library next
# to the reference dmodel on the host CPU, same files, same seeds, default move mix with annealing
# ------------------------------------------------------------------------------------------------
def mcmc_steps_per_s(workload, steps, ref_steps=12,
                     variants=(("gpu", []), ("gpu_dense_esc", ["--dense_esc"]), ("gpu_sparse_esc", ["--sparse_esc"])),
                     rank=0, world=1, dist=None):
    """Runs the C driver (one process per rank, started by every bench rank with its own RANK / LOCAL_RANK) on
    files written by rank 0; rank 0 returns the figures, the others None."""
    import re
    import shutil
    import uuid
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from chain_files import make_chain_inputs
    driver = os.path.join(ROOT, "mcqueen_b200", "host", "mcq_dmodel")
    if not os.path.exists(driver):
        return {"error": "mcqueen_b200/host/mcq_dmodel not built"}
    box = [None, None]
    if rank == 0:
        box = [tempfile.mkdtemp(prefix="mcq_mcmc_"), uuid.uuid4().hex[:12]]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    tmp, tag = box
    out = {"workload": workload, "steps": steps, "n_gpus": world,
           "variants": "gpu = driver defaults (every fourth column of F / B kept, refills when next read, escape-window "
                       "proposals re-evaluate carriers only when the shard is large enough to be bound by bytes); "
                       "dense_esc = every query, as the reference; whole_matrices = --checkpoint 1; reference_schedule = "
                       "whole matrices, every query, refill after every accept (what the reference does, on the GPU)",
           "timing": "the chain's own clock (MCMC loop line of the driver): the final state file is written after it stops"}
    try:
        paths = L = None
        if rank == 0:
            paths, L, _ = make_chain_inputs(os.path.join(tmp, "in"), workload, seed=SEED, query_model="dsim")
        if world > 1:
            pl = [paths, L]
            dist.broadcast_object_list(pl, src=0)
            paths, L = pl
        out["data"] = "queries drawn from the model itself (dsim recipe, mcqueen_b200/synth.py)"
        base = ["-H", paths["hla"], "-l", str(L), "-s", "1000000", paths["ref"], paths["query"]]
        for i, (key, extra) in enumerate(variants):
            env = dict(os.environ, MCQ_RUN_TAG=f"{tag}{i}")
            r = subprocess.run([driver, "--seed", "1", "--gsl_seed", "1", "-n", str(steps)] + extra + base +
                               [os.path.join(tmp, f"{key}_r{rank}")], capture_output=True, text=True, env=env)
            m = re.search(r"MCMC loop: (\d+) steps in ([0-9.]+) s = ([0-9.]+) steps/s", r.stdout)
            out[key + "_steps_per_s"] = float(m.group(3)) if m else None
            m2 = re.search(r"final log-likelihood: (\S+)", r.stdout)
            out[key + "_final_llk"] = float(m2.group(1)) if m2 else None
            if not m:
                out[key + "_error"] = (r.stdout + r.stderr)[-300:]
            if world > 1:
                dist.barrier()
        ref = os.path.join(ROOT, "oracle", "_ref", "dmodel_ref")
        if rank == 0 and os.path.exists(ref) and ref_steps > 0:
            env = dict(os.environ, MCQ_SEED="1", MCQ_GSL_SEED="1")
            t = []
            for n in (1, 1 + ref_steps):     # the difference removes the start-up (closest-L, initial F/B)
                t0 = time.perf_counter()
                subprocess.run([ref, "-n", str(n)] + base + [os.path.join(tmp, f"ref{n}")], capture_output=True,
                               text=True, env=env, check=True)
                t.append(time.perf_counter() - t0)
            out["reference_cpu_steps_per_s"] = ref_steps / max(t[1] - t[0], 1e-9)
            out["reference_cpu"] = f"unmodified reference dmodel, 1 core, {ref_steps} steps after start-up " \
                                   f"({t[0]:.1f} s start-up incl. closest-L and initial F/B)"
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=list(synth.CONFIGS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-mcmc", action="store_true", help="skip the dmodel MCMC steps/s measurement")
    ap.add_argument("--mcmc-steps", type=int, default=3000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = workload_dims(args.workload)
    Qtot, S, L, N, H = cfg["Q"], cfg["S"], cfg["L"], cfg["N"], cfg["H"]
    config = {"workload": f"{args.workload}: synthetic {Qtot} queries x L={L} x {S} codons, N={N} refs, H={H} HLA "
                          f"types; full likelihood evaluation (emission + forward fill + llk)",
              "queries": Qtot, "sites": S, "closest_n": L, "refs": N, "hla_types": H,
              "sharding": f"queries/{world}", "exchange": os.environ.get("MCQ_BENCH_COMM", "peer") if world > 1 else "none", "cache": "inputs+outputs per step (>20 GB at C4) exceed L2; no flush"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        ncpu = os.cpu_count() or 1
        # the reference arm needs a problem; build it without the GPU library where none is present
        p = build_problem_cpu(args.workload, ncpu)
        fd, path = tempfile.mkstemp(suffix=".bin")
        os.close(fd)
        vals = []
        desc = None
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        write_ref_problem(p, path, p.Q)
        for i in range(args.warmup + args.steps):
            v, desc = run_ref_bench(p, ncpu, per_step, path, total_queries=Qtot)
            if v is None:
                print(json.dumps({"impl": "reference", "unavailable": desc["error"]}))
                return 0
            if i >= args.warmup:
                vals.append(v)
        os.unlink(path)
        value = float(np.mean(vals))
        desc["value"] = value
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": Qtot * S * L / value * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "cpu_baseline": desc,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist
    from mcqueen_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the McQueen path has no CPU fallback")
    # stdout carries the one JSON line and nothing else: whatever the libraries print there (NCCL writes its log
    # to stdout) is sent to stderr, the line goes out through a duplicate of the original descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # NCCL's INFO log (communicator size, transports) is part of the record (it lands on stderr, see below)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = local_rank
    # the clock sampler starts before the set-up: nvidia-smi needs a second or two before its first line
    clk = ClockSampler(dev)
    if rank == 0:
        clk.__enter__()
    q0, q1 = Qtot * rank // world, Qtot * (rank + 1) // world
    p = build_shard(args.workload, q0, q1, dev, dist=dist if world > 1 else None, world=world)
    ctx = api.Context.from_problem(p, device=dev)
    comm_kind = os.environ.get("MCQ_BENCH_COMM", "peer") if world > 1 else "none"
    if world > 1:
        # both exchanges are attached: the headline runs on `comm_kind`, the other one is timed beside it
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_handle())
        ctx.attach_peers(world, rank, handles)
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.attach_comm(world, rank, ids[0])
        ctx.set_option(api.OPT_EXCHANGE, 1 if comm_kind == "nccl" else 0)
    log(f"[bench] rank {rank}: context holds {ctx.device_bytes() / 2**30:.1f} GiB")

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()

    llk0 = ctx.init_state()
    tuned = ctx.tune()        # one-off: resident blocks per SM of the sweep kernels for this shard size
    config["resident_blocks_per_sm"] = dict(zip(["forward", "forward_llk_only", "backward"], tuned))
    for _ in range(args.warmup):
        ctx.full_llk(True)

    # ---- per-kernel shares of a step (bracketing every launch costs host time: outside the timed region)
    ctx.enable_timing(2)
    ctx.kernel_time(api.K_FORWARD, reset=True)
    for _ in range(3):
        ctx.full_llk(True)
    emi_ms, emi_n = ctx.kernel_time(api.K_EMISSION)
    emi_ms, emi_n = emi_ms / 3, emi_n // 3

    # ---- timed region: K steps, device time via CUDA events on the library's stream; the forward launches
    # are bracketed by event pairs that are read out after the region (no host wait inside a step)
    ctx.enable_timing(3)
    ctx.kernel_time(api.K_FORWARD, reset=True)
    barrier()
    if True:
        mark0 = clk.mark()
        t_wall = time.perf_counter()
        ctx.timer_start()
        launches = 0
        for _ in range(args.steps):
            llk = ctx.full_llk(True)
            launches += ctx.last_timing()[1]
        ms_dev = ctx.timer_stop()
        barrier()
        t_wall = (time.perf_counter() - t_wall) * 1e3
        mark1 = clk.mark()
    clk.__exit__()
    fwd_ms, fwd_n = ctx.kernel_time(api.K_FORWARD)
    ctx.enable_timing(0)
    clocks = clk.summary(mark0, mark1)
    assert abs(llk - llk0) <= 1e-12 * abs(llk0), (llk, llk0)

    # ---- the other exchange of the shard totals, timed beside the headline (same steps, same problem):
    # north_star names one NCCL all-reduce, the default is the peer-mailbox store fused into the reduction kernel
    exchange_cmp = None
    if world > 1:
        other = "nccl" if comm_kind == "peer" else "peer"
        ctx.set_option(api.OPT_EXCHANGE, 1 if other == "nccl" else 0)
        for _ in range(3):
            ctx.full_llk(True)
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            llk_other = ctx.full_llk(True)
        ms_other = ctx.timer_stop()
        barrier()
        ctx.set_option(api.OPT_EXCHANGE, 1 if comm_kind == "nccl" else 0)
        assert abs(llk_other - llk0) <= 1e-12 * abs(llk0), (llk_other, llk0)
        exchange_cmp = (other, ms_other, llk_other)

    # ---- the sharded evaluation IS the single-GPU evaluation: (a) the total against the value recorded from a
    # one-GPU run of this workload (profiles/llk_single_gpu.json), (b) a sample of queries spread over all shards,
    # re-evaluated in a one-GPU context of their own on rank 0, against the shards' per-query values bit for bit
    per_q_own = ctx.query_llk().copy()
    per_q_all = [per_q_own]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_q_all[0])
        per_q_all = gathered
    identity = None
    if rank == 0:
        per_q_all = np.concatenate(per_q_all)
        sample = np.unique(np.linspace(0, Qtot - 1, min(Qtot, CPU_SAMPLE_QUERIES)).astype(np.int64))
        f = p.full
        ps = prep.build_problem(f["refs"], np.ascontiguousarray(f["queries"][sample]), f["hla"][sample],
                                np.ascontiguousarray(f["closest"][sample]),
                                cons_q=api.consensus_batch(f["refs"], np.ascontiguousarray(f["closest"][sample]), device=dev))
        prep.random_parameters(ps, seed=1)
        cs = api.Context.from_problem(ps, device=dev)
        cs.full_llk(True)
        per_s = cs.query_llk()
        cs.close()
        identity = {"sample_queries": int(sample.size), "sample_bit_identical": bool(np.array_equal(per_s, per_q_all[sample])),
                    "sample_max_rel_diff": float(np.max(np.abs(per_s - per_q_all[sample]) / np.abs(per_s)))}
        rec_path = os.path.join(ROOT, "profiles", "llk_single_gpu.json")
        rec = json.load(open(rec_path)).get(args.workload) if os.path.exists(rec_path) else None
        if rec is not None and rec.get("seed") == SEED:
            identity["single_gpu_llk"] = rec["llk"]
            identity["total_rel_diff"] = abs(llk0 - rec["llk"]) / abs(rec["llk"])
            identity["total_matches"] = bool(identity["total_rel_diff"] <= 1e-12)
        log(f"[bench] sharded == single-GPU: {identity}")
    seqs_sample = (p.full["refs"], np.ascontiguousarray(p.full["queries"][q0:q0 + 1024]))
    p.full = None

    # ---- e2e: everything from (pinned) host buffers every step, per-query results back
    host_arrays = [p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla]
    for a in host_arrays:
        api.host_register(a)
    rate_out = p.rate_out_of_codons()
    # per rank; with the peer exchange attached each rank DMAs 1/world of the reference panel and reads the
    # rest from its peers over NVLink (MCQ_OPT_SHARED_PANEL)
    panel_share = p.ref_triplets.nbytes // world
    h2d = panel_share + sum(a.nbytes for a in host_arrays[1:]) + p.consensus_codons.nbytes + \
        p.n_site_codons.nbytes + p.site_codons.nbytes + 8 * (64 + 2 * S + 64 * S + H * S + S)
    d2h = 8 * p.Q + 8

    e2e_parts = np.zeros(4)

    def e2e_step():
        t0 = time.perf_counter()
        ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla,
                          p.consensus_codons, p.n_site_codons, p.site_codons)
        t1 = time.perf_counter()
        ctx.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                       rate_out=rate_out, esc=p.esc, rev=p.rev[0])
        t2 = time.perf_counter()
        tot = ctx.full_llk(True)
        t3 = time.perf_counter()
        per_q = ctx.query_llk()
        t4 = time.perf_counter()
        e2e_parts[:] += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
        return tot, per_q

    # the gather's share of an upload (bracketing kernels costs host time: outside the timed loop)
    ctx.enable_timing(2)
    ctx.kernel_time(api.K_GATHER, reset=True)
    for _ in range(3):
        ctx.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla,
                          p.consensus_codons, p.n_site_codons, p.site_codons)
    gather_ms = ctx.kernel_time(api.K_GATHER)[0] / 3
    ctx.enable_timing(0)
    # the upload only queues its DMAs (the host buffers stay untouched, they are the same every step); the
    # evaluation follows the transfers and the device-side set-up chunk by chunk
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 2)
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_parts[:] = 0
    ctx.timer_start()
    t_e2e = time.perf_counter()
    for _ in range(args.steps):
        tot, per_q = e2e_step()
    ms_e2e_dev = ctx.timer_stop()
    barrier()
    t_e2e = (time.perf_counter() - t_e2e) * 1e3
    ctx.set_option(api.OPT_ASYNC_UPLOAD, 0)
    for a in host_arrays:
        api.host_unregister(a)
    assert abs(tot - llk0) <= 1e-12 * abs(llk0)

    # ---- extras: the other kernels of the path (reported, not the headline)
    extras = {}
    if not args.no_extras:
        ctx.init_state()
        ctx.enable_timing(1)

        def timed(fn, reps=3):
            best = 1e30
            for _ in range(reps):
                fn()
                best = min(best, ctx.last_timing()[0])
            return best

        cells = p.Q * S * L
        # the same evaluation without writing F (mcq_ctx_llk_only): the recursion then runs at the
        # instruction-issue / FP64 rate, 1 B of G per cell from HBM
        t = timed(lambda: ctx.llk_only(True))
        extras["llk_only_ms"] = t
        extras["llk_only_cells_per_s"] = cells / t * 1e3
        ctx.init_state()
        t = timed(ctx.backward_fill)
        extras["backward_fill_cells_per_s"] = cells / t * 1e3
        w0, w1 = S // 3, S // 3 + S // 6 - 1
        t = timed(lambda: ctx.propose(api.PROP_REV, w0, w1, value0=1.1))
        extras["window_proposal_ms"] = t
        extras["window_proposal_cells_per_s"] = p.Q * (w1 - w0 + 1) * L / t * 1e3
        t = timed(lambda: ctx.propose(api.PROP_REC, S // 2, value0=0.02), reps=5)
        extras["single_site_proposal_ms"] = t
        # the same proposal, every kernel bracketed by its own pair of events (the figure above spans the call: it
        # ends when the host, having seen the result, records the closing event)
        ctx.enable_timing(2)
        ctx.kernel_time(api.K_FORWARD, reset=True)
        for _ in range(5):
            ctx.propose(api.PROP_REC, S // 2, value0=0.02)
        kf, nf = ctx.kernel_time(api.K_FORWARD)
        ks, ns = ctx.kernel_time(api.K_SUM)
        extras["single_site_kernels_ms"] = {"column": kf / max(nf, 1), "sum_and_exchange": ks / max(ns, 1)}
        # 25 L bytes per query move for one column (F in, tmp_F out, B in, 8 B each, + the G row)
        extras["single_site_roofline_frac"] = 25.0 * L * p.Q / ((kf / max(nf, 1) + ks / max(ns, 1)) * 1e-3) / 1e9 / \
            (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
             if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0)
        ctx.enable_timing(1)
        best = 1e30
        for v in (1.1, 1.2, 1.3):        # (the first accept of a process also loads the kernel: best of three)
            ctx.propose(api.PROP_REV, w0, w1, value0=v)
            ctx.accept()
            best = min(best, ctx.last_timing()[0])
        extras["window_accept_ms"] = best
        extras["closest_L"] = p.closest_stats
        ctx.enable_timing(0)
        # the same evaluation on checkpointed matrices (MCQ_OPT_CHECKPOINT_STRIDE = 4: every fourth column of F kept,
        # what the MCMC driver runs on by default): reported beside the headline, which keeps whole matrices
        ck = api.Context(p.S, p.L, p.N, p.Q, p.H, device=dev)
        ck.set_option(api.OPT_CHECKPOINT_STRIDE, 4)
        ck.upload_static(p.ref_triplets, p.closest_refs, p.cons_query_nucls, p.query_codons, p.hla, p.consensus_codons,
                         p.n_site_codons, p.site_codons)
        ck.set_params(mu=p.mu, kappa=p.kappa, phi=p.phi, prior=p.prior, R=p.R, omega=p.omega,
                      rate_out=p.rate_out_of_codons(), esc=p.esc, rev=p.rev[0])
        ck.init_state()
        # (this context is not attached to the other ranks: its values are this shard's, compared per query)
        ck_same = bool(np.array_equal(ck.query_llk(), per_q_own))
        ck.enable_timing(1)

        def timed_ck(fn, reps=3):
            best = 1e30
            for _ in range(reps):
                fn()
                best = min(best, ck.last_timing()[0])
            return best

        t = timed_ck(lambda: ck.full_llk(True))
        tw = timed_ck(lambda: ck.propose(api.PROP_REV, w0, w1, value0=1.1))
        extras["checkpoint_stride_4"] = {
            "full_llk_ms": t, "full_llk_cells_per_s": cells / t * 1e3, "llk_equals_whole_matrices": ck_same,
            "algorithmic_bytes_per_cell": 1.0 + 8.0 / 4 + 520.0 / L, "window_proposal_ms": tw,
            "window_proposal_cells_per_s": p.Q * (w1 - w0 + 1) * L / tw * 1e3,
            "device_gib": ck.device_bytes() / 2**30, "device_gib_whole_matrices": ctx.device_bytes() / 2**30}
        ck.close()

    # ---- gather over ranks (max time)
    stats = np.array([ms_dev, t_wall, ms_e2e_dev, t_e2e, fwd_ms / max(fwd_n, 1),
                      exchange_cmp[1] if exchange_cmp else 0.0], np.float64)
    per_rank = [stats.tolist()]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, stats.tolist())
        per_rank = gathered
        tt = torch.from_numpy(stats)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stats = tt.numpy()
    ms_step = max(stats[0], 0.0) / args.steps
    ms_step_wall = stats[1] / args.steps
    ms_e2e = max(stats[2], stats[3]) / args.steps
    fwd_ms_avg = stats[4]
    ms_other_max = stats[5]

    if rank == 0:
        total_cells = Qtot * S * L
        value = total_cells / (ms_step * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bytes_per_cell = 9.0 + 520.0 / L
        alg_bytes = bytes_per_cell * p.Q * S * L
        achieved = alg_bytes / (fwd_ms_avg * 1e-3) / 1e9
        traffic = traffic_src = None
        tpath = os.path.join(ROOT, "profiles", "forward_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("workload") == args.workload and tj.get("queries_per_gpu") == p.Q:
                traffic = tj["dram_bytes_per_launch"]
                traffic_src = "recorded, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one " \
                              "k_forward launch of this shape under ncu --set full (" + tj.get("source", "profiles/") + ")"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total_cells / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                    "breakdown_ms": dict(zip(["upload_static", "set_params", "full_llk", "read_back"],
                                             (e2e_parts / args.steps * 1e3).round(3).tolist()), gather_kernel=round(gather_ms, 3)),
                    "what": "mcq_ctx_upload_static + mcq_ctx_set_params + mcq_ctx_full_llk + per-query llk "
                            "read-back, all inputs from pinned host buffers every step, MCQ_OPT_ASYNC_UPLOAD=2: the calls queue "
                            "the DMAs and the evaluation follows them and the device-side set-up chunk by chunk (bytes per rank" +
                            ("; the reference panel is DMA'd in 1/N slices and completed over NVLink)" if world > 1 else ")")},
            "roofline": {"bound": "hbm", "kernel": "k_forward", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_cell": bytes_per_cell, "launch_ms": fwd_ms_avg,
                         "cells_per_launch": p.Q * S * L},
            "ms_per_step_wall": ms_step_wall, "llk": llk0, "extras": extras,
            "per_rank": [dict(zip(["timed_ms", "timed_wall_ms", "e2e_ms", "e2e_wall_ms", "forward_launch_ms"], r))
                         for r in per_rank],
            "kernel_share": {"forward_ms_per_step": fwd_ms / max(args.steps, 1),
                             "emission_ms_per_step": emi_ms},
        }
        # (a workload without a recorded one-GPU total - config 5 does not fit one GPU - is judged on the sample alone,
        # and says so)
        if "total_matches" in identity:
            line["llk_matches_single_gpu"] = bool(identity["total_matches"] and identity["sample_bit_identical"])
        else:
            line["llk_matches_single_gpu"] = bool(identity["sample_bit_identical"])
            identity["total"] = "no one-GPU total recorded for this workload: the 200-query sample is the check"
        line["single_gpu_identity"] = identity
        if exchange_cmp is not None:
            other, ms_other, llk_other = exchange_cmp
            line["exchange_comparison"] = {
                "headline": comm_kind, "other": other, "other_ms_per_step": ms_other_max / args.steps,
                "other_value": total_cells / (ms_other_max / args.steps * 1e-3), "other_llk": llk_other,
                "nccl_nranks": world,
                "what": "the same K steps with the shard totals exchanged the other way (peer = NVLink stores into the "
                        "peers' mailboxes inside the reduction kernel; nccl = k_sum + one ncclAllReduce of one double + "
                        "publish kernel); NCCL's INFO log of this communicator is on stderr"}
        if world == 1 and args.cpu_seconds > 0:
            line["cpu_baseline"] = cpu_baseline_suite(p, args.cpu_seconds, seqs_sample, Qtot)
    # ---- dmodel MCMC steps/s (the second half of the metric) at this number of GPUs: every rank starts the
    # driver for its GPU once the bench context has released the memory
    if not args.no_mcmc:
        ctx.close()
        mc = mcmc_steps_per_s("C2", args.mcmc_steps, ref_steps=12 if world == 1 else 0, rank=rank, world=world,
                              dist=dist if world > 1 else None,
                              variants=(("gpu", []), ("gpu_dense_esc", ["--dense_esc"]), ("gpu_sparse_esc", ["--sparse_esc"]),
                                        ("gpu_whole_matrices", ["--checkpoint", "1"]))
                              if world == 1 else (("gpu", []), ("gpu_whole_matrices", ["--checkpoint", "1"])))
        mc4 = mcmc_steps_per_s(args.workload, 600, ref_steps=0, rank=rank, world=world, dist=dist if world > 1 else None,
                               variants=(("gpu", []), ("gpu_dense_esc", ["--dense_esc"]),
                                         ("gpu_sparse_esc", ["--sparse_esc"]),
                                         ("gpu_whole_matrices", ["--checkpoint", "1"]),
                                         ("gpu_reference_schedule", ["--checkpoint", "1", "--dense_esc", "--eager_refill"]))
                               if world == 1 else
                               (("gpu", []), ("gpu_dense_esc", ["--dense_esc"]), ("gpu_whole_matrices", ["--checkpoint", "1"])))
        if rank == 0:
            line["mcmc"], line["mcmc_c4"] = mc, mc4
    if rank == 0:
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def build_problem_cpu(name, ncpu):
    """Reference arm only: a problem for ref_bench on a sample of the workload's queries, built WITHOUT the product
    library: distances and closest-L lists come from the compiled reference (oracle/_ref/libmcq_ref.so, else the C
    oracle), Hamming rows over host threads, the tie shuffles in query order."""
    from concurrent.futures import ThreadPoolExecutor
    cfg = workload_dims(name)
    L = cfg["L"]
    data = make_data(name)
    nq = min(cfg["Q"], max(CPU_SAMPLE_QUERIES, ncpu * 4))
    refs, qs, hla = data["ref_seqs"], np.ascontiguousarray(data["query_seqs"][:nq]), data["hla"][:nq]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_libs
    orc = cpu_libs.Checker("ref" if cpu_libs.have_ref() else "oracle")
    skip = name in synth.SELF_REFERENCE
    cpu_libs.srandom(SEED & 0x7fffffff)
    closest = np.zeros((nq, L), np.int32)
    with ThreadPoolExecutor(max_workers=ncpu) as ex:      # ctypes releases the GIL: one Hamming row per thread
        for q0 in range(0, nq, 4 * ncpu):
            rows = list(ex.map(lambda q: orc.hamming(refs, qs[q]), range(q0, min(nq, q0 + 4 * ncpu))))
            for i, ham in enumerate(rows):
                c = orc.closest(ham, L + 1 if skip else L)
                closest[q0 + i] = c[1:] if skip else c
    p = prep.build_problem(refs, qs, hla, closest)
    prep.random_parameters(p, seed=1)
    p.seqs = (refs, qs)
    return p


if __name__ == "__main__":
    sys.exit(main())
