"""Times the closest-L path (packing, distances, selection, consensus) at a workload's shape:
python tools/closest_timing.py C3 [queries]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcqueen_b200 import api, synth  # noqa: E402

name = sys.argv[1]
cfg = bench.workload_dims(name)
data = bench.make_data(name)
refs, qs = data["ref_seqs"], data["query_seqs"]
if len(sys.argv) > 2:
    qs = np.ascontiguousarray(qs[:int(sys.argv[2])])
L = cfg["L"]
api.closest_batch(refs[:1000], qs[:2], 8)
for rep in range(3):
    t0 = time.time()
    cl = api.closest_batch(refs, qs, L, skip_first=name in synth.SELF_REFERENCE)
    t1 = time.time()
    ham, sel = api.closest_last_timing()
    cons = api.consensus_batch(refs, cl)
    t2 = time.time()
    print(f"{name}: {refs.shape[0]} refs x {qs.shape[0]} queries x {refs.shape[1]} nt, L={L}: hamming {ham:.2f} ms "
          f"({refs.shape[0] * qs.shape[0] * refs.shape[1] / ham / 1e9:.1f} Gnt-compares/ms... {refs.shape[0] * qs.shape[0] * refs.shape[1] / (ham * 1e-3):.3e}/s), "
          f"selection {sel:.2f} ms, closest wall {t1 - t0:.3f} s; consensus kernel {api.consensus_last_timing():.2f} ms, wall {t2 - t1:.3f} s",
          flush=True)
