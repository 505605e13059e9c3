#!/usr/bin/env python3
"""Time one batched search (library device time vs wall) against a bench workload's index."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import kmerseek_b200 as K  # noqa: E402
from kmerseek_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="target_100m_dayhoff_k16_s1")
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--hits", action="store_true")
a = ap.parse_args()
cfg = bench.WORKLOADS[a.workload]
res, offs = synth.proteome(cfg["n_residues"], cfg["seed"])
prot = K.Proteome.from_packed(res, offs)
qres, qoffs, _ = synth.queries(res, offs, a.queries, 79)
queries = K.Proteome.from_packed(qres, qoffs)
with K.ProteomeIndex("probe", cfg["k"], cfg["scaled"], cfg["moltype"]) as idx:
    idx.add_proteome(prot)
    idx.finalize()
    for i in range(a.reps):
        t0 = time.perf_counter()
        r = K.search(idx, queries, hits=a.hits)
        wall = (time.perf_counter() - t0) * 1e3
        del r.pairs, r.hits
        r_ms = r.ms_device
        t1 = time.perf_counter()
        import ctypes as C
        from kmerseek_b200 import _ffi
        from kmerseek_b200.errors import check
        out = C.POINTER(_ffi.ks_search_result)()
        check(_ffi.lib().ks_search_batch(idx._h, queries._h, 0, C.byref(out)))
        t2 = time.perf_counter()
        _ffi.lib().ks_search_result_free(out)
        print(f"   raw ks_search_batch: {(t2 - t1) * 1e3:.2f} ms", flush=True)
        print(f"search {i}: wall {wall:.2f} ms, device {r_ms:.2f} ms", flush=True)
