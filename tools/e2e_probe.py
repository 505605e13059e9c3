#!/usr/bin/env python3
"""Host-side split of the end-to-end build (host buffers in, index out): wall time of add_proteome (upload pipelined with
the sketch) and of finalize, next to the device stage times.  python tools/e2e_probe.py [workload] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import kmerseek_b200 as K  # noqa: E402
from kmerseek_b200 import _ffi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2_swissprot_hp_k24_s1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = bench.WORKLOADS[wl]
res, offs = synth.proteome(cfg["n_residues"], cfg["seed"])
L = _ffi.lib()
chk = K.errors.check
prot = K.Proteome.from_packed(res, offs)
idx = K.ProteomeIndex("q", cfg["k"], cfg["scaled"], cfg["moltype"])
rows = []
for i in range(steps + 3):
    t0 = time.perf_counter()
    chk(L.ks_index_clear(idx._h))
    t1 = time.perf_counter()
    chk(L.ks_index_add_proteome(idx._h, prot._h))
    t2 = time.perf_counter()
    chk(L.ks_index_finalize(idx._h))
    t3 = time.perf_counter()
    if i >= 3:
        s = idx.stats()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, s["ms_upload"], s["ms_sketch"],
                     s["ms_sort_partition"], s["ms_sort_bucket"], s["ms_csr"]))
m = np.mean(rows, axis=0)
print(f"{wl}: clear {m[0]:.3f}  add {m[1]:.3f}  finalize {m[2]:.3f}  total {m[3]:.3f} ms | device: upload {m[4]:.3f} sketch {m[5]:.3f} "
      f"partition {m[6]:.3f} bucket {m[7]:.3f} csr {m[8]:.3f}", flush=True)
