#!/usr/bin/env python3
"""Kernel A/B helper: resident build time and stage times of one workload (python tools/quick_build_bench.py [workload] [steps]).
KS_LIB_PATH picks the library variant (tools/build_variants.sh)."""
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
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cfg = bench.WORKLOADS[wl]
cache = f"/dev/shm/ks_synth_{cfg['n_residues']}_{cfg['seed']}.npz"
if os.path.exists(cache):
    z = np.load(cache)
    res, offs = z["res"], z["offs"]
else:
    res, offs = synth.proteome(cfg["n_residues"], cfg["seed"])
    np.savez(cache, res=res, offs=offs)
if frac < 1.0:
    p = int((len(offs) - 1) * frac)
    res, offs = res[:int(offs[p])], offs[:p + 1]
L = _ffi.lib()
chk = K.errors.check
prot = K.Proteome.from_packed(res, offs)
idx = K.ProteomeIndex("q", cfg["k"], cfg["scaled"], cfg["moltype"])
chk(L.ks_index_upload(idx._h, prot._h))
ms = []
stages = []
for i in range(steps + 3):
    t0 = time.perf_counter()
    chk(L.ks_index_clear(idx._h)); chk(L.ks_index_sketch_resident(idx._h)); chk(L.ks_index_finalize(idx._h))
    chk(L.ks_index_sync(idx._h))
    dt = (time.perf_counter() - t0) * 1e3
    if i >= 3:
        ms.append(dt)
        s = idx.stats()
        stages.append((s["ms_sketch"], s["ms_sort_partition"], s["ms_sort_bucket"], s["ms_csr"]))
st = np.mean(stages, axis=0)
s = idx.stats()
print(f"{os.path.basename(os.environ.get('KS_LIB_PATH', 'default')):28s} {wl} x{frac}: wall {np.mean(ms):.3f} ms (min {np.min(ms):.3f})  sketch {st[0]:.3f}  partition {st[1]:.3f}  "
      f"bucket {st[2]:.3f}  dir {st[3]:.3f}  path {s['build_path']} U {s['n_unique_hashes']} G {s['n_groups']}", flush=True)
