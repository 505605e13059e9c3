#!/usr/bin/env python3
"""Benchmark of the kmerseek sketch-and-search hot path on B200 (contract: see the task brief, section 4).

One "step" = one index build of the workload's proteome that is already resident in HBM:
    fused sketch kernel -> sort by hash -> CSR build (ks_index_clear + ks_index_sketch_resident +
    ks_index_finalize through the C ABI).  The library picks the build path from the parameters and the data
    (`config.build_path`: dense k-mer space path for hp with k <= 24; general path with the unstable or the stable
    partition otherwise); the stage names in `roofline.stages` follow it.
`value` = residues/s over all ranks with inputs resident in HBM; `e2e` = the same metric through the public
host API with HOST buffers (pinned 5-bit packed residues + offsets -> H2D -> build -> stats read back) inside the timed
region.  A batched search of 10 000 planted query domains against the built index is timed beside it
(`search`: query x proteome residues/s).  Multi-GPU: the proteome is sharded by protein, one rank per GPU,
no data-path collective in the build (weak scaling: every rank builds a shard of the workload's size);
the search gathers per-shard pair lists to rank 0 over NCCL.

--impl reference times the CPU restatement of the reference's own algorithm (oracle/, the reference is
Rust and cannot be built in this image) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: Swiss-Prot-sized index build, hp k=24 scaled=1 (the config the metric is quoted on)
    "c2_swissprot_hp_k24_s1": dict(n_residues=200_000_000, k=24, moltype="hp", scaled=1, seed=20260102),
    # BASELINE.json configs[2] on one GPU: 10 000 planted query domains against the C2-sized proteome, dayhoff k=16
    "c3_search_dayhoff_k16_s1": dict(n_residues=200_000_000, k=16, moltype="dayhoff", scaled=1, seed=20260102),
    # the north_star target run
    "target_100m_dayhoff_k16_s1": dict(n_residues=100_000_000, k=16, moltype="dayhoff", scaled=1, seed=20260103),
    "c4_slice_protein_k7_s10": dict(n_residues=1_000_000_000, k=7, moltype="protein", scaled=10, seed=20260104),
    "small": dict(n_residues=5_000_000, k=24, moltype="hp", scaled=1, seed=20260102),
    # one eighth of the target run: what a rank holds when the 100 M-residue proteome is sharded over 8 GPUs
    "target_shard_12m_dayhoff_k16_s1": dict(n_residues=12_500_000, k=16, moltype="dayhoff", scaled=1, seed=20260103),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of the C2
# workload (profiles/r01_*): only meaningful for that workload, null otherwise.
TRAFFIC = {  # bytes per launch, C2 workload, profiles/r01_d_ncu_full_raw_rep_c2.csv, r01_e_ncu_full_raw_dense_*.csv
    "bucket_sort_rep_kernel (+ fused CSR write, directory)": 2_991_587_000 + 2_473_064_000,
    "dense_bucket_kernel (sort + CSR write)": 1_638_007_000 + 3_419_163_000,
    "sketch_dense_kernel (ranks + first scatter level)": 2_342_600_000 + 1_481_121_000,
    "sketch_quad_kernel": 205_893_000 + 2_934_864_000,
}


def algorithmic_bytes(n_res, n_prot, n_tuples, n_unique):
    """BASELINE.md section 3."""
    sketch = n_res + (n_prot + 1) * 8 + n_tuples * 16
    build = n_tuples * 16 + n_tuples * 8 + n_unique * 8 + (n_unique + 1) * 4
    return sketch, build


def make_workload(cfg, rank):
    from kmerseek_b200 import synth
    res, offs = synth.proteome(cfg["n_residues"], cfg["seed"] + 7919 * rank)
    return res, offs


def cpu_baseline_run(cfg, faithful, sample_residues, threads):
    from oracle import oracle as O
    from kmerseek_b200 import synth
    res, offs = synth.proteome(sample_residues, cfg["seed"])
    t0 = time.perf_counter()
    n, uniq, kept = O.cpu_baseline(res, offs, cfg["k"], cfg["moltype"], cfg["scaled"], faithful=faithful, n_threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, dt, uniq, kept


def run_reference(args, cfg, wname):
    """The reference arm: CPU restatement of the reference's own algorithm (two passes per protein, linear
    `contains`, sorted-Vec combined insert -- src/rust/index.rs:749-830), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.reference_sample
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, uniq, kept = cpu_baseline_run(cfg, True, sample, threads)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals]) * 1e3)
    sample_txt = (f"first {sample} residues of the synthetic workload; reference cost structure kept (the combined-"
                  f"sketch insert is O(U) per new hash, so residues/s falls as the sample grows)")
    line = {
        "impl": "reference", "metric": "residues/s sketched+indexed", "value": v, "unit": "residues/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wname, "k": cfg["k"], "moltype": cfg["moltype"], "scaled": cfg["scaled"],
                   "sample_residues": sample},
        "cpu_baseline": {"value": v, "unit": "residues/s", "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": v, "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_swissprot_hp_k24_s1", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--reference-sample", type=int, default=150_000)
    ap.add_argument("--cpu-fast-sample", type=int, default=20_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, cfg, args.workload)

    import torch
    import torch.distributed as dist
    import kmerseek_b200 as K
    from kmerseek_b200 import _ffi, shard, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (kmerseek_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: this rank's shard (weak scaling: each rank builds a shard of the workload's size) ----
    res, offs = make_workload(cfg, rank)
    prot = K.Proteome.from_packed(res, offs)
    n_res, n_prot = prot.n_residues, prot.n_proteins
    idx = K.ProteomeIndex("bench", cfg["k"], cfg["scaled"], cfg["moltype"], device=local)
    L = _ffi.lib()
    stream = torch.cuda.ExternalStream(L.ks_index_stream(idx._h), device=torch.device("cuda", local))
    chk = K.errors.check

    def build_resident():
        chk(L.ks_index_clear(idx._h))
        chk(L.ks_index_sketch_resident(idx._h))
        chk(L.ks_index_finalize(idx._h))

    def build_from_host():
        chk(L.ks_index_clear(idx._h))
        chk(L.ks_index_add_proteome(idx._h, prot._h))  # pinned host buffers -> HBM (chunked) overlapped with the sketch
        chk(L.ks_index_finalize(idx._h))
        return idx.stats()  # the step's result read back on the host

    def timed(fn, steps, per_step=None):
        """EXACTLY `steps` calls between barrier+synchronize on both sides; device time by CUDA events on the
        library's stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
            if per_step:
                per_step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    chk(L.ks_index_upload(idx._h, prot._h))
    for _ in range(W):
        build_resident()
    st0 = idx.stats()
    stage_ms = {"sketch": [], "partition": [], "bucket": [], "csr": []}

    def collect():
        s = idx.stats()
        stage_ms["sketch"].append(s["ms_sketch"]); stage_ms["partition"].append(s["ms_sort_partition"])
        stage_ms["bucket"].append(s["ms_sort_bucket"]); stage_ms["csr"].append(s["ms_csr"])

    ms_total = timed(build_resident, args.steps, per_step=collect)
    st1 = idx.stats()
    launches = sum(st1[k] - st0[k] for k in ("sketch_launches", "sort_launches", "csr_launches"))
    ms_step = ms_total / args.steps
    value = world * n_res / (ms_step * 1e-3)

    # ---- e2e: host buffers -> H2D -> build -> stats back, through the public API ----
    for _ in range(2):
        build_from_host()
    ms_e2e = timed(build_from_host, args.steps) / args.steps
    e2e_value = world * n_res / (ms_e2e * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    st = idx.stats()
    n_tuples, n_unique = st["n_tuples"], st["n_unique_hashes"]

    # ---- search: planted query domains against the built index; per-shard results gathered to rank 0 ----
    qres, qoffs, _ = synth.queries(res, offs, args.queries, 77 + 2) if rank == 0 or world == 1 else (None, None, None)
    if world > 1:
        qres, qoffs = shard.broadcast_queries(qres, qoffs)
    queries = K.Proteome.from_packed(qres, qoffs)
    q_residues = queries.n_residues
    search_ms, lib_ms = [], []
    n_pairs_total = n_hits_total = 0
    for i in range(W + args.steps):
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gathered = shard.search_and_gather(idx, queries, pid_base=0, hits=False)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(ms, 0.0)
        if world > 1:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        if i >= W:
            search_ms.append((ms, wall))
        if rank == 0:
            n_pairs_total = gathered["n_pairs"]
            if world == 1 and i >= W:
                lib_ms.append(float(gathered["result"].ms_device))
    s_ms = float(np.mean([x[1] for x in search_ms]))  # wall: includes the H2D of the queries, NCCL gather and D2H
    search_value = q_residues * (world * n_res) / (s_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage times: CUDA events the library records on its own stream) ----
    peak, peak_src = peaks()
    n_groups = st["n_groups"]
    ms = {k: float(np.mean(v)) for k, v in stage_ms.items()}
    alphabet = {"protein": 20.0, "dayhoff": 6.0, "hp": 2.0}[cfg["moltype"]]
    repeat_heavy = n_tuples > 0.25 * alphabet ** cfg["k"] / cfg["scaled"]  # the library's choice of bucket-sort variant
    if st["build_path"] in (0, 3):
        bucket_name = ("bucket_sort_rep_kernel" if repeat_heavy else "bucket_sort_bin_kernel") + " (+ fused CSR write, directory)"
        alg = {
            # BASELINE.md section 3; per launch = per step (every stage runs once per step)
            "sketch_quad_kernel": ("sketch", n_res + (n_prot + 1) * 8 + n_tuples * 16),
            ("partition (2 library onesweep passes)" if st["build_path"] == 0 else
             "pair_partition_kernel (second scatter level; the first is fused into the sketch kernel)"):
                ("partition", n_tuples * 16 * 2),                                                # one read + one write
            # one read of every tuple, one write of its payload (loc), plus the CSR arrays the kernel emits (keys, key_grp,
            # grp_start) and the bucket directory (about one entry per 4 tuples)
            bucket_name: ("bucket", n_tuples * 16 + n_tuples * 8 + n_groups * 4 + n_unique * 12 + (n_tuples // 4) * 4),
        }
    else:
        # dense k-mer space path (hp, small k): tuples are 8-byte rank keys.  The algorithmic bytes stay SURVEY 8(d)'s
        # (16 B per tuple out of the sketch, 16 B read + 8 B payload written by the build): what the path does not
        # move shows up as a higher achieved figure, which is the point of it.
        alg = {
            "sketch_dense_kernel (ranks + first scatter level)": ("sketch", n_res + (n_prot + 1) * 8 + n_tuples * 16),
            ("dense_partition_kernel (second scatter level)" if st["build_path"] == 1 else "library key sort (3 onesweep passes)"):
                ("partition", n_tuples * 8 * 2),
            ("dense_bucket_kernel (sort + CSR write)" if st["build_path"] == 1 else "dense_count_kernel + dense_write_kernel"):
                ("bucket", n_tuples * 16 + n_tuples * 8 + n_groups * 4 + n_unique * 12),
            "dir_kernel": ("csr", n_unique * 8 + (n_tuples // 4) * 4),
        }
    stages = {name: {"ms": ms[key], "algorithmic_bytes": int(b), "achieved_gbs": b / ms[key] / 1e6 if ms[key] > 0 else 0.0}
              for name, (key, b) in alg.items()}
    own = {k: v for k, v in stages.items() if "library" not in k}
    dom = max(own, key=lambda k: own[k]["ms"])
    sk_bytes, bd_bytes = algorithmic_bytes(n_res, n_prot, n_tuples, n_unique)
    roof = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": stages[dom]["achieved_gbs"] / peak,
            "traffic": TRAFFIC.get(dom) if args.workload == "c2_swissprot_hp_k24_s1" else None, "peak_source": peak_src,
            "stages": {k: {"ms": round(v["ms"], 4), "achieved_gbs": round(v["achieved_gbs"], 1),
                           "frac": round(v["achieved_gbs"] / peak, 4)} for k, v in stages.items()},
            "whole_step": {"algorithmic_bytes": sk_bytes + bd_bytes,
                           "achieved_gbs": (sk_bytes + bd_bytes) / ms_step / 1e6,
                           "frac": (sk_bytes + bd_bytes) / ms_step / 1e6 / peak}}

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, _, _ = cpu_baseline_run(cfg, True, args.reference_sample, threads)
        vf, dtf, _, _ = cpu_baseline_run(cfg, False, args.cpu_fast_sample, threads)
        cpu = {"value": v, "unit": "residues/s", "cores": threads, "kind": "port",
               "sample": f"first {args.reference_sample} residues, reference cost structure (two passes, linear contains, "
                         f"sorted-Vec combined insert; quadratic, so the rate depends on the sample), {dt:.1f} s",
               "fast_variant": {"value": vf, "unit": "residues/s", "cores": threads,
                                "sample": f"{args.cpu_fast_sample} residues, binary-search membership and no combined "
                                          f"insert (a linear-cost CPU variant), {dtf:.1f} s"}}

    line = {
        "metric": "residues/s sketched+indexed", "value": value, "unit": "residues/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "k": cfg["k"], "moltype": cfg["moltype"], "scaled": cfg["scaled"],
                   "residues_per_gpu": n_res, "proteins_per_gpu": n_prot, "tuples_per_gpu": n_tuples,
                   "unique_hashes_per_gpu": n_unique, "parallelism": f"protein-sharded x{world}",
                   "build_path": {0: "general", 1: "dense k-mer space", 2: "dense k-mer space, library key sort",
                                  3: "general, unstable partition"}[st["build_path"]],
                   "l2": "inputs (0.2 GB residues, 3 GB tuples) exceed the 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": "residues/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int((n_res + 7) // 8 * 5 + 72 + (n_prot + 1) * 8), "d2h_bytes_per_step": 8 + 16 + 136},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "search": {"metric": "query x proteome residues/s searched", "value": search_value, "unit": "residue pairs/s",
                   "ms_per_batch_wall": s_ms, "ms_per_batch_device": float(np.mean([x[0] for x in search_ms])),
                   "ms_per_batch_kernels": float(np.mean(lib_ms)) if lib_ms else None,  # the library's own events: query
                                                                                       # sketch .. scores, no copies
                   "queries": args.queries, "query_residues": int(q_residues), "pairs": int(n_pairs_total),
                   "includes": "query H2D, sketch, lookup, aggregation, scores, NCCL gather to rank 0, D2H"},
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
