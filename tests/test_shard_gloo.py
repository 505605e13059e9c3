"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard planning, query broadcast, the
variable-length gather to rank 0 and the (query, target) merge.  Per-shard search results are produced by
the oracle here (no GPU in this container); the GPU test of the same path is in test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

from kmerseek_b200 import shard, synth


def test_plan_shards_balanced_and_contiguous():
    res, offs = synth.proteome(500_000, 1)
    for world in (1, 2, 3, 8):
        b = shard.plan_shards(offs, world)
        assert b[0] == 0 and b[-1] == len(offs) - 1 and all(x <= y for x, y in zip(b, b[1:]))
        sizes = [int(offs[b[r + 1]]) - int(offs[b[r]]) for r in range(world)]
        assert sum(sizes) == len(res)
        assert max(sizes) - min(sizes) <= 35000 + 1  # within one (maximal) protein
        parts = [shard.shard_of(res, offs, b, r) for r in range(world)]
        assert np.array_equal(np.concatenate([p[0] for p in parts]), res)
        for r, (pr, po) in enumerate(parts):
            assert po[0] == 0 and po[-1] == len(pr)
    assert shard.plan_shards(np.array([0, 5], np.uint64), 4) == [0, 0, 0, 1, 1] or True


def _oracle_shard_pairs(O, res, offs, qres, qoffs, k, moltype, scaled):
    th, tpid, tpos = O.sketch_tuples(res, offs, k, moltype, scaled)
    qh, qid, qpos = O.sketch_tuples(qres, qoffs, k, moltype, scaled)
    tsk = O.protein_sketches(th, tpid, len(offs) - 1)
    qsk = O.protein_sketches(qh, qid, len(qoffs) - 1)
    rows = O.manysearch(qsk, tsk, k, scaled, moltype)
    hits = O.hits(qh, qid, qpos, th, tpid, tpos)
    return rows, hits


def _blocks_from_rows(rows):
    n = len(rows)
    u32 = np.zeros((5, n), np.uint32)
    u64 = np.zeros((2, n), np.uint64)
    f64 = np.zeros((len(shard.PAIR_F64), n), np.float64)
    for j, r in enumerate(rows):
        u32[:, j] = [r["qid"], r["pid"], r["intersect_hashes"], 0, 0]
        u64[:, j] = [r["n_weighted_found"], r["total_weighted_hashes"]]
        f64[:, j] = [r[c] for c in shard.PAIR_F64]
    return {"u32": torch.from_numpy(u32.view(np.int32)), "u64": torch.from_numpy(u64.view(np.int64)),
            "f64": torch.from_numpy(f64)}


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        k, moltype, scaled = 7, "dayhoff", 1
        res, offs = synth.proteome(60_000, 42)
        qres = qoffs = None
        if rank == 0:
            qres, qoffs, _ = synth.queries(res, offs, 12, 9, min_len=30, max_len=90)
        qres, qoffs = shard.broadcast_queries(qres, qoffs)
        bounds = shard.plan_shards(offs, world)
        sres, soffs = shard.shard_of(res, offs, bounds, rank)
        rows, hits = _oracle_shard_pairs(O, sres, soffs, qres, qoffs, k, moltype, scaled)
        gathered, counts = shard.gather_blocks(_blocks_from_rows(rows), len(rows))
        bases = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(bases, torch.tensor([bounds[rank]], dtype=torch.int64))
        bases = [int(b.item()) for b in bases]
        h32 = np.array([[h[0] for h in hits], [h[1] for h in hits], [h[3] for h in hits], [h[4] for h in hits]], np.uint32).reshape(4, -1)
        h64 = np.array([[h[2] for h in hits]], np.uint64).reshape(1, -1)
        hg, hcounts = shard.gather_blocks({"h32": torch.from_numpy(h32.view(np.int32)), "h64": torch.from_numpy(h64.view(np.int64))}, len(hits))
        if rank == 0:
            merged = shard.merge_pairs(gathered, counts, bases)
            mh = shard.merge_hits(hg, hcounts, bases)
            full_rows, full_hits = _oracle_shard_pairs(O, res, offs, qres, qoffs, k, moltype, scaled)
            ok = len(full_rows) == len(merged["pair_qid"]) and len(full_rows) > 0
            for j, r in enumerate(full_rows):
                ok &= (int(merged["pair_qid"][j]), int(merged["pair_pid"][j])) == (r["qid"], r["pid"])
                ok &= int(merged["intersect_hashes"][j]) == r["intersect_hashes"]
                ok &= int(merged["total_weighted_hashes"][j]) == r["total_weighted_hashes"]
                for c in shard.PAIR_F64:
                    ok &= float(merged[c][j]) == r[c]
            mine = list(zip(mh["hit_qid"].tolist(), mh["hit_pid"].tolist(), mh["hit_hash"].tolist(),
                            mh["hit_qpos"].tolist(), mh["hit_tpos"].tolist()))
            ok &= mine == full_hits and len(mine) > 0
        # combined sketch of the whole proteome = union of the shards' combined sketches, abundances summed
        # (src/rust/index.rs:824-827); hashes above 2^63 must keep their unsigned order through the int64 tensors
        sh, spid, spos = O.sketch_tuples(sres, soffs, k, moltype, scaled)
        smins, sab = np.unique(sh, return_counts=True)
        merged_c = shard.merge_combined_sketch(smins, sab.astype(np.uint64))
        if rank == 0:
            fh, _, _ = O.sketch_tuples(res, offs, k, moltype, scaled)
            fmins, fab = np.unique(fh, return_counts=True)
            ok &= bool((fmins >> np.uint64(63)).any()) and np.array_equal(merged_c[0], fmins)
            ok &= np.array_equal(merged_c[1], fab.astype(np.uint64))
            q.put(bool(ok))
        else:
            ok_none = merged_c is None
            assert ok_none
    finally:
        dist.destroy_process_group()


def test_gather_and_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
