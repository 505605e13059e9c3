"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard planning, query broadcast, and the counting merge
rule of ks_shard_search_batch (shard.merge_positions, the numpy statement of the merge kernels) applied to per-shard
oracle results (no GPU in this container); the GPU test of the real path (NCCL + kernels) is test_gpu_multi.py."""
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


def _unit_offsets(units, n_units):
    """exclusive offsets per unit (length n_units + 1) of rows that are ordered by unit"""
    return np.concatenate([[0], np.cumsum(np.bincount(np.asarray(units, dtype=np.int64), minlength=n_units))]).astype(np.int64)


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
        nq = len(qoffs) - 1
        bounds = shard.plan_shards(offs, world)
        sres, soffs = shard.shard_of(res, offs, bounds, rank)
        rows, hits = _oracle_shard_pairs(O, sres, soffs, qres, qoffs, k, moltype, scaled)
        # what ks_shard_search_batch does on the GPUs, with the oracle as the per-shard search: every shard's rows (protein
        # ids made index-wide) reach rank 0 and land at the positions of the counting merge (shard.merge_positions is the
        # numpy statement of merge_pairs_kernel / merge_hits_kernel)
        rows = [dict(r, pid=r["pid"] + bounds[rank]) for r in rows]
        hits = [(h[0], h[1] + bounds[rank], h[2], h[3], h[4]) for h in hits]
        all_rows = [None] * world if rank == 0 else None
        all_hits = [None] * world if rank == 0 else None
        dist.gather_object(rows, all_rows, dst=0)
        dist.gather_object(hits, all_hits, dst=0)
        ok = True
        if rank == 0:
            dst = shard.merge_positions([_unit_offsets([r["qid"] for r in rs], nq) for rs in all_rows])
            merged = [None] * sum(len(rs) for rs in all_rows)
            for rs, d in zip(all_rows, dst):
                for r, j in zip(rs, d.tolist()):
                    assert merged[j] is None
                    merged[j] = r
            full_rows, full_hits = _oracle_shard_pairs(O, res, offs, qres, qoffs, k, moltype, scaled)
            ok = len(full_rows) == len(merged) and len(full_rows) > 0
            for m, r in zip(merged, full_rows):
                ok &= (m["qid"], m["pid"]) == (r["qid"], r["pid"]) and m["intersect_hashes"] == r["intersect_hashes"]
                ok &= m["total_weighted_hashes"] == r["total_weighted_hashes"]
                for c in shard.PAIR_F64:
                    ok &= m[c] == r[c]
            # hits: the same rule per (query, window) unit
            qo = np.asarray(qoffs, dtype=np.int64)
            n_units = int(qo[-1])
            dsth = shard.merge_positions([_unit_offsets([qo[h[0]] + h[3] for h in hs], n_units) for hs in all_hits])
            mh = [None] * sum(len(hs) for hs in all_hits)
            for hs, d in zip(all_hits, dsth):
                for h, j in zip(hs, d.tolist()):
                    mh[j] = h
            ok &= mh == full_hits and len(mh) > 0
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
