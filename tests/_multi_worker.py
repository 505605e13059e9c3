"""torchrun worker of tests/test_gpu_multi.py: protein-sharded build on every rank, replicated queries,
ks_shard_search_batch (NCCL all-gather of counts + grouped send/recv of result blocks + counting merge kernel on
rank 0), merged result checked against the oracle on the unsharded proteome."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kmerseek_b200 as K  # noqa: E402
from kmerseek_b200 import shard, synth  # noqa: E402

# (k, alphabet, scaled, KS_DENSE, expected build paths): the general path with the unstable partition, the dense k-mer
# space path (forced on: the shards are far below its coverage condition), and scaled > 1
CONFIGS = [(16, "dayhoff", 1, "0", (0, 3)), (24, "hp", 1, "1", (1,)), (7, "protein", 10, "0", (0, 3))]


def check_config(comm, rank, world, local, res, offs, qres, qoffs, k, moltype, scaled, dense, paths):
    os.environ["KS_DENSE"] = dense
    bounds = shard.plan_shards(offs, world)
    sres, soffs = shard.shard_of(res, offs, bounds, rank)
    idx = K.ProteomeIndex("shard", k, scaled, moltype, device=local)
    idx.add_proteome(K.Proteome.from_packed(sres, soffs))
    idx.finalize()
    ok = idx.stats()["build_path"] in paths
    out = shard.search_and_gather(idx, K.Proteome.from_packed(qres, qoffs), comm, pid_base=bounds[rank], hits=True,
                                  query_sketches=True)
    combined = shard.combined_minhash(idx)  # union of the shards' combined sketches (gather + merge on rank 0)
    n_rows = n_hits = 0
    if rank == 0:
        from oracle import oracle as O
        th, tpid, tpos = O.sketch_tuples(res, offs, k, moltype, scaled)
        qh, qid, qpos = O.sketch_tuples(qres, qoffs, k, moltype, scaled)
        qsk = O.protein_sketches(qh, qid, len(qoffs) - 1)
        rows = O.manysearch(qsk, O.protein_sketches(th, tpid, len(offs) - 1), k, scaled, moltype)
        ohits = O.hits(qh, qid, qpos, th, tpid, tpos)
        p, h = out["pairs"], out["hits"]
        ok &= len(rows) == len(p["pair_qid"]) and len(rows) > 0
        for j, r in enumerate(rows if ok else []):
            ok &= (int(p["pair_qid"][j]), int(p["pair_pid"][j])) == (r["qid"], r["pid"])
            ok &= int(p["intersect_hashes"][j]) == r["intersect_hashes"]
            ok &= int(p["n_weighted_found"][j]) == r["n_weighted_found"]
            ok &= int(p["total_weighted_hashes"][j]) == r["total_weighted_hashes"]
            for c in shard.PAIR_F64:
                ok &= abs(float(p[c][j]) - r[c]) <= 1e-6 * abs(r[c])
        mine = list(zip(h["hit_qid"].tolist(), h["hit_pid"].tolist(), h["hit_hash"].tolist(), h["hit_qpos"].tolist(),
                        h["hit_tpos"].tolist()))
        ok &= mine == ohits and len(mine) > 0
        for (m, a), (om, oa) in zip(out["query_sketches"], qsk):
            ok &= np.array_equal(m, om) and np.array_equal(a, oa)
        fmins, fab = np.unique(th, return_counts=True)
        ok &= np.array_equal(combined[0], fmins) and np.array_equal(combined[1], fab.astype(np.uint64))
        n_rows, n_hits = len(rows), len(ohits)
    else:
        ok &= "pairs" not in out
    idx.close()
    if rank == 0:
        print(f"multi-gpu check: world={world} {moltype} k={k} scaled={scaled} pairs={n_rows} hits={n_hits} ok={bool(ok)}",
              flush=True)
    return bool(ok)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = shard.Comm(local)
    res, offs = synth.proteome(400_000, 4321)
    qres = qoffs = None
    if rank == 0:
        qres, qoffs, _ = synth.queries(res, offs, 48, 91, min_len=40, max_len=160)
    qres, qoffs = shard.broadcast_queries(qres, qoffs)
    ok = True
    for cfg in CONFIGS:
        ok &= check_config(comm, rank, world, local, res, offs, qres, qoffs, *cfg)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi-gpu check: world={world} all ok={int(flag.item()) == 1}", flush=True)
    comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
