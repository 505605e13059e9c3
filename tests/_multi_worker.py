"""torchrun worker of tests/test_gpu_multi.py: protein-sharded build on every rank, replicated queries,
NCCL gather to rank 0, merged result checked against the oracle on the unsharded proteome."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import kmerseek_b200 as K  # noqa: E402
from kmerseek_b200 import shard, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k, moltype, scaled = 16, "dayhoff", 1
    res, offs = synth.proteome(400_000, 4321)
    qres = qoffs = None
    if rank == 0:
        qres, qoffs, _ = synth.queries(res, offs, 48, 91, min_len=40, max_len=160)
    qres, qoffs = shard.broadcast_queries(qres, qoffs)
    bounds = shard.plan_shards(offs, world)
    sres, soffs = shard.shard_of(res, offs, bounds, rank)
    idx = K.ProteomeIndex("shard", k, scaled, moltype, device=local)
    idx.add_proteome(K.Proteome.from_packed(sres, soffs))
    idx.finalize()
    out = shard.search_and_gather(idx, K.Proteome.from_packed(qres, qoffs), pid_base=bounds[rank], hits=True)
    ok = True
    combined = shard.combined_minhash(idx)  # union of the shards' combined sketches (NCCL gather + merge on rank 0)
    if rank == 0:
        from oracle import oracle as O
        th, tpid, tpos = O.sketch_tuples(res, offs, k, moltype, scaled)
        qh, qid, qpos = O.sketch_tuples(qres, qoffs, k, moltype, scaled)
        rows = O.manysearch(O.protein_sketches(qh, qid, len(qoffs) - 1), O.protein_sketches(th, tpid, len(offs) - 1),
                            k, scaled, moltype)
        ohits = O.hits(qh, qid, qpos, th, tpid, tpos)
        p, h = out["pairs"], out["hits"]
        ok &= len(rows) == len(p["pair_qid"]) and len(rows) > 0
        for j, r in enumerate(rows):
            ok &= (int(p["pair_qid"][j]), int(p["pair_pid"][j])) == (r["qid"], r["pid"])
            ok &= int(p["intersect_hashes"][j]) == r["intersect_hashes"]
            ok &= int(p["total_weighted_hashes"][j]) == r["total_weighted_hashes"]
            ok &= abs(float(p["containment"][j]) - r["containment"]) <= 1e-6 * r["containment"]
            ok &= abs(float(p["max_containment_ani"][j]) - r["max_containment_ani"]) <= 1e-6 * r["max_containment_ani"]
        mine = list(zip(h["hit_qid"].tolist(), h["hit_pid"].tolist(), h["hit_hash"].tolist(), h["hit_qpos"].tolist(),
                        h["hit_tpos"].tolist()))
        ok &= mine == ohits and len(mine) > 0
        fmins, fab = np.unique(th, return_counts=True)
        ok &= np.array_equal(combined[0], fmins) and np.array_equal(combined[1], fab.astype(np.uint64))
        print(f"multi-gpu check: world={world} pairs={len(rows)} hits={len(ohits)} ok={bool(ok)}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    idx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
