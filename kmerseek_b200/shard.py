"""Multi-GPU host logic: one process per GPU (torchrun), the proteome sharded by protein, queries replicated.

The build has no data-path collective: every rank sketches, sorts and indexes its own contiguous protein
range.  A target protein lives on exactly one shard, so a (query, target) pair is scored entirely by the
owning rank (|T| and the abundances are local) and the merge is a concatenation ordered by (query, target):
no reduction is needed.  The search exchange itself lives below the C ABI
(ks_shard_search_batch: NCCL all-gather of counts, grouped send/recv of result blocks, a counting merge kernel on
rank 0); this module shards the input, replicates the queries and bootstraps the library's communicator.
"""
import ctypes as C

import numpy as np

from . import _ffi
from .errors import check

PAIR_F64 = list(_ffi.SCORE_COLUMNS)


def plan_shards(offsets, world):
    """Contiguous protein ranges balanced by residue count: rank r owns proteins [b[r], b[r+1])."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_prot = len(offsets) - 1
    total = int(offsets[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        p = int(np.searchsorted(offsets, target, side="left"))
        bounds.append(min(max(p, bounds[-1]), n_prot))
    bounds.append(n_prot)
    return bounds


def shard_of(residues, offsets, bounds, rank):
    """(residues, offsets) of rank's protein range, offsets rebased to 0."""
    a, b = bounds[rank], bounds[rank + 1]
    lo, hi = int(offsets[a]), int(offsets[b])
    return residues[lo:hi], (np.asarray(offsets[a:b + 1], dtype=np.uint64) - np.uint64(lo))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def broadcast_queries(qres, qoffs, device=None):
    """Replicate the query batch from rank 0 (2 broadcasts: sizes, payload)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return qres, qoffs
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    rank = dist.get_rank()
    sizes = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == 0:
        sizes[0], sizes[1] = len(qres), len(qoffs)
    dist.broadcast(sizes, 0)
    n_res, n_off = int(sizes[0]), int(sizes[1])
    buf = torch.zeros(n_res + 8 * n_off, dtype=torch.uint8, device=dev)
    if rank == 0:
        payload = np.concatenate([np.asarray(qres, dtype=np.uint8), np.asarray(qoffs, dtype=np.uint64).view(np.uint8)])
        buf.copy_(torch.from_numpy(payload))
    dist.broadcast(buf, 0)
    host = buf.cpu().numpy()
    return host[:n_res].copy(), host[n_res:].view(np.uint64).copy()


class Comm:
    """ks_comm: the library's own NCCL communicator (ncclCommInitRank inside libkmerseek_b200.so).  torch.distributed is
    only the side channel that hands rank 0's ncclUniqueId to the other ranks."""

    def __init__(self, device=None):
        import torch
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        self.device = torch.cuda.current_device() if device is None else device
        L = _ffi.lib()
        ident = (C.c_uint8 * _ffi.KS_COMM_ID_BYTES)()
        if self.rank == 0:
            check(L.ks_comm_unique_id(ident))
        if self.world > 1:
            dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            ident = (C.c_uint8 * _ffi.KS_COMM_ID_BYTES)(*t.cpu().tolist())
        h = C.c_void_p()
        check(L.ks_comm_create(ident, self.rank, self.world, self.device, C.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            _ffi.lib().ks_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_positions(pair_offs):
    """The counting merge that `merge_pairs_kernel` (csrc/search.cu) runs on rank 0, stated in numpy.
    pair_offs[s] = shard s's exclusive pair offsets per query (length nq + 1).  Shards hold ascending protein ranges and
    each is ordered by (query, target), so row j of shard s (its query q = the q with off_s[q] <= j < off_s[q + 1]) goes to
        sum_s' off_s'[q]  +  sum_{s' < s} (off_s'[q + 1] - off_s'[q])  +  (j - off_s[q]).
    Returns one destination array per shard.  No sort anywhere."""
    offs = [np.asarray(o, dtype=np.int64) for o in pair_offs]
    base = np.sum(offs, axis=0)  # merged exclusive offset of every query
    out = []
    before = np.zeros(len(base) - 1, dtype=np.int64)  # pairs of query q in the shards before s
    for o in offs:
        counts = np.diff(o)
        q = np.repeat(np.arange(len(counts)), counts)
        j = np.arange(int(o[-1]))
        out.append(base[q] + before[q] + (j - o[q]))
        before = before + counts
    return out


def search_and_gather(index, queries, comm=None, pid_base=0, hits=False, query_sketches=False):
    """ks_shard_search_batch: search this rank's shard and merge on rank 0 (one NCCL all-gather of the counts, one
    grouped send/recv of the shards' result blocks, a counting merge kernel).  Returns on rank 0 a dict with `pairs`
    (and `hits`) as host column dicts with index-wide protein ids, plus `result` (the SearchResult that owns them);
    elsewhere {"n_pairs": local count}."""
    from .search import _collect
    L = _ffi.lib()
    flags = (_ffi.KS_SEARCH_HITS if hits else 0) | (_ffi.KS_SEARCH_QUERY_SKETCHES if query_sketches else 0)
    index.finalize()
    out = C.POINTER(_ffi.ks_search_result)()
    if comm is None or comm.world == 1:
        check(L.ks_search_batch(index._h, queries._h, flags, C.byref(out)))
        rank = 0
    else:
        check(L.ks_shard_search_batch(index._h, comm._h, queries._h, flags, pid_base, C.byref(out)))
        rank = comm.rank
    if rank != 0:
        n = int(out.contents.n_pairs)
        L.ks_search_result_free(out)
        return {"n_pairs": n}
    try:
        res = _collect(out.contents, hits, owner=out)
    except Exception:
        L.ks_search_result_free(out)
        raise
    if (comm is None or comm.world == 1) and pid_base:
        res.pairs["pair_pid"] = res.pairs["pair_pid"] + np.uint32(pid_base)
        if hits:
            res.hits["hit_pid"] = res.hits["hit_pid"] + np.uint32(pid_base)
    return {"n_pairs": res.n_pairs, "pairs": res.pairs, "hits": res.hits, "query_sketches": res.query_sketches,
            "result": res}


def _to_host(t):
    """Device tensor -> numpy through a pinned staging tensor (pageable D2H is several times slower)."""
    import torch
    if t.device.type == "cpu":
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def gather_blocks(blocks, count):
    """blocks: {key: tensor [n_cols, count]} on every rank (same keys/dtypes) -> on rank 0 {key: [n_cols, total]}
    concatenated in rank order, plus the per-rank counts; None elsewhere.  (Combined-sketch merge only: it is not on the
    search path.)"""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return blocks, [count]
    world, rank = dist.get_world_size(), dist.get_rank()
    any_t = next(iter(blocks.values()))
    cnt = torch.tensor([count], dtype=torch.int64, device=any_t.device)
    counts = [torch.zeros(1, dtype=torch.int64, device=any_t.device) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    mx = max(counts + [1])
    out = {}
    for key, t in blocks.items():
        pad = torch.zeros((t.shape[0], mx), dtype=t.dtype, device=t.device)
        pad[:, :count] = t
        recv = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, recv, dst=0)
        if rank == 0:
            out[key] = torch.cat([r[:, :c] for r, c in zip(recv, counts)], dim=1)
    return (out if rank == 0 else None), counts


def merge_combined_sketch(mins, abunds, device=None):
    """Union of the shards' combined sketches with abundances summed (combined_minhash of the whole proteome,
    src/rust/index.rs:824-827): all_gather of the sizes, one padded gather of (hash, count) to
    rank 0, then sort + segmented sum there.  `mins` / `abunds`: this shard's sorted unique hashes and their counts
    (uint64).  Returns (mins, abunds) on rank 0, None elsewhere.  Not on the search path: the per-shard indexes never
    need it."""
    import torch
    dist = _dist()
    mins = np.ascontiguousarray(mins, dtype=np.uint64)
    abunds = np.ascontiguousarray(abunds, dtype=np.uint64)
    if dist is None or dist.get_world_size() == 1:
        return mins, abunds
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    block = torch.from_numpy(np.stack([mins, abunds]).view(np.int64)).to(device)
    gathered, counts = gather_blocks({"c": block}, len(mins))
    if dist.get_rank() != 0:
        return None
    g = gathered["c"]
    # uint64 order through int64 tensors: flip the sign bit, sort, flip back
    flipped = g[0] ^ torch.iinfo(torch.int64).min
    order = torch.argsort(flipped, stable=True)
    keys, inverse = torch.unique_consecutive(flipped[order], return_inverse=True)
    sums = torch.zeros(keys.shape[0], dtype=torch.int64, device=g.device).index_add_(0, inverse, g[1][order])
    keys = keys ^ torch.iinfo(torch.int64).min
    return _to_host(keys).view(np.uint64).copy(), _to_host(sums).view(np.uint64).copy()


def combined_minhash(index):
    """(mins, abunds) of the combined sketch over all shards, on rank 0 (`index` is this rank's shard)."""
    mins, abunds = index.get_combined_minhash()
    return merge_combined_sketch(mins, abunds)

