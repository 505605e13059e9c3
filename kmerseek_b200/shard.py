"""Multi-GPU host logic: one process per GPU (torchrun), the proteome sharded by protein, queries replicated,
per-shard scored pairs / hit lists gathered to rank 0 with NCCL (torch.distributed is the plumbing).

The build has no data-path collective: every rank sketches, sorts and indexes its own contiguous protein
range.  A target protein lives on exactly one shard, so a (query, target) pair is scored entirely by the
owning rank (|T| and the abundances are local) and the merge is a concatenation ordered by (query, target):
no reduction is needed.
"""
import ctypes as C

import numpy as np

from . import _ffi
from .errors import check

PAIR_U32 = ["pair_qid", "pair_pid", "intersect_hashes", "q_size", "t_size"]
PAIR_U64 = ["n_weighted_found", "total_weighted_hashes"]
PAIR_F64 = list(_ffi.SCORE_COLUMNS)
HIT_U32 = ["hit_qid", "hit_pid", "hit_qpos", "hit_tpos"]
HIT_U64 = ["hit_hash"]


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


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def _device_columns(res_ptr, names, n, typestr, torch_dtype):
    import torch
    L = _ffi.lib()
    cols = []
    for name in names:
        if n == 0:
            cols.append(torch.zeros(0, dtype=torch_dtype, device="cuda"))
            continue
        p = L.ks_search_result_device_column(res_ptr, name.encode())
        t = torch.as_tensor(_DevArray(p, n, typestr), device="cuda")
        cols.append(t.view(torch_dtype) if t.dtype != torch_dtype else t)
    return torch.stack(cols) if cols else None


def gather_blocks(blocks, count):
    """blocks: {key: tensor [n_cols, count]} on every rank (same keys/dtypes) -> on rank 0 {key: [n_cols, total]}
    concatenated in rank order, plus the per-rank counts; None elsewhere.  1 all_gather (counts) + 1 gather per block."""
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


def _to_host(t):
    """Device tensor -> numpy through a pinned staging tensor (pageable D2H is several times slower)."""
    import torch
    if t.device.type == "cpu":
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def _base_vector(counts, pid_bases, device):
    import torch
    return torch.repeat_interleave(torch.tensor(pid_bases, dtype=torch.int32, device=device),
                                   torch.tensor(counts, dtype=torch.int64, device=device))


def merge_pairs(gathered, counts, pid_bases):
    """Rank-0 merge, on the device the blocks live on: globalise protein ids with each shard's base and order
    rows by (query, target).  Shards arrive in rank order = ascending protein ranges and every shard is already
    ordered by (query, target), so one stable sort by query is the whole merge."""
    import torch
    u32, u64, f64 = gathered["u32"], gathered["u64"], gathered["f64"]
    u32 = u32.clone()
    u32[1] += _base_vector(counts, pid_bases, u32.device)
    order = torch.sort(u32[0], stable=True).indices
    u32h = _to_host(u32[:, order].contiguous()).view(np.uint32)
    u64h = _to_host(u64[:, order].contiguous()).view(np.uint64)
    f64h = _to_host(f64[:, order].contiguous())
    cols = {n: u32h[i] for i, n in enumerate(PAIR_U32)}
    cols.update({n: u64h[i] for i, n in enumerate(PAIR_U64)})
    cols.update({n: f64h[i] for i, n in enumerate(PAIR_F64)})
    return cols


def merge_hits(gathered, counts, pid_bases):
    """Hits are ordered by (query, qpos, target, tpos) inside a shard; across shards targets ascend with the rank,
    so a stable sort by (query, qpos) merges them."""
    import torch
    h32, h64 = gathered["h32"].clone(), gathered["h64"]
    h32[1] += _base_vector(counts, pid_bases, h32.device)
    key = (h32[0].to(torch.int64) << 32) | (h32[2].to(torch.int64) & 0xffffffff)
    order = torch.sort(key, stable=True).indices
    h32h = _to_host(h32[:, order].contiguous()).view(np.uint32)
    h64h = _to_host(h64[:, order].contiguous()).view(np.uint64)
    cols = {n: h32h[i] for i, n in enumerate(HIT_U32)}
    cols["hit_hash"] = h64h[0]
    return cols


def search_and_gather(index, queries, pid_base=0, hits=False):
    """Search this rank's shard and gather to rank 0.  Returns on rank 0 a dict with `pairs` (and `hits`)
    as host column dicts with index-wide protein ids; elsewhere {"n_pairs": local count}."""
    import torch
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    L = _ffi.lib()
    flags = (_ffi.KS_SEARCH_HITS if hits else 0) | (_ffi.KS_SEARCH_DEVICE_ONLY if world > 1 else 0)
    index.finalize()
    out = C.POINTER(_ffi.ks_search_result)()
    check(L.ks_search_batch(index._h, queries._h, flags, C.byref(out)))
    try:
        r = out.contents
        if world == 1:
            from .search import _collect
            res = _collect(r, hits, owner=out)
            out = None  # ownership moved to the SearchResult
            if pid_base:
                res.pairs["pair_pid"] = res.pairs["pair_pid"] + np.uint32(pid_base)
                if hits:
                    res.hits["hit_pid"] = res.hits["hit_pid"] + np.uint32(pid_base)
            return {"n_pairs": res.n_pairs, "pairs": res.pairs, "hits": res.hits, "query_sketches": res.query_sketches,
                    "result": res}
        n = int(r.n_pairs)
        # integer columns travel as their signed twins (same bits); merge_* views them back as unsigned
        blocks = {"u32": _device_columns(out, PAIR_U32, n, "<i4", torch.int32),
                  "u64": _device_columns(out, PAIR_U64, n, "<i8", torch.int64),
                  "f64": _device_columns(out, PAIR_F64, n, "<f8", torch.float64)}
        bases = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(bases, torch.tensor([pid_base], dtype=torch.int64, device="cuda"))
        bases = [int(b.item()) for b in bases]
        gathered, counts = gather_blocks(blocks, n)
        result = {"n_pairs": n}
        if hits:
            nh = int(r.n_hits)
            hb = {"h32": _device_columns(out, HIT_U32, nh, "<i4", torch.int32),
                  "h64": _device_columns(out, HIT_U64, nh, "<i8", torch.int64)}
            hg, hcounts = gather_blocks(hb, nh)
        if dist.get_rank() == 0:
            result["pairs"] = merge_pairs(gathered, counts, bases)
            result["n_pairs"] = len(result["pairs"]["pair_qid"])
            if hits:
                result["hits"] = merge_hits(hg, hcounts, bases)
        return result
    finally:
        if out is not None:
            L.ks_search_result_free(out)


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

