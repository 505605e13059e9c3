"""CPU oracle for the kmerseek sketch-and-search hot path (numpy + the C restatement).

TEST INFRASTRUCTURE ONLY -- see the header of kmerseek_oracle.c.  The product package
(kmerseek_b200/) never imports this module; tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.

Parity status: PINNED against the reference's golden vectors (tests/golden/, see
tests/test_oracle_golden.py).  File:line citations are relative to /root/reference.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmerseek_oracle.so")
_SRC = os.path.join(_HERE, "kmerseek_oracle.c")
MOLTYPES = {"protein": 0, "raw": 0, "dayhoff": 1, "hp": 2}
SEED = 42  # src/rust/signature.rs:12
_lib = None


def build(force=False):
    """gcc the C restatement into oracle/libkmerseek_oracle.so (git-ignored, travels to the GPU box)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-pthread", "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p, u32p, u64p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint32, ctypes.c_uint64))
        L.kso_murmur64.restype = ctypes.c_uint64
        L.kso_murmur64.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64]
        L.kso_max_hash.restype = ctypes.c_uint64
        L.kso_max_hash.argtypes = [ctypes.c_uint32]
        L.kso_translate.restype = ctypes.c_uint8
        L.kso_translate.argtypes = [ctypes.c_uint8, ctypes.c_int]
        L.kso_normalize.restype = ctypes.c_int64
        L.kso_normalize.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                    ctypes.c_char_p, u8p, u64p]
        L.kso_sketch_tuples.restype = ctypes.c_uint64
        L.kso_sketch_tuples.argtypes = [u8p, u64p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                        u64p, u32p, u32p, ctypes.c_uint64]
        L.kso_cpu_baseline.restype = ctypes.c_uint64
        L.kso_cpu_baseline.argtypes = [u8p, u64p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_uint64, u64p, u64p]
        L.kso_allpairs_intersect.restype = None
        L.kso_allpairs_intersect.argtypes = [u64p, u64p, ctypes.c_uint64, u64p, u64p, ctypes.c_uint64, u32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


# ---------------------------------------------------------------------------------------------
# scalar pieces
# ---------------------------------------------------------------------------------------------
def murmur64(data: bytes, seed: int = SEED) -> int:
    """sourmash::_hash_murmur (src/rust/index.rs:766): MurmurHash3_x64_128 low word."""
    return lib().kso_murmur64(data, len(data), seed)


def murmur64_py(data: bytes, seed: int = SEED) -> int:
    """Pure-Python MurmurHash3_x64_128 low word; an independent cross-check of the C code."""
    M = (1 << 64) - 1
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & M

    def fmix(k):
        k ^= k >> 33
        k = (k * 0xFF51AFD7ED558CCD) & M
        k ^= k >> 33
        k = (k * 0xC4CEB9FE1A85EC53) & M
        return k ^ (k >> 33)

    h1 = h2 = seed
    n = len(data)
    for b in range(n // 16):
        k1 = int.from_bytes(data[16 * b:16 * b + 8], "little")
        k2 = int.from_bytes(data[16 * b + 8:16 * b + 16], "little")
        k1 = (rotl((k1 * c1) & M, 31) * c2) & M
        h1 = (rotl(h1 ^ k1, 27) + h2) & M
        h1 = (h1 * 5 + 0x52DCE729) & M
        k2 = (rotl((k2 * c2) & M, 33) * c1) & M
        h2 = (rotl(h2 ^ k2, 31) + h1) & M
        h2 = (h2 * 5 + 0x38495AB5) & M
    tail = data[(n // 16) * 16:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        h2 ^= (rotl((k2 * c2) & M, 33) * c1) & M
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        h1 ^= (rotl((k1 * c1) & M, 31) * c2) & M
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & M
    h2 = (h2 + h1) & M
    h1, h2 = fmix(h1), fmix(h2)
    return (h1 + h2) & M


def max_hash(scaled: int) -> int:
    return lib().kso_max_hash(scaled)


def translate(seq: str, moltype: str) -> str:
    """src/rust/encoding.rs:92-105 with the sourmash tables."""
    m = MOLTYPES[moltype]
    return "".join(chr(lib().kso_translate(ord(c), m)) for c in seq)


class InvalidAminoAcid(ValueError):
    """src/rust/errors.rs:14-15"""

    def __init__(self, char, pos):
        super().__init__(f"Invalid amino acid '{char}' found at position {pos}")
        self.char, self.pos = char, pos


def normalize(seq: str, protein_index: int = 0, ambig_seed: int = 0, mode: str = "kmerseek") -> str:
    """to_uppercase (src/rust/index.rs:1000) + validate_and_resolve (src/rust/aminoacid.rs:74-105).
    mode "sourmash": what the `kmerseek search` path does to its inputs (src/python/kmerseek/sketch.py:28-40 -> sourmash
    add_protein): upper-case only."""
    if mode == "sourmash":
        return "".join(chr(ord(c) - 32) if "a" <= c <= "z" else c for c in seq)
    raw = seq.encode()
    out = ctypes.create_string_buffer(len(raw) + 1)
    bc, bp = ctypes.c_uint8(0), ctypes.c_uint64(0)
    n = lib().kso_normalize(raw, len(raw), protein_index, ambig_seed, out, ctypes.byref(bc), ctypes.byref(bp))
    if n < 0:
        raise InvalidAminoAcid(chr(bc.value), bp.value)
    return out.raw[:n].decode()


def pack(seqs):
    """Concatenate already-normalised sequences into (residues u8[N], offsets u64[P+1])."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    residues = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).copy()
    return residues, offsets


def read_fasta(path):
    """needletail semantics used at src/rust/index.rs:920-935: id = whole header line, sequence lines joined."""
    import gzip
    op = gzip.open if str(path).endswith(".gz") else open
    names, seqs, cur = [], [], None
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                names.append(line[1:])
                cur = []
                seqs.append(cur)
            elif cur is not None:
                cur.append(line)
    return names, ["".join(s) for s in seqs]


# ---------------------------------------------------------------------------------------------
# sketch
# ---------------------------------------------------------------------------------------------
def sketch_tuples(residues, offsets, k, moltype, scaled):
    """Every kept (hash, pid, pos) in (pid, pos) order -- kso_sketch_tuples."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_prot = len(offsets) - 1
    if len(residues) == 0:
        residues = np.zeros(1, dtype=np.uint8)
    lens = (offsets[1:] - offsets[:-1]).astype(np.int64)
    cap = int(np.maximum(lens - k + 1, 0).sum())
    h = np.empty(max(cap, 1), dtype=np.uint64)
    pid = np.empty(max(cap, 1), dtype=np.uint32)
    pos = np.empty(max(cap, 1), dtype=np.uint32)
    n = lib().kso_sketch_tuples(_p(residues, ctypes.c_uint8), _p(offsets, ctypes.c_uint64), n_prot, k,
                                MOLTYPES[moltype], scaled, _p(h, ctypes.c_uint64), _p(pid, ctypes.c_uint32),
                                _p(pos, ctypes.c_uint32), cap)
    assert n <= cap
    return h[:n].copy(), pid[:n].copy(), pos[:n].copy()


def protein_sketches(h, pid, n_prot):
    """Per-protein (mins sorted distinct, abunds) -- KmerMinHash state after add_protein
    (src/rust/signature.rs:273-274)."""
    order = np.lexsort((h, pid))
    hs, ps = h[order], pid[order]
    out = []
    bounds = np.searchsorted(ps, np.arange(n_prot + 1, dtype=np.uint64))
    for p in range(n_prot):
        seg = hs[bounds[p]:bounds[p + 1]]
        mins, abunds = np.unique(seg, return_counts=True)
        out.append((mins.astype(np.uint64), abunds.astype(np.uint64)))
    return out


def signature_id(mins) -> str:
    """kmerseek's id string: hex of the wrapping sum of mins (src/rust/signature.rs:277-279)."""
    s = 0
    for m in mins:
        s = (s + int(m)) & ((1 << 64) - 1)
    return format(s, "x")


def md5sum(mins, k) -> str:
    """sourmash KmerMinHash::md5sum: MD5 of ascii(3k) then ascii(each min)."""
    m = hashlib.md5()
    m.update(str(3 * k).encode())
    for x in mins:
        m.update(str(int(x)).encode())
    return m.hexdigest()


def combined_sketch(sketches):
    """combined_minhash after store_signatures (src/rust/index.rs:802-827): multiset union with abundance sum."""
    if not sketches:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    allm = np.concatenate([s[0] for s in sketches])
    alla = np.concatenate([s[1] for s in sketches])
    mins, inv = np.unique(allm, return_inverse=True)
    ab = np.zeros(len(mins), dtype=np.uint64)
    np.add.at(ab, inv, alla)
    return mins, ab


def kmer_infos(seq, k, moltype, scaled):
    """kmer_infos of one protein (src/rust/index.rs:749-786): hash -> (encoded, {orig: [pos...]})."""
    res, offs = pack([seq])
    h, _, pos = sketch_tuples(res, offs, k, moltype, scaled)
    out = {}
    for hv, p in zip(h.tolist(), pos.tolist()):
        orig = seq[p:p + k]
        ent = out.setdefault(hv, (translate(orig, moltype), {}))
        ent[1].setdefault(orig, []).append(p)
    return out


# ---------------------------------------------------------------------------------------------
# index + search
# ---------------------------------------------------------------------------------------------
def build_index(h, pid, pos):
    """Sorted-hash CSR index: tuples ordered by (hash, pid, pos); unique keys, row_ptr, payload."""
    order = np.lexsort((pos, pid, h))
    hs, ps, qs = h[order], pid[order], pos[order]
    keys, start = np.unique(hs, return_index=True)
    row_ptr = np.append(start, len(hs)).astype(np.uint64)
    return keys, row_ptr, ps, qs


def hits(qh, qid, qpos, th, tpid, tpos):
    """Hit positions: inner join of query and target kept windows on the hash
    (src/python/kmerseek/search.py:204-213).  Rows (qid, pid, hash, qpos, tpos)
    in the boundary's canonical order: (qid, qpos, pid, tpos)."""
    keys, row_ptr, ps, ts = build_index(th, tpid, tpos)
    out = []
    idx = np.searchsorted(keys, qh)
    for i in range(len(qh)):
        j = idx[i]
        if j < len(keys) and keys[j] == qh[i]:
            a, b = int(row_ptr[j]), int(row_ptr[j + 1])
            for t in range(a, b):
                out.append((int(qid[i]), int(ps[t]), int(qh[i]), int(qpos[i]), int(ts[t])))
    out.sort(key=lambda r: (r[0], r[3], r[1], r[4]))
    return out


MANYSEARCH_COLUMNS = [
    "query_name", "query_md5", "match_name", "containment", "intersect_hashes", "ksize", "scaled", "moltype",
    "match_md5", "jaccard", "max_containment", "average_abund", "median_abund", "std_abund",
    "query_containment_ani", "match_containment_ani", "average_containment_ani", "max_containment_ani",
    "n_weighted_found", "total_weighted_hashes", "containment_target_in_query", "f_weighted_target_in_query",
]  # tests/test_search.py:33


def manysearch(q_sketches, t_sketches, k, scaled, moltype, q_names=None, t_names=None):
    """branchwater do_manysearch as called at src/python/kmerseek/search.py:125-141 (threshold 0,
    abundance on).  Formulae: SURVEY App. A.6.  Returns rows (dicts) ordered by (query, target)."""
    rows = []
    K = float(3 * k)
    for qi, (qm, _qa) in enumerate(q_sketches):
        if len(qm) == 0:
            continue
        for ti, (tm, ta) in enumerate(t_sketches):
            common, _, t_idx = np.intersect1d(qm, tm, assume_unique=True, return_indices=True)
            I = len(common)
            if I == 0 or len(tm) == 0:
                continue
            A = ta[t_idx].astype(np.float64)
            cont = I / len(qm)
            cont_t = I / len(tm)
            qani = cont ** (1.0 / K)
            mani = cont_t ** (1.0 / K)
            total_w = float(ta.sum())
            rows.append({
                "qid": qi, "pid": ti,
                "query_name": q_names[qi] if q_names else str(qi),
                "query_md5": md5sum(qm, k),
                "match_name": t_names[ti] if t_names else str(ti),
                "containment": cont, "intersect_hashes": I, "ksize": 3 * k, "scaled": scaled, "moltype": moltype,
                "match_md5": md5sum(tm, k),
                "jaccard": I / (len(qm) + len(tm) - I),
                "max_containment": max(cont, cont_t),
                "average_abund": float(A.sum() / I),
                "median_abund": float(np.median(A)),
                "std_abund": float(np.std(A)),
                "query_containment_ani": qani, "match_containment_ani": mani,
                "average_containment_ani": (qani + mani) / 2.0, "max_containment_ani": max(qani, mani),
                "n_weighted_found": int(A.sum()), "total_weighted_hashes": int(total_w),
                "containment_target_in_query": cont_t,
                "f_weighted_target_in_query": float(A.sum() / total_w),
            })
    return rows


def manysearch_indexed(q_sketches, th, tpid, k, scaled, moltype):
    """The rows of manysearch() for large target sets: the same formulae, but the overlaps come from a
    sorted-hash index over the target tuples (th, tpid: every kept window's hash and protein) instead of from all-pairs
    list intersections -- 1 000 queries against 10 M residues finish in seconds.  tests/test_oracle_golden.py checks it
    against manysearch() itself.  Returns rows (qid, pid, numeric columns) ordered by (query, target)."""
    th = np.asarray(th, dtype=np.uint64)
    tpid = np.asarray(tpid, dtype=np.int64)
    n_prot = int(tpid.max()) + 1 if len(tpid) else 0
    # per (hash, protein): abundance; per protein: distinct hashes |T| and kept windows
    order = np.lexsort((tpid, th))
    hs, ps = th[order], tpid[order]
    head = np.ones(len(hs), dtype=bool)
    head[1:] = (hs[1:] != hs[:-1]) | (ps[1:] != ps[:-1])
    g_start = np.flatnonzero(head)
    g_hash, g_pid = hs[g_start], ps[g_start]
    g_abund = np.diff(np.append(g_start, len(hs)))
    t_size = np.bincount(g_pid, minlength=n_prot)
    t_abund = np.bincount(tpid, minlength=n_prot)
    keys, k_start = np.unique(g_hash, return_index=True)
    k_end = np.append(k_start[1:], len(g_hash))
    rows = []
    K = float(3 * k)
    for qi, (qm, _qa) in enumerate(q_sketches):
        if len(qm) == 0:
            continue
        at = np.searchsorted(keys, qm)
        ok = (at < len(keys))
        ok[ok] &= keys[at[ok]] == qm[ok]
        if not ok.any():
            continue
        segs = [np.arange(k_start[j], k_end[j]) for j in at[ok]]
        g = np.concatenate(segs)
        pids, ab = g_pid[g], g_abund[g].astype(np.float64)
        o = np.lexsort((ab, pids))
        pids, ab = pids[o], ab[o]
        bounds = np.flatnonzero(np.append(True, pids[1:] != pids[:-1]))
        ends = np.append(bounds[1:], len(pids))
        for a, b in zip(bounds, ends):
            pid = int(pids[a])
            A = ab[a:b]
            I = b - a
            cont, cont_t = I / len(qm), I / t_size[pid]
            qani, mani = cont ** (1.0 / K), cont_t ** (1.0 / K)
            tw = float(t_abund[pid])
            rows.append({
                "qid": qi, "pid": pid, "containment": cont, "intersect_hashes": int(I), "jaccard": I / (len(qm) + t_size[pid] - I),
                "max_containment": max(cont, cont_t), "average_abund": float(A.sum() / I), "median_abund": float(np.median(A)),
                "std_abund": float(np.std(A)), "query_containment_ani": qani, "match_containment_ani": mani,
                "average_containment_ani": (qani + mani) / 2.0, "max_containment_ani": max(qani, mani),
                "n_weighted_found": int(A.sum()), "total_weighted_hashes": int(tw), "containment_target_in_query": cont_t,
                "f_weighted_target_in_query": float(A.sum() / tw)})
    return rows


def allpairs_intersect(q_sketches, t_sketches):
    """manysearch's all-pairs merge in C (timing baseline)."""
    def csr(sk):
        ptr = np.zeros(len(sk) + 1, dtype=np.uint64)
        np.cumsum([len(s[0]) for s in sk], out=ptr[1:])
        m = np.concatenate([s[0] for s in sk]) if sk else np.zeros(0, np.uint64)
        return np.ascontiguousarray(m if len(m) else np.zeros(1, np.uint64), dtype=np.uint64), ptr
    qm, qp = csr(q_sketches)
    tm, tp = csr(t_sketches)
    out = np.zeros(len(q_sketches) * len(t_sketches), dtype=np.uint32)
    lib().kso_allpairs_intersect(_p(qm, ctypes.c_uint64), _p(qp, ctypes.c_uint64), len(q_sketches),
                                 _p(tm, ctypes.c_uint64), _p(tp, ctypes.c_uint64), len(t_sketches),
                                 _p(out, ctypes.c_uint32))
    return out.reshape(len(q_sketches), len(t_sketches))


def cpu_baseline(residues, offsets, k, moltype, scaled, faithful=True, n_threads=None, batch_size=1000):
    """Time-able CPU restatement of process_fasta's compute (kso_cpu_baseline).  Returns
    (residues processed, combined sketch size, recorded positions)."""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    uniq, kept = ctypes.c_uint64(0), ctypes.c_uint64(0)
    n = lib().kso_cpu_baseline(_p(residues, ctypes.c_uint8), _p(offsets, ctypes.c_uint64), len(offsets) - 1, k,
                               MOLTYPES[moltype], scaled, 1 if faithful else 0, n_threads, batch_size,
                               ctypes.byref(uniq), ctypes.byref(kept))
    return n, uniq.value, kept.value


# ---------------------------------------------------------------------------------------------
# k-mer stitching: src/python/kmerseek/search.py:37-121
# ---------------------------------------------------------------------------------------------
def _stitch(kmers, starts):
    out, prev = "", 0
    for i, (s, km) in enumerate(zip(starts, kmers)):
        if i == 0:
            out = km
        else:
            d = s - prev
            out += km[-d:] if d != 0 else km  # python's kmer[-0:] is the whole k-mer
        prev = s
    return out


def stitch_pair(rows):
    """rows: list of (qpos, tpos, q_kmer, t_kmer, encoded) for one (query, match) pair.
    Follows stitch_kmers_in_query_match_pair literally, quirks included (the query string is
    stitched with the MATCH starts, search.py:78)."""
    rows = sorted(rows, key=lambda r: r[0])
    q = _stitch([r[2] for r in rows], [r[1] for r in rows])
    a = _stitch([r[4] for r in rows], [r[0] for r in rows])
    m = _stitch([r[3] for r in rows], [r[1] for r in rows])
    assert len(q) == len(a) == len(m)
    qs, ms = min(r[0] for r in rows), min(r[1] for r in rows)
    return {"query_start": qs, "query_end": qs + len(q), "query": q, "match_start": ms,
            "match_end": ms + len(q), "match": m, "encoded": a, "length": len(q)}
