"""Search over a GPU-resident ProteomeIndex: the host side of ks_search_batch.

Mirrors what `kmerseek search` computes (src/python/kmerseek/search.py:125-141 -> branchwater manysearch,
:198-240 -> k-mer join and stitching) with the reference's column names (tests/test_search.py:33).
"""
import ctypes as C

import numpy as np

from . import _ffi
from .errors import check
from .index import Proteome, ProteomeIndex, _np, md5_of_mins, translate

MANYSEARCH_COLUMNS = [
    "query_name", "query_md5", "match_name", "containment", "intersect_hashes", "ksize", "scaled", "moltype",
    "match_md5", "jaccard", "max_containment", "average_abund", "median_abund", "std_abund",
    "query_containment_ani", "match_containment_ani", "average_containment_ani", "max_containment_ani",
    "n_weighted_found", "total_weighted_hashes", "containment_target_in_query", "f_weighted_target_in_query",
]

PAIR_INT_COLUMNS = {"pair_qid": np.uint32, "pair_pid": np.uint32, "intersect_hashes": np.uint32, "q_size": np.uint32,
                    "t_size": np.uint32, "n_weighted_found": np.uint64, "total_weighted_hashes": np.uint64}
HIT_COLUMNS = {"hit_qid": np.uint32, "hit_pid": np.uint32, "hit_qpos": np.uint32, "hit_tpos": np.uint32,
               "hit_hash": np.uint64}


class SearchResult:
    """A ks_search_result on the host: `pairs` (dict of columns, ordered by (query, target)), `hits`
    (dict of columns, ordered by (query, qpos, target, tpos)), `query_sketches` [(mins, abunds)].
    The pair / hit columns are numpy views of the library's pinned result block (no second copy of what can be
    a gigabyte); the block is released when this object is."""

    def __init__(self, pairs, hits, query_sketches, ms_device, owner=None, q_sizes=None, n_pairs=None, n_hits=None):
        self.pairs, self.hits, self.query_sketches, self.ms_device = pairs, hits, query_sketches, ms_device
        self._n_pairs, self._n_hits = n_pairs, n_hits
        self.q_sizes = q_sizes  # |Q| per query (always there; the sketches themselves only when asked for)
        self._owner = owner  # POINTER(ks_search_result) kept alive for the views

    def close(self):
        """Detach from the pinned block: columns become owned copies."""
        if self._owner is not None:
            self.pairs = {k: v.copy() for k, v in self.pairs.items()} if self.pairs else self.pairs
            self.hits = {k: v.copy() for k, v in self.hits.items()} if self.hits else self.hits
            q = self.query_sketches
            if isinstance(q, _QuerySketches):
                self.query_sketches = _QuerySketches(q._p, q._m.copy(), q._a.copy())
            self._owner = None

    @property
    def n_pairs(self):
        return self._n_pairs if self._n_pairs is not None else len(self.pairs["pair_qid"])

    @property
    def n_hits(self):
        if self._n_hits is not None:
            return self._n_hits
        return len(self.hits["hit_qid"]) if self.hits else 0


class _QuerySketches:
    """Sequence of (mins, abunds) per query over the concatenated arrays; a batch of 10 000 queries should not
    pay for 10 000 slice pairs it may never look at."""

    def __init__(self, sig_ptr, mins, abunds):
        self._p, self._m, self._a = sig_ptr, mins, abunds

    def __len__(self):
        return len(self._p) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        a, b = int(self._p[i]), int(self._p[i + 1])
        return self._m[a:b], self._a[a:b]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _Block:
    """Frees a ks_search_result when the last numpy view of its pinned block is gone."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr is not None:
                _ffi.lib().ks_search_result_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class _LazyColumns(dict):
    """Column dict whose numpy views of the pinned block are made on first use: a result has 19 pair columns (+ 5 hit
    columns) and a caller usually reads a few -- building all of them costs more host time than a small batch's kernels.
    Behaves like the plain dict it replaces (assignment, items(), iteration, `in`)."""

    def __init__(self, spec):
        super().__init__()
        self._spec = spec  # name -> (pointer, n, dtype, block)

    def __missing__(self, k):
        p, n, dt, blk = self._spec.pop(k)  # KeyError for an unknown column
        v = _view(p, n, dt, blk) if blk is not None else _np(p, n, dt)
        super().__setitem__(k, v)
        return v

    def __setitem__(self, k, v):
        self._spec.pop(k, None)
        super().__setitem__(k, v)

    def _all(self):
        for k in list(self._spec):
            self[k]
        return self

    def __contains__(self, k):
        return k in self._spec or super().__contains__(k)

    def __iter__(self):
        return super(_LazyColumns, self._all()).__iter__()

    def __len__(self):
        return len(self._spec) + super().__len__()

    def keys(self):
        return super(_LazyColumns, self._all()).keys()

    def values(self):
        return super(_LazyColumns, self._all()).values()

    def items(self):
        return super(_LazyColumns, self._all()).items()

    def get(self, k, default=None):
        return self[k] if k in self else default


def _view(ptr, n, dtype, block):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    ctype = {np.uint32: C.c_uint32, np.uint64: C.c_uint64, np.float64: C.c_double}[dtype]
    arr = (ctype * n).from_address(C.addressof(ptr.contents))
    arr._block = block  # the view's base keeps the block (and with it the pinned memory) alive
    return np.frombuffer(arr, dtype=dtype)


def _collect(r, want_hits, owner=None):
    """owner given: columns are views of the pinned block (owner is freed with the SearchResult); else copies."""
    block = _Block(owner) if owner is not None else None
    try:
        get = (lambda p, n, dt: _view(p, n, dt, block)) if owner is not None else _np
        np_, nh, nq = int(r.n_pairs), int(r.n_hits), int(r.n_queries)
        spec = {n: (getattr(r, n), np_, dt, block) for n, dt in PAIR_INT_COLUMNS.items()}
        for n in _ffi.SCORE_COLUMNS:
            spec[n] = (getattr(r, n), np_, np.float64, block)
        pairs = _LazyColumns(spec)
        hits = _LazyColumns({n: (getattr(r, n), nh, dt, block) for n, dt in HIT_COLUMNS.items()}) if want_hits else None
        if owner is None:  # copies: the caller frees the result right after this call
            pairs._all()
            if hits is not None:
                hits._all()
        sig_ptr = _np(r.q_sig_ptr, nq + 1, np.uint64)
        sketches = None
        if r.q_mins:  # KS_SEARCH_QUERY_SKETCHES
            E = int(sig_ptr[-1]) if nq else 0
            sketches = _QuerySketches(sig_ptr, get(r.q_mins, E, np.uint64), get(r.q_abunds, E, np.uint64))
        return SearchResult(pairs, hits, sketches, r.ms_device, block, q_sizes=np.diff(sig_ptr).astype(np.uint32),
                            n_pairs=np_, n_hits=nh if want_hits else 0)
    except Exception:
        if block is not None:
            block.ptr = None  # the caller frees the result on this path: exactly one owner of the free
        raise


def search(index: ProteomeIndex, queries: Proteome, hits=True, query_sketches=True) -> SearchResult:
    """One batched search of `queries` against a finalized index (ks_search_batch).  `query_sketches=False` leaves the
    queries' mins / abundances on the device (they are only needed for the query_md5 column)."""
    index.finalize()
    flags = (_ffi.KS_SEARCH_HITS if hits else 0) | (_ffi.KS_SEARCH_QUERY_SKETCHES if query_sketches else 0)
    out = C.POINTER(_ffi.ks_search_result)()
    check(_ffi.lib().ks_search_batch(index._h, queries._h, flags, C.byref(out)))
    try:
        return _collect(out.contents, hits, owner=out)
    except Exception:
        _ffi.lib().ks_search_result_free(out)
        raise


def manysearch_rows(result: SearchResult, index: ProteomeIndex, query_names, target_names=None, target_sketches=None,
                    targets: Proteome = None):
    """The 22-column manysearch table (tests/test_search.py:33) as a list of dicts.  `match_md5` needs the matched
    targets' sketches: pass `targets` (the indexed proteome) and only the matched proteins are sketched again, or
    `target_sketches`; otherwise every sketch of the index is exported."""
    target_names = target_names if target_names is not None else index.names()
    p = result.pairs
    if target_sketches is None and targets is not None:
        matched = np.unique(p["pair_pid"]).astype(np.int64)
        offs, res = targets.offsets.astype(np.int64), targets.residues
        lens = offs[matched + 1] - offs[matched]
        sub_offs = np.zeros(len(matched) + 1, dtype=np.uint64)
        np.cumsum(lens, out=sub_offs[1:])
        idx = np.repeat(offs[matched] - sub_offs[:-1].astype(np.int64), lens) + np.arange(int(sub_offs[-1]))
        sub = Proteome.from_packed(res[idx], sub_offs)
        target_sketches = dict(zip(matched.tolist(), index.sketch_proteome(sub)))
        sub.close()
    if target_sketches is None:
        target_sketches = index.export_sketches()
    rows = []
    qmd5 = {}
    for j in range(result.n_pairs):
        q, t = int(p["pair_qid"][j]), int(p["pair_pid"][j])
        if q not in qmd5:
            qmd5[q] = md5_of_mins(result.query_sketches[q][0], index.ksize)
        row = {
            "query_name": query_names[q], "query_md5": qmd5[q], "match_name": target_names[t],
            "intersect_hashes": int(p["intersect_hashes"][j]), "ksize": 3 * index.ksize, "scaled": index.scaled,
            "moltype": index.moltype, "match_md5": md5_of_mins(target_sketches[t][0], index.ksize),
            "n_weighted_found": int(p["n_weighted_found"][j]), "total_weighted_hashes": int(p["total_weighted_hashes"][j]),
        }
        for c in _ffi.SCORE_COLUMNS:
            row[c] = float(p[c][j])
        rows.append({c: row[c] for c in MANYSEARCH_COLUMNS})
    return rows


def _stitch(kmers, starts):
    # single_stitch_together_kmers, src/python/kmerseek/search.py:37-61
    out, prev = "", 0
    for i, (s, km) in enumerate(zip(starts, kmers)):
        if i == 0:
            out = km
        else:
            d = s - prev
            out += km[-d:] if d != 0 else km
        prev = s
    return out


def stitch_hits(result: SearchResult, index: ProteomeIndex, query_seqs, target_seqs, query_names, target_names):
    """The stitched region table of `kmerseek search --extract-kmers` (src/python/kmerseek/search.py:64-121), quirks
    included: the reference groups the joined k-mer rows by `match_name` ONLY (search.py:222-240, stitch_kmers_per_gene)
    and labels the group with its first row's query_name after the sort by query start (search.py:70-74), so with several
    queries the rows of different queries that hit one match are stitched together; the query is stitched with the MATCH starts (search.py:78).
    SURVEY section 8f row N4."""
    h = result.hits
    k = index.ksize
    groups = {}
    for i in range(result.n_hits):
        q, t = int(h["hit_qid"][i]), int(h["hit_pid"][i])
        groups.setdefault(target_names[t], []).append((int(h["hit_qpos"][i]), int(h["hit_tpos"][i]), q, t))
    out = []
    for name, lst in groups.items():
        lst.sort(key=lambda r: r[0])  # df.sort("start_query"), search.py:70
        first_q = lst[0][2]           # df["query_name"][0], search.py:74: the row with the smallest query start
        qk = [query_seqs[q][a:a + k] for a, _, q, _ in lst]
        tk = [target_seqs[t][b:b + k] for _, b, _, t in lst]
        ek = [translate(x, index.moltype) for x in qk]
        qs_, ts_ = [r[0] for r in lst], [r[1] for r in lst]
        query = _stitch(qk, ts_)  # the reference stitches the query with the match starts (search.py:78)
        alpha = _stitch(ek, qs_)
        match = _stitch(tk, ts_)
        if not (len(query) == len(alpha) == len(match)):
            raise AssertionError("stitched lengths differ (the reference asserts the same, search.py:87-88)")
        out.append({"match_name": name, "query_name": query_names[first_q], "query_start": min(qs_),
                    "query_end": min(qs_) + len(query), "query": query, "match_start": min(ts_),
                    "match_end": min(ts_) + len(query), "match": match, "encoded": alpha, "length": len(query)})
    out.sort(key=lambda r: (r["query_start"], r["query_end"]))
    return out
