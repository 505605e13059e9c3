"""Synthetic proteomes and query sets: numpy PCG64, residues i.i.d. from the
Swiss-Prot background over the 20 standard letters, lengths lognormal(5.6, 0.65) clipped to [30, 35000] and
rescaled to the residue target.  No B/Z/J (the reference resolves them at random, src/rust/aminoacid.rs:45-54)."""
import numpy as np

LETTERS = np.frombuffer(b"ARNDCQEGHILKMFPSTWYV", dtype=np.uint8)
FREQ = np.array([8.25, 5.53, 4.06, 5.46, 1.38, 3.93, 6.72, 7.07, 2.27, 5.91, 9.65, 5.80, 2.41, 3.86, 4.74, 6.65,
                 5.36, 1.10, 2.92, 6.85], dtype=np.float64)
FREQ /= FREQ.sum()


def proteome_lengths(n_residues, seed, mean_len=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    est = max(1, int(n_residues / 350.0))
    lens = np.clip(rng.lognormal(5.6, 0.65, size=est), 30, 35000)
    lens = np.maximum(30, np.floor(lens * (n_residues / lens.sum()))).astype(np.int64)
    diff = int(n_residues - lens.sum())
    if diff > 0:
        lens[-1] += diff
    elif diff < 0:  # trim from the end
        i = len(lens) - 1
        while diff < 0 and i >= 0:
            take = min(int(lens[i]) - 30, -diff)
            lens[i] -= take
            diff += take
            i -= 1
        if diff < 0:
            lens = lens[: max(1, len(lens) + diff // 30)]
    return lens


def residues_iid(n, rng):
    cdf = np.cumsum(FREQ)
    cdf[-1] = 1.0
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for a in range(0, n, step):
        b = min(n, a + step)
        u = rng.random(b - a, dtype=np.float32)
        out[a:b] = LETTERS[np.searchsorted(cdf, u, side="right").clip(0, 19)]
    return out


def proteome(n_residues, seed):
    """-> (residues u8[N], offsets u64[P+1])"""
    lens = proteome_lengths(n_residues, seed)
    offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    rng = np.random.Generator(np.random.PCG64(seed + 1_000_003))
    return residues_iid(int(offsets[-1]), rng), offsets


def queries(residues, offsets, n_queries, seed, min_len=50, max_len=300, sub_rate=0.10):
    """Planted query domains: slices of random proteins with i.i.d. substitutions -> (residues, offsets, source pid)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    P = len(offsets) - 1
    src = rng.integers(0, P, size=n_queries)
    want = rng.integers(min_len, max_len + 1, size=n_queries)
    plen = (offsets[1:] - offsets[:-1]).astype(np.int64)[src]
    qlen = np.minimum(want, plen)
    start = (rng.random(n_queries) * (plen - qlen + 1)).astype(np.int64)
    qoffs = np.zeros(n_queries + 1, dtype=np.uint64)
    np.cumsum(qlen, out=qoffs[1:])
    total = int(qoffs[-1])
    idx = np.repeat(offsets[:-1].astype(np.int64)[src] + start - qoffs[:-1].astype(np.int64), qlen) + np.arange(total)
    qres = residues[idx].copy()
    mut = rng.random(total) < sub_rate
    qres[mut] = residues_iid(int(mut.sum()), rng)
    return qres, qoffs, src


def names(n, prefix="syn"):
    return [f"{prefix}|{i:09d}" for i in range(n)]


def write_fasta(path, residues, offsets, width=60, prefix="syn"):
    """Plain FASTA of a packed proteome: `>syn|000000012` headers, sequences wrapped at `width` columns.  Assembled with
    numpy in slabs of proteins (a 200 M-residue proteome is written in a few seconds)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    P = len(offsets) - 1
    with open(path, "wb") as f:
        step = 50_000
        for a in range(0, P, step):
            b = min(P, a + step)
            lens = offsets[a + 1:b + 1] - offsets[a:b]
            lines = (lens + width - 1) // width  # newlines inside / after a sequence (an empty sequence has none)
            head = len(prefix) + 1 + 9 + 2  # '>' + prefix + '|' + 9 digits + '\n'
            rec = head + lens + lines
            base = np.concatenate([[0], np.cumsum(rec)])
            out = np.full(int(base[-1]), ord("\n"), dtype=np.uint8)
            hdr = np.frombuffer("".join(f">{prefix}|{i:09d}\n" for i in range(a, b)).encode(), dtype=np.uint8).reshape(b - a, head)
            hpos = (base[:-1, None] + np.arange(head)[None, :]).ravel()
            out[hpos] = hdr.ravel()
            n = int(offsets[b] - offsets[a])
            within = np.arange(n) - np.repeat(offsets[a:b] - offsets[a], lens)
            pos = np.repeat(base[:-1] + head, lens) + within + within // width
            out[pos] = residues[int(offsets[a]):int(offsets[b])]
            f.write(out.tobytes())
