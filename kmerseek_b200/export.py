"""sourmash / branchwater compatible files from GPU results (formats: App. C.2-C.3).

  write_sig_zip        per-protein sketches -> `{fasta}.{moltype}.k{k}.scaled{s}.sig.zip`
                       (src/python/kmerseek/sketch.py:23-40: what do_manysketch(singleton=True) produces)
  write_manysearch_csv scored pairs -> the 22-column CSV of do_manysearch (tests/test_search.py:33)
  write_manysketch_csv / write_siglist   the small side files (sketch.py:13-20, index.py:44-48)
Pure host-side formatting: the numbers come from ks_index_export / ks_search_batch.
"""
import csv
import gzip
import io
import json
import os
import zipfile

from .index import md5_of_mins
from .search import MANYSEARCH_COLUMNS

MANIFEST_COLUMNS = ["internal_location", "md5", "md5short", "ksize", "moltype", "num", "scaled", "n_hashes",
                    "with_abundance", "name", "filename"]


def sig_filename(fasta, moltype, ksize, scaled):
    """src/python/kmerseek/sketch.py:23-25"""
    return f"{fasta}.{moltype}.k{ksize}.scaled{scaled}.sig.zip"


def signature_json(name, mins, abunds, ksize, scaled, moltype, filename, max_hash):
    """One sourmash signature document: JSON ksize is 3k, seed 42, num 0."""
    md5 = md5_of_mins(mins, ksize)
    return md5, [{
        "class": "sourmash_signature", "email": "", "hash_function": "0.murmur64", "filename": filename, "name": name,
        "license": "CC0",
        "signatures": [{"num": 0, "ksize": 3 * ksize, "seed": 42, "max_hash": int(max_hash),
                        "mins": [int(x) for x in mins], "md5sum": md5, "abundances": [int(x) for x in abunds],
                        "molecule": moltype}],
        "version": 0.4,
    }]


def manifest_row(md5, ksize, scaled, moltype, n_hashes, name, filename):
    """Manifest ksize is k (the JSON carries 3k)."""
    return [f"signatures/{md5}.sig.gz", md5, md5[:8], ksize, moltype, 0, scaled, n_hashes, 1, name, filename]


def write_sig_zip(path, sketches, names, ksize, scaled, moltype, filename, max_hash):
    """sketches: [(mins, abunds)] in protein order.  Returns the manifest text."""
    man = io.StringIO()
    man.write("# SOURMASH-MANIFEST-VERSION: 1.0\n")
    w = csv.writer(man, lineterminator="\n")
    w.writerow(MANIFEST_COLUMNS)
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_STORED) as z:
        seen = set()
        for (mins, abunds), name in zip(sketches, names):
            md5, doc = signature_json(name, mins, abunds, ksize, scaled, moltype, filename, max_hash)
            w.writerow(manifest_row(md5, ksize, scaled, moltype, len(mins), name, filename))
            if md5 in seen:  # identical sketches share a file, like sourmash's zip storage
                continue
            seen.add(md5)
            z.writestr(f"signatures/{md5}.sig.gz", gzip.compress(json.dumps(doc, separators=(",", ":")).encode(), mtime=0))
        z.writestr("SOURMASH-MANIFEST.csv", man.getvalue())
    return man.getvalue()


def write_manysketch_csv(fasta):
    """src/python/kmerseek/sketch.py:13-20 (asserted at tests/test_index.py:15-19)"""
    path = f"{fasta}.manysketch.csv"
    with open(path, "w") as f:
        f.write("name,genome_filename,protein_filename\n")
        f.write(f"{os.path.basename(fasta)},,{fasta}\n")
    return path


def write_siglist(sig):
    """src/python/kmerseek/index.py:44-48: the path, no trailing newline"""
    path = f"{sig}.siglist"
    with open(path, "w") as f:
        f.write(f"{sig}")
    return path


def _fmt(v):
    return repr(v) if isinstance(v, float) else str(v)


def manysearch_csv(rows):
    """22-column CSV text; floats in shortest round-trip form (what Rust's Display and the golden file show)."""
    out = io.StringIO()
    w = csv.writer(out, lineterminator="\n")
    w.writerow(MANYSEARCH_COLUMNS)
    for r in rows:
        w.writerow([_fmt(r[c]) for c in MANYSEARCH_COLUMNS])
    return out.getvalue()


def write_manysearch_csv(path, rows):
    with open(path, "w") as f:
        f.write(manysearch_csv(rows))
    return path
