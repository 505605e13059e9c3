#!/usr/bin/env python3
"""Extract the reference's golden vectors for the sketch-and-search path into small fixtures.

Runs only in the build container (reads /root/reference, which does not exist on the GPU
box).  Output lives beside this script and is committed; tests read the output only.

What is extracted:
  G1-G3  src/rust/index.rs:1084-1103,1187-1205,1309-1326  (hash, kmer[, encoded], pos) tables
  G4-G6  src/rust/index.rs test fns: per-test {ksize, moltype, scaled, id -> n_kmers, combined size}
  G7     src/rust/encoding.rs:195,209   translations of LIVINGALIVE
  G8     src/rust/index.rs:2031-2046    invalid-residue messages
  G9     tests/testdata/**/*.sig.zip    75 full sourmash sketches
  G10    tests/testdata/**/*.kmers.pq   (name, start, hashval, kmer, encoded) rows
  G11    tests/test_search.py:33-39     manysearch rows (22 columns)
  G12    tests/test_search.py:88-94     stitched regions
  inputs the FASTA files those tests run on (data, not source)
"""
import gzip
import json
import os
import re
import shutil
import sys
import zipfile

REF = os.environ.get("KMERSEEK_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def fn_bodies(src):
    """Yield (name, first_line_no, body_text) for each `fn test_*` in a Rust file."""
    lines = src.split("\n")
    starts = [i for i, l in enumerate(lines) if re.search(r"\bfn test_\w+\(", l)]
    for n, s in enumerate(starts):
        e = starts[n + 1] if n + 1 < len(starts) else len(lines)
        name = re.search(r"fn (test_\w+)\(", lines[s]).group(1)
        yield name, s + 1, "\n".join(lines[s:e])


def rust_goldens():
    src = open(os.path.join(REF, "src/rust/index.rs")).read()
    out = {"source": "src/rust/index.rs", "kmer_tables": {}, "index_tests": {}, "errors": []}
    t1 = re.compile(r'\((\d+), \("(\w+)", \[([\d, ]+)\]\)\)')
    t2 = re.compile(r'\((\d+), \("(\w+)", "(\w+)", \[([\d, ]+)\]\)\)')
    t3 = re.compile(r'\((\d+), \("(\w+)", vec!\[([^\]]*)\], vec!\[([\d, ]+)\]\)\)')
    for name, line, body in fn_bodies(src):
        ks = re.search(r"let protein_ksize = (\d+);", body)
        mt = re.search(r'let moltype = "(\w+)";', body)
        if name.startswith("test_process_kmers_moltype_"):
            moltype = name.rsplit("_", 1)[1]
            rows = []
            for m in t1.finditer(body):
                rows.append({"hash": m.group(1), "encoded": m.group(2) if moltype == "protein" else None,
                             "originals": [m.group(2)], "positions": [int(x) for x in m.group(3).split(",")]})
            for m in t2.finditer(body):
                rows.append({"hash": m.group(1), "encoded": m.group(2), "originals": [m.group(3)],
                             "positions": [int(x) for x in m.group(4).split(",")]})
            for m in t3.finditer(body):
                rows.append({"hash": m.group(1), "encoded": m.group(2),
                             "originals": re.findall(r'"(\w+)"', m.group(3)),
                             "positions": [int(x) for x in m.group(4).split(",")]})
            out["kmer_tables"][moltype] = {"line": line, "ksize": int(ks.group(1)), "scaled": 1,
                                           "sequence": "PLANTANDANIMALGENQMES", "rows": rows}
        ids = re.findall(r'md5sum == "([0-9a-f]+)"[^;]*?len\(\) == (\d+)', body, flags=re.S)
        comb = re.search(r"combined_minhash\.size\(\) == (\d+)", body)
        nsig = re.search(r"signatures\.len\(\), (\d+)", body)
        if (ids or comb) and ks and mt:
            out["index_tests"][name] = {
                "line": line, "ksize": int(ks.group(1)), "moltype": mt.group(1), "scaled": 1,
                "ids": {i: int(n) for i, n in ids},
                "combined_size": int(comb.group(1)) if comb else None,
                "n_signatures": int(nsig.group(1)) if nsig else None,
            }
        if name == "test_manual_vs_auto_index_equivalence":
            out["index_tests"][name] = {"line": line, "ksize": 16, "moltype": "hp", "scaled": 5,
                                        "ids": {}, "n_signatures": 25,
                                        "combined_size": int(re.search(r"== (\d+),\s*\n\s*\"Manual index should have \d+ hashes", body).group(1))
                                        if re.search(r"== (\d+),\s*\n\s*\"Manual index should have \d+ hashes", body) else None}
        for m in re.finditer(r'\("(PLANT[^"]+)", "(Invalid amino acid \'.\')"\)', body):
            out["errors"].append({"test": name, "sequence": m.group(1), "message": m.group(2)})
    # combined size 1603 (index.rs:2414-2418) is asserted through combined_minhash_size()
    m = re.search(r"combined_minhash_size\(\) == (\d+)", src)
    if m:
        out["index_tests"]["test_manual_vs_auto_index_equivalence"]["combined_size"] = int(m.group(1))
    enc = open(os.path.join(REF, "src/rust/encoding.rs")).read()
    out["translations"] = {
        "sequence": "LIVINGALIVE",
        "dayhoff": re.search(r'"(e[a-f]+)"', enc).group(1),
        "hp": re.search(r'"([hp]{11})"', enc).group(1),
    }
    fx = open(os.path.join(REF, "src/rust/tests/test_fixtures.rs")).read()
    out["fixtures"] = {k: json.loads('"' + v + '"') for k, v in re.findall(r'pub const (\w+): &str =\s*"([^"]*)";', fx)}
    return out


def sig_zip(path):
    sigs = []
    with zipfile.ZipFile(path) as z:
        manifest = z.read("SOURMASH-MANIFEST.csv").decode()
        order = [l.split(",")[0] for l in manifest.split("\n")[2:] if l]
        for n in order:
            d = json.loads(gzip.decompress(z.read(n)))
            assert len(d) == 1 and len(d[0]["signatures"]) == 1
            s = d[0]["signatures"][0]
            sigs.append({"name": d[0]["name"], "md5sum": s["md5sum"], "ksize": s["ksize"], "seed": s["seed"],
                         "max_hash": str(s["max_hash"]), "molecule": s["molecule"], "num": s["num"],
                         "mins": [str(x) for x in s["mins"]], "abundances": s["abundances"]})
    return {"manifest": manifest, "signatures": sigs}


def kmers_pq(path):
    import pyarrow.parquet as pq
    t = pq.read_table(path).to_pydict()
    rows = sorted(zip(t["sequence_name"], t["start"], t["hashval"], t["kmer"], t["encoded"]))
    return [{"name": a, "start": int(b), "hashval": str(c), "kmer": d, "encoded": e} for a, b, c, d, e in rows]


def search_goldens():
    src = open(os.path.join(REF, "tests/test_search.py")).read()
    blocks = re.findall(r'StringIO\(\s*"""(.*?)"""', src, flags=re.S)
    return {"manysearch_csv": blocks[0], "stitched_csv": blocks[1],
            "source": "tests/test_search.py:33-39,88-94"}


def main():
    td = os.path.join(REF, "tests/testdata")
    b25 = "bcl2_first25_uniprotkb_accession_O43236_OR_accession_2025_02_06.fasta.gz"
    os.makedirs(os.path.join(OUT, "fasta"), exist_ok=True)
    for src, dst in [
        (f"fasta/{b25}", "bcl2_first25.fasta.gz"),
        ("fasta/ced9.fasta", "ced9.fasta"),
        ("fasta/test_compression.fasta", "test_compression.fasta"),
        ("fasta/test_compression.fasta.zst", "test_compression.fasta.zst"),
        ("fasta/uniprotkb_BCL2_AND_model_organism_9606_2025_02_06.fasta.gz", "bcl2_all300.fasta.gz"),
    ]:
        shutil.copyfile(os.path.join(td, src), os.path.join(OUT, "fasta", dst))
        os.chmod(os.path.join(OUT, "fasta", dst), 0o644)
    shutil.copyfile(os.path.join(REF, "test_mixed_case.fasta"), os.path.join(OUT, "fasta", "test_mixed_case.fasta"))
    os.chmod(os.path.join(OUT, "fasta", "test_mixed_case.fasta"), 0o644)

    json.dump(rust_goldens(), open(os.path.join(OUT, "rust_tests.json"), "w"), indent=1, sort_keys=True)
    json.dump(search_goldens(), open(os.path.join(OUT, "search.json"), "w"), indent=1, sort_keys=True)

    sigs = {
        "hp.k16.scaled5": sig_zip(os.path.join(td, f"index/{b25}.hp.k16.scaled5.sig.zip")),
        "hp.k15.scaled5": sig_zip(os.path.join(td, f"index/{b25}.hp.k15.scaled5.sig.zip")),
        "hp.k24.scaled5": sig_zip(os.path.join(td, f"fasta/{b25}.hp.k24.scaled5.sig.TRUE.zip")),
    }
    with gzip.open(os.path.join(OUT, "sigs.json.gz"), "wt", compresslevel=9) as f:
        json.dump(sigs, f, sort_keys=True)
    kmers = {
        "hp.k16.scaled5": kmers_pq(os.path.join(td, f"index/{b25}.hp.k16.scaled5.sig.zip.kmers.pq")),
        "hp.k15.scaled5": kmers_pq(os.path.join(td, f"index/{b25}.hp.k15.scaled5.sig.zip.kmers.pq")),
        "hp.k24.scaled5": kmers_pq(os.path.join(td, f"fasta/{b25}.hp.k24.scaled5.sig.TRUE.zip.kmers.pq")),
    }
    with gzip.open(os.path.join(OUT, "kmers.json.gz"), "wt", compresslevel=9) as f:
        json.dump(kmers, f, sort_keys=True)
    for k in sigs:
        print(k, len(sigs[k]["signatures"]), "sigs", len(kmers[k]), "kmer rows")
    r = json.load(open(os.path.join(OUT, "rust_tests.json")))
    print({k: len(v["rows"]) for k, v in r["kmer_tables"].items()})
    print({k: (v["ids"], v["combined_size"]) for k, v in r["index_tests"].items()})
    print(r["errors"][:3], r["translations"])


if __name__ == "__main__":
    sys.exit(main())
