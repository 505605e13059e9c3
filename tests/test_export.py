"""N1: sourmash-compatible .sig.zip and manysearch CSV writers against the reference's golden files (decoded).
CPU-only: the sketches fed to the writers are the golden ones; the GPU test in test_gpu_parity.py checks that the
GPU produces exactly those."""
import csv
import gzip
import io
import json
import zipfile

import numpy as np
import pytest

import kmerseek_b200 as K
from kmerseek_b200 import export


@pytest.mark.parametrize("key,k", [("hp.k16.scaled5", 16), ("hp.k15.scaled5", 15), ("hp.k24.scaled5", 24)])
def test_sig_zip_matches_golden(golden_sigs, key, k, tmp_path):
    g = golden_sigs[key]
    sigs = g["signatures"]
    gold_lines = [l for l in g["manifest"].split("\n") if l]
    filename = list(csv.reader([gold_lines[2]]))[0][-1]
    sketches = [(np.array([int(x) for x in s["mins"]], np.uint64), np.array(s["abundances"], np.uint64)) for s in sigs]
    names = [s["name"] for s in sigs]
    path = tmp_path / "out.sig.zip"
    man = export.write_sig_zip(path, sketches, names, k, 5, "hp", filename, K.max_hash(5))
    assert sorted(l for l in man.split("\n") if l) == sorted(gold_lines)  # same rows (the reference's order is rayon's)
    with zipfile.ZipFile(path) as z:
        assert "SOURMASH-MANIFEST.csv" in z.namelist() and len(z.namelist()) == 26
        for s in sigs:
            doc = json.loads(gzip.decompress(z.read(f"signatures/{s['md5sum']}.sig.gz")))
            assert doc[0]["name"] == s["name"] and doc[0]["hash_function"] == "0.murmur64" and doc[0]["version"] == 0.4
            sk = doc[0]["signatures"][0]
            assert sk["md5sum"] == s["md5sum"] and sk["ksize"] == 3 * k and sk["seed"] == 42 and sk["num"] == 0
            assert sk["max_hash"] == int(s["max_hash"]) and sk["molecule"] == "hp"
            assert [str(x) for x in sk["mins"]] == s["mins"] and sk["abundances"] == s["abundances"]


def test_manysearch_csv_matches_golden(golden_search):
    gold = golden_search["manysearch_csv"]
    rows = []
    for r in csv.DictReader(io.StringIO(gold)):
        row = {}
        for c, v in r.items():
            if c in ("intersect_hashes", "ksize", "scaled", "n_weighted_found", "total_weighted_hashes"):
                row[c] = int(v)
            elif c in ("query_name", "query_md5", "match_name", "moltype", "match_md5"):
                row[c] = v
            else:
                row[c] = float(v)
        rows.append(row)
    assert export.manysearch_csv(rows) == gold  # byte for byte, quoting of the name with a comma included


def test_side_files(tmp_path):
    f = tmp_path / "x.fasta"
    f.write_text(">a\nACD\n")
    p = export.write_manysketch_csv(str(f))
    assert open(p).readlines() == ["name,genome_filename,protein_filename\n", f"x.fasta,,{f}\n"]  # tests/test_index.py:15-19
    sig = export.sig_filename(str(f), "hp", 24, 5)
    assert sig.endswith("x.fasta.hp.k24.scaled5.sig.zip")
    assert open(export.write_siglist(sig)).read() == sig  # tests/test_index.py:27-28
