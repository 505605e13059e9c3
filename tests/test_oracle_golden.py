"""Pin the CPU oracle against every golden vector the reference holds for this path
(extracted by tests/golden/make_golden.py).  CPU only."""
import csv
import io

import numpy as np
import pytest

from conftest import fasta_path
from oracle import oracle as O


def _load(name, protein_index_base=0):
    names, seqs = O.read_fasta(fasta_path(name))
    return names, [O.normalize(s, i) for i, s in enumerate(seqs)]


def _sketch_all(seqs, k, moltype, scaled):
    res, offs = O.pack(seqs)
    h, pid, pos = O.sketch_tuples(res, offs, k, moltype, scaled)
    return h, pid, pos, O.protein_sketches(h, pid, len(seqs))


# G1-G3: src/rust/index.rs:1084-1103,1187-1205,1309-1326 -------------------------------------
@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_known_answer_hashes_and_positions(golden_rust, moltype):
    g = golden_rust["kmer_tables"][moltype]
    infos = O.kmer_infos(g["sequence"], g["ksize"], moltype, g["scaled"])
    assert len(infos) == len(g["rows"])
    for row in g["rows"]:
        enc, originals = infos[int(row["hash"])]
        if row["encoded"] is not None:
            assert enc == row["encoded"]
        assert set(originals) == set(row["originals"])
        assert sorted(p for v in originals.values() for p in v) == sorted(row["positions"])
        for o in row["originals"]:
            assert O.murmur64(O.translate(o, moltype).encode()) == int(row["hash"])
            assert O.murmur64_py(O.translate(o, moltype).encode()) == int(row["hash"])


def test_known_answer_single_hashes():
    # SURVEY App. A.4
    assert O.murmur64(b"PLANT") == 5893010049374798421
    assert O.murmur64(b"bebcb") == 5045972850709227854
    assert O.murmur64(b"hhhpp") == 4230974618842309829


# G7: src/rust/encoding.rs:195,209 ---------------------------------------------------------
def test_translations(golden_rust):
    t = golden_rust["translations"]
    assert O.translate(t["sequence"], "dayhoff") == t["dayhoff"]
    assert O.translate(t["sequence"], "hp") == t["hp"]
    assert O.translate("PLANT", "dayhoff") == "bebcb" and O.translate("PLANT", "hp") == "hhhpp"
    assert O.translate(t["sequence"], "protein") == t["sequence"]


def test_max_hash():
    # SURVEY App. A.1; scaled=5 value is the golden JSON max_hash
    assert O.max_hash(1) == 2**64 - 1
    assert O.max_hash(2) == 9223372036854775808
    assert O.max_hash(5) == 3689348814741910528
    assert O.max_hash(10) == 1844674407370955264
    assert O.max_hash(100) == 184467440737095520
    assert O.max_hash(1000) == 18446744073709552
    assert O.max_hash(0) == 0


# G4-G6: index-level ids, sizes -------------------------------------------------------------
def _index_case(golden_rust, test, fasta=None, content=None):
    g = golden_rust["index_tests"][test]
    if content is not None:
        recs = content.split(">")[1:]
        seqs = [O.normalize("".join(r.split("\n")[1:]), i) for i, r in enumerate(recs)]
    else:
        _, seqs = _load(fasta)
    h, pid, pos, sk = _sketch_all(seqs, g["ksize"], g["moltype"], g["scaled"])
    ids = {O.signature_id(m): len(m) for m, _ in sk}
    for i, n in g["ids"].items():
        assert ids[i] == n, (test, i)
    if g["n_signatures"] is not None:
        assert len(set(ids)) == g["n_signatures"]
    if g["combined_size"] is not None:
        assert len(O.combined_sketch(sk)[0]) == g["combined_size"]


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_two_record_fasta(golden_rust, moltype):
    _index_case(golden_rust, f"test_process_fasta_moltype_{moltype}", content=golden_rust["fixtures"]["TEST_FASTA_CONTENT"])


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_bcl2_first25_sizes(golden_rust, moltype):
    _index_case(golden_rust, f"test_process_fasta_gz_moltype_{moltype}", fasta="bcl2_first25.fasta.gz")


def test_hp_k16_scaled5_combined_1603(golden_rust):
    _index_case(golden_rust, "test_manual_vs_auto_index_equivalence", fasta="bcl2_first25.fasta.gz")


def test_validation_ids(golden_rust):
    g = golden_rust["index_tests"]["test_create_protein_signature_amino_acid_validation_moltype_protein"]
    for seq in ["PLANTANDANIMALGENQMES", "ACDEFGHIKLMNPQRSTVWY"]:
        _, _, _, sk = _sketch_all([O.normalize(seq)], 5, "protein", 1)
        assert g["ids"][O.signature_id(sk[0][0])] == len(sk[0][0])


# G8: errors, stop codon, ambiguity --------------------------------------------------------
def test_invalid_residue_errors(golden_rust):
    assert golden_rust["errors"]
    for e in golden_rust["errors"]:
        with pytest.raises(O.InvalidAminoAcid) as ei:
            O.normalize(e["sequence"])
        assert e["message"] in str(ei.value)
        assert ei.value.pos == 18


def test_stop_codon_and_case_and_ambiguity():
    assert O.normalize("ACDEF*GHIK") == "ACDEF*"  # src/rust/aminoacid.rs:79-83,194-211
    assert O.normalize("mAaGgCcTt") == "MAAGGCCTT"  # src/rust/index.rs:2847-2934
    for i in range(32):
        r = O.normalize("ACBZJXUO", i)
        assert r[2] in "DN" and r[3] in "EQ" and r[4] in "IL" and r[5:] == "XUO"
    # the draw depends on the seed and the position in the sequence, not on the record's index: a copy of a sequence
    # resolves like the original (a query cut from a target matches it), different seeds give different outcomes
    assert len({O.normalize("BBBBBBBBBBBBBBBB", i) for i in range(8)}) == 1
    assert len({O.normalize("BBBBBBBBBBBBBBBB", 0, seed) for seed in range(8)}) > 1
    assert set(O.normalize("BBBBBBBBBBBBBBBB")) == {"D", "N"}
    assert O.normalize("acb*1zJ-", mode="sourmash") == "ACB*1ZJ-"


# G9: full sketches ------------------------------------------------------------------------
@pytest.mark.parametrize("key,k", [("hp.k16.scaled5", 16), ("hp.k15.scaled5", 15), ("hp.k24.scaled5", 24)])
def test_golden_sig_zip(golden_sigs, key, k):
    names, seqs = _load("bcl2_first25.fasta.gz")
    _, _, _, sk = _sketch_all(seqs, k, "hp", 5)
    by_name = {n: s for n, s in zip(names, sk)}
    sigs = golden_sigs[key]["signatures"]
    assert len(sigs) == 25
    for g in sigs:
        mins, abunds = by_name[g["name"]]
        assert g["ksize"] == 3 * k and g["seed"] == 42 and g["molecule"] == "hp" and g["num"] == 0
        assert int(g["max_hash"]) == O.max_hash(5)
        assert [int(x) for x in g["mins"]] == mins.tolist()
        assert g["abundances"] == abunds.tolist()
        assert g["md5sum"] == O.md5sum(mins, k)


# G10: k-mer tables ------------------------------------------------------------------------
@pytest.mark.parametrize("key,k", [("hp.k16.scaled5", 16), ("hp.k15.scaled5", 15), ("hp.k24.scaled5", 24)])
def test_golden_kmer_rows(golden_kmers, key, k):
    names, seqs = _load("bcl2_first25.fasta.gz")
    h, pid, pos, _ = _sketch_all(seqs, k, "hp", 5)
    mine = set()
    for hv, p, s in zip(h.tolist(), pid.tolist(), pos.tolist()):
        km = seqs[p][s:s + k]
        # hashval is stored as int64 in the parquet
        mine.add((names[p], s, hv if hv < 2**63 else hv - 2**64, km, O.translate(km, "hp")))
    gold = [(r["name"], r["start"], int(r["hashval"]), r["kmer"], r["encoded"]) for r in golden_kmers[key]]
    assert set(gold) == mine
    # the golden table repeats a row once per repeated (name, kmer) occurrence pair (sig2kmer.py:143-146 join)
    assert len(gold) >= len(mine)


# G11: manysearch rows ---------------------------------------------------------------------
def test_golden_manysearch(golden_search):
    qn, qs = _load("ced9.fasta")
    tn, ts = _load("bcl2_first25.fasta.gz")
    _, _, _, qsk = _sketch_all(qs, 16, "hp", 5)
    _, _, _, tsk = _sketch_all(ts, 16, "hp", 5)
    rows = O.manysearch(qsk, tsk, 16, 5, "hp", qn, tn)
    gold = list(csv.DictReader(io.StringIO(golden_search["manysearch_csv"])))
    assert len(gold) == 5 and len(rows) == 5
    assert len(qsk[0][0]) == 49
    by = {r["match_name"]: r for r in rows}
    for g in gold:
        r = by[g["match_name"]]
        for col in O.MANYSEARCH_COLUMNS:
            if isinstance(r[col], float):
                assert repr(r[col]) == repr(float(g[col])), col  # to the last printed digit
            else:
                assert str(r[col]) == g[col], col


# G12: stitched regions --------------------------------------------------------------------
def test_golden_stitched(golden_search):
    qn, qs = _load("ced9.fasta")
    tn, ts = _load("bcl2_first25.fasta.gz")
    qh, qid, qpos, _ = _sketch_all(qs, 16, "hp", 5)
    th, tpid, tpos, _ = _sketch_all(ts, 16, "hp", 5)
    hl = O.hits(qh, qid, qpos, th, tpid, tpos)
    pairs = {}
    for q, p, h, a, b in hl:
        qk, tk = qs[q][a:a + 16], ts[p][b:b + 16]
        pairs.setdefault((q, p), []).append((a, b, qk, tk, O.translate(qk, "hp")))
    gold = list(csv.DictReader(io.StringIO(golden_search["stitched_csv"])))
    assert len(gold) == len(pairs) == 5
    by = {tn[p]: O.stitch_pair(v) for (q, p), v in pairs.items()}
    for g in gold:
        r = by[g["match_name"]]
        for col in ["query_start", "query_end", "query", "match_start", "match_end", "match", "encoded", "length"]:
            assert str(r[col]) == g[col], (g["match_name"], col)


def test_hits_are_the_join():
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=60)) for _ in range(6)]
    q = [seqs[2][10:40], seqs[4][5:50]]
    qh, qid, qpos, _ = _sketch_all(q, 7, "dayhoff", 1)
    th, tpid, tpos, _ = _sketch_all(seqs, 7, "dayhoff", 1)
    hl = O.hits(qh, qid, qpos, th, tpid, tpos)
    brute = sorted(((int(a), int(b), int(h), int(c), int(d))
                    for h, a, c in zip(qh, qid, qpos) for h2, b, d in zip(th, tpid, tpos) if h == h2),
                   key=lambda r: (r[0], r[3], r[1], r[4]))
    assert hl == brute and len(hl) >= 24 + 39


def test_cpu_baseline_modes_agree():
    rng = np.random.default_rng(11)
    seqs = ["".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWY"), size=int(n))) for n in rng.integers(3, 300, size=200)]
    res, offs = O.pack(seqs)
    h, pid, pos = O.sketch_tuples(res, offs, 6, "hp", 1)
    n, uniq, kept = O.cpu_baseline(res, offs, 6, "hp", 1, faithful=True, n_threads=2)
    assert n == len(res) and uniq == len(np.unique(h)) and kept == len(h)
    n2, _, kept2 = O.cpu_baseline(res, offs, 6, "hp", 1, faithful=False, n_threads=1)
    assert n2 == n and kept2 == kept


def test_indexed_manysearch_equals_all_pairs():
    """manysearch_indexed (used at full size) against manysearch (the all-pairs restatement pinned by the golden CSV)."""
    from kmerseek_b200 import synth
    res, offs = synth.proteome(60_000, 11)
    qres, qoffs, _ = synth.queries(res, offs, 20, 5, min_len=40, max_len=150)
    for k, moltype, scaled in ((8, "hp", 1), (7, "dayhoff", 1), (5, "protein", 2)):
        th, tpid, _ = O.sketch_tuples(res, offs, k, moltype, scaled)
        qh, qid, _ = O.sketch_tuples(qres, qoffs, k, moltype, scaled)
        qsk = O.protein_sketches(qh, qid, len(qoffs) - 1)
        a = O.manysearch(qsk, O.protein_sketches(th, tpid, len(offs) - 1), k, scaled, moltype)
        b = O.manysearch_indexed(qsk, th, tpid, k, scaled, moltype)
        assert len(a) == len(b) > 0
        for x, y in zip(a, b):
            for c, v in y.items():
                assert x[c] == pytest.approx(v, rel=1e-12, abs=0), c
