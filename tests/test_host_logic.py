"""CPU-only tests: the C-ABI library loads and exports every declared symbol, and the host-side logic
(normalisation, FASTA ingest, md5/id strings, builder and error conventions) matches the oracle and the
reference's conventions.  No compute entry point is called here (there is no GPU in this container)."""
import gzip
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import ROOT, fasta_path, has_gpu
import kmerseek_b200 as K
from kmerseek_b200 import _ffi, synth
from oracle import oracle as O


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kmerseek_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ks_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    L = _ffi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in the header but not exported"
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    assert L.ks_abi_version() == 1


def test_scalars_match_oracle():
    for s in (0, 1, 2, 5, 10, 100, 1000, 7, 3, 4096):
        assert K.max_hash(s) == O.max_hash(s)
    L = _ffi.lib()
    for m, name in ((0, "protein"), (1, "dayhoff"), (2, "hp")):
        for c in range(256):
            assert L.ks_translate_residue(c, m) == O.lib().kso_translate(c, O.MOLTYPES[name])
    assert K.translate("LIVINGALIVE", "dayhoff") == "eeeecbbeeec"  # src/rust/encoding.rs:195
    assert K.translate("LIVINGALIVE", "hp") == "hhhhphhhhhp"  # src/rust/encoding.rs:209
    assert K.translate("AX*UOB", "hp") == "hX*XXX" and K.translate("AX*", "protein") == "AX*"


def test_md5_and_id_strings(golden_sigs):
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 100, 1000):
        mins = np.sort(rng.integers(0, 2**63, size=n, dtype=np.uint64))
        for k in (5, 16, 24):
            m = hashlib.md5(str(3 * k).encode() + b"".join(str(int(x)).encode() for x in mins)).hexdigest()
            assert K.md5_of_mins(mins, k) == m == O.md5sum(mins, k)
        assert K.id_of_mins(mins) == O.signature_id(mins)
    g = golden_sigs["hp.k16.scaled5"]["signatures"][0]
    assert K.md5_of_mins(np.array([int(x) for x in g["mins"]], dtype=np.uint64), 16) == g["md5sum"]
    assert K.id_of_mins(np.array([2**63, 2**63, 5], dtype=np.uint64)) == "5"
    assert K.id_of_mins(np.zeros(0, dtype=np.uint64)) == "0"


def test_normalisation_matches_oracle():
    rng = np.random.default_rng(7)
    alphabet = list("ACDEFGHIKLMNPQRSTVWYXUOBZJacdefghiklmnpqrstvwyxuobzj*")
    seqs = ["".join(rng.choice(alphabet, size=int(n))) for n in rng.integers(0, 120, size=300)]
    for seed in (0, 12345):
        p = K.Proteome.from_sequences(seqs, None, ambig_seed=seed)
        for i, s in enumerate(seqs):
            assert p.sequence(i) == O.normalize(s, i, seed)
        assert p.n_residues == sum(len(O.normalize(s, i, seed)) for i, s in enumerate(seqs))
        assert p.offsets[0] == 0 and p.offsets[-1] == p.n_residues


def test_invalid_residue_error_convention(golden_rust):
    for e in golden_rust["errors"]:  # src/rust/index.rs:2031-2046
        with pytest.raises(K.InvalidAminoAcid) as ei:
            K.Proteome.from_sequences(["ACDEF", e["sequence"]])
        assert e["message"] in str(ei.value)
        assert str(ei.value) == f"Invalid amino acid '{e['sequence'][17]}' found at position 18"
        assert (ei.value.pos, ei.value.protein_index) == (18, 1)
    with pytest.raises(K.InvalidAminoAcid) as ei:
        K.Proteome.from_sequences(["acd1"])
    assert ei.value.char == "1" and ei.value.pos == 4
    # '*' ends the sequence before the bad character is ever looked at (src/rust/aminoacid.rs:79-83)
    assert K.Proteome.from_sequences(["ACD*1"]).sequence(0) == "ACD*"


def test_fasta_ingest_matches_oracle(tmp_path):
    for name in ("bcl2_first25.fasta.gz", "ced9.fasta", "test_compression.fasta", "test_mixed_case.fasta",
                 "bcl2_all300.fasta.gz"):
        p = K.Proteome.from_fasta(fasta_path(name))
        names, seqs = O.read_fasta(fasta_path(name))
        assert p.names == names
        for i, s in enumerate(seqs):
            assert p.sequence(i) == O.normalize(s, i)
    p = K.Proteome.from_fasta(fasta_path("bcl2_first25.fasta.gz"))
    assert (p.n_proteins, p.n_residues) == (25, 9288)  # SURVEY section 4 fixtures
    f = tmp_path / "crlf.fasta"
    f.write_bytes(b">a b c\r\nACDE\r\nFGH\r\n\r\n>second\r\n>third\r\nKLMN")
    p = K.Proteome.from_fasta(f)
    assert p.names == ["a b c", "second", "third"]
    assert [p.sequence(i) for i in range(3)] == ["ACDEFGH", "", "KLMN"]
    g = tmp_path / "x.fasta.gz"
    with gzip.open(g, "wb") as fh:
        fh.write(b">p1\nPLANTANDANIMALGENQMES\n>p2\nlivingalive\n")
    p = K.Proteome.from_fasta(g)
    assert [p.sequence(i) for i in range(2)] == ["PLANTANDANIMALGENQMES", "LIVINGALIVE"]


def test_compressed_fasta_like_niffler(tmp_path):
    """niffler sniffs gzip / zstd / bzip2 / xz (src/rust/index.rs:920; zstd fixture: src/rust/index.rs:1734-1789)."""
    import bz2
    import lzma
    p = K.Proteome.from_fasta(fasta_path("test_compression.fasta.zst"))
    assert p.names == ["test_protein1", "test_protein2"]
    assert [p.sequence(i) for i in range(2)] == ["PLANTANDANIMALGENQMES", "LIVINGALIVE"]
    raw = gzip.open(fasta_path("bcl2_first25.fasta.gz"), "rb").read()
    ref = K.Proteome.from_fasta(fasta_path("bcl2_first25.fasta.gz"))
    for name, data in (("a.fasta.bz2", bz2.compress(raw)), ("a.fasta.xz", lzma.compress(raw)), ("a.fasta", raw)):
        f = tmp_path / name
        f.write_bytes(data)
        p = K.Proteome.from_fasta(f)
        assert p.names == ref.names and np.array_equal(p.residues, ref.residues) and np.array_equal(p.offsets, ref.offsets)


def test_fasta_errors(tmp_path):
    with pytest.raises(K.ParseError):
        K.Proteome.from_fasta(tmp_path / "missing.fasta")
    e = tmp_path / "empty.fasta"
    e.write_bytes(b"")
    with pytest.raises(K.ParseError):
        K.Proteome.from_fasta(e)
    b = tmp_path / "bad.fasta"
    b.write_bytes(b"ACDEFG\n")
    with pytest.raises(K.ParseError):
        K.Proteome.from_fasta(b)
    z = tmp_path / "z.fasta.zst"
    z.write_bytes(b"\x28\xb5\x2f\xfd" + b"\xff" * 32)
    with pytest.raises(K.ParseError):  # corrupt stream
        K.Proteome.from_fasta(z)
    bad = tmp_path / "bad_res.fasta"
    bad.write_bytes(b">ok\nACDEF\n>bad\nPLANTANDANIMALGEN1MES\n")
    with pytest.raises(K.InvalidAminoAcid, match="Invalid amino acid '1' found at position 18"):
        K.Proteome.from_fasta(bad)  # src/rust/index.rs:2251-2282


def test_packed_validation():
    with pytest.raises(K.ValidationError):
        K.Proteome.from_packed(np.zeros(4, np.uint8), np.array([1, 4], np.uint64))
    with pytest.raises(K.ValidationError):
        K.Proteome.from_packed(np.zeros(4, np.uint8), np.array([0, 3, 2], np.uint64))
    p = K.Proteome.from_packed(np.frombuffer(b"ACDEFG", np.uint8), np.array([0, 2, 2, 6], np.uint64))
    assert [p.sequence(i) for i in range(3)] == ["AC", "", "DEFG"] and p.names == ["", "", ""]


def test_builder_and_moltype_conventions():
    B = K.ProteomeIndex.builder
    for fn, msg in [(lambda: B().ksize(5).scaled(1).moltype("hp").build(), "Database path is required"),
                    (lambda: B().ksize(5).scaled(1).moltype("hp").build_with_auto_filename(), "Base path is required"),
                    (lambda: B().path("x").scaled(1).moltype("hp").build(), "K-mer size is required"),
                    (lambda: B().path("x").ksize(5).moltype("hp").build(), "Scaled value is required"),
                    (lambda: B().path("x").ksize(5).scaled(1).build(), "Molecular type is required")]:
        with pytest.raises(K.BuilderError) as ei:  # src/rust/index.rs:3021-3036
            fn()
        assert str(ei.value) == f"Builder error: {msg}"
    with pytest.raises(K.InvalidMoltype) as ei:  # src/rust/encoding.rs:22-25
        K.ProteomeIndex("x", 5, 1, "dna")
    assert str(ei.value) == "Invalid moltype: dna, only 'protein', 'hp', or 'dayhoff' are supported"


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a device")
def test_no_device_fails_loudly():
    with pytest.raises(K.NoDevice, match="no CPU fallback"):
        K.ProteomeIndex("x", 5, 1, "hp")
    with pytest.raises(K.NoDevice):
        K.ProteomeIndex.builder().path("x").ksize(5).scaled(1).moltype("raw").build()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kmerseek_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def test_synth_generator_is_deterministic_and_shaped():
    r1, o1 = synth.proteome(200_000, 20260102)
    r2, o2 = synth.proteome(200_000, 20260102)
    assert np.array_equal(r1, r2) and np.array_equal(o1, o2)
    assert o1[-1] == len(r1) == 200_000
    lens = np.diff(o1.astype(np.int64))
    assert lens.min() >= 30 and lens.max() <= 35000
    assert set(np.unique(r1)) <= set(b"ACDEFGHIKLMNPQRSTVWY")
    q, qo, src = synth.queries(r1, o1, 50, 78)
    assert len(qo) == 51 and qo[-1] == len(q) and (np.diff(qo.astype(np.int64)) >= 30).all()


def test_ctypes_mirrors_match_the_header(tmp_path):
    """The ctypes structures of kmerseek_b200/_ffi.py must have the size and field offsets of the C structs in
    include/kmerseek_b200.h (compiled here with gcc: the header is plain C)."""
    import ctypes as C
    import shutil
    import subprocess
    from kmerseek_b200 import _ffi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"ks_stats": _ffi.ks_stats, "ks_sketch": _ffi.ks_sketch, "ks_csr": _ffi.ks_csr, "ks_params": _ffi.ks_params}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "kmerseek_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(cls, field).offset, f"{name}.{field}"


def test_search_result_columns_are_built_on_first_use():
    """kmerseek_b200.search._LazyColumns: the dict of result columns makes its numpy arrays on first access and otherwise
    behaves like the dict it replaced (items, iteration, `in`, assignment)."""
    import ctypes as C
    from kmerseek_b200.search import _LazyColumns
    a = (C.c_uint32 * 4)(1, 2, 3, 4)
    b = (C.c_double * 4)(0.5, 1.5, 2.5, 3.5)
    cols = _LazyColumns({"pair_qid": (C.cast(a, C.POINTER(C.c_uint32)), 4, np.uint32, None),
                         "jaccard": (C.cast(b, C.POINTER(C.c_double)), 4, np.float64, None)})
    assert len(cols) == 2 and "jaccard" in cols and "nope" not in cols and bool(cols)
    assert dict.__len__(cols) == 0  # nothing built yet
    assert cols["pair_qid"].tolist() == [1, 2, 3, 4] and dict.__len__(cols) == 1
    cols["pair_qid"] = cols["pair_qid"] + np.uint32(10)
    assert cols["pair_qid"].tolist() == [11, 12, 13, 14]
    assert sorted(cols) == ["jaccard", "pair_qid"]
    assert {k: v.tolist() for k, v in cols.items()} == {"pair_qid": [11, 12, 13, 14], "jaccard": [0.5, 1.5, 2.5, 3.5]}
    assert cols.get("nope") is None and cols.get("jaccard")[0] == 0.5
    with pytest.raises(KeyError):
        cols["nope"]


def _np_pack5(res):
    """5-bit upload format restated in numpy: A-Z = 1..26, '*' = 27, 8 residues per 5 bytes little-endian, + 72 zero bytes."""
    n = len(res)
    code = np.where((res >= 65) & (res <= 90), res - 64, np.where(res == 42, 27, 255)).astype(np.uint64)
    assert int(code.max(initial=0)) < 32
    groups = (n + 7) // 8
    c = np.zeros(groups * 8, dtype=np.uint64)
    c[:n] = code
    c = c.reshape(groups, 8)
    v = np.zeros(groups, dtype=np.uint64)
    for i in range(8):
        v |= c[:, i] << np.uint64(5 * i)
    out = np.zeros(groups * 5 + 72, dtype=np.uint8)
    for b in range(5):
        out[b:groups * 5:5] = ((v >> np.uint64(8 * b)) & np.uint64(0xff)).astype(np.uint8)
    return out


def _packed_of(prot):
    import ctypes as C
    n = C.c_uint64(0)
    ptr = _ffi.lib().ks_proteome_packed(prot._h, C.byref(n))
    if not ptr or n.value == 0:
        return None
    return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_fasta_ingest_packs_the_upload_copy_in_the_same_pass(tmp_path, monkeypatch, threads):
    """The FASTA parser writes the 5-bit upload copy while it fills the residues (every thread packs the groups inside its
    own residue range, the straddling groups afterwards): byte-identical to the separate-pass packer behind
    ks_proteome_from_packed and to a numpy restatement, for thread counts that put chunk boundaries at every offset mod 8."""
    monkeypatch.setenv("KS_INGEST_THREADS", threads)
    rng = np.random.default_rng(7 + int(threads))
    aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXUO", dtype=np.uint8)
    path = str(tmp_path / "p.fasta")
    want = []
    N_REC = 48_000  # ~17 MB: the parser takes one thread per MB of file, so 16 threads means 16 chunks here
    with open(path, "w") as f:
        for i in range(N_REC):
            n = int(rng.integers(0, 700)) if i % 97 else 0  # a few empty records
            seq = aa[rng.integers(0, len(aa), n)].tobytes().decode()
            if i % 211 == 5 and n > 10:
                seq = seq[:n // 2] + "*" + seq[n // 2:]  # truncated after the '*' (kept)
            if i % 5 == 0:
                seq = seq.lower()
            f.write(f">r{i}\n")
            for j in range(0, len(seq), 61):
                f.write(seq[j:j + 61] + "\n")
            seq = seq.upper()
            want.append(seq[:seq.index("*") + 1] if "*" in seq else seq)
    # the parser reads its thread count once per process: run it in a child so that every parameter gets its own
    import subprocess, sys, textwrap
    code = textwrap.dedent(f"""
        import sys, ctypes as C, numpy as np
        sys.path.insert(0, {ROOT!r})
        import kmerseek_b200 as K
        from kmerseek_b200 import _ffi
        p = K.Proteome.from_fasta({path!r})
        n = C.c_uint64(0)
        ptr = _ffi.lib().ks_proteome_packed(p._h, C.byref(n))
        np.save({str(tmp_path / 'packed.npy')!r}, np.ctypeslib.as_array(ptr, shape=(n.value,)).copy())
        np.save({str(tmp_path / 'res.npy')!r}, np.array(p.residues))
        np.save({str(tmp_path / 'offs.npy')!r}, np.array(p.offsets))
    """)
    env = dict(os.environ, KS_INGEST_THREADS=threads)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    packed, res, offs = (np.load(str(tmp_path / f)) for f in ("packed.npy", "res.npy", "offs.npy"))
    assert res.tobytes().decode() == "".join(want) and len(offs) == N_REC + 1
    assert np.array_equal(packed, _np_pack5(res))
    ref = K.Proteome.from_packed(res, offs)  # the separate-pass packer
    assert np.array_equal(_packed_of(ref), packed)
    ref.close()


def test_unpackable_bytes_leave_no_packed_copy(tmp_path):
    """sourmash-mode normalisation keeps any byte: one outside A-Z and '*' means the residues themselves are uploaded."""
    path = str(tmp_path / "q.fasta")
    open(path, "w").write(">a\nACDE-FGH\n>b\nKLMN\n")
    p = K.Proteome.from_fasta(path, mode="sourmash")
    assert _packed_of(p) is None and bytes(p.residues) == b"ACDE-FGHKLMN"
    p.close()
    open(path, "w").write(">a\nACDEFGH\n>b\nKLMN\n")
    p = K.Proteome.from_fasta(path, mode="sourmash")
    assert np.array_equal(_packed_of(p), _np_pack5(np.frombuffer(b"ACDEFGHKLMN", dtype=np.uint8)))
    p.close()
