"""BASELINE.json configs[4] (C5): k 5..24 x {protein, dayhoff, hp}, scaled 1 -- the reference's own matrix is
benches/benchmark.rs:13-21 (moltype x k in {5, 10, 20}).  For every one of the 60 cells the index of a 10 M-residue
synthetic proteome (keys, row pointers, postings: every retained k-mer with its protein and position) must equal the
oracle's bit for bit, and a planted-query search -- pairs with their scores, hit list, query sketches -- must equal the
oracle's on a sub-proteome sized so that the hit list stays bounded (a two-letter 5-mer has 32 possible hashes: against
10 M residues one query would have tens of millions of hits)."""
import numpy as np
import pytest

from test_gpu_parity import _search_equals_oracle

pytestmark = pytest.mark.gpu

ALPHABET = {"protein": 20.0, "dayhoff": 6.0, "hp": 2.0}
N_RESIDUES = 10_000_000


@pytest.fixture(scope="module")
def K():
    import kmerseek_b200
    return kmerseek_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def proteome():
    from kmerseek_b200 import synth
    return synth.proteome(N_RESIDUES, 20260105)


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_c5_sweep_index_and_hits_equal_oracle(K, O, proteome, moltype):
    from kmerseek_b200 import synth
    res, offs = proteome
    prot = K.Proteome.from_packed(res, offs)
    for k in range(5, 25):
        # 1. the whole index at 10 M residues
        with K.ProteomeIndex("c5", k, 1, moltype) as idx:
            idx.add_proteome(prot)
            idx.finalize()
            keys, row_ptr, pid, pos = idx.csr()
            st = idx.stats()
        oh, opid, opos = O.sketch_tuples(res, offs, k, moltype, 1)
        okeys, orow, ops, oqs = O.build_index(oh, opid, opos)
        assert st["n_tuples"] == len(oh) and st["n_unique_hashes"] == len(okeys), (moltype, k)
        assert np.array_equal(keys, okeys) and np.array_equal(row_ptr, orow), (moltype, k)
        assert np.array_equal(pid, ops) and np.array_equal(pos, oqs), (moltype, k)
        del keys, row_ptr, pid, pos, oh, opid, opos, okeys, orow, ops, oqs
        # 2. planted queries against a sub-proteome: expected hits = query windows x sub-proteome windows / k-mer space
        space = ALPHABET[moltype] ** k
        n_q = 12
        sub_res = int(min(150_000, max(20_000, 150_000 * space / (n_q * 120.0))))
        p = int(np.searchsorted(offs, sub_res))
        sres, soffs = res[:int(offs[p])], offs[:p + 1]
        qres, qoffs, _ = synth.queries(sres, soffs, n_q, 77 + k, min_len=max(40, k + 8), max_len=160, sub_rate=0.03)
        with K.ProteomeIndex("c5s", k, 1, moltype) as idx:
            idx.add_proteome(K.Proteome.from_packed(sres, soffs))
            idx.finalize()
            r = _search_equals_oracle(K, O, idx, sres, soffs, qres, qoffs, k, moltype, 1, hits=True)
            assert r.n_pairs >= n_q // 2 and r.n_hits > 0, (moltype, k)  # (3 % substitutions: most queries keep a k-mer)
    prot.close()
