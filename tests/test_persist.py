"""N2 (partial): the bincode payloads of the reference's save_state round-trip through the decoder and have the
documented structure.  CPU-only: signatures are assembled from the oracle.  UNVERIFIED against the reference's
load_state (no RocksDB, no Rust here)."""
import struct

import numpy as np

from kmerseek_b200 import persist
from kmerseek_b200.index import KmerInfo
from oracle import oracle as O


class _Sig:
    def __init__(self, name, seq, k, moltype, scaled, raw):
        res, offs = O.pack([seq])
        h, pid, pos = O.sketch_tuples(res, offs, k, moltype, scaled)
        (self._mins, self._abunds), = O.protein_sketches(h, pid, 1)
        self.name, self.md5sum, self._raw = name, O.signature_id(self._mins), seq if raw else None
        self._infos = {}
        for hv, (enc, origs) in O.kmer_infos(seq, k, moltype, scaled).items():
            ki = KmerInfo(k, hv, enc)
            ki.original_kmer_to_position = origs
            self._infos[hv] = ki

    def mins(self): return self._mins
    def abunds(self): return self._abunds
    def kmer_infos(self): return self._infos
    def get_raw_sequence(self): return self._raw


def test_signature_chunk_round_trip_and_layout():
    sigs = [_Sig("test_protein1", "PLANTANDANIMALGENQMES", 5, "hp", 1, True), _Sig("test_protein2", "LIVINGALIVE", 5, "hp", 1, True)]
    blob = persist.encode_signature_chunk(sigs, include_raw_sequence=True)
    assert struct.unpack_from("<Q", blob, 0)[0] == 2  # Vec length first
    assert blob[8:16] == struct.pack("<Q", len("test_protein1")) and blob[16:29] == b"test_protein1"
    dec = persist.decode_signature_chunk(blob)
    assert [d["name"] for d in dec] == ["test_protein1", "test_protein2"]
    for d, s in zip(dec, sigs):
        assert d["mins"] == s.mins().tolist() and d["abunds"] == s.abunds().tolist()
        assert d["raw_sequence"] == s.get_raw_sequence()
        assert set(d["kmer_infos"]) == set(s.kmer_infos())
        for h, ki in d["kmer_infos"].items():
            assert ki["ksize"] == 5 and ki["hashval"] == h and ki["encoded_kmer"] == s.kmer_infos()[h].encoded_kmer
            assert ki["original_kmer_to_position"] == s.kmer_infos()[h].original_kmer_to_position
    # hp k5 of PLANT...: 14 entries, three with two originals (src/rust/index.rs:1309-1326)
    assert len(dec[0]["kmer_infos"]) == 14
    assert sum(len(k["original_kmer_to_position"]) == 2 for k in dec[0]["kmer_infos"].values()) == 3
    no_raw = persist.decode_signature_chunk(persist.encode_signature_chunk(sigs, include_raw_sequence=False))
    assert all(d["raw_sequence"] is None for d in no_raw)


def test_metadata_round_trip():
    mins = np.array([1, 5, 2**63 + 7], dtype=np.uint64)
    blob = persist.encode_metadata(250, mins, np.array([1, 2, 3], np.uint64), "dayhoff", 10, 5, False)
    d = persist.decode_metadata(blob)
    assert d == {"total_signatures": 250, "chunk_count": 3, "combined_mins": mins.tolist(), "combined_abunds": [1, 2, 3],
                 "moltype": "dayhoff", "ksize": 10, "scaled": 5, "store_raw_sequences": False}
    assert len(blob) == 8 + 8 + (8 + 24) + (1 + 8 + 24) + (8 + 7) + 4 + 4 + 1
