"""bincode 1.3 payloads of the reference's `save_state` (layout: App. C.1).

`save_state` (src/rust/index.rs:227-265) puts two kinds of values into RocksDB:
    "index_metadata"        -> bincode(ProteomeIndexMetadata)           (src/rust/index.rs:46-56)
    "signatures_chunk_{i}"  -> bincode(Vec<ProteinSignatureData>), 100 per chunk (src/rust/signature.rs:324-335,
                               src/rust/kmer.rs:7-12)
This module produces exactly those byte strings from GPU results and writes them as one file per key
(`{dir}/index_metadata.bin`, `{dir}/signatures_chunk_{i}.bin`) for a `db.put(key, value)` loader.

NOT built: the RocksDB container itself (SST / MANIFEST / WAL) -- there is no RocksDB library in this image and
no way to run the reference to verify a hand-written table file.  The payloads are checked by a round-trip decoder
and by their structure only (tests/test_persist.py): UNVERIFIED against the reference's `load_state`.

bincode 1.3 defaults: little endian, fixed-width integers, usize -> u64, String / Vec / HashMap = u64 length + items,
Option = u8 tag, bool = u8, struct fields in declaration order.
"""
import os
import struct

CHUNK_SIZE = 100  # src/rust/index.rs:240


def _u64(x):
    return struct.pack("<Q", int(x))


def _u32(x):
    return struct.pack("<I", int(x))


def _str(s):
    b = s.encode()
    return _u64(len(b)) + b


def _vec_u64(v):
    return _u64(len(v)) + b"".join(struct.pack("<Q", int(x)) for x in v)


def _opt(payload):
    return b"\x00" if payload is None else b"\x01" + payload


def encode_kmer_info(ki):
    """KmerInfo { ksize: usize, hashval: u64, encoded_kmer: String, original_kmer_to_position: HashMap<String, Vec<usize>> }"""
    out = [_u64(ki.ksize), _u64(ki.hashval), _str(ki.encoded_kmer), _u64(len(ki.original_kmer_to_position))]
    for orig, pos in ki.original_kmer_to_position.items():
        out.append(_str(orig))
        out.append(_vec_u64(pos))
    return b"".join(out)


def encode_signature_data(sig, include_raw_sequence):
    """ProteinSignatureData { name, mins, abunds: Option<Vec<u64>>, kmer_infos: HashMap<u64, KmerInfo>, raw_sequence: Option<String> }"""
    infos = sig.kmer_infos()
    out = [_str(sig.name), _vec_u64(sig.mins()), _opt(_vec_u64(sig.abunds())), _u64(len(infos))]
    for h, ki in infos.items():
        out.append(_u64(h))
        out.append(encode_kmer_info(ki))
    raw = sig.get_raw_sequence() if include_raw_sequence else None
    out.append(_opt(_str(raw) if raw is not None else None))
    return b"".join(out)


def encode_signature_chunk(sigs, include_raw_sequence=False):
    """Vec<ProteinSignatureData> (one `signatures_chunk_{i}` value)."""
    return _u64(len(sigs)) + b"".join(encode_signature_data(s, include_raw_sequence) for s in sigs)


def encode_metadata(total_signatures, combined_mins, combined_abunds, moltype, ksize, scaled, store_raw_sequences):
    """ProteomeIndexMetadata (the `index_metadata` value)."""
    chunk_count = (total_signatures + CHUNK_SIZE - 1) // CHUNK_SIZE
    return b"".join([_u64(total_signatures), _u64(chunk_count), _vec_u64(combined_mins), _opt(_vec_u64(combined_abunds)),
                     _str(moltype), _u32(ksize), _u32(scaled), b"\x01" if store_raw_sequences else b"\x00"])


def save_state_blobs(directory, index, signatures):
    """What save_state would `put`, one file per key.  `signatures` are the ProteinSignature objects to store (the
    reference stores one per distinct id: equal ids overwrite, src/rust/index.rs:817-820)."""
    os.makedirs(directory, exist_ok=True)
    by_id = {}
    for s in signatures:
        by_id[s.md5sum] = s
    sigs = list(by_id.values())
    keys = []
    for i in range(0, len(sigs), CHUNK_SIZE):
        key = f"signatures_chunk_{i // CHUNK_SIZE}"
        with open(os.path.join(directory, key + ".bin"), "wb") as f:
            f.write(encode_signature_chunk(sigs[i:i + CHUNK_SIZE], index.store_raw_sequences()))
        keys.append(key)
    mins, abunds = index.get_combined_minhash()
    with open(os.path.join(directory, "index_metadata.bin"), "wb") as f:
        f.write(encode_metadata(len(sigs), mins, abunds, index.moltype, index.ksize, index.scaled, index.store_raw_sequences()))
    keys.append("index_metadata")
    return keys


# --- decoder (tests; also documents the layout) -------------------------------------------------------------
class _R:
    def __init__(self, b):
        self.b, self.i = b, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.i)[0]
        self.i += 8
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.i)[0]
        self.i += 4
        return v

    def u8(self):
        v = self.b[self.i]
        self.i += 1
        return v

    def string(self):
        n = self.u64()
        s = self.b[self.i:self.i + n].decode()
        self.i += n
        return s

    def vec_u64(self):
        return [self.u64() for _ in range(self.u64())]

    def opt(self, fn):
        return fn() if self.u8() else None


def decode_metadata(b):
    r = _R(b)
    d = {"total_signatures": r.u64(), "chunk_count": r.u64(), "combined_mins": r.vec_u64(),
         "combined_abunds": r.opt(r.vec_u64), "moltype": r.string(), "ksize": r.u32(), "scaled": r.u32(),
         "store_raw_sequences": bool(r.u8())}
    assert r.i == len(b)
    return d


def decode_signature_chunk(b):
    r = _R(b)
    out = []
    for _ in range(r.u64()):
        d = {"name": r.string(), "mins": r.vec_u64(), "abunds": r.opt(r.vec_u64), "kmer_infos": {}}
        for _ in range(r.u64()):
            h = r.u64()
            ki = {"ksize": r.u64(), "hashval": r.u64(), "encoded_kmer": r.string(), "original_kmer_to_position": {}}
            for _ in range(r.u64()):
                o = r.string()
                ki["original_kmer_to_position"][o] = r.vec_u64()
            d["kmer_infos"][h] = ki
        d["raw_sequence"] = r.opt(r.string)
        out.append(d)
    assert r.i == len(b)
    return out
