"""kmerseek_b200 -- B200-native sketch-and-search hot path of kmerseek behind a C ABI.

Layout: csrc/ (CUDA kernels + C ABI, built into libkmerseek_b200.so), _ffi.py (ctypes binding),
index.py / search.py (host mirror of the reference's ProteomeIndex and search), shard.py (multi-GPU),
synth.py (synthetic proteomes for tests and bench).  There is no CPU fallback.
"""
from .errors import (BuilderError, CapacityError, CudaError, IndexError_, InvalidAminoAcid, InvalidKsize,  # noqa: F401
                     InvalidMoltype, NoDevice, NotFinalized, ParseError, ValidationError)
from .index import (KmerInfo, ProteinSignature, Proteome, ProteomeIndex, ProteomeIndexBuilder, id_of_mins,  # noqa: F401
                    max_hash, md5_of_mins, translate)
from .search import SearchResult, manysearch_rows, search, stitch_hits  # noqa: F401

__version__ = "0.1.0"
