"""IndexError variants of the reference (src/rust/errors.rs:4-55) plus the device-side codes of the C ABI."""
from . import _ffi


class IndexError_(Exception):
    """Base of every error the host layer raises (IndexError in the reference; renamed to avoid
    shadowing Python's builtin)."""
    status = None


class InvalidMoltype(IndexError_):
    status = 1


class InvalidAminoAcid(IndexError_):
    """errors.rs:14-15 -- "Invalid amino acid '{0}' found at position {1}" (1-based position)."""
    status = 2

    def __init__(self, message, char=None, pos=None, protein_index=None):
        super().__init__(message)
        self.char, self.pos, self.protein_index = char, pos, protein_index


class InvalidKsize(IndexError_):
    status = 3


class NoSavedState(IndexError_):
    status = 4


class IoError(IndexError_):
    status = 5


class Utf8Error(IndexError_):
    status = 6


class ParseError(IndexError_):
    status = 7


class BuilderError(IndexError_):
    status = 8

    def __init__(self, message):
        super().__init__(message if message.startswith("Builder error") else f"Builder error: {message}")


class ValidationError(IndexError_):
    status = 9


class NotFinalized(IndexError_):
    status = 10


class CudaError(IndexError_):
    status = 100


class NcclError(IndexError_):
    status = 101


class OutOfMemory(IndexError_):
    status = 102


class NoDevice(IndexError_):
    status = 103


class CapacityError(IndexError_):
    status = 104


_BY_STATUS = {c.status: c for c in (InvalidMoltype, InvalidAminoAcid, InvalidKsize, NoSavedState, IoError, Utf8Error,
                                    ParseError, BuilderError, ValidationError, NotFinalized, CudaError, NcclError,
                                    OutOfMemory, NoDevice, CapacityError)}


def check(status):
    """Raise the exception matching a non-zero ks_status, with the library's message."""
    if status == _ffi.KS_OK:
        return
    import ctypes as C
    L = _ffi.lib()
    msg = (L.ks_last_error_message() or b"").decode("utf-8", "replace")
    cls = _BY_STATUS.get(status, IndexError_)
    if cls is InvalidAminoAcid:
        ch, pos, prot = C.c_uint32(0), C.c_uint64(0), C.c_uint64(0)
        L.ks_last_error_detail(C.byref(ch), C.byref(pos), C.byref(prot))
        raise InvalidAminoAcid(msg, chr(ch.value), pos.value, prot.value)
    if cls is BuilderError:
        raise BuilderError(msg)
    raise cls(msg or _ffi.STATUS_NAMES.get(status, str(status)))
