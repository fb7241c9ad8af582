"""Minimal ctypes binding of the four NCCL calls a host needs to give cxxdes_b200_reduce_allgpus a communicator
(ncclGetUniqueId on one rank, the 128 bytes handed to the others by whatever channel the host has, ncclCommInitRank
on every rank) -- what examples/mm1_multi_gpu.cpp does from C++.  The library itself resolves ncclAllReduce from the
same libnccl.so.2 at its first reduce_allgpus call; nothing here computes anything."""
import ctypes as C

UNIQUE_ID_BYTES = 128


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * UNIQUE_ID_BYTES)]


_lib = None


def load():
    """The libnccl.so.2 already in the process (PyTorch's bundled one once torch is imported), else the system's."""
    global _lib
    if _lib is not None:
        return _lib
    last = None
    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            lib = C.CDLL(name, mode=C.RTLD_GLOBAL)
            break
        except OSError as e:   # pragma: no cover
            last = e
    else:
        raise ImportError(f"libnccl.so.2 not found: {last}")
    lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    lib.ncclCommCount.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.ncclGetErrorString.restype = C.c_char_p
    lib.ncclGetErrorString.argtypes = [C.c_int]
    for f in ("ncclGetUniqueId", "ncclCommInitRank", "ncclCommDestroy", "ncclCommCount"):
        getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib.ncclGetErrorString(rc).decode()}")


def get_unique_id():
    lib = load()
    uid = UniqueId()
    _check(lib, lib.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
    return bytes(uid.internal)


def comm_init_rank(n_ranks, unique_id, rank):
    """ncclCommInitRank on the CURRENT CUDA device; every rank must call it with the same 128 bytes."""
    lib = load()
    uid = UniqueId()
    C.memmove(uid.internal, bytes(unique_id), UNIQUE_ID_BYTES)
    comm = C.c_void_p()
    _check(lib, lib.ncclCommInitRank(C.byref(comm), int(n_ranks), uid, int(rank)), "ncclCommInitRank")
    return comm.value


def comm_count(comm):
    lib = load()
    n = C.c_int()
    _check(lib, lib.ncclCommCount(C.c_void_p(comm), C.byref(n)), "ncclCommCount")
    return n.value


def comm_destroy(comm):
    if comm:
        load().ncclCommDestroy(C.c_void_p(comm))
