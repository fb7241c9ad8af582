"""Multi-GPU plumbing of the ensemble: which replications a rank owns and how the per-shard
accumulators are summed.  Replications are independent (one `environment` each in the reference,
no shared state), so the data path needs no collective; the only exchange is one all-reduce of
the 88-byte accumulator struct at the end of a run (torch.distributed: NCCL on GPUs, gloo in the
CPU tests)."""
import ctypes as C

from . import abi

N_ACC_WORDS = C.sizeof(abi.Accumulators) // 8
N_INT_WORDS = 5 + abi.N_I64            # n_replications .. i64_sum[] are integers, the rest doubles


def shard_of(rank, world, total_replications):
    """Contiguous split of [0, total) over `world` ranks; the first `total % world` ranks get one more."""
    base, extra = divmod(int(total_replications), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def weak_shard_of(rank, replications_per_rank):
    """Weak scaling: every rank owns the same number of replications, global ids are disjoint."""
    return rank * int(replications_per_rank), int(replications_per_rank)


def accumulators_to_words(acc):
    """The struct as 11 little-endian 64-bit words (7 integer words, 4 IEEE doubles)."""
    return list((C.c_int64 * N_ACC_WORDS).from_buffer_copy(bytes(acc)))


def words_to_accumulators(words):
    buf = (C.c_int64 * N_ACC_WORDS)(*[int(w) for w in words])
    return abi.Accumulators.from_buffer_copy(bytes(buf))


def all_reduce_words(words_tensor, dist):
    """Sum an int64 tensor holding the accumulator words across ranks, in place: integer words
    with wrapping int64 addition (exact), the trailing words reinterpreted as float64."""
    import torch
    dist.all_reduce(words_tensor[:N_INT_WORDS], op=dist.ReduceOp.SUM)
    dist.all_reduce(words_tensor[N_INT_WORDS:].view(torch.float64), op=dist.ReduceOp.SUM)
    return words_tensor


class DevicePointerView:
    """Zero-copy torch view (via __cuda_array_interface__) of the accumulator struct the library
    keeps in device memory, so NCCL reduces it where it lies."""

    def __init__(self, ptr, n_int64=N_ACC_WORDS):
        self.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
