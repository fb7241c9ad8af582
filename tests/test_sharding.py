"""CPU tests of the multi-GPU plumbing: shard arithmetic and the accumulator all-reduce
(world_size 2, gloo backend, 127.0.0.1 rendezvous)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracles
from libcxxdes_b200 import abi, sharding


def test_shard_of_covers_everything():
    for total, world in [(10, 3), (1 << 20, 8), (7, 8), (0, 2), (5, 1)]:
        spans = [sharding.shard_of(r, world, total) for r in range(world)]
        assert spans[0][0] == 0
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1
        assert spans[-1][0] + spans[-1][1] == total
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert sharding.weak_shard_of(3, 1 << 20) == (3 << 20, 1 << 20)


def test_accumulator_words_round_trip():
    acc = abi.Accumulators()
    acc.n_replications, acc.n_events, acc.n_errors = 5, 1 << 40, 0
    acc.trace_hash_sum = (1 << 64) - 3
    acc.end_time_sum = -17
    acc.i64_sum[0], acc.i64_sum[1] = 11, -12
    acc.f64_sum[0], acc.f64_sum[1] = 0.125, -2.5
    acc.f64_sumsq[0], acc.f64_sumsq[1] = 1e300, 3.0
    back = sharding.words_to_accumulators(sharding.accumulators_to_words(acc))
    assert bytes(back) == bytes(acc)
    assert sharding.N_ACC_WORDS == 11 and sharding.N_INT_WORDS == 7


def _host_accumulators(cols):
    acc = abi.Accumulators()
    acc.n_replications = cols.n
    acc.n_events = int(cols.n_events.sum())
    acc.n_errors = int((cols.status != 0).sum())
    acc.trace_hash_sum = int(cols.trace_hash.sum(dtype=np.uint64))
    acc.end_time_sum = int(cols.end_time.sum())
    for k in range(abi.N_I64):
        acc.i64_sum[k] = int(cols.i64[k].sum())
    for k in range(abi.N_F64):
        acc.f64_sum[k] = float(cols.f64[k].sum())
        acc.f64_sumsq[k] = float((cols.f64[k] ** 2).sum())
    return acc


def _worker(rank, world, port_no, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = sharding.shard_of(rank, world, total)
    p = abi.MM1Params(5.0, 10.0, 300)
    cols = oracles.port_oracle().run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, first, count)
    words = torch.tensor(sharding.accumulators_to_words(_host_accumulators(cols)), dtype=torch.int64)
    sharding.all_reduce_words(words, dist)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), words.numpy())
    dist.destroy_process_group()


def test_two_rank_all_reduce_matches_single_shard(tmp_path, built):
    """Two gloo ranks each simulate half of the replications (on the CPU oracle -- this test is about
    the plumbing, not the kernels); the reduced accumulators equal those of the unsplit ensemble."""
    total = 101
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port_no, total, str(tmp_path)), nprocs=2, join=True)
    w0 = np.load(tmp_path / "rank0.npy")
    w1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(w0, w1)
    got = sharding.words_to_accumulators(w0.tolist())
    p = abi.MM1Params(5.0, 10.0, 300)
    want = _host_accumulators(oracles.port_oracle().run(abi.MODEL_MM1, p, oracles.MM1_CLOCK, 0x5EED, 0, total))
    assert got.n_replications == total and got.n_events == want.n_events
    assert got.trace_hash_sum == want.trace_hash_sum and got.end_time_sum == want.end_time_sum
    assert got.i64_sum[0] == want.i64_sum[0] and got.i64_sum[1] == want.i64_sum[1]
    assert abs(got.f64_sum[0] - want.f64_sum[0]) <= 1e-9 * abs(want.f64_sum[0])
    assert abs(got.f64_sumsq[0] - want.f64_sumsq[0]) <= 1e-9 * abs(want.f64_sumsq[0])
