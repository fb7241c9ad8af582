"""CPU tests of the arithmetic shared by host and device (include/cxxdes_b200/shared)."""
import ctypes as C
import math
import struct

import numpy as np
import pytest

# Philox4x32-10 known-answer vectors (Random123 kat_vectors), ctr | key -> out
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.cxxdes_oracle_philox(c, k, o)
    return tuple(o)


def test_philox_known_answers(port):
    for ctr, key, out in PHILOX_KAT:
        assert _philox(port.lib, ctr, key) == out


def test_stream_addressing(port):
    """draw d = words {0,1} (even) / {2,3} (odd) of block d>>1 with the documented key/counter."""
    seed, rep, stream = 0x0123456789ABCDEF, 0x1_0000_0042, 7
    n = 16
    out = np.zeros(n, np.uint64)
    port.lib.cxxdes_oracle_draw_bits(C.c_uint64(seed), C.c_uint64(rep), C.c_uint32(stream), C.c_uint64(0),
                                     C.c_uint64(n), out.ctypes.data_as(C.c_void_p))
    for d in range(n):
        blk = d >> 1
        w = _philox(port.lib, (blk & 0xFFFFFFFF, stream | ((blk >> 32) << 16), rep & 0xFFFFFFFF, rep >> 32),
                    (seed & 0xFFFFFFFF, seed >> 32))
        lo, hi = (w[2], w[3]) if d & 1 else (w[0], w[1])
        assert int(out[d]) == (hi << 32) | lo


def test_unit_complement_range(port):
    bits = np.array([0, 1 << 12, (1 << 64) - 1, 1 << 63, 0xFFF], np.uint64)
    out = np.zeros(len(bits))
    port.lib.cxxdes_oracle_unit_complement(bits.ctypes.data_as(C.c_void_p), C.c_uint64(len(bits)),
                                           out.ctypes.data_as(C.c_void_p))
    assert out[0] == 1.0                      # u = 0 -> w = 1 -> variate 0
    assert out[1] == 1.0 - 2.0 ** -52
    assert out[2] == 2.0 ** -52               # never 0: log stays finite
    assert out[3] == 0.5
    assert out[4] == 1.0                      # low 12 bits are discarded


def _log(port, w):
    w = np.ascontiguousarray(w, np.float64)
    out = np.zeros_like(w)
    port.lib.cxxdes_oracle_log(w.ctypes.data_as(C.c_void_p), C.c_uint64(len(w)), out.ctypes.data_as(C.c_void_p))
    return out


def test_log_accuracy(port):
    """Shared log vs mpmath: < 1.5 ulp away from the table edges near 1, tiny absolute error everywhere."""
    import mpmath
    mpmath.mp.prec = 120
    rng = np.random.default_rng(7)
    w = np.concatenate([rng.random(20000), 2.0 ** -rng.integers(0, 52, 2000) * (0.5 + rng.random(2000) / 2),
                        1.0 - rng.random(2000) * 2.0 ** -10, [1.0, 2.0 ** -52, 0.5, 0.6875, 0.7, 1.0 - 2.0 ** -52]])
    w = w[(w > 0) & (w <= 1.0)]
    got = _log(port, w)
    worst_ulp, worst_abs = 0.0, 0.0
    for x, g in zip(w.tolist(), got.tolist()):
        exact = mpmath.log(mpmath.mpf(x))
        err = abs(mpmath.mpf(g) - exact)
        if x > 0.68:
            worst_abs = max(worst_abs, float(err))
        else:
            ulp = math.ulp(float(exact))
            worst_ulp = max(worst_ulp, float(err) / ulp)
    assert worst_ulp < 1.5, worst_ulp
    assert worst_abs < 2.0 ** -54, worst_abs
    assert _log(port, [1.0])[0] == 0.0


def test_markstein_divide_equals_ieee(port):
    rng = np.random.default_rng(11)
    for d in (5.0, 10.0, 9.9, 0.1, 3.0, 1.0 / 3.0, 7.123456789, 2.0 / 64.0):
        x = np.concatenate([rng.random(200000) * 40.0, rng.random(1000) * 1e-12, [0.0, 36.04365338911715]])
        out = np.zeros_like(x)
        port.lib.cxxdes_oracle_divide(x.ctypes.data_as(C.c_void_p), C.c_double(d), C.c_uint64(len(x)),
                                      out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out.view(np.uint64), (x / d).view(np.uint64)), d


def test_exponential_moments(port):
    n = 400000
    v = np.zeros(n)
    t = np.zeros(n, np.int64)
    port.lib.cxxdes_oracle_exponential_ticks(C.c_uint64(0x5EED), C.c_uint64(3), C.c_uint32(1), C.c_double(5.0),
                                             C.c_int64(1000), C.c_uint64(0), C.c_uint64(n),
                                             v.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    assert abs(v.mean() - 0.2) < 0.002 and abs(v.var() - 0.04) < 0.002
    assert np.array_equal(t, (v * 1000.0).astype(np.int64))   # real_to_sim truncates toward zero
    assert v.min() >= 0.0


def test_trace_hash_is_order_sensitive(port):
    def h(events):
        t = np.array([e[0] for e in events], np.int64)
        p = np.array([e[1] for e in events], np.int64)
        k = np.array([e[2] for e in events], np.uint32)
        pr = np.array([e[3] for e in events], np.uint32)
        f = port.lib.cxxdes_oracle_trace_hash
        f.restype = C.c_uint64
        return f(t.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                 pr.ctypes.data_as(C.c_void_p), C.c_uint64(len(events)))
    a = [(0, 0, 0, 1), (0, 0, 0, 2), (5, 0, 1, 0)]
    assert h(a) != h([a[1], a[0], a[2]])          # same multiset, different order
    assert h(a) != h([(0, 0, 0, 1), (0, 0, 0, 2), (5, 1, 1, 0)])
    assert h(a) == h(list(a))


def test_restated_heap_matches_libstdcxx(port):
    """The restated __push_heap/__adjust_heap pops exactly what std::priority_queue pops, ties included."""
    f = port.lib.cxxdes_oracle_heap_check
    f.restype = C.c_uint64
    for seed, key_range, max_size in [(1, 4, 3), (2, 8, 64), (3, 3, 65), (4, 1000, 48), (5, 2, 7), (6, 16, 200)]:
        assert f(C.c_uint64(seed), C.c_uint64(200000), C.c_uint32(key_range), C.c_uint32(max_size)) == 0


def test_device_heaps_match_std_priority_queue(tmp_path):
    """The event heaps the CUDA kernels run (csrc/device/engine.cuh, engine32.cuh with its two-level look-ahead),
    compiled with g++: every pop returns the same token as std::priority_queue with the reference's comparator
    (environment.ipp:247-263) on 7.7 M random operations with heavy key ties -- the tie order is libstdc++'s."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "device_heap_host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unused-function", "-Wno-unused-variable",
                    f"-I{root / 'include'}", f"-I{root / 'libcxxdes_b200' / 'csrc'}",
                    str(root / "tests" / "native" / "device_heap_host.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatching runs" in r.stdout
