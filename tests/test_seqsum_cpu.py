"""The integer recurrence behind the GPU's exact sequential sums (gridcount_b200/csrc/seqsum.cuh), checked on
the CPU: the same term/then/apply primitives, driven chunk by chunk, must reproduce a plain sequential
f32 / f64 accumulation (a_ri3Dc.c:499-508, 633) bit for bit -- ties, saturation and subnormals included."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("seqsum") / "libseqsum_harness.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                           "-I", os.path.join(ROOT, "gridcount_b200", "csrc"),
                           os.path.join(ROOT, "tests", "seqsum_harness.cpp"), "-o", str(so)])
    L = C.CDLL(str(so))
    fp = C.POINTER(C.c_float)
    for name, res in (("ss_chain_f32", C.c_float), ("ss_chain_f64", C.c_double)):
        f = getattr(L, name)
        f.argtypes = [fp, C.c_size_t, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_int)]
        f.restype = res
    L.ss_inc_chain_f32.argtypes = [C.c_float, C.c_double, C.c_longlong]
    L.ss_inc_chain_f32.restype = C.c_float
    return L


def seq32(x):
    return np.add.accumulate(x.astype(np.float32), dtype=np.float32)[-1] if len(x) else np.float32(0)


def seq64(x):
    return np.add.accumulate(x.astype(np.float64), dtype=np.float64)[-1] if len(x) else np.float64(0)


def cases():
    rng = np.random.default_rng(20261019)
    n = 200_000
    yield "uniform", rng.random(n).astype(np.float32)
    yield "tiny", (rng.random(n) * 1e-3).astype(np.float32)
    yield "sparse", np.where(rng.random(n) < 0.8, 0, rng.random(n) * 100).astype(np.float32)
    yield "halves (ties everywhere)", (rng.integers(0, 7, n) * 0.5).astype(np.float32)
    yield "ones past 2^24 (f32 saturates)", np.ones((1 << 24) + 4097, np.float32)
    yield "threes past 2^24", np.full((1 << 23) + 999, 3.0, np.float32)
    yield "wide range", np.exp(rng.uniform(-60, 20, n)).astype(np.float32)
    yield "subnormal start", np.concatenate([np.zeros(77, np.float32), np.full(300, 1e-42, np.float32),
                                             (rng.random(3000) * 1e-38).astype(np.float32),
                                             rng.random(5000).astype(np.float32)])
    yield "big first", np.concatenate([np.full(3, 2.0 ** 24, np.float32), np.ones(5000, np.float32),
                                       np.full(5000, 3.0, np.float32), np.full(100, 1.5, np.float32)])
    yield "24-bit integers / 1024", (rng.integers(0, 2 ** 24, n).astype(np.float32) * np.float32(2.0 ** -10))
    yield "minus zero", np.where(rng.random(5000) < 0.5, -0.0, rng.random(5000)).astype(np.float32)
    yield "density-like", (rng.poisson(3.0, n) / np.float32(0.001 * 40.0)).astype(np.float32)
    yield "empty", np.zeros(0, np.float32)
    yield "all zero", np.zeros(1000, np.float32)


@pytest.mark.parametrize("name,x", list(cases()), ids=[c[0] for c in cases()])
@pytest.mark.parametrize("chunk", [1, 7, 256, 8192])
def test_chunked_recurrence_equals_sequential_sum(lib, name, x, chunk):
    if chunk == 1 and len(x) > 300_000:
        pytest.skip("covered by the larger chunks")
    x = np.ascontiguousarray(x, np.float32)
    cr, bad = C.c_long(0), C.c_int(0)
    p = x.ctypes.data_as(C.POINTER(C.c_float))
    got32 = np.float32(lib.ss_chain_f32(p, len(x), chunk, C.byref(cr), C.byref(bad)))
    assert not bad.value
    assert got32.view(np.uint32) == np.float32(seq32(x)).view(np.uint32), (name, got32, seq32(x), cr.value)
    got64 = np.float64(lib.ss_chain_f64(p, len(x), chunk, C.byref(cr), C.byref(bad)))
    assert not bad.value
    assert got64.view(np.uint64) == np.float64(seq64(x)).view(np.uint64), (name, got64, seq64(x), cr.value)
    assert cr.value <= 2200          # at most one real addition per exponent of the accumulator


@pytest.mark.parametrize("poison", [-1.0, np.inf, np.nan])
def test_precondition_violations_are_reported(lib, poison):
    x = np.ones(1000, np.float32)
    x[500] = poison
    cr, bad = C.c_long(0), C.c_int(0)
    lib.ss_chain_f32(x.ctypes.data_as(C.POINTER(C.c_float)), len(x), 64, C.byref(cr), C.byref(bad))
    assert bad.value == 1


def test_overflow_is_reported(lib):
    x = np.full(10, 3e38, np.float32)
    cr, bad = C.c_long(0), C.c_int(0)
    lib.ss_chain_f32(x.ctypes.data_as(C.POINTER(C.c_float)), len(x), 4, C.byref(cr), C.byref(bad))
    assert bad.value == 1


@pytest.mark.parametrize("c", [1.0 / 134.0, 1.0 / 400.0, 1.0 / 60.0, 0.5, 1.0, 1.0 / 3.0, 2.0 ** -30, 7.25, 1e-40])
@pytest.mark.parametrize("n", [0, 1, 5, 1000, 123457, 3_000_000])
def test_constant_increment_chain_closed_form(lib, c, n):
    """rweight[r] += 1.0/(real)mz, once per cell (a_ri3Dc.c:637): an f64 addition narrowed to f32 each time"""
    s = np.float32(0)
    inc = np.float64(c)
    for _ in range(min(n, 3_000_000)):
        s = np.float32(np.float64(s) + inc)
    got = np.float32(lib.ss_inc_chain_f32(0.0, c, n))
    assert got.view(np.uint32) == s.view(np.uint32), (c, n, got, s)
