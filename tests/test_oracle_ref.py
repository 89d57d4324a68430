"""CPU tests: the restatement against the LIVE compiled reference (oracle/_ref/libgcref.so).
Skipped where oracle/_ref has not been built (it is built only where /root/reference exists;
the prebuilt binaries travel to the GPU box)."""
import ctypes as C

import numpy as np
import pytest

import gcutil as u

pytestmark = pytest.mark.skipif(not u.have_ref(), reason="oracle/_ref not built")


def _same_geom(g, cav, mx, a, b, c):
    f = lambda v: np.asarray(v[:], np.float32).view(np.uint32)
    return (list(g.mx) == list(mx) and np.array_equal(f(g.a), a.view(np.uint32))
            and np.array_equal(f(g.b), b.view(np.uint32)) and np.float32(cav.radius) == c[3]
            and np.float32(cav.z1) == c[4] and np.float32(cav.z2) == c[5]
            and np.float32(cav.vol).view(np.uint32) == c[6:7].view(np.uint32)[0])


def test_geometry_random_boxes():
    rng = np.random.default_rng(7)
    r = u.ref()
    for it in range(4000):
        L = rng.uniform(1, 20, 3).astype(np.float32)
        dl = rng.uniform(0.02, 0.25, 3).astype(np.float32)
        if it % 3 == 0:
            dl[:] = dl[0]
        if it % 5 == 0:
            L[1] = L[0]
        kw = {}
        if it % 4 == 1:
            kw = dict(setmask=u.SET_R | u.SET_Z1 | u.SET_Z2, radius=float(rng.uniform(0.3, 0.5) * min(L[0], L[1])),
                      z1=float(rng.uniform(0, 0.4) * L[2]), z2=float(rng.uniform(0.6, 1.0) * L[2]))
        if it % 4 == 3:
            kw = dict(setmask=u.SET_CPOINT | u.SET_Z1 | u.SET_Z2, cpoint=tuple(rng.uniform(0.3, 0.7, 3) * L),
                      z1=float(L[2]), z2=0.1)  # swapped on purpose (count.c:53-59)
        g, cav = u.oracle_geometry(L, dl, **kw)
        h, mx, a, b, d, c = u.ref_geometry(L, dl, **kw)
        ok = _same_geom(g, cav, mx, a, b, c)
        r.gcref_grid_free(h)
        assert ok, (L, dl, kw)


def test_named_baseline_geometries():
    for L, dl, want in [((3, 3, 3), 0.05, (60, 60, 60)), ((8, 8, 13.4), 0.1, (80, 80, 134)),
                        ((20, 20, 20), 0.05, (400, 400, 400)), ((10, 10, 10), 0.02, (500, 500, 500))]:
        g, cav = u.oracle_geometry(L, (dl,) * 3)
        assert g.shape() == want
        assert u.oracle().gco_edge_vulnerable(C.byref(g)) == 0


@pytest.mark.parametrize("w", [1.0, 2.0, 0.1, 0.002, 0.3333])
def test_binning_constant_weight(w):
    """f32 accumulation of the reference == seq_add(count) and == oracle float path, bit for bit"""
    o, r = u.oracle(), u.ref()
    L, dl = (2.98, 3.31, 2.5), (0.31, 0.29, 0.5)
    natoms, nframes = 600, 300
    gen = u.make_gen(99, L)
    x = u.gen_frames(gen, 0, nframes, natoms)
    x[3, 5, 0] = -0.5      # below the grid
    x[4, 6, 2] = 7.0       # above the grid
    gindex = np.arange(0, natoms, 2, dtype=np.int32)
    weights = np.full(nframes, w, np.float32)
    g, cav = u.oracle_geometry(L, dl)
    h, mx, a, b, d, c = u.ref_geometry(L, dl)
    sp_ref = C.c_double(0)
    r.gcref_bin_frames(h, u.fptr(x), nframes, natoms, u.iptr(gindex), len(gindex), u.fptr(weights), C.byref(sp_ref))
    ref_grid = np.ctypeslib.as_array(r.gcref_grid_data(h), shape=(g.cells(),)).copy()
    r.gcref_grid_free(h)

    grid = np.zeros(g.cells(), np.float32)
    sp = C.c_double(0)
    hits = np.zeros(nframes, np.uint64)
    o.gco_bin_frames(C.byref(g), u.fptr(grid), u.fptr(x), nframes, natoms, u.iptr(gindex), len(gindex),
                     u.fptr(weights), C.byref(sp), hits.ctypes.data_as(u.c_u64_p), None)
    u.assert_bits_equal(grid, ref_grid, "float accumulation")
    assert sp.value == sp_ref.value

    counts = np.zeros(g.cells(), np.uint32)
    o.gco_count_frames(C.byref(g), counts.ctypes.data_as(u.c_u32_p), u.fptr(x), nframes, natoms, u.iptr(gindex),
                       len(gindex), None, None)
    assert counts.sum() == hits.sum()
    emu = np.array([o.gco_seq_add_f32(0.0, w, int(n)) for n in counts], np.float32)
    u.assert_bits_equal(emu, ref_grid, "seq_add_f32(count) vs reference accumulation")
    # sumP: frame-ordered f64 chain rebuilt from per-frame hit counts
    s = 0.0
    for f in range(nframes):
        s = o.gco_seq_add_f64(s, float(np.float32(w)), int(hits[f]))
    assert s == sp_ref.value


def test_seq_add_matches_bruteforce():
    o = u.oracle()
    rng = np.random.default_rng(3)
    for w in [0.1, 1.0, 3.0, 0.002, 1e-3, 0.75, 5e-8, 1.1920929e-07, 16777216.0, 0.30000001192092896]:
        for v0 in [0.0, 1.0, 0.5, 1023.9999, 16777215.0, 3.3e-39]:
            for n in [0, 1, 2, 3, 7, 100, 4097, 20000, 70001]:
                v = np.float32(v0)
                ww = np.float32(w)
                for _ in range(n):
                    v = np.float32(v + ww)
                got = np.float32(o.gco_seq_add_f32(float(np.float32(v0)), float(ww), n))
                assert got.view(np.uint32) == v.view(np.uint32), (w, v0, n, got, v)
    # saturation far away: 2^24 additions of 1.0 and beyond
    assert o.gco_seq_add_f32(0.0, 1.0, 20_000_000) == 16777216.0
    for w, v0, n in [(0.1, 0.0, 10**6), (0.002, 0.0, 15625), (0.1, 7.0, 123457)]:
        v = np.float64(v0)
        for _ in range(n):
            v = v + np.float64(np.float32(w))
        assert o.gco_seq_add_f64(v0, float(np.float32(w)), n) == v


def test_xdr_file_read_by_reference(tmp_path):
    """a file written by the oracle encoder is read back identically by the reference's grid_read"""
    o, r = u.oracle(), u.ref()
    g, cav = u.oracle_geometry((2.0, 2.2, 3.1), (0.2, 0.2, 0.31))
    rng = np.random.default_rng(5)
    grid = rng.random(g.shape(), dtype=np.float32)
    g.tweight = 123.5
    hdr = b"hello{T}1ps"
    n = o.gco_xdr_size(C.byref(g), hdr)
    buf = np.zeros(n, np.uint8)
    assert o.gco_xdr_encode(C.byref(g), u.fptr(grid), hdr, buf.ctypes.data_as(u.c_u8_p), n) == n
    p = tmp_path / "o.dat"
    p.write_bytes(buf.tobytes())
    hb = C.create_string_buffer(300)
    h = r.gcref_grid_read(str(p).encode(), hb)
    assert h
    assert hb.value == hdr
    mx = np.zeros(3, np.int32); a = np.zeros(3, np.float32); b = np.zeros(3, np.float32); d = np.zeros(3, np.float32)
    r.gcref_grid_info(h, u.iptr(mx), u.fptr(a), u.fptr(b), u.fptr(d), None)
    assert list(mx) == list(g.mx)
    got = np.ctypeslib.as_array(r.gcref_grid_data(h), shape=(g.cells(),)).reshape(g.shape())
    u.assert_bits_equal(got, grid, "grid_read of oracle-encoded file")
    # and the reference's own writer reproduces the oracle's bytes
    q = tmp_path / "r.dat"
    r.gcref_grid_set_tweight(h, 123.5)
    assert r.gcref_grid_write(h, str(q).encode(), hb)
    assert q.read_bytes() == buf.tobytes()
    r.gcref_grid_free(h)
