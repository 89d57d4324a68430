"""CPU tests of the product library's host side: symbol table, geometry, frame weights,
header, grid files -- against the oracle and the reference's golden files.  No GPU calls."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import gcutil as u
import golden_cases as gc
import gridcount_b200 as g
from gridcount_b200 import _cabi as K

ROOT = u.ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gridcount_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gcb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", K.LIB_PATH]).decode()
    exported = set(re.findall(r" T (gcb_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    # and the ctypes table binds all of them
    assert sorted(K.SIGNATURES) == names
    assert K.lib.gcb_abi_version() == 1


@pytest.mark.skipif(g.device_count() > 0, reason="GPU present")
def test_no_cpu_fallback():
    geom = g.setup_geometry((3, 3, 3), 0.25)
    with pytest.raises(K.GcbError) as e:
        g.GridCounter(geom, np.arange(10), 10)
    assert e.value.code == -2
    assert b"no CPU fallback" in K.lib.gcb_last_error()


def test_geometry_matches_oracle_random():
    rng = np.random.default_rng(21)
    for it in range(3000):
        L = rng.uniform(1, 20, 3).astype(np.float32)
        dl = rng.uniform(0.02, 0.25, 3).astype(np.float32)
        kw, kwo = {}, {}
        if it % 3 == 1:
            R = float(rng.uniform(0.3, 0.5) * min(L[0], L[1])); z1 = float(0.2 * L[2]); z2 = float(0.9 * L[2])
            kw = dict(radius=R, z1=z1, z2=z2)
            kwo = dict(setmask=u.SET_R | u.SET_Z1 | u.SET_Z2, radius=R, z1=z1, z2=z2)
        if it % 3 == 2:
            cp = tuple(float(v) for v in rng.uniform(0.3, 0.7, 3) * L)
            kw = dict(cpoint=cp, z1=float(L[2]), z2=0.25)
            kwo = dict(setmask=u.SET_CPOINT | u.SET_Z1 | u.SET_Z2, cpoint=cp, z1=float(L[2]), z2=0.25)
        geom = g.setup_geometry(L, dl, **kw)
        og, ocav = u.oracle_geometry(L, dl, **kwo)
        assert bytes(geom.tg)[:48] == bytes(og)[:48], (L, dl, kw)      # mx, a, b, delta
        assert bytes(geom.cav)[12:] == bytes(ocav)[12:], (L, dl, kw)   # cpoint, radius, z1, z2, vol
        assert geom.edge_vulnerable == bool(u.oracle().gco_edge_vulnerable(C.byref(og)))


def test_frame_weights_and_header_match_oracle():
    o = u.oracle()
    rng = np.random.default_rng(4)
    for dtw in (0, 1):
        t = (np.float32(7.0) + np.arange(50, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
        t[20:] += np.float32(0.05)
        w, T = g.frame_weights(t, dtw)
        ow = np.zeros_like(w)
        oT = C.c_double(0)
        o.gco_frame_weights(u.fptr(t), len(t), dtw, u.fptr(ow), C.byref(oT))
        u.assert_bits_equal(w, ow, "weights")
        assert T == oT.value
        geom = g.setup_geometry((2.98, 3.31, 4.07), (0.2, 0.2, 0.25))
        og, ocav = u.oracle_geometry((2.98, 3.31, 4.07), (0.2, 0.2, 0.25))
        sumP = float(rng.uniform(10, 1e6))
        want = C.create_string_buffer(256)
        o.gco_header(want, b"sub title", T, b"OW", C.byref(ocav), C.byref(og), sumP, b"traj.xtc", dtw)
        assert g.header_string(geom, T, sumP, "OW", "traj.xtc", "sub title", dtw) == want.value
    # streaming in pieces gives the same weights as one call
    t = np.cumsum(rng.uniform(0.5, 1.5, 40)).astype(np.float32)
    w_all, T_all = g.frame_weights(t, 1)
    tl = C.c_float(t[0] - (t[1] - t[0]))
    T = C.c_double(0)
    parts = []
    for lo, hi in ((0, 13), (13, 14), (14, 40)):
        w = np.zeros(hi - lo, np.float32)
        seg = np.ascontiguousarray(t[lo:hi])
        K.check(K.lib.gcb_frame_weights(u.fptr(seg), hi - lo, 1, C.byref(tl), u.fptr(w), C.byref(T)), "fw")
        parts.append(w)
    u.assert_bits_equal(np.concatenate(parts), w_all, "streamed weights")
    assert T.value == T_all


@pytest.mark.parametrize("name", list(gc.CASES) + ["gridcalc"])
def test_grid_file_roundtrip_byte_exact(name, tmp_path):
    fn = "avg.dat" if name == "gridcalc" else "grid.dat"
    src = os.path.join(u.GOLDEN_DIR, name, fn)
    raw = open(src, "rb").read()
    tg, hdr, grid = g.grid_read(src)
    assert g.xdr_encode(tg, hdr, grid_zfast=grid) == raw
    xfast = np.ascontiguousarray(grid.transpose(2, 1, 0))
    assert g.xdr_encode(tg, hdr, grid_xfast=xfast) == raw
    p = tmp_path / "out.dat"
    g.grid_write(p, tg, hdr, grid_zfast=grid)
    assert p.read_bytes() == raw


def test_grid_read_rejects_bad_files(tmp_path):
    raw = bytearray(open(os.path.join(u.GOLDEN_DIR, "cube", "grid.dat"), "rb").read())
    bad = tmp_path / "v1.dat"
    v1 = bytearray(raw); v1[3] = 1
    bad.write_bytes(bytes(v1))
    with pytest.raises(K.GcbError) as e:
        g.grid_read(bad)
    assert e.value.code == -5                      # version_check: v1 is rejected (xdr_grid.c:28-34)
    v3 = bytearray(raw); v3[3] = 3
    bad.write_bytes(bytes(v3))
    with pytest.raises(K.GcbError):
        g.grid_read(bad)
    bad.write_bytes(bytes(raw[:len(raw) // 2]))
    with pytest.raises(K.GcbError):
        g.grid_read(bad)
    with pytest.raises(K.GcbError) as e:
        g.grid_read(tmp_path / "missing.dat")
    assert e.value.code == -4


@pytest.mark.parametrize("name,run", [("cube", "spc"), ("pore", "molar_crop")])
def test_plt_and_ascii_match_reference(name, run, tmp_path):
    import test_oracle_golden as tog
    og, grid, aa, arrs, scal = tog.run_oracle_analysis(name, run)     # oracle supplies the (cropped) grid
    tg = K.TGrid.from_buffer_copy(bytes(og))
    p = tmp_path / "plt.dat"
    K.check(K.lib.gcb_plt_write(str(p).encode(), C.byref(tg), u.fptr(grid), scal["unit_value"]), "plt")
    assert p.read_bytes() == open(os.path.join(u.GOLDEN_DIR, name, run, "plt.dat"), "rb").read()
    golden_ascii = os.path.join(u.GOLDEN_DIR, name, run, "gridasc.dat")
    if os.path.exists(golden_ascii):
        q = tmp_path / "asc.dat"
        K.check(K.lib.gcb_ascii_write(str(q).encode(), C.byref(tg), u.fptr(grid), scal["unit_value"]), "ascii")
        assert q.read_bytes() == open(golden_ascii, "rb").read()


def test_dens_units():
    tg = g.setup_geometry((3, 3, 3), 0.25).tg
    assert [K.lib.gcb_dens_unit(i, C.byref(tg)) for i in range(4)] == [1.0, np.float32(32.3227), np.float32(0.602214), 1000.0]
    assert K.lib.gcb_dens_unit(4, C.byref(tg)) == np.float32(0.25 ** 3)


@pytest.mark.parametrize("threads", [0, 1, 3, 16])
def test_pack_frames_copies_the_index_group(threads):
    """gcb_pack_frames: out[f][i] = x[f][gindex[i]], the reference loop's reads (g_ri3Dc.c:425-426) laid out densely;
    pure host code, the packing step of gcb_counter_add_frames by itself"""
    import gridcount_b200 as g
    rng = np.random.default_rng(3)
    cases = [(2700, np.arange(0, 2700, 3)), (2700, np.arange(2700)), (1000, np.sort(rng.choice(1000, 137, replace=False))),
             (5, np.array([4, 0, 0, 2])), (60000, np.arange(1, 60000, 3)), (300, np.array([], np.int64)), (7, np.array([6]))]
    for natoms, gi in cases:
        for nframes in (0, 1, 7, 130):
            x = rng.random((nframes, natoms, 3), dtype=np.float32)
            out = g.pack_frames(x, gi.astype(np.int32), threads)
            assert out.shape == (nframes, len(gi), 3)
            assert np.array_equal(out.view(np.uint32), x[:, gi.astype(np.int64), :].view(np.uint32)), (natoms, nframes)
    x = np.zeros((2, 10, 3), np.float32)
    with pytest.raises(Exception):
        g.pack_frames(x, np.array([3, 10], np.int32), threads)       # entry outside [0, natoms)


def test_pack_frames_from_many_threads_at_once():
    """the pool takes one job at a time: callers on several threads (one counter per GPU in host/g_ri3Dc -gpus) queue up"""
    import threading
    import gridcount_b200 as g
    rng = np.random.default_rng(4)
    x = rng.random((40, 30000, 3), dtype=np.float32)
    gi = np.arange(0, 30000, 3, dtype=np.int32)
    want = x[:, gi, :]
    bad = []

    def work():
        for _ in range(20):
            if not np.array_equal(g.pack_frames(x, gi, 4), want):
                bad.append(1)

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not bad
