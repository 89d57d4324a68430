"""CPU tests: the plain-C restatement (oracle/gc_oracle.c) against golden outputs
of the UNMODIFIED reference (tests/golden, made by tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import gcutil as u
import golden_cases as gc


def _read(path):
    with open(path, "rb") as f:
        return f.read()


def _f64(path):
    return np.fromfile(path, dtype=np.float64)


def oracle_density(name):
    """Run the whole g_ri3Dc path through the oracle -> (geom, cavity, grid f32 zfast, header bytes, xdr bytes)"""
    o = u.oracle()
    c = gc.CASES[name]
    ga = gc.parse_gargs(c["gargs"])
    t, boxes, x, gindex = gc.case_inputs(name)
    g, cav = u.oracle_geometry(c["L"], c["delta"], ga["setmask"], ga["cpoint"], ga["R"], ga["z1"], ga["z2"])
    w = np.zeros(c["nframes"], np.float32)
    T = C.c_double(0)
    o.gco_frame_weights(u.fptr(t), c["nframes"], ga["dtweight"], u.fptr(w), C.byref(T))
    grid = np.zeros(g.shape(), np.float32)
    sumP = C.c_double(0)
    ovf = C.c_uint64(0)
    o.gco_bin_frames(C.byref(g), u.fptr(grid), u.fptr(x), c["nframes"], c["natoms"], u.iptr(gindex), len(gindex),
                     u.fptr(w), C.byref(sumP), None, C.byref(ovf))
    assert ovf.value == 0
    o.gco_normalise(C.byref(g), u.fptr(grid), T.value)
    hdr = C.create_string_buffer(256)
    grp = b"OW" if c["stride"] != 1 else b"System"
    o.gco_header(hdr, ga["subtitle"].encode(), T.value, grp, C.byref(cav), C.byref(g), sumP.value, b"traj.raw",
                 ga["dtweight"])
    n = o.gco_xdr_size(C.byref(g), hdr.value)
    buf = np.zeros(n, np.uint8)
    m = o.gco_xdr_encode(C.byref(g), u.fptr(grid), hdr.value, buf.ctypes.data_as(u.c_u8_p), n)
    assert m == n
    return g, cav, grid, hdr.value, buf.tobytes()


def decode_golden(path):
    o = u.oracle()
    raw = np.frombuffer(_read(path), np.uint8).copy()
    g = u.Geom()
    hdr = C.create_string_buffer(257)
    gp = u.c_float_p()
    rc = o.gco_xdr_decode(raw.ctypes.data_as(u.c_u8_p), len(raw), C.byref(g), hdr, C.byref(gp))
    assert rc == 0
    grid = np.ctypeslib.as_array(gp, shape=(g.cells(),)).reshape(g.shape()).copy()
    o.free(gp)
    return g, hdr.value, grid


@pytest.mark.parametrize("name", list(gc.CASES))
def test_inputs_unchanged(name):
    want = open(os.path.join(u.GOLDEN_DIR, name, "inputs.sha256")).read().strip()
    assert gc.inputs_digest(name) == want, "synthetic generator drifted: golden inputs no longer reproducible"


@pytest.mark.parametrize("name", list(gc.CASES))
def test_density_file_byte_exact(name):
    """geometry + frame weights + f32 accumulation + normalisation + header + XDR == reference file"""
    _, _, _, _, xdr = oracle_density(name)
    want = _read(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    assert len(xdr) == len(want)
    assert xdr == want


@pytest.mark.parametrize("name", list(gc.CASES))
def test_xdr_roundtrip(name):
    o = u.oracle()
    want = _read(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    g, hdr, grid = decode_golden(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    n = o.gco_xdr_size(C.byref(g), hdr)
    buf = np.zeros(n, np.uint8)
    assert o.gco_xdr_encode(C.byref(g), u.fptr(grid), hdr, buf.ctypes.data_as(u.c_u8_p), n) == n
    assert buf.tobytes() == want


def run_oracle_analysis(name, run):
    o = u.oracle()
    aa = gc.parse_aargs(gc.CASES[name]["analyses"][run])
    g, hdr, grid = decode_golden(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    if aa["setmask"]:
        user = u.Cavity()
        user.radius, user.z1, user.z2 = aa["R"], aa["z1"], aa["z2"]
        # hand the grid to C-owned memory because gco_readjust frees/replaces it
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p
        mem = libc.malloc(grid.nbytes)
        C.memmove(mem, grid.ctypes.data, grid.nbytes)
        gp = C.cast(mem, u.c_float_p)
        rc = o.gco_readjust(C.byref(g), C.byref(gp), C.byref(user), aa["setmask"])
        assert rc == 1
        grid = np.ctypeslib.as_array(gp, shape=(g.cells(),)).reshape(g.shape()).copy()
        o.free(gp)
    an = u.Analysis()
    an.unit = u.UNITS[aa["unit"]]
    an.min_occ = aa["minocc"]
    an.rad_ba_nr, an.rad_ba_step = aa["radnr"], aa["radstep"]
    assert o.gco_analyse(C.byref(g), u.fptr(grid), C.byref(an)) == 0
    arrs = u.analysis_arrays(an, g)
    scal = dict(volume=an.volume, reff=an.reff, sumP=an.sumP, sumH=an.sumH, dV=an.dV, average=an.average,
                ivol=an.ivol, iradNR=an.iradNR, radius=an.geometry.radius, unit_value=an.unit_value)
    o.gco_analysis_free(C.byref(an))
    return g, grid, aa, arrs, scal


ANALYSES = [(n, r) for n, c in gc.CASES.items() for r in c["analyses"]]


def check_analysis_against_golden(g, grid, aa, arrs, scal, rdir):
    """Compare arrays/scalars with the full-precision capture of the reference's own output."""
    mx, my, mz = g.mx[0], g.mx[1], g.mx[2]
    nr = scal["iradNR"]
    d = lambda a: np.asarray(a, np.float32).astype(np.float64)
    so = _f64(os.path.join(rdir, "stdout.f64"))
    # DATA_RESULT: V R* sumP sumH dV / average iVolume   (a_ri3Dc.c:768-770)
    want = [d(scal["volume"]), d(scal["reff"]), scal["sumP"], scal["sumH"], scal["dV"], d(scal["average"])]
    u.assert_bits_equal(np.array(want, np.float64), so[:6], "DATA_RESULT scalars")
    assert so[6] == np.float64(np.float32(scal["ivol"])) * scal["dV"]
    # profile: z, R*, R*/R
    prof = _f64(os.path.join(rdir, "profile.xvg.f64")).reshape(mz, 3)
    u.assert_bits_equal(d(arrs["profile"]), prof[:, :2], "profile")
    u.assert_bits_equal((arrs["profile"][:, 1] / np.float32(scal["radius"])).astype(np.float64), prof[:, 2], "R*/R")
    u.assert_bits_equal(d(arrs["xyp"]).T, _f64(os.path.join(rdir, "xyp.dat.f64")).reshape(my, mx), "xyp")
    if "xzp" in aa["files"]:
        u.assert_bits_equal(d(arrs["xzp"]).T, _f64(os.path.join(rdir, "xzp.dat.f64")).reshape(mz, mx), "xzp")
    if "yzp" in aa["files"]:
        u.assert_bits_equal(d(arrs["yzp"]).T, _f64(os.path.join(rdir, "yzp.dat.f64")).reshape(mz, my), "yzp")
    if "hxyp" in aa["files"]:
        u.assert_bits_equal(d(arrs["hxyp"]).T, _f64(os.path.join(rdir, "hxyp.dat.f64")).reshape(my, mx), "hxyp")
    cols = list(range(-nr + 1, nr)) if aa["mirror"] else list(range(nr))
    idx = np.abs(np.array(cols))
    u.assert_bits_equal(d(arrs["rzp"])[idx, :].T, _f64(os.path.join(rdir, "rzp.dat.f64")).reshape(mz, len(cols)), "rzp")
    if "hrzp" in aa["files"]:
        u.assert_bits_equal(d(arrs["hrzp"])[idx, :].T,
                            _f64(os.path.join(rdir, "hrzp.dat.f64")).reshape(mz, len(cols)), "hrzp")
    zdf = _f64(os.path.join(rdir, "zdf.xvg.f64")).reshape(mz, 3)
    u.assert_bits_equal(d(arrs["zdf"]), zdf[:, :2], "zdf")
    lz = _f64(os.path.join(rdir, "lzdf.xvg.f64")).reshape(mz, 4)
    u.assert_bits_equal(d(arrs["lzdf"]), lz[:, :2], "lzdf")
    u.assert_bits_equal(d(arrs["rdf"]), _f64(os.path.join(rdir, "rdf.xvg.f64")).reshape(nr, 2), "rdf")
    if "hrdf" in aa["files"]:
        u.assert_bits_equal(d(arrs["hrdf"]), _f64(os.path.join(rdir, "hrdf.xvg.f64")).reshape(nr, 2), "hrdf")


@pytest.mark.parametrize("name,run", ANALYSES)
def test_analysis_bit_exact(name, run):
    g, grid, aa, arrs, scal = run_oracle_analysis(name, run)
    check_analysis_against_golden(g, grid, aa, arrs, scal, os.path.join(u.GOLDEN_DIR, name, run))


@pytest.mark.parametrize("name,run", [(n, r) for n, r in ANALYSES if "-plt" in gc.CASES[n]["analyses"][r]])
def test_plt_byte_exact(name, run):
    o = u.oracle()
    g, grid, aa, arrs, scal = run_oracle_analysis(name, run)
    n = 4 * (11 + g.cells())
    buf = np.zeros(n, np.uint8)
    assert o.gco_plt_encode(C.byref(g), u.fptr(grid), scal["unit_value"], buf.ctypes.data_as(u.c_u8_p), n) == n
    assert buf.tobytes() == _read(os.path.join(u.GOLDEN_DIR, name, run, "plt.dat"))


def test_gridcalc_byte_exact():
    o = u.oracle()
    grids, tws, geom = [], [], None
    for n in gc.GRIDCALC["inputs"]:
        g, hdr, grid = decode_golden(os.path.join(u.GOLDEN_DIR, n, "grid.dat"))
        grids.append(np.ascontiguousarray(grid))
        tws.append(g.tweight)
        geom = g
    ptrs = (u.c_float_p * len(grids))(*[u.fptr(a) for a in grids])
    tw = np.asarray(tws, np.float32)
    avg = np.zeros(geom.shape(), np.float32)
    two = C.c_float(0)
    o.gco_gridcalc(len(grids), ptrs, u.fptr(tw), geom.mx[0], geom.mx[1], geom.mx[2], u.fptr(avg), C.byref(two))
    want_g, want_hdr, want_grid = decode_golden(os.path.join(u.GOLDEN_DIR, "gridcalc", "avg.dat"))
    assert want_g.tweight == two.value
    u.assert_bits_equal(avg, want_grid, "a_gridcalc average")
