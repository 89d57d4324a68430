"""GPU parity tests for K3: gcb_analyse / gcb_gridcalc through the C ABI against the golden output
of the UNMODIFIED reference's a_ri3Dc / a_gridcalc, and against the CPU oracle on larger grids."""
import ctypes as C
import os

import numpy as np
import pytest

import gcutil as u
import golden_cases as gc
import gridcount_b200 as g
from test_oracle_golden import ANALYSES, check_analysis_against_golden, decode_golden, _read

pytestmark = pytest.mark.gpu

ARRAYS = ["xyp", "xzp", "yzp", "hxyp", "zdf", "lzdf", "profile", "rweight", "rdf", "hrdf", "rzp", "hrzp"]


def gpu_analysis(tg, grid, aa):
    kw = {}
    if aa["setmask"] & u.SET_R: kw["radius"] = aa["R"]
    if aa["setmask"] & u.SET_Z1: kw["z1"] = aa["z1"]
    if aa["setmask"] & u.SET_Z2: kw["z2"] = aa["z2"]
    r = g.analyse(tg, grid, unit=aa["unit"], min_occ=aa["minocc"], rad_ba_nr=aa["radnr"], rad_ba_step=aa["radstep"],
                  **kw)
    scal = dict(volume=r["volume"], reff=r["reff"], sumP=r["sumP"], sumH=r["sumH"], dV=r["dV"], average=r["average"],
                ivol=r["ivol"], iradNR=r["iradNR"], radius=r["geometry"].radius, unit_value=r["unit_value"])
    return r, scal


@pytest.mark.parametrize("name,run", ANALYSES)
def test_analysis_matches_reference_output(name, run):
    """grid file -> gcb_grid_read -> gcb_analyse == what the reference's a_ri3Dc printed, bit for bit"""
    aa = gc.parse_aargs(gc.CASES[name]["analyses"][run])
    tg, hdr, grid = g.grid_read(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    r, scal = gpu_analysis(tg, grid, aa)
    og = u.Geom.from_buffer_copy(bytes(r["tg"]))
    check_analysis_against_golden(og, r["grid"], aa, {k: r[k] for k in ARRAYS}, scal,
                                  os.path.join(u.GOLDEN_DIR, name, run))
    assert r["kernel_launches"] >= 6


@pytest.mark.parametrize("name,run", [(n, r) for n, r in ANALYSES if "-plt" in gc.CASES[n]["analyses"][r]])
def test_plt_of_cropped_grid_byte_exact(name, run, tmp_path):
    aa = gc.parse_aargs(gc.CASES[name]["analyses"][run])
    tg, hdr, grid = g.grid_read(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    r, scal = gpu_analysis(tg, grid, aa)
    path = tmp_path / "out.plt"
    g._cabi.check(g._cabi.lib.gcb_plt_write(str(path).encode(), C.byref(r["tg"]), u.fptr(r["grid"]), r["unit_value"]),
                  "gcb_plt_write")
    assert _read(path) == _read(os.path.join(u.GOLDEN_DIR, name, run, "plt.dat"))


def oracle_analysis(og, grid, unit, min_occ):
    o = u.oracle()
    an = u.Analysis()
    an.unit = u.UNITS[unit]
    an.min_occ = min_occ
    an.rad_ba_nr, an.rad_ba_step = 0, 1
    grid = np.ascontiguousarray(grid, np.float32)
    assert o.gco_analyse(C.byref(og), u.fptr(grid), C.byref(an)) == 0
    arrs = u.analysis_arrays(an, og)
    scal = dict(volume=an.volume, reff=an.reff, sumP=an.sumP, sumH=an.sumH, dV=an.dV, average=an.average,
                ivol=an.ivol, iradNR=an.iradNR)
    o.gco_analysis_free(C.byref(an))
    return arrs, scal


@pytest.mark.parametrize("cfg", [((8.0, 8.0, 13.4), 0.1, (80, 80, 134), "SPC", 0.0),
                                 ((8.0, 8.0, 13.4), 0.1, (80, 80, 134), "molar", 0.7),
                                 ((3.0, 3.0, 3.0), 0.05, (60, 60, 60), "Voxel", 0.001),
                                 ((5.03, 4.41, 6.2), (0.1, 0.09, 0.2), None, "unity", 0.0)])
def test_analysis_bit_exact_vs_oracle_on_binned_density(cfg):
    """BASELINE config-2 / config-1 sized grids: bin a synthetic pore trajectory on the GPU, analyse the
    density on the GPU, compare every output with the oracle's sequential f32 sums bit for bit"""
    L, delta, shape, unit, minocc = cfg
    geom = g.setup_geometry(L, delta)
    if shape:
        assert geom.shape == shape
    natoms, nframes = 6000, 40
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    gen = g.make_gen(99, L, g._cabi.GEN_PORE, slab=(0.35 * L[2], 0.65 * L[2]), pore_r=0.8)
    dev = g.DeviceFrames(nframes, natoms).generate(gen)
    ctr = g.GridCounter(geom, gindex, natoms)
    ctr.add_device_frames(dev, weights=np.full(nframes, 2.0, np.float32))
    tg, dens, _ = ctr.density(2.0 * nframes)
    dens_dev = ctr.density_device(2.0 * nframes)
    ctr.sync()
    r = g.analyse(tg, dens, unit=unit, min_occ=minocc)
    # the same with the density left in HBM: K1 -> K2 -> K3 without the grid crossing PCIe
    rd = g.analyse(tg, None, unit=unit, min_occ=minocc, grid_dev=dens_dev)
    assert rd["grid"] is None
    for k in ARRAYS:
        u.assert_bits_equal(rd[k], r[k], "device-resident " + k)
    for k in ("volume", "reff", "average", "sumP", "sumH"):
        assert np.float64(rd[k]).view(np.uint64) == np.float64(r[k]).view(np.uint64), k
    ctr.close(); dev.free()
    og = u.Geom.from_buffer_copy(bytes(tg))
    u.oracle().gco_update_b(C.byref(og))
    arrs, scal = oracle_analysis(og, dens, unit, minocc)
    for k in ARRAYS:
        u.assert_bits_equal(r[k], arrs[k], k)
    for k in ("volume", "reff", "average"):
        u.assert_bits_equal(np.float32(r[k]), np.float32(scal[k]), k)
    for k in ("sumP", "sumH", "dV"):
        u.assert_bits_equal(np.float64(r[k]), np.float64(scal[k]), k)
    assert r["ivol"] == scal["ivol"] and r["iradNR"] == scal["iradNR"]


def test_analysis_crop_errors():
    tg, hdr, grid = g.grid_read(os.path.join(u.GOLDEN_DIR, "pore", "grid.dat"))
    with pytest.raises(g._cabi.GcbError):
        g.analyse(tg, grid, z1=5.0, z2=2.0)            # z2 <= z1 (a_ri3Dc.c:163-165)


def test_gridcalc_matches_reference_output():
    grids, tws, tg = [], [], None
    for n in gc.GRIDCALC["inputs"]:
        tg, hdr, grid = g.grid_read(os.path.join(u.GOLDEN_DIR, n, "grid.dat"))
        grids.append(grid)
        tws.append(tg.tweight)
    avg, tw = g.gridcalc(tg, grids, tws)
    want_g, want_hdr, want_grid = decode_golden(os.path.join(u.GOLDEN_DIR, "gridcalc", "avg.dat"))
    assert want_g.tweight == tw
    u.assert_bits_equal(avg, want_grid, "a_gridcalc average")


def _synthetic(kind, shape, rng):
    n = int(np.prod(shape))
    if kind == "ones":                      # f32 `average` saturates at 2^24
        return np.ones(shape, np.float32)
    if kind == "halves":                    # every addition is a tie once the sum is large enough
        return (rng.integers(0, 7, n) * 0.5).astype(np.float32).reshape(shape)
    if kind == "wide":
        return np.exp(rng.uniform(-40, 12, n)).astype(np.float32).reshape(shape)
    if kind == "sparse":
        return np.where(rng.random(n) < 0.9, 0, rng.random(n) * 55.5).astype(np.float32).reshape(shape)
    if kind == "negative":                  # not a density: the serial chain takes over
        x = rng.random(n).astype(np.float32).reshape(shape)
        x[shape[0] // 2, shape[1] // 2, shape[2] // 3] = -2.5
        return x
    raise ValueError(kind)


@pytest.mark.parametrize("kind,L,delta", [("ones", (13.0, 13.0, 13.0), 0.05),       # 260^3 = 17.6 M cells > 2^24
                                          ("halves", (6.0, 6.0, 9.0), 0.05),
                                          ("wide", (6.0, 5.0, 4.0), 0.05),
                                          ("sparse", (8.0, 8.0, 13.4), 0.1),
                                          ("negative", (3.0, 3.0, 3.0), 0.05)])
def test_whole_grid_chains_bit_exact_on_adversarial_grids(kind, L, delta):
    """sumP (f64) and average (f32) are evaluated by a parallel scan of the rounding recurrence
    (csrc/seqsum.cuh); they must equal the oracle's one-by-one sums whatever the cell values do"""
    geom = g.setup_geometry(L, delta)
    rng = np.random.default_rng(7)
    grid = _synthetic(kind, geom.shape, rng)
    tg = geom.tg
    r = g.analyse(tg, grid, unit="unity", min_occ=0.25)
    og = u.Geom.from_buffer_copy(bytes(tg))
    u.oracle().gco_update_b(C.byref(og))
    arrs, scal = oracle_analysis(og, grid, "unity", 0.25)
    u.assert_bits_equal(np.float32(r["average"]), np.float32(scal["average"]), "average")
    u.assert_bits_equal(np.float64(r["sumP"]), np.float64(scal["sumP"]), "sumP")
    u.assert_bits_equal(np.float64(r["sumH"]), np.float64(scal["sumH"]), "sumH")
    for k in ARRAYS:
        u.assert_bits_equal(r[k], arrs[k], k)


@pytest.mark.parametrize("shape_L,delta", [((3.0, 3.0, 3.0), 0.05), ((1.0, 1.0, 1.0), 0.5), ((0.9, 0.9, 5.0), (0.3, 0.3, 0.01))])
def test_analysis_of_an_empty_grid_and_tiny_grids(shape_L, delta):
    """all-zero density (0/0 = NaN in lzdf/profile, a_ri3Dc.c:672-695, must come out as NaN in the same places),
    a 2x2x2 grid, a 3x3x500 needle: same bits as the oracle"""
    geom = g.setup_geometry(shape_L, delta)
    for fill in (0.0, 1.5):
        grid = np.full(geom.shape, fill, np.float32)
        r = g.analyse(geom.tg, grid, unit="unity", min_occ=0.0)
        og = u.Geom.from_buffer_copy(bytes(geom.tg))
        u.oracle().gco_update_b(C.byref(og))
        arrs, scal = oracle_analysis(og, grid, "unity", 0.0)
        for k in ARRAYS:
            u.assert_bits_equal(r[k], arrs[k], "%s (fill %g)" % (k, fill))
        for k in ("volume", "reff", "average"):
            u.assert_bits_equal(np.float32(r[k]), np.float32(scal[k]), k)
        for k in ("sumP", "sumH"):
            u.assert_bits_equal(np.float64(r[k]), np.float64(scal[k]), k)


@pytest.mark.parametrize("name,run", [(n, r) for n, r in ANALYSES
                                      if gc.parse_aargs(gc.CASES[n]["analyses"][r])["setmask"]][:3])
def test_device_resident_grid_with_crop_equals_host_path(name, run):
    """gcb_analyse_device on a grid that already sits in HBM, with -R/-z1/-z2 cropping (readjust_tgrid on the
    device): same cropped grid and same profiles as the host-grid call"""
    aa = gc.parse_aargs(gc.CASES[name]["analyses"][run])
    tg, hdr, grid = g.grid_read(os.path.join(u.GOLDEN_DIR, name, "grid.dat"))
    r, _ = gpu_analysis(tg, grid, aa)
    lib = g._cabi.lib
    p = C.c_void_p()
    g._cabi.check(lib.gcb_device_malloc(C.byref(p), grid.nbytes, 0), "gcb_device_malloc")
    try:
        g._cabi.check(lib.gcb_memcpy_h2d(p, grid.ctypes.data, grid.nbytes, 0), "gcb_memcpy_h2d")
        kw = {}
        if aa["setmask"] & u.SET_R: kw["radius"] = aa["R"]
        if aa["setmask"] & u.SET_Z1: kw["z1"] = aa["z1"]
        if aa["setmask"] & u.SET_Z2: kw["z2"] = aa["z2"]
        rd = g.analyse(tg, None, unit=aa["unit"], min_occ=aa["minocc"], rad_ba_nr=aa["radnr"], rad_ba_step=aa["radstep"],
                       grid_dev=p, **kw)
    finally:
        lib.gcb_device_free(p, 0)
    assert rd["grid"] is not None and rd["grid"].shape == r["grid"].shape
    u.assert_bits_equal(rd["grid"], r["grid"], "cropped grid")
    for k in ARRAYS:
        u.assert_bits_equal(rd[k], r[k], k)
    assert bytes(rd["tg"]) == bytes(r["tg"])


def test_c3_sized_grid_bit_exact_against_the_oracle():
    """BASELINE config 3's grid (400^3 = 64 M cells, 256 MB): every output of the analysis equals the oracle's
    sequential chains bit for bit (the f32 `average` over 64 M cells saturates in the reference; so must ours), from
    a host grid and from the same grid already in device memory."""
    L, delta = (20.0, 20.0, 20.0), 0.05
    geom = g.setup_geometry(L, delta)
    assert geom.shape == (400, 400, 400)
    rng = np.random.default_rng(400)
    dens = (rng.poisson(2.0, geom.shape) * np.float32(0.41)).astype(np.float32)
    tg = geom.tg
    og = u.Geom.from_buffer_copy(bytes(tg))
    u.oracle().gco_update_b(C.byref(og))
    an = u.Analysis()
    an.unit = u.UNITS["SPC"]; an.rad_ba_step = 1
    u.oracle().gco_analyse(C.byref(og), u.fptr(dens), C.byref(an))
    want = u.analysis_arrays(an, og)
    scal = (np.float32(an.average), an.sumP, an.sumH, np.float32(an.volume), np.float32(an.reff), an.ivol)
    u.oracle().gco_analysis_free(C.byref(an))
    dev = g.DeviceFrames(1, dens.size // 3 + 1)
    g._cabi.check(g._cabi.lib.gcb_memcpy_h2d(dev.ptr, dens.ctypes.data, dens.nbytes, 0), "h2d")
    for r in (g.analyse(tg, dens, unit="SPC"), g.analyse(tg, None, unit="SPC", grid_dev=dev.ptr)):
        for k in ("rzp", "hrzp", "zdf", "lzdf", "profile", "rdf", "hrdf", "xyp", "hxyp", "xzp", "yzp", "rweight"):
            p, q = np.asarray(want[k], np.float32), np.asarray(r[k], np.float32)
            assert np.array_equal(np.isnan(p), np.isnan(q)), k
            assert ((p.view(np.uint32) == q.view(np.uint32)) | np.isnan(p)).all(), k
        got = (np.float32(r["average"]), r["sumP"], r["sumH"], np.float32(r["volume"]), np.float32(r["reff"]), r["ivol"])
        assert got == scal, (got, scal)
    dev.free()


def test_release_cache_between_analyses():
    """gcb_analyse keeps its device arena, streams, tables and pinned staging per device; giving them back and
    analysing again (another geometry in between) must change nothing"""
    tg, hdr, dens = g.grid_read(os.path.join(u.GOLDEN_DIR, "pore", "grid.dat"))
    r1 = g.analyse(tg, dens, unit="SPC")
    g._cabi.lib.gcb_analyse_release_cache(0)
    geom = g.setup_geometry((2.0, 2.0, 3.0), 0.1)
    rng = np.random.default_rng(1)
    other = (rng.poisson(2.0, geom.shape) * 0.5).astype(np.float32)
    g.analyse(geom.tg, other, unit="unity")
    r2 = g.analyse(tg, dens, unit="SPC")
    for k in ("rzp", "zdf", "lzdf", "rdf", "xyp", "xzp", "yzp", "profile"):
        assert np.array_equal(np.asarray(r1[k]).view(np.uint32), np.asarray(r2[k]).view(np.uint32)), k
    assert r1["sumP"] == r2["sumP"] and np.float32(r1["average"]) == np.float32(r2["average"])
