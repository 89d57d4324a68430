"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the CPU
oracle on the same seeded inputs, and against the reference's golden files."""
import ctypes as C
import os

import numpy as np
import pytest

import gcutil as u
import golden_cases as gc
import gridcount_b200 as g
from gridcount_b200 import _cabi as K

pytestmark = pytest.mark.gpu

KERNELS = [K.KERNEL_GATHER, K.KERNEL_STREAM, K.KERNEL_STREAM_INDEXED, K.KERNEL_PACKED, K.KERNEL_PRIVATE]


def to_ogeom(geom):
    return u.Geom.from_buffer_copy(bytes(geom.tg))


def oracle_counts(geom, x, gindex):
    o = u.oracle()
    og = to_ogeom(geom)
    x = np.ascontiguousarray(x, np.float32)
    gi = np.ascontiguousarray(gindex, np.int32)
    counts = np.zeros(geom.shape, np.uint32)
    hits = np.zeros(x.shape[0], np.uint64)
    ovf = C.c_uint64(0)
    o.gco_count_frames(C.byref(og), counts.ctypes.data_as(u.c_u32_p), u.fptr(x), x.shape[0], x.shape[1], u.iptr(gi),
                       len(gi), hits.ctypes.data_as(u.c_u64_p), C.byref(ovf))
    return counts, hits, ovf.value


def oracle_occupancy(geom, x, gindex, w):
    o = u.oracle()
    og = to_ogeom(geom)
    x = np.ascontiguousarray(x, np.float32)
    gi = np.ascontiguousarray(gindex, np.int32)
    w = np.ascontiguousarray(w, np.float32)
    grid = np.zeros(geom.shape, np.float32)
    sp = C.c_double(0)
    o.gco_bin_frames(C.byref(og), u.fptr(grid), u.fptr(x), x.shape[0], x.shape[1], u.iptr(gi), len(gi), u.fptr(w),
                     C.byref(sp), None, None)
    return grid, sp.value


# ----------------------------------------------------------------------------- generator
@pytest.mark.parametrize("mode,kw", [(0, {}), (1, {}), (2, dict(slab=(2.2, 4.5), pore_r=0.8)),
                                     (3, dict(nsites=7, sigma=0.1)), (0, dict(box_period=10, box_amp=0.002))])
def test_generator_bits_match_oracle(mode, kw):
    L = (4.0, 4.0, 6.7)
    natoms, nframes = 777, 23
    dev = g.DeviceFrames(nframes, natoms).generate(g.make_gen(42, L, mode, **kw), frame0=5)
    got = dev.download()
    want = u.gen_frames(u.make_gen(42, L, mode, **kw), 5, nframes, natoms)
    u.assert_bits_equal(got, want, "generated frames")
    dev.free()


# ----------------------------------------------------------------------------- golden cases end to end
@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", list(gc.CASES))
def test_golden_density_file_byte_exact(name, kernel):
    """frames -> GPU binning -> density -> XDR bytes == the file the reference wrote"""
    c = gc.CASES[name]
    ga = gc.parse_gargs(c["gargs"])
    t, boxes, x, gindex = gc.case_inputs(name)
    kw = {}
    if ga["setmask"] & u.SET_CPOINT: kw["cpoint"] = ga["cpoint"]
    if ga["setmask"] & u.SET_R: kw["radius"] = ga["R"]
    if ga["setmask"] & u.SET_Z1: kw["z1"] = ga["z1"]
    if ga["setmask"] & u.SET_Z2: kw["z2"] = ga["z2"]
    geom = g.setup_geometry(boxes[0], c["delta"], **kw)
    w, T = g.frame_weights(t, ga["dtweight"])
    ctr = g.GridCounter(geom, gindex, c["natoms"], kernel=kernel, frames_per_batch=7)
    ctr.add_frames(x, w)
    st = ctr.stats()
    assert st.frames == c["nframes"] and st.edge_overflow == 0
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    occ, osumP = oracle_occupancy(geom, x, gindex, w)
    u.assert_bits_equal(ctr.occupancy(), occ, "f32 occupancy")
    assert st.sumP == osumP
    tg, gz, gx = ctr.density(T, zfast=True, xfast=True)
    grp = "OW" if c["stride"] != 1 else "System"
    hdr = g.header_string(geom, T, st.sumP, grp, "traj.raw", ga["subtitle"], ga["dtweight"])
    want = open(os.path.join(u.GOLDEN_DIR, name, "grid.dat"), "rb").read()
    assert g.xdr_encode(tg, hdr, grid_xfast=gx) == want
    assert g.xdr_encode(tg, hdr, grid_zfast=gz) == want
    # the same file image with the reordering and the XDR byte order done on the device
    tg2, image = ctr.xdr(T, hdr)
    assert image == want and bytes(tg2) == bytes(tg)
    # gOpenMol plt of the density: device-encoded cells == the host writer's (plt.c:42-77)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        g._cabi.check(g._cabi.lib.gcb_plt_write(os.path.join(d, "a.plt").encode(), C.byref(tg), u.fptr(gz), 0.75), "gcb_plt_write")
        assert ctr.plt(T, unit=0.75) == open(os.path.join(d, "a.plt"), "rb").read()
        ctr.grid_write(T, os.path.join(d, "g.dat"), hdr)
        assert open(os.path.join(d, "g.dat"), "rb").read() == want
    ctr.close()


# ----------------------------------------------------------------------------- randomised shapes
SHAPES = [
    # natoms, gindex builder, nframes, box, delta
    (2700, lambda n, r: np.arange(0, n, 3), 40, (3.0, 3.0, 3.0), 0.05),
    (2049, lambda n, r: np.arange(n), 9, (2.98, 3.31, 4.07), (0.2, 0.2, 0.25)),          # natoms % 4 != 0
    (61, lambda n, r: np.arange(0, n, 2), 300, (2.0, 2.0, 2.0), 0.1),                      # many frames per tile
    (5000, lambda n, r: r.permutation(n)[:1234], 13, (4.0, 5.0, 6.0), (0.11, 0.13, 0.17)),  # unsorted group
    (4099, lambda n, r: r.integers(0, n, 3000), 11, (4.0, 4.0, 4.0), 0.08),                # duplicates
    (10000, lambda n, r: np.array([17, 4242, 9999]), 25, (5.0, 5.0, 5.0), 0.1),            # sparse (ions)
    (3, lambda n, r: np.array([1]), 1, (1.0, 1.0, 1.0), 0.5),
    (8192, lambda n, r: np.arange(n), 4, (6.0, 6.0, 6.0), 0.05),                            # exact tiles, no tail
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("shape", range(len(SHAPES)))
def test_counts_bit_exact_random_shapes(shape, kernel):
    natoms, mk, nframes, L, delta = SHAPES[shape]
    rng = np.random.default_rng(100 + shape)
    gindex = mk(natoms, rng).astype(np.int32)
    # positions reach a little outside the box so that drops happen on every side
    x = u.gen_frames(u.make_gen(500 + shape, tuple(1.1 * v for v in L)), 0, nframes, natoms)
    x -= np.float32(0.05) * np.asarray(L, np.float32)
    geom = g.setup_geometry(L, delta)
    ocounts, ohits, oovf = oracle_counts(geom, x, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    ctr.add_frames(x)                                    # host path, pageable memory
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    st = ctr.stats()
    assert st.hits == ohits.sum() == ocounts.sum(dtype=np.uint64)
    assert st.edge_overflow == oovf
    # device-resident path gives the same grid (and doubles it)
    dev = g.DeviceFrames(nframes, natoms).upload(x)
    ctr.add_device_frames(dev)
    np.testing.assert_array_equal(ctr.counts(), 2 * ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), np.concatenate([ohits, ohits]))
    dev.free()
    ctr.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_special_values_and_edge_overflow(kernel):
    # a geometry where b - 1ulp bins to mx: the hit is dropped and counted
    L, delta = (9.4928503, 9.4928503, 3.0), (0.100521415, 0.100521415, 0.1)
    geom = g.setup_geometry(L, delta)
    assert geom.edge_vulnerable
    natoms, nframes = 64, 3
    x = u.gen_frames(u.make_gen(9, L), 0, nframes, natoms)
    b = np.array(geom.tg.b[:], np.float32)
    a = np.array(geom.tg.a[:], np.float32)
    x[0, 0] = (np.nextafter(b[0], np.float32(-1)), 1.0, 1.0)        # edge overflow in x
    x[0, 1] = (np.nan, 1.0, 1.0)                                    # NaN passes both bounds tests
    x[0, 2] = (np.inf, 1.0, 1.0)
    x[0, 3] = (-np.inf, 1.0, 1.0)
    x[0, 4] = (a[0], a[1], a[2])                                    # exactly on the lower corner: inside
    x[0, 5] = (b[0], 1.0, 1.0)                                      # exactly on the upper edge: outside
    x[0, 6] = (np.nan, -5.0, 1.0)                                   # outside wins over NaN
    x[1, 7] = (1.0, 1.0, np.float32(-0.0))
    gindex = np.arange(natoms, dtype=np.int32)
    ocounts, ohits, oovf = oracle_counts(geom, x, gindex)
    assert oovf == 2
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    ctr.add_frames(x)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    assert ctr.stats().edge_overflow == oovf
    ctr.close()


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("geo", [((3.0, 3.0, 3.0), 0.05), ((8.0, 8.0, 13.4), 0.1), ((2.98, 3.31, 4.07), (0.2, 0.21, 0.25)),
                                 ((20.137, 20.137, 20.137), 0.05)])
def test_cell_boundaries_bit_exact(geo, kernel):
    """atoms placed on (and one ulp either side of) every cell boundary: the estimate-and-verify bin
    conversion must agree with the reference's IEEE division exactly where it matters"""
    L, delta = geo
    geom = g.setup_geometry(L, delta)
    a = np.array(geom.tg.a[:], np.float32)
    dl = np.array(geom.tg.delta[:], np.float32)
    mx = np.array(geom.tg.mx[:])
    rng = np.random.default_rng(11)
    pts = []
    for d in range(3):
        k = np.arange(0, mx[d] + 1, dtype=np.float32)
        edge = (a[d] + k * dl[d]).astype(np.float32)                 # f32 product and sum, like a boundary estimate
        for cand in (edge, np.nextafter(edge, np.float32(-1e9)), np.nextafter(edge, np.float32(1e9)),
                     np.nextafter(np.nextafter(edge, np.float32(1e9)), np.float32(1e9))):
            p = (a + rng.random((len(cand), 3), dtype=np.float32) * (mx * dl)).astype(np.float32)
            p[:, d] = cand
            pts.append(p)
    x = np.concatenate(pts).astype(np.float32)
    natoms = len(x)
    x = x[None, :, :].repeat(2, axis=0)
    x[1] = x[1, ::-1]
    gindex = np.arange(natoms, dtype=np.int32)
    ocounts, ohits, oovf = oracle_counts(geom, x, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    dev = g.DeviceFrames(2, natoms).upload(x)
    ctr.add_device_frames(dev)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    assert ctr.stats().edge_overflow == oovf
    dev.free(); ctr.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_pbc_rect_wrap(kernel):
    o = u.oracle()
    L = (3.0, 3.5, 4.0)
    natoms, nframes = 3000, 6
    x = u.gen_frames(u.make_gen(77, tuple(1.6 * v for v in L)), 0, nframes, natoms)
    x -= np.float32(0.3) * np.asarray(L, np.float32)                # spans about [-0.3 L, 1.3 L)
    boxes = np.tile(u.box9(L), (nframes, 1)).astype(np.float32)
    boxes[:, 0] *= np.linspace(1.0, 1.01, nframes, dtype=np.float32)     # breathing box: wrap uses each frame's box
    geom = g.setup_geometry(L, 0.1)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    xw = x.copy()
    o.gco_wrap_frames(u.fptr(xw), nframes, natoms, u.fptr(boxes))
    ocounts, ohits, _ = oracle_counts(geom, xw, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel, pbc=K.PBC_RECT)
    ctr.add_frames(x, box=boxes)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    # without the opt-in the reference behaviour (drop) is kept
    ctr2 = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    ctr2.add_frames(x)
    np.testing.assert_array_equal(ctr2.counts(), oracle_counts(geom, x, gindex)[0])
    ctr.close(); ctr2.close()


# ----------------------------------------------------------------------------- f32 accumulation semantics
def _weights(kind, n, rng):
    if kind == "const01":
        return np.full(n, 0.1, np.float32)
    if kind in ("const1", "const075", "const2p-100"):     # short significands: the exact-product shortcut of seq_add_dev
        return np.full(n, {"const1": 1.0, "const075": 0.75, "const2p-100": 2.0 ** -100}[kind], np.float32)
    if kind == "jitter":              # dt = t_f - t_{f-1} of f32 times f*0.1: varies in the last bits
        t = (np.arange(n + 1, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
        return (t[1:] - t[:-1]).astype(np.float32)
    if kind == "blocks":              # long runs of equal weight -> integer path + fold
        return np.repeat(np.array([0.1, 0.3, 0.1, 2.0], np.float32), (n + 3) // 4)[:n]
    if kind == "mixed":
        w = np.repeat(np.array([1.0, 0.7], np.float32), n // 2)
        w[n // 3] = 0.123
        return np.resize(w, n).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("kind", ["const01", "const1", "const075", "const2p-100", "jitter", "blocks", "mixed"])
def test_occupancy_reproduces_reference_f32_accumulation(kind, kernel):
    """`grid[i][j][k] += weight` in f32, frame after frame (g_ri3Dc.c:149): bit-exact for any weights"""
    rng = np.random.default_rng(5)
    L = (1.5, 1.5, 1.5)
    natoms, nframes = 6000, 96                   # ~400 hits per cell: deep into f32 rounding for w = 0.1
    geom = g.setup_geometry(L, 0.25)
    x = u.gen_frames(u.make_gen(31, L), 0, nframes, natoms)
    gindex = np.arange(natoms, dtype=np.int32)
    w = _weights(kind, nframes, rng)
    occ, osumP = oracle_occupancy(geom, x, gindex, w)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel, frames_per_batch=10)
    ctr.add_frames(x[:50], w[:50])
    ctr.add_frames(x[50:], w[50:])
    u.assert_bits_equal(ctr.occupancy(), occ, "occupancy (%s)" % kind)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    st = ctr.stats()
    assert st.sumP == osumP
    assert bool(st.uniform_weight) == kind.startswith("const")
    # density = occupancy / ((double)(float)dV * T)
    T = float(np.sum(w.astype(np.float64)))
    og = to_ogeom(geom)
    dens = occ.copy()
    u.oracle().gco_normalise(C.byref(og), u.fptr(dens), T)
    tg, gz, gx = ctr.density(T, zfast=True, xfast=True)
    u.assert_bits_equal(gz, dens, "density")
    u.assert_bits_equal(gx, np.ascontiguousarray(dens.transpose(2, 1, 0)), "density, file order")
    assert tg.tweight == np.float32(T)
    ctr.close()


def test_reset_and_slot_interface():
    L = (2.0, 2.0, 2.0)
    natoms, nframes = 1000, 12
    geom = g.setup_geometry(L, 0.1)
    x = u.gen_frames(u.make_gen(3, L), 0, nframes, natoms)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, frames_per_batch=5, nslots=2)
    ctr.add_frames(x)
    ctr.reset()
    assert ctr.counts().sum() == 0 and ctr.stats().frames == 0
    # zero-copy producer interface: fill the pinned slots directly
    done = 0
    slot = 0
    while done < nframes:
        p = K.c_float_p()
        cap = C.c_long(0)
        K.check(K.lib.gcb_counter_slot(ctr.h, slot, C.byref(p), C.byref(cap)), "slot")
        n = min(cap.value, nframes - done)
        C.memmove(p, x[done:done + n].ctypes.data, n * natoms * 12)
        K.check(K.lib.gcb_counter_submit_slot(ctr.h, slot, n, None, None), "submit")
        done += n
        slot = (slot + 1) % 2
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    # pinned caller memory is DMA'd directly
    pin = g.PinnedFrames(nframes, natoms)
    pin.array[:] = x
    ctr.add_frames(pin.array)
    np.testing.assert_array_equal(ctr.counts(), 2 * ocounts)
    pin.free()
    ctr.close()


# ----------------------------------------------------------------------------- full-size properties
def test_c2_shape_512_frames_all_kernels_agree():
    """BASELINE config 2 shape (60k atoms, 20k OW, 80x80x134) on 512 generated frames: size-independent
    invariants + a replayable prefix against the oracle."""
    L = (8.0, 8.0, 13.4)
    natoms, nframes = 60000, 512
    gen = g.make_gen(20261021, L, K.GEN_PORE, slab=(4.7, 8.7), pore_r=0.8)
    dev = g.DeviceFrames(nframes, natoms).generate(gen)
    geom = g.setup_geometry(L, 0.1)
    assert geom.shape == (80, 80, 134) and not geom.edge_vulnerable
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    res = {}
    for kernel in KERNELS:
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        ctr.add_device_frames(dev)
        res[kernel] = (ctr.counts(), ctr.hits_per_frame())
        st = ctr.stats()
        assert st.hits == res[kernel][0].sum(dtype=np.uint64) == res[kernel][1].sum()
        assert st.hits == nframes * len(gindex)            # every generated atom is inside this grid
        # linearity: binning the two halves separately adds up
        a = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        a.add_device_frames(dev, 0, 200)
        b = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        b.add_device_frames(dev, 200, nframes - 200)
        np.testing.assert_array_equal(a.counts() + b.counts(), res[kernel][0])
        a.close(); b.close(); ctr.close()
    np.testing.assert_array_equal(res[K.KERNEL_GATHER][0], res[K.KERNEL_STREAM][0])
    # replayable prefix on the CPU oracle
    x = u.gen_frames(u.make_gen(20261021, L, 2, slab=(4.7, 8.7), pore_r=0.8), 0, 24, natoms)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    ctr = g.GridCounter(geom, gindex, natoms)
    ctr.add_device_frames(dev, 0, 24)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    # the membrane slab is empty outside the pore
    cz = ocounts.sum(axis=(0, 1))
    assert cz[50:85].max() < cz[:40].min()
    ctr.close(); dev.free()


@pytest.mark.parametrize("cfg", [((20.0, 20.0, 20.0), 0.05, (400, 400, 400), 300000, 6),
                                 ((10.0, 10.0, 10.0), 0.02, (500, 500, 500), 99999, 10)])
def test_large_grids(cfg):
    """BASELINE configs 3 and 4 geometry (grid larger than L2): counts against the oracle on a prefix"""
    L, delta, shape, natoms, nframes = cfg
    geom = g.setup_geometry(L, delta)
    assert geom.shape == shape
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    dev = g.DeviceFrames(nframes, natoms).generate(g.make_gen(7, L))
    x = dev.download()
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    for kernel in KERNELS:
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        ctr.add_device_frames(dev)
        np.testing.assert_array_equal(ctr.counts(), ocounts)
        np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
        ctr.close()
    dev.free()


def test_full_size_c2_properties():
    """BASELINE config 2 at FULL size (60 000 atoms, 20 000 OW, 10 000 frames, 80x80x134): too large for the oracle,
    so size-independent properties: every generated atom lies in the grid (total = frames*gnx, per-frame hits = gnx);
    direct, packed and gather kernels agree cell for cell; counting is additive over frame ranges; and the first
    64 frames match the oracle."""
    L, natoms, F = (8.0, 8.0, 13.4), 60000, 10000
    geom = g.setup_geometry(L, 0.1)
    assert geom.shape == (80, 80, 134)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    gen = g.make_gen(20261021, L, K.GEN_PORE, slab=(4.7, 8.7), pore_r=0.8)
    dev = g.DeviceFrames(F, natoms).generate(gen)
    ref = None
    for kernel in (K.KERNEL_STREAM, K.KERNEL_PACKED, K.KERNEL_GATHER):
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        ctr.add_device_frames(dev)
        c = ctr.counts()
        st = ctr.stats()
        assert st.hits == F * len(gindex) == int(c.sum(dtype=np.uint64)) and st.edge_overflow == 0
        assert np.all(ctr.hits_per_frame() == len(gindex))
        if ref is None:
            ref = c
        else:
            assert np.array_equal(c, ref), "kernel %d differs from the streaming kernel" % kernel
        ctr.close()
    # additivity over frame ranges (odd split so that tiles straddle the cut)
    cut = 3333
    a = g.GridCounter(geom, gindex, natoms)
    a.add_device_frames(dev, first=0, count=cut)
    b = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_GATHER)     # the second range starts unaligned for TMA
    b.add_device_frames(dev, first=cut, count=F - cut)
    assert np.array_equal(a.counts() + b.counts(), ref)
    a.close(); b.close()
    # oracle on a prefix
    x = dev.download(0, 64)
    oc, oh, _ = oracle_counts(geom, x, gindex)
    p = g.GridCounter(geom, gindex, natoms)
    p.add_device_frames(dev, first=0, count=64)
    assert np.array_equal(p.counts(), oc)
    p.close(); dev.free()


@pytest.mark.parametrize("name", ["C3", "C4"])
def test_named_large_shapes_with_the_kernel_auto_picks(name):
    """BASELINE configs 3 and 4 at their named frame size (1 M atoms on 400^3; 100 k atoms on 500^3) through
    GCB_KERNEL_AUTO, i.e. the packed-counter path with its persisting-L2 window: counts and per-frame hits against the
    oracle for a uniform box (no round may need the redo) and for atoms piled onto eight small clusters (every packed
    round overflows its lanes: the digit-sum check must catch it, the uint32 kernel redoes it, the host pauses the
    packed passes) -- and a uniform batch afterwards still comes out exact."""
    L, delta, shape, natoms, nframes = {"C3": ((20.0, 20.0, 20.0), 0.05, (400, 400, 400), 1000000, 16),
                                        "C4": ((10.0, 10.0, 10.0), 0.02, (500, 500, 500), 100000, 96)}[name]
    geom = g.setup_geometry(L, delta)
    assert geom.shape == shape
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    ctr = g.GridCounter(geom, gindex, natoms)
    dev = g.DeviceFrames(nframes, natoms).generate(g.make_gen(31, L))
    ctr.add_device_frames(dev)
    want, whits, _ = oracle_counts(geom, dev.download(), gindex)
    np.testing.assert_array_equal(ctr.counts(), want)
    np.testing.assert_array_equal(ctr.hits_per_frame(), whits)
    st = ctr.stats()
    assert st.packed_rounds >= 1 and st.packed_redone == 0
    dev.generate(g.make_gen(32, L, K.GEN_CLUSTER, nsites=8, sigma=0.01))
    total = want
    NCALLS = 40                                        # enough calls for the back-off to reach the host-side pause, on the
    for rep in range(NCALLS):                          # 500^3 grid by way of the byte-lane fallback (which overflows as well)
        ctr.add_device_frames(dev)
        ctr.sync()                                     # (the host learns of the back-off from a mirror copied behind each call)
    w2, h2, _ = oracle_counts(geom, dev.download(), gindex)
    total = total + NCALLS * w2
    np.testing.assert_array_equal(ctr.counts(), total)
    st2 = ctr.stats()
    assert st2.packed_redone >= 1 and st2.hits == int(whits.sum()) + NCALLS * int(h2.sum())
    assert st2.packed_rounds < st.packed_rounds + NCALLS, "the host never paused the packed passes"
    ctr.close(); dev.free()


# ----------------------------------------------------------------------------- packed counters under pressure
@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("kind", ["one cell", "64 sites", "warm + background", "uniform"])
def test_packed_counters_overflow_and_backoff_stay_exact(kind, bits, monkeypatch):
    """GCB_KERNEL_PACKED keeps 8- or 4-bit counters packed into words and updates them with non-returning REDs; a
    round is accepted only if the digit sum of the packed counters equals the hits it reported, otherwise it is
    thrown away and redone by the uint32 kernel, and the next round(s) go straight there (back-off).  Counts and
    per-frame hits must be exact whichever way every round went; a spread-out input must never take the redo."""
    monkeypatch.setenv("GCB_PACKED_BITS", str(bits))
    monkeypatch.setenv("GCB_PACKED_UNITS", str(8192 * 8))           # 8 frames per round: six rounds per call
    L, delta = (4.0, 4.0, 4.0), 0.1
    geom = g.setup_geometry(L, delta)
    natoms, nframes = 8192, 48
    rng = np.random.default_rng(11)
    x = (rng.random((nframes, natoms, 3)) * np.asarray(L)).astype(np.float32)
    if kind == "one cell":                       # 65 536 hits on a single lane per round: every packed pass must fail
        x[:] = np.asarray([1.234, 2.345, 3.456], np.float32)
    elif kind == "64 sites":                     # ~1 000 hits per cell and round, all atoms of a frame on 64 cells
        sites = (rng.random((64, 3)) * np.asarray(L)).astype(np.float32)
        x[:] = sites[rng.integers(0, 64, (nframes, natoms))]
    elif kind == "warm + background":            # every 16th atom sits on one of 64 sites: ~8 hits per frame and site
        sites = (rng.random((64, 3)) * np.asarray(L)).astype(np.float32)
        x[:, ::16] = sites[rng.integers(0, 64, (nframes, natoms // 16))]
    gindex = np.arange(natoms, dtype=np.int32)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_PACKED)
    ctr.add_frames(x, np.ones(nframes, np.float32))
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    st = ctr.stats()
    assert st.packed_rounds == 6 and st.hits == int(ohits.sum())
    if kind in ("one cell", "64 sites"):
        assert st.packed_redone == 6, "overflowing lanes were not caught by the digit-sum check"
    if kind == "uniform":                        # 64 000 cells, 65 536 hits per round: mean 1 per cell
        assert st.packed_redone == 0
    # a second batch on top: the packed counters must have been left all-zero, the statistics keep counting
    ctr.add_frames(x[:8], np.ones(8, np.float32))
    o2, _, _ = oracle_counts(geom, x[:8], gindex)
    np.testing.assert_array_equal(ctr.counts(), ocounts + o2)
    assert ctr.stats().packed_rounds == 7
    ctr.close()


def test_packed_backoff_recovers(monkeypatch):
    """Dense frames followed by spread-out ones: the dense rounds fail verification and switch the following rounds
    to the direct kernel; the back-off halves on every verified round, so the packed path comes back."""
    monkeypatch.setenv("GCB_PACKED_UNITS", str(8192 * 4))
    L = (4.0, 4.0, 4.0)
    geom = g.setup_geometry(L, 0.1)
    natoms, nframes = 8192, 128
    rng = np.random.default_rng(12)
    x = (rng.random((nframes, natoms, 3)) * np.asarray(L)).astype(np.float32)
    x[:16] = np.asarray([0.5, 0.5, 0.5], np.float32)
    gindex = np.arange(natoms, dtype=np.int32)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_PACKED)
    ctr.add_frames(x, np.ones(nframes, np.float32))
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    st = ctr.stats()
    assert st.packed_rounds == 32
    assert 4 <= st.packed_redone < 16, st.packed_redone      # the four dense rounds and a few skipped ones, not the rest
    ctr.close()


# ----------------------------------------------------------------------------- empty and degenerate inputs
@pytest.mark.parametrize("kernel", KERNELS)
def test_empty_and_degenerate_inputs(kernel):
    """empty index group, zero frames, every atom outside the grid, a one-cell grid: nothing is counted that the
    reference would not count (g_ri3Dc.c:145 drops outside atoms silently), nothing crashes"""
    L = (2.0, 2.0, 2.0)
    geom = g.setup_geometry(L, 0.1)
    natoms = 500
    rng = np.random.default_rng(3)
    x = (rng.random((6, natoms, 3)) * 2.0).astype(np.float32)
    # (1) empty index group
    ctr = g.GridCounter(geom, np.zeros(0, np.int32), natoms, kernel=kernel)
    ctr.add_frames(x, np.ones(6, np.float32))
    assert ctr.counts().sum() == 0 and ctr.stats().hits == 0 and ctr.stats().frames == 6
    np.testing.assert_array_equal(ctr.hits_per_frame(), np.zeros(6, np.uint64))
    ctr.close()
    # (2) zero frames, then a real batch, then zero frames again
    gindex = np.arange(natoms, dtype=np.int32)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    ctr.add_frames(x[:0], np.ones(0, np.float32))
    assert ctr.stats().frames == 0 and ctr.counts().sum() == 0
    ctr.add_frames(x, np.ones(6, np.float32))
    ctr.add_frames(x[:0], np.ones(0, np.float32))
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    ctr.close()
    # (3) every atom outside [a, b): dropped, density identically zero
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
    far = x + np.float32(5.0)
    far[0, :10] = -far[0, :10]
    ctr.add_frames(far, np.ones(6, np.float32))
    st = ctr.stats()
    assert st.hits == 0 and st.edge_overflow == 0 and st.sumP == 0.0
    tg, dens, _ = ctr.density(6.0)
    assert not dens.any()
    ctr.close()
    # (4) one cell
    one = g.setup_geometry(L, 2.0)
    assert one.shape == (1, 1, 1)
    ctr = g.GridCounter(one, gindex, natoms, kernel=kernel)
    ctr.add_frames(x, np.ones(6, np.float32))
    oc, oh, _ = oracle_counts(one, x, gindex)
    np.testing.assert_array_equal(ctr.counts(), oc)
    np.testing.assert_array_equal(ctr.hits_per_frame(), oh)
    ctr.close()


# ----------------------------------------------------------------------------- saturation guard of the uint32 cells
def test_saturation_guard_with_lowered_thresholds(monkeypatch):
    """The library clamps cell counts at 2^27 whenever the uint32 range could otherwise run out (the reference's f32
    accumulator has stopped moving long before, so densities keep the reference's bits).  With the thresholds lowered
    through the environment the mechanism is visible: a cell that keeps collecting hits is held between the clamp
    value and the ceiling and flagged, every other cell stays exact; a well-spread run is never touched."""
    monkeypatch.setenv("GCB_CLAMP_AT", "1000")
    monkeypatch.setenv("GCB_CLAMP_CEILING", "40000")
    L = (2.0, 2.0, 2.0)
    geom = g.setup_geometry(L, 0.1)
    natoms, nframes = 4096, 40
    rng = np.random.default_rng(5)
    x = (rng.random((nframes, natoms, 3)) * 2.0).astype(np.float32)
    gindex = np.arange(natoms, dtype=np.int32)
    for kernel in (K.KERNEL_STREAM, K.KERNEL_GATHER, K.KERNEL_PACKED):
        # (a) spread out: bounds force clamp passes (4096 possible hits per frame against a ceiling of 40 000), no cell is near 1000
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel, frames_per_batch=4)
        ctr.add_frames(x, np.ones(nframes, np.float32))
        oc, oh, _ = oracle_counts(geom, x, gindex)
        np.testing.assert_array_equal(ctr.counts(), oc)
        assert ctr.stats().saturated == 0 and oc.max() < 1000
        ctr.close()
        # (b) a quarter of the atoms parked in one cell: 1024 hits per frame on it
        y = x.copy()
        y[:, ::4] = np.asarray([0.55, 0.65, 0.75], np.float32)
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel, frames_per_batch=4)
        ctr.add_frames(y, np.ones(nframes, np.float32))
        oc, oh, _ = oracle_counts(geom, y, gindex)
        got = ctr.counts()
        hot = np.unravel_index(np.argmax(oc), oc.shape)
        assert oc[hot] >= 1024 * nframes
        assert 1000 <= got[hot] <= 40000, got[hot]
        mask = np.ones(oc.shape, bool); mask[hot] = False
        np.testing.assert_array_equal(got[mask], oc[mask])
        np.testing.assert_array_equal(ctr.hits_per_frame(), oh)      # per-frame statistics are 64-bit and unaffected
        assert ctr.stats().saturated == 1
        ctr.close()


# ----------------------------------------------------------------------------- private shared-memory grids
@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("kind", ["uniform", "one cell", "64 sites", "warm + background"])
def test_private_grids_stay_exact_through_overflow_flushes_and_reset(kind, bits, monkeypatch):
    """GCB_KERNEL_PRIVATE: every CTA bins with shared-memory atomics into its own 8- or 4-bit copy of the grid; the
    copies rest in global memory between launches and are folded into the uint32 grid lazily.  A CTA whose digit sum
    does not equal the hits it holds (a lane overflowed) bins its tiles again directly and the next launches back
    off.  Counts and per-frame hits must equal the oracle whichever way every launch went, over several launches,
    with forced flushes in between, after a reset, and for frames whose tiles straddle frame boundaries."""
    monkeypatch.setenv("GCB_PRIVATE_BITS", str(bits))
    L, delta = (4.0, 4.0, 4.0), 0.1
    geom = g.setup_geometry(L, delta)                # 40^3 = 64 000 cells
    natoms, nframes = 8190, 40                       # 8190 atoms: tiles of 3072 atoms straddle frames
    rng = np.random.default_rng(13)
    x = (rng.random((nframes, natoms, 3)) * np.asarray(L)).astype(np.float32)
    if kind == "one cell":                           # every CTA gets thousands of hits on one lane: must fail and redo
        x[:] = np.asarray([1.234, 2.345, 3.456], np.float32)
    elif kind == "64 sites":
        sites = (rng.random((64, 3)) * np.asarray(L)).astype(np.float32)
        x[:] = sites[rng.integers(0, 64, (nframes, natoms))]
    elif kind == "warm + background":
        sites = (rng.random((64, 3)) * np.asarray(L)).astype(np.float32)
        x[:, ::16] = sites[rng.integers(0, 64, (nframes, x[:, ::16].shape[1]))]
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    dev = g.DeviceFrames(nframes, natoms)
    dev.upload(x)
    for limit in (None, "3000"):                     # default flush threshold / a flush every launch or two
        monkeypatch.delenv("GCB_PRIVATE_LIMIT", raising=False)
        if limit:
            monkeypatch.setenv("GCB_PRIVATE_LIMIT", limit)
        ctr = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_PRIVATE)
        for first in range(0, nframes, 10):          # four launches, nothing read in between
            ctr.add_device_frames(dev, first, 10)
        np.testing.assert_array_equal(ctr.counts(), ocounts)
        np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
        st = ctr.stats()
        assert st.hits == int(ohits.sum()) and st.private_launches == 4
        if kind == "one cell":
            assert st.private_redone >= 1, "overflowing lanes were not caught by the digit-sum check"
        if kind == "uniform":
            assert st.private_redone == 0
        # more frames on top of a grid that has been read (the copies were folded in and cleared)
        ctr.add_device_frames(dev, 0, 7)
        o2, h2, _ = oracle_counts(geom, x[:7], gindex)
        np.testing.assert_array_equal(ctr.counts(), ocounts + o2)
        # density (K2) straight after binning: the lazy flush must come first
        ctr.reset()
        ctr.add_device_frames(dev, 3, 9)
        o3, _, _ = oracle_counts(geom, x[3:12], gindex)
        occ = ctr.occupancy()
        np.testing.assert_array_equal(occ, o3.astype(np.float32))
        # and a reset drops what the copies hold
        ctr.add_device_frames(dev, 0, 5)
        ctr.reset()
        ctr.add_device_frames(dev, 20, 4)
        o4, _, _ = oracle_counts(geom, x[20:24], gindex)
        np.testing.assert_array_equal(ctr.counts(), o4)
        ctr.close()
    dev.free()


def test_private_backoff_recovers():
    """after a launch that overflowed, the next launches go to the uint32 grid; a spread-out input later returns to
    the private grids (the back-off halves while launches pass)"""
    L, delta = (4.0, 4.0, 4.0), 0.1
    geom = g.setup_geometry(L, delta)
    natoms = 8192
    rng = np.random.default_rng(3)
    spread = (rng.random((12, natoms, 3)) * np.asarray(L)).astype(np.float32)
    piled = np.broadcast_to(np.asarray([1.234, 2.345, 3.456], np.float32), (12, natoms, 3)).copy()
    gindex = np.arange(natoms, dtype=np.int32)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_PRIVATE)
    total = np.zeros(geom.shape, np.uint32)
    for batch in (spread, piled, spread, spread, spread, spread, spread):
        ctr.add_frames(batch)
        total += oracle_counts(geom, batch, gindex)[0]
    np.testing.assert_array_equal(ctr.counts(), total)
    st = ctr.stats()
    assert st.private_redone >= 1 and st.private_launches == 7
    ctr.close()


def test_auto_takes_the_private_path_on_the_c1_shape():
    """BASELINE config 1 (2700 atoms, 900 OW, 60^3 cells as 4-bit private counters): AUTO bins a large batch through
    the private grids, a small one straight into the uint32 grid, and mixing both is exact"""
    L = (3.0, 3.0, 3.0)
    natoms = 2700
    geom = g.setup_geometry(L, 0.05)
    assert geom.shape == (60, 60, 60)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    big, small = 16384, 64
    dev = g.DeviceFrames(big, natoms).generate(g.make_gen(77, L))
    ctr = g.GridCounter(geom, gindex, natoms)
    ctr.add_device_frames(dev)                      # 14.7 M atom-frames: private
    ctr.add_device_frames(dev, 0, small)            # 57 600: direct
    st = ctr.stats()
    assert st.private_launches == 1 and st.private_redone == 0
    assert st.hits == (big + small) * len(gindex)
    c = ctr.counts()
    assert int(c.sum(dtype=np.uint64)) == st.hits
    want, _, _ = oracle_counts(geom, dev.download(0, small), gindex)
    ref = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_STREAM)
    ref.add_device_frames(dev)
    np.testing.assert_array_equal(c, ref.counts() + want)
    ref.close(); ctr.close(); dev.free()


def test_nibble_rounds_through_the_byte_grid_stay_exact(monkeypatch):
    """4-bit packed rounds are added to a byte grid for up to 16 rounds and folded into the uint32 grid lazily:
    40 rounds with a read in the middle, a reset, and GCB_NO_MID as the cross-check"""
    monkeypatch.setenv("GCB_PACKED_BITS", "4")
    monkeypatch.setenv("GCB_PACKED_UNITS", str(8192 * 4))            # 4 frames per round
    L, delta = (4.0, 4.0, 4.0), 0.1
    geom = g.setup_geometry(L, delta)
    natoms, nframes = 8192, 160
    x = u.gen_frames(u.make_gen(77, L), 0, nframes, natoms)
    gindex = np.arange(natoms, dtype=np.int32)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    o_half, _, _ = oracle_counts(geom, x[:100], gindex)
    for nomid in (False, True):
        if nomid:
            monkeypatch.setenv("GCB_NO_MID", "1")
        ctr = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_PACKED)
        ctr.add_frames(x[:100], np.ones(100, np.float32))            # 25 rounds: the byte grid is folded in once on the way
        np.testing.assert_array_equal(ctr.counts(), o_half)          # ... and again by this read
        ctr.add_frames(x[100:], np.ones(60, np.float32))
        np.testing.assert_array_equal(ctr.occupancy(), ocounts.astype(np.float32))
        np.testing.assert_array_equal(ctr.counts(), ocounts)
        st = ctr.stats()
        assert st.packed_rounds == 40 and st.packed_redone == 0 and st.hits == int(ohits.sum())
        ctr.reset()
        ctr.add_frames(x[:8], np.ones(8, np.float32))
        ctr.reset()                                                  # drops what the byte grid holds
        ctr.add_frames(x[8:20], np.ones(12, np.float32))
        np.testing.assert_array_equal(ctr.counts(), oracle_counts(geom, x[8:20], gindex)[0])
        ctr.close()


def test_pinned_source_may_be_refilled_as_soon_as_add_frames_returns():
    """gcb_counter_add_frames DMAs pinned memory in place; its contract is that x may be refilled when the call returns
    (the natural read loop reuses one buffer).  Frames are overwritten with rubbish right after every call."""
    L = (4.0, 4.0, 4.0)
    natoms, chunk, nchunks = 50000, 40, 6            # 24 MB per call: several staging batches in flight
    geom = g.setup_geometry(L, 0.1)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    x = u.gen_frames(u.make_gen(91, L), 0, chunk * nchunks, natoms)
    ocounts, ohits, _ = oracle_counts(geom, x, gindex)
    pin = g.PinnedFrames(chunk, natoms)
    ctr = g.GridCounter(geom, gindex, natoms, frames_per_batch=4)
    for c in range(nchunks):
        pin.array[:] = x[c * chunk:(c + 1) * chunk]
        ctr.add_host_ptr(pin.ptr, chunk)
        pin.array[:] = -1.0e9                         # the next chunk is "read" into the same buffer at once
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    ctr.close(); pin.free()


def test_nibble_grid_falls_back_to_byte_lanes_on_clustered_input():
    """500^3 (nibble lanes by default): atoms piled around 256 sites overflow the nibbles; AUTO then retries with byte
    lanes and short rounds, which hold (peak density far below 255 per round), instead of giving the packed path up;
    counts stay exact through the switch, and uniform input afterwards goes back to the nibbles"""
    L, delta, natoms, nframes = (10.0, 10.0, 10.0), 0.02, 100000, 960
    geom = g.setup_geometry(L, delta)
    assert geom.shape == (500, 500, 500)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    dev = g.DeviceFrames(nframes, natoms).generate(g.make_gen(41, L, K.GEN_CLUSTER, nsites=256, sigma=0.1))
    ctr = g.GridCounter(geom, gindex, natoms)
    NC = 14
    for _ in range(NC):
        ctr.add_device_frames(dev)
        ctr.sync()
    st = ctr.stats()
    assert st.hits == NC * nframes * len(gindex)
    want, whits, _ = oracle_counts(geom, dev.download(0, 64), gindex)
    ref = g.GridCounter(geom, gindex, natoms, kernel=K.KERNEL_STREAM)
    ref.add_device_frames(dev)
    c1 = ref.counts()
    np.testing.assert_array_equal(ctr.counts(), NC * c1)
    ref.reset(); ref.add_device_frames(dev, 0, 64)
    np.testing.assert_array_equal(ref.counts(), want)                 # the direct kernel itself against the oracle
    redone_before = st.packed_redone
    for _ in range(4):                                                # by now on byte lanes: rounds hold
        ctr.add_device_frames(dev)
        ctr.sync()
    st2 = ctr.stats()
    assert st2.packed_rounds > st.packed_rounds and st2.packed_redone == redone_before, (st2.packed_rounds, st2.packed_redone, redone_before)
    np.testing.assert_array_equal(ctr.counts(), (NC + 4) * c1)
    ref.close(); ctr.close(); dev.free()
