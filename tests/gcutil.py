"""Test helpers: ctypes access to the CPU oracle (oracle/libgcoracle.so), to the
compiled UNMODIFIED reference (oracle/_ref, when present) and small file writers
for the raw trajectory / topology stub / .ndx formats the reference shim reads.

TEST INFRASTRUCTURE ONLY -- nothing in the product package imports this."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u64_p = C.POINTER(C.c_uint64)
c_u32_p = C.POINTER(C.c_uint32)
c_u8_p = C.POINTER(C.c_uint8)
c_double_p = C.POINTER(C.c_double)


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


class Cavity(C.Structure):
    _fields_ = [("axis", C.c_float * 3), ("cpoint", C.c_float * 3), ("radius", C.c_float),
                ("z1", C.c_float), ("z2", C.c_float), ("vol", C.c_float)]


class Geom(C.Structure):
    _fields_ = [("mx", C.c_int * 3), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("delta", C.c_float * 3), ("tweight", C.c_float)]

    def cells(self):
        return self.mx[0] * self.mx[1] * self.mx[2]

    def shape(self):
        return (self.mx[0], self.mx[1], self.mx[2])


class Analysis(C.Structure):
    _fields_ = [("unit", C.c_int), ("min_occ", C.c_float), ("rad_ba_nr", C.c_int), ("rad_ba_step", C.c_int),
                ("geometry", Cavity), ("DeltaR", C.c_float), ("dV", C.c_double), ("iradNR", C.c_int),
                ("unit_value", C.c_float), ("sumP", C.c_double), ("sumH", C.c_double),
                ("volume", C.c_float), ("reff", C.c_float), ("average", C.c_float), ("height", C.c_float),
                ("ivol", C.c_int),
                ("xyp", c_float_p), ("xzp", c_float_p), ("yzp", c_float_p), ("hxyp", c_float_p),
                ("zdf", c_float_p), ("lzdf", c_float_p), ("profile", c_float_p), ("rweight", c_float_p),
                ("rdf", c_float_p), ("hrdf", c_float_p), ("rzp", c_float_p), ("hrzp", c_float_p)]


class Gen(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mode", C.c_int), ("L", C.c_float * 3),
                ("slab_lo", C.c_float), ("slab_hi", C.c_float), ("pore_r", C.c_float),
                ("nsites", C.c_int), ("sigma", C.c_float), ("box_period", C.c_int), ("box_amp", C.c_float)]


UNITS = {"unity": 0, "SPC": 1, "molar": 2, "Angstrom": 3, "Voxel": 4}
SET_CPOINT, SET_R, SET_Z1, SET_Z2 = 1, 2, 4, 8

_oracle = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def oracle():
    """The plain-C restatement (oracle/gc_oracle.c)."""
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(ORACLE_DIR, "libgcoracle.so")
    src = os.path.join(ORACLE_DIR, "gc_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build_oracle()
    L = C.CDLL(path)
    L.gco_autoset_cavity.argtypes = [C.POINTER(Cavity), c_float_p, C.c_int]
    L.gco_setup_grid.argtypes = [C.POINTER(Geom), C.POINTER(Cavity), c_float_p]
    L.gco_delta_v.argtypes = [C.POINTER(Geom)]
    L.gco_delta_v.restype = C.c_double
    L.gco_edge_vulnerable.argtypes = [C.POINTER(Geom)]
    L.gco_frame_weights.argtypes = [c_float_p, C.c_long, C.c_int, c_float_p, c_double_p]
    L.gco_bin_frames.argtypes = [C.POINTER(Geom), c_float_p, c_float_p, C.c_long, C.c_int, c_int_p, C.c_int,
                                 c_float_p, c_double_p, c_u64_p, c_u64_p]
    L.gco_bin_frames_f64.argtypes = [C.POINTER(Geom), c_double_p, c_float_p, C.c_long, C.c_int, c_int_p, C.c_int, c_float_p]
    L.gco_count_frames.argtypes = [C.POINTER(Geom), c_u32_p, c_float_p, C.c_long, C.c_int, c_int_p, C.c_int,
                                   c_u64_p, c_u64_p]
    L.gco_wrap_frames.argtypes = [c_float_p, C.c_long, C.c_int, c_float_p]
    L.gco_seq_add_f32.argtypes = [C.c_float, C.c_float, C.c_uint64]
    L.gco_seq_add_f32.restype = C.c_float
    L.gco_seq_add_f64.argtypes = [C.c_double, C.c_double, C.c_uint64]
    L.gco_seq_add_f64.restype = C.c_double
    L.gco_normalise.argtypes = [C.POINTER(Geom), c_float_p, C.c_double]
    L.gco_header.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_char_p, C.POINTER(Cavity), C.POINTER(Geom),
                             C.c_double, C.c_char_p, C.c_int]
    L.gco_xdr_size.argtypes = [C.POINTER(Geom), C.c_char_p]
    L.gco_xdr_size.restype = C.c_size_t
    L.gco_xdr_encode.argtypes = [C.POINTER(Geom), c_float_p, C.c_char_p, c_u8_p, C.c_size_t]
    L.gco_xdr_encode.restype = C.c_size_t
    L.gco_xdr_decode.argtypes = [c_u8_p, C.c_size_t, C.POINTER(Geom), C.c_char_p, C.POINTER(c_float_p)]
    L.gco_gridcalc.argtypes = [C.c_int, C.POINTER(c_float_p), c_float_p, C.c_size_t, C.c_size_t, C.c_size_t,
                               c_float_p, c_float_p]
    L.gco_plt_encode.argtypes = [C.POINTER(Geom), c_float_p, C.c_float, c_u8_p, C.c_size_t]
    L.gco_plt_encode.restype = C.c_size_t
    L.gco_update_b.argtypes = [C.POINTER(Geom)]
    L.gco_tgrid2cavity.argtypes = [C.POINTER(Geom), C.POINTER(Cavity)]
    L.gco_readjust.argtypes = [C.POINTER(Geom), C.POINTER(c_float_p), C.POINTER(Cavity), C.c_int]
    L.gco_analyse.argtypes = [C.POINTER(Geom), c_float_p, C.POINTER(Analysis)]
    L.gco_analysis_free.argtypes = [C.POINTER(Analysis)]
    L.gco_gen_frames.argtypes = [C.POINTER(Gen), C.c_long, C.c_long, C.c_int, c_float_p]
    L.gco_gen_box.argtypes = [C.POINTER(Gen), C.c_long, c_float_p]
    L.gco_bench.argtypes = [C.POINTER(Geom), c_float_p, C.c_long, C.c_int, c_int_p, C.c_int, C.c_float, C.c_int,
                            c_double_p]
    L.gco_bench.restype = C.c_double
    L.free = C.CDLL(None).free
    L.free.argtypes = [C.c_void_p]
    _oracle = L
    return L


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libgcref.so")) and os.path.exists(
        os.path.join(REF_DIR, "g_ri3Dc_ref"))


_ref = None


def ref():
    """The UNMODIFIED reference compiled by oracle/Makefile (function-level harness)."""
    global _ref
    if _ref is not None:
        return _ref
    L = C.CDLL(os.path.join(REF_DIR, "libgcref.so"))
    L.gcref_grid_create.argtypes = [c_float_p, c_float_p, C.c_int, c_float_p, C.c_float, C.c_float, C.c_float]
    L.gcref_grid_create.restype = C.c_void_p
    L.gcref_grid_info.argtypes = [C.c_void_p, c_int_p, c_float_p, c_float_p, c_float_p, c_float_p]
    L.gcref_grid_data.argtypes = [C.c_void_p]
    L.gcref_grid_data.restype = c_float_p
    L.gcref_grid_free.argtypes = [C.c_void_p]
    L.gcref_grid_set_tweight.argtypes = [C.c_void_p, C.c_float]
    L.gcref_deltaV.argtypes = [C.c_void_p]
    L.gcref_deltaV.restype = C.c_double
    L.gcref_bin_frames.argtypes = [C.c_void_p, c_float_p, C.c_long, C.c_int, c_int_p, C.c_int, c_float_p,
                                   c_double_p]
    L.gcref_grid_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.gcref_grid_read.argtypes = [C.c_char_p, C.c_char_p]
    L.gcref_grid_read.restype = C.c_void_p
    L.gcref_bench.argtypes = [c_float_p, c_float_p, c_float_p, C.c_long, C.c_int, c_int_p, C.c_int, C.c_float,
                              C.c_int, c_double_p, c_float_p]
    L.gcref_bench.restype = C.c_double
    _ref = L
    return L


# ---------------------------------------------------------------- oracle wrappers
def box9(L):
    b = np.zeros(9, np.float32)
    b[0], b[4], b[8] = L
    return b


def oracle_geometry(L, delta, setmask=0, cpoint=(0, 0, 0), radius=0.0, z1=0.0, z2=0.0):
    o = oracle()
    cav = Cavity()
    cav.axis[2] = 1
    for d in range(3):
        cav.cpoint[d] = cpoint[d]
    cav.radius, cav.z1, cav.z2 = radius, z1, z2
    b9 = box9(L)
    o.gco_autoset_cavity(C.byref(cav), fptr(b9), setmask)
    g = Geom()
    dl = np.asarray(delta, np.float32)
    ok = o.gco_setup_grid(C.byref(g), C.byref(cav), fptr(dl))
    assert ok
    return g, cav


def ref_geometry(L, delta, setmask=0, cpoint=(0, 0, 0), radius=0.0, z1=0.0, z2=0.0):
    r = ref()
    b9 = box9(L)
    dl = np.asarray(delta, np.float32)
    cp = np.asarray(cpoint, np.float32)
    h = r.gcref_grid_create(fptr(b9), fptr(dl), setmask, fptr(cp), radius, z1, z2)
    assert h
    mx = np.zeros(3, np.int32)
    a = np.zeros(3, np.float32)
    b = np.zeros(3, np.float32)
    d = np.zeros(3, np.float32)
    cav = np.zeros(7, np.float32)
    r.gcref_grid_info(h, iptr(mx), fptr(a), fptr(b), fptr(d), fptr(cav))
    return h, mx, a, b, d, cav


def gen_frames(gen, frame0, nframes, natoms):
    x = np.empty((nframes, natoms, 3), np.float32)
    oracle().gco_gen_frames(C.byref(gen), frame0, nframes, natoms, fptr(x))
    return x


def make_gen(seed, L, mode=0, slab=(0, 0), pore_r=0.0, nsites=1, sigma=0.0, box_period=0, box_amp=0.0):
    g = Gen()
    g.seed = seed
    g.mode = mode
    for d in range(3):
        g.L[d] = L[d]
    g.slab_lo, g.slab_hi, g.pore_r = slab[0], slab[1], pore_r
    g.nsites, g.sigma = nsites, sigma
    g.box_period, g.box_amp = box_period, box_amp
    return g


def analysis_arrays(an, g):
    """Copy the arrays of a gco_analysis into numpy (before gco_analysis_free)."""
    mx, my, mz = g.mx[0], g.mx[1], g.mx[2]
    nr = an.iradNR

    def arr(p, shape):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, np.float32)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape).copy()

    return dict(xyp=arr(an.xyp, (mx, my)), xzp=arr(an.xzp, (mx, mz)), yzp=arr(an.yzp, (my, mz)),
                hxyp=arr(an.hxyp, (mx, my)), zdf=arr(an.zdf, (mz, 2)), lzdf=arr(an.lzdf, (mz, 2)),
                profile=arr(an.profile, (mz, 2)), rweight=arr(an.rweight, (nr,)), rdf=arr(an.rdf, (nr, 2)),
                hrdf=arr(an.hrdf, (nr, 2)), rzp=arr(an.rzp, (nr, mz)), hrzp=arr(an.hrzp, (nr, mz)))


# ---------------------------------------------------------------- files for the reference binaries
def write_raw_traj(path, t, boxes, x):
    """GCTRJ1 raw trajectory (oracle/shim/fakegmx.c)."""
    nframes, natoms = x.shape[0], x.shape[1]
    with open(path, "wb") as f:
        f.write(b"GCTRJ1\n\0")
        f.write(struct.pack("<ii", natoms, nframes))
        for i in range(nframes):
            f.write(np.float32(t[i]).tobytes())
            f.write(np.asarray(boxes[i], np.float32).reshape(9).tobytes())
            f.write(np.ascontiguousarray(x[i], np.float32).tobytes())


def write_top(path, natoms):
    with open(path, "w") as f:
        f.write("GCTOP1 %d\n" % natoms)


def write_ndx(path, name, index0):
    """.ndx: 1-based, 15 per line (count.c:158-176)."""
    with open(path, "w") as f:
        f.write("[ %s ]\n" % name)
        for i, v in enumerate(index0):
            f.write("%5d " % (int(v) + 1))
            if (i + 1) % 15 == 0:
                f.write("\n")
        f.write("\n")


def run_ref(tool, args, cwd, capture=False):
    env = dict(os.environ)
    if capture:
        env["GCSHIM_CAPTURE"] = "1"
    p = subprocess.run([os.path.join(REF_DIR, tool)] + [str(a) for a in args], cwd=cwd, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError("%s failed:\n%s" % (tool, p.stderr.decode(errors="replace")))
    return p.stdout.decode(errors="replace"), p.stderr.decode(errors="replace")


def bits(a):
    """Bit pattern view for exact float comparisons (NaN-safe)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def assert_bits_equal(x, y, what=""):
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y)
    assert x.shape == y.shape, "%s: shape %s vs %s" % (what, x.shape, y.shape)
    bx, by = bits(x), bits(y)
    # all NaNs compare equal to each other (payload/sign of 0/0 differs between compilers and devices)
    if x.dtype.kind == "f":
        nx, ny = np.isnan(x), np.isnan(y)
        assert np.array_equal(nx, ny), "%s: NaN positions differ" % what
        ok = (bx == by) | nx
    else:
        ok = bx == by
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError("%s: %d of %d elements differ; first at %s: %r vs %r"
                             % (what, len(bad), x.size, i, x[i], y[i]))
