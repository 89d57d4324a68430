"""Thin object layer over the C ABI (names follow the reference's domain:
cavity, tgrid, index group, frames, density, header)."""
import ctypes as C

import numpy as np

from . import _cabi as K
from ._cabi import lib, check


def _fptr(a):
    return a.ctypes.data_as(K.c_float_p) if a is not None else None


def device_count():
    return lib.gcb_device_count()


class Geometry:
    """t_cavity + t_tgrid as set up by autoset_cavity/setup_tgrid (count.c:25-64, grid3D.c:125-203)."""

    def __init__(self, tg, cav):
        self.tg, self.cav = tg, cav

    @property
    def shape(self):
        return self.tg.shape

    @property
    def cells(self):
        return self.tg.cells

    @property
    def dV(self):
        return lib.gcb_delta_v(C.byref(self.tg))

    @property
    def edge_vulnerable(self):
        return bool(lib.gcb_edge_vulnerable(C.byref(self.tg)))


def setup_geometry(box, delta, cpoint=None, radius=None, z1=None, z2=None):
    """box: 3 edge lengths or 9-float matrix of the FIRST frame; options as g_ri3Dc's -cpoint/-R/-z1/-z2."""
    b = np.zeros(9, np.float32)
    box = np.asarray(box, np.float32).ravel()
    if box.size == 3:
        b[0], b[4], b[8] = box
    else:
        b[:] = box
    cav = K.Cavity()
    mask = 0
    if cpoint is not None:
        mask |= K.SET_CPOINT
        for d in range(3):
            cav.cpoint[d] = cpoint[d]
    if radius is not None:
        mask |= K.SET_R
        cav.radius = radius
    if z1 is not None:
        mask |= K.SET_Z1
        cav.z1 = z1
    if z2 is not None:
        mask |= K.SET_Z2
        cav.z2 = z2
    check(lib.gcb_autoset_cavity(C.byref(cav), _fptr(b), mask), "gcb_autoset_cavity")
    tg = K.TGrid()
    dl = np.asarray(delta, np.float32).ravel()
    if dl.size == 1:
        dl = np.repeat(dl, 3)
    check(lib.gcb_setup_tgrid(C.byref(tg), C.byref(cav), _fptr(dl)), "gcb_setup_tgrid")
    return Geometry(tg, cav)


def frame_weights(t, dtweight=True):
    """-> (weights f32[F], T_tot) as the frame loop computes them (g_ri3Dc.c:390-394, 413-416)."""
    t = np.ascontiguousarray(t, np.float32)
    w = np.zeros(len(t), np.float32)
    T = C.c_double(0)
    check(lib.gcb_frame_weights(_fptr(t), len(t), int(bool(dtweight)), None, _fptr(w), C.byref(T)),
          "gcb_frame_weights")
    return w, T.value


def header_string(geom, T_tot, sumP, grpname, trxname, subtitle="", dtweight=True):
    out = C.create_string_buffer(K.HEADER_MAX)
    check(lib.gcb_header(out, subtitle.encode(), T_tot, grpname.encode(), C.byref(geom.cav), C.byref(geom.tg), sumP,
                         trxname.encode(), int(bool(dtweight))), "gcb_header")
    return out.value


def xdr_encode(tg, header, grid_zfast=None, grid_xfast=None):
    n = lib.gcb_xdr_size(C.byref(tg), header)
    buf = np.zeros(n, np.uint8)
    wrote = C.c_size_t(0)
    check(lib.gcb_xdr_encode(C.byref(tg), _fptr(grid_zfast), _fptr(grid_xfast), header,
                             buf.ctypes.data_as(K.c_u8_p), n, C.byref(wrote)), "gcb_xdr_encode")
    return buf[:wrote.value].tobytes()


def grid_write(path, tg, header, grid_zfast=None, grid_xfast=None):
    check(lib.gcb_grid_write(str(path).encode(), C.byref(tg), _fptr(grid_zfast), _fptr(grid_xfast), header),
          "gcb_grid_write")


def grid_read(path):
    tg = K.TGrid()
    hdr = C.create_string_buffer(K.HEADER_MAX + 1)
    gp = K.c_float_p()
    check(lib.gcb_grid_read(str(path).encode(), C.byref(tg), hdr, C.byref(gp)), "gcb_grid_read")
    grid = np.ctypeslib.as_array(gp, shape=(tg.cells,)).reshape(tg.shape).copy()
    lib.gcb_free(gp)
    return tg, hdr.value, grid


def make_gen(seed, L, mode=K.GEN_UNIFORM, slab=(0.0, 0.0), pore_r=0.0, nsites=1, sigma=0.0, box_period=0,
             box_amp=0.0):
    g = K.Gen()
    g.seed, g.mode = seed, mode
    for d in range(3):
        g.L[d] = L[d]
    g.slab_lo, g.slab_hi, g.pore_r = slab[0], slab[1], pore_r
    g.nsites, g.sigma, g.box_period, g.box_amp = nsites, sigma, box_period, box_amp
    return g


class DeviceFrames:
    """x[nframes][natoms][3] f32 resident in HBM (filled by the synthetic generator or from the host)."""

    def __init__(self, nframes, natoms, device=0):
        self.nframes, self.natoms, self.device = nframes, natoms, device
        self.nbytes = nframes * natoms * 12
        p = C.c_void_p()
        check(lib.gcb_device_malloc(C.byref(p), max(self.nbytes, 16), device), "gcb_device_malloc")
        self.ptr = p

    def generate(self, gen, frame0=0):
        check(lib.gcb_gen_device_frames(C.byref(gen), frame0, self.nframes, self.natoms, self.ptr, self.device),
              "gcb_gen_device_frames")
        return self

    def upload(self, x):
        x = np.ascontiguousarray(x, np.float32)
        assert x.nbytes == self.nbytes
        check(lib.gcb_memcpy_h2d(self.ptr, x.ctypes.data, x.nbytes, self.device), "gcb_memcpy_h2d")
        return self

    def download(self, first=0, count=None):
        count = self.nframes - first if count is None else count
        out = np.empty((count, self.natoms, 3), np.float32)
        src = C.c_void_p(self.ptr.value + first * self.natoms * 12)
        check(lib.gcb_memcpy_d2h(out.ctypes.data, src, out.nbytes, self.device), "gcb_memcpy_d2h")
        return out

    def frame_ptr(self, first):
        return C.c_void_p(self.ptr.value + first * self.natoms * 12)

    def free(self):
        if self.ptr:
            lib.gcb_device_free(self.ptr, self.device)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedFrames:
    """Page-locked host frames, the source of the end-to-end (host buffer) path."""

    def __init__(self, nframes, natoms):
        self.nframes, self.natoms = nframes, natoms
        self.nbytes = nframes * natoms * 12
        p = C.c_void_p()
        check(lib.gcb_host_alloc_pinned(C.byref(p), max(self.nbytes, 16)), "gcb_host_alloc_pinned")
        self.ptr = p
        buf = (C.c_float * (nframes * natoms * 3)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.float32).reshape(nframes, natoms, 3)

    def free(self):
        if self.ptr:
            self.array = None
            lib.gcb_host_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class XtcBuffer:
    """The bytes of an .xtc file (or any run of whole frames) plus the table of its frame headers."""

    def __init__(self, data, pinned=False):
        data = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
        self.nbytes = int(data.nbytes)
        self._pin = None
        if pinned:
            p = C.c_void_p()
            check(lib.gcb_host_alloc_pinned(C.byref(p), max(self.nbytes, 16)), "gcb_host_alloc_pinned")
            C.memmove(p.value, data.ctypes.data, self.nbytes)
            self._pin = p
            self.ptr = p
            self._keep = None
        else:
            self._keep = np.ascontiguousarray(data)
            self.ptr = C.c_void_p(self._keep.ctypes.data)
        n = C.c_long(0)
        used = C.c_size_t(0)
        check(lib.gcb_xtc_scan(self.ptr, self.nbytes, None, 0, C.byref(n), C.byref(used)), "gcb_xtc_scan")
        self.nframes = n.value
        self.frames = (K.XtcFrame * max(self.nframes, 1))()
        check(lib.gcb_xtc_scan(self.ptr, self.nbytes, self.frames, self.nframes, C.byref(n), C.byref(used)), "gcb_xtc_scan")
        self.consumed = used.value
        self.natoms = self.frames[0].natoms if self.nframes else 0

    @property
    def times(self):
        return np.array([self.frames[i].time for i in range(self.nframes)], np.float32)

    def decode_to_device(self, device=0, dev=None):
        """decoded frames x[nframes][natoms][3] in device memory (`dev`: an existing DeviceFrames to decode into)"""
        dev = dev if dev is not None else DeviceFrames(self.nframes, self.natoms, device)
        check(lib.gcb_xtc_decode_device(self.ptr, self.nbytes, self.frames, self.nframes, self.natoms, dev.ptr, device),
              "gcb_xtc_decode_device")
        return dev

    def free(self):
        if self._pin:
            lib.gcb_host_free_pinned(self._pin)
            self._pin = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GridCounter:
    """The g_ri3Dc binning loop (g_ri3Dc.c:410-429) on one GPU."""

    def __init__(self, geom, gindex, natoms, device=0, pbc=K.PBC_NONE, kernel=K.KERNEL_AUTO, nslots=0,
                 frames_per_batch=0):
        self.geom = geom
        self.natoms = natoms
        self.gindex = np.ascontiguousarray(gindex, np.int32)
        opts = K.CounterOpts(device, pbc, kernel, nslots, frames_per_batch)
        h = C.c_void_p()
        check(lib.gcb_counter_create(C.byref(h), C.byref(geom.tg), self.gindex.ctypes.data_as(K.c_int_p),
                                     len(self.gindex), natoms, C.byref(opts)), "gcb_counter_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.gcb_counter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(lib.gcb_counter_reset(self.h), "gcb_counter_reset")

    reset_grid = reset

    def density_device(self, T_tot):
        """K2 only; the density stays in HBM (returns the device pointer)"""
        p = C.c_void_p()
        check(lib.gcb_counter_device_density(self.h, T_tot, C.byref(p)), "gcb_counter_device_density")
        return p

    def density_into(self, T_tot, out_zfast):
        check(lib.gcb_counter_density(self.h, T_tot, _fptr(out_zfast), None, None), "gcb_counter_density")

    def add_frames(self, x, weights=None, box=None):
        """x: host array [F][natoms][3] (numpy or PinnedFrames.array)"""
        x = np.ascontiguousarray(x, np.float32)
        nframes = x.shape[0]
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        b = None if box is None else np.ascontiguousarray(box, np.float32)
        check(lib.gcb_counter_add_frames(self.h, x.ctypes.data, nframes, _fptr(w), _fptr(b)),
              "gcb_counter_add_frames")

    def add_host_ptr(self, ptr, nframes, weights=None):
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        check(lib.gcb_counter_add_frames(self.h, ptr, nframes, _fptr(w), None), "gcb_counter_add_frames")

    def add_device_frames(self, dev, first=0, count=None, weights=None, box=None):
        count = dev.nframes - first if count is None else count
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        b = None if box is None else np.ascontiguousarray(box, np.float32)
        check(lib.gcb_counter_add_device_frames(self.h, dev.frame_ptr(first), count, _fptr(w), _fptr(b)),
              "gcb_counter_add_device_frames")

    def add_xtc(self, xtc, weights=None):
        """xtc: XtcBuffer (compressed frames in host memory); decoded and binned on the GPU"""
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        check(lib.gcb_counter_add_xtc(self.h, xtc.ptr, xtc.nbytes, xtc.frames, xtc.nframes, _fptr(w)),
              "gcb_counter_add_xtc")

    def sync(self):
        check(lib.gcb_counter_sync(self.h), "gcb_counter_sync")

    def set_host_pack(self, mode):
        """K.HOST_PACK_AUTO / _NEVER / _ALWAYS: how add_frames stages host frames (see the header)"""
        check(lib.gcb_counter_set_host_pack(self.h, mode), "gcb_counter_set_host_pack")

    def host_staging(self):
        s = K.HostStaging()
        check(lib.gcb_counter_host_staging(self.h, C.byref(s)), "gcb_counter_host_staging")
        return s

    def stats(self):
        s = K.Stats()
        check(lib.gcb_counter_stats(self.h, C.byref(s)), "gcb_counter_stats")
        return s

    def hits_per_frame(self):
        n = C.c_long(0)
        check(lib.gcb_counter_hits_per_frame(self.h, None, 0, C.byref(n)), "gcb_counter_hits_per_frame")
        out = np.zeros(n.value, np.uint64)
        check(lib.gcb_counter_hits_per_frame(self.h, out.ctypes.data_as(K.c_u64_p), n.value, C.byref(n)),
              "gcb_counter_hits_per_frame")
        return out

    def counts(self):
        out = np.zeros(self.geom.shape, np.uint32)
        check(lib.gcb_counter_counts(self.h, out.ctypes.data_as(K.c_u32_p)), "gcb_counter_counts")
        return out

    def occupancy(self):
        out = np.zeros(self.geom.shape, np.float32)
        check(lib.gcb_counter_occupancy(self.h, _fptr(out)), "gcb_counter_occupancy")
        return out

    def density(self, T_tot, zfast=True, xfast=False):
        """-> (tgrid with tweight, grid z-fastest [mx,my,mz] or None, grid in file order [mz,my,mx] or None)"""
        mx, my, mz = self.geom.shape
        gz = np.zeros((mx, my, mz), np.float32) if zfast else None
        gx = np.zeros((mz, my, mx), np.float32) if xfast else None
        tg = K.TGrid()
        check(lib.gcb_counter_density(self.h, T_tot, _fptr(gz), _fptr(gx), C.byref(tg)), "gcb_counter_density")
        return tg, gz, gx

    def xdr(self, T_tot, header=b""):
        """the finished density as the grid file image, reordered and byte-swapped on the device -> (tgrid, bytes)"""
        tg = K.TGrid.from_buffer_copy(self.geom.tg)
        n = lib.gcb_xdr_size(C.byref(tg), header)
        buf = np.zeros(n, np.uint8)
        wrote = C.c_size_t(0)
        check(lib.gcb_counter_xdr(self.h, T_tot, header, buf.ctypes.data_as(K.c_u8_p), n, C.byref(wrote), C.byref(tg)),
              "gcb_counter_xdr")
        return tg, buf[:wrote.value].tobytes()

    def grid_write(self, T_tot, path, header=b""):
        tg = K.TGrid()
        check(lib.gcb_counter_grid_write(self.h, T_tot, str(path).encode(), header, C.byref(tg)), "gcb_counter_grid_write")
        return tg

    def plt(self, T_tot, unit=1.0):
        n = 44 + 4 * self.geom.cells
        buf = np.zeros(n, np.uint8)
        wrote = C.c_size_t(0)
        check(lib.gcb_counter_plt(self.h, T_tot, unit, buf.ctypes.data_as(K.c_u8_p), n, C.byref(wrote)), "gcb_counter_plt")
        return buf[:wrote.value].tobytes()

    def event_record(self, i):
        check(lib.gcb_counter_event_record(self.h, i), "gcb_counter_event_record")

    def event_elapsed(self, i, j):
        ms = C.c_float(0)
        check(lib.gcb_counter_event_elapsed(self.h, i, j, C.byref(ms)), "gcb_counter_event_elapsed")
        return ms.value

    def profile(self, enable=True):
        """event pairs around every launch of the streaming kernel from now on (gcb_counter_profile)"""
        check(lib.gcb_counter_profile(self.h, int(bool(enable))), "gcb_counter_profile")

    def profile_read(self):
        """-> (summed launch durations in ms, launches, atom-frames those launches covered)"""
        ms, n, units = C.c_double(0), C.c_long(0), C.c_double(0)
        check(lib.gcb_counter_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(units)), "gcb_counter_profile_read")
        return ms.value, n.value, units.value

    def profile_kinds(self):
        """-> {kind: (summed ms, launches, atom-frames)} for 'direct', 'packed', 'redo' (the guarded uint32 pass
        behind a packed pass), 'other' (verify/apply passes, gather tails, fused xtc kernel) and 'private'"""
        ms, n, units = (C.c_double * 5)(), (C.c_long * 5)(), (C.c_double * 5)()
        check(lib.gcb_counter_profile_kinds(self.h, ms, n, units), "gcb_counter_profile_kinds")
        return {k: (ms[i], n[i], units[i]) for i, k in enumerate(("direct", "packed", "redo", "other", "private"))}

    def timer_start(self):
        check(lib.gcb_counter_timer_start(self.h), "gcb_counter_timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        check(lib.gcb_counter_timer_stop(self.h, C.byref(ms)), "gcb_counter_timer_stop")
        return ms.value

    def comm_init(self, uid, rank, nranks):
        check(lib.gcb_counter_comm_init(self.h, uid, rank, nranks), "gcb_counter_comm_init")

    def allreduce(self):
        check(lib.gcb_counter_allreduce(self.h), "gcb_counter_allreduce")


def pack_frames(x, gindex, threads=0):
    """out[f][i] = x[f][gindex[i]] by the library's host thread pool (the packing step of add_frames by itself)"""
    x = np.ascontiguousarray(x, np.float32)
    gi = np.ascontiguousarray(gindex, np.int32)
    out = np.empty((x.shape[0], len(gi), 3), np.float32)
    check(lib.gcb_pack_frames(_fptr(x), x.shape[0], x.shape[1], gi.ctypes.data_as(K.c_int_p), len(gi), _fptr(out), threads),
          "gcb_pack_frames")
    return out


def selftest_l2(window_bytes, mix, device=0, stream_bytes=0):
    """gcb_selftest_l2: what the memory system sustains for K1's traffic mix -> (G REDs/s, GB/s streamed).
    mix 0: random REDs only; 1: + one RED per 36 streamed bytes; 2: per 12 bytes; 3: the stream alone."""
    red, rd = C.c_double(0), C.c_double(0)
    check(lib.gcb_selftest_l2(device, window_bytes, mix, stream_bytes, C.byref(red), C.byref(rd)), "gcb_selftest_l2")
    return red.value, rd.value


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    check(lib.gcb_nccl_unique_id(buf), "gcb_nccl_unique_id")
    return buf.raw


ANALYSIS_2D = dict(xyp=("mx", "my"), xzp=("mx", "mz"), yzp=("my", "mz"), hxyp=("mx", "my"), zdf=("mz", 2),
                   lzdf=("mz", 2), profile=("mz", 2), rweight=("nr",), rdf=("nr", 2), hrdf=("nr", 2),
                   rzp=("nr", "mz"), hrzp=("nr", "mz"))


def analyse(tg, grid_zfast, unit="unity", min_occ=0.0, radius=None, z1=None, z2=None, rad_ba_nr=0, rad_ba_step=1,
            device=0, grid_dev=None):
    """a_ri3Dc's reductions (a_ri3Dc.c:488-762) on the GPU -> dict of arrays and scalars.
    grid_dev: device pointer of a z-fastest grid (GridCounter.density_device) instead of the host array."""
    grid = None if grid_dev is not None else np.ascontiguousarray(grid_zfast, np.float32)
    o = K.AnalysisOpts()
    o.unit = K.UNITS[unit]
    o.min_occ = min_occ
    o.rad_ba_nr, o.rad_ba_step = rad_ba_nr, rad_ba_step
    o.device = device
    if radius is not None:
        o.setmask |= K.SET_R
        o.radius = radius
    if z1 is not None:
        o.setmask |= K.SET_Z1
        o.z1 = z1
    if z2 is not None:
        o.setmask |= K.SET_Z2
        o.z2 = z2
    a = K.Analysis()
    if grid_dev is not None:
        check(lib.gcb_analyse_device(C.byref(tg), grid_dev, C.byref(o), C.byref(a)), "gcb_analyse_device")
    else:
        check(lib.gcb_analyse(C.byref(tg), _fptr(grid), C.byref(o), C.byref(a)), "gcb_analyse")
    try:
        dims = dict(mx=a.tg.mx[0], my=a.tg.mx[1], mz=a.tg.mx[2], nr=a.iradNR)
        out = {}
        for name, shp in ANALYSIS_2D.items():
            shape = tuple(dims[s] if isinstance(s, str) else s for s in shp)
            n = int(np.prod(shape))
            p = getattr(a, name)
            out[name] = (np.ctypeslib.as_array(p, shape=(n,)).reshape(shape).copy() if n
                         else np.zeros(shape, np.float32))
        n = a.tg.cells
        # the cropped grid, else the grid that was passed in (None for a device-resident one)
        out["grid"] = np.ctypeslib.as_array(a.grid, shape=(n,)).reshape(a.tg.shape).copy() if a.grid else grid
        tg2 = K.TGrid.from_buffer_copy(a.tg)
        cav = K.Cavity.from_buffer_copy(a.geometry)
        out.update(tg=tg2, geometry=cav, DeltaR=a.DeltaR, dV=a.dV, iradNR=a.iradNR, unit_value=a.unit_value,
                   sumP=a.sumP, sumH=a.sumH, volume=a.volume, reff=a.reff, average=a.average, height=a.height,
                   ivol=a.ivol, kernel_launches=a.kernel_launches)
        return out
    finally:
        lib.gcb_analysis_free(C.byref(a))


def gridcalc(tg, grids_zfast, tweights, device=0):
    """a_gridcalc's time-weighted average (a_gridcalc.c:176-189) -> (avg [mx,my,mz] f32, total weight)"""
    grids = [np.ascontiguousarray(a, np.float32) for a in grids_zfast]
    ptrs = (K.c_float_p * len(grids))(*[_fptr(a) for a in grids])
    tw = np.ascontiguousarray(tweights, np.float32)
    avg = np.zeros(tg.shape, np.float32)
    two = C.c_float(0)
    check(lib.gcb_gridcalc(len(grids), ptrs, _fptr(tw), C.byref(tg), _fptr(avg), C.byref(two), device),
          "gcb_gridcalc")
    return avg, two.value
