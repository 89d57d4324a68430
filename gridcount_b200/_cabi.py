"""ctypes binding of libgridcount_b200.so (include/gridcount_b200.h).

The library is built in-tree by `make -C gridcount_b200/csrc` (see
__graft_entry__.build).  There is no fallback: if the shared object is
missing, importing this module raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCB_LIB") or os.path.join(HERE, "libgridcount_b200.so")   # GCB_LIB: tools/sweep_k1.sh variants

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_u8_p = C.POINTER(C.c_uint8)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)
c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)

HEADER_MAX = 256
SET_CPOINT, SET_R, SET_Z1, SET_Z2 = 1, 2, 4, 8
PBC_NONE, PBC_RECT = 0, 1
HOST_PACK_AUTO, HOST_PACK_NEVER, HOST_PACK_ALWAYS = -1, 0, 1
KERNEL_AUTO, KERNEL_GATHER, KERNEL_STREAM, KERNEL_STREAM_INDEXED, KERNEL_PACKED, KERNEL_PRIVATE = 0, 1, 2, 3, 5, 6
UNITS = {"unity": 0, "SPC": 1, "molar": 2, "Angstrom": 3, "Voxel": 4}
GEN_UNIFORM, GEN_QUANTISED, GEN_PORE, GEN_CLUSTER = 0, 1, 2, 3


class Cavity(C.Structure):
    _fields_ = [("axis", C.c_float * 3), ("cpoint", C.c_float * 3), ("radius", C.c_float),
                ("z1", C.c_float), ("z2", C.c_float), ("vol", C.c_float)]


class TGrid(C.Structure):
    _fields_ = [("mx", C.c_int * 3), ("a", C.c_float * 3), ("b", C.c_float * 3),
                ("delta", C.c_float * 3), ("tweight", C.c_float)]

    @property
    def shape(self):
        return (self.mx[0], self.mx[1], self.mx[2])

    @property
    def cells(self):
        return self.mx[0] * self.mx[1] * self.mx[2]


class CounterOpts(C.Structure):
    _fields_ = [("device", C.c_int), ("pbc", C.c_int), ("kernel", C.c_int), ("nslots", C.c_int),
                ("frames_per_batch", C.c_long)]


class Stats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("atom_frames", C.c_uint64), ("hits", C.c_uint64),
                ("edge_overflow", C.c_uint64), ("sumP", C.c_double), ("uniform_weight", C.c_int),
                ("weight", C.c_float), ("kernel_launches", C.c_int64), ("packed_rounds", C.c_uint64),
                ("packed_redone", C.c_uint64), ("saturated", C.c_int),
                ("private_launches", C.c_uint64), ("private_redone", C.c_uint64), ("packed_bits", C.c_int)]


class HostStaging(C.Structure):
    _fields_ = [("mode", C.c_int), ("threads", C.c_int), ("h2d_bytes", C.c_uint64), ("frames_packed", C.c_uint64),
                ("frames_in_place", C.c_uint64), ("pack_bytes_per_s", C.c_double), ("lanes", C.c_int),
                ("two_lane_frames_per_s", C.c_double), ("in_place_frames_per_s", C.c_double)]


class AnalysisOpts(C.Structure):
    _fields_ = [("unit", C.c_int), ("min_occ", C.c_float), ("rad_ba_nr", C.c_int), ("rad_ba_step", C.c_int),
                ("setmask", C.c_int), ("radius", C.c_float), ("z1", C.c_float), ("z2", C.c_float),
                ("device", C.c_int)]


class Analysis(C.Structure):
    _fields_ = [("tg", TGrid), ("geometry", Cavity), ("DeltaR", C.c_float), ("dV", C.c_double),
                ("iradNR", C.c_int), ("unit_value", C.c_float), ("sumP", C.c_double), ("sumH", C.c_double),
                ("volume", C.c_float), ("reff", C.c_float), ("average", C.c_float), ("height", C.c_float),
                ("ivol", C.c_int), ("grid", c_float_p),
                ("xyp", c_float_p), ("xzp", c_float_p), ("yzp", c_float_p), ("hxyp", c_float_p),
                ("zdf", c_float_p), ("lzdf", c_float_p), ("profile", c_float_p), ("rweight", c_float_p),
                ("rdf", c_float_p), ("hrdf", c_float_p), ("rzp", c_float_p), ("hrzp", c_float_p),
                ("kernel_launches", C.c_int64)]


class XtcFrame(C.Structure):
    _fields_ = [("payload_offset", C.c_uint64), ("payload_bytes", C.c_uint32), ("natoms", C.c_int), ("step", C.c_int),
                ("time", C.c_float), ("box", C.c_float * 9), ("precision", C.c_float),
                ("minint", C.c_int * 3), ("maxint", C.c_int * 3), ("smallidx", C.c_int)]


class Gen(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mode", C.c_int), ("L", C.c_float * 3),
                ("slab_lo", C.c_float), ("slab_hi", C.c_float), ("pore_r", C.c_float),
                ("nsites", C.c_int), ("sigma", C.c_float), ("box_period", C.c_int), ("box_amp", C.c_float)]


# every symbol include/gridcount_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gcb_last_error": (C.c_char_p, []),
    "gcb_abi_version": (C.c_int, []),
    "gcb_autoset_cavity": (C.c_int, [C.POINTER(Cavity), c_float_p, C.c_int]),
    "gcb_setup_tgrid": (C.c_int, [C.POINTER(TGrid), C.POINTER(Cavity), c_float_p]),
    "gcb_delta_v": (C.c_double, [C.POINTER(TGrid)]),
    "gcb_update_tgrid": (None, [C.POINTER(TGrid)]),
    "gcb_tgrid2cavity": (None, [C.POINTER(TGrid), C.POINTER(Cavity)]),
    "gcb_edge_vulnerable": (C.c_int, [C.POINTER(TGrid)]),
    "gcb_frame_weights": (C.c_int, [c_float_p, C.c_long, C.c_int, c_float_p, c_float_p, c_double_p]),
    "gcb_counter_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(TGrid), c_int_p, C.c_int, C.c_int,
                                     C.POINTER(CounterOpts)]),
    "gcb_counter_destroy": (None, [C.c_void_p]),
    "gcb_counter_reset": (C.c_int, [C.c_void_p]),
    "gcb_counter_add_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, c_float_p, c_float_p]),
    "gcb_counter_set_host_pack": (C.c_int, [C.c_void_p, C.c_int]),
    "gcb_counter_host_staging": (C.c_int, [C.c_void_p, C.POINTER(HostStaging)]),
    "gcb_pack_frames": (C.c_int, [c_float_p, C.c_long, C.c_int, c_int_p, C.c_int, c_float_p, C.c_int]),
    "gcb_counter_add_device_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_long, c_float_p, c_float_p]),
    "gcb_counter_slot": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(c_float_p), c_long_p]),
    "gcb_counter_submit_slot": (C.c_int, [C.c_void_p, C.c_int, C.c_long, c_float_p, c_float_p]),
    "gcb_counter_sync": (C.c_int, [C.c_void_p]),
    "gcb_counter_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "gcb_counter_hits_per_frame": (C.c_int, [C.c_void_p, c_u64_p, C.c_long, c_long_p]),
    "gcb_counter_counts": (C.c_int, [C.c_void_p, c_u32_p]),
    "gcb_counter_occupancy": (C.c_int, [C.c_void_p, c_float_p]),
    "gcb_counter_density": (C.c_int, [C.c_void_p, C.c_double, c_float_p, c_float_p, C.POINTER(TGrid)]),
    "gcb_counter_device_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "gcb_counter_device_density": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_void_p)]),
    "gcb_counter_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "gcb_counter_event_elapsed": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_float_p]),
    "gcb_counter_timer_start": (C.c_int, [C.c_void_p]),
    "gcb_counter_timer_stop": (C.c_int, [C.c_void_p, c_float_p]),
    "gcb_counter_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "gcb_counter_profile_read": (C.c_int, [C.c_void_p, c_double_p, c_long_p, c_double_p]),
    "gcb_counter_profile_kinds": (C.c_int, [C.c_void_p, c_double_p, c_long_p, c_double_p]),
    "gcb_selftest_l2": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_size_t, c_double_p, c_double_p]),
    "gcb_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "gcb_counter_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "gcb_counter_allreduce": (C.c_int, [C.c_void_p]),
    "gcb_ctl_open": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gcb_ctl_allgather": (C.c_int, [C.c_void_p, c_u64_p, c_u64_p, C.c_int]),
    "gcb_ctl_close": (None, [C.c_void_p]),
    "gcb_local_comm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "gcb_counter_comm_init_local": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "gcb_local_comm_destroy": (None, [C.c_void_p]),
    "gcb_xtc_scan": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(XtcFrame), C.c_long, c_long_p, C.POINTER(C.c_size_t)]),
    "gcb_xtc_decode_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(XtcFrame), C.c_long, C.c_int, C.c_void_p, C.c_int]),
    "gcb_counter_add_xtc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(XtcFrame), C.c_long, c_float_p]),
    "gcb_header": (C.c_int, [C.c_char_p, C.c_char_p, C.c_double, C.c_char_p, C.POINTER(Cavity), C.POINTER(TGrid),
                             C.c_double, C.c_char_p, C.c_int]),
    "gcb_xdr_size": (C.c_size_t, [C.POINTER(TGrid), C.c_char_p]),
    "gcb_xdr_encode": (C.c_int, [C.POINTER(TGrid), c_float_p, c_float_p, C.c_char_p, c_u8_p, C.c_size_t,
                                 C.POINTER(C.c_size_t)]),
    "gcb_grid_write": (C.c_int, [C.c_char_p, C.POINTER(TGrid), c_float_p, c_float_p, C.c_char_p]),
    "gcb_counter_xdr": (C.c_int, [C.c_void_p, C.c_double, C.c_char_p, c_u8_p, C.c_size_t, C.POINTER(C.c_size_t),
                                  C.POINTER(TGrid)]),
    "gcb_counter_grid_write": (C.c_int, [C.c_void_p, C.c_double, C.c_char_p, C.c_char_p, C.POINTER(TGrid)]),
    "gcb_counter_plt": (C.c_int, [C.c_void_p, C.c_double, C.c_float, c_u8_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gcb_grid_read": (C.c_int, [C.c_char_p, C.POINTER(TGrid), C.c_char_p, C.POINTER(c_float_p)]),
    "gcb_plt_write": (C.c_int, [C.c_char_p, C.POINTER(TGrid), c_float_p, C.c_float]),
    "gcb_ascii_write": (C.c_int, [C.c_char_p, C.POINTER(TGrid), c_float_p, C.c_float]),
    "gcb_free": (None, [C.c_void_p]),
    "gcb_dens_unit": (C.c_float, [C.c_int, C.POINTER(TGrid)]),
    "gcb_analyse": (C.c_int, [C.POINTER(TGrid), c_float_p, C.POINTER(AnalysisOpts), C.POINTER(Analysis)]),
    "gcb_analyse_device": (C.c_int, [C.POINTER(TGrid), C.c_void_p, C.POINTER(AnalysisOpts), C.POINTER(Analysis)]),
    "gcb_analysis_free": (None, [C.POINTER(Analysis)]),
    "gcb_analyse_release_cache": (None, [C.c_int]),
    "gcb_gridcalc": (C.c_int, [C.c_int, C.POINTER(c_float_p), c_float_p, C.POINTER(TGrid), c_float_p, c_float_p,
                               C.c_int]),
    "gcb_gen_device_frames": (C.c_int, [C.POINTER(Gen), C.c_long, C.c_long, C.c_int, C.c_void_p, C.c_int]),
    "gcb_device_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "gcb_device_free": (C.c_int, [C.c_void_p, C.c_int]),
    "gcb_host_alloc_pinned": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "gcb_host_free_pinned": (C.c_int, [C.c_void_p]),
    "gcb_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "gcb_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "gcb_device_count": (C.c_int, []),
}

if not os.path.exists(LIB_PATH):
    raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                      "`make -C gridcount_b200/csrc`; there is no CPU fallback" % LIB_PATH)

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class GcbError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.gcb_last_error()
        super().__init__("%s failed (%d): %s" % (where, code, msg.decode(errors="replace") if msg else ""))


def check(rc, where):
    if rc != 0:
        raise GcbError(rc, where)
