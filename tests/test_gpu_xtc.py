"""GPU XTC decoding (gcb_xtc_decode_device, gcb_counter_add_xtc) against the host decoder of host/trajio.c,
bit for bit, and end to end against the oracle on the decoded coordinates."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gcutil as u
import gridcount_b200 as g
from gridcount_b200 import _cabi as K
from test_host_tools_cpu import FrameInfo, water_like, HOST

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", HOST, "libgcbhost.so"])
    L = C.CDLL(os.path.join(HOST, "libgcbhost.so"))
    L.gh_xtc_encode_frame.argtypes = [C.c_int, C.c_int, C.c_float, u.c_float_p, u.c_float_p, C.c_float, u.c_u8_p, C.c_size_t]
    L.gh_xtc_encode_frame.restype = C.c_size_t
    L.gh_xtc_decode_frame.argtypes = [u.c_u8_p, C.c_size_t, C.POINTER(FrameInfo), u.c_float_p, C.c_int]
    L.gh_xtc_decode_frame.restype = C.c_size_t
    return L


def encode_traj(host, x, t, box=(3, 3, 3), prec=1000.0):
    out = []
    b9 = u.box9(box)
    for f in range(len(x)):
        xf = np.ascontiguousarray(x[f], np.float32)
        buf = np.zeros(xf.shape[0] * 16 + 256, np.uint8)
        n = host.gh_xtc_encode_frame(xf.shape[0], f, float(t[f]), u.fptr(b9), u.fptr(xf), prec, buf.ctypes.data_as(u.c_u8_p), len(buf))
        assert n > 0
        out.append(buf[:n].tobytes())
    return b"".join(out)


def host_decode(host, data, natoms, nframes):
    buf = np.frombuffer(data, np.uint8).copy()
    off = 0
    out = np.zeros((nframes, natoms, 3), np.float32)
    info = FrameInfo()
    for f in range(nframes):
        view = buf[off:]
        n = host.gh_xtc_decode_frame(view.ctypes.data_as(u.c_u8_p), len(view), C.byref(info), u.fptr(out[f]), natoms)
        assert n > 0
        off += n
    assert off == len(buf)
    return out


def make_frames(kind, natoms, nframes, seed):
    rng = np.random.default_rng(seed)
    if kind == "water":
        return np.stack([water_like(natoms // 3, 3.0, rng) for _ in range(nframes)])
    if kind == "uniform":
        return (rng.random((nframes, natoms, 3)) * 8.0).astype(np.float32)
    if kind == "negative":
        return ((rng.random((nframes, natoms, 3)) - 0.5) * 20.0).astype(np.float32)
    if kind == "huge":
        return ((rng.random((nframes, natoms, 3)) - 0.5) * 40000.0).astype(np.float32)
    if kind == "clustered":                 # long runs of small differences, smallidx moves both ways
        base = rng.random((nframes, 1, 3)) * 3.0
        steps = rng.normal(0, 0.004, (nframes, natoms, 3)) * (1 + 20 * (rng.random((nframes, natoms, 1)) < 0.02))
        return (base + np.cumsum(steps, axis=1)).astype(np.float32)
    return (rng.random((nframes, natoms, 3)) * 3.0).astype(np.float32)      # "tiny"


@pytest.mark.parametrize("kind,natoms,nframes", [("water", 2700, 40), ("water", 30, 5), ("uniform", 1000, 33), ("negative", 300, 7),
                                                 ("huge", 64, 9), ("clustered", 5000, 12), ("tiny", 9, 6), ("tiny", 2, 3),
                                                 ("water", 60000, 70)])
def test_gpu_decode_equals_host_decode(host, kind, natoms, nframes):
    x = make_frames(kind, natoms, nframes, natoms + nframes)
    t = np.arange(nframes, dtype=np.float32) * 0.5
    data = encode_traj(host, x, t)
    xb = g.XtcBuffer(data)
    assert xb.nframes == nframes and xb.natoms == natoms and xb.consumed == len(data)
    assert np.array_equal(xb.times, t)
    want = host_decode(host, data, natoms, nframes)
    dev = xb.decode_to_device()
    got = dev.download()
    dev.free()
    u.assert_bits_equal(got, want, "GPU-decoded xtc frames")


def test_add_xtc_counts_equal_oracle(host):
    L, natoms, nframes = (3.0, 3.0, 3.0), 2700, 130
    rng = np.random.default_rng(5)
    x = np.stack([np.mod(water_like(natoms // 3, 3.0, rng), 3.0).astype(np.float32) for _ in range(nframes)])
    t = np.arange(nframes, dtype=np.float32) * 2.0
    data = encode_traj(host, x, t)
    xd = host_decode(host, data, natoms, nframes)
    geom = g.setup_geometry(L, 0.05)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    w, T = g.frame_weights(t)
    og = u.Geom.from_buffer_copy(bytes(geom.tg))
    want = np.zeros(geom.shape, np.uint32)
    hits = np.zeros(nframes, np.uint64)
    u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(xd), nframes, natoms, u.iptr(gindex),
                                len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
    for pinned in (False, True):
        xb = g.XtcBuffer(data, pinned=pinned)
        for batch in ("8", "64", None):
            if batch:
                os.environ["GCB_XTC_BATCH"] = batch
            else:
                os.environ.pop("GCB_XTC_BATCH", None)
            ctr = g.GridCounter(geom, gindex, natoms)
            ctr.add_xtc(xb, w)
            assert np.array_equal(ctr.counts(), want)
            assert np.array_equal(ctr.hits_per_frame(), hits)
            ctr.close()
        xb.free()
    os.environ.pop("GCB_XTC_BATCH", None)


def test_damaged_stream_is_reported(host):
    natoms, nframes = 600, 6
    x = make_frames("water", natoms, nframes, 1)
    data = bytearray(encode_traj(host, x, np.arange(nframes, dtype=np.float32)))
    xb0 = g.XtcBuffer(bytes(data))
    fr = xb0.frames[3]
    data[fr.payload_offset - 8:fr.payload_offset - 4] = (70).to_bytes(4, "big")     # smallidx far too large for the stream
    for k in range(fr.payload_offset, fr.payload_offset + 64):
        data[k] = 0xFF
    xb = g.XtcBuffer(bytes(data))
    with pytest.raises(g._cabi.GcbError):
        xb.decode_to_device()
    bad = bytearray(encode_traj(host, x, np.arange(nframes, dtype=np.float32)))
    bad[2] ^= 0x40
    with pytest.raises(g._cabi.GcbError):
        g.XtcBuffer(bytes(bad))


@pytest.mark.parametrize("case", ["stride3", "all", "indexed+duplicates", "pbc", "jittered weights", "tiny frames"])
def test_binning_out_of_the_decoder_equals_the_two_step_path(host, case, monkeypatch):
    """gcb_counter_add_xtc bins a batch straight out of the segment decoder (xtc_bin_kernel) when the batch is one run
    of integer counts; otherwise, and with GCB_XTC_UNFUSED, it decodes to x[frame][atom][3] first and runs the
    binning kernels.  Both must give the oracle's counts, per-frame hits, sumP and occupancy, bit for bit."""
    L, natoms, nframes = (3.0, 3.0, 3.0), 2700, 70
    if case == "tiny frames":
        natoms = 9
    rng = np.random.default_rng(11)
    if case == "pbc":           # atoms outside the box on both sides: the single-image wrap of count.c:88-92 must act
        x = (np.stack([water_like(natoms // 3, 3.0, rng) for _ in range(nframes)]) + rng.normal(0, 0.6, (nframes, 1, 3))).astype(np.float32)
    elif case == "tiny frames":
        x = (rng.random((nframes, natoms, 3)) * 3.0).astype(np.float32)
    else:
        x = np.stack([np.mod(water_like(natoms // 3, 3.0, rng), 3.0).astype(np.float32) for _ in range(nframes)])
    t = (np.arange(nframes + 1, dtype=np.float64) * 0.1 + 100.0).astype(np.float32)[1:]
    data = encode_traj(host, x, t)
    xd = host_decode(host, data, natoms, nframes)
    geom = g.setup_geometry(L, 0.05)
    if case == "indexed+duplicates":
        gindex = np.concatenate([rng.choice(natoms, 900, replace=False), [5, 5, 17]]).astype(np.int32)
    elif case in ("all", "tiny frames"):
        gindex = np.arange(natoms, dtype=np.int32)
    else:
        gindex = np.arange(0, natoms, 3, dtype=np.int32)
    if case == "jittered weights":
        w, T = g.frame_weights(t)              # dt in f32: several distinct weights
        assert len(np.unique(w)) > 1
    else:
        w = np.full(nframes, 0.1, np.float32)
    pbc = K.PBC_RECT if case == "pbc" else K.PBC_NONE
    xo = xd.copy()
    b9 = np.tile(u.box9(L), (nframes, 1)).astype(np.float32)
    og = u.Geom.from_buffer_copy(bytes(geom.tg))
    if case == "pbc":
        u.oracle().gco_wrap_frames(u.fptr(xo), nframes, natoms, u.fptr(b9))
    want = np.zeros(geom.shape, np.uint32)
    hits = np.zeros(nframes, np.uint64)
    u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(xo), nframes, natoms, u.iptr(gindex),
                                len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
    occ = np.zeros(geom.shape, np.float32)
    sp = C.c_double(0)
    u.oracle().gco_bin_frames(C.byref(og), u.fptr(occ), u.fptr(xo), nframes, natoms, u.iptr(gindex), len(gindex), u.fptr(w),
                              C.byref(sp), None, None)
    assert int(hits.sum()) > 0
    xb = g.XtcBuffer(data, pinned=True)
    for unfused in (False, True):
        for batch in ("12", None):
            if unfused:
                monkeypatch.setenv("GCB_XTC_UNFUSED", "1")
            else:
                monkeypatch.delenv("GCB_XTC_UNFUSED", raising=False)
            if batch:
                monkeypatch.setenv("GCB_XTC_BATCH", batch)
            else:
                monkeypatch.delenv("GCB_XTC_BATCH", raising=False)
            ctr = g.GridCounter(geom, gindex, natoms, pbc=pbc)
            ctr.add_xtc(xb, w)
            np.testing.assert_array_equal(ctr.counts(), want)
            np.testing.assert_array_equal(ctr.hits_per_frame(), hits)
            st = ctr.stats()
            assert st.hits == int(hits.sum()) and st.sumP == sp.value
            np.testing.assert_array_equal(ctr.occupancy().view(np.uint32), occ.view(np.uint32))
            ctr.close()
    xb.free()


def test_a_damaged_frame_fails_the_binning_call(host):
    natoms, nframes = 600, 6
    x = make_frames("water", natoms, nframes, 1)
    data = bytearray(encode_traj(host, x, np.arange(nframes, dtype=np.float32)))
    fr = g.XtcBuffer(bytes(data)).frames[3]
    for k in range(fr.payload_offset, fr.payload_offset + 64):
        data[k] = 0xFF
    xb = g.XtcBuffer(bytes(data))
    geom = g.setup_geometry((3.0, 3.0, 3.0), 0.1)
    ctr = g.GridCounter(geom, np.arange(0, natoms, 3, dtype=np.int32), natoms)
    with pytest.raises(g._cabi.GcbError, match="damaged"):
        ctr.add_xtc(xb, np.ones(nframes, np.float32))
    ctr.close()


def test_xtc_into_a_grid_on_the_packed_counter_path(host):
    """a grid whose uint32 cells exceed the L2 (234^3 = 12.8 M cells): gcb_counter_add_xtc decodes to x[frame][atom][3]
    and bins through the packed-counter rounds instead of binning out of the decoder; counts equal the oracle's on the
    decoded coordinates"""
    L, natoms, nframes = (5.85, 5.85, 5.85), 60000, 40
    rng = np.random.default_rng(21)
    x = (rng.random((nframes, natoms, 3)) * np.asarray(L)).astype(np.float32)
    t = np.arange(nframes, dtype=np.float32)
    data = encode_traj(host, x, t)
    xd = host_decode(host, data, natoms, nframes)
    geom = g.setup_geometry(L, 0.025)
    assert geom.cells > 12_000_000
    gindex = np.arange(natoms, dtype=np.int32)
    og = u.Geom.from_buffer_copy(bytes(geom.tg))
    want = np.zeros(geom.shape, np.uint32)
    hits = np.zeros(nframes, np.uint64)
    u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(xd), nframes, natoms, u.iptr(gindex),
                                len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
    xb = g.XtcBuffer(data, pinned=True)
    ctr = g.GridCounter(geom, gindex, natoms)
    ctr.add_xtc(xb, np.ones(nframes, np.float32))
    np.testing.assert_array_equal(ctr.counts(), want)
    np.testing.assert_array_equal(ctr.hits_per_frame(), hits)
    assert ctr.stats().packed_rounds >= 1
    ctr.close(); xb.free()
