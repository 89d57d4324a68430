"""CPU tests of the plain-C host layer (host/libgcbhost.so): the XTC codec (round trips and the
quantisation invariants -- parity with Gromacs' own codec is unpinned, see host/trajio.c), the raw and
.gro readers and the command-line conventions of the drop-in tools."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gcutil as u

ROOT = u.ROOT
HOST = os.path.join(ROOT, "host")


class FrameInfo(C.Structure):
    _fields_ = [("natoms", C.c_int), ("step", C.c_int), ("time", C.c_float), ("box", C.c_float * 9), ("prec", C.c_float)]


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-s", "-C", HOST, "libgcbhost.so", "gcb_xtc"])
    L = C.CDLL(os.path.join(HOST, "libgcbhost.so"))
    L.gh_xtc_encode_frame.argtypes = [C.c_int, C.c_int, C.c_float, u.c_float_p, u.c_float_p, C.c_float, u.c_u8_p, C.c_size_t]
    L.gh_xtc_encode_frame.restype = C.c_size_t
    L.gh_xtc_decode_frame.argtypes = [u.c_u8_p, C.c_size_t, C.POINTER(FrameInfo), u.c_float_p, C.c_int]
    L.gh_xtc_decode_frame.restype = C.c_size_t
    L.gh_traj_open.argtypes = [C.c_char_p]
    L.gh_traj_open.restype = C.c_void_p
    L.gh_traj_natoms.argtypes = [C.c_void_p]
    L.gh_traj_next.argtypes = [C.c_void_p, C.POINTER(FrameInfo), u.c_float_p]
    L.gh_traj_close.argtypes = [C.c_void_p]
    return L


def encode(host, x, step=7, time=1.5, box=(3, 3, 3), prec=1000.0):
    x = np.ascontiguousarray(x, np.float32)
    b9 = u.box9(box)
    buf = np.zeros(len(x) * 16 + 256, np.uint8)
    n = host.gh_xtc_encode_frame(len(x), step, time, u.fptr(b9), u.fptr(x), prec, buf.ctypes.data_as(u.c_u8_p), len(buf))
    assert n > 0
    return buf[:n].copy()


def decode(host, buf, natoms):
    info = FrameInfo()
    x = np.zeros((natoms, 3), np.float32)
    n = host.gh_xtc_decode_frame(buf.ctypes.data_as(u.c_u8_p), len(buf), C.byref(info), u.fptr(x), natoms)
    assert n == len(buf), "decoder consumed %d of %d bytes" % (n, len(buf))
    return info, x


def water_like(n_mol, L, rng):
    """3-site molecules: O uniform in the box, H within 0.1 nm: exercises the small-difference runs and the swap"""
    o = rng.random((n_mol, 3)) * L
    h1 = o + rng.normal(0, 0.05, (n_mol, 3))
    h2 = o + rng.normal(0, 0.05, (n_mol, 3))
    return np.stack([o, h1, h2], axis=1).reshape(-1, 3).astype(np.float32)


@pytest.mark.parametrize("kind,natoms", [("water", 2700), ("water", 30), ("uniform", 1000), ("uniform", 10), ("line", 500),
                                         ("tiny", 9), ("tiny", 1), ("negative", 300), ("huge", 64)])
def test_xtc_round_trip(host, kind, natoms):
    rng = np.random.default_rng(natoms)
    if kind == "water":
        x = water_like(natoms // 3, 3.0, rng)
    elif kind == "uniform":
        x = (rng.random((natoms, 3)) * 8.0).astype(np.float32)
    elif kind == "line":
        x = np.stack([np.arange(natoms) * 0.001, np.zeros(natoms), np.full(natoms, 2.5)], axis=1).astype(np.float32)
    elif kind == "tiny":
        x = (rng.random((natoms, 3)) * 3.0).astype(np.float32)
    elif kind == "negative":
        x = ((rng.random((natoms, 3)) - 0.5) * 20.0).astype(np.float32)
    else:                                   # extents beyond 2^24 grid units: components are sent separately
        x = ((rng.random((natoms, 3)) - 0.5) * 40000.0).astype(np.float32)
    buf = encode(host, x)
    info, y = decode(host, buf, len(x))
    assert info.natoms == len(x) and info.step == 7 and info.time == np.float32(1.5)
    assert list(info.box) == list(u.box9((3, 3, 3)))
    if len(x) <= 9:
        u.assert_bits_equal(y, x, "uncompressed small frame")
        return
    # the decoded value is round(x * prec) * fl(1/prec), in f32, exactly
    lf = x * np.float32(1000.0)
    q = np.where(lf >= 0, (lf + np.float32(0.5)).astype(np.int32), (lf - np.float32(0.5)).astype(np.int32))
    want = q.astype(np.float32) * (np.float32(1.0) / np.float32(1000.0))
    u.assert_bits_equal(y, want, "decoded coordinates")
    if kind == "huge":
        return                              # q * fl(1/prec) * prec does not round back to q beyond ~2^21 grid units
    # and the codec is idempotent on its own output
    buf2 = encode(host, y)
    info2, y2 = decode(host, buf2, len(x))
    u.assert_bits_equal(y2, y, "second generation")


def test_xtc_compresses_water(host):
    rng = np.random.default_rng(1)
    x = water_like(3000, 4.0, rng)
    buf = encode(host, x)
    assert len(buf) < 0.45 * x.nbytes, "no compression: %d of %d bytes" % (len(buf), x.nbytes)


def test_xtc_rejects_damage(host):
    rng = np.random.default_rng(2)
    x = water_like(100, 3.0, rng)
    buf = encode(host, x)
    info = FrameInfo()
    y = np.zeros((300, 3), np.float32)
    bad = buf.copy(); bad[3] ^= 0xFF                                     # magic
    assert host.gh_xtc_decode_frame(bad.ctypes.data_as(u.c_u8_p), len(bad), C.byref(info), u.fptr(y), 300) == 0
    short = buf[:len(buf) - 8].copy()                                    # truncated
    assert host.gh_xtc_decode_frame(short.ctypes.data_as(u.c_u8_p), len(short), C.byref(info), u.fptr(y), 300) == 0
    # a crafted header with maxint = minint - 1 makes the range wrap to zero: must be refused, not divide by zero
    for d in range(3):
        wrap = buf.copy()
        lo = int.from_bytes(bytes(wrap[60 + 4 * d:64 + 4 * d]), "big", signed=True)
        wrap[72 + 4 * d:76 + 4 * d] = np.frombuffer((lo - 1).to_bytes(4, "big", signed=True), np.uint8)
        assert host.gh_xtc_decode_frame(wrap.ctypes.data_as(u.c_u8_p), len(wrap), C.byref(info), u.fptr(y), 300) == 0


def read_all(host, path):
    t = host.gh_traj_open(path.encode())
    assert t
    n = host.gh_traj_natoms(t)
    frames, times = [], []
    info = FrameInfo()
    x = np.zeros((n, 3), np.float32)
    while host.gh_traj_next(t, C.byref(info), u.fptr(x)) == 1:
        frames.append(x.copy()); times.append(info.time)
    host.gh_traj_close(t)
    return np.array(times, np.float32), np.array(frames)


def test_raw_to_xtc_file_round_trip(host, tmp_path):
    rng = np.random.default_rng(3)
    F, natoms = 12, 600
    x = np.stack([water_like(natoms // 3, 3.0, rng) for _ in range(F)])
    t = (np.arange(F) * 0.5).astype(np.float32)
    boxes = np.tile(u.box9((3, 3, 3)), (F, 1))
    raw = str(tmp_path / "traj.raw")
    u.write_raw_traj(raw, t, boxes, x)
    tr, xr = read_all(host, raw)
    u.assert_bits_equal(xr, x, "raw reader")
    xtc = str(tmp_path / "traj.xtc")
    out = subprocess.check_output([os.path.join(HOST, "gcb_xtc"), "convert", raw, xtc]).decode()
    assert out.split()[0] == str(F)
    tx, xx = read_all(host, xtc)
    assert np.array_equal(tx, t)
    assert np.abs(xx - x).max() <= 0.0005 + 1e-6
    assert "12 frames" in subprocess.check_output([os.path.join(HOST, "gcb_xtc"), "info", xtc]).decode()


def test_gro_reader(host, tmp_path):
    p = tmp_path / "conf.gro"
    with open(p, "w") as f:
        for fr in range(2):
            f.write("water t= %8.3f\n    3\n" % (fr * 2.0))
            for i in range(3):
                f.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (1, "SOL", "OW", i + 1, 0.1 * i + fr, 1.234, -0.5))
            f.write("   3.00000   3.10000   3.20000\n")
    t, x = read_all(host, str(p))
    assert list(t) == [0.0, 2.0] and x.shape == (2, 3, 3)
    assert x[1, 2, 0] == np.float32(1.2) and x[0, 0, 2] == np.float32(-0.5)


@pytest.mark.parametrize("tool", ["g_ri3Dc", "a_ri3Dc", "a_gridcalc"])
def test_tools_link_and_follow_the_option_conventions(tool, tmp_path):
    """-h works without a GPU; unknown options and missing values are fatal (exit code 1, message on stderr)"""
    subprocess.check_call(["make", "-s", "-C", HOST])
    exe = os.path.join(HOST, tool)
    out = subprocess.check_output([exe, "-h"]).decode()
    assert tool in out and "Options:" in out
    p = subprocess.run([exe, "-nonsense"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=str(tmp_path))
    assert p.returncode == 1 and b"Invalid command line argument" in p.stderr
    if tool == "g_ri3Dc":
        p = subprocess.run([exe, "-delta"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=str(tmp_path))
        assert p.returncode == 1 and b"needs a value" in p.stderr
        p = subprocess.run([exe, "-m"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=str(tmp_path))
        assert p.returncode == 1 and b"-m option not working" in p.stderr          # g_ri3Dc.c:357-358


def test_xtc_scan_reads_the_frame_table(host):
    """gcb_xtc_scan (product library, host code): offsets, sizes and header fields of every frame; an
    incomplete trailing frame is left for the next read; a wrong magic number is an error"""
    import gridcount_b200 as g
    from gridcount_b200 import _cabi as K
    rng = np.random.default_rng(9)
    parts, sizes = [], []
    for f, n in enumerate((300, 300, 300)):
        buf = encode(host, water_like(n // 3, 3.0, rng), step=f, time=0.25 * f, box=(3, 4, 5))
        parts.append(buf.tobytes()); sizes.append(len(buf))
    tiny = encode(host, (rng.random((4, 3)) * 3).astype(np.float32), step=9, time=9.0)
    data = b"".join(parts)
    xb = g.XtcBuffer(data + data[:100])            # a truncated fourth frame
    assert xb.nframes == 3 and xb.consumed == len(data) and xb.natoms == 300
    off = 0
    for f in range(3):
        fr = xb.frames[f]
        assert fr.payload_offset == off + 92 and fr.step == f and fr.time == np.float32(0.25 * f) and fr.precision == 1000.0
        assert (fr.payload_bytes + 3) // 4 * 4 == sizes[f] - 92 and list(fr.box)[::4] == [3.0, 4.0, 5.0]
        assert all(fr.minint[d] <= fr.maxint[d] for d in range(3)) and 9 <= fr.smallidx <= 72
        off += sizes[f]
    small = g.XtcBuffer(tiny.tobytes())
    assert small.nframes == 1 and small.frames[0].payload_bytes == 0 and small.frames[0].natoms == 4
    bad = bytearray(data); bad[1] = 7
    with pytest.raises(K.GcbError):
        g.XtcBuffer(bytes(bad))
