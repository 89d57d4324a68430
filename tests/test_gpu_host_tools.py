"""GPU tests of the drop-in command-line tools (host/g_ri3Dc, host/a_ri3Dc, host/a_gridcalc): same command
lines as tests/golden/make_golden.py gave the UNMODIFIED reference binaries, outputs compared with what those
binaries wrote -- grid files byte for byte, text files character for character (xvg time stamps aside)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import gcutil as u
import golden_cases as gc
from test_oracle_golden import ANALYSES

pytestmark = pytest.mark.gpu
HOST = os.path.join(u.ROOT, "host")


@pytest.fixture(scope="module", autouse=True)
def tools():
    subprocess.check_call(["make", "-s", "-C", HOST])


def run_tool(name, args, cwd, stdin=None, env=None):
    p = subprocess.run([os.path.join(HOST, name)] + [str(a) for a in args], cwd=cwd, input=stdin,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600,
                       env=dict(os.environ, **env) if env else None)
    assert p.returncode == 0, "%s failed:\n%s" % (name, p.stderr.decode(errors="replace")[-2000:])
    return p.stdout.decode(errors="replace"), p.stderr.decode(errors="replace")


def write_inputs(name, tmp):
    c = gc.CASES[name]
    t, boxes, x, gindex = gc.case_inputs(name)
    u.write_raw_traj(os.path.join(tmp, "traj.raw"), t, boxes, x)
    u.write_top(os.path.join(tmp, "top.stub"), c["natoms"])
    args = ["-f", "traj.raw", "-s", "top.stub", "-grid", "grid.dat", "-delta"] + list(c["delta"]) + c["gargs"]
    if c["stride"] != 1:
        u.write_ndx(os.path.join(tmp, "grp.ndx"), "OW", gindex)
        args += ["-n", "grp.ndx"]
    return args


@pytest.mark.parametrize("name", list(gc.CASES))
def test_g_ri3Dc_grid_file_equals_reference(name, tmp_path):
    args = write_inputs(name, str(tmp_path))
    run_tool("g_ri3Dc", args, str(tmp_path))
    got = open(tmp_path / "grid.dat", "rb").read()
    want = open(os.path.join(u.GOLDEN_DIR, name, "grid.dat"), "rb").read()
    assert got == want


def body(path):
    """file content without the xvg lines that carry the program name and the time of day"""
    with open(path, "r", errors="replace") as f:
        return [ln for ln in f.read().splitlines() if not ln.startswith("# ")]


@pytest.mark.parametrize("name,run", ANALYSES)
def test_a_ri3Dc_outputs_equal_reference(name, run, tmp_path):
    rdir = os.path.join(u.GOLDEN_DIR, name, run)
    shutil.copy(os.path.join(u.GOLDEN_DIR, name, "grid.dat"), tmp_path / "grid.dat")
    out, err = run_tool("a_ri3Dc", ["-grid", "grid.dat"] + gc.CASES[name]["analyses"][run], str(tmp_path))
    skip = {"stdout.txt", "stderr.txt", "stdout.f64"}
    checked = 0
    for fn in sorted(os.listdir(rdir)):
        if fn in skip or fn.endswith(".f64"):
            continue
        assert os.path.exists(tmp_path / fn), "a_ri3Dc did not write %s" % fn
        if fn in ("plt.dat",):
            assert open(tmp_path / fn, "rb").read() == open(os.path.join(rdir, fn), "rb").read(), fn
        else:
            assert body(tmp_path / fn) == body(os.path.join(rdir, fn)), fn
        checked += 1
    assert checked >= 7
    want = [ln for ln in open(os.path.join(rdir, "stdout.txt"), errors="replace").read().splitlines() if ln.startswith("DATA_RESULT")]
    got = [ln for ln in out.splitlines() if ln.startswith("DATA_RESULT")]
    assert got == want


def test_a_gridcalc_equals_reference(tmp_path):
    ins = [os.path.join(u.GOLDEN_DIR, n, "grid.dat") for n in gc.GRIDCALC["inputs"]]
    run_tool("a_gridcalc", ["-f"] + ins + ["-o", "avg.dat"], str(tmp_path))
    assert open(tmp_path / "avg.dat", "rb").read() == open(os.path.join(u.GOLDEN_DIR, "gridcalc", "avg.dat"), "rb").read()


def test_g_ri3Dc_from_xtc_counts_decoded_coordinates(tmp_path):
    """raw -> xtc (own codec) -> g_ri3Dc -f traj.xtc: the density equals the oracle's on the DECODED coordinates;
    two index groups in the file exercise the interactive group selection"""
    import gridcount_b200 as g
    name = "pore"
    c = gc.CASES[name]
    t, boxes, x, gindex = gc.case_inputs(name)
    tmp = str(tmp_path)
    u.write_raw_traj(os.path.join(tmp, "traj.raw"), t, boxes, x)
    subprocess.check_call([os.path.join(HOST, "gcb_xtc"), "convert", "traj.raw", "traj.xtc"], cwd=tmp, stdout=subprocess.DEVNULL)
    with open(os.path.join(tmp, "grp.ndx"), "w") as f:
        f.write("[ System ]\n" + " ".join(str(i + 1) for i in range(c["natoms"])) + "\n")
        f.write("[ OW ]\n" + "\n".join(" ".join(str(i + 1) for i in gindex[k:k + 15]) for k in range(0, len(gindex), 15)) + "\n")
    run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", "grid.dat", "-delta", "0.2", "-R", "1.5", "-z1", "1.0",
                         "-z2", "5.5", "-cpoint", "2.05", "1.95", "3.0"], tmp, stdin=b"1\n")
    tg, hdr, dens = g.grid_read(os.path.join(tmp, "grid.dat"))
    assert b"{grp}OW" in hdr and b"{XTC}traj.xtc" in hdr
    # frames decoded on the host instead of on the GPU: the same file
    run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", "grid_host.dat", "-delta", "0.2", "-R", "1.5", "-z1", "1.0",
                         "-z2", "5.5", "-cpoint", "2.05", "1.95", "3.0", "-hostdecode"], tmp, stdin=b"OW\n")
    assert open(os.path.join(tmp, "grid.dat"), "rb").read() == open(os.path.join(tmp, "grid_host.dat"), "rb").read()
    # the file read in small chunks (a frame cut by a chunk's end is carried over; weights and T_tot carry across calls)
    fsize = os.path.getsize(os.path.join(tmp, "traj.xtc"))
    run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", "grid_chunks.dat", "-delta", "0.2", "-R", "1.5", "-z1", "1.0",
                         "-z2", "5.5", "-cpoint", "2.05", "1.95", "3.0"], tmp, stdin=b"1\n",
             env={"GCB_XTC_CHUNK_BYTES": str(max(1, fsize // 7))})
    assert open(os.path.join(tmp, "grid.dat"), "rb").read() == open(os.path.join(tmp, "grid_chunks.dat"), "rb").read()
    # the frames sharded over "two GPUs" (twice this one: the in-process reduce accepts that), each shard read by its
    # own thread in small chunks.  Without -dtweight every frame weighs 1: bit-identical to the single-GPU file.
    run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", "one.dat", "-delta", "0.2", "-nodtweight"], tmp, stdin=b"1\n")
    import gridcount_b200 as g2
    real = [("0,1", "two_real.dat")] if g2.device_count() >= 2 else []          # two different devices where the box has them
    for gpus, out in [("0,0", "two.dat"), ("0,0,0", "three.dat")] + real:
        so, se = run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", out, "-delta", "0.2", "-nodtweight", "-gpus", gpus],
                          tmp, stdin=b"1\n", env={"GCB_XTC_CHUNK_BYTES": str(max(1, fsize // 5))})
        assert "Sharding" in se
        assert open(os.path.join(tmp, "one.dat"), "rb").read() == open(os.path.join(tmp, out), "rb").read()
    # with the time step as weight (f32-jittered dt) the shards' f32 accumulators are summed: equal to 1e-6
    run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", "two_dt.dat", "-delta", "0.2", "-R", "1.5", "-z1", "1.0",
                         "-z2", "5.5", "-cpoint", "2.05", "1.95", "3.0", "-gpus", "0,0"], tmp, stdin=b"1\n")
    tg2, hdr2, dens2 = g.grid_read(os.path.join(tmp, "two_dt.dat"))
    np.testing.assert_allclose(dens2, dens, rtol=1e-6, atol=0)
    # -b/-e select the same frames on both paths
    for extra, out in ((["-b", "10", "-e", "30"], "sel_gpu.dat"), (["-b", "10", "-e", "30", "-hostdecode"], "sel_host.dat")):
        run_tool("g_ri3Dc", ["-f", "traj.xtc", "-n", "grp.ndx", "-grid", out, "-delta", "0.2"] + extra, tmp, stdin=b"1\n")
    assert open(os.path.join(tmp, "sel_gpu.dat"), "rb").read() == open(os.path.join(tmp, "sel_host.dat"), "rb").read()
    # decoded coordinates through the same reader
    from test_host_tools_cpu import FrameInfo
    L = C.CDLL(os.path.join(HOST, "libgcbhost.so"))
    L.gh_traj_open.argtypes = [C.c_char_p]; L.gh_traj_open.restype = C.c_void_p
    L.gh_traj_next.argtypes = [C.c_void_p, C.POINTER(FrameInfo), u.c_float_p]
    L.gh_traj_close.argtypes = [C.c_void_p]
    h = L.gh_traj_open(os.path.join(tmp, "traj.xtc").encode())
    xd = np.zeros_like(x)
    info = FrameInfo()
    for f in range(len(x)):
        buf = np.zeros((c["natoms"], 3), np.float32)
        assert L.gh_traj_next(h, C.byref(info), u.fptr(buf)) == 1
        xd[f] = buf
    L.gh_traj_close(h)
    o = u.oracle()
    og, cav = u.oracle_geometry(c["L"], (0.2, 0.2, 0.2), u.SET_R | u.SET_Z1 | u.SET_Z2 | u.SET_CPOINT, (2.05, 1.95, 3.0), 1.5, 1.0, 5.5)
    w = np.zeros(len(t), np.float32)
    T = C.c_double(0)
    o.gco_frame_weights(u.fptr(t), len(t), 1, u.fptr(w), C.byref(T))
    occ = np.zeros(og.shape(), np.float32)
    sp = C.c_double(0)
    o.gco_bin_frames(C.byref(og), u.fptr(occ), u.fptr(xd), len(t), c["natoms"], u.iptr(gindex), len(gindex), u.fptr(w),
                     C.byref(sp), None, None)
    o.gco_normalise(C.byref(og), u.fptr(occ), T.value)
    u.assert_bits_equal(dens, occ, "density from the XTC trajectory")
