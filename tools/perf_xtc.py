#!/usr/bin/env python
"""Decode throughput of the GPU XTC decoder vs the host decoder (development aid).
Usage: python tools/perf_xtc.py [water|uniform] [natoms] [frames] [replicas]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gcutil as u  # noqa: E402
import gridcount_b200 as g  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "water"
    natoms = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    nf = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    rep = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    H = C.CDLL(os.path.join(ROOT, "host", "libgcbhost.so"))
    H.gh_xtc_encode_frame.argtypes = [C.c_int, C.c_int, C.c_float, u.c_float_p, u.c_float_p, C.c_float, u.c_u8_p, C.c_size_t]
    H.gh_xtc_encode_frame.restype = C.c_size_t
    rng = np.random.default_rng(0)
    L = 8.0
    parts = []
    buf = np.zeros(natoms * 16 + 256, np.uint8)
    b9 = u.box9((L, L, L))
    for f in range(nf):
        if kind == "water":
            o = rng.random((natoms // 3, 3)) * L
            x = np.stack([o, o + rng.normal(0, 0.05, o.shape), o + rng.normal(0, 0.05, o.shape)], axis=1).reshape(-1, 3)
        else:
            x = rng.random((natoms, 3)) * L
        x = np.ascontiguousarray(x, np.float32)
        n = H.gh_xtc_encode_frame(len(x), f, float(f), u.fptr(b9), u.fptr(x), 1000.0, buf.ctypes.data_as(u.c_u8_p), len(buf))
        parts.append(buf[:n].tobytes())
    data = b"".join(parts)
    natoms = len(x)
    xb = g.XtcBuffer(data * rep, pinned=True)
    print("%s: %d atoms, %d frames, %.2f bytes/atom" % (kind, natoms, xb.nframes, len(data) / (nf * natoms)))
    for _ in range(2):
        t0 = time.perf_counter()
        d = xb.decode_to_device()
        dt = time.perf_counter() - t0
        d.free()
    print("gpu decode incl. H2D+alloc: %.1f ms, %.2f G atoms/s" % (dt * 1e3, xb.nframes * natoms / dt * 1e-9))
    geom = g.setup_geometry((L, L, L), 0.1)
    ctr = g.GridCounter(geom, np.arange(0, natoms, 3, dtype=np.int32), natoms)
    w = np.ones(xb.nframes, np.float32)
    for _ in range(2):
        ctr.reset_grid()
        t0 = time.perf_counter()
        ctr.add_xtc(xb, w)
        ctr.sync()
        dt = time.perf_counter() - t0
    print("add_xtc (H2D + decode + bin, pipelined): %.1f ms, %.2f G atoms/s decoded, %.2f GB/s compressed"
          % (dt * 1e3, xb.nframes * natoms / dt * 1e-9, xb.nbytes / dt * 1e-9))


if __name__ == "__main__":
    main()
