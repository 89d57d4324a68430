#!/usr/bin/env python
"""Wall time of gcb_analyse (K3) on BASELINE-sized grids, with the oracle's CPU time beside it.
Usage: python tools/perf_k3.py [c2|c3|c1]"""
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

SHAPES = {"c1": ((3.0, 3.0, 3.0), 0.05), "c2": ((8.0, 8.0, 13.4), 0.1), "c3": ((20.0, 20.0, 20.0), 0.05)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    L, delta = SHAPES[name]
    geom = g.setup_geometry(L, delta)
    rng = np.random.default_rng(0)
    dens = (rng.poisson(3.0, geom.shape) * 0.37).astype(np.float32)
    tg = geom.tg
    g.analyse(tg, dens, unit="SPC")                       # warm-up (context, allocations)
    t0 = time.perf_counter()
    r = g.analyse(tg, dens, unit="SPC")
    gpu_s = time.perf_counter() - t0
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        g.analyse(tg, dens, unit="SPC")
        ts.append(time.perf_counter() - t0)
    dev = g.DeviceFrames(1, dens.size // 3 + 1)                 # any device buffer of the grid's size
    g._cabi.check(g._cabi.lib.gcb_memcpy_h2d(dev.ptr, dens.ctypes.data, dens.nbytes, 0), "h2d")
    g.analyse(tg, None, unit="SPC", grid_dev=dev.ptr)
    td = []
    for _ in range(5):
        t0 = time.perf_counter()
        g.analyse(tg, None, unit="SPC", grid_dev=dev.ptr)
        td.append(time.perf_counter() - t0)
    print("%s: host grid in %.2f ms (median of 5), grid already in HBM %.2f ms" % (name, 1e3 * sorted(ts)[2], 1e3 * sorted(td)[2]))
    og = u.Geom.from_buffer_copy(bytes(tg))
    u.oracle().gco_update_b(C.byref(og))
    an = u.Analysis()
    an.unit = u.UNITS["SPC"]; an.rad_ba_step = 1
    t0 = time.perf_counter()
    u.oracle().gco_analyse(C.byref(og), u.fptr(dens), C.byref(an))
    cpu_s = time.perf_counter() - t0
    same = np.array_equal(u.analysis_arrays(an, og)["rzp"].view(np.uint32), r["rzp"].view(np.uint32))
    u.oracle().gco_analysis_free(C.byref(an))
    print("%s grid %s: gcb_analyse %.1f ms (incl. H2D of the grid and D2H of all outputs), oracle on one host thread %.1f ms, rzp bit-equal: %s"
          % (name, geom.shape, gpu_s * 1e3, cpu_s * 1e3, same))


if __name__ == "__main__":
    main()
