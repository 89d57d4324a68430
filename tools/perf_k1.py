#!/usr/bin/env python
"""Quick device-resident timing of the binning kernels on a named shape (development aid;
bench.py is the contract).  Usage: python tools/perf_k1.py [c1|c2|c2all|c3|c4] [frames]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridcount_b200 as g  # noqa: E402
from gridcount_b200 import _cabi as K  # noqa: E402

SHAPES = {
    "c1": ((3.0, 3.0, 3.0), 0.05, 2700, 3, 20000),
    "c1long": ((3.0, 3.0, 3.0), 0.05, 2700, 3, 200000),
    "pore": ((4.0, 4.0, 6.0), 0.1, 30000, 3, 20000),          # a typical pore analysis: 96 000 cells, 10 000 OW
    "c2": ((8.0, 8.0, 13.4), 0.1, 60000, 3, 4000),
    "c2all": ((8.0, 8.0, 13.4), 0.1, 60000, 1, 2000),
    "c3": ((20.0, 20.0, 20.0), 0.05, 1000000, 3, 1152),
    "c3all": ((20.0, 20.0, 20.0), 0.05, 1000000, 1, 384),
    "c4": ((10.0, 10.0, 10.0), 0.02, 100000, 3, 9600),
    "c4k": ((10.0, 10.0, 10.0), 0.02, 100000, 3, 9600),     # clustered: Gaussians (sigma 0.1 nm) around 256 sites (SURVEY 8d "K")
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    L, delta, natoms, stride, nframes = SHAPES[name]
    if len(sys.argv) > 2:
        nframes = int(sys.argv[2])
    geom = g.setup_geometry(L, delta)
    gindex = np.arange(0, natoms, stride, dtype=np.int32)
    gen = g.make_gen(1, L, K.GEN_CLUSTER, nsites=256, sigma=0.1) if name.endswith("k") else g.make_gen(1, L)
    dev = g.DeviceFrames(nframes, natoms).generate(gen)
    kernels = (("gather", K.KERNEL_GATHER), ("stream", K.KERNEL_STREAM), ("packed", K.KERNEL_PACKED), ("private", K.KERNEL_PRIVATE),
               ("auto", K.KERNEL_AUTO))
    only = os.environ.get("PERF_KERNELS")
    for kname, kernel in kernels:
        if only and kname not in only.split(","):
            continue
        ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel)
        for _ in range(3):
            ctr.add_device_frames(dev)
        ctr.sync()
        times = []
        for _ in range(10):
            ctr.timer_start()
            ctr.add_device_frames(dev)
            times.append(ctr.timer_stop())
        ctr.sync()
        af = nframes * len(gindex)
        best, med = min(times), float(np.median(times))
        print("%s %s grid=%s atom-frames=%.3g  best %.3f ms (%.1f G/s, %.1f GB/s algorithmic)  median %.3f ms (%.1f G/s)"
              % (name, kname, geom.shape, af, best, af / best * 1e-6, af * 12 / best * 1e-6, med, af / med * 1e-6))
        st = ctr.stats()
        if st.packed_rounds:
            print("    packed rounds %d, redone %d" % (st.packed_rounds, st.packed_redone))
        if st.private_launches:
            print("    private launches %d, redone %d" % (st.private_launches, st.private_redone))
        ctr.close()
    dev.free()


if __name__ == "__main__":
    main()
