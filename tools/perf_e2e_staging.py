#!/usr/bin/env python
"""End-to-end rate of gcb_counter_add_frames from PINNED host frames under the three staging modes
(development aid; bench.py is the contract): NEVER = every batch DMA-ed out of the caller's memory, ALWAYS = every
batch packed by the host threads, AUTO = both lanes at once.  Same step as bench.py's e2e: reset, F frames through
the C ABI, density read back.  No torch.  Usage: python tools/perf_e2e_staging.py [c2|c3] [calls] [steps]"""
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridcount_b200 as g  # noqa: E402
from gridcount_b200 import _cabi as K  # noqa: E402

SHAPES = {"c2": ((8.0, 8.0, 13.4), 0.1, 60000, 1536), "c3": ((20.0, 20.0, 20.0), 0.05, 1000000, 96)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    L, delta, natoms, hostf = SHAPES[name]
    geom = g.setup_geometry(L, delta)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    dev = g.DeviceFrames(hostf, natoms).generate(g.make_gen(1, L))
    pin = g.PinnedFrames(hostf, natoms)
    for f0 in range(0, hostf, 32):
        n = min(32, hostf - f0)
        pin.array[f0:f0 + n] = dev.download(f0, n)
    dev.free()
    w = np.ones(hostf, np.float32)
    dens = np.zeros(geom.shape, np.float32)
    units = float(len(gindex)) * hostf * calls
    crc = {}
    for mname, mode in (("never", K.HOST_PACK_NEVER), ("auto", K.HOST_PACK_AUTO), ("always", K.HOST_PACK_ALWAYS), ("auto", K.HOST_PACK_AUTO)):
        ctr = g.GridCounter(geom, gindex, natoms)
        ctr.set_host_pack(mode)

        def step():
            ctr.reset_grid()
            for _ in range(calls):
                ctr.add_host_ptr(pin.ptr, hostf, w)
            ctr.density_into(float(hostf * calls), dens)

        step()
        h0 = ctr.host_staging()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
        h1 = ctr.host_staging()
        st = ctr.stats()
        assert st.hits == st.atom_frames == int(units), (st.hits, st.atom_frames, units)
        crc[mname] = zlib.crc32(dens.tobytes())
        print("%s %-6s %.1f ms/step  %.3f G atom-frames/s  h2d %.2f GB/step  packed %d in-place %d frames/step  pack %.1f GB/s on %d threads  "
              "trial: lanes=%d two-lane %.0f in-place %.0f frames/s  crc %08x"
              % (name, mname, dt * 1e3, units / dt * 1e-9, (h1.h2d_bytes - h0.h2d_bytes) / steps * 1e-9,
                 (h1.frames_packed - h0.frames_packed) // steps, (h1.frames_in_place - h0.frames_in_place) // steps,
                 h1.pack_bytes_per_s * 1e-9, h1.threads, h1.lanes, h1.two_lane_frames_per_s, h1.in_place_frames_per_s,
                 crc[mname]), flush=True)
        ctr.close()
    assert len(set(crc.values())) == 1, "density differs between staging modes"
    pin.free()


if __name__ == "__main__":
    main()
