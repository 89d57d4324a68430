#!/usr/bin/env python
"""What the host-to-device link of this box delivers for one big pinned copy, next to the end-to-end binning path
driven from the same pinned memory with different staging-batch sizes (development aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridcount_b200 as g  # noqa: E402

x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
print("raw pinned H2D: %.1f GB/s" % (5 * (1 << 30) / (time.perf_counter() - t0) * 1e-9))
del x, d

L, delta, natoms = (20.0, 20.0, 20.0), 0.05, 1000000
geom = g.setup_geometry(L, delta)
gindex = np.arange(0, natoms, 3, dtype=np.int32)
NF = 96
dev = g.DeviceFrames(NF, natoms).generate(g.make_gen(1, L))
pin = g.PinnedFrames(NF, natoms)
for f0 in range(0, NF, 32):
    pin.array[f0:f0 + 32] = dev.download(f0, 32)
dev.free()
w = np.ones(NF, np.float32)
dens = np.zeros(geom.shape, np.float32)
for fpb, nslots in ((0, 0), (4, 3), (8, 3), (16, 3), (8, 4), (2, 6)):
    ctr = g.GridCounter(geom, gindex, natoms, frames_per_batch=fpb, nslots=nslots)

    def step():
        ctr.reset_grid()
        for _ in range(6):
            ctr.add_host_ptr(pin.ptr, NF, w)
        ctr.density_into(float(NF * 6), dens)
    step()
    t0 = time.perf_counter()
    for _ in range(2):
        step()
    dt = (time.perf_counter() - t0) / 2
    print("frames_per_batch %2d slots %d: %.3f G atom-frames/s, %.1f GB/s of H2D" % (fpb, nslots, NF * 6 * len(gindex) / dt * 1e-9,
                                                                                 NF * 6 * natoms * 12 / dt * 1e-9))
    ctr.close()
pin.free()
