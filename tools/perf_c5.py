#!/usr/bin/env python
"""BASELINE config 5 at a size that fits a measurement: frame-sharded run over N GPUs on 1 M atoms per frame
(333 334 water oxygens counted, 20 nm box breathing with pressure coupling, 0.05 nm grid = 400^3 cells, 256 MB
of uint32 per rank), NCCL integer grid all-reduce, density, and a_ri3Dc's slice averages over the reduced grid.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/perf_c5.py [frames_per_rank]

Every rank keeps `frames_per_rank` generated frames resident in HBM (12 MB each) and bins them REPS times per
timed job, so a job is REPS * frames_per_rank frames per rank.  Prints one JSON line on rank 0.  Size-independent
checks: the reduced grid's total equals the gathered per-frame hit counts, and the first frames of rank 0 are
replayed on the CPU oracle."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gcutil as u  # noqa: E402
import gridcount_b200 as g  # noqa: E402
from gridcount_b200.api import nccl_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    REPS, JOBS = 4, 3
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, delta, natoms = (20.0, 20.0, 20.0), 0.05, 1_000_000
    geom = g.setup_geometry(L, delta)
    assert geom.shape == (400, 400, 400)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    gen = g.make_gen(20261019 + 5, L, g._cabi.GEN_UNIFORM, box_period=1000, box_amp=0.002)
    dev = g.DeviceFrames(F, natoms, local).generate(gen, frame0=rank * F)
    ctr = g.GridCounter(geom, gindex, natoms, device=local)
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctr.comm_init(uid[0], rank, world)
    w = np.ones(F, np.float32)
    T_tot = float(F * REPS * world)

    def job(ev=None):
        ctr.reset_grid()
        if ev is not None:
            ctr.event_record(ev)
        for _ in range(REPS):
            ctr.add_device_frames(dev, weights=w)
        if ev is not None:
            ctr.event_record(ev + 1)
        if world > 1:
            ctr.allreduce()
        if ev is not None:
            ctr.event_record(ev + 2)
        return ctr.density_device(T_tot)

    job()
    ctr.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctr.event_record(2)
    for j in range(JOBS):
        dptr = job(ev=10 + 4 * j)
    ctr.event_record(3)
    ctr.sync()
    if world > 1:
        dist.barrier()
    ms = ctr.event_elapsed(2, 3) / JOBS
    bin_ms = float(np.mean([ctr.event_elapsed(10 + 4 * j, 11 + 4 * j) for j in range(JOBS)]))
    red_ms = float(np.mean([ctr.event_elapsed(11 + 4 * j, 12 + 4 * j) for j in range(JOBS)]))
    if world > 1:
        t = torch.tensor([ms, bin_ms, red_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, bin_ms, red_ms = (float(v) for v in t.tolist())
    units = float(len(gindex)) * F * REPS * world
    st = ctr.stats()
    counts_total = int(ctr.counts().astype(np.uint64).sum()) if rank == 0 else 0
    line = None
    if rank == 0:
        # a_ri3Dc's slice averages over the reduced density, which stays in HBM
        t0 = time.perf_counter()
        r = g.analyse(geom.tg, None, unit="SPC", grid_dev=dptr, device=local)
        k3_ms = (time.perf_counter() - t0) * 1e3
        # replay of rank 0's first 2 frames on the oracle: their hit counts must match the gathered ones
        x = dev.download(0, 2)
        og = u.Geom.from_buffer_copy(bytes(geom.tg))
        want = np.zeros(geom.shape, np.uint32)
        hits = np.zeros(2, np.uint64)
        u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(x), 2, natoms, u.iptr(gindex),
                                    len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
        mine = ctr.hits_per_frame()
        line = {"config": "C5 shape: 1M atoms/frame, 333334 OW counted, 400^3 grid, breathing box", "n_gpus": world,
                "frames_per_job": F * REPS * world, "G_atom_frames_per_s": units / (ms * 1e-3) * 1e-9, "ms_per_job": ms,
                "binning_ms": bin_ms, "allreduce_256MB_ms": red_ms, "hits": int(st.hits), "atom_frames": int(st.atom_frames),
                "grid_total_equals_hits": counts_total == int(st.hits),
                "oracle_prefix_hits_equal": bool(np.array_equal(mine[:2], hits)),
                "a_ri3Dc_on_reduced_grid_ms": k3_ms, "zdf_mid": float(r["zdf"][200, 1]), "sumP_density": float(r["sumP"])}
        print(json.dumps(line), flush=True)
    ctr.close()
    dev.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
