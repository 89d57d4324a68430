"""Worker for tests/test_gpu_multi.py: run under torchrun with one rank per GPU.  Every rank bins its
own frame shard, the grids are combined with the NCCL integer all-reduce (gcb_counter_allreduce), and
rank 0 compares the global counts, statistics and density with the oracle run over ALL frames."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gcutil as u  # noqa: E402
import gridcount_b200 as g  # noqa: E402
from gridcount_b200.api import nccl_unique_id  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, delta, natoms, F = (4.0, 4.0, 6.7), 0.1, 3000, 64          # frames per rank
    geom = g.setup_geometry(L, delta)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    # breathing box (BASELINE config 5: pressure coupling): the grid is fixed by the first frame's box, atoms that
    # end up outside are dropped by the reference and by every rank alike
    gen = g.make_gen(5, L, g._cabi.GEN_PORE, slab=(2.2, 4.5), pore_r=0.8, box_period=16, box_amp=0.01)
    dev = g.DeviceFrames(F, natoms, local).generate(gen, frame0=rank * F)
    w = np.full(F, 0.1, np.float32)
    ctr = g.GridCounter(geom, gindex, natoms, device=local)
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctr.comm_init(uid[0], rank, world)
    ctr.add_device_frames(dev, weights=w)
    ctr.allreduce()
    counts = ctr.counts()
    st = ctr.stats()
    local_hits = ctr.hits_per_frame()           # this rank's shard; filled from the gathered counters
    tg, dens, _ = ctr.density(0.1 * F * world)
    ok = True
    if rank == 0:
        og = u.Geom.from_buffer_copy(bytes(geom.tg))
        x = u.gen_frames(u.make_gen(5, L, 2, slab=(2.2, 4.5), pore_r=0.8, box_period=16, box_amp=0.01), 0, F * world, natoms)
        want = np.zeros(geom.shape, np.uint32)
        hits = np.zeros(F * world, np.uint64)
        u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(x), F * world, natoms,
                                    u.iptr(gindex), len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
        occ = np.zeros(geom.shape, np.float32)
        sp = C.c_double(0)
        wall = np.full(F * world, 0.1, np.float32)
        u.oracle().gco_bin_frames(C.byref(og), u.fptr(occ), u.fptr(x), F * world, natoms, u.iptr(gindex), len(gindex),
                                  u.fptr(wall), C.byref(sp), None, None)
        u.oracle().gco_normalise(C.byref(og), u.fptr(occ), 0.1 * F * world)
        assert 0 < int(hits.sum()) < F * world * len(gindex), "the breathing box should push some atoms out of the grid"
        ok = (np.array_equal(counts, want) and st.frames == F * world and st.hits == int(hits.sum())
              and np.array_equal(dens.view(np.uint32), occ.view(np.uint32)) and st.sumP == sp.value
              and np.array_equal(local_hits, hits[:F]))
        print("MULTI_GPU_UNIFORM", "OK" if ok else "FAIL", "world", world, "hits", st.hits, flush=True)
    ctr.close(); dev.free()

    # ---- second job: ragged shards and a weight that differs between ranks ------------------------------
    # rank r holds 16 + 8 r frames of weight 0.1 (even r) or 0.25 (odd r).  Counts, per-frame hits and the
    # frame-ordered sumP stay exact; the f32 occupancy becomes a sum of per-rank accumulators, which is not the
    # reference's sequential order any more: 1e-6 relative.
    Fr = [16 + 8 * r for r in range(world)]
    off = [sum(Fr[:r]) for r in range(world)]
    wr = [np.float32(0.1 if r % 2 == 0 else 0.25) for r in range(world)]
    dev = g.DeviceFrames(Fr[rank], natoms, local).generate(gen, frame0=off[rank])
    ctr = g.GridCounter(geom, gindex, natoms, device=local)
    uid2 = [nccl_unique_id() if rank == 0 else None]          # a unique id initialises one communicator
    dist.broadcast_object_list(uid2, src=0)
    ctr.comm_init(uid2[0], rank, world)
    ctr.add_device_frames(dev, weights=np.full(Fr[rank], wr[rank], np.float32))
    ctr.allreduce()
    counts = ctr.counts()
    st = ctr.stats()
    Ttot = float(sum(float(wr[r]) * Fr[r] for r in range(world)))
    tg, dens, _ = ctr.density(Ttot)
    local_hits = ctr.hits_per_frame()
    if rank == 0:
        Fall = sum(Fr)
        x = u.gen_frames(u.make_gen(5, L, 2, slab=(2.2, 4.5), pore_r=0.8, box_period=16, box_amp=0.01), 0, Fall, natoms)
        want = np.zeros(geom.shape, np.uint32)
        hits = np.zeros(Fall, np.uint64)
        u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(x), Fall, natoms,
                                    u.iptr(gindex), len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
        wall = np.concatenate([np.full(Fr[r], wr[r], np.float32) for r in range(world)])
        occ = np.zeros(geom.shape, np.float32)
        sp = C.c_double(0)
        u.oracle().gco_bin_frames(C.byref(og), u.fptr(occ), u.fptr(x), Fall, natoms, u.iptr(gindex), len(gindex),
                                  u.fptr(wall), C.byref(sp), None, None)
        u.oracle().gco_normalise(C.byref(og), u.fptr(occ), Ttot)
        close = np.allclose(dens, occ, rtol=1e-6, atol=0.0)
        ok2 = (np.array_equal(counts, want) and st.frames == Fall and st.hits == int(hits.sum()) and st.sumP == sp.value
               and np.array_equal(local_hits, hits[:Fr[0]]) and close)
        print("MULTI_GPU_RAGGED", "OK" if ok2 else "FAIL", "world", world, "frames", Fr, "density close", close,
              "sumP", st.sumP, sp.value, flush=True)
        ok = ok and ok2
    ctr.close(); dev.free()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if ok else "FAIL", "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
