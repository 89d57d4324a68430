"""CPU tests (gloo, world_size 2 and 3) of the host logic of the frame-sharded multi-GPU run:
shard ranges, frame weights across shard boundaries, frame-ordered job statistics."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import gcutil as u

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import gcutil as uu
    from gridcount_b200 import sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, x, gindex, L, delta, dtweight = case
    F = len(t)
    a, b = sh.frame_shard(F, world, rank)
    w, dts = sh.shard_frame_weights(t[a:b], t[a - 1] if a > 0 else None, dtweight, t[b] if b < F else None)
    # the per-frame hit counts of this shard (here from the CPU oracle; on the GPU box they come from the counter)
    og, _ = uu.oracle_geometry(L, (delta,) * 3)
    counts = np.zeros(og.shape(), np.uint32)
    hits = np.zeros(b - a, np.uint64)
    xs = np.ascontiguousarray(x[a:b])
    uu.oracle().gco_count_frames(C.byref(og), counts.ctypes.data_as(uu.c_u32_p), uu.fptr(xs), b - a, x.shape[1],
                                 uu.iptr(gindex), len(gindex), hits.ctypes.data_as(uu.c_u64_p), None)
    sumP, T_tot, nh = sh.gather_job_statistics(hits, w, dts)
    q.put((rank, a, b, w.tobytes(), sumP, T_tot, nh))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dtweight,dt", [(2, 1, 0.1), (3, 1, 1.0), (2, 0, 0.1)])
def test_sharded_statistics_equal_single_run(world, dtweight, dt):
    L, delta, natoms, F = (3.0, 3.0, 3.0), 0.25, 300, 11
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    x = u.gen_frames(u.make_gen(3, (3.3, 3.3, 3.3)), 0, F, natoms)          # box larger than the grid: some misses
    t = (np.float32(5.0) + np.arange(F, dtype=np.float32) * np.float32(dt)).astype(np.float32)
    # single run through the oracle
    o = u.oracle()
    og, _ = u.oracle_geometry(L, (delta,) * 3)
    w1 = np.zeros(F, np.float32)
    T1 = C.c_double(0)
    o.gco_frame_weights(u.fptr(t), F, dtweight, u.fptr(w1), C.byref(T1))
    grid = np.zeros(og.shape(), np.float32)
    sp = C.c_double(0)
    hits = np.zeros(F, np.uint64)
    o.gco_bin_frames(C.byref(og), u.fptr(grid), u.fptr(x), F, natoms, u.iptr(gindex), len(gindex), u.fptr(w1),
                     C.byref(sp), hits.ctypes.data_as(u.c_u64_p), None)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + dtweight
    procs = [ctx.Process(target=_worker, args=(r, world, port, (t, x, gindex, L, delta, dtweight), q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wcat = b"".join(r[3] for r in res)
    assert wcat == w1.tobytes(), "frame weights across shard boundaries differ from the single run"
    assert [r[1] for r in res][0] == 0 and res[-1][2] == F and all(res[i][2] == res[i + 1][1] for i in range(world - 1))
    for r in res:
        assert r[4] == sp.value, "sumP"
        assert r[5] == T1.value, "T_tot"
        assert r[6] == int(hits.sum())


def test_frame_shard_partition():
    sys.path.insert(0, os.path.dirname(HERE))
    from gridcount_b200 import sharding as sh
    for F in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            ranges = [sh.frame_shard(F, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == F
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
