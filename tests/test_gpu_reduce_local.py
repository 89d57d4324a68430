"""The frame-sharded reduce (gcb_counter_allreduce) on ONE GPU: the ranks are counters driven by threads of this
process over the in-process transport (gcb_counter_comm_init_local).  Same protocol as the NCCL transport -- rank
table through the shared-memory control plane, saturation guard, integer grid sum, in-place gather of the per-frame
statistics, frame-ordered sumP chain -- so what tests/test_gpu_multi.py checks on a multi-GPU box is checked here
on any box.  The reference's manual equivalent is a_gridcalc's merge (a_gridcalc.c:180-189); the oracle is the
reference's loop over ALL frames in order (g_ri3Dc.c:410-429)."""
import ctypes as C
import threading

import numpy as np
import pytest

import gcutil as u
import gridcount_b200 as g
from gridcount_b200 import _cabi as K
from gridcount_b200._cabi import check, lib

pytestmark = pytest.mark.gpu

L, DELTA, NATOMS = (4.0, 4.0, 6.7), 0.1, 3000
GEN = dict(slab=(2.2, 4.5), pore_r=0.8, box_period=16, box_amp=0.01)


def run_ranks(world, body):
    """body(rank, ctr_factory) on `world` threads; the factory attaches the counter to the shared communicator"""
    lc = C.c_void_p()
    check(lib.gcb_local_comm_create(C.byref(lc), world), "gcb_local_comm_create")
    out, err = [None] * world, [None] * world

    def attach(ctr, rank):
        check(lib.gcb_counter_comm_init_local(ctr.h, lc, rank), "gcb_counter_comm_init_local")
        return ctr

    def run(rank):
        try:
            out[rank] = body(rank, lambda ctr: attach(ctr, rank))
        except BaseException as e:          # noqa: BLE001  (reported to the main thread)
            err[rank] = e

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank is stuck in the reduce"
    lib.gcb_local_comm_destroy(lc)
    for e in err:
        if e is not None:
            raise e
    return out


def oracle_all(geom, x, gindex, weights, T_tot):
    og = u.Geom.from_buffer_copy(bytes(geom.tg))
    n = x.shape[0]
    want = np.zeros(geom.shape, np.uint32)
    hits = np.zeros(n, np.uint64)
    u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(x), n, x.shape[1],
                                u.iptr(gindex), len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
    occ = np.zeros(geom.shape, np.float32)
    sp = C.c_double(0)
    u.oracle().gco_bin_frames(C.byref(og), u.fptr(occ), u.fptr(x), n, x.shape[1], u.iptr(gindex), len(gindex),
                              u.fptr(weights), C.byref(sp), None, None)
    occ64 = np.zeros(geom.shape, np.float64)
    u.oracle().gco_bin_frames_f64(C.byref(og), occ64.ctypes.data_as(u.c_double_p), u.fptr(x), n, x.shape[1],
                                  u.iptr(gindex), len(gindex), u.fptr(weights))
    dens = occ.copy()
    u.oracle().gco_normalise(C.byref(og), u.fptr(dens), T_tot)
    return want, hits, occ, occ64, dens, sp.value


@pytest.mark.parametrize("world", [2, 3, 4])
def test_uniform_weight_reduce_is_bit_exact(world):
    """one weight for the whole job, breathing box (atoms leave the grid): counts, per-frame hits, sumP, occupancy and
    density of every rank equal the oracle over all frames, bit for bit; two jobs on the same communicator"""
    F = 48
    geom = g.setup_geometry(L, DELTA)
    gindex = np.arange(0, NATOMS, 3, dtype=np.int32)
    gen = g.make_gen(5, L, K.GEN_PORE, **GEN)
    x = u.gen_frames(u.make_gen(5, L, 2, **GEN), 0, F * world, NATOMS)
    wall = np.full(F * world, 0.1, np.float32)
    want, hits, occ, _, dens, sumP = oracle_all(geom, x, gindex, wall, 0.1 * F * world)
    assert 0 < int(hits.sum()) < F * world * len(gindex)

    ndev = g.device_count()                      # the ranks spread over the box's GPUs (all on one where there is one)

    def body(rank, attach):
        dev = g.DeviceFrames(F, NATOMS, rank % ndev).generate(gen, frame0=rank * F)
        ctr = attach(g.GridCounter(geom, gindex, NATOMS, device=rank % ndev))
        res = []
        for job in range(2):
            ctr.reset()
            ctr.add_device_frames(dev, weights=wall[:F])
            ctr.allreduce()
            st = ctr.stats()
            res.append((ctr.counts(), ctr.hits_per_frame(), st.frames, st.hits, st.sumP, ctr.occupancy(),
                        ctr.density(0.1 * F * world)[1]))
        ctr.close(); dev.free()
        return res

    for rank, res in enumerate(run_ranks(world, body)):
        for counts, lh, frames, nh, sp, o, d in res:
            np.testing.assert_array_equal(counts, want)
            np.testing.assert_array_equal(lh, hits[rank * F:(rank + 1) * F])
            assert frames == F * world and nh == int(hits.sum()) and sp == sumP
            np.testing.assert_array_equal(o.view(np.uint32), occ.view(np.uint32))
            np.testing.assert_array_equal(d.view(np.uint32), dens.view(np.uint32))


def test_ragged_shards_host_frames_and_an_idle_rank():
    """ranks hold 0, 24, 40 frames; one rank feeds host frames through the staging slots, one device frames; the rank
    without frames still takes part in the reduce"""
    world, Fr = 3, [0, 24, 40]
    off = [0, 0, 24]
    geom = g.setup_geometry(L, DELTA)
    gindex = np.arange(0, NATOMS, 3, dtype=np.int32)
    gen = g.make_gen(5, L, K.GEN_PORE, **GEN)
    x = u.gen_frames(u.make_gen(5, L, 2, **GEN), 0, sum(Fr), NATOMS)
    wall = np.full(sum(Fr), 0.25, np.float32)
    want, hits, occ, _, dens, sumP = oracle_all(geom, x, gindex, wall, 0.25 * sum(Fr))

    def body(rank, attach):
        ctr = attach(g.GridCounter(geom, gindex, NATOMS, frames_per_batch=7))
        if rank == 1:
            ctr.add_frames(x[off[1]:off[1] + Fr[1]], weights=wall[:Fr[1]])
        elif rank == 2:
            dev = g.DeviceFrames(Fr[2], NATOMS).generate(gen, frame0=off[2])
            ctr.add_device_frames(dev, weights=wall[:Fr[2]])
        ctr.allreduce()
        st = ctr.stats()
        res = (ctr.counts(), ctr.hits_per_frame(), st.frames, st.hits, st.sumP, ctr.density(0.25 * sum(Fr))[1])
        ctr.close()
        return res

    for rank, (counts, lh, frames, nh, sp, d) in enumerate(run_ranks(world, body)):
        np.testing.assert_array_equal(counts, want)
        np.testing.assert_array_equal(lh, hits[off[rank]:off[rank] + Fr[rank]])
        assert frames == sum(Fr) and nh == int(hits.sum()) and sp == sumP
        np.testing.assert_array_equal(d.view(np.uint32), dens.view(np.uint32))


@pytest.mark.parametrize("world", [2, 4])
def test_jittered_weights_across_ranks_stay_within_the_f32_summation_bound(world):
    """Frame weights that differ from frame to frame (f32-jittered XTC time steps, the default -dtweight case) on
    several ranks: counts, per-frame hits and sumP are exact; the occupancy is a sum of per-rank f32 chains, which is
    NOT the reference's single frame-ordered chain (gridcount_b200.h says so).  With ~10^4 hits in the fullest cells
    both are measured against the `real = double` accumulation: the multi-rank result must stay inside the standard
    bound of sequential f32 summation of n positive terms, (n - 1) * 2^-24 relative, and is reported next to the
    reference chain's own error."""
    natoms, F = 6000, 1536                       # per rank; a coarse grid, so cells collect thousands of hits
    Lc = (2.0, 2.0, 2.0)
    geom = g.setup_geometry(Lc, 0.25)            # 8 x 8 x 8 cells
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    gen = g.make_gen(77, Lc, K.GEN_UNIFORM)
    x = u.gen_frames(u.make_gen(77, Lc, 0), 0, F * world, natoms)
    # time stamps k * 0.1 stored in f32, dt = t[k] - t[k-1] in f32 (g_ri3Dc.c:413): a handful of distinct weights
    t = (np.arange(F * world + 1, dtype=np.float64) * 0.1).astype(np.float32)
    wall = (t[1:] - t[:-1]).astype(np.float32)
    assert len(np.unique(wall)) > 2
    T_tot = float(wall.astype(np.float64).sum())
    want, hits, occ, occ64, dens, sumP = oracle_all(geom, x, gindex, wall, T_tot)
    assert want.max() > 10000

    def body(rank, attach):
        dev = g.DeviceFrames(F, natoms).generate(gen, frame0=rank * F)
        ctr = attach(g.GridCounter(geom, gindex, natoms))
        ctr.add_device_frames(dev, weights=wall[rank * F:(rank + 1) * F])
        ctr.allreduce()
        st = ctr.stats()
        res = (ctr.counts(), st.hits, st.sumP, st.uniform_weight, ctr.occupancy())
        ctr.close(); dev.free()
        return res

    outs = run_ranks(world, body)
    err_ref = np.abs(occ.astype(np.float64) - occ64) / occ64
    for counts, nh, sp, uni, o in outs:
        np.testing.assert_array_equal(counts, want)
        assert nh == int(hits.sum()) and sp == sumP and uni == 0
        err = np.abs(o.astype(np.float64) - occ64) / occ64
        bound = (want.astype(np.float64) - 1.0) * 2.0 ** -24
        assert (err <= bound + 2.0 ** -24).all(), (err.max(), bound.max())
        # same order as the reference chain's own distance from the exact sum (printed with -s)
        print("world %d: max rel. error vs f64: %d ranks %.3g, reference f32 chain %.3g, bound %.3g"
              % (world, world, err.max(), err_ref.max(), bound.max()))
        np.testing.assert_array_equal(o.view(np.uint32), outs[0][4].view(np.uint32))      # every rank holds the same bits
