"""How host frames reach the device (gcb_host_staging): packed staging by host threads, DMA out of the caller's
pinned memory, and both lanes at once must give the oracle's counts, per-frame hits, sumP and f32 occupancy bit for
bit.  Replaces the atom loop's `x[index[i]]` reads (g_ri3Dc.c:425-426)."""
import numpy as np
import pytest

import gcutil as u
import gridcount_b200 as g
from gridcount_b200 import _cabi as K
from test_gpu_binning import oracle_counts, oracle_occupancy

pytestmark = pytest.mark.gpu

MODES = [K.HOST_PACK_NEVER, K.HOST_PACK_ALWAYS, K.HOST_PACK_AUTO]


def jittered_weights(nframes, seed):
    rng = np.random.default_rng(seed)
    t = np.cumsum(np.float32(2.0) + rng.integers(-2, 3, nframes).astype(np.float32) * np.float32(2.0 ** -18)).astype(np.float32)
    return g.frame_weights(t, True)


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kernel", [K.KERNEL_AUTO, K.KERNEL_GATHER, K.KERNEL_STREAM_INDEXED])
@pytest.mark.parametrize("shape", [0, 1, 2, 3, 4, 5])
def test_every_staging_mode_equals_the_oracle(shape, kernel, mode, pinned):
    natoms, gindex, nframes, L, delta = [
        (3000, np.arange(0, 3000, 3), 61, (3.0, 3.0, 3.0), 0.1),                     # water oxygens
        (1500, np.arange(1, 1500, 7), 90, (2.5, 3.0, 3.5), (0.1, 0.125, 0.07)),      # g0 = 1, fewer than 256 selected
        (4099, np.random.default_rng(5).integers(0, 4099, 1500), 23, (4.0, 4.0, 4.0), 0.08),   # unsorted, duplicates
        (9000, np.array([17, 4242, 8999]), 40, (5.0, 5.0, 5.0), 0.25),               # three ions
        (100, np.random.default_rng(6).integers(0, 100, 250), 50, (2.0, 2.0, 2.0), 0.2),       # more entries than atoms
        (10, np.random.default_rng(7).integers(0, 10, 100), 30, (2.0, 2.0, 2.0), 0.2),         # ... so many that no packed frame fits a slot
    ][shape]
    gindex = gindex.astype(np.int32)
    x = u.gen_frames(u.make_gen(900 + shape, tuple(1.1 * v for v in L)), 0, nframes, natoms)
    x -= np.float32(0.05) * np.asarray(L, np.float32)
    geom = g.setup_geometry(L, delta)
    w, T = jittered_weights(nframes, 7 + shape)
    ocounts, ohits, oovf = oracle_counts(geom, x, gindex)
    occ, osumP = oracle_occupancy(geom, x, gindex, w)
    ctr = g.GridCounter(geom, gindex, natoms, kernel=kernel, frames_per_batch=8)
    ctr.set_host_pack(mode)
    pin = None
    if pinned:
        pin = g.PinnedFrames(nframes, natoms)
        pin.array[:] = x
        ctr.add_host_ptr(pin.ptr, nframes, w)
        pin.array[:] = 0                                  # the call has consumed x
    else:
        ctr.add_frames(x, w)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    u.assert_bits_equal(ctr.occupancy(), occ, "f32 occupancy")
    st = ctr.stats()
    assert st.sumP == osumP and st.edge_overflow == oovf and st.frames == nframes
    hs = ctr.host_staging()
    assert hs.mode == mode
    fits = (8 * natoms) // len(gindex) > 0                # packed frames per slot (frames_per_batch = 8)
    if not fits:
        assert hs.frames_packed == 0
    elif mode == K.HOST_PACK_ALWAYS or (mode == K.HOST_PACK_AUTO and not pinned):
        assert hs.frames_packed == nframes and hs.h2d_bytes == nframes * len(gindex) * 12
    if mode == K.HOST_PACK_NEVER:
        assert hs.frames_packed == 0 and hs.h2d_bytes == nframes * natoms * 12
    assert hs.frames_packed + hs.frames_in_place == (nframes if pinned or hs.frames_packed else 0)
    ctr.close()
    if pin:
        pin.free()


@pytest.mark.parametrize("mode", MODES)
def test_staging_modes_with_the_rectangular_wrap(mode):
    """GCB_PBC_RECT reads the per-frame box by frame number: the packed layout must keep frames apart"""
    o = u.oracle()
    L = (3.0, 3.5, 4.0)
    natoms, nframes = 1200, 37
    x = u.gen_frames(u.make_gen(78, tuple(1.6 * v for v in L)), 0, nframes, natoms)
    x -= np.float32(0.3) * np.asarray(L, np.float32)
    boxes = np.tile(u.box9(L), (nframes, 1)).astype(np.float32)
    boxes[:, 0] *= np.linspace(1.0, 1.05, nframes, dtype=np.float32)     # breathing box: wrap uses each frame's box
    geom = g.setup_geometry(L, 0.1)
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    xw = x.copy()
    o.gco_wrap_frames(u.fptr(xw), nframes, natoms, u.fptr(boxes))
    ocounts, ohits, _ = oracle_counts(geom, xw, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, pbc=K.PBC_RECT, frames_per_batch=5)
    ctr.set_host_pack(mode)
    ctr.add_frames(x, box=boxes)
    np.testing.assert_array_equal(ctr.counts(), ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), ohits)
    ctr.close()


def test_both_lanes_carry_pinned_frames():
    """AUTO with pinned frames: copies out of the caller's memory and packed batches are in flight at the same time;
    every frame takes exactly one of the two lanes and the grid is the oracle's"""
    natoms, nframes, L = 300000, 96, (6.0, 6.0, 6.0)      # 3.6 MB per frame: an in-place copy outlasts the next submissions
    gindex = np.arange(0, natoms, 3, dtype=np.int32)
    dev = g.DeviceFrames(nframes, natoms).generate(g.make_gen(77, L))
    pin = g.PinnedFrames(nframes, natoms)
    pin.array[:] = dev.download()
    dev.free()
    geom = g.setup_geometry(L, 0.1)
    ocounts, ohits, _ = oracle_counts(geom, pin.array, gindex)
    ctr = g.GridCounter(geom, gindex, natoms, frames_per_batch=4)
    for rep in range(3):                                  # later calls pack with the measured rate
        ctr.add_host_ptr(pin.ptr, nframes)
    np.testing.assert_array_equal(ctr.counts(), 3 * ocounts)
    np.testing.assert_array_equal(ctr.hits_per_frame(), np.tile(ohits, 3))
    hs = ctr.host_staging()
    assert hs.frames_packed + hs.frames_in_place == 3 * nframes
    assert hs.frames_in_place > 0 and hs.frames_packed > 0, (hs.frames_in_place, hs.frames_packed)
    assert hs.h2d_bytes == hs.frames_in_place * natoms * 12 + hs.frames_packed * len(gindex) * 12
    assert hs.pack_bytes_per_s > 0 and hs.threads >= 1
    # the counter's trial: after ~40 ms of calls on the two lanes and ~40 ms of plain DMA it settles on the faster way
    assert hs.lanes == -1
    more = 0
    while ctr.host_staging().lanes < 0 and more < 200:
        ctr.add_host_ptr(pin.ptr, nframes)
        more += 1
    hs = ctr.host_staging()
    assert hs.lanes in (0, 1) and hs.two_lane_frames_per_s > 0 and hs.in_place_frames_per_s > 0, (more, hs.lanes)
    assert (hs.two_lane_frames_per_s > 1.05 * hs.in_place_frames_per_s) == (hs.lanes == 1)
    before = (hs.frames_packed, hs.frames_in_place)
    ctr.add_host_ptr(pin.ptr, nframes)
    hs = ctr.host_staging()
    assert hs.frames_packed + hs.frames_in_place == before[0] + before[1] + nframes
    if hs.lanes == 0:
        assert hs.frames_packed == before[0]
    np.testing.assert_array_equal(ctr.counts(), (3 + more + 1) * ocounts)
    assert ctr.stats().frames == (3 + more + 1) * nframes
    # an explicit slot count keeps the round-1 staging (the caller sized the pipeline)
    ctr2 = g.GridCounter(geom, gindex, natoms, frames_per_batch=4, nslots=2)
    ctr2.add_host_ptr(pin.ptr, nframes)
    np.testing.assert_array_equal(ctr2.counts(), ocounts)
    assert ctr2.host_staging().frames_packed == 0
    ctr2.close()
    ctr.close()
    pin.free()
