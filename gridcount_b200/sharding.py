"""Host-side logic of the frame-sharded multi-GPU run: which frames a rank bins, the
frame weights at shard boundaries, and the frame-ordered statistics of the whole job.

The reference has no multi-process mode; its manual equivalent is one g_ri3Dc run per trajectory
segment merged by a_gridcalc (a_gridcalc.c:180-189).  Here the ranks reproduce ONE g_ri3Dc run over
all frames: `dt = t - tlast` of a shard's first frame needs the previous shard's last time
(g_ri3Dc.c:413-414), and `T_tot += dt`, `sumP += weight` are f64 chains in frame order (:415, :426).
Pure host code: works on any torch.distributed backend (the CPU tests use gloo)."""
import ctypes as C

import numpy as np

from . import _cabi as K
from ._cabi import lib, check


def frame_shard(nframes, world, rank):
    """contiguous frame range [start, stop) of `rank`; ranges are in rank order and differ by at most one frame"""
    base, extra = divmod(nframes, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_frame_weights(t_shard, t_before=None, dtweight=True, t_after=None):
    """Frame weights of a shard exactly as the single run computes them.
    t_before: time of the frame preceding the shard (None for the first shard, whose tlast is bootstrapped as
    t[0] - (t[1] - t[0]), g_ri3Dc.c:390-394; a one-frame first shard needs t_after = time of the next frame).
    -> (weights f32[n], partial T_tot as the list of f32 dt values to be chained in frame order)"""
    t = np.ascontiguousarray(t_shard, np.float32)
    n = len(t)
    w = np.zeros(n, np.float32)
    if n == 0:
        return w, np.zeros(0, np.float32)
    T = C.c_double(0)
    if t_before is None:
        if n == 1 and t_after is not None:
            t2 = np.array([t[0], np.float32(t_after)], np.float32)
            w2 = np.zeros(2, np.float32)
            check(lib.gcb_frame_weights(t2.ctypes.data_as(K.c_float_p), 2, int(bool(dtweight)), None,
                                        w2.ctypes.data_as(K.c_float_p), C.byref(T)), "gcb_frame_weights")
            w[0] = w2[0]
        else:
            check(lib.gcb_frame_weights(t.ctypes.data_as(K.c_float_p), n, int(bool(dtweight)), None,
                                        w.ctypes.data_as(K.c_float_p), C.byref(T)), "gcb_frame_weights")
    else:
        tlast = C.c_float(np.float32(t_before))
        check(lib.gcb_frame_weights(t.ctypes.data_as(K.c_float_p), n, int(bool(dtweight)), C.byref(tlast),
                                    w.ctypes.data_as(K.c_float_p), C.byref(T)), "gcb_frame_weights")
    # dt per frame (the T_tot increments), independent of -nodtweight
    prev = np.concatenate([[np.float32(t_before) if t_before is not None else np.float32(0)], t[:-1]]).astype(np.float32)
    dt = (t - prev).astype(np.float32)
    if t_before is None:
        nxt = t[1] if n > 1 else (np.float32(t_after) if t_after is not None else t[0])
        dt[0] = np.float32(t[0] - np.float32(t[0] - np.float32(nxt - t[0])))
    return w, dt


def gather_job_statistics(hits_per_frame, weights, dts, group=None):
    """All-gather the per-frame hit counts, weights and time steps of every rank (rank order = frame
    order) and redo the reference's f64 chains over the whole job -> (sumP, T_tot, total hits)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.stack([torch.from_numpy(np.asarray(hits_per_frame, np.float64).copy()),
                        torch.from_numpy(np.asarray(weights, np.float32).astype(np.float64)),
                        torch.from_numpy(np.asarray(dts, np.float32).astype(np.float64))])
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([mine.shape[1]], dtype=torch.int64), group=group)
    nmax = max(int(s.item()) for s in sizes)
    pad = torch.zeros(3, nmax, dtype=torch.float64)
    pad[:, :mine.shape[1]] = mine
    parts = [torch.zeros(3, nmax, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    sumP, T_tot, hits = 0.0, 0.0, 0
    for p, s in zip(parts, sizes):
        n = int(s.item())
        for f in range(n):
            h, w, dt = int(p[0, f].item()), float(p[1, f].item()), float(p[2, f].item())
            T_tot += dt                      # T_tot += (double)dt
            sumP = _seq_add(sumP, w, h)      # sumP += (double)weight, h times
            hits += h
    return sumP, T_tot, hits


def _seq_add(v, w, n):
    """n sequential f64 additions of w to v; exact shortcut while the partial sums stay exactly representable"""
    if n == 0:
        return v
    # weights are f32 values: n*w is exact in f64 for n < 2^29, and adding a multiple of ulp(w) to a sum of such
    # multiples below 2^53 ulp(w) is exact, so the chain equals one multiply-add; otherwise fall back to the loop
    import math
    if w != 0.0 and n < (1 << 29):
        m, e = math.frexp(w)
        q = math.ldexp(1.0, e - 24)          # ulp(w) as an f32 value
        if abs(v) < q * (1 << 52) and abs(v + n * w) < q * (1 << 52) and math.fmod(v, q) == 0.0:
            return v + n * w
    for _ in range(n):
        v += w
    return v
