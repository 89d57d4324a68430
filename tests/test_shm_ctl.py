"""The shared-memory control plane of the multi-GPU reduce (gridcount_b200/csrc/shm_ctl.cpp): an all-gather of a few
words between the ranks of one node.  No GPU involved: real processes attach to one segment and exchange tables."""
import ctypes as C
import multiprocessing as mp
import os

import numpy as np

from gridcount_b200 import _cabi as K
from gridcount_b200._cabi import lib

WORDS = 8


def _rank(job_id, rank, nranks, rounds, q):
    h = C.c_void_p()
    rc = lib.gcb_ctl_open(C.byref(h), job_id, rank, nranks, WORDS, 20000)
    if rc:
        q.put((rank, "open failed: %s" % lib.gcb_last_error().decode()))
        return
    bad = None
    mine = np.zeros(WORDS, np.uint64)
    everything = np.zeros((nranks, WORDS), np.uint64)
    for k in range(rounds):
        mine[:] = [rank, k, rank * 1000003 + k, 7, 0, 0, 0, (rank + 1) * (k + 1)]
        rc = lib.gcb_ctl_allgather(h, mine.ctypes.data_as(K.c_u64_p), everything.ctypes.data_as(K.c_u64_p), 20000)
        if rc:
            bad = "allgather %d failed: %s" % (k, lib.gcb_last_error().decode())
            break
        for r in range(nranks):
            want = [r, k, r * 1000003 + k, 7, 0, 0, 0, (r + 1) * (k + 1)]
            if list(everything[r]) != want:
                bad = "round %d: rank %d saw %s for rank %d" % (k, rank, list(everything[r]), r)
                break
        if bad:
            break
    lib.gcb_ctl_close(h)
    q.put((rank, bad))


def test_allgather_between_processes():
    nranks, rounds = 4, 2000
    job_id = os.urandom(128)
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rank, args=(job_id, r, nranks, rounds, q)) for r in range(nranks)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    assert sorted(r for r, _ in got) == list(range(nranks))
    assert all(b is None for _, b in got), got


def test_open_times_out_when_a_rank_is_missing():
    h = C.c_void_p()
    rc = lib.gcb_ctl_open(C.byref(h), os.urandom(128), 0, 2, WORDS, 300)     # rank 1 never comes
    assert rc != 0 and "attached" in lib.gcb_last_error().decode()
    rc = lib.gcb_ctl_open(C.byref(h), os.urandom(128), 1, 2, WORDS, 300)     # rank 0 never creates the block
    assert rc != 0 and "did not appear" in lib.gcb_last_error().decode()
    assert lib.gcb_ctl_open(C.byref(h), os.urandom(128), 2, 2, WORDS, 300) != 0
