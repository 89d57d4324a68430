#!/usr/bin/env python
"""bench.py -- headline benchmark of the gridcount hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): hydrophobic-nanopore membrane system,
60 000 atoms of which the 20 000 water oxygens (every third atom) are counted, 10 000 frames,
0.1 nm grid -> 80 x 80 x 134 cells.  Synthetic pore-shaped positions from the counter-based
generator (bit-reproducible on the CPU oracle).

One STEP is one complete g_ri3Dc job over the rank's 10 000-frame shard: zero the grid, bin all
frames (K1), combine the per-GPU grids with one NCCL integer all-reduce when N > 1, normalise to
the number density (K2).  `value` is binned atom-frames per second with the frames resident in HBM;
`e2e` is the same job driven through the C ABI from pinned HOST frames (H2D staging in batches on
CUDA streams inside the timed region, density grid read back).  N > 1 is frame sharding: every
rank gets its own 10 000 frames (weak scaling).

Beside the contract's keys the line carries, at N = 1: `a_ri3Dc` (K3, gcb_analyse over the job's density: wall time
from a host grid and from the density left in HBM, the oracle's single-thread time, bit-equality), `xtc` (host
decode rate, GPU decode rate, the job driven from compressed bytes) and, in `roofline`, the DRAM fraction the
kernel really reaches (36 B of traffic per unit) and the fraction of the fitted L2 ceiling.

--impl reference times the reference's own CPU implementation of the path (the compiled,
unmodified gridcount() loop in oracle/_ref/libgcref.so; the oracle's C port if that binary is
absent) on all host threads, on bounded samples of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = dict(name="C2 nanopore: 60000 atoms, 20000 OW counted, 10000 frames, 0.1 nm grid 80x80x134",
                L=(8.0, 8.0, 13.4), delta=0.1, natoms=60000, stride=3, nframes=10000,
                slab=(4.7, 8.7), pore_r=0.8, seed=20261019 + 2, dt=1.0)
ALGO_BYTES_PER_UNIT = 12          # three f32 coordinates of one selected atom
METRIC = "binned atom-frames/sec"
UNIT = "G atom-frames/s"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic():
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [v.strip() for v in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples in the upper half of the observed power range
        lo, hi = min(power), max(power)
        load = [s for s, p in zip(sm, power) if p >= lo + 0.5 * (hi - lo)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "power_w_max": hi,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def cpu_reference(sample_frames, repeats, threads, frames_host, gindex, L, delta, natoms):
    """Times the reference's atom loop (g_ri3Dc.c:425-426) over `sample_frames` frames held in memory,
    `repeats` times, sharded over `threads` threads with private grids.  -> (atom-frames/s, kind)"""
    import gcutil as u
    b9 = u.box9(L)
    dl = np.asarray((delta,) * 3, np.float32)
    gi = np.ascontiguousarray(gindex, np.int32)
    x = np.ascontiguousarray(frames_host[:sample_frames], np.float32)
    total = 0.0
    if u.have_ref():
        r = u.ref()
        kind = "reference"
        for _ in range(repeats):
            sp = C.c_double(0)
            total += r.gcref_bench(u.fptr(b9), u.fptr(dl), u.fptr(x), sample_frames, natoms, u.iptr(gi), len(gi),
                                   1.0, threads, C.byref(sp), None)
    else:
        o = u.oracle()
        kind = "port"
        og, _ = u.oracle_geometry(L, (delta,) * 3)
        for _ in range(repeats):
            sp = C.c_double(0)
            total += o.gco_bench(C.byref(og), u.fptr(x), sample_frames, natoms, u.iptr(gi), len(gi), 1.0, threads,
                                 C.byref(sp))
    return sample_frames * len(gi) * repeats / total, kind


def host_sample_frames(nframes, wl):
    import gcutil as u
    gen = u.make_gen(wl["seed"], wl["L"], 2, slab=wl["slab"], pore_r=wl["pore_r"])
    return u.gen_frames(gen, 0, nframes, wl["natoms"])


def run_reference_arm(args, rank, world):
    wl = WORKLOAD
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    gindex = np.arange(0, wl["natoms"], wl["stride"], dtype=np.int32)
    sample = 256                                   # frames held in memory (18 MB), re-binned `repeats` times
    x = host_sample_frames(sample, wl)
    # size a step to ~2 s of wall time on this host
    rate, kind = cpu_reference(sample, 1, threads, x, gindex, wl["L"], wl["delta"], wl["natoms"])
    per_rep = sample * len(gindex) / rate
    repeats = int(max(1, min(200, round(2.0 / max(per_rep, 1e-6)))))
    for _ in range(args.warmup):
        cpu_reference(sample, 1, threads, x, gindex, wl["L"], wl["delta"], wl["natoms"])
    t_units = 0.0
    t_secs = 0.0
    for _ in range(args.steps):
        r, kind = cpu_reference(sample, repeats, threads, x, gindex, wl["L"], wl["delta"], wl["natoms"])
        units = sample * len(gindex) * repeats
        t_units += units
        t_secs += units / r
    value = t_units / t_secs * 1e-9
    sample_txt = ("%d generated frames of the workload held in host memory, binned %d times per step by %d threads "
                  "with private grids (decode excluded)" % (sample, repeats, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "grid": [80, 80, 134], "distribution": "pore",
                       "inputs": "host memory"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample_txt},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def xtc_section(g, ctr, dev, wl, geom, gindex):
    """256 frames of the workload written as XTC (precision 1000) by the host codec (host/libgcbhost.so):
    single-thread host decode rate, GPU decode rate, and the binning job driven from pinned XTC bytes."""
    import gcutil as u
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "libgcbhost.so"])
    H = C.CDLL(os.path.join(ROOT, "host", "libgcbhost.so"))
    H.gh_xtc_encode_frame.argtypes = [C.c_int, C.c_int, C.c_float, u.c_float_p, u.c_float_p, C.c_float, u.c_u8_p, C.c_size_t]
    H.gh_xtc_encode_frame.restype = C.c_size_t
    H.gh_xtc_decode_frame.argtypes = [u.c_u8_p, C.c_size_t, C.c_void_p, u.c_float_p, C.c_int]
    H.gh_xtc_decode_frame.restype = C.c_size_t
    natoms, NF = wl["natoms"], 256
    x = dev.download(0, NF)
    b9 = u.box9(wl["L"])
    buf = np.zeros(natoms * 16 + 256, np.uint8)
    parts = []
    for f in range(NF):
        n = H.gh_xtc_encode_frame(natoms, f, float(f), u.fptr(b9), u.fptr(x[f]), 1000.0, buf.ctypes.data_as(u.c_u8_p), len(buf))
        assert n > 0
        parts.append(buf[:n].tobytes())
    data = b"".join(parts)
    raw = np.frombuffer(data, np.uint8).copy()
    info = (C.c_char * 64)()
    out = np.zeros((natoms, 3), np.float32)
    t0 = time.perf_counter()
    off = 0
    for f in range(NF):
        view = raw[off:]
        off += H.gh_xtc_decode_frame(view.ctypes.data_as(u.c_u8_p), len(view), C.cast(info, C.c_void_p), u.fptr(out), natoms)
    host_s = time.perf_counter() - t0
    REP = 16                                     # 4096 frames per call: enough frames for one-thread-per-frame decoding
    xb = g.XtcBuffer(data * REP, pinned=True)
    w = np.ones(xb.nframes, np.float32)
    d = xb.decode_to_device(); d.free()          # warm-up
    t0 = time.perf_counter()
    d = xb.decode_to_device()
    gpu_decode_s = time.perf_counter() - t0
    d.free()
    ctr.reset_grid(); ctr.add_xtc(xb, w); ctr.sync()
    t0 = time.perf_counter()
    ctr.reset_grid(); ctr.add_xtc(xb, w); ctr.sync()
    e2e_s = time.perf_counter() - t0
    res = {"frames": xb.nframes, "compressed_bytes_per_atom": len(data) / (NF * natoms),
           "host_decode_atoms_per_s": NF * natoms / host_s, "host_decode_threads": 1,
           "gpu_decode_atoms_per_s": xb.nframes * natoms / gpu_decode_s,
           "gpu_decode_note": "includes H2D of the compressed bytes and the device allocation",
           "e2e_from_xtc_bytes_G_atom_frames_per_s": xb.nframes * len(gindex) / e2e_s * 1e-9,
           "data": "synthetic uniform positions compress to ~5 B/atom; real water compresses further"}
    xb.free()
    ctr.reset_grid()
    return res


def k3_section(g, ctr, T_tot):
    """a_ri3Dc's reductions (K3, gcb_analyse) over the job's density: wall time through the C ABI from a HOST grid
    (H2D of the grid, all kernels, D2H of every output), with the oracle's single-thread time beside it.
    Reported on its own line, not folded into the metric."""
    import gcutil as u
    tg, dens, _ = ctr.density(T_tot)
    g.analyse(tg, dens, unit="SPC")                       # warm-up
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = g.analyse(tg, dens, unit="SPC")
        ts.append(time.perf_counter() - t0)
    # the same with the density left in HBM (gcb_counter_device_density -> gcb_analyse_device): no grid over PCIe
    dptr = ctr.density_device(T_tot)
    ctr.sync()
    g.analyse(tg, None, unit="SPC", grid_dev=dptr)
    td = []
    for _ in range(5):
        t0 = time.perf_counter()
        g.analyse(tg, None, unit="SPC", grid_dev=dptr)
        td.append(time.perf_counter() - t0)
    og = u.Geom.from_buffer_copy(bytes(tg))
    u.oracle().gco_update_b(C.byref(og))
    an = u.Analysis()
    an.unit = u.UNITS["SPC"]; an.rad_ba_step = 1
    t0 = time.perf_counter()
    u.oracle().gco_analyse(C.byref(og), u.fptr(dens), C.byref(an))
    cpu_s = time.perf_counter() - t0
    arrs = u.analysis_arrays(an, og)
    def _same(p, q):
        return np.array_equal(np.isnan(p), np.isnan(q)) and bool(((p.view(np.uint32) == q.view(np.uint32)) | np.isnan(p)).all())
    same = all(_same(arrs[k], r[k]) for k in ("rzp", "zdf", "rdf", "xyp"))
    same = same and np.float32(an.average).view(np.uint32) == np.float32(r["average"]).view(np.uint32) and an.sumP == r["sumP"]
    u.oracle().gco_analysis_free(C.byref(an))
    cells = int(np.prod(dens.shape))
    return {"grid": list(dens.shape), "ms": float(np.median(ts)) * 1e3, "cells_per_s": cells / float(np.median(ts)),
            "ms_device_resident_grid": float(np.median(td)) * 1e3,
            "kernel_launches": int(r["kernel_launches"]), "oracle_ms_1_thread": cpu_s * 1e3, "bit_equal_to_oracle": bool(same),
            "note": "host grid in, all profiles out; every sum is the reference's sequential f32/f64 chain, evaluated in parallel"}


def bind_to_gpu_numa_node(device):
    """Pin this rank's threads (and with them its pinned staging memory, which is placed where the allocating
    thread runs) to the NUMA node the GPU hangs off, so that H2D DMA reads do not cross the socket interconnect.
    Best effort: returns a short description, or None when the topology cannot be read."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id if hasattr(torch.cuda.get_device_properties(device), "pci_bus_id") else None
        if bus is None:
            out = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)], text=True)
            bus = out.strip()
        bus = bus.lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:      # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import gridcount_b200 as g
    from gridcount_b200 import _cabi as K
    from gridcount_b200.api import nccl_unique_id

    wl = WORKLOAD
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    numa = bind_to_gpu_numa_node(device) if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    natoms, F = wl["natoms"], wl["nframes"]
    gindex = np.arange(0, natoms, wl["stride"], dtype=np.int32)
    gnx = len(gindex)
    geom = g.setup_geometry(wl["L"], wl["delta"])
    assert geom.shape == (80, 80, 134)
    gen = g.make_gen(wl["seed"], wl["L"], K.GEN_PORE, slab=wl["slab"], pore_r=wl["pore_r"])
    dev = g.DeviceFrames(F, natoms, device).generate(gen, frame0=rank * F)       # 7.2 GB resident in HBM
    weights = np.full(F, wl["dt"], np.float32)
    T_tot = float(F * world) * wl["dt"]

    ctr = g.GridCounter(geom, gindex, natoms, device=device)
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctr.comm_init(uid[0], rank, world)

    def step_device(ev=None):
        ctr.reset_grid()
        if ev is not None:
            ctr.event_record(ev)
        ctr.add_device_frames(dev, weights=weights)
        if ev is not None:
            ctr.event_record(ev + 1)
        if world > 1:
            ctr.allreduce()
        ctr.density_device(T_tot)

    # ---- device-resident value -----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    ctr.sync()
    launches0 = ctr.stats().kernel_launches
    sampler = ClockSampler(device) if rank == 0 else None     # one poller per job: nvidia-smi polling perturbs launches
    barrier()
    if sampler:
        sampler.start()
    ctr.event_record(2)
    for s in range(args.steps):
        step_device(ev=4 + 2 * s)
    ctr.event_record(3)
    ctr.sync()
    barrier()
    # the device-timed region lasts tens of milliseconds; the sampler keeps running through the end-to-end region
    # below (the same kernels fed over PCIe) so that the clocks line rests on more than one or two samples
    ms_total = ctr.event_elapsed(2, 3)
    k1_ms = [ctr.event_elapsed(4 + 2 * s, 5 + 2 * s) for s in range(args.steps)]
    launches = ctr.stats().kernel_launches - launches0
    if dist is not None:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        ll = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ll)
        launches = int(ll.item())
    units_per_step = float(gnx) * F * world
    value = units_per_step * args.steps / (ms_total * 1e-3) * 1e-9

    # parity spot check inside the bench: global hit count equals frames*gnx (all generated atoms are in-grid)
    st = ctr.stats()
    assert st.hits == st.atom_frames, "bench grid lost hits"

    # ---- roofline of the dominant kernel (K1, the streaming binning kernel) -----------------
    peak, peak_src = read_peaks()
    k1_avg_ms = float(np.mean(k1_ms))
    achieved = ALGO_BYTES_PER_UNIT * float(gnx) * F / (k1_avg_ms * 1e-3) * 1e-9      # GB/s, this rank's launch
    traffic = read_traffic()
    roofline = {"bound": "hbm", "kernel": "bin_stream_kernel (K1)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UNIT * gnx * F,
                "launch_ms_avg": k1_avg_ms,
                "l2_atomic_peak_gops": (traffic or {}).get("l2_atomic_peak_gops"),
                "frac_of_l2_atomic_peak": ((float(gnx) * F / (k1_avg_ms * 1e-3) * 1e-9) /
                                           traffic["l2_atomic_peak_gops"]) if traffic and
                traffic.get("l2_atomic_peak_gops") else None}
    if traffic:
        # what the kernel really moves: every 32-byte sector of a frame holds part of a counted atom when one atom in
        # three is counted, so DRAM traffic is 36 B per unit whatever the kernel does
        roofline["dram_frac"] = traffic["dram_bytes_per_launch"] / (k1_avg_ms * 1e-3) * 1e-9 / peak
        roofline["dram_bytes_per_unit"] = traffic["dram_bytes_per_launch"] / (float(gnx) * F)
    # L2 work model fitted to three measurements: a RED costs 1.625 read sectors, the L2
    # sustains 310 G sector-equivalents/s; this workload needs 36/32 read sectors + 1 RED per unit
    l2_ceiling = 310.0 / (36.0 / 32.0 + 1.625)
    roofline["l2_model_ceiling_g_units_s"] = l2_ceiling
    roofline["frac_of_l2_model"] = (float(gnx) * F / (k1_avg_ms * 1e-3) * 1e-9) / l2_ceiling

    # ---- end to end: pinned host frames through the C ABI ----------------------------------
    HOSTF = 1000                                     # distinct frames kept in pinned memory (720 MB)
    pin = g.PinnedFrames(HOSTF, natoms)
    chunk = 250
    for f0 in range(0, HOSTF, chunk):
        pin.array[f0:f0 + chunk] = dev.download(f0, chunk)
    wh = np.full(HOSTF, wl["dt"], np.float32)
    dens_host = np.zeros(geom.shape, np.float32)

    def step_e2e():
        ctr.reset_grid()
        for _ in range(F // HOSTF):
            ctr.add_host_ptr(pin.ptr, HOSTF, wh)        # H2D in batches on copy streams, overlapped with K1
        if world > 1:
            ctr.allreduce()
        ctr.density_into(T_tot, dens_host)              # K2 + D2H of the density grid

    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "device-timed steps + end-to-end steps"
    e2e = {"value": units_per_step * e2e_steps / e2e_s * 1e-9, "unit": UNIT,
           "h2d_bytes_per_step": int(F) * natoms * 12 * world,
           "d2h_bytes_per_step": int(geom.cells * 4 + F * 8) * world, "steps": e2e_steps,
           "ms_per_step": e2e_s / e2e_steps * 1e3,
           "path": "gcb_counter_add_frames(pinned host x[F][natoms][3]) -> gcb_counter_density(host grid)"}
    pin.free()

    # ---- a_ri3Dc's reductions over the job's density (K3), on its own line -----------------------
    k3 = None
    if rank == 0 and world == 1:
        try:
            k3 = k3_section(g, ctr, T_tot)
        except Exception as e:
            k3 = {"error": str(e)[:200]}

    # ---- XTC decode, timed separately (north star): host decoder vs GPU decoder on the same frames ----
    xtc = None
    if rank == 0 and world == 1:
        try:
            xtc = xtc_section(g, ctr, dev, wl, geom, gindex)
        except Exception as e:          # the section is informative; never let it take the headline down
            xtc = {"error": str(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        sample = 256
        xh = dev.download(0, sample)
        r1, kind = cpu_reference(sample, 2, 1, xh, gindex, wl["L"], wl["delta"], natoms)
        per = sample * gnx / r1
        rep1 = int(max(1, min(20, round(4.0 / per))))
        r1, kind = cpu_reference(sample, rep1, 1, xh, gindex, wl["L"], wl["delta"], natoms)
        rp, kind = cpu_reference(sample, 4, threads, xh, gindex, wl["L"], wl["delta"], natoms)
        repp = int(max(1, min(400, round(6.0 * rp / (sample * gnx)))))
        rp, kind = cpu_reference(sample, repp, threads, xh, gindex, wl["L"], wl["delta"], natoms)
        cpu = {"value": rp * 1e-9, "unit": UNIT, "cores": threads, "kind": kind,
               "single_core_value": r1 * 1e-9,
               "sample": "%d frames of the workload (from the same generator) in host memory; 1 thread x %d passes "
                         "and %d threads x %d passes with private grids; trajectory decode excluded"
                         % (sample, rep1, threads, repp)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "grid": list(geom.shape), "distribution": "pore (generator mode 2)",
                           "frames_resident_gb": F * natoms * 12 / 1e9, "l2_policy": "inputs (7.2 GB) larger than L2",
                           "parallelism": "frame-sharded x%d, one NCCL u32 all-reduce per job" % world,
                           "host_binding": numa,
                           "step": "reset + K1 bin 10000 frames + (all-reduce) + K2 density"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "a_ri3Dc": k3, "xtc": xtc}
        print(json.dumps(line), flush=True)
    ctr.close()
    dev.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
