#!/usr/bin/env python
"""bench.py -- benchmark of the gridcount hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[2], SURVEY.md 8d "C3" -- the largest single-GPU configuration):
1M-atom solvated box of 20 nm, the 333 334 water oxygens (every third atom) counted, 0.05 nm grid = 400^3 cells
(256 MB of uint32 counters, beyond the L2).  Synthetic uniform positions from the counter-based generator
(bit-reproducible on the CPU oracle).  Each rank keeps RESIDENT (1152) frames in HBM, 13.8 GB, far beyond the L2.

One STEP = one batch of the job on every rank: zero the grid, bin PASSES x RESIDENT frames (K1), combine the per-GPU
grids with one NCCL integer all-reduce when N > 1, normalise to the number density (K2).  `value` is binned
atom-frames per second with the frames resident in HBM; `e2e` is the same path driven through the C ABI from
pinned HOST frames (H2D staging in batches on CUDA streams inside the timed region, density grid read back).
N > 1 is frame sharding: every rank gets its own frames (weak scaling).

Beside the contract's keys the line carries
  roofline      K1's dominant kernel: achieved = 12 B x atom-frames / summed launch durations, measured live with CUDA
                events around every launch; the memory-system ceilings for K1's traffic mix measured in the same run
                by gcb_selftest_l2 (random-RED peak, and a read stream with one RED per 36 bytes)
  per_config    (N = 1) the other named shapes C1, C2, C4 (uniform and clustered) and a pore-sized grid P1, each with
                value, roofline, the kernel that ran and a bit-exact oracle check of a prefix of its frames;
                (N > 1) C1, C2, C4 frame-sharded with the NCCL all-reduce, value only
  e2e_xtc       the end-to-end job from XTC bytes (compressed bytes over PCIe, decoded and binned on the GPU)
  c5            (N > 1) the C5 shape: breathing box, 256 MB NCCL all-reduce, with its own oracle parity check
  parity_checked / parity   counts and per-frame hits of a frame prefix (of every rank's shard) equal the CPU oracle's
  a_ri3Dc, xtc  (N = 1) K3 over the job's density; XTC decode rates

--impl reference times the reference's own CPU implementation of the path (the compiled, unmodified gridcount()
loop in oracle/_ref/libgcref.so; the oracle's C port if that binary is absent) on all host threads, on bounded
samples of the same workload.
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

SEED = 20261019
CONFIGS = {
    "C1": dict(title="C1 SPC water box 3 nm: 2700 atoms, 900 OW counted, 0.05 nm grid 60x60x60",
               L=(3.0, 3.0, 3.0), delta=0.05, natoms=2700, stride=3, named_frames=1000, resident=20000, gen="uniform",
               grid=(60, 60, 60), seed=SEED + 1, prefix=64, reset_each_pass=True,
               note="the named job (1000 frames, 32 MB) would be L2-resident and launch-bound; measured over 20000 "
                    "distinct frames (648 MB)"),
    "P1": dict(title="P1 (not a BASELINE config) pore region 4 x 4 x 6 nm: 30000 atoms, 10000 OW counted, 0.1 nm grid "
                     "40x40x60 = 96000 cells, which fit one CTA's shared memory as 8-bit private counters",
               L=(4.0, 4.0, 6.0), delta=0.1, natoms=30000, stride=3, named_frames=20000, resident=20000, gen="uniform",
               grid=(40, 40, 60), seed=SEED + 6, prefix=64, reset_each_pass=True),
    "C2": dict(title="C2 nanopore: 60000 atoms, 20000 OW counted, 10000 frames, 0.1 nm grid 80x80x134",
               L=(8.0, 8.0, 13.4), delta=0.1, natoms=60000, stride=3, named_frames=10000, resident=10000, gen="pore",
               slab=(4.7, 8.7), pore_r=0.8, grid=(80, 80, 134), seed=SEED + 2, prefix=64),
    "C3": dict(title="C3 1M-atom solvated box 20 nm: 333334 OW counted, 0.05 nm grid 400x400x400 (grid beyond L2)",
               L=(20.0, 20.0, 20.0), delta=0.05, natoms=1000000, stride=3, named_frames=100000, resident=1152,
               gen="uniform", grid=(400, 400, 400), seed=SEED + 3, prefix=8),
    "C4": dict(title="C4 fine grid: 10 nm box, 100000 atoms, 33334 OW counted, 0.02 nm grid 500x500x500",
               L=(10.0, 10.0, 10.0), delta=0.02, natoms=100000, stride=3, named_frames=50000, resident=9600,
               gen="uniform", grid=(500, 500, 500), seed=SEED + 4, prefix=32),
    "C4K": dict(title="C4 clustered (atomic-contention stress): as C4, Gaussians of sigma 0.1 nm around 256 sites",
                L=(10.0, 10.0, 10.0), delta=0.02, natoms=100000, stride=3, named_frames=50000, resident=9600,
                gen="cluster", nsites=256, sigma=0.1, grid=(500, 500, 500), seed=SEED + 4, prefix=32),
    "C5": dict(title="C5 frame-sharded run: 1M atoms per frame, 333334 OW counted, pressure-coupled (breathing) box, "
                     "400x400x400 grid, NCCL uint32 grid all-reduce",
               L=(20.0, 20.0, 20.0), delta=0.05, natoms=1000000, stride=3, named_frames=1000000, resident=1152,
               gen="uniform", box_period=1000, box_amp=0.002, grid=(400, 400, 400), seed=SEED + 5, prefix=4),
}
HEADLINE = "C3"
PASSES = 8                         # passes over the resident frames per step: 9216 frames per rank and step
ALGO_BYTES_PER_UNIT = 12           # three f32 coordinates of one selected atom
METRIC = "binned atom-frames/sec"
UNIT = "G atom-frames/s"
GEN_MODES = {"uniform": 0, "quantised": 1, "pore": 2, "cluster": 3}


def cfg_gindex(cfg):
    return np.arange(0, cfg["natoms"], cfg["stride"], dtype=np.int32)


def cfg_gen(mod, cfg):
    """mod: gridcount_b200 (device generator) or gcutil (the oracle's generator); same bits"""
    return mod.make_gen(cfg["seed"], cfg["L"], GEN_MODES[cfg["gen"]], slab=cfg.get("slab", (0.0, 0.0)),
                        pore_r=cfg.get("pore_r", 0.0), nsites=cfg.get("nsites", 1), sigma=cfg.get("sigma", 0.0),
                        box_period=cfg.get("box_period", 0), box_amp=cfg.get("box_amp", 0.0))


def headline_config(world):
    """the `config` object of the JSON line: identical in both arms"""
    cfg = CONFIGS[HEADLINE]
    return {"workload": cfg["title"], "grid": list(cfg["grid"]), "atoms_per_frame": cfg["natoms"],
            "counted_per_frame": len(cfg_gindex(cfg)), "distribution": cfg["gen"] + " (counter-based generator)",
            "frames_per_step_per_gpu": cfg["resident"] * PASSES,
            "step": "zero the grid + bin %d x %d frames (K1) + grid all-reduce if N > 1 + density (K2)"
                    % (PASSES, cfg["resident"]),
            "parallelism": "frame-sharded x%d" % world}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic(kernel_key):
    """per-launch DRAM bytes of a kernel from the committed ncu --set full capture of this round, if any
    (profiles/k1_traffic.json: {kernel_key: {dram_bytes_per_launch, atom_frames_per_launch, source}})"""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
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


def run_reference_arm(args, rank, world):
    """the reference's gridcount() loop on all host threads over a bounded sample of the headline workload"""
    import gcutil as u
    cfg = CONFIGS[HEADLINE]
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    gindex = cfg_gindex(cfg)
    sample = 2 * threads                           # frames held in memory (12 MB each), one or two per thread
    x = u.gen_frames(cfg_gen(u, cfg), 0, sample, cfg["natoms"])
    # size a step to ~3 s of WALL time on this host (the first call also allocates the per-thread grids)
    cpu_reference(sample, 1, threads, x, gindex, cfg["L"], cfg["delta"], cfg["natoms"])
    t0 = time.perf_counter()
    rate, kind = cpu_reference(sample, 1, threads, x, gindex, cfg["L"], cfg["delta"], cfg["natoms"])
    per_rep = max(time.perf_counter() - t0, 1e-3)
    repeats = int(max(1, min(50, round(3.0 / per_rep), 120.0 / (max(args.steps, 1) * per_rep))))     # and ~2 minutes in all
    for _ in range(min(args.warmup, 2)):
        cpu_reference(sample, 1, threads, x, gindex, cfg["L"], cfg["delta"], cfg["natoms"])
    t_units = 0.0
    t_secs = 0.0
    for _ in range(args.steps):
        r, kind = cpu_reference(sample, repeats, threads, x, gindex, cfg["L"], cfg["delta"], cfg["natoms"])
        units = sample * len(gindex) * repeats
        t_units += units
        t_secs += units / r
    value = t_units / t_secs * 1e-9
    sample_txt = ("%d generated frames of the workload (%d MB) held in host memory, binned %d times per step by %d threads, "
                  "each into a private 256 MB grid and a contiguous frame range; grid allocation and trajectory decode "
                  "excluded from the timed region" % (sample, sample * cfg["natoms"] * 12 >> 20, repeats, threads))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample_txt},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def xtc_encode(dev, wl, nf, build=True):
    """`nf` frames of the resident batch written as XTC (precision 1000) by the host codec (host/libgcbhost.so)
    -> (bytes, the codec library)"""
    import gcutil as u
    if build:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "libgcbhost.so"])
    H = C.CDLL(os.path.join(ROOT, "host", "libgcbhost.so"))
    H.gh_xtc_encode_frame.argtypes = [C.c_int, C.c_int, C.c_float, u.c_float_p, u.c_float_p, C.c_float, u.c_u8_p, C.c_size_t]
    H.gh_xtc_encode_frame.restype = C.c_size_t
    H.gh_xtc_decode_frame.argtypes = [u.c_u8_p, C.c_size_t, C.c_void_p, u.c_float_p, C.c_int]
    H.gh_xtc_decode_frame.restype = C.c_size_t
    natoms = wl["natoms"]
    x = dev.download(0, nf)
    b9 = u.box9(wl["L"])
    buf = np.zeros(natoms * 16 + 256, np.uint8)
    parts = []
    for f in range(nf):
        n = H.gh_xtc_encode_frame(natoms, f, float(f), u.fptr(b9), u.fptr(x[f]), 1000.0, buf.ctypes.data_as(u.c_u8_p), len(buf))
        assert n > 0
        parts.append(buf[:n].tobytes())
    return b"".join(parts), H


def xtc_section(g, ctr, dev, wl, geom, gindex):
    """256 frames of the workload written as XTC (precision 1000) by the host codec (host/libgcbhost.so):
    single-thread host decode rate, GPU decode rate, and the binning job driven from pinned XTC bytes."""
    import gcutil as u
    natoms, NF = wl["natoms"], 256
    data, H = xtc_encode(dev, wl, NF)
    raw = np.frombuffer(data, np.uint8).copy()
    info = (C.c_char * 64)()
    out = np.zeros((natoms, 3), np.float32)
    t0 = time.perf_counter()
    off = 0
    for f in range(NF):
        view = raw[off:]
        off += H.gh_xtc_decode_frame(view.ctypes.data_as(u.c_u8_p), len(view), C.cast(info, C.c_void_p), u.fptr(out), natoms)
    host_s = time.perf_counter() - t0
    REP = 16                                     # 4096 frames per call
    xb = g.XtcBuffer(data * REP, pinned=True)
    w = np.ones(xb.nframes, np.float32)
    d = xb.decode_to_device()                    # warm-up; the buffer is reused below
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        xb.decode_to_device(dev=d)
        ts.append(time.perf_counter() - t0)
    gpu_decode_s = float(np.median(ts))
    d.free()
    dens = np.zeros(geom.shape, np.float32)
    T = float(xb.nframes)

    def step():
        ctr.reset_grid()
        ctr.add_xtc(xb, w)                       # compressed bytes H2D in batches + skim + decode-and-bin kernels
        ctr.density_into(T, dens)                # K2 + D2H of the density grid

    step()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    e2e_s = float(np.median(ts))
    # parity inside the bench: the fused decode-and-bin path against decode-to-x + the binning kernels
    got = ctr.counts()
    d2 = xb.decode_to_device()
    c2 = g.GridCounter(geom, gindex, natoms, device=ctr.device, kernel=g._cabi.KERNEL_STREAM)
    c2.add_device_frames(d2, weights=w)
    same = bool(np.array_equal(got, c2.counts()))
    c2.close(); d2.free()
    res = {"frames": xb.nframes, "compressed_bytes_per_atom": len(data) / (NF * natoms),
           "host_decode_atoms_per_s": NF * natoms / host_s, "host_decode_threads": 1,
           "gpu_decode_atoms_per_s": xb.nframes * natoms / gpu_decode_s,
           "gpu_decode_note": "pinned compressed bytes -> x[frame][atom][3] in HBM (H2D + skim + segment decode), buffer preallocated",
           "e2e_xtc": {"value": xb.nframes * len(gindex) / e2e_s * 1e-9, "unit": UNIT, "h2d_bytes_per_step": int(xb.nbytes),
                       "d2h_bytes_per_step": int(geom.cells * 4), "ms_per_step": e2e_s * 1e3,
                       "atoms_decoded_per_s": xb.nframes * natoms / e2e_s,
                       "pcie_bound_at_55GBs": 55e9 / (len(data) / (NF * natoms)) / (natoms / len(gindex)) * 1e-9,
                       "counts_equal_two_step_path": same,
                       "path": "gcb_counter_add_xtc(pinned xtc bytes): the frames are binned straight out of the GPU "
                               "decoder, no coordinates stored -> gcb_counter_density(host grid)"},
           "data": "synthetic positions of this workload, precision 1000: ~5 B/atom compressed; real water compresses further"}
    xb.free()
    ctr.reset_grid()
    return res


def k3_section(g, ctr, T_tot, oracle_too=True):
    """a_ri3Dc's reductions (K3, gcb_analyse) over the job's density: wall time through the C ABI from a HOST grid
    (H2D of the grid, all kernels, D2H of every output) and from the density left in HBM, with the oracle's
    single-thread time beside it.  Reported on its own line, not folded into the metric."""
    import gcutil as u
    tg, dens, _ = ctr.density(T_tot)
    g.analyse(tg, dens, unit="SPC")                       # warm-up
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = g.analyse(tg, dens, unit="SPC")
        ts.append(time.perf_counter() - t0)
    dptr = ctr.density_device(T_tot)
    ctr.sync()
    g.analyse(tg, None, unit="SPC", grid_dev=dptr)
    td = []
    for _ in range(5):
        t0 = time.perf_counter()
        g.analyse(tg, None, unit="SPC", grid_dev=dptr)
        td.append(time.perf_counter() - t0)
    cells = int(np.prod(dens.shape))
    out = {"grid": list(dens.shape), "ms": float(np.median(ts)) * 1e3, "cells_per_s": cells / float(np.median(ts)),
           "ms_device_resident_grid": float(np.median(td)) * 1e3,
           "roofline_frac_device_resident": cells * 4 / float(np.median(td)) * 1e-9 / read_peaks()[0],
           "kernel_launches": int(r["kernel_launches"]),
           "note": "host grid in, all profiles out; every sum is the reference's sequential f32/f64 chain, evaluated in "
                   "parallel; roofline: 4 B per cell read once"}
    if oracle_too:
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
        same = all(_same(arrs[k], r[k]) for k in ("rzp", "zdf", "rdf", "xyp", "xzp", "yzp", "lzdf"))
        same = same and np.float32(an.average).view(np.uint32) == np.float32(r["average"]).view(np.uint32) and an.sumP == r["sumP"]
        u.oracle().gco_analysis_free(C.byref(an))
        out["oracle_ms_1_thread"] = cpu_s * 1e3
        out["bit_equal_to_oracle"] = bool(same)
    return out


def oracle_counts(tg, x, gindex):
    """the CPU oracle's integer counts and per-frame hits for host frames x (checker only)"""
    import gcutil as u
    og = u.Geom.from_buffer_copy(bytes(tg))
    shape = (tg.mx[0], tg.mx[1], tg.mx[2])
    want = np.zeros(shape, np.uint32)
    hits = np.zeros(x.shape[0], np.uint64)
    u.oracle().gco_count_frames(C.byref(og), want.ctypes.data_as(u.c_u32_p), u.fptr(x), x.shape[0], x.shape[1],
                                u.iptr(gindex), len(gindex), hits.ctypes.data_as(u.c_u64_p), None)
    return want, hits


def prefix_parity(g, cfg, geom, gindex, dev, device):
    """Bit-exactness inside the bench: a prefix of the resident frames is binned by a fresh counter with the kernel
    AUTO picks for this shape and compared, cell by cell and frame by frame, with the CPU oracle on the same frames."""
    n = min(cfg["prefix"], dev.nframes)
    ctr = g.GridCounter(geom, gindex, cfg["natoms"], device=device)
    ctr.add_device_frames(dev, first=0, count=n)
    got = ctr.counts()
    hits = ctr.hits_per_frame()
    st = ctr.stats()
    ctr.close()
    want, whits = oracle_counts(geom.tg, dev.download(0, n), gindex)
    ok = bool(np.array_equal(got, want) and np.array_equal(hits, whits))
    return {"ok": ok, "frames": int(n), "hits": int(whits.sum()), "packed_rounds": int(st.packed_rounds),
            "packed_redone": int(st.packed_redone), "against": "CPU oracle (oracle/gc_oracle.c), counts and per-frame hits"}


def l2_ceilings(g, window_bytes, stride, device):
    """the memory system's ceilings for K1's traffic on this shape, measured now (gcb_selftest_l2)"""
    red_peak, _ = g.selftest_l2(window_bytes, 0, device)
    mix, rd = g.selftest_l2(window_bytes, 1 if stride == 3 else 2, device)
    _, alone = g.selftest_l2(window_bytes, 3, device)
    return {"window_mb": window_bytes / 2.0 ** 20, "red_peak_gops": red_peak,
            "stream_plus_red_gops": mix, "stream_plus_red_read_gbs": rd, "stream_alone_read_gbs": alone,
            "stream_alone_gops": alone / (12.0 * stride),
            "mix": "one random RED per %d streamed bytes" % (36 if stride == 3 else 12)}


def roofline_of(g, ctr, stride, cells, peak, peak_src, device, kernel_key, l2=None):
    """K1's dominant kernel from the event pairs gcb_counter_profile put around every K1 launch: the kind of
    binning kernel (direct uint32 REDs, packed-counter pass, or the guarded uint32 pass that redoes a failed packed
    round) with the largest summed duration; the other K1 launches of the same frames are listed beside it"""
    kinds = ctr.profile_kinds()
    dom = max(("direct", "packed", "redo", "private"), key=lambda k: kinds[k][0])
    k1_ms, nl, units = kinds[dom]
    bits = ctr.stats().packed_bits or (8 if cells <= (80 << 20) else 4)
    names = {"direct": "bin_stream_kernel<DirectSink> (one RED per hit into the uint32 grid)",
             "packed": "bin_stream_kernel<PackedSink<%d>> (non-returning REDs into %d-bit L2-resident counters)" % (bits, bits),
             "private": "bin_stream_kernel<PrivateSink> (shared-memory atomics into per-CTA private copies of the grid, "
                        "folded into the uint32 grid when it is read)",
             "redo": "bin_stream_kernel<DirectSink>, guarded pass (one RED per hit into the uint32 grid, after a packed "
                     "round failed its digit-sum verification)"}
    gps = units / (k1_ms * 1e-3) * 1e-9
    achieved = ALGO_BYTES_PER_UNIT * gps
    tr = read_traffic(kernel_key) if dom != "redo" else None
    all_ms = sum(v[0] for v in kinds.values())
    out = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
           "peak_source": peak_src, "launch_ms_avg": k1_ms / max(nl, 1), "launches_timed": int(nl),
           "algorithmic_bytes_per_launch": ALGO_BYTES_PER_UNIT * units / max(nl, 1),
           "kernel_g_atom_frames_per_s": gps, "share_of_k1_time": k1_ms / all_ms if all_ms else None,
           "k1_launch_kinds_ms": {k: round(v[0], 4) for k, v in kinds.items()},
           "traffic": None, "hbm_ceiling_note": "one atom in %d is counted out of x[atom][3]: every 32 B sector of a frame "
                                                "holds part of a counted atom, so DRAM moves %d B per unit and frac <= %.2f"
                                                % (stride, 12 * stride, 1.0 / stride) if stride > 1 else None}
    if tr and tr.get("kernel", "").find("Packed") >= 0 and dom != "packed":
        tr = None                                  # the committed capture is of the other kernel
    if tr:
        out["traffic"] = tr["dram_bytes_per_launch"] * (units / max(nl, 1)) / tr["atom_frames_per_launch"]
        out["traffic_source"] = "%s: %.2f B of DRAM traffic per atom-frame, scaled to this launch size" % (
            tr.get("source"), tr["dram_bytes_per_launch"] / tr["atom_frames_per_launch"])
        out["traffic_bytes_per_unit"] = tr["dram_bytes_per_launch"] / tr["atom_frames_per_launch"]
    if l2:
        out["l2"] = dict(l2)
        out["l2"]["frac_of_red_peak"] = gps / l2["red_peak_gops"]
        out["l2"]["frac_of_stream_plus_red"] = gps / l2["stream_plus_red_gops"]
    return out


def measure_config(g, name, cfg, device, peak, peak_src, reps=10, extra=None):
    """device-resident throughput of one named shape on one GPU, with roofline and oracle parity"""
    geom = g.setup_geometry(cfg["L"], cfg["delta"])
    assert geom.shape == tuple(cfg["grid"]), (geom.shape, cfg["grid"])
    gindex = cfg_gindex(cfg)
    gnx, F = len(gindex), cfg["resident"]
    dev = g.DeviceFrames(F, cfg["natoms"], device).generate(cfg_gen(g, cfg))
    ctr = g.GridCounter(geom, gindex, cfg["natoms"], device=device)
    w = np.ones(F, np.float32)
    for i in range(max(3, reps)):                 # also sizes the counter's per-call buffers for `reps` calls in flight
        if cfg.get("reset_each_pass"):
            ctr.reset_grid()
        ctr.add_device_frames(dev, weights=w)
        if i < 8:
            ctr.sync()                            # (the host-side policy of the packed path learns from finished calls)
    ctr.sync()
    ctr.reset_grid()
    ctr.event_record(0)
    for _ in range(reps):
        if cfg.get("reset_each_pass"):            # (private counters accumulate over launches: identical frames binned again
            ctr.reset_grid()                      #  and again would pile up as no trajectory does)
        ctr.add_device_frames(dev, weights=w)
    ctr.event_record(1)
    ms = ctr.event_elapsed(0, 1)
    st = ctr.stats()
    lost = st.hits != st.atom_frames
    ctr.profile(True)
    for _ in range(2):
        if cfg.get("reset_each_pass"):
            ctr.reset_grid()
        ctr.add_device_frames(dev, weights=w)
    packed_bytes = geom.cells if geom.cells <= (80 << 20) else geom.cells // 2
    rl = roofline_of(g, ctr, cfg["stride"], geom.cells, peak, peak_src, device, name)
    if "PackedSink<8>" in rl["kernel"]:
        packed_bytes = geom.cells
    window = packed_bytes if "PackedSink" in rl["kernel"] else geom.cells * 4       # what the dominant kernel's REDs address
    ctr.profile(False)
    res = {"workload": cfg["title"], "grid": list(geom.shape), "frames_resident": F,
           "frames_resident_gb": F * cfg["natoms"] * 12 / 1e9, "passes_timed": reps,
           "value": float(gnx) * F * reps / (ms * 1e-3) * 1e-9, "unit": UNIT,
           "ms_per_pass": ms / reps, "all_hits_in_grid": not lost,
           "packed_rounds": int(st.packed_rounds), "packed_redone": int(st.packed_redone), "packed_bits": int(st.packed_bits),
           "private_launches": int(st.private_launches), "private_redone": int(st.private_redone)}
    if cfg.get("reset_each_pass"):
        res["pass"] = "zero the grid + bin the resident frames"
    if cfg.get("note"):
        res["note"] = cfg["note"]
    if extra:
        try:
            res.update(extra(ctr, dev, geom, gindex))
        except Exception as e:
            res["extra_error"] = str(e)[:200]
    ctr.close()
    par = prefix_parity(g, cfg, geom, gindex, dev, device)
    dev.free()
    l2 = l2_ceilings(g, window, cfg["stride"], device)        # after the counters are gone: it borrows the L2 set-aside
    rl["l2"] = dict(l2)
    rl["l2"]["frac_of_red_peak"] = rl["kernel_g_atom_frames_per_s"] / l2["red_peak_gops"]
    rl["l2"]["frac_of_stream_plus_red"] = rl["kernel_g_atom_frames_per_s"] / l2["stream_plus_red_gops"]
    rl["l2"]["frac_of_stream_alone"] = rl["kernel_g_atom_frames_per_s"] / l2["stream_alone_gops"]
    res["roofline"] = rl
    res["parity_checked"] = par["ok"]
    res["parity"] = par
    return res


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
        bind_to_gpu_numa_node.node, bind_to_gpu_numa_node.ncpus = node, len(cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None



# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import gcutil as u
    import gridcount_b200 as g
    from gridcount_b200.api import nccl_unique_id

    cfg = CONFIGS[HEADLINE]
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    numa = bind_to_gpu_numa_node(device) if world > 1 else None
    if world > 1 and "GCB_PACK_THREADS" not in os.environ:
        # host threads that pack frames for this rank (gcb_host_staging): the CPUs this rank may run on, shared
        # with the other ranks bound to the same NUMA node (the library's own default cannot know those)
        try:
            mine = (getattr(bind_to_gpu_numa_node, "node", -1), getattr(bind_to_gpu_numa_node, "ncpus", len(os.sched_getaffinity(0))))
            every = [None] * world
            dist.all_gather_object(every, mine)
            sharing = sum(1 for e in every if e[0] == mine[0]) if mine[0] >= 0 else world
            os.environ["GCB_PACK_THREADS"] = str(max(1, min(16, mine[1] // max(1, sharing))))
        except Exception:
            pass

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        tt = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    peak, peak_src = read_peaks()
    natoms, F = cfg["natoms"], cfg["resident"]
    gindex = cfg_gindex(cfg)
    gnx = len(gindex)
    geom = g.setup_geometry(cfg["L"], cfg["delta"])
    assert geom.shape == tuple(cfg["grid"])
    packed_bytes = geom.cells                         # 8-bit lanes for 400^3
    # the memory system's ceilings for this traffic mix, measured before the counter takes the L2 set-aside
    l2 = l2_ceilings(g, packed_bytes, cfg["stride"], device) if rank == 0 else None
    dev = g.DeviceFrames(F, natoms, device).generate(cfg_gen(g, cfg), frame0=rank * F)       # 13.8 GB resident in HBM
    weights = np.ones(F, np.float32)
    T_tot = float(F * PASSES * world)

    ctr = g.GridCounter(geom, gindex, natoms, device=device)
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctr.comm_init(uid[0], rank, world)

    def step_device():
        ctr.reset_grid()
        for _ in range(PASSES):
            ctr.add_device_frames(dev, weights=weights)
        if world > 1:
            ctr.allreduce()
        ctr.density_device(T_tot)

    # ---- device-resident value -----------------------------------------------------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    ctr.sync()
    launches0 = ctr.stats().kernel_launches
    sampler = ClockSampler(device) if rank == 0 else None     # one poller per job: nvidia-smi polling perturbs launches
    barrier()
    if sampler:
        sampler.start()
    ctr.event_record(2)
    for s in range(args.steps):
        step_device()
    ctr.event_record(3)
    ctr.sync()
    barrier()
    ms_total = max_over_ranks(ctr.event_elapsed(2, 3))
    st = ctr.stats()
    launches = st.kernel_launches - launches0
    if dist is not None:
        ll = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(ll)
        launches = int(ll.item())
    units_per_step = float(gnx) * F * PASSES * world
    value = units_per_step * args.steps / (ms_total * 1e-3) * 1e-9
    # size-independent property of the full-size run: every generated atom lies in the grid, none was lost
    assert st.hits == st.atom_frames == int(gnx) * F * PASSES * world, "bench grid lost hits"
    packed_info = {"rounds": int(st.packed_rounds), "redone_by_direct_kernel": int(st.packed_redone)}

    # ---- roofline of the dominant kernel: event pairs around every launch of one more step ----------
    ctr.profile(True)
    step_device()
    roofline = roofline_of(g, ctr, cfg["stride"], geom.cells, peak, peak_src, device, HEADLINE, l2)
    ctr.profile(False)

    # ---- end to end: pinned host frames through the C ABI ----------------------------------
    HOSTF = 96                                        # distinct frames kept in pinned memory (1.15 GB)
    pin = g.PinnedFrames(HOSTF, natoms)
    for f0 in range(0, HOSTF, 32):
        pin.array[f0:f0 + 32] = dev.download(f0, 32)
    wh = np.ones(HOSTF, np.float32)
    dens_host = np.zeros(geom.shape, np.float32)
    E2E_CALLS = F // HOSTF                            # 12 calls: the resident batch's 1152 frames, once

    def step_e2e():
        ctr.reset_grid()
        for _ in range(E2E_CALLS):
            ctr.add_host_ptr(pin.ptr, HOSTF, wh)        # H2D in batches on copy streams, overlapped with K1
        if world > 1:
            ctr.allreduce()
        ctr.density_into(float(HOSTF * E2E_CALLS * world), dens_host)     # K2 + D2H of the density grid

    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    hs0 = ctr.host_staging()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    hs1 = ctr.host_staging()
    # bytes of frames that actually crossed PCIe, counted by the library where it queues the copies (this rank; the
    # other ranks stage the same way): frames DMA-ed out of the pinned buffer carry every atom, frames packed by the
    # host threads only the index group
    h2d_step = (hs1.h2d_bytes - hs0.h2d_bytes) / e2e_steps
    fr_packed = (hs1.frames_packed - hs0.frames_packed) / e2e_steps
    fr_in_place = (hs1.frames_in_place - hs0.frames_in_place) / e2e_steps
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "device-timed steps + roofline step + end-to-end steps"
    e2e_units = float(gnx) * HOSTF * E2E_CALLS * world
    e2e = {"value": e2e_units * e2e_steps / e2e_s * 1e-9, "unit": UNIT,
           "h2d_bytes_per_step": int(h2d_step) * world,
           "host_bytes_read_per_step": int(HOSTF * E2E_CALLS) * natoms * 12 * world,
           "staging": {"frames_dma_in_place_per_step": fr_in_place, "frames_packed_by_host_threads_per_step": fr_packed,
                       "pack_threads": int(hs1.threads), "pack_source_gbs": hs1.pack_bytes_per_s * 1e-9,
                       "trial": {"two_lanes_kept": int(hs1.lanes), "two_lane_frames_per_s": hs1.two_lane_frames_per_s,
                                 "dma_in_place_frames_per_s": hs1.in_place_frames_per_s,
                                 "note": "the counter's first calls (warm-up step) time both ways and keep the faster"},
                       "note": "two lanes at once: batches DMA-ed out of the pinned buffer (12 B per atom) and batches whose "
                               "index group host threads copy into pinned slots (12 B per counted atom); rank 0's split"},
           "d2h_bytes_per_step": int(geom.cells * 4 + HOSTF * E2E_CALLS * 8) * world, "steps": e2e_steps,
           "frames_per_step_per_gpu": HOSTF * E2E_CALLS, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "path": "gcb_counter_add_frames(pinned host x[F][natoms][3]) -> gcb_counter_density(host grid)"}
    pin.free()

    # ---- parity inside the bench ---------------------------------------------------------------
    parity = None
    c5 = None
    e2e_xtc_multi = None
    per_config_n = None
    if world == 1:
        parity = prefix_parity(g, cfg, geom, gindex, dev, device)
    else:
        c5, parity = c5_section(g, u, torch, dist, ctr, dev, geom, rank, world, device, barrier, max_over_ranks)
        try:
            per_config_n = per_config_multi(g, torch, dist, rank, world, device, barrier, max_over_ranks)
        except Exception as e:
            per_config_n = {"error": str(e)[:200]}
        try:
            e2e_xtc_multi = xtc_multi_section(g, torch, dist, rank, world, device, barrier, max_over_ranks)
        except Exception as e:                        # (every rank fails alike or not at all: the section is collective)
            e2e_xtc_multi = {"error": str(e)[:200]}

    # ---- a_ri3Dc's reductions over the job's density (K3), on its own line -----------------------
    k3 = None
    if rank == 0 and world == 1:
        try:
            ctr.reset_grid()
            ctr.add_device_frames(dev, weights=weights)
            k3 = k3_section(g, ctr, float(F))
        except Exception as e:
            k3 = {"error": str(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference's loop on this box's cores --------
    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        sample = 2 * threads
        xh = dev.download(0, sample)
        r1, kind = cpu_reference(4, 1, 1, xh, gindex, cfg["L"], cfg["delta"], natoms)
        rep1 = int(max(1, min(8, round(4.0 * r1 / (4 * gnx)))))
        r1, kind = cpu_reference(4, rep1, 1, xh, gindex, cfg["L"], cfg["delta"], natoms)
        cpu = None
        try:
            # all host threads: the reference arm itself, in a process of its own (the same measurement the driver
            # takes with --impl reference; inside this process, next to 14 GB of CUDA allocations, the same loop
            # measures 40 % slower)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=300, text=True)
            ref = json.loads(p.stdout.strip().splitlines()[-1])
            cpu = dict(ref["cpu_baseline"])
            cpu["single_core_value"] = r1 * 1e-9
            cpu["sample"] += "; one thread: 4 frames x %d passes, measured inside the GPU process" % rep1
        except Exception:
            cpu = None
        if cpu is None:
            rp, kind = cpu_reference(sample, 1, threads, xh, gindex, cfg["L"], cfg["delta"], natoms)
            repp = int(max(1, min(40, round(8.0 * rp / (sample * gnx)))))
            rp, kind = cpu_reference(sample, repp, threads, xh, gindex, cfg["L"], cfg["delta"], natoms)
            cpu = {"value": rp * 1e-9, "unit": UNIT, "cores": threads, "kind": kind, "single_core_value": r1 * 1e-9,
                   "sample": "frames of the workload (from the same generator) in host memory, binned into private 256 MB "
                             "grids: 1 thread x 4 frames x %d passes, and %d threads x %d frames x %d passes; grid "
                             "allocation and trajectory decode excluded" % (rep1, threads, sample, repp)}
    ctr.close()
    dev.free()

    # ---- the other named shapes (N = 1): C1, C2 (+ its a_ri3Dc and XTC sections), C4 uniform and clustered ----
    per_config = None
    xtc = None
    k3_c2 = None
    if rank == 0 and world == 1:
        per_config = {}
        for name in ("C1", "P1", "C2", "C4", "C4K"):
            c = CONFIGS[name]
            extra = None
            if name == "C2":
                def extra(cc, dd, gg, gi, c=c):
                    out = {}
                    cc.reset_grid()
                    cc.add_device_frames(dd, weights=np.ones(dd.nframes, np.float32))
                    out["a_ri3Dc"] = k3_section(g, cc, float(dd.nframes))
                    out["xtc"] = xtc_section(g, cc, dd, c, gg, gi)
                    return out
            try:
                per_config[name] = measure_config(g, name, c, device, peak, peak_src, extra=extra)
            except Exception as e:
                per_config[name] = {"error": str(e)[:300]}
        xtc = per_config.get("C2", {}).pop("xtc", None)
        k3_c2 = per_config.get("C2", {}).pop("a_ri3Dc", None)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": headline_config(world),
                "timing": {"frames_resident_gb_per_gpu": F * natoms * 12 / 1e9,
                           "l2_policy": "inputs (13.8 GB per GPU) larger than L2", "host_binding": numa,
                           "clock": "CUDA events on the launching stream, max over ranks"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "packed_counters": packed_info,
                "parity_checked": bool(parity and parity["ok"]), "parity": parity,
                "cpu_baseline": cpu, "a_ri3Dc": {"C3": k3, "C2": k3_c2} if world == 1 else None, "xtc": xtc,
                "per_config": per_config if world == 1 else per_config_n, "c5": c5, "e2e_xtc": e2e_xtc_multi if world > 1 else (xtc or {}).get("e2e_xtc")}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def xtc_multi_section(g, torch, dist, rank, world, device, barrier, max_over_ranks):
    """N > 1: the end-to-end job from XTC bytes on every rank (C2 shape, 4096 frames per rank and step): pinned
    compressed bytes -> H2D -> GPU decode-and-bin -> NCCL grid all-reduce -> density on the host."""
    from gridcount_b200.api import nccl_unique_id
    cfg = CONFIGS["C2"]
    natoms, NF, REP = cfg["natoms"], 256, 16
    geom = g.setup_geometry(cfg["L"], cfg["delta"])
    gindex = cfg_gindex(cfg)
    dev = g.DeviceFrames(NF, natoms, device).generate(cfg_gen(g, cfg), frame0=rank * NF)
    if rank == 0:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "libgcbhost.so"])
    barrier()
    data, _ = xtc_encode(dev, cfg, NF, build=False)
    dev.free()
    xb = g.XtcBuffer(data * REP, pinned=True)
    w = np.ones(xb.nframes, np.float32)
    ctr = g.GridCounter(geom, gindex, natoms, device=device)
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctr.comm_init(uid[0], rank, world)
    dens = np.zeros(geom.shape, np.float32)
    T = float(xb.nframes * world)

    def step():
        ctr.reset_grid()
        ctr.add_xtc(xb, w)
        ctr.allreduce()
        ctr.density_into(T, dens)

    step()
    barrier()
    STEPS = 3
    t0 = time.perf_counter()
    for _ in range(STEPS):
        step()
    barrier()
    secs = max_over_ranks(time.perf_counter() - t0) / STEPS
    st = ctr.stats()
    res = None
    if rank == 0:
        res = {"value": xb.nframes * len(gindex) * world / secs * 1e-9, "unit": UNIT, "n_gpus": world,
               "h2d_bytes_per_step": int(xb.nbytes) * world, "d2h_bytes_per_step": int(geom.cells * 4) * world,
               "ms_per_step": secs * 1e3, "frames_per_step": int(st.frames), "hits": int(st.hits),
               "workload": cfg["title"] + ", %d frames per rank and step as XTC (%.2f B/atom)" % (xb.nframes, len(data) / (NF * natoms)),
               "path": "gcb_counter_add_xtc(pinned xtc bytes) -> gcb_counter_allreduce -> gcb_counter_density(host grid)"}
    ctr.close(); xb.free()
    return res


def per_config_multi(g, torch, dist, rank, world, device, barrier, max_over_ranks):
    """N > 1: the other named shapes, frame-sharded like the headline -- every rank bins its own resident frames
    (10 passes = one job), the grids are combined by the NCCL all-reduce, the density is formed; `value` counts the
    atom-frames of all ranks.  (Rooflines, oracle parity and the private/packed accounting of these shapes are in the
    N = 1 line.)"""
    from gridcount_b200.api import nccl_unique_id
    out = {}
    for name in ("C1", "C2", "C4"):
        cfg = CONFIGS[name]
        geom = g.setup_geometry(cfg["L"], cfg["delta"])
        gindex = cfg_gindex(cfg)
        F, reps = cfg["resident"], 10
        dev = g.DeviceFrames(F, cfg["natoms"], device).generate(cfg_gen(g, cfg), frame0=rank * F)
        ctr = g.GridCounter(geom, gindex, cfg["natoms"], device=device)
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctr.comm_init(uid[0], rank, world)
        w = np.ones(F, np.float32)
        T = float(F * reps * world)

        def job():
            ctr.reset_grid()
            for _ in range(reps):
                if cfg.get("reset_each_pass"):
                    ctr.reset_grid()
                ctr.add_device_frames(dev, weights=w)
            ctr.allreduce()
            ctr.density_device(T)

        job()
        ctr.sync()
        barrier()
        ctr.event_record(0)
        job()
        ctr.event_record(1)
        ctr.sync()
        barrier()
        ms = max_over_ranks(ctr.event_elapsed(0, 1))
        st = ctr.stats()
        kept = 1 if cfg.get("reset_each_pass") else reps        # passes whose hits the final grid holds
        if rank == 0:
            out[name] = {"workload": cfg["title"], "n_gpus": world, "value": float(len(gindex)) * F * reps * world / (ms * 1e-3) * 1e-9,
                         "unit": UNIT, "ms_per_job": ms, "frames_per_job": F * reps * world,
                         "all_hits_in_grid": bool(st.hits == int(len(gindex)) * F * kept * world),
                         "job": "%d passes over %d resident frames per rank + NCCL grid all-reduce + density" % (reps, F)}
        ctr.close(); dev.free()
    return out if rank == 0 else None


def c5_section(g, u, torch, dist, ctr, dev, geom, rank, world, device, barrier, max_over_ranks):
    """N > 1: the C5 shape on the same counter -- every rank's frames come from a breathing box (the grid stays the
    first frame's, atoms that end up outside are dropped as the reference drops them), one 256 MB NCCL uint32
    all-reduce per job.  First a parity job: a short prefix of EVERY rank's shard is binned and all-reduced, and rank 0
    compares the reduced counts with the CPU oracle run over the same frames; then the timed jobs."""
    cfg = CONFIGS["C5"]
    natoms, F = cfg["natoms"], cfg["resident"]
    gindex = cfg_gindex(cfg)
    gnx = len(gindex)
    dev.generate(cfg_gen(g, cfg), frame0=rank * F)
    npre = cfg["prefix"]
    ctr.reset_grid()
    ctr.add_device_frames(dev, first=0, count=npre)
    ctr.allreduce()
    st = ctr.stats()
    parity = None
    if rank == 0:
        got = ctr.counts()
        want = np.zeros(geom.shape, np.uint32)
        nh = 0
        for r in range(world):
            x = u.gen_frames(cfg_gen(u, cfg), r * F, npre, natoms)
            wr, hr = oracle_counts(geom.tg, x, gindex)
            want += wr
            nh += int(hr.sum())
        parity = {"ok": bool(np.array_equal(got, want) and st.hits == nh and st.frames == npre * world),
                  "frames": npre * world, "hits": nh, "atom_frames": npre * world * gnx,
                  "against": "CPU oracle over the first %d frames of every rank's shard (breathing box): NCCL-reduced "
                             "uint32 grid, global hit count and frame count" % npre}
    else:
        ctr.counts()
    barrier()
    REPS, JOBS = 4, 3
    w = np.ones(F, np.float32)

    def job(ev=None):
        ctr.reset_grid()
        if ev is not None:
            ctr.event_record(ev)
        for _ in range(REPS):
            ctr.add_device_frames(dev, weights=w)
        if ev is not None:
            ctr.event_record(ev + 1)
        ctr.allreduce()
        if ev is not None:
            ctr.event_record(ev + 2)
        ctr.density_device(float(F * REPS * world))

    job()
    ctr.sync()
    barrier()
    ctr.event_record(2)
    for j in range(JOBS):
        job(ev=10 + 4 * j)
    ctr.event_record(3)
    ctr.sync()
    barrier()
    ms = max_over_ranks(ctr.event_elapsed(2, 3) / JOBS)
    bin_ms = max_over_ranks(float(np.mean([ctr.event_elapsed(10 + 4 * j, 11 + 4 * j) for j in range(JOBS)])))
    red_ms = max_over_ranks(float(np.mean([ctr.event_elapsed(11 + 4 * j, 12 + 4 * j) for j in range(JOBS)])))
    st = ctr.stats()
    c5 = None
    if rank == 0:
        c5 = {"workload": cfg["title"], "n_gpus": world, "frames_per_job": F * REPS * world,
              "value": float(gnx) * F * REPS * world / (ms * 1e-3) * 1e-9, "unit": UNIT, "ms_per_job": ms,
              "binning_ms": bin_ms, "allreduce_256MB_ms": red_ms, "hits": int(st.hits), "atom_frames": int(st.atom_frames),
              "note": "a job here is %d frames per rank; the named job is 125000 per rank, so the all-reduce's share is "
                      "overstated %dx" % (F * REPS, 125000 // (F * REPS)),
              "parity_checked": bool(parity["ok"])}
    return c5, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
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
