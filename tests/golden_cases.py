"""Definitions of the golden parity cases.  The inputs are regenerated from the
counter-based generator of the oracle (and guarded by a stored SHA-256); the
outputs under tests/golden/<case>/ were produced by the UNMODIFIED reference
binaries (oracle/_ref) through tests/golden/make_golden.py."""
import hashlib

import numpy as np

import gcutil as u

CASES = {
    # cubic box, dt = 1 ps: float accumulation is exact integer arithmetic
    "cube": dict(L=(3.0, 3.0, 3.0), natoms=300, stride=3, nframes=20, t0=0.0, dt=1.0, seed=11,
                 gen=dict(mode=0), delta=(0.25, 0.25, 0.25), gargs=[],
                 analyses={"spc": ["-unit", "SPC", "-xzp", "-yzp", "-hxyp", "-hrzp", "-hrdf", "-plt", "-dump"],
                           "default": []}),
    # second data set for a_gridcalc
    "cube2": dict(L=(3.0, 3.0, 3.0), natoms=300, stride=3, nframes=10, t0=100.0, dt=2.0, seed=12,
                  gen=dict(mode=0), delta=(0.25, 0.25, 0.25), gargs=[], analyses={}),
    # non-commensurate box (origin shifts, atoms outside the grid are dropped),
    # all atoms counted (no index file), f32-jittered times => varying dt
    "shifted": dict(L=(2.98, 3.31, 4.07), natoms=240, stride=1, nframes=16, t0=10.0, dt=0.1, seed=13,
                    gen=dict(mode=0), delta=(0.2, 0.2, 0.25), gargs=[],
                    analyses={"voxel": ["-unit", "Voxel", "-minocc", "0.01", "-xzp", "-yzp"],
                              "default": []}),
    # user-set cylinder, pore-shaped density, -nodtweight, subtitle
    "pore": dict(L=(4.0, 4.0, 6.7), natoms=900, stride=3, nframes=24, t0=0.0, dt=2.0, seed=14,
                 gen=dict(mode=2, slab=(2.2, 4.5), pore_r=0.8), delta=(0.2, 0.2, 0.2),
                 gargs=["-R", "1.5", "-z1", "1.0", "-z2", "5.5", "-cpoint", "2.05", "1.95", "3.0",
                        "-nodtweight", "-subtitle", "pore test"],
                 analyses={"spc": ["-unit", "SPC", "-xzp", "-yzp", "-hxyp", "-hrzp", "-hrdf"],
                           "molar_crop": ["-unit", "molar", "-R", "1.0", "-z1", "2.0", "-z2", "5.0",
                                          "-nomirror", "-minocc", "0.5", "-hxyp", "-hrdf", "-hrzp", "-plt"],
                           "ang_ba": ["-unit", "Angstrom", "-radnr", "4", "-radstep", "2"]}),
}

GRIDCALC = dict(inputs=["cube", "cube2"], subtitle=None)

SETMASK = {"-cpoint": u.SET_CPOINT, "-R": u.SET_R, "-z1": u.SET_Z1, "-z2": u.SET_Z2}


def case_inputs(name):
    """(times f32[F], boxes f32[F,9], x f32[F,natoms,3], gindex int32[gnx])"""
    c = CASES[name]
    gen = u.make_gen(c["seed"], c["L"], **c["gen"])
    x = u.gen_frames(gen, 0, c["nframes"], c["natoms"])
    t = (np.float32(c["t0"]) + np.arange(c["nframes"], dtype=np.float32) * np.float32(c["dt"])).astype(np.float32)
    boxes = np.tile(u.box9(c["L"]), (c["nframes"], 1))
    gindex = np.arange(0, c["natoms"], c["stride"], dtype=np.int32)
    return t, boxes, x, gindex


def inputs_digest(name):
    t, boxes, x, gindex = case_inputs(name)
    h = hashlib.sha256()
    for a in (t, boxes, x, gindex):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def parse_gargs(gargs):
    """-> dict(setmask, cpoint, R, z1, z2, dtweight, subtitle) from g_ri3Dc style options"""
    out = dict(setmask=0, cpoint=(0.0, 0.0, 0.0), R=0.0, z1=0.0, z2=0.0, dtweight=1, subtitle="")
    i = 0
    while i < len(gargs):
        a = gargs[i]
        if a == "-cpoint":
            out["cpoint"] = tuple(float(v) for v in gargs[i + 1:i + 4]); out["setmask"] |= u.SET_CPOINT; i += 4
        elif a == "-R":
            out["R"] = float(gargs[i + 1]); out["setmask"] |= u.SET_R; i += 2
        elif a == "-z1":
            out["z1"] = float(gargs[i + 1]); out["setmask"] |= u.SET_Z1; i += 2
        elif a == "-z2":
            out["z2"] = float(gargs[i + 1]); out["setmask"] |= u.SET_Z2; i += 2
        elif a == "-nodtweight":
            out["dtweight"] = 0; i += 1
        elif a == "-subtitle":
            out["subtitle"] = gargs[i + 1]; i += 2
        else:
            raise ValueError(a)
    return out


def parse_aargs(aargs):
    out = dict(unit="unity", setmask=0, R=0.0, z1=0.0, z2=0.0, mirror=True, minocc=0.0, radnr=0, radstep=1,
               files=set())
    i = 0
    while i < len(aargs):
        a = aargs[i]
        if a == "-unit":
            out["unit"] = aargs[i + 1]; i += 2
        elif a in ("-R", "-z1", "-z2"):
            out[a[1:]] = float(aargs[i + 1]); out["setmask"] |= SETMASK[a]; i += 2
        elif a == "-nomirror":
            out["mirror"] = False; i += 1
        elif a == "-minocc":
            out["minocc"] = float(aargs[i + 1]); i += 2
        elif a == "-radnr":
            out["radnr"] = int(aargs[i + 1]); i += 2
        elif a == "-radstep":
            out["radstep"] = int(aargs[i + 1]); i += 2
        elif a in ("-xzp", "-yzp", "-hxyp", "-hrzp", "-hrdf", "-plt", "-dump"):
            out["files"].add(a[1:]); i += 1
        else:
            raise ValueError(a)
    return out
