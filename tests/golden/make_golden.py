#!/usr/bin/env python
"""Regenerate tests/golden/ by running the UNMODIFIED reference (oracle/_ref,
built from /root/reference by `make -C oracle ref`) on the cases of
tests/golden_cases.py.  Run from the repo root in the build container:

    python tests/golden/make_golden.py

Per case it stores: grid.dat (the XDR grid file g_ri3Dc wrote), g_stderr.txt,
inputs.sha256, and per analysis run <run>/ with every output file of a_ri3Dc,
its stdout and the full-precision capture <file>.f64 of each printed float
(oracle/shim/capture.h)."""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import gcutil as u  # noqa: E402
import golden_cases as gc  # noqa: E402


def main():
    assert u.have_ref(), "oracle/_ref missing: run `make -C oracle ref` where /root/reference exists"
    for name, c in gc.CASES.items():
        out = os.path.join(HERE, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        t, boxes, x, gindex = gc.case_inputs(name)
        with tempfile.TemporaryDirectory() as tmp:
            u.write_raw_traj(os.path.join(tmp, "traj.raw"), t, boxes, x)
            u.write_top(os.path.join(tmp, "top.stub"), c["natoms"])
            args = ["-f", "traj.raw", "-s", "top.stub", "-grid", "grid.dat", "-delta"] + list(c["delta"]) + c["gargs"]
            if c["stride"] != 1:
                u.write_ndx(os.path.join(tmp, "grp.ndx"), "OW", gindex)
                args += ["-n", "grp.ndx"]
            _, err = u.run_ref("g_ri3Dc_ref", args, tmp)
            shutil.copy(os.path.join(tmp, "grid.dat"), os.path.join(out, "grid.dat"))
            open(os.path.join(out, "g_stderr.txt"), "w").write(err)
            open(os.path.join(out, "inputs.sha256"), "w").write(gc.inputs_digest(name) + "\n")
            for run, aargs in c["analyses"].items():
                rdir = os.path.join(out, run)
                os.makedirs(rdir)
                shutil.copy(os.path.join(tmp, "grid.dat"), os.path.join(rdir, "grid.dat"))
                so, se = u.run_ref("a_ri3Dc_ref", ["-grid", "grid.dat"] + aargs, rdir, capture=True)
                os.remove(os.path.join(rdir, "grid.dat"))
                open(os.path.join(rdir, "stdout.txt"), "w").write(so)
                open(os.path.join(rdir, "stderr.txt"), "w").write(se)
        print("golden:", name, sorted(os.listdir(out)))
    # a_gridcalc
    out = os.path.join(HERE, "gridcalc")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    ins = [os.path.join(HERE, n, "grid.dat") for n in gc.GRIDCALC["inputs"]]
    u.run_ref("a_gridcalc_ref", ["-f"] + ins + ["-o", "avg.dat"], out)
    print("golden: gridcalc", os.listdir(out))


if __name__ == "__main__":
    main()
