#!/usr/bin/env python
"""Condensed view of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_x.log"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("headline %.1f %s on %d GPU(s), %.2f ms/step; e2e %.3f; kernel %.1f G/s frac %.3f; parity %s; launches %s" % (
    d["value"], d["unit"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_g_atom_frames_per_s", 0), r.get("frac", 0),
    d.get("parity_checked"), d.get("gpu_launches")))
if r.get("l2"):
    print("  l2 ceilings:", {k: round(v, 2) if isinstance(v, float) else v for k, v in r["l2"].items()})
if d.get("cpu_baseline"):
    print("  cpu: %.3f G/s on %d threads (%s), 1 thread %.4f" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"],
                                                                 d["cpu_baseline"]["kind"], d["cpu_baseline"].get("single_core_value", 0)))
for n, c in (d.get("per_config") or {}).items():
    if "error" in c:
        print("  %s ERROR %s" % (n, c["error"]))
        continue
    if "roofline" not in c:
        print("  %-4s %7.1f G/s on %d GPUs, %.2f ms per job, all hits in grid: %s" % (n, c["value"], c["n_gpus"], c["ms_per_job"], c["all_hits_in_grid"]))
        continue
    rl = c["roofline"]
    print("  %-4s %6.1f G/s  kernel %6.1f frac %.3f  [%s]  kinds %s  parity %s  packed %d/%d private %d/%d  of stream+red %.2f of stream %.2f" % (
        n, c["value"], rl["kernel_g_atom_frames_per_s"], rl["frac"], rl["kernel"][:32], rl.get("k1_launch_kinds_ms"), c["parity_checked"],
        c["packed_rounds"], c["packed_redone"], c.get("private_launches", 0), c.get("private_redone", 0),
        rl["l2"]["frac_of_stream_plus_red"], rl["l2"].get("frac_of_stream_alone", 0)))
for k in ("a_ri3Dc", "xtc", "c5"):
    if d.get(k):
        print("  %s: %s" % (k, json.dumps(d[k])[:900]))
