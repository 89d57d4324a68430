#!/usr/bin/env python
"""Condense the ncu exports that tools/gpu_profile.sh leaves in gpurun_out/ into the small, tracked
files under profiles/:  python tools/summarize_ncu.py TAG "COMMAND that was profiled" [CONFIG atom_frames_per_launch]
  profiles/launches_TAG.csv      the launch list of the command (gpu__time_duration.sum)
  profiles/launches_TAG.md       per-kernel totals and shares of that run
  profiles/TAG_summary.json      selected `ncu --set full` metrics of the captured kernel (one launch)
  profiles/TAG_hot_sass.txt      the 40 SASS lines with most stall samples
  profiles/k1_traffic.json       with CONFIG given: DRAM bytes per launch of that config's dominant kernel (read by bench.py)"""
import csv
import json
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "lts__t_requests.sum",
        "lts__t_requests_srcunit_tex_op_red.sum", "lts__t_requests_srcunit_ltcfabric.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum",
        "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_red_lookup_hit.sum",
        "lts__t_sectors_op_red_lookup_miss.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    tag = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
    config = sys.argv[3] if len(sys.argv) > 3 else None
    per_launch = float(sys.argv[4]) if len(sys.argv) > 4 else None
    os.makedirs(OUT, exist_ok=True)
    # ---- launch list
    lp = os.path.join(SRC, "launches_%s.csv" % tag)
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(OUT, "launches_%s.csv" % tag))
        rows = [r for r in csv.reader(open(lp)) if len(r) > 14 and r[0].isdigit()]
        tot = defaultdict(lambda: [0, 0.0])
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            tot[name][0] += 1
            tot[name][1] += float(r[14]) * 1e-6
        allms = sum(v[1] for v in tot.values())
        with open(os.path.join(OUT, "launches_%s.md" % tag), "w") as f:
            f.write("# ncu launch list, `%s` (%s)\n\n" % (cmd, tag))
            f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache\n"
                    "and serialised, so read the SHARES.  The frame generator is outside every timed region but inside\n"
                    "this list.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
            for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                f.write("| `%s` | %d | %.3f | %.1f %% |\n" % (name, n, ms, 100 * ms / allms))
    # ---- full capture of the dominant kernel
    rp = os.path.join(SRC, "raw_%s.csv" % tag)
    if os.path.exists(rp):
        rows = list(csv.reader(open(rp)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(hdr)}
        summ = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None, "tag": tag,
                "how": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 " + cmd}
        for k in KEYS:
            if k in col:
                try:
                    summ[k] = {"value": float(vals[col[k]]), "unit": units[col[k]]}
                except ValueError:
                    summ[k] = {"value": vals[col[k]], "unit": units[col[k]]}
        json.dump(summ, open(os.path.join(OUT, "%s_summary.json" % tag), "w"), indent=1)
        rd, wr = summ.get("dram__bytes_read.sum"), summ.get("dram__bytes_write.sum")
        if rd and wr and config:
            traffic = rd["value"] * TO_BYTES[rd["unit"]] + wr["value"] * TO_BYTES[wr["unit"]]
            tj = os.path.join(OUT, "k1_traffic.json")
            old = json.load(open(tj)) if os.path.exists(tj) else {}
            old = {k: v for k, v in old.items() if isinstance(v, dict)}        # drop the round-1 flat layout
            old[config] = {"dram_bytes_per_launch": traffic, "atom_frames_per_launch": per_launch,
                           "source": "profiles/%s_summary.json (ncu --set full, one launch)" % tag,
                           "kernel": summ["kernel"].split("(")[0] if summ["kernel"] else None}
            json.dump(old, open(tj, "w"), indent=1)
    sp = os.path.join(SRC, "src_%s.csv" % tag)
    if os.path.exists(sp):
        rows = list(csv.reader(open(sp)))
        data = [r for r in rows[2:] if len(r) > 6 and r[2].isdigit()]
        tot = sum(int(r[2]) for r in data)
        ti = sum(int(r[5]) for r in data)
        with open(os.path.join(OUT, "%s_hot_sass.txt" % tag), "w") as f:
            f.write("%s\n%d stall samples, %d warp instructions executed in this launch\n\nsamples  warp-instr  SASS\n"
                    % (rows[0][1], tot, ti))
            for r in sorted(data, key=lambda r: -int(r[2]))[:40]:
                f.write("%7s %11s  %s\n" % (r[2], r[5], r[1].strip()))
    print("profiles/ updated for", tag)


if __name__ == "__main__":
    main()
