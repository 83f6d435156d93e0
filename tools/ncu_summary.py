"""Turns ncu reports brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md>     per-kernel launch list / shares
  python tools/ncu_summary.py kernel <prof.ncu-rep> <out.md> <out.json> <units_per_launch>
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print(open(out).read())


def kernel(rep, out_md, out_json, units):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit = rows[0], rows[1]
    data = rows[2:]
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i], unit[i])
        res.append(d)
    d = res[-1]
    traffic = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
    inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
    summ = {"kernel": d["kernel"], "launches_captured": len(res), "dram_bytes_per_launch": traffic,
            "units_per_launch": units, "dram_bytes_per_unit": traffic / units,
            "warp_instructions_per_unit": inst / units,
            "duration": d["gpu__time_duration.sum"]}
    json.dump(summ, open(out_json, "w"), indent=1)
    with open(out_md, "w") as f:
        f.write("`%s` (last of %d captured launches; %d stream-frames per launch)\n\n| metric | value |\n|---|---|\n" % (
            d["kernel"], len(res), units))
        for k in KEYS:
            if k in d:
                f.write("| %s | %s %s |\n" % (k, d[k][0], d[k][1]))
        f.write("| DRAM traffic per launch (read+write) | %.1f MB = %.0f B per stream-frame |\n" % (traffic / 1e6, traffic / units))
        f.write("| warp instructions per stream-frame | %.0f |\n" % (inst / units))
    print(open(out_md).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]))
