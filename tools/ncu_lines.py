"""Per-source-line instruction / stall-sample table from an ncu report (source page, cuda,sass view).

  python tools/ncu_lines.py <prof.ncu-rep> <units_per_launch> [min_instr_per_unit]
"""
import csv, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 8
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; hdr = None; out = []
for r in csv.reader(txt.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try: n = int(d["Instructions Executed"])
        except ValueError: continue
        out.append((cur, int(r[0]), n / units, int(d["# Samples"] or 0),
                    int(d.get("L1 Wavefronts Shared Excessive", "0") or 0) / units, r[1]))
tot = sum(o[2] for o in out); ts = sum(o[3] for o in out)
print("total warp instr / unit: %.0f   samples %d" % (tot, ts))
for f, ln, n, s, ex, src in out:
    if n >= thr or s >= ts * 0.005:
        print("%-16s %4d  %7.1f instr  %5.2f%% smp  %5.1f xs-wf  %s" % (f[:16], ln, n, 100.0 * s / ts, ex, src.strip()[:90]))
