"""Dynamic opcode mix of a kernel from an ncu report (source page, SASS view): warp instructions per unit by opcode.

  python tools/ncu_opcodes.py <prof.ncu-rep> <units_per_launch> [top_n]
"""
import collections, csv, re, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
hdr = None
ops = collections.Counter()
for r in csv.reader(txt.splitlines()):
    if len(r) >= 2 and "Instructions Executed" in r and "Source" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    try:
        n = int(d["Instructions Executed"])
    except ValueError:
        continue
    t = re.sub(r"^@!?U?P\d+\s+", "", d["Source"].strip())
    ops[t.split()[0].split(".")[0]] += n
tot = sum(ops.values())
print("total warp instr / unit: %.0f" % (tot / units))
for k, v in ops.most_common(top):
    print("  %-10s %8.1f  %5.1f%%" % (k, v / units, 100.0 * v / tot))
