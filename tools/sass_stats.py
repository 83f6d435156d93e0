"""Static look at one kernel's SASS: instruction count, opcode mix, and how many instructions
consume the result of the instruction right before them (a proxy for fixed-latency 'wait' stalls).

  python tools/sass_stats.py <lib.so> <kernel-name-substring> [start_line end_line source-file]
"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
blk = [b for b in blocks if pat in b.split("\n")[0]]
assert blk, "kernel not found"
ins = []
for ln in blk[0].splitlines():
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", ln)
    if m: ins.append(m.group(2).strip())
ops = collections.Counter()
dep1 = dep2 = 0
prev_dst = [None, None]
for t in ins:
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0].split(".")[0]
    ops[op] += 1
    regs = re.findall(r"\bR(\d+)\b", t2)
    dst = regs[0] if regs and op not in ("STS", "STG", "BRA", "ISETP", "FSETP", "BSSY", "BSYNC", "STL") else None
    srcs = set(regs[1:] if dst else regs)
    if prev_dst[0] in srcs: dep1 += 1
    elif prev_dst[1] in srcs: dep2 += 1
    prev_dst = [dst, prev_dst[0]]
n = len(ins)
print("%s: %d instructions; consume result of previous instr: %.1f%%, of the one before: %.1f%%" % (
    blk[0].split("\n")[0][:70], n, 100.0 * dep1 / n, 100.0 * dep2 / n))
print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(14)))
