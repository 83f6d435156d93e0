"""Build-time audit of the packed fp32 arithmetic (f32x2) in the float kernels.

ptxas 12.9 contracts mul.rn.f32x2 -> add.rn.f32x2 into one FFMA2 even under --fmad false (and sees through
fma(a, b, -0)), which would break the bit-for-bit recursion.  The kernels therefore route
every product that feeds a sum through scalar additions (ns_warp.cuh vmadd / vmmadd / cmul_parts), which ptxas
leaves alone.  This script proves none slipped through: with -lineinfo every SASS instruction carries the source
line it came from, so every FFMA2 must come from the __ffma2_rn intrinsic itself; one attributed to __fadd2_rn or
__fmul2_rn is a contraction.  Contractions whose multiplier is a literal power of two are exact (0.5 * x + y has
the bits of the two-step form) and are listed as such.

  python tools/check_packed_fusion.py [out.md]        (needs cuobjdump / nvdisasm; seconds, no GPU)
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audiosignalprocess_b200", "libwebrtc_ns_b200.so")
HDR = "/usr/local/cuda/targets/x86_64-linux/include/crt/sm_100_rt.hpp"


def intrinsic_lines():
    """line of sm_100_rt.hpp -> intrinsic whose body it is"""
    out, cur = {}, None
    for i, ln in enumerate(open(HDR), 1):
        m = re.search(r"float2 (__f\w+2_r[nzdu])\(", ln)
        if m and "impl" not in ln:
            cur = m.group(1)
        if cur:
            out[i] = cur
        if ln.startswith("}"):
            cur = None
    return out


def pow2(tok):
    try:
        v = abs(float(tok))
    except ValueError:
        return False
    return v > 0 and (v.hex().startswith("0x1.0000000000000p"))


def main():
    lines_of = intrinsic_lines()
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, check=True, capture_output=True)
        cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True, check=True).stdout
    fn, cur = None, None
    stats = collections.defaultdict(collections.Counter)
    bad = []
    for ln in dis.splitlines():
        m = re.search(r"\.section\s+\.text\.(\w+)", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = lines_of.get(int(m.group(2))) if m.group(1).endswith("sm_100_rt.hpp") else "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
            continue
        m = re.search(r"\b(FFMA2|FMUL2|FADD2)\s+(.*?);", ln)
        if not m or fn is None:
            continue
        op = m.group(1)
        stats[fn][op] += 1
        if op == "FFMA2" and cur != "__ffma2_rn":
            ops = [t.strip() for t in m.group(2).split(",")]
            if any(pow2(t) for t in ops[1:3]):
                stats[fn]["exact"] += 1
            else:
                stats[fn]["contracted"] += 1
                bad.append((fn, cur, ln.strip()))
    rows = ["| kernel | FFMA2 | FMUL2 | FADD2 | FFMA2 not from `__ffma2_rn`: by a power of two (exact) | other (contraction) |",
            "|---|---:|---:|---:|---:|---:|"]
    for k in sorted(stats):
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        c = stats[k]
        rows.append("| `%s` | %d | %d | %d | %d | %d |" % (name, c["FFMA2"], c["FMUL2"], c["FADD2"], c["exact"], c["contracted"]))
    text = "\n".join(rows)
    print(text)
    for b in bad:
        print("CONTRACTED:", *b)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write("Packed fp32 instructions per kernel (static SASS) and the contraction audit of "
                                     "`tools/check_packed_fusion.py`.\n\n" + text + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
