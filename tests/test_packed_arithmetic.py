"""The float kernels use Blackwell's packed fp32 instructions (FADD2 / FMUL2 / FFMA2) on slot pairs and complex
pairs.  ptxas contracts a packed product into a following packed sum whatever --fmad says, which would break the
bit-for-bit recursion, so the sources route such products through fma(a, b, -0) (ns_warp.cuh vmul_o).  This is the
build-time audit: every FFMA2 in the library's SASS must come from an explicit fma in the source (or multiply by a
power of two, where contraction is exact).  The device-side counterpart is WebRtcNsB200_SelfTest."""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvdisasm") is None or shutil.which("cuobjdump") is None, reason="CUDA binary utilities not installed")
def test_no_contracted_packed_multiply_add(capsys):
    import __graft_entry__ as g
    g.build()
    spec = importlib.util.spec_from_file_location("check_packed_fusion", os.path.join(ROOT, "tools", "check_packed_fusion.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import sys
    argv, sys.argv = sys.argv, ["check_packed_fusion.py"]
    try:
        rc = mod.main()
    finally:
        sys.argv = argv
    out = capsys.readouterr().out
    assert rc == 0, out
    # and the packed instructions are really there: the headline kernel
    row = [l for l in out.splitlines() if "nsf_process_kernel<256, 1, true, false>" in l]
    assert row, out
    ffma2, fmul2, fadd2 = (int(x) for x in row[0].split("|")[2:5])
    assert ffma2 > 50 and fadd2 > 50 and fmul2 > 20, row[0]
