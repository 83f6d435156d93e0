"""Builds libwebrtc_ns_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libwebrtc_ns_b200.so")
SRC = os.path.join(HERE, "csrc", "ns_capi.cu")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-enum-compare",
]


def _sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(HERE, "..", "include", "webrtc_ns_b200.h"))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB


def build_examples():
    """examples/test_ns_module: the Linux clone of the reference CLI driver; examples/apm_ns_block: the C++ mirror
    of the author's APM_NS class (include/apm_ns_b200.h) driven block by block.  Both link against the library."""
    out = []
    hdrs = [os.path.join(HERE, "..", "include", h) for h in ("webrtc_ns_b200.h", "apm_ns_b200.h")]
    for name in ("test_ns_module", "apm_ns_block"):
        src = os.path.join(HERE, "..", "examples", name + ".cpp")
        exe = os.path.join(HERE, name)
        out.append(exe)
        if os.path.exists(exe) and os.path.getmtime(exe) > max([os.path.getmtime(src), os.path.getmtime(LIB)] +
                                                               [os.path.getmtime(h) for h in hdrs]):
            continue
        cmd = ["g++", "-O2", "-std=c++11", "-Wall", "-o", exe, src, "-L" + HERE, "-lwebrtc_ns_b200", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
