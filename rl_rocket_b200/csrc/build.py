"""In-tree nvcc build of librlrocket.so for sm_100a (no torch / pybind dependency: plain C ABI).

    python -m rl_rocket_b200.csrc.build [--force]

kernels_exact.cu is compiled with -fmad=false (op-for-op IEEE sequence of the reference);
kernels_fast.cu with FMA contraction.  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_C")
_V = os.environ.get("RLR_VARIANT", "A")
OUT = os.path.join(OUT_DIR, "librlrocket.so" if _V == "A" else f"librlrocket_{_V}.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
VARIANT = os.environ.get("RLR_DEFS", "").split()     # tuning builds: RLR_VARIANT=<tag> RLR_DEFS="-DRLR_REGS6=144 ..." -> librlrocket_<tag>.so
COMMON = VARIANT + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE]
UNITS = [("kernels_exact.cu", ["-fmad=false"]), ("kernels_fast.cu", ["-fmad=true"]), ("fma_peak.cu", []), ("abi.cu", []), ("host_pipe.cu", []), ("exchange.cu", [])]


def _sources():
    return [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))] + [os.path.join(ROOT, "include", "rl_rocket_b200.h")]


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in _sources()):
        return OUT
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(OUT_DIR, src.replace(".cu", f".{_V}.o"))
        cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            print(" ".join(cmd))
            print(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([NVCC, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
