"""Builds librgbdslam_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

The CAPE translation units are compiled with -fmad=false: the reference's x86-64 build has no FMA contraction
(CMakeLists.txt:14, no -march), and the integer labels only come out bit-exact when every FP32 product / FP64 sum
rounds the same way. The pose solver keeps FMA contraction (tolerance-based parity, 1e-4 relative)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librgbdslam_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]

UNITS = [
    ("cape_cell_fit.cu", ["-fmad=false"]),
    ("cape_segment.cu", ["-fmad=false"]),
    ("rectify.cu", ["-fmad=false"]),
    ("api_cape.cu", ["-fmad=false"]),
    ("kalman.cu", ["-fmad=false"]),
    ("api_kalman.cu", []),
    ("plane_match.cu", ["-fmad=false"]),
    ("pose_solve.cu", []),
    ("pose_chain.cu", []),
    ("pose_wide.cu", []),
    ("api_pose.cu", []),
]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "rgbdslam_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    logs = []
    for name, extra in UNITS:
        src = os.path.join(CSRC, name)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, name.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append((name, r.stderr))
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + name)
    if force or _newer(objs, OUT):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed")
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        for name, log in logs:
            f.write("==== %s\n%s\n" % (name, log))
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
