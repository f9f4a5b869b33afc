"""CPU: the C-ABI shared library builds, loads, and exports exactly the entry points include/rgbdslam_b200.h declares.
No compute call is made here (there is no GPU in the CPU test tier; the library has no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rgbd_slam_b200 as rs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rgbdslam_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = C.CDLL(rs.lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in the header but not exported: %s" % missing


def test_struct_layouts_match_header():
    """abi.py's numpy dtypes against the sizes the C compiler gives the header's structs."""
    import subprocess
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "rgbdslam_b200.h"
int main(void){
 printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rs_cell_out), sizeof(rs_plane_out), sizeof(rs_cyl_out),
        sizeof(rs_cape_frame_info), sizeof(rs_match), sizeof(rs_pose_out), sizeof(rs_pose_opts), sizeof(rs_cape_outputs));
 printf("%zu %zu %zu\n", offsetof(rs_pose_out, pose), offsetof(rs_cyl_out, kept), offsetof(rs_plane_out, n_boundary));
 return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(v) for v in out]
    a = rs.abi
    assert sizes[:6] == [a.cell_dtype.itemsize, a.plane_dtype.itemsize, a.cyl_dtype.itemsize, a.info_dtype.itemsize,
                         a.match_dtype.itemsize, a.pose_out_dtype.itemsize]
    assert sizes[6] == C.sizeof(a.PoseOpts) and sizes[7] == C.sizeof(a.CapeOutputs)
    assert sizes[8] == a.pose_out_dtype.fields["pose"][1]
    assert sizes[9] == a.cyl_dtype.fields["kept"][1]
    assert sizes[10] == a.plane_dtype.fields["n_boundary"][1]


def test_no_device_means_loud_failure():
    """Without a GPU the product must refuse to run (never fall back to a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rs.RsError) as e:
        rs.PrimitiveDetection(640, 480, 20)
    assert "no CUDA device" in str(e.value) or "CPU" in str(e.value)
    with pytest.raises(rs.RsError):
        rs.PoseOptimization(max_batch=1, max_matches=16)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rgb-d-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text, f
                assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), f


def test_synthetic_inputs_are_deterministic():
    a = rs.synth.scene_v0_depth(5)
    b = rs.synth.scene_v0_depth(5)
    assert a.dtype == np.float32 and a.shape == (480, 640) and np.array_equal(a, b)
    assert 0.015 < (a == 0).mean() < 0.03
    t, g, m = rs.synth.pose_correspondences(3)
    assert len(m) == 320 and (m["type"] == rs.abi.RS_FEAT_PLANE).sum() == 20
