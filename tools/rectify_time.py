"""Times rectify_depth alone (memset + rectify_scatter_kernel + rectify_resolve_kernel per frame group) on 256 resident 640x480
frames, identity and an offset extrinsic: 10 calls between two CUDA events after 2 warm-up calls.
Usage (GPU box): python tools/rectify_time.py [frames]; under ncu for the per-kernel split (profiles/r02_rectify_*)."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch, rgbd_slam_b200 as rs
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
depth = np.stack([rs.synth.scene_v0_depth(i) for i in range(32)]); depth = np.tile(depth, (F // 32, 1, 1))
det = rs.PrimitiveDetection(640, 480, 20, max_batch=F)
d = torch.from_numpy(depth).cuda(); out = torch.empty_like(d)
s = torch.cuda.current_stream().cuda_stream
T = np.eye(4); T[:3, 3] = (25.0, -3.0, 4.0)
R = np.eye(4); c, sn = np.cos(0.02), np.sin(0.02)
R[:3, :3] = np.array([[c, 0, sn], [0, 1, 0], [-sn, 0, c]]); R[:3, 3] = (25.0, -3.0, 4.0)
for name, ext in (("identity", None), ("offset 25 mm", T), ("rotated 0.02 rad + offset", R)):
    det.set_rectification(ext, enable=True)
    for _ in range(2): det.rectify_device(d.data_ptr(), F, out.data_ptr(), stream=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): det.rectify_device(d.data_ptr(), F, out.data_ptr(), stream=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("rectify_depth %s: %.4f ms per %d frames = %.0f GB/s of the algorithmic 8 B/pixel" % (name, ms, F, 8 * 640 * 480 * F / ms / 1e6))
