"""Experiment: L independent lanes (context pair + stream pair each), steps dealt round-robin, no cross-lane sync.
Usage (GPU box): python tools/exp_two_lanes.py [lanes] [steps] [frames]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rs = importlib.import_module("rgb-d-slam_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
F = int(sys.argv[3]) if len(sys.argv) > 3 else 256
M = 320
depth = rs.synth.scene_v0_batch(0, 8)
depth = np.concatenate([depth] * (F // 8))
truth, cur, matches, n = rs.synth.pose_batch(0, 8, M)
cur = np.concatenate([cur] * (F // 8)); matches = np.concatenate([matches] * (F // 8)); n = np.concatenate([n] * (F // 8))
lanes = []
for l in range(L):
    det = rs.PrimitiveDetection(640, 480, 20, max_batch=F)
    sol = rs.PoseOptimization(max_batch=F, max_matches=M, max_iterations=119, max_variance=100)
    opts = sol.options(seed=1234, rng_mode=rs.abi.RS_RNG_DEVICE)
    sol.upload(cur, matches, n)
    d = torch.from_numpy(depth).cuda()
    s, p = torch.cuda.Stream(), torch.cuda.Stream()
    lanes.append((det, sol, opts, d, s, p))
def step(i):
    det, sol, opts, d, s, p = lanes[i % L]
    det.run_device(d.data_ptr(), F, seed=0, stream=s.cuda_stream)
    det.stream_wait_fit(p.cuda_stream)
    sol.solve_device(F, opts, stream=p.cuda_stream)
    s.wait_stream(p)
for i in range(2 * L + 2):
    step(i)
torch.cuda.synchronize()
main = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(main)
for l in lanes:
    l[4].wait_stream(main)
for i in range(steps):
    step(i)
for l in lanes:
    main.wait_stream(l[4])
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("lanes %d frames %d: %.3f ms/step, %.1f k frames/s" % (L, F, ms / steps, F * steps / ms))
