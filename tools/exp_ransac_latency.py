"""Experiment: where does the latency of pose_ransac go? Times the pose kernels of a 256-frame batch for a few option sets."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rs = importlib.import_module("rgb-d-slam_b200")
F, M = 256, 320
truth, cur, matches, n = rs.synth.pose_batch(0, F, M)
sol = rs.PoseOptimization(max_batch=F, max_matches=M, max_iterations=1024, max_variance=100)
sol.upload(cur, matches, n)
s = torch.cuda.current_stream().cuda_stream
for name, kw in [("default 119 hyp", {}), ("1 hypothesis", dict(max_iterations=1)), ("4 hypotheses", dict(max_iterations=4)),
                 ("8 hypotheses", dict(max_iterations=8)), ("16 hypotheses", dict(max_iterations=16)),
                 ("119 hyp, maxfev 40", dict(lm_max_fev=40)), ("119 hyp, maxfev 80", dict(lm_max_fev=80)),
                 ("119 hyp, maxfev 160", dict(lm_max_fev=160)), ("no variance", dict(n_variance=0))]:
    opts = sol.options(seed=1234, rng_mode=rs.abi.RS_RNG_DEVICE, **kw)
    for _ in range(3):
        sol.solve_device(F, opts, stream=s)
    torch.cuda.synchronize()
    sol.set_timing(10)
    for _ in range(10):
        sol.solve_device(F, opts, stream=s)
    torch.cuda.synchronize()
    ms = np.mean([sol.kernel_ms(i) for i in range(10)], axis=0)
    out, _ = sol.download(F)
    sol.set_timing(0)
    print("%-22s prepare %.3f ransac %.3f variance %.3f cov %.3f | ok %.2f iterations_run mean %.1f max %d, best_iteration mean %.1f" % (
        name, ms[0], ms[1], ms[2], ms[3], (out["status"] == 1).mean(), out["iterations_run"].mean(), out["iterations_run"].max(), out["best_iteration"].mean()))
