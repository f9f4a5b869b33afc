"""Per-frame timeline of the fused pose kernel (rs_pose_debug_frame_times) on the bench workload, solve kernel alone."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rgbd_slam_b200 as rs

F, M = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 320
hyp = int(sys.argv[2]) if len(sys.argv) > 2 else 119
outl = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
truth, cur, m, n = rs.synth.pose_batch(0, F, M, n_points=300, n_planes=20, outlier_frac=outl)
solver = rs.PoseOptimization(max_batch=F, max_matches=M, max_iterations=hyp, max_variance=100)
solver.upload(cur, m, n)
settings = [dict(), dict(RS_POSE_MC_CAP="0"), dict(RS_POSE_HELP_MIN="1000"), dict(RS_POSE_MC_CAP="0", RS_POSE_HELP_MIN="1000")]
if os.environ.get("EXP_SETTINGS"):
    settings = [dict(kv.split("=") for kv in grp.split(",") if kv) for grp in os.environ["EXP_SETTINGS"].split(";")]
for env in settings:
  for k in ("RS_POSE_MC_CAP", "RS_POSE_HELP_MIN"):
    os.environ.pop(k, None)
  os.environ.update(env)
  print("==== settings", env)
  for nv in ((100, 0) if not env else (100,)):
      opts = solver.options(max_iterations=hyp, n_variance=nv, seed=1234, rng_mode=rs.abi.RS_RNG_DEVICE,
                              worker_ctas_per_sm=int(os.environ.get("EXP_CTAS", "0")))
      s = torch.cuda.Stream()
      for _ in range(3):
          solver.solve_device(F, opts, stream=s.cuda_stream)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(s)
      for _ in range(10):
          solver.solve_device(F, opts, stream=s.cuda_stream)
      e1.record(s)
      torch.cuda.synchronize()
      t = solver.frame_times(F)
      print("n_variance", nv, "ms per solve (events, incl. prepare):", e0.elapsed_time(e1) / 10, "phase_ms", solver.phase_ms(), solver.work_counters())
      for k, name in enumerate(("hyp start", "stage closed", "final done", "cov done")):
          v = t[:, k][t[:, k] >= 0]
          if len(v):
              print("  %-13s min %.3f  p50 %.3f  p90 %.3f  p99 %.3f  max %.3f" % (name, v.min(), np.percentile(v, 50), np.percentile(v, 90), np.percentile(v, 99), v.max()))
      d1 = t[:, 1] - t[:, 0]; d2 = t[:, 2] - t[:, 1]; d3 = t[:, 3] - t[:, 2]
      print("  hyp stage  p50 %.3f p90 %.3f max %.3f | final LM p50 %.3f p90 %.3f max %.3f | MC+cov p50 %.3f max %.3f" % (
          np.percentile(d1, 50), np.percentile(d1, 90), d1.max(), np.percentile(d2, 50), np.percentile(d2, 90), d2.max(), np.percentile(d3, 50), d3.max()))
      out, _ = solver.download(F)
      print("  iterations_run: mean %.2f max %d" % (out["iterations_run"].mean(), out["iterations_run"].max()))
