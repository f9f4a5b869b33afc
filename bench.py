#!/usr/bin/env python
"""Benchmark of the RGB-D per-frame hot path (CAPE plane/cylinder segmentation + RANSAC/LM pose solve).

  python bench.py --gpus N --steps K --warmup W            # this framework on N B200s (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU path (restated oracle) on the host cores
  python bench.py --config 5 ...                           # BASELINE configs[4]: 1280x960 / 40 px cells / 1024 hypotheses

A "step" = one pass of the hot path over one batch of synthetic RGB-D frames per GPU. The headline workload is BASELINE.json
configs[2] on configs[3]'s batch (640x480, 20 px cells, 300-point / 20-plane RANSAC-LM solve with its 100-sample
Monte-Carlo covariance, 256 frames per GPU per step); `--config 5` times configs[4] as the headline instead, and a default
run carries a shorter measurement of it under the `config5` key. Frames shard across ranks with no data-path collective;
the single collective is the all-gather of the per-frame poses (weak scaling: frames per GPU fixed).
Prints ONE JSON line on rank 0 (see README / DESIGN.md "Measurement")."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, N_PLANES = 300, 20
MAX_MATCHES = N_POINTS + N_PLANES


class Workload:
    """Geometry + solver sizes of one BASELINE config."""

    def __init__(self, name, width, height, cell, hypotheses, frames_per_gpu, outlier_frac=0.1):
        self.name, self.W, self.H, self.cell, self.hypotheses, self.F = name, width, height, cell, hypotheses, frames_per_gpu
        # share of wrong matches among the correspondences. The reference's RANSAC stops as soon as (from the fourth
        # hypothesis on) one has more than 80 % inliers: with 10 % outliers that is after 4-7 hypotheses whatever the cap is,
        # so the "1024 hypotheses per frame" config is given 30 % outliers - the early stop can then never fire and every
        # frame evaluates all 1024 hypotheses (check.mean_ransac_iterations in the output says what ran).
        self.outlier_frac = outlier_frac
        self.scale = width / 640.0
        self.K = (550.0 * self.scale, 550.0 * self.scale, 320.0 * self.scale, 240.0 * self.scale)
        self.n_cells = (width // cell) * (height // cell)
        # depth read once + one 160-byte record per cell
        self.k1_bytes_per_frame = 4 * width * height + 160 * self.n_cells

    def config(self):
        """The `config` object of the JSON line: the workload only, identical for both arms (--impl b200 / reference)."""
        return {
            "workload": "%dx%d full frame: CAPE plane+cylinder extraction (%d px cells) + %d-point/%d-plane RANSAC-LM pose solve "
                        "(%d hypotheses) + 100-sample covariance (%s), batch of %d frames per GPU per step"
                        % (self.W, self.H, self.cell, N_POINTS, N_PLANES, self.hypotheses, self.name, self.F),
            "width": self.W, "height": self.H, "cell_px": self.cell, "cells_per_frame": self.n_cells,
            "points": N_POINTS, "planes": N_PLANES, "outlier_fraction": self.outlier_frac,
            "ransac_hypotheses": self.hypotheses, "n_variance": 100,
            "frames_per_gpu": self.F,
            "l2": "inputs larger than L2 (%.0f MB of depth per step per GPU vs 126 MB)" % (4e-6 * self.W * self.H * self.F),
        }


def workload_for(args):
    if args.config == 5:
        wl = Workload("BASELINE configs[4]", 1280, 960, 40, 1024, args.frames_per_gpu or 64, outlier_frac=0.3)
    else:
        wl = Workload("BASELINE configs[2] on configs[3]'s batch", 640, 480, 20, 119, args.frames_per_gpu or 256)
    if args.width or args.height or args.cell or args.hypotheses:
        wl = Workload("custom", args.width or wl.W, args.height or wl.H, args.cell or wl.cell, args.hypotheses or wl.hypotheses, wl.F,
                      wl.outlier_frac)
    return wl


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", type=int, default=4, choices=[4, 5],
                   help="BASELINE.md §4 row: 4 = 640x480 / 20 px / 119 hypotheses / 256 frames per GPU (default), "
                        "5 = 1280x960 / 40 px / 1024 hypotheses / 64 frames per GPU")
    p.add_argument("--width", type=int, default=0)
    p.add_argument("--height", type=int, default=0)
    p.add_argument("--cell", type=int, default=0)
    p.add_argument("--hypotheses", type=int, default=0)
    p.add_argument("--frames-per-gpu", type=int, default=0)
    p.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 10))")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--solver", default="auto", choices=["auto", "chain", "fused", "wide"],
                   help="rs_pose_opts.solver of the timed step: the three-launch chain, the fused persistent kernel, or by shape")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-e2e-lanes", action="store_true", help="skip the two-batches-in-flight variant of the host-buffer leg")
    p.add_argument("--no-config5", action="store_true", help="skip the secondary configs[4] measurement of a default run")
    p.add_argument("--no-extras", action="store_true", help="skip the informational legs (single frame, rectify, Kalman, lanes)")
    return p.parse_args()


def make_inputs(wl, first_frame, n_frames):
    import rgbd_slam_b200 as rs
    depth = np.empty((n_frames, wl.H, wl.W), dtype=np.float32)
    for i in range(n_frames):
        depth[i] = rs.synth.scene_v0_depth(first_frame + i, wl.W, wl.H)
    truth, cur, matches, n = pose_inputs(wl, first_frame, n_frames)
    return depth, truth, cur, matches, n


def pose_inputs(wl, first_frame, n_frames):
    import rgbd_slam_b200 as rs
    return rs.synth.pose_batch(first_frame, n_frames, MAX_MATCHES, n_points=N_POINTS, n_planes=N_PLANES, scale=wl.scale,
                               outlier_frac=wl.outlier_frac)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [v.strip() for v in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_port_frames_per_s(wl, depth, cur, matches, n, n_threads, frames=None):
    """Times the CPU oracle (restated reference path: CAPE + RANSAC/LM pose + covariance) on `frames` frames."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    lib = ol.load()
    B = len(depth) if frames is None else min(frames, len(depth))
    K = np.array(wl.K)
    poses = np.zeros((B, 7))
    sec = lib.orc_process_frames(wl.W, wl.H, wl.cell, K.ctypes.data, depth.ctypes.data, cur.ctypes.data, matches.ctypes.data,
                                 n.ctypes.data, MAX_MATCHES, B, 1, 1, wl.hypotheses, 100, 0, n_threads, poses.ctypes.data)
    return B / sec, sec, B


def cpu_compiled_reference_frames_per_s(wl, depth, cur, matches, n, frames=2):
    """Informational: the reference's OWN CAPE and pose-solve translation units as oracle/ref_shim compiles them
    (oracle/_ref/libref_cape.so + libref_pose.so, shipped as files), one host thread, `frames` frames. Those builds exist to pin
    the oracle bit for bit; their stand-in Eigen (an eager matrix class on the heap) makes them several times slower than the
    restated port, so they are NOT the CPU baseline - timing them only shows that the port is the faster (conservative) arm."""
    if wl.W != 640 or wl.H != 480 or wl.cell != 20 or wl.hypotheses != 119:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if not (os.path.exists(ol.REF_LIB) and os.path.exists(ol.REF_POSE_LIB)):
        return None
    frames = min(frames, len(depth))
    t0 = time.perf_counter()
    for b in range(frames):
        ol.ref_cape_run(depth[b])
        ol.ref_pose_solve(cur[b], matches[b][:n[b]])
    sec = time.perf_counter() - t0
    return {"value_1_thread": frames / sec, "frames": frames, "seconds": sec,
            "note": "compiled reference sources against stand-in third-party headers (correctness pin, see DESIGN.md section 2); "
                    "slower than the restated port timed above, hence not used as the baseline"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. The reference binary cannot be built
    (Eigen / OpenCV C++ / TBB / boost / flann absent, no network), so this is the restated oracle, all host threads
    over the frame loop (the reference's TBB build parallelises the RANSAC / variance loops instead). The reference's own
    translation units do compile against stand-in third-party headers (oracle/_ref): that build pins the oracle bit for bit
    but runs several times slower than the port (heap-allocating stand-in matrices), so timing it would flatter the GPU arm.
    Protocol: >= 10 warm-up frames, then `steps` timed samples whose MEDIAN gives the value, >= 100 timed
    frames in total; a sample = `sample` frames of the arm's workload, sized so that the run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_for(args)
    threads = os.cpu_count() or 1
    sample = min(wl.F, 64 if wl.W <= 640 else 16)
    steps = max(args.steps, (100 + sample - 1) // sample)
    depth, truth, cur, matches, n = make_inputs(wl, 0, sample)
    for _ in range(max(args.warmup, (10 + sample - 1) // sample)):
        cpu_port_frames_per_s(wl, depth, cur, matches, n, threads)
    secs = []
    t0 = time.perf_counter()
    for _ in range(steps):
        secs.append(cpu_port_frames_per_s(wl, depth, cur, matches, n, threads)[1])
    wall = time.perf_counter() - t0
    med = float(np.median(secs))
    fps = sample / med
    line = {
        "impl": "reference", "metric": "RGB-D frames/sec (%dx%d)" % (wl.W, wl.H), "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * med, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wl.config(),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d synthetic frames per timed sample, frame loop over %d host threads; median of %d samples "
                                   "(%d timed frames, %.1f s wall; mean-based value %.1f frames/s)"
                                   % (sample, threads, steps, sample * steps, wall, sample * steps / float(np.sum(secs)))},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


class _DevPtr:  # zero-copy torch view of a library-owned device buffer
    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


def bind_host_threads(local_rank, world):
    """Spreads the ranks of one box over disjoint sets of host cores (the e2e leg is fed by host threads: pinned-buffer
    writers, the copy engines' submission threads). Returns the cores this rank may use."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(len(cores) // max(world, 1), 1)
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return len(mine)
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def measure_resident(wl, args, rs, torch, dist, rank, world, local_rank, steps, warmup, extras):
    """The HBM-resident leg of one workload: inputs live on the device before the timed region. Returns a dict."""
    F = wl.F
    depth, truth, cur, matches, n = make_inputs(wl, rank * F, F)
    det = rs.PrimitiveDetection(wl.W, wl.H, wl.cell, *wl.K, max_batch=F, device=local_rank)
    solver = rs.PoseOptimization(max_batch=F, max_matches=MAX_MATCHES, max_iterations=wl.hypotheses, max_variance=100, device=local_rank)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    d_depth = torch.from_numpy(depth).cuda()
    solver.upload(cur, matches, n)
    poses_view = torch.as_tensor(_DevPtr(solver.device_poses_ptr(), (F, 7)), device="cuda")
    pose_stream = torch.cuda.Stream()
    pptr = pose_stream.cuda_stream
    # The collective is off the step's critical path: the poses of step i are copied (7 doubles per frame) into one of two
    # staging buffers at the end of the pose chain, and the all-gather of that buffer is issued on a side stream that
    # nothing waits for until the buffer comes round again two steps later - the next step's K1 starts at once.
    comm_stream = torch.cuda.Stream() if world > 1 else None
    staged = [torch.zeros((F, 7), dtype=torch.float64, device="cuda") for _ in range(2)]
    gathered = [torch.zeros((world, F, 7), dtype=torch.float64, device="cuda") for _ in range(2)]
    comm_done = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    comm_start = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    counter = [0]
    comm_ms = []

    solver_choice = {"auto": rs.abi.RS_SOLVER_AUTO, "chain": rs.abi.RS_SOLVER_CHAIN, "fused": rs.abi.RS_SOLVER_FUSED,
                     "wide": rs.abi.RS_SOLVER_WIDE}[args.solver]
    opts = solver.options(max_iterations=wl.hypotheses, seed=1234 + rank, rng_mode=rs.abi.RS_RNG_DEVICE, intrinsics=wl.K,
                          solver=solver_choice)

    def step():
        # CAPE and the pose solve of a frame are independent (the reference runs find_primitives on its own thread):
        # K1a (HBM bound) runs alone, then the latency-bound segmentation (main stream) and the pose solve (pose stream)
        # share the SMs; the main stream joins the pose stream before the next step. (Measured and not done: enqueueing the
        # solve's preparation kernel ahead of K1a - the RANSAC kernel then starts beside K1b and loses more than the 0.04 ms.)
        i = counter[0] & 1
        counter[0] += 1
        det.run_device(d_depth.data_ptr(), F, seed=0, stream=sptr)
        det.stream_wait_fit(pptr)
        solver.solve_device(F, opts, stream=pptr)
        if world > 1:
            with torch.cuda.stream(pose_stream):
                if counter[0] > 2:
                    pose_stream.wait_event(comm_done[i])   # the gather that read staged[i] two steps ago
                staged[i].copy_(poses_view)
            comm_stream.wait_stream(pose_stream)
            with torch.cuda.stream(comm_stream):
                comm_start[i].record(comm_stream)
                dist.all_gather_into_tensor(gathered[i].view(-1), staged[i].view(-1))
                comm_done[i].record(comm_stream)
        stream.wait_stream(pose_stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    det.set_timing(steps)
    solver.set_timing(steps)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = rs.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    if world > 1:
        stream.wait_stream(comm_stream)   # the timed region ends when the last gather has landed
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = rs.launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        comm_ms = [comm_start[k].elapsed_time(comm_done[k]) for k in range(2)]
    k1_ms = float(np.mean([det.kernel_ms(s)[0] for s in range(steps)]))
    seg_ms = float(np.mean([det.kernel_ms(s)[1] for s in range(steps)]))
    pose_ms = np.mean([solver.kernel_ms(s) for s in range(steps)], axis=0)
    det.set_timing(0)
    solver.set_timing(0)
    phase_ms = solver.phase_ms()   # device-timer view of the last step's solve kernel
    work_counters = solver.work_counters()
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * F * steps / (ms_total * 1e-3)

    # sanity: the timed work produced valid poses close to the synthetic truth
    out, _ = solver.download(F)
    ok_frac = float((out["status"] == 1).mean())
    pos_err = float(np.median(np.linalg.norm(out["pose"][:, :3] - truth[:, :3], axis=1)))
    check = {"frames_with_valid_pose": ok_frac, "median_position_error_mm": pos_err,
             "mean_ransac_iterations": float(out["iterations_run"].mean())}

    # ---- N > 1: the collective delivered what the ranks computed, and a shard equals a single-GPU run of the same frames ----
    if world > 1:
        last = (counter[0] - 1) & 1
        mine_ok = bool(torch.equal(gathered[last][rank], poses_view))
        # every rank holds the same gathered tensor (bitwise): compare a checksum of the bytes across ranks
        bits = gathered[last].view(torch.int64)
        digest = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device="cuda").view_as(bits)).sum()])
        dmin, dmax = digest.clone(), digest.clone()
        dist.all_reduce(dmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        same_everywhere = bool(torch.equal(dmin, dmax))
        flags = torch.tensor([int(mine_ok)], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        shard_ok = None
        if rank == 0:
            # rank 0 re-solves the shard of the LAST rank (its frame indices, its seed) on its own GPU: byte-identical poses
            r = world - 1
            _, cur_r, m_r, n_r = pose_inputs(wl, r * F, F)
            solver.upload(cur_r, m_r, n_r)
            opts_r = solver.options(max_iterations=wl.hypotheses, seed=1234 + r, rng_mode=rs.abi.RS_RNG_DEVICE, intrinsics=wl.K,
                                    solver=solver_choice)
            solver.solve_device(F, opts_r, stream=pptr)
            torch.cuda.synchronize()
            shard_ok = bool(torch.equal(gathered[last][r], poses_view))
            solver.upload(cur, matches, n)
        check["multi_gpu"] = {"gathered_rank_slices_equal_local_poses_on_every_rank": bool(flags.item() == 1),
                              "gathered_tensor_bitwise_equal_on_every_rank": same_everywhere,
                              "last_rank_shard_equals_single_gpu_rerun_on_rank0": shard_ok}
        if not (flags.item() == 1 and same_everywhere and shard_ok in (None, True)):
            raise SystemExit("multi-GPU correctness check failed: %r" % (check["multi_gpu"],))

    res = {"wl": wl, "value": value, "ms_per_step": ms_total / steps, "steps": steps, "launches": int(launches), "clocks": clocks,
           "k1_ms": k1_ms, "seg_ms": seg_ms, "pose_ms": [float(v) for v in pose_ms], "phase_ms": phase_ms, "work_counters": work_counters, "check": check,
           "comm_ms": comm_ms, "F": F}
    ctx = dict(det=det, solver=solver, opts=opts, d_depth=d_depth, depth=depth, truth=truth, cur=cur, matches=matches, n=n,
               stream=stream, pose_stream=pose_stream, poses_view=poses_view)
    if extras:
        return res, ctx
    det.close()
    solver.close()
    del d_depth
    return res, None


def roofline_of(wl, res):
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = wl.k1_bytes_per_frame * res["F"] / (res["k1_ms"] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json" if wl.cell == 20 else "k1_traffic_cell%d.json" % wl.cell)
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("frames_per_launch") == res["F"]:
            traffic = tj.get("dram_bytes_per_launch")
    return {"kernel": "cape_cell_fit (K1a + K1b)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "bytes_per_launch": wl.k1_bytes_per_frame * res["F"], "ms_per_launch": res["k1_ms"]}


def kernels_of(res):
    p = res["pose_ms"]
    if p[2] > 0:   # the three-launch chain
        return {"cape_cell_fit": res["k1_ms"], "cape_segment": res["seg_ms"], "pose_prepare (enqueued ahead of K1a)": p[0],
                "pose_ransac_final_lm": p[1], "pose_variance_covariance": p[2]}
    return {"cape_cell_fit": res["k1_ms"], "cape_segment": res["seg_ms"], "pose_prepare": p[0],
            "pose_solve (RANSAC + final LM + Monte-Carlo + covariance, one persistent kernel)": p[1],
            "pose_solve_ransac_phase (device timer, last step: first CTA in -> last final LM out)": res["phase_ms"][0],
            "pose_solve_device_timer_total (last step)": res["phase_ms"][1], "pose_solve_work_counters (last step)": res["work_counters"]}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import rgbd_slam_b200 as rs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    host_cores = bind_host_threads(local_rank, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = workload_for(args)
    F = wl.F
    W, H = wl.W, wl.H
    steps, warmup = args.steps, max(args.warmup, 3)
    res, ctx = measure_resident(wl, args, rs, torch, dist, rank, world, local_rank, steps, warmup, extras=True)
    det, solver, opts, d_depth = ctx["det"], ctx["solver"], ctx["opts"], ctx["d_depth"]
    depth, cur, matches, n = ctx["depth"], ctx["cur"], ctx["matches"], ctx["n"]
    stream, pose_stream = ctx["stream"], ctx["pose_stream"]
    sptr, pptr = stream.cuda_stream, pose_stream.cuda_stream
    extras = rank == 0 and not args.no_extras

    # ---- one frame at a time (BASELINE configs[1] / [2] at batch = 1: what the reference's per-frame track() call sees) ----
    single = None
    if extras:
        def one_frame(cape=True, pose=True):
            if cape:
                det.run_device(d_depth.data_ptr(), 1, seed=0, stream=sptr)
            if pose:
                if cape:
                    det.stream_wait_fit(pptr)
                else:
                    pose_stream.wait_stream(stream)
                solver.solve_device(1, opts, stream=pptr)
                stream.wait_stream(pose_stream)

        def time_frames(nf, **kw):
            for _ in range(5):
                one_frame(**kw)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(nf):
                one_frame(**kw)
            b_.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b_) / nf
        ms_cape, ms_pose, ms_full = time_frames(50, pose=False), time_frames(50, cape=False), time_frames(50)
        single = {"cape_ms": ms_cape, "pose_ms": ms_pose, "full_frame_ms": ms_full, "frames_per_s": 1e3 / ms_full,
                  "note": "batch = 1, inputs resident, back-to-back frames on one GPU: CAPE plane + cylinder extraction (configs[1]), "
                          "the RANSAC-LM solve with its covariance, and both overlapped (configs[2])"}

    # ---- the pose solve alone, inputs resident: the three-launch chain against the fused persistent kernel ----
    pose_alone = None
    if extras:
        pose_alone = {}
        for name, choice in (("chain", rs.abi.RS_SOLVER_CHAIN), ("fused", rs.abi.RS_SOLVER_FUSED)):
            o = solver.options(max_iterations=wl.hypotheses, seed=1234, rng_mode=rs.abi.RS_RNG_DEVICE, intrinsics=wl.K, solver=choice)
            for _ in range(3):
                solver.solve_device(F, o, stream=pptr)
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(pose_stream)
            for _ in range(10):
                solver.solve_device(F, o, stream=pptr)
            b_.record(pose_stream)
            torch.cuda.synchronize()
            pose_alone[name + "_ms_per_batch"] = a.elapsed_time(b_) / 10
            if name == "fused":
                pose_alone["fused_ransac_phase_ms"] = solver.phase_ms()[0]
        pose_alone["note"] = ("rs_pose_solve_device alone (preparation included), %d frames: three launches (per-frame RANSAC kernel, then the "
                              "Monte-Carlo kernel) against one persistent kernel with per-frame hand-over" % F)

    # ---- rectify_depth (the step in front of the path, off in the headline workload as in examples/main_TUM.cpp) ----
    rect = None
    if extras:
        d_rect = torch.empty_like(d_depth)
        rot = np.eye(4)
        cr, sr = np.cos(0.02), np.sin(0.02)
        rot[:3, :3] = np.array([[cr, 0.0, sr], [0.0, 1.0, 0.0], [-sr, 0.0, cr]])
        rot[:3, 3] = (25.0, -3.0, 4.0)
        rect_times = {}
        for name, ext in (("identity", None), ("rotated", rot)):
            det.set_rectification(ext, enable=True)
            for _ in range(2):
                det.rectify_device(d_depth.data_ptr(), F, d_rect.data_ptr(), stream=sptr)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(5):
                det.rectify_device(d_depth.data_ptr(), F, d_rect.data_ptr(), stream=sptr)
            r1.record(stream)
            torch.cuda.synchronize()
            rect_times[name] = r0.elapsed_time(r1) / 5
        rect_ms = rect_times["identity"]
        rect_alg = 8 * W * H * F   # algorithmic: the depth image read once, the rectified image written once
        rect = {"ms_per_batch": rect_ms, "frames": F, "algorithmic_bytes_per_pixel": 8,
                "algorithmic_GBps": rect_alg / (rect_ms * 1e-3) / 1e9,
                "general_extrinsics": {"ms_per_batch": rect_times["rotated"],
                                       "algorithmic_GBps": rect_alg / (rect_times["rotated"] * 1e-3) / 1e9,
                                       "note": "camera pair rotated by 0.02 rad and shifted by 25 mm: the kernels' general instantiation"},
                "note": "rs_cape_rectify_device; achieved = 8 B/pixel (depth read + rectified depth written) / time; ms_per_batch is the "
                        "reference's default calibration (coincident cameras: the instantiation for an identity rotation block). The "
                        "winners are found in the output image itself (32-bit keys, no scratch); both kernels are bound by instruction "
                        "issue (the reference's double-precision projection of every pixel), not by DRAM"}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm = float(json.load(open(peaks_path))["hbm_gbs"])
            rect["frac_of_hbm_peak"] = rect["algorithmic_GBps"] / hbm
            rect["general_extrinsics"]["frac_of_hbm_peak"] = rect["general_extrinsics"]["algorithmic_GBps"] / hbm
        det.set_rectification(None, enable=False)
        del d_rect

    # ---- plane matching against the local map (the step between find_primitives and the pose solve) ----
    plane_match = None
    if extras:
        pm = rs.synth.plane_match_problem(11, n_frames=F, n_det=8, n_extra_map=3, max_vertices=32)
        rs.plane_match(*pm[:-1], det_matched=pm[-1], device=local_rank)   # warm-up
        t0 = time.perf_counter()
        for _ in range(3):
            sel, _inter = rs.plane_match(*pm[:-1], det_matched=pm[-1], device=local_rank)
        pm_ms = (time.perf_counter() - t0) / 3 * 1e3
        plane_match = {"ms_per_batch": pm_ms, "frames": F, "map_planes": int(len(pm[4])), "detections": int(len(pm[1])),
                       "matched": int((sel >= 0).sum()),
                       "note": "rs_plane_match with host pointers (uploads, kernel and downloads inside): MapPlane::find_matches for "
                               "every map plane of a %d-frame batch, polygons of 3-32 vertices" % F}

    # ---- Kalman update of the matched map features (the step after the pose solve) ----
    kalman = None
    if extras:
        lib = rs.load()
        rng = np.random.default_rng(0)
        npt, npl = F * N_POINTS, F * N_PLANES

        def spd(cnt, d, scale, floor):
            a = rng.standard_normal((cnt, d, d)) * scale
            return a @ a.transpose(0, 2, 1) + np.eye(d) * floor
        px = rng.uniform(-3000, 3000, (npt, 3))
        pz = px + rng.standard_normal((npt, 3)) * 4
        host = [px, spd(npt, 3, 2.0, 0.1), pz, spd(npt, 3, 2.0, 0.1)]
        nrm = rng.standard_normal((npl, 3))
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        qx = np.concatenate([nrm, rng.uniform(500, 3000, (npl, 1))], axis=1)
        hostp = [qx, spd(npl, 4, 0.05, 1e-3), qx + rng.standard_normal((npl, 4)) * 0.01, spd(npl, 4, 0.05, 1e-3)]
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in host]
        devp = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in hostp]
        o_x, o_P, o_s = torch.empty_like(dev[0]), torch.empty_like(dev[1]), torch.empty(npt, dtype=torch.float64, device="cuda")
        o_m, o_st = torch.empty(npt, dtype=torch.uint8, device="cuda"), torch.empty(npt, dtype=torch.int32, device="cuda")
        q_x, q_P, q_s = torch.empty_like(devp[0]), torch.empty_like(devp[1]), torch.empty(npl, dtype=torch.float64, device="cuda")
        q_st = torch.empty(npl, dtype=torch.int32, device="cuda")

        def kalman_step():
            lib.rs_kalman_track_points_device(npt, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), 0.001,
                                              o_x.data_ptr(), o_P.data_ptr(), o_s.data_ptr(), o_m.data_ptr(), o_st.data_ptr(), sptr)
            lib.rs_kalman_track_planes_device(npl, devp[0].data_ptr(), devp[1].data_ptr(), devp[2].data_ptr(), devp[3].data_ptr(), 1e-6,
                                              q_x.data_ptr(), q_P.data_ptr(), q_s.data_ptr(), q_st.data_ptr(), sptr)
        for _ in range(3):
            kalman_step()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(10):
            kalman_step()
        k1.record(stream)
        torch.cuda.synchronize()
        kal_ms = k0.elapsed_time(k1) / 10
        kalman = {"ms_per_batch": kal_ms, "points": npt, "planes": npl, "features_per_s": (npt + npl) / (kal_ms * 1e-3),
                  "valid": float((o_st == 0).float().mean().item()),
                  "note": "rs_kalman_track_points_device + rs_kalman_track_planes_device on the matched features of one %d-frame batch" % F}

    # ---- informational: two batches in flight (a second context pair, steps dealt alternately, no cross-lane sync):
    # the throughput-bound kernels of one batch fill the SMs the latency-bound ones of the other leave idle. Not the
    # headline: a step's latency doubles and K1 no longer runs alone. ----
    pipelined = None
    if extras and world == 1 and not args.no_e2e:
        det2 = rs.PrimitiveDetection(W, H, wl.cell, *wl.K, max_batch=F, device=local_rank)
        solver2 = rs.PoseOptimization(max_batch=F, max_matches=MAX_MATCHES, max_iterations=wl.hypotheses, max_variance=100, device=local_rank)
        solver2.upload(cur, matches, n)
        d_depth2 = d_depth.clone()
        s2, p2 = torch.cuda.Stream(), torch.cuda.Stream()
        lanes = [(det, solver, d_depth, stream, pose_stream), (det2, solver2, d_depth2, s2, p2)]

        def lane_step(i):
            dt, sv, dd, ms_, ps_ = lanes[i & 1]
            dt.run_device(dd.data_ptr(), F, seed=0, stream=ms_.cuda_stream)
            dt.stream_wait_fit(ps_.cuda_stream)
            sv.solve_device(F, opts, stream=ps_.cuda_stream)
            ms_.wait_stream(ps_)
        for i in range(6):
            lane_step(i)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record(stream)
        s2.wait_stream(stream)
        nsteps2 = steps + (steps & 1)
        for i in range(nsteps2):
            lane_step(i)
        stream.wait_stream(s2)
        q1.record(stream)
        torch.cuda.synchronize()
        ms2 = q0.elapsed_time(q1)
        pipelined = {"value": F * nsteps2 / (ms2 * 1e-3), "unit": "frames/s", "ms_per_step": ms2 / nsteps2, "steps": nsteps2,
                     "note": "two %d-frame batches in flight on two context / stream pairs, inputs resident" % F}
        det2.close()
        solver2.close()
        del d_depth2

    # ---- end-to-end leg: host (pinned) buffers through the public host API, copies inside the timed region.
    # The SAME two modes are measured at every N: one blocking call at a time, and two batches in flight (two host threads per
    # rank, each calling the blocking entry points on its own contexts and pinned buffers); e2e.value is the larger. ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(steps, 10)
        wanted = ("plane_labels", "cyl_labels", "planes", "cyls", "boundary_xyz", "info")

        class Lane:
            def __init__(self, dt, sv):
                self.det, self.solver = dt, sv
                self.h_depth = torch.from_numpy(depth).pin_memory()
                self.h_depth_np = self.h_depth.numpy()
                self.h_d16 = torch.from_numpy(np.clip(np.rint(depth), 0, 65535).astype(np.uint16)).pin_memory()
                self.h_d16_np = self.h_d16.numpy()
                arrs, _ = rs.abi.alloc_cape_outputs(F, dt.n_cells, dt.max_boundary)
                self.pinned = {}
                for k in wanted:  # results land in pinned host memory
                    tns = torch.empty(arrs[k].nbytes, dtype=torch.uint8).pin_memory()
                    self.pinned[k] = tns
                    arrs[k] = tns.numpy().view(arrs[k].dtype).reshape(arrs[k].shape)
                self.arrs = arrs
                self.st = rs.abi.CapeOutputs(**{k: arrs[k].ctypes.data for k in wanted})
                self.h_matches = torch.from_numpy(matches.view(np.uint8).reshape(F, -1).copy()).pin_memory()
                self.h_matches_np = self.h_matches.numpy().view(rs.abi.match_dtype).reshape(F, MAX_MATCHES)
                self.h_pose_out = torch.empty(F * rs.abi.pose_out_dtype.itemsize, dtype=torch.uint8).pin_memory()
                self.h_pose_out_np = self.h_pose_out.numpy().view(rs.abi.pose_out_dtype)
                self.h_mask = torch.empty((F, MAX_MATCHES), dtype=torch.uint8).pin_memory()
                self.h_mask_np = self.h_mask.numpy()

            def step(self, u16, o=opts):
                # the public host API: the pose solve is enqueued first (its copies and kernels run in the shadow of the
                # depth upload), find_primitives streams the batch through the GPU in chunks, then the solve is joined
                self.solver.compute_optimized_pose_begin(cur, self.h_matches_np, n, o, out=self.h_pose_out_np, mask=self.h_mask_np)
                if u16:
                    self.det.find_primitives_u16(self.h_d16_np, alpha=1.0, seed=0, out=(self.arrs, self.st))
                else:
                    self.det.find_primitives(self.h_depth_np, seed=0, out=(self.arrs, self.st))
                return self.solver.compute_optimized_pose_end()[0]

        lane_a = Lane(det, solver)

        def timed(fn, nsteps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            fn(nsteps)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        def one_at_a_time(u16, o=opts):
            def run(nsteps):
                for _ in range(nsteps):
                    lane_a.step(u16, o)
            return run

        timed(one_at_a_time(False), 2)
        sec = timed(one_at_a_time(False), e2e_steps)
        timed(one_at_a_time(True), 2)
        sec16 = timed(one_at_a_time(True), e2e_steps)
        h2d = int(lane_a.h_depth_np.nbytes + lane_a.h_matches_np.nbytes + cur.nbytes + n.nbytes)
        h2d16 = int(lane_a.h_d16_np.nbytes + lane_a.h_matches_np.nbytes + cur.nbytes + n.nbytes)
        d2h = int(sum(lane_a.arrs[k].nbytes for k in wanted) + lane_a.h_pose_out_np.nbytes + F * MAX_MATCHES)
        one = {"value": world * F * e2e_steps / sec, "u16_depth": world * F * e2e_steps / sec16,
               "h2d_GBps_per_rank": h2d * e2e_steps / sec / 1e9, "u16_h2d_GBps_per_rank": h2d16 * e2e_steps / sec16 / 1e9}
        # the adaptor's default RNG mode (RS_RNG_REFERENCE: host std::mt19937 draws, one host round trip between the RANSAC
        # and the Monte-Carlo kernels) through the same call sequence, float depth
        opts_ref = solver.options(max_iterations=wl.hypotheses, seed=1234 + rank, rng_mode=rs.abi.RS_RNG_REFERENCE, intrinsics=wl.K)
        timed(one_at_a_time(False, opts_ref), 1)
        sec_ref = timed(one_at_a_time(False, opts_ref), max(e2e_steps // 2, 2))
        one["rng_reference_mode"] = {"value": world * F * max(e2e_steps // 2, 2) / sec_ref,
                                     "note": "RS_RNG_REFERENCE (the integration adaptor's default: the reference's own mt19937 "
                                             "shuffle / normal draws on the host, one host round trip inside the solve), float depth, one call at a time"}

        # link roofline of this box at this N: every rank copies its pinned depth batch to its GPU at the same time
        d_sink = torch.empty_like(d_depth)

        def pure_h2d(nsteps):
            for _ in range(nsteps):
                d_sink.copy_(lane_a.h_depth, non_blocking=True)
        timed(pure_h2d, 2)
        sec_link = timed(pure_h2d, e2e_steps)
        link = lane_a.h_depth_np.nbytes * e2e_steps / sec_link / 1e9
        del d_sink

        e2e = {"value": one["value"], "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "mode": "one blocking call at a time", "one_call_at_a_time": one,
               "u16_depth": {"value": one["u16_depth"], "unit": "frames/s", "h2d_bytes_per_step": h2d16,
                             "note": "rs_cape_run_u16: CV_16U sensor image, convertTo(CV_32F) on the device"},
               "h2d_link": {"GBps_per_rank_all_ranks_copying": link, "aggregate_GBps": link * world,
                            "note": "plain cudaMemcpyAsync of the pinned depth batch on every rank at once (max over ranks): what the "
                                    "box's host memory / PCIe path delivers at this N, the ceiling of any host-fed leg"},
               "host_cores_per_rank": host_cores,
               "timing": "host wall clock around the C-ABI calls (pose solve begin -> find_primitives, chunk-pipelined copies -> pose solve end), max over ranks"}

        if not args.no_e2e_lanes:
            det_b = rs.PrimitiveDetection(W, H, wl.cell, *wl.K, max_batch=F, device=local_rank)
            solver_b = rs.PoseOptimization(max_batch=F, max_matches=MAX_MATCHES, max_iterations=wl.hypotheses, max_variance=100,
                                           device=local_rank)
            lanes_h = [lane_a, Lane(det_b, solver_b)]

            def two_lanes(u16):
                def run(nsteps):
                    def worker(lane):
                        torch.cuda.set_device(local_rank)
                        for _ in range(nsteps):
                            lane.step(u16)
                    th = [threading.Thread(target=worker, args=(ln,)) for ln in lanes_h]
                    for t_ in th:
                        t_.start()
                    for t_ in th:
                        t_.join()
                return run

            timed(two_lanes(False), 2)
            sec_l = timed(two_lanes(False), e2e_steps)
            timed(two_lanes(True), 2)
            sec_l16 = timed(two_lanes(True), e2e_steps)
            two = {"value": world * 2 * F * e2e_steps / sec_l, "u16_depth": world * 2 * F * e2e_steps / sec_l16,
                   "h2d_GBps_per_rank": 2 * h2d * e2e_steps / sec_l / 1e9, "u16_h2d_GBps_per_rank": 2 * h2d16 * e2e_steps / sec_l16 / 1e9,
                   "steps": 2 * e2e_steps,
                   "note": "two host threads per rank, each calling the blocking host API on its own contexts and pinned buffers "
                           "(two batches in flight); same bytes per step"}
            e2e["two_lanes"] = two
            if two["value"] > e2e["value"]:
                e2e["value"] = two["value"]
                e2e["mode"] = "two batches in flight (two host threads per rank on the blocking host API); one call at a time: see one_call_at_a_time"
            if two["u16_depth"] > e2e["u16_depth"]["value"]:
                e2e["u16_depth"]["value"] = two["u16_depth"]
            e2e["frac_of_h2d_link"] = (e2e["value"] / world) * (h2d / F) / 1e9 / link
            lanes_h[1] = None
            det_b.close()
            solver_b.close()
        lane_a = None

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        nsample = F if W <= 640 else min(F, 32)
        fps_all, sec_all, nf = cpu_port_frames_per_s(wl, depth, cur, matches, n, threads, frames=nsample)
        fps_1, sec_1, nf1 = cpu_port_frames_per_s(wl, depth, cur, matches, n, 1, frames=32 if W <= 640 else 8)
        cpu = {"value": fps_all, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames of one step, frame loop over %d host threads (%.2f s); 1 thread on %d frames: %.1f frames/s"
                         % (nf, threads, sec_all, nf1, fps_1),
               "value_1_thread": fps_1}
        try:
            ref_build = cpu_compiled_reference_frames_per_s(wl, depth, cur, matches, n)
        except Exception as e:   # informational leg: never fails the bench
            ref_build = {"error": str(e)[:200]}
        if ref_build:
            cpu["compiled_reference_sources"] = ref_build

    det.close()
    solver.close()
    del d_depth, ctx

    # ---- BASELINE configs[4] (1280x960, 40 px cells, 1024 hypotheses) beside the headline, same ranks, shorter run ----
    config5 = None
    if args.config == 4 and not args.no_config5 and not (args.width or args.height or args.cell or args.hypotheses):
        wl5 = Workload("BASELINE configs[4]", 1280, 960, 40, 1024, 64, outlier_frac=0.3)
        steps5 = max(min(steps, 10), 3)
        r5, _ = measure_resident(wl5, args, rs, torch, dist, rank, world, local_rank, steps5, 3, extras=False)
        config5 = {"metric": "RGB-D frames/sec (1280x960)", "value": r5["value"], "unit": "frames/s", "n_gpus": world,
                   "ms_per_step": r5["ms_per_step"], "steps": steps5, "warmup": 3, "config": wl5.config(),
                   "roofline": roofline_of(wl5, r5), "kernels_ms_per_step": kernels_of(r5), "gpu_launches": r5["launches"],
                   "check": r5["check"], "clocks": r5["clocks"]}
        if world > 1:
            config5["comm_ms"] = r5["comm_ms"]

    if rank == 0:
        line = {
            "metric": "RGB-D frames/sec (%dx%d)" % (W, H), "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": wl.config(),
            "notes": {
                "rng": "RS_RNG_DEVICE (counter-based on-device draws)", "solver": args.solver,
                "streams": "K1 alone, then cape_segment (main stream) beside the pose chain (pose stream); kernels_ms_per_step are "
                           "per-kernel event times and overlap",
                "collective": ("all-gather of [frames x 7] f64 poses per step, issued on a side stream from a double-buffered copy of "
                               "the poses (nothing waits for it until the buffer is reused two steps later)") if world > 1 else "none",
            },
            "roofline": roofline_of(wl, res),
            "kernels_ms_per_step": kernels_of(res),
            "gpu_launches": res["launches"], "clocks": res["clocks"],
            "check": res["check"],
        }
        if world > 1:
            line["comm_ms"] = {"all_gather_last_two_steps": res["comm_ms"],
                               "note": "NCCL all-gather on the side stream (start -> done events); includes waiting for the slowest rank"}
        if rect is not None:
            line["rectify_depth"] = rect
        if single is not None:
            line["single_frame"] = single
        if pose_alone is not None:
            line["pose_solve_alone"] = pose_alone
        if plane_match is not None:
            line["plane_match"] = plane_match
        if kalman is not None:
            line["kalman_update"] = kalman
        if pipelined is not None:
            line["two_batches_in_flight"] = pipelined
        if e2e is not None:
            line["e2e"] = e2e
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if config5 is not None:
            line["config5"] = config5
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
