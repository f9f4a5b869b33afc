"""Times K1 alone (K1a cape_cell_fit_kernel + K1b cape_cell_finish_kernel) on 256 resident 640x480 frames: 20 launches between two
CUDA events after 3 warm-up launches. Usage (GPU box): python tools/k1_time.py. Used by tools/exp_k1_variants.sh."""
import sys; sys.path.insert(0,'.')
import numpy as np, torch, rgbd_slam_b200 as rs
F=256
depth=np.stack([rs.synth.scene_v0_depth(i) for i in range(32)]); depth=np.tile(depth,(8,1,1))
det=rs.PrimitiveDetection(640,480,20,max_batch=F)
d=torch.from_numpy(depth).cuda()
s=torch.cuda.current_stream().cuda_stream
for _ in range(3): det.cell_fit_device(d.data_ptr(),F,stream=s)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): det.cell_fit_device(d.data_ptr(),F,stream=s)
e1.record(); torch.cuda.synchronize()
print("K1 ms/launch", e0.elapsed_time(e1)/20)
