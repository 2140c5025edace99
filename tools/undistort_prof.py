"""Runs lgx_undistort on B stereo frames (for ncu / timing): python tools/undistort_prof.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import iotool
from bench import synth_camera
W, H = 2448, 2048
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
maps = iotool.CameraMaps.from_params([synth_camera(W, H, 11), synth_camera(W, H, 12)], W, H)
frames = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device="cuda")
idx = (torch.arange(B, device="cuda") % 2).to(torch.int32)
out = torch.empty_like(frames)
for _ in range(2):
    iotool.undistort_device(frames, maps, idx, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    iotool.undistort_device(frames, maps, idx, out=out)
e1.record(); torch.cuda.synchronize()
print(f"undistort {B} frames: {e0.elapsed_time(e1) / 5 / B * 1e3:.2f} us/frame")
