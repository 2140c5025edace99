"""Runs the front-end on B frames with a chosen ridge instantiation (for ncu / timing): python tools/ridge_ws_prof.py B NWARPS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth
W, H = 2448, 2048
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NWARPS = int(sys.argv[2]) if len(sys.argv) > 2 else 16
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
fe = lgx.Frontend(W, H, chunk_frames=B)
fe.set_ridge_warps(NWARPS)
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
frames = fe.render_noisy(base, B)
fe.run(frames, masks=False)
torch.cuda.synchronize()
fe.set_timing(True); fe.stats(reset=True)
for _ in range(REPS):
    fe.run(frames, masks=False)
torch.cuda.synchronize()
ms, chunks, launches = fe.stats()
print(f"[ridge_warps={NWARPS}] {B} frames: us/frame", {k: round(v / REPS / B * 1e3, 1) for k, v in zip(("blur5", "ridge", "sauvola", "open_hv", "joints"), ms)})
