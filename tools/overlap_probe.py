import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth, _lib
W, H, B = 2448, 2048, 256
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
dev2 = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
for chunk in (32, 128):
    fe = lgx.Frontend(W, H, chunk_frames=chunk)
    frames = fe.render_noisy(base, B)
    fe.run(frames, masks=False, max_centroids=65536); torch.cuda.synchronize()
    fe.set_timing(True); fe.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fe.run(frames, masks=False, max_centroids=65536)
    e1.record(); torch.cuda.synchronize()
    ms, chunks, launches = fe.stats()
    print(f"chunk {chunk}: device-resident (masks off) {e0.elapsed_time(e1)/3:.2f} ms per 256 frames; us/frame", {k: round(v / 3 / B * 1e3, 1) for k, v in zip(("blur5", "ridge", "sauvola", "open_hv", "joints"), ms)})
    fe.set_timing(False)
    # the same with a concurrent H2D copy stream running
    s2 = torch.cuda.Stream()
    e0.record()
    with torch.cuda.stream(s2):
        for _ in range(2):
            dev2.copy_(host, non_blocking=True)
    for _ in range(3):
        fe.run(frames, masks=False, max_centroids=65536)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk {chunk}: device-resident with a concurrent 2.56 GB H2D: {e0.elapsed_time(e1)/3:.2f} ms per 256 frames")
    del fe
