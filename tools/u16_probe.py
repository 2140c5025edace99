import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth
kw4 = {k: v for k, v in synth.CYLINDER_4096.items() if k not in ("width", "height", "noise")}
base4 = torch.stack([synth.render_base_torch(4096, 3000, device="cuda", **kw4)])
B = 64
fe4 = lgx.Frontend(4096, 3000, chunk_frames=64)  # 25 bands x 64 frames: pipeline ridge kernel
for bits in (16, 8):
    f = fe4.render_noisy(base4, B, bits=bits)
    fe4.run(f, masks=True, max_centroids=262144); torch.cuda.synchronize()
    fe4.set_timing(True); fe4.stats(reset=True)
    for _ in range(3):
        r = fe4.run(f, masks=True, max_centroids=262144)
    torch.cuda.synchronize()
    ms, chunks, _ = fe4.stats()
    tot = sum(ms) / 3
    print(f"4096x3000 u{bits}: {tot:.1f} ms/step = {B/tot*1e3:.0f} frames/s; us/frame:",
          {k: round(v / 3 / B * 1e3, 1) for k, v in zip(("blur5", "ridge", "sauvola", "open_hv", "joints"), ms)})
    fe4.set_timing(False)
