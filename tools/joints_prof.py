"""Per-kernel-group device times of the front-end with either first pass of the contour stage.
    python tools/joints_prof.py [frames=128] [chunk=128]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cylinder_pose_estimation_b200 as lgx  # noqa: E402
from cylinder_pose_estimation_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 128
W, H = 2448, 2048
fe = lgx.Frontend(W, H, chunk_frames=chunk)
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, shift=s, device="cuda", **kw) for s in (0.0, -37.0)])
frames = fe.render_noisy(base, B, sigma=1.0, seed0=7)
res = {}
for glob in (True, False):
    fe.set_joints_global(glob)
    for _ in range(2):
        r = fe.run(frames, masks=True, max_centroids=65536)
    torch.cuda.synchronize()
    fe.stats(reset=True)
    fe.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = fe.run(frames, masks=True, max_centroids=65536)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    kms, chunks, launches = fe.stats(reset=True)
    fe.set_timing(False)
    print(f"{fe.last_joints_kernel()}: {ms:.2f} ms per {B} frames (chunks of {chunk}) = {B / ms * 1e3:.0f} frames/s; per frame us: "
          + ", ".join(f"{n} {v / 5 / B * 1e3:.1f}" for n, v in zip(("blur", "ridge", "sauvola", "morph", "joints"), kms)))
    res[glob] = r.centroid_lists()
print("identical centroid lists:", res[True] == res[False])
