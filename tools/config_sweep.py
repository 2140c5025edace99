"""Throughput of the other BASELINE.json configs (device-resident, CUDA events): single 2448x2048 frame latency
(config 2), 4096x3000 u16 batch (config 4 shape), dense multi-cylinder 4096x3000 u8 (config 5 shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


# config 2: one frame, latency incl. host round trip through the reference-named functions
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(2448, 2048, device="cuda", **kw)])
fe1 = lgx.Frontend(2448, 2048, chunk_frames=1)
one = fe1.render_noisy(base, 1)
ms, r = timed(lambda: fe1.run(one, masks=True), 20)
print(f"config2 single 2448x2048 u8 frame, device-resident: {ms*1e3:.0f} us ({int(r.counts[0])} centroids)")
img = one[0].cpu().numpy()
lgx.load_and_preprocess_image(img)
t = time.perf_counter()
for _ in range(5):
    o = lgx.load_and_preprocess_image(img); h, v, c = lgx.extract_joints(o[3])
print(f"config2 load_and_preprocess_image + extract_joints (NumPy in/out): {(time.perf_counter()-t)/5*1e3:.1f} ms, {len(c)} centroids")
del fe1

# config 4 shape: 4096x3000 u16
kw4 = {k: v for k, v in synth.CYLINDER_4096.items() if k not in ("width", "height", "noise")}
base4 = torch.stack([synth.render_base_torch(4096, 3000, device="cuda", **kw4)])
B = 64
fe4 = lgx.Frontend(4096, 3000, chunk_frames=32)
f16 = fe4.render_noisy(base4, B, bits=16)
ms, r = timed(lambda: fe4.run(f16, masks=True, max_centroids=262144), 3)
print(f"config4 shape: {B} x 4096x3000 u16: {ms:.1f} ms/step = {B/ms*1e3:.0f} frames/s, {r.counts.float().mean().item():.0f} centroids/frame, flags {int(r.flags.max())}")
f8 = fe4.render_noisy(base4, B, bits=8)
ms, r = timed(lambda: fe4.run(f8, masks=True, max_centroids=262144), 3)
print(f"same scene u8: {ms:.1f} ms/step = {B/ms*1e3:.0f} frames/s, flags {int(r.flags.max())}, generic frames {int((r.flags & 2).ne(0).sum())}")
# config 5 shape: dense multi-cylinder, one host-rendered scene + device noise
dense = torch.from_numpy(synth.render_multi_cylinder(4096, 3000, seed=1).astype(np.float32))[None].cuda()
f5 = fe4.render_noisy(dense, B, sigma=0.5, bits=8)
ms, r = timed(lambda: fe4.run(f5, masks=True, max_centroids=262144), 3)
print(f"config5 shape: {B} x 4096x3000 u8 dense multi-cylinder: {ms:.1f} ms/step = {B/ms*1e3:.0f} frames/s, {r.counts.float().mean().item():.0f} centroids/frame, flags {int(r.flags.max())}")
