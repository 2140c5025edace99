import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth
W, H, B = 2448, 2048, 256
host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
dev = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
for _ in range(2):
    t = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"H2D pinned 1.28 GB: {dt*1e3:.1f} ms = {host.numel()/dt/1e9:.1f} GB/s")
for _ in range(2):
    t = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"D2H pinned 1.28 GB: {dt*1e3:.1f} ms = {host.numel()/dt/1e9:.1f} GB/s")
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
for chunk in (32, 64):
    fe = lgx.Frontend(W, H, chunk_frames=chunk)
    base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
    frames = fe.render_noisy(base, B)
    host.copy_(frames); torch.cuda.synchronize()
    hf = host.numpy()
    fe.run_host(hf, masks=False, max_centroids=65536)
    for rep in range(2):
        t = time.perf_counter(); out = fe.run_host(hf, masks=False, max_centroids=65536); dt = time.perf_counter() - t
        print(f"chunk {chunk}: run_host {dt*1e3:.1f} ms -> {B/dt:.0f} fps")
    # device-resident for comparison at this chunk
    fe.run(frames, masks=False, max_centroids=65536); torch.cuda.synchronize()
    t = time.perf_counter(); fe.run(frames, masks=False, max_centroids=65536); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"chunk {chunk}: device-resident {dt*1e3:.1f} ms -> {B/dt:.0f} fps")
    # raw C call with preallocated pinned outputs
    import ctypes as C
    cent = torch.empty((B, 65536, 2), dtype=torch.int32).pin_memory().numpy()
    counts = torch.empty((B,), dtype=torch.int32).pin_memory().numpy()
    flags = torch.empty((B,), dtype=torch.int32).pin_memory().numpy()
    P = lambda a: C.c_void_p(a.ctypes.data)
    for rep in range(2):
        t = time.perf_counter()
        rc = fe._lib.lgx_frontend_host(fe._h, P(hf), 8, B, H, W, None, None, None, None, P(cent), None, 65536, P(counts), P(flags), None)
        dt = time.perf_counter() - t
        print(f"chunk {chunk}: raw lgx_frontend_host pinned outputs rc={rc} {dt*1e3:.1f} ms -> {B/dt:.0f} fps")
    del fe
