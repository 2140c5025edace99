"""End-to-end probe of lgx_frontend_host (pinned host frames in, centroid lists out): frames/s for a few chunk sizes with
and without the split first chunk (LGX_OPT_HOST_SPLIT_FIRST), alternating, same process.  python tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth, _lib
W, H, B = 2448, 2048, 256
host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
for chunk in (32, 64):
    fe = lgx.Frontend(W, H, chunk_frames=chunk)
    if chunk == 32:
        host.copy_(fe.render_noisy(base, B)); torch.cuda.synchronize()
    hf = host.numpy()
    bufs = fe.host_buffers(B, H, W, masks=False, max_centroids=65536)
    ref = None
    for rep in range(3):
        for split in (0, 1):
            _lib.check(fe._lib.lgx_set_option(fe._h, _lib.LGX_OPT_HOST_SPLIT_FIRST, split))
            fe.run_host(hf, buffers=bufs)
            t = time.perf_counter()
            for _ in range(3):
                out = fe.run_host(hf, buffers=bufs)
            dt = (time.perf_counter() - t) / 3
            cl = [c.copy() for c in out["centroids"][:4]]
            if ref is None:
                ref = cl
            assert all(np.array_equal(a, b) for a, b in zip(ref, cl))
            print(f"chunk {chunk} split_first={split}: {dt*1e3:.2f} ms -> {B/dt:.0f} frames/s")
    del fe, bufs
