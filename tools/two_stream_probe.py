"""Experiment: two lgx handles on two CUDA streams, each owning half of the batch, with the persistent ridge
kernel restricted to N SMs (LGX_OPT_RIDGE_SMS): does the issue-bound ridge of one stream overlap with the
DRAM / latency-bound kernels of the other on the remaining SMs?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth, _lib
W, H, B = 2448, 2048, 256
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
for chunk, sms in ((128, 0), (64, 0), (64, 120), (64, 104), (64, 90), (32, 104)):
    fes = [lgx.Frontend(W, H, chunk_frames=chunk) for _ in range(2)]
    for fe in fes:
        _lib.check(fe._lib.lgx_set_option(fe._h, _lib.LGX_OPT_RIDGE_SMS, sms))
    frames = fes[0].render_noisy(base, B)
    halves = [frames[:B // 2].contiguous(), frames[B // 2:].contiguous()]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()

    def one_stream():
        return fes[0].run(frames, masks=True, max_centroids=65536)

    def two_streams():
        out = []
        for fe, h, s in zip(fes, halves, streams):
            with torch.cuda.stream(s):
                out.append(fe.run(h, masks=True, max_centroids=65536))
        return out
    for name, fn in (("one stream", one_stream), ("two streams", two_streams)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / 5
        print(f"chunk {chunk} ridge SMs {sms or 'all'} {name:12s}: {dt*1e3:6.1f} ms/step  {B/dt:7.0f} frames/s", flush=True)
    del fes
