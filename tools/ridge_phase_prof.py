"""Per-phase cycle breakdown of the ridge kernel (debug option LGX_OPT_RIDGE_PROF)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cylinder_pose_estimation_b200 as lgx
from cylinder_pose_estimation_b200 import synth, _lib
W, H, B = 2448, 2048, int(sys.argv[1]) if len(sys.argv) > 1 else 64
NWARPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fe = lgx.Frontend(W, H, chunk_frames=B)
_lib.check(fe._lib.lgx_set_option(fe._h, _lib.LGX_OPT_RIDGE_WARPS, NWARPS))
kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
base = torch.stack([synth.render_base_torch(W, H, device="cuda", **kw)])
frames = fe.render_noisy(base, B)
for _ in range(2):
    fe.run(frames, masks=False)
torch.cuda.synchronize()
_lib.check(fe._lib.lgx_set_option(fe._h, _lib.LGX_OPT_RIDGE_PROF, 1))
fe.set_timing(True); fe.stats(reset=True)
fe.run(frames, masks=False)
torch.cuda.synchronize()
ms, chunks, launches = fe.stats()
out = (C.c_ulonglong * 16)()
_lib.check(fe._lib.lgx_get_ridge_prof(fe._h, out, 1))
v = list(out)
ctas = v[5]; nchunks = (W - 8 + 31) // 32 + 1
names = ["S2 vertical", "S3 horizontal", "S4 hessian", "S5 chain+fill", "top barrier"]
tot = sum(v[:5])
ctas = max(ctas, 1); tot = max(tot, 1)
print(f"[{NWARPS} warps/CTA] blur {ms[0]/B*1e3:.1f} + ridge {ms[1]/B*1e3:.1f} us/frame ({B} frames); CTAs {ctas}, steps/CTA {nchunks}")
for n, c in zip(names, v[:5]):
    print(f"  {n:16s} {c/ctas/nchunks:9.0f} cycles/step  {c/tot:6.1%}")
print(f"  total            {tot/ctas/nchunks:9.0f} cycles/step")
if NWARPS == 16 or v[15]:
    w = list(out)[8:]
    tot, ctas = w[6], w[7]
    bands = (H + 123) // 124
    steps = B * bands * nchunks / ctas          # sweep steps per persistent CTA
    print(f"pipeline kernel: {ctas} persistent CTAs x {steps:.0f} steps, {tot/ctas/12/steps:.0f} cycles/step per warp")
    for i, role in enumerate("VHE"):
        print(f"  {role}: waits for input {w[2*i]/ctas/4/steps:7.0f}  for output buffer {w[2*i+1]/ctas/4/steps:7.0f} cycles/step")
