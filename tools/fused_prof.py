"""Times (CUDA events) the fused ridge + sauvola kernel alone through lgx_ridge_sauvola; the command ncu wraps.
    python tools/fused_prof.py [frames=16] [W=2448] [H=2048] [reps=3] [max_ctas=0]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cylinder_pose_estimation_b200 as lgx  # noqa: E402
from cylinder_pose_estimation_b200 import _lib, synth  # noqa: E402
from cylinder_pose_estimation_b200._lib import check  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:]]
    B, W, H, reps, ctas = (a + [16, 2448, 2048, 3, 0][len(a):])[:5]
    fe = lgx.Frontend(W, H, chunk_frames=B)
    kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
    base = torch.stack([synth.render_base_torch(W, H, shift=s, device="cuda", **kw) for s in (0.0, -37.0)])
    frames = fe.render_noisy(base, B, sigma=1.0, seed0=7)
    lib = fe._lib
    if ctas:
        check(lib.lgx_set_option(fe._h, _lib.LGX_OPT_RIDGE_SMS, ctas))
    WW = lib.lgx_bits_pitch(W)
    wbits = torch.zeros((B, H, WW), dtype=torch.int32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    call = lambda: check(lib.lgx_ridge_sauvola(fe._h, P(frames), 8, B, H, W, W, H * W, None, None, None, P(wbits), None), "lgx_ridge_sauvola")
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{B} x {W}x{H} ctas={ctas or 'all'}: blur5 + fused {ms:.3f} ms = {ms / B * 1e3:.1f} us/frame")
    out = (C.c_ulonglong * 32)()
    check(lib.lgx_debug_fused_prof(out, 1))
    v = list(out)
    if v[5]:
        n = reps + 1
        nv, ne = v[5], v[17]
        print(f"  VH warps {nv // n}: cycles per warp per launch {v[4] / nv:.0f}: wait tile {v[0] / v[4]:.1%}, vertical {v[1] / v[4]:.1%}, "
              f"wait free group {v[2] / v[4]:.1%}, horizontal {v[3] / v[4]:.1%}")
        print(f"  EC warps {ne // n}: cycles per warp per launch {v[16] / ne:.0f}: wait g {v[8] / v[16]:.1%}, wait up/band {v[9] / v[16]:.1%} "
              f"(band, of EC_0's time: {4 * v[18] / v[16]:.1%}), wait down {v[10] / v[16]:.1%}, release {v[11] / v[16]:.1%}, phases 1-2 {v[12] / v[16]:.1%}, "
              f"hand-over wait {v[13] / v[16]:.1%}, phase 3 {v[14] / v[16]:.1%}, phase 4 {v[15] / v[16]:.1%}")


if __name__ == "__main__":
    main()
