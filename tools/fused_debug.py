"""GPU diagnostic for the fused ridge + sauvola kernel (csrc/lgx_fused.cu): every case against oracle/restate.py with
mismatch counts and the first differing coordinates per plane, then a timing A/B against the three-kernel path.
    python tools/fused_debug.py [quick|full|time]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import _cases  # noqa: E402
import cylinder_pose_estimation_b200 as lgx  # noqa: E402
from cylinder_pose_estimation_b200._lib import check  # noqa: E402
from oracle import restate  # noqa: E402


def run_fused(fe, imgs):
    lib = fe._lib
    B, H, W = imgs.shape
    bits = 8 if imgs.dtype == np.uint8 else 16
    Wp, WW = lib.lgx_plane_pitch(W), lib.lgx_bits_pitch(W)
    d = torch.from_numpy(imgs).cuda()
    b = torch.full((B, H, Wp), float("nan"), dtype=torch.float64, device="cuda")
    T = torch.full((B, H, Wp), float("nan"), dtype=torch.float64, device="cuda")
    binary = torch.full((B, H, W), 77, dtype=torch.uint8, device="cuda")
    wbits = torch.zeros((B, H, WW), dtype=torch.int32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    es = bits // 8
    check(lib.lgx_ridge_sauvola(fe._h, P(d), bits, B, H, W, W * es, H * W * es, P(b), P(T), P(binary), P(wbits), None), "lgx_ridge_sauvola")
    torch.cuda.synchronize()
    return b.cpu().numpy()[:, :, :W], T.cpu().numpy()[:, :, :W], binary.cpu().numpy(), wbits.cpu().numpy()


def diff(name, got, want):
    bad = np.argwhere(got.view(np.uint64) != want.view(np.uint64)) if got.dtype == np.float64 else np.argwhere(got != want)
    if len(bad):
        ys, xs = bad[:, 0], bad[:, 1]
        print(f"    {name}: {len(bad)} differ; rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; first {bad[:6].tolist()}"
              f" got {got[tuple(bad[0])]!r} want {want[tuple(bad[0])]!r}")
    return len(bad)


def case(fe, img, label, **kw):
    r = restate.frontend(img, **kw)
    try:
        b, T, binary, wbits = run_fused(fe, img[None])
    except Exception as e:
        print(f"{label}: EXCEPTION {e}")
        return False
    n = diff("b", b[0], r["b"]) + diff("T", T[0], r["T"]) + diff("binary", binary[0], r["binary"])
    packed = np.packbits(r["binary"] > 0, axis=1, bitorder="little")
    n += diff("bits", wbits[0].view(np.uint8)[:, :packed.shape[1]], packed)
    print(f"{label}: {'ok' if n == 0 else 'MISMATCH'}")
    return n == 0


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    fe = lgx.Frontend(2448, 2048, chunk_frames=16)
    ok = True
    sizes = [(64, 16), (96, 40), (200, 37), (97, 131), (72, 124), (96, 117), (80, 118), (70, 248), (65, 152), (130, 260), (333, 257)]
    if mode in ("quick", "full"):
        for (w, h) in sizes:
            for kind in ("grid_u8", "noise_u8") + (("grid_u16",) if mode == "full" else ()):
                img = {"grid_u8": _cases.grid_u8, "noise_u8": _cases.noise_u8, "grid_u16": _cases.grid_u16}[kind](w, h, seed=w * 17 + h)
                ok &= case(fe, img, f"{kind} {w}x{h}")
                if not ok and mode == "quick":
                    return 1
        a = np.zeros((140, 200), np.uint8)
        a[30:90, 50:150] = np.random.default_rng(9).integers(0, 256, (60, 100), dtype=np.uint8)
        ok &= case(fe, a, "black border 200x140")
        ok &= case(fe, np.full((64, 96), 255, np.uint8), "flat 96x64")
        fe.set_mixed_from_cols(False)
        ok &= case(fe, _cases.grid_u8(333, 257, seed=5), "mixed=0 333x257", mixed_from_cols=False)
        fe.set_mixed_from_cols(True)
        # batches: several items per CTA, bands of a frame on different CTAs
        imgs = np.stack([_cases.grid_u8(333, 257, seed=100 + s) for s in range(64)])
        fe.set_fused(2)
        t = time.time()
        out = fe.run_host(imgs, masks=True)
        fe.set_fused(0)
        ref = fe.run_host(imgs, masks=True)
        fe.set_fused(1)
        nb = sum(int((out["binary"][i] != ref["binary"][i]).sum()) for i in range(len(imgs)))
        nc = sum(int(not np.array_equal(out["centroids"][i], ref["centroids"][i])) for i in range(len(imgs)))
        print(f"batch 64 x 333x257 (192 items): binary pixels differing {nb}, centroid lists differing {nc}, kernel {fe.last_ridge_kernel()}")
        ok &= nb == 0 and nc == 0
    if mode in ("full", "time"):
        from cylinder_pose_estimation_b200 import synth
        kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
        W, H, B = 2448, 2048, 128
        big = lgx.Frontend(W, H, chunk_frames=128)
        base = torch.stack([synth.render_base_torch(W, H, shift=s, device="cuda", **kw) for s in (0.0, -37.0)])
        frames = big.render_noisy(base, B, sigma=1.0, seed0=7)
        res = {}
        for fused in (0, 1):
            big.set_fused(fused)
            for _ in range(2):
                r = big.run(frames, masks=True, max_centroids=65536)
            torch.cuda.synchronize()
            big.stats(reset=True)
            big.set_timing(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                r = big.run(frames, masks=True, max_centroids=65536)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            kms, chunks, launches = big.stats(reset=True)
            big.set_timing(False)
            print(f"fused={fused}: {ms:.2f} ms per {B} frames = {B / ms * 1e3:.0f} frames/s; per frame us: "
                  + ", ".join(f"{n} {v / chunks / B * 1e3:.1f}" for n, v in zip(("blur", "ridge", "sauvola", "morph", "joints"), kms))
                  + f"; kernel {big.last_ridge_kernel()}")
            res[fused] = (r.binary.clone(), r.centroid_lists())
        same = torch.equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
        print("2448x2048 x128: fused == three-kernel path:", same)
        ok &= same
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
