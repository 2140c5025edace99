#!/usr/bin/env python
"""bench.py — grid-detected frames/s of the laser-grid point extractor (stages 1-2) on B200.

    python bench.py --gpus N --steps K --warmup W            # the B200 implementation (lgx)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path on the host cores

One "step" = one pass of stages 1-2 over one batch of synthetic frames (BASELINE.json configs[2]:
256 frames of 2448x2048 u8, laser grid on a cylinder, stereo L/R pairs, per-frame noise).  Frames are
rendered on the device (csrc/lgx_synth.cu) and stay resident in HBM for `value`; `e2e` repeats the step
through the host-buffer C-ABI call (lgx_frontend_host) with pinned host frames in and centroid lists out.
N > 1: one process per GPU (torchrun), every rank runs the same per-GPU batch (weak scaling, no
collective on the data path); the timed region is bracketed by barrier + synchronize and the max over
ranks is reported.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 2448, 2048
BATCH = 256
METRIC = "grid-detected frames/s at 2448x2048 (stages 1-2, centroid lists delivered)"
UNIT = "frames/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per ridge-kernel launch from the committed ncu capture, if any (profiles/ridge_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "ridge_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's path (oracle/ref_port.py = the reference's own library calls) on host cores
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, base_path = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_port
    from cylinder_pose_estimation_b200 import synth
    base = np.load(base_path, mmap_mode="r")
    img = synth.add_noise_u8(np.asarray(base[seed % base.shape[0]]), seed)
    t = time.perf_counter()
    _, s2 = ref_port.frontend(img, as_reference=True)      # every array pass the reference makes (incl. the eigenvalue it discards)
    return time.perf_counter() - t, len(s2.centroids)


def cpu_frames_per_s(n_frames, cores, base_path):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, base_path)] * cores)          # warm-up: imports, page-in
        t = time.perf_counter()
        res = pool.map(_cpu_worker, [(s, base_path) for s in range(n_frames)], chunksize=1)
        wall = time.perf_counter() - t
    return n_frames / wall, wall, float(np.mean([r[0] for r in res])), int(np.mean([r[1] for r in res]))


def _check_worker(args):
    """oracle results of one benchmarked frame, as bit-packed planes written to shared memory + the centroid list"""
    fi, frames_path, planes_path, shape = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_port
    B, Hh, Ww = shape
    frames = np.memmap(frames_path, dtype=np.uint8, mode="r", shape=shape)
    planes = np.memmap(planes_path, dtype=np.uint8, mode="r+", shape=(B, 3, Hh, Ww // 8))
    s1, s2 = ref_port.frontend(np.asarray(frames[fi]))
    for k, m in enumerate((s1.binary, s2.hmask, s2.vmask)):
        planes[fi, k] = np.packbits(m > 0, axis=1, bitorder="little")
    return fi, np.asarray(s2.centroids, dtype=np.int32).reshape(-1, 2)


def check_batch(res, hf, host_lists, n_check, torch):
    """Compares `n_check` frames (all of them by default) of the benchmarked batch with the CPU oracle: binary / hmask / vmask
    byte for byte (as bit planes, on the device) and both centroid lists.  CPU pool, outside every timed region."""
    import multiprocessing as mp
    B, Hh, Ww = hf.shape
    assert Ww % 8 == 0
    idx = list(range(B)) if n_check >= B else [(i * 97) % B for i in range(n_check)]
    tag = f"{os.getpid()}"
    need = hf.nbytes + B * 3 * Hh * (Ww // 8)
    shm = "/dev/shm"
    try:
        st = os.statvfs(shm)
        if st.f_bavail * st.f_frsize < need + (64 << 20):
            shm = "/tmp"
    except OSError:
        shm = "/tmp"
    frames_path, planes_path = f"{shm}/lgx_check_frames_{tag}", f"{shm}/lgx_check_planes_{tag}"
    try:
        fm = np.memmap(frames_path, dtype=np.uint8, mode="w+", shape=hf.shape)
        fm[:] = hf
        fm.flush()
        np.memmap(planes_path, dtype=np.uint8, mode="w+", shape=(B, 3, Hh, Ww // 8)).flush()
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as pool:
            results = pool.map(_check_worker, [(fi, frames_path, planes_path, hf.shape) for fi in idx], chunksize=1)
        planes = np.memmap(planes_path, dtype=np.uint8, mode="r", shape=(B, 3, Hh, Ww // 8))
        cl = res.centroid_lists()
        weights = (1 << torch.arange(8, device=res.binary.device, dtype=torch.int32)).to(torch.uint8)

        def packed(t):
            return ((t.view(Hh, Ww // 8, 8) > 0).to(torch.uint8) * weights).sum(-1, dtype=torch.uint8)
        for fi, cents in results:
            want = torch.from_numpy(np.array(planes[fi])).to(res.binary.device)
            for k, plane in enumerate((res.binary, res.hmask, res.vmask)):
                assert torch.equal(packed(plane[fi]), want[k]), f"{('binary', 'hmask', 'vmask')[k]} mismatch in frame {fi}"
            assert np.array_equal(np.asarray(cl[fi], dtype=np.int32).reshape(-1, 2), cents), f"centroid list mismatch frame {fi}"
            assert np.array_equal(np.asarray(host_lists[fi], dtype=np.int32).reshape(-1, 2), cents), f"host-path centroid list mismatch frame {fi}"
        return len(results)
    finally:
        for q in (frames_path, planes_path):
            if os.path.exists(q):
                os.remove(q)


def cpu_as_shipped(base_path, n=3):
    """The path as the reference ships it: ONE process, cv2's default thread pool, stages 1-2 of
    `detect_grid` on n frames; frames/s."""
    import cv2
    from oracle import ref_port
    from cylinder_pose_estimation_b200 import synth
    base = np.load(base_path, mmap_mode="r")
    imgs = [synth.add_noise_u8(np.asarray(base[i % base.shape[0]]), 500 + i) for i in range(n + 1)]
    ref_port.frontend(imgs[0], as_reference=True)
    t = time.perf_counter()
    for im in imgs[1:]:
        ref_port.frontend(im, as_reference=True)
    return n / (time.perf_counter() - t), cv2.getNumThreads()


def host_bases():
    """two noise-free scenes (L / R: horizontal disparity) for the CPU arm, cached under /tmp"""
    from cylinder_pose_estimation_b200 import synth
    path = f"/tmp/lgx_bases_{W}x{H}.npy"
    if not os.path.exists(path):
        kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
        np.save(path, np.stack([synth.render_base(W, H, shift=s, dtype=np.float32, **kw) for s in (0.0, -37.0)]))
    return path


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    base_path = host_bases()
    frames_per_step = cores
    times = []
    for i in range(args.warmup + args.steps):
        fps, wall, per_frame, ncent = cpu_frames_per_s(frames_per_step, cores, base_path)
        if i >= args.warmup:
            times.append(wall)
    T = float(np.mean(times))
    v = frames_per_step / T
    import cv2, scipy
    out = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": T * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{BATCH}x {W}x{H} u8 cylinder frames (BASELINE.json configs[2]); each step is a "
                                  f"bounded SAMPLE of {frames_per_step} frames of it (one per host core), not the 256: "
                                  f"frames/s is per frame, so the unit is the same", "frames_per_step": frames_per_step,
                      "sample_of_workload": True},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{frames_per_step} frames/step over {cores} processes (cv2 single-threaded per "
                                      f"process), oracle/ref_port.py = the reference's own cv2/scipy/numpy calls; "
                                      f"cv2 {cv2.__version__}, scipy {scipy.__version__}, numpy {np.__version__}",
                            "s_per_frame_per_core": per_frame, "centroids_per_frame": ncent},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def synth_camera(w, h, seed):
    """a plausible calibration in the reference's JSON layout (2 radial + 2 tangential coefficients)"""
    rng = np.random.default_rng(seed)
    f = 1.1 * w + rng.normal() * 20
    return {"IntrinsicMatrix": [[f, 0.0, w / 2 + rng.normal() * 6], [0.0, f * (1 + rng.normal() * 1e-3), h / 2 + rng.normal() * 6],
                                [0.0, 0.0, 1.0]],
            "RadialDistortion": [float(-0.18 + rng.normal() * 0.01), float(0.11 + rng.normal() * 0.01)],
            "TangentialDistortion": [float(rng.normal() * 8e-4), float(rng.normal() * 8e-4)]}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_lgx(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import cylinder_pose_estimation_b200 as lgx
    from cylinder_pose_estimation_b200 import synth

    batch, chunk = args.batch, args.chunk
    fe = lgx.Frontend(W, H, chunk_frames=chunk, device=local_rank)
    kw = {k: v for k, v in synth.CYLINDER_2448.items() if k not in ("width", "height", "noise")}
    base = torch.stack([synth.render_base_torch(W, H, shift=s, device=dev, **kw) for s in (0.0, -37.0)])
    frames = fe.render_noisy(base, batch, sigma=1.0, seed0=1000 * rank, bits=8)     # resident in HBM
    torch.cuda.synchronize()
    maxc = 65536

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return fe.run(frames, masks=True, max_centroids=maxc)

    sampler = ClockSampler(torch.cuda.current_device())     # nvidia-smi needs ~0.3 s to start: begin before warm-up
    sampler.start()
    for _ in range(args.warmup):
        res = step()
    torch.cuda.synchronize()
    n_cent = int(res.counts.sum().item())
    bad = int((res.flags & 12).ne(0).sum().item())
    assert bad == 0, "capacity overflow"

    # ---- timed region: K steps, inputs resident in HBM (1.28 GB > 126 MB L2, so every step re-reads HBM)
    for _ in range(200):                                     # make sure the sampler is delivering before timing
        if sampler.lines:
            break
        time.sleep(0.01)
    sampler.lines.clear()
    fe.stats(reset=True)
    fe.set_timing(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kms, kchunks, launches = fe.stats(reset=True)
    fe.set_timing(False)
    ridge_kernel_name = fe.last_ridge_kernel()             # what the last chunk of the timed region launched (lgx_last_ridge_kernel)
    joints_kernel_name = fe.last_joints_kernel()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- e2e: host frames (pinned) -> lgx_frontend_host -> centroid lists on the host
    host_frames = torch.empty((batch, H, W), dtype=torch.uint8).pin_memory()
    host_frames.copy_(frames)
    hf = host_frames.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    fe_full = fe
    fe = lgx.Frontend(W, H, chunk_frames=args.e2e_chunk, device=local_rank)   # finer chunks: copy/compute overlap
    bufs = fe.host_buffers(batch, H, W, masks=False, max_centroids=maxc)        # page-locked outputs, allocated once
    out = fe.run_host(hf, buffers=bufs)                              # warm-up (allocates device mirrors)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = fe.run_host(hf, buffers=bufs)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    d2h = int(sum(c.nbytes for c in out["centroids"]) + out["counts"].nbytes + out["flags"].nbytes)
    # same, with the three u8 planes the reference's later stages read also copied back
    bufs_full = fe.host_buffers(batch, H, W, masks=True, max_centroids=maxc)
    fe.run_host(hf, buffers=bufs_full)                               # warm-up of the larger mirrors
    t0 = time.perf_counter()
    fe.run_host(hf, buffers=bufs_full)
    torch.cuda.synchronize()
    e2e_full_s = time.perf_counter() - t0
    del bufs_full
    # same, the three masks as bit planes (LGX_OPT_PACKED_MASKS): what a batch caller of the reference's stages 3-6 needs
    bufs_packed = fe.host_buffers(batch, H, W, masks=True, max_centroids=maxc, packed=True)
    fe.run_host(hf, buffers=bufs_packed)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        outp = fe.run_host(hf, buffers=bufs_packed)
    torch.cuda.synchronize()
    e2e_packed_s = (time.perf_counter() - t0) / e2e_steps
    packed_bytes = int(sum(outp[k].nbytes for k in ("binary", "hmask", "vmask")))
    if rank == 0 and args.check != 0:
        assert np.array_equal(lgx.unpack_mask(outp["hmask"][:2], W), res.hmask[:2].cpu().numpy())
        assert np.array_equal(lgx.unpack_mask(outp["binary"][:2], W), res.binary[:2].cpu().numpy())
    del bufs_packed, outp
    te = torch.tensor([e2e_s, e2e_full_s, e2e_packed_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, e2e_full_s, e2e_packed_s = (float(x) for x in te.tolist())

    # ---- input side: lgx_undistort on the same resident batch, stereo L/R maps (own roofline)
    und = None
    try:
        from cylinder_pose_estimation_b200 import iotool
        cams = [synth_camera(W, H, 11), synth_camera(W, H, 12)]
        maps = iotool.CameraMaps.from_params(cams, W, H)
        cam_idx = (torch.arange(batch, device=dev) % 2).to(torch.int32)
        und_out = torch.empty_like(frames)
        for _ in range(3):
            iotool.undistort_device(frames, maps, cam_idx, out=und_out)
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        u0.record()
        for _ in range(reps):
            iotool.undistort_device(frames, maps, cam_idx, out=und_out)
        u1.record()
        torch.cuda.synchronize()
        und_ms = u0.elapsed_time(u1) / reps
        und_checked = 0
        if rank == 0 and args.check != 0:
            from oracle import ref_port
            for fi in (0, 1):
                assert np.array_equal(und_out[fi].cpu().numpy(), ref_port.undistort_image(frames[fi].cpu().numpy(), cams[fi % 2]))
                und_checked += 1
        und = {"ms_per_launch": und_ms, "frames_per_launch": batch, "parity_checked_frames": und_checked}
        del und_out, maps
    except Exception as e:          # the headline path must still report if this optional row fails
        und = {"error": f"{type(e).__name__}: {e}"}

    # ---- parity of the benchmarked batch against the CPU oracle (every frame by default; outside the timed region)
    checked = 0
    if rank == 0 and args.check != 0:
        checked = check_batch(res, hf, out["centroids"], batch if args.check < 0 else min(args.check, batch), torch)

    if rank != 0:
        return
    total_frames = batch * world * args.steps
    value = total_frames / (ms_max * 1e-3)
    alg_bytes_frame = 4 * W * H + 8 * (n_cent / batch) + 4
    peak, peak_src = measured_peak()
    ridge_ms = kms[1] / max(kchunks, 1)                               # average launch duration of the dominant kernel (ridge alone)
    frames_per_launch = min(chunk, batch)
    achieved = frames_per_launch * alg_bytes_frame / (ridge_ms * 1e-3) / 1e9
    tr = ncu_traffic()
    # CPU baseline: bounded sample on the host cores (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        fps, wall, per_frame, ncent = cpu_frames_per_s(2 * cores, cores, host_bases())
        import cv2, scipy
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{2 * cores} frames of the same workload over {cores} processes in {wall:.1f} s; "
                         f"oracle/ref_port.py (the reference's own cv2/scipy/numpy calls), cv2 {cv2.__version__}, "
                         f"scipy {scipy.__version__}, numpy {np.__version__}",
               "s_per_frame_per_core": per_frame}
        shipped, nthr = cpu_as_shipped(host_bases())
        cpu["as_shipped_single_process"] = {"value": shipped, "unit": UNIT, "cv2_threads": nthr,
                                            "what": "one process, cv2 default threads, stages 1-2 (python_grid_detection_cylinder.py:77,82)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{batch}x {W}x{H} u8 cylinder frames per GPU (BASELINE.json configs[2]), stereo L/R "
                               f"scenes + per-frame noise rendered on device", "frames_per_gpu": batch,
                   "chunk_frames": chunk, "sharding": f"frames x{world} (no collective on the data path)",
                   "l2": "inputs 1.28 GB per step > 126 MB L2 (no flush needed)",
                   "parity_checked_frames": checked, "centroids_per_frame": n_cent / batch},
        "grid_points_per_s": value * n_cent / batch,
        "e2e": {"value": batch * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hf.nbytes),
                "d2h_bytes_per_step": d2h, "call": "lgx_frontend_host (pinned host frames in, centroid lists out; "
                                                   f"three rotating device slots, chunks of {args.e2e_chunk} frames)",
                "with_u8_planes_back": {"value": batch * world / e2e_full_s, "unit": UNIT,
                                        "d2h_bytes_per_step": d2h + 3 * int(hf.nbytes)},
                "with_bit_planes_back": {"value": batch * world / e2e_packed_s, "unit": UNIT,
                                         "d2h_bytes_per_step": d2h + packed_bytes,
                                         "what": "binary / hmask / vmask as bit planes (LGX_OPT_PACKED_MASKS), everything the "
                                                 "reference's stages 3-6 read"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (tr["dram_bytes_per_frame"] * frames_per_launch) if tr else None,
                     "traffic_source": (tr or {}).get("source"),
                     "kernel": ridge_kernel_name, "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": alg_bytes_frame, "frames_per_launch": frames_per_launch,
                     "launch_ms": ridge_ms,
                     "kernel_ms_share": {k: v / max(sum(kms), 1e-9) for k, v in zip(("blur5", "ridge", "sauvola", "open_hv", "joints"), kms)},
                     "joints_first_pass": joints_kernel_name,
                     "binding_bound": "issue slots of the FP64 stencil (no FMA allowed: ~104 f64 instr/px at 2 issue cycles each + ~55 others); see DESIGN.md",
                     "whole_path_frac": value / world * alg_bytes_frame / 1e9 / peak},
        "cpu_baseline": cpu,
    }
    if und and "ms_per_launch" in und:
        # algorithmic bytes: 1 B/px read + 1 B/px written; the 6 B/px of maps are per camera (L2-resident across a batch)
        ub = 2.0 * W * H * batch
        und.update({"value": batch / (und["ms_per_launch"] * 1e-3), "unit": UNIT, "kernel": "undistort_kernel<1>",
                    "roofline": {"bound": "hbm", "achieved": ub / (und["ms_per_launch"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": ub / (und["ms_per_launch"] * 1e-3) / 1e9 / peak,
                                 "algorithmic_bytes_per_frame": 2 * W * H, "with_maps_bytes_per_frame": 8 * W * H}})
    line["undistort"] = und
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[3] / configs[4]: ONE job of `total` frames sharded over the GPUs (strong scaling)
# ---------------------------------------------------------------------------------------------------
def render_multi_cylinder_base_torch(width, height, device):
    """noise-free config-5 scene on the device (synth.render_multi_cylinder without its rng draw)"""
    import torch
    from cylinder_pose_estimation_b200 import synth
    parts = [dict(n=120, pitch=9.0, lw=1.3, curv=9e-6, shift=-1100.0), dict(n=120, pitch=9.0, lw=1.3, curv=-7e-6, shift=1100.0),
             dict(n=200, pitch=7.0, lw=1.2, curv=5e-6, shift=0.0)]
    img = torch.full((height, width), 12.0, dtype=torch.float32, device=device)
    for q in parts:
        layer = synth.render_base_torch(width, height, spot=(q["shift"] == 0.0), device=device, **q)
        img = torch.where(layer > 21.0, layer, img)
    return img


def run_sharded(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import cylinder_pose_estimation_b200 as lgx
    from cylinder_pose_estimation_b200 import shard, synth
    Wc, Hc = 4096, 3000
    if args.config == 4:
        bits, total, name = 16, args.total or 8192, "BASELINE.json configs[3]: 8192 x 4096x3000 u16 cylinder frames"
        kw = {k: v for k, v in synth.CYLINDER_4096.items() if k not in ("width", "height", "noise")}
        base = torch.stack([synth.render_base_torch(Wc, Hc, shift=sft, device=dev, **kw) for sft in (0.0, -61.0)])
    else:
        bits, total, name = 8, args.total or 2048, "BASELINE.json configs[4]: 3-cylinder dense scene (occlusion, sensor noise) 4096x3000 u8"
        base = render_multi_cylinder_base_torch(Wc, Hc, dev)[None]
    lo, hi = shard.frame_range(rank, world, total)
    mine = hi - lo
    chunk = min(args.chunk, 64)
    pool_n = max(chunk, min(args.pool, mine))
    fe = lgx.Frontend(Wc, Hc, chunk_frames=chunk, device=local_rank)
    pool = fe.render_noisy(base, pool_n, sigma=1.0, seed0=lo, bits=bits)       # frame g of the job = pool[(g - lo) % pool_n]
    torch.cuda.synchronize()
    maxc = 131072

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(gather):
        """this rank's shard, chunk by chunk; point lists compacted on the device and gathered on rank 0"""
        flat, counts = [], []
        done = 0
        while done < mine:
            n = min(chunk, mine - done)
            p0 = done % pool_n
            if p0 + n > pool_n:
                n = pool_n - p0
            r = fe.run(pool[p0:p0 + n], masks=False, max_centroids=maxc)
            if gather:
                keep = torch.arange(maxc, device=dev)[None, :] < r.counts[:, None]
                flat.append(r.centroids[keep])
                counts.append(r.counts)
            done += n
        if not gather:
            return r, None
        pts, cnt = shard.gather_points_tensors(torch.cat(flat), torch.cat(counts), dst=0)
        return r, (pts, cnt)

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    for _ in range(max(1, args.warmup)):
        res, got = step(True)
    torch.cuda.synchronize()
    assert int((res.flags & 12).ne(0).sum().item()) == 0, "capacity overflow"
    for _ in range(200):
        if sampler.lines:
            break
        time.sleep(0.01)
    sampler.lines.clear()
    fe.stats(reset=True)
    fe.set_timing(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res, got = step(True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kms, kchunks, launches = fe.stats(reset=True)
    fe.set_timing(False)
    ridge_kernel_name = fe.last_ridge_kernel()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- e2e: this rank's shard from pinned host frames through lgx_frontend_host, point lists on the host
    hn = min(pool_n, 2 * chunk)
    host_pool = torch.empty((hn, Hc, Wc), dtype=pool.dtype).pin_memory()
    host_pool.copy_(pool[:hn])
    hp = host_pool.numpy()
    fe_h = lgx.Frontend(Wc, Hc, chunk_frames=min(args.e2e_chunk, 16), device=local_rank)
    bufs = fe_h.host_buffers(hn, Hc, Wc, dtype=hp.dtype, masks=False, max_centroids=maxc)
    out = fe_h.run_host(hp, buffers=bufs)
    barrier()
    t0 = time.perf_counter()
    done = 0
    while done < mine:
        out = fe_h.run_host(hp, buffers=bufs)
        done += hn
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) * (mine / done)          # (the last call may run past the shard: scaled to the shard)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    d2h = int(sum(c.nbytes for c in out["centroids"]) + out["counts"].nbytes + out["flags"].nbytes) * (mine // hn)

    # ---- parity: frames of rank 0's pool against the CPU oracle, outside the timed region
    checked = 0
    if rank == 0 and args.check != 0:
        from oracle import ref_port
        ncheck = 2 if args.check < 0 else min(args.check, hn)
        r = fe.run(pool[:chunk], masks=True, max_centroids=maxc)
        cl = r.centroid_lists()
        for i in range(ncheck):
            fi = (i * 37) % min(chunk, hn)
            s1, s2 = ref_port.frontend(hp[fi])
            assert np.array_equal(r.binary[fi].cpu().numpy(), s1.binary), f"binary mismatch frame {fi}"
            assert np.array_equal(r.hmask[fi].cpu().numpy(), s2.hmask) and np.array_equal(r.vmask[fi].cpu().numpy(), s2.vmask)
            assert cl[fi] == s2.centroids and [tuple(map(int, c)) for c in out["centroids"][fi]] == s2.centroids
            checked += 1
    if rank != 0:
        return
    pts, cnt = got
    assert cnt.shape[0] == total and int(cnt.sum().item()) == pts.shape[0], (cnt.shape, total, pts.shape)
    n_cent = float(cnt.sum().item()) / total
    value = total * args.steps / (ms_max * 1e-3)
    alg_bytes_frame = ((bits // 8) + 3) * Wc * Hc + 8 * n_cent + 4          # frame in, binary / hmask / vmask, 8 B per point
    peak, peak_src = measured_peak()
    ridge_ms = kms[1] / max(kchunks, 1)
    achieved = chunk * alg_bytes_frame / (ridge_ms * 1e-3) / 1e9
    line = {
        "metric": f"grid-detected frames/s at {Wc}x{Hc} (stages 1-2, point lists gathered on rank 0)", "value": value, "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}; job of {total} frames sharded by shard.frame_range, {mine} on rank 0; each rank cycles through "
                               f"{pool_n} distinct resident frames (per-frame seeds = global frame index) in chunks of {chunk}",
                   "total_frames": total, "sharding": f"frames / {world} (no collective on the data path; one gather of the point lists per step, "
                                                      f"inside the timed region)",
                   "l2": f"{pool_n * Wc * Hc * (bits // 8) / 1e9:.1f} GB of resident frames per GPU > 126 MB L2",
                   "parity_checked_frames": checked, "centroids_per_frame": n_cent},
        "grid_points_per_s": value * n_cent, "points_gathered_per_step": int(pts.shape[0]),
        "e2e": {"value": total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(hp.nbytes // hn) * mine, "d2h_bytes_per_step": d2h,
                "call": "lgx_frontend_host per rank over its shard (pinned host frames in, point lists out), max over ranks; no gather"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": ridge_kernel_name, "peak_source": peak_src, "algorithmic_bytes_per_frame": alg_bytes_frame,
                     "frames_per_launch": chunk, "launch_ms": ridge_ms,
                     "kernel_ms_share": {k: v / max(sum(kms), 1e-9) for k, v in zip(("blur5", "ridge", "sauvola", "open_hv", "joints"), kms)},
                     "whole_path_frac": value / world * alg_bytes_frame / 1e9 / peak},
        "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lgx", choices=["lgx", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--chunk", type=int, default=128)
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--check", type=int, default=-1, help="frames of the benchmarked batch verified against the CPU oracle after timing "
                                                         "(-1 = all of rank 0's batch, 0 = none)")
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="BASELINE.json config (1-based): 3 = 256 x 2448x2048 u8 per GPU (default, weak scaling), 4 = 8192 x 4096x3000 u16 "
                         "sharded over the GPUs (strong scaling), 5 = dense 3-cylinder scene 4096x3000 sharded over the GPUs")
    ap.add_argument("--total", type=int, default=0, help="configs 4/5: frames of the whole job (default 8192 / 2048)")
    ap.add_argument("--pool", type=int, default=128, help="configs 4/5: distinct frames resident per GPU (the shard cycles through them)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--affinity", action="store_true", help="bind each rank to the CPUs local to its GPU (NVML) before allocating pinned buffers")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.affinity and args.impl == "lgx":
        # this rank on the CPUs NVML reports as local to its GPU, before any page-locked buffer is allocated (first touch)
        try:
            import pynvml
            pynvml.nvmlInit()
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank), (os.cpu_count() + 63) // 64)
            cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception as e:
            print(f"affinity not set: {e}", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config in (4, 5):
        run_sharded(args, rank, world, local_rank)
    else:
        run_lgx(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
