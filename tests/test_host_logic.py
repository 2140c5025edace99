"""CPU: host-side logic — embedded constants, sharding + the world_size-2 gather (gloo), the drop-in wiring of
the reference's stages 3-6 around the lgx front-end, and the synthetic generator."""
import json
import os
import re
import sys

import numpy as np
import pytest

import _cases
from oracle import import_reference, ref_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_embedded_gauss_weights_are_scipys():
    from scipy.ndimage import _filters
    w = _filters._gaussian_kernel1d(3.0, 0, 12)
    src = open(os.path.join(ROOT, "cylinder-pose-estimation_b200", "csrc", "lgx_capi.cu")).read()
    body = src[src.index("kGaussW[13]"):]
    vals = [float.fromhex(v) for v in re.findall(r"0x1\.[0-9a-f]+p-\d+", body)[:13]]
    assert vals == [float(x) for x in w[:13]]
    assert np.array_equal(w, w[::-1])


def test_frame_ranges_cover_and_keep_pairs(lgx):
    from cylinder_pose_estimation_b200 import shard
    for total in (0, 1, 2, 7, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            r = [shard.frame_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(lo % 2 == 0 for lo, hi in r if lo < total)
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 3      # one stereo pair, plus an odd tail frame


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from cylinder_pose_estimation_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    total = 10
    lo, hi = shard.frame_range(rank, world, total)
    rng = [np.random.default_rng(f) for f in range(lo, hi)]
    lists = [r.integers(0, 4096, (int(r.integers(0, 50)), 2)).astype(np.int32) for r in rng]
    out = shard.gather_point_lists(lists, dst=0)
    if rank == 0:
        q.put([o.tolist() for o in out])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo(lgx):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = []
    for f in range(10):
        r = np.random.default_rng(f)
        want.append(r.integers(0, 4096, (int(r.integers(0, 50)), 2)).astype(np.int32).tolist())
    assert got == want


def _gather_tensor_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from cylinder_pose_estimation_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    total = 11                                   # odd: the last rank gets the tail frame
    lo, hi = shard.frame_range(rank, world, total)
    lists = [np.random.default_rng(100 + f).integers(0, 4096, (f % 4 * 7, 2)).astype(np.int32) for f in range(lo, hi)]
    pts = torch.from_numpy(np.concatenate(lists, axis=0)) if lists else torch.zeros((0, 2), dtype=torch.int32)
    cnt = torch.tensor([len(a) for a in lists], dtype=torch.int32)
    gp, gc = shard.gather_points_tensors(pts, cnt, dst=0)
    if rank == 0:
        q.put((gp.tolist(), gc.tolist()))
    else:
        assert gp is None and gc is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tensor_gather_gloo(lgx):
    """shard.gather_points_tensors (what bench.py --config 4 | 5 calls inside the timed region, over NCCL there): ragged shards,
    frames without points, global frame order"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_tensor_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    pts, cnt = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [np.random.default_rng(100 + f).integers(0, 4096, (f % 4 * 7, 2)).astype(np.int32) for f in range(11)]
    assert cnt == [len(a) for a in want]
    assert pts == np.concatenate(want, axis=0).tolist()


def test_synthetic_frames_are_deterministic(lgx):
    a = lgx.synth.render_u8(200, 160, seed=5, n=7, pitch=14.0)
    b = lgx.synth.render_u8(200, 160, seed=5, n=7, pitch=14.0)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.max() == 255
    c = lgx.synth.render_u16(200, 160, seed=5, n=7, pitch=14.0)
    assert c.dtype == np.uint16 and c.max() == 65535


@pytest.mark.skipif(not import_reference.available(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("which", ["cylinder", "plane"])
def test_dropin_wiring_reproduces_reference_json(which, monkeypatch, lgx):
    """The drop-in detect_grid = lgx stages 1-2 + the reference's own stages 3-6.  Here (no GPU) the two lgx
    functions are replaced by the CPU oracle *in the test only*, which checks the wiring: module lookup by the
    reference's names, argument order of every stage, list order hand-over and the JSON the MATLAB side decodes.
    On the GPU box the same two functions are checked against the oracle directly (tests/test_gpu_parity.py)."""
    import_reference.load()            # puts the shim + reference on sys.path
    monkeypatch.setenv("LGX_REFERENCE_ROOT", import_reference.REFERENCE_ROOT)
    from cylinder_pose_estimation_b200 import frontend, _refbridge

    def fake_stage1(img):
        s = ref_port.stage1(np.asarray(img))
        return s.original, s.gray, s.blurred, s.binary

    def fake_stage2(binary):
        s = ref_port.stage2(binary)
        return s.hmask, s.vmask, s.centroids
    monkeypatch.setattr(frontend, "load_and_preprocess_image", fake_stage1)
    monkeypatch.setattr(frontend, "extract_joints", fake_stage2)
    _refbridge._loaded.clear()
    name = "python_grid_detection_" + which
    sys.modules.pop("cylinder_pose_estimation_b200." + name, None)
    import importlib
    mod = importlib.import_module("cylinder_pose_estimation_b200." + name)
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", ("cyl" if which == "cylinder" else "plane") + "_u8_960x768_full.npz"))
        res = mod.detect_grid(g["image"])
        assert res is not None and len(res) == 4
        col_img, result_json, rows, cols = res
        assert col_img.shape == (768, 960, 3) and isinstance(result_json, str)
        assert json.loads(result_json) == json.loads(bytes(g["result_json"]).decode())
        # error convention: any exception is swallowed, None is returned
        assert mod.detect_grid(np.zeros((4, 4, 3, 1), np.uint8)) is None
        # batched entry: one device pass for stages 1-2 (faked here), reference stages 3-6 per frame

        def fake_batch(frames, chunk_frames=8):
            out = []
            for f in frames:
                a, b2 = ref_port.frontend(f)
                out.append((a.original, a.gray, a.blurred, a.binary, b2.hmask, b2.vmask, b2.centroids))
            return out
        monkeypatch.setattr(frontend, "stage12_batch", fake_batch)
        batch = mod.detect_grid_batch(np.stack([g["image"], np.zeros_like(g["image"])]))
        assert len(batch) == 2 and batch[1] is None            # a black frame has no grid: the reference raises inside
        assert json.loads(batch[0][1]) == json.loads(bytes(g["result_json"]).decode())
    finally:
        _refbridge._loaded.clear()
        sys.modules.pop("cylinder_pose_estimation_b200." + name, None)


def test_u16_to_float_without_division_is_the_ieee_quotient():
    """LGX_OPT_FLOAT_DIV mode: lgx_ridge_ws.cu converts u16 pixels with q0 = v*RN(1/65535); rem = fma(-q0, 65535, v);
    q = fma(rem, RN(1/65535), q0) instead of a division: replayed here in exact rational arithmetic (one rounding per
    operation, as the device's DMUL / DFMA do), it equals v / 65535.0 for every 16-bit value.  (The default mode is q0
    itself: scikit-image 0.19 multiplies by the reciprocal, tests/test_oracle.py::test_img_as_float_multiplies_...)"""
    from fractions import Fraction as F
    r = 1.0 / 65535.0
    for v in range(65536):
        q0 = float(v) * r
        rem = float(F(v) - F(q0) * 65535)
        q = float(F(q0) + F(rem) * F(r))
        assert q == v / 65535.0, v


@pytest.mark.skipif(not import_reference.available(), reason="reference checkout only exists in the build container")
@pytest.mark.parametrize("which", ["cylinder", "plane"])
def test_dropin_folder_cli_matches_the_reference_cli(which, monkeypatch, tmp_path, lgx):
    """process_images_in_folder of the drop-in module against the reference's own (python_grid_detection_cylinder.py:12-64)
    on the same folder: same files written, same processed_images_data.json, same return value.  As in the wiring test
    above, the three device functions are replaced by the CPU oracle in this test only (no GPU here); on the GPU box they
    are checked against the oracle directly."""
    import cv2
    import importlib
    cyl, pla = import_reference.load()
    ref_mod = cyl if which == "cylinder" else pla
    monkeypatch.setenv("LGX_REFERENCE_ROOT", import_reference.REFERENCE_ROOT)
    from cylinder_pose_estimation_b200 import frontend, _refbridge, iotool

    def fake_stage1(img):
        s = ref_port.stage1(np.asarray(img))
        return s.original, s.gray, s.blurred, s.binary

    def fake_stage2(binary):
        s = ref_port.stage2(binary)
        return s.hmask, s.vmask, s.centroids
    monkeypatch.setattr(frontend, "load_and_preprocess_image", fake_stage1)
    monkeypatch.setattr(frontend, "extract_joints", fake_stage2)
    monkeypatch.setattr(iotool, "undistort_image", ref_port.undistort_image)
    monkeypatch.setattr(iotool, "undistort_batch", lambda images, cams: [ref_port.undistort_image(i, c) for i, c in zip(images, cams)])
    monkeypatch.setattr(frontend, "prime_stage12", lambda images, chunk_frames=8: 0)      # (device pass: nothing primed on the CPU)
    g = np.load(os.path.join(ROOT, "tests", "golden", ("cyl" if which == "cylinder" else "plane") + "_u8_960x768_full.npz"))
    img = g["image"]
    h, w = img.shape
    cam = {"IntrinsicMatrix": [[1050.0, 0.0, w / 2 + 0.25], [0.0, 1049.5, h / 2 - 0.5], [0.0, 0.0, 1.0]],
           "RadialDistortion": [-0.012, 0.004], "TangentialDistortion": [0.0002, -0.0001]}
    (tmp_path / "cams.json").write_text(json.dumps({"LeftCamera": cam, "RightCamera": cam}))
    src = tmp_path / "in"
    src.mkdir()
    cv2.imwrite(str(src / "pair0_L.png"), img)
    (src / "notes.txt").write_text("not an image")
    _refbridge._loaded.clear()
    name = "cylinder_pose_estimation_b200.python_grid_detection_" + which
    sys.modules.pop(name, None)
    mod = importlib.import_module(name)
    try:
        out_ref, out_new = tmp_path / "ref", tmp_path / "new"
        ret_ref = ref_mod.process_images_in_folder(str(tmp_path / "cams.json"), str(src), str(out_ref))
        ret_new = mod.process_images_in_folder(str(tmp_path / "cams.json"), str(src), str(out_new))
        assert sorted(os.listdir(out_ref)) == sorted(os.listdir(out_new)) == ["pair0_L_arc.png", "processed_images_data.json"]
        assert json.loads(ret_new) == json.loads(ret_ref) and len(json.loads(ret_new)["pair0_L"]["points"]) > 100
        assert (out_new / "processed_images_data.json").read_text() == (out_ref / "processed_images_data.json").read_text()
        # the overlay image draws its lines in random saturation / value (util_cylinder.py:1600-1601): compare where no line is drawn
        a, b = cv2.imread(str(out_new / "pair0_L_arc.png")), cv2.imread(str(out_ref / "pair0_L_arc.png"))
        gray_px = (a[..., 0] == a[..., 1]) & (a[..., 1] == a[..., 2]) & (b[..., 0] == b[..., 1]) & (b[..., 1] == b[..., 2])
        assert a.shape == b.shape and gray_px.mean() > 0.5 and np.array_equal(a[gray_px], b[gray_px])
        # a folder without images: message and None, like the reference
        empty = tmp_path / "empty"
        empty.mkdir()
        assert mod.process_images_in_folder(str(tmp_path / "cams.json"), str(empty)) is None
    finally:
        _refbridge._loaded.clear()
        sys.modules.pop(name, None)


def test_host_buffers_are_validated_before_any_copy():
    """run_host(buffers=...) must refuse any buffer a device-to-host copy could overrun (wrong shape, dtype, list length)"""
    from cylinder_pose_estimation_b200.frontend import _validate_host_buffers
    B, H, W, n = 2, 10, 12, 64

    def bufs(**over):
        b = dict(binary=np.empty((B, H, W), np.uint8), hmask=np.empty((B, H, W), np.uint8), vmask=None,
                 blurred=np.empty((B, H, W), np.uint16), cent=np.empty((B, n, 2), np.int32), centf=None,
                 counts=np.empty((B,), np.int32), flags=np.empty((B,), np.uint32), n=n)
        b.update(over)
        return b
    _validate_host_buffers(bufs(), B, H, W, np.dtype(np.uint16))
    _validate_host_buffers(bufs(cent=np.empty((B + 3, n, 2), np.int32), counts=np.empty((B + 3,), np.int32)), B, H, W, np.dtype(np.uint16))
    bad = [bufs(blurred=np.empty((B, H, W), np.uint8)),                 # u16 frames, u8 blurred buffer
           bufs(hmask=np.empty((B, H, W - 1), np.uint8)),
           bufs(binary=np.empty((B, H, W), np.int8)),
           bufs(cent=np.empty((B, n - 1, 2), np.int32)),                  # n of the buffers disagrees with cent.shape[1]
           bufs(cent=np.empty((B - 1, n, 2), np.int32)),
           bufs(centf=np.empty((B, n, 2), np.float32)),
           bufs(counts=np.empty((B - 1,), np.int32)),
           bufs(flags=np.empty((B,), np.int32)),
           bufs(cent=None),
           bufs(hmask=np.empty((B, H, 2 * W), np.uint8)[:, :, ::2])]     # not contiguous
    for b in bad:
        with pytest.raises(ValueError):
            _validate_host_buffers(b, B, H, W, np.dtype(np.uint16))


def test_unpack_mask_is_the_inverse_of_the_bit_plane_layout():
    """LGX_OPT_PACKED_MASKS layout (bit i of word w = pixel 32 w + i, rows padded to whole words): unpack_mask restores the
    reference's u8 planes for any width, batched or not"""
    from cylinder_pose_estimation_b200.frontend import unpack_mask
    rng = np.random.default_rng(0)
    for w in (1, 31, 32, 33, 95, 2448):
        m = (rng.random((3, 5, w)) < 0.4)
        ww = (w + 31) // 32
        padded = np.zeros((3, 5, ww * 32), bool)
        padded[..., :w] = m
        words = np.packbits(padded, axis=-1, bitorder="little").view(np.uint32)
        assert words.shape == (3, 5, ww)
        assert words[0, 0, 0] & 1 == int(m[0, 0, 0])
        out = unpack_mask(words, w)
        assert out.dtype == np.uint8 and out.shape == (3, 5, w) and np.array_equal(out, m.astype(np.uint8) * 255)
        assert np.array_equal(unpack_mask(words[1], w), m[1].astype(np.uint8) * 255)


def test_stage_caches_answer_only_for_the_unchanged_object():
    """frontend's stage-1 (primed by the batched folder pre-pass) and stage-2 caches: a hit needs the same object with the same
    bytes, is consumed by the hit, and an in-place edit or a copy is a miss (the reference recomputes on every call)"""
    from cylinder_pose_estimation_b200 import frontend
    rng = np.random.default_rng(1)
    binary = (rng.random((40, 50)) < 0.5).astype(np.uint8) * 255
    hm, vm, cents = binary.copy(), binary.copy(), [(1, 2), (3, 4)]
    frontend._stage2_cache.clear()
    frontend._remember(binary, hm, vm, cents)
    assert frontend._recall(binary.copy()) is None                       # another object
    hit = frontend._recall(binary)
    assert hit is not None and hit[0] is hm and hit[2] is cents
    assert frontend._recall(binary) is None                              # consumed
    frontend._remember(binary, hm, vm, cents)
    binary[3, 4] ^= 255                                                  # edited in place
    assert frontend._recall(binary) is None
    binary[3, 4] ^= 255
    assert frontend._recall(binary) is not None
    # two edits that keep the plain sum (one pixel on, one off) are caught by the second checksum
    frontend._remember(binary, hm, vm, cents)
    on, off = np.argwhere(binary == 0)[0], np.argwhere(binary == 255)[0]
    binary[tuple(on)], binary[tuple(off)] = 255, 0
    assert frontend._recall(binary) is None
    frontend._stage2_cache.clear()
    # stage 1
    img = rng.integers(0, 256, (40, 50, 3), dtype=np.uint8)
    frontend._stage1_cache[:] = [(frontend.weakref.ref(img), frontend._checksums(img), ("o", "g", "bl", "bi", "h", "v", []))]
    assert frontend._recall1(img.copy()) is None and frontend._recall1(img) == ("o", "g", "bl", "bi", "h", "v", [])
    assert frontend._recall1(img) is None
