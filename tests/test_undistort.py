"""Input side: cv2.undistort of the reference's utils/iotool.py:22-39.
CPU: the restatement of the remap and the stripe-wise maps against cv2 itself and against the unmodified
reference function; GPU: lgx_undistort through the C ABI against the oracle (byte-equal)."""
import json
import os

import numpy as np
import pytest

from oracle import import_reference, ref_port, restate

cv2 = pytest.importorskip("cv2")


def camera(w, h, seed, strength=1.0):
    """a plausible calibration in the reference's JSON layout (MATLAB stereoParameters export: 2 radial + 2 tangential)"""
    rng = np.random.default_rng(seed)
    f = 1.1 * w + rng.normal() * 20
    return {"IntrinsicMatrix": [[f, 0.0, w / 2 + rng.normal() * 6], [0.0, f * (1 + rng.normal() * 1e-3), h / 2 + rng.normal() * 6],
                                [0.0, 0.0, 1.0]],
            "RadialDistortion": [float(-0.18 * strength + rng.normal() * 0.01), float(0.11 * strength + rng.normal() * 0.01)],
            "TangentialDistortion": [float(rng.normal() * 8e-4), float(rng.normal() * 8e-4)]}


def image(w, h, seed, channels=1):
    rng = np.random.default_rng(seed)
    shape = (h, w) if channels == 1 else (h, w, channels)
    return rng.integers(0, 256, shape, dtype=np.uint8)


SIZES = [(5, 4), (37, 29), (160, 120), (333, 257), (640, 480), (5000, 9)]


@pytest.mark.parametrize("size", SIZES)
def test_remap_restatement_equals_cv2_remap(size):
    w, h = size
    rng = np.random.default_rng(w + h)
    for channels in (1, 3):
        src = image(w, h, w * h, channels)
        m1 = np.stack([rng.integers(-3, w + 3, (h, w)), rng.integers(-3, h + 3, (h, w))], -1).astype(np.int16)
        m2 = rng.integers(0, 1024, (h, w)).astype(np.uint16)
        ref = cv2.remap(src, m1, m2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(restate.remap_bilinear_fixed(src, m1, m2), ref)


@pytest.mark.parametrize("size", SIZES + [(2448, 64)])
@pytest.mark.parametrize("strength", [1.0, 3.0])
def test_stripewise_maps_reproduce_cv2_undistort(lgx, size, strength):
    """host logic of the product (iotool.undistort_maps) + the restated remap == cv2.undistort == ref_port"""
    w, h = size
    cam = camera(w, h, seed=w + 7 * h, strength=strength)
    K, d = lgx.iotool.camera_arrays(cam)
    mxy, mfr = lgx.iotool.undistort_maps(K, d, w, h)
    assert mxy.shape == (h, w, 2) and mxy.dtype == np.int16 and mfr.shape == (h, w) and mfr.dtype == np.uint16
    for channels in (1, 3):
        src = image(w, h, 3 * w + h, channels)
        ref = ref_port.undistort_image(src, cam)
        assert np.array_equal(restate.remap_bilinear_fixed(src, mxy, mfr), ref)


@pytest.mark.skipif(not import_reference.available(), reason="reference checkout not present")
def test_ref_port_equals_the_reference_function():
    import sys
    sys.path.insert(0, import_reference.REFERENCE_ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_iotool", os.path.join(import_reference.REFERENCE_ROOT, "utils", "iotool.py"))
    ref_iotool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_iotool)
    for (w, h), channels in [((160, 120), 3), ((333, 257), 1), ((640, 480), 3)]:
        cam = camera(w, h, seed=w)
        src = image(w, h, h, channels)
        assert np.array_equal(ref_iotool.undistort_image(src, cam), ref_port.undistort_image(src, cam))


def test_golden_undistort_vector():
    """committed output of the UNMODIFIED reference function (oracle/make_golden.py)"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "undistort_u8_333x257.npz"))
    cam = json.loads(bytes(g["camera_json"]).decode())
    assert np.array_equal(ref_port.undistort_image(g["image"], cam), g["undistorted"])
    assert np.array_equal(ref_port.undistort_image(g["image_bgr"], cam), g["undistorted_bgr"])


def test_load_camera_data_roundtrip(lgx, tmp_path):
    cams = {"LeftCamera": camera(64, 48, 1), "RightCamera": camera(64, 48, 2)}
    p = tmp_path / "cams.json"
    p.write_text(json.dumps(cams))
    left, right = lgx.iotool.load_camera_data(str(p))
    assert left == cams["LeftCamera"] and right == cams["RightCamera"]


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("size", SIZES + [(1280, 1024)])
@pytest.mark.parametrize("channels", [1, 3])
def test_undistort_image_gpu(lgx, size, channels):
    w, h = size
    for strength in (1.0, 3.0):
        cam = camera(w, h, seed=w + 7 * h, strength=strength)
        src = image(w, h, 3 * w + h, channels)
        out = lgx.iotool.undistort_image(src, cam)
        assert out.shape == src.shape and out.dtype == np.uint8
        assert np.array_equal(out, ref_port.undistort_image(src, cam))


@pytest.mark.gpu
def test_undistort_golden_gpu(lgx):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "undistort_u8_333x257.npz"))
    cam = json.loads(bytes(g["camera_json"]).decode())
    assert np.array_equal(lgx.iotool.undistort_image(g["image"], cam), g["undistorted"])
    assert np.array_equal(lgx.iotool.undistort_image(g["image_bgr"], cam), g["undistorted_bgr"])


@pytest.mark.gpu
def test_undistort_stereo_batch_full_size(lgx):
    """2448x2048 L/R batch on the device with per-frame camera selection, strided input rows, then the front-end:
    every frame byte-equal to cv2.undistort, and detect points of an undistorted frame == the oracle's."""
    import torch
    from cylinder_pose_estimation_b200 import synth
    W, H, B = 2448, 2048, 4
    cams = [camera(W, H, 11), camera(W, H, 12)]
    maps = lgx.iotool.CameraMaps.from_params(cams, W, H)
    frames = np.stack([synth.render_u8(seed=s, **synth.CYLINDER_2448) for s in range(2)] * 2)
    padded = torch.zeros((B, H, W + 48), dtype=torch.uint8, device="cuda")
    padded[:, :, :W] = torch.from_numpy(frames).cuda()
    idx = torch.tensor([0, 1, 1, 0], dtype=torch.int32)
    und = lgx.iotool.undistort_device(padded[:, :, :W], maps, idx)
    got = und.cpu().numpy()
    for i in range(B):
        assert np.array_equal(got[i], ref_port.undistort_image(frames[i], cams[int(idx[i])])), i
    fe = lgx.Frontend(W, H, chunk_frames=2)
    res = fe.run(und, masks=False)
    s1, s2 = ref_port.frontend(got[1])
    assert res.centroid_lists()[1] == s2.centroids


@pytest.mark.gpu
def test_undistort_rejects_what_it_cannot_do(lgx):
    cam = camera(64, 48, 1)
    with pytest.raises(TypeError):
        lgx.iotool.undistort_image(np.zeros((48, 64), np.uint16), cam)
    with pytest.raises(TypeError):
        lgx.iotool.undistort_image(np.zeros((48, 64, 4), np.uint8), cam)


@pytest.mark.gpu
def test_process_images_in_folder_gpu(lgx, tmp_path):
    """utils/iotool.py:41-71: every .png with an L / R in its name is undistorted with that camera and written under
    the same name; other files are skipped"""
    w, h = 200, 150
    cams = {"LeftCamera": camera(w, h, 21), "RightCamera": camera(w, h, 22)}
    (tmp_path / "cams.json").write_text(json.dumps(cams))
    src_dir, out_dir = tmp_path / "in", tmp_path / "out"
    src_dir.mkdir()
    imgs = {"a_L.png": image(w, h, 1, 3), "a_R.png": image(w, h, 2, 3), "b_x.png": image(w, h, 3, 3), "c_L.jpg": image(w, h, 4, 3)}
    for name, im in imgs.items():
        cv2.imwrite(str(src_dir / name), im)
    lgx.iotool.process_images_in_folder(str(tmp_path / "cams.json"), str(src_dir), str(out_dir))
    assert sorted(os.listdir(out_dir)) == ["a_L.png", "a_R.png"]
    for name, cam in (("a_L.png", cams["LeftCamera"]), ("a_R.png", cams["RightCamera"])):
        assert np.array_equal(cv2.imread(str(out_dir / name)), ref_port.undistort_image(imgs[name], cam))


@pytest.mark.skipif(not import_reference.available(), reason="reference checkout not present")
def test_undistort_folder_cli_matches_the_reference(lgx, monkeypatch, tmp_path):
    """iotool.process_images_in_folder against utils/iotool.py:41-71 on the same folder (the device function is replaced
    by the CPU oracle in this test only; the GPU test above checks the real one)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_iotool", os.path.join(import_reference.REFERENCE_ROOT, "utils", "iotool.py"))
    ref_iotool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_iotool)
    monkeypatch.setattr(lgx.iotool, "undistort_image", ref_port.undistort_image)
    w, h = 120, 90
    cams = {"LeftCamera": camera(w, h, 31), "RightCamera": camera(w, h, 32)}
    (tmp_path / "cams.json").write_text(json.dumps(cams))
    src = tmp_path / "in"
    src.mkdir()
    for name, seed in (("f0_L.png", 1), ("f0_R.png", 2), ("other.png", 3), ("f1_L.bmp", 4)):
        cv2.imwrite(str(src / name), image(w, h, seed, 3))
    ref_iotool.process_images_in_folder(str(tmp_path / "cams.json"), str(src), str(tmp_path / "ref"))
    lgx.iotool.process_images_in_folder(str(tmp_path / "cams.json"), str(src), str(tmp_path / "new"))
    assert sorted(os.listdir(tmp_path / "ref")) == sorted(os.listdir(tmp_path / "new")) == ["f0_L.png", "f0_R.png"]
    for name in ("f0_L.png", "f0_R.png"):
        assert np.array_equal(cv2.imread(str(tmp_path / "ref" / name)), cv2.imread(str(tmp_path / "new" / name)))


@pytest.mark.gpu
def test_folder_cli_batched_prepass_equals_per_file_path(tmp_path, lgx):
    """iotool.grid_folder: the batched pre-pass (one undistort call + one stage-1/2 pass per batch of files, results primed for
    the per-file detect_grid) against the per-file path on the same folder: same undistorted images, same stage-1/2 results,
    same report; files of another size, without a camera letter or unreadable keep the per-file behaviour."""
    import json
    import cv2
    from cylinder_pose_estimation_b200 import frontend, iotool
    import _cases
    rng = np.random.default_rng(3)
    cam = lambda seed: {"IntrinsicMatrix": [[700.0 + seed, 0.0, 160.3], [0.0, 701.0, 127.6], [0.0, 0.0, 1.0]],
                        "RadialDistortion": [-0.11 + 0.01 * seed, 0.04], "TangentialDistortion": [0.0004, -0.0003]}
    (tmp_path / "cams.json").write_text(json.dumps({"LeftCamera": cam(0), "RightCamera": cam(1)}))
    src = tmp_path / "in"
    src.mkdir()
    for i in range(5):
        cv2.imwrite(str(src / f"p{i}_{'LR'[i % 2]}.png"), _cases.grid_u8(320, 256, seed=20 + i))
    cv2.imwrite(str(src / "big_L.png"), _cases.grid_u8(400, 300, seed=31))          # another size: its own batch
    cv2.imwrite(str(src / "colour_R.png"), rng.integers(0, 256, (256, 320, 3), dtype=np.uint8))
    (src / "broken_L.png").write_bytes(b"not a png")
    calls = {}

    def detect_grid(img):
        original, gray, blurred, binary = frontend.load_and_preprocess_image(img)
        hmask, vmask, cents = frontend.extract_joints(binary)
        calls[len(calls)] = (img.copy(), gray, blurred, binary, hmask, vmask, cents)
        return original, json.dumps({"n": len(cents), "first": list(cents[0]) if cents else None}), None, None

    runs = []
    for batch_files in (4, 1):
        calls.clear()
        ret = iotool.grid_folder(str(tmp_path / "cams.json"), str(src), str(tmp_path / f"out{batch_files}"), detect_grid,
                                 tolerate_errors=True, batch_files=batch_files)
        runs.append((json.loads(ret), dict(calls)))
    (rep_a, calls_a), (rep_b, calls_b) = runs
    assert rep_a == rep_b and "error" in rep_a["broken_L"] and rep_a["p0_L"]["n"] > 50
    assert len(calls_a) == len(calls_b) == 7
    for k in calls_a:
        for x, y in zip(calls_a[k][:6], calls_b[k][:6]):
            assert np.array_equal(x, y)
        assert calls_a[k][6] == calls_b[k][6]
    # and the per-file path is the oracle's: undistort + stages 1-2 of one file
    img = cv2.imread(str(src / "p1_R.png"))
    und = ref_port.undistort_image(img, cam(1))
    k = [k for k in calls_a if calls_a[k][0].shape == und.shape and np.array_equal(calls_a[k][0], und)]
    assert len(k) == 1
    s1, s2 = ref_port.frontend(und)
    assert np.array_equal(calls_a[k[0]][3], s1.binary) and calls_a[k[0]][6] == s2.centroids
