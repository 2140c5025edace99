"""Generates tests/golden/*.npz by running the UNMODIFIED reference (through oracle/import_reference.py)
on seeded synthetic frames.  Build-container only (needs /root/reference); the vectors travel to the GPU box.

    python oracle/make_golden.py

Each file holds the input frame and what the reference returned for it at the stage-1/2 boundary
(binary / horizontal / vertical masks bit-packed, the centroid list in the reference's order) and, where the
reference's full detect_grid succeeds, its result JSON.  MANIFEST.json records library versions.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import import_reference  # noqa: E402
from cylinder_pose_estimation_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (module, frame)
    "cyl_u8_320x256": ("cyl", lambda: synth.render_u8(320, 256, seed=0, n=9, pitch=14.0)),
    "cyl_u8_333x257": ("cyl", lambda: synth.render_u8(333, 257, seed=1, n=9, pitch=14.0, curv=3e-5)),
    "plane_u8_320x256": ("plane", lambda: synth.render_u8(320, 256, seed=2, n=9, pitch=14.0, curv=0.0)),
    "cyl_u16_256x200": ("cyl", lambda: synth.render_u16(256, 200, seed=3, n=7, pitch=14.0)),
    "cyl_u8_24x25": ("cyl", lambda: synth.render_u8(24, 25, seed=4, n=1, pitch=8.0)),
    "noise_u8_97x131": ("cyl", lambda: np.random.default_rng(5).integers(0, 256, (131, 97), dtype=np.uint8)),
    "cyl_u8_960x768_full": ("cyl", lambda: synth.render_u8(960, 768, seed=6, n=21, pitch=28.0)),
    "plane_u8_960x768_full": ("plane", lambda: synth.render_u8(960, 768, seed=7, n=21, pitch=28.0, curv=0.0)),
    # BASELINE.json configs[0]: python_grid_detection_plane.py on one synthetic 1280x1024 8-bit plane image
    "plane_u8_1280x1024_full": ("plane", lambda: synth.render_u8(seed=1, **synth.PLANE_1280)),
}


def undistort_case():
    """utils/iotool.py:undistort_image of the unmodified reference on a seeded gray and a seeded BGR image."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_iotool", os.path.join(import_reference.REFERENCE_ROOT, "utils", "iotool.py"))
    ref_iotool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_iotool)
    w, h = 333, 257
    cam = {"IntrinsicMatrix": [[371.25, 0.0, 170.5], [0.0, 370.75, 124.25], [0.0, 0.0, 1.0]],
           "RadialDistortion": [-0.21, 0.13], "TangentialDistortion": [0.0007, -0.0004]}
    img = synth.render_u8(w, h, seed=8, n=9, pitch=14.0)
    bgr = np.random.default_rng(9).integers(0, 256, (h, w, 3), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "undistort_u8_333x257.npz"), image=img, image_bgr=bgr,
                        camera_json=np.frombuffer(json.dumps(cam).encode(), dtype=np.uint8),
                        undistorted=ref_iotool.undistort_image(img, cam), undistorted_bgr=ref_iotool.undistort_image(bgr, cam))
    return {"module": "utils/iotool.py:undistort_image", "shape": [h, w], "dtype": "uint8"}


# ---- full-size frames of BASELINE.json configs 2 / 4 / 5: too large to commit, so the fixture holds digests of what the
# unmodified reference returned for them (tests/golden/full_size_digests.json); the frames are regenerated from their seeds
FULL_SIZE = {
    "config2_cyl_u8_2448x2048": ("cyl", lambda: synth.render_u8(seed=0, **synth.CYLINDER_2448), True),
    "config4_cyl_u16_4096x3000": ("cyl", lambda: synth.render_u16(4096, 3000, seed=2, **{k: v for k, v in synth.CYLINDER_4096.items()
                                                                                        if k not in ("width", "height", "noise")}), False),
    "config5_dense_u8_4096x3000": ("cyl", lambda: synth.render_multi_cylinder(4096, 3000, seed=1), False),
}


def digest(image, binary, hmask, vmask, centroids):
    """sha256 of each array in a fixed layout (C order; centroids as int32 [n,2])"""
    import hashlib
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    c = np.asarray(centroids, dtype=np.int32).reshape(-1, 2)
    return {"image": h(image), "binary": h(np.asarray(binary, np.uint8)), "hmask": h(np.asarray(hmask, np.uint8)),
            "vmask": h(np.asarray(vmask, np.uint8)), "n_centroids": int(len(c)), "centroids": h(c)}


def full_size_digests():
    cyl, pla = import_reference.load()
    out = {}
    for name, (which, make, full) in FULL_SIZE.items():
        img = make()
        util = cyl.util_cylinder
        _original, _gray, _blurred, binary = util.load_and_preprocess_image(img)
        hmask, vmask, cents = util.extract_joints(binary)
        rec = digest(img, binary, hmask, vmask, cents)
        rec.update(shape=list(img.shape), dtype=str(img.dtype))
        if full:
            res = cyl.detect_grid(img)
            rec["result_json"] = json.loads(res[1]) if res is not None else None
        out[name] = rec
        print(name, {k: v for k, v in rec.items() if k != "result_json"})
    json.dump(out, open(os.path.join(OUT, "full_size_digests.json"), "w"), indent=1)


def main():
    import cv2, scipy
    cyl, pla = import_reference.load()
    os.makedirs(OUT, exist_ok=True)
    manifest = {"generator": "oracle/make_golden.py", "reference": import_reference.REFERENCE_ROOT,
                "versions": {"numpy": np.__version__, "cv2": cv2.__version__, "scipy": scipy.__version__,
                             "skimage": "shim (oracle/refshim/skimage): 0.19.x semantics - img_as_float multiplies by 1/imax, order='rc' forms Hrc = d(g_c)/dr; "
                                        "every vector is bit-identical under the other variants (division, d(g_r)/dc)"}, "cases": {}}
    for name, (which, make) in CASES.items():
        img = make()
        mod = cyl if which == "cyl" else pla
        util = mod.util_cylinder if which == "cyl" else mod.util_plane
        original, gray, blurred, binary = util.load_and_preprocess_image(img)
        hmask, vmask, cents = util.extract_joints(binary)
        rec = dict(image=img, blurred=blurred,
                   binary=np.packbits(binary > 0, axis=1, bitorder="little"),
                   hmask=np.packbits(hmask > 0, axis=1, bitorder="little"),
                   vmask=np.packbits(vmask > 0, axis=1, bitorder="little"),
                   centroids=np.array(cents, dtype=np.int32).reshape(-1, 2))
        info = {"module": which, "shape": list(img.shape), "dtype": str(img.dtype), "centroids": len(cents)}
        if name.endswith("_full"):
            res = mod.detect_grid(img)
            if res is not None:
                rec["result_json"] = np.frombuffer(res[1].encode(), dtype=np.uint8)
                info["grid_points"] = len(json.loads(res[1])["points"])
            else:
                info["grid_points"] = None
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        manifest["cases"][name] = info
        print(name, info)
    manifest["cases"]["undistort_u8_333x257"] = undistort_case()
    full_size_digests()
    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
