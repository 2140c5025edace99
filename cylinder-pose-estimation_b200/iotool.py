"""Drop-in for the reference's utils/iotool.py (/root/reference/utils/iotool.py): same three names, arguments and
return values; `undistort_image` runs its per-frame work on the B200 (lgx_undistort, csrc/lgx_undistort.cu).

cv2.undistort (iotool.py:38) = cv2.initUndistortRectifyMap(CV_16SC2) + cv2.remap(INTER_LINEAR, BORDER_CONSTANT).
The maps depend on the camera and the image size only: `undistort_maps` builds them once per camera on the host with
OpenCV's own initUndistortRectifyMap, stripe by stripe exactly as cv2.undistort does internally (the stripes shift
the principal point, which decides the last bit of a fixed-point coordinate), caches them and keeps a device copy.
The remap of every frame is the CUDA kernel.  8-bit images only (cv2.imread, the reference's only source of
frames at python_grid_detection_cylinder.py:34 and iotool.py:59, always returns 8-bit); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _lib
from ._lib import check

_maps_cache = {}


def load_camera_data(json_path):
    """(LeftCamera, RightCamera) parameter dicts of the calibration JSON (utils/iotool.py:8-20)."""
    with open(json_path, "r") as fh:
        rig = json.load(fh)
    return rig["LeftCamera"], rig["RightCamera"]


def camera_for(filename, left, right):
    """The reference's rule for picking a camera: an 'L' anywhere in the file name means left, otherwise an 'R' means
    right, otherwise there is none (utils/iotool.py:62-68, python_grid_detection_cylinder.py:36-41)."""
    if "L" in filename:
        return left
    return right if "R" in filename else None


def camera_arrays(camera_params):
    """(intrinsic 3x3 f64, distortion coefficients) exactly as utils/iotool.py:33-36 forms them
    (radial then tangential, np.hstack)."""
    intrinsic_matrix = np.array(camera_params["IntrinsicMatrix"])
    distortion_coeffs = np.hstack((camera_params["RadialDistortion"], camera_params["TangentialDistortion"]))
    return intrinsic_matrix, distortion_coeffs


def undistort_maps(intrinsic_matrix, distortion_coeffs, width, height):
    """Fixed-point maps of cv2.undistort for a (height, width) image: map_xy int16 [H,W,2], map_frac uint16 [H,W].
    cv2.undistort processes stripes of min(max(1, 4096 // width), height) rows, each through
    initUndistortRectifyMap with the new camera matrix's cy shifted by the stripe's first row."""
    import cv2
    A = np.asarray(intrinsic_matrix, dtype=np.float64)
    dist = np.asarray(distortion_coeffs, dtype=np.float64)
    stripe0 = min(max(1, (1 << 12) // max(width, 1)), height)
    map_xy = np.empty((height, width, 2), np.int16)
    map_frac = np.empty((height, width), np.uint16)
    eye = np.eye(3)
    v0 = A[1, 2]
    for y in range(0, height, stripe0):
        rows = min(stripe0, height - y)
        Ar = A.copy()
        Ar[1, 2] = v0 - y
        m1, m2 = cv2.initUndistortRectifyMap(A, dist, eye, Ar, (width, rows), cv2.CV_16SC2)
        map_xy[y:y + rows] = m1
        map_frac[y:y + rows] = m2
    return map_xy, map_frac


class CameraMaps:
    """The maps of one or more cameras for one image size, on the host and (lazily) on a device.
    `cameras` = list of (intrinsic_matrix, distortion_coeffs)."""

    def __init__(self, cameras, width, height):
        self.width, self.height = int(width), int(height)
        maps = [undistort_maps(K, d, self.width, self.height) for K, d in cameras]
        self.map_xy = np.ascontiguousarray(np.stack([m[0] for m in maps]))       # [ncam, H, W, 2] int16
        self.map_frac = np.ascontiguousarray(np.stack([m[1] for m in maps]))     # [ncam, H, W] uint16
        self._dev = {}

    @classmethod
    def from_params(cls, camera_params_list, width, height):
        return cls([camera_arrays(p) for p in camera_params_list], width, height)

    def device(self, device=None):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        if dev not in self._dev:
            self._dev[dev] = (torch.from_numpy(self.map_xy).to(dev),
                              torch.from_numpy(self.map_frac.view(np.int16)).to(dev))
        return self._dev[dev]


def undistort_device(frames, maps: CameraMaps, cam_index=None, out=None):
    """Batched, device-resident: frames = torch uint8 CUDA tensor [B,H,W] or [B,H,W,3] (rows may be strided),
    cam_index = None (camera 0) or an int32 tensor [B].  Returns a dense tensor of the same shape."""
    import torch
    if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() not in (3, 4):
        raise TypeError("frames must be a CUDA uint8 tensor [B,H,W] or [B,H,W,3]")
    channels = 1 if frames.dim() == 3 else int(frames.shape[3])
    B, H, W = (int(v) for v in frames.shape[:3])
    if (H, W) != (maps.height, maps.width) or channels not in (1, 3):
        raise ValueError("frame size / channel count does not match the maps")
    if frames.stride(2) != channels or (channels == 3 and frames.stride(3) != 1) or frames.stride(1) < W * channels \
            or (B > 1 and frames.stride(0) < frames.stride(1) * H):
        frames = frames.contiguous()
    mxy, mfr = maps.device(frames.device.index)
    if out is None:
        out = torch.empty(tuple(frames.shape), dtype=torch.uint8, device=frames.device)
    ci = None
    if cam_index is not None:
        ci = cam_index.to(device=frames.device, dtype=torch.int32).contiguous()
        if ci.numel() != B or (B and (int(ci.min()) < 0 or int(ci.max()) >= maps.map_xy.shape[0])):
            raise ValueError("cam_index must hold one valid camera number per frame")
    lib = _lib.load()
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    check(lib.lgx_undistort(C.c_void_p(frames.data_ptr()), channels, B, H, W, frames.stride(1), frames.stride(0) if B > 1 else frames.stride(1) * H,
                            C.c_void_p(mxy.data_ptr()), C.c_void_p(mfr.data_ptr()),
                            C.c_void_p(ci.data_ptr()) if ci is not None else None, C.c_void_p(out.data_ptr()),
                            C.c_void_p(stream)), "lgx_undistort")
    return out


def _maps_for(camera_params, width, height):
    K, d = camera_arrays(camera_params)
    key = (K.tobytes(), np.asarray(d, np.float64).tobytes(), width, height)
    m = _maps_cache.get(key)
    if m is None:
        if len(_maps_cache) >= 8:
            _maps_cache.pop(next(iter(_maps_cache)))
        m = _maps_cache[key] = CameraMaps([(K, d)], width, height)
    return m


def undistort_image(image, camera_params):
    """utils/iotool.py:22-39: undistorted copy of one image (numpy uint8 [H,W] or [H,W,3]), same shape and dtype."""
    import torch
    image = np.asarray(image)
    if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] != 3):
        raise TypeError("undistort_image: 8-bit images with 1 or 3 channels only (lgx has no CPU path)")
    H, W = image.shape[:2]
    maps = _maps_for(camera_params, W, H)
    d = torch.from_numpy(np.ascontiguousarray(image)).cuda()
    return undistort_device(d[None], maps)[0].cpu().numpy()


def undistort_batch(images, cameras):
    """undistort_image for a list of same-shape 8-bit images with one camera parameter dict each, as ONE device call
    (one upload, lgx_undistort with a per-frame camera index, one download).  Returns the list of undistorted arrays."""
    import torch
    images = [np.ascontiguousarray(im) for im in images]
    H, W = images[0].shape[:2]
    uniq = []
    idx = []
    for cam in cameras:
        key = json.dumps(cam, sort_keys=True)
        if key not in [k for k, _ in uniq]:
            uniq.append((key, cam))
        idx.append([k for k, _ in uniq].index(key))
    mkey = ("batch", tuple(k for k, _ in uniq), W, H)
    maps = _maps_cache.get(mkey)
    if maps is None:
        if len(_maps_cache) >= 8:
            _maps_cache.pop(next(iter(_maps_cache)))
        maps = _maps_cache[mkey] = CameraMaps.from_params([c for _, c in uniq], W, H)
    d = torch.from_numpy(np.stack(images)).cuda()
    out = undistort_device(d, maps, torch.tensor(idx, dtype=torch.int32, device=d.device)).cpu().numpy()
    return [out[i] for i in range(len(images))]


def process_images_in_folder(json_path, input_folder, output_folder):
    """utils/iotool.py:41-71: undistort every .png of a folder with the camera its file name selects and write the result
    under the same name; files without an L / R are reported and skipped."""
    import cv2
    cameras = load_camera_data(json_path)
    os.makedirs(output_folder, exist_ok=True)
    for name in os.listdir(input_folder):
        if not name.endswith(".png"):
            continue
        cam = camera_for(name, *cameras)
        if cam is None:
            print(f"Skipped {name}: no L or R in the file name")
            continue
        target = os.path.join(output_folder, name)
        cv2.imwrite(target, undistort_image(cv2.imread(os.path.join(input_folder, name)), cam))
        print(f"Processed {name} -> Saved to {target}")


_GRID_IMAGE_EXTS = (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff")


def _prepare_batch(folder_path, names, cameras):
    """Batched pre-pass of the folder CLIs: decodes the next files, undistorts those of the first
    file's size in one device call and runs stages 1-2 for them in one device pass (frontend.prime_stage12), so that the
    per-file detect_grid that follows finds its stage-1/2 results ready.  Files it cannot take (unreadable, no camera, other
    size) are left to the per-file path, which reports them exactly as the reference does."""
    import cv2
    from . import frontend
    picked = []
    for name in names:
        cam = camera_for(name, *cameras)
        img = cv2.imread(os.path.join(folder_path, name)) if cam is not None else None
        if img is None or img.dtype != np.uint8 or (picked and img.shape != picked[0][1].shape):
            if not picked:
                return {}
            continue
        picked.append((name, img, cam))
    if not picked:
        return {}
    und = undistort_batch([im for _, im, _ in picked], [cam for _, _, cam in picked])
    frontend.prime_stage12(und)
    return {name: u for (name, _, _), u in zip(picked, und)}


def grid_folder(json_path, folder_path, output_folder, detect_grid, tolerate_errors, batch_files=8):
    """Shared body of the folder CLIs of the two drop-in modules (reference: python_grid_detection_cylinder.py:12-64,
    python_grid_detection_plane.py:13-73): every image of the folder (os.listdir order, the order of the keys of the
    report) is undistorted with the camera its name selects and handed to `detect_grid`; the overlay goes to
    `<stem>_arc<ext>`, the decoded result JSON into `processed_images_data.json` under the file's stem, and the report is
    returned as JSON text.  tolerate_errors=False (cylinder) lets a failing file raise, as the reference does;
    True (plane) records {'error': message} for it and goes on."""
    import cv2
    from tqdm import tqdm
    cameras = load_camera_data(json_path)
    out_dir = folder_path if output_folder is None else output_folder
    os.makedirs(out_dir, exist_ok=True)
    names = [n for n in os.listdir(folder_path) if n.lower().endswith(_GRID_IMAGE_EXTS)]
    if not names:
        print(f"No images found in folder: {folder_path}")
        return None
    report = {}
    ready = {}                    # name -> undistorted image whose stages 1-2 are already computed (batched pre-pass below)
    for pos, name in enumerate(tqdm(names, desc="Processing images")):
        stem, ext = os.path.splitext(name)
        source = os.path.join(folder_path, name)
        if name not in ready:
            ready.clear()
            ready.update(_prepare_batch(folder_path, names[pos:pos + batch_files], cameras))
        try:
            cam = camera_for(name, *cameras)
            if cam is None:
                raise ValueError(f"Unknown camera type in filename: {name}")
            image = ready.pop(name) if name in ready else undistort_image(cv2.imread(source), cam)
            overlay, result_json, _, _ = detect_grid(image)
            try:
                report[stem] = json.loads(result_json)
            except json.JSONDecodeError:
                print(f"Invalid JSON data for image {name}. Skipping.")
                continue
            cv2.imwrite(os.path.join(out_dir, f"{stem}_arc{ext}"), overlay)
        except Exception as exc:
            if not tolerate_errors:
                raise
            print(f"Error processing {source}: {exc}")
            report[stem] = {"error": str(exc)}
    report_path = os.path.join(out_dir, "processed_images_data.json")
    with open(report_path, "w") as fh:
        json.dump(report, fh, indent=4)
    print(f"Data saved to {report_path}")
    return json.dumps(report)
