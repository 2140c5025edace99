"""Drop-in for the reference's python_grid_detection_cylinder (detect_grid at
/root/reference/python_grid_detection_cylinder.py:68-112): same name, argument, 4-tuple return and
swallow-and-return-None error behaviour; stages 1-2 run on the B200 through lgx, stages 3-6 are the
reference's own util_cylinder functions.  MATLAB binds it by module name (utils/makePyGridPts.m:6,15,29):
point pyEnv.utilsDir at this directory and keep pyEnv.modulename."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
import cylinder_pose_estimation_b200 as _lgx            # noqa: E402
from cylinder_pose_estimation_b200 import _refbridge    # noqa: E402

util_cylinder = _refbridge.load_reference_utils("util_cylinder")


def detect_grid(input_img):
    try:
        u = util_cylinder
        original, gray, _blurred, binary = u.load_and_preprocess_image(input_img)           # stage 1 (lgx)
        hmask, vmask, centroids = u.extract_joints(binary)                                  # stage 2 (lgx)
        contour, contour_mask = u.detect_largest_blob(original, binary, clipLimit=4.5)      # stage 3
        _img, cyl_centroids, center, _radius = u.find_cylinder_centroids_and_center(        # stage 4
            centroids, contour, gray, original)
        roi_h, roi_v, spot_radius = u.mask_roi_around_center(hmask, vmask, contour_mask, original)   # stage 5
        return u.color_and_expand_lines(roi_h, roi_v, spot_radius, center, contour, contour_mask,    # stage 6
                                        original, cyl_centroids)
    except Exception as e:   # the reference prints and returns None (python_grid_detection_cylinder.py:111-112)
        print(f"Error in detect_grid: {e}")
        return None


def detect_grid_batch(frames, chunk_frames=8):
    """Additive: detect_grid for a stack of gray frames [B,H,W].  Stages 1-2 run as one batched device pass
    (frontend.stage12_batch); the reference's stages 3-6 then run per frame.  Returns a list with one
    detect_grid result (4-tuple or None) per frame."""
    results = []
    u = util_cylinder
    for original, gray, _blurred, binary, hmask, vmask, centroids in _lgx.frontend.stage12_batch(frames, chunk_frames):
        try:
            contour, contour_mask = u.detect_largest_blob(original, binary, clipLimit=4.5)
            _img, cents, center, _radius = u.find_cylinder_centroids_and_center(centroids, contour, gray, original)
            roi_h, roi_v, spot_radius = u.mask_roi_around_center(hmask, vmask, contour_mask, original)
            results.append(u.color_and_expand_lines(roi_h, roi_v, spot_radius, center, contour, contour_mask, original, cents))
        except Exception as e:
            print(f"Error in detect_grid: {e}")
            results.append(None)
    return results


def detect_points_batch(frames, chunk_frames=8):
    """Additive: batched stages 1-2 only (see frontend.detect_points_batch)."""
    return _lgx.detect_points_batch(frames, chunk_frames)


def process_images_in_folder(json_path, folder_path, output_folder=None):
    """Folder CLI of the reference (python_grid_detection_cylinder.py:12-64): same files written (`<stem>_arc<ext>`,
    `processed_images_data.json`), same return value and error behaviour; undistortion and stages 1-2 on the device
    (iotool.grid_folder)."""
    from cylinder_pose_estimation_b200 import iotool
    return iotool.grid_folder(json_path, folder_path, output_folder, detect_grid, tolerate_errors=False)
