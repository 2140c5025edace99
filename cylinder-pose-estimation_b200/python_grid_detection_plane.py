"""Drop-in for the reference's python_grid_detection_plane (detect_grid at
/root/reference/python_grid_detection_plane.py:74-119).  Stages 1-2 are byte-identical in util_plane and
util_cylinder, so the same lgx front-end serves both; stages 3-6 are the reference's own
util_plane functions (convex-hull ROI instead of the blob detector, no centre point in stage 6)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
import cylinder_pose_estimation_b200 as _lgx            # noqa: E402
from cylinder_pose_estimation_b200 import _refbridge    # noqa: E402

util_plane = _refbridge.load_reference_utils("util_plane")


def detect_grid(input_img):
    try:
        u = util_plane
        original, gray, _blurred, binary = u.load_and_preprocess_image(input_img)           # stage 1 (lgx)
        hmask, vmask, centroids = u.extract_joints(binary)                                  # stage 2 (lgx)
        contour, contour_mask = u.get_convex_hull(original, expansion_pixels=5, visualize=False)   # stage 3
        _img, plane_centroids, _center, _radius = u.find_cylinder_centroids_and_center(     # stage 4
            centroids, contour, gray, original)
        roi_h, roi_v, spot_radius = u.mask_roi_around_center(hmask, vmask, contour_mask, original)   # stage 5
        return u.color_and_expand_lines(roi_h, roi_v, spot_radius, contour, contour_mask,   # stage 6
                                        original, plane_centroids)
    except Exception as e:   # the reference prints and returns None (python_grid_detection_plane.py:118-119)
        print(f"Error in detect_grid: {e}")
        return None


def detect_grid_batch(frames, chunk_frames=8):
    """Additive: detect_grid for a stack of gray frames [B,H,W].  Stages 1-2 run as one batched device pass
    (frontend.stage12_batch); the reference's stages 3-6 then run per frame.  Returns a list with one
    detect_grid result (4-tuple or None) per frame."""
    results = []
    u = util_plane
    for original, gray, _blurred, binary, hmask, vmask, centroids in _lgx.frontend.stage12_batch(frames, chunk_frames):
        try:
            contour, contour_mask = u.get_convex_hull(original, expansion_pixels=5, visualize=False)
            _img, cents, center, _radius = u.find_cylinder_centroids_and_center(centroids, contour, gray, original)
            roi_h, roi_v, spot_radius = u.mask_roi_around_center(hmask, vmask, contour_mask, original)
            results.append(u.color_and_expand_lines(roi_h, roi_v, spot_radius, contour, contour_mask, original, cents))
        except Exception as e:
            print(f"Error in detect_grid: {e}")
            results.append(None)
    return results


def detect_points_batch(frames, chunk_frames=8):
    return _lgx.detect_points_batch(frames, chunk_frames)


def process_images_in_folder(json_path, folder_path, output_folder=None):
    """Folder CLI of the reference (python_grid_detection_plane.py:13-73): same files written (`<stem>_arc<ext>`,
    `processed_images_data.json`), same return value and error behaviour; undistortion and stages 1-2 on the device
    (iotool.grid_folder)."""
    from cylinder_pose_estimation_b200 import iotool
    return iotool.grid_folder(json_path, folder_path, output_folder, detect_grid, tolerate_errors=True)
