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
    """Folder CLI of the reference (python_grid_detection_plane.py:13-73): imread -> undistort with the left / right camera chosen
    by an 'L' / 'R' in the file name (lgx_undistort on the device) -> detect_grid; writes `<name>_arc<ext>` and
    `processed_images_data.json` to the output folder and returns the JSON text.  Same control flow and error
    behaviour as the reference."""
    import json
    import cv2
    from tqdm import tqdm
    from cylinder_pose_estimation_b200.iotool import undistort_image, load_camera_data
    left_camera_params, right_camera_params = load_camera_data(json_path)
    if output_folder is None:
        output_folder = folder_path
    if not os.path.exists(output_folder):
        os.makedirs(output_folder)
    valid_exts = ('.png', '.jpg', '.jpeg', '.bmp', '.tif', '.tiff')
    image_files = [f for f in os.listdir(folder_path) if f.lower().endswith(valid_exts)]
    if not image_files:
        print(f"No images found in folder: {folder_path}")
        return
    images_json_data = {}
    for filename in tqdm(image_files, desc="Processing images"):
        image_path = os.path.join(folder_path, filename)
        original_img = cv2.imread(image_path)
        base_name = os.path.splitext(filename)[0]
        try:
            if 'L' in filename:
                undistorted_image = undistort_image(original_img, left_camera_params)
            elif 'R' in filename:
                undistorted_image = undistort_image(original_img, right_camera_params)
            else:
                raise ValueError(f"Unknown camera type in filename: {filename}")
            img, result_json, _, _ = detect_grid(undistorted_image)
            try:
                images_json_data[base_name] = json.loads(result_json)
            except json.JSONDecodeError:
                print(f"Invalid JSON data for image {filename}. Skipping.")
                continue
            cv2.imwrite(os.path.join(output_folder, f"{base_name}_arc{os.path.splitext(filename)[1]}"), img)
        except Exception as e:
            print(f"Error processing {image_path}: {e}")
            images_json_data[base_name] = {'error': str(e)}
            continue
    output_json_path = os.path.join(output_folder, "processed_images_data.json")
    with open(output_json_path, 'w') as json_file:
        json.dump(images_json_data, json_file, indent=4)
    print(f"Data saved to {output_json_path}")
    return json.dumps(images_json_data)
