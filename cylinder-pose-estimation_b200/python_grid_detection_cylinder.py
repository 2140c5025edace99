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
    """Folder CLI of the reference (python_grid_detection_cylinder.py:12-64): imread -> undistort with the left / right camera chosen
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
        if 'L' in filename:
            undistorted_image = undistort_image(original_img, left_camera_params)
        elif 'R' in filename:
            undistorted_image = undistort_image(original_img, right_camera_params)
        else:
            raise ValueError(f"Unknown camera type in filename: {filename}")
        img, result_json, _, _ = detect_grid(undistorted_image)
        base_name = os.path.splitext(filename)[0]
        try:
            images_json_data[base_name] = json.loads(result_json)
        except json.JSONDecodeError:
            print(f"Invalid JSON data for image {filename}. Skipping.")
            continue
        cv2.imwrite(os.path.join(output_folder, f"{base_name}_arc{os.path.splitext(filename)[1]}"), img)
    output_json_path = os.path.join(output_folder, "processed_images_data.json")
    with open(output_json_path, 'w') as json_file:
        json.dump(images_json_data, json_file, indent=4)
    print(f"Data saved to {output_json_path}")
    return json.dumps(images_json_data)
