"""Scene normalisation and the testbed folder layout (first half of the Python stages).

Host-side mirror of the reference's `rnb_neus2/prepare.py` + `rnb_neus2/scaling.py` — same entry points, argument meaning, results and
files (pinned against the reference's own modules: tests/golden/make_prepare_golden.py, tests/test_prepare.py):

    prepare_testbed_data(data, output_folder, logger, scaling_mode="auto", sphere_scale=1.0, margin_px=20) -> dict
    compute_unit_sphere_scaling(points_3d, sphere_scale)                    # scaling.py:9-34
    compute_scaling_from_silhouettes(cameras, masks, sphere_scale, fg_area_ratio)      # :37-103
    compute_scaling_from_silhouettes_v2(cameras, masks, sphere_scale, margin_px, percentile)   # :145-253
    extract_cameras_for_scaling(data)                                       # :256-305

`data` is the standard dict of the reference's dataloaders (dataloaders/base.py): views with `c2w`, `K`, `normal_path`, `albedo_path`,
`mask_path`; optional `landmarks`; `image_width` / `image_height`.  The output is what `rnb_load_dataset_images` / `dataset.load_transforms`
(and the reference's `load_nerf`) read: `transform.json` with `from_na`, `n2w`, per-frame `transform_matrix` / `intrinsic_matrix`, and RGBA
PNGs under `normals/` and `albedos/` whose alpha is the thresholded mask.

This stage is small-array numpy and file conversion — nothing here runs on the GPU, and nothing needs to: it is provided so that the
whole data path up to the training loop exists on this side of the boundary.
"""
import json
import os

import numpy as np

_SCALING_MODES = ("auto", "pcd", "silhouettes", "silhouettes_v2", "cameras", "none")


# ------------------------------------------------------------------------------------------------------------------------- scaling
def _homogeneous_scale(center, factor):
    """4x4 float32 `x -> factor * (x - center)`"""
    m = np.eye(4, dtype=np.float32)
    for a in range(3):
        m[a, a] = factor
        m[a, 3] = -center[a] * factor
    return m


def compute_unit_sphere_scaling(points_3d, sphere_scale=1.0):
    """centre and scale that put the points (99th-percentile inliers around their centroid) inside a sphere of radius sphere_scale"""
    pts = np.asarray(points_3d)
    d0 = np.linalg.norm(pts - np.mean(pts, axis=0), axis=1)
    keep = pts[d0 <= np.percentile(d0, 99)]
    center = np.mean(keep, axis=0)
    factor = sphere_scale / np.max(np.linalg.norm(keep - center, axis=1))
    return center, factor, _homogeneous_scale(center, factor)


def _mask_ray(cam, mask):
    """world-space unit ray through the mask's centre of mass, or None when the mask is empty / degenerate"""
    from scipy.ndimage import center_of_mass
    com = center_of_mass(mask.astype(np.float64))
    if np.any(np.isnan(com)):
        return None
    K = np.array([[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]])
    d = np.linalg.inv(K) @ np.array([com[1], com[0], 1.0])
    n = np.linalg.norm(d)
    if n < 1e-12:
        return None
    return cam["R_cam2world"] @ (d / n)


def _closest_point_to_rays(cameras, masks, skip_degenerate):
    """least-squares point closest to the rays (camera centre -> mask centre of mass): sum_i (I - m m^T) x = sum_i (I - m m^T) o_i"""
    A = np.zeros((3, 3)); b = np.zeros(3)
    for cam, mask in zip(cameras, masks):
        m = _mask_ray(cam, mask)
        if m is None:
            if skip_degenerate:
                continue
            m = np.full(3, np.nan)          # the first variant of the reference does not guard: the NaN propagates, as there
        P = np.eye(3) - np.outer(m, m)
        A += P
        b += P @ cam["center"]
    return A, b


def compute_scaling_from_silhouettes(cameras, masks, sphere_scale=1.0, fg_area_ratio=1.5):
    """centre by triangulating the mask centres of mass; radius from matching the projected sphere area to the foreground area"""
    A, b = _closest_point_to_rays(cameras, masks, skip_degenerate=False)
    center = np.linalg.lstsq(A, b, rcond=None)[0]
    area = 0; inv_depth2 = 0
    for cam, mask in zip(cameras, masks):
        area += mask.sum()
        z = (cam["R_cam2world"].T @ (center - cam["center"]))[2]
        if abs(z) < 1e-8:
            z = 1e-8
        inv_depth2 += (cam["fx"] / z) ** 2
    radius = np.sqrt(fg_area_ratio * area / (np.pi * inv_depth2))
    if radius < 1e-8:
        radius = 1.0
    return center, float(sphere_scale / radius)


def _contour_samples(mask, percentile, max_points):
    """outer contour pixels of the mask (x, y), trimmed to the given percentile of their distance to the centre of mass and thinned to
    about max_points while keeping the convex hull"""
    import cv2
    from scipy.ndimage import center_of_mass
    contours, _ = cv2.findContours((mask > 0.5).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not contours:
        return None
    pts = np.vstack(contours).squeeze().astype(np.float64)
    if pts.ndim == 1:
        return None
    if percentile < 100:
        com = center_of_mass(mask.astype(np.float64))
        if not np.any(np.isnan(com)):
            d = np.linalg.norm(pts - np.array([com[1], com[0]]), axis=1)
            pts = pts[d <= np.percentile(d, percentile)]
            if len(pts) == 0:
                return None
    if len(pts) > max_points:
        try:
            hull = cv2.convexHull(pts.astype(np.float32)).squeeze().astype(np.float64)
            if hull.ndim == 1:
                hull = hull.reshape(1, 2)
        except cv2.error:
            hull = pts[:0]
        thin = pts[::max(1, len(pts) // (max_points - len(hull)))]
        pts = np.vstack([hull, thin]) if len(hull) > 0 else thin
    return pts


def compute_scaling_from_silhouettes_v2(cameras, masks, sphere_scale=1.0, margin_px=20, percentile=99):
    """smallest sphere whose projection contains every (trimmed) mask contour with a pixel margin: Nelder-Mead over the centre, the
    radius follows from the centre"""
    from scipy.optimize import minimize
    A, b = _closest_point_to_rays(cameras, masks, skip_degenerate=True)
    try:
        start = np.linalg.lstsq(A, b, rcond=None)[0]
    except np.linalg.LinAlgError:
        start = np.array([cam["center"] for cam in cameras]).mean(axis=0)
    views = []
    for cam, mask in zip(cameras, masks):
        pts = _contour_samples(mask, percentile, 2000)
        if pts is None:
            continue
        R = cam["R_cam2world"].T
        views.append((cam["fx"], cam["fy"], cam["cx"], cam["cy"], R, -R @ cam["center"], pts))
    if not views:
        return start, float(sphere_scale)

    def radius_needed(c):
        worst = 0.0
        for fx, fy, cx, cy, R, t, pts in views:
            p = R @ c + t
            z = p[2]
            if z <= 1e-6:
                return 1e12
            u = fx * p[0] / z + cx
            v = fy * p[1] / z + cy
            ex = (pts[:, 0] - u) * z / fx
            ey = (pts[:, 1] - v) * z / fy
            r = np.sqrt(ex**2 + ey**2)
            worst = max(worst, r.max() + margin_px * z / ((fx + fy) * 0.5))
        return worst

    best = minimize(radius_needed, start, method="Nelder-Mead", options={"maxiter": 5000, "xatol": 1e-4, "fatol": 1e-6}).x
    return best.astype(np.float32), float(sphere_scale / radius_needed(best))


def extract_cameras_for_scaling(data, mask_folder_path=""):
    """camera dicts (fx, fy, cx, cy, R_cam2world, center) and thresholded float masks of the views that have a readable mask"""
    import cv2
    cameras, masks = [], []
    for view in data["views"]:
        path = view["mask_path"]
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED) if path and os.path.exists(path) else None
        if img is None:
            continue
        if img.ndim == 3:
            img = img[:, :, 0]
        K = view["K"]; c2w = view["c2w"]
        cameras.append({"fx": float(K[0, 0]), "fy": float(K[1, 1]), "cx": float(K[0, 2]), "cy": float(K[1, 2]),
                        "R_cam2world": c2w[:3, :3].astype(np.float64), "center": c2w[:3, 3].astype(np.float64)})
        masks.append((img > (125 if img.dtype == np.uint8 else 30000)).astype(np.float32))
    return cameras, masks


def _scene_scaling(data, mode, sphere_scale, margin_px, logger):
    """(centre, factor, 4x4) by the first source that applies: silhouettes, landmarks, camera centres"""
    center = np.zeros(3, dtype=np.float32); factor = 1.0; matrix = np.eye(4, dtype=np.float32)
    if mode == "none":
        return center, factor, matrix
    done = False
    if mode in ("auto", "silhouettes", "silhouettes_v2"):
        cams, masks = extract_cameras_for_scaling(data)
        if cams and masks:
            if mode in ("auto", "silhouettes_v2"):
                logger.info("Scaling from silhouettes_v2 (min enclosing sphere): {} views".format(len(cams)))
                center, factor = compute_scaling_from_silhouettes_v2(cams, masks, sphere_scale=sphere_scale, margin_px=margin_px)
            else:
                logger.info("Scaling from silhouettes: {} views".format(len(cams)))
                center, factor = compute_scaling_from_silhouettes(cams, masks, sphere_scale=sphere_scale)
            center = center.astype(np.float32)
            matrix = _homogeneous_scale(center, factor)
            done = True
    if not done and mode in ("auto", "pcd"):
        lm = data.get("landmarks")
        if lm is not None and len(lm) > 0:
            logger.info("Scaling from landmarks: {} points".format(len(lm)))
            center, factor, matrix = compute_unit_sphere_scaling(lm, sphere_scale)
            done = True
    if not done and mode in ("auto", "cameras"):
        centers = [v["c2w"][:3, 3].copy() for v in data["views"]]
        if centers:
            logger.info("Scaling from camera centers: {} cameras".format(len(centers)))
            center, factor, matrix = compute_unit_sphere_scaling(np.array(centers, dtype=np.float32), sphere_scale)
            done = True
    if not done:
        raise RuntimeError("No data for scaling. Use scaling_mode='none' to disable.")
    logger.info("Scene center: {}".format(center.tolist()))
    logger.info("Scale factor: {:.6f}".format(factor))
    return center, factor, matrix


# --------------------------------------------------------------------------------------------------------------------------- images
def _alpha_from_mask(mask_path, shape, bits):
    """mask file thresholded to {0, max} of the target bit depth (float masks at 0.5, 8-bit at 125, 16-bit at 30000); all-opaque
    when there is no readable mask"""
    import cv2
    top = 65535 if bits == 16 else 255
    dtype = np.uint16 if bits == 16 else np.uint8
    img = cv2.imread(mask_path, cv2.IMREAD_UNCHANGED) if mask_path and os.path.exists(mask_path) else None
    if img is None:
        return np.ones(shape, dtype=dtype) * top
    if img.ndim == 3:
        img = img[:, :, 0]
    if img.dtype == np.float32:
        on = (img > 0.5).astype(np.float64)
    else:
        on = np.where(img > (125 if img.dtype == np.uint8 else 30000), 1.0, 0.0)
    return (on * top).astype(dtype)


def _rgb_integer(img, signed_unit_range):
    """float EXR content -> uint16 ([-1, 1] normals or [0, 1] albedos), alpha dropped; integer images pass through"""
    if img.dtype == np.float32:
        img = np.clip((img + 1.0) / 2.0, 0, 1) if signed_unit_range else np.clip(img, 0, 1)
        img = (img * 65535).astype(np.uint16)
    if img.ndim == 3 and img.shape[2] == 4:
        img = img[:, :, :3]
    return img


def prepare_testbed_data(data, output_folder, logger, scaling_mode="auto", sphere_scale=1.0, margin_px=20):
    """writes output_folder/{transform.json, normals/NNNNN.png, albedos/NNNNN.png}; returns scene_center, scale_factor, scale_matrix, n2w,
    n_frames (prepare.py:116-257)"""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    center, factor, matrix = _scene_scaling(data, scaling_mode, sphere_scale, margin_px, logger)
    normals_dir = os.path.join(output_folder, "normals"); albedos_dir = os.path.join(output_folder, "albedos")
    os.makedirs(albedos_dir, exist_ok=True); os.makedirs(normals_dir, exist_ok=True)
    frames = []
    for idx, view in enumerate(data["views"]):
        pose = view["c2w"].copy()
        pose[:3, 3] = factor * (pose[:3, 3].copy() - center)
        path = view["normal_path"]
        if not os.path.exists(path):
            logger.warning("Normal not found: {}, skipping".format(path))
            continue
        normal = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if normal is None:
            logger.warning("Cannot read: {}".format(path))
            continue
        normal = _rgb_integer(normal, signed_unit_range=True)
        bits = 16 if normal.dtype == np.uint16 else 8
        albedo = None
        apath = view.get("albedo_path")
        if apath and os.path.exists(apath):
            albedo = cv2.imread(apath, cv2.IMREAD_UNCHANGED)
            if albedo is not None:
                albedo = _rgb_integer(albedo, signed_unit_range=False)
        if albedo is None:
            albedo = (np.ones_like(normal) * (65535 if bits == 16 else 255)).astype(normal.dtype)
        # the alpha channel carries the mask at the bit depth of the image it is attached to (normals and albedos may differ)
        n_alpha = _alpha_from_mask(view.get("mask_path"), normal.shape[:2], bits)
        a_bits = 16 if albedo.dtype == np.uint16 else 8
        a_alpha = n_alpha if a_bits == bits else _alpha_from_mask(view.get("mask_path"), albedo.shape[:2], a_bits)
        name = "{:05d}.png".format(idx)
        cv2.imwrite(os.path.join(normals_dir, name), np.concatenate([normal, n_alpha[:, :, np.newaxis]], axis=-1))
        cv2.imwrite(os.path.join(albedos_dir, name), np.concatenate([albedo, a_alpha[:, :, np.newaxis]], axis=-1))
        frames.append({"albedo_path": "albedos/{}".format(name), "normal_path": "normals/{}".format(name),
                       "transform_matrix": pose.tolist(), "intrinsic_matrix": view["K"].tolist()})
    if not frames:
        raise RuntimeError("No valid frames could be processed")
    logger.info("Processed {} frames".format(len(frames)))
    n2w = np.linalg.inv(matrix)
    doc = {"w": data["image_width"], "h": data["image_height"], "aabb_scale": 1.0, "scale": 0.5, "offset": [0.5, 0.5, 0.5], "from_na": True,
           "n2w": n2w.tolist(), "frames": frames}
    out = os.path.join(output_folder, "transform.json")
    with open(out, "w") as f:
        json.dump(doc, f, indent=4)
    logger.info("Saved transform.json to {}".format(out))
    return {"scene_center": center, "scale_factor": factor, "scale_matrix": matrix, "n2w": n2w, "n_frames": len(frames)}
