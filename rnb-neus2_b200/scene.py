"""Synthetic normal+albedo scenes in the layout the reference loader produces on the device
(reference src/nerf_loader.cu: uint16 RGBA per view, camera-to-world in the NGP frame after nerf_matrix_to_ngp,
nerf_loader.h:186-207; normal-map sign convention of src/testbed_nerf.cu:1507-1509).

No dataset is available offline (DiLiGenT-MV "bear" in BASELINE.json), so benchmarks and tests use an analytic
ellipsoid: closed-form ray intersection, analytic normals, procedural albedo.
"""
import numpy as np

CENTER = np.array([0.5, 0.5, 0.5], np.float32)
AXES = np.array([0.30, 0.22, 0.26], np.float32)


def look_at(pos, target=CENTER):
    """3x4 camera-to-world, column-major flat[12]: columns right, down, forward, position (ray = R @ (x, y, 1))."""
    pos = np.asarray(pos, np.float32)
    fwd = target - pos; fwd /= np.linalg.norm(fwd)
    up = np.array([0, 0, 1], np.float32)
    if abs(fwd @ up) > 0.99:
        up = np.array([0, 1, 0], np.float32)
    right = np.cross(fwd, up); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    return np.concatenate([right, down, fwd, pos]).astype(np.float32)


def camera_ring(n_views, radius=1.15, elevations=(-20.0, 10.0, 40.0)):
    poses = []
    per = max(1, n_views // len(elevations))
    for i in range(n_views):
        el = np.deg2rad(elevations[min(i // per, len(elevations) - 1)])
        az = 2 * np.pi * (i % per) / per + 0.1 * (i // per)
        p = CENTER + radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], np.float32)
        poses.append(look_at(p))
    return poses


def render_view(xform, w, h, focal, with_albedo=True):
    """Returns (normal uint16[h,w,4], albedo uint16[h,w,4] or None)."""
    R = xform[:9].reshape(3, 3).T.astype(np.float32)     # columns are xform[0:3], [3:6], [6:9]
    o = xform[9:12].astype(np.float32)
    xs = (np.arange(w, dtype=np.float32) + 0.5) / w
    ys = (np.arange(h, dtype=np.float32) + 0.5) / h
    dx = (xs - 0.5) * w / focal
    dy = (ys - 0.5) * h / focal
    dcx, dcy = np.meshgrid(dx, dy)
    d = dcx[..., None] * R[:, 0] + dcy[..., None] * R[:, 1] + R[:, 2]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    oc = (o - CENTER) / AXES
    dd = d / AXES
    a = (dd * dd).sum(-1); b = 2 * (dd * oc).sum(-1); c = (oc * oc).sum() - 1.0
    disc = b * b - 4 * a * c
    hit = disc > 0
    t = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a)
    hit &= t > 0
    p = o + t[..., None] * d
    n = (p - CENTER) / (AXES * AXES)
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-12)
    n_cam = n @ R                                        # R^T n
    m = np.stack([n_cam[..., 0], -n_cam[..., 1], -n_cam[..., 2]], -1) * 0.5 + 0.5
    normal = np.zeros((h, w, 4), np.uint16)
    normal[..., :3] = np.where(hit[..., None], np.clip(np.rint(m * 65535.0), 1, 65535), 0).astype(np.uint16)
    normal[..., 3] = np.where(hit, 65535, 0).astype(np.uint16)
    albedo = None
    if with_albedo:
        q = (p - CENTER) * 9.0
        col = 0.2 + 0.7 * (0.5 + 0.5 * np.sin(np.stack([q[..., 0] + 0.3, q[..., 1] * 1.3 + 1.1, q[..., 2] * 0.7 + 2.3], -1)))
        albedo = np.zeros((h, w, 4), np.uint16)
        albedo[..., :3] = np.where(hit[..., None], np.rint(col * 65535.0), 0).astype(np.uint16)
        albedo[..., 3] = normal[..., 3]
    return normal, albedo


def make_scene(n_views=8, w=256, h=256, with_albedo=True, focal_scale=1.37):
    """List of view dicts accepted by Testbed.load_training_data and the oracle binding."""
    views = []
    focal = focal_scale * w
    for xf in camera_ring(n_views):
        nm, al = render_view(xf, w, h, focal, with_albedo)
        views.append(dict(normal=nm, albedo=al, fx=focal, fy=focal, cx=0.5, cy=0.5, xform=xf, w=w, h=h))
    return views
