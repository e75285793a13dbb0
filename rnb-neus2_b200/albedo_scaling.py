"""Albedo scaling between the two training phases (BASELINE configs[2] "--has-albedo two-phase").

Host-side mirror of the reference's `rnb_neus2/albedo_scaling.py` — same entry points, argument meaning and result:

    compute_albedo_scale_ratios(albedo_path, camera_source, mesh_path, n_samples=2000, logger=None) -> (n_views, 3)
    scale_and_save_albedos(albedo_path, output_albedo_path, scale_ratios, bit_depth=None, logger=None)

(`run_with_albedo_scaling`, reference rnb_neus2/pipeline.py:106-170, calls exactly these two between phase 1 and phase 2.)

The reference traces the rays with trimesh on the CPU (`mesh.ray.intersects_location`, albedo_scaling.py:285-289,316-329).  Here
the two ray/mesh queries run on the GPU through `rnb_raymesh_*` (csrc/rnb_raymesh.cu: uniform cell grid + 3D DDA, binary64); all
views are traced in three launches (first hits of every view, then the occlusion rays towards the right and the left neighbours)
instead of 3 x n_views trimesh calls.  The per-point bookkeeping (projection, bilinear look-up, medians) is numpy in binary64, as
in the reference.  There is no CPU path for the ray queries: without the CUDA library / a device this module raises.
"""
import json
import os
from pathlib import Path

import numpy as np

from . import RayMesh

_IMAGE_EXT = (".png", ".exr")


# ---------------------------------------------------------------------------------------------------------------- images / mesh
def load_image(path):
    """(H, W, C) float32 RGB(A); 8/16-bit PNG normalised to [0, 1], float EXR as is (reference image_io.py:14-44)"""
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
    import cv2
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError("Cannot read image: {}".format(path))
    if img.dtype == np.uint8:
        img = img.astype(np.float32) / 255.0
    elif img.dtype == np.uint16:
        img = img.astype(np.float32) / 65535.0
    elif img.dtype != np.float32:
        raise ValueError("Unsupported dtype: {}".format(img.dtype))
    if img.ndim == 3 and img.shape[2] >= 3:
        img = np.ascontiguousarray(np.concatenate([img[:, :, 2::-1], img[:, :, 3:]], axis=2))      # BGR(A) -> RGB(A)
    return img


def save_image(image, path, bit_depth=16):
    """RGB(A) float image in [0, 1] -> 8/16-bit PNG, truncating like the reference (image_io.py:47-70)"""
    import cv2
    img = np.clip(np.nan_to_num(image, nan=0.0), 0.0, 1.0) * float(2 ** bit_depth - 1)
    img = img.astype(np.uint8 if bit_depth == 8 else np.uint16)
    if img.ndim == 3 and img.shape[2] >= 3:
        img = np.ascontiguousarray(np.concatenate([img[:, :, 2::-1], img[:, :, 3:]], axis=2))
    cv2.imwrite(str(path), img, [cv2.IMWRITE_PNG_COMPRESSION, 0])


def load_mesh(path):
    """vertices (n, 3) float32 and triangles (m, 3) uint32 of a Wavefront OBJ or an ASCII PLY (what rnb_save_mesh / the reference's
    save_mesh write: "v x y z [r g b]", "f a//a b//b c//c"; polygons are fanned)."""
    path = str(path)
    with open(path, "r") as f:
        text = f.read()
    if path.lower().endswith(".ply"):
        head, _, body = text.partition("end_header")
        nv = nf = 0
        for line in head.splitlines():
            p = line.split()
            if len(p) == 3 and p[0] == "element":
                if p[1] == "vertex": nv = int(p[2])
                if p[1] == "face": nf = int(p[2])
        if "format ascii" not in head:
            raise ValueError("only ASCII PLY is supported: {}".format(path))
        lines = body.strip().splitlines()
        verts = np.array([ln.split()[:3] for ln in lines[:nv]], dtype=np.float64)
        tris = []
        for ln in lines[nv:nv + nf]:
            p = [int(x) for x in ln.split()]
            for k in range(1, p[0] - 1):
                tris.append((p[1], p[1 + k], p[2 + k]))
        return verts.astype(np.float32), np.array(tris, dtype=np.uint32).reshape(-1, 3)
    vs, fs = [], []
    for line in text.splitlines():
        if line.startswith("v "):
            vs.append(line.split()[1:4])
        elif line.startswith("f "):
            idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
            for k in range(1, len(idx) - 1):
                fs.append((idx[0], idx[k], idx[k + 1]))
    verts = np.array(vs, dtype=np.float64).reshape(-1, 3)
    tris = np.array(fs, dtype=np.int64).reshape(-1, 3)
    tris = np.where(tris < 0, tris + len(verts), tris - 1)          # OBJ indices are 1-based, negative = relative to the end
    return verts.astype(np.float32), tris.astype(np.uint32)


# -------------------------------------------------------------------------------------------------------------------- cameras
def load_K_Rt_from_P(P):
    """intrinsics (4x4) and camera-to-world pose (4x4) of a 3x4 projection matrix (albedo_scaling.py:26-41)"""
    import cv2
    K, R, t = cv2.decomposeProjectionMatrix(P)[:3]
    intr = np.eye(4); intr[:3, :3] = K / K[2, 2]
    pose = np.eye(4, dtype=np.float32); pose[:3, :3] = R.transpose(); pose[:3, 3] = (t[:3] / t[3])[:, 0]
    return intr, pose


def load_cameras_from_npz(npz_path, n_views, logger=None):
    cams = np.load(npz_path)
    K, R, C = [], [], []
    for k in range(n_views):
        intr, pose = load_K_Rt_from_P(cams["world_mat_{}".format(k)][:3, :])
        K.append(intr[:3, :3]); R.append(pose[:3, :3]); C.append(pose[:3, [3]])
    return np.array(K), np.array(R), np.array(C)


def load_cameras_from_transform_json(json_path, albedo_images, logger=None):
    """K, R_c2w (float32) and centres (float32, 3x1) per albedo image; cameras go to world space when the file has `n2w`
    (albedo_scaling.py:128-193)"""
    with open(json_path, "r") as f:
        data = json.load(f)
    n2w = np.array(data["n2w"], dtype=np.float64) if "n2w" in data else None
    frames = {}
    for fr in data["frames"]:
        frames.setdefault(Path(fr["albedo_path"]).stem, fr)          # the reference takes the first frame with a matching stem
    g_fx = data.get("fl_x", None); g_fy = data.get("fl_y", g_fx); g_cx = data.get("cx", None); g_cy = data.get("cy", None)
    K_all, R_all, C_all = [], [], []
    for name in albedo_images:
        fr = frames.get(Path(name).stem)
        if fr is None:
            raise RuntimeError("No frame for albedo image: {}".format(name))
        K = np.eye(3, dtype=np.float32)
        if "intrinsic_matrix" in fr:
            K[:3, :3] = np.array(fr["intrinsic_matrix"], dtype=np.float32)[:3, :3]
        else:
            fx = fr.get("fl_x", g_fx or 500.0)
            K[0, 0] = fx
            K[1, 1] = fr.get("fl_y", g_fy or fx)
            K[0, 2] = fr.get("cx", g_cx or data.get("w", 512) / 2)
            K[1, 2] = fr.get("cy", g_cy or data.get("h", 512) / 2)
        c2w = np.array(fr["transform_matrix"], dtype=np.float64)
        if n2w is not None:
            c2w = n2w @ c2w
        K_all.append(K); R_all.append(c2w[:3, :3].astype(np.float32)); C_all.append(c2w[:3, [3]].astype(np.float32))
    if logger:
        logger.info("Loaded {} cameras from transform.json".format(len(K_all)))
    return np.array(K_all), np.array(R_all), np.array(C_all)


def load_cameras(camera_source, albedo_images, logger=None):
    """auto-detect by suffix like the reference (:196-211); Meshroom .sfm needs pyalicevision, which this image does not have"""
    p = Path(camera_source)
    if p.suffix == ".npz":
        return load_cameras_from_npz(p, len(albedo_images), logger)
    if p.suffix == ".json":
        return load_cameras_from_transform_json(p, albedo_images, logger)
    if p.suffix == ".sfm":
        raise RuntimeError("sfmData cameras need pyalicevision (not available); export a transform.json instead")
    raise ValueError("Unsupported camera format: {}".format(p.suffix))


# --------------------------------------------------------------------------------------------------------------------- ratios
def _bilinear(img, yx):
    """linear interpolation of img[h, w] at fractional (row, col), 0 <= row < h-1, 0 <= col < w-1 (scipy RegularGridInterpolator on
    the integer lattice, :347-358)"""
    y0 = np.floor(yx[:, 0]).astype(np.int64); x0 = np.floor(yx[:, 1]).astype(np.int64)
    fy = yx[:, 0] - y0; fx = yx[:, 1] - x0
    a = img[y0, x0].astype(np.float64); b = img[y0, x0 + 1].astype(np.float64); c = img[y0 + 1, x0].astype(np.float64); d = img[y0 + 1, x0 + 1].astype(np.float64)
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def albedo_scale_ratios_from_arrays(albedos, masks, K_array, R_c2w_array, centers_array, verts, tris, n_samples=2000, choose=None, raymesh=None):
    """the computation of compute_albedo_scale_ratios on arrays already in memory.  `choose(n_pixels, n_good)` picks the sampled pixels
    (default: np.random.choice(n_pixels, n_good, replace=False), the reference's draw on the global generator)."""
    if choose is None:
        choose = lambda n, k: np.random.choice(n, k, replace=False)
    albedos = np.asarray(albedos); masks = np.asarray(masks)
    V, h, w, _ = albedos.shape
    rm = raymesh if raymesh is not None else RayMesh(verts, tris)
    # ---- pixel rays of every view, one first-hit launch (:255-289)
    first = []          # per view: (ray slot within the view, albedo value)
    org_all, dir_all = [], []
    for cam in range(V):
        rows, cols = np.where(masks[cam].astype(bool))
        n_good = min(n_samples, rows.shape[0])
        sel = choose(rows.shape[0], n_good)
        px = np.stack([cols[sel], rows[sel], np.ones(n_good, dtype=np.int64)], axis=0).astype(np.float64)
        K, R, C = K_array[cam], R_c2w_array[cam], centers_array[cam]
        on_ray = (R @ (np.linalg.inv(K) @ px) + C).T
        org = np.tile(C.T, (n_good, 1))
        d = on_ray - org
        d /= np.linalg.norm(d, axis=1)[:, None]
        first.append(albedos[cam, rows[sel], cols[sel], :])
        org_all.append(org.astype(np.float64)); dir_all.append(d)
    counts = np.array([o.shape[0] for o in org_all]); starts = np.concatenate([[0], np.cumsum(counts)])
    t_first, tri_first = rm.first_hit(np.concatenate(org_all), np.concatenate(dir_all))
    hit_all = tri_first != RayMesh.NO_TRI
    loc_all = np.concatenate(org_all) + np.concatenate(dir_all) * np.where(hit_all, t_first, 0.0)[:, None]
    # ---- occlusion rays from every surface point to both neighbour cameras, one launch per side (:298-329)
    blocked = []
    for side in (+1, -1):
        o_l, d_l, tm_l = [], [], []
        for cam in range(V):
            s = slice(starts[cam], starts[cam + 1]); m = hit_all[s]
            nC = centers_array[(cam + side) % V]
            nd = nC.T - loc_all[s][m]
            dist = np.linalg.norm(nd, axis=1, keepdims=True)
            nd = nd / dist
            eps = np.maximum(dist.flatten() * 1e-4, 1e-2)
            o_l.append(loc_all[s][m] + eps[:, None] * nd); d_l.append(nd); tm_l.append(dist.flatten() - eps)
        blocked.append(rm.any_hit(np.concatenate(o_l), np.concatenate(d_l), np.concatenate(tm_l)))
    # ---- per view and side: project the visible points into the neighbour and compare albedos (:331-364)
    ratios = np.zeros((V, n_samples, 3, 2), dtype=np.float32)
    found = np.zeros((V, n_samples, 2), dtype=bool)
    hit_starts = np.concatenate([[0], np.cumsum([hit_all[starts[c]:starts[c + 1]].sum() for c in range(V)])])
    for kk, side in enumerate((+1, -1)):
        for cam in range(V):
            s = slice(starts[cam], starts[cam + 1]); m = hit_all[s]
            index_ray = np.where(m)[0]
            vis = ~blocked[kk][hit_starts[cam]:hit_starts[cam + 1]]
            pts = loc_all[s][m][vis]; idx = index_ray[vis]; va = first[cam][m][vis]
            nb = (cam + side) % V
            pc = R_c2w_array[nb].T @ (pts.T - centers_array[nb])
            pr = (K_array[nb] @ pc).T
            pr = pr[:, :2] / pr[:, 2][:, None]
            ok = (0 <= pr[:, 1]) & (pr[:, 1] < h - 1) & (0 <= pr[:, 0]) & (pr[:, 0] < w - 1)
            pr, idx, va = pr[ok], idx[ok], va[ok]
            an = albedos[nb].astype(np.float32)
            yx = pr[:, ::-1]
            look = np.stack([_bilinear(an[:, :, c], yx) for c in range(3)], axis=1)
            nz = ~np.any(look == 0, axis=1)
            ratios[cam, idx[nz], :, kk] = va[nz] / look[nz]
            found[cam, idx[nz], kk] = True
    # ---- medians over both directions of every neighbouring pair, chained and normalised (:366-383)
    left_r = np.roll(ratios[:, :, :, 1], -1, axis=0); left_f = np.roll(found[:, :, 1], -1, axis=0)
    med = np.zeros((V, 3))
    for cam in range(V):
        med[cam] = np.median(np.concatenate((ratios[cam, found[cam, :, 0], :, 0], 1 / left_r[cam, left_f[cam]]), axis=0), axis=0)
    prop = np.cumprod(np.concatenate([np.ones((1, 3)), med[:-1]], axis=0), axis=0)
    return prop / np.mean(prop, axis=0)


def compute_albedo_scale_ratios(albedo_path, camera_source, mesh_path, n_samples=2000, logger=None):
    """per-view albedo scaling factors from multi-view consistency, (n_views, 3) (albedo_scaling.py:214-383)"""
    log = (lambda m: logger.info(m)) if logger else (lambda m: None)
    names = sorted(f for f in os.listdir(albedo_path) if f.lower().endswith(_IMAGE_EXT))
    log("Loading {} albedo images...".format(len(names)))
    albedos, masks = [], []
    for name in names:
        img = load_image(os.path.join(albedo_path, name))
        masks.append(img[:, :, 3] if img.shape[2] == 4 else np.ones(img.shape[:2]))
        albedos.append(img[:, :, :3])
    K, R, C = load_cameras(camera_source, names, logger)
    log("Loading mesh from {}...".format(mesh_path))
    verts, tris = load_mesh(mesh_path)
    out = albedo_scale_ratios_from_arrays(np.array(albedos), np.array(masks), K, R, C, verts, tris, n_samples)
    log("Scale ratios: {}".format(out))
    return out


def scale_and_save_albedos(albedo_path, output_albedo_path, scale_ratios, bit_depth=None, logger=None):
    """multiply every albedo image by its view's factors and write it under the same name (albedo_scaling.py:386-436)"""
    import cv2
    log = (lambda m: logger.info(m)) if logger else (lambda m: None)
    os.makedirs(output_albedo_path, exist_ok=True)
    names = sorted(f for f in os.listdir(albedo_path) if f.lower().endswith(_IMAGE_EXT))
    if bit_depth is None:
        probe = cv2.imread(os.path.join(albedo_path, names[0]), cv2.IMREAD_UNCHANGED)
        bit_depth = 8 if probe.dtype == np.uint8 else 16
        log("Auto-detected bit depth: {}".format(bit_depth))
    for i, name in enumerate(names):
        img = load_image(os.path.join(albedo_path, name))
        alpha = img[:, :, 3] if img.shape[2] == 4 else np.ones(img.shape[:2])
        rgba = np.concatenate((img[:, :, :3] * scale_ratios[i], alpha[:, :, None]), axis=-1)
        save_image(rgba, os.path.join(output_albedo_path, name), bit_depth=bit_depth)
        log("Saved {}/{}: {}".format(i + 1, len(names), name))
