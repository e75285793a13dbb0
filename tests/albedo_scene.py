"""Synthetic scene for the albedo-scaling tests: a sphere mesh, a ring of pinhole cameras looking at it, and albedo images that show one
view-independent texture multiplied by a per-view, per-channel gain.  The stage under test must recover the gains."""
import numpy as np


def icosphere(subdiv=3, radius=1.0, center=(0.0, 0.0, 0.0)):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}; nf = []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius + np.array(center)).astype(np.float32), np.array(f, dtype=np.uint32)


def texture(p):
    """smooth, strictly positive RGB texture of the surface point (n, 3) -> (n, 3)"""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return np.stack([0.45 + 0.25 * np.sin(2.1 * x + 0.3) * np.cos(1.7 * y), 0.5 + 0.2 * np.sin(1.3 * y - 0.8 * z), 0.4 + 0.25 * np.cos(1.9 * z + 0.6 * x)], axis=1)


def ring_cameras(n_views, w, h, dist=3.2, focal=None, elev=0.25):
    """K (float32), R_c2w (float32, columns right/down/forward), centres (float32 3x1)"""
    focal = focal or 1.25 * w
    Ks, Rs, Cs = [], [], []
    for i in range(n_views):
        a = 2 * np.pi * i / n_views
        C = np.array([dist * np.cos(a), dist * np.sin(a), dist * elev * np.sin(2 * a + 0.4)])
        fwd = -C / np.linalg.norm(C)
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0])); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        Ks.append(np.array([[focal, 0, w / 2], [0, focal, h / 2], [0, 0, 1]], dtype=np.float32))
        Rs.append(np.stack([right, down, fwd], axis=1).astype(np.float32)); Cs.append(C.reshape(3, 1).astype(np.float32))
    return np.array(Ks), np.array(Rs), np.array(Cs)


def render_views(Ks, Rs, Cs, w, h, gains, radius=1.0):
    """albedos [V, h, w, 3] float32 = gain_v * texture(first hit on the analytic sphere), masks [V, h, w]"""
    V = len(Ks)
    alb = np.zeros((V, h, w, 3), np.float32); msk = np.zeros((V, h, w), np.float32)
    uu, vv = np.meshgrid(np.arange(w), np.arange(h))
    px = np.stack([uu.ravel(), vv.ravel(), np.ones(w * h)], axis=0).astype(np.float64)
    for i in range(V):
        d = (Rs[i].astype(np.float64) @ (np.linalg.inv(Ks[i].astype(np.float64)) @ px)).T
        d /= np.linalg.norm(d, axis=1)[:, None]
        o = Cs[i].astype(np.float64).reshape(1, 3)
        b = (d * o).sum(1); c = (o * o).sum() - radius * radius
        disc = b * b - c
        hit = disc > 0
        t = -b - np.sqrt(np.where(hit, disc, 0.0))
        p = o + d * t[:, None]
        col = texture(p) * gains[i][None, :]
        alb[i] = np.where(hit[:, None], col, 0.0).reshape(h, w, 3).astype(np.float32)
        msk[i] = hit.reshape(h, w).astype(np.float32)
    return alb, msk


def expected_factors(gains):
    inv = 1.0 / np.asarray(gains, dtype=np.float64)
    inv = inv / inv[0]
    return inv / inv.mean(axis=0)


# ---- the fixed scene of the albedo-scaling tests, as files (the reference's inputs: albedo folder, transform.json, mesh) --------------
V, W, H = 8, 96, 80
GAINS = np.array([[1.0, 1.0, 1.0], [0.8, 0.9, 1.1], [1.2, 0.7, 0.95], [0.9, 1.1, 1.3], [1.05, 0.85, 0.75], [0.7, 1.2, 1.0], [1.3, 1.0, 0.9], [0.95, 0.95, 1.15]])


def fixed_scene():
    verts, tris = icosphere(3)
    K, R, Cc = ring_cameras(V, W, H)
    alb, msk = render_views(K, R, Cc, W, H, GAINS)
    return verts, tris, K, R, Cc, alb, msk


def write_scene_files(root):
    """16-bit RGBA PNGs (alpha = mask) under root/albedos, root/transform.json (with an n2w that is not the identity), root/mesh_0.obj in
    world space.  Returns the image names and a hash of everything written."""
    import hashlib
    import json
    import os
    import cv2
    verts, tris, K, R, Cc, alb, msk = fixed_scene()
    os.makedirs(os.path.join(root, "albedos"), exist_ok=True)
    n2w = np.diag([1.5, 1.5, 1.5, 1.0]); n2w[:3, 3] = [0.3, -0.2, 0.1]
    w2n = np.linalg.inv(n2w)
    h = hashlib.sha256()
    names, frames = [], []
    for i in range(V):
        rgba = np.concatenate([alb[i], msk[i][:, :, None]], axis=2)
        img = (np.clip(rgba, 0.0, 1.0) * 65535.0).astype(np.uint16)
        name = "%05d.png" % i
        cv2.imwrite(os.path.join(root, "albedos", name), np.ascontiguousarray(np.concatenate([img[:, :, 2::-1], img[:, :, 3:]], axis=2)), [cv2.IMWRITE_PNG_COMPRESSION, 0])
        h.update(img.tobytes())
        c2w = np.eye(4); c2w[:3, :3] = R[i]; c2w[:3, 3] = Cc[i][:, 0]
        c2n = w2n @ c2w                                   # the file holds normalised-space cameras; n2w takes them back to the world
        frames.append({"albedo_path": "albedos/" + name, "normal_path": "normals/" + name, "transform_matrix": c2n.tolist(), "intrinsic_matrix": K[i].tolist()})
        names.append(name)
    tj = json.dumps({"w": W, "h": H, "n2w": n2w.tolist(), "frames": frames})
    open(os.path.join(root, "transform.json"), "w").write(tj); h.update(tj.encode())
    with open(os.path.join(root, "mesh_0.obj"), "w") as f:
        for v in verts:
            f.write("v %0.5f %0.5f %0.5f 0.500 0.500 0.500\n" % tuple(v))
        for t in tris:
            f.write("f %d//%d %d//%d %d//%d\n" % (t[0] + 1, t[0] + 1, t[1] + 1, t[1] + 1, t[2] + 1, t[2] + 1))
    h.update(open(os.path.join(root, "mesh_0.obj"), "rb").read())
    return {"names": names, "sha256": h.hexdigest()}


# ---- input of the prepare stage: the standard dict of the reference's dataloaders (dataloaders/base.py) over files on disk ------------
def write_prepare_inputs(root):
    """normal maps (16-bit RGB PNG), albedos (8-bit for odd views, 16-bit otherwise, none for view 5), 8-bit masks, view 6 without its
    normal file (must be skipped), landmarks on the sphere; returns the data dict"""
    import os
    import cv2
    verts, tris, K, R, Cc, alb, msk = fixed_scene()
    src = os.path.join(root, "src"); os.makedirs(src, exist_ok=True)
    rng = np.random.default_rng(21)
    views = []
    for i in range(V):
        nrm = (np.clip(rng.normal(0.5, 0.2, size=(H, W, 3)), 0, 1) * 65535).astype(np.uint16) * (msk[i][:, :, None] > 0)
        npath = os.path.join(src, "n%02d.png" % i)
        if i != 6:
            cv2.imwrite(npath, nrm)
        apath = None
        if i != 5:
            apath = os.path.join(src, "a%02d.png" % i)
            a = np.clip(alb[i][:, :, ::-1], 0, 1)
            cv2.imwrite(apath, (a * 255).astype(np.uint8) if i % 2 else (a * 65535).astype(np.uint16))
        mpath = os.path.join(src, "m%02d.png" % i)
        cv2.imwrite(mpath, (msk[i] * 255).astype(np.uint8))
        c2w = np.eye(4); c2w[:3, :3] = R[i]; c2w[:3, 3] = Cc[i][:, 0] * 37.0 + np.array([120.0, -40.0, 15.0])      # an unnormalised world
        K4 = np.eye(4); K4[:3, :3] = K[i]
        views.append({"c2w": c2w, "K": K4, "normal_path": npath, "albedo_path": apath, "mask_path": mpath if i != 2 else None, "pose_id": str(i)})
    lm = rng.normal(size=(500, 3)); lm = lm / np.linalg.norm(lm, axis=1)[:, None] * 37.0 + np.array([120.0, -40.0, 15.0])
    lm[:5] *= 3.0                                                    # outliers for the percentile cut
    return {"views": views, "landmarks": lm, "image_width": W, "image_height": H, "scale_mat": None}
