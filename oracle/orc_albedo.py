"""ORACLE (test infrastructure only) — numpy restatement of the reference's albedo-scaling stage.

Follows rnb_neus2/albedo_scaling.py (reference):
  * cameras from transform.json with the n2w world conversion          :128-193
  * per view: sampled masked pixels -> pixel rays -> first mesh hit     :255-289
  * both neighbours: occlusion test, projection, bilinear albedo lookup,
    per-channel ratio                                                   :298-364
  * medians, chained product, normalisation by the mean                 :366-383
  * scale_and_save_albedos                                              :386-436

Pinning.  The reference delegates the two ray/mesh queries to trimesh (`mesh.ray.intersects_location`; `trimesh`, unpinned in the
reference's setup.py:14, NOT in this image).  Everything else of the module IS run here: tests/golden/make_albedo_golden.py
imports the reference's rnb_neus2/albedo_scaling.py with a stand-in `trimesh` module that has exactly the surface the
reference uses (`load_mesh(path).ray.intersects_location(...)`) and answers with brute force over all triangles in binary64
(Moeller-Trumbore, the published algorithm behind trimesh's default `ray_triangle` engine).  The resulting fixture
(tests/golden/ref_albedo_scaling.npz: ratios for two seeds, cameras, the scaled PNGs) pins this restatement to 1e-9 and the
product's host logic to 1e-9 / byte-identical files (tests/test_albedo_scaling.py).  **Unpinned: trimesh's own intersector**
(its tolerance handling on edges and its hit ordering); the known-answer scene (per-view gains applied to a view-independent
texture are recovered) bounds what that could change.

Only tests/ may import this module.
"""
import numpy as np
from scipy.interpolate import RegularGridInterpolator

NO_TRI = 0xFFFFFFFF


def ray_params_bruteforce(verts, tris, origins, dirs, chunk=256):
    """t[i, j] = parameter at which ray i crosses triangle j (nan where it does not), binary64, two-sided."""
    v0 = verts[tris[:, 0]].astype(np.float64)
    e1 = verts[tris[:, 1]].astype(np.float64) - v0
    e2 = verts[tris[:, 2]].astype(np.float64) - v0
    n = origins.shape[0]
    out = np.full((n, tris.shape[0]), np.nan)
    eps = 1e-9
    for s in range(0, n, chunk):
        o = origins[s:s + chunk, None, :].astype(np.float64)
        d = dirs[s:s + chunk, None, :].astype(np.float64)
        p = np.cross(d, e2[None])
        det = np.einsum("tk,rtk->rt", e1, p)
        ok = np.abs(det) > 1e-300
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        sv = o - v0[None]
        u = np.einsum("rtk,rtk->rt", sv, p) * inv
        q = np.cross(sv, e1[None])
        v = np.einsum("rtk,rtk->rt", np.broadcast_to(d, q.shape), q) * inv
        t = np.einsum("tk,rtk->rt", e2, q) * inv
        hit = ok & (u >= -eps) & (u <= 1 + eps) & (v >= -eps) & (u + v <= 1 + eps)
        out[s:s + chunk] = np.where(hit, t, np.nan)
    return out


def first_hit(verts, tris, origins, dirs):
    """closest crossing with t > 0: (t, tri) with t = inf / tri = NO_TRI for a miss"""
    t = ray_params_bruteforce(verts, tris, origins, dirs)
    t = np.where(t > 0, t, np.inf)
    t = np.where(np.isnan(t), np.inf, t)
    j = np.argmin(t, axis=1)
    tm = t[np.arange(t.shape[0]), j]
    return tm, np.where(np.isfinite(tm), j, NO_TRI).astype(np.uint32)


def any_hit(verts, tris, origins, dirs, t_max):
    """is there a crossing with 0 < t < t_max[i]"""
    t = ray_params_bruteforce(verts, tris, origins, dirs)
    with np.errstate(invalid="ignore"):
        return np.any((t > 0) & (t < t_max[:, None]), axis=1)


def cameras_from_transform(data, stems):
    """K (float32 3x3), R_c2w (float32 3x3), centre (float32 3x1) per albedo image stem (:128-193)"""
    import os
    n2w = np.array(data["n2w"], dtype=np.float64) if "n2w" in data else None
    by_stem = {os.path.splitext(os.path.basename(f["albedo_path"]))[0]: f for f in data["frames"]}
    Ks, Rs, Cs = [], [], []
    for s in stems:
        f = by_stem[s]
        K = np.eye(3, dtype=np.float32)
        if "intrinsic_matrix" in f:
            K[:3, :3] = np.array(f["intrinsic_matrix"], dtype=np.float32)[:3, :3]
        else:
            fx = f.get("fl_x", data.get("fl_x") or 500.0)
            K[0, 0] = fx
            K[1, 1] = f.get("fl_y", data.get("fl_y", data.get("fl_x")) or fx)
            K[0, 2] = f.get("cx", data.get("cx") or data.get("w", 512) / 2)
            K[1, 2] = f.get("cy", data.get("cy") or data.get("h", 512) / 2)
        c2w = np.array(f["transform_matrix"], dtype=np.float64)
        if n2w is not None:
            c2w = n2w @ c2w
        Ks.append(K); Rs.append(c2w[:3, :3].astype(np.float32)); Cs.append(c2w[:3, [3]].astype(np.float32))
    return np.array(Ks), np.array(Rs), np.array(Cs)


def albedo_scale_ratios(albedos, masks, Ks, Rs, Cs, verts, tris, n_samples, choose):
    """albedos [V,h,w,3] in [0,1], masks [V,h,w]; `choose(n_pixels, n_good)` returns the sampled pixel subset (the reference
    calls np.random.choice(n_pixels, n_good, replace=False) on the global generator).  Returns the (V, 3) factors (:241-383)."""
    V, h, w, _ = albedos.shape
    ratios = np.zeros((V, n_samples, 3, 2), dtype=np.float32)
    found = np.zeros((V, n_samples, 2), dtype=bool)
    for cam in range(V):
        rows, cols = np.where(masks[cam].astype(bool))
        pix = np.stack([cols, rows], axis=1)
        vals = albedos[cam, rows, cols, :]
        K, R, C = Ks[cam], Rs[cam], Cs[cam]
        n_good = min(n_samples, pix.shape[0])
        sel = choose(pix.shape[0], n_good)
        pix, vals = pix[sel], vals[sel]
        org = np.tile(C.T, (n_good, 1))
        on_ray = (R @ (np.linalg.inv(K) @ np.concatenate((pix, np.ones((n_good, 1))), axis=1).T) + C).T
        dirs = on_ray - org
        dirs /= np.linalg.norm(dirs, axis=1)[:, None]
        t, tri = first_hit(verts, tris, org, dirs)
        index_ray = np.where(tri != NO_TRI)[0]
        loc = org[index_ray].astype(np.float64) + dirs[index_ray].astype(np.float64) * t[index_ray, None]
        vals = vals[index_ray]
        for kk, nb in enumerate([(cam + 1) % V, (cam - 1) % V]):
            nK, nR, nC = Ks[nb], Rs[nb], Cs[nb]
            nd = nC.T - loc
            dist = np.linalg.norm(nd, axis=1, keepdims=True)
            nd = nd / dist
            eps = np.maximum(dist.flatten() * 1e-4, 1e-2)
            no = loc + eps[:, None] * nd
            blocked = any_hit(verts, tris, no, nd, dist.flatten() - eps) if len(loc) else np.zeros(0, dtype=bool)
            pts = loc[~blocked]; idx = index_ray[~blocked]; va = vals[~blocked]
            pc = nR.T @ (pts.T - nC)
            pr = (nK @ pc).T
            pr /= pr[:, 2][:, None]
            pr = pr[:, :2]
            ok = (0 <= pr[:, 1]) & (pr[:, 1] < h - 1) & (0 <= pr[:, 0]) & (pr[:, 0] < w - 1)
            pr, idx, va = pr[ok], idx[ok], va[ok]
            an = albedos[nb].astype(np.float32)
            yx = np.stack([pr[:, 1], pr[:, 0]], axis=1)
            look = np.stack([RegularGridInterpolator((np.arange(h), np.arange(w)), an[:, :, c])(yx) for c in range(3)], axis=1)
            nz = ~np.any(look == 0, axis=1)
            ratios[cam, idx[nz], :, kk] = va[nz] / look[nz]
            found[cam, idx[nz], kk] = True
    med = np.zeros((V, 3))
    left_r = np.roll(ratios[:, :, :, 1], -1, axis=0); left_f = np.roll(found[:, :, 1], -1, axis=0)
    for cam in range(V):
        allr = np.concatenate((ratios[cam, found[cam, :, 0], :, 0], 1 / left_r[cam, left_f[cam]]), axis=0)
        med[cam] = np.median(allr, axis=0)
    prop = np.ones((V, 3))
    for i in range(V - 1):
        prop[i + 1] = prop[i] * med[i]
    return prop / np.mean(prop, axis=0)
