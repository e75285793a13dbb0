"""Dataset ingest: the transform.json half of load_nerf (reference src/nerf_loader.cu:355-700) on the host.

    meta = load_transforms("scene/transform.json")      # per-view intrinsics, camera matrices in the NGP frame, file paths, n2w, scale/offset
    testbed.load_training_data_dir("scene/")            # + PNG decode on host threads and upload (csrc/rnb_dataset.cu)

Every arithmetic step is done in binary32 in the reference's order so that the values handed to the kernels are the same bits:
`focal = float(K[0][0])`, `principal = float(K[0][2]) / float(w)`, translation `t * scale + offset`, the axis conventions of
NerfDataset::nerf_matrix_to_ngp (nerf_loader.h:180-202: flip y/z columns, then undo it for `from_na`, negate x/z for Mitsuba, cycle
the rows xyz <- yzx otherwise).  Only what the RNb-NeuS2 path reads is handled: frames with `normal_path`, `albedo_path`,
`intrinsic_matrix`, `transform_matrix`; no distortion, rolling shutter, depth, rays or environment maps."""
import json
import os

import numpy as np

NERF_SCALE = np.float32(0.33)          # nerf_loader.h:  default scale / offset of a NerfDataset
f32 = np.float32


def nerf_matrix_to_ngp(m, scale, offset, from_na, from_mitsuba=False):
    """m: [3,4] float32 camera-to-world from transform.json -> NGP frame (nerf_loader.h:180-202)."""
    r = np.array(m, np.float32, copy=True)
    r[:, 1] *= f32(-1); r[:, 2] *= f32(-1)
    r[:, 3] = (r[:, 3] * f32(scale)).astype(np.float32) + np.asarray(offset, np.float32)
    if from_na:
        r[:, 1] *= f32(-1); r[:, 2] *= f32(-1)
    elif from_mitsuba:
        r[:, 0] *= f32(-1); r[:, 2] *= f32(-1)
    else:
        r = r[[1, 2, 0], :]
    return r


def load_transforms(path):
    """path: transform.json or the scene directory.  Returns dict(views=[...], scale, offset, aabb_scale, from_na, n2w_s, n2w_t, w, h)."""
    if os.path.isdir(path):
        path = os.path.join(path, "transform.json")
    base = os.path.dirname(os.path.abspath(path))
    with open(path) as f:
        j = json.load(f)
    if "frames" not in j or not j["frames"]:
        raise ValueError("No training images were found for NeRF training!")
    from_mitsuba = "normal_mts_args" in j
    from_na = "from_na" in j                                   # presence, not value (nerf_loader.cu:392-394)
    scale = NERF_SCALE; offset = np.array([0.5, 0.5, 0.5], np.float32)
    if from_mitsuba:
        scale = f32(0.66); offset = np.full(3, f32(0.25) * scale, np.float32)
    if "scale" in j:
        scale = f32(j["scale"])
    if "offset" in j:
        o = j["offset"]
        offset = np.array(o if isinstance(o, list) else [o, o, o], np.float32)
    if "aabb" in j:
        a = np.array(j["aabb"], np.float32)
        length = max(f32(0.000001), np.abs(a[1] - a[0]).max())
        scale = f32(1.0) / f32(length)
        offset = (((a[1] + a[0]) * f32(0.5)) * -scale + f32(0.5)).astype(np.float32)
    w, h = j["w"], j["h"]
    n2w_s = f32(1.0); n2w_t = np.zeros(3, np.float32)
    if "n2w" in j:
        n2w_t = np.array([j["n2w"][m][3] for m in range(3)], np.float32); n2w_s = f32(j["n2w"][0][0])
    views = []
    for i, fr in enumerate(j["frames"]):
        m = np.array(fr["transform_matrix_start"] if "transform_matrix_start" in fr else fr["transform_matrix"], np.float64)[:3, :4].astype(np.float32)
        x = nerf_matrix_to_ngp(m, scale, offset, from_na, from_mitsuba)
        K = fr["intrinsic_matrix"]

        def img(key, required):
            p = fr.get(key, "") or ""
            if p == "":
                if required:      # the reference falls back to a per-part default name (nerf_loader.cu:588-592) that the RNb pipeline never writes
                    raise ValueError("transform.json frame %d has no %s" % (i, key))
                return None
            p = os.path.join(base, p)
            if os.path.splitext(p)[1] == "":
                p += ".png"
            return p.replace("\\", "/")             # every single backslash, like replace(..., '\\', '/') in nerf_loader.cu:604,645
        views.append(dict(normal_path=img("normal_path", True), albedo_path=img("albedo_path", False), w=int(w), h=int(h),
                          fx=f32(K[0][0]), fy=f32(K[1][1]), cx=f32(K[0][2]) / f32(w), cy=f32(K[1][2]) / f32(h),
                          xform=x.T.reshape(-1).copy()))           # column-major 3x4, as rnb_view.xform
    return dict(views=views, scale=float(scale), offset=tuple(float(v) for v in offset), aabb_scale=int(j.get("aabb_scale", 1)), from_na=from_na,
                n2w_s=float(n2w_s), n2w_t=tuple(float(v) for v in n2w_t), w=int(w), h=int(h))
