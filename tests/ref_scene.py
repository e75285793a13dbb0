"""Scene files for the reference harness (oracle/_ref/bin/ref_harness, the reference's own nerf_loader.cu) and the
reader for what the harness dumps.  TEST INFRASTRUCTURE.

write_scene() stores the synthetic views of rnb-neus2_b200/scene.py in the on-disk layout the reference pipeline produces
(reference rnb_neus2/prepare.py:215-244: transform.json with from_na/scale 0.5/offset 0.5, normals/*.png and
albedos/*.png as 16-bit RGBA).  load_dataset_dump() reads dataset.bin back (what the reference loader actually put on the
device), so that both sides of a parity check see bit-identical pixels and camera parameters.
"""
import json
import os
import zlib
import struct
import numpy as np


def _png16_rgba(path, img):
    """Minimal 16-bit RGBA PNG writer (no cv2 dependency on the GPU box)."""
    h, w, _ = img.shape
    raw = np.ascontiguousarray(img).astype(">u2").tobytes()
    stride = w * 8
    rows = b"".join(b"\x00" + raw[y * stride:(y + 1) * stride] for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(rows, 1)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def _png_job(a):
    _png16_rgba(*a)


def write_scene(out_dir, views, workers=1, n2w=None):
    os.makedirs(os.path.join(out_dir, "normals"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "albedos"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "output"), exist_ok=True)
    frames = []; jobs = []
    w, h = views[0]["w"], views[0]["h"]
    for i, v in enumerate(views):
        name = "%05d.png" % i
        alb = v["albedo"]
        if alb is None:
            alb = np.full_like(v["normal"], 65535); alb[..., 3] = v["normal"][..., 3]
        jobs.append((os.path.join(out_dir, "normals", name), v["normal"])); jobs.append((os.path.join(out_dir, "albedos", name), alb))
        xf = np.asarray(v["xform"], np.float64)
        R = xf[:9].reshape(3, 3).T            # columns right, down, forward (NGP frame; from_na undoes the y/z flip)
        t = (xf[9:12] - 0.5) / 0.5            # nerf_matrix_to_ngp: t * scale + offset (nerf_loader.h:186-190)
        c2w = np.eye(4); c2w[:3, :3] = R; c2w[:3, 3] = t
        K = [[float(v["fx"]), 0.0, float(v["cx"]) * w], [0.0, float(v["fy"]), float(v["cy"]) * h], [0.0, 0.0, 1.0]]
        frames.append({"albedo_path": "albedos/" + name, "normal_path": "normals/" + name, "transform_matrix": c2w.tolist(), "intrinsic_matrix": K})
    if workers > 1:
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=workers) as ex:
            list(ex.map(_png_job, jobs))
    else:
        for j in jobs:
            _png_job(j)
    tj = {"w": w, "h": h, "aabb_scale": 1.0, "scale": 0.5, "offset": [0.5, 0.5, 0.5], "from_na": True, "n2w": (np.eye(4) if n2w is None else np.asarray(n2w)).tolist(), "frames": frames}
    with open(os.path.join(out_dir, "transform.json"), "w") as f:
        json.dump(tj, f)


def load_dataset_dump(path):
    """dataset.bin of ref_harness -> list of view dicts (normal/albedo uint16[h,w,4], fx, fy, cx, cy, xform[12], w, h)."""
    b = open(path, "rb").read()
    views = []; o = 0
    while o < len(b):
        w, h = struct.unpack_from("<ii", b, o); o += 8
        fx, fy, cx, cy = struct.unpack_from("<4f", b, o); o += 16
        xf = np.frombuffer(b, np.float32, 12, o).copy(); o += 48
        n = w * h * 4
        nm = np.frombuffer(b, np.uint16, n, o).reshape(h, w, 4).copy(); o += n * 2
        al = np.frombuffer(b, np.uint16, n, o).reshape(h, w, 4).copy(); o += n * 2
        views.append(dict(normal=nm, albedo=al, fx=fx, fy=fy, cx=cx, cy=cy, xform=xf, w=w, h=h))
    return views


def read_meta(path):
    d = {}
    for line in open(path):
        if "=" in line:
            k, v = line.strip().split("=", 1); d[k] = v
    return d


def small_network_config(path, n_levels=8, log2_hashmap=14, n_neurons=32, rgb_hidden=1):
    """BASELINE configs[0] as a testbed --config file: the shipped configs/nerf/base.json with four values changed."""
    base = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "configs", "nerf", "base.json")
    cfg = json.load(open(base))
    cfg["encoding"]["n_levels"] = n_levels
    cfg["encoding"]["log2_hashmap_size"] = log2_hashmap
    cfg["network"]["n_neurons"] = n_neurons
    cfg["rgb_network"]["n_neurons"] = n_neurons
    cfg["rgb_network"]["n_hidden_layers"] = rgb_hidden
    json.dump(cfg, open(path, "w"), indent=1)
    return path
