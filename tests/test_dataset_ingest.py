"""Dataset ingest on the CPU: the library's PNG decoder (host code behind the C ABI, stbi_load_16(..., 4)
conventions) and the transform.json reader (rnb-neus2_b200/dataset.py, nerf_loader.cu:355-700).

The reference side of this path was pinned on the B200 box by tests/ref_pin.py: the reference's own loader, run on a scene written
by tests/ref_scene.py, hands back exactly the pixels and camera matrices the scene was written from (dataset_roundtrip in
tests/golden/ref_pin_summary_*.json: pixels_equal, xform_max_abs 0.0).  So "our loader on that scene == the arrays it was written
from" is "our loader == the reference's loader"."""
import json
import os
import struct
import zlib

import numpy as np
import pytest
import rnb_loader
import ref_scene

pkg = rnb_loader.load_package()
from rnb_neus2_b200 import dataset as ds      # noqa: E402


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)


def _write_png(path, rows, w, h, depth, ctype, extra=b"", filters=None, interlace=0):
    """rows: list of h byte strings (unfiltered scanlines); filters: per-row PNG filter type applied here."""
    bpp = max(1, {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype] * depth // 8)
    raw = b""; prev = bytes(len(rows[0]))
    for y, r in enumerate(rows):
        ft = 0 if filters is None else filters[y % len(filters)]
        a = np.frombuffer(r, np.uint8).astype(np.int32); b = np.frombuffer(prev, np.uint8).astype(np.int32)
        left = np.concatenate([np.zeros(bpp, np.int32), a[:-bpp]]); ul = np.concatenate([np.zeros(bpp, np.int32), b[:-bpp]])
        if ft == 0:
            f = a
        elif ft == 1:
            f = a - left
        elif ft == 2:
            f = a - b
        elif ft == 3:
            f = a - ((left + b) >> 1)
        else:
            p = left + b - ul; pa = np.abs(p - left); pb = np.abs(p - b); pc = np.abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, b, ul)); f = a - pred
        raw += bytes([ft]) + (f & 255).astype(np.uint8).tobytes(); prev = r
    ihdr = struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)
    data = zlib.compress(raw, 6)
    half = len(data) // 2
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + extra + _chunk(b"IDAT", data[:half]) + _chunk(b"IDAT", data[half:]) + _chunk(b"IEND", b""))


def _expected_rgba16(a, depth, ctype, key=None):
    """a: [h, w, channels] samples -> stbi_load_16(..., 4) result"""
    a = a.astype(np.uint32)
    wide = a if depth == 16 else a * 257
    h, w = a.shape[:2]
    out = np.zeros((h, w, 4), np.uint16)
    if ctype in (0, 4):
        out[..., :3] = wide[..., :1]
        out[..., 3] = wide[..., 1] if ctype == 4 else 65535
    else:
        out[..., :3] = wide[..., :3]
        out[..., 3] = wide[..., 3] if ctype == 6 else 65535
    if key is not None:
        kw = np.array(key, np.uint32) * (1 if depth == 16 else 257)
        hit = np.all(wide[..., :len(key)] == kw, axis=-1)
        out[..., 3] = np.where(hit, 0, out[..., 3])
    return out


@pytest.mark.parametrize("depth", [8, 16])
@pytest.mark.parametrize("ctype", [0, 2, 4, 6])
def test_png_decoder_all_formats_and_filters(tmp_path, depth, ctype):
    rs = np.random.RandomState(depth * 10 + ctype)
    w, h, ch = 37, 23, {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    hi = 255 if depth == 8 else 65535
    a = rs.randint(0, hi + 1, (h, w, ch)).astype(np.uint16)
    a[5:12] = ((np.arange(w)[None, :, None] * 3 + np.arange(7)[:, None, None]).astype(np.int64) % (hi + 1)).astype(np.uint16)      # smooth rows: filters matter
    rows = [(a[y].astype(">u2").tobytes() if depth == 16 else a[y].astype(np.uint8).tobytes()) for y in range(h)]
    path = tmp_path / "t.png"
    _write_png(path, rows, w, h, depth, ctype, filters=[0, 1, 2, 3, 4])
    got = pkg.load_png_rgba16(path)
    assert got.dtype == np.uint16 and got.shape == (h, w, 4)
    assert np.array_equal(got, _expected_rgba16(a, depth, ctype))
    try:
        import cv2
    except ImportError:
        return
    ref = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)             # an independent decoder on the same file
    ref = ref.reshape(h, w, -1)
    if ref.shape[2] >= 3:
        ref = ref[..., [2, 1, 0] + ([3] if ref.shape[2] == 4 else [])]
    if ctype == 4 and ref.shape[2] != 2:
        return                                                     # OpenCV expands grey+alpha its own way
    assert np.array_equal(ref.astype(np.uint16), a[..., :ref.shape[2]])


def test_png_colour_key_palette_and_errors(tmp_path):
    w, h = 8, 4
    g = (np.arange(w * h).reshape(h, w, 1) % 7 * 30).astype(np.uint16)
    p = tmp_path / "k.png"
    _write_png(p, [g[y].astype(np.uint8).tobytes() for y in range(h)], w, h, 8, 0, extra=_chunk(b"tRNS", struct.pack(">H", 60)))
    assert np.array_equal(pkg.load_png_rgba16(p), _expected_rgba16(g, 8, 0, key=[60]))
    c = np.random.RandomState(1).randint(0, 4, (h, w, 3)).astype(np.uint16) * 20000
    _write_png(p, [c[y].astype(">u2").tobytes() for y in range(h)], w, h, 16, 2, extra=_chunk(b"tRNS", struct.pack(">HHH", 20000, 40000, 0)))
    assert np.array_equal(pkg.load_png_rgba16(p), _expected_rgba16(c, 16, 2, key=[20000, 40000, 0]))
    pal = np.array([[255, 0, 0], [0, 128, 0], [10, 20, 30]], np.uint8); idx = (np.arange(w * h).reshape(h, w) % 3).astype(np.uint8)
    _write_png(p, [idx[y].tobytes() for y in range(h)], w, h, 8, 3, extra=_chunk(b"PLTE", pal.tobytes()) + _chunk(b"tRNS", bytes([255, 7])))
    exp = np.zeros((h, w, 4), np.uint16); exp[..., :3] = pal[idx].astype(np.uint16) * 257; exp[..., 3] = np.array([255, 7, 255], np.uint16)[idx] * 257
    assert np.array_equal(pkg.load_png_rgba16(p), exp)
    with pytest.raises(pkg.RnbError, match="image not found"):
        pkg.load_png_rgba16(tmp_path / "missing.png")
    open(tmp_path / "x.png", "wb").write(b"not a png at all")
    with pytest.raises(pkg.RnbError, match="not a PNG"):
        pkg.load_png_rgba16(tmp_path / "x.png")
    _write_png(p, [idx[y].tobytes() for y in range(h)], w, h, 8, 0, interlace=1)
    with pytest.raises(pkg.RnbError, match="interlaced"):
        pkg.load_png_rgba16(p)
    good = open(p, "rb").read()
    _write_png(p, [idx[y].tobytes() for y in range(h)], w, h, 8, 0)
    b = bytearray(open(p, "rb").read()); b[60] ^= 0x55
    open(p, "wb").write(bytes(b))
    with pytest.raises(pkg.RnbError):
        pkg.load_png_rgba16(p)
    del good


def test_transform_json_reader_matches_the_scene_it_was_written_from(tmp_path, scene_mod):
    views0 = scene_mod.make_scene(5, 48, 40, with_albedo=True)
    n2w = np.eye(4); n2w[:3, :3] *= 1.75; n2w[:3, 3] = (0.125, -2.5, 31.0)
    ref_scene.write_scene(str(tmp_path), views0, n2w=n2w)
    meta = ds.load_transforms(str(tmp_path))
    assert meta["from_na"] and meta["scale"] == 0.5 and meta["offset"] == (0.5, 0.5, 0.5) and meta["aabb_scale"] == 1
    assert meta["n2w_s"] == 1.75 and meta["n2w_t"] == (0.125, -2.5, 31.0) and len(meta["views"]) == 5
    for a, b in zip(views0, meta["views"]):
        assert np.array_equal(np.asarray(a["xform"], np.float32), b["xform"])                  # bit-exact camera matrices
        assert np.float32(a["fx"]) == b["fx"] and np.float32(a["fy"]) == b["fy"]
        assert abs(float(a["cx"]) - float(b["cx"])) < 1e-7 and abs(float(a["cy"]) - float(b["cy"])) < 1e-7
        assert (b["w"], b["h"]) == (a["w"], a["h"])
        assert np.array_equal(pkg.load_png_rgba16(b["normal_path"]), a["normal"])
        assert np.array_equal(pkg.load_png_rgba16(b["albedo_path"]), a["albedo"])


def test_axis_conventions_and_aabb(tmp_path):
    m = np.arange(12, dtype=np.float32).reshape(3, 4) + 1
    x = ds.nerf_matrix_to_ngp(m, 0.5, (0.5, 0.5, 0.5), from_na=False)
    t = m[:, 3] * np.float32(0.5) + np.float32(0.5)
    exp = np.stack([m[:, 0], -m[:, 1], -m[:, 2], t], 1)[[1, 2, 0]]
    assert np.array_equal(x, exp)
    x = ds.nerf_matrix_to_ngp(m, 0.5, (0.5, 0.5, 0.5), from_na=True)
    assert np.array_equal(x, np.stack([m[:, 0], m[:, 1], m[:, 2], t], 1))
    x = ds.nerf_matrix_to_ngp(m, 0.66, (0.165,) * 3, from_na=False, from_mitsuba=True)
    assert np.array_equal(x[:, :3], np.stack([-m[:, 0], -m[:, 1], m[:, 2]], 1))
    K = [[100.0, 0, 24], [0, 101.0, 20], [0, 0, 1]]
    j = {"w": 48, "h": 40, "aabb": [[-1, -2, -3], [3, 0, 1]], "frames": [{"normal_path": "n/0", "albedo_path": "", "intrinsic_matrix": K, "transform_matrix": np.eye(4).tolist()}]}
    json.dump(j, open(tmp_path / "transform.json", "w"))
    meta = ds.load_transforms(str(tmp_path / "transform.json"))
    assert meta["scale"] == 0.25 and meta["offset"] == (0.25, 0.75, 0.75) and not meta["from_na"]
    v = meta["views"][0]
    assert v["normal_path"].endswith(os.path.join("n", "0.png")) and v["albedo_path"] is None and v["cx"] == np.float32(0.5) and v["cy"] == np.float32(0.5)
    with pytest.raises(ValueError, match="No training images"):
        json.dump({"w": 1, "h": 1, "frames": []}, open(tmp_path / "transform.json", "w")); ds.load_transforms(str(tmp_path))


@pytest.mark.parametrize("seed", [1, 2])
def test_png_decoder_survives_corrupt_files(tmp_path, seed):
    """2 x 600 corrupt PNGs (tests/fuzz/png_fuzz.py) through the C ABI in a child process: it must exit normally — a hostile IHDR once
    drove a std::bad_alloc out of the library and aborted the process"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz", "png_fuzz.py"), str(seed), str(tmp_path), "600"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-600:])
    assert "rejected" in r.stdout and int(r.stdout.split("rejected")[1].split()[0]) > 400


def test_most_compressible_image_is_not_mistaken_for_a_hostile_header(tmp_path):
    """the decoder refuses headers that promise more pixels than deflate's 1032:1 bound allows for the compressed data; a constant
    full-size map at maximum compression (ratio ~1028:1) is the closest legitimate file"""
    import cv2
    pkg = rnb_loader.load_package()
    for val in (0, 65535):
        p = str(tmp_path / ("const%d.png" % val))
        cv2.imwrite(p, np.full((1200, 1600, 4), val, np.uint16), [cv2.IMWRITE_PNG_COMPRESSION, 9])
        assert (1600 * 8 + 1) * 1200 / os.path.getsize(p) > 800
        a = pkg.load_png_rgba16(p)
        assert a.shape == (1200, 1600, 4) and int(a.min()) == val and int(a.max()) == val


def test_transform_json_windows_paths_and_missing_normal(tmp_path):
    """replace(path, '\\\\', '/') converts EVERY backslash (ref:src/nerf_loader.cu:604,645); a frame without a normal map is a clear error, an empty
    albedo path means "no albedo map" (None), never an AttributeError further down"""
    fr = {"normal_path": "normals\\sub\\00000.png", "albedo_path": "", "transform_matrix": np.eye(4).tolist(), "intrinsic_matrix": [[10, 0, 2], [0, 10, 2], [0, 0, 1]]}
    tj = {"w": 4, "h": 4, "aabb_scale": 1, "scale": 0.5, "offset": [0.5] * 3, "from_na": True, "frames": [fr]}
    (tmp_path / "transform.json").write_text(json.dumps(tj))
    m = ds.load_transforms(str(tmp_path))
    assert m["views"][0]["normal_path"].endswith("normals/sub/00000.png") and "\\" not in m["views"][0]["normal_path"]
    assert m["views"][0]["albedo_path"] is None
    fr["normal_path"] = ""
    (tmp_path / "transform.json").write_text(json.dumps(tj))
    with pytest.raises(ValueError, match="normal_path"):
        ds.load_transforms(str(tmp_path))
