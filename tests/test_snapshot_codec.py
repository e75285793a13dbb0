"""Host-side snapshot format (rnb-neus2_b200/snapshot.py) on the CPU: the MessagePack subset nlohmann::json
emits, checked against the `msgpack` package in both directions, and the snapshot object's layout / error behaviour
(src/testbed.cu:3280-3390).  The file written by the reference build itself is checked on the B200 box (tests/ref_pin_snapshot.py,
summary under tests/golden/)."""
import numpy as np
import pytest
import rnb_loader

rnb_loader.load_package()
from rnb_neus2_b200 import snapshot as snap      # noqa: E402

msgpack = pytest.importorskip("msgpack")


def _sample():
    return {"encoding": {"otype": "HashGrid", "n_levels": 14, "per_level_scale": 1.3819, "base_resolution": 16, "log2_hashmap_size": 19},
            "optimizer": {"nested": {"learning_rate": 1e-3, "beta1": 0.9, "epsilon": 1e-15, "l2_reg": 1e-6}, "decay_start": 20000, "otype": "Ema"},
            "ints": [0, 1, 127, 128, 255, 256, 65535, 65536, 2 ** 32 - 1, 2 ** 32, 2 ** 63, -1, -32, -33, -128, -129, -32768, -32769, -2 ** 31, -2 ** 31 - 1],
            "floats": [0.0, 1.0, 0.5, 0.1, 1e-15, 3.5e38, 1e39, -2.25], "flags": [True, False, None], "empty": {}, "s": "x" * 40, "long": "y" * 300,
            "list17": list(range(17)), "map17": {"k%02d" % i: i for i in range(17)}, "blob": bytes(range(200)), "blob2": bytes(70000), "blob0": b""}


def test_codec_round_trip_and_cross_check_with_msgpack():
    o = _sample()
    b = snap.packb(o)
    assert snap.unpackb(b) == o
    assert msgpack.unpackb(b, raw=False, strict_map_key=False) == o                      # the package reads what we write
    b2 = msgpack.packb(o, use_bin_type=True)
    assert snap.unpackb(b2) == o                                                          # and we read what it writes
    # same bytes as the package for everything but floats (it always writes float64, nlohmann shortens exact binary32 values)
    o2 = {k: v for k, v in o.items() if k not in ("floats", "encoding", "optimizer")}
    assert snap.packb(o2) == msgpack.packb({k: o2[k] for k in sorted(o2)}, use_bin_type=True)
    assert snap.packb(0.5) == b"\xca" + np.array(0.5, ">f4").tobytes()
    assert snap.packb(0.1)[0] == 0xCB and snap.packb(1e39)[0] == 0xCB
    with pytest.raises(ValueError):
        snap.unpackb(b[:-3])
    with pytest.raises(ValueError):
        snap.unpackb(b + b"\x00")


def test_snapshot_object_layout_and_errors(tmp_path):
    rs = np.random.RandomState(0)
    p = rs.normal(0, 0.1, 1000).astype(np.float16); g = rs.uniform(-1, 5, snap.GRID_SIZE ** 3).astype(np.float32)
    cfg = snap.build_snapshot({"network": {"n_neurons": 64}, "snapshot": {"stale": 1}}, p, g, 1234, 0.0123, 4096, 200000, 850000)
    s = cfg["snapshot"]
    assert "stale" not in s and cfg["network"] == {"n_neurons": 64}
    assert s["n_params"] == 1000 and s["density_grid_size"] == 128 and len(s["density_grid_binary"]) == 2 * 128 ** 3 and len(s["params_binary"]) == 2000
    assert len(s["rotation"]) == 24 and len(s["transition"]) == 8
    path = tmp_path / "s.msgpack"
    snap.write_snapshot(path, cfg)
    d = snap.parse_snapshot(snap.read_snapshot(path))
    assert np.array_equal(d["params_fp16"], p) and np.array_equal(d["density_grid"], g.astype(np.float16).astype(np.float32))
    assert (d["training_step"], d["rays_per_batch"], d["measured_batch_size"], d["measured_batch_size_before_compaction"], d["aabb_scale"]) == (1234, 4096, 200000, 850000, 1)
    assert d["loss"] == float(np.float32(0.0123))
    with pytest.raises(ValueError, match="does not contain a snapshot"):
        snap.parse_snapshot({"network": {}})
    bad = snap.build_snapshot({}, p, g, 0, 0.0, 1, 1, 1); bad["snapshot"]["density_grid_size"] = 64
    with pytest.raises(ValueError, match="Incompatible grid size"):
        snap.parse_snapshot(bad)
    bad = snap.build_snapshot({}, p, g[:1000], 0, 0.0, 1, 1, 1)
    with pytest.raises(ValueError, match="cascades"):
        snap.parse_snapshot(bad)
    empty = snap.build_snapshot({}, p, np.zeros(0, np.float32), 0, 0.0, 1, 1, 1)            # untrained model: empty grid is valid
    assert snap.parse_snapshot(empty)["density_grid"].size == 0


def test_reference_snapshot_fixture():
    """tests/golden/ref_snapshot_small.npz: the file Testbed::save_snapshot of the reference build wrote on the B200 box
    (tests/ref_pin_snapshot.py), its two large blobs cut to 4 KiB.  Our codec must decode it and re-encode it byte for byte, the
    object must parse, and the movement blobs must be the static-scene defaults this repo writes."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_snapshot_small.npz"))
    raw = g["truncated"].tobytes()
    cfg = snap.unpackb(raw)
    assert snap.packb(cfg) == raw
    assert msgpack.unpackb(raw, raw=False, strict_map_key=False) == cfg
    s = cfg["snapshot"]
    assert sorted(s) == ["density_grid_binary", "density_grid_size", "local_rotation", "local_transition", "loss", "n_params", "nerf", "params_binary", "rotation",
                         "training_step", "transition"]
    assert sorted(s["nerf"]) == ["aabb_scale", "dataset", "rgb"] and sorted(s["nerf"]["rgb"]) == ["measured_batch_size", "measured_batch_size_before_compaction", "rays_per_batch"]
    assert s["density_grid_size"] == 128 and s["n_params"] == 241156 and s["training_step"] == 40
    assert np.array_equal(np.frombuffer(s["params_binary"], np.uint16), g["params_head"])
    assert np.array_equal(np.frombuffer(s["density_grid_binary"], np.float16), g["grid_head"].astype(np.float16))
    for k, v in snap.movement_defaults().items():
        assert s[k] == v, k
    for k in ("encoding", "network", "rgb_network", "optimizer", "loss", "dir_encoding"):
        assert k in cfg
