"""Snapshot save / load on the CUDA path (Testbed::save_snapshot / load_snapshot, src/testbed.cu:3280-3390) through
the C ABI's export / import calls and the host-side codec.  Against the reference's own files: tests/ref_pin_snapshot.py."""
import json
import numpy as np
import pytest
from common import SMALL, product_config

pytestmark = pytest.mark.gpu


def test_snapshot_round_trip_and_training_continues(pkg, scene_mod, tmp_path):
    from rnb_neus2_b200 import snapshot as snap
    views = scene_mod.make_scene(6, 96, 96, with_albedo=True)
    mk = lambda: pkg.Testbed(product_config(pkg, SMALL, rays_per_batch=512, pin_rays_per_batch=1), pkg.default_flags(no_albedo=0, light_mode=-2))
    a = mk(); a.init_params(); a.load_training_data(views)
    for _ in range(24):
        st = a.train()
    net_cfg = {"encoding": {"otype": "HashGrid", "n_levels": 8}, "network": {"n_neurons": 32}, "optimizer": {"otype": "Ema", "decay": 0.95}}
    path = tmp_path / "snap.msgpack"
    a.save_snapshot(path, net_cfg)
    cfg = snap.read_snapshot(path)
    assert {k: cfg[k] for k in net_cfg} == net_cfg
    s = cfg["snapshot"]
    assert s["training_step"] == 24 and s["n_params"] == a.n_params and s["nerf"]["rgb"]["measured_batch_size"] == st.n_samples_compacted
    assert s["loss"] == float(np.float32(st.loss))
    ema = a.export_params_fp16(use_ema=True)
    assert s["params_binary"] == ema.tobytes()                                     # the INFERENCE parameters are what a snapshot holds
    b = mk(); b.load_training_data(views)
    b.load_snapshot(path)
    assert np.array_equal(b.export_params_fp16(use_ema=True).view(np.uint16), ema.view(np.uint16))
    assert np.array_equal(b.export_params_fp16(use_ema=False).view(np.uint16), ema.view(np.uint16))
    assert np.array_equal(b.get_params(), ema.astype(np.float32))                 # fp32 master re-derived from binary16 (trainer.h:263-275)
    ga, _ = a.export_density_grid(); gb, _ = b.export_density_grid()
    assert np.array_equal(gb, ga.astype(np.float16).astype(np.float32))
    # bitfield follows the loaded grid (update_density_grid_mean_and_bitfield, src/testbed.cu:3371): identical wherever binary16
    # rounding does not move a cell across the threshold
    ba, bb = np.unpackbits(a.get_bitfield()[:128 ** 3 // 8]), np.unpackbits(b.get_bitfield()[:128 ** 3 // 8])
    assert (ba != bb).mean() < 1e-3
    assert b.get_train_state()[:2] == [24, a.get_train_state()[1]]
    l0 = b.train().loss
    assert np.isfinite(l0) and abs(l0 - st.loss) < 0.5 * max(st.loss, 1e-3) + 1e-3      # picks up where the writer stopped
    with pytest.raises(pkg.RnbError):
        c = pkg.Testbed(product_config(pkg, dict(SMALL, log2_hashmap=13)))
        c.load_snapshot(path)                                                      # wrong parameter count


def test_snapshot_of_untrained_model_has_empty_grid(pkg, tmp_path):
    from rnb_neus2_b200 import snapshot as snap
    t = pkg.Testbed(product_config(pkg, SMALL)); t.init_params()
    path = tmp_path / "s0.msgpack"
    t.save_snapshot(path, {})
    d = snap.parse_snapshot(snap.read_snapshot(path))
    assert d["training_step"] == 0 and d["density_grid"].size == 0
    u = pkg.Testbed(product_config(pkg, SMALL)); u.load_snapshot(path)
    assert np.array_equal(u.export_params_fp16().view(np.uint16), t.export_params_fp16(use_ema=True).view(np.uint16))
