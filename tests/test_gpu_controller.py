"""The adaptive rays-per-batch controller on the GPU (pin_rays_per_batch = 0 — what shim/rnb_testbed_shim.h and therefore ./build/testbed
run): Counters::update_after_training (ref:src/testbed_nerf.cu:3532-3558), the max_inference clamp of train_nerf_step (:3891-3896), ray
scratch regrowth and the 2^18-ray cap (:3555).  Against the oracle in lock step and as exact properties of the product's own counters."""
import numpy as np
import pytest
from common import SMALL, make_pair
from oracle_binding import default_flags as orc_flags

pytestmark = pytest.mark.gpu


def next_multiple(v, d):
    return (v + d - 1) // d * d


def controller(R, target, compacted):
    """ref:src/testbed_nerf.cu:3554-3555 in binary32"""
    r = int(np.float32(np.float32(R) * np.float32(target)) / np.float32(compacted))
    return min(next_multiple(r, 128), 1 << 18)


def test_unpinned_controller_tracks_the_oracle(pkg, scene_mod):
    views = scene_mod.make_scene(6, 96, 96, with_albedo=True)
    target = 1 << 16
    f = orc_flags(no_albedo=0, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=256, pin_rays_per_batch=0, target_batch_size=target)
    o.set_train_state(training_step=0, rays_per_batch=256, pin_rays=0, target_batch=target)
    t.set_train_state(0, 256)
    t.set_rng(o.get_rng())
    resyncs = 0; seen_R = set(); prev_before = None; clamped = 0
    for it in range(24):
        a = o.train_step()
        b = t.train()
        seen_R.add(int(b.n_rays))
        if prev_before is not None and b.n_samples > next_multiple(prev_before, 128):
            clamped += 1                      # this step marched more samples than the clamp admits: rays were dropped, on both sides alike
        assert b.training_step == it + 1
        # the product's own arithmetic, exactly
        assert b.n_samples_compacted > 0
        assert b.rays_per_batch_next == controller(int(b.n_rays), target, int(b.n_samples_compacted)), (it, b.n_rays, b.n_samples_compacted, b.rays_per_batch_next)
        assert b.n_samples_trained == min(int(b.n_samples_compacted), target)
        # n_samples is numsteps_counter: every marched ray adds its samples to it BEFORE the capacity check (:1352-1357), so it also counts
        # the rays dropped by the max_inference clamp (:3891-3896) and may exceed last step's count; the oracle comparison below covers it
        prev_before = int(b.n_samples)
        # against the oracle: identical rays => identical sample counts unless an occupancy bit at the density threshold flipped
        assert abs(int(a.n_samples) - int(b.n_samples)) <= 0.01 * a.n_samples + 8, (it, a.n_samples, b.n_samples)
        assert abs(int(a.n_compacted) - int(b.n_samples_compacted)) <= 0.01 * a.n_compacted + 8
        assert abs(a.loss - b.loss) <= 0.03 * abs(a.loss) + 1e-6
        assert abs(int(a.rays_per_batch_next) - int(b.rays_per_batch_next)) <= 128, (it, a.rays_per_batch_next, b.rays_per_batch_next)
        if a.rays_per_batch_next != b.rays_per_batch_next:       # one 128-ray granule apart (count differed by a few samples): back to lock step
            st = t.get_train_state()
            t.set_train_state(st[0], int(a.rays_per_batch_next), st[2], st[3])
            resyncs += 1
    assert resyncs <= 4, resyncs
    assert len(seen_R) >= 2, seen_R           # the batch size moved (this small scene settles at the 128-ray granule quickly)
    print("steps whose sample count exceeded the max_inference clamp:", clamped)
    st = t.get_train_state()
    assert st[0] == 24 and st[3] == prev_before


def test_unpinned_controller_settles_at_the_sample_budget(pkg, scene_mod):
    """default sample budget (2^18): the controller's own arithmetic on every step, and the compacted count held near the target"""
    views = scene_mod.make_scene(8, 128, 128, with_albedo=False)
    f = orc_flags(no_albedo=1, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=4096, pin_rays_per_batch=0)
    del o
    t.set_train_state(0, 4096)
    target = 1 << 18; Rs = []
    for it in range(120):
        b = t.train()
        Rs.append(int(b.n_rays))
        assert b.n_samples_compacted > 0 and np.isfinite(b.loss)
        assert b.rays_per_batch_next == controller(int(b.n_rays), target, int(b.n_samples_compacted))
        assert b.n_samples_trained == min(int(b.n_samples_compacted), target)
    assert Rs[-1] % 128 == 0 and len(set(Rs)) > 3
    assert 0.5 * target < b.n_samples_compacted < 1.6 * target, b.n_samples_compacted


def test_controller_ray_cap_and_empty_steps(pkg, scene_mod):
    """a nearly empty occupancy grid: few samples => the batch grows to the 2^18-ray cap (:3555); a completely empty grid produces 0
    samples: both measured sizes are zeroed and rays_per_batch stays (Counters::update_after_training returns early, :3540-3542)"""
    views = scene_mod.make_scene(6, 96, 96, with_albedo=False)
    f = orc_flags(no_albedo=1, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=1024, pin_rays_per_batch=0)
    del o
    bf = np.zeros(128 ** 3, np.uint8)
    # a 16^3 block of cells around the scene centre in mip 0 (Morton order, common_device.h:338-363)
    def morton(x, y, z):
        def e(v):
            v = (v * 0x00010001) & 0xFF0000FF; v = (v * 0x00000101) & 0x0F00F00F; v = (v * 0x00000011) & 0xC30C30C3; v = (v * 0x00000005) & 0x49249249; return v
        return e(x) | (e(y) << 1) | (e(z) << 2)
    for x in range(56, 72):
        for y in range(56, 72):
            for z in range(56, 72):
                i = morton(x, y, z); bf[i >> 3] |= 1 << (i & 7)
    t.set_bitfield(bf)
    t.set_train_state(1, 1024)               # step > 0: rnb_train_step does not refresh the grid
    Rs = []
    for it in range(4):
        b = t.train_nerf()
        Rs.append(int(b.n_rays))
        assert b.n_samples_compacted > 0
        assert np.isfinite(b.loss)
    # 1024 -> (ray scratch regrown past its initial 4096 rays) -> the cap: at least one full step ran with 2^18 rays
    assert (1 << 18) in Rs and Rs[1] > 4096, (Rs, b.rays_per_batch_next)
    r_before = int(b.rays_per_batch_next)
    t.set_bitfield(np.zeros(128 ** 3, np.uint8))
    b = t.train_nerf()
    assert b.n_samples == 0 and b.n_samples_compacted == 0 and b.rays_per_batch_next == r_before
    st = t.get_train_state()
    assert st[3] == 0                         # measured_batch_size_before_compaction zeroed


def test_resume_like_load_snapshot_rebuilds_the_occupancy_grid(pkg, scene_mod):
    """Testbed::load_snapshot restores m_training_step but leaves m_canonical_training_step at 0 and (in a fresh process) n_images_for_training_prev
    at 0 (src/testbed.cu:2451,3333-3390; testbed.h:578,907): the first Testbed::train after it refreshes the grid at once (cadence src/testbed.cu:2805),
    in bootstrap mode (testbed_nerf.cu:4133), from an EMPTIED grid (:3446-3452), and restarts n_rays_total (:3906).  Library vs oracle, and vs a
    run that simply continues (which must NOT do any of that)."""
    views = scene_mod.make_scene(6, 96, 96, with_albedo=False)
    f = orc_flags(no_albedo=1, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=512)
    o.set_train_state(training_step=0, rays_per_batch=512, pin_rays=1); t.set_train_state(0, 512)
    t.set_rng(o.get_rng())
    for _ in range(6):
        a = o.train_step(); b = t.train()
    # a state in the middle of a run (step 300: refresh every 16 steps, 300 % 16 != 0): the step just trains
    st = t.get_train_state()
    o.set_train_state(training_step=300, rays_per_batch=512, n_rays_total=st[2], measured_before=st[3], pin_rays=1); t.set_train_state(300, 512, st[2], st[3])
    a = o.train_step(); b = t.train()
    assert b.training_step == 301 and b.density_grid_updated == 0 and t.get_train_state()[2] == st[2] + 512
    assert abs(int(a.n_samples) - int(b.n_samples)) <= 0.01 * a.n_samples + 8
    g_before, ema_before = t.export_density_grid()
    # "load_snapshot": same parameters / grid / step, canonical step and previous image count forgotten
    o.set_canonical_state(0, 0); t.set_canonical_state(0, 0)
    a = o.train_step(); b = t.train()
    assert b.density_grid_updated == 1 and b.training_step == 302
    assert abs(int(a.n_samples) - int(b.n_samples)) <= 0.01 * a.n_samples + 8 and abs(a.loss - b.loss) <= 0.03 * abs(a.loss) + 1e-6
    g_after, ema_after = t.export_density_grid()
    g_ref = o.get_density_grid()
    assert ema_after == ema_before + 1                                     # the EMA step is NOT reset (only at training step 0)
    assert np.linalg.norm(g_after - g_ref) <= 1e-3 * np.linalg.norm(g_ref)
    # from an emptied grid: no cell keeps more than the one new sample; the continued run would hold max(0.95 * old, new) >= 0.95 * old
    assert (g_after < 0.95 * g_before - 1e-6).any()
    assert t.get_train_state()[2] == 512                                   # n_rays_total restarted at this step
