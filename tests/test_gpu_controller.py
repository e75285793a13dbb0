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
    target = 1 << 14
    f = orc_flags(no_albedo=0, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=256, pin_rays_per_batch=0, target_batch_size=target)
    o.set_train_state(training_step=0, rays_per_batch=256, pin_rays=0, target_batch=target)
    t.set_train_state(0, 256)
    t.set_rng(o.get_rng())
    resyncs = 0; seen_R = set(); prev_before = None; clamped = 0
    for it in range(40):
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
    assert len(seen_R) > 3                    # the batch size really moved
    assert clamped > 0                        # and the max_inference clamp was exercised
    st = t.get_train_state()
    assert st[0] == 40 and st[3] == prev_before


def test_unpinned_controller_regrows_ray_scratch(pkg, scene_mod):
    """default sample budget (2^18): the batch grows past the initial ray capacity (4096) as the occupancy grid is carved"""
    views = scene_mod.make_scene(8, 128, 128, with_albedo=False)
    f = orc_flags(no_albedo=1, light_mode=-2)
    o, t = make_pair(pkg, SMALL, views=views, flags=f, rays_per_batch=4096, pin_rays_per_batch=0)
    del o
    t.set_train_state(0, 4096)
    target = 1 << 18; prev_before = None; Rs = []
    for it in range(120):
        b = t.train()
        Rs.append(int(b.n_rays))
        assert b.n_samples_compacted > 0 and np.isfinite(b.loss)
        assert b.rays_per_batch_next == controller(int(b.n_rays), target, int(b.n_samples_compacted))
        assert b.n_samples_trained == min(int(b.n_samples_compacted), target)
        prev_before = int(b.n_samples)
    assert max(Rs) > 4096, max(Rs)             # ensure_ray_capacity ran
    assert Rs[-1] % 128 == 0
    # the controller holds the compacted count near the target once it has settled
    assert 0.5 * target < b.n_samples_compacted < 1.6 * target, b.n_samples_compacted


def test_controller_ray_cap_and_empty_steps(pkg, scene_mod):
    """a nearly empty occupancy grid: almost no samples => the batch grows to the 2^18-ray cap (:3555); a completely empty grid produces 0
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
    assert b.rays_per_batch_next == 1 << 18 and Rs[-1] == 1 << 18, (Rs, b.rays_per_batch_next)
    b = t.train_nerf()                        # one full step at the cap
    assert b.n_rays == 1 << 18 and np.isfinite(b.loss)
    t.set_bitfield(np.zeros(128 ** 3, np.uint8))
    b = t.train_nerf()
    assert b.n_samples == 0 and b.n_samples_compacted == 0 and b.rays_per_batch_next == 1 << 18
    st = t.get_train_state()
    assert st[3] == 0                         # measured_batch_size_before_compaction zeroed
