// ORACLE — TEST INFRASTRUCTURE ONLY.  See orc_common.h for scope and parity status (pinned against the reference build run on B200).
//
// C entry points (ctypes) over the CPU restatement.  Stage functions mirror the reference call graph of
// Testbed::train -> training_prep_nerf / train_nerf -> train_nerf_step (src/testbed.cu:2776-2872,
// src/testbed_nerf.cu:3424-3558,3560-3668,3844-4138) and the optimizer chain Ema -> ExponentialDecay -> Adam
// (tcnn optimizers/ema.h:116-152, exponential_decay.h:61-72, adam.h:51-202).
#include "orc_common.h"
#include "orc_network.h"
#include "orc_render.h"
#include "orc_mesh.h"
#include <random>
#include <limits>
#include <cstdio>

using namespace orc;

struct OrcStats {
	float loss, ek_loss, mask_loss;
	uint32_t n_rays_kept, n_samples, n_compacted, n_emitted, rays_per_batch_next;
};

struct Oracle {
	ModelDesc m;
	int threads = 1;
	std::vector<float> master;        // fp32 master weights
	std::vector<float> pv;            // binary16 training weights, widened
	std::vector<float> ema;           // binary16 EMA (inference) weights, widened
	std::vector<float> m1, m2;
	std::vector<uint32_t> psteps;
	std::vector<float> grads;         // fp32 accumulators, × loss scale
	// optimizer
	float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.99f, eps = 1e-15f, l2 = 1e-6f, ema_decay = 0.95f;
	uint32_t decay_start = 20000, decay_interval = 10000; float decay_base = 0.33f;
	uint32_t opt_step = 0; float lr_factor = 1.0f; int only_sdf = 0;
	float loss_scale = 128.f;
	// occupancy
	std::vector<float> density_grid; std::vector<uint8_t> bitfield; uint32_t density_ema_step = 0; float density_mean = 0.f; float density_decay = 0.95f;
	Pcg32 rng, density_rng;
	// training state
	uint32_t training_step = 0, rays_per_batch = 4096, n_rays_total = 0, target_batch = 1u << 18;
	uint32_t measured_before = 0, measured = 0; int pin_rays = 1;
	// Testbed::m_canonical_training_step (testbed.h:907; 0 after load_snapshot -> reset_network, src/testbed.cu:2451; = training step after every step,
	// testbed_nerf.cu:3646) and Training::n_images_for_training_prev (testbed.h:578; ~0u = adopt the current image count)
	uint32_t canonical_step = 0, n_images_prev = ~0u;
	Flags flags;
	std::vector<View> views;
	// scratch kept for inspection
	std::vector<uint32_t> ray_indices, numsteps; std::vector<float> rays, coords;
	std::vector<float> last_loss, last_ek, last_mask;   // per kept ray, in ray_indices order
	// data-parallel restatement: rank r of `world` marches rays i == r (mod world); sums[] are all-reduced with the gradients
	uint32_t world = 1, rank = 0; bool in_step = false; uint32_t step_R = 0;
	// dp_exact: the shards share ONE sample order — every clamp, truncation and roll-over multiplicity is taken at the sample's index in the
	// GLOBAL batch (the order a single process would have produced), so that the sum of the shard gradients IS the single-process gradient.
	// The restatement gets there by marching / compacting the whole batch on every rank and keeping loss + backward of its own rays only
	// (the product exchanges two per-ray prefix tables instead).  false: every rank clamps / truncates / pads its own shard (target / world).
	bool dp_exact = false;
	size_t shard_begin = 0, shard_end = 0;      // data-parallel optimizer shard (end == 0: all parameters)
	double sums[4] = {0, 0, 0, 0};   // loss, ek, mask, compacted samples
	uint32_t cnt_kept = 0, cnt_samples = 0, cnt_total = 0, cnt_trained = 0;
};

static void sync_half(Oracle* o) { for (size_t i = 0; i < o->m.n_params; ++i) o->pv[i] = hq(o->master[i]); }

extern "C" {

Oracle* orc_create(uint32_t n_levels, uint32_t log2_hashmap, uint32_t base_res, float per_level_scale, uint32_t sdf_width, uint32_t sdf_hidden, uint32_t rgb_width, uint32_t rgb_hidden, float sdf_bias, int threads) {
	Oracle* o = new Oracle();
	o->m.n_levels = n_levels; o->m.log2_hashmap = log2_hashmap; o->m.base_res = base_res; o->m.per_level_scale = per_level_scale;
	o->m.sdf_width = sdf_width; o->m.sdf_hidden = sdf_hidden; o->m.rgb_width = rgb_width; o->m.rgb_hidden = rgb_hidden; o->m.sdf_bias = sdf_bias;
	o->m.finalize();
	o->threads = threads < 1 ? 1 : threads;
	size_t n = o->m.n_params;
	o->master.assign(n, 0.f); o->pv.assign(n, 0.f); o->ema.assign(n, 0.f); o->m1.assign(n, 0.f); o->m2.assign(n, 0.f); o->psteps.assign(n, 0); o->grads.assign(n, 0.f);
	o->density_grid.assign(GRIDSIZE * GRIDSIZE * GRIDSIZE, 0.f);
	o->bitfield.assign((size_t)GRIDSIZE * GRIDSIZE * GRIDSIZE * CASCADES / 8, 0);
	o->rng = Pcg32(1337);
	Pcg32 tmp = o->rng;
	o->density_rng = Pcg32(tmp.next_uint());     // src/testbed.cu:2223,2236 (copy: m_rng itself is re-seeded at :2490)
	return o;
}
void orc_destroy(Oracle* o) { delete o; }
float orc_per_level_scale(float top_res, float aabb_scale, uint32_t base_res, uint32_t n_levels) {
	return std::exp(std::log(top_res * aabb_scale / (float)base_res) / (n_levels - 1));      // src/testbed.cu:2321
}
uint64_t orc_n_params(Oracle* o) { return o->m.n_params; }
void orc_layout(Oracle* o, uint64_t* out /*off_sdf, off_rgb, off_grid, off_var, n_params, sdf_in, rgb_in*/) {
	out[0] = o->m.off_sdf; out[1] = o->m.off_rgb; out[2] = o->m.off_grid; out[3] = o->m.off_var; out[4] = o->m.n_params; out[5] = o->m.sdf_in; out[6] = o->m.rgb_in;
}
void orc_grid_meta(Oracle* o, uint32_t* offsets /*L+1*/, uint32_t* res /*L*/, float* scale /*L*/) {
	for (uint32_t i = 0; i <= o->m.n_levels; ++i) offsets[i] = o->m.offsets[i];
	for (uint32_t i = 0; i < o->m.n_levels; ++i) { res[i] = o->m.res[i]; scale[i] = o->m.scale[i]; }
}
uint32_t orc_grid_index(uint32_t hashmap_size, uint32_t res, uint32_t x, uint32_t y, uint32_t z) { uint32_t p[3] = {x, y, z}; return grid_index(hashmap_size, res, p); }
uint32_t orc_valid_level(Oracle* o, int step) { return valid_level_for_step(o->m, step); }
uint32_t orc_morton3D(uint32_t x, uint32_t y, uint32_t z) { return morton3D(x, y, z); }
uint32_t orc_morton3D_invert(uint32_t x) { return morton3D_invert(x); }
void orc_pcg32(uint64_t seed, int64_t advance, uint32_t n, uint32_t* out_u, float* out_f) {
	Pcg32 r(seed); r.advance(advance);
	Pcg32 r2 = r;
	for (uint32_t i = 0; i < n; ++i) { if (out_u) out_u[i] = r.next_uint(); if (out_f) out_f[i] = r2.next_float(); }
}
void orc_set_threads(Oracle* o, int t) { o->threads = t < 1 ? 1 : t; }

// Parameter initialisation in the reference order: Trainer rng = pcg32{seed_seq{seed}.front()} (trainer.h:54-60);
// SDF MLP xavier (then overwritten by the geometric-init file), colour MLP xavier (gpu_matrix.h:292-304), hash grid
// U(-1e-4,1e-4) through generate_random_kernel's strided layout (random.h:67-93, grid.h:1379-1384), variance 0.3.
void orc_init_params(Oracle* o, uint32_t seed, const float* sdf_init, uint64_t n_sdf_init) {
	std::seed_seq seq{seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	Pcg32 rnd(seeds.front());
	auto xavier = [&](const std::vector<Layer>& ls) {
		for (const Layer& L : ls) {
			float scale = std::sqrt(6.0f / (float)(L.cols + L.rows));
			for (size_t i = 0; i < (size_t)L.rows * L.cols; ++i) o->master[L.off + i] = rnd.next_float() * 2.0f * scale - scale;
		}
	};
	xavier(o->m.sdf_layers);
	if (sdf_init) {
		size_t n_sdf = o->m.off_rgb - o->m.off_sdf;
		for (size_t i = 0; i < n_sdf; ++i) o->master[o->m.off_sdf + i] = i < n_sdf_init ? sdf_init[i] : 0.f;   // short file -> zeros (nerf_network.h:587,616-619)
	}
	xavier(o->m.rgb_layers);
	{
		const size_t n = o->m.n_grid_params;
		const size_t n_thr_needed = (n + 3) / 4;
		const size_t n_threads = ((n_thr_needed + 127) / 128) * 128;
		float* out = o->master.data() + o->m.off_grid;
		for (size_t i = 0; i < n_threads; ++i) {
			Pcg32 r = rnd; r.advance((int64_t)(i * 4));
			for (size_t j = 0; j < 4; ++j) {
				size_t idx = i + n_threads * j;
				if (idx >= n) break;
				out[idx] = r.next_float() * (1e-4f - (-1e-4f)) + (-1e-4f);
			}
		}
		rnd.advance((int64_t)n);
	}
	for (int i = 0; i < 4; ++i) o->master[o->m.off_var + i] = 0.3f;
	sync_half(o);
	std::fill(o->ema.begin(), o->ema.end(), 0.f);
	std::fill(o->m1.begin(), o->m1.end(), 0.f); std::fill(o->m2.begin(), o->m2.end(), 0.f); std::fill(o->psteps.begin(), o->psteps.end(), 0u);
	o->opt_step = 0; o->lr_factor = 1.f;
}
void orc_set_params(Oracle* o, const float* p) { std::memcpy(o->master.data(), p, o->m.n_params * 4); sync_half(o); }
void orc_get_params(Oracle* o, float* master, float* half_widened, float* ema) {
	if (master) std::memcpy(master, o->master.data(), o->m.n_params * 4);
	if (half_widened) std::memcpy(half_widened, o->pv.data(), o->m.n_params * 4);
	if (ema) std::memcpy(ema, o->ema.data(), o->m.n_params * 4);
}
void orc_get_grads(Oracle* o, float* g) { std::memcpy(g, o->grads.data(), o->m.n_params * 4); }
void orc_set_grads(Oracle* o, const float* g) { std::memcpy(o->grads.data(), g, o->m.n_params * 4); }
void orc_get_opt_state(Oracle* o, float* m1, float* m2, uint32_t* steps) {
	std::memcpy(m1, o->m1.data(), o->m.n_params * 4); std::memcpy(m2, o->m2.data(), o->m.n_params * 4); std::memcpy(steps, o->psteps.data(), o->m.n_params * 4);
}

void orc_set_views(Oracle* o, const View* v, uint32_t n) { o->views.assign(v, v + n); }
void orc_set_flags(Oracle* o, const Flags* f) { o->flags = *f; }
void orc_set_train_state(Oracle* o, uint32_t training_step, uint32_t rays_per_batch, uint32_t n_rays_total, uint32_t measured_before, int pin_rays, uint32_t target_batch) {
	o->training_step = training_step; o->rays_per_batch = rays_per_batch; o->n_rays_total = n_rays_total; o->measured_before = measured_before; o->pin_rays = pin_rays; o->target_batch = target_batch;
	o->canonical_step = training_step;
}
// the two members Testbed::load_snapshot does not restore (src/testbed.cu:3333-3390)
void orc_set_canonical_state(Oracle* o, uint32_t canonical_step, uint32_t n_images_prev) { o->canonical_step = canonical_step; o->n_images_prev = n_images_prev; }
void orc_set_world(Oracle* o, uint32_t world, uint32_t rank) { o->world = world ? world : 1; o->rank = rank; }
void orc_set_dp_exact(Oracle* o, int on) { o->dp_exact = on != 0; }
// data-parallel restatement of the sharded optimizer: Adam/EMA on [begin, end) only; the binary16 training weights of the other
// shards are installed by the caller from the all-gather (orc_set_half_params)
void orc_set_opt_shard(Oracle* o, uint64_t begin, uint64_t end) { o->shard_begin = (size_t)begin; o->shard_end = (size_t)end; }
void orc_set_half_params(Oracle* o, const float* half_widened) { std::memcpy(o->pv.data(), half_widened, o->m.n_params * 4); }
void orc_set_rng(Oracle* o, uint64_t state, uint64_t inc, uint64_t dstate, uint64_t dinc) { o->rng.state = state; o->rng.inc = inc; o->density_rng.state = dstate; o->density_rng.inc = dinc; }
void orc_get_rng(Oracle* o, uint64_t* out) { out[0] = o->rng.state; out[1] = o->rng.inc; out[2] = o->density_rng.state; out[3] = o->density_rng.inc; }
void orc_get_bitfield(Oracle* o, uint8_t* out) { std::memcpy(out, o->bitfield.data(), o->bitfield.size()); }
void orc_set_bitfield(Oracle* o, const uint8_t* in) { std::memcpy(o->bitfield.data(), in, o->bitfield.size()); }
void orc_get_density_grid(Oracle* o, float* out) { std::memcpy(out, o->density_grid.data(), o->density_grid.size() * 4); }
void orc_set_density_grid(Oracle* o, const float* in, uint32_t ema_step) { std::memcpy(o->density_grid.data(), in, o->density_grid.size() * 4); o->density_ema_step = ema_step; }

// ----------------------------------------------------------------------------------------------------------
// Stage: ray generation + occupancy marching (generate_training_samples_nerf, testbed_nerf.cu:1216-1387).
// Sample slots are handed out in ray order (the reference hands them out in atomicAdd arrival order).
// counters: [0] rays kept, [1] samples counted (includes rays dropped by the max_samples guard, like the reference)
// ----------------------------------------------------------------------------------------------------------
void orc_generate_samples(Oracle* o, uint32_t n_rays, uint32_t n_rays_total, uint32_t max_samples,
                          uint32_t* ray_indices, float* rays /*6 per ray: o, d_unnormalised*/, uint32_t* numsteps /*2 per ray*/, float* coords /*7 per sample*/, uint32_t* counters) {
	std::vector<RayGen> rg(n_rays);
	parallel_for(o->threads, n_rays, [&](int, size_t b, size_t e) {
		for (size_t i = b; i < e; ++i) {
			if (!o->dp_exact && i % o->world != o->rank) { rg[i].valid = false; continue; }      // ray shard of this rank
			ray_setup((uint32_t)i, n_rays, n_rays_total, o->rng, o->views.data(), (uint32_t)o->views.size(), o->bitfield.data(), rg[i]);
		}
	});
	uint32_t n_kept = 0, counter = 0;
	std::vector<uint32_t> slot(n_rays, 0xFFFFFFFFu);
	for (uint32_t i = 0; i < n_rays; ++i) {
		if (!rg[i].valid) continue;
		uint32_t base = counter; counter += rg[i].numsteps;
		if (base + rg[i].numsteps > max_samples) continue;
		uint32_t k = n_kept++;
		slot[i] = k;
		ray_indices[k] = i;
		for (int c = 0; c < 3; ++c) { rays[k * 6 + c] = rg[i].o[c]; rays[k * 6 + 3 + c] = rg[i].d_un[c]; }
		numsteps[k * 2] = rg[i].numsteps; numsteps[k * 2 + 1] = base;
	}
	parallel_for(o->threads, n_rays, [&](int, size_t b, size_t e) {
		for (size_t i = b; i < e; ++i) if (slot[i] != 0xFFFFFFFFu) ray_emit(rg[i], o->bitfield.data(), coords + (size_t)numsteps[slot[i] * 2 + 1] * 7);
	});
	counters[0] = n_kept; counters[1] = counter;
}

// Stage: network forward on arbitrary coords (NerfNetwork::forward_impl).  out: 16 floats per sample (binary16 values).
void orc_network_forward(Oracle* o, const float* coords, uint64_t n, uint32_t valid_level, int use_ema, int with_rgb, float* out, float* normal_f32 /*optional 3 per sample*/) {
	Net<true> net(o->m, use_ema ? o->ema.data() : o->pv.data(), valid_level);
	parallel_for(o->threads, n, [&](int, size_t b, size_t e) {
		Net<true>::Ctx c;
		for (size_t i = b; i < e; ++i) {
			net.forward(coords + i * 7, c, with_rgb != 0);
			for (int k = 0; k < 16; ++k) out[i * 16 + k] = c.out[k];
			if (normal_f32) for (int d = 0; d < 3; ++d) normal_f32[i * 3 + d] = c.normal[d];
		}
	});
}
// hash encoding alone: enc (2L binary16 values) and dy/dx (2L x 3 fp32) per sample
void orc_encode(Oracle* o, const float* xyz, uint64_t n, uint32_t valid_level, float* enc, float* dydx) {
	Net<true> net(o->m, o->pv.data(), valid_level);
	for (size_t i = 0; i < n; ++i) {
		Net<true>::Ctx c;
		float p[3] = {xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]};
		for (uint32_t l = 0; l < o->m.n_levels; ++l) {
			float e[2]; net.encode_level(l, p, c, e);
			enc[i * o->m.n_enc + 2 * l] = e[0]; enc[i * o->m.n_enc + 2 * l + 1] = e[1];
		}
		if (dydx) for (uint32_t k = 0; k < o->m.n_enc; ++k) for (int d = 0; d < 3; ++d) dydx[(i * o->m.n_enc + k) * 3 + d] = c.dydx[k][d];
	}
}
// NerfNetwork::sdf / density (nerf_network.h:454-537; sdf_to_density_variance_buffer common_operation.cuh:310-328)
void orc_eval_sdf(Oracle* o, const float* xyz, uint64_t n, uint32_t valid_level, int use_ema, float* sdf_out, float* density_out) {
	const float* P = use_ema ? o->ema.data() : o->pv.data();
	Net<true> net(o->m, P, valid_level);
	parallel_for(o->threads, n, [&](int, size_t b, size_t e) {
		for (size_t i = b; i < e; ++i) {
			float s = net.sdf_only(xyz + i * 3);
			if (sdf_out) sdf_out[i] = s;
			if (density_out) {
				float var = P[o->m.off_var];
				float sc = hq(std::exp(hmul(var, 10.0f)));
				float sg = hq(logistic(hmul(s, sc)));
				density_out[i] = hmul(hmul(sc, sg), hsub(1.0f, sg));
			}
		}
	});
}

// Stage: per-ray transmittance compaction on pass-A outputs.  numsteps[2k], [2k+1] = (count, base) from generation.
// Writes n_fwd[k] (samples surviving T >= 1e-4), cbase[k] (prefix over rays, untruncated), n_emit[k] (truncated to max_compacted).
// Returns the untruncated compacted total (what the reference's compacted counter holds).
uint32_t orc_compact(Oracle* o, const float* out_a, const uint32_t* numsteps, uint32_t n_kept, uint32_t max_compacted, uint32_t* n_fwd, uint32_t* cbase, uint32_t* n_emit) {
	const float dt = MIN_STEP();
	parallel_for(o->threads, n_kept, [&](int, size_t b, size_t e) {
		for (size_t k = b; k < e; ++k) {
			const float* oa = out_a + (size_t)numsteps[k * 2 + 1] * 16;
			n_fwd[k] = ray_compacted_count(oa, numsteps[k * 2], oa, dt, o->flags.cos_anneal_ratio);
		}
	});
	uint32_t total = 0;
	for (uint32_t k = 0; k < n_kept; ++k) {
		cbase[k] = total;
		n_emit[k] = std::min(max_compacted - std::min(max_compacted, total), n_fwd[k]);
		total += n_fwd[k];
	}
	return total;
}

// Stage: loss + dL/d(out) per ray on the ray's compacted outputs (out_c laid out at cbase[k]).
void orc_loss(Oracle* o, const float* out_c, const uint32_t* ray_indices, const uint32_t* n_fwd, const uint32_t* cbase, const uint32_t* n_emit, uint32_t n_kept,
              uint32_t n_rays, uint32_t n_rays_total, uint32_t step, float* dout /*16 per compacted sample*/, float* loss, float* ek_loss, float* mask_loss /*per kept ray*/) {
	const float dt = MIN_STEP();
	parallel_for(o->threads, n_kept, [&](int, size_t b, size_t e) {
		for (size_t k = b; k < e; ++k) {
			loss[k] = ek_loss[k] = mask_loss[k] = 0.f;
			if (n_emit[k] == 0) continue;
			if (o->dp_exact && ray_indices[k] % o->world != o->rank) continue;      // another rank's ray: its dL/d(out) stays zero here
			RayTarget T; ray_target(ray_indices[k], n_rays, n_rays_total, o->rng, o->views.data(), (uint32_t)o->views.size(), o->flags, step, T);
			RayLoss RL;
			ray_loss(out_c + (size_t)cbase[k] * 16, n_fwd[k], n_emit[k], T, o->flags, dt, n_rays, o->loss_scale, dout + (size_t)cbase[k] * 16, RL);
			loss[k] = RL.loss; ek_loss[k] = RL.ek_loss; mask_loss[k] = RL.mask_loss;
		}
	});
}

// roll-over multiplicity weight (fill_rollover_and_rescale, common_device.h:525-535)
float orc_rollover_weight(uint32_t s, uint32_t n_in, uint32_t n_batch) {
	if (n_in == 0 || n_in >= n_batch) return 1.0f;
	uint32_t c = (n_batch - 1 - s) / n_in;
	return 1.0f + (float)c * ((float)n_in / (float)n_batch);
}

// Stage: network forward + backward (first and second order) on n compacted samples; accumulates into o->grads (overwrites).
static void network_backward_impl(Oracle* o, const float* coords, const float* dout, uint64_t n, uint32_t n_in_for_rollover, uint32_t n_roll, uint32_t n_batch, uint32_t valid_level, const uint8_t* foreign = nullptr);
void orc_network_backward(Oracle* o, const float* coords, const float* dout, uint64_t n, uint32_t n_in_for_rollover, uint32_t n_batch, uint32_t valid_level) {
	network_backward_impl(o, coords, dout, n, n_in_for_rollover, n_batch, n_batch, valid_level);
}
// n_roll: size the compacted batch is padded to by roll-over (per rank: target / world); n_batch: global Eikonal divisor
// foreign (dp_exact): samples of other ranks' rays — they hold their place in the global order (roll-over index) and are skipped
static void network_backward_impl(Oracle* o, const float* coords, const float* dout, uint64_t n, uint32_t n_in_for_rollover, uint32_t n_roll, uint32_t n_batch, uint32_t valid_level, const uint8_t* foreign) {
	const size_t n_mlp = o->m.off_grid;
	std::fill(o->grads.begin(), o->grads.end(), 0.f);
	const int T = o->threads;
	std::vector<std::vector<float>> mlp_g(T > 1 ? T : 0, std::vector<float>(n_mlp, 0.f));
	Net<true> net(o->m, o->pv.data(), valid_level);
	float* G = o->grads.data();
	parallel_for(T, n, [&](int t, size_t b, size_t e) {
		Net<true>::Ctx c;
		float* Gm = T > 1 ? mlp_g[t].data() : G;     // MLP gradients: thread-private, folded below; hash gradients: shared + atomic
		float gvar = 0.f;
		for (size_t i = b; i < e; ++i) {
			if (foreign && foreign[i]) continue;
			net.forward(coords + i * 7, c, true);
			float w = orc_rollover_weight((uint32_t)i, n_in_for_rollover, n_roll);
			net.backward(c, dout + i * 16, w, n_batch, Gm, G, &gvar, T > 1);
		}
		atomic_add_f32(&G[o->m.off_var], gvar);
	});
	for (int t = 0; t < (int)mlp_g.size(); ++t) for (size_t i = 0; i < n_mlp; ++i) G[i] += mlp_g[t][i];
}

// Optimizer: Ema(ExponentialDecay(Adam)) on o->grads.  adam.h:51-202, exponential_decay.h:61-72, ema.h:64-78,116-152
void orc_optimizer_step(Oracle* o) {
	if (o->opt_step == 0) o->lr_factor = 1.0f;
	if (o->opt_step >= o->decay_start && (o->opt_step - o->decay_start) % o->decay_interval == 0) o->lr_factor *= o->decay_base;
	const float base_lr = o->lr * o->lr_factor;
	++o->opt_step;
	const size_t n = o->m.n_params, n_mat = o->m.off_grid;
	const size_t rgb_b = o->m.off_rgb, rgb_e = o->m.off_grid;
	const size_t sb = o->shard_end ? o->shard_begin : 0, se = o->shard_end ? std::min(o->shard_end, n) : n;
	parallel_for(o->threads, n, [&](int, size_t b, size_t e) {
		for (size_t i = b; i < e; ++i) {
			if (i < sb || i >= se) continue;
			float gradient = hq(o->grads[i]) / o->loss_scale;
			const bool is_mat = i < n_mat;
			if (!is_mat && gradient == 0) continue;
			if (o->only_sdf && i >= rgb_b && i < rgb_e) continue;
			const float w = o->master[i];
			if (is_mat) gradient += o->l2 * w;
			const float gsq = gradient * gradient;
			float fm = o->m1[i] = o->beta1 * o->m1[i] + (1 - o->beta1) * gradient;
			const float sm = o->m2[i] = o->beta2 * o->m2[i] + (1 - o->beta2) * gsq;
			float lr = base_lr;
			const uint32_t cs = ++o->psteps[i];
			lr *= std::sqrt(1 - std::pow(o->beta2, (float)cs)) / (1 - std::pow(o->beta1, (float)cs));
			const float eff = std::fmin(std::fmax(lr / (std::sqrt(sm) + o->eps), 0.f), std::numeric_limits<float>::max());
			const float nw = w - eff * fm;
			o->master[i] = nw;
			o->pv[i] = hq(nw);
		}
	});
	const float d_old = 1 - (float)std::pow(o->ema_decay, o->opt_step - 1);
	const float d_new = 1.0f / (1 - (float)std::pow(o->ema_decay, o->opt_step));
	parallel_for(o->threads, n, [&](int, size_t b, size_t e) {
		for (size_t i = b; i < e; ++i) if (i >= sb && i < se) o->ema[i] = hq((o->ema[i] * o->ema_decay * d_old + o->pv[i] * (1 - o->ema_decay)) * d_new);
	});
}

// Occupancy refresh: update_density_grid_nerf + update_density_grid_mean_and_bitfield (testbed_nerf.cu:3424-3517) and
// kernels :585-614,616-635,655-685,693-740.  With aabb_scale 1 there is one cascade of densities; mips are OR-pooled.
void orc_density_update(Oracle* o, uint32_t n_uniform, uint32_t n_nonuniform, uint32_t valid_level) {
	const uint32_t NE = GRIDSIZE * GRIDSIZE * GRIDSIZE;
	if (o->n_images_prev == ~0u) o->n_images_prev = (uint32_t)o->views.size();
	if (o->training_step == 0 || (uint32_t)o->views.size() != o->n_images_prev) {      // testbed_nerf.cu:3446-3452
		o->n_images_prev = (uint32_t)o->views.size();
		if (o->training_step == 0) o->density_ema_step = 0;
		std::fill(o->density_grid.begin(), o->density_grid.end(), 0.f);
	}
	std::vector<float> tmp(NE, 0.f);
	const uint32_t n_total = n_uniform + n_nonuniform;
	std::vector<float> pos((size_t)n_total * 3); std::vector<uint32_t> idxs(n_total);
	auto gen = [&](uint32_t n_el, uint32_t off, float thresh) {
		Pcg32 base = o->density_rng;
		parallel_for(o->threads, n_el, [&](int, size_t b, size_t e) {
			for (size_t i = b; i < e; ++i) {
				Pcg32 r = base; r.advance((int64_t)i * 4);
				uint32_t level = (uint32_t)(r.next_float() * 1) % 1;
				uint32_t idx = 0;
				for (uint32_t j = 0; j < 10; ++j) {
					idx = (((uint32_t)i + o->density_ema_step * n_el) * 56924617u + j * 19349663u + 96925573u) % NE;
					idx += level * NE;
					if (o->density_grid[idx] > thresh) break;
				}
				uint32_t pi = idx % NE;
				uint32_t x = morton3D_invert(pi >> 0), y = morton3D_invert(pi >> 1), z = morton3D_invert(pi >> 2);
				float rx = r.next_float(), ry = r.next_float(), rz = r.next_float();
				float* p = &pos[(off + i) * 3];
				p[0] = (((float)x + rx) / (float)GRIDSIZE - 0.5f) * std::scalbn(1.0f, (int)level) + 0.5f;
				p[1] = (((float)y + ry) / (float)GRIDSIZE - 0.5f) * std::scalbn(1.0f, (int)level) + 0.5f;
				p[2] = (((float)z + rz) / (float)GRIDSIZE - 0.5f) * std::scalbn(1.0f, (int)level) + 0.5f;
				idxs[off + i] = idx;
			}
		});
		o->density_rng.advance();
	};
	gen(n_uniform, 0, -0.01f);
	gen(n_nonuniform, n_uniform, MIN_OPTICAL_THICKNESS);
	std::vector<float> dens(n_total);
	orc_eval_sdf(o, pos.data(), n_total, valid_level, 0, nullptr, dens.data());
	for (uint32_t i = 0; i < n_total; ++i) {        // atomicMax on uint-punned floats (:632-634)
		uint32_t a, b; std::memcpy(&a, &tmp[idxs[i]], 4); std::memcpy(&b, &dens[i], 4);
		if (b > a) tmp[idxs[i]] = dens[i];
	}
	for (uint32_t i = 0; i < NE; ++i) { float pv = o->density_grid[i]; o->density_grid[i] = pv < 0.f ? pv : std::fmax(pv * o->density_decay, tmp[i]); }
	++o->density_ema_step;
	double acc = 0; for (uint32_t i = 0; i < NE; ++i) acc += (double)(std::fmax(o->density_grid[i], 0.f) / (float)NE);
	o->density_mean = (float)acc;
	const float thresh = std::min(MIN_OPTICAL_THICKNESS, o->density_mean);
	std::fill(o->bitfield.begin(), o->bitfield.end(), 0);
	for (uint32_t i = 0; i < NE / 8; ++i) { uint8_t bits = 0; for (int j = 0; j < 8; ++j) bits |= o->density_grid[i * 8 + j] > thresh ? (uint8_t)(1 << j) : 0; o->bitfield[i] = bits; }
	for (uint32_t lvl = 1; lvl < CASCADES; ++lvl) {
		const uint8_t* prev = o->bitfield.data() + (size_t)NE * (lvl - 1) / 8; uint8_t* next = o->bitfield.data() + (size_t)NE * lvl / 8;
		for (uint32_t i = 0; i < NE / 64; ++i) {
			uint8_t bits = 0; for (int j = 0; j < 8; ++j) bits |= prev[i * 8 + j] > 0 ? (uint8_t)(1 << j) : 0;
			uint32_t x = morton3D_invert(i >> 0) + GRIDSIZE / 8, y = morton3D_invert(i >> 1) + GRIDSIZE / 8, z = morton3D_invert(i >> 2) + GRIDSIZE / 8;
			next[morton3D(x, y, z)] |= bits;
		}
	}
}
float orc_density_mean(Oracle* o) { return o->density_mean; }

// training_prep_nerf cadence (src/testbed.cu:2805-2806, testbed_nerf.cu:4125-4138)
int orc_prep_if_due(Oracle* o) {
	uint32_t skip = std::min(std::max(o->canonical_step / 16u, 1u), 16u);      // src/testbed.cu:2805-2806
	if (o->canonical_step % skip != 0) return 0;
	uint32_t vl = valid_level_for_step(o->m, (int)o->training_step);
	const uint32_t NE = GRIDSIZE * GRIDSIZE * GRIDSIZE;
	if (o->canonical_step < 256) orc_density_update(o, NE, 0, vl); else orc_density_update(o, NE / 4, NE / 4, vl);      // testbed_nerf.cu:4133
	return 1;
}

// One full training step: Testbed::train (prep cadence) + train_nerf + optimizer + controller.
// [begin: everything up to and including backward] -> (data parallel: all-reduce o->grads and o->sums) -> [end: optimizer + controller]
void orc_train_step_begin(Oracle* o) {
	orc_prep_if_due(o);
	const uint32_t vl = valid_level_for_step(o->m, (int)o->training_step);
	const uint32_t R = o->rays_per_batch;
	const uint32_t max_samples = o->target_batch * 16;
	const bool exact = o->dp_exact && o->world > 1;
	const uint32_t local_target = exact ? o->target_batch : o->target_batch / o->world;
	uint32_t max_inference;
	if (o->measured_before == 0) { o->measured_before = max_inference = max_samples; }
	else max_inference = next_multiple(std::min(o->measured_before, max_samples), 128u);
	if (o->training_step == 0 || o->canonical_step == 0) o->n_rays_total = 0;      // testbed_nerf.cu:3906
	const uint32_t nrt = o->n_rays_total; o->n_rays_total += R;
	o->ray_indices.assign(R, 0); o->rays.assign((size_t)R * 6, 0.f); o->numsteps.assign((size_t)R * 2, 0); o->coords.resize((size_t)max_inference * 7);
	uint32_t counters[2];
	orc_generate_samples(o, R, nrt, max_inference, o->ray_indices.data(), o->rays.data(), o->numsteps.data(), o->coords.data(), counters);
	const uint32_t K = counters[0];
	uint32_t n_emitted_a = 0; for (uint32_t k = 0; k < K; ++k) n_emitted_a = std::max(n_emitted_a, o->numsteps[k * 2 + 1] + o->numsteps[k * 2]);
	std::vector<float> out_a((size_t)n_emitted_a * 16);
	orc_network_forward(o, o->coords.data(), n_emitted_a, vl, 0, o->flags.no_albedo ? 0 : 1, out_a.data(), nullptr);
	std::vector<uint32_t> n_fwd(K), cbase(K), n_emit(K);
	uint32_t total = orc_compact(o, out_a.data(), o->numsteps.data(), K, local_target, n_fwd.data(), cbase.data(), n_emit.data());
	// gather compacted coords / outputs
	std::vector<float> cc((size_t)total * 7), oc((size_t)total * 16), dout((size_t)total * 16, 0.f);
	for (uint32_t k = 0; k < K; ++k) {
		std::memcpy(&cc[(size_t)cbase[k] * 7], &o->coords[(size_t)o->numsteps[k * 2 + 1] * 7], (size_t)n_fwd[k] * 7 * 4);
		std::memcpy(&oc[(size_t)cbase[k] * 16], &out_a[(size_t)o->numsteps[k * 2 + 1] * 16], (size_t)n_fwd[k] * 16 * 4);
	}
	std::vector<float> loss(K), ek(K), ml(K);
	orc_loss(o, oc.data(), o->ray_indices.data(), n_fwd.data(), cbase.data(), n_emit.data(), K, R, nrt, o->training_step, dout.data(), loss.data(), ek.data(), ml.data());
	o->last_loss = loss; o->last_ek = ek; o->last_mask = ml; o->ray_indices.resize(K);
	const uint32_t n_in = std::min(total, local_target);
	std::vector<uint8_t> foreign;
	uint32_t total_own = total;
	if (exact) {
		foreign.assign(total, 0); total_own = 0;
		for (uint32_t k = 0; k < K; ++k) {
			if (o->ray_indices[k] % o->world == o->rank) total_own += n_fwd[k];
			else std::fill(foreign.begin() + cbase[k], foreign.begin() + cbase[k] + n_fwd[k], (uint8_t)1);
		}
	}
	network_backward_impl(o, cc.data(), dout.data(), n_in, n_in, local_target, o->target_batch, vl, exact ? foreign.data() : nullptr);
	o->rng.advance();
	double sl = 0, se = 0, sm = 0; for (uint32_t k = 0; k < K; ++k) { sl += loss[k]; se += ek[k]; sm += ml[k]; }
	o->sums[0] = sl; o->sums[1] = se; o->sums[2] = sm; o->sums[3] = (double)total_own;      // all-reduced: the global compacted count either way
	o->cnt_kept = K; o->cnt_samples = counters[1]; o->cnt_total = total; o->cnt_trained = n_in; o->step_R = R;
	o->in_step = true;
}

void orc_train_step_end(Oracle* o, OrcStats* st) {
	orc_optimizer_step(o);
	++o->training_step;
	o->canonical_step = o->training_step;
	o->in_step = false;
	const uint32_t total = o->cnt_total, R = o->step_R;
	if (o->cnt_samples == 0 || total == 0) { o->measured_before = 0; o->measured = 0; }      // Counters::update_after_training, testbed_nerf.cu:3540-3542
	else { o->measured_before = o->cnt_samples; o->measured = total; }
	const float f = (float)o->sums[3] / (float)o->target_batch;     // sums[3] is the global compacted count after the all-reduce
	if (st) { st->loss = (float)o->sums[0] * f; st->ek_loss = (float)o->sums[1] * f; st->mask_loss = (float)o->sums[2] * f; st->n_rays_kept = o->cnt_kept; st->n_samples = o->cnt_samples; st->n_compacted = total; st->n_emitted = o->cnt_trained; }
	// data parallel: sums[3] is the all-reduced compacted count, so that every rank derives the same next batch size
	const uint32_t total_global = o->world > 1 ? (uint32_t)(o->sums[3] + 0.5) : total;
	if (!o->pin_rays && total_global > 0 && (o->world > 1 || o->cnt_samples != 0)) {
		uint32_t r = (uint32_t)((float)R * (float)o->target_batch / (float)total_global);
		o->rays_per_batch = std::min(next_multiple(r, 128u), 1u << 18);
	}
	if (st) st->rays_per_batch_next = o->rays_per_batch;
}

void orc_train_step(Oracle* o, OrcStats* st) { orc_train_step_begin(o); orc_train_step_end(o, st); }
void orc_get_sums(Oracle* o, double* out) { for (int i = 0; i < 4; ++i) out[i] = o->sums[i]; }
void orc_set_sums(Oracle* o, const double* in) { for (int i = 0; i < 4; ++i) o->sums[i] = in[i]; }

// per-ray loss terms of the last orc_train_step (compute_loss_kernel's loss_output / ek_loss_output / mask_loss_output)
uint32_t orc_get_last_losses(Oracle* o, uint32_t cap, uint32_t* ray_idx, float* loss, float* ek, float* mask) {
	const uint32_t K = (uint32_t)std::min<size_t>(cap, o->last_loss.size());
	for (uint32_t k = 0; k < K; ++k) { ray_idx[k] = o->ray_indices[k]; loss[k] = o->last_loss[k]; ek[k] = o->last_ek[k]; mask[k] = o->last_mask[k]; }
	return K;
}

// ---- double-precision (smooth) entry points for finite-difference checks only ----
void orc_forward_f64(Oracle* o, const double* params, const float* coords, uint64_t n, uint32_t valid_level, double* out) {
	Net<false> net(o->m, params, valid_level);
	Net<false>::Ctx c;
	for (size_t i = 0; i < n; ++i) { net.forward(coords + i * 7, c, true); for (int k = 0; k < 16; ++k) out[i * 16 + k] = c.out[k]; }
}
void orc_backward_f64(Oracle* o, const double* params, const float* coords, const double* dout, uint64_t n, uint32_t n_batch, uint32_t valid_level, double* grads) {
	Net<false> net(o->m, params, valid_level);
	Net<false>::Ctx c;
	std::fill(grads, grads + o->m.n_params, 0.0);
	double gvar = 0;
	for (size_t i = 0; i < n; ++i) { net.forward(coords + i * 7, c, true); net.backward(c, dout + i * 16, 1.0, n_batch, grads, grads, &gvar, false); }
	grads[o->m.off_var] += gvar;
}

} // extern "C"

extern "C" void orc_pcg32_seq(uint64_t seed, uint64_t seq, int64_t advance, uint32_t n, uint32_t* out_u) {
	orc::Pcg32 r(seed, seq); r.advance(advance);
	for (uint32_t i = 0; i < n; ++i) out_u[i] = r.next_uint();
}

// ---- mesh path (orc_mesh.h) ---------------------------------------------------------------------------------------------
struct OrcMesh { orc_mesh::Mesh m; };
extern "C" OrcMesh* orc_marching_cubes(const float* density, const uint32_t res[3], const float mn[3], const float mx[3], float thresh, uint32_t counts[3]) {
	OrcMesh* h = new OrcMesh{orc_mesh::marching_cubes(density, res, mn, mx, thresh)};
	counts[0] = h->m.n_verts; counts[1] = (uint32_t)(h->m.verts.size() / 3); counts[2] = (uint32_t)h->m.indices.size();
	return h;
}
extern "C" void orc_mesh_get(OrcMesh* h, float* verts, float* normals, uint32_t* indices) {
	memcpy(verts, h->m.verts.data(), h->m.verts.size() * 4); memcpy(normals, h->m.normals.data(), h->m.normals.size() * 4);
	memcpy(indices, h->m.indices.data(), h->m.indices.size() * 4);
}
extern "C" void orc_mesh_free(OrcMesh* h) { delete h; }
extern "C" int orc_save_mesh(const float* verts, const float* normals, const float* colors, const uint32_t* indices, uint32_t n_verts, uint32_t n_indices, const char* path,
                             float nerf_scale, const float off[3], float n2w_s, const float n2w_t[3], int invert_normals) {
	return orc_mesh::save_mesh(verts, normals, colors, indices, n_verts, n_indices, path, nerf_scale, off, n2w_s, n2w_t, invert_normals);
}
