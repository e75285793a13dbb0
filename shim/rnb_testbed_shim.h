// rnb_testbed_shim.h — the reference-side binding of librnb_b200.so, as code that COMPILES against the reference's own headers.
//
// A maintainer of RobinBruneau/RNb-NeuS2 includes this file at the top of src/testbed.cu and src/testbed_nerf.cu (after testbed.h) and
// adds six one-line calls behind `#ifdef NGP_USE_RNB_B200` (INTEGRATION.md §2 lists them; oracle/shim_patch.py applies exactly those to
// a scratch copy of the two files and `make -C oracle -f Makefile.ref shim` builds oracle/_ref/bin/testbed_rnb from it — the reference
// sources themselves are never modified or copied into this repository).
//
//   Testbed::reset_network()            ... end:    rnb_shim::on_reset_network(*this);
//   Testbed::load_nerf()                ... end:    rnb_shim::on_dataset(*this);
//   Testbed::train(batch)               ... start:  if (m_testbed_mode == ETestbedMode::Nerf && rnb_shim::train(*this)) { update_loss_graph(); return; }
//   Testbed::compute_and_save_marching_cubes_mesh(...) start:  if (... rnb_shim::compute_and_save_mesh(*this, filename, res3d, aabb, thresh, unwrap_it)) return;
//   Testbed::save_snapshot(...)         ... start:  rnb_shim::push_state(*this);
//   Testbed::load_snapshot(...)         ... end:    rnb_shim::pull_state(*this);
//
// Design: the library owns the training state while training runs; `push_state` copies it into the reference's own objects (trainer
// parameters, density grid + bitfield, step counters) so that EVERY stock consumer — snapshot writer, renderer, the stock marching
// cubes — keeps working on trained weights without knowing about the library; `pull_state` goes the other way after a snapshot load.
// One Testbed per process (./build/testbed), hence one context.  Data parallel: N processes, one per GPU, RNB_WORLD_SIZE / RNB_RANK / RNB_COMM_ID_FILE
// (install_communicator below; run on 2 B200s: profiles/r02_shim_dp_record.json).
#pragma once
#ifdef NGP_USE_RNB_B200

#include <rnb_b200.h>
#include <neural-graphics-primitives/testbed.h>
#include <neural-graphics-primitives/nerf_network.h>
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/trainer.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace rnb_shim {

inline rnb_ctx*& ctx() { static rnb_ctx* c = nullptr; return c; }
inline bool& state_dirty() { static bool d = false; return d; }          // the library holds newer weights than the reference's trainer

inline void check(int rc) {      // CUDA_CHECK_THROW behaviour: a failing call ends the run with the library's message
	if (rc != RNB_OK) throw std::runtime_error{std::string{"rnb_b200: "} + rnb_last_error()};
}

inline uint32_t grid_cells() { return ngp::NERF_GRIDSIZE() * ngp::NERF_GRIDSIZE() * ngp::NERF_GRIDSIZE(); }

// --- data parallel: N copies of ./build/testbed, one per GPU (CUDA_VISIBLE_DEVICES), same command line, each with its own output directory ----------
//   RNB_WORLD_SIZE=N RNB_RANK=r RNB_COMM_ID_FILE=/shared/path ./build/testbed --scene ... (rank r marches rays i = r mod N of the SAME global batch,
//   gradients are exchanged inside rnb_train; with one sample order the N-GPU run reproduces the single-GPU run)
// The 128-byte NCCL id travels through a file: rank 0 writes <file>.<generation> (temporary name + rename), the others wait for it.  The generation
// counts the networks this process has created (load_snapshot -> reset_network creates a second one), the same on every rank.  The launcher removes
// <file>.* of an earlier run before it starts the ranks (a stale id file would be read as it is).
inline uint32_t env_u32(const char* name, uint32_t dflt) { const char* v = std::getenv(name); return v && *v ? (uint32_t)std::strtoul(v, nullptr, 10) : dflt; }
inline uint32_t world_size() { static const uint32_t w = std::max(env_u32("RNB_WORLD_SIZE", 1u), 1u); return w; }
inline uint32_t world_rank() { static const uint32_t r = env_u32("RNB_RANK", 0u); return r; }
inline void install_communicator() {
	static uint32_t generation = 0;
	const char* base = std::getenv("RNB_COMM_ID_FILE");
	if (!base || !*base) throw std::runtime_error{"rnb_b200: RNB_WORLD_SIZE > 1 needs RNB_COMM_ID_FILE (a path every rank can read)"};
	const std::string path = std::string{base} + "." + std::to_string(generation++);
	uint8_t id[RNB_COMM_ID_BYTES];
	if (world_rank() == 0) {
		check(rnb_comm_unique_id(id));
		const std::string tmp = path + ".tmp";
		FILE* f = std::fopen(tmp.c_str(), "wb");
		if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) { if (f) std::fclose(f); throw std::runtime_error{"rnb_b200: cannot write " + tmp}; }
		std::fclose(f);
		if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error{"rnb_b200: cannot publish " + path};
	} else {
		const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(env_u32("RNB_COMM_TIMEOUT_S", 120u));
		for (;;) {
			if (FILE* f = std::fopen(path.c_str(), "rb")) {
				const size_t got = std::fread(id, 1, sizeof(id), f);
				std::fclose(f);
				if (got == sizeof(id)) break;
			}
			if (std::chrono::steady_clock::now() > deadline) throw std::runtime_error{"rnb_b200: rank 0 did not publish " + path};
			std::this_thread::sleep_for(std::chrono::milliseconds(20));
		}
	}
	check(rnb_comm_init(ctx(), id));
}

// --- dataset: device pointers stay owned by the reference's loader (metadata_normal / metadata_albedo, nerf_loader.h:88-89) -----------
inline void on_dataset(ngp::Testbed& t) {
	if (!ctx() || t.m_testbed_mode != ngp::ETestbedMode::Nerf) return;
	const auto& ds = t.m_nerf.training.dataset;
	if (ds.n_images == 0 || ds.metadata_normal.size() < ds.n_images || t.m_nerf.training.transforms.size() < ds.n_images) return;
	std::vector<rnb_view> views(ds.n_images);
	for (size_t i = 0; i < views.size(); ++i) {
		const auto& mn = ds.metadata_normal[i];
		views[i].normal_px = mn.pixels;
		views[i].albedo_px = i < ds.metadata_albedo.size() ? ds.metadata_albedo[i].pixels : nullptr;
		views[i].w = mn.resolution.x(); views[i].h = mn.resolution.y();
		views[i].fx = mn.focal_length.x(); views[i].fy = mn.focal_length.y();
		views[i].cx = mn.principal_point.x(); views[i].cy = mn.principal_point.y();
		std::memcpy(views[i].xform, t.m_nerf.training.transforms[i].start.data(), 12 * sizeof(float));      // Eigen 3x4, column-major
	}
	check(rnb_set_dataset(ctx(), views.data(), (uint32_t)views.size()));
}

// --- network: same configuration, and the reference's OWN initial parameters (no second initialiser to keep in sync) ----------------
inline void on_reset_network(ngp::Testbed& t) {
	if (t.m_testbed_mode != ngp::ETestbedMode::Nerf) return;
	using json = nlohmann::json;
	const json& cfgj = t.m_network_config;
	const json& enc = cfgj.contains("encoding") ? cfgj["encoding"] : json::object();
	const json& net = cfgj.contains("network") ? cfgj["network"] : json::object();
	const json& rgb = cfgj.contains("rgb_network") ? cfgj["rgb_network"] : json::object();
	rnb_config cfg; rnb_default_config(&cfg);
	cfg.n_levels = enc.value("n_levels", 14u);
	cfg.log2_hashmap_size = enc.value("log2_hashmap_size", 19u);
	cfg.base_resolution = t.m_base_grid_resolution;
	cfg.per_level_scale = t.m_per_level_scale;
	cfg.top_resolution = enc.value("top_resolution", 2048.0f);
	cfg.base_valid_level_scale = enc.value("base_valid_level_scale", 0.2f);
	cfg.valid_level_scale = enc.value("valid_level_scale", 0.02f);
	cfg.base_training_step = enc.value("base_training_step", 100u);
	cfg.sdf_n_neurons = net.value("n_neurons", 64u);
	cfg.sdf_n_hidden_layers = net.value("n_hidden_layers", 1u);
	cfg.sdf_bias = net.value("sdf_bias", -0.1f);
	cfg.rgb_n_neurons = rgb.value("n_neurons", 64u);
	cfg.rgb_n_hidden_layers = rgb.value("n_hidden_layers", 2u);
	// optimizer: Ema(ExponentialDecay(Adam)) as in configs/nerf/base.json:5-29 — walk the nesting, take what each level defines
	const json* o = cfgj.contains("optimizer") ? &cfgj["optimizer"] : nullptr;
	while (o) {
		const std::string otype = o->value("otype", std::string{});
		if (otype == "Ema") cfg.ema_decay = o->value("decay", cfg.ema_decay);
		else if (otype == "ExponentialDecay") {
			cfg.lr_decay_start = o->value("decay_start", cfg.lr_decay_start);
			cfg.lr_decay_interval = o->value("decay_interval", cfg.lr_decay_interval);
			cfg.lr_decay_base = o->value("decay_base", cfg.lr_decay_base);
		} else if (otype == "Adam") {
			cfg.learning_rate = o->value("learning_rate", cfg.learning_rate);
			cfg.beta1 = o->value("beta1", cfg.beta1); cfg.beta2 = o->value("beta2", cfg.beta2);
			cfg.epsilon = o->value("epsilon", cfg.epsilon); cfg.l2_reg = o->value("l2_reg", cfg.l2_reg);
		}
		o = o->contains("nested") ? &(*o)["nested"] : nullptr;
	}
	cfg.seed = t.m_seed;
	cfg.rays_per_batch = t.m_nerf.training.counters_rgb.rays_per_batch;
	cfg.pin_rays_per_batch = 0;                                    // the reference's adaptive controller stays in charge
	cfg.density_grid_decay = t.m_nerf.training.density_grid_decay;
	cfg.world_size = world_size(); cfg.rank = world_rank();      // the global batch (rays per step, 2^18-sample budget) is the reference's; the ranks share it
	if (cfg.rank >= cfg.world_size) throw std::runtime_error{"rnb_b200: RNB_RANK must be below RNB_WORLD_SIZE"};
	if (ctx()) { rnb_destroy(ctx()); ctx() = nullptr; }
	check(rnb_create(&cfg, &ctx()));
	if (cfg.world_size > 1) install_communicator();
	// initial parameters: whatever the reference's Trainer just initialised (fp32 master copy), in the same order (nerf_network.h:539-583)
	uint64_t layout[5]; check(rnb_param_layout(ctx(), layout));
	const size_t n = t.m_network->n_params();
	if (n != layout[4]) throw std::runtime_error{"rnb_b200: parameter count differs from the reference network (" + std::to_string(n) + " vs " + std::to_string(layout[4]) + ")"};
	// binary16 copy first (training + inference parameters, Adam state cleared), then the exact fp32 master weights on top
	std::vector<uint16_t> h16(n);
	CUDA_CHECK_THROW(cudaMemcpy(h16.data(), t.m_trainer->params(), n * sizeof(uint16_t), cudaMemcpyDeviceToHost));
	check(rnb_import_params_fp16(ctx(), h16.data(), n));
	std::vector<float> w(n);
	CUDA_CHECK_THROW(cudaMemcpy(w.data(), t.m_trainer->params_full_precision(), n * sizeof(float), cudaMemcpyDeviceToHost));
	check(rnb_set_params_fp32(ctx(), w.data(), n));
	state_dirty() = false;
	on_dataset(t);
}

inline rnb_flags flags_of(ngp::Testbed& t) {
	rnb_flags f; rnb_default_flags(&f);
	f.apply_L2 = t.m_apply_L2; f.apply_supernormal = t.m_apply_supernormal; f.apply_rgbplus = t.m_apply_rgbplus; f.apply_relu = t.m_apply_relu;
	f.apply_bce = t.m_apply_bce; f.light_opti = t.m_light_opti; f.no_albedo = t.m_no_albedo;
	f.mask_loss_weight = t.m_mask_loss_weight; f.ek_loss_weight = t.m_ek_loss_weight;
	f.cos_anneal_ratio = t.m_nerf_network->cos_anneal_ratio();
	// --fractional-training warm-up (Testbed::frame, src/testbed.cu:1886-1895): geometry only, the colour MLP is frozen
	f.only_sdf_training = (t.m_fractional_training && t.m_training_step < t.m_fractional) ? 1 : 0;
	return f;
}

// --- one call of Testbed::train: occupancy refresh cadence + train_nerf + optimizer step (src/testbed.cu:2776-2872) -----------------
inline bool train(ngp::Testbed& t) {
	if (!ctx()) return false;
	rnb_flags f = flags_of(t);
	check(rnb_set_flags(ctx(), &f));
	rnb_step_stats st;
	check(rnb_train(ctx(), t.m_training_stream, &st));              // returns with the stream synchronised, like the reference
	state_dirty() = true;
	t.m_training_step = st.training_step;
	t.m_canonical_training_step = (int)st.training_step;
	if (st.density_grid_updated) t.m_nerf.training.n_images_for_training_prev = t.m_nerf.training.n_images_for_training;      // update_density_grid_nerf :3447
	t.m_nerf_network->m_training_step = st.training_step;
	t.m_loss_scalar.update(st.loss); t.m_ek_loss_scalar.update(st.ek_loss); t.m_mask_loss_scalar.update(st.mask_loss);
	auto& c = t.m_nerf.training.counters_rgb;
	c.rays_per_batch = st.rays_per_batch_next;
	c.measured_batch_size = st.n_samples_compacted;
	c.measured_batch_size_before_compaction = st.n_samples;
	c.n_rays_total += st.n_rays;
	if (st.n_samples_compacted == 0 && world_size() == 1) {      // data parallel: the count is this rank's; a rank must not leave the collective on its own
		tlog::warning() << "Nerf training generated 0 samples. Aborting training.";
		t.m_train = false;
	}
	return true;
}

// --- library -> reference objects: after this every stock consumer (snapshot, renderer, stock marching cubes) sees the trained state ---
inline void push_state(ngp::Testbed& t) {
	if (!ctx() || !state_dirty()) return;
	using precision_t = tcnn::network_precision_t;
	static_assert(sizeof(precision_t) == 2, "the library exchanges binary16 parameters");
	const size_t n = t.m_network->n_params();
	std::vector<uint16_t> h(n);
	if (world_size() > 1) {      // sharded optimizer: gather the per-shard EMA copy (collective: every rank saves at the same point)
		check(rnb_comm_sync_ema(ctx(), t.m_training_stream)); CUDA_CHECK_THROW(cudaStreamSynchronize(t.m_training_stream));
	}
	check(rnb_export_params_fp16(ctx(), h.data(), n, /*use_ema=*/1));
	CUDA_CHECK_THROW(cudaMemcpy(t.m_trainer->params_inference(), h.data(), n * 2, cudaMemcpyHostToDevice));
	check(rnb_export_params_fp16(ctx(), h.data(), n, /*use_ema=*/0));
	CUDA_CHECK_THROW(cudaMemcpy(t.m_trainer->params(), h.data(), n * 2, cudaMemcpyHostToDevice));
	std::vector<float> w(n);
	check(rnb_get_params_fp32(ctx(), w.data(), n));
	CUDA_CHECK_THROW(cudaMemcpy(t.m_trainer->params_full_precision(), w.data(), n * 4, cudaMemcpyHostToDevice));
	const uint32_t cells = grid_cells();
	std::vector<float> grid(cells); uint32_t ema_step = 0;
	check(rnb_export_density_grid(ctx(), grid.data(), cells, &ema_step));
	if (t.m_nerf.density_grid.size() < cells) t.m_nerf.density_grid.resize(cells);
	CUDA_CHECK_THROW(cudaMemcpy(t.m_nerf.density_grid.data(), grid.data(), (size_t)cells * 4, cudaMemcpyHostToDevice));
	t.m_nerf.density_grid_ema_step = ema_step;
	std::vector<uint8_t> bits(cells);                                // 8 mips x 128^3 bits
	check(rnb_get_bitfield(ctx(), bits.data(), cells));
	const size_t nb = std::min<size_t>(t.m_nerf.density_grid_bitfield.size(), bits.size());
	if (nb) CUDA_CHECK_THROW(cudaMemcpy(t.m_nerf.density_grid_bitfield.data(), bits.data(), nb, cudaMemcpyHostToDevice));
	uint32_t ts[4]; check(rnb_get_train_state(ctx(), ts));
	t.m_training_step = ts[0];
	state_dirty() = false;
}

// --- reference objects -> library: after Testbed::load_snapshot --------------------------------------------------------------------
inline void pull_state(ngp::Testbed& t) {
	if (!ctx() || t.m_testbed_mode != ngp::ETestbedMode::Nerf) return;
	const size_t n = t.m_network->n_params();
	std::vector<uint16_t> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), t.m_trainer->params(), n * 2, cudaMemcpyDeviceToHost));
	check(rnb_import_params_fp16(ctx(), h.data(), n));               // Adam moments restart, as in the reference (trainer.h:263-275)
	const uint32_t cells = grid_cells();
	if (t.m_nerf.density_grid.size() >= cells) {
		std::vector<float> grid(cells);
		CUDA_CHECK_THROW(cudaMemcpy(grid.data(), t.m_nerf.density_grid.data(), (size_t)cells * 4, cudaMemcpyDeviceToHost));
		check(rnb_import_density_grid(ctx(), grid.data(), cells, t.m_nerf.density_grid_ema_step));
	}
	const auto& c = t.m_nerf.training.counters_rgb;
	check(rnb_set_train_state(ctx(), (uint32_t)t.m_training_step, c.rays_per_batch, c.n_rays_total, c.measured_batch_size_before_compaction));
	// what load_snapshot does NOT restore: the canonical step (0 after reset_network) and the image count of the last occupancy refresh
	check(rnb_set_canonical_state(ctx(), (uint32_t)t.m_canonical_training_step, (uint32_t)t.m_nerf.training.n_images_for_training_prev));
	state_dirty() = false;
}

// --- Testbed::compute_and_save_marching_cubes_mesh (src/testbed.cu:369-381): sweep + extraction + text on the GPU -----------------------
inline bool compute_and_save_mesh(ngp::Testbed& t, const char* filename, Eigen::Vector3i res3d, const ngp::BoundingBox& aabb, float thresh, bool unwrap_it) {
	if (!ctx()) return false;
	if (unwrap_it) { push_state(t); return false; }                  // the unwrap / texture variant stays with the reference, on the trained weights
	if (thresh == std::numeric_limits<float>::max()) thresh = t.m_mesh.thresh;
	const uint32_t r[3] = {(uint32_t)res3d.x(), (uint32_t)res3d.y(), (uint32_t)res3d.z()};      // rounded up to multiples of 16 inside, like :4298-4300
	rnb_mesh_info mi;
	if (world_size() > 1) { check(rnb_comm_sync_ema(ctx(), t.m_training_stream)); CUDA_CHECK_THROW(cudaStreamSynchronize(t.m_training_stream)); }
	check(rnb_marching_cubes(ctx(), r, aabb.min.data(), aabb.max.data(), thresh, /*use_ema=*/1, t.m_inference_stream, &mi));
	float *v = nullptr, *nrm = nullptr, *col = nullptr; uint32_t* idx = nullptr;
	check(rnb_mesh_buffers(ctx(), &v, &nrm, &col, &idx, nullptr));
	const auto& ds = t.m_nerf.training.dataset;
	check(rnb_save_mesh(v, nrm, col, idx, mi.n_verts_padded, mi.n_indices, filename, ds.scale, ds.offset.data(), ds.n2w_s, ds.n2w_t.data(), ds.from_na ? 1 : 0, t.m_inference_stream, nullptr));
	return true;
}

} // namespace rnb_shim
#endif // NGP_USE_RNB_B200
