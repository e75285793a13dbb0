// ref_harness.cu — drives the UNMODIFIED reference (RobinBruneau/RNb-NeuS2) on the GPU box and dumps golden vectors.
//
// TEST INFRASTRUCTURE ONLY (oracle/): linked by oracle/Makefile.ref against the reference's own objects
// (src/testbed*.cu, tiny-cuda-nn) compiled from /root/reference; never part of the product library.
// It calls the reference's public Testbed API exactly like src/main.cu does (load_training_data,
// reload_network_from_file, apply_* setters, train) and, around every Testbed::train call, copies the state that the
// hot path reads and writes to plain binary files:
//
//   dataset.bin            per view: w h fx fy cx cy xform[12] (fp32, as loaded by nerf_loader.cu) + raw uint16 RGBA pixels
//   step<k>_in_*.bin       params fp32, density grid, bitfield, pcg32 states, controller counters  (state BEFORE step k)
//   step<k>_out_*.bin      per-ray loss / ek / mask arrays, sample counters, fp16 gradient buffer (loss-scaled, before Adam
//                          consumes it: tcnn keeps it until the next backward), params fp32 + EMA fp16 AFTER Adam
//   probe_*.bin            NerfNetwork::inference (16-wide fp16 rows) and sdf() at harness-chosen positions
//   meta.txt               key=value lines
//
// The only behavioural pin is the light index of the loss kernel (clock64-seeded curand -> ray index % 3), injected by
// oracle/ref_prelude.h when compiling src/testbed_nerf.cu; see there.
#include <neural-graphics-primitives/testbed.h>
#include <neural-graphics-primitives/nerf_network.h>
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/trainer.h>
#include <tiny-cuda-nn/network.h>
#include <neural-graphics-primitives/marching_cubes.h>
#include <filesystem/path.h>
#include <chrono>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <unistd.h>
#include <limits.h>

using namespace ngp;
using namespace tcnn;
namespace fs = ::filesystem;

static fs::path g_exe_dir;
fs::path get_executable_dir_global() { return g_exe_dir; }      // the reference resolves utils/ relative to this (nerf_network.h:591)

static std::string g_out;

template <typename T>
static void dump_dev(const std::string& name, const T* dev, size_t n) {
	std::vector<T> h(n);
	CUDA_CHECK_THROW(cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
	FILE* f = fopen((g_out + "/" + name).c_str(), "wb");
	if (!f) { fprintf(stderr, "cannot write %s\n", name.c_str()); exit(2); }
	fwrite(h.data(), sizeof(T), n, f);
	fclose(f);
}
static void dump_host(const std::string& name, const void* p, size_t bytes) {
	FILE* f = fopen((g_out + "/" + name).c_str(), "wb");
	if (!f) { fprintf(stderr, "cannot write %s\n", name.c_str()); exit(2); }
	fwrite(p, 1, bytes, f);
	fclose(f);
}

int main(int argc, char** argv) {
	if (argc < 5) {
		fprintf(stderr, "usage: ref_harness <scene_dir> <network_config.json> <out_dir> <n_steps> [--no-albedo] [--supernormal] [--opti-lights] [--l1] [--no-rgbplus] [--dump-every K | --dump-steps a,b,c] [--time-only] [--time-from K] [--pin-rays N] [--mesh RES] [--save-snapshot FILE] [--load-snapshot FILE] [--print-every K]\n");
		return 1;
	}
	char buf[PATH_MAX]; ssize_t cnt = readlink("/proc/self/exe", buf, PATH_MAX);
	if (cnt > 0) { buf[cnt] = 0; g_exe_dir = fs::path(buf).parent_path(); } else g_exe_dir = fs::path(".");
	const std::string scene = argv[1], config = argv[2];
	g_out = argv[3];
	const int n_steps = atoi(argv[4]);
	bool no_albedo = false, supernormal = false, opti = false, l1 = false, rgbplus = true, time_only = false;
	int dump_every = 1; uint32_t pin_rays = 0; std::vector<int> dump_steps; int time_from = -1; int mesh_res = 0; int print_every = 50; std::string save_snapshot, load_snapshot;
	for (int i = 5; i < argc; ++i) {
		std::string a = argv[i];
		if (a == "--no-albedo") no_albedo = true; else if (a == "--supernormal") supernormal = true; else if (a == "--opti-lights") opti = true;
		else if (a == "--l1") l1 = true; else if (a == "--no-rgbplus") rgbplus = false; else if (a == "--time-only") time_only = true;
		else if (a == "--dump-every" && i + 1 < argc) dump_every = atoi(argv[++i]);
		else if (a == "--pin-rays" && i + 1 < argc) pin_rays = (uint32_t)atoi(argv[++i]);
		else if (a == "--time-from" && i + 1 < argc) time_from = atoi(argv[++i]);
		else if (a == "--mesh" && i + 1 < argc) mesh_res = atoi(argv[++i]);
		else if (a == "--print-every" && i + 1 < argc) print_every = std::max(1, atoi(argv[++i]));
		else if (a == "--save-snapshot" && i + 1 < argc) save_snapshot = argv[++i];
		else if (a == "--load-snapshot" && i + 1 < argc) load_snapshot = argv[++i];
		else if (a == "--dump-steps" && i + 1 < argc) { std::string l = argv[++i]; size_t p0 = 0; while (p0 < l.size()) { size_t q = l.find(',', p0); if (q == std::string::npos) q = l.size(); dump_steps.push_back(atoi(l.substr(p0, q - p0).c_str())); p0 = q + 1; } }
	}

	Testbed tb{ETestbedMode::Nerf};
	tb.set_max_iter((uint32_t)n_steps + 1000000u);
	const auto load_t0 = std::chrono::steady_clock::now();
	tb.load_training_data(scene);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	const double load_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - load_t0).count();      // Testbed::load_training_data = load_nerf: JSON + image decode + upload
	tb.reload_network_from_file(config);
	tb.m_train = true;
	// same order as src/main.cu:369-398
	if (!l1) tb.apply_L2();
	if (supernormal) tb.apply_supernormal();
	if (rgbplus) tb.apply_rgbplus();
	if (opti) tb.apply_light_opti();
	if (no_albedo) tb.apply_no_albedo(true);
	// snapshot hand-off: Testbed::load_snapshot exactly as src/main.cu:312 calls it, after the dataset and the network config
	if (!load_snapshot.empty()) tb.load_snapshot(load_snapshot);

	auto& tr = tb.m_nerf.training;
	const size_t n_params = tb.m_network->n_params();
	FILE* meta = fopen((g_out + "/meta.txt").c_str(), "w");
	fprintf(meta, "n_params=%zu\nn_images=%zu\nn_steps=%d\nno_albedo=%d\nsupernormal=%d\nopti_lights=%d\nl1=%d\nrgbplus=%d\n", n_params, tr.dataset.n_images, n_steps, no_albedo, supernormal, opti, l1, rgbplus);
	fprintf(meta, "load_seconds=%.4f\n", load_seconds);
	fprintf(meta, "mask_loss_weight=%g\nek_loss_weight=%g\naabb_min=%g %g %g\naabb_max=%g %g %g\n", tb.m_mask_loss_weight, tb.m_ek_loss_weight,
	        tb.m_aabb.min.x(), tb.m_aabb.min.y(), tb.m_aabb.min.z(), tb.m_aabb.max.x(), tb.m_aabb.max.y(), tb.m_aabb.max.z());

	if (!time_only) {
		// dataset exactly as the loader left it on the device
		std::vector<uint8_t> ds;
		auto put = [&](const void* p, size_t b) { const uint8_t* q = (const uint8_t*)p; ds.insert(ds.end(), q, q + b); };
		for (size_t i = 0; i < tr.dataset.n_images; ++i) {
			const auto& mn = tr.dataset.metadata_normal[i]; const auto& ma = tr.dataset.metadata_albedo[i];
			int32_t wh[2] = {mn.resolution.x(), mn.resolution.y()};
			float k[4] = {mn.focal_length.x(), mn.focal_length.y(), mn.principal_point.x(), mn.principal_point.y()};
			float xf[12];
			for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) xf[c * 3 + r] = tr.dataset.xforms[i].start(r, c);   // column-major 3x4
			put(wh, 8); put(k, 16); put(xf, 48);
			const size_t bytes = (size_t)wh[0] * wh[1] * 8;
			std::vector<uint8_t> px(bytes);
			CUDA_CHECK_THROW(cudaMemcpy(px.data(), mn.pixels, bytes, cudaMemcpyDeviceToHost)); put(px.data(), bytes);
			CUDA_CHECK_THROW(cudaMemcpy(px.data(), ma.pixels, bytes, cudaMemcpyDeviceToHost)); put(px.data(), bytes);
		}
		dump_host("dataset.bin", ds.data(), ds.size());
	}

	auto dump_state = [&](const std::string& tag) {
		dump_dev(tag + "_params_fp32.bin", tb.m_trainer->params_full_precision(), n_params);
		dump_dev(tag + "_density_grid.bin", tb.m_nerf.density_grid.data(), tb.m_nerf.density_grid.size());
		dump_dev(tag + "_bitfield.bin", tb.m_nerf.density_grid_bitfield.data(), tb.m_nerf.density_grid_bitfield.size());
		uint64_t st[8] = {tb.m_rng.state, tb.m_rng.inc, tr.density_grid_rng.state, tr.density_grid_rng.inc,
		                  tb.m_training_step, tr.counters_rgb.rays_per_batch, tr.counters_rgb.n_rays_total, tr.counters_rgb.measured_batch_size_before_compaction};
		uint64_t st2[2] = {tb.m_nerf.density_grid_ema_step, tr.counters_rgb.measured_batch_size};
		std::vector<uint64_t> all(st, st + 8); all.push_back(st2[0]); all.push_back(st2[1]);
		dump_host(tag + "_state.bin", all.data(), all.size() * 8);
	};

	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	double total_ms = 0; uint64_t total_rays = 0; int timed_steps = 0;
	for (int k = 0; k < n_steps; ++k) {
		bool dump = !time_only && (k % dump_every == 0 || k == n_steps - 1);
		if (!time_only && !dump_steps.empty()) { dump = false; for (int d : dump_steps) dump |= (d == k); }
		const std::string tag = "step" + std::to_string(k);
		// benchmark / parity knob: overwrite the batch-size controller's output (a public member) so that every step marches the
		// same number of rays; with N <= 256 the compacted count can never exceed 2^18 and no arrival-order truncation occurs
		// (and the clamp of the sample budget to last step's count, whose victims also depend on arrival order: testbed_nerf.cu:3891-3896,1354-1357)
		if (pin_rays) { tr.counters_rgb.rays_per_batch = pin_rays; if (!time_only) tr.counters_rgb.measured_batch_size_before_compaction = 0; }
		if (dump) dump_state(tag + "_in");
		const uint32_t R = tr.counters_rgb.rays_per_batch;
		cudaEventRecord(e0, tb.m_training_stream);
		tb.train(1u << 18);
		cudaEventRecord(e1, tb.m_training_stream);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
		if (k >= (time_from >= 0 ? time_from : n_steps / 2)) { total_ms += ms; total_rays += R; ++timed_steps; }
		if (dump) {
			dump_dev(tag + "_out_loss.bin", tr.counters_rgb.loss.data(), R);
			dump_dev(tag + "_out_ek_loss.bin", tr.counters_rgb.ek_loss.data(), R);
			dump_dev(tag + "_out_mask_loss.bin", tr.counters_rgb.mask_loss.data(), R);
			dump_dev(tag + "_out_grads_fp16.bin", (const uint16_t*)tb.m_trainer->param_gradients(), n_params);
			dump_dev(tag + "_out_params_fp32.bin", tb.m_trainer->params_full_precision(), n_params);
			dump_dev(tag + "_out_params_ema_fp16.bin", (const uint16_t*)tb.m_trainer->params_inference(), n_params);
			uint32_t c[2]; CUDA_CHECK_THROW(cudaMemcpy(&c[0], tr.counters_rgb.numsteps_counter.data(), 4, cudaMemcpyDeviceToHost));
			CUDA_CHECK_THROW(cudaMemcpy(&c[1], tr.counters_rgb.numsteps_counter_compacted.data(), 4, cudaMemcpyDeviceToHost));
			float sc[3] = {tb.m_loss_scalar.val(), tb.m_ek_loss_scalar.val(), tb.m_mask_loss_scalar.val()};
			uint64_t o[8] = {R, c[0], c[1], tr.counters_rgb.rays_per_batch, tr.counters_rgb.measured_batch_size, tr.counters_rgb.measured_batch_size_before_compaction, 0, 0};
			memcpy(&o[6], sc, 12);
			dump_host(tag + "_out_counters.bin", o, sizeof(o));
		}
		if (k % print_every == 0 || k == n_steps - 1)
			printf("ref step %d rays %u samples %u compacted %u loss %g  %.3f ms\n", k, R, tr.counters_rgb.measured_batch_size_before_compaction, tr.counters_rgb.measured_batch_size, tb.m_loss_scalar.val(), ms);
	}
	if (!time_only) dump_state("final");
	fprintf(meta, "timed_steps=%d\ntimed_ms=%.4f\ntimed_rays=%llu\nrays_per_second=%.1f\n", timed_steps, total_ms, (unsigned long long)total_rays, total_ms > 0 ? total_rays / (total_ms * 1e-3) : 0.0);

	// network probes on the final parameters: training weights (use_inference_params=false), positions on a fixed lattice
	if (!time_only) {
		const uint32_t n = 4096;
		std::vector<float> coords(n * 7);
		pcg32 rng(42);
		for (uint32_t i = 0; i < n; ++i) {
			coords[i * 7 + 0] = 0.25f + 0.5f * rng.next_float(); coords[i * 7 + 1] = 0.25f + 0.5f * rng.next_float(); coords[i * 7 + 2] = 0.25f + 0.5f * rng.next_float();
			coords[i * 7 + 3] = 0.f; coords[i * 7 + 4] = 0.5f; coords[i * 7 + 5] = 0.25f; coords[i * 7 + 6] = 0.75f;
		}
		GPUMemory<float> dc(n * 7); dc.copy_from_host(coords);
		GPUMemory<precision_t> out(n * 16);
		GPUMatrix<float> in_m(dc.data(), 7, n);
		GPUMatrix<precision_t> out_m(out.data(), 16, n);
		tb.m_network->inference_mixed_precision(tb.m_training_stream, in_m, out_m, false);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		dump_host("probe_coords.bin", coords.data(), coords.size() * 4);
		dump_dev("probe_out_fp16.bin", (const uint16_t*)out.data(), (size_t)n * 16);
	}
	if (!save_snapshot.empty()) {          // src/main.cu:465-468
		tb.save_snapshot(save_snapshot, false);
		dump_dev("snapshot_params_inference_fp16.bin", (const uint16_t*)tb.m_trainer->params_inference(), n_params);
		dump_dev("snapshot_density_grid.bin", tb.m_nerf.density_grid.data(), tb.m_nerf.density_grid.size());
	}
	// mesh path: the SDF lattice of get_density_on_grid, then compute_and_save_marching_cubes_mesh exactly as
	// src/main.cu:460 calls it (empty aabb -> m_render_aabb, threshold 0, no unwrap), and the mesh arrays it leaves in m_mesh
	if (mesh_res > 0) {
		const BoundingBox aabb = tb.m_render_aabb;
		const uint32_t r16 = next_multiple((uint32_t)mesh_res, 16u);
		GPUMemory<float> dens = tb.get_density_on_grid(Eigen::Vector3i::Constant((int)r16), aabb);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		dump_dev("mesh_density.bin", dens.data(), dens.size());
		dump_dev("mesh_params_inference_fp16.bin", (const uint16_t*)tb.m_trainer->params_inference(), n_params);
		const std::string obj = g_out + "/ref_mesh.obj", ply = g_out + "/ref_mesh.ply";
		cudaEvent_t m0, m1; cudaEventCreate(&m0); cudaEventCreate(&m1);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		const auto w0 = std::chrono::steady_clock::now();
		tb.compute_and_save_marching_cubes_mesh(obj.c_str(), Eigen::Vector3i::Constant(mesh_res), {}, 0.0f, false);
		CUDA_CHECK_THROW(cudaDeviceSynchronize());
		const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
		dump_dev("mesh_verts.bin", (const float*)tb.m_mesh.verts.data(), tb.m_mesh.verts.size() * 3);
		dump_dev("mesh_normals.bin", (const float*)tb.m_mesh.vert_normals.data(), tb.m_mesh.vert_normals.size() * 3);
		dump_dev("mesh_colors.bin", (const float*)tb.m_mesh.vert_colors.data(), tb.m_mesh.vert_colors.size() * 3);
		dump_dev("mesh_indices.bin", tb.m_mesh.indices.data(), tb.m_mesh.indices.size());
		const auto& ds = tr.dataset;
		save_mesh(tb.m_mesh.verts, tb.m_mesh.vert_normals, tb.m_mesh.vert_colors, tb.m_mesh.indices, ply.c_str(), false, ds.scale, ds.offset, ds.n2w_s, ds.n2w_t, ds.from_na);
		fprintf(meta, "mesh_res=%u\nmesh_aabb_min=%g %g %g\nmesh_aabb_max=%g %g %g\nmesh_verts=%zu\nmesh_indices=%zu\nmesh_seconds=%.4f\n", r16, aabb.min.x(), aabb.min.y(), aabb.min.z(),
		        aabb.max.x(), aabb.max.y(), aabb.max.z(), tb.m_mesh.verts.size(), tb.m_mesh.indices.size(), wall);
		fprintf(meta, "dataset_scale=%.9g\ndataset_offset=%.9g %.9g %.9g\nn2w_s=%.9g\nn2w_t=%.9g %.9g %.9g\nfrom_na=%d\nmesh_training_step=%u\n", ds.scale, ds.offset.x(), ds.offset.y(), ds.offset.z(),
		        ds.n2w_s, ds.n2w_t.x(), ds.n2w_t.y(), ds.n2w_t.z(), (int)ds.from_na, (unsigned)tb.m_training_step);
	}
	fclose(meta);
	printf("ref_harness done\n");
	return 0;
}
