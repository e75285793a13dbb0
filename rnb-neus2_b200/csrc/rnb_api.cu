// rnb_api.cu — context object and C ABI (include/rnb_b200.h).  Host-side orchestration of one training step:
// the B200-native counterpart of Testbed::train / training_prep_nerf / train_nerf / train_nerf_step
// (reference src/testbed.cu:2776-2872, src/testbed_nerf.cu:3424-3668,3844-4138) and Trainer::optimizer_step.
//
// Differences in structure (not in results): no per-step host synchronisation (sample counts stay on the device and
// kernels read them there; the reference copies counters to the host and reduces the variance gradient through the
// host every step), one scratch arena allocated up front, ordered scans instead of atomicAdd slot hand-out.
#include "rnb_common.cuh"
#include <vector>
#include <string>
#include <random>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges are no-ops unless a profiler (nsys / ncu --nvtx) is attached
#include <nccl.h>      // types and prototypes only: libnccl.so.2 is opened at run time by rnb_comm_* (single-GPU users never need it)

namespace rnb {
// rnb_march.cu
void launch_march(cudaStream_t, uint32_t, uint32_t, uint32_t, uint32_t, Pcg32, const ViewDev*, uint32_t, const uint8_t*, uint32_t*, float*, float*, uint32_t = 0);
void launch_scan_rays(cudaStream_t, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t*, uint32_t = 1, uint32_t = 0, const uint32_t* = nullptr, uint32_t* = nullptr);
void launch_prefix_positions(cudaStream_t, uint32_t, uint32_t, uint32_t, const uint32_t*, const uint32_t*, uint32_t*);
void launch_emit(cudaStream_t, uint32_t, const uint32_t*, uint32_t, const uint32_t*, const uint32_t*, const float*, const float*, float4*, float* = nullptr);
// rnb_network_simt.cu
void launch_forward_simt(cudaStream_t, const ModelDev&, const __half*, uint32_t, int, const float4*, const uint32_t*, uint32_t, const float*, __half*, float*, float*, float*);
void launch_backward_simt(cudaStream_t, const ModelDev&, const __half*, uint32_t, const float4*, const __half*, const uint32_t*, uint32_t, uint32_t, uint32_t, const uint32_t*, const uint32_t*, float*, __half*, float*);
size_t backward_simt_scratch_halfs(const ModelDev&);
// rnb_network_mma.cu
bool mma_supported(const ModelDev&);
size_t mma_pack_u32(const ModelDev&);
void launch_mma(int, cudaStream_t, const ModelDev&, const __half*, uint32_t*, uint32_t, const float4*, const uint32_t*, uint32_t, const float*, __half*, const __half*, uint32_t, uint32_t, const uint32_t*, float*, int);
bool tc_supported(const ModelDev&);
void set_bw_debug(int);
void set_bw_scatter_groups(int);
void launch_tc_sdf_grid(cudaStream_t, const ModelDev&, const __half*, const uint8_t*, uint32_t, const uint32_t[3], const float[3], const float[3], float*, int);
void launch_tc_backward(cudaStream_t, const ModelDev&, const __half*, const uint8_t*, uint32_t, const float4*, const __half*, const uint32_t*, uint32_t, uint32_t, uint32_t, const uint32_t*, float*, int, const uint32_t* = nullptr);
size_t tc_blob_bytes(const ModelDev&);
void launch_tc(int, cudaStream_t, const ModelDev&, const __half*, uint8_t*, uint32_t, const float4*, const uint32_t*, uint32_t, __half*, float*, float*, int, const float* = nullptr);
// rnb_loss.cu
void launch_ray_dirw(cudaStream_t, uint32_t, const uint32_t*, const uint32_t*, const float*, float*);
void launch_compact_count(cudaStream_t, uint32_t, const uint32_t*, const uint32_t*, const __half*, const float*, const __half*, uint32_t, float, uint32_t*);
void launch_scan_compact(cudaStream_t, uint32_t*, uint32_t, const uint32_t*, uint32_t*, uint32_t*, float*, const uint32_t* = nullptr, uint32_t = 0, uint32_t = 1, uint32_t = 0, const uint32_t* = nullptr, uint32_t* = nullptr);
void launch_gather_compacted(cudaStream_t, uint32_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, const float4*, float4*);
void launch_loss(cudaStream_t, uint32_t, const rnb_flags&, uint32_t, uint32_t, uint32_t, float, const uint32_t*, Pcg32, const ViewDev*, uint32_t, const uint32_t*, const float*,
                 const uint32_t*, const uint32_t*, const uint32_t*, const __half*, __half*, float*, float*);
// rnb_optim.cu
struct AdamParams {
	float base_lr, beta1, beta2, eps, l2, loss_scale, ema_decay, ema_debias_old, ema_debias_new;
	uint32_t n_params, n_matrix, rgb_begin, rgb_end; int only_sdf; float log2_beta1, log2_beta2;
	uint32_t shard_begin, shard_end; const float* gsrc;      // data-parallel optimizer shard (rnb_optim.cu)
	const __half* gsrc16;                                   // binary16 gradient exchange (rnb_optim.cu)
	uint32_t first, last;                                   // parameter range of this launch
};
void launch_pack_grads(cudaStream_t, uint32_t, float*, __half*);
void launch_adam_ema(cudaStream_t, const AdamParams&, float*, __half*, __half*, float*, float*, float*, uint32_t*);
void launch_cast_params(cudaStream_t, uint32_t, const float*, __half*);
void launch_widen_params(cudaStream_t, uint32_t, const __half*, float*);
void launch_init_grid(cudaStream_t, Pcg32, uint64_t, float*);
void launch_grid_samples(cudaStream_t, uint32_t, Pcg32, uint32_t, const float*, float4*, uint32_t*, float);
void launch_grid_finish(cudaStream_t, uint32_t, const uint32_t*, const float*, float, float*, float*, double*, float*, uint8_t*);
// rnb_dataset.cu
std::string load_png_rgba16_host(const char*, uint32_t*, uint32_t*, uint16_t**);
std::string load_images_to_device(cudaStream_t, uint32_t, const char* const*, uint32_t, void**, void**, uint32_t*, void**, size_t*);
// rnb_mesh.cu
std::string mesh_extract(cudaStream_t, const float*, const uint32_t[3], const float[3], const float[3], float, void**, size_t*, float**, float**, uint32_t**, uint32_t*, uint32_t*, uint32_t*, float[2], uint64_t*);
void launch_mesh_color_inputs(cudaStream_t, uint32_t, const float*, float4*, float*);
void launch_mesh_colors(cudaStream_t, uint32_t, const __half*, float*);
std::string mesh_write(cudaStream_t, const float*, const float*, const float*, const uint32_t*, uint32_t, uint32_t, const char*, float, const float[3], float, const float[3], int, uint64_t*, uint64_t*);
}

using namespace rnb;

#ifndef RNB_PRELAUNCH_AT_DEFAULT
#define RNB_PRELAUNCH_AT_DEFAULT 3      /* A/B on B200 (profiles/r01_ab_prelaunch_at.txt): step 0.822 (0) / 0.847 (1) / 0.832 (2) / 0.809 ms (3) */
#endif
#ifndef RNB_SCATTER_AGG_DEFAULT
#define RNB_SCATTER_AGG_DEFAULT 5u      /* A/B on B200 (profiles/r01_ab_scatter_agg.txt): backward 0.290 -> 0.276 ms at 5 levels, 0.286 at 8 */
#endif
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
// every multi-statement entry point is a function-try-block: no C++ exception (std::bad_alloc from a scratch vector, std::length_error)
// crosses the C ABI; it becomes a status code with the message in rnb_last_error()
#define RNB_API_CATCH catch (const std::bad_alloc&) { return fail(RNB_ERR_NOMEM, "out of host memory"); } \
                      catch (const std::exception& ex_) { return fail(RNB_ERR_INVALID, std::string("internal error: ") + ex_.what()); }
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(RNB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// ---- NCCL, opened at run time ---------------------------------------------------------------------------------------------
// The single-GPU library has no link-time dependency on NCCL.  rnb_comm_* dlopen libnccl.so.2: inside a process that has already loaded
// one (PyTorch bundles its own) the loader hands back that copy, otherwise the system library.
struct NcclApi {
	void* handle = nullptr;
	decltype(&ncclGetUniqueId) GetUniqueId = nullptr; decltype(&ncclCommInitRank) CommInitRank = nullptr; decltype(&ncclCommDestroy) CommDestroy = nullptr;
	decltype(&ncclAllReduce) AllReduce = nullptr; decltype(&ncclReduceScatter) ReduceScatter = nullptr; decltype(&ncclAllGather) AllGather = nullptr;
	decltype(&ncclGroupStart) GroupStart = nullptr; decltype(&ncclGroupEnd) GroupEnd = nullptr; decltype(&ncclGetErrorString) GetErrorString = nullptr;
	decltype(&ncclCommCount) CommCount = nullptr; decltype(&ncclCommUserRank) CommUserRank = nullptr; decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi* nccl_api() {
	static NcclApi api; static bool tried = false;
	if (tried) return api.handle ? &api : nullptr;
	tried = true;
	const char* names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
	if (!api.handle) return nullptr;
	bool ok = true;
	auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) ok = false; return p; };
	api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId"); api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
	api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy"); api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
	api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter"); api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
	api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart"); api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
	api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString"); api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
	api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank"); api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
	if (!ok) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
	return &api;
}
#define NC(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return fail(RNB_ERR_CUDA, std::string(#x) + ": " + N->GetErrorString(r_)); } while (0)

struct rnb_ctx {
	rnb_config cfg; rnb_flags flags; ModelDev M;
	uint32_t off_sdf = 0, off_rgb = 0;
	// parameters / optimizer state
	float* master = nullptr; __half* params = nullptr; __half* ema = nullptr; float* grads = nullptr; float* m1 = nullptr; float* m2 = nullptr; uint32_t* steps = nullptr;
	uint32_t opt_step = 0; float lr_factor = 1.f;
	size_t np_padded = 0;                              // per-parameter arrays are allocated (and zeroed) up to a multiple of 512 elements: equal shards for any power-of-two world
	uint32_t shard_begin = 0, shard_end = 0; const float* shard_grads = nullptr;      // rnb_set_optimizer_shard; end == 0: the whole range
	// occupancy
	float* density_grid = nullptr; float* density_tmp = nullptr; uint8_t* bitfield = nullptr; double* mean_acc = nullptr; float* mean = nullptr;
	float4* gpos = nullptr; uint32_t* gidx = nullptr; float* gdens = nullptr;
	uint32_t density_ema_step = 0;
	// dataset
	ViewDev* views_dev = nullptr; uint32_t n_views = 0; std::vector<void*> owned;
	// step scratch
	uint32_t cap_rays = 0, max_samples = 0, cap_compact = 0;
	uint32_t *ray_n = nullptr, *ray_indices = nullptr, *numsteps = nullptr, *counters = nullptr, *n_fwd = nullptr, *cbase = nullptr, *n_emit = nullptr;
	float *ray_geom = nullptr, *ts = nullptr, *ray_dirw = nullptr, *loss_out = nullptr, *stats = nullptr;
	float4 *pos4 = nullptr, *cpos4 = nullptr;
	__half *outA = nullptr, *out16 = nullptr, *dout16 = nullptr, *bw_scratch = nullptr; float* bw_front = nullptr;
	uint32_t* counters_host = nullptr; float* stats_host = nullptr;   // pinned
	cudaEvent_t ev_counters = nullptr; bool counters_pending = false; bool async_end = true;      // RNB_ASYNC_END=0: wait for every step even without stats (A/B)  // asynchronous read-back of the step counters (rnb_train_step_end without stats)
	// data parallelism behind the boundary (rnb_comm_*): one NCCL communicator per context, binary16 gradient exchange buffer
	ncclComm_t comm = nullptr; bool comm_owned = false; __half* grads16 = nullptr; int dp_sharded = 0;
	// one sample order over all ranks (dp_exact; RNB_DP_EXACT=0 selects per-rank clamp / truncation / roll-over): per step two all-gathers of a per-ray-position
	// prefix table (marched samples before scan/emit, compacted samples before the truncation) give every ray its place in the batch a single GPU would have built
	bool dp_exact = false; uint32_t *dp_xg = nullptr, *slot_of_pos = nullptr, *goff = nullptr;
	bool ema_stale = false;      // sharded optimizer: the EMA copy of the other ranks' shards is out of date until rnb_comm_sync_ema
	const __half* xch16 = nullptr; uint32_t xch_begin = 0, xch_end = 0;      // result of this step's gradient exchange, consumed by rnb_train_step_end
	// chunked all-reduce on a communication stream, pipelined with Adam / EMA on the caller's stream (RNB_DP_CHUNKS; default 1 = one all-reduce on the caller's stream: at N = 2 four chunks cost 0.898 ms per step against 0.835, profiles/r02_dp_n2.txt)
	static constexpr uint32_t MAX_CHUNKS = 16;
	cudaStream_t comm_stream = nullptr; cudaEvent_t ev_pack = nullptr, ev_chunk[MAX_CHUNKS] = {}; uint32_t dp_chunks = 1, xch_chunks = 0, xch_chunk_elems = 0;
	// training state
	Pcg32 rng, density_rng;
	uint32_t training_step = 0, rays_per_batch = 4096, n_rays_total = 0, measured_before = 0, measured = 0;
	// Testbed::m_canonical_training_step (testbed.h:907): drives the occupancy-refresh cadence and mode (src/testbed.cu:2805-2806, testbed_nerf.cu:4133)
	// and the n_rays_total reset (:3906); it equals the training step after every step (:3646) but is 0 after load_snapshot (reset_network, src/testbed.cu:2451).
	// n_images_prev = Training::n_images_for_training_prev (testbed.h:578): a refresh that sees another image count starts from an empty grid (:3446-3452);
	// ~0u = "adopt the current dataset" (a context that continues a run whose state was imported piecewise)
	uint32_t canonical_step = 0, n_images_prev = ~0u;
	uint32_t step_R = 0, step_nrt = 0; bool in_step = false;
	// software pipelining of the ray march (it reads only the bitfield, the dataset and the rng, never the parameters): with a
	// pinned batch size the march of step N+1 is launched on a side stream when the backward of step N has finished, so that it
	// shares the SMs with the (HBM-bound) optimizer / the gradient all-reduce instead of running alone
	cudaStream_t side = nullptr; cudaEvent_t ev_bwd = nullptr, ev_march = nullptr;
	// in-memory checkpoint (rnb_checkpoint_save / _restore): one device-side slot of everything a step reads and writes
	struct Ckpt { void* buf = nullptr; size_t bytes = 0; bool valid = false; uint32_t opt_step, density_ema_step, training_step, rays_per_batch, n_rays_total, measured_before, measured, canonical_step, n_images_prev; float lr_factor; Pcg32 rng, density_rng; } ck;
	uint32_t pre_ctas_per_sm = 0;               // size of the pre-launched march (CTAs of 256 threads per SM; 0 = one warp per ray at once); RNB_PRELAUNCH_CTAS (A/B: capping it is slower, profiles/r02_ab_async_prelaunch.txt)
	int pre_at = RNB_PRELAUNCH_AT_DEFAULT;      // where the next march may start: 0 behind the backward, 1 behind the loss, 2 behind pass A, 3 behind scan/emit (last readers of the ray buffers)
	int pre_at_env = -1;                        // RNB_PRELAUNCH_AT; unset: 3 when the host waits for every step, 2 when it runs ahead (profiles/r02_ab_async_prelaunch.txt)
	bool step_waits = true;                     // this step's rnb_train_step_end will synchronise (stats requested / adaptive controller)
	bool pre_armed = false;                     // this step will pre-launch (decided before its kernels are queued)
	bool prelaunch = true; bool pre_valid = false; uint32_t pre_R = 0, pre_nrt = 0; uint64_t pre_rng_state = 0, pre_rng_inc = 0;
	// last extracted mesh (rnb_marching_cubes*): MeshState verts / vert_normals / vert_colors / indices (testbed.h:418-447), device memory
	struct Mesh { float *verts = nullptr, *normals = nullptr, *colors = nullptr; uint32_t* indices = nullptr; uint32_t n_verts = 0, n_verts_padded = 0, n_indices = 0; float ms[4] = {0, 0, 0, 0}; } mesh;
	void* pin_stage = nullptr; size_t pin_stage_bytes = 0;      // pinned staging of the dataset loader (grow-only)
	void* mesh_ws = nullptr; size_t mesh_ws_bytes = 0; float* mesh_density = nullptr; size_t mesh_density_bytes = 0;    // grow-only scratch of the mesh path
	bool use_mma = false; uint32_t* wpack = nullptr; int n_sm = 148;
	bool use_tc = false, use_tc_bwd = false; uint8_t* wtc = nullptr;      // tcgen05 / TMEM kernels (rnb_network_tc.cu) for pass A and the SDF probes
	// instrumentation: kernel launch counter and optional per-stage CUDA-event timing (bench.py roofline)
	uint64_t launches = 0;
	bool prof = false;
	struct ProfAcc { std::string name; double ms = 0; uint64_t calls = 0; };
	struct ProfPending { int acc; cudaEvent_t e0, e1; };
	std::vector<ProfAcc> prof_acc; std::vector<ProfPending> prof_pending; std::vector<cudaEvent_t> ev_pool;
};

static void prof_begin(rnb_ctx* c, cudaStream_t st, const char* name) {
	if (!c->prof) return;
	int a = -1;
	for (size_t i = 0; i < c->prof_acc.size(); ++i) if (c->prof_acc[i].name == name) { a = (int)i; break; }
	if (a < 0) { c->prof_acc.push_back({name, 0, 0}); a = (int)c->prof_acc.size() - 1; }
	cudaEvent_t e[2];
	for (int k = 0; k < 2; ++k) { if (c->ev_pool.empty()) cudaEventCreate(&e[k]); else { e[k] = c->ev_pool.back(); c->ev_pool.pop_back(); } }
	cudaEventRecord(e[0], st);
	c->prof_pending.push_back({a, e[0], e[1]});
}
static void prof_end(rnb_ctx* c, cudaStream_t st) { if (c->prof) cudaEventRecord(c->prof_pending.back().e1, st); }
static void prof_resolve(rnb_ctx* c) {      // call after the stream has been synchronised
	for (auto& p : c->prof_pending) {
		float ms = 0; if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { c->prof_acc[p.acc].ms += ms; c->prof_acc[p.acc].calls++; }
		c->ev_pool.push_back(p.e0); c->ev_pool.push_back(p.e1);
	}
	c->prof_pending.clear();
}
// every stage of the step is an NVTX range ("rnb/<stage>") around its launches: `ncu --nvtx --nvtx-include "rnb/backward/"` and nsys timelines
// select stages by name (the reference has no tracing hooks beyond wall-clock EMAs)
struct NvtxRange { explicit NvtxRange(const char* n) { char b[64]; snprintf(b, sizeof(b), "rnb/%s", n); nvtxRangePushA(b); } ~NvtxRange() { nvtxRangePop(); } };
#define KT(name, nk, call) do { NvtxRange nvtx_(name); prof_begin(c, st, name); call; prof_end(c, st); c->launches += (nk); } while (0)

static uint32_t next_multiple(uint32_t v, uint32_t d) { return ((v + d - 1) / d) * d; }
// a pair of timing events that is released on every return path
struct EventPair { cudaEvent_t e[2] = {nullptr, nullptr}; ~EventPair() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); } cudaEvent_t& operator[](int i) { return e[i]; } };

// a pre-launched march is only usable if nothing it depends on changed: drop it (after it has drained) otherwise
static void drop_prelaunch(rnb_ctx* c) { if (c->pre_valid) { cudaStreamSynchronize(c->side); c->pre_valid = false; } }

// host mirror of the step counters.  update_after_training's rule (:3540-3545): both measured sizes are zeroed by a step without samples
static void apply_counters(rnb_ctx* c) {
	// counters[5] is the device's copy of the rule (k_scan_compact): marched samples of the step — of all ranks when they share one sample order — or 0
	const uint32_t before = c->counters_host[5], total = c->counters_host[2];
	if (before == 0 || total == 0) { c->measured_before = 0; c->measured = 0; } else { c->measured_before = before; c->measured = total; }
}
// rnb_train_step_end without stats leaves the read-back of the counters in flight: whoever needs the host copy waits for it here
static void pull_counters(rnb_ctx* c) {
	if (!c->counters_pending) return;
	cudaEventSynchronize(c->ev_counters);
	apply_counters(c);
	c->counters_pending = false;
}
// the clamp of the next step's sample budget lives on the device (counters[5]); host-side state changes are pushed to it
static cudaError_t push_measured(rnb_ctx* c) { return cudaMemcpy(c->counters + 5, &c->measured_before, 4, cudaMemcpyHostToDevice); }

// sharded optimizer: readers of the inference (EMA) parameters must not see a copy whose foreign shards are stale — fail loudly instead
#define RNB_EMA_READY(c, use_ema) do { if ((use_ema) && (c)->ema_stale) return fail(RNB_ERR_STATE, "sharded optimizer: the EMA parameters of the other ranks' shards are stale; call rnb_comm_sync_ema (collective) first"); } while (0)

static uint32_t valid_level_for_step(const rnb_ctx* c, int step) {   // grid.h:1430-1437
	if (step <= 0) return c->cfg.n_levels;
	float v = c->cfg.base_valid_level_scale * (float)c->cfg.n_levels + c->cfg.valid_level_scale * (float)std::max(0, (int)((uint32_t)step - c->cfg.base_training_step));
	return std::min(c->cfg.n_levels, (uint32_t)std::ceil(v));
}

static int ensure_ray_capacity(rnb_ctx* c, uint32_t R) {
	if (R <= c->cap_rays) return 0;
	// grow with head-room: the adaptive controller moves the batch size by a few rays per step, and a reallocation is a device synchronisation
	// (N = 8 adaptive record of session r2dp: 4.4 ms per step while the capacity followed the batch size ray by ray)
	uint32_t cap = std::min(std::max(R + R / 2, 4096u), 1u << 18);
	cudaFree(c->ray_n); cudaFree(c->ray_indices); cudaFree(c->numsteps); cudaFree(c->n_fwd); cudaFree(c->cbase); cudaFree(c->n_emit);
	cudaFree(c->ray_geom); cudaFree(c->ts); cudaFree(c->ray_dirw); cudaFree(c->loss_out);
	cudaFree(c->dp_xg); cudaFree(c->slot_of_pos); cudaFree(c->goff); c->dp_xg = c->slot_of_pos = c->goff = nullptr;
	if (c->cfg.world_size > 1) {      // two tables of world x ceil(R / world) prefixes, the slot of every ray position, the global offset of every kept ray
		CU(cudaMalloc(&c->dp_xg, (size_t)2 * (cap + c->cfg.world_size) * 4)); CU(cudaMalloc(&c->slot_of_pos, cap * 4)); CU(cudaMalloc(&c->goff, cap * 4));
	}
	CU(cudaMalloc(&c->ray_n, cap * 4)); CU(cudaMalloc(&c->ray_indices, cap * 4)); CU(cudaMalloc(&c->numsteps, cap * 8));
	CU(cudaMalloc(&c->n_fwd, cap * 4)); CU(cudaMalloc(&c->cbase, cap * 4)); CU(cudaMalloc(&c->n_emit, cap * 4));
	CU(cudaMalloc(&c->ray_geom, (size_t)cap * 9 * 4));
	const uint32_t local = (cap + c->cfg.world_size - 1) / c->cfg.world_size;
	CU(cudaMalloc(&c->ts, (size_t)local * MAX_STEPS * 4));
	CU(cudaMalloc(&c->ray_dirw, (size_t)cap * 3 * 4)); CU(cudaMalloc(&c->loss_out, (size_t)cap * 3 * 4));
	CU(cudaMemset(c->ray_n, 0, cap * 4));
	c->cap_rays = cap;
	return 0;
}

extern "C" {

const char* rnb_last_error(void) { return g_err.c_str(); }
int rnb_set_error_(int code, const char* msg) { return fail(code, msg ? msg : ""); }      // for the other translation units of the library (rnb_raymesh.cu)
uint32_t rnb_abi_version(void) { return RNB_ABI_VERSION; }

void rnb_default_config(rnb_config* c) {
	memset(c, 0, sizeof(*c));
	c->abi_version = RNB_ABI_VERSION;
	c->n_levels = 14; c->log2_hashmap_size = 19; c->base_resolution = 16; c->per_level_scale = 0.f; c->top_resolution = 2048.f;
	c->base_valid_level_scale = 0.2f; c->valid_level_scale = 0.02f; c->base_training_step = 100;
	c->sdf_n_neurons = 64; c->sdf_n_hidden_layers = 1; c->rgb_n_neurons = 64; c->rgb_n_hidden_layers = 2; c->sdf_bias = -0.1f;
	c->learning_rate = 1e-3f; c->beta1 = 0.9f; c->beta2 = 0.99f; c->epsilon = 1e-15f; c->l2_reg = 1e-6f; c->ema_decay = 0.95f;
	c->lr_decay_start = 20000; c->lr_decay_interval = 10000; c->lr_decay_base = 0.33f; c->loss_scale = 128.f;
	c->target_batch_size = 1u << 18; c->rays_per_batch = 4096; c->pin_rays_per_batch = 1; c->seed = 1337; c->density_grid_decay = 0.95f;
	c->world_size = 1; c->rank = 0;
}
void rnb_default_flags(rnb_flags* f) {
	memset(f, 0, sizeof(*f));
	f->apply_L2 = 1; f->apply_rgbplus = 1; f->no_albedo = 1; f->mask_loss_weight = 1.0f; f->ek_loss_weight = 0.01f; f->cos_anneal_ratio = 1.0f; f->light_mode = -1;
}

int rnb_create(const rnb_config* cfg, rnb_ctx** out) try {
	if (!cfg || !out) return fail(RNB_ERR_INVALID, "null argument");
	if (cfg->abi_version != RNB_ABI_VERSION) return fail(RNB_ERR_INVALID, "ABI version mismatch");
	if (cfg->n_levels < 1 || cfg->n_levels > 14) return fail(RNB_ERR_INVALID, "n_levels must be in 1..14 (SDF-MLP input width must be 32 or 48, nerf_network.h:594-604)");
	if (cfg->sdf_n_hidden_layers != 1) return fail(RNB_ERR_INVALID, "sdf network: exactly 1 hidden layer is supported (geometric initialisation file covers 1 layer)");
	if (cfg->rgb_n_hidden_layers < 1 || cfg->rgb_n_hidden_layers > 2) return fail(RNB_ERR_INVALID, "rgb network: 1 or 2 hidden layers");
	if ((cfg->sdf_n_neurons != 32 && cfg->sdf_n_neurons != 64) || (cfg->rgb_n_neurons != 32 && cfg->rgb_n_neurons != 64)) return fail(RNB_ERR_INVALID, "n_neurons must be 32 or 64");
	if (cfg->world_size < 1 || cfg->rank >= cfg->world_size) return fail(RNB_ERR_INVALID, "bad rank / world_size");
	{   // values that would otherwise reach the device as shifts past the word size, zero divisors or NaN resolutions
		auto pos = [](float v) { return std::isfinite(v) && v > 0.f; };
		auto unit = [](float v) { return std::isfinite(v) && v >= 0.f && v < 1.f; };
		if (cfg->log2_hashmap_size < 4 || cfg->log2_hashmap_size > 24) return fail(RNB_ERR_INVALID, "log2_hashmap_size must be in 4..24");
		if (cfg->base_resolution < 2 || cfg->base_resolution > 4096) return fail(RNB_ERR_INVALID, "base_resolution must be in 2..4096");
		if (cfg->per_level_scale > 0.f ? !(std::isfinite(cfg->per_level_scale) && cfg->per_level_scale <= 16.f)
		                               : (cfg->n_levels > 1 && !(pos(cfg->top_resolution) && cfg->top_resolution >= (float)cfg->base_resolution && cfg->top_resolution <= 65536.f)))
			return fail(RNB_ERR_INVALID, "per_level_scale must be in (0, 16], or top_resolution in [base_resolution, 65536] to derive it");
		if (std::isnan(cfg->per_level_scale)) return fail(RNB_ERR_INVALID, "per_level_scale is NaN");
		if (cfg->target_batch_size < cfg->world_size || cfg->target_batch_size > (1u << 24)) return fail(RNB_ERR_INVALID, "target_batch_size must be in world_size..2^24");
		if (cfg->rays_per_batch < 1 || cfg->rays_per_batch > (1u << 18)) return fail(RNB_ERR_INVALID, "rays_per_batch must be in 1..2^18 (the reference's cap, testbed_nerf.cu:3555)");
		if (!pos(cfg->loss_scale)) return fail(RNB_ERR_INVALID, "loss_scale must be positive");
		if (!(std::isfinite(cfg->learning_rate) && cfg->learning_rate >= 0.f) || !unit(cfg->beta1) || !unit(cfg->beta2) || !pos(cfg->epsilon) || !(std::isfinite(cfg->l2_reg) && cfg->l2_reg >= 0.f))
			return fail(RNB_ERR_INVALID, "optimizer: learning_rate >= 0, beta1 / beta2 in [0, 1), epsilon > 0, l2_reg >= 0");
		if (!unit(cfg->ema_decay) || (cfg->lr_decay_interval && !pos(cfg->lr_decay_base))) return fail(RNB_ERR_INVALID, "optimizer: ema_decay in [0, 1), lr_decay_base > 0");
		if (!(pos(cfg->density_grid_decay) && cfg->density_grid_decay <= 1.f)) return fail(RNB_ERR_INVALID, "density_grid_decay must be in (0, 1]");
		if (!std::isfinite(cfg->sdf_bias) || !std::isfinite(cfg->base_valid_level_scale) || !std::isfinite(cfg->valid_level_scale)) return fail(RNB_ERR_INVALID, "non-finite sdf_bias / valid_level_scale");
	}
	int dev = 0; CU(cudaGetDevice(&dev));
	rnb_ctx* c = new rnb_ctx();
	struct Guard { rnb_ctx* p; ~Guard() { if (p) rnb_destroy(p); } } guard{c};      // any early return below releases what has been allocated so far
	c->cfg = *cfg; rnb_default_flags(&c->flags);
	if (c->cfg.per_level_scale <= 0.f)
		c->cfg.per_level_scale = c->cfg.n_levels > 1 ? std::exp(std::log(c->cfg.top_resolution * 1.0f / (float)c->cfg.base_resolution) / (c->cfg.n_levels - 1)) : 1.0f;   // src/testbed.cu:2321
	ModelDev& M = c->M; memset(&M, 0, sizeof(M));
	M.n_levels = cfg->n_levels; M.n_enc = 2 * cfg->n_levels; M.sdf_width = cfg->sdf_n_neurons; M.rgb_width = cfg->rgb_n_neurons; M.sdf_bias = cfg->sdf_bias;
	uint32_t offset = 0;
	for (uint32_t i = 0; i < M.n_levels; ++i) {      // grid.h:977-1013
		const float s = exp2f(i * std::log2(c->cfg.per_level_scale)) * c->cfg.base_resolution - 1.0f;
		const uint32_t r = (uint32_t)(ceilf(s)) + 1;
		M.scale[i] = (float)(r - 1); M.res[i] = r;
		uint32_t max_params = 0xFFFFFFFFu / 2;
		uint32_t pil = std::pow((float)r, 3.f) > (float)max_params ? max_params : r * r * r;
		pil = next_multiple(pil, 8u);
		pil = std::min(pil, 1u << cfg->log2_hashmap_size);
		M.offsets[i] = offset; offset += pil;
		if ((uint64_t)r * r * r > (uint64_t)pil) M.hashed_mask |= 1u << i;
	}
	M.offsets[M.n_levels] = offset;
	M.sdf_in = next_multiple(3 + M.n_enc, 16u); M.rgb_in = next_multiple(3 + 3 + 16 + 16, 16u);
	uint32_t off = 0; c->off_sdf = 0;
	auto add = [&](LayerDesc* ls, uint32_t& nl, uint32_t in, uint32_t width, uint32_t hidden) {
		nl = 0; ls[nl++] = {width, in, off}; off += width * in;
		for (uint32_t h = 1; h < hidden; ++h) { ls[nl++] = {width, width, off}; off += width * width; }
		ls[nl++] = {16, width, off}; off += 16 * width;
	};
	add(M.sdf_layers, M.n_sdf_layers, M.sdf_in, M.sdf_width, cfg->sdf_n_hidden_layers);
	c->off_rgb = off;
	add(M.rgb_layers, M.n_rgb_layers, M.rgb_in, M.rgb_width, cfg->rgb_n_hidden_layers);
	M.off_grid = off; off += offset * 2; M.off_var = off; off += 4; M.n_params = off;
	// paired 16-byte gradient atomics need every level's first entry on a 16-byte boundary of the fp32 gradient buffer
	M.scatter_pair = (M.off_grid % 4 == 0) ? 1u : 0u;
	for (uint32_t i = 0; i < M.n_levels; ++i) if (M.offsets[i] % 2) M.scatter_pair = 0;
	if (const char* e = getenv("RNB_SCATTER_PAIR")) M.scatter_pair = M.scatter_pair && atoi(e) != 0;
	// warp-aggregated scatter on the coarse levels (adjacent lanes = adjacent samples of a ray share the cell); RNB_SCATTER_AGG=n overrides
	M.scatter_agg = RNB_SCATTER_AGG_DEFAULT;
	if (const char* e = getenv("RNB_SCATTER_AGG")) M.scatter_agg = (uint32_t)std::max(0, atoi(e));
	M.scatter_agg = std::min(M.scatter_agg, M.n_levels);
	for (uint32_t i = 0; i < M.scatter_agg; ++i) if (M.res[i] > 1024u) { M.scatter_agg = i; break; }      // the 10-bit cell key
	const size_t np = ((size_t)M.n_params + 511) / 512 * 512;      // padded capacity (the tail stays zero): collectives over equal shards
	c->np_padded = np;
	CU(cudaMalloc(&c->master, np * 4)); CU(cudaMalloc(&c->params, np * 2)); CU(cudaMalloc(&c->ema, np * 2)); CU(cudaMalloc(&c->grads, np * 4));
	CU(cudaMalloc(&c->m1, np * 4)); CU(cudaMalloc(&c->m2, np * 4)); CU(cudaMalloc(&c->steps, np * 4));
	CU(cudaMemset(c->master, 0, np * 4)); CU(cudaMemset(c->params, 0, np * 2)); CU(cudaMemset(c->ema, 0, np * 2)); CU(cudaMemset(c->grads, 0, np * 4));
	CU(cudaMemset(c->m1, 0, np * 4)); CU(cudaMemset(c->m2, 0, np * 4)); CU(cudaMemset(c->steps, 0, np * 4));
	CU(cudaMalloc(&c->density_grid, GRID_CELLS * 4)); CU(cudaMalloc(&c->density_tmp, GRID_CELLS * 4)); CU(cudaMalloc(&c->bitfield, GRID_CELLS));
	CU(cudaMalloc(&c->mean_acc, 8)); CU(cudaMalloc(&c->mean, 4));
	CU(cudaMemset(c->density_grid, 0, GRID_CELLS * 4)); CU(cudaMemset(c->density_tmp, 0, GRID_CELLS * 4)); CU(cudaMemset(c->bitfield, 0, GRID_CELLS)); CU(cudaMemset(c->mean, 0, 4));
	CU(cudaMalloc(&c->gpos, (size_t)GRID_CELLS * 16)); CU(cudaMalloc(&c->gidx, (size_t)GRID_CELLS * 4)); CU(cudaMalloc(&c->gdens, (size_t)GRID_CELLS * 4));
	c->max_samples = cfg->target_batch_size * 16;                 // testbed_nerf.cu:3846
	c->cap_compact = cfg->target_batch_size + MAX_STEPS;          // the straddling ray is forwarded in full
	CU(cudaMalloc(&c->pos4, (size_t)c->max_samples * 16)); CU(cudaMalloc(&c->outA, (size_t)c->max_samples * 8));
	CU(cudaMalloc(&c->cpos4, (size_t)c->cap_compact * 16)); CU(cudaMalloc(&c->out16, (size_t)c->cap_compact * 32)); CU(cudaMalloc(&c->dout16, (size_t)c->cap_compact * 32));
	CU(cudaMemset(c->dout16, 0, (size_t)c->cap_compact * 32));
	CU(cudaMalloc(&c->bw_scratch, (size_t)cfg->target_batch_size * backward_simt_scratch_halfs(M) * 2)); CU(cudaMalloc(&c->bw_front, (size_t)cfg->target_batch_size * M.sdf_width * 4));
	CU(cudaMalloc(&c->counters, 16 * 4)); CU(cudaMemset(c->counters, 0, 16 * 4));
	CU(cudaMalloc(&c->stats, 8 * 4)); CU(cudaMemset(c->stats, 0, 8 * 4));
	CU(cudaMallocHost(&c->counters_host, 16 * 4)); CU(cudaMallocHost(&c->stats_host, 8 * 4));
	CU(cudaEventCreateWithFlags(&c->ev_counters, cudaEventDisableTiming));
	{   // the side stream has the LOWEST priority: when the caller trains on a stream of higher priority, the pre-launched march only takes what the step's own kernels leave free
		int prio_lo = 0, prio_hi = 0; CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CU(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_lo));
	}
	CU(cudaEventCreateWithFlags(&c->ev_bwd, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->ev_march, cudaEventDisableTiming));
	{
		cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, dev)); c->n_sm = prop.multiProcessorCount;
		const char* e = getenv("RNB_NETWORK");        // "simt" selects the CUDA-core kernels (cross-check path)
		c->use_mma = mma_supported(M) && !(e && std::string(e) == "simt");
		CU(cudaMalloc(&c->wpack, mma_pack_u32(M) * 4)); CU(cudaMemset(c->wpack, 0, mma_pack_u32(M) * 4));
		// RNB_NETWORK=mma keeps the mma.sync tile kernels everywhere; default: tcgen05 kernels where they exist
		if (const char* d = getenv("RNB_BW_DEBUG")) set_bw_debug(atoi(d));
		if (const char* d = getenv("RNB_BW_SCATTER_WG")) set_bw_scatter_groups(atoi(d));      // scatter warpgroups of the tcgen05 backward (1 default, 2)
		if (const char* d = getenv("RNB_PRELAUNCH")) c->prelaunch = atoi(d) != 0;
		if (const char* d = getenv("RNB_ASYNC_END")) c->async_end = atoi(d) != 0;
		if (const char* d = getenv("RNB_PRELAUNCH_CTAS")) c->pre_ctas_per_sm = (uint32_t)std::max(0, atoi(d));
		if (const char* d = getenv("RNB_PRELAUNCH_AT")) c->pre_at_env = std::min(std::max(atoi(d), 0), 3);
		c->use_tc = c->use_mma && tc_supported(M) && !(e && std::string(e) == "mma");
		c->use_tc_bwd = c->use_tc && !(getenv("RNB_BACKWARD") && std::string(getenv("RNB_BACKWARD")) == "mma");     // RNB_BACKWARD=mma: mma.sync backward
		CU(cudaMalloc(&c->wtc, tc_blob_bytes(M))); CU(cudaMemset(c->wtc, 0, tc_blob_bytes(M)));
	}
	c->rays_per_batch = cfg->rays_per_batch;
	int rc = ensure_ray_capacity(c, c->rays_per_batch); if (rc) return rc;
	c->rng = Pcg32(cfg->seed);
	{ Pcg32 t = c->rng; c->density_rng = Pcg32(t.next_uint()); }     // src/testbed.cu:2223,2236,2490
	guard.p = nullptr;
	*out = c;
	return RNB_OK;
} RNB_API_CATCH

int rnb_destroy(rnb_ctx* c) try {
	if (!c) return RNB_OK;
	void* ptrs[] = {c->master, c->params, c->ema, c->grads, c->m1, c->m2, c->steps, c->density_grid, c->density_tmp, c->bitfield, c->mean_acc, c->mean, c->gpos, c->gidx, c->gdens,
	                c->views_dev, c->ray_n, c->ray_indices, c->numsteps, c->counters, c->n_fwd, c->cbase, c->n_emit, c->ray_geom, c->ts, c->ray_dirw, c->loss_out, c->stats,
	                c->pos4, c->cpos4, c->outA, c->out16, c->dout16, c->bw_scratch, c->bw_front, c->wpack, c->wtc};
	for (void* p : ptrs) cudaFree(p);
	for (void* p : c->owned) cudaFree(p);
	cudaFreeHost(c->counters_host); cudaFreeHost(c->stats_host);
	cudaFree(c->ck.buf);
	cudaFree(c->mesh.verts); cudaFree(c->mesh.normals); cudaFree(c->mesh.colors); cudaFree(c->mesh.indices); cudaFree(c->mesh_ws); cudaFree(c->mesh_density); cudaFreeHost(c->pin_stage);
	if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
	if (c->ev_bwd) cudaEventDestroy(c->ev_bwd);
	if (c->ev_march) cudaEventDestroy(c->ev_march);
	if (c->ev_counters) cudaEventDestroy(c->ev_counters);
	cudaFree(c->grads16); cudaFree(c->dp_xg); cudaFree(c->slot_of_pos); cudaFree(c->goff);
	if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); }
	if (c->ev_pack) cudaEventDestroy(c->ev_pack);
	for (cudaEvent_t e : c->ev_chunk) if (e) cudaEventDestroy(e);
	if (c->comm && c->comm_owned) { if (NcclApi* N = nccl_api()) N->CommDestroy(c->comm); }
	delete c;
	return RNB_OK;
} RNB_API_CATCH

int rnb_param_layout(rnb_ctx* c, uint64_t out[5]) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	out[0] = c->off_sdf; out[1] = c->off_rgb; out[2] = c->M.off_grid; out[3] = c->M.off_var; out[4] = c->M.n_params;
	return RNB_OK;
} RNB_API_CATCH

// Built-in geometric initialisation used when the reference's utils/mlp_weights*.txt is not supplied: a bias-free
// one-hidden-layer ReLU network whose output 0 approximates |x - 0.5| * k (sum of ReLUs over random directions).
static void builtin_sphere_init(const rnb_ctx* c, std::vector<float>& w) {
	const ModelDev& M = c->M;
	const uint32_t W = M.sdf_width, IN = M.sdf_in;
	w.assign((size_t)W * IN + 16 * W, 0.f);
	Pcg32 r(c->cfg.seed + 7);
	for (uint32_t i = 0; i < W; ++i) {
		float x, y, z, n;
		do { x = r.next_float() * 2 - 1; y = r.next_float() * 2 - 1; z = r.next_float() * 2 - 1; n = std::sqrt(x * x + y * y + z * z); } while (n < 1e-3f || n > 1.f);
		w[(size_t)i * IN + 0] = x / n; w[(size_t)i * IN + 1] = y / n; w[(size_t)i * IN + 2] = z / n;
		for (uint32_t k = 3; k < 3 + M.n_enc; ++k) w[(size_t)i * IN + k] = (r.next_float() * 2 - 1) * 0.05f;
	}
	// E[relu(w.x)] = |x|/4 over the unit sphere; radius 0.25 in the unit cube with the -0.1 bias: k*0.25 = 0.1
	const float k = -c->cfg.sdf_bias / 0.25f;
	for (uint32_t i = 0; i < W; ++i) w[(size_t)W * IN + i] = k * 4.0f / (float)W;
	for (uint32_t o = 1; o < 16; ++o) for (uint32_t i = 0; i < W; ++i) w[(size_t)W * IN + (size_t)o * W + i] = (r.next_float() * 2 - 1) * 0.1f;
}

int rnb_init_params(rnb_ctx* c, const float* sdf_init, size_t n_sdf_init) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	const ModelDev& M = c->M;
	std::seed_seq seq{c->cfg.seed};                 // trainer.h:54-60
	std::vector<uint32_t> seeds(2); seq.generate(seeds.begin(), seeds.end());
	Pcg32 rnd(seeds.front());
	std::vector<float> mlp(M.off_grid, 0.f);
	auto xavier = [&](const LayerDesc* ls, uint32_t nl) {   // gpu_matrix.h:292-304
		for (uint32_t l = 0; l < nl; ++l) {
			const float scale = std::sqrt(6.0f / (float)(ls[l].cols + ls[l].rows));
			for (size_t i = 0; i < (size_t)ls[l].rows * ls[l].cols; ++i) mlp[ls[l].off + i] = rnd.next_float() * 2.0f * scale - scale;
		}
	};
	xavier(M.sdf_layers, M.n_sdf_layers);
	const size_t n_sdf = c->off_rgb - c->off_sdf;
	std::vector<float> builtin;
	if (!sdf_init) { builtin_sphere_init(c, builtin); sdf_init = builtin.data(); n_sdf_init = builtin.size(); }
	for (size_t i = 0; i < n_sdf; ++i) mlp[c->off_sdf + i] = i < n_sdf_init ? sdf_init[i] : 0.f;
	xavier(M.rgb_layers, M.n_rgb_layers);
	CU(cudaMemcpy(c->master, mlp.data(), mlp.size() * 4, cudaMemcpyHostToDevice));
	const uint64_t n_grid = (uint64_t)M.off_var - M.off_grid;
	launch_init_grid(0, rnd, n_grid, c->master + M.off_grid);
	rnd.advance((int64_t)n_grid);
	const float var[4] = {0.3f, 0.3f, 0.3f, 0.3f};
	CU(cudaMemcpy(c->master + M.off_var, var, 16, cudaMemcpyHostToDevice));
	launch_cast_params(0, M.n_params, c->master, c->params);
	// before the first optimizer step the inference parameters are the training parameters (trainer.h:100-109); the EMA recurrence
	// ignores this content at its first step (debias factor 1 - decay^0 = 0, ema.h:121-122)
	CU(cudaMemcpy(c->ema, c->params, (size_t)M.n_params * 2, cudaMemcpyDeviceToDevice));
	CU(cudaMemset(c->m1, 0, (size_t)M.n_params * 4)); CU(cudaMemset(c->m2, 0, (size_t)M.n_params * 4));
	CU(cudaMemset(c->steps, 0, (size_t)M.n_params * 4)); CU(cudaMemset(c->grads, 0, (size_t)M.n_params * 4));
	c->opt_step = 0; c->lr_factor = 1.f;
	CU(cudaDeviceSynchronize());
	return RNB_OK;
} RNB_API_CATCH

int rnb_set_params_fp32(rnb_ctx* c, const float* p, size_t n) try {
	if (!c || !p || n != c->M.n_params) return fail(RNB_ERR_INVALID, "bad parameter buffer");
	CU(cudaMemcpy(c->master, p, n * 4, cudaMemcpyHostToDevice));
	launch_cast_params(0, c->M.n_params, c->master, c->params);
	CU(cudaDeviceSynchronize());
	return RNB_OK;
} RNB_API_CATCH
int rnb_get_params_fp32(rnb_ctx* c, float* p, size_t n) try {
	if (!c || !p || n != c->M.n_params) return fail(RNB_ERR_INVALID, "bad parameter buffer");
	CU(cudaMemcpy(p, c->master, n * 4, cudaMemcpyDeviceToHost));
	return RNB_OK;
} RNB_API_CATCH
int rnb_export_params_fp16(rnb_ctx* c, uint16_t* host, size_t n, int use_ema) try {
	if (!c || !host || n != c->M.n_params) return fail(RNB_ERR_INVALID, "bad parameter buffer");
	RNB_EMA_READY(c, use_ema);
	CU(cudaMemcpy(host, use_ema ? c->ema : c->params, n * 2, cudaMemcpyDeviceToHost));
	return RNB_OK;
} RNB_API_CATCH
// Trainer::deserialize (trainer.h:263-275): fp16 params in, fp32 master re-derived, optimizer moments restart
int rnb_import_params_fp16(rnb_ctx* c, const uint16_t* host, size_t n) try {
	if (!c || !host || n != c->M.n_params) return fail(RNB_ERR_INVALID, "bad parameter buffer");
	CU(cudaMemcpy(c->params, host, n * 2, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->ema, host, n * 2, cudaMemcpyHostToDevice));
	launch_widen_params(0, c->M.n_params, c->params, c->master);
	CU(cudaMemset(c->m1, 0, n * 4)); CU(cudaMemset(c->m2, 0, n * 4)); CU(cudaMemset(c->steps, 0, n * 4));
	c->opt_step = 0; c->lr_factor = 1.f;
	CU(cudaDeviceSynchronize());
	return RNB_OK;
} RNB_API_CATCH
int rnb_export_density_grid(rnb_ctx* c, float* host, size_t n, uint32_t* ema_step) try {
	if (!c || !host || n != GRID_CELLS) return fail(RNB_ERR_INVALID, "density grid is 128^3 floats");
	CU(cudaMemcpy(host, c->density_grid, n * 4, cudaMemcpyDeviceToHost));
	if (ema_step) *ema_step = c->density_ema_step;
	return RNB_OK;
} RNB_API_CATCH
int rnb_get_bitfield(rnb_ctx* c, uint8_t* host, size_t n) try {
	if (!c || !host || n != GRID_CELLS) return fail(RNB_ERR_INVALID, "bitfield is 128^3 bytes");
	CU(cudaMemcpy(host, c->bitfield, n, cudaMemcpyDeviceToHost));
	return RNB_OK;
} RNB_API_CATCH
int rnb_set_bitfield(rnb_ctx* c, const uint8_t* host, size_t n) try {
	if (!c || !host || n != GRID_CELLS) return fail(RNB_ERR_INVALID, "bitfield is 128^3 bytes");
	drop_prelaunch(c);
	CU(cudaMemcpy(c->bitfield, host, n, cudaMemcpyHostToDevice));
	return RNB_OK;
} RNB_API_CATCH
int rnb_get_train_state(rnb_ctx* c, uint32_t out[4]) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	pull_counters(c);
	out[0] = c->training_step; out[1] = c->rays_per_batch; out[2] = c->n_rays_total; out[3] = c->measured_before;
	return RNB_OK;
} RNB_API_CATCH
int rnb_set_train_state(rnb_ctx* c, uint32_t training_step, uint32_t rays_per_batch, uint32_t n_rays_total, uint32_t measured_before) try {
	if (c) drop_prelaunch(c);
	if (!c || rays_per_batch == 0 || rays_per_batch > (1u << 18)) return fail(RNB_ERR_INVALID, "bad train state");
	pull_counters(c);
	c->training_step = training_step; c->rays_per_batch = rays_per_batch; c->n_rays_total = n_rays_total; c->measured_before = measured_before;
	CU(push_measured(c));
	c->canonical_step = training_step;                 // a state in the middle of a run; rnb_set_canonical_state says otherwise (snapshot load)
	return ensure_ray_capacity(c, rays_per_batch);
} RNB_API_CATCH
// the two members Testbed::load_snapshot does NOT restore (see rnb_ctx): the reference resumes stage 2 with m_canonical_training_step == 0 and, in a
// fresh process, n_images_for_training_prev == 0, so its first Testbed::train refreshes the occupancy grid at once, in bootstrap mode, from an EMPTIED grid
int rnb_set_canonical_state(rnb_ctx* c, uint32_t canonical_training_step, uint32_t n_images_prev) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	drop_prelaunch(c);
	c->canonical_step = canonical_training_step; c->n_images_prev = n_images_prev;
	return RNB_OK;
} RNB_API_CATCH
int rnb_get_rng(rnb_ctx* c, uint64_t out[4]) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	out[0] = c->rng.state; out[1] = c->rng.inc; out[2] = c->density_rng.state; out[3] = c->density_rng.inc; return RNB_OK;
} RNB_API_CATCH
int rnb_set_rng(rnb_ctx* c, const uint64_t in[4]) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	drop_prelaunch(c);
	c->rng.state = in[0]; c->rng.inc = in[1]; c->density_rng.state = in[2]; c->density_rng.inc = in[3]; return RNB_OK;
} RNB_API_CATCH

static int set_views(rnb_ctx* c, const rnb_view* views, uint32_t n, bool upload) {
	if (!c || !views || n == 0) return fail(RNB_ERR_INVALID, "no views");
	// validate everything first, build the new allocations and the new view table in locals, and only then swap them into the context:
	// a failure leaves the previous dataset installed and intact (nothing the old ViewDev table points at has been freed)
	for (uint32_t i = 0; i < n; ++i) if (!views[i].normal_px || views[i].w <= 0 || views[i].h <= 0) return fail(RNB_ERR_INVALID, "view without normal map");
	drop_prelaunch(c);
	std::vector<void*> fresh;
	auto release = [&]() { for (void* p : fresh) cudaFree(p); fresh.clear(); };
	std::vector<ViewDev> vd(n);
	for (uint32_t i = 0; i < n; ++i) {
		const rnb_view& v = views[i];
		const size_t bytes = (size_t)v.w * v.h * 8;
		const void* np = v.normal_px; const void* ap = v.albedo_px;
		if (upload) {
			for (int k = 0; k < 2; ++k) {
				const void* src = k == 0 ? v.normal_px : v.albedo_px;
				if (!src) continue;
				void* d = nullptr;
				cudaError_t e = cudaMalloc(&d, bytes);
				if (e == cudaSuccess) { fresh.push_back(d); e = cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice); }
				if (e != cudaSuccess) { release(); return fail(e == cudaErrorMemoryAllocation ? RNB_ERR_NOMEM : RNB_ERR_CUDA, std::string("dataset upload: ") + cudaGetErrorString(e)); }
				(k == 0 ? np : ap) = d;
			}
		}
		vd[i].normal_px = (const uint2*)np; vd[i].albedo_px = (const uint2*)ap; vd[i].w = v.w; vd[i].h = v.h;
		vd[i].fx = v.fx; vd[i].fy = v.fy; vd[i].cx = v.cx; vd[i].cy = v.cy;
		memcpy(vd[i].xform, v.xform, sizeof(v.xform));
	}
	ViewDev* table = nullptr;
	cudaError_t e = cudaMalloc(&table, n * sizeof(ViewDev));
	if (e == cudaSuccess) e = cudaMemcpy(table, vd.data(), n * sizeof(ViewDev), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(table); release(); return fail(RNB_ERR_CUDA, std::string("dataset view table: ") + cudaGetErrorString(e)); }
	for (void* p : c->owned) cudaFree(p);
	cudaFree(c->views_dev);
	c->owned = std::move(fresh); c->views_dev = table; c->n_views = n;
	return RNB_OK;
}
int rnb_set_dataset(rnb_ctx* c, const rnb_view* v, uint32_t n) { return set_views(c, v, n, false); }
int rnb_upload_dataset(rnb_ctx* c, const rnb_view* v, uint32_t n) { return set_views(c, v, n, true); }

// ---- dataset ingest -----------------------------------------------------------------------------------------
// stbi_load_16(path, &w, &h, &comp, 4) as load_nerf uses it (src/nerf_loader.cu:612,653): any PNG -> 16-bit RGBA in host memory
int rnb_load_png_rgba16(const char* path, uint32_t* w, uint32_t* h, uint16_t** pixels_host) try {
	if (!path || !w || !h || !pixels_host) return fail(RNB_ERR_INVALID, "null argument");
	const std::string e = load_png_rgba16_host(path, w, h, pixels_host);
	if (!e.empty()) return fail(RNB_ERR_INVALID, e);
	return RNB_OK;
} RNB_API_CATCH
void rnb_free_host(void* p) { free(p); }

// the image half of load_nerf (src/nerf_loader.cu:556-760): decode every normal / albedo map on host threads into pinned staging
// and upload; meta[i] carries intrinsics and the camera matrix (pixel pointers ignored; w/h checked against the files when non-zero)
int rnb_load_dataset_images(rnb_ctx* c, const rnb_view* meta, uint32_t n, const char* const* normal_paths, const char* const* albedo_paths, uint32_t threads, void* stream) try {
	if (!c || !meta || !normal_paths || n == 0) return fail(RNB_ERR_INVALID, "no views");
	std::vector<const char*> paths(2 * (size_t)n, nullptr);
	for (uint32_t i = 0; i < n; ++i) {
		if (!normal_paths[i]) return fail(RNB_ERR_INVALID, "view without normal map");
		paths[i] = normal_paths[i]; paths[n + i] = albedo_paths ? albedo_paths[i] : nullptr;
	}
	std::vector<void*> dev(2 * (size_t)n, nullptr); std::vector<uint32_t> wh(4 * (size_t)n, 0);
	void* arena = nullptr;
	const std::string e = load_images_to_device((cudaStream_t)stream, 2 * n, paths.data(), threads, &arena, dev.data(), wh.data(), &c->pin_stage, &c->pin_stage_bytes);
	if (!e.empty()) return fail(RNB_ERR_INVALID, e);
	std::vector<rnb_view> v(meta, meta + n);
	std::string bad;
	for (uint32_t i = 0; i < n && bad.empty(); ++i) {
		const uint32_t w = wh[2 * i], h = wh[2 * i + 1];
		if (dev[n + i] && (wh[2 * (n + i)] != w || wh[2 * (n + i) + 1] != h)) bad = std::string("normal and albedo map differ in size: ") + normal_paths[i];
		if ((v[i].w > 0 && (uint32_t)v[i].w != w) || (v[i].h > 0 && (uint32_t)v[i].h != h)) bad = std::string("image size does not match the metadata: ") + normal_paths[i];
		v[i].normal_px = dev[i]; v[i].albedo_px = dev[n + i]; v[i].w = (int32_t)w; v[i].h = (int32_t)h;
	}
	int rc = bad.empty() ? set_views(c, v.data(), n, false) : fail(RNB_ERR_INVALID, bad);
	if (rc != RNB_OK) { cudaFree(arena); return rc; }
	c->owned.push_back(arena);                                  // the context owns the uploaded pixels (freed by the next dataset call / rnb_destroy)
	return RNB_OK;
} RNB_API_CATCH

int rnb_set_flags(rnb_ctx* c, const rnb_flags* f) { if (!c || !f) return fail(RNB_ERR_INVALID, "null argument"); c->flags = *f; return RNB_OK; }

int rnb_import_density_grid(rnb_ctx* c, const float* host, size_t n, uint32_t ema_step) try {
	if (!c || !host || n != GRID_CELLS) return fail(RNB_ERR_INVALID, "density grid is 128^3 floats");
	drop_prelaunch(c);
	CU(cudaMemcpy(c->density_grid, host, n * 4, cudaMemcpyHostToDevice));
	c->density_ema_step = ema_step;
	// update_density_grid_mean_and_bitfield (testbed_nerf.cu:3497-3517): bitfield follows the imported grid
	CU(cudaMemset(c->density_tmp, 0, GRID_CELLS * 4));
	launch_grid_finish(0, 0, c->gidx, c->gdens, 1.0f, c->density_grid, c->density_tmp, c->mean_acc, c->mean, c->bitfield);
	CU(cudaDeviceSynchronize());
	return RNB_OK;
} RNB_API_CATCH

static void net_density(rnb_ctx* c, cudaStream_t st, uint32_t vl, uint32_t n) {
	if (c->use_tc) {
		launch_tc(0, st, c->M, c->params, c->wtc, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, c->n_sm);
		launch_tc(2, st, c->M, c->params, c->wtc, vl, c->gpos, nullptr, n, nullptr, nullptr, c->gdens, c->n_sm);
	} else if (c->use_mma) {
		launch_mma(0, st, c->M, c->params, c->wpack, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, c->n_sm);
		launch_mma(4, st, c->M, c->params, c->wpack, vl, c->gpos, nullptr, n, nullptr, (__half*)c->gdens, nullptr, 0, 0, nullptr, nullptr, c->n_sm);
	} else launch_forward_simt(st, c->M, c->params, vl, 2, c->gpos, nullptr, n, nullptr, nullptr, nullptr, nullptr, c->gdens);
}

// ---- occupancy refresh: training_prep_nerf / update_density_grid_nerf --------------------------------------------
static int density_update(rnb_ctx* c, cudaStream_t st, uint32_t n_uniform, uint32_t n_nonuniform) {
	if (c->n_images_prev == ~0u) c->n_images_prev = c->n_views;
	if (c->training_step == 0 || c->n_views != c->n_images_prev) {      // testbed_nerf.cu:3446-3452
		c->n_images_prev = c->n_views;
		if (c->training_step == 0) c->density_ema_step = 0;
		CU(cudaMemsetAsync(c->density_grid, 0, GRID_CELLS * 4, st));
	}
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	const uint32_t n = n_uniform + n_nonuniform;
	KT("grid_update", 14, (launch_grid_samples(st, n_uniform, c->density_rng, c->density_ema_step, c->density_grid, c->gpos, c->gidx, -0.01f),
	    c->density_rng.advance(),
	    launch_grid_samples(st, n_nonuniform, c->density_rng, c->density_ema_step, c->density_grid, c->gpos + n_uniform, c->gidx + n_uniform, MIN_OPTICAL_THICKNESS),
	    c->density_rng.advance(),
	    net_density(c, st, vl, n),
	    launch_grid_finish(st, n, c->gidx, c->gdens, c->cfg.density_grid_decay, c->density_grid, c->density_tmp, c->mean_acc, c->mean, c->bitfield)));
	++c->density_ema_step;
	CU(cudaGetLastError());
	return RNB_OK;
}
int rnb_prep(rnb_ctx* c, void* stream) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	drop_prelaunch(c);
	if (c->canonical_step < 256) return density_update(c, (cudaStream_t)stream, GRID_CELLS, 0);      // testbed_nerf.cu:4133
	return density_update(c, (cudaStream_t)stream, GRID_CELLS / 4, GRID_CELLS / 4);
} RNB_API_CATCH

// network stage dispatch: tensor-core tile kernels (default) or the CUDA-core kernels
static void net_pack(rnb_ctx* c, cudaStream_t st, const __half* P) {
	if (c->use_mma) launch_mma(0, st, c->M, P, c->wpack, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, nullptr, nullptr, c->n_sm);
	if (c->use_tc) launch_tc(0, st, c->M, P, c->wtc, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, c->n_sm);
}
static void net_pass_a(rnb_ctx* c, cudaStream_t st, uint32_t vl, const float4* pos, const uint32_t* n_ptr, uint32_t n_max) {
	if (c->use_tc) launch_tc(1, st, c->M, c->params, c->wtc, vl, pos, n_ptr, n_max, c->outA, nullptr, nullptr, c->n_sm);
	else if (c->use_mma) launch_mma(1, st, c->M, c->params, c->wpack, vl, pos, n_ptr, n_max, nullptr, c->outA, nullptr, 0, 0, nullptr, nullptr, c->n_sm);
	else launch_forward_simt(st, c->M, c->params, vl, 0, pos, n_ptr, n_max, c->ray_dirw, c->outA, nullptr, nullptr, nullptr);
}
static void net_pass_b(rnb_ctx* c, cudaStream_t st, const __half* P, uint32_t vl, const float4* pos, const uint32_t* n_ptr, uint32_t n_max, const float* dirw) {
	if (c->use_tc && P == c->params) launch_tc(3, st, c->M, P, c->wtc, vl, pos, n_ptr, n_max, c->out16, nullptr, nullptr, c->n_sm, dirw);
	else if (c->use_mma) launch_mma(2, st, c->M, P, c->wpack, vl, pos, n_ptr, n_max, dirw, c->out16, nullptr, 0, 0, nullptr, nullptr, c->n_sm);
	else launch_forward_simt(st, c->M, P, vl, 1, pos, n_ptr, n_max, dirw, c->out16, nullptr, nullptr, nullptr);
}
static void net_backward(rnb_ctx* c, cudaStream_t st, uint32_t vl, const uint32_t* n_ptr, uint32_t n_max, uint32_t n_roll, const uint32_t* n_in_ptr, const uint32_t* goff = nullptr) {
	if (c->use_tc_bwd) launch_tc_backward(st, c->M, c->params, c->wtc, vl, c->cpos4, c->dout16, n_ptr, n_max, n_roll, c->cfg.target_batch_size, n_in_ptr, c->grads, c->n_sm, goff);
	else if (c->use_mma) launch_mma(3, st, c->M, c->params, c->wpack, vl, c->cpos4, n_ptr, n_max, nullptr, nullptr, c->dout16, n_roll, c->cfg.target_batch_size, n_in_ptr, c->grads, c->n_sm);
	else launch_backward_simt(st, c->M, c->params, vl, c->cpos4, c->dout16, n_ptr, n_max, n_roll, c->cfg.target_batch_size, n_in_ptr, nullptr, c->grads, c->bw_scratch, c->bw_front);
}

// ---- one training step ----------------------------------------------------------------------------------------------
// counters (device, uint32): [0] kept rays  [1] samples before compaction  [2] compacted (untruncated)  [3] trained = min([2], target)
//                            [4] samples forwarded in pass B  [5] samples before compaction of the previous step (0: no clamp)  [6] sample slots pass A forwards
//                            data parallel with one sample order: [1]-[4], [6] are this rank's; [3] = this rank's samples inside the global target,
//                            [8] = min(compacted samples of all ranks, target) (roll-over), [9] = samples of all ranks before compaction ([5] is taken from it)
static int step_front(rnb_ctx* c, cudaStream_t st, uint32_t R, uint32_t nrt) {
	const uint32_t max_inference = c->max_samples;      // capacity; the clamp to last step's sample count happens on the device (k_scan_rays, counters[5] -> counters[6])
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	const ModelDev& M = c->M;
	const uint32_t G = c->cfg.world_size;
	// data parallel: either every rank clamps / truncates / pads its own shard against target / world, or (exact: library communicator + tcgen05 backward) all
	// ranks share the sample order of the single-GPU batch
	const bool exact = c->dp_exact && G > 1 && c->comm && c->use_tc_bwd;
	const uint32_t local_target = exact ? c->cfg.target_batch_size : c->cfg.target_batch_size / G;
	const uint32_t L = (R + G - 1) / G;
	uint32_t* xg_march = exact ? c->dp_xg : nullptr; uint32_t* xg_comp = exact ? c->dp_xg + (size_t)G * L : nullptr;
	NcclApi* N = exact ? nccl_api() : nullptr;
	if (c->pre_valid && c->pre_R == R && c->pre_nrt == nrt && c->pre_rng_state == c->rng.state && c->pre_rng_inc == c->rng.inc) {
		CU(cudaStreamWaitEvent(st, c->ev_march, 0));            // this step's march already ran in the shadow of the previous optimizer step
		c->pre_valid = false;
	} else {
		drop_prelaunch(c);
		KT("march", 1, launch_march(st, R, G, c->cfg.rank, nrt, c->rng, c->views_dev, c->n_views, c->bitfield, c->ray_n, c->ray_geom, c->ts));
	}
	if (exact) {
		int nrc = 0;
		KT("prefix_exchange", 2, (launch_prefix_positions(st, R, G, c->cfg.rank, c->ray_n, nullptr, xg_march + (size_t)c->cfg.rank * L),
		                         nrc = (int)N->AllGather(xg_march + (size_t)c->cfg.rank * L, xg_march, L, ncclUint32, c->comm, st)));
		if (nrc) return fail(RNB_ERR_CUDA, std::string("ncclAllGather: ") + N->GetErrorString((ncclResult_t)nrc));
	}
	KT("scan_emit", 2, (launch_scan_rays(st, R, max_inference, c->counters + 5, c->ray_n, c->ray_indices, c->numsteps, c->counters, G, c->cfg.rank, xg_march, exact ? c->slot_of_pos : nullptr),
	                   launch_emit(st, (R + G - 1) / G, c->counters, G, c->ray_indices, c->numsteps, c->ray_geom, c->ts, c->pos4, c->ray_dirw)));
	if (c->pre_armed && c->pre_at == 3) CU(cudaEventRecord(c->ev_bwd, st));      // ray_n / ray_geom / ts have been consumed: the next march may overwrite them
	// weight blobs for this step's kernels: the mma.sync panel copy only when one of its kernels runs in the step (cross-check paths)
	const bool all_tc = c->use_tc && c->use_tc_bwd;
	KT("pass_a_sdf_normal", all_tc ? 2 : 3, ((all_tc ? (void)launch_tc(0, st, c->M, c->params, c->wtc, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, c->n_sm) : net_pack(c, st, c->params)),
	                                         net_pass_a(c, st, vl, c->pos4, c->counters + 6, max_inference)));
	if (c->pre_armed && c->pre_at == 2) CU(cudaEventRecord(c->ev_bwd, st));
	if (exact) {
		int nrc = 0;
		KT("compact_count", 1, launch_compact_count(st, L, c->counters, c->numsteps, c->outA, c->ray_dirw, c->params, M.off_var, c->flags.cos_anneal_ratio, c->n_fwd));
		KT("prefix_exchange", 2, (launch_prefix_positions(st, R, G, c->cfg.rank, c->n_fwd, c->slot_of_pos, xg_comp + (size_t)c->cfg.rank * L),
		                         nrc = (int)N->AllGather(xg_comp + (size_t)c->cfg.rank * L, xg_comp, L, ncclUint32, c->comm, st)));
		if (nrc) return fail(RNB_ERR_CUDA, std::string("ncclAllGather: ") + N->GetErrorString((ncclResult_t)nrc));
		KT("compact", 2, (launch_scan_compact(st, c->counters, local_target, c->n_fwd, c->cbase, c->n_emit, c->stats, xg_comp, L, G, c->cfg.rank, c->ray_indices, c->goff),
		                 launch_gather_compacted(st, L, c->counters, c->numsteps, c->n_fwd, c->cbase, c->n_emit, c->pos4, c->cpos4)));
	} else
	KT("compact", 3, (launch_compact_count(st, (R + G - 1) / G, c->counters, c->numsteps, c->outA, c->ray_dirw, c->params, M.off_var, c->flags.cos_anneal_ratio, c->n_fwd),
	                 launch_scan_compact(st, c->counters, local_target, c->n_fwd, c->cbase, c->n_emit, c->stats),
	                 launch_gather_compacted(st, (R + G - 1) / G, c->counters, c->numsteps, c->n_fwd, c->cbase, c->n_emit, c->pos4, c->cpos4)));
	KT("pass_b_forward", 1, net_pass_b(c, st, c->params, vl, c->cpos4, c->counters + 4, c->cap_compact, c->ray_dirw));
	// the per-ray loss terms are summed only when somebody will read the sums: the caller (stats), the adaptive controller's host path, or the other ranks
	const bool want_sums = c->step_waits || G > 1 || c->prof;
	KT("loss", want_sums ? 2 : 1, launch_loss(st, (R + G - 1) / G, c->flags, R, nrt, c->training_step, c->cfg.loss_scale, c->counters, c->rng, c->views_dev, c->n_views, c->ray_indices, c->ray_dirw, c->n_fwd, c->cbase, c->n_emit,
	            c->out16, c->dout16, c->loss_out, want_sums ? c->stats : nullptr));
	if (c->pre_armed && c->pre_at == 1) CU(cudaEventRecord(c->ev_bwd, st));
	if (exact) KT("backward", 1, net_backward(c, st, vl, c->counters + 3, c->cap_compact, local_target, c->counters + 8, c->goff));
	else KT("backward", c->use_mma ? 1 : 9, net_backward(c, st, vl, c->counters + 3, local_target, local_target, c->counters + 3));
	CU(cudaGetLastError());
	return RNB_OK;
}

static int optimizer_step(rnb_ctx* c, cudaStream_t st, const __half* gsrc16 = nullptr, uint32_t sh_begin = 0, uint32_t sh_end = 0) {
	if (c->opt_step == 0) c->lr_factor = 1.0f;      // exponential_decay.h:61-72
	if (c->opt_step >= c->cfg.lr_decay_start && c->cfg.lr_decay_interval && (c->opt_step - c->cfg.lr_decay_start) % c->cfg.lr_decay_interval == 0) c->lr_factor *= c->cfg.lr_decay_base;
	++c->opt_step;
	AdamParams A;
	A.base_lr = c->cfg.learning_rate * c->lr_factor; A.beta1 = c->cfg.beta1; A.beta2 = c->cfg.beta2; A.eps = c->cfg.epsilon; A.l2 = c->cfg.l2_reg; A.loss_scale = c->cfg.loss_scale;
	A.ema_decay = c->cfg.ema_decay;
	A.ema_debias_old = 1 - (float)std::pow(c->cfg.ema_decay, c->opt_step - 1);       // ema.h:121-122
	A.ema_debias_new = 1.0f / (1 - (float)std::pow(c->cfg.ema_decay, c->opt_step));
	A.n_params = c->M.n_params; A.n_matrix = c->M.off_grid; A.rgb_begin = c->off_rgb; A.rgb_end = c->M.off_grid; A.only_sdf = c->flags.only_sdf_training;
	A.log2_beta1 = (float)std::log2((double)c->cfg.beta1); A.log2_beta2 = (float)std::log2((double)c->cfg.beta2);
	A.shard_begin = c->shard_end ? c->shard_begin : 0u; A.shard_end = c->shard_end ? std::min(c->shard_end, c->M.n_params) : c->M.n_params; A.gsrc = c->shard_end ? c->shard_grads : nullptr;
	if (gsrc16 && sh_end) { A.shard_begin = sh_begin; A.shard_end = std::min(sh_end, c->M.n_params); A.gsrc = nullptr; }      // sharded optimizer on the library's own communicator
	A.gsrc16 = gsrc16;
	A.first = 0; A.last = c->M.n_params;
	if (c->xch_chunks > 1) {
		// chunked exchange: the all-reduce of chunk k + 1 (communication stream) runs under Adam / EMA of chunk k (this stream)
		NvtxRange nvtx_("adam_ema_pipelined"); prof_begin(c, st, "adam_ema");
		for (uint32_t k = 0; k < c->xch_chunks; ++k) {
			CU(cudaStreamWaitEvent(st, c->ev_chunk[k], 0));
			A.first = k * c->xch_chunk_elems; A.last = std::min<uint32_t>((k + 1) * c->xch_chunk_elems, c->M.n_params);
			launch_adam_ema(st, A, c->master, c->params, c->ema, c->grads, c->m1, c->m2, c->steps);
		}
		prof_end(c, st); c->launches += c->xch_chunks;
	} else {
		KT("adam_ema", 1, launch_adam_ema(st, A, c->master, c->params, c->ema, c->grads, c->m1, c->m2, c->steps));
	}
	CU(cudaGetLastError());
	return RNB_OK;
}

int rnb_train_step_begin(rnb_ctx* c, void* stream) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->views_dev) return fail(RNB_ERR_STATE, "no dataset: call rnb_set_dataset / rnb_upload_dataset first");
	if (c->in_step) return fail(RNB_ERR_STATE, "rnb_train_step_begin called twice");
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t R = c->rays_per_batch;
	int rc = ensure_ray_capacity(c, R); if (rc) return rc;
	if (c->training_step == 0 || c->canonical_step == 0) c->n_rays_total = 0;      // :3906-3911
	const uint32_t nrt = c->n_rays_total; c->n_rays_total += R;
	c->step_R = R; c->step_nrt = nrt;
	// pre-launch the NEXT step's march on the side stream (pinned batch size only: the adaptive controller fixes the next batch size
	// after this step's counters are read; not on steps that refresh the occupancy grid first; not while profiling stages)
	{
		const uint32_t ts_next = c->training_step + 1, skip = std::min(std::max(ts_next / 16u, 1u), 16u);
		c->pre_armed = c->cfg.pin_rays_per_batch && !c->prof && c->prelaunch && ts_next % skip != 0 && R <= c->cap_rays;
		// With a host that waits for every step the march's launch reaches the device ~0.1 ms into the step and fills the gaps behind pass A by itself;
		// when the host runs ahead it is eligible the moment its event fires: behind scan/emit it would take the SMs before pass A is resident
		// (pass A and the backward use every register of an SM), so it is released behind pass A instead, beside the small compaction / loss kernels
		c->pre_at = c->pre_at_env >= 0 ? c->pre_at_env : (c->step_waits ? 3 : 2);
	}
	rc = step_front(c, st, R, nrt); if (rc) return rc;
	c->rng.advance();                                               // :4118
	c->in_step = true;
	{
		if (c->pre_armed) {      // the side stream waits for the point chosen by pre_at (recorded inside step_front, or here: behind the backward)
			if (c->pre_at == 0) CU(cudaEventRecord(c->ev_bwd, st));
			CU(cudaStreamWaitEvent(c->side, c->ev_bwd, 0));
			launch_march(c->side, R, c->cfg.world_size, c->cfg.rank, c->n_rays_total, c->rng, c->views_dev, c->n_views, c->bitfield, c->ray_n, c->ray_geom, c->ts, (uint32_t)c->n_sm * c->pre_ctas_per_sm);
			CU(cudaEventRecord(c->ev_march, c->side));
			c->launches += 1;
			c->pre_valid = true; c->pre_R = R; c->pre_nrt = c->n_rays_total; c->pre_rng_state = c->rng.state; c->pre_rng_inc = c->rng.inc;
		}
	}
	return RNB_OK;
} RNB_API_CATCH

int rnb_train_step_end(rnb_ctx* c, void* stream, rnb_step_stats* stats) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->in_step) return fail(RNB_ERR_STATE, "rnb_train_step_end without begin");
	cudaStream_t st = (cudaStream_t)stream;
	int rc = c->xch16 ? optimizer_step(c, st, c->xch16, c->xch_begin, c->xch_end) : optimizer_step(c, st); if (rc) return rc;
	if (c->xch16 && c->xch_end) {      // sharded: every rank's next forward needs all updated shards of the binary16 training parameters
		NcclApi* N = nccl_api();
		const size_t shard = c->np_padded / c->cfg.world_size;
		c->ema_stale = true;
		KT("param_allgather", 1, rc = (int)N->AllGather(c->params + (size_t)c->cfg.rank * shard, c->params, shard, ncclHalf, c->comm, st));
		if (rc) return fail(RNB_ERR_CUDA, std::string("ncclAllGather: ") + N->GetErrorString((ncclResult_t)rc));
	}
	c->xch16 = nullptr; c->xch_begin = c->xch_end = 0; c->xch_chunks = 0;
	++c->training_step;
	c->canonical_step = c->training_step;              // testbed_nerf.cu:3646 (static scene: no global-movement phase)
	c->in_step = false;
	// Counters::update_after_training (:3532-3558).  The clamp of the next step's sample budget stays on the device (k_scan_compact -> counters[5]
	// -> k_scan_rays), so with a pinned batch size nothing of this step is needed on the host before the next one is queued: without `stats`
	// the call returns with the read-back in flight (pull_counters) — no host synchronisation.  The adaptive controller sizes the next
	// launch from the compacted count and therefore waits, like the reference does every step (:3535-3551, src/testbed.cu:2866).
	CU(cudaMemcpyAsync(c->counters_host, c->counters, 8 * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(c->stats_host, c->stats, 8 * 4, cudaMemcpyDeviceToHost, st));
	const uint32_t R = c->step_R;
	if (!stats && c->cfg.pin_rays_per_batch && !c->prof && c->async_end) {
		CU(cudaEventRecord(c->ev_counters, st));
		c->counters_pending = true;
		return RNB_OK;
	}
	CU(cudaStreamSynchronize(st));
	c->counters_pending = false;
	prof_resolve(c);
	const uint32_t total = c->counters_host[2];
	// data parallel: every rank must derive the SAME next batch size, so the controller is driven by the all-reduced compacted count
	// (stats[3], summed over ranks together with the losses) against the global target; single GPU: the local count
	const uint32_t total_global = c->cfg.world_size > 1 ? (uint32_t)(c->stats_host[3] + 0.5f) : total;
	apply_counters(c);
	if (!c->cfg.pin_rays_per_batch && total_global > 0 && (c->cfg.world_size > 1 || c->counters_host[1] != 0)) {
		uint32_t r = (uint32_t)((float)R * (float)c->cfg.target_batch_size / (float)total_global);
		c->rays_per_batch = std::min(next_multiple(r, 128u), 1u << 18);
	}
	if (stats) {
		// stats_host[3] = compacted sample count as float: summed over ranks together with the losses when data-parallel
		const float f = c->stats_host[3] / (float)c->cfg.target_batch_size;
		stats->loss = c->stats_host[0] * f; stats->ek_loss = c->stats_host[1] * f; stats->mask_loss = c->stats_host[2] * f;
		stats->n_rays = R; stats->n_rays_kept = c->counters_host[0]; stats->n_samples = c->counters_host[1]; stats->n_samples_compacted = total;
		stats->n_samples_trained = c->counters_host[3]; stats->rays_per_batch_next = c->rays_per_batch; stats->training_step = c->training_step; stats->density_grid_updated = 0;
	}
	return RNB_OK;
} RNB_API_CATCH

// ---- data parallelism behind the boundary ----------------------------------------------------------------------------------
// The reference is single-GPU; its gradient buffer is one contiguous binary16 array between backward and optimizer_step (trainer.h:78-84,
// testbed_nerf.cu:4068 -> :3624): that is where the collective goes.  Rays i = rank (mod world) are marched by this context (rnb_config);
// after the backward the fp32 accumulators are rounded to binary16 ONCE (the reference's gradient format), summed over the ranks with one
// ncclAllReduce on the caller's stream (21 MB instead of the 42 MB of the fp32 buffer; in-switch reduction on NVSwitch systems), together with
// the 8 floats of loss sums / sample counts in the same NCCL group, and Adam/EMA read the binary16 sum.  RNB_DP=sharded: reduce-scatter ->
// Adam/EMA on 1/world of the parameters -> all-gather of the binary16 training parameters (same bytes on the wire, 1/world of the optimizer).
int rnb_comm_unique_id(uint8_t id_out[RNB_COMM_ID_BYTES]) try {
	if (!id_out) return fail(RNB_ERR_INVALID, "null argument");
	NcclApi* N = nccl_api(); if (!N) return fail(RNB_ERR_STATE, "libnccl.so.2 not found");
	static_assert(RNB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
	ncclUniqueId id; NC(N->GetUniqueId(&id));
	memcpy(id_out, id.internal, RNB_COMM_ID_BYTES);
	return RNB_OK;
} RNB_API_CATCH
static int comm_buffers(rnb_ctx* c) {
	if (!c->grads16) { CU(cudaMalloc(&c->grads16, c->np_padded * 2)); CU(cudaMemset(c->grads16, 0, c->np_padded * 2)); }
	// default: sharded optimizer (reduce-scatter -> Adam / EMA on 1 / world of the parameters -> all-gather of the binary16 parameters): 0.825 vs 0.835 ms per
	// step at N = 2 and 0.857 vs 0.929 ms at N = 8 against the single all-reduce (profiles/r02_dp_n2.txt, r02_dp_n8.txt).  RNB_DP=allreduce selects the latter.
	const char* e = getenv("RNB_DP");
	c->dp_sharded = (!(e && std::string(e) == "allreduce") && c->np_padded % ((size_t)c->cfg.world_size * 8) == 0) ? 1 : 0;
	{ const char* x = getenv("RNB_DP_EXACT"); c->dp_exact = !(x && atoi(x) == 0); }
	if (const char* d = getenv("RNB_DP_CHUNKS")) c->dp_chunks = (uint32_t)std::min<int>(std::max(atoi(d), 1), (int)rnb_ctx::MAX_CHUNKS);
	if (!c->comm_stream) {
		int prio_lo = 0, prio_hi = 0; CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
		CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi));
		CU(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
		for (uint32_t k = 0; k < rnb_ctx::MAX_CHUNKS; ++k) CU(cudaEventCreateWithFlags(&c->ev_chunk[k], cudaEventDisableTiming));
	}
	return RNB_OK;
}
int rnb_comm_init(rnb_ctx* c, const uint8_t id_in[RNB_COMM_ID_BYTES]) try {
	if (!c || !id_in) return fail(RNB_ERR_INVALID, "null argument");
	if (c->comm) return fail(RNB_ERR_STATE, "communicator already set");
	if (c->in_step) return fail(RNB_ERR_STATE, "communicator changed inside a step");
	NcclApi* N = nccl_api(); if (!N) return fail(RNB_ERR_STATE, "libnccl.so.2 not found");
	ncclUniqueId id; memcpy(id.internal, id_in, RNB_COMM_ID_BYTES);
	NC(N->CommInitRank(&c->comm, (int)c->cfg.world_size, id, (int)c->cfg.rank));
	c->comm_owned = true;
	return comm_buffers(c);
} RNB_API_CATCH
int rnb_comm_adopt(rnb_ctx* c, void* nccl_comm) try {
	if (!c || !nccl_comm) return fail(RNB_ERR_INVALID, "null argument");
	if (c->comm) return fail(RNB_ERR_STATE, "communicator already set");
	NcclApi* N = nccl_api(); if (!N) return fail(RNB_ERR_STATE, "libnccl.so.2 not found");
	int n = 0, r = -1; NC(N->CommCount((ncclComm_t)nccl_comm, &n)); NC(N->CommUserRank((ncclComm_t)nccl_comm, &r));
	if ((uint32_t)n != c->cfg.world_size || (uint32_t)r != c->cfg.rank) return fail(RNB_ERR_INVALID, "communicator size / rank differ from rnb_config.world_size / rank");
	c->comm = (ncclComm_t)nccl_comm; c->comm_owned = false;
	return comm_buffers(c);
} RNB_API_CATCH
int rnb_comm_destroy(rnb_ctx* c) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (c->in_step) return fail(RNB_ERR_STATE, "communicator changed inside a step");
	if (c->comm && c->comm_owned) { NcclApi* N = nccl_api(); CU(cudaDeviceSynchronize()); if (N) N->CommDestroy(c->comm); }
	c->comm = nullptr; c->comm_owned = false;
	return RNB_OK;
} RNB_API_CATCH
// sharded optimizer only: the EMA (inference) parameters are maintained per shard; before a snapshot export or a mesh extraction every rank
// calls this once (collective) to gather all shards.  No-op with the all-reduce protocol.
int rnb_comm_sync_ema(rnb_ctx* c, void* stream) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->comm || !c->dp_sharded) return RNB_OK;
	NcclApi* N = nccl_api(); if (!N) return fail(RNB_ERR_STATE, "libnccl.so.2 not found");
	const size_t shard = c->np_padded / c->cfg.world_size;
	NC(N->AllGather(c->ema + (size_t)c->cfg.rank * shard, c->ema, shard, ncclHalf, c->comm, (cudaStream_t)stream));
	c->ema_stale = false;
	return RNB_OK;
} RNB_API_CATCH
int rnb_comm_info(rnb_ctx* c, uint32_t out[4]) try {
	if (!c || !out) return fail(RNB_ERR_INVALID, "null argument");
	NcclApi* N = nccl_api(); int v = 0; if (N) N->GetVersion(&v);
	out[0] = c->comm ? 1u : 0u; out[1] = (uint32_t)v; out[2] = (uint32_t)c->dp_sharded | ((c->comm && c->dp_exact && c->use_tc_bwd) ? 2u : 0u); out[3] = c->cfg.world_size;
	return RNB_OK;
} RNB_API_CATCH

// between rnb_train_step_begin and rnb_train_step_end
static int exchange_gradients(rnb_ctx* c, cudaStream_t st) {
	NcclApi* N = nccl_api(); if (!N) return fail(RNB_ERR_STATE, "libnccl.so.2 not found");
	const uint32_t np = (uint32_t)c->np_padded;
	KT("grad_pack", 1, launch_pack_grads(st, np, c->grads, c->grads16));
	CU(cudaGetLastError());
	c->xch_chunks = 0;
	if (!c->dp_sharded && c->dp_chunks > 1 && !c->prof) {
		// chunked + pipelined: chunk boundaries are multiples of 512 parameters; the statistics ride with chunk 0
		const uint32_t per = (uint32_t)(((c->np_padded + c->dp_chunks - 1) / c->dp_chunks + 511) / 512 * 512);
		c->xch_chunk_elems = per; c->xch_chunks = (uint32_t)((c->np_padded + per - 1) / per);
		CU(cudaEventRecord(c->ev_pack, st));
		CU(cudaStreamWaitEvent(c->comm_stream, c->ev_pack, 0));
		for (uint32_t k = 0; k < c->xch_chunks; ++k) {
			const size_t off = (size_t)k * per, cnt = std::min<size_t>(per, c->np_padded - off);
			NC(N->GroupStart());
			NC(N->AllReduce(c->grads16 + off, c->grads16 + off, cnt, ncclHalf, ncclSum, c->comm, c->comm_stream));
			if (k == 0) NC(N->AllReduce(c->stats, c->stats, 8, ncclFloat, ncclSum, c->comm, c->comm_stream));
			NC(N->GroupEnd());
			CU(cudaEventRecord(c->ev_chunk[k], c->comm_stream));
		}
		c->xch16 = c->grads16; c->xch_begin = 0; c->xch_end = 0;
		c->launches += c->xch_chunks;
		return RNB_OK;
	}
	prof_begin(c, st, "grad_exchange");
	NC(N->GroupStart());
	if (c->dp_sharded) {
		const size_t shard = c->np_padded / c->cfg.world_size;
		NC(N->ReduceScatter(c->grads16, c->grads16 + (size_t)c->cfg.rank * shard, shard, ncclHalf, ncclSum, c->comm, st));
		c->xch16 = c->grads16 + (size_t)c->cfg.rank * shard; c->xch_begin = (uint32_t)(c->cfg.rank * shard); c->xch_end = (uint32_t)((c->cfg.rank + 1) * shard);
	} else {
		NC(N->AllReduce(c->grads16, c->grads16, np, ncclHalf, ncclSum, c->comm, st));
		c->xch16 = c->grads16; c->xch_begin = 0; c->xch_end = 0;
	}
	NC(N->AllReduce(c->stats, c->stats, 8, ncclFloat, ncclSum, c->comm, st));
	NC(N->GroupEnd());
	prof_end(c, st); c->launches += 1;
	return RNB_OK;
}

int rnb_train_step(rnb_ctx* c, void* stream, rnb_step_stats* stats) try {
	// refused before anything is queued: a data-parallel step whose gradients nobody exchanges would silently train on 1 / world of the rays
	if (c && c->cfg.world_size > 1 && !c->comm) return fail(RNB_ERR_STATE, "world_size > 1: call rnb_comm_init / rnb_comm_adopt first, or drive rnb_train_step_begin / _end with your own collective");
	if (c) c->step_waits = stats != nullptr || !c->cfg.pin_rays_per_batch || !c->async_end;
	int rc = rnb_train_step_begin(c, stream);
	if (c) c->step_waits = true;
	if (rc) return rc;
	if (c->cfg.world_size > 1) {
		rc = exchange_gradients(c, (cudaStream_t)stream); if (rc) { c->in_step = false; return rc; }
	}
	return rnb_train_step_end(c, stream, stats);
} RNB_API_CATCH

int rnb_train(rnb_ctx* c, void* stream, rnb_step_stats* stats) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (c->cfg.world_size > 1 && !c->comm) return fail(RNB_ERR_STATE, "world_size > 1: call rnb_comm_init / rnb_comm_adopt first, or drive rnb_train_step_begin / _end with your own collective");
	NvtxRange nvtx_("train");
	const uint32_t skip = std::min(std::max(c->canonical_step / 16u, 1u), 16u);    // src/testbed.cu:2805-2806
	uint32_t updated = 0;
	if (c->canonical_step % skip == 0) { int rc = rnb_prep(c, stream); if (rc) return rc; updated = 1; }
	int rc = rnb_train_step(c, stream, stats);
	if (!rc && stats) stats->density_grid_updated = updated;
	return rc;
} RNB_API_CATCH

// ---- in-memory checkpoint / resume (one slot, device side) -------------------------------------------------------------
// Everything Testbed::train reads and writes between steps: fp32 master / fp16 / EMA parameters, Adam moments and per-parameter
// step counters, density grid + bitfield, both pcg32 streams, the controller counters.  (The gradient buffer is zero between
// steps.)  Used by bench.py to time `value` and `e2e` on the SAME training steps; the file-level snapshot of the reference
// (src/testbed.cu:3280-3390) keeps only the EMA fp16 weights and the density grid and goes through rnb_export_* / rnb_import_*.
int rnb_checkpoint_save(rnb_ctx* c) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (c->in_step) return fail(RNB_ERR_STATE, "checkpoint inside a step");
	drop_prelaunch(c);
	pull_counters(c);
	const size_t np = c->M.n_params, need = np * 20 + (size_t)GRID_CELLS * 4 + GRID_CELLS;
	if (c->ck.bytes < need) { cudaFree(c->ck.buf); c->ck.buf = nullptr; CU(cudaMalloc(&c->ck.buf, need)); c->ck.bytes = need; }
	CU(cudaDeviceSynchronize());
	uint8_t* b = (uint8_t*)c->ck.buf;
	CU(cudaMemcpy(b, c->master, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(b, c->m1, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(b, c->m2, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(b, c->steps, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(b, c->params, np * 2, cudaMemcpyDeviceToDevice)); b += np * 2;
	CU(cudaMemcpy(b, c->ema, np * 2, cudaMemcpyDeviceToDevice)); b += np * 2;
	CU(cudaMemcpy(b, c->density_grid, (size_t)GRID_CELLS * 4, cudaMemcpyDeviceToDevice)); b += (size_t)GRID_CELLS * 4;
	CU(cudaMemcpy(b, c->bitfield, GRID_CELLS, cudaMemcpyDeviceToDevice));
	auto& k = c->ck;
	k.opt_step = c->opt_step; k.lr_factor = c->lr_factor; k.density_ema_step = c->density_ema_step; k.training_step = c->training_step; k.rays_per_batch = c->rays_per_batch;
	k.n_rays_total = c->n_rays_total; k.measured_before = c->measured_before; k.measured = c->measured; k.rng = c->rng; k.density_rng = c->density_rng;
	k.canonical_step = c->canonical_step; k.n_images_prev = c->n_images_prev;
	k.valid = true;
	return RNB_OK;
} RNB_API_CATCH
int rnb_checkpoint_restore(rnb_ctx* c) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->ck.valid) return fail(RNB_ERR_STATE, "no checkpoint");
	if (c->in_step) return fail(RNB_ERR_STATE, "restore inside a step");
	drop_prelaunch(c);
	pull_counters(c);
	CU(cudaDeviceSynchronize());
	const size_t np = c->M.n_params;
	const uint8_t* b = (const uint8_t*)c->ck.buf;
	CU(cudaMemcpy(c->master, b, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(c->m1, b, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(c->m2, b, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(c->steps, b, np * 4, cudaMemcpyDeviceToDevice)); b += np * 4;
	CU(cudaMemcpy(c->params, b, np * 2, cudaMemcpyDeviceToDevice)); b += np * 2;
	CU(cudaMemcpy(c->ema, b, np * 2, cudaMemcpyDeviceToDevice)); b += np * 2;
	CU(cudaMemcpy(c->density_grid, b, (size_t)GRID_CELLS * 4, cudaMemcpyDeviceToDevice)); b += (size_t)GRID_CELLS * 4;
	CU(cudaMemcpy(c->bitfield, b, GRID_CELLS, cudaMemcpyDeviceToDevice));
	CU(cudaMemset(c->grads, 0, np * 4));
	const auto& k = c->ck;
	c->opt_step = k.opt_step; c->lr_factor = k.lr_factor; c->density_ema_step = k.density_ema_step; c->training_step = k.training_step; c->rays_per_batch = k.rays_per_batch;
	c->n_rays_total = k.n_rays_total; c->measured_before = k.measured_before; c->measured = k.measured; c->rng = k.rng; c->density_rng = k.density_rng;
	c->canonical_step = k.canonical_step; c->n_images_prev = k.n_images_prev;
	CU(push_measured(c));
	return RNB_OK;
} RNB_API_CATCH

int rnb_profile_enable(rnb_ctx* c, int on) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	c->prof = on != 0;
	if (on) { for (auto& a : c->prof_acc) { a.ms = 0; a.calls = 0; } }
	return RNB_OK;
} RNB_API_CATCH
// names_buf receives ';'-separated stage names; ms / calls are parallel arrays of capacity *n (in) and count (out)
int rnb_profile_read(rnb_ctx* c, char* names_buf, size_t names_cap, double* ms, uint64_t* calls, uint32_t* n) try {
	if (!c || !n) return fail(RNB_ERR_INVALID, "null argument");
	std::string names; uint32_t k = 0;
	for (auto& a : c->prof_acc) { if (k >= *n) break; names += a.name; names += ';'; ms[k] = a.ms; calls[k] = a.calls; ++k; }
	*n = k;
	if (names_buf && names_cap) { strncpy(names_buf, names.c_str(), names_cap - 1); names_buf[names_cap - 1] = 0; }
	return RNB_OK;
} RNB_API_CATCH
int rnb_launch_count(rnb_ctx* c, uint64_t* out) { if (!c || !out) return fail(RNB_ERR_INVALID, "null argument"); *out = c->launches; return RNB_OK; }

int rnb_grad_buffer(rnb_ctx* c, float** g, uint64_t* n) { if (!c) return fail(RNB_ERR_INVALID, "null ctx"); *g = c->grads; *n = c->M.n_params; return RNB_OK; }
// host copies for parity tests: the fp32 gradient accumulators (valid between rnb_train_step_begin and _end) and the per-ray
// loss terms of the last step (loss_output / ek_loss_output / mask_loss_output of compute_loss_kernel, testbed_nerf.cu:1396-2097)
int rnb_get_grads_fp32(rnb_ctx* c, float* host, size_t n) try {
	if (!c || !host || n != c->M.n_params) return fail(RNB_ERR_INVALID, "bad gradient buffer");
	CU(cudaMemcpy(host, c->grads, n * 4, cudaMemcpyDeviceToHost));
	return RNB_OK;
} RNB_API_CATCH
int rnb_get_ray_losses(rnb_ctx* c, uint32_t cap, uint32_t* ray_idx, float* loss3, uint32_t* n_out) try {
	if (!c || !ray_idx || !loss3 || !n_out) return fail(RNB_ERR_INVALID, "null argument");
	uint32_t K = 0; CU(cudaMemcpy(&K, c->counters, 4, cudaMemcpyDeviceToHost));
	K = std::min(K, std::min(cap, c->cap_rays));
	CU(cudaMemcpy(ray_idx, c->ray_indices, (size_t)K * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(loss3, c->loss_out, (size_t)K * 12, cudaMemcpyDeviceToHost));
	*n_out = K;
	return RNB_OK;
} RNB_API_CATCH
// per kept ray of the last step: samples marched and samples kept by the transmittance cut (profiling / tests)
int rnb_get_ray_counts(rnb_ctx* c, uint32_t cap, uint32_t* marched, uint32_t* kept, uint32_t* n_out) try {
	if (!c || !marched || !kept || !n_out) return fail(RNB_ERR_INVALID, "null argument");
	uint32_t K = 0; CU(cudaMemcpy(&K, c->counters, 4, cudaMemcpyDeviceToHost));
	K = std::min(K, std::min(cap, c->cap_rays));
	std::vector<uint32_t> ns((size_t)K * 2);
	CU(cudaMemcpy(ns.data(), c->numsteps, (size_t)K * 8, cudaMemcpyDeviceToHost));
	for (uint32_t k = 0; k < K; ++k) marched[k] = ns[2 * k];
	CU(cudaMemcpy(kept, c->n_fwd, (size_t)K * 4, cudaMemcpyDeviceToHost));
	*n_out = K;
	return RNB_OK;
} RNB_API_CATCH
// ---- data-parallel optimizer sharding (new) ---------------------------------------------------------------
int rnb_param_buffers(rnb_ctx* c, void** params_fp16, void** ema_fp16, uint64_t* n_params, uint64_t* n_padded) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (params_fp16) *params_fp16 = c->params;
	if (ema_fp16) *ema_fp16 = c->ema;
	if (n_params) *n_params = c->M.n_params;
	if (n_padded) *n_padded = c->np_padded;
	return RNB_OK;
} RNB_API_CATCH
int rnb_set_optimizer_shard(rnb_ctx* c, uint64_t begin, uint64_t end, const float* reduced_grads_dev) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (c->in_step) return fail(RNB_ERR_STATE, "optimizer shard changed inside a step");
	if (end == 0) { c->shard_begin = c->shard_end = 0; c->shard_grads = nullptr; return RNB_OK; }
	if (begin % 8 || end % 8 || begin >= end || end > c->np_padded) return fail(RNB_ERR_INVALID, "optimizer shard must be a non-empty [begin, end) of multiples of 8 inside the padded parameter range");
	c->shard_begin = (uint32_t)begin; c->shard_end = (uint32_t)end; c->shard_grads = reduced_grads_dev;
	return RNB_OK;
} RNB_API_CATCH
int rnb_stat_buffer(rnb_ctx* c, float** s, uint64_t* n) { if (!c) return fail(RNB_ERR_INVALID, "null ctx"); *s = c->stats; *n = 8; return RNB_OK; }

int rnb_eval_sdf(rnb_ctx* c, const float* xyz_dev, size_t n, float* sdf_dev, float* normal_dev, float* density_dev, int use_ema, void* stream) try {
	if (!c || !xyz_dev) return fail(RNB_ERR_INVALID, "null argument");
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	float4* tmp = nullptr;
	const size_t CH = 1u << 20;
	RNB_EMA_READY(c, use_ema);
	const __half* P = use_ema ? c->ema : c->params;
	if (c->use_tc && !normal_dev) launch_tc(0, st, c->M, P, c->wtc, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, c->n_sm);     // weight blob of the requested parameter set (re-packed by the next training step)
	CU(cudaMallocAsync(&tmp, CH * 16, st));
	for (size_t o = 0; o < n; o += CH) {
		const size_t m = std::min(CH, n - o);
		CU(cudaMemcpy2DAsync(tmp, 16, xyz_dev + o * 3, 12, 12, m, cudaMemcpyDeviceToDevice, st));
		// sdf / density only (marching-cubes sweep): the tcgen05 probe kernel; with normals: the CUDA-core kernel
		if (c->use_tc && !normal_dev) launch_tc(2, st, c->M, P, c->wtc, vl, tmp, nullptr, (uint32_t)m, nullptr, sdf_dev ? sdf_dev + o : nullptr, density_dev ? density_dev + o : nullptr, c->n_sm);
		else launch_forward_simt(st, c->M, P, vl, 2, tmp, nullptr, (uint32_t)m, nullptr, nullptr, sdf_dev ? sdf_dev + o : nullptr, normal_dev ? normal_dev + o * 3 : nullptr, density_dev ? density_dev + o : nullptr);
	}
	CU(cudaFreeAsync(tmp, st));
	CU(cudaGetLastError());
	return RNB_OK;
} RNB_API_CATCH

// SDF on a lattice — Testbed::get_density_on_grid (src/testbed_nerf.cu:4218-4269: generate_grid_samples_nerf_uniform + NerfNetwork::sdf in
// 1 M-point batches + grid_samples_half_to_float) in one launch: the lattice positions are generated inside the tcgen05 probe kernel
// (the reference materialises 12 B per point first: 12.9 GB at 1024^3).  out_dev[x + y rx + z rx ry] = sdf (incl. bias), fp32.
int rnb_sdf_on_grid(rnb_ctx* c, const uint32_t res[3], const float aabb_min[3], const float aabb_max[3], float* out_dev, int use_ema, void* stream) try {
	if (!c || !res || !aabb_min || !aabb_max || !out_dev) return fail(RNB_ERR_INVALID, "null argument");
	if ((uint64_t)res[0] * res[1] * res[2] > 0xFFFFFFFFull) return fail(RNB_ERR_INVALID, "lattice larger than 2^32 points");
	if (!c->use_tc) return fail(RNB_ERR_STATE, "rnb_sdf_on_grid needs the tcgen05 network path");
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	RNB_EMA_READY(c, use_ema);
	const __half* P = use_ema ? c->ema : c->params;
	launch_tc(0, st, c->M, P, c->wtc, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, c->n_sm);
	launch_tc_sdf_grid(st, c->M, P, c->wtc, vl, res, aabb_min, aabb_max, out_dev, c->n_sm);
	CU(cudaGetLastError());
	return RNB_OK;
} RNB_API_CATCH

// ---- mesh extraction ----------------------------------------------------------------------------------------
static void mesh_free(rnb_ctx* c) {
	cudaFree(c->mesh.verts); cudaFree(c->mesh.normals); cudaFree(c->mesh.colors); cudaFree(c->mesh.indices);
	c->mesh = rnb_ctx::Mesh();
}

// marching_cubes_gpu + compute_mesh_1ring + compute_mesh_vertex_colors on a caller-provided lattice of SDF values
// (src/marching_cubes.cu:794-822, :722-728, src/testbed_nerf.cu:4193-4216).  with_colors == 0 skips the network pass (colours zero).
int rnb_marching_cubes_from_density(rnb_ctx* c, const float* density_dev, const uint32_t res[3], const float aabb_min[3], const float aabb_max[3], float thresh,
                                    int with_colors, int use_ema, void* stream, rnb_mesh_info* info) try {
	if (!c || !density_dev || !res || !aabb_min || !aabb_max) return fail(RNB_ERR_INVALID, "null argument");
	if (res[0] == 0 || res[1] == 0 || res[2] == 0 || res[0] % 16 != 0) return fail(RNB_ERR_INVALID, "lattice x resolution must be a positive multiple of 16");
	if ((uint64_t)res[0] * res[1] * res[2] > 0xFFFFFFFFull) return fail(RNB_ERR_INVALID, "lattice larger than 2^32 points");
	if ((uintptr_t)density_dev & 3u) return fail(RNB_ERR_INVALID, "density lattice must be 4-byte aligned");
	if (c->in_step) return fail(RNB_ERR_STATE, "training step in flight");
	cudaStream_t st = (cudaStream_t)stream;
	mesh_free(c);
	rnb_ctx::Mesh& m = c->mesh;
	const std::string e = mesh_extract(st, density_dev, res, aabb_min, aabb_max, thresh, &c->mesh_ws, &c->mesh_ws_bytes, &m.verts, &m.normals, &m.indices, &m.n_verts, &m.n_verts_padded, &m.n_indices, m.ms + 1, &c->launches);
	m.ms[0] = 0.f;
	if (!e.empty()) { mesh_free(c); return fail(RNB_ERR_CUDA, e); }
	const size_t nvp = std::max<uint32_t>(m.n_verts_padded, 1);
	CU(cudaMalloc(&m.colors, nvp * 12));
	CU(cudaMemsetAsync(m.colors, 0, nvp * 12, st));
	EventPair ec;
	CU(cudaEventCreate(&ec[0])); CU(cudaEventCreate(&ec[1]));
	CU(cudaEventRecord(ec[0], st));
	if (with_colors && m.n_verts_padded) {
		// the padding vertices (zeros) go through the network as well, as in the reference
		const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
		RNB_EMA_READY(c, use_ema);
		const __half* P = use_ema ? c->ema : c->params;
		const uint32_t CH = c->cap_compact;
		float* dirw = nullptr;
		CU(cudaMalloc(&dirw, (size_t)std::min(CH, m.n_verts_padded) * 12));
		net_pack(c, st, P);
		for (uint32_t o = 0; o < m.n_verts_padded; o += CH) {
			const uint32_t k = std::min(CH, m.n_verts_padded - o);
			launch_mesh_color_inputs(st, k, m.verts + (size_t)o * 3, c->cpos4, dirw);
			net_pass_b(c, st, P, vl, c->cpos4, nullptr, k, dirw);
			launch_mesh_colors(st, k, c->out16, m.colors + (size_t)o * 3);
			c->launches += 3;
		}
		if (P != c->params) net_pack(c, st, c->params);
		CU(cudaStreamSynchronize(st));
		CU(cudaFree(dirw));
	}
	CU(cudaEventRecord(ec[1], st));
	CU(cudaStreamSynchronize(st));
	CU(cudaGetLastError());
	cudaEventElapsedTime(&m.ms[3], ec[0], ec[1]);
	if (info) {
		info->n_verts = m.n_verts; info->n_verts_padded = m.n_verts_padded; info->n_indices = m.n_indices; info->res[0] = res[0]; info->res[1] = res[1]; info->res[2] = res[2];
		for (int k = 0; k < 4; ++k) info->stage_ms[k] = m.ms[k];
	}
	return RNB_OK;
} RNB_API_CATCH

// Testbed::marching_cubes (src/testbed_nerf.cu:4297-4348): resolution rounded up to multiples of 16, SDF sweep, extraction, normals, colours.
int rnb_marching_cubes(rnb_ctx* c, const uint32_t res_in[3], const float aabb_min[3], const float aabb_max[3], float thresh, int use_ema, void* stream, rnb_mesh_info* info) try {
	if (!c || !res_in || !aabb_min || !aabb_max) return fail(RNB_ERR_INVALID, "null argument");
	const uint32_t res[3] = {next_multiple(res_in[0], 16u), next_multiple(res_in[1], 16u), next_multiple(res_in[2], 16u)};
	const uint64_t n = (uint64_t)res[0] * res[1] * res[2];
	if (n == 0 || n > 0xFFFFFFFFull) return fail(RNB_ERR_INVALID, "lattice empty or larger than 2^32 points");
	if (c->mesh_density_bytes < n * 4) { cudaFree(c->mesh_density); c->mesh_density = nullptr; c->mesh_density_bytes = 0; CU(cudaMalloc(&c->mesh_density, n * 4)); c->mesh_density_bytes = n * 4; }
	EventPair es;
	CU(cudaEventCreate(&es[0])); CU(cudaEventCreate(&es[1]));
	CU(cudaEventRecord(es[0], (cudaStream_t)stream));
	int rc = rnb_sdf_on_grid(c, res, aabb_min, aabb_max, c->mesh_density, use_ema, stream);
	CU(cudaEventRecord(es[1], (cudaStream_t)stream));
	c->launches += 2;
	if (rc == RNB_OK) rc = rnb_marching_cubes_from_density(c, c->mesh_density, res, aabb_min, aabb_max, thresh, 1, use_ema, stream, info);
	if (rc == RNB_OK) { cudaEventElapsedTime(&c->mesh.ms[0], es[0], es[1]); if (info) info->stage_ms[0] = c->mesh.ms[0]; }
	return rc;
} RNB_API_CATCH

int rnb_mesh_buffers(rnb_ctx* c, float** verts, float** normals, float** colors, uint32_t** indices, rnb_mesh_info* info) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->mesh.verts) return fail(RNB_ERR_STATE, "no mesh: call rnb_marching_cubes first");
	if (verts) *verts = c->mesh.verts;
	if (normals) *normals = c->mesh.normals;
	if (colors) *colors = c->mesh.colors;
	if (indices) *indices = c->mesh.indices;
	if (info) { info->n_verts = c->mesh.n_verts; info->n_verts_padded = c->mesh.n_verts_padded; info->n_indices = c->mesh.n_indices; info->res[0] = info->res[1] = info->res[2] = 0; for (int k = 0; k < 4; ++k) info->stage_ms[k] = c->mesh.ms[k]; }
	return RNB_OK;
} RNB_API_CATCH

// host copies of the mesh (Testbed::compute_marching_cubes_mesh, src/python_api.cu:99-130); each pointer may be null
int rnb_mesh_download(rnb_ctx* c, float* verts, float* normals, float* colors, uint32_t* indices) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (!c->mesh.verts) return fail(RNB_ERR_STATE, "no mesh: call rnb_marching_cubes first");
	const size_t vb = (size_t)c->mesh.n_verts_padded * 12;
	if (verts) CU(cudaMemcpy(verts, c->mesh.verts, vb, cudaMemcpyDeviceToHost));
	if (normals) CU(cudaMemcpy(normals, c->mesh.normals, vb, cudaMemcpyDeviceToHost));
	if (colors) CU(cudaMemcpy(colors, c->mesh.colors, vb, cudaMemcpyDeviceToHost));
	if (indices) CU(cudaMemcpy(indices, c->mesh.indices, (size_t)c->mesh.n_indices * 4, cudaMemcpyDeviceToHost));
	return RNB_OK;
} RNB_API_CATCH

// save_mesh (src/marching_cubes.cu:824-982) on device arrays; no context needed
int rnb_save_mesh(const float* verts_dev, const float* normals_dev, const float* colors_dev, const uint32_t* indices_dev, uint32_t n_verts, uint32_t n_indices, const char* path,
                  float nerf_scale, const float nerf_offset[3], float n2w_s, const float n2w_t[3], int invert_normals, void* stream, uint64_t* bytes_written) try {
	if (!path || !nerf_offset || !n2w_t) return fail(RNB_ERR_INVALID, "null argument");
	if ((n_verts && (!verts_dev || !normals_dev || !colors_dev)) || (n_indices && !indices_dev)) return fail(RNB_ERR_INVALID, "null mesh array");
	if (n_indices % 3) return fail(RNB_ERR_INVALID, "index count is not a multiple of 3");
	uint64_t launches = 0;
	const std::string e = mesh_write((cudaStream_t)stream, verts_dev, normals_dev, colors_dev, indices_dev, n_verts, n_indices, path, nerf_scale, nerf_offset, n2w_s, n2w_t, invert_normals, bytes_written, &launches);
	if (!e.empty()) return fail(RNB_ERR_CUDA, e);
	return RNB_OK;
} RNB_API_CATCH

// ---- stage-level entry points (host buffers) ------------------------------------------------------------------------
int rnb_stage_generate(rnb_ctx* c, uint32_t n_rays, uint32_t n_rays_total, uint32_t max_samples, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t counters[2]) try {
	if (!c || !c->views_dev) return fail(RNB_ERR_STATE, "no dataset");
	if (max_samples > c->max_samples) return fail(RNB_ERR_INVALID, "max_samples exceeds capacity");
	drop_prelaunch(c);
	int rc = ensure_ray_capacity(c, n_rays); if (rc) return rc;
	launch_march(0, n_rays, c->cfg.world_size, c->cfg.rank, n_rays_total, c->rng, c->views_dev, c->n_views, c->bitfield, c->ray_n, c->ray_geom, c->ts);
	launch_scan_rays(0, n_rays, max_samples, nullptr, c->ray_n, c->ray_indices, c->numsteps, c->counters);
	launch_emit(0, n_rays, c->counters, c->cfg.world_size, c->ray_indices, c->numsteps, c->ray_geom, c->ts, c->pos4);
	CU(cudaDeviceSynchronize());
	uint32_t cnt[2]; CU(cudaMemcpy(cnt, c->counters, 8, cudaMemcpyDeviceToHost));
	counters[0] = cnt[0]; counters[1] = cnt[1];
	const uint32_t K = cnt[0];
	std::vector<float> geom((size_t)n_rays * 9);
	CU(cudaMemcpy(ray_indices, c->ray_indices, K * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(numsteps, c->numsteps, K * 8, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(geom.data(), c->ray_geom, geom.size() * 4, cudaMemcpyDeviceToHost));
	uint32_t n_emitted = 0;
	for (uint32_t k = 0; k < K; ++k) n_emitted = std::max(n_emitted, numsteps[2 * k] + numsteps[2 * k + 1]);
	std::vector<float4> p4(n_emitted);
	CU(cudaMemcpy(p4.data(), c->pos4, (size_t)n_emitted * 16, cudaMemcpyDeviceToHost));
	for (uint32_t k = 0; k < K; ++k) {
		const float* g = &geom[(size_t)ray_indices[k] * 9];
		for (int q = 0; q < 6; ++q) rays[k * 6 + q] = g[q];
		for (uint32_t j = 0; j < numsteps[2 * k]; ++j) {
			const size_t s = numsteps[2 * k + 1] + j;
			float* o = coords + s * 7;
			o[0] = p4[s].x; o[1] = p4[s].y; o[2] = p4[s].z; o[3] = 0.f;
			o[4] = (g[6] + 1.0f) * 0.5f; o[5] = (g[7] + 1.0f) * 0.5f; o[6] = (g[8] + 1.0f) * 0.5f;
		}
	}
	return RNB_OK;
} RNB_API_CATCH

static int upload_coords(rnb_ctx* c, const float* coords, size_t n, float4* dst_pos, float** dirw_out) {
	std::vector<float4> p4(n); std::vector<float> dw(n * 3);
	for (size_t i = 0; i < n; ++i) {
		uint32_t slot = (uint32_t)i; float w; memcpy(&w, &slot, 4);
		p4[i] = make_float4(coords[i * 7], coords[i * 7 + 1], coords[i * 7 + 2], w);
		dw[i * 3] = coords[i * 7 + 4]; dw[i * 3 + 1] = coords[i * 7 + 5]; dw[i * 3 + 2] = coords[i * 7 + 6];
	}
	CU(cudaMemcpy(dst_pos, p4.data(), n * 16, cudaMemcpyHostToDevice));
	float* d = nullptr; CU(cudaMalloc(&d, std::max<size_t>(n, 1) * 12));
	CU(cudaMemcpy(d, dw.data(), n * 12, cudaMemcpyHostToDevice));
	*dirw_out = d;
	return RNB_OK;
}

int rnb_stage_forward(rnb_ctx* c, const float* coords, size_t n, int use_ema, float* out16, float* normal) try {
	if (!c || !coords || !out16) return fail(RNB_ERR_INVALID, "null argument");
	if (n > c->cap_compact) return fail(RNB_ERR_INVALID, "too many samples for one stage call");
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	float* dirw = nullptr;
	int rc = upload_coords(c, coords, n, c->cpos4, &dirw); if (rc) return rc;
	RNB_EMA_READY(c, use_ema);
	const __half* P = use_ema ? c->ema : c->params;
	net_pack(c, 0, P);
	net_pass_b(c, 0, P, vl, c->cpos4, nullptr, (uint32_t)n, dirw);
	float* nrm_dev = nullptr;
	if (normal) { CU(cudaMalloc(&nrm_dev, std::max<size_t>(n, 1) * 12)); launch_forward_simt(0, c->M, P, vl, 2, c->cpos4, nullptr, (uint32_t)n, nullptr, nullptr, nullptr, nrm_dev, nullptr); }
	CU(cudaDeviceSynchronize());
	std::vector<__half> h(n * 16);
	CU(cudaMemcpy(h.data(), c->out16, n * 32, cudaMemcpyDeviceToHost));
	for (size_t i = 0; i < n * 16; ++i) out16[i] = __half2float(h[i]);
	if (normal) { CU(cudaMemcpy(normal, nrm_dev, n * 12, cudaMemcpyDeviceToHost)); cudaFree(nrm_dev); }
	cudaFree(dirw);
	return RNB_OK;
} RNB_API_CATCH

int rnb_stage_loss(rnb_ctx* c, const float* out16_c, const uint32_t* ray_indices, const uint32_t* n_fwd, const uint32_t* cbase, const uint32_t* n_emit,
                   uint32_t K, uint32_t n_rays, uint32_t n_rays_total, float* dout16, float* loss, float* ek, float* mask) try {
	if (!c || !c->views_dev) return fail(RNB_ERR_STATE, "no dataset");
	int rc = ensure_ray_capacity(c, std::max(K, n_rays)); if (rc) return rc;
	size_t total = 0; for (uint32_t k = 0; k < K; ++k) total = std::max<size_t>(total, (size_t)cbase[k] + n_fwd[k]);
	if (total > c->cap_compact) return fail(RNB_ERR_INVALID, "too many compacted samples");
	std::vector<__half> h(total * 16); std::vector<float> dirw((size_t)K * 3);
	for (size_t i = 0; i < total * 16; ++i) h[i] = __float2half_rn(out16_c[i]);
	for (uint32_t k = 0; k < K; ++k) for (int d = 0; d < 3; ++d) dirw[k * 3 + d] = out16_c[(size_t)cbase[k] * 16 + 8 + d];   // already binary16 values
	CU(cudaMemcpy(c->out16, h.data(), total * 32, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->ray_dirw, dirw.data(), (size_t)K * 12, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->ray_indices, ray_indices, K * 4, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->n_fwd, n_fwd, K * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(c->cbase, cbase, K * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(c->n_emit, n_emit, K * 4, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(c->counters, &K, 4, cudaMemcpyHostToDevice));
	CU(cudaMemset(c->dout16, 0, total * 32));
	launch_loss(0, K, c->flags, n_rays, n_rays_total, c->training_step, c->cfg.loss_scale, c->counters, c->rng, c->views_dev, c->n_views, c->ray_indices, c->ray_dirw, c->n_fwd, c->cbase, c->n_emit,
	            c->out16, c->dout16, c->loss_out, nullptr);
	CU(cudaDeviceSynchronize());
	CU(cudaMemcpy(h.data(), c->dout16, total * 32, cudaMemcpyDeviceToHost));
	for (size_t i = 0; i < total * 16; ++i) dout16[i] = __half2float(h[i]);
	std::vector<float> lo((size_t)K * 3);
	CU(cudaMemcpy(lo.data(), c->loss_out, (size_t)K * 12, cudaMemcpyDeviceToHost));
	for (uint32_t k = 0; k < K; ++k) { loss[k] = lo[3 * k]; ek[k] = lo[3 * k + 1]; mask[k] = lo[3 * k + 2]; }
	return RNB_OK;
} RNB_API_CATCH

int rnb_stage_backward(rnb_ctx* c, const float* coords, const float* dout16, size_t n, uint32_t n_in_rollover, float* grads) try {
	if (!c || !coords || !dout16 || !grads) return fail(RNB_ERR_INVALID, "null argument");
	if (n > c->cfg.target_batch_size) return fail(RNB_ERR_INVALID, "too many samples");
	const uint32_t vl = valid_level_for_step(c, (int)c->training_step);
	float* dirw = nullptr;
	int rc = upload_coords(c, coords, n, c->cpos4, &dirw); if (rc) return rc;
	std::vector<__half> h(n * 16);
	for (size_t i = 0; i < n * 16; ++i) h[i] = __float2half_rn(dout16[i]);
	CU(cudaMemcpy(c->dout16, h.data(), n * 32, cudaMemcpyHostToDevice));
	CU(cudaMemset(c->grads, 0, (size_t)c->M.n_params * 4));
	uint32_t* nin = nullptr; CU(cudaMalloc(&nin, 4)); CU(cudaMemcpy(nin, &n_in_rollover, 4, cudaMemcpyHostToDevice));
	net_pack(c, 0, c->params);
	net_backward(c, 0, vl, nullptr, (uint32_t)n, c->cfg.target_batch_size, nin);
	CU(cudaDeviceSynchronize());
	CU(cudaMemcpy(grads, c->grads, (size_t)c->M.n_params * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemset(c->grads, 0, (size_t)c->M.n_params * 4));
	cudaFree(nin); cudaFree(dirw);
	return RNB_OK;
} RNB_API_CATCH

int rnb_stage_optimizer(rnb_ctx* c, const float* grads_host) try {
	if (!c) return fail(RNB_ERR_INVALID, "null ctx");
	if (grads_host) CU(cudaMemcpy(c->grads, grads_host, (size_t)c->M.n_params * 4, cudaMemcpyHostToDevice));
	int rc = optimizer_step(c, 0); if (rc) return rc;
	CU(cudaDeviceSynchronize());
	return RNB_OK;
} RNB_API_CATCH

} // extern "C"
