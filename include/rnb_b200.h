/*
 * rnb_b200.h — C ABI of the Blackwell-native NeuS2 / RNb-NeuS2 training inner loop.
 *
 * The reference (RobinBruneau/RNb-NeuS2) has no plugin/FFI seam on this path: L2–L0 are C++ templates linked into
 * ./build/testbed (CMakeLists.txt:322-333).  This header introduces the seam at the Testbed -> step boundary.
 * Every entry point names the reference interface it replaces (file:line in the reference tree).
 *
 * Conventions
 *   - plain C types only; no C++ exceptions cross the boundary.  Every call returns 0 on success or a negative
 *     rnb_status; rnb_last_error() returns a thread-local message (reference behaviour: CUDA_CHECK_THROW ->
 *     std::runtime_error; the caller turns non-zero into a throw).
 *   - `stream` arguments are cudaStream_t passed as void* (NULL = legacy default stream).
 *   - device pointers are marked _dev, host pointers _host.
 *   - one rnb_ctx per CUDA device / process; calls on one ctx must come from one host thread at a time
 *     (the reference drives training from a single host thread, src/testbed.cu:2776).
 */
#ifndef RNB_B200_H
#define RNB_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNB_ABI_VERSION 1

typedef enum rnb_status {
	RNB_OK = 0,
	RNB_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
	RNB_ERR_CUDA = -2,        /* a CUDA runtime call failed */
	RNB_ERR_STATE = -3,       /* call order violated (e.g. training without a dataset) */
	RNB_ERR_NOMEM = -4
} rnb_status;

typedef struct rnb_ctx rnb_ctx;

/* Network / optimizer / sampling configuration: the values of configs/nerf/base.json (:5-80) plus the Testbed
 * constants that reach the step (include/neural-graphics-primitives/testbed.h:237,550,633; src/testbed.cu:2256). */
typedef struct rnb_config {
	uint32_t abi_version;           /* RNB_ABI_VERSION */
	/* encoding: configs/nerf/base.json:30-40, src/testbed.cu:2300-2325 */
	uint32_t n_levels;              /* 14 */
	uint32_t log2_hashmap_size;     /* 19 */
	uint32_t base_resolution;       /* 16 */
	float    per_level_scale;       /* <=0: derive from top_resolution like src/testbed.cu:2321 */
	float    top_resolution;        /* 2048 */
	float    base_valid_level_scale;/* 0.2  (progressive levels, grid.h:1430-1437) */
	float    valid_level_scale;     /* 0.02 */
	uint32_t base_training_step;    /* 100 */
	/* MLPs: base.json:41-48,64-70 — FullyFusedMLP, ReLU, no output activation, no bias */
	uint32_t sdf_n_neurons;         /* 64 (32 supported) */
	uint32_t sdf_n_hidden_layers;   /* 1 */
	uint32_t rgb_n_neurons;         /* 64 (32 supported) */
	uint32_t rgb_n_hidden_layers;   /* 2 (1 supported) */
	float    sdf_bias;              /* -0.1 */
	/* optimizer: base.json:5-29 */
	float    learning_rate;         /* 1e-3 */
	float    beta1, beta2, epsilon; /* 0.9, 0.99, 1e-15 */
	float    l2_reg;                /* 1e-6 (MLP weights only) */
	float    ema_decay;             /* 0.95 */
	uint32_t lr_decay_start;        /* 20000 */
	uint32_t lr_decay_interval;     /* 10000 */
	float    lr_decay_base;         /* 0.33 */
	float    loss_scale;            /* 128 (testbed.h:237) */
	/* sampling / batching */
	uint32_t target_batch_size;     /* 1<<18 (src/testbed.cu:2256) */
	uint32_t rays_per_batch;        /* 4096 initial (testbed.h:633) */
	uint32_t pin_rays_per_batch;    /* 1 = benchmark knob: disable the controller of testbed_nerf.cu:3554-3555 */
	uint32_t seed;                  /* 1337 (testbed.h:550) */
	float    density_grid_decay;    /* 0.95 (testbed.h:671) */
	/* data parallelism: rays i with i % world_size == rank are processed by this ctx */
	uint32_t world_size;            /* 1 */
	uint32_t rank;                  /* 0 */
} rnb_config;

/* Loss / shading switches: Testbed::m_apply_* etc. (testbed.h:490-521), passed by value to
 * compute_loss_kernel_train_nerf (src/testbed_nerf.cu:4030-4039). */
typedef struct rnb_flags {
	int32_t apply_L2;               /* 1 unless --lone */
	int32_t apply_supernormal;      /* --supernormal: identity light basis */
	int32_t apply_rgbplus;          /* 1 unless --no-rgbplus */
	int32_t apply_relu;
	int32_t apply_bce;
	int32_t light_opti;             /* --opti-lights */
	int32_t no_albedo;              /* --no-albedo */
	float   mask_loss_weight;       /* 1.0 */
	float   ek_loss_weight;         /* 0.01 */
	float   cos_anneal_ratio;       /* 1.0 (anneal_end == 0) */
	int32_t light_mode;             /* -1: hashed per (ray, step) [the reference uses curand_init(clock64())]; -2: ray index % 3; 0..2 pinned */
	int32_t only_sdf_training;      /* Optimizer::only_sdf_training (src/testbed.cu:1886-1895) */
} rnb_flags;

/* One training view = TrainingImageMetadata + TrainingXForm (nerf_loader.h:32-49, common.h:171-174).
 * Pixels are uint16 RGBA, 8 bytes per pixel, as uploaded by src/nerf_loader.cu and read by read_rgba
 * (common_device.cuh:665-700). */
typedef struct rnb_view {
	const void* normal_px;          /* device (rnb_set_dataset) or host (rnb_upload_dataset) pointer */
	const void* albedo_px;          /* may be NULL: alpha of the normal map is used as the colour mask */
	int32_t w, h;
	float fx, fy;                   /* focal length in pixels */
	float cx, cy;                   /* principal point / resolution */
	float xform[12];                /* 3x4 camera-to-world in the NGP frame, column-major (nerf_matrix_to_ngp applied) */
} rnb_view;

/* Per-step read-back: Counters::update_after_training (src/testbed_nerf.cu:3532-3558) and m_loss_scalar & co. */
typedef struct rnb_step_stats {
	float    loss, ek_loss, mask_loss;
	uint32_t n_rays;                /* rays_per_batch used by this step (global, all ranks) */
	uint32_t n_rays_kept;           /* rays that produced samples (this rank) */
	uint32_t n_samples;             /* samples before compaction  = numsteps_counter */
	uint32_t n_samples_compacted;   /* samples after compaction   = numsteps_counter_compacted */
	uint32_t n_samples_trained;     /* min(compacted, target_batch_size) */
	uint32_t rays_per_batch_next;
	uint32_t training_step;         /* value after the step */
	uint32_t density_grid_updated;  /* 1 if training_prep_nerf ran inside this call */
} rnb_step_stats;

/* error reporting: every entry point returns an RNB_* code; the message of the last failure on the calling thread is kept here.  Replaces the reference's
 * CUDA_CHECK_THROW -> std::runtime_error (tiny-cuda-nn/common.h:202-208): the caller converts a non-zero code into its own throw (shim: rnb_shim::check). */
const char* rnb_last_error(void);
uint32_t    rnb_abi_version(void);

/* replaces Testbed::reset_network (src/testbed.cu:2220-2485): builds encoding tables, allocates
 * [fp32 | fp16 | fp16 ema | grads | adam state] (trainer.h:72-109) and all step scratch on the current device. */
int rnb_create(const rnb_config* cfg, rnb_ctx** out);
int rnb_destroy(rnb_ctx* ctx);
void rnb_default_config(rnb_config* cfg);
void rnb_default_flags(rnb_flags* f);

/* parameter layout (same order as NerfNetwork::set_params, nerf_network.h:539-583):
 * [sdf mlp | rgb mlp | hash grid | variance(4)].  out[0..4] = off_sdf, off_rgb, off_grid, off_var, n_params */
int rnb_param_layout(rnb_ctx* ctx, uint64_t out[5]);

/* replaces Trainer::initialize_params + NerfNetwork::initialize_params (trainer.h:72-109, nerf_network.h:625-694).
 * sdf_init_host: contents of utils/mlp_weights*.txt (geometric initialisation, nerf_network.h:585-623) or NULL for
 * the built-in sphere initialisation. */
int rnb_init_params(rnb_ctx* ctx, const float* sdf_init_host, size_t n_sdf_init);
int rnb_set_params_fp32(rnb_ctx* ctx, const float* params_host, size_t n);
int rnb_get_params_fp32(rnb_ctx* ctx, float* params_host, size_t n);

/* snapshot hand-off: Trainer::serialize / deserialize (trainer.h:263-304), src/testbed.cu:3280-3390.
 * fp16 views use IEEE binary16 bit patterns. use_ema selects the inference (EMA) copy. */
int rnb_export_params_fp16(rnb_ctx* ctx, uint16_t* host, size_t n, int use_ema);
int rnb_import_params_fp16(rnb_ctx* ctx, const uint16_t* host, size_t n);
int rnb_export_density_grid(rnb_ctx* ctx, float* host, size_t n /* 128^3 */, uint32_t* ema_step);
int rnb_import_density_grid(rnb_ctx* ctx, const float* host, size_t n, uint32_t ema_step);
int rnb_get_bitfield(rnb_ctx* ctx, uint8_t* host, size_t n /* 128^3 (8 mips x 128^3 bits) */);
int rnb_set_bitfield(rnb_ctx* ctx, const uint8_t* host, size_t n);
int rnb_get_train_state(rnb_ctx* ctx, uint32_t out[4] /* training_step, rays_per_batch, n_rays_total, measured_before_compaction */);
int rnb_set_train_state(rnb_ctx* ctx, uint32_t training_step, uint32_t rays_per_batch, uint32_t n_rays_total, uint32_t measured_before_compaction);
/* Testbed::load_snapshot restores m_training_step but neither m_canonical_training_step (testbed.h:907; reset_network zeroes it, src/testbed.cu:2451)
 * nor Training::n_images_for_training_prev (testbed.h:578).  Both steer the occupancy refresh: cadence and bootstrap mode follow the canonical step
 * (src/testbed.cu:2805-2806, src/testbed_nerf.cu:4133), and a refresh that sees an image count different from the previous one starts from an EMPTY
 * grid (:3446-3452).  A stage-2 process of run_pipeline.py therefore discards the snapshot's grid at its first step and rebuilds it from one full sweep.
 * rnb_set_train_state leaves canonical == training_step (a state in the middle of a run); a caller that mirrors load_snapshot passes the reference's
 * own two values here afterwards (shim: pull_state).  n_images_prev == 0xffffffff: adopt the current dataset (never empty the grid on that account). */
int rnb_set_canonical_state(rnb_ctx* ctx, uint32_t canonical_training_step, uint32_t n_images_prev);
int rnb_get_rng(rnb_ctx* ctx, uint64_t out[4] /* m_rng state, inc, density_grid_rng state, inc */);
int rnb_set_rng(rnb_ctx* ctx, const uint64_t in[4]);

/* replaces Testbed::load_nerf's device-side result (metadata_normal_gpu / metadata_albedo_gpu / transforms_gpu,
 * testbed.h:591-592).  rnb_set_dataset: pixel pointers are device pointers owned by the caller;
 * rnb_upload_dataset: pixel pointers are host pointers, copied into device memory owned by ctx. */
int rnb_set_dataset(rnb_ctx* ctx, const rnb_view* views_host, uint32_t n_views);
int rnb_upload_dataset(rnb_ctx* ctx, const rnb_view* views_host, uint32_t n_views);
/* ---- dataset ingest: the image half of load_nerf (src/nerf_loader.cu:556-760) ----
 * rnb_load_png_rgba16 replaces stbi_load_16(path, &w, &h, &comp, 4) (:612, :653): any non-interlaced PNG (grey / RGB / palette / +alpha,
 * 8 or 16 bit, tRNS keys) as 16-bit RGBA in host memory (8-bit samples widened to v * 257, missing alpha 65535); free with rnb_free_host.
 * rnb_load_dataset_images decodes all normal / albedo maps on `threads` host threads (0 = all cores) into pinned staging, uploads them
 * as they finish and installs the views (meta[i]: intrinsics + camera matrix as for rnb_upload_dataset; pixel pointers ignored). */
int  rnb_load_png_rgba16(const char* path, uint32_t* w, uint32_t* h, uint16_t** pixels_host);
void rnb_free_host(void* p);
int  rnb_load_dataset_images(rnb_ctx* ctx, const rnb_view* meta, uint32_t n_views, const char* const* normal_paths, const char* const* albedo_paths /* NULL or entries NULL: no albedo */,
                             uint32_t threads, void* stream);
/* replaces the Testbed members the CLI sets (testbed.h:503-521) and train_nerf_step passes by value to the loss kernel every step (src/testbed_nerf.cu:4030-4039) */
int rnb_set_flags(rnb_ctx* ctx, const rnb_flags* flags);

/* replaces Testbed::training_prep_nerf (src/testbed_nerf.cu:4125-4138): one occupancy-grid refresh. */
int rnb_prep(rnb_ctx* ctx, void* stream);
/* replaces Testbed::train_nerf (src/testbed_nerf.cu:3560-3668): generate samples, forward, loss, backward,
 * optimizer step, counters.  stats may be NULL: with a pinned batch size the call then returns without any host synchronisation (the clamp
 * of the next step's sample budget stays on the device); the adaptive controller needs the compacted count on the host and waits for it,
 * as the reference does after every step (src/testbed_nerf.cu:3535-3551, src/testbed.cu:2866). */
int rnb_train_step(rnb_ctx* ctx, void* stream, rnb_step_stats* stats);
/* replaces Testbed::train (src/testbed.cu:2776-2872): progressive-level update, prep cadence, train_nerf. */
int rnb_train(rnb_ctx* ctx, void* stream, rnb_step_stats* stats);

/* data-parallel split of rnb_train_step: [begin: everything up to and including backward] -> the caller all-reduces
 * rnb_grad_buffer (sum, fp32, n_params elements) and rnb_stat_buffer (sum, 8 floats: loss sums and sample counts) over its communicator
 * -> [end: optimizer + controller].  The reference has no collective; this is where one goes (trainer.h:78-84). */
int rnb_train_step_begin(rnb_ctx* ctx, void* stream);
int rnb_train_step_end(rnb_ctx* ctx, void* stream, rnb_step_stats* stats);
int rnb_grad_buffer(rnb_ctx* ctx, float** grads_dev, uint64_t* n);
/* Data parallelism behind the boundary (new; the reference is single-GPU, its gradient buffer between backward and optimizer_step is one contiguous
 * binary16 array: trainer.h:78-84, src/testbed_nerf.cu:4068 -> :3624).  With a communicator installed and rnb_config.world_size > 1,
 * rnb_train_step / rnb_train do the exchange themselves on the caller's stream: the fp32 accumulators are rounded to binary16 once (21 MB at the
 * default configuration), reduce-scattered (grouped with the 8 floats of loss sums / counts), Adam / EMA run on this rank's 1 / world of the parameters
 * and the binary16 training parameters are all-gathered (sharded optimizer, the measured default; the EMA copy is per shard until rnb_comm_sync_ema).
 * Environment RNB_DP=allreduce selects ONE ncclAllReduce of the binary16 buffer with replicated Adam / EMA instead.
 * NCCL is opened at run time (dlopen of libnccl.so.2), so single-GPU users have no dependency on it.
 *   rnb_comm_unique_id: ncclGetUniqueId on one rank; the caller's launcher carries the 128 bytes to the other ranks (file, socket, MPI, torchrun store)
 *   rnb_comm_init:      ncclCommInitRank(world_size, id, rank) on the current device, communicator owned by the context
 *   rnb_comm_adopt:     use a caller-owned ncclComm_t (its size and rank must equal rnb_config's)
 *   rnb_comm_info:      out = { communicator installed, NCCL version code, bit 0: sharded optimizer | bit 1: one sample order, world_size }
 * One sample order: with a communicator installed the ranks also exchange two per-ray prefix tables per step (ncclAllGather, 4 bytes per ray), so that the
 * slot guard, the 2^18-sample truncation and the roll-over multiplicity are those of the batch ONE GPU would have built from the same rays: the data-parallel
 * run reproduces the single-GPU run.  Environment RNB_DP_EXACT=0 (read by rnb_comm_init / rnb_comm_adopt), or driving
 * rnb_train_step_begin / _end with an external collective and no communicator, applies those rules per rank against target / world instead (3-6 % faster, the
 * trajectories drift apart by a few per cent; with the adaptive controller the budget is met, not exceeded). */
#define RNB_COMM_ID_BYTES 128
int rnb_comm_unique_id(uint8_t id_out[RNB_COMM_ID_BYTES]);
int rnb_comm_init(rnb_ctx* ctx, const uint8_t id[RNB_COMM_ID_BYTES]);
int rnb_comm_adopt(rnb_ctx* ctx, void* nccl_comm);
int rnb_comm_destroy(rnb_ctx* ctx);
int rnb_comm_info(rnb_ctx* ctx, uint32_t out[4]);
/* sharded optimizer only (collective; no-op otherwise): gather the per-shard EMA (inference) parameters before rnb_export_params_fp16(use_ema) / a mesh extraction */
int rnb_comm_sync_ema(rnb_ctx* ctx, void* stream);
int rnb_stat_buffer(rnb_ctx* ctx, float** stats_dev, uint64_t* n);
/* Sharded optimizer for data parallelism (new; the reference is single-GPU): instead of all-reducing the fp32 gradient buffer and running
 * Adam/EMA on every rank, rank r owns parameters [begin, end): the caller REDUCE-SCATTERS the gradient buffer (rnb_grad_buffer; the arrays
 * are allocated and zero up to n_padded, a multiple of 512, so that equal shards exist for every power-of-two world), rnb_train_step_end
 * updates only the shard (fp32 master weights and Adam moments of the other shards are never touched on this rank) and clears the
 * whole gradient buffer, and the caller ALL-GATHERS the binary16 training parameters (params_fp16_dev, n_padded elements).  The EMA copy
 * is maintained per shard: all-gather ema_fp16_dev once before exporting a snapshot / extracting a mesh.
 * reduced_grads_dev: where the reduced shard lives (end - begin floats; NULL = in place in the gradient buffer).  end == 0 switches back. */
int rnb_param_buffers(rnb_ctx* ctx, void** params_fp16_dev, void** ema_fp16_dev, uint64_t* n_params, uint64_t* n_padded);
int rnb_set_optimizer_shard(rnb_ctx* ctx, uint64_t begin, uint64_t end, const float* reduced_grads_dev);
/* host copies for parity checks: gradient accumulators (fp32, loss-scaled; the reference's Trainer::param_gradients(),
 * trainer.h:240, is the fp16 equivalent) and, per kept ray, its index and (loss, ek_loss, mask_loss) as written by
 * compute_loss_kernel_train_nerf (src/testbed_nerf.cu:1833-1841) */
int rnb_get_grads_fp32(rnb_ctx* ctx, float* host, size_t n);
int rnb_get_ray_losses(rnb_ctx* ctx, uint32_t cap, uint32_t* ray_idx_host, float* loss3_host, uint32_t* n_out);
int rnb_get_ray_counts(rnb_ctx* ctx, uint32_t cap, uint32_t* marched_host, uint32_t* kept_host, uint32_t* n_out);

/* in-memory checkpoint / resume of the complete training state (one device-side slot): parameters (fp32 master, fp16, EMA), Adam
 * moments and per-parameter step counters (tcnn adam.h), density grid + bitfield, both pcg32 streams, controller counters.
 * The reference's file snapshot (src/testbed.cu:3280-3390) is served by rnb_export_* / rnb_import_* instead. */
int rnb_checkpoint_save(rnb_ctx* ctx);
int rnb_checkpoint_restore(rnb_ctx* ctx);

/* instrumentation for bench.py: per-stage CUDA-event timing (events recorded on the caller's stream around each stage of
 * the step; resolved at the end of rnb_train_step_end) and a count of kernels launched by this ctx.  The reference's counterpart are the wall-clock
 * EMAs m_training_prep_ms / m_training_ms (testbed.h:863-864, fed by ScopeGuards at src/testbed.cu:2807-2810,2853-2856). */
int rnb_profile_enable(rnb_ctx* ctx, int on);
int rnb_profile_read(rnb_ctx* ctx, char* names_buf, size_t names_cap, double* ms, uint64_t* calls, uint32_t* n);
int rnb_launch_count(rnb_ctx* ctx, uint64_t* out);

/* replaces NerfNetwork::sdf / density (nerf_network.h:454-537) used by marching cubes (src/testbed_nerf.cu:4252)
 * and the grid refresh.  xyz_dev: n x 3 floats in [0,1]^3; outputs may be NULL. */
int rnb_eval_sdf(rnb_ctx* ctx, const float* xyz_dev, size_t n, float* sdf_dev, float* normal_dev, float* density_dev, int use_ema, void* stream);
/* replaces Testbed::get_density_on_grid (src/testbed_nerf.cu:4218-4269), the marching-cubes sweep: SDF (incl. bias, fp32) at the lattice
 * points idx / res * (aabb_max - aabb_min) + aabb_min, out_dev[x + y*res[0] + z*res[0]*res[1]]; positions are generated in the kernel */
int rnb_sdf_on_grid(rnb_ctx* ctx, const uint32_t res[3], const float aabb_min[3], const float aabb_max[3], float* out_dev, int use_ema, void* stream);

/* ---- mesh extraction and export ----------------------------------------------------------------
 * The mesh lives in device memory owned by the context (MeshState verts / vert_normals / vert_colors / indices,
 * include/neural-graphics-primitives/testbed.h:418-447).  Vertices are numbered in lattice order (point x + y rx + z rx ry, its
 * +x, +y, +z edge), triangles in cell order: the reference's mesh up to the permutation its atomicAdd slot hand-out picks. */
typedef struct rnb_mesh_info {
	uint32_t n_verts;          /* vertices on the iso-surface */
	uint32_t n_verts_padded;   /* array length: rounded up to 128, padding zero (src/marching_cubes.cu:810-812); this is what gets saved */
	uint32_t n_indices;        /* 3 x triangles */
	uint32_t res[3];           /* lattice actually used */
	float stage_ms[4];         /* device time of the last extraction (CUDA events): SDF sweep, sign bits + count + scan, vertices + normals + faces, colours */
} rnb_mesh_info;
/* replaces Testbed::marching_cubes (src/testbed_nerf.cu:4297-4348): res rounded up to multiples of 16, SDF sweep over the aabb,
 * marching_cubes_gpu, area-weighted vertex normals (compute_mesh_1ring), vertex colours (compute_mesh_vertex_colors).
 * Synchronises the stream (the vertex count sizes the arrays, as in the reference). */
int rnb_marching_cubes(rnb_ctx* ctx, const uint32_t res[3], const float aabb_min[3], const float aabb_max[3], float thresh, int use_ema, void* stream, rnb_mesh_info* info);
/* replaces marching_cubes_gpu (src/marching_cubes.cu:794-822) + compute_mesh_1ring (:722-728) on a caller-provided lattice of values
 * density_dev[x + y res[0] + z res[0] res[1]] (res[0] a multiple of 16, as every lattice the reference builds); with_colors != 0 also runs the colour network at the vertices */
int rnb_marching_cubes_from_density(rnb_ctx* ctx, const float* density_dev, const uint32_t res[3], const float aabb_min[3], const float aabb_max[3], float thresh,
                                    int with_colors, int use_ema, void* stream, rnb_mesh_info* info);
/* device pointers of the current mesh = MeshState::verts / vert_normals / vert_colors / indices (testbed.h:418-447); valid until the next extraction or
 * rnb_destroy; any out pointer may be NULL */
int rnb_mesh_buffers(rnb_ctx* ctx, float** verts_dev, float** normals_dev, float** colors_dev, uint32_t** indices_dev, rnb_mesh_info* info);
/* replaces the copies of Testbed::compute_marching_cubes_mesh (src/python_api.cu:99-130): host arrays of n_verts_padded x 3 floats and
 * n_indices uint32; any pointer may be NULL */
int rnb_mesh_download(rnb_ctx* ctx, float* verts_host, float* normals_host, float* colors_host, uint32_t* indices_host);
/* replaces save_mesh (src/marching_cubes.cu:824-982) for device arrays: Wavefront OBJ ("v x y z r g b", "vn", "f a//a b//b c//c"; the
 * unwrap/texture variant is not provided) or ASCII PLY when the path ends in "ply".  Byte-identical to the reference's fprintf output
 * for the same arrays: p = n2w_s * ((v - nerf_offset) / nerf_scale) + n2w_t, "%0.5f" / "%0.3f", normals normalised, triangles reversed
 * unless invert_normals.  The text is formatted on the GPU. */
int rnb_save_mesh(const float* verts_dev, const float* normals_dev, const float* colors_dev, const uint32_t* indices_dev, uint32_t n_verts, uint32_t n_indices,
                  const char* path, float nerf_scale, const float nerf_offset[3], float n2w_s, const float n2w_t[3], int invert_normals, void* stream, uint64_t* bytes_written);

/* ---- ray / mesh queries of the albedo-scaling stage --------------------------------------------------
 * Replace the two trimesh calls of compute_albedo_scale_ratios (rnb_neus2/albedo_scaling.py:285-289 and :316-329): a triangle mesh
 * (host arrays, e.g. the stage-1 mesh read back from its OBJ or from rnb_mesh_download) is binned into a uniform cell grid once;
 * rays are traced on the GPU in binary64.  There is no CPU path: rnb_raymesh_create fails with RNB_ERR_CUDA without a device. */
typedef struct rnb_raymesh rnb_raymesh;
/* grid_res: cells along the longest axis of the mesh box, 0 = chosen from the triangle count */
int rnb_raymesh_create(const float* verts_host /* n_verts x 3 */, uint32_t n_verts, const uint32_t* indices_host /* n_tris x 3 */, uint32_t n_tris, uint32_t grid_res, rnb_raymesh** out);
int rnb_raymesh_destroy(rnb_raymesh* rm);
int rnb_raymesh_info(rnb_raymesh* rm, uint32_t res_out[3], uint64_t* n_cell_refs);
/* t_max_host == NULL: closest intersection with t > 0 (mesh.ray.intersects_location(..., multiple_hits=False)):
 *     t_out[i] = ray parameter (distance when the direction has unit length) or +inf, tri_out[i] = triangle or 0xffffffff.
 * t_max_host != NULL: occlusion test, is there ANY intersection with 0 < t < t_max[i] (what :321-329 derives from the multiple_hits=True
 *     list): tri_out[i] = a blocking triangle or 0xffffffff, t_out[i] = its parameter or +inf.
 * origins / dirs: n x 3 binary64, host memory; results are complete when the call returns. */
int rnb_raymesh_intersect(rnb_raymesh* rm, const double* origins_host, const double* dirs_host, const double* t_max_host, uint32_t n,
                          double* t_out_host, uint32_t* tri_out_host, void* stream);

/* ---- stage-level entry points (host buffers; parity tests and micro-benchmarks) --------------------------------
 * Each mirrors one reference kernel / call; see DESIGN.md for the mapping. */
/* generate_training_samples_nerf (src/testbed_nerf.cu:1216-1387) */
int rnb_stage_generate(rnb_ctx* ctx, uint32_t n_rays, uint32_t n_rays_total, uint32_t max_samples,
                       uint32_t* ray_indices_host, float* rays_host /*6/ray*/, uint32_t* numsteps_host /*2/ray*/, float* coords_host /*7/sample*/, uint32_t counters_host[2]);
/* NerfNetwork::forward_impl / inference_mixed_precision (nerf_network.h:87-253): coords 7 floats/sample -> 16 binary16 as float */
int rnb_stage_forward(rnb_ctx* ctx, const float* coords_host, size_t n, int use_ema, float* out16_host, float* normal_host);
/* compute_loss_kernel_train_nerf (src/testbed_nerf.cu:1396-2097) on the samples produced by rnb_stage_generate */
int rnb_stage_loss(rnb_ctx* ctx, const float* out16_compacted_host, const uint32_t* ray_indices_host, const uint32_t* n_fwd, const uint32_t* cbase, const uint32_t* n_emit,
                   uint32_t n_kept, uint32_t n_rays, uint32_t n_rays_total, float* dout16_host, float* loss_host, float* ek_host, float* mask_host);
/* NerfNetwork::forward + backward (nerf_network.h:257-452): gradients (x loss scale) into the fp32 gradient buffer */
int rnb_stage_backward(rnb_ctx* ctx, const float* coords_host, const float* dout16_host, size_t n, uint32_t n_in_rollover, float* grads_host);
/* Trainer::optimizer_step (trainer.h:170; adam.h:51-202, ema.h:116-152, exponential_decay.h:61-72) on given gradients */
int rnb_stage_optimizer(rnb_ctx* ctx, const float* grads_host /* NULL: use the device gradient buffer */);

#ifdef __cplusplus
}
#endif
#endif /* RNB_B200_H */
