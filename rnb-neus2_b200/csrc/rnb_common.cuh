// rnb_common.cuh — shared device/host definitions for the sm_100a NeuS2 training step.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/rnb_b200.h"

namespace rnb {

constexpr uint32_t GRIDSIZE = 128;          // NERF_GRIDSIZE, nerf.h:24
constexpr uint32_t GRID_CELLS = GRIDSIZE * GRIDSIZE * GRIDSIZE;
constexpr uint32_t CASCADES = 8;            // NERF_CASCADES, testbed_nerf.cu:50
constexpr uint32_t MAX_STEPS = 1024;        // NERF_STEPS, testbed_nerf.cu:49
constexpr uint32_t RNG_PER_RAY = 8;         // N_MAX_RANDOM_SAMPLES_PER_RAY, testbed_nerf.cu:58
constexpr float SQRT3 = 1.73205080757f;
constexpr float DT = SQRT3 / 1024.0f;       // STEPSIZE == MIN_CONE_STEPSIZE; constant because cone_angle_constant == 0 (testbed_nerf.cu:3214)
constexpr float MIN_OPTICAL_THICKNESS = 0.1f;
constexpr int MAX_LEVELS = 16;
constexpr int MAX_MLP_LAYERS = 4;

struct LayerDesc { uint32_t rows, cols, off; };

// Everything a kernel needs to know about the model; passed by value (fits in constant bank).
struct ModelDev {
	uint32_t n_levels, n_enc, sdf_in, rgb_in, sdf_width, rgb_width;
	uint32_t n_sdf_layers, n_rgb_layers;        // matrices per MLP (hidden + 1)
	LayerDesc sdf_layers[MAX_MLP_LAYERS], rgb_layers[MAX_MLP_LAYERS];
	uint32_t offsets[MAX_LEVELS + 1];           // hashmap_offset_table (entries, not halfs)
	uint32_t res[MAX_LEVELS];
	float scale[MAX_LEVELS];
	uint32_t off_grid, off_var, n_params;
	uint32_t hashed_mask;                       // bit l: level l is hashed (res^3 > entries), else dense
	uint32_t scatter_pair;                      // 1: x-neighbour corners whose entries are adjacent are reduced with one 16-byte atomic (needs 16-byte aligned level bases)
	uint32_t scatter_agg;                       // levels [0, scatter_agg) use the warp-aggregated scatter in the tcgen05 backward (rnb_encode.cuh: scatter_level_agg)
	float sdf_bias;
};

struct ViewDev {
	const uint2* normal_px; const uint2* albedo_px;
	int32_t w, h; float fx, fy, cx, cy; float xform[12];
};

// ---- pcg32 (PCG-XSH-RR 64/32, public algorithm; the reference vendors tcnn/dependencies/pcg32/pcg32.h) ----
struct Pcg32 {
	uint64_t state, inc;
	static constexpr uint64_t MULT = 0x5851f42d4c957f2dULL;
	__host__ __device__ Pcg32() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
	__host__ __device__ explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
		state = 0u; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint();
	}
	__host__ __device__ uint32_t next_uint() {
		uint64_t old = state;
		state = old * MULT + inc;
		uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t)(old >> 59u);
		return (xs >> rot) | (xs << ((~rot + 1u) & 31));
	}
	__host__ __device__ float next_float() {
		uint32_t u = (next_uint() >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
		return __uint_as_float(u) - 1.0f;
#else
		float f; memcpy(&f, &u, 4); return f - 1.0f;
#endif
	}
	__host__ __device__ void advance(int64_t delta_ = (1ll << 32)) {
		uint64_t cur_mult = MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u, delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta >>= 1;
		}
		state = acc_mult * state + acc_plus;
	}
};

__host__ __device__ inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
__host__ __device__ inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
__host__ __device__ inline uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

// hash-grid cell index (entry index, multiply by 2 for the half offset) — tcnn grid.h:113-148
__device__ __forceinline__ uint32_t grid_entry(uint32_t hsz, uint32_t res, uint32_t x, uint32_t y, uint32_t z) {
	uint32_t stride = 1, index = 0;
	index += x * stride; stride *= res;
	if (stride <= hsz) { index += y * stride; stride *= res; if (stride <= hsz) { index += z * stride; stride *= res; } }
	if (hsz < stride) index = x ^ (y * 2654435761u) ^ (z * 805459861u);
	return index % hsz;
}

__device__ __forceinline__ float hq(float x) { return __half2float(__float2half_rn(x)); }
// tcnn::logistic (common_device.h:52-54): full-precision expf — the reference is not built with --use_fast_math
__device__ __forceinline__ float logisticf(float x) { return 1.0f / (1.0f + expf(-x)); }

// image_idx — testbed_nerf.cu:1194-1214 (uint32 wrap-around intended)
__host__ __device__ inline uint32_t image_idx(uint32_t base, uint32_t n_rays, uint32_t n_rays_total, uint32_t n_images) {
	return (((base + n_rays_total) * n_images) / n_rays) % n_images;
}

__device__ __forceinline__ float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : powf((s + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * powf(l, 0.41666f) - 0.055f; }

// read_rgba — common_device.cuh:665-700 (uint16 RGBA, premultiplied linear rgb)
__device__ __forceinline__ float4 read_rgba(const uint2* px, int w, int h, float x, float y) {
	int ix = max(0, min(w - 1, (int)(x * (float)w)));
	int iy = max(0, min(h - 1, (int)(y * (float)h)));
	uint2 raw = __ldg(&px[(size_t)ix + (size_t)iy * w]);
	if (raw.x == 0x00FF00FFu && raw.y == 0u) return make_float4(-1.f, -1.f, -1.f, -1.f);
	float a = (float)(raw.y >> 16) * (1.0f / 65535.0f);
	return make_float4(srgb_to_linear((float)(raw.x & 0xFFFF) * (1.0f / 65535.0f)) * a,
	                   srgb_to_linear((float)(raw.x >> 16) * (1.0f / 65535.0f)) * a,
	                   srgb_to_linear((float)(raw.y & 0xFFFF) * (1.0f / 65535.0f)) * a, a);
}

// pixel position for a training ray — nerf_random_image_pos_training, testbed_nerf.cu:1171-1192 (no CDF, snap on)
__device__ __forceinline__ float2 pixel_pos(Pcg32& rng, int w, int h) {
	float u = rng.next_float(), v = rng.next_float();
	float px = fminf(fmaxf(u * (float)w, 0.0f), (float)(w - 1));
	float py = fminf(fmaxf(v * (float)h, 0.0f), (float)(h - 1));
	return make_float2((px + 0.5f) / (float)w, (py + 0.5f) / (float)h);
}

__host__ __device__ inline uint32_t hashed_light(uint32_t ray_idx, uint32_t step) {
	uint32_t h = ray_idx * 0x9E3779B1u ^ (step * 0x85EBCA77u) ^ 0xC2B2AE3Du;
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h % 3u;
}

// Data parallel with ONE sample order (rnb_api.cu, dp_exact): xg is the all-gathered table of every rank's inclusive per-position prefix
// (k_prefix_positions; position m of rank q = global ray m * world + q, L positions per rank).  The samples in front of global ray (m, rank) are
// those of positions <= m on the ranks before this one and of positions < m on the ranks behind it, plus this rank's own exclusive prefix.
__device__ __forceinline__ uint32_t foreign_prefix(const uint32_t* __restrict__ xg, uint32_t L, uint32_t world, uint32_t rank, uint32_t m) {
	uint32_t s = 0;
	for (uint32_t q = 0; q < rank; ++q) s += xg[q * L + m];
	if (m) for (uint32_t q = rank + 1; q < world; ++q) s += xg[q * L + m - 1];
	return s;
}
__device__ __forceinline__ uint32_t global_total(const uint32_t* __restrict__ xg, uint32_t L, uint32_t world) {
	uint32_t s = 0;
	for (uint32_t q = 0; q < world; ++q) s += xg[q * L + L - 1];
	return s;
}
__host__ __device__ inline float rollover_weight(uint32_t s, uint32_t n_in, uint32_t n_batch) {
	if (n_in == 0 || n_in >= n_batch) return 1.0f;
	uint32_t c = (n_batch - 1 - s) / n_in;
	return 1.0f + (float)c * ((float)n_in / (float)n_batch);
}

} // namespace rnb
