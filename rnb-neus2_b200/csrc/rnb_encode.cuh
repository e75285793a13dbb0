// rnb_encode.cuh — multiresolution hash-grid gather shared by the SIMT and tensor-core network kernels.
// Semantics: kernel_grid (reference tcnn encodings/grid.h:169-364), pos_fract (common_device.h:415-424).
#pragma once
#include "rnb_common.cuh"

namespace rnb {

struct LevelGeom { float fx, fy, fz; uint32_t gx, gy, gz; };

__device__ __forceinline__ LevelGeom level_geom(float scale, float x, float y, float z) {
	LevelGeom g;
	const float px = fmaf(x, scale, 0.5f), py = fmaf(y, scale, 0.5f), pz = fmaf(z, scale, 0.5f);
	const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
	g.gx = (uint32_t)(int)fx; g.gy = (uint32_t)(int)fy; g.gz = (uint32_t)(int)fz;
	g.fx = px - fx; g.fy = py - fy; g.fz = pz - fz;
	return g;
}

// The 8 corner entries of one cell, corner c = bx + 2 by + 4 bz.  Same values as grid_entry() (tcnn grid.h:113-148)
// without its per-corner branches and runtime modulo:
//  * a hashed level has hsz == 2^log2_hashmap_size, so `% hsz` is a mask and the three products are shared by the corners;
//  * a dense level has index < res^3 + res^2 + res < 2 hsz, so `% hsz` is one conditional subtraction, and it can only
//    trigger when the cell touches the upper face of the grid (coordinate + 1 == res).
__device__ __forceinline__ void corner_entries(bool hashed, uint32_t hsz, uint32_t res, const LevelGeom& g, uint32_t (&e)[8]) {
	if (hashed) {
		const uint32_t m = hsz - 1u;
		const uint32_t x0 = g.gx, x1 = g.gx + 1u;
		const uint32_t y0 = g.gy * 2654435761u, y1 = y0 + 2654435761u;
		const uint32_t z0 = g.gz * 805459861u, z1 = z0 + 805459861u;
		const uint32_t a0 = y0 ^ z0, a1 = y1 ^ z0, a2 = y0 ^ z1, a3 = y1 ^ z1;
		e[0] = (x0 ^ a0) & m; e[1] = (x1 ^ a0) & m; e[2] = (x0 ^ a1) & m; e[3] = (x1 ^ a1) & m;
		e[4] = (x0 ^ a2) & m; e[5] = (x1 ^ a2) & m; e[6] = (x0 ^ a3) & m; e[7] = (x1 ^ a3) & m;
	} else {
		const uint32_t r2 = res * res;
		const uint32_t b = g.gx + g.gy * res + g.gz * r2;
		e[0] = b; e[1] = b + 1u; e[2] = b + res; e[3] = b + res + 1u;
		e[4] = b + r2; e[5] = b + r2 + 1u; e[6] = b + r2 + res; e[7] = b + r2 + res + 1u;
		if (max(g.gx, max(g.gy, g.gz)) + 1u >= res || b >= hsz) {
			#pragma unroll
			for (int c = 0; c < 8; ++c) e[c] = e[c] % hsz;
		}
	}
}

// One level for one sample: the 8 corners are gathered once (the reference re-gathers them per axis); the encoding is
// accumulated in binary16 in corner order (grid.h:291-315), dy/dx in fp32 (grid.h:324-363).
// Returns the two features packed as half2; dy = {d f0/dx, d f0/dy, d f0/dz, d f1/dx, d f1/dy, d f1/dz}.
__device__ __forceinline__ __half2 encode_level_packed(const ModelDev& M, const __half* __restrict__ P, uint32_t l, float x, float y, float z, float* __restrict__ dy) {
	const uint32_t off = M.offsets[l];
	const __half2* grid = reinterpret_cast<const __half2*>(P + M.off_grid);      // 32-bit entry index + one wide multiply-add per corner
	const uint32_t hsz = M.offsets[l + 1] - off, res = M.res[l];
	const float scale = M.scale[l];
	const LevelGeom g = level_geom(scale, x, y, z);
	uint32_t e[8];
	corner_entries((M.hashed_mask >> l) & 1u, hsz, res, g, e);
	float2 v[8];
	#pragma unroll
	for (int c = 0; c < 8; ++c) v[c] = __half22float2(__ldg(grid + (e[c] + off)));
	const float wx[2] = {1.f - g.fx, g.fx}, wy[2] = {1.f - g.fy, g.fy}, wz[2] = {1.f - g.fz, g.fz};
	const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};
	__half2 r = __floats2half2_rn(0.f, 0.f);
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const float w = wxy[c & 3] * wz[c >> 2];
		r = __hadd2(r, __floats2half2_rn(w * v[c].x, w * v[c].y));
	}
	if (dy) {
		const float sx[2] = {scale * wx[0], scale * wx[1]}, sy[2] = {scale * wy[0], scale * wy[1]};
		#pragma unroll
		for (int d = 0; d < 3; ++d) {
			float a0 = 0.f, a1 = 0.f;
			#pragma unroll
			for (int idx = 0; idx < 4; ++idx) {
				int c; float w;
				if (d == 0) { c = (idx & 1) * 2 + (idx >> 1) * 4; w = sy[idx & 1] * wz[idx >> 1]; }
				else if (d == 1) { c = (idx & 1) * 1 + (idx >> 1) * 4; w = sx[idx & 1] * wz[idx >> 1]; }
				else { c = (idx & 1) * 1 + (idx >> 1) * 2; w = sx[idx & 1] * wy[idx >> 1]; }
				const int cr = c | (1 << d);
				a0 += w * (v[cr].x - v[c].x);
				a1 += w * (v[cr].y - v[c].y);
			}
			dy[d] = a0; dy[3 + d] = a1;
		}
	}
	return r;
}

// ---- batched gather -----------------------------------------------------------------------------------------------------
// Memory-level parallelism for the thread-per-sample kernels: the corner loads of LB levels (8 * LB independent 4-byte loads)
// are all issued before the first one is consumed, so one sample pays the L2 round trip once per batch instead of once
// per level.  Arithmetic identical to encode_level_packed().
struct LevelLoads { uint32_t raw[8]; float fx, fy, fz; };

// stab / n_stage_levels (optional): the first n_stage_levels levels of the table staged in shared memory by the CTA (rnb_network_tc.cu: one bulk-async
// copy per CTA); their corner reads are shared-memory loads.  The branch is uniform (it depends on the level only).
__device__ __forceinline__ void level_issue(const ModelDev& M, const __half* __restrict__ P, uint32_t l, float x, float y, float z, LevelLoads& Q,
                                            const uint32_t* __restrict__ stab = nullptr, uint32_t n_stage_levels = 0) {
	const uint32_t off = M.offsets[l];
	const uint32_t* grid = reinterpret_cast<const uint32_t*>(P + M.off_grid);
	const uint32_t hsz = M.offsets[l + 1] - off, res = M.res[l];
	const LevelGeom g = level_geom(M.scale[l], x, y, z);
	uint32_t e[8];
	corner_entries((M.hashed_mask >> l) & 1u, hsz, res, g, e);
	if (l < n_stage_levels) {
		#pragma unroll
		for (int c = 0; c < 8; ++c) Q.raw[c] = stab[e[c] + off];
	} else {
		#pragma unroll
		for (int c = 0; c < 8; ++c) Q.raw[c] = __ldg(grid + (e[c] + off));
	}
	Q.fx = g.fx; Q.fy = g.fy; Q.fz = g.fz;
}

__device__ __forceinline__ __half2 level_finish(float scale, const LevelLoads& Q, float* __restrict__ dy) {
	float2 v[8];
	#pragma unroll
	for (int c = 0; c < 8; ++c) v[c] = __half22float2(*reinterpret_cast<const __half2*>(&Q.raw[c]));
	const float wx[2] = {1.f - Q.fx, Q.fx}, wy[2] = {1.f - Q.fy, Q.fy}, wz[2] = {1.f - Q.fz, Q.fz};
	const float wxy[4] = {wx[0] * wy[0], wx[1] * wy[0], wx[0] * wy[1], wx[1] * wy[1]};
	__half2 r = __floats2half2_rn(0.f, 0.f);
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const float w = wxy[c & 3] * wz[c >> 2];
		r = __hadd2(r, __floats2half2_rn(w * v[c].x, w * v[c].y));
	}
	if (dy) {
		const float sx[2] = {scale * wx[0], scale * wx[1]}, sy[2] = {scale * wy[0], scale * wy[1]};
		#pragma unroll
		for (int d = 0; d < 3; ++d) {
			float a0 = 0.f, a1 = 0.f;
			#pragma unroll
			for (int idx = 0; idx < 4; ++idx) {
				int c; float w;
				if (d == 0) { c = (idx & 1) * 2 + (idx >> 1) * 4; w = sy[idx & 1] * wz[idx >> 1]; }
				else if (d == 1) { c = (idx & 1) * 1 + (idx >> 1) * 4; w = sx[idx & 1] * wz[idx >> 1]; }
				else { c = (idx & 1) * 1 + (idx >> 1) * 2; w = sx[idx & 1] * wy[idx >> 1]; }
				const int cr = c | (1 << d);
				a0 += w * (v[cr].x - v[c].x);
				a1 += w * (v[cr].y - v[c].y);
			}
			dy[d] = a0; dy[3 + d] = a1;
		}
	}
	return r;
}

// Merged first- and second-order gradient scatter for one (sample, level):
//   dgrid[corner] += dL/denc * w_c  +  dsdf/denc * scale * sum_dim gn[dim] * (+-1) * w_c^(dim)
// (kernel_grid_backward grid.h:366-495 and kernel_grid_backward_input_backward_grid grid.h:556-683 hit the same 8 corners).
// scatter_values(): the 8 corner entries and the 8 (2-feature) contributions; `cell` identifies the lattice cell within the level.
__device__ __forceinline__ void scatter_values(const ModelDev& M, uint32_t l, float x, float y, float z, float d10, float d11, float ge0, float ge1,
                                               float gnx, float gny, float gnz, uint32_t (&e)[8], float2 (&v)[8], uint32_t& cell) {
	const uint32_t off = M.offsets[l];
	const uint32_t hsz = M.offsets[l + 1] - off, res = M.res[l];
	const float scale = M.scale[l];
	const LevelGeom g = level_geom(scale, x, y, z);
	corner_entries((M.hashed_mask >> l) & 1u, hsz, res, g, e);
	cell = g.gx | (g.gy << 10) | (g.gz << 20);            // unique for res <= 1024 (the host enables aggregation only there)
	const float wx[2] = {1.f - g.fx, g.fx}, wy[2] = {1.f - g.fy, g.fy}, wz[2] = {1.f - g.fz, g.fz};
	const float sgx = scale * gnx, sgy = scale * gny, sgz = scale * gnz;
	// w1 = wx wy wz ; w2 = (+-sgx) wy wz + (+-sgy) wx wz + (+-sgz) wx wy   (sign = + on the upper corner of that axis)
	float ax[4], ay[4], az[4];   // ax[by + 2 bz], ay[bx + 2 bz], az[bx + 2 by]
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		ax[i] = sgx * (wy[i & 1] * wz[i >> 1]);
		ay[i] = sgy * (wx[i & 1] * wz[i >> 1]);
		az[i] = sgz * (wx[i & 1] * wy[i >> 1]);
	}
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
		const float w1 = wx[bx] * wy[by] * wz[bz];
		const float w2 = (bx ? ax[by + 2 * bz] : -ax[by + 2 * bz]) + (by ? ay[bx + 2 * bz] : -ay[bx + 2 * bz]) + (bz ? az[bx + 2 * by] : -az[bx + 2 * by]);
		v[c] = make_float2(d10 * w1 + ge0 * w2, d11 * w1 + ge1 * w2);
	}
}

// scatter_emit(): the reductions into the fp32 gradient buffer.
__device__ __forceinline__ void scatter_emit(const ModelDev& M, float* __restrict__ G, uint32_t l, const uint32_t (&e)[8], const float2 (&v)[8]) {
	const uint32_t off = M.offsets[l];
	float2* gg = reinterpret_cast<float2*>(G + M.off_grid);
	// The two corners of an x-edge are neighbours in memory whenever their entry indices differ only in bit 0: always in a dense
	// level with an even base index, and in a hashed level whenever the cell's x is even (the x prime is 1, so x ^ (x + 1) == 1).
	// Those pairs go out as ONE 16-byte reduction instead of two 8-byte ones: the backward is bound by the number of atomics.
	#pragma unroll
	for (int c = 0; c < 8; c += 2) {
		const bool z0 = v[c].x == 0.f && v[c].y == 0.f, z1 = v[c + 1].x == 0.f && v[c + 1].y == 0.f;
		if (z0 && z1) continue;
		if (M.scatter_pair && ((e[c] ^ e[c + 1]) == 1u)) {
			const bool lo = e[c] < e[c + 1];
			const float2 a = lo ? v[c] : v[c + 1], b = lo ? v[c + 1] : v[c];
			atomicAdd(reinterpret_cast<float4*>(gg + (min(e[c], e[c + 1]) + off)), make_float4(a.x, a.y, b.x, b.y));
		} else {
			if (!z0) atomicAdd(gg + (e[c] + off), v[c]);
			if (!z1) atomicAdd(gg + (e[c + 1] + off), v[c + 1]);
		}
	}
}

__device__ __forceinline__ void scatter_level(const ModelDev& M, float* __restrict__ G, uint32_t l, float x, float y, float z,
                                              float d10, float d11, float ge0, float ge1, float gnx, float gny, float gnz) {
	uint32_t e[8], cell; float2 v[8];
	scatter_values(M, l, x, y, z, d10, d11, ge0, ge1, gnx, gny, gnz, e, v, cell);
	scatter_emit(M, G, l, e, v);
}

// Warp-aggregated scatter.  Consecutive compacted samples are consecutive lattice points of one ray (step sqrt(3)/1024), so in a
// coarse level a run of adjacent lanes falls into the same cell (26 lanes at res 16, 5 at res 72, 2-3 at res 151) and would send
// its 8 reductions to the same 8 addresses.  A run of adjacent lanes with an identical cell is a segment: the contributions
// are summed with a segmented shuffle reduction (log2(longest run) steps, warp-uniform) and only the segment's first lane issues
// the reductions.  Must be called by all 32 lanes; lanes that are not `live` contribute nothing.
__device__ __forceinline__ void scatter_level_agg(const ModelDev& M, float* __restrict__ G, uint32_t l, bool live, float x, float y, float z,
                                                  float d10, float d11, float ge0, float ge1, float gnx, float gny, float gnz) {
	constexpr uint32_t FULL = 0xffffffffu;
	uint32_t e[8], cell; float2 v[8];
	scatter_values(M, l, x, y, z, d10, d11, ge0, ge1, gnx, gny, gnz, e, v, cell);
	if (!live) {
		cell = 0xffffffffu;
		#pragma unroll
		for (int c = 0; c < 8; ++c) v[c] = make_float2(0.f, 0.f);
	}
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t prev = __shfl_up_sync(FULL, cell, 1);
	const bool head = lane == 0u || prev != cell;
	const uint32_t heads = __ballot_sync(FULL, head);
	if (32 - __popc(heads) >= 4) {                       // warp-uniform: worth it once a few lanes' reductions are saved
		const uint32_t above = lane == 31u ? 0u : (heads >> (lane + 1u));
		const uint32_t seg_last = above ? lane + (uint32_t)__ffs(above) - 1u : 31u;
		const uint32_t run = __reduce_max_sync(FULL, seg_last - lane);        // longest segment - 1
		for (uint32_t d = 1; d <= run; d <<= 1) {        // before the step lane i holds the sum over [i, min(i + d - 1, seg_last)]
			const bool take = lane + d <= seg_last;
			#pragma unroll
			for (int c = 0; c < 8; ++c) {
				const float ox = __shfl_down_sync(FULL, v[c].x, d), oy = __shfl_down_sync(FULL, v[c].y, d);
				if (take) { v[c].x += ox; v[c].y += oy; }
			}
		}
		if (head) scatter_emit(M, G, l, e, v);
	} else {
		scatter_emit(M, G, l, e, v);
	}
}

} // namespace rnb
