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

// One level for one sample: the 8 corners are gathered once (the reference re-gathers them per axis); the encoding is
// accumulated in binary16 in corner order (grid.h:291-315), dy/dx in fp32 (grid.h:324-363).
// Returns the two features packed as half2; dy = {d f0/dx, d f0/dy, d f0/dz, d f1/dx, d f1/dy, d f1/dz}.
__device__ __forceinline__ __half2 encode_level_packed(const ModelDev& M, const __half* __restrict__ P, uint32_t l, float x, float y, float z, float* __restrict__ dy) {
	const __half2* grid = reinterpret_cast<const __half2*>(P + M.off_grid) + M.offsets[l];
	const uint32_t hsz = M.offsets[l + 1] - M.offsets[l], res = M.res[l];
	const float scale = M.scale[l];
	const LevelGeom g = level_geom(scale, x, y, z);
	float2 v[8];
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const uint32_t e = grid_entry(hsz, res, g.gx + (c & 1), g.gy + ((c >> 1) & 1), g.gz + ((c >> 2) & 1));
		v[c] = __half22float2(__ldg(&grid[e]));
	}
	const float wx[2] = {1.f - g.fx, g.fx}, wy[2] = {1.f - g.fy, g.fy}, wz[2] = {1.f - g.fz, g.fz};
	__half r0 = __float2half_rn(0.f), r1 = r0;
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const float w = wx[c & 1] * wy[(c >> 1) & 1] * wz[(c >> 2) & 1];
		r0 = __hadd(r0, __float2half_rn(w * v[c].x));
		r1 = __hadd(r1, __float2half_rn(w * v[c].y));
	}
	if (dy) {
		#pragma unroll
		for (int d = 0; d < 3; ++d) {
			float a0 = 0.f, a1 = 0.f;
			#pragma unroll
			for (int idx = 0; idx < 4; ++idx) {
				int c; float w = scale;
				if (d == 0) { c = (idx & 1) * 2 + (idx >> 1) * 4; w *= wy[idx & 1]; w *= wz[idx >> 1]; }
				else if (d == 1) { c = (idx & 1) * 1 + (idx >> 1) * 4; w *= wx[idx & 1]; w *= wz[idx >> 1]; }
				else { c = (idx & 1) * 1 + (idx >> 1) * 2; w *= wx[idx & 1]; w *= wy[idx >> 1]; }
				const int cr = c | (1 << d);
				a0 += w * (v[cr].x - v[c].x);
				a1 += w * (v[cr].y - v[c].y);
			}
			dy[d] = a0; dy[3 + d] = a1;
		}
	}
	return __halves2half2(r0, r1);
}

// Merged first- and second-order gradient scatter for one (sample, level):
//   dgrid[corner] += dL/denc * w_c  +  dsdf/denc * scale * sum_dim gn[dim] * (+-1) * w_c^(dim)
// (kernel_grid_backward grid.h:366-495 and kernel_grid_backward_input_backward_grid grid.h:556-683 hit the same 8 corners).
__device__ __forceinline__ void scatter_level(const ModelDev& M, float* __restrict__ G, uint32_t l, float x, float y, float z,
                                              float d10, float d11, float ge0, float ge1, float gnx, float gny, float gnz) {
	float* gg = G + M.off_grid + (size_t)M.offsets[l] * 2;
	const uint32_t hsz = M.offsets[l + 1] - M.offsets[l], res = M.res[l];
	const float scale = M.scale[l];
	const LevelGeom g = level_geom(scale, x, y, z);
	const float wx[2] = {1.f - g.fx, g.fx}, wy[2] = {1.f - g.fy, g.fy}, wz[2] = {1.f - g.fz, g.fz};
	#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
		const float w1 = wx[bx] * wy[by] * wz[bz];
		const float w2 = scale * (gnx * (bx ? 1.f : -1.f) * wy[by] * wz[bz] + gny * (by ? 1.f : -1.f) * wx[bx] * wz[bz] + gnz * (bz ? 1.f : -1.f) * wx[bx] * wy[by]);
		const uint32_t e = grid_entry(hsz, res, g.gx + bx, g.gy + by, g.gz + bz);
		const float v0 = d10 * w1 + ge0 * w2, v1 = d11 * w1 + ge1 * w2;
		if (v0 != 0.f || v1 != 0.f) atomicAdd(reinterpret_cast<float2*>(gg + 2 * e), make_float2(v0, v1));
	}
}

} // namespace rnb
