// rnb_raymesh.cuh — ray / triangle-mesh intersection over a uniform cell grid (the part of the albedo-scaling stage that the
// reference hands to trimesh: rnb_neus2/albedo_scaling.py:285-289 first hit, :316-329 occlusion test).
// The same source is compiled for the device (rnb_raymesh.cu) and, for the CPU check of the traversal logic, for the host
// (tests/cuda/raymesh_host.cpp).  Arithmetic is binary64 like trimesh's; vertices are stored as binary32 (mesh files carry ~7 digits).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RNB_HD __host__ __device__ __forceinline__
#else
#define RNB_HD inline
#endif

namespace rnb { namespace raymesh {

constexpr uint32_t NO_TRI = 0xffffffffu;
constexpr double BARY_EPS = 1e-9;        // barycentric slack: rays through a shared edge hit one of the two triangles, not neither

// Device/host view of the acceleration structure.  Cell (ix, iy, iz) -> id = ix + res[0] * (iy + res[1] * iz); the triangles whose
// (slightly inflated) bounding box overlaps the cell are cell_tris[cell_start[id] .. cell_start[id + 1]).
struct GridView {
	double bmin[3], cell[3], inv_cell[3];
	int32_t res[3];
	const uint32_t* cell_start;
	const uint32_t* cell_tris;
	const float* tri_verts;              // 9 floats per triangle: v0 v1 v2
	uint32_t n_tris;
};

// Moeller-Trumbore, two-sided.  Returns true and the ray parameter t (in units of |d|) when the line o + t d crosses the triangle.
RNB_HD bool tri_param(const float* __restrict__ T, const double o[3], const double d[3], double& t) {
	const double v0x = T[0], v0y = T[1], v0z = T[2];
	const double e1x = (double)T[3] - v0x, e1y = (double)T[4] - v0y, e1z = (double)T[5] - v0z;
	const double e2x = (double)T[6] - v0x, e2y = (double)T[7] - v0y, e2z = (double)T[8] - v0z;
	const double px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
	const double det = e1x * px + e1y * py + e1z * pz;
	if (fabs(det) < 1e-300) return false;
	const double inv = 1.0 / det;
	const double sx = o[0] - v0x, sy = o[1] - v0y, sz = o[2] - v0z;
	const double u = (sx * px + sy * py + sz * pz) * inv;
	if (u < -BARY_EPS || u > 1.0 + BARY_EPS) return false;
	const double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
	const double v = (d[0] * qx + d[1] * qy + d[2] * qz) * inv;
	if (v < -BARY_EPS || u + v > 1.0 + BARY_EPS) return false;
	t = (e2x * qx + e2y * qy + e2z * qz) * inv;
	return true;
}

// Walk the cells pierced by the ray (3D DDA) between t_min and t_max.
//  ANY == false: closest intersection with t_min < t < t_max  -> t_hit / tri_hit
//  ANY == true : stops at the first intersection found in that interval (occlusion test)
template <bool ANY>
RNB_HD bool trace(const GridView& G, const double o[3], const double d[3], double t_min, double t_max, double& t_hit, uint32_t& tri_hit) {
	t_hit = INFINITY; tri_hit = NO_TRI;
	// clip against the grid box
	double tn = t_min, tf = t_max;
	for (int a = 0; a < 3; ++a) {
		const double lo = G.bmin[a], hi = G.bmin[a] + G.cell[a] * G.res[a];
		if (d[a] == 0.0) { if (o[a] < lo || o[a] > hi) return false; continue; }
		double t0 = (lo - o[a]) / d[a], t1 = (hi - o[a]) / d[a];
		if (t0 > t1) { const double s = t0; t0 = t1; t1 = s; }
		tn = fmax(tn, t0); tf = fmin(tf, t1);
	}
	if (!(tn <= tf)) return false;
	int32_t ix[3], step[3]; double t_next[3], t_delta[3];
	for (int a = 0; a < 3; ++a) {
		const double p = o[a] + d[a] * tn;
		int32_t i = (int32_t)floor((p - G.bmin[a]) * G.inv_cell[a]);
		i = i < 0 ? 0 : (i >= G.res[a] ? G.res[a] - 1 : i);
		ix[a] = i;
		if (d[a] > 0.0) { step[a] = 1; t_next[a] = (G.bmin[a] + (i + 1) * G.cell[a] - o[a]) / d[a]; t_delta[a] = G.cell[a] / d[a]; }
		else if (d[a] < 0.0) { step[a] = -1; t_next[a] = (G.bmin[a] + i * G.cell[a] - o[a]) / d[a]; t_delta[a] = -G.cell[a] / d[a]; }
		else { step[a] = 0; t_next[a] = INFINITY; t_delta[a] = INFINITY; }
	}
	const int32_t max_iter = G.res[0] + G.res[1] + G.res[2] + 3;       // a straight line crosses at most that many cells
	for (int32_t it = 0; it < max_iter; ++it) {
		const uint32_t id = (uint32_t)ix[0] + (uint32_t)G.res[0] * ((uint32_t)ix[1] + (uint32_t)G.res[1] * (uint32_t)ix[2]);
		const uint32_t k0 = G.cell_start[id], k1 = G.cell_start[id + 1];
		for (uint32_t k = k0; k < k1; ++k) {
			const uint32_t tri = G.cell_tris[k];
			double t;
			if (tri_param(G.tri_verts + (size_t)tri * 9, o, d, t) && t > t_min && t < t_max && t < t_hit) {
				t_hit = t; tri_hit = tri;
				if (ANY) return true;
			}
		}
		const int a = t_next[0] <= t_next[1] ? (t_next[0] <= t_next[2] ? 0 : 2) : (t_next[1] <= t_next[2] ? 1 : 2);
		const double t_exit = t_next[a];
		if (t_hit <= t_exit || t_exit > tf) break;          // the closest hit lies inside the cells walked so far / the ray left the interval
		ix[a] += step[a];
		if (ix[a] < 0 || ix[a] >= G.res[a]) break;
		t_next[a] += t_delta[a];
	}
	return tri_hit != NO_TRI;
}

}} // namespace rnb::raymesh
