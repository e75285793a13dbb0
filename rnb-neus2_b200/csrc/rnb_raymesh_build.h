// rnb_raymesh_build.h — host-side construction of the uniform cell grid over a triangle mesh (two counting passes, no sorting).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "rnb_raymesh.cuh"

namespace rnb { namespace raymesh {

struct HostGrid {
	GridView view;                       // pointers are filled by the owner (host vectors below or their device copies)
	std::vector<uint32_t> cell_start, cell_tris;
	std::vector<float> tri_verts;
};

// grid_res: cells along the longest box axis (0: chosen from the triangle count so that a cell holds a handful of triangles;
// a surface mesh with n triangles occupies ~n / k cells of an r^3 grid when r ~ sqrt(n / k'))
inline void build_grid(const float* verts, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris, uint32_t grid_res, HostGrid& H) {
	double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
	for (uint32_t i = 0; i < n_verts; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], (double)verts[3 * i + a]); hi[a] = std::max(hi[a], (double)verts[3 * i + a]); }
	double ext = 0.0;
	for (int a = 0; a < 3; ++a) ext = std::max(ext, hi[a] - lo[a]);
	if (!(ext > 0.0)) ext = 1.0;
	const double pad = 1e-3 * ext;
	if (grid_res == 0) grid_res = (uint32_t)std::lround(std::sqrt((double)n_tris / 6.0));
	grid_res = std::min(std::max(grid_res, 4u), 320u);
	GridView& G = H.view;
	G.n_tris = n_tris;
	H.tri_verts.resize((size_t)n_tris * 9);
	for (uint32_t t = 0; t < n_tris; ++t) for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) H.tri_verts[(size_t)t * 9 + 3 * k + a] = verts[3 * (size_t)indices[3 * t + k] + a];
	size_t n_cells = 1;
	auto lay_out = [&](uint32_t r) {
		const double h = (ext + 2 * pad) / r;
		n_cells = 1;
		for (int a = 0; a < 3; ++a) {
			G.bmin[a] = lo[a] - pad;
			G.res[a] = std::max(1, (int)std::ceil((hi[a] - lo[a] + 2 * pad) / h));
			G.cell[a] = h; G.inv_cell[a] = 1.0 / h;
			n_cells *= (size_t)G.res[a];
		}
	};
	lay_out(grid_res);
	// cell range of a triangle: its bounding box, inflated so that hits computed a rounding error outside a cell are still found there
	const double slack = 1e-6 * ext;
	auto range = [&](uint32_t t, int32_t c0[3], int32_t c1[3]) {
		const float* T = &H.tri_verts[(size_t)t * 9];
		for (int a = 0; a < 3; ++a) {
			const double mn = std::min({(double)T[a], (double)T[3 + a], (double)T[6 + a]}) - slack, mx = std::max({(double)T[a], (double)T[3 + a], (double)T[6 + a]}) + slack;
			c0[a] = std::min(std::max((int32_t)std::floor((mn - G.bmin[a]) * G.inv_cell[a]), 0), G.res[a] - 1);
			c1[a] = std::min(std::max((int32_t)std::floor((mx - G.bmin[a]) * G.inv_cell[a]), 0), G.res[a] - 1);
		}
	};
	// a few triangles much larger than a cell (a ground plane next to a scanned object) would be referenced from every cell they span:
	// coarsen the grid until the reference list is at most a small multiple of the triangle count
	for (;;) {
		size_t refs = 0;
		for (uint32_t t = 0; t < n_tris; ++t) {
			int32_t c0[3], c1[3]; range(t, c0, c1);
			refs += (size_t)(c1[0] - c0[0] + 1) * (size_t)(c1[1] - c0[1] + 1) * (size_t)(c1[2] - c0[2] + 1);
		}
		if (refs <= std::max<size_t>((size_t)64 * n_tris, (size_t)1 << 22) || grid_res <= 4) break;
		grid_res = std::max(4u, grid_res / 2);
		lay_out(grid_res);
	}
	H.cell_start.assign(n_cells + 1, 0u);
	for (uint32_t t = 0; t < n_tris; ++t) {
		int32_t c0[3], c1[3]; range(t, c0, c1);
		for (int32_t z = c0[2]; z <= c1[2]; ++z) for (int32_t y = c0[1]; y <= c1[1]; ++y) for (int32_t x = c0[0]; x <= c1[0]; ++x)
			++H.cell_start[(size_t)x + (size_t)G.res[0] * ((size_t)y + (size_t)G.res[1] * (size_t)z) + 1];
	}
	for (size_t i = 0; i < n_cells; ++i) H.cell_start[i + 1] += H.cell_start[i];
	H.cell_tris.resize(H.cell_start[n_cells]);
	std::vector<uint32_t> fill(H.cell_start.begin(), H.cell_start.end() - 1);
	for (uint32_t t = 0; t < n_tris; ++t) {
		int32_t c0[3], c1[3]; range(t, c0, c1);
		for (int32_t z = c0[2]; z <= c1[2]; ++z) for (int32_t y = c0[1]; y <= c1[1]; ++y) for (int32_t x = c0[0]; x <= c1[0]; ++x)
			H.cell_tris[fill[(size_t)x + (size_t)G.res[0] * ((size_t)y + (size_t)G.res[1] * (size_t)z)]++] = t;
	}
	G.cell_start = H.cell_start.data(); G.cell_tris = H.cell_tris.data(); G.tri_verts = H.tri_verts.data();
}

}} // namespace rnb::raymesh
