// TEST INFRASTRUCTURE (oracle) — not part of the product; only tests/, __graft_entry__.smoke() and bench.py's CPU baseline
// may use it.  Sequential CPU restatement of the reference's mesh path:
//   gen_vertices / gen_faces / marching_cubes_gpu   src/marching_cubes.cu:276-330, 377-720, 794-822
//   accumulate_1ring (vertex normals)                src/marching_cubes.cu:332-364
//   save_mesh (OBJ without unwrap, ASCII PLY)        src/marching_cubes.cu:824-982
// The reference hands out vertex and triangle slots with atomicAdd in whatever order its threads retire; here the kernels'
// thread bodies run in lattice order (x fastest), which fixes one of the orders the reference can produce.  Device
// arithmetic that nvcc fuses is written with fmaf(); host arithmetic (save_mesh) is left unfused (-ffp-contract=off) and the
// text is produced by the C library's fprintf, exactly as in the reference.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "orc_mc_tables.h"

namespace orc_mesh {

struct Lattice { uint32_t rx, ry, rz; float s[3], o[3], thresh; };

inline Lattice make_lattice(const uint32_t res[3], const float mn[3], const float mx[3], float thresh) {
	Lattice L; L.rx = res[0]; L.ry = res[1]; L.rz = res[2]; L.thresh = thresh;
	for (int d = 0; d < 3; ++d) { L.s[d] = (mx[d] - mn[d]) / (float)res[d]; L.o[d] = mn[d]; }      // :280-281
	return L;
}

// gen_vertices (:276-330), one lattice point; counter = the atomicAdd target
inline void gen_vertices_point(const Lattice& L, const float* density, uint32_t x, uint32_t y, uint32_t z, int* vertidx_grid, std::vector<float>* verts, uint32_t& counter) {
	const uint64_t res2 = (uint64_t)L.rx * L.ry, res3 = res2 * L.rz, idx = x + (uint64_t)y * L.rx + z * res2;
	const float f0 = density[idx];
	const bool inside = f0 > L.thresh;
	const uint32_t lim[3] = {L.rx, L.ry, L.rz}, p[3] = {x, y, z};
	const uint64_t step[3] = {1, L.rx, res2};
	for (int a = 0; a < 3; ++a) {
		if (p[a] >= lim[a] - 1) continue;
		const float f1 = density[idx + step[a]];
		if (inside == (f1 > L.thresh)) continue;
		const uint32_t vidx = counter++;
		if (!verts) continue;
		vertidx_grid[idx + res3 * a] = (int)vidx + 1;
		const float dt = (L.thresh - f0) / (f1 - f0);
		float l[3] = {(float)x, (float)y, (float)z};
		l[a] = l[a] + dt;
		for (int d = 0; d < 3; ++d) (*verts)[(size_t)vidx * 3 + d] = std::fmaf(l[d], L.s[d], L.o[d]);   // cwiseProduct(scale) + offset, fused by nvcc
	}
}

// gen_faces (:377-720), one cell
inline void gen_faces_cell(const Lattice& L, const float* density, uint32_t x, uint32_t y, uint32_t z, const int* vertidx_grid, std::vector<uint32_t>* indices, uint32_t& counter) {
	if (x >= L.rx - 1 || y >= L.ry - 1 || z >= L.rz - 1) return;
	const uint64_t res1 = L.rx, res2 = (uint64_t)L.rx * L.ry, res3 = res2 * L.rz;
	uint64_t idx = x + (uint64_t)y * res1 + z * res2;
	const uint64_t idx_x = idx, idx_y = idx + res3, idx_z = idx + res3 * 2;
	int mask = 0;
	if (density[idx] > L.thresh) mask |= 1;
	if (density[idx + 1] > L.thresh) mask |= 2;
	if (density[idx + 1 + res1] > L.thresh) mask |= 4;
	if (density[idx + res1] > L.thresh) mask |= 8;
	idx += res2;
	if (density[idx] > L.thresh) mask |= 16;
	if (density[idx + 1] > L.thresh) mask |= 32;
	if (density[idx + 1 + res1] > L.thresh) mask |= 64;
	if (density[idx + res1] > L.thresh) mask |= 128;
	if (!mask || mask == 255) return;
	int local_edges[12] = {0};
	if (vertidx_grid) {
		local_edges[0] = vertidx_grid[idx_x]; local_edges[1] = vertidx_grid[idx_y + 1]; local_edges[2] = vertidx_grid[idx_x + res1]; local_edges[3] = vertidx_grid[idx_y];
		local_edges[4] = vertidx_grid[idx_x + res2]; local_edges[5] = vertidx_grid[idx_y + 1 + res2]; local_edges[6] = vertidx_grid[idx_x + res1 + res2]; local_edges[7] = vertidx_grid[idx_y + res2];
		local_edges[8] = vertidx_grid[idx_z]; local_edges[9] = vertidx_grid[idx_z + 1]; local_edges[10] = vertidx_grid[idx_z + 1 + res1]; local_edges[11] = vertidx_grid[idx_z + res1];
	}
	const int8_t* tri = ORC_MC_TRIANGLES[mask];
	uint32_t tricount = 0;
	for (; tricount < 15; tricount += 3) if (tri[tricount] < 0) break;
	const uint32_t tidx = counter; counter += tricount;
	if (indices) for (int i = 0; i < 15; ++i) { const int j = tri[i]; if (j < 0) break; (*indices)[tidx + i] = (uint32_t)(local_edges[j] - 1); }
}

struct Mesh { std::vector<float> verts, normals; std::vector<uint32_t> indices; uint32_t n_verts = 0; };

// marching_cubes_gpu (:794-822): count pass, arrays sized (vertex count rounded up to 128, zero filled), generate pass; then
// compute_mesh_1ring's normals (:332-364, 722-728)
inline Mesh marching_cubes(const float* density, const uint32_t res[3], const float mn[3], const float mx[3], float thresh) {
	const Lattice L = make_lattice(res, mn, mx, thresh);
	uint32_t counters[4] = {0, 0, 0, 0};
	for (uint32_t z = 0; z < L.rz; ++z) for (uint32_t y = 0; y < L.ry; ++y) for (uint32_t x = 0; x < L.rx; ++x) gen_vertices_point(L, density, x, y, z, nullptr, nullptr, counters[0]);
	for (uint32_t z = 0; z < L.rz; ++z) for (uint32_t y = 0; y < L.ry; ++y) for (uint32_t x = 0; x < L.rx; ++x) gen_faces_cell(L, density, x, y, z, nullptr, nullptr, counters[1]);
	Mesh M; M.n_verts = counters[0];
	const uint32_t n_verts = (counters[0] + 127u) & ~127u;
	M.verts.assign((size_t)n_verts * 3, 0.f);
	M.indices.assign(counters[1], 0u);
	std::vector<int> grid((size_t)L.rx * L.ry * L.rz * 3, -1);
	for (uint32_t z = 0; z < L.rz; ++z) for (uint32_t y = 0; y < L.ry; ++y) for (uint32_t x = 0; x < L.rx; ++x) gen_vertices_point(L, density, x, y, z, grid.data(), &M.verts, counters[2]);
	for (uint32_t z = 0; z < L.rz; ++z) for (uint32_t y = 0; y < L.ry; ++y) for (uint32_t x = 0; x < L.rx; ++x) gen_faces_cell(L, density, x, y, z, grid.data(), &M.indices, counters[3]);
	// accumulate_1ring: normals_out[i{a,b,c}] += (pb - pa) x (pa - pc), triangle after triangle
	M.normals.assign((size_t)n_verts * 3, 0.f);
	for (size_t t = 0; t + 2 < M.indices.size(); t += 3) {
		const uint32_t ia = M.indices[t], ib = M.indices[t + 1], ic = M.indices[t + 2];
		const float* pa = &M.verts[(size_t)ia * 3]; const float* pb = &M.verts[(size_t)ib * 3]; const float* pc = &M.verts[(size_t)ic * 3];
		const float u[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]}, v[3] = {pa[0] - pc[0], pa[1] - pc[1], pa[2] - pc[2]};
		const float n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
		for (uint32_t i : {ia, ib, ic}) for (int d = 0; d < 3; ++d) M.normals[(size_t)i * 3 + d] += n[d];
	}
	return M;
}

inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }       // tcnn::clamp, common.h
inline void normalized3(float n[3]) {                                                              // Eigen MatrixBase::normalized(), Dot.h:126-136
	const float z = n[0] * n[0] + (n[1] * n[1] + n[2] * n[2]);
	if (z > 0.f) { const float r = std::sqrt(z); n[0] /= r; n[1] /= r; n[2] /= r; }
}

// save_mesh (:824-982) without the unwrap branch
inline int save_mesh(const float* verts, const float* normals, const float* colors, const uint32_t* indices, uint32_t n_verts, uint32_t n_indices, const char* path,
                     float nerf_scale, const float off[3], float n2w_s, const float n2w_t[3], int invert_normals) {
	FILE* f = fopen(path, "wb");
	if (!f) return -1;
	const std::string p(path);
	const size_t dot = p.find_last_of('.');
	const bool ply = dot != std::string::npos && p.substr(dot + 1) == "ply";
	auto world = [&](uint32_t i, float out[3]) { for (int d = 0; d < 3; ++d) { const float q = (verts[(size_t)i * 3 + d] - off[d]) / nerf_scale; out[d] = n2w_s * q + n2w_t[d]; } };
	if (ply) {
		fprintf(f, "ply\nformat ascii 1.0\ncomment output from https://github.com/NVlabs/instant-ngp\nelement vertex %u\nproperty float x\nproperty float y\nproperty float z\n"
		           "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face %u\n"
		           "property list uchar int vertex_index\nend_header\n", (unsigned)n_verts, (unsigned)n_indices / 3);
		for (uint32_t i = 0; i < n_verts; ++i) {
			float w[3]; world(i, w);
			float n[3] = {normals[(size_t)i * 3], normals[(size_t)i * 3 + 1], normals[(size_t)i * 3 + 2]}; normalized3(n);
			const float* c = colors + (size_t)i * 3;
			unsigned char c8[3] = {(unsigned char)clampf(c[0] * 255.f, 0.f, 255.f), (unsigned char)clampf(c[1] * 255.f, 0.f, 255.f), (unsigned char)clampf(c[2] * 255.f, 0.f, 255.f)};
			fprintf(f, "%0.5f %0.5f %0.5f %0.3f %0.3f %0.3f %d %d %d\n", w[0], w[1], w[2], n[0], n[1], n[2], c8[0], c8[1], c8[2]);
		}
		for (size_t i = 0; i < n_indices; i += 3) {
			if (invert_normals) fprintf(f, "3 %d %d %d\n", indices[i + 0], indices[i + 1], indices[i + 2]);
			else fprintf(f, "3 %d %d %d\n", indices[i + 2], indices[i + 1], indices[i + 0]);
		}
	} else {
		for (uint32_t i = 0; i < n_verts; ++i) {
			float w[3]; world(i, w);
			const float* c = colors + (size_t)i * 3;
			fprintf(f, "v %0.5f %0.5f %0.5f %0.3f %0.3f %0.3f\n", w[0], w[1], w[2], clampf(c[0], 0.f, 1.f), clampf(c[1], 0.f, 1.f), clampf(c[2], 0.f, 1.f));
		}
		for (uint32_t i = 0; i < n_verts; ++i) {
			float n[3] = {n2w_s * normals[(size_t)i * 3], n2w_s * normals[(size_t)i * 3 + 1], n2w_s * normals[(size_t)i * 3 + 2]}; normalized3(n);
			fprintf(f, "vn %0.5f %0.5f %0.5f\n", n[0], n[1], n[2]);
		}
		for (size_t i = 0; i < n_indices; i += 3) {
			if (invert_normals) fprintf(f, "f %u//%u %u//%u %u//%u\n", indices[i] + 1, indices[i] + 1, indices[i + 1] + 1, indices[i + 1] + 1, indices[i + 2] + 1, indices[i + 2] + 1);
			else fprintf(f, "f %u//%u %u//%u %u//%u\n", indices[i + 2] + 1, indices[i + 2] + 1, indices[i + 1] + 1, indices[i + 1] + 1, indices[i] + 1, indices[i] + 1);
		}
	}
	fclose(f);
	return 0;
}

} // namespace orc_mesh
