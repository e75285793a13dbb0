// rnb_mesh.cu — iso-surface extraction and mesh export: the B200-native counterpart of
// marching_cubes_gpu (reference src/marching_cubes.cu:276-330,377-720,794-822), compute_mesh_1ring (:332-364,722-728),
// Testbed::compute_mesh_vertex_colors (src/testbed_nerf.cu:4193-4216) and save_mesh (src/marching_cubes.cu:824-982).
//
// Differences in structure (not in results):
//  * vertex and triangle slots come from ordered scans over the lattice, not from atomicAdd hand-out: the numbering is the
//    lexicographic one (lattice point x + y rx + z rx ry; its +x, +y, +z edge in that order; triangles in table order), the
//    same on every run.  The reference's numbering is whatever order its atomics retire in, so its output is this mesh up
//    to a permutation of the vertices and of the triangles;
//  * the lattice is read once, into one sign bit per point; counting, numbering and case lookup run bit-parallel on 16 points per
//    thread, and the fp32 values are touched again only at the crossing edges (the reference reads the lattice in four passes
//    of 4-7 loads per point);
//  * one 32-bit word per vertex-owning lattice point (first vertex id << 2 | x-edge crossed | y-edge crossed << 1) replaces the
//    reference's three ints per point (12.9 GB at 1024^3 written and cleared -> 4.3 GB reserved, only the surface touched);
//  * vertex normals are gathered per vertex from the (at most four) cells around its edge in a fixed order instead of
//    being scattered with float atomics: same addends, deterministic sum;
//  * the OBJ text is formatted on the GPU (exact "%0.5f" / "%0.3f" / "%u" of glibc in integer arithmetic, ordered scan of the
//    line lengths, staged through shared memory) and written with one fwrite per chunk; the reference runs ~3 fprintf per
//    vertex on one host thread.
// Compiled with -fmad=false: every fused multiply-add below is written out, where nvcc fuses it in the reference's kernels.
#include "rnb_common.cuh"
#include "rnb_mc_tables.cuh"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>

namespace rnb {

struct McGrid {
	uint32_t rx, ry, rz, n;        // lattice size, n = rx ry rz (< 2^32)
	float sx, sy, sz, ox, oy, oz;  // vertex = lattice coordinate * s + o  (s = (aabb.max - aabb.min) / res, :280-281)
	float thresh;
};

// exclusive prefix over the block (in thread order) + block total; red: >= blockDim/32 words of shared memory
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* red, uint32_t& total) {
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint32_t inc = v;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += t; }
	__syncthreads();
	if (lane == 31) red[w] = inc;
	__syncthreads();
	uint32_t base = 0, tot = 0;
	for (uint32_t i = 0; i < nw; ++i) { const uint32_t r = red[i]; if (i < w) base += r; tot += r; }
	total = tot;
	return base + inc - v;
}

// Sign bits.  Everything marching cubes decides (which edges carry a vertex, which case a cell is) depends on the lattice
// only through density > thresh, so pass 0 reads the lattice ONCE, perfectly coalesced, and leaves one bit per point
// (bit idx & 31 of word idx >> 5; 134 MB at 1024^3, L2-resident); the counting, vertex and face passes work on 16 points per
// thread in bit-parallel form and touch the fp32 lattice again only at the (sparse) crossing edges.
__global__ void __launch_bounds__(256) k_mc_bits(uint32_t n, float thresh, const float* __restrict__ D, uint32_t* __restrict__ bits) {
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint64_t base = warp * 128;                     // 128 points (four words) per warp
	if (base >= n) return;
	uint32_t mine = 0;
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const uint64_t i = base + k * 32 + lane;
		const bool above = i < n && __ldg(D + i) > thresh;
		const uint32_t b = __ballot_sync(0xFFFFFFFFu, above);
		if (lane == (uint32_t)k) mine = b;
	}
	if (lane < 4 && base + lane * 32 < n) bits[(base >> 5) + lane] = mine;
}

constexpr int MC_THREADS = 256, MC_PER_THREAD = 16, MC_BLOCK = MC_THREADS * MC_PER_THREAD;

// The 16 lattice points a thread owns (consecutive in x; rx % 16 == 0 keeps them in one row).  b[dz][dy]: bit i = sign bit of
// point (x0 + i, y + dy, z + dz), i = 0..16 (bit 16 = first point of the next group, 0 past the end of the row).
struct Group {
	uint32_t x0, y, z, idx0;
	uint32_t b00, b10, b01, b11;
	uint32_t edge_x;        // points that have a +x neighbour
	bool has_y, has_z;
};
__device__ __forceinline__ uint32_t bits17(const uint32_t* __restrict__ bits, uint32_t idx, bool more) {
	const uint32_t w = __ldg(bits + (idx >> 5));
	if ((idx & 31u) == 0) return more ? (w & 0x1FFFFu) : (w & 0xFFFFu);
	uint32_t r = w >> 16;
	if (more) r |= (__ldg(bits + (idx >> 5) + 1) & 1u) << 16;
	return r;
}
__device__ __forceinline__ bool load_group(const McGrid& g, const uint32_t* __restrict__ bits, uint32_t q, Group& G) {
	const uint64_t i0 = (uint64_t)q * MC_PER_THREAD;
	if (i0 >= g.n) return false;
	const uint32_t idx = (uint32_t)i0, rxy = g.rx * g.ry;
	G.idx0 = idx; G.z = idx / rxy; const uint32_t r = idx - G.z * rxy; G.y = r / g.rx; G.x0 = r - G.y * g.rx;
	const bool more = G.x0 + 16 < g.rx;
	G.has_y = G.y + 1 < g.ry; G.has_z = G.z + 1 < g.rz;
	G.edge_x = more ? 0xFFFFu : 0x7FFFu;
	G.b00 = bits17(bits, idx, more);
	G.b10 = G.has_y ? bits17(bits, idx + g.rx, more) : 0u;
	G.b01 = G.has_z ? bits17(bits, idx + rxy, more) : 0u;
	G.b11 = (G.has_y && G.has_z) ? bits17(bits, idx + rxy + g.rx, more) : 0u;
	return true;
}
// crossed +x / +y / +z edges per point (gen_vertices, :289-327), one bit per point
__device__ __forceinline__ void group_cross(const Group& G, uint32_t& cx, uint32_t& cy, uint32_t& cz) {
	cx = (G.b00 ^ (G.b00 >> 1)) & G.edge_x;
	cy = G.has_y ? ((G.b00 ^ G.b10) & 0xFFFFu) : 0u;
	cz = G.has_z ? ((G.b00 ^ G.b01) & 0xFFFFu) : 0u;
}
// cells (by origin point) that are neither all-below nor all-above
__device__ __forceinline__ uint32_t group_active_cells(const Group& G) {
	if (!G.has_y || !G.has_z) return 0u;
	const uint32_t s00 = G.b00 >> 1, s10 = G.b10 >> 1, s01 = G.b01 >> 1, s11 = G.b11 >> 1;
	const uint32_t any = G.b00 | s00 | G.b10 | s10 | G.b01 | s01 | G.b11 | s11, all = G.b00 & s00 & G.b10 & s10 & G.b01 & s01 & G.b11 & s11;
	return (any & ~all) & G.edge_x;
}
// case mask of the cell with origin at point i (gen_faces, :668-679)
__device__ __forceinline__ uint32_t group_mask(const Group& G, int i) {
	const int j = i + 1;
	return ((G.b00 >> i) & 1u) | (((G.b00 >> j) & 1u) << 1) | (((G.b10 >> j) & 1u) << 2) | (((G.b10 >> i) & 1u) << 3) |
	       (((G.b01 >> i) & 1u) << 4) | (((G.b01 >> j) & 1u) << 5) | (((G.b11 >> j) & 1u) << 6) | (((G.b11 >> i) & 1u) << 7);
}
__device__ __forceinline__ uint32_t group_index_count(const Group& G, uint32_t active) {
	uint32_t ni = 0;
	while (active) { const int i = __ffs(active) - 1; active &= active - 1; ni += mc_index_count(MC_TRIANGLES[group_mask(G, i)]); }
	return ni;
}

// pass 1: vertices and triangle indices per block of MC_BLOCK lattice points
__global__ void __launch_bounds__(MC_THREADS) k_mc_count(McGrid g, const uint32_t* __restrict__ bits, uint32_t* __restrict__ blockV, uint32_t* __restrict__ blockI) {
	__shared__ uint32_t red[MC_THREADS / 32];
	Group G; uint32_t nv = 0, ni = 0;
	if (load_group(g, bits, blockIdx.x * MC_THREADS + threadIdx.x, G)) {
		uint32_t cx, cy, cz; group_cross(G, cx, cy, cz);
		nv = __popc(cx) + __popc(cy) + __popc(cz);
		ni = group_index_count(G, group_active_cells(G));
	}
	uint32_t tv, ti;
	block_exclusive(nv, red, tv);
	block_exclusive(ni, red, ti);
	if (threadIdx.x == 0) { blockV[blockIdx.x] = tv; blockI[blockIdx.x] = ti; }
}

// exclusive scan of per-block counts by one CTA per array (a few thousand to a million entries: L2-resident), 16 entries per
// thread and iteration; CTA b scans in + b * stride -> out + b * stride, total -> total[b].  in may alias out.
template <typename TOut>
__global__ void __launch_bounds__(1024) k_scan_counts(const uint32_t* in, TOut* out, uint64_t n, uint64_t stride, TOut* total) {
	__shared__ TOut red[32];
	__shared__ TOut carry_s;
	in += blockIdx.x * stride; out += blockIdx.x * stride;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	constexpr int PER = 16;
	for (uint64_t base = 0; base < n; base += 1024 * PER) {
		const uint64_t i = base + (uint64_t)threadIdx.x * PER;
		uint32_t v[PER];
		TOut mine = 0;
		#pragma unroll
		for (int k = 0; k < PER; ++k) { v[k] = (i + k < n) ? in[i + k] : 0u; mine += v[k]; }
		TOut inc = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const TOut t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= (uint32_t)d) inc += t; }
		if (lane == 31) red[w] = inc;
		__syncthreads();
		// warp 0 turns the 32 warp totals into exclusive prefixes; red[31] keeps the grand total of the iteration in lane 31's copy
		TOut wt = red[lane], winc = wt;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const TOut t = __shfl_up_sync(0xFFFFFFFFu, winc, d); if (lane >= (uint32_t)d) winc += t; }
		const TOut wbase = __shfl_sync(0xFFFFFFFFu, winc - wt, w);
		const TOut tot = __shfl_sync(0xFFFFFFFFu, winc, 31);
		TOut run = carry_s + wbase + inc - mine;
		#pragma unroll
		for (int k = 0; k < PER; ++k) { if (i + k < n) out[i + k] = run; run += v[k]; }
		__syncthreads();
		if (threadIdx.x == 0) carry_s += tot;
		__syncthreads();
	}
	if (threadIdx.x == 0) total[blockIdx.x] = carry_s;
}

// cube edge -> (owner point offset dx | dy << 1 | dz << 2 | axis << 3), five bits per edge (numbering: rnb_mc_tables.cuh)
constexpr uint64_t MC_EDGE_OWNER = (0ull) | (9ull << 5) | (2ull << 10) | (8ull << 15) | (4ull << 20) | (13ull << 25) | (6ull << 30) | (12ull << 35) |
                                   (16ull << 40) | (17ull << 45) | (19ull << 50) | (18ull << 55);
__device__ __forceinline__ uint32_t edge_owner(uint32_t e) { return (uint32_t)(MC_EDGE_OWNER >> (5 * e)) & 31u; }

// pass 3: triangle indices (gen_faces, :680-719)
__global__ void __launch_bounds__(MC_THREADS) k_mc_faces(McGrid g, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ blockI,
                                                          const uint32_t* __restrict__ point_word, uint32_t* __restrict__ indices) {
	__shared__ uint32_t red[MC_THREADS / 32];
	Group G; uint32_t active = 0, ni = 0;
	const bool live = load_group(g, bits, blockIdx.x * MC_THREADS + threadIdx.x, G);
	if (live) { active = group_active_cells(G); ni = group_index_count(G, active); }
	uint32_t tot;
	uint32_t t = blockI[blockIdx.x] + block_exclusive(ni, red, tot);
	if (!ni) return;
	const uint32_t rxy = g.rx * g.ry;
	while (active) {
		const int i = __ffs(active) - 1; active &= active - 1;
		uint64_t row = MC_TRIANGLES[group_mask(G, i)];
		const uint32_t cnt = mc_index_count(row);
		for (uint32_t k = 0; k < cnt; ++k, row >>= 4) {
			const uint32_t o = edge_owner((uint32_t)row & 15u);
			const uint32_t w = __ldg(point_word + (G.idx0 + i + (o & 1u) + ((o >> 1) & 1u) * g.rx + ((o >> 2) & 1u) * rxy));
			const uint32_t axis = o >> 3;
			indices[t++] = (w >> 2) + (axis == 0 ? 0u : axis == 1 ? (w & 1u) : (w & 1u) + ((w >> 1) & 1u));
		}
	}
}

// ---- vertex normals (accumulate_1ring, :332-364: sum over the triangles of a vertex of (pb - pa) x (pa - pc)) --------------------
struct CellCorners { float d[8]; uint32_t x, y, z; };     // d[dx + 2 dy + 4 dz]
__device__ __forceinline__ void edge_position(const McGrid& g, const CellCorners& C, uint32_t e, float p[3]) {
	const uint32_t o = edge_owner(e), dx = o & 1u, dy = (o >> 1) & 1u, dz = (o >> 2) & 1u, axis = o >> 3;
	const float f0 = C.d[dx + 2 * dy + 4 * dz], f1 = C.d[(dx + 2 * dy + 4 * dz) + (1u << axis)];
	const float dt = (g.thresh - f0) / (f1 - f0);
	float l[3] = {(float)(C.x + dx), (float)(C.y + dy), (float)(C.z + dz)};
	l[axis] += dt;
	p[0] = fmaf(l[0], g.sx, g.ox); p[1] = fmaf(l[1], g.sy, g.oy); p[2] = fmaf(l[2], g.sz, g.oz);
}
// local edge id of an axis-a edge in the cell that lies (j, k) cells below it along the two other axes (in x<y<z order)
__device__ __forceinline__ uint32_t local_edge(uint32_t axis, uint32_t j, uint32_t k) {
	const uint32_t t = j + 2 * k;
	return axis == 0 ? ((0x6420u >> (4 * t)) & 15u) : axis == 1 ? ((0x5713u >> (4 * t)) & 15u) : ((0xAB98u >> (4 * t)) & 15u);   // x: 0,2,4,6  y: 3,1,7,5  z: 8,9,11,10
}
__device__ void vertex_normal(const McGrid& g, const float* __restrict__ D, uint32_t px, uint32_t py, uint32_t pz, uint32_t axis, float n[3]) {
	n[0] = n[1] = n[2] = 0.f;
	const uint32_t rxy = g.rx * g.ry;
	// the four cells around the edge in ascending lattice order (t = 3, 2, 1, 0) and their triangles in table order: the order in
	// which a triangle-after-triangle accumulation (accumulate_1ring run sequentially) adds them up
	for (int t = 3; t >= 0; --t) {
		const uint32_t j = (uint32_t)t & 1u, k = (uint32_t)t >> 1;
		// cell origin: the edge's owner moved down by j along the first other axis and by k along the second
		uint32_t c[3] = {px, py, pz};
		const uint32_t a1 = axis == 0 ? 1u : 0u, a2 = axis == 2 ? 1u : 2u;
		if (c[a1] < j || c[a2] < k) continue;
		c[a1] -= j; c[a2] -= k;
		if (c[0] + 1 >= g.rx || c[1] + 1 >= g.ry || c[2] + 1 >= g.rz) continue;
		CellCorners C; C.x = c[0]; C.y = c[1]; C.z = c[2];
		const uint32_t base = c[0] + c[1] * g.rx + c[2] * rxy;
		uint32_t mask = 0;
		#pragma unroll
		for (uint32_t q = 0; q < 8; ++q) C.d[q] = __ldg(D + (base + (q & 1u) + ((q >> 1) & 1u) * g.rx + (q >> 2) * rxy));
		// corner numbering of the case table: 0,1,2,3 = (0,0),(1,0),(1,1),(0,1) in (dx,dy)
		mask = (uint32_t)(C.d[0] > g.thresh) | ((uint32_t)(C.d[1] > g.thresh) << 1) | ((uint32_t)(C.d[3] > g.thresh) << 2) | ((uint32_t)(C.d[2] > g.thresh) << 3) |
		       ((uint32_t)(C.d[4] > g.thresh) << 4) | ((uint32_t)(C.d[5] > g.thresh) << 5) | ((uint32_t)(C.d[7] > g.thresh) << 6) | ((uint32_t)(C.d[6] > g.thresh) << 7);
		const uint32_t me = local_edge(axis, j, k);
		uint64_t row = MC_TRIANGLES[mask];
		const uint32_t cnt = mc_index_count(row);
		for (uint32_t q = 0; q < cnt; q += 3, row >>= 12) {
			const uint32_t ea = (uint32_t)row & 15u, eb = (uint32_t)(row >> 4) & 15u, ec = (uint32_t)(row >> 8) & 15u;
			if (ea != me && eb != me && ec != me) continue;
			float pa[3], pb[3], pc[3];
			edge_position(g, C, ea, pa); edge_position(g, C, eb, pb); edge_position(g, C, ec, pc);
			const float u0 = pb[0] - pa[0], u1 = pb[1] - pa[1], u2 = pb[2] - pa[2];
			const float v0 = pa[0] - pc[0], v1 = pa[1] - pc[1], v2 = pa[2] - pc[2];
			n[0] += u1 * v2 - u2 * v1;
			n[1] += u2 * v0 - u0 * v2;
			n[2] += u0 * v1 - u1 * v0;
		}
	}
}
// pass 2: vertex positions and, for every point that owns a vertex, its word (first vertex id << 2 | +x crossed | +y crossed << 1);
// points without crossings are never looked up by the face pass.  The slot of the vertex normal receives the edge the vertex
// sits on (owner point, axis) for k_mc_normals, which runs one thread per vertex: the surface is sparse in the lattice, and
// a thread per lattice group would leave 31 lanes of a warp waiting for the one that found vertices.
__global__ void __launch_bounds__(MC_THREADS) k_mc_vertices(McGrid g, const float* __restrict__ D, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ blockV,
                                                             uint32_t* __restrict__ point_word, float* __restrict__ verts, float* __restrict__ normals) {
	__shared__ uint32_t red[MC_THREADS / 32];
	Group G; uint32_t cx = 0, cy = 0, cz = 0, nv = 0;
	const bool live = load_group(g, bits, blockIdx.x * MC_THREADS + threadIdx.x, G);
	if (live) { group_cross(G, cx, cy, cz); nv = __popc(cx) + __popc(cy) + __popc(cz); }
	uint32_t tot;
	uint32_t v = blockV[blockIdx.x] + block_exclusive(nv, red, tot);
	if (!nv) return;
	const uint32_t rxy = g.rx * g.ry;
	const float fy = (float)G.y, fz = (float)G.z;
	uint32_t any = cx | cy | cz;
	while (any) {
		const int i = __ffs(any) - 1; any &= any - 1;
		const uint32_t p = G.idx0 + i, bx = (cx >> i) & 1u, by = (cy >> i) & 1u, bz = (cz >> i) & 1u;
		point_word[p] = (v << 2) | bx | (by << 1);
		const float fx = (float)(G.x0 + i), f0 = __ldg(D + p);
		// dt = (thresh - f0) / (f1 - f0); vertex = (lattice + dt e_axis) * scale + offset   (:296-300)
		if (bx) {
			const float dt = (g.thresh - f0) / (__ldg(D + p + 1) - f0); float* o = verts + (size_t)v * 3; float* nn = normals + (size_t)v * 3;
			o[0] = fmaf(fx + dt, g.sx, g.ox); o[1] = fmaf(fy, g.sy, g.oy); o[2] = fmaf(fz, g.sz, g.oz);
			nn[0] = __uint_as_float(p); nn[1] = __uint_as_float(0u); ++v;
		}
		if (by) {
			const float dt = (g.thresh - f0) / (__ldg(D + p + g.rx) - f0); float* o = verts + (size_t)v * 3; float* nn = normals + (size_t)v * 3;
			o[0] = fmaf(fx, g.sx, g.ox); o[1] = fmaf(fy + dt, g.sy, g.oy); o[2] = fmaf(fz, g.sz, g.oz);
			nn[0] = __uint_as_float(p); nn[1] = __uint_as_float(1u); ++v;
		}
		if (bz) {
			const float dt = (g.thresh - f0) / (__ldg(D + p + rxy) - f0); float* o = verts + (size_t)v * 3; float* nn = normals + (size_t)v * 3;
			o[0] = fmaf(fx, g.sx, g.ox); o[1] = fmaf(fy, g.sy, g.oy); o[2] = fmaf(fz + dt, g.sz, g.oz);
			nn[0] = __uint_as_float(p); nn[1] = __uint_as_float(2u); ++v;
		}
	}
}

__global__ void __launch_bounds__(256) k_mc_normals(McGrid g, const float* __restrict__ D, uint32_t n_verts, float* __restrict__ normals) {
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n_verts) return;
	float* nn = normals + (size_t)v * 3;
	const uint32_t p = __float_as_uint(nn[0]), axis = __float_as_uint(nn[1]), rxy = g.rx * g.ry;
	const uint32_t z = p / rxy, r = p - z * rxy, y = r / g.rx, x = r - y * g.rx;
	float nl[3];
	vertex_normal(g, D, x, y, z, axis, nl);
	nn[0] = nl[0]; nn[1] = nl[1]; nn[2] = nl[2];
}

// ---- vertex colours: network inputs (generate_nerf_network_inputs_from_positions, src/testbed_nerf.cu:793-799) and output
// activation (extract_srgb_with_activation :477-490 with the Logistic rgb activation of an LDR dataset, :3121) ----------------------
__global__ void k_mesh_color_inputs(uint32_t n, const float* __restrict__ verts, float4* __restrict__ pos4, float* __restrict__ dirw) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float x = verts[(size_t)i * 3], y = verts[(size_t)i * 3 + 1], z = verts[(size_t)i * 3 + 2];
	float dx = x - 0.5f, dy = y - 0.5f, dz = z - 0.5f;
	const float s = dx * dx + (dy * dy + dz * dz);
	if (s > 0.f) { const float r = sqrtf(s); dx /= r; dy /= r; dz /= r; }
	pos4[i] = make_float4(x, y, z, __uint_as_float(i));
	dirw[(size_t)i * 3] = (dx + 1.f) * 0.5f; dirw[(size_t)i * 3 + 1] = (dy + 1.f) * 0.5f; dirw[(size_t)i * 3 + 2] = (dz + 1.f) * 0.5f;   // warp_direction
}
__global__ void k_mesh_colors(uint32_t n, const __half* __restrict__ out16, float* __restrict__ colors) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * 3) return;
	const uint32_t v = i / 3, d = i - v * 3;
	colors[i] = logisticf(__half2float(out16[(size_t)v * 16 + d]));
}

// ---- OBJ / PLY text --------------------------------------------------------------------------------------------------------
// printf("%0.<D>f", (double)v) for a binary32 v exactly as glibc prints it: the decimal expansion of the exact binary value,
// rounded half-to-even at D fractional digits.  |v| = m 2^e with a 24-bit m, so round(|v| 10^D) is one shift of m 10^D
// (< 2^41) with its remainder compared against one half; values >= 2^24 are integers and print all their digits.
template <int DIGITS>
__device__ uint32_t fmt_fixed(float v, char* out) {
	constexpr uint32_t P = DIGITS == 5 ? 100000u : DIGITS == 3 ? 1000u : 1u;
	const uint32_t bits = __float_as_uint(v), ex = (bits >> 23) & 0xFFu, man = bits & 0x7FFFFFu;
	uint32_t len = 0;
	if (bits >> 31) out[len++] = '-';
	if (ex == 0xFFu) { const char* s = man ? "nan" : "inf"; out[len++] = s[0]; out[len++] = s[1]; out[len++] = s[2]; return len; }
	const uint64_t m = ex ? (uint64_t)(man | 0x800000u) : (uint64_t)man;
	const int e = (int)(ex ? ex : 1u) - 150;
	char tmp[40]; int nd = 0;
	uint32_t frac = 0;
	if (e >= 0) {                                         // an integer: m << e < 2^128
		unsigned __int128 ip = (unsigned __int128)m << e;
		do { tmp[nd++] = (char)('0' + (int)(ip % 10)); ip /= 10; } while (ip != 0);
	} else {
		const int s = -e;
		uint64_t q = 0;
		if (s < 64) {
			const uint64_t X = m * P, half = 1ull << (s - 1), r = X & ((half << 1) - 1ull);
			q = X >> s;
			if (r > half || (r == half && (q & 1ull))) ++q;
		}
		uint64_t ip = q / P; frac = (uint32_t)(q - ip * P);
		do { tmp[nd++] = (char)('0' + (int)(ip % 10)); ip /= 10; } while (ip != 0);
	}
	while (nd) out[len++] = tmp[--nd];
	if (DIGITS > 0) {
		out[len++] = '.';
		uint32_t div = P / 10;
		#pragma unroll
		for (int k = 0; k < DIGITS; ++k) { const uint32_t d = frac / div; out[len++] = (char)('0' + d); frac -= d * div; div = div > 1 ? div / 10 : 1; }
	}
	return len;
}
__device__ __forceinline__ uint32_t fmt_u32(uint32_t v, char* out) {
	char tmp[10]; int nd = 0;
	do { tmp[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
	uint32_t len = 0;
	while (nd) out[len++] = tmp[--nd];
	return len;
}
__device__ __forceinline__ uint32_t fmt_i32(int v, char* out) {
	uint32_t len = 0;
	if (v < 0) { out[len++] = '-'; return len + fmt_u32((uint32_t)(-(int64_t)v), out + len); }
	return fmt_u32((uint32_t)v, out);
}

struct MeshText {
	const float* verts; const float* normals; const float* colors; const uint32_t* indices;
	uint32_t n_verts, n_tris;
	float nerf_scale, off[3], n2w_s, n2w_t[3];
	int invert, kind;       // kind: 0 OBJ "v", 1 OBJ "vn", 2 OBJ "f", 3 PLY vertex, 4 PLY face
};
constexpr int TXT_THREADS = 128, TXT_MAX = 192;      // longest possible record: a "v" line of six finite binary32 values (162 bytes)

__device__ __forceinline__ float clamp01(float v, float hi) { return v < 0.f ? 0.f : (hi < v ? hi : v); }     // tcnn::clamp(v, 0, hi), NaN passes through

// one record of the file (save_mesh, src/marching_cubes.cu:926-979) into `o`; returns its length
__device__ uint32_t format_record(const MeshText& T, uint32_t i, char* o) {
	uint32_t n = 0;
	if (T.kind == 0 || T.kind == 3) {
		float p[3];
		#pragma unroll
		for (int d = 0; d < 3; ++d) p[d] = T.n2w_s * ((T.verts[(size_t)i * 3 + d] - T.off[d]) / T.nerf_scale) + T.n2w_t[d];     // two roundings, as on the host (no fma)
		const float c[3] = {T.colors[(size_t)i * 3], T.colors[(size_t)i * 3 + 1], T.colors[(size_t)i * 3 + 2]};
		if (T.kind == 0) {
			o[n++] = 'v';
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_fixed<5>(p[d], o + n); }
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_fixed<3>(clamp01(c[d], 1.f), o + n); }
		} else {
			float nn[3] = {T.normals[(size_t)i * 3], T.normals[(size_t)i * 3 + 1], T.normals[(size_t)i * 3 + 2]};
			const float z = nn[0] * nn[0] + (nn[1] * nn[1] + nn[2] * nn[2]);      // Eigen squaredNorm of a Vector3f: x^2 + (y^2 + z^2)
			if (z > 0.f) { const float r = sqrtf(z); nn[0] /= r; nn[1] /= r; nn[2] /= r; }
			for (int d = 0; d < 3; ++d) { if (d) o[n++] = ' '; n += fmt_fixed<5>(p[d], o + n); }
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_fixed<3>(nn[d], o + n); }
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_i32((int)(unsigned char)clamp01(c[d] * 255.f, 255.f), o + n); }
		}
	} else if (T.kind == 1) {
		float nn[3];
		#pragma unroll
		for (int d = 0; d < 3; ++d) nn[d] = T.n2w_s * T.normals[(size_t)i * 3 + d];
		const float z = nn[0] * nn[0] + (nn[1] * nn[1] + nn[2] * nn[2]);
		if (z > 0.f) { const float r = sqrtf(z); nn[0] /= r; nn[1] /= r; nn[2] /= r; }
		o[n++] = 'v'; o[n++] = 'n';
		for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_fixed<5>(nn[d], o + n); }
	} else {
		uint32_t a = T.indices[(size_t)i * 3], b = T.indices[(size_t)i * 3 + 1], c = T.indices[(size_t)i * 3 + 2];
		if (!T.invert) { const uint32_t t = a; a = c; c = t; }
		if (T.kind == 2) {
			o[n++] = 'f';
			const uint32_t id[3] = {a + 1u, b + 1u, c + 1u};
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_u32(id[d], o + n); o[n++] = '/'; o[n++] = '/'; n += fmt_u32(id[d], o + n); }
		} else {
			o[n++] = '3';
			const int id[3] = {(int)a, (int)b, (int)c};
			for (int d = 0; d < 3; ++d) { o[n++] = ' '; n += fmt_i32(id[d], o + n); }
		}
	}
	o[n++] = '\n';
	return n;
}

// pass 1 (text == nullptr): bytes per block of TXT_THREADS records; pass 2: the bytes, at the scanned block offsets
__global__ void __launch_bounds__(TXT_THREADS) k_mesh_text(MeshText T, uint32_t first, uint32_t n_rec, uint32_t* __restrict__ block_bytes,
                                                            const uint64_t* __restrict__ block_off, char* __restrict__ text) {
	__shared__ uint32_t red[TXT_THREADS / 32];
	__shared__ char stage[TXT_THREADS * TXT_MAX];
	const uint32_t i = blockIdx.x * TXT_THREADS + threadIdx.x;
	char line[TXT_MAX];
	const uint32_t len = i < n_rec ? format_record(T, first + i, line) : 0u;
	uint32_t tot;
	const uint32_t at = block_exclusive(len, red, tot);
	if (!text) { if (threadIdx.x == 0) block_bytes[blockIdx.x] = tot; return; }
	for (uint32_t k = 0; k < len; ++k) stage[at + k] = line[k];
	__syncthreads();
	char* dst = text + block_off[blockIdx.x];
	for (uint32_t k = threadIdx.x; k < tot; k += TXT_THREADS) dst[k] = stage[k];
}

// ---- host side ---------------------------------------------------------------------------------------------------------
#define MCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); goto done; } } while (0)

McGrid make_mc_grid(const uint32_t res[3], const float mn[3], const float mx[3], float thresh) {
	McGrid g;
	g.rx = res[0]; g.ry = res[1]; g.rz = res[2]; g.n = res[0] * res[1] * res[2];
	g.sx = (mx[0] - mn[0]) / (float)res[0]; g.sy = (mx[1] - mn[1]) / (float)res[1]; g.sz = (mx[2] - mn[2]) / (float)res[2];
	g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.thresh = thresh;
	return g;
}

// marching_cubes_gpu (:794-822) + compute_mesh_1ring normals.  Allocates *verts / *normals (n_verts rounded up to 128, the
// padding zeroed, :810-812) and *indices with cudaMalloc; the caller frees them.  ws / ws_bytes: grow-only scratch kept by the
// caller between calls (one word and one bit per lattice point + the block counters).  ms[0] = sign bits + count + scan, ms[1] = vertices +
// normals + faces (CUDA events).  Returns "" or an error message.  launches += kernels run.
std::string mesh_extract(cudaStream_t st, const float* density, const uint32_t res[3], const float mn[3], const float mx[3], float thresh, void** ws, size_t* ws_bytes,
                         float** verts_out, float** normals_out, uint32_t** indices_out, uint32_t* n_verts, uint32_t* n_verts_padded, uint32_t* n_indices, float ms[2], uint64_t* launches) {
	std::string err;
	const McGrid g = make_mc_grid(res, mn, mx, thresh);
	const uint32_t nb = (uint32_t)(((uint64_t)g.n + MC_BLOCK - 1) / MC_BLOCK);
	const size_t nb_pad = ((size_t)nb + 63) & ~(size_t)63;
	const size_t n_bitwords = (((size_t)g.n + 31) / 32 + 64) & ~(size_t)63;
	const size_t need = (size_t)g.n * 4 + n_bitwords * 4 + nb_pad * 8 + 256;
	uint32_t *blockV = nullptr, *blockI = nullptr, *totals = nullptr, *words = nullptr, *bits = nullptr, *indices = nullptr;
	float *verts = nullptr, *normals = nullptr;
	uint32_t tot[2] = {0, 0}, nvp = 0;
	cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
	ms[0] = ms[1] = 0.f;
	if (*ws_bytes < need) { cudaFree(*ws); *ws = nullptr; *ws_bytes = 0; MCU(cudaMalloc(ws, need)); *ws_bytes = need; }
	words = (uint32_t*)*ws; bits = words + g.n; blockV = bits + n_bitwords; blockI = blockV + nb_pad; totals = blockI + nb_pad;
	for (auto& e : ev) MCU(cudaEventCreate(&e));
	MCU(cudaEventRecord(ev[0], st));
	k_mc_bits<<<(uint32_t)((((uint64_t)g.n + 127) / 128 * 32 + 255) / 256), 256, 0, st>>>(g.n, thresh, density, bits);
	k_mc_count<<<nb, MC_THREADS, 0, st>>>(g, bits, blockV, blockI);
	k_scan_counts<uint32_t><<<2, 1024, 0, st>>>(blockV, blockV, nb, nb_pad, totals);
	MCU(cudaEventRecord(ev[1], st));
	MCU(cudaMemcpyAsync(tot, totals, 8, cudaMemcpyDeviceToHost, st));
	MCU(cudaStreamSynchronize(st));
	*launches += 3;
	if (tot[0] >= (1u << 30)) { err = "more than 2^30 vertices"; goto done; }
	nvp = (tot[0] + 127u) & ~127u;
	MCU(cudaMalloc(&verts, std::max<size_t>(nvp, 1) * 12)); MCU(cudaMalloc(&normals, std::max<size_t>(nvp, 1) * 12));
	MCU(cudaMalloc(&indices, std::max<size_t>(tot[1], 1) * 4));
	MCU(cudaMemsetAsync(verts, 0, std::max<size_t>(nvp, 1) * 12, st)); MCU(cudaMemsetAsync(normals, 0, std::max<size_t>(nvp, 1) * 12, st));
	if (tot[0]) {
		k_mc_vertices<<<nb, MC_THREADS, 0, st>>>(g, density, bits, blockV, words, verts, normals);
		k_mc_faces<<<nb, MC_THREADS, 0, st>>>(g, bits, blockI, words, indices);
		k_mc_normals<<<(tot[0] + 255) / 256, 256, 0, st>>>(g, density, tot[0], normals);
		*launches += 3;
	}
	MCU(cudaEventRecord(ev[2], st));
	MCU(cudaGetLastError());
	MCU(cudaStreamSynchronize(st));
	cudaEventElapsedTime(&ms[0], ev[0], ev[1]); cudaEventElapsedTime(&ms[1], ev[1], ev[2]);
	*verts_out = verts; *normals_out = normals; *indices_out = indices; verts = normals = nullptr; indices = nullptr;
	*n_verts = tot[0]; *n_verts_padded = nvp; *n_indices = tot[1];
done:
	for (auto& e : ev) if (e) cudaEventDestroy(e);
	if (!err.empty()) { cudaFree(verts); cudaFree(normals); cudaFree(indices); }
	return err;
}

void launch_mesh_color_inputs(cudaStream_t st, uint32_t n, const float* verts, float4* pos4, float* dirw) { if (n) k_mesh_color_inputs<<<(n + 255) / 256, 256, 0, st>>>(n, verts, pos4, dirw); }
void launch_mesh_colors(cudaStream_t st, uint32_t n, const __half* out16, float* colors) { if (n) k_mesh_colors<<<(n * 3 + 255) / 256, 256, 0, st>>>(n, out16, colors); }

// Pinned staging for the text on its way to the file: two slices, so that the device->host copy of slice k+1 runs while the
// host writes slice k.  Process-wide, allocated on first use (pinning 32 MB costs more than formatting a whole mesh).
namespace {
constexpr size_t STAGE_BYTES = 16u << 20;
struct TextStaging { char* buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; std::mutex mu;
	~TextStaging() { for (int i = 0; i < 2; ++i) { if (buf[i]) cudaFreeHost(buf[i]); } } };
TextStaging g_staging;
}

// save_mesh (:824-982): OBJ (positions + colours, normals, faces; the unwrap/texture variant is not provided) or, for a path
// ending in "ply" as the reference tests it, ASCII PLY.  All arrays on the device.
std::string mesh_write(cudaStream_t st, const float* verts, const float* normals, const float* colors, const uint32_t* indices, uint32_t n_verts, uint32_t n_indices,
                       const char* path, float nerf_scale, const float off[3], float n2w_s, const float n2w_t[3], int invert, uint64_t* bytes_out, uint64_t* launches) {
	std::string err;
	MeshText T; T.verts = verts; T.normals = normals; T.colors = colors; T.indices = indices; T.n_verts = n_verts; T.n_tris = n_indices / 3;
	T.nerf_scale = nerf_scale; T.n2w_s = n2w_s; T.invert = invert;
	for (int d = 0; d < 3; ++d) { T.off[d] = off[d]; T.n2w_t[d] = n2w_t[d]; }
	const std::string p(path);
	const size_t dot = p.find_last_of('.');
	const bool ply = dot != std::string::npos && p.substr(dot + 1) == "ply";
	std::lock_guard<std::mutex> lock(g_staging.mu);
	FILE* f = fopen(path, "wb");
	if (!f) return std::string("Failed to open ") + path + " for writing.";
	uint64_t written = 0;
	const uint32_t CH = 1u << 21;                          // records per chunk: bounds the text buffer (<= 2 M x 162 B)
	uint32_t* bb = nullptr; uint64_t* bo = nullptr; uint64_t* total_dev = nullptr; char* text = nullptr;
	size_t text_cap = 0;
	const uint32_t nbmax = (CH + TXT_THREADS - 1) / TXT_THREADS;
	const int kinds_obj[3] = {0, 1, 2}, kinds_ply[2] = {3, 4};
	const int* kinds = ply ? kinds_ply : kinds_obj; const int nk = ply ? 2 : 3;
	MCU(cudaMalloc(&bb, (size_t)nbmax * 4)); MCU(cudaMalloc(&bo, (size_t)nbmax * 8)); MCU(cudaMalloc(&total_dev, 8));
	for (int i = 0; i < 2; ++i) {
		if (!g_staging.buf[i]) MCU(cudaMallocHost(&g_staging.buf[i], STAGE_BYTES));
		if (!g_staging.ev[i]) MCU(cudaEventCreateWithFlags(&g_staging.ev[i], cudaEventDisableTiming));
	}
	if (ply) {
		written += (uint64_t)fprintf(f, "ply\nformat ascii 1.0\ncomment output from https://github.com/NVlabs/instant-ngp\nelement vertex %u\nproperty float x\nproperty float y\nproperty float z\n"
		                                "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nelement face %u\n"
		                                "property list uchar int vertex_index\nend_header\n", n_verts, n_indices / 3);
	}
	for (int s = 0; s < nk; ++s) {
		T.kind = kinds[s];
		const uint32_t n_rec = (T.kind == 2 || T.kind == 4) ? T.n_tris : n_verts;
		for (uint32_t first = 0; first < n_rec; first += CH) {
			const uint32_t m = std::min(CH, n_rec - first), nb = (m + TXT_THREADS - 1) / TXT_THREADS;
			uint64_t total = 0;
			k_mesh_text<<<nb, TXT_THREADS, 0, st>>>(T, first, m, bb, nullptr, nullptr);
			k_scan_counts<uint64_t><<<1, 1024, 0, st>>>(bb, bo, nb, 0, total_dev);
			MCU(cudaMemcpyAsync(&total, total_dev, 8, cudaMemcpyDeviceToHost, st));
			MCU(cudaStreamSynchronize(st));
			if (total > text_cap) { cudaFree(text); text = nullptr; text_cap = 0; MCU(cudaMalloc(&text, (size_t)total + (total >> 2))); text_cap = (size_t)total + (total >> 2); }
			k_mesh_text<<<nb, TXT_THREADS, 0, st>>>(T, first, m, bb, bo, text);
			*launches += 3;
			// device -> pinned slice -> file, the copy of the next slice in flight while this one is written
			const uint64_t n_slices = (total + STAGE_BYTES - 1) / STAGE_BYTES;
			auto issue = [&](uint64_t k) -> cudaError_t {
				const uint64_t o = k * STAGE_BYTES, len = std::min<uint64_t>(STAGE_BYTES, total - o);
				cudaError_t e = cudaMemcpyAsync(g_staging.buf[k & 1], text + o, len, cudaMemcpyDeviceToHost, st);
				return e != cudaSuccess ? e : cudaEventRecord(g_staging.ev[k & 1], st);
			};
			if (n_slices) MCU(issue(0));
			for (uint64_t k = 0; k < n_slices; ++k) {
				if (k + 1 < n_slices) MCU(issue(k + 1));
				MCU(cudaEventSynchronize(g_staging.ev[k & 1]));
				const uint64_t len = std::min<uint64_t>(STAGE_BYTES, total - k * STAGE_BYTES);
				if (fwrite(g_staging.buf[k & 1], 1, len, f) != len) { err = std::string("short write to ") + path; goto done; }
			}
			written += total;
		}
	}
	MCU(cudaGetLastError());
done:
	cudaStreamSynchronize(st);
	if (fclose(f) != 0 && err.empty()) err = std::string("close failed for ") + path;
	cudaFree(bb); cudaFree(bo); cudaFree(total_dev); cudaFree(text);
	if (bytes_out) *bytes_out = written;
	return err;
}

} // namespace rnb
