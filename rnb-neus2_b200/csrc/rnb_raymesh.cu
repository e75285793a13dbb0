// rnb_raymesh.cu — ray / mesh queries of the albedo-scaling stage (reference rnb_neus2/albedo_scaling.py:214-383 calls
// trimesh's mesh.ray.intersects_location twice per view pair: first hit of the sampled pixel rays, then an occlusion test towards
// both neighbour cameras).  The mesh of a finished stage-1 run has millions of marching-cubes triangles of lattice size, so the
// acceleration structure is a uniform cell grid built once on the host in two counting passes (triangles overlap 1-8 cells) and
// walked on the device with a 3D DDA, one thread per ray.  C ABI: rnb_raymesh_create / _intersect / _destroy (include/rnb_b200.h).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <string>
#include <vector>
#include "../../include/rnb_b200.h"
#include "rnb_raymesh.cuh"
#include "rnb_raymesh_build.h"      // host-side grid build, shared with the CPU check of the traversal (tests/cuda/raymesh_host.cpp)

namespace rnb { namespace raymesh {

template <bool ANY>
__global__ void __launch_bounds__(128) k_trace(GridView G, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ t_max,
                                               uint32_t n, double* __restrict__ t_out, uint32_t* __restrict__ tri_out) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const double o[3] = {org[3 * i], org[3 * i + 1], org[3 * i + 2]}, d[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
		double t; uint32_t tri;
		trace<ANY>(G, o, d, 0.0, ANY ? t_max[i] : (double)INFINITY, t, tri);
		t_out[i] = t; tri_out[i] = tri;
	}
}

}} // namespace rnb::raymesh

using namespace rnb::raymesh;

struct rnb_raymesh {
	GridView view;
	uint32_t* cell_start = nullptr; uint32_t* cell_tris = nullptr; float* tri_verts = nullptr;
	double* d_org = nullptr; double* d_dir = nullptr; double* d_tmax = nullptr; double* d_t = nullptr; uint32_t* d_tri = nullptr; size_t cap = 0;
	uint64_t n_refs = 0; int n_sm = 148;
};

extern "C" {

// the error string is shared with rnb_api.cu
int rnb_set_error_(int code, const char* msg);

int rnb_raymesh_destroy(rnb_raymesh* r);
#define RM_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return rnb_set_error_(RNB_ERR_CUDA, (std::string(#x) + ": " + cudaGetErrorString(e_)).c_str()); } while (0)

int rnb_raymesh_create(const float* verts, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris, uint32_t grid_res, rnb_raymesh** out) {
	if (!verts || !indices || !out) return rnb_set_error_(RNB_ERR_INVALID, "null argument");
	if (n_verts == 0 || n_tris == 0) return rnb_set_error_(RNB_ERR_INVALID, "empty mesh");
	for (size_t i = 0; i < (size_t)3 * n_tris; ++i) if (indices[i] >= n_verts) return rnb_set_error_(RNB_ERR_INVALID, "triangle index out of range");
	for (size_t i = 0; i < (size_t)3 * n_verts; ++i) if (!std::isfinite(verts[i])) return rnb_set_error_(RNB_ERR_INVALID, "mesh vertex is not finite");
	int dev_count = 0;
	if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) return rnb_set_error_(RNB_ERR_CUDA, "no CUDA device: the ray/mesh queries have no CPU path");
	HostGrid H;
	try { build_grid(verts, n_verts, indices, n_tris, grid_res, H); }      // no exception leaves the C ABI
	catch (const std::exception& ex) { return rnb_set_error_(RNB_ERR_NOMEM, (std::string("building the cell grid failed: ") + ex.what()).c_str()); }
	rnb_raymesh* r = new rnb_raymesh();
	struct Guard { rnb_raymesh* r; ~Guard() { if (r) rnb_raymesh_destroy(r); } } guard{r};      // released on success
	r->view = H.view; r->n_refs = H.cell_tris.size();
	RM_CU(cudaMalloc(&r->cell_start, H.cell_start.size() * 4)); RM_CU(cudaMalloc(&r->cell_tris, std::max<size_t>(H.cell_tris.size(), 1) * 4)); RM_CU(cudaMalloc(&r->tri_verts, H.tri_verts.size() * 4));
	RM_CU(cudaMemcpy(r->cell_start, H.cell_start.data(), H.cell_start.size() * 4, cudaMemcpyHostToDevice));
	RM_CU(cudaMemcpy(r->cell_tris, H.cell_tris.data(), H.cell_tris.size() * 4, cudaMemcpyHostToDevice));
	RM_CU(cudaMemcpy(r->tri_verts, H.tri_verts.data(), H.tri_verts.size() * 4, cudaMemcpyHostToDevice));
	r->view.cell_start = r->cell_start; r->view.cell_tris = r->cell_tris; r->view.tri_verts = r->tri_verts;
	int dev = 0; cudaGetDevice(&dev); cudaDeviceProp prop; RM_CU(cudaGetDeviceProperties(&prop, dev)); r->n_sm = prop.multiProcessorCount;
	*out = r; guard.r = nullptr;
	return RNB_OK;
}

int rnb_raymesh_destroy(rnb_raymesh* r) {
	if (!r) return RNB_OK;
	cudaFree(r->cell_start); cudaFree(r->cell_tris); cudaFree(r->tri_verts);
	cudaFree(r->d_org); cudaFree(r->d_dir); cudaFree(r->d_tmax); cudaFree(r->d_t); cudaFree(r->d_tri);
	delete r;
	return RNB_OK;
}

int rnb_raymesh_info(rnb_raymesh* r, uint32_t res_out[3], uint64_t* n_refs) {
	if (!r) return rnb_set_error_(RNB_ERR_INVALID, "null handle");
	if (res_out) for (int a = 0; a < 3; ++a) res_out[a] = (uint32_t)r->view.res[a];
	if (n_refs) *n_refs = r->n_refs;
	return RNB_OK;
}

int rnb_raymesh_intersect(rnb_raymesh* r, const double* origins, const double* dirs, const double* t_max, uint32_t n, double* t_out, uint32_t* tri_out, void* stream) {
	if (!r) return rnb_set_error_(RNB_ERR_INVALID, "null handle");
	if (n == 0) return RNB_OK;
	if (!origins || !dirs || !t_out || !tri_out) return rnb_set_error_(RNB_ERR_INVALID, "null argument");
	cudaStream_t st = (cudaStream_t)stream;
	if (r->cap < n) {
		cudaFree(r->d_org); cudaFree(r->d_dir); cudaFree(r->d_tmax); cudaFree(r->d_t); cudaFree(r->d_tri);
		r->d_org = r->d_dir = r->d_tmax = r->d_t = nullptr; r->d_tri = nullptr; r->cap = 0;
		const size_t cap = std::max<size_t>(n, 4096);
		RM_CU(cudaMalloc(&r->d_org, cap * 24)); RM_CU(cudaMalloc(&r->d_dir, cap * 24)); RM_CU(cudaMalloc(&r->d_tmax, cap * 8)); RM_CU(cudaMalloc(&r->d_t, cap * 8)); RM_CU(cudaMalloc(&r->d_tri, cap * 4));
		r->cap = cap;
	}
	RM_CU(cudaMemcpyAsync(r->d_org, origins, (size_t)n * 24, cudaMemcpyHostToDevice, st));
	RM_CU(cudaMemcpyAsync(r->d_dir, dirs, (size_t)n * 24, cudaMemcpyHostToDevice, st));
	if (t_max) RM_CU(cudaMemcpyAsync(r->d_tmax, t_max, (size_t)n * 8, cudaMemcpyHostToDevice, st));
	const uint32_t blocks = std::min<uint32_t>((n + 127) / 128, (uint32_t)r->n_sm * 16);
	if (t_max) k_trace<true><<<blocks, 128, 0, st>>>(r->view, r->d_org, r->d_dir, r->d_tmax, n, r->d_t, r->d_tri);
	else k_trace<false><<<blocks, 128, 0, st>>>(r->view, r->d_org, r->d_dir, nullptr, n, r->d_t, r->d_tri);
	RM_CU(cudaGetLastError());
	RM_CU(cudaMemcpyAsync(t_out, r->d_t, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
	RM_CU(cudaMemcpyAsync(tri_out, r->d_tri, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
	RM_CU(cudaStreamSynchronize(st));
	return RNB_OK;
}

} // extern "C"
