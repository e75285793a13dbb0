// CPU check of the ray/mesh traversal (TEST INFRASTRUCTURE): compiles the SAME grid build and DDA walk that the CUDA kernel runs
// (rnb-neus2_b200/csrc/rnb_raymesh.cuh, rnb_raymesh_build.h) for the host, so that tests/test_raymesh_host.py can compare the
// traversal with brute force over all triangles (oracle/orc_albedo.py) without a GPU.  Not part of the product library.
#include "../../rnb-neus2_b200/csrc/rnb_raymesh_build.h"

using namespace rnb::raymesh;

extern "C" int raymesh_host_trace(const float* verts, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris, uint32_t grid_res,
                                  const double* org, const double* dir, const double* t_max, uint32_t n, double* t_out, uint32_t* tri_out, uint32_t res_out[3]) {
	HostGrid H;
	build_grid(verts, n_verts, indices, n_tris, grid_res, H);
	for (int a = 0; a < 3; ++a) res_out[a] = (uint32_t)H.view.res[a];
	for (uint32_t i = 0; i < n; ++i) {
		double t; uint32_t tri;
		if (t_max) trace<true>(H.view, org + 3 * i, dir + 3 * i, 0.0, t_max[i], t, tri);
		else trace<false>(H.view, org + 3 * i, dir + 3 * i, 0.0, (double)INFINITY, t, tri);
		t_out[i] = t; tri_out[i] = tri;
	}
	return 0;
}
