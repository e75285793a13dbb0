// rnb_network_simt.cu — one-thread-per-sample network kernels (CUDA cores).
//
// These are the straightforward kernels: used for the occupancy-grid density sweep (SDF only, tiny MLP work per sample),
// for rnb_eval_sdf, and as the cross-check path for the tensor-core tile kernels in rnb_network_mma.cu
// (RNB_NETWORK=simt selects them for the whole step).
//
// Semantics follow NerfNetwork::forward_impl / backward_impl (reference include/neural-graphics-primitives/nerf_network.h:97-452),
// kernel_grid & friends (tcnn encodings/grid.h:169-364,366-495,556-683,858-883) and FullyFusedMLP (tcnn src/fully_fused_mlp.cu):
// binary16 parameters and activations, fp32 accumulation, binary16 rounding at every layer output.
#include "rnb_common.cuh"

namespace rnb {

struct LevelGeom { float fx, fy, fz; uint32_t gx, gy, gz; };

__device__ __forceinline__ LevelGeom level_geom(float scale, float x, float y, float z) {   // pos_fract, common_device.h:415-424
	LevelGeom g;
	float px = fmaf(x, scale, 0.5f), py = fmaf(y, scale, 0.5f), pz = fmaf(z, scale, 0.5f);
	float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
	g.gx = (uint32_t)(int)fx; g.gy = (uint32_t)(int)fy; g.gz = (uint32_t)(int)fz;
	g.fx = px - fx; g.fy = py - fy; g.fz = pz - fz;
	return g;
}

// Encode one level: binary16 accumulation in corner order (grid.h:291-315) and fp32 dy/dx (grid.h:324-363).
// The 8 corners are gathered once; the reference re-gathers them per axis.
__device__ __forceinline__ void encode_level(const ModelDev& M, const __half* __restrict__ P, uint32_t l, float x, float y, float z,
                                             float& e0, float& e1, float dy0[3], float dy1[3]) {
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
	e0 = __half2float(r0); e1 = __half2float(r1);
	// d/dx: pairs (c, c|1); weights scale*wy*wz in idx order (y bit, z bit)
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
		dy0[d] = a0; dy1[d] = a1;
	}
}

// y[r] = hq( sum_c W[r][c] x[c] ), optional ReLU
__device__ __forceinline__ void matvec(const __half* __restrict__ W, int rows, int cols, const float* x, float* y, bool relu) {
	for (int r = 0; r < rows; ++r) {
		const __half2* w2 = reinterpret_cast<const __half2*>(W + (size_t)r * cols);
		float acc = 0.f;
		for (int c = 0; c < cols; c += 2) { float2 w = __half22float2(__ldg(&w2[c >> 1])); acc = fmaf(w.x, x[c], acc); acc = fmaf(w.y, x[c + 1], acc); }
		if (relu && acc < 0.f) acc = 0.f;
		y[r] = hq(acc);
	}
}
// y[c] = hq( sum_r W[r][c] x[r] ), optionally masked by act[c] > 0
__device__ __forceinline__ void matvec_t(const __half* __restrict__ W, int rows, int cols, const float* x, float* y, const float* act) {
	for (int c = 0; c < cols; ++c) y[c] = 0.f;
	for (int r = 0; r < rows; ++r) {
		const float xr = x[r];
		if (xr == 0.f) continue;
		const __half2* w2 = reinterpret_cast<const __half2*>(W + (size_t)r * cols);
		for (int c = 0; c < cols; c += 2) { float2 w = __half22float2(__ldg(&w2[c >> 1])); y[c] = fmaf(w.x, xr, y[c]); y[c + 1] = fmaf(w.y, xr, y[c + 1]); }
	}
	for (int c = 0; c < cols; ++c) { float v = y[c]; if (act && !(act[c] > 0.f)) v = 0.f; y[c] = hq(v); }
}

struct SampleFwd {
	float u[48];          // SDF-MLP input  [x-0.5 | enc | 0]
	float hs[64];         // SDF hidden activation
	float y[16];
	float tm[64];         // relu'(hs) * W_out[0,:]
	float g[48];          // d sdf / d u
	float nrm[3];
};

// SDF branch: encoding, SDF MLP, analytic normal (nerf_network.h:139-189).  dydx (2L x 3) is optional.
__device__ void sdf_branch(const ModelDev& M, const __half* __restrict__ P, uint32_t valid_level, float x, float y, float z, SampleFwd& S, float (*dydx)[3]) {
	for (uint32_t i = 0; i < M.sdf_in; ++i) S.u[i] = 0.f;
	S.u[0] = __half2float(__hsub(__float2half_rn(x), __float2half_rn(0.5f)));
	S.u[1] = __half2float(__hsub(__float2half_rn(y), __float2half_rn(0.5f)));
	S.u[2] = __half2float(__hsub(__float2half_rn(z), __float2half_rn(0.5f)));
	float nacc[3] = {0.f, 0.f, 0.f};
	float dloc[32][3];
	for (uint32_t l = 0; l < M.n_levels; ++l) {
		float e0 = 0.f, e1 = 0.f, d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0};
		if (l <= valid_level) encode_level(M, P, l, x, y, z, e0, e1, d0, d1);
		S.u[3 + 2 * l] = e0; S.u[4 + 2 * l] = e1;
		for (int d = 0; d < 3; ++d) { dloc[2 * l][d] = d0[d]; dloc[2 * l + 1][d] = d1[d]; }
	}
	const LayerDesc& L0 = M.sdf_layers[0]; const LayerDesc& L1 = M.sdf_layers[1];
	matvec(P + L0.off, L0.rows, L0.cols, S.u, S.hs, true);
	matvec(P + L1.off, L1.rows, L1.cols, S.hs, S.y, false);
	for (uint32_t c = 0; c < L1.cols; ++c) S.tm[c] = S.hs[c] > 0.f ? __half2float(__ldg(P + L1.off + c)) : 0.f;
	matvec_t(P + L0.off, L0.rows, L0.cols, S.tm, S.g, nullptr);
	for (uint32_t k = 0; k < M.n_enc; ++k) for (int d = 0; d < 3; ++d) nacc[d] = fmaf(S.g[3 + k], dloc[k][d], nacc[d]);
	for (int d = 0; d < 3; ++d) S.nrm[d] = nacc[d] + S.g[d];
	if (dydx) for (uint32_t k = 0; k < M.n_enc; ++k) for (int d = 0; d < 3; ++d) dydx[k][d] = dloc[k][d];
}

__device__ __forceinline__ void store_out16(__half* out, const float* v) {
	__align__(16) __half h[16];
	for (int i = 0; i < 16; ++i) h[i] = __float2half_rn(v[i]);
	reinterpret_cast<uint4*>(out)[0] = reinterpret_cast<uint4*>(h)[0];
	reinterpret_cast<uint4*>(out)[1] = reinterpret_cast<uint4*>(h)[1];
}

// mode 0: pass A — only (sdf+bias, normal) as 4 binary16 into outA[n][4]
// mode 1: pass B — full 16-wide output row (nerf_network.h:221-250)
// mode 2: SDF probe — sdf / normal / density as fp32 (NerfNetwork::sdf, ::density, nerf_network.h:454-537)
__global__ void __launch_bounds__(128) k_forward_simt(ModelDev M, const __half* __restrict__ P, uint32_t valid_level, int mode,
                                                      const float4* __restrict__ pos4, const uint32_t* __restrict__ n_ptr, uint32_t n_max,
                                                      const float* __restrict__ ray_dirw /*3 per ray slot*/, __half* __restrict__ out,
                                                      float* __restrict__ sdf_f, float* __restrict__ nrm_f, float* __restrict__ dens_f) {
	const uint32_t n = n_ptr ? min(*n_ptr, n_max) : n_max;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
	const float4 p = pos4[i];
	SampleFwd S;
	sdf_branch(M, P, valid_level, p.x, p.y, p.z, S, nullptr);
	const float sdfb = __half2float(__hadd(__float2half_rn(S.y[0]), __float2half_rn(M.sdf_bias)));
	if (mode == 0) {
		__half2 a = __floats2half2_rn(sdfb, S.nrm[0]), b = __floats2half2_rn(S.nrm[1], S.nrm[2]);
		uint2 v; v.x = *reinterpret_cast<uint32_t*>(&a); v.y = *reinterpret_cast<uint32_t*>(&b);
		reinterpret_cast<uint2*>(out)[i] = v;
		continue;
	}
	if (mode == 2) {
		if (sdf_f) sdf_f[i] = sdfb;
		if (nrm_f) { nrm_f[3 * i] = S.nrm[0]; nrm_f[3 * i + 1] = S.nrm[1]; nrm_f[3 * i + 2] = S.nrm[2]; }
		if (dens_f) {   // sdf_to_density_variance_buffer, common_operation.cuh:310-328 (all binary16 arithmetic)
			const __half var = __ldg(P + M.off_var);
			const __half s = __float2half_rn(__expf(__half2float(__hmul(var, __float2half_rn(10.0f)))));
			const __half sg = __float2half_rn(1.0f / (1.0f + expf(-__half2float(__hmul(__float2half_rn(sdfb), s)))));
			dens_f[i] = __half2float(__hmul(__hmul(s, sg), __hsub(__float2half_rn(1.0f), sg)));
		}
		continue;
	}
	float rin[48], h1[64], h2[64], c[16], o[16];
	for (uint32_t k = 0; k < M.rgb_in; ++k) rin[k] = 0.f;
	for (int k = 0; k < 16; ++k) rin[k] = S.y[k];
	rin[32] = hq(p.x); rin[33] = hq(p.y); rin[34] = hq(p.z);
	rin[35] = hq(S.nrm[0]); rin[36] = hq(S.nrm[1]); rin[37] = hq(S.nrm[2]);
	const LayerDesc* L = M.rgb_layers;
	matvec(P + L[0].off, L[0].rows, L[0].cols, rin, h1, true);
	if (M.n_rgb_layers == 3) { matvec(P + L[1].off, L[1].rows, L[1].cols, h1, h2, true); matvec(P + L[2].off, L[2].rows, L[2].cols, h2, c, false); }
	else matvec(P + L[1].off, L[1].rows, L[1].cols, h1, c, false);
	for (int k = 0; k < 16; ++k) o[k] = c[k];
	o[3] = sdfb; o[4] = S.nrm[0]; o[5] = S.nrm[1]; o[6] = S.nrm[2];
	o[7] = __half2float(__ldg(P + M.off_var));
	const uint32_t slot = __float_as_uint(p.w);
	o[8] = ray_dirw[3 * slot]; o[9] = ray_dirw[3 * slot + 1]; o[10] = ray_dirw[3 * slot + 2];
	store_out16(out + (size_t)i * 16, o);
	}
}

// Scratch row layout (binary16) written by the SIMT backward for the weight-gradient GEMMs.
struct BwdScratch { __half *dc, *h2, *dh2, *h1, *dh1, *rin, *dy, *hs, *dhs, *u, *tm, *v; float* front1; };

// Forward recompute + backward for one sample (nerf_network.h:257-452): data gradients through both MLPs, merged first- and
// second-order hash-grid scatter, operands for the weight-gradient GEMMs.
__global__ void __launch_bounds__(128) k_backward_simt(ModelDev M, const __half* __restrict__ P, uint32_t valid_level,
                                                       const float4* __restrict__ pos4, const __half* __restrict__ dout16, const uint32_t* __restrict__ n_ptr, uint32_t n_max,
                                                       uint32_t n_roll /*roll-over batch (target / world)*/, uint32_t n_batch /*Eikonal divisor (global target)*/, const uint32_t* __restrict__ n_in_ptr, const uint32_t* __restrict__ gidx,
                                                       float* __restrict__ G, BwdScratch B) {
	const uint32_t n = n_ptr ? min(*n_ptr, n_max) : n_max;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
	const float4 p = pos4[i];
	SampleFwd S;
	float dydx[32][3];
	sdf_branch(M, P, valid_level, p.x, p.y, p.z, S, dydx);
	float rin[48], h1[64], h2[64], c[16];
	for (uint32_t k = 0; k < M.rgb_in; ++k) rin[k] = 0.f;
	for (int k = 0; k < 16; ++k) rin[k] = S.y[k];
	rin[32] = hq(p.x); rin[33] = hq(p.y); rin[34] = hq(p.z);
	rin[35] = hq(S.nrm[0]); rin[36] = hq(S.nrm[1]); rin[37] = hq(S.nrm[2]);
	const LayerDesc* L = M.rgb_layers;
	const bool three = M.n_rgb_layers == 3;
	matvec(P + L[0].off, L[0].rows, L[0].cols, rin, h1, true);
	if (three) { matvec(P + L[1].off, L[1].rows, L[1].cols, h1, h2, true); matvec(P + L[2].off, L[2].rows, L[2].cols, h2, c, false); }
	else matvec(P + L[1].off, L[1].rows, L[1].cols, h1, c, false);
	// incoming gradient, scaled by the roll-over multiplicity (fill_rollover_and_rescale, common_device.h:525-535)
	const uint32_t n_in = n_in_ptr ? *n_in_ptr : n;
	const float w = rollover_weight(gidx ? gidx[i] : i, min(n_in, n_roll), n_roll);
	float dout[16];
	{
		__align__(16) __half hraw[16];
		reinterpret_cast<uint4*>(hraw)[0] = reinterpret_cast<const uint4*>(dout16 + (size_t)i * 16)[0];
		reinterpret_cast<uint4*>(hraw)[1] = reinterpret_cast<const uint4*>(dout16 + (size_t)i * 16)[1];
		for (int k = 0; k < 16; ++k) dout[k] = hq(__half2float(hraw[k]) * w);
	}
	float dc[16]; for (int k = 0; k < 16; ++k) dc[k] = 0.f;
	dc[0] = dout[0]; dc[1] = dout[1]; dc[2] = dout[2];
	float dh2[64], dh1[64], drin[48];
	const LayerDesc& Ll = L[M.n_rgb_layers - 1];
	if (three) { matvec_t(P + Ll.off, Ll.rows, Ll.cols, dc, dh2, h2); matvec_t(P + L[1].off, L[1].rows, L[1].cols, dh2, dh1, h1); }
	else matvec_t(P + Ll.off, Ll.rows, Ll.cols, dc, dh1, h1);
	matvec_t(P + L[0].off, L[0].rows, L[0].cols, dh1, drin, nullptr);
	float dy[16];
	for (int k = 0; k < 16; ++k) dy[k] = drin[k];
	dy[0] = __half2float(__hadd(__float2half_rn(dy[0]), __float2half_rn(dout[3])));
	const LayerDesc& S0 = M.sdf_layers[0]; const LayerDesc& S1 = M.sdf_layers[1];
	float dhs[64], du[48];
	matvec_t(P + S1.off, S1.rows, S1.cols, dy, dhs, S.hs);
	matvec_t(P + S0.off, S0.rows, S0.cols, dhs, du, nullptr);
	atomicAdd(&G[M.off_var], dout[7]);
	float gn[3];
	for (int d = 0; d < 3; ++d) gn[d] = drin[35 + d] + dout[4 + d] / (float)n_batch + dout[8 + d];
	// hash-grid gradients, first + second order merged per corner
	for (uint32_t l = 0; l < M.n_levels && l <= valid_level; ++l) {
		float* gg = G + M.off_grid + (size_t)M.offsets[l] * 2;
		const uint32_t hsz = M.offsets[l + 1] - M.offsets[l], res = M.res[l];
		const float scale = M.scale[l];
		const LevelGeom g = level_geom(scale, p.x, p.y, p.z);
		const float d10 = du[3 + 2 * l], d11 = du[4 + 2 * l], ge0 = S.g[3 + 2 * l], ge1 = S.g[4 + 2 * l];
		const float wx[2] = {1.f - g.fx, g.fx}, wy[2] = {1.f - g.fy, g.fy}, wz[2] = {1.f - g.fz, g.fz};
		#pragma unroll
		for (int cidx = 0; cidx < 8; ++cidx) {
			const int bx = cidx & 1, by = (cidx >> 1) & 1, bz = (cidx >> 2) & 1;
			const float w1 = wx[bx] * wy[by] * wz[bz];
			const float w2 = scale * (gn[0] * (bx ? 1.f : -1.f) * wy[by] * wz[bz] + gn[1] * (by ? 1.f : -1.f) * wx[bx] * wz[bz] + gn[2] * (bz ? 1.f : -1.f) * wx[bx] * wy[by]);
			const uint32_t e = grid_entry(hsz, res, g.gx + bx, g.gy + by, g.gz + bz);
			const float v0 = d10 * w1 + ge0 * w2, v1 = d11 * w1 + ge1 * w2;
			if (v0 != 0.f || v1 != 0.f) atomicAdd(reinterpret_cast<float2*>(gg + 2 * e), make_float2(v0, v1));
		}
	}
	// second order through the SDF MLP (fully_fused_mlp.cu:1036-1142)
	float v[48];
	for (uint32_t k = 0; k < M.sdf_in; ++k) v[k] = 0.f;
	for (int d = 0; d < 3; ++d) v[d] = hq(gn[d]);
	for (uint32_t k = 0; k < M.n_enc; ++k) v[3 + k] = hq(dydx[k][0] * gn[0] + dydx[k][1] * gn[1] + dydx[k][2] * gn[2]);
	float f1[64];
	{
		const __half* W = P + S0.off;
		for (uint32_t r = 0; r < S0.rows; ++r) {
			float acc = 0.f;
			for (uint32_t cc = 0; cc < S0.cols; ++cc) acc = fmaf(__half2float(__ldg(W + (size_t)r * S0.cols + cc)), v[cc], acc);
			f1[r] = S.hs[r] > 0.f ? hq(acc) : 0.f;
		}
	}
	// operands for the weight-gradient GEMMs
	auto st = [&](__half* dst, const float* src, uint32_t wdt) { if (dst) for (uint32_t k = 0; k < wdt; ++k) dst[(size_t)i * wdt + k] = __float2half_rn(src[k]); };
	st(B.dc, dc, 16); st(B.h1, h1, M.rgb_width); st(B.dh1, dh1, M.rgb_width); st(B.rin, rin, M.rgb_in);
	if (three) { st(B.h2, h2, M.rgb_width); st(B.dh2, dh2, M.rgb_width); }
	st(B.dy, dy, 16); st(B.hs, S.hs, M.sdf_width); st(B.dhs, dhs, M.sdf_width); st(B.u, S.u, M.sdf_in);
	st(B.tm, S.tm, M.sdf_width); st(B.v, v, M.sdf_in);
	for (uint32_t k = 0; k < M.sdf_width; ++k) B.front1[(size_t)i * M.sdf_width + k] = f1[k];
	}
}

// dW[r][c] += sum_m dY[m][r] * X[m][c]  (fp32 accumulation over binary16 operands), one CTA per 1024-sample chunk.
__global__ void __launch_bounds__(256) k_dw_gemm(const __half* __restrict__ dY, const __half* __restrict__ X, uint32_t R, uint32_t C, const uint32_t* __restrict__ n_ptr, uint32_t n_max, float* __restrict__ dW) {
	__shared__ float sY[16][64], sX[16][64];
	const uint32_t n = n_ptr ? min(*n_ptr, n_max) : n_max;
	const uint32_t m0 = blockIdx.x * 1024, m1 = min(n, m0 + 1024);
	if (m0 >= n) return;
	float acc[16];
	#pragma unroll
	for (int q = 0; q < 16; ++q) acc[q] = 0.f;
	const uint32_t tid = threadIdx.x, nout = R * C;
	for (uint32_t m = m0; m < m1; m += 16) {
		for (uint32_t e = tid; e < 16 * R; e += 256) { uint32_t mm = e / R, r = e % R; sY[mm][r] = (m + mm < m1) ? __half2float(dY[(size_t)(m + mm) * R + r]) : 0.f; }
		for (uint32_t e = tid; e < 16 * C; e += 256) { uint32_t mm = e / C, c = e % C; sX[mm][c] = (m + mm < m1) ? __half2float(X[(size_t)(m + mm) * C + c]) : 0.f; }
		__syncthreads();
		#pragma unroll
		for (int q = 0; q < 16; ++q) {
			const uint32_t o = tid + 256 * q;
			if (o < nout) { const uint32_t r = o / C, c = o % C; float a = acc[q]; for (int mm = 0; mm < 16; ++mm) a = fmaf(sY[mm][r], sX[mm][c], a); acc[q] = a; }
		}
		__syncthreads();
	}
	#pragma unroll
	for (int q = 0; q < 16; ++q) { const uint32_t o = tid + 256 * q; if (o < nout && acc[q] != 0.f) atomicAdd(&dW[o], acc[q]); }
}
// dW[0][c] += sum_m F[m][c]  (second-order gradient of the SDF output row)
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ F, uint32_t C, const uint32_t* __restrict__ n_ptr, uint32_t n_max, float* __restrict__ dW) {
	const uint32_t n = n_ptr ? min(*n_ptr, n_max) : n_max;
	const uint32_t m0 = blockIdx.x * 1024, m1 = min(n, m0 + 1024);
	if (m0 >= n) return;
	__shared__ float s[256];
	const uint32_t c = threadIdx.x % C, part = threadIdx.x / C, nparts = 256 / C;
	float a = 0.f;
	for (uint32_t m = m0 + part; m < m1; m += nparts) a += F[(size_t)m * C + c];
	s[threadIdx.x] = a;
	__syncthreads();
	if (threadIdx.x < C) { float t = 0.f; for (uint32_t q = 0; q < nparts; ++q) t += s[q * C + threadIdx.x]; atomicAdd(&dW[threadIdx.x], t); }
}

void launch_forward_simt(cudaStream_t st, const ModelDev& M, const __half* P, uint32_t valid_level, int mode, const float4* pos4, const uint32_t* n_ptr, uint32_t n_max,
                         const float* ray_dirw, __half* out, float* sdf_f, float* nrm_f, float* dens_f) {
	if (!n_max) return;
	k_forward_simt<<<min((n_max + 127) / 128, 148u * 16u), 128, 0, st>>>(M, P, valid_level, mode, pos4, n_ptr, n_max, ray_dirw, out, sdf_f, nrm_f, dens_f);
}

void launch_backward_simt(cudaStream_t st, const ModelDev& M, const __half* P, uint32_t valid_level, const float4* pos4, const __half* dout16, const uint32_t* n_ptr, uint32_t n_max,
                          uint32_t n_roll, uint32_t n_batch, const uint32_t* n_in_ptr, const uint32_t* gidx, float* G, __half* scratch_h, float* scratch_f) {
	if (!n_max) return;
	BwdScratch B;
	size_t o = 0; const size_t N = n_max;
	auto take = [&](uint32_t w) { __half* p = scratch_h + o; o += N * w; return p; };
	B.dc = take(16); B.h2 = take(M.rgb_width); B.dh2 = take(M.rgb_width); B.h1 = take(M.rgb_width); B.dh1 = take(M.rgb_width); B.rin = take(M.rgb_in);
	B.dy = take(16); B.hs = take(M.sdf_width); B.dhs = take(M.sdf_width); B.u = take(M.sdf_in); B.tm = take(M.sdf_width); B.v = take(M.sdf_in);
	B.front1 = scratch_f;
	k_backward_simt<<<min((n_max + 127) / 128, 148u * 16u), 128, 0, st>>>(M, P, valid_level, pos4, dout16, n_ptr, n_max, n_roll, n_batch, n_in_ptr, gidx, G, B);
	const uint32_t nb = (n_max + 1023) / 1024;
	const LayerDesc* L = M.rgb_layers;
	const bool three = M.n_rgb_layers == 3;
	if (three) {
		k_dw_gemm<<<nb, 256, 0, st>>>(B.dc, B.h2, 16, M.rgb_width, n_ptr, n_max, G + L[2].off);
		k_dw_gemm<<<nb, 256, 0, st>>>(B.dh2, B.h1, M.rgb_width, M.rgb_width, n_ptr, n_max, G + L[1].off);
	} else k_dw_gemm<<<nb, 256, 0, st>>>(B.dc, B.h1, 16, M.rgb_width, n_ptr, n_max, G + L[1].off);
	k_dw_gemm<<<nb, 256, 0, st>>>(B.dh1, B.rin, M.rgb_width, M.rgb_in, n_ptr, n_max, G + L[0].off);
	k_dw_gemm<<<nb, 256, 0, st>>>(B.dy, B.hs, 16, M.sdf_width, n_ptr, n_max, G + M.sdf_layers[1].off);
	k_dw_gemm<<<nb, 256, 0, st>>>(B.dhs, B.u, M.sdf_width, M.sdf_in, n_ptr, n_max, G + M.sdf_layers[0].off);
	k_dw_gemm<<<nb, 256, 0, st>>>(B.tm, B.v, M.sdf_width, M.sdf_in, n_ptr, n_max, G + M.sdf_layers[0].off);
	k_colsum<<<nb, 256, 0, st>>>(B.front1, M.sdf_width, n_ptr, n_max, G + M.sdf_layers[1].off);
}
size_t backward_simt_scratch_halfs(const ModelDev& M) { return 16 + 4 * M.rgb_width + M.rgb_in + 16 + 3 * M.sdf_width + 2 * M.sdf_in; }

} // namespace rnb
