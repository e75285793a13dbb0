// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h header).
//
// Restatement of the network composition on the hot path:
//   hash-grid encoding + analytic dy/dx   tcnn encodings/grid.h:113-148,169-364
//   fully fused MLPs (no bias, ReLU)       tcnn src/fully_fused_mlp.cu:624-911
//   NerfNetwork::forward_impl              include/neural-graphics-primitives/nerf_network.h:97-253
//   NerfNetwork::backward_impl (+2nd order) nerf_network.h:257-452, grid.h:366-495,556-683,858-883,
//                                          fully_fused_mlp.cu:913-1031,1036-1142
// Template parameter Q: true  = reference numerics (fp32 math, explicit binary16 rounding points)
//                       false = smooth double precision (used only by the finite-difference checks)
#pragma once
#include "orc_common.h"

namespace orc {

struct Layer { uint32_t rows, cols; size_t off; };

struct ModelDesc {
	uint32_t n_levels = 14, log2_hashmap = 19, base_res = 16;
	float per_level_scale = 1.4524226f;
	uint32_t sdf_width = 64, sdf_hidden = 1, rgb_width = 64, rgb_hidden = 2;
	float sdf_bias = -0.1f;
	// derived
	uint32_t n_enc = 28, sdf_in = 32, rgb_in = 48;
	std::vector<uint32_t> offsets, res;
	std::vector<float> scale;
	std::vector<Layer> sdf_layers, rgb_layers;
	size_t off_sdf = 0, off_rgb = 0, off_grid = 0, off_var = 0, n_params = 0, n_grid_params = 0;

	// per-level tables: grid.h:977-1013; per_level_scale from src/testbed.cu:2319-2323
	void finalize() {
		n_enc = n_levels * 2;
		offsets.assign(n_levels + 1, 0); res.assign(n_levels, 0); scale.assign(n_levels, 0.f);
		uint32_t offset = 0;
		for (uint32_t i = 0; i < n_levels; ++i) {
			const float s = exp2f(i * std::log2(per_level_scale)) * base_res - 1.0f;
			const uint32_t r = (uint32_t)(ceilf(s)) + 1;
			scale[i] = (float)(r - 1);
			res[i] = r;
			uint32_t max_params = 0xFFFFFFFFu / 2;
			uint32_t params_in_level = std::pow((float)r, 3.f) > (float)max_params ? max_params : r * r * r;
			params_in_level = next_multiple(params_in_level, 8u);
			params_in_level = std::min(params_in_level, 1u << log2_hashmap);
			offsets[i] = offset;
			offset += params_in_level;
		}
		offsets[n_levels] = offset;
		n_grid_params = (size_t)offset * 2;
		sdf_in = next_multiple(3 + n_enc, 16u);           // nerf_network.h:47
		rgb_in = next_multiple(3 + 3 + 16 + 16, 16u);     // nerf_network.h:60 (dir-encoding slot is 16 wide, zeros)
		size_t off = 0;
		off_sdf = off;
		sdf_layers.clear(); rgb_layers.clear();
		auto add_mlp = [&](std::vector<Layer>& ls, uint32_t in, uint32_t width, uint32_t hidden) {
			ls.push_back({width, in, off}); off += (size_t)width * in;
			for (uint32_t h = 1; h < hidden; ++h) { ls.push_back({width, width, off}); off += (size_t)width * width; }
			ls.push_back({16, width, off}); off += (size_t)16 * width;
		};
		add_mlp(sdf_layers, sdf_in, sdf_width, sdf_hidden);
		off_rgb = off;
		add_mlp(rgb_layers, rgb_in, rgb_width, rgb_hidden);
		off_grid = off; off += n_grid_params;
		off_var = off; off += 4;                           // TrainableBuffer<1,1,T>{4}: nerf_network.h:70
		n_params = off;
	}
};

// progressive level schedule — grid.h:1430-1437
static inline uint32_t valid_level_for_step(const ModelDesc& m, int step, float base_scale = 0.2f, float lvl_scale = 0.02f, uint32_t base_step = 100) {
	if (step <= 0) return m.n_levels;
	float v = base_scale * (float)m.n_levels + lvl_scale * (float)std::max(0, (int)((uint32_t)step - base_step));
	return std::min(m.n_levels, (uint32_t)std::ceil(v));
}

// grid.h:113-148
static inline uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
	uint32_t stride = 1, index = 0;
	for (uint32_t d = 0; d < 3 && stride <= hashmap_size; ++d) {
		index += p[d] * stride;
		stride *= res;
	}
	if (hashmap_size < stride) {
		index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
	}
	return (index % hashmap_size) * 2;
}

static const int MAXW = 64;   // widest layer / input
static const int MAXL = 16;   // max hash levels

template <bool Q>
struct Net {
	typedef typename std::conditional<Q, float, double>::type real;
	static inline real h(real x) { return Q ? (real)hq((float)x) : x; }

	const ModelDesc& m;
	const real* P;            // parameters as values (Q: the fp16 copy widened to float)
	uint32_t valid_level;
	Net(const ModelDesc& m_, const real* P_, uint32_t vl) : m(m_), P(P_), valid_level(vl) {}

	struct Ctx {
		real frac[MAXL][3]; uint32_t pg[MAXL][3];
		real u[MAXW];
		real act_sdf[3][MAXW];
		real y[16];
		real g[MAXW];                // dSDF/d(sdf_in) (one-hot backward)
		real tmask[3][MAXW];         // masked back-chain of the one-hot, per hidden layer
		real dydx[2 * MAXL][3];
		real normal[3];
		real rin[MAXW];
		real act_rgb[3][MAXW];
		real c[16];
		real out[16];
	};

	// y = W x ; W row-major [rows x cols]; fp32 accumulate, rounded at the layer output
	void matvec(const Layer& L, const real* x, real* y, bool relu) const {
		const real* W = P + L.off;
		for (uint32_t r = 0; r < L.rows; ++r) {
			real acc = 0;
			for (uint32_t c = 0; c < L.cols; ++c) acc += W[(size_t)r * L.cols + c] * x[c];
			if (relu && acc < 0) acc = 0;
			y[r] = h(acc);
		}
	}
	// y = W^T x, optionally masked by act>0
	void matvec_t(const Layer& L, const real* x, real* y, const real* act_mask) const {
		const real* W = P + L.off;
		for (uint32_t c = 0; c < L.cols; ++c) {
			real acc = 0;
			for (uint32_t r = 0; r < L.rows; ++r) acc += W[(size_t)r * L.cols + c] * x[r];
			if (act_mask && !(act_mask[c] > 0)) acc = 0;
			y[c] = h(acc);
		}
	}

	void mlp_forward(const std::vector<Layer>& ls, const real* in, real act[][MAXW], real* out) const {
		const real* cur = in;
		for (size_t i = 0; i + 1 < ls.size(); ++i) { matvec(ls[i], cur, act[i], true); cur = act[i]; }
		matvec(ls.back(), cur, out, false);
	}

	// encoding of one level: kernel_grid (grid.h:169-364)
	void encode_level(uint32_t l, const real xyz[3], Ctx& c, real enc[2]) const {
		if (l > valid_level) {
			enc[0] = enc[1] = 0;
			for (int f = 0; f < 2; ++f) for (int d = 0; d < 3; ++d) c.dydx[l * 2 + f][d] = 0;
			for (int d = 0; d < 3; ++d) { c.frac[l][d] = 0; c.pg[l][d] = 0; }
			return;
		}
		const real* grid = P + m.off_grid + (size_t)m.offsets[l] * 2;
		const uint32_t hsz = m.offsets[l + 1] - m.offsets[l];
		const real scale = (real)m.scale[l];
		const uint32_t res = m.res[l];
		for (int d = 0; d < 3; ++d) {          // pos_fract: common_device.h:415-424
			real p = std::fma(xyz[d], scale, (real)0.5);   // FMA: see orc_render.h ray_setup
			int t = (int)std::floor(p);
			c.pg[l][d] = (uint32_t)t;
			c.frac[l][d] = p - (real)t;
		}
		const real* fr = c.frac[l]; const uint32_t* pg = c.pg[l];
		real r0 = 0, r1 = 0;
		for (uint32_t idx = 0; idx < 8; ++idx) {
			real w = 1; uint32_t pl[3];
			for (uint32_t d = 0; d < 3; ++d) {
				if ((idx & (1u << d)) == 0) { w *= 1 - fr[d]; pl[d] = pg[d]; }
				else { w *= fr[d]; pl[d] = pg[d] + 1; }
			}
			uint32_t gi = grid_index(hsz, res, pl);
			// binary16 accumulation of (T)(weight*data): grid.h:309-314
			if (Q) { r0 = (real)hadd((float)r0, hq((float)(w * grid[gi]))); r1 = (real)hadd((float)r1, hq((float)(w * grid[gi + 1]))); }
			else { r0 += w * grid[gi]; r1 += w * grid[gi + 1]; }
		}
		enc[0] = r0; enc[1] = r1;
		// dy/dx: grid.h:324-363
		real gr[2][3] = {{0, 0, 0}, {0, 0, 0}};
		for (uint32_t gd = 0; gd < 3; ++gd) {
			for (uint32_t idx = 0; idx < 4; ++idx) {
				real w = scale; uint32_t pl[3];
				for (uint32_t nd = 0; nd < 2; ++nd) {
					const uint32_t d = nd >= gd ? nd + 1 : nd;
					if ((idx & (1u << nd)) == 0) { w *= 1 - fr[d]; pl[d] = pg[d]; }
					else { w *= fr[d]; pl[d] = pg[d] + 1; }
				}
				pl[gd] = pg[gd];     uint32_t il = grid_index(hsz, res, pl);
				pl[gd] = pg[gd] + 1; uint32_t ir = grid_index(hsz, res, pl);
				for (int f = 0; f < 2; ++f) gr[f][gd] += w * (grid[ir + f] - grid[il + f]) * (real)1;
			}
		}
		for (int f = 0; f < 2; ++f) for (int d = 0; d < 3; ++d) c.dydx[l * 2 + f][d] = gr[f][d];
	}

	// one-hot back chain through the SDF MLP: returns g = d(y0)/d(sdf_in); keeps masked chain in c.tmask
	void sdf_onehot_backward(Ctx& c) const {
		const auto& ls = m.sdf_layers;
		const size_t nh = ls.size() - 1;          // hidden layers
		real e0[16] = {0}; e0[0] = 1;
		const real* cur = e0;
		for (size_t i = nh; i >= 1; --i) {        // layer i (rows = 16 or width) transposed, masked by act[i-1]
			matvec_t(ls[i], cur, c.tmask[i - 1], c.act_sdf[i - 1]);
			cur = c.tmask[i - 1];
		}
		matvec_t(ls[0], cur, c.g, nullptr);
	}

	// NerfNetwork::forward_impl — nerf_network.h:97-253.  coord: pos3, dt, dir3 (all warped to [0,1])
	void forward(const float coord[7], Ctx& c, bool with_rgb = true) const {
		real xyz[3] = {(real)coord[0], (real)coord[1], (real)coord[2]};
		for (uint32_t i = 0; i < m.sdf_in; ++i) c.u[i] = 0;
		for (int d = 0; d < 3; ++d) c.u[d] = Q ? (real)hsub(hq((float)xyz[d]), 0.5f) : xyz[d] - (real)0.5;   // common_operation.cuh:186-199
		for (uint32_t l = 0; l < m.n_levels; ++l) { real e[2]; encode_level(l, xyz, c, e); c.u[3 + 2 * l] = e[0]; c.u[3 + 2 * l + 1] = e[1]; }
		mlp_forward(m.sdf_layers, c.u, c.act_sdf, c.y);
		sdf_onehot_backward(c);
		// normal = dy/dx^T g_enc (fp32, kernel_grid_backward_input grid.h:527-554) + g[0:3]
		real n[3] = {0, 0, 0};
		for (uint32_t k = 0; k < m.n_enc; ++k) for (int d = 0; d < 3; ++d) n[d] += c.g[3 + k] * c.dydx[k][d];
		for (int d = 0; d < 3; ++d) c.normal[d] = n[d] + c.g[d];
		for (uint32_t i = 0; i < m.rgb_in; ++i) c.rin[i] = 0;
		for (int i = 0; i < 16; ++i) c.rin[i] = c.y[i];
		for (int d = 0; d < 3; ++d) { c.rin[32 + d] = h(xyz[d]); c.rin[35 + d] = h(c.normal[d]); }
		if (with_rgb) mlp_forward(m.rgb_layers, c.rin, c.act_rgb, c.c);
		else for (int i = 0; i < 16; ++i) c.c[i] = 0;
		for (int i = 0; i < 16; ++i) c.out[i] = c.c[i];
		c.out[3] = Q ? (real)hadd((float)c.y[0], hq(m.sdf_bias)) : c.y[0] + (real)m.sdf_bias;
		for (int d = 0; d < 3; ++d) c.out[4 + d] = h(c.normal[d]);
		c.out[7] = P[m.off_var];
		for (int d = 0; d < 3; ++d) c.out[8 + d] = h((real)coord[4 + d]);
	}

	// NerfNetwork::sdf — nerf_network.h:454-520 (row 0 only): raw SDF + bias in binary16
	real sdf_only(const float xyz_f[3]) const {
		Ctx c;
		float coord[7] = {xyz_f[0], xyz_f[1], xyz_f[2], 0, 0, 0, 0};
		real xyz[3] = {(real)coord[0], (real)coord[1], (real)coord[2]};
		for (uint32_t i = 0; i < m.sdf_in; ++i) c.u[i] = 0;
		for (int d = 0; d < 3; ++d) c.u[d] = Q ? (real)hsub(hq((float)xyz[d]), 0.5f) : xyz[d] - (real)0.5;
		for (uint32_t l = 0; l < m.n_levels; ++l) { real e[2]; encode_level(l, xyz, c, e); c.u[3 + 2 * l] = e[0]; c.u[3 + 2 * l + 1] = e[1]; }
		mlp_forward(m.sdf_layers, c.u, c.act_sdf, c.y);
		return Q ? (real)hadd((float)c.y[0], hq(m.sdf_bias)) : c.y[0] + (real)m.sdf_bias;
	}

	// MLP backward: dW += d_out ⊗ in ; returns d_in (rounded).  fully_fused_mlp.cu:913-1031
	void mlp_backward(const std::vector<Layer>& ls, const real* in, real act[][MAXW], const real* dout, real* din, real* G, real wscale) const {
		real cur[MAXW], nxt[MAXW];
		const size_t nl = ls.size();
		for (uint32_t i = 0; i < ls.back().rows; ++i) cur[i] = dout[i];
		for (size_t li = nl; li-- > 0;) {
			const Layer& L = ls[li];
			const real* x = li == 0 ? in : act[li - 1];
			if (G) {
				real* g = G + L.off;
				for (uint32_t r = 0; r < L.rows; ++r) { real d = cur[r] * wscale; if (d == 0) continue; for (uint32_t cc = 0; cc < L.cols; ++cc) g[(size_t)r * L.cols + cc] += d * x[cc]; }
			}
			if (li == 0) { if (din) matvec_t(L, cur, din, nullptr); }
			else { matvec_t(L, cur, nxt, act[li - 1]); for (uint32_t i = 0; i < L.cols; ++i) cur[i] = nxt[i]; }
		}
	}

	// NerfNetwork::backward_impl for one sample.  dout: 16 values (binary16 in Q mode, already × loss scale).
	// w: roll-over multiplicity weight of the sample (common_device.h:525-535), applied to every gradient.
	// Gm: gradient accumulator indexed by MLP parameter offsets (may be a thread-private buffer);
	// G: accumulator indexed by full parameter offsets, used for the hash grid (atomic adds when `atomic`); gvar: variance grad.
	void backward(const Ctx& c_in, const real dout_in[16], real w, uint32_t n_batch, real* Gm, real* G, real* gvar, bool atomic) const {
		Ctx& c = const_cast<Ctx&>(c_in);
		real dout[16];
		for (int i = 0; i < 16; ++i) dout[i] = Q ? h(dout_in[i] * w) : dout_in[i] * w;   // fill_rollover_and_rescale rounds the scaled copy
		real dc[16] = {0}; dc[0] = dout[0]; dc[1] = dout[1]; dc[2] = dout[2];             // extract_rgb: common_operation.cuh:1010-1025
		real drin[MAXW];
		mlp_backward(m.rgb_layers, c.rin, c.act_rgb, dc, drin, Gm, 1);
		real dy[16];
		for (int i = 0; i < 16; ++i) dy[i] = drin[i];
		dy[0] = Q ? (real)hadd((float)dy[0], (float)dout[3]) : dy[0] + dout[3];           // add_density_gradient
		real du[MAXW];
		mlp_backward(m.sdf_layers, c.u, c.act_sdf, dy, du, Gm, 1);
		*gvar += dout[7];
		// g_n: nerf_network.h:343-373
		real gn[3];
		for (int d = 0; d < 3; ++d) gn[d] = drin[35 + d] + dout[4 + d] / (real)n_batch + dout[8 + d];
		// hash-grid gradient: first order (grid.h:366-495) + second order via dy/dx (grid.h:556-683), merged per corner
		for (uint32_t l = 0; l < m.n_levels; ++l) {
			if (l > valid_level) continue;
			real* gg = G + m.off_grid + (size_t)m.offsets[l] * 2;
			const uint32_t hsz = m.offsets[l + 1] - m.offsets[l];
			const uint32_t res = m.res[l];
			const real scale = (real)m.scale[l];
			const real* fr = c.frac[l]; const uint32_t* pg = c.pg[l];
			const real d1[2] = {du[3 + 2 * l], du[3 + 2 * l + 1]};      // dL/denc
			const real ge[2] = {c.g[3 + 2 * l], c.g[3 + 2 * l + 1]};    // dSDF/denc
			for (uint32_t idx = 0; idx < 8; ++idx) {
				uint32_t pl[3]; real wd[3]; real sg[3];
				for (uint32_t d = 0; d < 3; ++d) {
					if ((idx & (1u << d)) == 0) { wd[d] = 1 - fr[d]; pl[d] = pg[d]; sg[d] = -1; }
					else { wd[d] = fr[d]; pl[d] = pg[d] + 1; sg[d] = 1; }
				}
				const real w1 = wd[0] * wd[1] * wd[2];
				const real w2 = scale * (gn[0] * sg[0] * wd[1] * wd[2] + gn[1] * sg[1] * wd[0] * wd[2] + gn[2] * sg[2] * wd[0] * wd[1]);
				uint32_t gi = grid_index(hsz, res, pl);
				for (int f = 0; f < 2; ++f) {
					real v = d1[f] * w1 + ge[f] * w2;
					if (atomic && Q) atomic_add_f32((float*)&gg[gi + f], (float)v); else gg[gi + f] += v;
				}
			}
		}
		// second order through the SDF MLP: fully_fused_mlp.cu:1036-1142
		real v[MAXW];
		for (uint32_t i = 0; i < m.sdf_in; ++i) v[i] = 0;
		for (int d = 0; d < 3; ++d) v[d] = h(gn[d]);
		for (uint32_t k = 0; k < m.n_enc; ++k) {            // kernel_grid_backward_input_backward_dLdoutput grid.h:858-883
			real r = 0;
			for (int d = 0; d < 3; ++d) r += c.dydx[k][d] * gn[d];
			v[3 + k] = h(r);
		}
		const auto& ls = m.sdf_layers;
		const size_t nh = ls.size() - 1;
		real front[4][MAXW];
		for (uint32_t i = 0; i < m.sdf_in; ++i) front[0][i] = v[i];
		for (size_t i = 1; i <= nh; ++i) {                  // front[i] = relu'(act[i-1]) ⊙ (M_{i-1} front[i-1])
			const Layer& L = ls[i - 1];
			const real* W = P + L.off;
			for (uint32_t r = 0; r < L.rows; ++r) {
				real acc = 0;
				for (uint32_t cc = 0; cc < L.cols; ++cc) acc += W[(size_t)r * L.cols + cc] * front[i - 1][cc];
				if (!(c.act_sdf[i - 1][r] > 0)) acc = 0;
				front[i][r] = h(acc);
			}
		}
		// back[i+1] for matrix i: matrix nh (output) ← e0 ; matrix i<nh ← tmask[i]
		for (size_t i = 0; i <= nh; ++i) {
			const Layer& L = ls[i];
			real* g = Gm + L.off;
			if (i == nh) {
				for (uint32_t cc = 0; cc < L.cols; ++cc) g[cc] += front[i][cc];      // row 0 only (one-hot)
			} else {
				for (uint32_t r = 0; r < L.rows; ++r) { real b = c.tmask[i][r]; if (b == 0) continue; for (uint32_t cc = 0; cc < L.cols; ++cc) g[(size_t)r * L.cols + cc] += b * front[i][cc]; }
			}
		}
	}
};

} // namespace orc
