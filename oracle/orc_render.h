// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_common.h header).
//
// Restatement of the ray / compositing / loss side of the step:
//   image_idx, pixel sampling         src/testbed_nerf.cu:1171-1214
//   read_rgba, sRGB curves            include/neural-graphics-primitives/common_device.cuh:31-61,665-700
//   occupancy lookup + DDA marching   src/testbed_nerf.cu:140-155,301-323,439-465,569-583,1216-1387
//   NeuS compositing + RNb losses     src/testbed_nerf.cu:1396-2097
#pragma once
#include "orc_common.h"

namespace orc {

static constexpr uint32_t GRIDSIZE = 128, CASCADES = 8, NSTEPS = 1024, N_RNG_PER_RAY = 8;
static constexpr float SQRT3f = 1.73205080757f;
static inline float MIN_STEP() { return SQRT3f / NSTEPS; }
static inline float MAX_STEP() { return MIN_STEP() * (1 << (CASCADES - 1)) * NSTEPS / GRIDSIZE; }
static constexpr float MIN_OPTICAL_THICKNESS = 0.1f;

struct View {
	const void* normal_px;   // uint16 RGBA
	const void* albedo_px;   // uint16 RGBA (may be null when no_albedo)
	int32_t w, h;
	float fx, fy;            // focal length, pixels
	float cx, cy;            // principal point, normalised to [0,1]
	float xform[12];         // 3x4 camera-to-world (NGP frame), column-major
};

struct Flags {
	int32_t apply_L2 = 1, apply_supernormal = 0, apply_rgbplus = 1, apply_relu = 0, apply_bce = 0, light_opti = 0, no_albedo = 1;
	float mask_loss_weight = 1.0f, ek_loss_weight = 0.01f, cos_anneal_ratio = 1.0f;
	int32_t light_mode = -1;   // -1: hashed per (ray, step); -2: ray index % 3 (the pin oracle/ref_prelude.h gives the reference); 0..2: pinned light index
};

static inline float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : std::pow((s + 0.055f) / 1.055f, 2.4f); }
static inline float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * std::pow(l, 0.41666f) - 0.055f; }

static inline void read_rgba(const void* pixels, int w, int h, float x, float y, float out[4]) {
	int px = std::max(0, std::min(w - 1, (int)(x * (float)w)));
	int py = std::max(0, std::min(h - 1, (int)(y * (float)h)));
	uint16_t v[4];
	std::memcpy(v, (const uint8_t*)pixels + ((size_t)px + (size_t)py * w) * 8, 8);
	uint64_t raw; std::memcpy(&raw, v, 8);
	if (raw == 0x00FF00FFull) { out[0] = out[1] = out[2] = out[3] = -1.0f; return; }
	float a = (float)v[3] * (1.0f / 65535.0f);
	for (int c = 0; c < 3; ++c) out[c] = srgb_to_linear((float)v[c] * (1.0f / 65535.0f)) * a;
	out[3] = a;
}

static inline uint32_t image_idx(uint32_t base, uint32_t n_rays, uint32_t n_rays_total, uint32_t n_images) {
	return (((base + n_rays_total) * n_images) / n_rays) % n_images;   // uint32 wrap-around is intended
}

static inline void pixel_pos(Pcg32& rng, int w, int h, float xy[2]) {
	float u = rng.next_float(), v = rng.next_float();
	float px = std::min(std::max(u * (float)w, 0.0f), (float)(w - 1));
	float py = std::min(std::max(v * (float)h, 0.0f), (float)(h - 1));
	xy[0] = (px + 0.5f) / (float)w;
	xy[1] = (py + 0.5f) / (float)h;
}

static inline int mip_from_pos(const float p[3]) {
	float mx = std::max(std::fabs(p[0] - 0.5f), std::max(std::fabs(p[1] - 0.5f), std::fabs(p[2] - 0.5f)));
	int e; std::frexp(mx, &e);
	return std::min((int)CASCADES - 1, std::max(0, e + 1));
}
static inline int mip_from_dt(float dt, const float p[3]) {
	int mip = mip_from_pos(p);
	dt *= 2 * GRIDSIZE;
	if (dt < 1.f) return mip;
	int e; std::frexp(dt, &e);
	return std::min((int)CASCADES - 1, std::max(e, mip));
}
static inline uint32_t cascaded_grid_idx_at(const float pos[3], uint32_t mip) {
	float ms = std::scalbn(1.0f, -(int)mip);
	int i[3];
	for (int d = 0; d < 3; ++d) {
		float p = pos[d] - 0.5f; p *= ms; p += 0.5f;
		i[d] = (int)(p * (float)GRIDSIZE);
		i[d] = std::max(0, std::min((int)GRIDSIZE - 1, i[d]));
	}
	return morton3D((uint32_t)i[0], (uint32_t)i[1], (uint32_t)i[2]);
}
static inline bool occupied_at(const float pos[3], const uint8_t* bitfield, uint32_t mip) {
	uint32_t idx = cascaded_grid_idx_at(pos, mip);
	return bitfield[idx / 8 + (GRIDSIZE * GRIDSIZE * GRIDSIZE) * mip / 8] & (1 << (idx % 8));
}
static inline float sgn(float x) { return std::copysign(1.0f, x); }
static inline float advance_to_next_voxel(float t, const float pos[3], const float dir[3], const float idir[3], uint32_t res) {
	float tt[3];
	for (int d = 0; d < 3; ++d) {
		float p = (float)res * pos[d];
		tt[d] = (std::floor(p + 0.5f + 0.5f * sgn(dir[d])) - p) * idir[d];
	}
	float tm = std::min(std::min(tt[0], tt[1]), tt[2]);
	float t_target = t + std::max(tm / (float)res, 0.0f);
	do { t += MIN_STEP(); } while (t < t_target);   // calc_dt == MIN_STEP when cone_angle_constant == 0
	return t;
}
static inline bool in_unit_cube(const float p[3]) {
	return p[0] >= 0.f && p[0] <= 1.f && p[1] >= 0.f && p[1] <= 1.f && p[2] >= 0.f && p[2] <= 1.f;
}
// BoundingBox::ray_intersect for the unit cube — bounding_box.cuh:163-214
static inline void ray_unit_cube(const float o[3], const float d[3], float& tmin_o, float& tmax_o) {
	const float BIG = std::numeric_limits<float>::max();
	float tmin = (0.f - o[0]) / d[0], tmax = (1.f - o[0]) / d[0];
	if (tmin > tmax) std::swap(tmin, tmax);
	float tymin = (0.f - o[1]) / d[1], tymax = (1.f - o[1]) / d[1];
	if (tymin > tymax) std::swap(tymin, tymax);
	if (tmin > tymax || tymin > tmax) { tmin_o = tmax_o = BIG; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (0.f - o[2]) / d[2], tzmax = (1.f - o[2]) / d[2];
	if (tzmin > tzmax) std::swap(tzmin, tzmax);
	if (tmin > tzmax || tzmin > tmax) { tmin_o = tmax_o = BIG; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	tmin_o = tmin; tmax_o = tmax;
}

struct RayGen {
	bool valid; uint32_t img; float xy[2]; float o[3]; float d_un[3]; float dir[3]; float startt; uint32_t numsteps;
};

// first half of generate_training_samples_nerf (up to the counting march)
static inline void ray_setup(uint32_t i, uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng, const View* views, uint32_t n_views, const uint8_t* bitfield, RayGen& r) {
	r.valid = false; r.numsteps = 0;
	uint32_t img = image_idx(i, n_rays, n_rays_total, n_views);
	const View& v = views[img];
	rng.advance((int64_t)i * N_RNG_PER_RAY);
	pixel_pos(rng, v.w, v.h, r.xy);
	float px[4];
	read_rgba(v.normal_px, v.w, v.h, r.xy[0], r.xy[1], px);
	if (px[0] <= 0.0f && rng.next_float() >= 0.9) return;     // short-circuit: the draw happens only for background pixels
	(void)rng.next_float();                                       // motion-blur time (unused)
	float dc[3] = {(r.xy[0] - v.cx) * (float)v.w / v.fx, (r.xy[1] - v.cy) * (float)v.h / v.fy, 1.0f};
	const float* X = v.xform;
	// nvcc contracts a*b+c into FMA by default (-fmad=true); the restatement spells the fused form out so that sample
	// positions — and therefore occupancy cell indices — are bit-identical to the CUDA path.
	for (int k = 0; k < 3; ++k) { r.d_un[k] = std::fma(X[6 + k], dc[2], std::fma(X[3 + k], dc[1], X[0 + k] * dc[0])); r.o[k] = X[9 + k]; }
	float nrm = std::sqrt(std::fma(r.d_un[2], r.d_un[2], std::fma(r.d_un[1], r.d_un[1], r.d_un[0] * r.d_un[0])));
	for (int k = 0; k < 3; ++k) r.dir[k] = r.d_un[k] / nrm;
	float tmin, tmax; ray_unit_cube(r.o, r.dir, tmin, tmax);
	tmin = std::max(tmin, 0.0f);
	r.startt = std::fma(MIN_STEP(), rng.next_float(), tmin);
	r.img = img;
	float idir[3] = {1.0f / r.dir[0], 1.0f / r.dir[1], 1.0f / r.dir[2]};
	uint32_t j = 0; float t = r.startt; float pos[3];
	for (;;) {
		for (int k = 0; k < 3; ++k) pos[k] = std::fma(t, r.dir[k], r.o[k]);
		if (!in_unit_cube(pos) || j >= NSTEPS) break;
		float dt = MIN_STEP();
		uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
		if (occupied_at(pos, bitfield, mip)) { ++j; t += dt; }
		else t = advance_to_next_voxel(t, pos, r.dir, idir, GRIDSIZE >> mip);
	}
	r.numsteps = j;
	r.valid = j > 0;
}

// second march: emit coords (pos3, warped dt = 0, warped dir3)
static inline void ray_emit(const RayGen& r, const uint8_t* bitfield, float* coords /* 7*numsteps */) {
	float idir[3] = {1.0f / r.dir[0], 1.0f / r.dir[1], 1.0f / r.dir[2]};
	float wd[3] = {(r.dir[0] + 1.0f) * 0.5f, (r.dir[1] + 1.0f) * 0.5f, (r.dir[2] + 1.0f) * 0.5f};
	float max_step = MIN_STEP() * (1 << (CASCADES - 1));
	uint32_t j = 0; float t = r.startt; float pos[3];
	for (;;) {
		for (int k = 0; k < 3; ++k) pos[k] = std::fma(t, r.dir[k], r.o[k]);
		if (!in_unit_cube(pos) || j >= r.numsteps) break;
		float dt = MIN_STEP();
		uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
		if (occupied_at(pos, bitfield, mip)) {
			float* c = coords + (size_t)j * 7;
			c[0] = pos[0]; c[1] = pos[1]; c[2] = pos[2];
			c[3] = (dt - MIN_STEP()) / (max_step - MIN_STEP());
			c[4] = wd[0]; c[5] = wd[1]; c[6] = wd[2];
			++j; t += dt;
		} else t = advance_to_next_voxel(t, pos, r.dir, idir, GRIDSIZE >> mip);
	}
}

static inline uint32_t hashed_light(uint32_t ray_idx, uint32_t step) {
	uint32_t h = ray_idx * 0x9E3779B1u ^ (step * 0x85EBCA77u) ^ 0xC2B2AE3Du;
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h % 3u;
}

static inline float logistic(float x) { return 1.0f / (1.0f + std::exp(-x)); }
static inline float reluf(float x) { return x > 0.f ? x : 0.f; }

struct RayTarget { float rgbt[4]; float light[3]; float mask_certainty, mask_gt; };

// per-ray target + light set-up: testbed_nerf.cu:1485-1593
static inline void ray_target(uint32_t ray_idx, uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng, const View* views, uint32_t n_views, const Flags& F, uint32_t step, RayTarget& T) {
	rng.advance((int64_t)ray_idx * N_RNG_PER_RAY);
	uint32_t img = image_idx(ray_idx, n_rays, n_rays_total, n_views);
	const View& v = views[img];
	float xy[2]; pixel_pos(rng, v.w, v.h, xy);
	float tn[4], ta[4] = {0, 0, 0, 0};
	read_rgba(v.normal_px, v.w, v.h, xy[0], xy[1], tn);
	if (v.albedo_px) read_rgba(v.albedo_px, v.w, v.h, xy[0], xy[1], ta); else { ta[0] = ta[1] = ta[2] = tn[3]; ta[3] = tn[3]; }
	float nv[3];
	for (int c = 0; c < 3; ++c) nv[c] = linear_to_srgb(1.0f * tn[c]) * 2.0f - 1.0f;   // exposure scale = exp(0) = 1
	nv[1] *= -1; nv[2] *= -1;
	float nn = std::sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
	for (int c = 0; c < 3; ++c) nv[c] /= nn;
	float alb[4];
	if (F.no_albedo) { alb[0] = alb[1] = alb[2] = 1.0f; alb[3] = 0.0f; }
	else {
		float a3[3]; for (int c = 0; c < 3; ++c) a3[c] = linear_to_srgb(1.0f * ta[c]);
		alb[0] = a3[0]; alb[1] = a3[1]; alb[2] = a3[2]; alb[3] = 0.0f;
		if (F.apply_rgbplus) alb[3] = F.apply_L2 ? std::sqrt(std::max(0.0f, 3 - a3[0] * a3[0] - a3[1] * a3[1] - a3[2] * a3[2])) : 3 - std::fabs(a3[0]) - std::fabs(a3[1]) - std::fabs(a3[2]);
	}
	// light basis (camera frame): rows 0..2, column = light index.  testbed_nerf.cu:1537-1554
	float LD[3][3];
	const float slant = (float)(54.74f * M_PI / 180.0f);
	const float tilt[3] = {(float)(0.0f * M_PI / 180.0f), (float)(120.0f * M_PI / 180.0f), (float)(240.0f * M_PI / 180.0f)};
	for (int k = 0; k < 3; ++k) { LD[0][k] = -std::sin(slant) * std::cos(tilt[k]); LD[1][k] = -std::sin(slant) * std::sin(tilt[k]); LD[2][k] = -std::cos(slant); }
	if (F.apply_supernormal) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) LD[a][b] = a == b ? 1.f : 0.f;
	uint32_t li = F.light_mode >= 0 ? (uint32_t)F.light_mode % 3u : (F.light_mode == -2 ? ray_idx % 3u : hashed_light(ray_idx, step));
	if (F.light_opti) {   // Rodrigues alignment to the GT normal: testbed_nerf.cu:1563-1581
		float k[3] = {-nv[1], nv[0], 0.f};
		float kn = std::sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
		for (int c = 0; c < 3; ++c) k[c] /= kn;
		float ct = nv[2], st = std::sqrt(1 - ct * ct);
		float K[3][3] = {{0, -k[2], k[1]}, {k[2], 0, -k[0]}, {-k[1], k[0], 0}};
		float R[3][3];
		for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[a][b] = ct * (a == b ? 1.f : 0.f) + st * K[a][b] + (1 - ct) * k[a] * k[b];
		float L2[3][3];
		for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { float s = 0; for (int c = 0; c < 3; ++c) s += -R[a][c] * LD[c][b]; L2[a][b] = s; }
		std::memcpy(LD, L2, sizeof(LD));
	}
	float lc[3] = {LD[0][li], LD[1][li], LD[2][li]};
	const float* X = v.xform;
	for (int k = 0; k < 3; ++k) T.light[k] = X[0 + k] * lc[0] + X[3 + k] * lc[1] + X[6 + k] * lc[2];
	float sh = nv[0] * lc[0] + nv[1] * lc[1] + nv[2] * lc[2];
	if (F.apply_relu) sh = reluf(sh);
	for (int c = 0; c < 4; ++c) T.rgbt[c] = alb[c] * sh;
	T.mask_certainty = (float)(ta[3] > 0.99);
	T.mask_gt = (float)(tn[3] > 0.99);
}

static inline void albedo4(const float* out, const Flags& F, float a[4]) {
	if (F.no_albedo) { a[0] = a[1] = a[2] = 1.0f; a[3] = 0.0f; return; }
	for (int c = 0; c < 3; ++c) a[c] = logistic(out[c]);
	a[3] = 0.0f;
	if (F.apply_rgbplus) a[3] = F.apply_L2 ? std::sqrt(std::max(0.0f, 3 - a[0] * a[0] - a[1] * a[1] - a[2] * a[2])) : 3 - std::fabs(a[0]) - std::fabs(a[1]) - std::fabs(a[2]);
}

struct AlphaTerms { float inv_s, sdf, n[3], true_cos, iter_cos, next_sdf, p_div_c, alpha; };

// NeuS logistic alpha from the 16-wide network output: testbed_nerf.cu:1652-1677
static inline void neus_alpha(const float* out, const float dir[3], float dt, float car, AlphaTerms& A) {
	A.inv_s = std::exp(hmul(10.0f, out[7]));
	A.sdf = out[3];
	for (int d = 0; d < 3; ++d) A.n[d] = out[4 + d];
	A.true_cos = dir[0] * A.n[0] + dir[1] * A.n[1] + dir[2] * A.n[2];
	A.iter_cos = (float)-((double)reluf((float)(-A.true_cos * 0.5 + 0.5)) * (1.0 - car) + (double)reluf(-A.true_cos) * car);
	A.next_sdf = (float)(A.sdf + A.iter_cos * dt * 0.5);
	float prev_sdf = (float)(A.sdf - A.iter_cos * dt * 0.5);
	float next_cdf = logistic(A.next_sdf * A.inv_s), prev_cdf = logistic(prev_sdf * A.inv_s);
	A.p_div_c = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
	A.alpha = std::min(std::max(A.p_div_c, 0.0f), 1.0f);
}

// Transmittance scan: how many leading samples survive T >= 1e-4 (first loop of the loss kernel, :1608-1697)
static inline uint32_t ray_compacted_count(const float* out /*16*n*/, uint32_t n, const float* first_out, float dt, float car) {
	float dirw[3] = {first_out[8] * 2.0f - 1.0f, first_out[9] * 2.0f - 1.0f, first_out[10] * 2.0f - 1.0f};
	float dn = std::sqrt(dirw[0] * dirw[0] + dirw[1] * dirw[1] + dirw[2] * dirw[2]);
	float dir[3] = {dirw[0] / dn, dirw[1] / dn, dirw[2] / dn};
	float T = 1.f; uint32_t k = 0;
	for (; k < n; ++k) {
		if (T < 1e-4f) break;
		AlphaTerms A; neus_alpha(out + (size_t)k * 16, dir, dt, car, A);
		T *= (1.f - A.alpha);
	}
	return k;
}

struct RayLoss { float loss, ek_loss, mask_loss; };

// Composite + loss + gradient w.r.t. the 16 outputs for one ray.  out[0..n) are the ray's samples that survive
// the transmittance cut (first sweep); gradients are emitted for the first n_emit <= n of them (the reference
// truncates the last rays when the compacted batch is full, :1722-1728).
// dout receives n_emit*16 binary16-rounded values (already × loss_scale/n_rays).  testbed_nerf.cu:1608-2097
static inline void ray_loss(const float* out, uint32_t n, uint32_t n_emit, const RayTarget& Tg, const Flags& F, float dt, uint32_t n_rays, float loss_scale_total, float* dout, RayLoss& RL) {
	float dirw[3] = {out[8] * 2.0f - 1.0f, out[9] * 2.0f - 1.0f, out[10] * 2.0f - 1.0f};
	float dn = std::sqrt(dirw[0] * dirw[0] + dirw[1] * dirw[1] + dirw[2] * dirw[2]);
	float dir[3] = {dirw[0] / dn, dirw[1] / dn, dirw[2] / dn};
	const float car = F.cos_anneal_ratio;
	float rgb_ray[4] = {0, 0, 0, 0}, weight_sum = 0.f, T = 1.f;
	for (uint32_t k = 0; k < n; ++k) {
		const float* o = out + (size_t)k * 16;
		AlphaTerms A; neus_alpha(o, dir, dt, car, A);
		float alb[4]; albedo4(o, F, alb);
		float w = A.alpha * T;
		float sh = A.n[0] * Tg.light[0] + A.n[1] * Tg.light[1] + A.n[2] * Tg.light[2];
		if (F.apply_relu) sh = reluf(sh);
		for (int c = 0; c < 4; ++c) rgb_ray[c] += w * alb[c] * sh;
		weight_sum += w;
		T *= (1.f - A.alpha);
	}
	// loss + dL/drgb
	float grad[4], loss = 0.f;
	for (int c = 0; c < 4; ++c) {
		float d = rgb_ray[c] - Tg.rgbt[c];
		if (F.apply_L2) { loss += d * d; grad[c] = 2 * d; }
		else { loss += std::fabs(d); grad[c] = std::copysign(1.0f, d); }
	}
	if (F.apply_rgbplus) { loss /= 2; for (int c = 0; c < 4; ++c) grad[c] /= 2; }
	loss *= Tg.mask_certainty; for (int c = 0; c < 4; ++c) grad[c] *= Tg.mask_certainty;
	float gws;
	if (weight_sum >= 1.0 - 1e-4) { weight_sum = (float)(1.0 - 1e-4); gws = 0.0f; }
	else if (weight_sum <= 1e-4) { weight_sum = (float)1e-4; gws = 0.0f; }
	else {
		float sg = (float)(1.0f / (1.0f + std::exp(-weight_sum)));
		gws = F.apply_bce ? ((1 - Tg.mask_gt) / (1 - weight_sum) - Tg.mask_gt / weight_sum) * F.mask_loss_weight : (sg - Tg.mask_gt) * F.mask_loss_weight;
	}
	RL.loss = loss / (float)n_rays;
	{
		float sg = (float)(1.0f / (1.0f + std::exp(-weight_sum)));
		RL.mask_loss = F.apply_bce ? -(Tg.mask_gt * std::log(weight_sum) + (1 - Tg.mask_gt) * std::log(1 - weight_sum))
		                           : -(Tg.mask_gt * std::log(sg) + (1 - Tg.mask_gt) * std::log(1 - sg));
	}
	float ek_acc = 0.f;
	const float loss_scale = loss_scale_total / (float)n_rays;
	float rgb_ray2[4] = {0, 0, 0, 0}, weight_sum2 = 0.f; T = 1.f;
	for (uint32_t k = 0; k < n_emit; ++k) {
		const float* o = out + (size_t)k * 16;
		AlphaTerms A; neus_alpha(o, dir, dt, car, A);
		float alb[4]; albedo4(o, F, alb);
		const float alpha = A.alpha;
		const float w = alpha * T;
		float sh = A.n[0] * Tg.light[0] + A.n[1] * Tg.light[1] + A.n[2] * Tg.light[2];
		if (F.apply_relu) sh = reluf(sh);
		for (int c = 0; c < 4; ++c) rgb_ray2[c] += w * alb[c] * sh;
		weight_sum2 += w;
		T *= (1.f - alpha);
		float suffix[4]; for (int c = 0; c < 4; ++c) suffix[c] = rgb_ray[c] - rgb_ray2[c];
		float ag = alb[0] * grad[0] + alb[1] * grad[1] + alb[2] * grad[2] + alb[3] * grad[3];
		float dloss_dn[3]; for (int d = 0; d < 3; ++d) dloss_dn[d] = w * Tg.light[d] * ag;
		float jac3[3] = {0, 0, 0};
		if (F.apply_rgbplus) {
			if (F.apply_L2) for (int c = 0; c < 3; ++c) jac3[c] = (float)(-2 * alb[c] / (alb[3] + 1e-5));
			else for (int c = 0; c < 3; ++c) jac3[c] = -std::copysign(1.0f, alb[c]);
		}
		float drgb[3]; for (int c = 0; c < 3; ++c) drgb[c] = w * sh * (grad[c] + jac3[c] * grad[3]);
		float* L = dout + (size_t)k * 16;
		for (int c = 0; c < 16; ++c) L[c] = 0.f;
		const float opti_rgb = F.no_albedo ? 0.0f : 1.0f;
		for (int c = 0; c < 3; ++c) { float s = logistic(o[c]); L[c] = hq(opti_rgb * loss_scale * (drgb[c] * (s * (1 - s)))); }
		const float sum_weight_suffix = weight_sum - weight_sum2;
		float dot = 0.f; for (int c = 0; c < 4; ++c) dot += grad[c] * (T * alb[c] * sh - suffix[c]);
		float dloss_dalpha = (float)((dot + gws * (T - sum_weight_suffix)) / (1.0f - alpha + 1e-5));
		float da_dE = 0.f, dE_dsdf = 0.f, dE_dinvs = 0.f, da_dP = 0.f, dP_dinvs = 0.f, dP_dcos = 0.f, dE_dcos = 0.f;
		if (!(A.p_div_c <= 0.0f || A.p_div_c >= 1.0f)) {
			float P = std::exp(A.inv_s * A.iter_cos * dt);
			float E = std::exp(-A.next_sdf * A.inv_s);
			dE_dsdf = -A.inv_s * E;
			dE_dinvs = -A.next_sdf * E;
			float a = 1 + E, b = 1 + P * E;
			float c = (float)(1e-5 + 1 / (1 + P * E));
			float delta = a * (b * b) * (c * c);
			da_dE = -(P / delta - 1 / (a * a * c));
			da_dP = -E / delta;
			dP_dinvs = P * A.iter_cos * dt;
			dP_dcos = P * A.inv_s * dt;
			dE_dcos = (float)(-A.inv_s * E * dt * 0.5);
		}
		float dloss_dinvs = dloss_dalpha * (da_dE * dE_dinvs + da_dP * dP_dinvs);
		float dloss_dvar = dloss_dinvs * A.inv_s * 10;
		float dcos = A.true_cos >= 0 ? 0.0f : 1.0f;
		float gnorm = (float)std::sqrt((double)(A.n[0] * A.n[0] + A.n[1] * A.n[1] + A.n[2] * A.n[2]) + 1e-6);
		float gninv = 1 - 1 / gnorm;
		float dloss_dnn = dloss_dalpha * (da_dE * dE_dcos + dP_dcos * da_dP) * dcos;
		float dloss_dsdf = dloss_dalpha * da_dE * dE_dsdf;
		L[3] = hq(loss_scale * dloss_dsdf);
		ek_acc += (gnorm - 1.0f) * (gnorm - 1.0f);
		for (int d = 0; d < 3; ++d) L[4 + d] = hq(F.ek_loss_weight * 2 * loss_scale_total * gninv * A.n[d]);
		L[7] = hq(loss_scale * dloss_dvar);
		for (int d = 0; d < 3; ++d) L[8 + d] = hq(loss_scale * (dloss_dn[d] + dloss_dnn * dir[d]));
	}
	RL.ek_loss = n_emit ? ek_acc / ((float)n_emit * (float)n_rays) : 0.f;
}

} // namespace orc
