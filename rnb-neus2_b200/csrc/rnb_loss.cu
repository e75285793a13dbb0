// rnb_loss.cu — transmittance compaction, NeuS compositing and the RNb losses with their gradients.
//
// Replaces compute_loss_kernel_train_nerf_with_global_movement (reference src/testbed_nerf.cu:1396-2097).  The reference
// runs three serial sweeps per ray inside one thread.  Here the work is split in two kernels around the second network
// pass: k_compact_count (first sweep's transmittance cut, needs only sdf/normal of pass A) and k_loss (compositing,
// losses, gradients w.r.t. the 16 network outputs).  Sample slots of the compacted batch come from an ordered scan.
#include "rnb_common.cuh"

namespace rnb {

struct AlphaTerms { float inv_s, sdf, nx, ny, nz, true_cos, iter_cos, next_sdf, p_div_c, alpha; };

// NeuS logistic alpha — testbed_nerf.cu:1652-1677.  var_h: binary16 variance parameter.
__device__ __forceinline__ AlphaTerms neus_alpha(float sdf, float nx, float ny, float nz, __half var_h, float dx, float dy, float dz, float car) {
	AlphaTerms A;
	A.inv_s = __expf(__half2float(__hmul(__float2half_rn(10.0f), var_h)));
	A.sdf = sdf; A.nx = nx; A.ny = ny; A.nz = nz;
	A.true_cos = dx * nx + dy * ny + dz * nz;
	A.iter_cos = -(fmaxf(-A.true_cos * 0.5f + 0.5f, 0.f) * (1.0f - car) + fmaxf(-A.true_cos, 0.f) * car);
	const float hstep = A.iter_cos * DT * 0.5f;
	A.next_sdf = sdf + hstep;
	const float prev_sdf = sdf - hstep;
	const float next_cdf = logisticf(A.next_sdf * A.inv_s), prev_cdf = logisticf(prev_sdf * A.inv_s);
	A.p_div_c = (prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f);
	A.alpha = fminf(fmaxf(A.p_div_c, 0.0f), 1.0f);
	return A;
}

// bent view direction: unwarp(binary16(warp(dir))) normalised — testbed_nerf.cu:1645-1650
__device__ __forceinline__ void bent_dir(const float* dirw, float& dx, float& dy, float& dz) {
	const float a = hq(dirw[0]) * 2.0f - 1.0f, b = hq(dirw[1]) * 2.0f - 1.0f, c = hq(dirw[2]) * 2.0f - 1.0f;
	const float n = sqrtf(a * a + b * b + c * c);
	dx = a / n; dy = b / n; dz = c / n;
}

// One warp per kept ray.  n_fwd[k] = number of leading samples with T >= 1e-4 (testbed_nerf.cu:1608-1611).
// The product T *= (1-alpha) is evaluated in sample order by every lane (prefix chain), so the cut is order-exact.
__global__ void __launch_bounds__(256) k_compact_count(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ numsteps, const __half* __restrict__ outA /*4 per sample*/,
                                                       const float* __restrict__ ray_dirw, const __half* __restrict__ P, uint32_t off_var, float car, uint32_t* __restrict__ n_fwd) {
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (k >= counters[0]) return;
	const uint32_t n = numsteps[2 * k], base = numsteps[2 * k + 1];
	float dx, dy, dz; bent_dir(ray_dirw + 3 * k, dx, dy, dz);
	const __half var_h = __ldg(P + off_var);
	float T = 1.0f; uint32_t count = n;
	// the chunk loop is a chain (load -> alpha -> ordered product): the next chunk's load is issued before this chunk's arithmetic
	const uint2* src = reinterpret_cast<const uint2*>(outA) + base;
	uint2 raw_next = lane < n ? __ldg(src + lane) : make_uint2(0u, 0u);
	for (uint32_t c0 = 0; c0 < n; c0 += 32) {
		const uint32_t j = c0 + lane;
		const uint2 raw = raw_next;
		if (j + 32 < n) raw_next = __ldg(src + j + 32);
		float om = 1.0f;
		if (j < n) {
			const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
			om = 1.0f - neus_alpha(a.x, a.y, b.x, b.y, var_h, dx, dy, dz, car).alpha;
		}
		// T before sample c0+q, q = 0..31 (uniform across lanes)
		int stop = -1;
		#pragma unroll
		for (int q = 0; q < 32; ++q) {
			const float omq = __shfl_sync(0xffffffffu, om, q);
			if (stop < 0) { if (c0 + q < n && T < 1e-4f) stop = q; else T *= omq; }
		}
		if (stop >= 0) { count = c0 + stop; break; }
	}
	if (lane == 0) n_fwd[k] = count;
}

// Ordered scan over kept rays: compacted base (untruncated prefix), emitted count (truncated at max_compacted, :1722-1728),
// totals.  counters: [2] = compacted total (untruncated), [3] = trained samples min(total, max), [4] = samples to forward.
// Data parallel with one sample order (xg != nullptr; foreign_prefix, rnb_common.cuh): max_compacted is the GLOBAL target and every ray is cut where the
// single-process batch would cut it; goff[k] = (index of the ray's first sample in the global batch) - cbase[k] for the roll-over multiplicities;
// counters[3] = this rank's samples inside the target, counters[8] = min(samples of all ranks, target), counters[9] = marched samples of all ranks.
__global__ void __launch_bounds__(1024) k_scan_compact(uint32_t* __restrict__ counters, uint32_t max_compacted, const uint32_t* __restrict__ n_fwd,
                                                       uint32_t* __restrict__ cbase, uint32_t* __restrict__ n_emit, float* __restrict__ stats,
                                                       const uint32_t* __restrict__ xg, uint32_t L, uint32_t world, uint32_t rank, const uint32_t* __restrict__ ray_indices, uint32_t* __restrict__ goff) {
	__shared__ uint32_t s_a[32];
	__shared__ uint32_t carry, fwd_end, trained;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const uint32_t K = counters[0];
	if (tid == 0) { carry = 0; fwd_end = 0; trained = 0; }
	__syncthreads();
	for (uint32_t c0 = 0; c0 < K; c0 += 1024) {
		const uint32_t k = c0 + tid;
		const uint32_t n = k < K ? n_fwd[k] : 0;
		uint32_t a = n;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, a, o); if ((int)lane >= o) a += v; }
		if (lane == 31) s_a[wid] = a;
		__syncthreads();
		if (wid == 0) { uint32_t v = s_a[lane]; for (int o = 1; o < 32; o <<= 1) { uint32_t w = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += w; } s_a[lane] = v; }
		__syncthreads();
		const uint32_t incl = a + (wid ? s_a[wid - 1] : 0) + carry;
		const uint32_t base = incl - n;
		if (k < K) {
			cbase[k] = base;
			uint32_t gbase = base;
			if (xg) { const uint32_t f = foreign_prefix(xg, L, world, rank, ray_indices[k] / world); gbase += f; goff[k] = f; }
			const uint32_t e = min(max_compacted - min(max_compacted, gbase), n);
			n_emit[k] = e;
			if (e > 0) { atomicMax(&fwd_end, incl); if (xg) atomicAdd(&trained, e); }
		}
		__syncthreads();
		if (tid == 1023) carry = incl;
		__syncthreads();
	}
	if (tid == 0) {
		const uint32_t total_all = xg ? global_total(xg, L, world) : carry, marched_all = xg ? counters[9] : counters[1];
		counters[2] = carry; counters[3] = xg ? trained : min(carry, max_compacted); counters[4] = fwd_end;
		counters[5] = (total_all == 0u || marched_all == 0u) ? 0u : marched_all;      // measured_batch_size_before_compaction for the next step's clamp (Counters::update_after_training :3540-3545)
		counters[8] = min(total_all, max_compacted);
		if (stats) { stats[3] = (float)carry; stats[4] = (float)counters[1]; stats[5] = (float)K; }   // float copies: summed across ranks with the losses
	}
}

// Gather compacted samples: cpos4[cbase + j] = pos4[base + j] for j < n_fwd, rays with n_emit > 0 only.
__global__ void __launch_bounds__(256) k_gather_compacted(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ numsteps, const uint32_t* __restrict__ n_fwd,
                                                          const uint32_t* __restrict__ cbase, const uint32_t* __restrict__ n_emit, const float4* __restrict__ pos4, float4* __restrict__ cpos4) {
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (k >= counters[0] || n_emit[k] == 0) return;
	const uint32_t n = n_fwd[k], base = numsteps[2 * k + 1], cb = cbase[k];
	for (uint32_t j = lane; j < n; j += 32) cpos4[cb + j] = pos4[base + j];
}

struct LossParams {
	rnb_flags F; uint32_t n_rays, n_rays_total, step; float loss_scale;
};

__device__ __forceinline__ void albedo4(const float* o, const rnb_flags& F, float a[4]) {
	if (F.no_albedo) { a[0] = a[1] = a[2] = 1.0f; a[3] = 0.0f; return; }
	for (int c = 0; c < 3; ++c) a[c] = 1.0f / (1.0f + expf(-o[c]));
	a[3] = 0.0f;
	if (F.apply_rgbplus) a[3] = F.apply_L2 ? sqrtf(fmaxf(0.0f, 3 - a[0] * a[0] - a[1] * a[1] - a[2] * a[2])) : 3 - fabsf(a[0]) - fabsf(a[1]) - fabsf(a[2]);
}

__device__ __forceinline__ void unpack_out16(const uint4& lo, const uint4& hi, float* o) {
	const uint32_t u[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
	#pragma unroll
	for (int i = 0; i < 8; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[i])); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
}

__device__ __forceinline__ float warp_scan_add(float v, int lane) {
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
	return v;
}
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v *= t; }
	return v;
}
__device__ __forceinline__ float warp_sum(float v) {
	#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// One WARP per kept ray (the reference walks each ray serially in one thread): lanes own consecutive samples, the
// transmittance product and the running colour / weight sums become warp scans, every lane writes its sample's gradient row.
__global__ void __launch_bounds__(256) k_loss(LossParams LP, const uint32_t* __restrict__ counters, Pcg32 rng, const ViewDev* __restrict__ views, uint32_t n_views,
                                              const uint32_t* __restrict__ ray_indices, const float* __restrict__ ray_dirw,
                                              const uint32_t* __restrict__ n_fwd, const uint32_t* __restrict__ cbase, const uint32_t* __restrict__ n_emit,
                                              const __half* __restrict__ out16, __half* __restrict__ dout16, float* __restrict__ loss_out /*3 per kept ray*/) {
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (k >= counters[0]) return;
	const uint32_t ne = n_emit[k];
	if (ne == 0) { if (lane < 3) loss_out[3 * k + lane] = 0.f; return; }
	const uint32_t nf = n_fwd[k], cb = cbase[k];
	const rnb_flags& F = LP.F;
	const uint32_t ray_idx = ray_indices[k];
	// every chunk loop below is a chain (load -> alpha -> warp scans -> carries): the next chunk's rows are requested before this chunk's arithmetic,
	// the first chunk's before the per-ray target is fetched
	const uint4* rows = reinterpret_cast<const uint4*>(out16 + (size_t)cb * 16);
	uint4 lo_next = make_uint4(0u, 0u, 0u, 0u), hi_next = lo_next;
	if ((uint32_t)lane < nf) { lo_next = __ldg(rows + 2 * lane); hi_next = __ldg(rows + 2 * lane + 1); }
	// ---- per-ray target and light (testbed_nerf.cu:1485-1593) ----
	rng.advance((int64_t)ray_idx * RNG_PER_RAY);
	const uint32_t img = image_idx(ray_idx, LP.n_rays, LP.n_rays_total, n_views);
	const ViewDev& v = views[img];
	const float2 xy = pixel_pos(rng, v.w, v.h);
	const float4 tn = read_rgba(v.normal_px, v.w, v.h, xy.x, xy.y);
	float4 ta = v.albedo_px ? read_rgba(v.albedo_px, v.w, v.h, xy.x, xy.y) : make_float4(tn.w, tn.w, tn.w, tn.w);
	float nv[3] = {linear_to_srgb(tn.x) * 2.0f - 1.0f, -(linear_to_srgb(tn.y) * 2.0f - 1.0f), -(linear_to_srgb(tn.z) * 2.0f - 1.0f)};
	{ const float nn = sqrtf(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]); nv[0] /= nn; nv[1] /= nn; nv[2] /= nn; }
	float albt[4];
	if (F.no_albedo) { albt[0] = albt[1] = albt[2] = 1.0f; albt[3] = 0.0f; }
	else {
		albt[0] = linear_to_srgb(ta.x); albt[1] = linear_to_srgb(ta.y); albt[2] = linear_to_srgb(ta.z); albt[3] = 0.0f;
		if (F.apply_rgbplus) albt[3] = F.apply_L2 ? sqrtf(fmaxf(0.0f, 3 - albt[0] * albt[0] - albt[1] * albt[1] - albt[2] * albt[2])) : 3 - fabsf(albt[0]) - fabsf(albt[1]) - fabsf(albt[2]);
	}
	float LD[3][3];
	{
		const float slant = 54.74f * 3.14159265358979323846f / 180.0f;
		const float tilt[3] = {0.0f, 120.0f * 3.14159265358979323846f / 180.0f, 240.0f * 3.14159265358979323846f / 180.0f};
		for (int q = 0; q < 3; ++q) { LD[0][q] = -sinf(slant) * cosf(tilt[q]); LD[1][q] = -sinf(slant) * sinf(tilt[q]); LD[2][q] = -cosf(slant); }
		if (F.apply_supernormal) for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) LD[a][b] = a == b ? 1.f : 0.f;
	}
	const uint32_t li = F.light_mode >= 0 ? (uint32_t)F.light_mode % 3u : (F.light_mode == -2 ? ray_idx % 3u : hashed_light(ray_idx, LP.step));
	if (F.light_opti) {
		float kk[3] = {-nv[1], nv[0], 0.f};
		const float kn = sqrtf(kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2]);
		kk[0] /= kn; kk[1] /= kn; kk[2] /= kn;
		const float ct = nv[2], st = sqrtf(1 - ct * ct);
		const float Km[3][3] = {{0, -kk[2], kk[1]}, {kk[2], 0, -kk[0]}, {-kk[1], kk[0], 0}};
		float R[3][3], L2[3][3];
		for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) R[a][b] = ct * (a == b ? 1.f : 0.f) + st * Km[a][b] + (1 - ct) * kk[a] * kk[b];
		for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { float s = 0; for (int c = 0; c < 3; ++c) s += -R[a][c] * LD[c][b]; L2[a][b] = s; }
		for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) LD[a][b] = L2[a][b];
	}
	const float lc[3] = {LD[0][li], LD[1][li], LD[2][li]};
	const float* X = v.xform;
	const float light[3] = {X[0] * lc[0] + X[3] * lc[1] + X[6] * lc[2], X[1] * lc[0] + X[4] * lc[1] + X[7] * lc[2], X[2] * lc[0] + X[5] * lc[1] + X[8] * lc[2]};
	float sht = nv[0] * lc[0] + nv[1] * lc[1] + nv[2] * lc[2];
	if (F.apply_relu) sht = fmaxf(sht, 0.f);
	const float rgbt[4] = {albt[0] * sht, albt[1] * sht, albt[2] * sht, albt[3] * sht};
	const float mask_certainty = ta.w > 0.99f ? 1.f : 0.f, mask_gt = tn.w > 0.99f ? 1.f : 0.f;

	float dx, dy, dz; bent_dir(ray_dirw + 3 * k, dx, dy, dz);
	const float car = F.cos_anneal_ratio;
	// ---- sweep 1: composite (testbed_nerf.cu:1608-1697) ----
	float rgb_ray[4] = {0, 0, 0, 0}, weight_sum = 0.f, Tc = 1.f;
	for (uint32_t c0 = 0; c0 < nf; c0 += 32) {
		const uint32_t j = c0 + lane;
		const uint4 lo = lo_next, hi = hi_next;
		if (j + 32 < nf) { lo_next = __ldg(rows + 2 * (j + 32)); hi_next = __ldg(rows + 2 * (j + 32) + 1); }
		float om = 1.f, w = 0.f, alb[4] = {0, 0, 0, 0}, sh = 0.f, alpha = 0.f;
		if (j < nf) {
			float o[16]; unpack_out16(lo, hi, o);
			const AlphaTerms A = neus_alpha(o[3], o[4], o[5], o[6], __float2half_rn(o[7]), dx, dy, dz, car);
			albedo4(o, F, alb);
			sh = A.nx * light[0] + A.ny * light[1] + A.nz * light[2];
			if (F.apply_relu) sh = fmaxf(sh, 0.f);
			alpha = A.alpha; om = 1.f - alpha;
		}
		const float incl = warp_scan_mul(om, lane);
		float excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = 1.f;
		w = alpha * (Tc * excl);
		#pragma unroll
		for (int c = 0; c < 4; ++c) rgb_ray[c] += w * alb[c] * sh;
		weight_sum += w;
		Tc *= __shfl_sync(0xffffffffu, incl, 31);
	}
	lo_next = hi_next = make_uint4(0u, 0u, 0u, 0u);      // first chunk of sweep 2, requested before the reductions and the loss terms
	if ((uint32_t)lane < ne) { lo_next = __ldg(rows + 2 * lane); hi_next = __ldg(rows + 2 * lane + 1); }
	#pragma unroll
	for (int c = 0; c < 4; ++c) rgb_ray[c] = warp_sum(rgb_ray[c]);
	weight_sum = warp_sum(weight_sum);
	// ---- losses (testbed_nerf.cu:1737-1798) ----
	float grad[4], loss = 0.f;
	for (int c = 0; c < 4; ++c) {
		const float d = rgb_ray[c] - rgbt[c];
		if (F.apply_L2) { loss += d * d; grad[c] = 2 * d; } else { loss += fabsf(d); grad[c] = copysignf(1.0f, d); }
	}
	if (F.apply_rgbplus) { loss /= 2; for (int c = 0; c < 4; ++c) grad[c] /= 2; }
	loss *= mask_certainty; for (int c = 0; c < 4; ++c) grad[c] *= mask_certainty;
	float gws;
	if (weight_sum >= 1.0f - 1e-4f) { weight_sum = 1.0f - 1e-4f; gws = 0.0f; }
	else if (weight_sum <= 1e-4f) { weight_sum = 1e-4f; gws = 0.0f; }
	else {
		const float sg = 1.0f / (1.0f + expf(-weight_sum));
		gws = F.apply_bce ? ((1 - mask_gt) / (1 - weight_sum) - mask_gt / weight_sum) * F.mask_loss_weight : (sg - mask_gt) * F.mask_loss_weight;
	}
	if (lane == 0) {
		const float sg = 1.0f / (1.0f + expf(-weight_sum));
		loss_out[3 * k] = loss / (float)LP.n_rays;
		loss_out[3 * k + 2] = F.apply_bce ? -(mask_gt * logf(weight_sum) + (1 - mask_gt) * logf(1 - weight_sum)) : -(mask_gt * logf(sg) + (1 - mask_gt) * logf(1 - sg));
	}
	// ---- sweep 2: gradients (testbed_nerf.cu:1836-2091) ----
	const float loss_scale = LP.loss_scale / (float)LP.n_rays;
	float c_rgb[4] = {0, 0, 0, 0}, c_w = 0.f, ek_acc = 0.f; Tc = 1.f;
	for (uint32_t c0 = 0; c0 < ne; c0 += 32) {
		const uint32_t j = c0 + lane;
		const bool valid = j < ne;
		const uint4 lo = lo_next, hi = hi_next;          // all-zero rows (binary16 0) for lanes past the end
		lo_next = hi_next = make_uint4(0u, 0u, 0u, 0u);
		if (j + 32 < ne) { lo_next = __ldg(rows + 2 * (j + 32)); hi_next = __ldg(rows + 2 * (j + 32) + 1); }
		float o[16];
		unpack_out16(lo, hi, o);
		const AlphaTerms A = neus_alpha(o[3], o[4], o[5], o[6], __float2half_rn(o[7]), dx, dy, dz, car);
		float alb[4]; albedo4(o, F, alb);
		const float alpha = valid ? A.alpha : 0.f;
		float sh = A.nx * light[0] + A.ny * light[1] + A.nz * light[2];
		if (F.apply_relu) sh = fmaxf(sh, 0.f);
		const float incl = warp_scan_mul(1.f - alpha, lane);
		float excl = __shfl_up_sync(0xffffffffu, incl, 1); if (lane == 0) excl = 1.f;
		const float w = alpha * (Tc * excl);
		const float T = Tc * incl;                               // transmittance after this sample
		float rgb_ray2[4];
		#pragma unroll
		for (int c = 0; c < 4; ++c) rgb_ray2[c] = c_rgb[c] + warp_scan_add(w * alb[c] * sh, lane);
		const float weight_sum2 = c_w + warp_scan_add(w, lane);
		#pragma unroll
		for (int c = 0; c < 4; ++c) c_rgb[c] = __shfl_sync(0xffffffffu, rgb_ray2[c], 31);
		c_w = __shfl_sync(0xffffffffu, weight_sum2, 31);
		Tc *= __shfl_sync(0xffffffffu, incl, 31);
		if (!valid) continue;
		const float ag = alb[0] * grad[0] + alb[1] * grad[1] + alb[2] * grad[2] + alb[3] * grad[3];
		float jac3[3] = {0, 0, 0};
		if (F.apply_rgbplus) {
			if (F.apply_L2) for (int c = 0; c < 3; ++c) jac3[c] = -2 * alb[c] / (alb[3] + 1e-5f);
			else for (int c = 0; c < 3; ++c) jac3[c] = -copysignf(1.0f, alb[c]);
		}
		float L[16];
		#pragma unroll
		for (int c = 0; c < 16; ++c) L[c] = 0.f;
		const float opti_rgb = F.no_albedo ? 0.0f : 1.0f;
		for (int c = 0; c < 3; ++c) {
			const float drgb = w * sh * (grad[c] + jac3[c] * grad[3]);
			const float s = 1.0f / (1.0f + expf(-o[c]));
			L[c] = opti_rgb * loss_scale * (drgb * (s * (1 - s)));
		}
		float dot = 0.f;
		for (int c = 0; c < 4; ++c) dot += grad[c] * (T * alb[c] * sh - (rgb_ray[c] - rgb_ray2[c]));
		const float dloss_dalpha = (dot + gws * (T - (weight_sum - weight_sum2))) / (1.0f - alpha + 1e-5f);
		float da_dE = 0.f, dE_dsdf = 0.f, dE_dinvs = 0.f, da_dP = 0.f, dP_dinvs = 0.f, dP_dcos = 0.f, dE_dcos = 0.f;
		if (!(A.p_div_c <= 0.0f || A.p_div_c >= 1.0f)) {
			const float Pp = __expf(A.inv_s * A.iter_cos * DT);
			const float E = __expf(-A.next_sdf * A.inv_s);
			dE_dsdf = -A.inv_s * E;
			dE_dinvs = -A.next_sdf * E;
			const float a = 1 + E, b = 1 + Pp * E, cc = 1e-5f + 1 / (1 + Pp * E);
			const float delta = a * (b * b) * (cc * cc);
			da_dE = -(Pp / delta - 1 / (a * a * cc));
			da_dP = -E / delta;
			dP_dinvs = Pp * A.iter_cos * DT;
			dP_dcos = Pp * A.inv_s * DT;
			dE_dcos = -A.inv_s * E * DT * 0.5f;
		}
		const float dloss_dinvs = dloss_dalpha * (da_dE * dE_dinvs + da_dP * dP_dinvs);
		const float dloss_dvar = dloss_dinvs * A.inv_s * 10;
		const float dcos = A.true_cos >= 0 ? 0.0f : 1.0f;
		const float gnorm = sqrtf(A.nx * A.nx + A.ny * A.ny + A.nz * A.nz + 1e-6f);
		const float gninv = 1 - 1 / gnorm;
		const float dloss_dnn = dloss_dalpha * (da_dE * dE_dcos + dP_dcos * da_dP) * dcos;
		L[3] = loss_scale * (dloss_dalpha * da_dE * dE_dsdf);
		ek_acc += (gnorm - 1.0f) * (gnorm - 1.0f);
		const float ekc = F.ek_loss_weight * 2 * LP.loss_scale * gninv;
		L[4] = ekc * A.nx; L[5] = ekc * A.ny; L[6] = ekc * A.nz;
		L[7] = loss_scale * dloss_dvar;
		const float wag = w * ag;
		L[8] = loss_scale * (wag * light[0] + dloss_dnn * dx);
		L[9] = loss_scale * (wag * light[1] + dloss_dnn * dy);
		L[10] = loss_scale * (wag * light[2] + dloss_dnn * dz);
		__align__(16) __half h[16];
		#pragma unroll
		for (int c = 0; c < 16; ++c) h[c] = __float2half_rn(L[c]);
		uint4* dst = reinterpret_cast<uint4*>(dout16 + (size_t)(cb + j) * 16);
		dst[0] = reinterpret_cast<uint4*>(h)[0]; dst[1] = reinterpret_cast<uint4*>(h)[1];
	}
	ek_acc = warp_sum(ek_acc);
	if (lane == 0) loss_out[3 * k + 1] = ek_acc / ((float)ne * (float)LP.n_rays);
}

// Sum per-ray losses into stats[0..2] (reduce_sum, testbed_nerf.cu:3547-3552); one CTA.
__global__ void __launch_bounds__(1024) k_reduce_losses(const uint32_t* __restrict__ counters, const float* __restrict__ loss_out, float* __restrict__ stats) {
	__shared__ float s[3][32];
	const uint32_t K = counters[0];
	float a = 0.f, b = 0.f, c = 0.f;
	for (uint32_t k = threadIdx.x; k < K; k += 1024) { a += loss_out[3 * k]; b += loss_out[3 * k + 1]; c += loss_out[3 * k + 2]; }
	for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
	if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = a; s[1][threadIdx.x >> 5] = b; s[2][threadIdx.x >> 5] = c; }
	__syncthreads();
	if (threadIdx.x < 32) {
		a = s[0][threadIdx.x]; b = s[1][threadIdx.x]; c = s[2][threadIdx.x];
		for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
		if (threadIdx.x == 0) { stats[0] = a; stats[1] = b; stats[2] = c; }
	}
}

// per kept ray: warped direction (dir+1)/2 — warp_direction, testbed_nerf.cu:413-415
__global__ void k_ray_dirw(const uint32_t* __restrict__ counters, const uint32_t* __restrict__ ray_indices, const float* __restrict__ ray_geom, float* __restrict__ ray_dirw) {
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= counters[0]) return;
	const float* g = ray_geom + (size_t)ray_indices[k] * 9;
	ray_dirw[3 * k] = (g[6] + 1.0f) * 0.5f; ray_dirw[3 * k + 1] = (g[7] + 1.0f) * 0.5f; ray_dirw[3 * k + 2] = (g[8] + 1.0f) * 0.5f;
}

void launch_ray_dirw(cudaStream_t st, uint32_t n_upper, const uint32_t* counters, const uint32_t* ray_indices, const float* ray_geom, float* ray_dirw) {
	if (n_upper) k_ray_dirw<<<(n_upper + 255) / 256, 256, 0, st>>>(counters, ray_indices, ray_geom, ray_dirw);
}
void launch_compact_count(cudaStream_t st, uint32_t n_upper, const uint32_t* counters, const uint32_t* numsteps, const __half* outA, const float* ray_dirw, const __half* P, uint32_t off_var, float car, uint32_t* n_fwd) {
	if (n_upper) k_compact_count<<<(n_upper * 32 + 255) / 256, 256, 0, st>>>(counters, numsteps, outA, ray_dirw, P, off_var, car, n_fwd);
}
void launch_scan_compact(cudaStream_t st, uint32_t* counters, uint32_t max_compacted, const uint32_t* n_fwd, uint32_t* cbase, uint32_t* n_emit, float* stats,
                         const uint32_t* xg, uint32_t L, uint32_t world, uint32_t rank, const uint32_t* ray_indices, uint32_t* goff) {
	k_scan_compact<<<1, 1024, 0, st>>>(counters, max_compacted, n_fwd, cbase, n_emit, stats, xg, L, world ? world : 1u, rank, ray_indices, goff);
}
void launch_gather_compacted(cudaStream_t st, uint32_t n_upper, const uint32_t* counters, const uint32_t* numsteps, const uint32_t* n_fwd, const uint32_t* cbase, const uint32_t* n_emit, const float4* pos4, float4* cpos4) {
	if (n_upper) k_gather_compacted<<<(n_upper * 32 + 255) / 256, 256, 0, st>>>(counters, numsteps, n_fwd, cbase, n_emit, pos4, cpos4);
}
void launch_loss(cudaStream_t st, uint32_t n_upper, const rnb_flags& F, uint32_t n_rays, uint32_t n_rays_total, uint32_t step, float loss_scale, const uint32_t* counters, Pcg32 rng,
                 const ViewDev* views, uint32_t n_views, const uint32_t* ray_indices, const float* ray_dirw, const uint32_t* n_fwd, const uint32_t* cbase, const uint32_t* n_emit,
                 const __half* out16, __half* dout16, float* loss_out, float* stats) {
	if (!n_upper) return;
	LossParams LP; LP.F = F; LP.n_rays = n_rays; LP.n_rays_total = n_rays_total; LP.step = step; LP.loss_scale = loss_scale;
	k_loss<<<(n_upper * 32 + 255) / 256, 256, 0, st>>>(LP, counters, rng, views, n_views, ray_indices, ray_dirw, n_fwd, cbase, n_emit, out16, dout16, loss_out);
	if (stats) k_reduce_losses<<<1, 1024, 0, st>>>(counters, loss_out, stats);
}

} // namespace rnb
