// rnb_march.cu — ray generation and occupancy-grid marching, warp-per-ray.
//
// Replaces generate_training_samples_nerf_with_global_movement (reference src/testbed_nerf.cu:1216-1387), which runs one
// thread per ray and marches twice.  Here one warp owns one ray: the 32 lanes test 32 consecutive lattice points
// t_k = t_{k-1} + dt at once (the lattice is fixed per ray because dt is constant, testbed_nerf.cu:153-155,3214), the
// DDA skip rule (:301-323) is resolved with ballots, and the accepted t values are written once.  Sample slots are then
// handed out by an ordered scan (deterministic) instead of atomicAdd arrival order (:1352,1359).
#include "rnb_common.cuh"
#include <algorithm>

namespace rnb {

__device__ __forceinline__ int mip_from_pos(float x, float y, float z) {   // testbed_nerf.cu:569-574
	float mx = fmaxf(fabsf(x - 0.5f), fmaxf(fabsf(y - 0.5f), fabsf(z - 0.5f)));
	int e; frexpf(mx, &e);
	return min((int)CASCADES - 1, max(0, e + 1));
}
// mip_from_dt (:576-583) with dt == DT: DT*2*128 < 1, so the position decides.
__device__ __forceinline__ uint32_t cell_index(float x, float y, float z, uint32_t mip) {   // cascaded_grid_idx_at :439-459
	float ms = scalbnf(1.0f, -(int)mip);
	int ix = (int)(((x - 0.5f) * ms + 0.5f) * (float)GRIDSIZE);
	int iy = (int)(((y - 0.5f) * ms + 0.5f) * (float)GRIDSIZE);
	int iz = (int)(((z - 0.5f) * ms + 0.5f) * (float)GRIDSIZE);
	ix = max(0, min((int)GRIDSIZE - 1, ix)); iy = max(0, min((int)GRIDSIZE - 1, iy)); iz = max(0, min((int)GRIDSIZE - 1, iz));
	return morton3D((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
}

struct RaySetup { bool valid; float ox, oy, oz, ux, uy, uz, dx, dy, dz, startt; };

// Everything before the march in the reference kernel (:1254-1330).  Executed redundantly by all lanes of the warp.
__device__ __forceinline__ RaySetup ray_setup(uint32_t i, uint32_t n_rays, uint32_t n_rays_total, Pcg32 rng, const ViewDev* __restrict__ views, uint32_t n_views) {
	RaySetup r; r.valid = false;
	const uint32_t img = image_idx(i, n_rays, n_rays_total, n_views);
	const ViewDev& v = views[img];
	rng.advance((int64_t)i * RNG_PER_RAY);
	float2 xy = pixel_pos(rng, v.w, v.h);
	float4 px = read_rgba(v.normal_px, v.w, v.h, xy.x, xy.y);
	if (px.x <= 0.0f && (double)rng.next_float() >= 0.9) return r;   // background pixels: 10 % kept (the draw happens only for them)
	(void)rng.next_float();                                          // motion-blur time
	const float dcx = (xy.x - v.cx) * (float)v.w / v.fx, dcy = (xy.y - v.cy) * (float)v.h / v.fy;
	const float* X = v.xform;
	r.ux = fmaf(X[6], 1.0f, fmaf(X[3], dcy, X[0] * dcx));
	r.uy = fmaf(X[7], 1.0f, fmaf(X[4], dcy, X[1] * dcx));
	r.uz = fmaf(X[8], 1.0f, fmaf(X[5], dcy, X[2] * dcx));
	r.ox = X[9]; r.oy = X[10]; r.oz = X[11];
	const float nrm = sqrtf(fmaf(r.uz, r.uz, fmaf(r.uy, r.uy, r.ux * r.ux)));
	r.dx = r.ux / nrm; r.dy = r.uy / nrm; r.dz = r.uz / nrm;
	// BoundingBox::ray_intersect on the unit cube, bounding_box.cuh:163-214
	const float BIG = 3.402823466e+38f;
	float tmin = (0.f - r.ox) / r.dx, tmax = (1.f - r.ox) / r.dx;
	if (tmin > tmax) { float s = tmin; tmin = tmax; tmax = s; }
	float tymin = (0.f - r.oy) / r.dy, tymax = (1.f - r.oy) / r.dy;
	if (tymin > tymax) { float s = tymin; tymin = tymax; tymax = s; }
	bool miss = (tmin > tymax || tymin > tmax);
	if (!miss) {
		if (tymin > tmin) tmin = tymin;
		if (tymax < tmax) tmax = tymax;
		float tzmin = (0.f - r.oz) / r.dz, tzmax = (1.f - r.oz) / r.dz;
		if (tzmin > tzmax) { float s = tzmin; tzmin = tzmax; tzmax = s; }
		miss = (tmin > tzmax || tzmin > tmax);
		if (!miss) { if (tzmin > tmin) tmin = tzmin; }
	}
	if (miss) tmin = BIG;
	tmin = fmaxf(tmin, 0.0f);
	r.startt = fmaf(DT, rng.next_float(), tmin);
	r.valid = true;
	return r;
}

// grid: one warp per local ray.  ts: [n_local_rays][MAX_STEPS] accepted t values.
__global__ void __launch_bounds__(256) k_march(uint32_t n_rays, uint32_t world, uint32_t rank, uint32_t n_rays_total, Pcg32 rng,
                                               const ViewDev* __restrict__ views, uint32_t n_views, const uint8_t* __restrict__ bitfield,
                                               uint32_t* __restrict__ ray_n, float* __restrict__ ray_geom /*9 per global ray: o, d_un, dir*/, float* __restrict__ ts) {
	// grid-stride over rays (one warp per ray at a time): the launch of the NEXT step's march, which shares the SMs with the network kernels of the
	// current step, is sized to one CTA per SM so that it cannot keep their CTAs from becoming resident (launch_march: `max_ctas`)
	const uint32_t lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp * world + rank < n_rays; warp += n_warps) {
	const uint32_t i = warp * world + rank;
	const RaySetup r = ray_setup(i, n_rays, n_rays_total, rng, views, n_views);
	if (!r.valid) { if (lane == 0) ray_n[i] = 0; continue; }
	if (lane < 9) {
		const float g[9] = {r.ox, r.oy, r.oz, r.ux, r.uy, r.uz, r.dx, r.dy, r.dz};
		ray_geom[(size_t)i * 9 + lane] = g[lane];
	}
	const float idx_ = 1.0f / r.dx, idy_ = 1.0f / r.dy, idz_ = 1.0f / r.dz;
	float* tout = ts + (size_t)warp * MAX_STEPS;
	float t0 = r.startt, tstar = -3.402823466e+38f;
	uint32_t nsamp = 0;
	bool done = false;
	for (int chunk = 0; chunk < 80 && !done; ++chunk) {
		// Lattice t_{k+1} = fl(t_k + dt) (the reference adds dt step by step, testbed_nerf.cu:311-323,1337-1347).  Inside one binade
		// every t is a multiple of the same ulp, so the rounded sum advances the BIT PATTERN by a constant D = rn(dt / ulp) (a tie
		// would need ulp = 2^-32, excluded below): lane j takes bits(t0) + j D.  A chunk that leaves the binade (t crossing 0.5, 1, 2)
		// or starts below 2^-7 falls back to the sequential adds.
		float tl, tc;
		{
			const uint32_t b0 = __float_as_uint(t0), ebits = b0 & 0x7F800000u;
			const uint32_t D = __float_as_uint(__uint_as_float(ebits) + DT) - ebits;
			const uint32_t last = b0 + 32u * D;
			if (ebits >= 0x3C000000u && ebits < 0x7F000000u && (last & 0x7F800000u) == ebits && (int)b0 > 0) {
				tl = __uint_as_float(b0 + lane * D); tc = __uint_as_float(last);
			} else {
				tl = t0; tc = t0;
				#pragma unroll
				for (int j = 0; j < 32; ++j) { if (j == (int)lane) tl = tc; tc += DT; }
			}
		}
		t0 = tc;
		const float px = fmaf(tl, r.dx, r.ox), py = fmaf(tl, r.dy, r.oy), pz = fmaf(tl, r.dz, r.oz);
		const bool inside = px >= 0.f && px <= 1.f && py >= 0.f && py <= 1.f && pz >= 0.f && pz <= 1.f;
		bool occ = false; float target = tl;
		if (inside) {
			const uint32_t mip = (uint32_t)mip_from_pos(px, py, pz);
			const uint32_t idx = cell_index(px, py, pz, mip);
			occ = __ldg(&bitfield[idx / 8 + (GRID_CELLS / 8) * mip]) & (1u << (idx % 8));
			if (!occ) {   // distance_to_next_voxel / advance_to_next_voxel, :301-323
				const float res = (float)(GRIDSIZE >> mip);
				const float qx = res * px, qy = res * py, qz = res * pz;
				const float tx = (floorf(qx + 0.5f + 0.5f * copysignf(1.0f, r.dx)) - qx) * idx_;
				const float ty = (floorf(qy + 0.5f + 0.5f * copysignf(1.0f, r.dy)) - qy) * idy_;
				const float tz = (floorf(qz + 0.5f + 0.5f * copysignf(1.0f, r.dz)) - qz) * idz_;
				target = tl + fmaxf(fminf(fminf(tx, ty), tz) / res, 0.0f);
			}
		}
		const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ), in_mask = __ballot_sync(0xffffffffu, inside);
		int cur = 0;
		while (cur < 32) {
			const uint32_t ge = __ballot_sync(0xffffffffu, tl >= tstar) & (0xffffffffu << cur);
			if (!ge) break;
			const int f = __ffs(ge) - 1;
			if (!((in_mask >> f) & 1u)) { done = true; break; }
			const uint32_t rm = occ_mask >> f;
			int run = (rm == 0xffffffffu) ? 32 : (__ffs(~rm) - 1);
			run = min(run, 32 - f);
			run = min(run, (int)(MAX_STEPS - nsamp));
			if ((int)lane >= f && (int)lane < f + run) tout[nsamp + (lane - f)] = tl;
			nsamp += run;
			if (nsamp >= MAX_STEPS) { done = true; break; }
			cur = f + run;
			if (cur >= 32) break;
			if (!((in_mask >> cur) & 1u)) { done = true; break; }
			if ((occ_mask >> cur) & 1u) continue;      // run was cut by the 32-f limit only
			tstar = __shfl_sync(0xffffffffu, target, cur);
			cur += 1;
		}
	}
	if (lane == 0) ray_n[i] = nsamp;
	}
}

// Ordered hand-out of sample slots: base = exclusive prefix of numsteps over rays (what the reference's atomicAdd counter
// yields for in-order arrival), the max_samples guard (:1353-1355) and the kept-ray compaction (:1359-1364).  One CTA.
// prev (optional, device): last step's sample count; the sample budget of this step is clamped to it like train_nerf_step does on the host
// (max_inference, testbed_nerf.cu:3891-3896): 0 -> the whole buffer, else next_multiple(min(prev, capacity), 128).  Keeping the clamp on the device
// means no step ever needs last step's counters on the host.
// Data parallel: only rays i = rank (mod world) were marched by this rank; the scan walks those (in ray order), not the global ray array.
// out[m] = inclusive prefix over this rank's ray positions of val: val[ray] (slot_of_pos == nullptr) or val[slot_of_pos[m]] (0 for dropped rays)
__global__ void __launch_bounds__(1024) k_prefix_positions(uint32_t n_rays, uint32_t world, uint32_t rank, uint32_t L, const uint32_t* __restrict__ val,
                                                           const uint32_t* __restrict__ slot_of_pos, uint32_t* __restrict__ out) {
	__shared__ uint32_t s_a[32];
	__shared__ uint32_t carry;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (tid == 0) carry = 0;
	__syncthreads();
	for (uint32_t c0 = 0; c0 < L; c0 += 1024) {
		const uint32_t m = c0 + tid, i = m * world + rank;
		uint32_t n = 0;
		if (m < L && i < n_rays) {
			if (slot_of_pos) { const uint32_t k = slot_of_pos[m]; n = k != 0xFFFFFFFFu ? val[k] : 0u; }
			else n = val[i];
		}
		uint32_t a = n;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, a, o); if ((int)lane >= o) a += v; }
		if (lane == 31) s_a[wid] = a;
		__syncthreads();
		if (wid == 0) { uint32_t v = s_a[lane]; for (int o = 1; o < 32; o <<= 1) { uint32_t w = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += w; } s_a[lane] = v; }
		__syncthreads();
		const uint32_t incl = a + (wid ? s_a[wid - 1] : 0) + carry;
		if (m < L) out[m] = incl;
		__syncthreads();
		if (tid == 1023) carry = incl;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(1024) k_scan_rays(uint32_t n_rays, uint32_t max_samples, const uint32_t* __restrict__ prev, const uint32_t* __restrict__ ray_n,
                                                    uint32_t* __restrict__ ray_indices, uint32_t* __restrict__ numsteps /*2 per kept ray*/, uint32_t* __restrict__ counters /*[0]=kept,[1]=samples,[6]=samples to forward,[9]=samples of all ranks*/,
                                                    uint32_t world, uint32_t rank, const uint32_t* __restrict__ xg, uint32_t L, uint32_t* __restrict__ slot_of_pos) {
	__shared__ uint32_t s_a[32], s_b[32];
	__shared__ uint32_t carry_n, carry_k;
	const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	if (prev) { const uint32_t p = *prev; if (p) max_samples = min(max_samples, (min(p, max_samples) + 127u) / 128u * 128u); }
	if (tid == 0) { carry_n = 0; carry_k = 0; }
	__syncthreads();
	const uint32_t n_local = (n_rays + world - 1 - rank) / world;
	for (uint32_t c0 = 0; c0 < n_local; c0 += 1024) {
		const uint32_t li = c0 + tid, i = li * world + rank;
		const uint32_t n = (li < n_local && i < n_rays) ? ray_n[i] : 0;
		uint32_t a = n;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, a, o); if ((int)lane >= o) a += v; }
		if (lane == 31) s_a[wid] = a;
		__syncthreads();
		if (wid == 0) { uint32_t v = s_a[lane]; for (int o = 1; o < 32; o <<= 1) { uint32_t w = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += w; } s_a[lane] = v; }
		__syncthreads();
		const uint32_t incl = a + (wid ? s_a[wid - 1] : 0) + carry_n;
		const uint32_t base = incl - n;
		// the guard of the reference's slot counter (:1346-1350) looks at the ray's place in the whole batch
		const uint32_t gbase = base + ((xg && n > 0) ? foreign_prefix(xg, L, world, rank, li) : 0u);
		const uint32_t kept = (n > 0 && gbase + n <= max_samples) ? 1u : 0u;
		uint32_t b = kept;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, b, o); if ((int)lane >= o) b += v; }
		if (lane == 31) s_b[wid] = b;
		__syncthreads();
		if (wid == 0) { uint32_t v = s_b[lane]; for (int o = 1; o < 32; o <<= 1) { uint32_t w = __shfl_up_sync(0xffffffffu, v, o); if ((int)lane >= o) v += w; } s_b[lane] = v; }
		__syncthreads();
		const uint32_t kincl = b + (wid ? s_b[wid - 1] : 0) + carry_k;
		if (kept) { const uint32_t k = kincl - 1; ray_indices[k] = i; numsteps[2 * k] = n; numsteps[2 * k + 1] = base; }
		if (slot_of_pos && li < L) slot_of_pos[li] = kept ? kincl - 1 : 0xFFFFFFFFu;
		__syncthreads();
		if (tid == 1023) { carry_n = incl; carry_k = kincl; }
		__syncthreads();
	}
	if (tid == 0) { counters[0] = carry_k; counters[1] = carry_n; counters[6] = min(carry_n, max_samples); counters[9] = xg ? global_total(xg, L, world) : carry_n; }
}

// One warp per kept ray: pos4[base + j] = { o + t*dir , ray slot }.
// Also writes the ray's warped direction (dir + 1) / 2 (warp_direction, testbed_nerf.cu:413-415) when ray_dirw is given (one launch less per step).
__global__ void __launch_bounds__(256) k_emit(const uint32_t* __restrict__ counters, uint32_t world, const uint32_t* __restrict__ ray_indices, const uint32_t* __restrict__ numsteps,
                                              const float* __restrict__ ray_geom, const float* __restrict__ ts, float4* __restrict__ pos4, float* __restrict__ ray_dirw) {
	const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (k >= counters[0]) return;
	const uint32_t i = ray_indices[k], n = numsteps[2 * k], base = numsteps[2 * k + 1];
	const float* g = ray_geom + (size_t)i * 9;
	const float ox = g[0], oy = g[1], oz = g[2], dx = g[6], dy = g[7], dz = g[8];
	if (ray_dirw && lane < 3) ray_dirw[3 * k + lane] = (g[6 + lane] + 1.0f) * 0.5f;
	const float* tin = ts + (size_t)(i / world) * MAX_STEPS;
	for (uint32_t j = lane; j < n; j += 32) {
		const float t = tin[j];
		pos4[base + j] = make_float4(fmaf(t, dx, ox), fmaf(t, dy, oy), fmaf(t, dz, oz), __uint_as_float(k));
	}
}

void launch_march(cudaStream_t st, uint32_t n_rays, uint32_t world, uint32_t rank, uint32_t n_rays_total, Pcg32 rng, const ViewDev* views, uint32_t n_views,
                  const uint8_t* bitfield, uint32_t* ray_n, float* ray_geom, float* ts, uint32_t max_ctas) {
	const uint32_t n_local = (n_rays + world - 1 - rank) / world;
	if (!n_local) return;
	uint32_t grid = (n_local * 32 + 255) / 256;
	if (max_ctas) grid = std::min(grid, max_ctas);
	k_march<<<grid, 256, 0, st>>>(n_rays, world, rank, n_rays_total, rng, views, n_views, bitfield, ray_n, ray_geom, ts);
}
void launch_scan_rays(cudaStream_t st, uint32_t n_rays, uint32_t max_samples, const uint32_t* prev, const uint32_t* ray_n, uint32_t* ray_indices, uint32_t* numsteps, uint32_t* counters, uint32_t world, uint32_t rank,
                      const uint32_t* xg, uint32_t* slot_of_pos) {
	if (!world) world = 1;
	k_scan_rays<<<1, 1024, 0, st>>>(n_rays, max_samples, prev, ray_n, ray_indices, numsteps, counters, world, rank, xg, (n_rays + world - 1) / world, slot_of_pos);
}
void launch_prefix_positions(cudaStream_t st, uint32_t n_rays, uint32_t world, uint32_t rank, const uint32_t* val, const uint32_t* slot_of_pos, uint32_t* out) {
	const uint32_t L = (n_rays + world - 1) / world;
	if (L) k_prefix_positions<<<1, 1024, 0, st>>>(n_rays, world, rank, L, val, slot_of_pos, out);
}
void launch_emit(cudaStream_t st, uint32_t n_rays_upper, const uint32_t* counters, uint32_t world, const uint32_t* ray_indices, const uint32_t* numsteps, const float* ray_geom, const float* ts, float4* pos4, float* ray_dirw) {
	if (!n_rays_upper) return;
	k_emit<<<(n_rays_upper * 32 + 255) / 256, 256, 0, st>>>(counters, world, ray_indices, numsteps, ray_geom, ts, pos4, ray_dirw);
}

} // namespace rnb
