// rnb_optim.cu — optimizer (Adam + EMA, fused) and occupancy-grid maintenance kernels.
//
// Optimizer: Ema(ExponentialDecay(Adam)) of the reference (tcnn optimizers/adam.h:51-202, ema.h:64-78,116-152,
// exponential_decay.h:61-72) as ONE pass over the parameters: the reference launches adam_step and
// ema_step_half_precision separately.  Gradients arrive as fp32 accumulators (× loss scale); they are rounded to
// binary16 first because the reference's gradient buffer is binary16 (trainer.h:78-84) and its "gradient == 0 -> skip"
// rule for hash-grid entries (adam.h:111-115) depends on that underflow.
//
// Occupancy grid: update_density_grid_nerf / update_density_grid_mean_and_bitfield (src/testbed_nerf.cu:3424-3517) and
// kernels :585-614,616-635,655-685,693-740.
#include "rnb_common.cuh"
#include <algorithm>

namespace rnb {

struct AdamParams {
	float base_lr, beta1, beta2, eps, l2, loss_scale, ema_decay, ema_debias_old, ema_debias_new;
	uint32_t n_params, n_matrix, rgb_begin, rgb_end; int only_sdf; float log2_beta1, log2_beta2;
	// data-parallel optimizer shard: this rank updates parameters [shard_begin, shard_end) (multiples of 4) from gsrc[i - shard_begin]
	// (the reduce-scattered gradient sum; NULL: the gradient buffer itself) and only clears the gradient buffer elsewhere
	uint32_t shard_begin, shard_end; const float* gsrc;
	// binary16 gradient exchange (data parallel behind the C ABI): the all-reduced / reduce-scattered gradient sum as binary16 (gsrc16[i - shard_begin]);
	// the fp32 accumulators have already been cleared by k_pack_grads, so this pass does not touch them
	const __half* gsrc16;
	// this launch covers parameters [first, last) only (chunked exchange: Adam on chunk k runs while chunk k + 1 is still being all-reduced); multiples of 4
	uint32_t first, last;
};

// One thread owns 4 consecutive parameters: every array is moved with one 128-bit (fp32 / u32) or 64-bit (binary16) access.
// Fast path for the common case of an untouched hash-grid quad whose EMA copy has already converged to the weight:
// nothing to write (the EMA update (ema*d*old + w*(1-d))*new is a fixed point at ema == w because d*old + 1 - d == 1/new).
struct AdamLane { float w, m1, m2; uint32_t step; };

// bias-corrected learning rate of a parameter that has taken `cs` steps (adam.h:189-190).  Computed per parameter: caching it across the four parameters of a
// quad (they almost always share `cs`) was measured SLOWER (0.091 -> 0.104 ms: the compare-and-branch chain costs the instruction-level parallelism of four
// independent evaluations; sessions r2j / r2k)
__device__ __forceinline__ float adam_lr(const AdamParams& A, uint32_t cs) {
	return A.base_lr * (sqrtf(1 - exp2f((float)cs * A.log2_beta2)) / (1 - exp2f((float)cs * A.log2_beta1)));
}

__device__ __forceinline__ bool adam_one(const AdamParams& A, uint32_t i, float g32, AdamLane& S, __half& wh) {
	float gradient = hq(g32) / A.loss_scale;
	const bool is_mat = i < A.n_matrix;
	bool update = is_mat || gradient != 0.f;
	if (A.only_sdf && i >= A.rgb_begin && i < A.rgb_end) update = false;
	if (!update) return false;
	const float w = S.w;
	if (is_mat) gradient += A.l2 * w;
	const float fm = S.m1 = A.beta1 * S.m1 + (1 - A.beta1) * gradient;
	const float sm = S.m2 = A.beta2 * S.m2 + (1 - A.beta2) * (gradient * gradient);
	const uint32_t cs = ++S.step;
	const float lr = adam_lr(A, cs);
	const float eff = fminf(fmaxf(lr / (sqrtf(sm) + A.eps), 0.f), 3.402823466e+38f);
	const float nw = w - eff * fm;
	S.w = nw;
	wh = __float2half_rn(nw);
	return true;
}

// Two quads per thread, in three phases (gradient / binary16 copies of both quads -> optimizer state of the quads that need it -> arithmetic and stores): the loads
// of a phase are all in flight together, so a thread keeps 2 x (32 + 64) bytes outstanding instead of waiting twice for 32 + 64 (ncu r02: 362 MB of DRAM traffic
// per launch at 63 % of the copy bandwidth with one quad per thread).
constexpr int ADAM_U = 2;

__global__ void __launch_bounds__(256) k_adam_ema(AdamParams A, float* __restrict__ master, __half* __restrict__ params, __half* __restrict__ ema,
                                                  float* __restrict__ grads, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
	const uint32_t base = A.first + (blockIdx.x * blockDim.x + threadIdx.x) * (4 * ADAM_U);
	const uint32_t lim = min(A.n_params, A.last);
	float4 g[ADAM_U]; uint2 pw[ADAM_U], pe[ADAM_U]; int mode[ADAM_U];      // 0 nothing, 1 full quad in this rank's range, 2 partial tail quad, 3 another rank's quad
	// ---- phase 1
	#pragma unroll
	for (int u = 0; u < ADAM_U; ++u) {
		const uint32_t i0 = base + 4 * u;
		mode[u] = 0;
		if (i0 >= lim) continue;
		if (i0 < A.shard_begin || i0 >= A.shard_end) { mode[u] = 3; if (!A.gsrc16 && i0 + 4 <= A.n_params) g[u] = *reinterpret_cast<const float4*>(grads + i0); continue; }
		if (i0 + 4 > A.n_params) { mode[u] = 2; continue; }
		mode[u] = 1;
		if (A.gsrc16) {
			const uint2 h = *reinterpret_cast<const uint2*>(A.gsrc16 + (i0 - A.shard_begin));
			const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
			g[u] = make_float4(lo.x, lo.y, hi.x, hi.y);
		} else g[u] = *reinterpret_cast<const float4*>(grads + i0);
		pw[u] = *reinterpret_cast<const uint2*>(params + i0); pe[u] = *reinterpret_cast<const uint2*>(ema + i0);
	}
	// ---- phase 2
	float4 w4[ADAM_U], a4[ADAM_U], b4[ADAM_U]; uint4 s4[ADAM_U]; bool upd[ADAM_U];
	#pragma unroll
	for (int u = 0; u < ADAM_U; ++u) {
		const uint32_t i0 = base + 4 * u;
		upd[u] = false;
		if (mode[u] == 3) {      // another rank's parameters: drop this rank's partial gradient, the weights arrive with the all-gather
			if (!A.gsrc16) {
				if (i0 + 4 <= A.n_params) { if (g[u].x != 0.f || g[u].y != 0.f || g[u].z != 0.f || g[u].w != 0.f) *reinterpret_cast<float4*>(grads + i0) = make_float4(0.f, 0.f, 0.f, 0.f); }
				else for (uint32_t i = i0; i < A.n_params; ++i) grads[i] = 0.f;
			}
			mode[u] = 0;
			continue;
		}
		if (mode[u] != 1) continue;
		if (A.gsrc) {                                   // the reduced gradient lives in the caller's reduce-scatter output
			if (g[u].x != 0.f || g[u].y != 0.f || g[u].z != 0.f || g[u].w != 0.f) *reinterpret_cast<float4*>(grads + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
			g[u] = *reinterpret_cast<const float4*>(A.gsrc + (i0 - A.shard_begin));
		}
		const bool anyg = g[u].x != 0.f || g[u].y != 0.f || g[u].z != 0.f || g[u].w != 0.f;
		const bool mat = i0 < A.n_matrix;
		if (!anyg && !mat && pw[u].x == pe[u].x && pw[u].y == pe[u].y) { mode[u] = 0; continue; }      // untouched hash quad whose EMA has converged: nothing to write
		if (anyg || mat) {
			upd[u] = true;
			if (anyg && !A.gsrc && !A.gsrc16) *reinterpret_cast<float4*>(grads + i0) = make_float4(0.f, 0.f, 0.f, 0.f);     // consumed: ready for the next step's atomics
			w4[u] = *reinterpret_cast<const float4*>(master + i0); a4[u] = *reinterpret_cast<const float4*>(m1 + i0); b4[u] = *reinterpret_cast<const float4*>(m2 + i0);
			s4[u] = *reinterpret_cast<const uint4*>(steps + i0);
		}
	}
	// ---- phase 3
	#pragma unroll
	for (int u = 0; u < ADAM_U; ++u) {
		const uint32_t i0 = base + 4 * u;
		if (mode[u] == 1) {
			__half* wh = reinterpret_cast<__half*>(&pw[u]); __half* eh = reinterpret_cast<__half*>(&pe[u]);
			if (upd[u]) {
				AdamLane S[4] = {{w4[u].x, a4[u].x, b4[u].x, s4[u].x}, {w4[u].y, a4[u].y, b4[u].y, s4[u].y}, {w4[u].z, a4[u].z, b4[u].z, s4[u].z}, {w4[u].w, a4[u].w, b4[u].w, s4[u].w}};
				const float gg[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
				bool any = false;
				#pragma unroll
				for (int q = 0; q < 4; ++q) any |= adam_one(A, i0 + q, gg[q], S[q], wh[q]);
				if (any) {
					*reinterpret_cast<float4*>(master + i0) = make_float4(S[0].w, S[1].w, S[2].w, S[3].w);
					*reinterpret_cast<float4*>(m1 + i0) = make_float4(S[0].m1, S[1].m1, S[2].m1, S[3].m1);
					*reinterpret_cast<float4*>(m2 + i0) = make_float4(S[0].m2, S[1].m2, S[2].m2, S[3].m2);
					*reinterpret_cast<uint4*>(steps + i0) = make_uint4(S[0].step, S[1].step, S[2].step, S[3].step);
				}
			}
			#pragma unroll
			for (int q = 0; q < 4; ++q)
				eh[q] = __float2half_rn((__half2float(eh[q]) * A.ema_decay * A.ema_debias_old + __half2float(wh[q]) * (1 - A.ema_decay)) * A.ema_debias_new);
			*reinterpret_cast<uint2*>(params + i0) = pw[u]; *reinterpret_cast<uint2*>(ema + i0) = pe[u];
		} else if (mode[u] == 2) {      // the last, partial quad (parameter count not a multiple of 4)
			for (uint32_t i = i0; i < A.n_params; ++i) {
				const float g32 = A.gsrc16 ? __half2float(A.gsrc16[i - A.shard_begin]) : (A.gsrc ? A.gsrc[i - A.shard_begin] : grads[i]);
				if (!A.gsrc16) grads[i] = 0.f;
				__half wh = params[i];
				AdamLane S{master[i], m1[i], m2[i], steps[i]};
				if (adam_one(A, i, g32, S, wh)) { master[i] = S.w; m1[i] = S.m1; m2[i] = S.m2; steps[i] = S.step; }
				params[i] = wh;
				ema[i] = __float2half_rn((__half2float(ema[i]) * A.ema_decay * A.ema_debias_old + __half2float(wh) * (1 - A.ema_decay)) * A.ema_debias_new);
			}
		}
	}
}

// fp32 gradient accumulators -> binary16 (the reference's gradient format, trainer.h:78-84) for the data-parallel exchange; clears the accumulators.
// n is a multiple of 8 (the arrays are padded to 512 elements).  One thread moves 8 parameters: two 128-bit loads, one 128-bit store.
__global__ void __launch_bounds__(256) k_pack_grads(uint32_t n, float* __restrict__ grads, __half* __restrict__ out) {
	const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
	if (i0 >= n) return;
	const float4 a = *reinterpret_cast<const float4*>(grads + i0), b = *reinterpret_cast<const float4*>(grads + i0 + 4);
	const bool nz = a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f || b.x != 0.f || b.y != 0.f || b.z != 0.f || b.w != 0.f;
	uint4 h = make_uint4(0u, 0u, 0u, 0u);
	if (nz) {
		__half2 t;
		t = __floats2half2_rn(a.x, a.y); h.x = *reinterpret_cast<uint32_t*>(&t); t = __floats2half2_rn(a.z, a.w); h.y = *reinterpret_cast<uint32_t*>(&t);
		t = __floats2half2_rn(b.x, b.y); h.z = *reinterpret_cast<uint32_t*>(&t); t = __floats2half2_rn(b.z, b.w); h.w = *reinterpret_cast<uint32_t*>(&t);
		*reinterpret_cast<float4*>(grads + i0) = make_float4(0.f, 0.f, 0.f, 0.f); *reinterpret_cast<float4*>(grads + i0 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	*reinterpret_cast<uint4*>(out + i0) = h;
}

__global__ void k_cast_params(uint32_t n, const float* __restrict__ master, __half* __restrict__ params) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) params[i] = __float2half_rn(master[i]);
}
__global__ void k_widen_params(uint32_t n, const __half* __restrict__ params, float* __restrict__ master) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) master[i] = __half2float(params[i]);
}

// hash-grid initialisation in the reference's generate_random_kernel layout (tcnn random.h:67-93): thread i owns elements
// i + n_threads*j, j < 4, drawing 4 consecutive floats of the stream advanced by 4 i.
__global__ void k_init_grid(Pcg32 rng, uint64_t n, uint64_t n_threads, float* __restrict__ out) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_threads) return;
	rng.advance((int64_t)(i * 4));
	for (uint64_t j = 0; j < 4; ++j) {
		const uint64_t idx = i + n_threads * j;
		if (idx >= n) return;
		out[idx] = rng.next_float() * (1e-4f - (-1e-4f)) + (-1e-4f);
	}
}

// generate_grid_samples_nerf_nonuniform (:585-614), one cascade
__global__ void k_grid_samples(uint32_t n_elements, Pcg32 rng, uint32_t step, const float* __restrict__ grid_in, float4* __restrict__ pos_out, uint32_t* __restrict__ idx_out, float thresh) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	rng.advance((int64_t)i * 4);
	const uint32_t level = (uint32_t)(rng.next_float() * 1.0f) % 1u;
	uint32_t idx = 0;
	for (uint32_t j = 0; j < 10; ++j) {
		idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % GRID_CELLS;
		idx += level * GRID_CELLS;
		if (grid_in[idx] > thresh) break;
	}
	const uint32_t pi = idx % GRID_CELLS;
	const uint32_t x = morton3D_invert(pi >> 0), y = morton3D_invert(pi >> 1), z = morton3D_invert(pi >> 2);
	const float rx = rng.next_float(), ry = rng.next_float(), rz = rng.next_float();
	const float sc = scalbnf(1.0f, (int)level);
	pos_out[i] = make_float4((((float)x + rx) / (float)GRIDSIZE - 0.5f) * sc + 0.5f, (((float)y + ry) / (float)GRIDSIZE - 0.5f) * sc + 0.5f, (((float)z + rz) / (float)GRIDSIZE - 0.5f) * sc + 0.5f, 0.f);
	idx_out[i] = idx;
}
// splat_grid_samples_nerf_max_nearest_neighbor (:616-635)
__global__ void k_grid_splat(uint32_t n, const uint32_t* __restrict__ idx, const float* __restrict__ dens, float* __restrict__ tmp) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicMax(reinterpret_cast<uint32_t*>(&tmp[idx[i]]), __float_as_uint(dens[i]));
}
// ema_grid_samples_nerf (:655-685) fused with the mean reduction (:3509) — sum of max(v,0)/n in double
__global__ void __launch_bounds__(256) k_grid_ema_mean(uint32_t n, float decay, float* __restrict__ grid, float* __restrict__ tmp, double* __restrict__ mean_acc) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	double v = 0.0;
	if (i < n) {
		const float pv = grid[i];
		const float nv = pv < 0.f ? pv : fmaxf(pv * decay, tmp[i]);
		grid[i] = nv;
		tmp[i] = 0.f;                                 // ready for the next refresh
		v = (double)(fmaxf(nv, 0.f) / (float)n);
	}
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__shared__ double s[8];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0) { double t = 0; for (int q = 0; q < 8; ++q) t += s[q]; atomicAdd(mean_acc, t); }
}
// grid_to_bitfield (:693-717) for mip 0 + zero fill of the other mips
__global__ void k_grid_bitfield(const float* __restrict__ grid, uint8_t* __restrict__ bitfield, const double* __restrict__ mean_acc, float* __restrict__ mean_out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= GRID_CELLS / 8 * CASCADES) return;
	if (i >= GRID_CELLS / 8) { bitfield[i] = 0; return; }
	const float mean = (float)*mean_acc;
	if (i == 0) *mean_out = mean;
	const float thresh = fminf(MIN_OPTICAL_THICKNESS, mean);
	uint8_t bits = 0;
	#pragma unroll
	for (int j = 0; j < 8; ++j) bits |= grid[i * 8 + j] > thresh ? (uint8_t)(1 << j) : 0;
	bitfield[i] = bits;
}
// bitfield_max_pool (:719-740).  Distinct output bytes per thread: plain store.
__global__ void k_bitfield_pool(const uint8_t* __restrict__ prev, uint8_t* __restrict__ next) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= GRID_CELLS / 64) return;
	uint8_t bits = 0;
	#pragma unroll
	for (int j = 0; j < 8; ++j) bits |= prev[i * 8 + j] > 0 ? (uint8_t)(1 << j) : 0;
	const uint32_t x = morton3D_invert(i >> 0) + GRIDSIZE / 8, y = morton3D_invert(i >> 1) + GRIDSIZE / 8, z = morton3D_invert(i >> 2) + GRIDSIZE / 8;
	next[morton3D(x, y, z)] |= bits;
}

void launch_adam_ema(cudaStream_t st, const AdamParams& A, float* master, __half* params, __half* ema, float* grads, float* m1, float* m2, uint32_t* steps) {
	const uint32_t end = std::min(A.last, A.n_params);
	if (end <= A.first) return;
	k_adam_ema<<<((end - A.first + 4 * ADAM_U - 1) / (4 * ADAM_U) + 255) / 256, 256, 0, st>>>(A, master, params, ema, grads, m1, m2, steps);
}
void launch_pack_grads(cudaStream_t st, uint32_t n_padded, float* grads, __half* out) { k_pack_grads<<<(n_padded / 8 + 255) / 256, 256, 0, st>>>(n_padded, grads, out); }
void launch_cast_params(cudaStream_t st, uint32_t n, const float* master, __half* params) { k_cast_params<<<(n + 255) / 256, 256, 0, st>>>(n, master, params); }
void launch_widen_params(cudaStream_t st, uint32_t n, const __half* params, float* master) { k_widen_params<<<(n + 255) / 256, 256, 0, st>>>(n, params, master); }
void launch_init_grid(cudaStream_t st, Pcg32 rng, uint64_t n, float* out) {
	const uint64_t n_thr = (n + 3) / 4, n_threads = ((n_thr + 127) / 128) * 128;
	k_init_grid<<<(uint32_t)(n_threads / 128), 128, 0, st>>>(rng, n, n_threads, out);
}
void launch_grid_samples(cudaStream_t st, uint32_t n, Pcg32 rng, uint32_t step, const float* grid, float4* pos, uint32_t* idx, float thresh) {
	if (n) k_grid_samples<<<(n + 127) / 128, 128, 0, st>>>(n, rng, step, grid, pos, idx, thresh);
}
void launch_grid_finish(cudaStream_t st, uint32_t n_samples, const uint32_t* idx, const float* dens, float decay, float* grid, float* tmp, double* mean_acc, float* mean_out, uint8_t* bitfield) {
	cudaMemsetAsync(mean_acc, 0, sizeof(double), st);
	if (n_samples) k_grid_splat<<<(n_samples + 255) / 256, 256, 0, st>>>(n_samples, idx, dens, tmp);
	k_grid_ema_mean<<<GRID_CELLS / 256, 256, 0, st>>>(GRID_CELLS, decay, grid, tmp, mean_acc);
	k_grid_bitfield<<<(GRID_CELLS / 8 * CASCADES + 255) / 256, 256, 0, st>>>(grid, bitfield, mean_acc, mean_out);
	for (uint32_t lvl = 1; lvl < CASCADES; ++lvl)
		k_bitfield_pool<<<(GRID_CELLS / 64 + 255) / 256, 256, 0, st>>>(bitfield + (size_t)GRID_CELLS * (lvl - 1) / 8, bitfield + (size_t)GRID_CELLS * lvl / 8);
}

} // namespace rnb
