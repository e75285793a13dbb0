// umma_probe.cu — standalone check of the tcgen05 operand layouts used by rnb_network_tc.cu (run on the B200 box).
//
// The network kernels keep every activation tile in ONE "panel" layout: a [rows x cols] fp16 tile is stored as cols/8
// panels, panel j holding the 16-byte chunk (cols 8j..8j+7) of every row, rows 16 bytes apart.  This file verifies on
// hardware that the same bytes are consumed correctly by tcgen05.mma (kind::f16, no swizzle) as
//   T1  A K-major  x B K-major   (forward layer:     D[s][n]   = sum_k X[s][k]  W[n][k])
//   T2  A K-major  x B MN-major  (transposed layer:  D[s][i]   = sum_k G[s][k]  W[k][i], same W buffer as T1)
//   T3  A MN-major x B MN-major  (weight gradient:   D[o][i]   = sum_s dY[s][o] X[s][i], K = samples)
// and prints which TMEM lanes hold the rows of an M=64 accumulator.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
	d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
	return d;                        // layout_type 0 = no swizzle, base_offset 0
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
	return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}

struct TestCfg { int M, N, K; int a_mn, b_mn; uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep; };

// smem: A region 32 KB at 0, B region 32 KB at 32768 (raw bytes uploaded by the host in final layout)
__global__ void __launch_bounds__(128) k_probe(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, TestCfg c, float* __restrict__ out /*128 lanes x 64 cols*/) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint32_t tmem_base;
	__shared__ __align__(8) uint64_t bar;
	for (int i = threadIdx.x; i < 32768 / 16; i += 128) {
		reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
		reinterpret_cast<uint4*>(smem + 32768)[i] = reinterpret_cast<const uint4*>(b_img)[i];
	}
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tb = tmem_base;
	// zero the accumulator columns first so that unwritten lanes are visible as zeros
	{
		const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
		for (int col = 0; col < 64; col += 8) {
			asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + col), "r"(0u) : "memory");
		}
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (threadIdx.x == 0) {
		const uint32_t idesc = make_idesc(c.M, c.N, c.a_mn, c.b_mn);
		const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
		for (int k = 0; k < c.K / 16; ++k)
			umma(tb, make_desc(sa + k * c.a_kstep, c.a_lbo, c.a_sbo), make_desc(sb + k * c.b_kstep, c.b_lbo, c.b_sbo), idesc, k > 0 ? 1u : 0u);
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	{
		const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
		for (int col = 0; col < 64; col += 8) {
			uint32_t v[8];
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + col) : "memory");
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			for (int j = 0; j < 8; ++j) out[threadIdx.x * 64 + col + j] = __uint_as_float(v[j]);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64u) : "memory");
}


// ---- T5: A operand in TMEM (activations never leave the tensor-core datapath) ------------------------------------------
// Each thread (TMEM lane = row) stores its K binary16 values packed two per 32-bit column with tcgen05.st, then
// tcgen05.mma reads A from [a_tmem] (K = 16 per instruction = 8 columns) and B from shared memory.
__device__ __forceinline__ void umma_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__global__ void __launch_bounds__(128) k_probe_ta(const uint32_t* __restrict__ a_rows /*[128][32] packed half2*/, const uint8_t* __restrict__ b_img, TestCfg c, float* __restrict__ out) {
	extern __shared__ __align__(1024) uint8_t smem[];
	__shared__ uint32_t tmem_base;
	__shared__ __align__(8) uint64_t bar;
	for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4*>(smem + 32768)[i] = reinterpret_cast<const uint4*>(b_img)[i];
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tb = tmem_base;
	const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
	// D at columns 0..63 (zeroed), A at columns 64..95 (K = 64 halfs = 32 packed columns)
	for (int col = 0; col < 64; col += 8) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + col), "r"(0u) : "memory");
	{
		const uint32_t* ar = a_rows + threadIdx.x * 32;
		for (int col = 0; col < 32; col += 8)
			asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr + 64 + col), "r"(ar[col]), "r"(ar[col + 1]), "r"(ar[col + 2]), "r"(ar[col + 3]),
			             "r"(ar[col + 4]), "r"(ar[col + 5]), "r"(ar[col + 6]), "r"(ar[col + 7]) : "memory");
	}
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (threadIdx.x == 0) {
		const uint32_t idesc = make_idesc(c.M, c.N, 0, c.b_mn);
		const uint32_t sb = smem_u32(smem + 32768);
		for (int k = 0; k < c.K / 16; ++k) umma_ta(tb, tb + 64 + k * 8, make_desc(sb + k * c.b_kstep, c.b_lbo, c.b_sbo), idesc, k > 0 ? 1u : 0u);
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	uint32_t done = 0;
	while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int col = 0; col < 64; col += 8) {
		uint32_t v[8];
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + col) : "memory");
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int j = 0; j < 8; ++j) out[threadIdx.x * 64 + col + j] = __uint_as_float(v[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128u) : "memory");
}

static float h2f(__half h) { return __half2float(h); }

// panel layout: element (row, col) of a [rows x cols] tile -> byte offset
static size_t panel_off(int rows, int row, int col) { return (size_t)(col / 8) * rows * 16 + (size_t)row * 16 + (col % 8) * 2; }

int main() {
	uint8_t *da, *db; float* dout;
	cudaMalloc(&da, 32768); cudaMalloc(&db, 32768); cudaMalloc(&dout, 128 * 64 * 4);
	cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
	std::vector<uint8_t> A(32768), B(32768);
	std::vector<float> out(128 * 64);
	srand(1);
	auto rnd = []() { return (float)((rand() % 17) - 8) / 8.0f; };
	int fails = 0;
	auto run = [&](const char* name, TestCfg c, auto ref /*(m,n)->float*/, int rows_valid, bool lane_map, bool expect_ok = true) {
		cudaMemcpy(da, A.data(), 32768, cudaMemcpyHostToDevice); cudaMemcpy(db, B.data(), 32768, cudaMemcpyHostToDevice);
		cudaMemset(dout, 0, 128 * 64 * 4);
		k_probe<<<1, 128, 65536>>>(da, db, c, dout);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("%-28s CUDA ERROR %s\n", name, cudaGetErrorString(e)); fails++; return; }
		cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
		if (!lane_map) {
			double mx = 0;
			for (int m = 0; m < rows_valid; ++m) for (int n = 0; n < c.N; ++n) mx = fmax(mx, fabs(out[m * 64 + n] - ref(m, n)));
			printf("%-28s max abs err %.3g  %s\n", name, mx, mx < 1e-3 ? "OK" : (expect_ok ? "MISMATCH" : "mismatch (expected: negative control)"));
			if ((mx < 1e-3) != expect_ok) fails++;
		} else {
			// find for each logical row which lane holds it
			printf("%-28s lane of row r (M=64): ", name);
			int bad = 0;
			for (int m = 0; m < 64; ++m) {
				int found = -1;
				for (int lane = 0; lane < 128 && found < 0; ++lane) {
					bool ok = true;
					for (int n = 0; n < c.N && ok; ++n) ok = fabs(out[lane * 64 + n] - ref(m, n)) < 1e-3;
					if (ok) found = lane;
				}
				if (m % 8 == 0) printf("r%d->%d ", m, found);
				if (found < 0) bad++;
			}
			printf(" (%d rows not found)\n", bad);
		}
	};

	// ---- T1: A [128 x 32] K-major, B = W [64 x 32] K-major ----
	{
		std::vector<float> X(128 * 32), W(64 * 32);
		for (auto& v : X) v = rnd(); for (auto& v : W) v = rnd();
		std::fill(A.begin(), A.end(), 0); std::fill(B.begin(), B.end(), 0);
		for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) *(__half*)&A[panel_off(128, r, k)] = __float2half(X[r * 32 + k]);
		for (int n = 0; n < 64; ++n) for (int k = 0; k < 32; ++k) *(__half*)&B[panel_off(64, n, k)] = __float2half(W[n * 32 + k]);
		auto ref = [&](int m, int n) { float s = 0; for (int k = 0; k < 32; ++k) s += X[m * 32 + k] * W[n * 32 + k]; return s; };
		TestCfg c{128, 64, 32, 0, 0, /*a_lbo*/ 128 * 16, /*a_sbo*/ 128, /*a_kstep*/ 2 * 128 * 16, /*b_lbo*/ 64 * 16, /*b_sbo*/ 128, /*b_kstep*/ 2 * 64 * 16};
		run("T1 K-major x K-major", c, ref, 128, false);
		TestCfg cs = c; cs.a_lbo = 128; cs.a_sbo = 128 * 16; cs.b_lbo = 128; cs.b_sbo = 64 * 16;
		run("T1 (LBO/SBO swapped)", cs, ref, 128, false, false);
		// ---- T2: A = G [128 x 64] K-major, B = same W buffer read MN-major: D[s][i] = sum_h G[s][h] W[h][i], N = 32 ----
		std::vector<float> G(128 * 64);
		for (auto& v : G) v = rnd();
		std::fill(A.begin(), A.end(), 0);
		for (int r = 0; r < 128; ++r) for (int k = 0; k < 64; ++k) *(__half*)&A[panel_off(128, r, k)] = __float2half(G[r * 64 + k]);
		auto ref2 = [&](int m, int n) { float s = 0; for (int h = 0; h < 64; ++h) s += G[m * 64 + h] * W[h * 32 + n]; return s; };
		// B MN-major: MN groups (8 inputs) are one panel apart (64*16), K groups (8 hidden rows) are 128 B apart
		TestCfg c2{128, 32, 64, 0, 1, 128 * 16, 128, 2 * 128 * 16, /*b_lbo (K groups)*/ 128, /*b_sbo (MN groups)*/ 64 * 16, /*b_kstep: 16 hidden rows*/ 256};
		run("T2 K-major x MN-major", c2, ref2, 128, false);
		TestCfg c2s = c2; c2s.b_lbo = 64 * 16; c2s.b_sbo = 128;
		run("T2 (B LBO/SBO swapped)", c2s, ref2, 128, false, false);
	}
	// ---- T3: weight gradient, both MN-major, K = 128 samples ----
	{
		std::vector<float> dY(128 * 64), X(128 * 32);
		for (auto& v : dY) v = rnd(); for (auto& v : X) v = rnd();
		std::fill(A.begin(), A.end(), 0); std::fill(B.begin(), B.end(), 0);
		for (int s = 0; s < 128; ++s) for (int o = 0; o < 64; ++o) *(__half*)&A[panel_off(128, s, o)] = __float2half(dY[s * 64 + o]);   // 8 panels used, 8 zero
		for (int s = 0; s < 128; ++s) for (int i = 0; i < 32; ++i) *(__half*)&B[panel_off(128, s, i)] = __float2half(X[s * 32 + i]);
		auto ref3 = [&](int m, int n) { float a = 0; if (m >= 64) return 0.f; for (int s = 0; s < 128; ++s) a += dY[s * 64 + m] * X[s * 32 + n]; return a; };
		// MN-major: MN groups (8 features) one panel apart (128*16); K groups (8 samples) 128 B apart; 16 samples per instruction = 256 B
		TestCfg c3{128, 32, 128, 1, 1, /*a_lbo*/ 128, /*a_sbo*/ 128 * 16, /*a_kstep*/ 256, /*b_lbo*/ 128, /*b_sbo*/ 128 * 16, /*b_kstep*/ 256};
		run("T3 MN x MN (M=128, 64 used)", c3, ref3, 128, false);
		TestCfg c3s = c3; c3s.a_lbo = 128 * 16; c3s.a_sbo = 128; c3s.b_lbo = 128 * 16; c3s.b_sbo = 128;
		run("T3 (LBO/SBO swapped)", c3s, ref3, 128, false, false);
		TestCfg c4 = c3; c4.M = 64;
		run("T4 MN x MN (M=64)", c4, ref3, 64, true);
	}
	// ---- T5: A from TMEM (K = 64), B = W [64 x 32] panelised read MN-major (as T2) and K-major W' [32 x 64] ----
	{
		cudaFuncSetAttribute(k_probe_ta, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
		std::vector<float> G(128 * 64), W(64 * 32);
		for (auto& v : G) v = rnd(); for (auto& v : W) v = rnd();
		std::vector<uint32_t> arows(128 * 32);
		for (int r = 0; r < 128; ++r) for (int j = 0; j < 32; ++j) { __half2 h = __floats2half2_rn(G[r * 64 + 2 * j], G[r * 64 + 2 * j + 1]); arows[r * 32 + j] = *reinterpret_cast<uint32_t*>(&h); }
		uint32_t* darows; cudaMalloc(&darows, arows.size() * 4); cudaMemcpy(darows, arows.data(), arows.size() * 4, cudaMemcpyHostToDevice);
		std::fill(B.begin(), B.end(), 0);
		for (int n = 0; n < 64; ++n) for (int k = 0; k < 32; ++k) *(__half*)&B[panel_off(64, n, k)] = __float2half(W[n * 32 + k]);
		auto ref5 = [&](int m, int n) { float s = 0; for (int h = 0; h < 64; ++h) s += G[m * 64 + h] * W[h * 32 + n]; return s; };
		TestCfg c5{128, 32, 64, 0, 1, 0, 0, 0, /*b_lbo*/ 128, /*b_sbo*/ 64 * 16, /*b_kstep*/ 256};
		cudaMemcpy(db, B.data(), 32768, cudaMemcpyHostToDevice); cudaMemset(dout, 0, 128 * 64 * 4);
		k_probe_ta<<<1, 128, 65536>>>(darows, db, c5, dout);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("T5 A-in-TMEM                 CUDA ERROR %s\n", cudaGetErrorString(e)); fails++; }
		else {
			cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
			double mx = 0; for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) mx = fmax(mx, fabs(out[m * 64 + n] - ref5(m, n)));
			printf("T5 A-in-TMEM x MN-major      max abs err %.3g  %s\n", mx, mx < 1e-3 ? "OK" : "MISMATCH");
			if (!(mx < 1e-3)) { fails++; printf("   row0: got %g %g %g %g  want %g %g %g %g\n", out[0], out[1], out[2], out[3], ref5(0, 0), ref5(0, 1), ref5(0, 2), ref5(0, 3)); }
		}
	}
	printf(fails ? "UMMA PROBE: %d FAILED\n" : "UMMA PROBE: ALL OK\n", fails);
	return 0;
}
