// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the RNb-NeuS2 per-step training hot path (reference:
// /root/reference/src/testbed_nerf.cu, include/neural-graphics-primitives/nerf_network.h,
// dependencies/neus2_tcnn/...).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may load this library.
//
// PARITY STATUS: PINNED against outputs of the reference itself.  The reference ships no golden vectors, KATs or unit tests for this
// path, so the unmodified reference is compiled for sm_100 by a committed recipe (oracle/Makefile.ref -> oracle/_ref/,
// light draw pinned by oracle/ref_prelude.h), run on the B200 box by committed scripts (tests/ref_pin*.py, tools/resume_probe.py), and every
// dumped step is replayed on this restatement: initial parameters bit-exact, sample counts exact, per-ray losses / SDF / normals within 1e-3
// (tests/golden/ref_pin_summary_*.json; default network with all 14 levels live: tests/golden/ref_pin_summary_full_700steps.json and the
// fixture tests/golden/ref_full_probe.npz).  Committed fixtures carry the reference's vectors into the CPU suite (tests/golden/ref_small*.npz,
// tests/test_reference_golden.py).  Further pins: public known-answer vectors for pcg32 / Morton codes / hash indices (tests/test_oracle_kat.py)
// and finite-difference checks of the analytic first- and second-order backward (tests/test_oracle_gradcheck.py).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <thread>
#include <functional>

namespace orc {

typedef _Float16 half_t;

// round-to-nearest-even through IEEE binary16 and back
static inline float hq(float x) { return (float)(half_t)x; }
static inline float hq_d(double x) { return (float)(half_t)x; }
// binary16 arithmetic on operands that are already binary16 values.  The product of two 11-bit significands is exact
// in fp32; a sum is exact in fp32 unless the exponents differ by more than 12, in which case the fp32 result is
// already within a quarter binary16-ulp of the larger operand — so rounding fp32 -> binary16 (hardware F16C) gives the
// correctly rounded result except on measure-zero ties.
static inline float hadd(float a, float b) { return hq(a + b); }
static inline float hsub(float a, float b) { return hq(a - b); }
static inline float hmul(float a, float b) { return hq(a * b); }

static inline uint16_t half_bits(float x) { half_t h = (half_t)x; uint16_t u; std::memcpy(&u, &h, 2); return u; }
static inline float half_from_bits(uint16_t u) { half_t h; std::memcpy(&h, &u, 2); return (float)h; }

// ---------------------------------------------------------------------------------------------
// pcg32 — M.E. O'Neill's PCG-XSH-RR 64/32 (public algorithm, pcg-random.org), in the
// variant vendored by the reference at dependencies/neus2_tcnn/dependencies/pcg32/pcg32.h:40-170.
// ---------------------------------------------------------------------------------------------
struct Pcg32 {
	uint64_t state, inc;
	static constexpr uint64_t MULT = 0x5851f42d4c957f2dULL;
	Pcg32() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
	explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) { seed(initstate, initseq); }
	void seed(uint64_t initstate, uint64_t initseq = 1u) {
		state = 0u;
		inc = (initseq << 1u) | 1u;
		next_uint();
		state += initstate;
		next_uint();
	}
	uint32_t next_uint() {
		uint64_t old = state;
		state = old * MULT + inc;
		uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t)(old >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	float next_float() {
		union { uint32_t u; float f; } x;
		x.u = (next_uint() >> 9) | 0x3f800000u;
		return x.f - 1.0f;
	}
	void advance(int64_t delta_ = (1ll << 32)) {
		uint64_t cur_mult = MULT, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		uint64_t delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) {
				acc_mult *= cur_mult;
				acc_plus = acc_plus * cur_mult + cur_plus;
			}
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
};

// Morton helpers — tcnn common_device.h:338-363
static inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
	return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

static inline uint32_t next_multiple(uint32_t v, uint32_t d) { return ((v + d - 1) / d) * d; }

// simple static-partition parallel for (threads<=1 → serial, deterministic)
template <typename F>
static void parallel_for(int threads, size_t n, F&& f) {
	if (threads <= 1 || n < 2) { f(0, (size_t)0, n); return; }
	std::vector<std::thread> pool;
	size_t chunk = (n + threads - 1) / threads;
	for (int t = 0; t < threads; ++t) {
		size_t b = std::min(n, (size_t)t * chunk), e = std::min(n, b + chunk);
		if (b >= e) break;
		pool.emplace_back([&f, t, b, e]() { f(t, b, e); });
	}
	for (auto& th : pool) th.join();
}

static inline void atomic_add_f32(float* p, float v) {
	uint32_t* ip = (uint32_t*)p;
	uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED);
	for (;;) {
		float f; std::memcpy(&f, &old, 4);
		f += v;
		uint32_t nw; std::memcpy(&nw, &f, 4);
		if (__atomic_compare_exchange_n(ip, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return;
	}
}

} // namespace orc
