// rnb_dataset.cu — dataset ingest: the image half of load_nerf (reference src/nerf_loader.cu:556-760).
// Host code: a PNG decoder with the conventions of stbi_load_16(path, &w, &h, &comp, 4) as the loader calls it (:612, :653) —
// every source format ends up as 16-bit RGBA, 8-bit samples widened to v * 257, grey replicated to RGB, missing alpha = 65535,
// tRNS colour keys honoured — on a pool of host threads, decoding straight into pinned memory, each image handed to the copy
// engine as soon as it is decoded so that the upload of image k overlaps the inflate of image k+1.
// zlib (inflate) is the only dependency; the transform.json half stays with the caller (nlohmann::json in the reference,
// rnb-neus2_b200/dataset.py in this repo's Python mirror).
#include <cuda_runtime.h>
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <mutex>
#include <chrono>
#include <algorithm>

namespace rnb {

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Decodes `path` into out (w * h * 4 uint16, caller-provided through alloc(w, h)).  Returns "" or an error message.
constexpr uint32_t PNG_MAX_DIM = 1u << 24;      // stb_image's STBI_MAX_DIMENSIONS: larger headers are treated as corrupt, never allocated for

template <typename Alloc>
static std::string decode_png_rgba16_impl(const char* path, uint32_t* w_out, uint32_t* h_out, Alloc alloc) {
	FILE* f = fopen(path, "rb");
	if (!f) return std::string("image not found: ") + path;
	std::vector<uint8_t> file;
	{
		fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
		if (sz < 8) { fclose(f); return std::string("not a PNG file: ") + path; }
		file.resize((size_t)sz);
		const size_t got = fread(file.data(), 1, file.size(), f);
		fclose(f);
		if (got != file.size()) return std::string("short read: ") + path;
	}
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	if (memcmp(file.data(), sig, 8) != 0) return std::string("not a PNG file: ") + path;
	uint32_t w = 0, h = 0; int depth = 0, ctype = -1, interlace = 0;
	std::vector<uint8_t> idat, plte, trns;
	size_t o = 8; bool seen_end = false;
	while (o + 12 <= file.size() && !seen_end) {
		const uint32_t len = be32(&file[o]); const uint8_t* type = &file[o + 4];
		if (o + 12 + (size_t)len > file.size()) return std::string("truncated PNG chunk: ") + path;
		const uint8_t* data = &file[o + 8];
		if (!memcmp(type, "IHDR", 4)) {
			if (len < 13) return std::string("bad IHDR: ") + path;
			w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
		} else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
		else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
		else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
		else if (!memcmp(type, "IEND", 4)) seen_end = true;
		o += 12 + (size_t)len;
	}
	if (w == 0 || h == 0 || ctype < 0) return std::string("PNG without IHDR: ") + path;
	if (w > PNG_MAX_DIM || h > PNG_MAX_DIM) return std::string("PNG too large (corrupt header?): ") + path;
	if (interlace) return std::string("interlaced PNG is not supported: ") + path;
	int channels;
	switch (ctype) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break;
		default: return std::string("bad PNG colour type: ") + path; }
	if (!((depth == 8 || depth == 16) && ctype != 3) && !(ctype == 3 && depth == 8)) return std::string("unsupported PNG bit depth: ") + path;
	const size_t bpp = (size_t)channels * depth / 8, stride = (size_t)w * bpp;
	// a deflate stream expands at most 1032:1: a header that promises more pixels than the compressed data can hold is corrupt, and
	// must not drive an allocation
	if ((stride + 1) * (size_t)h > idat.size() * 1032 + 1024) return std::string("corrupt PNG data: ") + path;
	std::vector<uint8_t> raw((stride + 1) * h);
	{
		uLongf dl = (uLongf)raw.size();
		const int rc = uncompress(raw.data(), &dl, idat.data(), (uLong)idat.size());
		if (rc != Z_OK || dl != raw.size()) return std::string("corrupt PNG data: ") + path;
	}
	// undo the scanline filters in place (PNG spec 9.2); `prev` is the reconstructed row above
	std::vector<uint8_t> zero(stride, 0);
	for (uint32_t y = 0; y < h; ++y) {
		uint8_t* row = &raw[(stride + 1) * y + 1]; const uint8_t ft = row[-1];
		const uint8_t* prev = y ? &raw[(stride + 1) * (y - 1) + 1] : zero.data();
		switch (ft) {
			case 0: break;
			case 1: for (size_t i = bpp; i < stride; ++i) row[i] = (uint8_t)(row[i] + row[i - bpp]); break;
			case 2: for (size_t i = 0; i < stride; ++i) row[i] = (uint8_t)(row[i] + prev[i]); break;
			case 3:
				for (size_t i = 0; i < bpp; ++i) row[i] = (uint8_t)(row[i] + (prev[i] >> 1));
				for (size_t i = bpp; i < stride; ++i) row[i] = (uint8_t)(row[i] + ((row[i - bpp] + prev[i]) >> 1));
				break;
			case 4:
				for (size_t i = 0; i < bpp; ++i) row[i] = (uint8_t)(row[i] + prev[i]);
				for (size_t i = bpp; i < stride; ++i) {
					const int a = row[i - bpp], b = prev[i], c = prev[i - bpp], p = a + b - c;
					const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
					row[i] = (uint8_t)(row[i] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
				}
				break;
			default: return std::string("bad PNG filter type: ") + path;
		}
	}
	uint16_t* out = alloc(w, h);
	if (!out) return "out of memory";
	// sample -> 16 bit: big-endian words as they are, bytes widened to (v << 8) + v  (stbi__convert_8_to_16)
	auto sample = [&](const uint8_t* px, int c) -> uint16_t { return depth == 16 ? (uint16_t)((px[2 * c] << 8) | px[2 * c + 1]) : (uint16_t)(px[c] * 257u); };
	uint16_t key[3] = {0, 0, 0}; bool has_key = false;
	if ((ctype == 0 && trns.size() >= 2) || (ctype == 2 && trns.size() >= 6)) {      // colour key: 16-bit values; for 8-bit images the low byte, widened like the samples
		has_key = true;
		for (int c = 0; c < (ctype == 0 ? 1 : 3); ++c) { const uint16_t v = (uint16_t)((trns[2 * c] << 8) | trns[2 * c + 1]); key[c] = depth == 16 ? v : (uint16_t)((v & 255u) * 257u); }
	}
	for (uint32_t y = 0; y < h; ++y) {
		const uint8_t* row = &raw[(stride + 1) * y + 1];
		uint16_t* dst = out + (size_t)y * w * 4;
		for (uint32_t x = 0; x < w; ++x, dst += 4) {
			const uint8_t* px = row + (size_t)x * bpp;
			switch (ctype) {
				case 0: { const uint16_t g = sample(px, 0); dst[0] = dst[1] = dst[2] = g; dst[3] = (has_key && g == key[0]) ? 0 : 65535; } break;
				case 2: { dst[0] = sample(px, 0); dst[1] = sample(px, 1); dst[2] = sample(px, 2); dst[3] = (has_key && dst[0] == key[0] && dst[1] == key[1] && dst[2] == key[2]) ? 0 : 65535; } break;
				case 3: {
					const size_t i = px[0];
					if (3 * i + 2 >= plte.size()) return std::string("PNG palette index out of range: ") + path;
					dst[0] = (uint16_t)(plte[3 * i] * 257u); dst[1] = (uint16_t)(plte[3 * i + 1] * 257u); dst[2] = (uint16_t)(plte[3 * i + 2] * 257u);
					dst[3] = (uint16_t)((i < trns.size() ? trns[i] : 255u) * 257u);
				} break;
				case 4: { const uint16_t g = sample(px, 0); dst[0] = dst[1] = dst[2] = g; dst[3] = sample(px, 1); } break;
				default: { dst[0] = sample(px, 0); dst[1] = sample(px, 1); dst[2] = sample(px, 2); dst[3] = sample(px, 3); } break;
			}
		}
	}
	*w_out = w; *h_out = h;
	return "";
}

// the C ABI never lets an exception out: allocation failures of the decoder's scratch vectors become an error string
template <typename Alloc>
static std::string decode_png_rgba16(const char* path, uint32_t* w_out, uint32_t* h_out, Alloc alloc) {
	try { return decode_png_rgba16_impl(path, w_out, h_out, alloc); }
	catch (const std::exception& ex) { return std::string("out of memory while decoding ") + path + " (" + ex.what() + ")"; }
}

std::string load_png_rgba16_host(const char* path, uint32_t* w, uint32_t* h, uint16_t** pixels) {
	*pixels = nullptr;
	uint16_t* buf = nullptr;
	const std::string e = decode_png_rgba16(path, w, h, [&](uint32_t ww, uint32_t hh) { buf = (uint16_t*)malloc((size_t)ww * hh * 8); return buf; });
	if (!e.empty()) { free(buf); return e; }
	*pixels = buf;
	return "";
}

// width / height from the IHDR chunk (the first 33 bytes of a PNG)
static std::string png_size(const char* path, uint32_t* w, uint32_t* h) {
	FILE* f = fopen(path, "rb");
	if (!f) return std::string("image not found: ") + path;
	uint8_t hd[33]; const size_t got = fread(hd, 1, sizeof(hd), f); fclose(f);
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	if (got != sizeof(hd) || memcmp(hd, sig, 8) != 0 || memcmp(hd + 12, "IHDR", 4) != 0) return std::string("not a PNG file: ") + path;
	*w = be32(hd + 16); *h = be32(hd + 20);
	if (*w == 0 || *h == 0) return std::string("PNG without IHDR: ") + path;
	if (*w > PNG_MAX_DIM || *h > PNG_MAX_DIM) return std::string("PNG too large (corrupt header?): ") + path;
	return "";
}

// n images (paths[i] may be null: slot skipped) -> ONE device allocation *arena (cudaMalloc, caller owns) with image i at dev[i]
// (256-byte aligned), sizes wh[2i], wh[2i+1].  The headers are read first, so that device memory and the pinned staging (one slot
// per thread in *stage, grow-only, kept by the caller between calls) are allocated once; then `threads` host threads inflate, each
// into its own slot, and hand the image to the copy engine at once (cudaMemcpyAsync + an event guarding the slot's reuse).
std::string load_images_to_device(cudaStream_t st, uint32_t n, const char* const* paths, uint32_t threads, void** arena, void** dev, uint32_t* wh, void** stage, size_t* stage_bytes) {
	if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
	threads = std::min<uint32_t>(threads, std::max<uint32_t>(n, 1));
	const bool debug = getenv("RNB_DATASET_DEBUG") != nullptr;
	auto now_us = []() { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	const long long wall0 = now_us();
	*arena = nullptr;
	std::vector<size_t> off(n, 0); size_t total = 0, max_bytes = 0;
	for (uint32_t i = 0; i < n; ++i) {
		dev[i] = nullptr; wh[2 * i] = wh[2 * i + 1] = 0;
		if (!paths[i]) continue;
		const std::string e = png_size(paths[i], &wh[2 * i], &wh[2 * i + 1]);
		if (!e.empty()) return e;
		const size_t bytes = (size_t)wh[2 * i] * wh[2 * i + 1] * 8;
		off[i] = total; total += (bytes + 255) & ~(size_t)255; max_bytes = std::max(max_bytes, bytes);
	}
	if (total == 0) return "";
	if (cudaMalloc(arena, total) != cudaSuccess) { *arena = nullptr; return std::string("cudaMalloc of the dataset failed: ") + cudaGetErrorString(cudaGetLastError()); }
	const size_t slot = (max_bytes + 4095) & ~(size_t)4095, need = slot * threads;
	if (*stage_bytes < need) {
		if (*stage) cudaFreeHost(*stage);
		*stage = nullptr; *stage_bytes = 0;
		if (cudaMallocHost(stage, need) != cudaSuccess) { cudaFree(*arena); *arena = nullptr; return std::string("pinned staging allocation failed: ") + cudaGetErrorString(cudaGetLastError()); }
		*stage_bytes = need;
	}
	const long long setup_us = now_us() - wall0;
	std::atomic<uint32_t> next{0};
	std::atomic<bool> failed{false};
	std::mutex mu; std::string err;
	std::atomic<long long> us_wait{0}, us_decode{0}, us_enqueue{0};
	std::vector<std::thread> pool;
	int device = 0; cudaGetDevice(&device);
	for (uint32_t t = 0; t < threads; ++t) pool.emplace_back([&, device, t]() {
		cudaSetDevice(device);
		uint16_t* pinned = (uint16_t*)((char*)*stage + slot * t);
		cudaEvent_t done; cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
		bool in_flight = false;
		for (;;) {
			const uint32_t i = next.fetch_add(1);
			if (i >= n || failed.load()) break;
			if (!paths[i]) continue;
			uint32_t w = 0, h = 0; long long wait_us = 0; const long long t0 = now_us();
			std::string e = decode_png_rgba16(paths[i], &w, &h, [&](uint32_t ww, uint32_t hh) -> uint16_t* {
				if (ww != wh[2 * i] || hh != wh[2 * i + 1]) return nullptr;
				const long long p0 = now_us();
				if (in_flight) { cudaEventSynchronize(done); in_flight = false; }       // the previous image of this thread has left the slot
				wait_us = now_us() - p0;
				return pinned;
			});
			const long long t1 = now_us();
			us_wait += wait_us; us_decode += t1 - t0 - wait_us;
			if (e.empty()) {
				const size_t bytes = (size_t)w * h * 8;
				dev[i] = (char*)*arena + off[i];
				if (cudaMemcpyAsync(dev[i], pinned, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaEventRecord(done, st) != cudaSuccess) e = std::string("upload failed: ") + cudaGetErrorString(cudaGetLastError());
				else in_flight = true;
				us_enqueue += now_us() - t1;
			}
			if (!e.empty()) { std::lock_guard<std::mutex> l(mu); if (err.empty()) err = e; failed = true; }
		}
		if (in_flight) cudaEventSynchronize(done);
		cudaEventDestroy(done);
	});
	for (auto& th : pool) th.join();
	cudaStreamSynchronize(st);
	if (debug) fprintf(stderr, "[rnb dataset] %u images, %u threads: wall %.1f ms (headers + allocations %.1f ms); summed over threads: decode %.1f ms, waiting for the slot %.1f ms, enqueue %.1f ms\n",
	                   n, threads, (now_us() - wall0) * 1e-3, setup_us * 1e-3, us_decode.load() * 1e-3, us_wait.load() * 1e-3, us_enqueue.load() * 1e-3);
	if (!err.empty()) { cudaFree(*arena); *arena = nullptr; for (uint32_t i = 0; i < n; ++i) dev[i] = nullptr; }
	return err;
}

} // namespace rnb
