// Force-included (-include) in every reference .cu translation unit built by oracle/Makefile.ref.
// TEST INFRASTRUCTURE ONLY.  It changes no reference source; it only
//  (1) makes tcnn::parallel_for_gpu visible to unqualified calls inside `namespace ngp` templates, which nvcc 12.9
//      otherwise rejects (ref: include/neural-graphics-primitives/nerf_network.h:908,1026,1036,1068,1078);
//  (2) with -DRNB_PIN_LIGHT, replaces the clock64()-seeded light draw of the loss kernel
//      (ref: src/testbed_nerf.cu:1557-1561, `curand_init(clock64(), i, 0, &state); curand(&state) % 3`) by the
//      deterministic draw `ray_idx % 3` (`ray_idx` is the loss kernel's own local, in scope at that line; the slot `i`
//      comes from atomicAdd arrival order and is not reproducible), so that a step of the reference can be compared
//      with the oracle at all.
#pragma once
#include <tiny-cuda-nn/common.h>
namespace ngp { using tcnn::parallel_for_gpu; }
#ifdef RNB_PIN_LIGHT
#include <curand_kernel.h>
__device__ inline void rnb_pin_light_init(curandState* s, unsigned int ray) { s->d = ray; }
__device__ inline unsigned int rnb_pin_light_draw(curandState* s) { return s->d; }
#define curand_init(seed, seq, off, st) rnb_pin_light_init((st), ray_idx)
#define curand(st) rnb_pin_light_draw(st)
#endif
