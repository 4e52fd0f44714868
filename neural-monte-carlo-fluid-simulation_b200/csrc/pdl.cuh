// csrc/pdl.cuh -- programmatic dependent launch for the kernels of a fit iteration.
//
// A captured fit iteration is a chain of five or six short kernels (2 - 40 us) on one stream; each pays its own ramp -- block
// dispatch, first instruction fetch, parameter loads -- after the previous one has drained.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization a kernel may become resident while its predecessor still runs; its first
// instructions are gridEnter(): `griddepcontrol.launch_dependents` (lets ITS successor be scheduled as soon as every CTA of this
// grid has started) and `griddepcontrol.wait` (blocks until the predecessor grid has completed and its memory is visible).
// Nothing is read or written before the wait, so the chain computes what the serialised launches compute; only the ramps
// overlap.  A kernel launched without the attribute (or after a kernel that never triggers) sees both instructions as no-ops.
// NMC_PDL=0 launches without the attribute (A/B measurements).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#ifndef NMC_PDL_DEFAULT
#define NMC_PDL_DEFAULT 0 // flipped to 1 once the A/B on the B200 favours it (profiles/)
#endif

namespace nmc_pdl {

__device__ __forceinline__ void gridEnter() {
	asm volatile("griddepcontrol.launch_dependents;");
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool enabled() {
	static const bool on = [] { const char* e = getenv("NMC_PDL"); return e ? e[0] != '0' : NMC_PDL_DEFAULT != 0; }();
	return on;
}

inline cudaLaunchAttribute attribute() {
	cudaLaunchAttribute a = {};
	a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	a.val.programmaticStreamSerializationAllowed = 1;
	return a;
}

template <class... P, class... A>
inline cudaError_t launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute at[1] = {attribute()};
	cfg.attrs = at; cfg.numAttrs = enabled() ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, static_cast<A&&>(args)...);
}

} // namespace nmc_pdl
