// csrc/fit_glue.cu -- the per-iteration glue of the fit loops (_training_loop, src/2d/models/base.py:129-152) as single
// launches: drawing the batch (sample_in_training 'random', base.py:225-241, utils/model_utils.py:22-31) and gathering the
// projection fit's batch from the pressure samples (model_split.py:272-277).  With stock torch ops these are 3-8 launches of
// ~2 us each per iteration (rand, scale, shift, obstacle test, where, two index kernels), a fifth of a 100 us iteration.
// The draws are keyed by (seed, epoch, step, element) with the counters in device memory, so the captured iteration draws
// fresh numbers at every graph replay.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nmcfs_siren.h"
#include "pdl.cuh"

namespace nmc_siren_detail { void setError(const char* m); }

namespace {

int fail(const char* m) { nmc_siren_detail::setError(m); return 1; }

struct Box { float lo[3], ext[3], c[3], r; };

// splitmix64 finaliser: a bijection of the 64-bit key with full avalanche
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27))*0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t iterationKey(uint64_t seed, const long long* step, const long long* epoch) {
	return mix64(mix64(seed ^ ((uint64_t)*epoch << 32)) + (uint64_t)*step);
}
__device__ __forceinline__ float unit24(uint32_t bits) { return (float)(bits >> 8)*(1.0f/16777216.0f); }

__global__ void fitSampleUniform(int dim, Box b, long long n, float* __restrict__ out, const long long* __restrict__ step,
								 const long long* __restrict__ epoch, uint64_t seed) {
	const uint64_t key = iterationKey(seed, step, epoch);
	for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) {
		float x[3] = {0.0f, 0.0f, 0.0f};
		for (int draw = 0; draw < 2; draw++) {
			const uint64_t a = mix64(key + 4ull*(uint64_t)i + 2ull*(uint64_t)draw), c = mix64(a);
			const float u[3] = {unit24((uint32_t)a), unit24((uint32_t)(a >> 32)), unit24((uint32_t)c)};
			float d2 = 0.0f;
			for (int k = 0; k < dim; k++) {
				x[k] = u[k]*b.ext[k] + b.lo[k];
				d2 += (x[k] - b.c[k])*(x[k] - b.c[k]);
			}
			if (!(b.r > 0.0f) || sqrtf(d2) - b.r > 0.0f) break;   // outside the obstacle (or no obstacle): keep; inside: one redraw
		}
		for (int k = 0; k < dim; k++) out[i*dim + k] = x[k];
	}
}

__global__ void fitGather(int dim, long long n, const float* __restrict__ srcX, const float* __restrict__ srcG, const float* __restrict__ count,
						  long long cap, float* __restrict__ outX, float* __restrict__ outG, const long long* __restrict__ step,
						  const long long* __restrict__ epoch, uint64_t seed) {
	const uint64_t key = iterationKey(seed ^ 0x5851F42D4C957F2Dull, step, epoch);
	const float cnt = *count;
	for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) {
		const float u = unit24((uint32_t)mix64(key + (uint64_t)i));
		long long idx = (long long)(u*cnt);
		idx = idx < 0 ? 0 : (idx > cap - 1 ? cap - 1 : idx);
		for (int k = 0; k < dim; k++) { outX[i*dim + k] = __ldg(&srcX[idx*dim + k]); outG[i*dim + k] = __ldg(&srcG[idx*dim + k]); }
	}
}

// one slot of the target ring -> the fixed buffers a captured iteration reads; the slot is Adam's device-side step modulo `slots`
__global__ void fitFetch(long long count, int slots, const float* __restrict__ ringX, const float* __restrict__ ringT, const float* __restrict__ ringS,
						 const long long* __restrict__ step, float* __restrict__ outX, float* __restrict__ outT, float* __restrict__ outS, int vec) {
	nmc_pdl::gridEnter();
	const long long base = (*step % slots)*count;
	const long long tid = (long long)blockIdx.x*blockDim.x + threadIdx.x, nth = (long long)gridDim.x*blockDim.x;
	if (vec) {
		const float4* x4 = reinterpret_cast<const float4*>(ringX + base); const float4* t4 = reinterpret_cast<const float4*>(ringT + base);
		const float4* s4 = ringS ? reinterpret_cast<const float4*>(ringS + base) : nullptr;
		for (long long i = tid; i < (count >> 2); i += nth) {
			reinterpret_cast<float4*>(outX)[i] = x4[i]; reinterpret_cast<float4*>(outT)[i] = t4[i];
			if (s4) reinterpret_cast<float4*>(outS)[i] = s4[i];
		}
	} else {
		for (long long i = tid; i < count; i += nth) {
			outX[i] = ringX[base + i]; outT[i] = ringT[base + i];
			if (ringS) outS[i] = ringS[base + i];
		}
	}
}

unsigned gridFor(long long n) { long long b = (n + 255)/256; return (unsigned)(b < 1 ? 1 : (b > 1184 ? 1184 : b)); }

} // namespace

extern "C" int nmc_fit_sample_uniform(int dim, const float* lo, const float* hi, int64_t n, float* out, const long long* step,
									  const long long* epoch, uint64_t seed, const float* obstacle, void* stream) {
	if (dim < 1 || dim > 3) return fail("dim must be 1, 2 or 3");
	if (n <= 0) return 0;
	if (!lo || !hi || !out || !step || !epoch) return fail("null argument");
	Box b = {};
	for (int k = 0; k < dim; k++) { b.lo[k] = lo[k]; b.ext[k] = hi[k] - lo[k]; b.c[k] = obstacle ? obstacle[k] : 0.0f; }
	b.r = obstacle ? obstacle[dim] : 0.0f;
	fitSampleUniform<<<gridFor(n), 256, 0, (cudaStream_t)stream>>>(dim, b, n, out, step, epoch, seed);
	cudaError_t e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_fit_gather(int dim, int64_t n, const float* src_x, const float* src_g, const float* count, int64_t cap, float* out_x,
							  float* out_g, const long long* step, const long long* epoch, uint64_t seed, void* stream) {
	if (dim < 1 || dim > 3) return fail("dim must be 1, 2 or 3");
	if (n <= 0) return 0;
	if (!src_x || !src_g || !count || !out_x || !out_g || !step || !epoch || cap < 1) return fail("bad arguments");
	fitGather<<<gridFor(n), 256, 0, (cudaStream_t)stream>>>(dim, n, src_x, src_g, count, cap, out_x, out_g, step, epoch, seed);
	cudaError_t e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_fit_fetch(int64_t count, int slots, const float* ring_x, const float* ring_t, const float* ring_s, const long long* step,
							 float* out_x, float* out_t, float* out_s, void* stream) {
	if (count <= 0) return 0;
	if (slots < 1 || !ring_x || !ring_t || !step || !out_x || !out_t || (ring_s && !out_s)) return fail("bad arguments");
	const int vec = (count % 4 == 0) && ((((uintptr_t)ring_x | (uintptr_t)ring_t | (uintptr_t)ring_s | (uintptr_t)out_x | (uintptr_t)out_t | (uintptr_t)out_s) & 15) == 0);
	cudaError_t e = nmc_pdl::launch(fitFetch, dim3(gridFor(vec ? count/4 : count)), dim3(256), 0, (cudaStream_t)stream, (long long)count, slots, ring_x, ring_t, ring_s, step, out_x, out_t, out_s, vec);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}
