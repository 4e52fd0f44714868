// csrc/peaks.cu -- instruction-throughput micro-benchmarks for the issue-bound roofline of the walk kernel.
//
// The walk kernel is bound by instruction issue, not by HBM, so the denominator of its
// roofline is the rate at which the SMs can issue warp instructions.  MEASURED_PEAKS.json holds only HBM and
// bf16 tensor numbers; these kernels measure, on the device the bench runs on and at the clocks it runs at,
//   [0] fp32 FMA   warp-instructions / s  (one per scheduler per clock = the issue limit of an SM sub-partition)
//   [1] MUFU.EX2   warp-instructions / s  (the special-function pipe: exp / log / rsqrt / sin / cos)
//   [2] fp64 FMA   warp-instructions / s  (the deterministic mode's Bessel polynomials)
// Each thread runs 8 independent dependency chains, 32 warps per SM resident, so that the pipe and not latency is
// the limit.
#include "../../include/nmcfs.h"
#include <cuda_runtime.h>

namespace {

constexpr int kIters = 4096, kChains = 8;

__global__ void __launch_bounds__(256) ffmaPeak(float* out, float a, float b) {
	float v[kChains];
#pragma unroll
	for (int k = 0; k < kChains; k++) v[k] = a + (float)(threadIdx.x + k);
#pragma unroll 1
	for (int i = 0; i < kIters; i++) {
#pragma unroll
		for (int k = 0; k < kChains; k++) v[k] = fmaf(v[k], a, b);
	}
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < kChains; k++) s += v[k];
	if (s == 12345.678f) out[0] = s; // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) mufuPeak(float* out, float a) {
	float v[kChains];
#pragma unroll
	for (int k = 0; k < kChains; k++) v[k] = a*(float)(threadIdx.x + k + 1)*1e-3f;
#pragma unroll 1
	for (int i = 0; i < kIters; i++) {
#pragma unroll
		for (int k = 0; k < kChains; k++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[k]));
	}
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < kChains; k++) s += v[k];
	if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) dfmaPeak(float* out, double a, double b) {
	double v[kChains];
#pragma unroll
	for (int k = 0; k < kChains; k++) v[k] = a + (double)(threadIdx.x + k);
#pragma unroll 1
	for (int i = 0; i < kIters/16; i++) {
#pragma unroll
		for (int k = 0; k < kChains; k++) v[k] = fma(v[k], a, b);
	}
	double s = 0.0;
#pragma unroll
	for (int k = 0; k < kChains; k++) s += v[k];
	if (s == 12345.678) out[0] = (float)s;
}

template <class F>
cudaError_t timeKernel(F launch, double warpInstPerLaunch, float* rate) {
	cudaEvent_t e0, e1;
	cudaError_t e = cudaEventCreate(&e0);
	if (e != cudaSuccess) return e;
	e = cudaEventCreate(&e1);
	if (e != cudaSuccess) { cudaEventDestroy(e0); return e; }
	float best = 0.0f;
	for (int rep = 0; rep < 4; rep++) { // first repetition warms up
		cudaEventRecord(e0, 0);
		launch();
		cudaEventRecord(e1, 0);
		cudaEventSynchronize(e1);
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms > 0.0f) { float r = (float)(warpInstPerLaunch/(ms*1e-3)); if (r > best) best = r; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*rate = best;
	return cudaGetLastError();
}

} // namespace

extern "C" int nmc_measure_peaks(int device, float* out3) {
	if (!out3) return NMC_ERR_INVALID;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return NMC_ERR_NO_DEVICE; }
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return NMC_ERR_CUDA; }
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	float* d = nullptr;
	if (cudaMalloc((void**)&d, 16) != cudaSuccess) return NMC_ERR_CUDA;
	const int grid = sms*8, block = 256; // 8 CTAs x 8 warps resident per SM, 4 waves
	const double warps = (double)grid*(block/32);
	cudaError_t e = timeKernel([&] { ffmaPeak<<<grid*4, block>>>(d, 0.999f, 1e-3f); }, 4.0*warps*kIters*kChains, &out3[0]);
	if (e == cudaSuccess) e = timeKernel([&] { mufuPeak<<<grid, block>>>(d, 0.5f); }, warps*kIters*kChains, &out3[1]);
	if (e == cudaSuccess) e = timeKernel([&] { dfmaPeak<<<grid, block>>>(d, 0.999, 1e-3); }, warps*(kIters/16)*kChains, &out3[2]);
	cudaFree(d);
	return e == cudaSuccess ? NMC_OK : NMC_ERR_CUDA;
}
