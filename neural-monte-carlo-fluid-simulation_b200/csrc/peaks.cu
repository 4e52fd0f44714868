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

// SM clock under load: cycles (clock64) per nanosecond (globaltimer) over a busy window, on every SM's first CTA.
__global__ void __launch_bounds__(256) clockProbe(unsigned long long* out, float a, float b) {
	unsigned long long t0 = 0, c0 = 0;
	if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); c0 = clock64(); }
	float v[kChains];
#pragma unroll
	for (int k = 0; k < kChains; k++) v[k] = a + (float)(threadIdx.x + k);
#pragma unroll 1
	for (int i = 0; i < 4*kIters; i++) {
#pragma unroll
		for (int k = 0; k < kChains; k++) v[k] = fmaf(v[k], a, b);
	}
	float s = 0.0f;
#pragma unroll
	for (int k = 0; k < kChains; k++) s += v[k];
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long t1, c1 = clock64();
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		out[2*blockIdx.x] = c1 - c0; out[2*blockIdx.x + 1] = t1 - t0;
	}
	if (s == 12345.678f) out[0] = (unsigned long long)s;
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

// Issue limit of the device: every SM sub-partition (4 per SM) dispatches at most one warp instruction per clock, so
//   out2[0] = 4 x SMs x (SM clock measured under load)   [warp instructions / s],   out2[1] = that clock in Hz.
// (A pure FFMA stream reaches ~2/3 of it on B200 -- nmc_measure_peaks()[0] -- because the FMA pipe, not dispatch, limits it;
// a kernel that mixes FMA, ALU, XU and LSU instructions, like the walk kernel, can exceed the FFMA rate.)
extern "C" int nmc_measure_issue_peak(int device, float* out2) {
	if (!out2) return NMC_ERR_INVALID;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return NMC_ERR_NO_DEVICE; }
	if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return NMC_ERR_CUDA; }
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	const int grid = sms*8;
	unsigned long long* d = nullptr;
	if (cudaMalloc((void**)&d, sizeof(unsigned long long)*2*grid) != cudaSuccess) return NMC_ERR_CUDA;
	unsigned long long* hbuf = new unsigned long long[2*grid];
	double best = 0.0;
	cudaError_t e = cudaSuccess;
	for (int rep = 0; rep < 3 && e == cudaSuccess; rep++) {
		clockProbe<<<grid, 256>>>(d, 0.999f, 1e-3f);
		e = cudaMemcpy(hbuf, d, sizeof(unsigned long long)*2*grid, cudaMemcpyDeviceToHost);
		if (e != cudaSuccess) break;
		double cyc = 0.0, ns = 0.0;
		for (int i = 0; i < grid; i++) { cyc += (double)hbuf[2*i]; ns += (double)hbuf[2*i + 1]; }
		if (rep > 0 && ns > 0.0 && cyc/ns > best) best = cyc/ns; // GHz
	}
	delete[] hbuf;
	cudaFree(d);
	if (e != cudaSuccess) { cudaGetLastError(); return NMC_ERR_CUDA; }
	out2[1] = (float)(best*1e9);
	out2[0] = (float)(4.0*sms*best*1e9);
	return NMC_OK;
}
