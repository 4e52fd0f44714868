// csrc/fields.cu -- grid post-processing on the device (include/nmcfs_fields.h): semi-Lagrangian density advection
// (scipy.ndimage.map_coordinates(order=1) semantics) and the squared velocity error reduction.  Both are one
// pass over the grid: HBM-bound, thread per node, coalesced along the fastest axis; the gathers of the
// back-traced positions stay within a cell or two of the node (dt*|u| << grid spacing * n), so they hit L1/L2.
#include <cuda_runtime.h>
#include <cstdio>
#include "../../include/nmcfs_fields.h"

namespace {

thread_local char g_err[256] = "";
int fail(const char* what, cudaError_t e = cudaSuccess) {
	if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
	else snprintf(g_err, sizeof g_err, "%s", what);
	return 1;
}

struct Grid { int n[3]; float lo[3], scale[3]; }; // scale = n/extent

template <int DIM>
__global__ void advectDensity(Grid G, const float* __restrict__ din, const float* __restrict__ vel, float dt, int mode,
							  float* __restrict__ dout, long long total) {
	for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x*blockDim.x) {
		int idx[3];
		long long r = t;
		for (int a = DIM - 1; a >= 0; a--) { idx[a] = (int)(r % G.n[a]); r /= G.n[a]; }
		float pos[3]; bool outside = false;
		for (int a = 0; a < DIM; a++) {
			// x = lo + i/scale; pos = (x - dt u - lo)*scale = i - dt u scale   (evaluated like the reference, in that order)
			float x = (float)idx[a]/G.scale[a] + G.lo[a];
			float p = (x - dt*vel[t*DIM + a] - G.lo[a])*G.scale[a];
			const float hi = (float)(G.n[a] - 1);
			if (mode == 1) p = fminf(fmaxf(p, 0.0f), hi);
			else if (!(p >= 0.0f && p <= hi)) outside = true;
			pos[a] = p;
		}
		float v = 0.0f;
		if (!outside) {
			int i0[3]; float f[3];
			for (int a = 0; a < DIM; a++) {
				int b = (int)floorf(pos[a]);
				if (b > G.n[a] - 2) b = G.n[a] - 2 > 0 ? G.n[a] - 2 : 0; // pos == n-1: weight 1 on the last node
				i0[a] = b; f[a] = pos[a] - (float)b;
			}
			if (DIM == 2) {
				const float* p = din + (long long)i0[0]*G.n[1] + i0[1];
				const int s0 = G.n[0] > 1 ? G.n[1] : 0, s1 = G.n[1] > 1 ? 1 : 0;
				v = (1.0f - f[0])*((1.0f - f[1])*p[0] + f[1]*p[s1]) + f[0]*((1.0f - f[1])*p[s0] + f[1]*p[s0 + s1]);
			} else {
				const long long s0 = G.n[0] > 1 ? (long long)G.n[1]*G.n[2] : 0, s1 = G.n[1] > 1 ? G.n[2] : 0, s2 = G.n[2] > 1 ? 1 : 0;
				const float* p = din + ((long long)i0[0]*G.n[1] + i0[1])*G.n[2] + i0[2];
				float c00 = (1.0f - f[2])*p[0] + f[2]*p[s2], c01 = (1.0f - f[2])*p[s1] + f[2]*p[s1 + s2];
				float c10 = (1.0f - f[2])*p[s0] + f[2]*p[s0 + s2], c11 = (1.0f - f[2])*p[s0 + s1] + f[2]*p[s0 + s1 + s2];
				v = (1.0f - f[0])*((1.0f - f[1])*c00 + f[1]*c01) + f[0]*((1.0f - f[1])*c10 + f[1]*c11);
			}
		}
		dout[t] = v;
	}
}

__global__ void sumSquaredError(const float* __restrict__ u, const float* __restrict__ ref, long long nFloats, double* __restrict__ out) {
	double acc = 0.0;
	for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < nFloats; t += (long long)gridDim.x*blockDim.x) {
		float d = u[t] - ref[t];
		acc += (double)d*(double)d;
	}
	for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
	__shared__ double part[8];
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0.0;
		for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += part[w];
		atomicAdd(out, s);
	}
}

struct Box3 { float lo[3], hi[3]; };
// semi-Lagrangian back-trace of the fit loops: out = clamp(x - dt u, lo, hi)  (model_split.py:100-103)
__global__ void backtrace(const float* __restrict__ x, const float* __restrict__ u, long long count, int dim, float dt, Box3 b,
						  float* __restrict__ out) {
	for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < count; t += (long long)gridDim.x*blockDim.x) {
		const int a = (int)(t % dim);
		out[t] = fminf(fmaxf(x[t] - u[t]*dt, b.lo[a]), b.hi[a]);
	}
}

int smCount() {
	static int n = 0;
	if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
	return n;
}

} // namespace

extern "C" {

const char* nmc_fields_last_error(void) { return g_err; }

int nmc_advect_density(int dim, const int* shape, const float* d_in, const float* vel, float dt, const float* lo,
					   const float* extent, int mode, float* d_out, void* stream) {
	if (dim != 2 && dim != 3) return fail("dim must be 2 or 3");
	if (!shape || !d_in || !vel || !lo || !extent || !d_out) return fail("null argument");
	if (d_in == d_out) return fail("d_out must not alias d_in");
	if (mode != 0 && mode != 1) return fail("mode must be 0 (constant) or 1 (nearest)");
	Grid G; long long total = 1;
	for (int a = 0; a < 3; a++) { G.n[a] = 1; G.lo[a] = 0.0f; G.scale[a] = 1.0f; }
	for (int a = 0; a < dim; a++) {
		if (shape[a] <= 0 || !(extent[a] > 0.0f)) return fail("empty grid or non-positive extent");
		G.n[a] = shape[a]; G.lo[a] = lo[a]; G.scale[a] = (float)shape[a]/extent[a]; total *= shape[a];
	}
	const int block = 256;
	long long blocks = (total + block - 1)/block, cap = 32ll*smCount();
	unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
	if (dim == 2) advectDensity<2><<<grid, block, 0, (cudaStream_t)stream>>>(G, d_in, vel, dt, mode, d_out, total);
	else advectDensity<3><<<grid, block, 0, (cudaStream_t)stream>>>(G, d_in, vel, dt, mode, d_out, total);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : fail("advectDensity launch", e);
}

int nmc_backtrace(int dim, const float* x, const float* u, int64_t n, float dt, const float* lo, const float* hi, float* out, void* stream) {
	if (dim < 1 || dim > 3) return fail("dim must be 1..3");
	if (n <= 0) return 0;
	if (!x || !u || !lo || !hi || !out) return fail("null argument");
	Box3 b;
	for (int a = 0; a < 3; a++) { b.lo[a] = a < dim ? lo[a] : 0.0f; b.hi[a] = a < dim ? hi[a] : 0.0f; }
	const long long count = (long long)n*dim;
	const int block = 256;
	long long blocks = (count + block - 1)/block, cap = 16ll*smCount();
	backtrace<<<(unsigned)(blocks < cap ? blocks : cap), block, 0, (cudaStream_t)stream>>>(x, u, count, dim, dt, b, out);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : fail("backtrace launch", e);
}

int nmc_sum_squared_error(int dim, const float* u, const float* u_ref, int64_t n, double* out_sum, void* stream) {
	if (dim < 1 || dim > 3) return fail("dim must be 1..3");
	if (n <= 0) return 0;
	if (!u || !u_ref || !out_sum) return fail("null argument");
	const int block = 256;
	long long nf = (long long)n*dim, blocks = (nf + block - 1)/block, cap = 16ll*smCount();
	sumSquaredError<<<(unsigned)(blocks < cap ? blocks : cap), block, 0, (cudaStream_t)stream>>>(u, u_ref, nf, out_sum);
	cudaError_t e = cudaGetLastError();
	return e == cudaSuccess ? 0 : fail("sumSquaredError launch", e);
}

} // extern "C"
