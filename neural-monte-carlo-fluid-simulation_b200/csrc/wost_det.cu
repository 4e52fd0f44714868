// csrc/wost_det.cu -- deterministic-mode kernels (one sample point per thread) and the device probes.
// MUST be compiled with -fmad=false: the deterministic mode reproduces the reference's IEEE float
// arithmetic operation by operation (see nmc_math.cuh).
#include "nmc_device.h"
#include "../../include/nmcfs.h"

namespace nmc {

template <int DIM>
__global__ void __launch_bounds__(128)
detKernel(SceneView S, SolverParams o, const float* __restrict__ pts, long long n, unsigned long long indexOffset,
		  float* __restrict__ pOut, float* __restrict__ gOut, float* __restrict__ lhs,
		  Counters* __restrict__ counters, float* __restrict__ stats12) {
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	unsigned started = 0, completed = 0, steps = 0, active = 0;
	if (i < n) {
		V3 x = mk(pts[i*DIM], pts[i*DIM + 1], DIM == 3 ? pts[i*DIM + 2] : 0.0f);
		LhsScratch sc; sc.base = lhs + i; sc.stride = (size_t)n;
		PointResult r;
		detEstimatePoint<DIM>(S, o, x, indexOffset + (unsigned long long)i, sc, r);
		pOut[i] = r.p;
		for (int k = 0; k < DIM; k++) gOut[i*DIM + k] = r.g[k];
		started = r.walksStarted; completed = (unsigned)r.nSol; steps = r.steps; active = (unsigned)r.active;
		if (stats12) { // layout of oracle/ref_harness.cpp ref_wost stats
			float* t = stats12 + i*12;
			int nv = r.nSol - 1 > 1 ? r.nSol - 1 : 1;
			t[0] = r.solMean; t[1] = r.solM2/nv;
			for (int k = 0; k < 3; k++) { t[2 + k] = k < DIM ? r.gradMean[k] : 0.0f; t[5 + k] = k < DIM ? r.gradM2[k]/nv : 0.0f; }
			t[8] = r.meanFirstSource; t[9] = (float)r.nSol;
			t[10] = (float)r.totalWalkLength/(r.nSol > 1 ? r.nSol : 1); t[11] = (float)r.active;
			if (!r.active) for (int k = 0; k < 11; k++) t[k] = 0.0f;
		}
	}
	// warp-aggregate the counters, one atomic per warp
	for (int off = 16; off > 0; off >>= 1) {
		started += __shfl_down_sync(0xffffffffu, started, off);
		completed += __shfl_down_sync(0xffffffffu, completed, off);
		steps += __shfl_down_sync(0xffffffffu, steps, off);
		active += __shfl_down_sync(0xffffffffu, active, off);
	}
	if ((threadIdx.x & 31) == 0 && counters) {
		atomicAdd(&counters->walksStarted, (unsigned long long)started);
		atomicAdd(&counters->walksCompleted, (unsigned long long)completed);
		atomicAdd(&counters->steps, (unsigned long long)steps);
		atomicAdd(&counters->activePoints, (unsigned long long)active);
	}
}

size_t deterministicScratchFloats(int dim, const SolverParams& o, long long n) {
	int nPairs = o.nWalks;
	if (o.useGradientAntitheticVariates) nPairs = o.nWalks/2 > 1 ? o.nWalks/2 : 1;
	return (size_t)(dim - 1)*2*(size_t)nPairs*(size_t)n;
}

cudaError_t launchDeterministic(const SceneView& S, const SolverParams& o, const float* d_pts, long long n,
								unsigned long long indexOffset, float* d_p, float* d_g, float* d_lhs,
								Counters* d_counters, float* d_stats12, cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int block = 128;
	unsigned grid = (unsigned)((n + block - 1)/block);
	if (S.dim == 2) detKernel<2><<<grid, block, 0, stream>>>(S, o, d_pts, n, indexOffset, d_p, d_g, d_lhs, d_counters, d_stats12);
	else detKernel<3><<<grid, block, 0, stream>>>(S, o, d_pts, n, indexOffset, d_p, d_g, d_lhs, d_counters, d_stats12);
	return cudaGetLastError();
}

// ---- solution-only estimator at caller-given sample points (boundary value caching's cache points) ----------------------
template <int DIM>
__global__ void __launch_bounds__(128)
detSolutionKernel(SceneView S, SolverParams o, const float* __restrict__ pts, const float* __restrict__ normals,
				  const int* __restrict__ types, const int* __restrict__ aligned, long long n, int nWalks,
				  unsigned long long indexOffset, float* __restrict__ sol, float* __restrict__ stats4, Counters* __restrict__ counters) {
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	unsigned started = 0, completed = 0, steps = 0;
	if (i < n) {
		V3 x = mk(pts[i*DIM], pts[i*DIM + 1], DIM == 3 ? pts[i*DIM + 2] : 0.0f);
		V3 nr = normals ? mk(normals[i*DIM], normals[i*DIM + 1], DIM == 3 ? normals[i*DIM + 2] : 0.0f) : mk(0, 0, 0);
		float r4[4], firstR = 0.0f;
		detEstimateSolution<DIM>(S, o, x, nr, types ? types[i] : 0, aligned ? aligned[i] != 0 : false, nWalks,
								 indexOffset + (unsigned long long)i, r4, &firstR, started, steps);
		sol[i] = r4[0];
		completed = (unsigned)r4[2];
		if (stats4) { // layout of oracle/ref_harness.cpp ref_estimate_solution: variance, count, mean walk length, first sphere radius
			const int nSol = (int)r4[2];
			stats4[i*4] = r4[1]/(nSol - 1 > 1 ? nSol - 1 : 1); stats4[i*4 + 1] = r4[2];
			stats4[i*4 + 2] = r4[3]/(nSol > 1 ? nSol : 1); stats4[i*4 + 3] = firstR;
		}
	}
	for (int off = 16; off > 0; off >>= 1) {
		started += __shfl_down_sync(0xffffffffu, started, off);
		completed += __shfl_down_sync(0xffffffffu, completed, off);
		steps += __shfl_down_sync(0xffffffffu, steps, off);
	}
	if ((threadIdx.x & 31) == 0 && counters) {
		atomicAdd(&counters->walksStarted, (unsigned long long)started);
		atomicAdd(&counters->walksCompleted, (unsigned long long)completed);
		atomicAdd(&counters->steps, (unsigned long long)steps);
	}
}

cudaError_t launchSolutionEstimator(const SceneView& S, const SolverParams& o, const float* d_pts, const float* d_normals,
									const int* d_types, const int* d_aligned, long long n, int nWalks, unsigned long long indexOffset,
									float* d_sol, float* d_stats4, Counters* d_counters, cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int block = 128;
	unsigned grid = (unsigned)((n + block - 1)/block);
	if (S.dim == 2) detSolutionKernel<2><<<grid, block, 0, stream>>>(S, o, d_pts, d_normals, d_types, d_aligned, n, nWalks, indexOffset, d_sol, d_stats4, d_counters);
	else detSolutionKernel<3><<<grid, block, 0, stream>>>(S, o, d_pts, d_normals, d_types, d_aligned, n, nWalks, indexOffset, d_sol, d_stats4, d_counters);
	return cudaGetLastError();
}

// ---- probes ---------------------------------------------------------------------------------------
int probeWidth(int dim, int kind) {
	switch (kind) {
		case NMC_PROBE_RAY: case NMC_PROBE_RAY_PACKET: return 2 + 2*dim;
		case NMC_PROBE_CLOSEST_PACKET: return 2;
		case NMC_PROBE_GREENS: case NMC_PROBE_GREENS_FAST: return 10;
		case NMC_PROBE_SAMPLE_VOLUME: return 3;
		case NMC_PROBE_SAMPLE_RADIUS_FAST: return 2;
		default: return 1;
	}
}

template <int DIM>
__global__ void probeKernel(SceneView S, int kind, long long n, const float* __restrict__ pts, const float* __restrict__ a0,
							const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ a3,
							const float* __restrict__ params, float* __restrict__ out) {
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	typedef ExactMath M;
	V3 x = mk(0, 0, 0);
	if (pts) x = mk(pts[i*DIM], pts[i*DIM + 1], DIM == 3 ? pts[i*DIM + 2] : 0.0f);
	switch (kind) {
		case NMC_PROBE_DIST_NEUMANN: out[i] = distNeumann<DIM>(S, x, false); break;
		case NMC_PROBE_SIGNED_DIST_NEUMANN: out[i] = distNeumann<DIM>(S, x, true); break;
		case NMC_PROBE_DIST_DIRICHLET: out[i] = distDirichlet<DIM>(S, x); break;
		case NMC_PROBE_INSIDE_DOMAIN: out[i] = insideDomain<DIM>(S, x) ? 1.0f : 0.0f; break;
		case NMC_PROBE_STAR_RADIUS: out[i] = starRadius<DIM, M>(S, x, params[0], a0[i], params[1], params[2] != 0.0f); break;
		case NMC_PROBE_SOURCE: out[i] = sourceAt<DIM>(S, x); break;
		case NMC_PROBE_RAY: {
			V3 nn = mk(a0[i*DIM], a0[i*DIM + 1], DIM == 3 ? a0[i*DIM + 2] : 0.0f);
			V3 d = mk(a1[i*DIM], a1[i*DIM + 1], DIM == 3 ? a1[i*DIM + 2] : 0.0f);
			Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
			bool hit = intersectNeumann<DIM>(S, x, nn, d, a2[i], a3[i] != 0.0f, h);
			float* r = out + i*(2 + 2*DIM);
			r[0] = hit ? 1.0f : 0.0f; r[1] = h.d;
			r[2] = h.p.x; r[3] = h.p.y; if (DIM == 3) r[4] = h.p.z;
			r[2 + DIM] = h.n.x; r[3 + DIM] = h.n.y; if (DIM == 3) r[4 + DIM] = h.n.z;
		} break;
		case NMC_PROBE_GREENS: { // layout of greens_probe (oracle/ref_harness.cpp:107-127)
			float lambda = params[0], R = a0[i], rr = a1[i];
			BallExact<DIM> g; g.init(lambda > 0.0f, lambda);
			V3 c = mk(0, 0, 0); g.update(c, R); g.r = rr;
			V3 ex = mk(1, 0, 0), el = DIM == 2 ? mk(0, 1, 0) : mk(0, 0, 1);
			g.yVol = c + rr*ex; g.ySurf = c + R*el;
			float* o = out + i*10;
			o[0] = g.evaluate(); o[1] = g.norm_(); o[2] = g.gradientNorm(); o[3] = g.poissonKernel();
			o[4] = g.directionSampledPoissonKernel(g.yVol);
			V3 pg = g.poissonKernelGradient(); o[5] = DIM == 2 ? pg.y : pg.z;
			o[6] = g.evaluate(c + (0.25f*R)*el, g.yVol);
			o[7] = g.potential(); o[8] = g.gradient().x; o[9] = 0.0f;
		} break;
		case NMC_PROBE_SAMPLE_VOLUME: {
			float lambda = params[0], R = a0[i];
			const unsigned* sd = reinterpret_cast<const unsigned*>(a1);
			unsigned long long seed = (unsigned long long)sd[2*i] | ((unsigned long long)sd[2*i + 1] << 32);
			Pcg32 rng; rng.seed(seed, 1);
			BallExact<DIM> g; g.init(lambda > 0.0f, lambda);
			g.update(mk(0, 0, 0), R);
			float pdf = 0.0f;
			g.sampleVolume(mk(1, 0, 0), rng, pdf);
			Pcg32 t; t.seed(seed, 1);
			int k = 0; while (t.state != rng.state && k < 4096) { t.nextUInt(); k++; }
			out[i*3] = g.r; out[i*3 + 1] = pdf; out[i*3 + 2] = (float)k;
		} break;
		default: break;
	}
}

cudaError_t launchProbe(const SceneView& S, int kind, long long n, const float* d_pts, const float* a0, const float* a1,
						const float* a2, const float* a3, const float* params, float* d_out, cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	const int block = 128;
	unsigned grid = (unsigned)((n + block - 1)/block);
	if (S.dim == 2) probeKernel<2><<<grid, block, 0, stream>>>(S, kind, n, d_pts, a0, a1, a2, a3, params, d_out);
	else probeKernel<3><<<grid, block, 0, stream>>>(S, kind, n, d_pts, a0, a1, a2, a3, params, d_out);
	return cudaGetLastError();
}

} // namespace nmc
