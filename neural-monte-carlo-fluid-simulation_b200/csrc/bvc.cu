// csrc/bvc.cu -- boundary value caching on the GPU: the solver behind the bindings' second
// entry point, zombie_bindings.bvc(scene, solverConfig, outputConfig) (bindings/zombie/demo/demo.cpp:265-363,
// include/zombie/boundary_value_caching/{boundary_sampler,domain_sampler,splatter}.h), 2D only like the reference's module.
//
//   1. cache points: area-weighted, stratified samples on the boundary primitives (BoundarySampler::generateSamples,
//      boundary_sampler.h:309-403) and stratified samples in the solve region carrying the source term
//      (DomainSampler::generateSamples, domain_sampler.h:33-59) -- host side, a few thousand points;
//   2. the solution at every boundary cache point, estimated by walks that START ON the reflecting boundary
//      (EstimationQuantity::Solution, walk_on_stars.h:354-461): detSolutionKernel (wost_det.cu), one thread per point;
//   3. splatting: every evaluation point sums the free-space kernels of all cache points (Splatter::splat, splatter.h:53-116,
//      203-290): bvcSplatKernel, one thread per evaluation point, cache records staged in shared memory;
//   4. evaluation points closer to the absorbing boundary than the cut-off are re-estimated pointwise (splatter.h:160-180),
//      then the grid is masked like saveEvaluationGrid (demo/grid.h:388-411).
// In both bindings every primitive is reflecting with zero Neumann data and there is no Dirichlet geometry (scene.h:147-181),
// so a cache point carries only its estimated solution: the splat is  u(x) = mean_y( -P(x, y) u(y) / pdf ) + mean_z( G(x, z) f(z) / pdf ).
// Deviation from the reference, which does not change the estimator: the per-primitive sample counts are visited in
// primitive order (the reference iterates an std::unordered_map, boundary_sampler.h:330-340), so the cache points are the
// same in distribution, not draw for draw.  The wall-clock seeds of the samplers come from one pcg32 keyed by opts->seed.
#include "capi_scene.h"

#include <algorithm>
#include <cmath>
#include <map>

using namespace nmc;

namespace {

struct CachePoint { float x, y, nx, ny, value, aux, pdf, kind; }; // kind 0: boundary (value = solution, aux = normal derivative), 1: source (value = f)

// free-space Green's functions in 2D (distributions.h:85-119 harmonic, :168-219 Yukawa): G and the Poisson kernel dG/dn_y
__device__ __forceinline__ void freeSpace2D(float lambda, float sqrtLambda, float dx, float dy, float r, float nx, float ny, float& G, float& P) {
	const float twoPi = 6.28318530717958647692f;
	const float ndot = nx*dx + ny*dy; // n . (x - y)
	if (lambda > 0.0f) {
		const float mur = r*sqrtLambda;
		G = (float)(bessk0((double)mur)/(2.0*3.14159265358979323846));
		const float Qr = sqrtLambda*(float)bessk1((double)mur);
		P = ndot*Qr/(twoPi*r);
	} else {
		G = -logf(r)/twoPi;
		P = ndot/(twoPi*r*r);
	}
}

// One evaluation point per thread; cache records staged 256 at a time.  sums[3 groups: boundary, boundary normal-aligned,
// source] and counts mirror the three SampleStatistics of an EvaluationPoint (splatter.h:293-355): the estimate is the sum
// of the groups' means over the samples that were actually added (non-finite kernels are skipped, splatter.h:218-222).
__global__ void __launch_bounds__(256)
bvcSplatKernel(const float* __restrict__ evalPts, const float* __restrict__ evalDirichletDist, long long nEval,
			   const CachePoint* __restrict__ cache, int nCache, float lambda, float radiusClamp, float regularization,
			   float dirichletDistCutoff, float* __restrict__ out) {
	__shared__ CachePoint tile[256];
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	const bool live = i < nEval;
	float ex = 0.0f, ey = 0.0f;
	bool skip = true;
	if (live) { ex = evalPts[2*i]; ey = evalPts[2*i + 1]; skip = evalDirichletDist[i] < dirichletDistCutoff; }
	const float sqrtLambda = lambda > 0.0f ? sqrtf(lambda) : 0.0f;
	double sum[3] = {0.0, 0.0, 0.0};
	int cnt[3] = {0, 0, 0};
	for (int base = 0; base < nCache; base += 256) {
		__syncthreads();
		if (base + (int)threadIdx.x < nCache) tile[threadIdx.x] = cache[base + threadIdx.x];
		__syncthreads();
		const int m = nCache - base < 256 ? nCache - base : 256;
		if (skip) continue;
		for (int k = 0; k < m; k++) {
			const CachePoint c = tile[k];
			const float dx = ex - c.x, dy = ey - c.y;
			float r = fmaxf(radiusClamp, sqrtf(dx*dx + dy*dy));
			const bool aligned = c.kind == 2.0f;
			const float nx = aligned ? -c.nx : c.nx, ny = aligned ? -c.ny : c.ny;
			float G, P;
			freeSpace2D(lambda, sqrtLambda, dx, dy, r, nx, ny, G, P);
			if (c.kind == 1.0f) { // source sample (splatter.h:258-290)
				if (!isfinite(G) || !isfinite(1.0f/r)) continue;
				// regularisation of the 2D Green's function is the identity (splatter.h:14-17)
				sum[2] += (double)(G*c.value/c.pdf); cnt[2]++;
			} else { // boundary sample (splatter.h:203-255); alpha = 1: evaluation points are in the domain
				if (!isfinite(G) || !isfinite(P) || !isfinite(1.0f/(r*r))) continue;
				if (regularization > 0.0f) { const float rr = r/regularization; P *= 1.0f - expf(-rr*rr); }
				const int g = aligned ? 1 : 0;
				sum[g] += (double)((G*c.aux - P*c.value)/c.pdf); cnt[g]++;
			}
		}
	}
	if (live && !skip) {
		float v = 0.0f;
		for (int g = 0; g < 3; g++) if (cnt[g] > 0) v += (float)(sum[g]/cnt[g]);
		out[i] = v;
	}
}

// generateStratifiedSamples<D> (include/zombie/core/sampling.h:434-457) on the host
template <int D>
void stratifiedSamples(std::vector<float>& samples, int n, Pcg32& rng) {
	const float oneMinusEps = 1.0f - kEps;
	const float inv = 1.0f/n;
	samples.resize((size_t)D*n);
	for (int i = 0; i < n; i++) for (int j = 0; j < D; j++) samples[D*i + j] = std::min((i + rng.nextFloat())*inv, oneMinusEps);
	for (int i = 0; i < D; i++) for (int j = 0; j < n; j++) {
		int other = j + (int)rng.nextBounded((uint32_t)(n - j));
		std::swap(samples[D*j + i], samples[D*other + i]);
	}
}

// CDFTable (sampling.h:261-319)
struct CdfTable {
	std::vector<float> table;
	float build(const std::vector<float>& w) {
		const int n = (int)w.size();
		if (n == 0) return 0.0f;
		table.assign(n + 1, 0.0f);
		for (int i = 1; i <= n; i++) table[i] = table[i - 1] + w[i - 1];
		const float total = table[n];
		if (total == 0.0f) for (int i = 1; i <= n; i++) table[i] = (float)i/(float)n;
		else for (int i = 1; i <= n; i++) table[i] /= total;
		return total;
	}
	int sample(float u) const {
		int size = (int)table.size(), first = 0, len = size;
		while (len > 0) {
			int half = len >> 1, middle = first + half;
			if (table[middle] <= u) { first = middle + 1; len -= half + 1; } else len = half;
		}
		return std::min(std::max(first - 1, 0), size - 2);
	}
};

struct BoundarySample { float x, y, nx, ny, pdf; int aligned; };

#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return nmcFail(NMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

} // namespace

extern "C" int nmc_estimate_solution(nmc_scene* s, const nmc_solver_opts* opts, const float* pts, const float* normals,
									 const int* types, const int* aligned, int64_t n, int n_walks, uint64_t index_offset,
									 float* solution_out, float* stats4_out) {
	if (!s) return nmcFail(NMC_ERR_INVALID, "null scene");
	SolverParams p;
	int rc = nmcToParams(opts, p);
	if (rc != NMC_OK) return rc;
	if (n < 0 || n_walks < 0 || (n > 0 && (!pts || !solution_out))) return nmcFail(NMC_ERR_INVALID, "nmc_estimate_solution: bad arguments");
	if (types) for (int64_t i = 0; i < n; i++) if (types[i] != 0 && types[i] != 2)
		return nmcFail(NMC_ERR_UNSUPPORTED, "nmc_estimate_solution: sample type must be 0 (in the domain) or 2 (on the reflecting boundary); there is no Dirichlet geometry");
	if (n == 0) return NMC_OK;
	Lock lock(s->mu);
	CKB(cudaSetDevice(s->device));
	const int dim = s->flat.dim;
	const size_t nn = (size_t)n;
	float* d = nullptr;
	CKB(cudaMalloc((void**)&d, nn*(size_t)(2*dim + 2 + 1 + 4)*sizeof(float)));
	float* d_pts = d; float* d_nrm = d_pts + nn*dim; int* d_ty = reinterpret_cast<int*>(d_nrm + nn*dim); int* d_al = d_ty + nn;
	float* d_sol = reinterpret_cast<float*>(d_al + nn); float* d_st = d_sol + nn;
	cudaError_t e = cudaMemcpy(d_pts, pts, nn*dim*4, cudaMemcpyHostToDevice);
	if (!e && normals) e = cudaMemcpy(d_nrm, normals, nn*dim*4, cudaMemcpyHostToDevice);
	if (!e && types) e = cudaMemcpy(d_ty, types, nn*4, cudaMemcpyHostToDevice);
	if (!e && aligned) e = cudaMemcpy(d_al, aligned, nn*4, cudaMemcpyHostToDevice);
	if (!e) e = launchSolutionEstimator(s->view, p, d_pts, normals ? d_nrm : nullptr, types ? d_ty : nullptr, aligned ? d_al : nullptr,
										n, n_walks, index_offset, d_sol, stats4_out ? d_st : nullptr, nullptr, 0);
	if (!e) e = cudaMemcpy(solution_out, d_sol, nn*4, cudaMemcpyDeviceToHost);
	if (!e && stats4_out) e = cudaMemcpy(stats4_out, d_st, nn*16, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e) return nmcFail(NMC_ERR_CUDA, std::string("nmc_estimate_solution: ") + cudaGetErrorString(e));
	return NMC_OK;
}

extern "C" int nmc_bvc_splat(int dim, float absorption, const float* eval_pts, const float* eval_dirichlet_dist, int64_t n_eval,
							 const float* cache8, int n_cache, float radius_clamp, float regularization, float dirichlet_dist_cutoff,
							 float* out) {
	if (dim != 2) return nmcFail(NMC_ERR_UNSUPPORTED, "nmc_bvc_splat: boundary value caching is 2D only (as the reference's bindings)");
	if (n_eval < 0 || n_cache < 0 || (n_eval > 0 && (!eval_pts || !eval_dirichlet_dist || !out)) || (n_cache > 0 && !cache8)) return nmcFail(NMC_ERR_INVALID, "nmc_bvc_splat: bad arguments");
	if (n_eval == 0) return NMC_OK;
	const size_t ne = (size_t)n_eval;
	float* d = nullptr;
	CKB(cudaMalloc((void**)&d, (ne*4 + (size_t)n_cache*8 + 8)*sizeof(float)));
	float* d_pts = d; float* d_dd = d_pts + ne*2; float* d_out = d_dd + ne; CachePoint* d_c = reinterpret_cast<CachePoint*>(d_out + ne);
	cudaError_t e = cudaMemcpy(d_pts, eval_pts, ne*8, cudaMemcpyHostToDevice);
	if (!e) e = cudaMemcpy(d_dd, eval_dirichlet_dist, ne*4, cudaMemcpyHostToDevice);
	if (!e) e = cudaMemcpy(d_out, out, ne*4, cudaMemcpyHostToDevice); // entries below the cut-off keep the caller's value
	if (!e && n_cache) e = cudaMemcpy(d_c, cache8, (size_t)n_cache*32, cudaMemcpyHostToDevice);
	if (!e) {
		bvcSplatKernel<<<(unsigned)((ne + 255)/256), 256>>>(d_pts, d_dd, n_eval, d_c, n_cache, absorption, radius_clamp, regularization, dirichlet_dist_cutoff, d_out);
		e = cudaGetLastError();
	}
	if (!e) e = cudaMemcpy(out, d_out, ne*4, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e) return nmcFail(NMC_ERR_CUDA, std::string("nmc_bvc_splat: ") + cudaGetErrorString(e));
	return NMC_OK;
}

extern "C" int nmc_bvc_solve(nmc_scene* s, const nmc_solver_opts* opts, const nmc_bvc_opts* b, float* grid_out,
							 float* cache_out, int cache_cap, int* n_cache_out, int* n_domain_out) {
	if (!s || !opts || !b || !grid_out) return nmcFail(NMC_ERR_INVALID, "nmc_bvc_solve: null argument");
	if (s->flat.dim != 2) return nmcFail(NMC_ERR_UNSUPPORTED, "nmc_bvc_solve: boundary value caching is 2D only (as the reference's bindings)");
	if (b->gridRes <= 0 || b->boundaryCacheSize < 0 || b->domainCacheSize < 0) return nmcFail(NMC_ERR_INVALID, "nmc_bvc_solve: bad sizes");
	const SceneView& V = s->view;
	const bool doubleSided = V.doubleSided;
	const int nP = (int)s->prims.size()/2;
	const float* vx = s->verts.data();
	const int* pr = s->prims.data();
	const float lo[2] = {V.bboxLo[0], V.bboxLo[1]}, ext[2] = {V.bboxHi[0] - V.bboxLo[0], V.bboxHi[1] - V.bboxLo[1]};
	Pcg32 master; master.seed(pointSeed(opts->seed, 0x6276635F6D617374ull), 1);
	auto clockSeed = [&]() { return (uint64_t)master.nextUInt(); }; // stands in for the reference's system_clock reads
	auto inBox = [&](float x, float y) { return !outsideBox<2>(V, mk(x, y, 0.0f)); };

	// ---- evaluation grid (createEvaluationGrid, demo/grid.h:352-368) and its scene data -----------------------------------
	const int res = b->gridRes;
	const size_t nEval = (size_t)res*res;
	std::vector<float> evalPts(2*nEval), evalD(nEval), evalN(nEval), evalIn(nEval);
	for (int i = 0; i < res; i++) for (int j = 0; j < res; j++) {
		evalPts[2*((size_t)i*res + j)] = (i/float(res))*ext[0] + lo[0];
		evalPts[2*((size_t)i*res + j) + 1] = (j/float(res))*ext[1] + lo[1];
	}
	int rc;
	{
		Lock lock(s->mu);
		rc = nmcProbeUnlocked(s, NMC_PROBE_DIST_DIRICHLET, (int64_t)nEval, evalPts.data(), nullptr, nullptr, nullptr, nullptr, nullptr, evalD.data());
		if (rc == NMC_OK) rc = nmcProbeUnlocked(s, NMC_PROBE_DIST_NEUMANN, (int64_t)nEval, evalPts.data(), nullptr, nullptr, nullptr, nullptr, nullptr, evalN.data());
		if (rc == NMC_OK) rc = nmcProbeUnlocked(s, NMC_PROBE_INSIDE_DOMAIN, (int64_t)nEval, evalPts.data(), nullptr, nullptr, nullptr, nullptr, nullptr, evalIn.data());
	}
	if (rc != NMC_OK) return rc;

	// ---- boundary cache (BoundarySampler, boundary_sampler.h:107-145, 270-403) ----------------------------------------------
	Pcg32 brng; brng.seed(clockSeed(), 1);
	Pcg32 drng; drng.seed(clockSeed(), 1);
	std::vector<BoundarySample> bsamples;
	auto buildAndSample = [&](float normalOffset, int nSamples, bool alignedFlag, float& totalArea, bool sampleNow) {
		std::vector<float> w(nP, 0.0f);
		for (int i = 0; i < nP; i++) {
			const float ax = vx[2*pr[2*i]], ay = vx[2*pr[2*i] + 1], bx = vx[2*pr[2*i + 1]], by = vx[2*pr[2*i + 1] + 1];
			const float sx = bx - ax, sy = by - ay;
			float nx = sy, ny = -sx;
			const float nn = std::sqrt(nx*nx + ny*ny);
			const float ux = nx/nn, uy = ny/nn;
			const float mx = (ax + bx)/2.0f, my = (ay + by)/2.0f;
			if (inBox(mx + normalOffset*ux, my + normalOffset*uy)) w[i] = nn; // every primitive is reflecting: no displacement (:288-293)
		}
		CdfTable table;
		totalArea = table.build(w);
		if (!sampleNow) return;
		if (!(totalArea > 0.0f) || nSamples <= 0) return;
		const float pdf = 1.0f/totalArea;
		std::vector<float> strat;
		stratifiedSamples<1>(strat, nSamples, brng);
		std::map<int, int> count; // primitive -> number of samples (the reference: unordered_map, iteration order unspecified)
		for (int i = 0; i < nSamples; i++) count[table.sample(strat[i])]++;
		for (auto& kv : count) {
			std::vector<float> u;
			if (kv.second == 1) u.push_back(brng.nextFloat());
			else stratifiedSamples<1>(u, kv.second, brng);
			const int i = kv.first;
			const float ax = vx[2*pr[2*i]], ay = vx[2*pr[2*i] + 1], bx = vx[2*pr[2*i + 1]], by = vx[2*pr[2*i + 1] + 1];
			const float sx = bx - ax, sy = by - ay;
			float nx = sy, ny = -sx;
			const float nn = std::sqrt(nx*nx + ny*ny);
			nx /= nn; ny /= nn;
			for (int k = 0; k < kv.second; k++) { // sampleLineSegmentUniformly (sampling.h:213-224)
				BoundarySample bs; bs.x = ax + u[k]*sx; bs.y = ay + u[k]*sy; bs.nx = nx; bs.ny = ny; bs.pdf = pdf; bs.aligned = alignedFlag ? 1 : 0;
				bsamples.push_back(bs);
			}
		}
	};
	float area = 0.0f, areaAligned = 0.0f;
	const float off = b->normalOffsetForCachedDirichletSamples;
	if (doubleSided) { // :121-135: split the sample count by the two tables' areas
		buildAndSample(-off, 0, false, area, false);
		buildAndSample(off, 0, true, areaAligned, false);
		const float total = area + areaAligned;
		const int n0 = total > 0.0f ? (int)std::ceil(b->boundaryCacheSize*area/total) : 0;
		const int n1 = total > 0.0f ? (int)std::ceil(b->boundaryCacheSize*areaAligned/total) : 0;
		buildAndSample(-off, n0, false, area, true);
		buildAndSample(off, n1, true, areaAligned, true);
	} else buildAndSample(-off, b->boundaryCacheSize, false, area, true);
	const int nB = (int)bsamples.size();

	// ---- domain cache (DomainSampler::generateSamples, domain_sampler.h:33-59) -----------------------------------------------
	std::vector<float> dpts, dsrc;
	float domainPdf = 0.0f;
	if (!opts->ignoreSource && b->domainCacheSize > 0) {
		float volume = ext[0]*ext[1];
		if (!doubleSided) { // Scene::getSolveRegionVolume (scene.h:92-100): |sum of the primitives' signed areas|
			float v = 0.0f;
			for (int i = 0; i < nP; i++) v += 0.5f*(vx[2*pr[2*i]]*vx[2*pr[2*i + 1] + 1] - vx[2*pr[2*i] + 1]*vx[2*pr[2*i + 1]]);
			volume = std::fabs(v);
		}
		domainPdf = 1.0f/volume;
		int nStrat = b->domainCacheSize;
		if (volume > 0.0f) nStrat = (int)(nStrat*(ext[0]*ext[1]*domainPdf));
		std::vector<float> strat;
		if (nStrat > 0) stratifiedSamples<2>(strat, nStrat, drng);
		std::vector<float> cand((size_t)2*std::max(nStrat, 0)), inside(std::max(nStrat, 0)), src(std::max(nStrat, 0));
		for (int i = 0; i < nStrat; i++) { cand[2*i] = lo[0] + ext[0]*strat[2*i]; cand[2*i + 1] = lo[1] + ext[1]*strat[2*i + 1]; }
		if (nStrat > 0) {
			Lock lock(s->mu);
			rc = nmcProbeUnlocked(s, NMC_PROBE_INSIDE_DOMAIN, nStrat, cand.data(), nullptr, nullptr, nullptr, nullptr, nullptr, inside.data());
			if (rc == NMC_OK) rc = nmcProbeUnlocked(s, NMC_PROBE_SOURCE, nStrat, cand.data(), nullptr, nullptr, nullptr, nullptr, nullptr, src.data());
			if (rc != NMC_OK) return rc;
		}
		for (int i = 0; i < nStrat; i++) {
			const bool in = doubleSided ? inBox(cand[2*i], cand[2*i + 1]) : inside[i] != 0.0f;
			if (!in) continue;
			dpts.push_back(cand[2*i]); dpts.push_back(cand[2*i + 1]); dsrc.push_back(src[i]);
		}
	}
	const int nD = (int)dsrc.size();

	// ---- solution at the boundary cache points: walks that start on the reflecting boundary --------------------------------
	std::vector<float> bsol(nB, 0.0f);
	if (nB > 0) {
		std::vector<float> p(2*(size_t)nB), nr(2*(size_t)nB); std::vector<int> ty(nB, 2), al(nB, 0);
		for (int i = 0; i < nB; i++) { p[2*i] = bsamples[i].x; p[2*i + 1] = bsamples[i].y; nr[2*i] = bsamples[i].nx; nr[2*i + 1] = bsamples[i].ny; al[i] = bsamples[i].aligned; }
		nmc_solver_opts o2 = *opts;
		o2.seed = clockSeed() | (clockSeed() << 32); // the per-point streams: pointSeed(o2.seed, cache index)
		rc = nmc_estimate_solution(s, &o2, p.data(), nr.data(), ty.data(), al.data(), nB, b->nWalksForCachedSolutionEstimates, 0, bsol.data(), nullptr);
		if (rc != NMC_OK) return rc;
	}

	// ---- splat (splatter.h) ----------------------------------------------------------------------------------------------
	std::vector<float> cache8((size_t)8*(nB + nD));
	for (int i = 0; i < nB; i++) {
		float* c = &cache8[(size_t)8*i];
		c[0] = bsamples[i].x; c[1] = bsamples[i].y; c[2] = bsamples[i].nx; c[3] = bsamples[i].ny;
		c[4] = bsol[i]; c[5] = 0.0f; /* pde.neumann == 0 (scene.h:176-181) */ c[6] = bsamples[i].pdf; c[7] = bsamples[i].aligned ? 2.0f : 0.0f;
	}
	for (int i = 0; i < nD; i++) {
		float* c = &cache8[(size_t)8*(nB + i)];
		c[0] = dpts[2*i]; c[1] = dpts[2*i + 1]; c[2] = c[3] = 0.0f; c[4] = dsrc[i]; c[5] = 0.0f; c[6] = domainPdf; c[7] = 1.0f;
	}
	std::vector<float> value(nEval, 0.0f);
	{
		Lock lock(s->mu);
		CKB(cudaSetDevice(s->device));
		rc = nmc_bvc_splat(2, V.absorption, evalPts.data(), evalD.data(), (int64_t)nEval, cache8.data(), nB + nD, b->radiusClampForKernels,
						   b->regularizationForKernels, off, value.data());
	}
	if (rc != NMC_OK) return rc;

	// ---- pointwise estimates next to the absorbing boundary (splatter.h:160-180); cannot trigger without Dirichlet geometry
	// unless the bounding box is tiny, kept for faithfulness
	{
		std::vector<int64_t> near;
		for (size_t i = 0; i < nEval; i++) if (evalD[i] < off) near.push_back((int64_t)i);
		if (!near.empty()) {
			std::vector<float> p(2*near.size()), sol(near.size());
			for (size_t k = 0; k < near.size(); k++) { p[2*k] = evalPts[2*near[k]]; p[2*k + 1] = evalPts[2*near[k] + 1]; }
			nmc_solver_opts o2 = *opts;
			o2.seed = clockSeed() | (clockSeed() << 32);
			rc = nmc_estimate_solution(s, &o2, p.data(), nullptr, nullptr, nullptr, (int64_t)near.size(), b->nWalksForCachedSolutionEstimates, 0, sol.data(), nullptr);
			if (rc != NMC_OK) return rc;
			for (size_t k = 0; k < near.size(); k++) value[near[k]] = sol[k];
		}
	}

	// ---- mask (saveEvaluationGrid, demo/grid.h:388-411) --------------------------------------------------------------------
	for (size_t i = 0; i < nEval; i++) {
		const bool maskOut = (evalIn[i] == 0.0f && !doubleSided) || std::min(std::fabs(evalD[i]), std::fabs(evalN[i])) < opts->boundaryDistanceMask;
		grid_out[i] = maskOut ? 0.0f : value[i];
	}
	if (cache_out) for (int i = 0; i < nB && i < cache_cap; i++) {
		float* c = cache_out + (size_t)6*i;
		c[0] = bsamples[i].x; c[1] = bsamples[i].y; c[2] = bsamples[i].nx; c[3] = bsamples[i].ny; c[4] = bsol[i]; c[5] = bsamples[i].pdf;
	}
	if (n_cache_out) *n_cache_out = nB;
	if (n_domain_out) *n_domain_out = nD;
	return NMC_OK;
}
