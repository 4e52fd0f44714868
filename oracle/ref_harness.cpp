// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI driver around the UNMODIFIED reference solver headers, compiled from where they lie
// under /root/reference (recipe: oracle/Makefile, outputs only into oracle/_ref/).
// It plays the role of bindings/zombie{,3d}/demo/demo.cpp (runWalkOnStars_sampled :119-205 /
// runWalkOnStars_3d :15-116) without pybind11, and adds two things the reference lacks:
//
//  1. determinism.  The reference seeds every SamplePoint's pcg32 and every antithetic walk
//     pair from std::chrono::system_clock (walk_on_stars.h:498,639).  We do not patch or copy
//     that header; instead the token `system_clock` is re-pointed, for the duration of the
//     #include of walk_on_stars.h only, at a clock whose now().time_since_epoch().count()
//     returns the next 32-bit draw of the CURRENT point's own pcg32 stream.  Together with
//     per-point seeding  sampler.seed(nmc_point_seed(seed, global_index), 1)  this is the
//     deterministic rule of SURVEY.md section 8c / Appendix D.
//  2. probes: scalar entry points into Bessel / Green's functions / samplers / geometric
//     queries so the C restatement (oracle/nmc_oracle.c) and the CUDA path can be pinned
//     function by function.
//
// Parallelism: std::thread over points with an atomic work counter (the reference uses
// tbb::parallel_for over the same loop, walk_on_stars.h:91-103; the per-point solve() we call
// is the reference's own public overload, walk_on_stars.h:59-72).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "pcg32.h"
#include "tbb/parallel_for.h"
#include "tbb/blocked_range.h"
#include <zombie/core/pde.h>
#include <zombie/core/geometric_queries.h>
#include <zombie/core/distributions.h>
#include <zombie/utils/fcpw_scene_loader.h>

namespace nmc_shim {
thread_local pcg32* current_sampler = nullptr;
struct epoch_t { unsigned long long v; unsigned long long count() const { return v; } };
struct tp_t { unsigned long long v; epoch_t time_since_epoch() const { return epoch_t{v}; } };
struct det_clock {
	static tp_t now() { return tp_t{ current_sampler ? (unsigned long long)current_sampler->nextUInt() : 0ull }; }
};
}
namespace std { namespace chrono { using nmc_det_clock = ::nmc_shim::det_clock; } }

#define system_clock nmc_det_clock
#include <zombie/point_estimation/walk_on_stars.h>
#undef system_clock
#include <zombie/boundary_value_caching/splatter.h>

#include "grid.h"
#if REF_DIM == 2
#include "scene.h"
#else
#include "scene_3d.h"
#endif

static const int DIM = REF_DIM;
using VecD = zombie::Vector<REF_DIM>;

static inline uint64_t nmc_point_seed(uint64_t seed, uint64_t index) {
	// splitmix64 finaliser of (seed + golden*(index+1)); the same rule is restated in
	// oracle/nmc_oracle.c and in the CUDA library (include/nmcfs.h documents it).
	uint64_t z = seed + 0x9E3779B97F4A7C15ull*(index + 1ull);
	z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27))*0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

struct RefScene {
	Scene *scene;
};

template <typename F>
static void parallel_points(int n, int nthreads, F&& fn) {
	if (nthreads <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
	std::atomic<int> next(0);
	const int chunk = 16;
	std::vector<std::thread> pool;
	for (int t = 0; t < nthreads; t++) {
		pool.emplace_back([&]() {
			for (;;) {
				int b = next.fetch_add(chunk);
				if (b >= n) break;
				int e = std::min(n, b + chunk);
				for (int i = b; i < e; i++) fn(i);
			}
		});
	}
	for (auto& th : pool) th.join();
}

template <typename G>
static void greens_probe(G& g, float R, float r, float* o) {
	VecD c = VecD::Zero();
	g.updateBall(c, R);
	g.r = r;
	VecD ex = VecD::Zero(); ex[0] = 1.0f;
	VecD el = VecD::Zero(); el[DIM - 1] = 1.0f;
	g.yVol = c + r*ex;
	g.ySurf = c + R*el;
	o[0] = g.evaluate();
	o[1] = g.norm();
	o[2] = g.gradientNorm();
	o[3] = g.poissonKernel();
	o[4] = g.directionSampledPoissonKernel(g.yVol);
	o[5] = g.poissonKernelGradient()[DIM - 1];
	VecD x = c + 0.25f*R*el;
	o[6] = g.evaluate(x, g.yVol);
	o[7] = g.potential();
	o[8] = g.gradient()[0];
	o[9] = 0.0f;
}
extern "C" {

int ref_dim() { return REF_DIM; }
int ref_uses_enoki() {
#ifdef FCPW_USE_ENOKI
	return 1;
#else
	return 0;
#endif
}

// ---- scene -------------------------------------------------------------------------------
// 2D: src is row-major [n0=h][n1=w] (mat[i][j], i<->y, j<->x, image.h:59-75); n2 ignored.
// 3D: src is [n0][n1][n2] <-> (x,y,z) (scene_3d.h:120-126).
void* ref_scene_create(const char* scene_json, const float* src, int n0, int n1, int n2) {
	json cfg = json::parse(scene_json);
	std::streambuf* old = std::cout.rdbuf(nullptr); // FCPW prints build stats (sbvh.inl:482-493)
#if REF_DIM == 2
	std::vector<std::vector<float>> mat(n0, std::vector<float>(n1));
	for (int i = 0; i < n0; i++) for (int j = 0; j < n1; j++) mat[i][j] = src[(size_t)i*n1 + j];
	Scene *s = new Scene(cfg, mat);
#else
	std::vector<std::vector<std::vector<float>>> mat(n0, std::vector<std::vector<float>>(n1, std::vector<float>(n2)));
	for (int i = 0; i < n0; i++) for (int j = 0; j < n1; j++) for (int k = 0; k < n2; k++)
		mat[i][j][k] = src[((size_t)i*n1 + j)*n2 + k];
	Scene *s = new Scene(cfg, mat);
#endif
	std::cout.rdbuf(old);
	return new RefScene{s};
}

void ref_scene_destroy(void* h) {
	RefScene *rs = (RefScene*)h;
	delete rs->scene;
	delete rs;
}

void ref_scene_bbox(void* h, float* out /*2*DIM*/) {
	Scene *s = ((RefScene*)h)->scene;
	for (int i = 0; i < DIM; i++) { out[i] = s->bbox.pMin[i]; out[DIM + i] = s->bbox.pMax[i]; }
}

// ---- the solve: mirrors demo.cpp runWalkOnStars_sampled / runWalkOnStars_3d ------------------
// stats (optional, may be null): per point 12 floats:
//  [0] unmasked solution mean [1] solution variance [2..4] unmasked gradient mean
//  [5..7] gradient variance [8] mean first-source contribution [9] nSolutionEstimates
//  [10] mean walk length [11] estimation quantity (1 = SolutionAndGradient, 0 = None)
int ref_wost(void* h, const char* solver_json, const char* output_json,
			 const float* pts, int n, uint64_t seed, uint64_t index_offset, int nthreads,
			 float* p_out, float* grad_out, float* stats) {
	Scene& scene = *((RefScene*)h)->scene;
	json solverConfig = json::parse(solver_json);
	json outputConfig = json::parse(output_json);

	const bool disableGradientControlVariates = getOptional<bool>(solverConfig, "disableGradientControlVariates", false);
	const bool disableGradientAntitheticVariates = getOptional<bool>(solverConfig, "disableGradientAntitheticVariates", false);
	const bool useCosineSamplingForDirectionalDerivatives = getOptional<bool>(solverConfig, "useCosineSamplingForDirectionalDerivatives", false);
	const bool ignoreDirichlet = getOptional<bool>(solverConfig, "ignoreDirichlet", false);
	const bool ignoreNeumann = getOptional<bool>(solverConfig, "ignoreNeumann", false);
	const bool ignoreSource = getOptional<bool>(solverConfig, "ignoreSource", false);
	const int nWalks = getOptional<int>(solverConfig, "nWalks", 128);
	const int maxWalkLength = getOptional<int>(solverConfig, "maxWalkLength", 1024);
	const int stepsBeforeApplyingTikhonov = getOptional<int>(solverConfig, "setpsBeforeApplyingTikhonov", maxWalkLength);
	const int stepsBeforeUsingMaximalSpheres = getOptional<int>(solverConfig, "setpsBeforeUsingMaximalSpheres", maxWalkLength);
	const int gridRes = getRequired<int>(outputConfig, "gridRes");
	const float epsilonShell = getOptional<float>(solverConfig, "epsilonShell", 1e-3f);
	const float minStarRadius = getOptional<float>(solverConfig, "minStarRadius", 1e-3f);
	const float silhouettePrecision = getOptional<float>(solverConfig, "silhouettePrecision", 1e-3f);
	const float russianRouletteThreshold = getOptional<float>(solverConfig, "russianRouletteThreshold", 0.0f);

	const zombie::GeometricQueries<REF_DIM>& queries = scene.queries;
	const zombie::PDE<float, REF_DIM>& pde = scene.pde;
	bool solveDoubleSided = scene.isDoubleSided;

	std::vector<std::vector<float>> sample_points(n, std::vector<float>(DIM));
	for (int i = 0; i < n; i++) for (int k = 0; k < DIM; k++) sample_points[i][k] = pts[(size_t)i*DIM + k];

	std::vector<zombie::SamplePoint<float, REF_DIM>> samplePts;
	samplePts.reserve(n);
#if REF_DIM == 2
	createSolutionGrid(samplePts, queries, scene.bbox.pMin, scene.bbox.pMax, gridRes, sample_points);
#else
	createSolutionGrid_3d(samplePts, queries, scene.bbox.pMin, scene.bbox.pMax, gridRes, sample_points);
#endif

	std::vector<zombie::SampleEstimationData<REF_DIM>> est(n);
	for (int i = 0; i < n; i++) {
		est[i].nWalks = nWalks;
		est[i].estimationQuantity = queries.insideDomain(samplePts[i].pt) || solveDoubleSided ?
									zombie::EstimationQuantity::SolutionAndGradient :
									zombie::EstimationQuantity::None;
	}

	zombie::WalkSettings<float> walkSettings(0.0f, epsilonShell, minStarRadius,
											 silhouettePrecision, russianRouletteThreshold,
											 maxWalkLength, stepsBeforeApplyingTikhonov,
											 stepsBeforeUsingMaximalSpheres, solveDoubleSided,
											 !disableGradientControlVariates,
											 !disableGradientAntitheticVariates,
											 useCosineSamplingForDirectionalDerivatives,
											 ignoreDirichlet, ignoreNeumann, ignoreSource, false);
	zombie::WalkOnStars<float, REF_DIM> walkOnStars(queries);

	parallel_points(n, nthreads, [&](int i) {
		samplePts[i].sampler.seed(nmc_point_seed(seed, index_offset + (uint64_t)i), 1);
		nmc_shim::current_sampler = &samplePts[i].sampler;
		walkOnStars.solve(pde, walkSettings, est[i], samplePts[i]);
		nmc_shim::current_sampler = nullptr;
	});

#if REF_DIM == 2
	std::vector<float> solution = getSolution(samplePts, pde, queries, solveDoubleSided, outputConfig);
	std::vector<std::vector<float>> gradient = getGradient(samplePts, pde, queries, solveDoubleSided, outputConfig);
#else
	std::vector<float> solution = getSolution_3d(samplePts, pde, queries, solveDoubleSided, outputConfig);
	std::vector<std::vector<float>> gradient = getGradient_3d(samplePts, pde, queries, solveDoubleSided, outputConfig);
#endif
	for (int i = 0; i < n; i++) {
		p_out[i] = solution[i];
		for (int k = 0; k < DIM; k++) grad_out[(size_t)i*DIM + k] = gradient[i][k];
	}
	if (stats) {
		for (int i = 0; i < n; i++) {
			float *s = stats + (size_t)i*12;
			for (int k = 0; k < 12; k++) s[k] = 0.0f;
			s[11] = est[i].estimationQuantity == zombie::EstimationQuantity::None ? 0.0f : 1.0f;
			if (!samplePts[i].statistics) continue;
			auto& st = *samplePts[i].statistics;
			s[0] = st.getEstimatedSolution();
			s[1] = st.getEstimatedSolutionVariance();
			std::vector<float> gv = st.getEstimatedGradientVariance();
			for (int k = 0; k < DIM; k++) { s[2 + k] = st.getEstimatedGradient()[k]; s[5 + k] = gv[k]; }
			s[8] = st.getMeanFirstSourceContribution();
			s[9] = (float)st.getSolutionEstimateCount();
			s[10] = st.getMeanWalkLength();
		}
	}
	return 0;
}

// ---- solution-only estimator (walk_on_stars.h:354-461, EstimationQuantity::Solution) at caller-given points -------------
// types: 0 = InDomain, 2 = OnNeumannBoundary (zombie::SampleType); normals: n x DIM (used by boundary starts);
// aligned: estimateBoundaryNormalAligned per point (may be null).  stats (may be null): per point 4 floats
// [0] variance [1] number of estimates [2] mean walk length [3] firstSphereRadius
static zombie::WalkSettings<float> walk_settings_from(const json& solverConfig, bool solveDoubleSided) {
	const bool disableGradientControlVariates = getOptional<bool>(solverConfig, "disableGradientControlVariates", false);
	const bool disableGradientAntitheticVariates = getOptional<bool>(solverConfig, "disableGradientAntitheticVariates", false);
	const bool useCosineSamplingForDirectionalDerivatives = getOptional<bool>(solverConfig, "useCosineSamplingForDirectionalDerivatives", false);
	const bool ignoreDirichlet = getOptional<bool>(solverConfig, "ignoreDirichlet", false);
	const bool ignoreNeumann = getOptional<bool>(solverConfig, "ignoreNeumann", false);
	const bool ignoreSource = getOptional<bool>(solverConfig, "ignoreSource", false);
	const int maxWalkLength = getOptional<int>(solverConfig, "maxWalkLength", 1024);
	const int stepsBeforeApplyingTikhonov = getOptional<int>(solverConfig, "setpsBeforeApplyingTikhonov", maxWalkLength);
	const int stepsBeforeUsingMaximalSpheres = getOptional<int>(solverConfig, "setpsBeforeUsingMaximalSpheres", maxWalkLength);
	const float epsilonShell = getOptional<float>(solverConfig, "epsilonShell", 1e-3f);
	const float minStarRadius = getOptional<float>(solverConfig, "minStarRadius", 1e-3f);
	const float silhouettePrecision = getOptional<float>(solverConfig, "silhouettePrecision", 1e-3f);
	const float russianRouletteThreshold = getOptional<float>(solverConfig, "russianRouletteThreshold", 0.0f);
	return zombie::WalkSettings<float>(0.0f, epsilonShell, minStarRadius, silhouettePrecision, russianRouletteThreshold,
									   maxWalkLength, stepsBeforeApplyingTikhonov, stepsBeforeUsingMaximalSpheres, solveDoubleSided,
									   !disableGradientControlVariates, !disableGradientAntitheticVariates,
									   useCosineSamplingForDirectionalDerivatives, ignoreDirichlet, ignoreNeumann, ignoreSource, false);
}

int ref_estimate_solution(void* h, const char* solver_json, const float* pts, const float* normals, const int* types,
						  const int* aligned, int n, int nWalks, uint64_t seed, uint64_t index_offset, int nthreads,
						  float* sol_out, float* stats) {
	Scene& scene = *((RefScene*)h)->scene;
	json solverConfig = json::parse(solver_json);
	const zombie::GeometricQueries<REF_DIM>& queries = scene.queries;
	const zombie::PDE<float, REF_DIM>& pde = scene.pde;
	zombie::WalkSettings<float> walkSettings = walk_settings_from(solverConfig, scene.isDoubleSided);
	zombie::WalkOnStars<float, REF_DIM> walkOnStars(queries);
	std::vector<zombie::SamplePoint<float, REF_DIM>> samplePts;
	samplePts.reserve(n);
	for (int i = 0; i < n; i++) {
		VecD pt, nr = VecD::Zero();
		for (int k = 0; k < DIM; k++) { pt[k] = pts[(size_t)i*DIM + k]; if (normals) nr[k] = normals[(size_t)i*DIM + k]; }
		zombie::SampleType ty = types && types[i] == 2 ? zombie::SampleType::OnNeumannBoundary : zombie::SampleType::InDomain;
		float dDist = queries.computeDistToDirichlet(pt, false);
		float nDist = queries.computeDistToNeumann(pt, false);
		samplePts.emplace_back(zombie::SamplePoint<float, REF_DIM>(pt, nr, ty, 1.0f, dDist, nDist, 0.0f));
		if (aligned && aligned[i]) samplePts.back().estimateBoundaryNormalAligned = true;
	}
	zombie::SampleEstimationData<REF_DIM> est(nWalks, zombie::EstimationQuantity::Solution);
	parallel_points(n, nthreads, [&](int i) {
		samplePts[i].sampler.seed(nmc_point_seed(seed, index_offset + (uint64_t)i), 1);
		nmc_shim::current_sampler = &samplePts[i].sampler;
		walkOnStars.solve(pde, walkSettings, est, samplePts[i]);
		nmc_shim::current_sampler = nullptr;
	});
	for (int i = 0; i < n; i++) {
		auto& st = *samplePts[i].statistics;
		sol_out[i] = st.getEstimatedSolution();
		if (stats) {
			stats[(size_t)i*4 + 0] = st.getEstimatedSolutionVariance();
			stats[(size_t)i*4 + 1] = (float)st.getSolutionEstimateCount();
			stats[(size_t)i*4 + 2] = st.getMeanWalkLength();
			stats[(size_t)i*4 + 3] = samplePts[i].firstSphereRadius;
		}
	}
	return 0;
}

#if REF_DIM == 2
// ---- boundary value caching: mirrors demo.cpp runBoundaryValueCaching (:265-363) up to saveEvaluationGrid's masking (grid.h:388-411)
// grid_out: gridRes*gridRes masked solution values, index i*gridRes + j <-> point (i, j) of createEvaluationGrid (grid.h:352-368).
// cache_out (may be null): up to cache_cap boundary samples x 6 floats (x, y, nx, ny, estimated solution, pdf); *n_cache = their number.
// The samplers' wall-clock seeds (boundary_sampler.h:103, domain_sampler.h:26, every SamplePoint) come from one pcg32 seeded with `seed`.
int ref_bvc(void* h, const char* solver_json, const char* output_json, uint64_t seed, int nthreads,
			float* grid_out, float* cache_out, int cache_cap, int* n_cache, int* n_domain) {
	Scene& scene = *((RefScene*)h)->scene;
	json solverConfig = json::parse(solver_json);
	json outputConfig = json::parse(output_json);
	const bool useFiniteDifferencesForBoundaryDerivatives = getOptional<bool>(solverConfig, "useFiniteDifferencesForBoundaryDerivatives", false);
	const bool ignoreSource = getOptional<bool>(solverConfig, "ignoreSource", false);
	const int nWalksForCachedSolutionEstimates = getOptional<int>(solverConfig, "nWalksForCachedSolutionEstimates", 128);
	const int nWalksForCachedGradientEstimates = getOptional<int>(solverConfig, "nWalksForCachedGradientEstimates", 640);
	const int boundaryCacheSize = getOptional<int>(solverConfig, "boundaryCacheSize", 1024);
	const int domainCacheSize = getOptional<int>(solverConfig, "domainCacheSize", 1024);
	const int gridRes = getRequired<int>(outputConfig, "gridRes");
	const float epsilonShell = getOptional<float>(solverConfig, "epsilonShell", 1e-3f);
	const float normalOffsetForCachedDirichletSamples = getOptional<float>(solverConfig, "normalOffsetForCachedDirichletSamples", 5.0f*epsilonShell);
	const float radiusClampForKernels = getOptional<float>(solverConfig, "radiusClampForKernels", 1e-3f);
	const float regularizationForKernels = getOptional<float>(solverConfig, "regularizationForKernels", 0.0f);
	const float boundaryDistanceMask = getOptional<float>(outputConfig, "boundaryDistanceMask", 0.0);

	fcpw::BoundingBox<2> bbox = scene.bbox;
	const zombie::GeometricQueries<2>& queries = scene.queries;
	const zombie::PDE<float, 2>& pde = scene.pde;
	bool solveDoubleSided = scene.isDoubleSided;
	std::function<bool(const Vector2&)> insideSolveRegionBoundarySampler = [&queries](const Vector2& x) -> bool { return !queries.outsideBoundingDomain(x); };
	std::function<bool(const Vector2&)> insideSolveRegionDomainSampler = [&queries, solveDoubleSided](const Vector2& x) -> bool {
		return solveDoubleSided ? !queries.outsideBoundingDomain(x) : queries.insideDomain(x);
	};
	std::function<bool(const Vector2&)> onNeumannBoundary = [&scene](const Vector2 &x) -> bool { return scene.onNeumannBoundary(x); };

	pcg32 master(seed, 1);
	nmc_shim::current_sampler = &master; // every clock read below draws from `master` (single-threaded set-up)
	std::vector<zombie::SamplePoint<float, 2>> boundaryCache, boundaryCacheNormalAligned, domainCache;
	std::vector<zombie::EvaluationPoint<float, 2>> evalPts;
	createEvaluationGrid(evalPts, queries, bbox.pMin, bbox.pMax, gridRes);
	zombie::WalkOnStars<float, 2> walkOnStars(queries);
	zombie::BoundarySampler<float, 2> boundarySampler(scene.vertices, scene.segments, queries, walkOnStars,
													  insideSolveRegionBoundarySampler, onNeumannBoundary);
	zombie::DomainSampler<float, 2> domainSampler(queries, insideSolveRegionDomainSampler, bbox.pMin, bbox.pMax, scene.getSolveRegionVolume());
	boundarySampler.initialize(normalOffsetForCachedDirichletSamples, solveDoubleSided);
	boundarySampler.generateSamples(boundaryCacheSize, normalOffsetForCachedDirichletSamples, solveDoubleSided, 0.0f,
									boundaryCache, boundaryCacheNormalAligned);
	if (!ignoreSource) domainSampler.generateSamples(pde, domainCacheSize, domainCache);
	nmc_shim::current_sampler = nullptr;

	zombie::WalkSettings<float> walkSettings = walk_settings_from(solverConfig, solveDoubleSided);
	// computeEstimates (boundary_sampler.h:148-226) uses tbb::parallel_for inside solve(); the per-point samplers were seeded
	// from `master` at construction, so the estimates do not depend on the thread schedule
	boundarySampler.computeEstimates(pde, walkSettings, nWalksForCachedSolutionEstimates, nWalksForCachedGradientEstimates,
									 boundaryCache, useFiniteDifferencesForBoundaryDerivatives, nthreads <= 1);
	boundarySampler.computeEstimates(pde, walkSettings, nWalksForCachedSolutionEstimates, nWalksForCachedGradientEstimates,
									 boundaryCacheNormalAligned, useFiniteDifferencesForBoundaryDerivatives, nthreads <= 1);
	zombie::Splatter<float, 2> splatter(queries, walkOnStars);
	splatter.splat(pde, boundaryCache, radiusClampForKernels, regularizationForKernels, normalOffsetForCachedDirichletSamples, evalPts);
	splatter.splat(pde, boundaryCacheNormalAligned, radiusClampForKernels, regularizationForKernels, normalOffsetForCachedDirichletSamples, evalPts);
	splatter.splat(pde, domainCache, radiusClampForKernels, regularizationForKernels, normalOffsetForCachedDirichletSamples, evalPts);
	splatter.estimatePointwiseNearDirichletBoundary(pde, walkSettings, normalOffsetForCachedDirichletSamples,
													nWalksForCachedSolutionEstimates, evalPts, nthreads <= 1);
	for (int i = 0; i < gridRes; i++) for (int j = 0; j < gridRes; j++) {
		int idx = i*gridRes + j;
		float inDomain = queries.insideDomain(evalPts[idx].pt) ? 1 : 0;
		float value = evalPts[idx].getEstimatedSolution();
		bool maskOutValue = (!inDomain && !solveDoubleSided) ||
							std::min(std::abs(evalPts[idx].dirichletDist), std::abs(evalPts[idx].neumannDist)) < boundaryDistanceMask;
		grid_out[idx] = maskOutValue ? 0.0f : value;
	}
	int nb = 0;
	for (auto* cache : {&boundaryCache, &boundaryCacheNormalAligned}) for (auto& sp : *cache) {
		if (cache_out && nb < cache_cap) {
			float* c = cache_out + (size_t)nb*6;
			c[0] = sp.pt[0]; c[1] = sp.pt[1]; c[2] = sp.normal[0]; c[3] = sp.normal[1]; c[4] = sp.solution; c[5] = sp.pdf;
		}
		nb++;
	}
	if (n_cache) *n_cache = nb;
	if (n_domain) *n_domain = (int)domainCache.size();
	return 0;
}
#endif

// ---- probes: RNG (deps/pcg32/pcg32.h:40-112) ------------------------------------------------
void ref_pcg32_uint(uint64_t initstate, uint64_t initseq, int n, uint32_t* out) {
	pcg32 s(initstate, initseq);
	for (int i = 0; i < n; i++) out[i] = s.nextUInt();
}
void ref_pcg32_float(uint64_t initstate, uint64_t initseq, int n, float* out) {
	pcg32 s(initstate, initseq);
	for (int i = 0; i < n; i++) out[i] = s.nextFloat();
}
void ref_pcg32_bounded(uint64_t initstate, uint64_t initseq, const uint32_t* bounds, int n, uint32_t* out) {
	pcg32 s(initstate, initseq);
	for (int i = 0; i < n; i++) out[i] = s.nextUInt(bounds[i]);
}
uint64_t ref_point_seed(uint64_t seed, uint64_t index) { return nmc_point_seed(seed, index); }

// generateStratifiedSamples<DIM-1> (sampling.h:434-457); out has (DIM-1)*nSamples floats;
// state_out = {state, inc} after the call
void ref_stratified(uint64_t initstate, int nSamples, float* out, uint64_t* state_out) {
	pcg32 s(initstate, 1);
	std::vector<float> v;
	zombie::generateStratifiedSamples<REF_DIM - 1>(v, nSamples, s);
	std::copy(v.begin(), v.end(), out);
	state_out[0] = s.state; state_out[1] = s.inc;
}

// sampleUnitSphereUniform<DIM>(float* u) (sampling.h:29-45): n x (DIM-1) in, n x DIM out
void ref_sphere_dir(const float* u, int n, float* out) {
	for (int i = 0; i < n; i++) {
		float uu[2] = {u[(size_t)i*(DIM - 1)], DIM == 3 ? u[(size_t)i*(DIM - 1) + 1] : 0.0f};
		VecD d = zombie::sampleUnitSphereUniform<REF_DIM>(uu);
		for (int k = 0; k < DIM; k++) out[(size_t)i*DIM + k] = d[k];
	}
}

// ---- probes: Bessel (deps/bessel/bessel.hpp:373-556) ----------------------------------------
// kind: 0 = i0, 1 = i1, 2 = k0, 3 = k1, 4 = k2 (bessk(2, x)).  double in/out.
void ref_bessel(int kind, const double* x, int n, double* out) {
	for (int i = 0; i < n; i++) {
		switch (kind) {
			case 0: out[i] = bessel::bessi0(x[i]); break;
			case 1: out[i] = bessel::bessi1(x[i]); break;
			case 2: out[i] = bessel::bessk0(x[i]); break;
			case 3: out[i] = bessel::bessk1(x[i]); break;
			default: out[i] = bessel::bessk(2, x[i]); break;
		}
	}
}

// ---- probes: ball Green's functions (distributions.h:396-832) -------------------------------
// For each i: ball centred at the origin with radius R[i], sampled radius r[i] along +x,
// surface point along +y (2D) / +z (3D).  lambda > 0 -> Yukawa, lambda == 0 -> harmonic.
// out per i (10 floats): evaluate(), norm(), gradientNorm(), poissonKernel(),
//  directionSampledPoissonKernel(y = c + r*e_x), |poissonKernelGradient()| component along the
//  surface direction, evaluate(x = c + 0.25R e_y/z, y = c + r e_x), potential(), 0, 0
void ref_greens_ball(float lambda, const float* R, const float* r, int n, float* out) {
	for (int i = 0; i < n; i++) {
		if (lambda > 0.0f) { zombie::YukawaGreensFnBall<REF_DIM> g(lambda); greens_probe(g, R[i], r[i], out + (size_t)i*10); }
		else { zombie::HarmonicGreensFnBall<REF_DIM> g; greens_probe(g, R[i], r[i], out + (size_t)i*10); }
	}
}

// sampleVolume(dir = e_x, sampler, pdf): for each i a fresh pcg32(seeds[i]); outputs r, pdf and
// the number of nextFloat draws consumed (counted by replaying the stream).
void ref_sample_volume(float lambda, const float* R, const uint64_t* seeds, int n,
					   float* r_out, float* pdf_out, int* draws_out) {
	VecD ex = VecD::Zero(); ex[0] = 1.0f;
	for (int i = 0; i < n; i++) {
		pcg32 s(seeds[i], 1);
		float pdf = 0.0f, r = 0.0f;
		if (lambda > 0.0f) {
			zombie::YukawaGreensFnBall<REF_DIM> g(lambda);
			g.updateBall(VecD::Zero(), R[i]);
			g.sampleVolume(ex, s, pdf); r = g.r;
		} else {
			zombie::HarmonicGreensFnBall<REF_DIM> g;
			g.updateBall(VecD::Zero(), R[i]);
			g.sampleVolume(ex, s, pdf); r = g.r;
		}
		r_out[i] = r; pdf_out[i] = pdf;
		pcg32 t(seeds[i], 1);
		int k = 0;
		while (!(t.state == s.state) && k < 4096) { t.nextUInt(); k++; }
		draws_out[i] = k;
	}
}

// ---- probes: geometric queries (fcpw_scene_loader.h:292-652) --------------------------------
void ref_dist_neumann(void* h, const float* pts, int n, int signed_, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.queries.computeDistToNeumann(x, signed_ != 0);
	}
}
void ref_dist_dirichlet(void* h, const float* pts, int n, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.queries.computeDistToDirichlet(x, false);
	}
}
void ref_inside_domain(void* h, const float* pts, int n, int* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.queries.insideDomain(x) ? 1 : 0;
	}
}
void ref_outside_bbox(void* h, const float* pts, int n, int* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.queries.outsideBoundingDomain(x) ? 1 : 0;
	}
}
void ref_star_radius(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flip, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.queries.computeStarRadius(x, minR, maxR[i], prec, flip != 0);
	}
}
// out per ray: hit(0/1), dist, pt[DIM], normal[DIM]  -> 2 + 2*DIM floats
void ref_intersect_neumann(void* h, const float* org, const float* nrm, const float* dir, const float* tmax,
						   const int* onb, int n, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	const int W = 2 + 2*DIM;
	for (int i = 0; i < n; i++) {
		VecD o, nn, d;
		for (int k = 0; k < DIM; k++) { o[k] = org[(size_t)i*DIM + k]; nn[k] = nrm[(size_t)i*DIM + k]; d[k] = dir[(size_t)i*DIM + k]; }
		zombie::IntersectionPoint<REF_DIM> ip;
		bool hit = s.queries.intersectWithNeumann(o, nn, d, tmax[i], onb[i] != 0, ip);
		float *r = out + (size_t)i*W;
		r[0] = hit ? 1.0f : 0.0f; r[1] = ip.dist;
		for (int k = 0; k < DIM; k++) { r[2 + k] = ip.pt[k]; r[2 + DIM + k] = ip.normal[k]; }
	}
}
// visibility: returns 1 when the segment xi->xj is BLOCKED (intersectsWithNeumann)
void ref_blocked(void* h, const float* xi, const float* xj, const float* ni, const float* nj,
				 const int* offi, const int* offj, int n, int* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD a, b, na, nb;
		for (int k = 0; k < DIM; k++) { a[k] = xi[(size_t)i*DIM + k]; b[k] = xj[(size_t)i*DIM + k]; na[k] = ni[(size_t)i*DIM + k]; nb[k] = nj[(size_t)i*DIM + k]; }
		out[i] = s.queries.intersectsWithNeumann(a, b, na, nb, offi[i] != 0, offj[i] != 0) ? 1 : 0;
	}
}
// out per query: found(0/1), pdf, pt[DIM], normal[DIM]
void ref_sample_neumann(void* h, const float* pts, const float* radius, const float* rnd /*n x DIM*/, int n, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	const int W = 2 + 2*DIM;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		float rn[3] = {0, 0, 0};
		for (int k = 0; k < DIM; k++) rn[k] = rnd[(size_t)i*DIM + k];
		zombie::BoundarySample<REF_DIM> bs;
		bool found = s.queries.sampleNeumann(x, radius[i], rn, bs);
		float *r = out + (size_t)i*W;
		r[0] = found ? 1.0f : 0.0f; r[1] = bs.pdf;
		for (int k = 0; k < DIM; k++) { r[2 + k] = bs.pt[k]; r[2 + DIM + k] = bs.normal[k]; }
	}
}
void ref_offset_point(const float* p, const float* nrm, int n, float* out) {
	for (int i = 0; i < n; i++) {
		VecD a, b;
		for (int k = 0; k < DIM; k++) { a[k] = p[(size_t)i*DIM + k]; b[k] = nrm[(size_t)i*DIM + k]; }
		VecD r = zombie::offsetPointAlongDirection<REF_DIM>(a, b);
		for (int k = 0; k < DIM; k++) out[(size_t)i*DIM + k] = r[k];
	}
}
// pde.source(x) (scene.h:194-198 / scene_3d.h:120-126)
void ref_source(void* h, const float* pts, int n, float* out) {
	Scene& s = *((RefScene*)h)->scene;
	for (int i = 0; i < n; i++) {
		VecD x; for (int k = 0; k < DIM; k++) x[k] = pts[(size_t)i*DIM + k];
		out[i] = s.pde.source(x);
	}
}

} // extern "C"
