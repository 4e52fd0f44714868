// csrc/nmc_estimator.cuh -- deterministic-mode estimator for ONE sample point: the reflecting
// walk-on-stars loop and the antithetic / control-variate solution-and-gradient estimator.
//
// Replaces zombie::WalkOnStars<float,DIM>::walk (include/zombie/point_estimation/walk_on_stars.h:135-329)
// and ::estimateSolutionAndGradient (:466-617) plus SampleStatistics (:744-877) and the per-point
// set-up / masking of the bindings (demo/grid.h:69-102,155-237, demo/demo.cpp:152-158).
// RNG consumption order is the reference's; the two wall-clock seeds of the
// reference are replaced by the rule in include/nmcfs.h (NMC_MODE_DETERMINISTIC).
#pragma once
#include "nmc_geom.cuh"
#include "nmc_ball.cuh"

namespace nmc {

struct SolverParams {
	int nWalks, maxWalkLength, stepsBeforeApplyingTikhonov, stepsBeforeUsingMaximalSpheres;
	float epsilonShell, minStarRadius, silhouettePrecision, russianRouletteThreshold;
	int useGradientControlVariates, useGradientAntitheticVariates, useCosineSampling;
	int ignoreDirichlet, ignoreNeumann, ignoreSource;
	float boundaryDistanceMask;
	uint64_t seed;
};

struct PointResult {
	float p, g[3];            // masked outputs (getSolution / getGradient)
	float solMean, solM2, gradMean[3], gradM2[3], meanFirstSource;
	int nSol; int totalWalkLength; int active;
	unsigned walksStarted, steps;
};

struct WalkState { // WalkState, walk_on_stars.h:880-913
	V3 pt, normal, prevDir, srcGradDir, bdyGradDir;
	float prevDist, throughput;
	bool onNeumann;
	float totalNeumann, totalSource, firstSource;
	int walkLength;
};

enum WalkCode { kReachedDirichlet = 0, kRussianRoulette = 1, kExceededLength = 2, kEscaped = 3 };

NMC_HD void welford(float est, float& mean, float& M2, int N) { // :863-868
	float delta = est - mean;
	mean += delta/N;
	float delta2 = est - mean;
	M2 += delta*delta2;
}

// walk() :135-329.  firstSphereRadius == 0 on the solution-and-gradient path (:580); estimateSolution passes the radius it
// precomputed for the sample point (:405-425), which replaces the first step's star radius (:148-149).
template <int DIM>
NMC_HD int detWalk(const SceneView& S, const SolverParams& o, float dirichletDist, Pcg32& rng,
				   BallExact<DIM>& g, WalkState& st, unsigned& steps, float firstSphereRadius = 0.0f) {
	typedef ExactMath M;
	bool firstStep = true;
	while (dirichletDist > o.epsilonShell) {
		steps++;
		float starR;
		if (firstStep && firstSphereRadius > 0.0f) starR = firstSphereRadius;
		else {
			bool flipOrient = false;
			if (S.doubleSided && st.onNeumann) { // :154-160
				if (st.prevDist > 0.0f && dot(st.prevDir, st.normal) < 0.0f) { st.normal = st.normal*-1.0f; flipOrient = true; }
			}
			if (o.stepsBeforeUsingMaximalSpheres <= st.walkLength) starR = dirichletDist;
			else {
				starR = starRadius<DIM, M>(S, st.pt, o.minStarRadius, dirichletDist, o.silhouettePrecision, flipOrient);
				if (o.minStarRadius <= dirichletDist) starR = maxS(kShrink*starR, o.minStarRadius);
			}
		}
		firstStep = false;
		g.update(st.pt, starR);

		float u0 = rng.nextFloat();
		float u1 = DIM == 3 ? rng.nextFloat() : 0.0f;
		V3 dir = sphereDir<DIM, M>(u0, u1);
		if (st.onNeumann && dot(st.normal, dir) > 0.0f) dir = dir*-1.0f;

		Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
		bool hit = intersectNeumann<DIM>(S, st.pt, st.normal, dir, starR, st.onNeumann, h);
		V3 ipt, inrm = mk(0, 0, 0); float idist;
		if (hit) { ipt = h.p; inrm = h.n; idist = h.d; if (DIM == 2) { ipt.z = 0.0f; inrm.z = 0.0f; } }
		else {
			V3 cp = st.onNeumann ? offsetPoint<DIM>(st.pt, neg(st.normal)) : st.pt;
			ipt = cp + starR*dir;
			idist = starR;
		}
		if (!o.ignoreNeumann) { // :212-260; Neumann data is identically zero in the bindings, only the draws are observable
			for (int k = 0; k < DIM; k++) (void)rng.nextFloat();
		}
		if (!o.ignoreSource) { // :262-276
			float pdf;
			g.sampleVolume(dir, rng, pdf);
			if (g.r <= idist) {
				float sc = g.norm_()*sourceAt<DIM>(S, g.yVol);
				st.totalSource += st.throughput*sc;
			}
		}
		if (!hit && outsideBox<DIM>(S, ipt)) return kEscaped;

		st.prevDist = idist; st.prevDir = dir; st.pt = ipt; st.normal = inrm; st.onNeumann = hit;

		st.throughput *= g.directionSampledPoissonKernel(st.pt);
		if (st.throughput < o.russianRouletteThreshold) {
			float survival = st.throughput/o.russianRouletteThreshold;
			if (survival < rng.nextFloat()) { st.throughput = 0.0f; return kRussianRoulette; }
			st.throughput = o.russianRouletteThreshold;
		}
		st.walkLength++;
		if (st.walkLength > o.maxWalkLength) return kExceededLength;
		if (S.absorption > 0.0f && o.stepsBeforeApplyingTikhonov == st.walkLength) g.init(true, S.absorption); // :319-321
		dirichletDist = distDirichlet<DIM>(S, st.pt);
	}
	return kReachedDirichlet;
}

// Scratch accessor for the Latin-hypercube samples of one point: element j lives at base[j*stride]
// (stride = number of points in the launch so that a warp touches consecutive addresses).
struct LhsScratch {
	float* base; size_t stride;
	NMC_HD float& at(int j) const { return base[(size_t)j*stride]; }
};

// generateStratifiedSamples<DIM-1> (include/zombie/core/sampling.h:434-457)
template <int D>
NMC_HD void stratify(const LhsScratch& s, int n, Pcg32& rng) {
	const float oneMinusEps = 1.0f - kEps;
	float inv = 1.0f/n;
	for (int i = 0; i < n; i++) for (int j = 0; j < D; j++) {
		float sj = (i + rng.nextFloat())*inv;
		s.at(D*i + j) = minS(sj, oneMinusEps);
	}
	for (int i = 0; i < D; i++) for (int j = 0; j < n; j++) {
		int other = j + (int)rng.nextBounded((uint32_t)(n - j));
		float t = s.at(D*j + i); s.at(D*j + i) = s.at(D*other + i); s.at(D*other + i) = t;
	}
}

template <int DIM>
NMC_HD void detEstimatePoint(const SceneView& S, const SolverParams& o, V3 x, uint64_t globalIndex,
							 const LhsScratch& lhs, PointResult& out) {
	typedef ExactMath M;
	// createSolutionGrid demo/grid.h:69-102 ; estimationQuantity demo.cpp:152-158
	float dDist = distDirichlet<DIM>(S, x);
	float nDist = distNeumann<DIM>(S, x, false);
	bool inside = insideDomain<DIM>(S, x);
	bool active = inside || S.doubleSided;

	float solMean = 0, solM2 = 0, gradMean[3] = {0, 0, 0}, gradM2[3] = {0, 0, 0}, totalFirstSource = 0;
	int nSol = 0, nGrad = 0, totalWalkLength = 0;
	unsigned walksStarted = 0, steps = 0;

	if (active) {
		Pcg32 rng; rng.seed(pointSeed(o.seed, globalIndex), 1);
		int nWalks = o.nWalks, nAnti = 1;
		if (o.useGradientAntitheticVariates) { nWalks = nWalks/2 > 1 ? nWalks/2 : 1; nAnti = 2; }
		float firstR = kShrink*minS(dDist, nDist);
		const int D = DIM - 1;
		stratify<DIM - 1>(lhs, 2*nWalks, rng);

		for (int w = 0; w < nWalks; w++) {
			float boundaryPdf = 0, sourcePdf = 0;
			V3 boundaryPt = mk(0, 0, 0), sourcePt = mk(0, 0, 0);
			uint32_t seed = rng.nextUInt(); // stands in for the clock read at :498
			float bcv = 0.0f, scv = 0.0f;
			if (o.useGradientControlVariates) { bcv = solMean; scv = totalFirstSource/(nSol > 1 ? nSol : 1); }
			for (int a = 0; a < nAnti; a++) {
				BallExact<DIM> g; g.init(S.absorption > 0.0f && o.stepsBeforeApplyingTikhonov == 0, S.absorption);
				WalkState st;
				st.pt = x; st.normal = st.prevDir = st.srcGradDir = st.bdyGradDir = mk(0, 0, 0);
				st.prevDist = 0.0f; st.throughput = 1.0f; st.onNeumann = false;
				st.totalNeumann = st.totalSource = st.firstSource = 0.0f; st.walkLength = 0;
				g.update(st.pt, firstR);
				if (!o.ignoreSource) { // :526-543
					if (a == 0) {
						V3 sd = sphereDir<DIM, M>(lhs.at(D*(2*w)), DIM == 3 ? lhs.at(D*(2*w) + D - 1) : 0.0f);
						g.sampleVolume(sd, rng, sourcePdf);
						sourcePt = g.yVol;
					} else {
						V3 sd = sourcePt - st.pt;
						g.yVol = st.pt - sd;
						g.r = norm(sd);
					}
					float gn = g.norm_();
					float sc = gn*sourceAt<DIM>(S, g.yVol);
					st.totalSource += st.throughput*sc;
					st.firstSource = sc;
					st.srcGradDir = g.gradient()/(sourcePdf*gn);
				}
				if (a == 0) { // :547-567
					const float ub0 = lhs.at(D*(2*w + 1)), ub1 = DIM == 3 ? lhs.at(D*(2*w + 1) + D - 1) : 0.0f;
					V3 bd;
					if (o.useCosineSampling) { // :550-554: cosine lobe around +/- directionForDerivative = e_x (SampleEstimationData's default, :680-683)
						bd = cosineHemisphere<DIM, M>(ub0, ub1);
						float* last = DIM == 2 ? &bd.y : &bd.z;
						if (rng.nextFloat() < 0.5f) *last *= -1.0f;
						boundaryPdf = 0.5f*pdfCosineHemisphere<DIM>(fabsf(*last));
						bd = toFrame<DIM>(mk(1.0f, 0.0f, 0.0f), bd);
					} else {
						bd = sphereDir<DIM, M>(ub0, ub1);
						boundaryPdf = pdfSphere<DIM>(1.0f);
					}
					g.ySurf = g.c + g.R*bd;
					boundaryPt = g.ySurf;
				} else {
					V3 bd = boundaryPt - st.pt;
					g.ySurf = st.pt - bd;
				}
				st.prevDist = g.R;
				st.prevDir = (g.ySurf - st.pt)/g.R;
				st.pt = g.ySurf;
				st.throughput *= g.poissonKernel()/boundaryPdf;
				st.bdyGradDir = g.poissonKernelGradient()/(boundaryPdf*st.throughput);

				float dirichletDist = distDirichlet<DIM>(S, st.pt);
				rng.seed(seed, 1);
				walksStarted++;
				int code = detWalk<DIM>(S, o, dirichletDist, rng, g, st, steps);
				if (code == kReachedDirichlet || code == kRussianRoulette) { // :583-614
					float terminal = 0.0f; // pde.dirichlet == 0 and initVal == 0 (:331-351)
					float total = st.throughput*terminal + st.totalNeumann + st.totalSource;
					float bContribution = total - st.firstSource;
					float bE[3], sE[3];
					bE[0] = (bContribution - bcv)*st.bdyGradDir.x; sE[0] = (st.firstSource - scv)*st.srcGradDir.x;
					bE[1] = (bContribution - bcv)*st.bdyGradDir.y; sE[1] = (st.firstSource - scv)*st.srcGradDir.y;
					bE[2] = (bContribution - bcv)*st.bdyGradDir.z; sE[2] = (st.firstSource - scv)*st.srcGradDir.z;
					nSol += 1; welford(total, solMean, solM2, nSol);
					totalFirstSource += st.firstSource;
					nGrad += 1;
					for (int i = 0; i < DIM; i++) welford(bE[i] + sE[i], gradMean[i], gradM2[i], nGrad);
					totalWalkLength += st.walkLength;
				}
			}
		}
	}
	// getSolution / getGradient demo/grid.h:155-179, 207-237
	bool maskP = fabsf(nDist) < o.boundaryDistanceMask;
	bool maskG = (!inside && !S.doubleSided) || maskP;
	out.p = maskP ? 0.0f : solMean;
	for (int k = 0; k < 3; k++) out.g[k] = maskG ? 0.0f : gradMean[k];
	out.solMean = solMean; out.solM2 = solM2;
	for (int k = 0; k < 3; k++) { out.gradMean[k] = gradMean[k]; out.gradM2[k] = gradM2[k]; }
	out.meanFirstSource = totalFirstSource/(nSol > 1 ? nSol : 1);
	out.nSol = nSol; out.totalWalkLength = totalWalkLength; out.active = active ? 1 : 0;
	out.walksStarted = walksStarted; out.steps = steps;
}

// estimateSolution (walk_on_stars.h:354-461): EstimationQuantity::Solution at one sample point that lies in the domain
// (type 0) or ON the reflecting boundary (type 2 = SampleType::OnNeumannBoundary, with its unit normal) -- the estimator
// boundary value caching runs at its cache points (boundary_sampler.h:148-185).  The bindings have no Dirichlet geometry,
// so SampleType::OnDirichletBoundary cannot occur.  One pcg32 stream per point, not re-seeded between walks.
// out4: mean, M2, number of estimates, summed walk length; *firstR receives firstSphereRadius.
template <int DIM>
NMC_HD void detEstimateSolution(const SceneView& S, const SolverParams& o, V3 x, V3 nrm, int type, bool normalAligned,
								int nWalks, uint64_t globalIndex, float* out4, float* firstR, unsigned& walksStarted, unsigned& steps) {
	typedef ExactMath M;
	const float dDist = distDirichlet<DIM>(S, x);
	if (dDist <= o.epsilonShell) nWalks = 1; // :382-385
	V3 currentNormal = nrm, prevDir = nrm;
	bool flip = false;
	if (S.doubleSided && type == 2 && normalAligned) { currentNormal = nrm*-1.0f; prevDir = nrm*-1.0f; flip = true; } // :395-401
	float firstSphereRadius;
	if (dDist > o.epsilonShell && o.stepsBeforeUsingMaximalSpheres != 0) { // :405-425
		float starR = starRadius<DIM, M>(S, x, o.minStarRadius, dDist, o.silhouettePrecision, flip);
		if (o.minStarRadius <= dDist) starR = maxS(kShrink*starR, o.minStarRadius);
		firstSphereRadius = starR;
	} else firstSphereRadius = dDist;
	float mean = 0.0f, M2 = 0.0f;
	int nSol = 0, totalLen = 0;
	Pcg32 rng; rng.seed(pointSeed(o.seed, globalIndex), 1);
	for (int w = 0; w < nWalks; w++) {
		BallExact<DIM> g; g.init(S.absorption > 0.0f && o.stepsBeforeApplyingTikhonov == 0, S.absorption);
		WalkState st;
		st.pt = x; st.normal = currentNormal; st.prevDir = prevDir; st.srcGradDir = st.bdyGradDir = mk(0, 0, 0);
		st.prevDist = kMaxF; st.throughput = 1.0f; st.onNeumann = type == 2;
		st.totalNeumann = st.totalSource = st.firstSource = 0.0f; st.walkLength = 0;
		walksStarted++;
		int code = detWalk<DIM>(S, o, dDist, rng, g, st, steps, firstSphereRadius);
		if (code == kReachedDirichlet || code == kRussianRoulette) { // :446-458; terminal value 0 (pde.dirichlet == 0, initVal == 0)
			float total = st.throughput*0.0f + st.totalNeumann + st.totalSource;
			nSol += 1; welford(total, mean, M2, nSol);
			totalLen += st.walkLength;
		}
	}
	out4[0] = mean; out4[1] = M2; out4[2] = (float)nSol; out4[3] = (float)totalLen;
	if (firstR) *firstR = firstSphereRadius;
}

} // namespace nmc
