// csrc/wost_fast.cu -- default-mode estimator: persistent walk kernel for sm_100a.
//
// Work decomposition (replaces tbb::parallel_for over points, walk_on_stars.h:91-103, and the
// sequential per-point loop of estimateSolutionAndGradient, :466-617):
//   * the grid is persistent: smCount x (resident CTAs per SM) CTAs, every WARP pulls sample points
//     from a global atomic queue, so long and short points balance across the 148 SMs;
//   * inside a warp one LANE runs one walk at a time.  A point's walks (two per antithetic pair) are handed
//     out to lanes through a ballot/popc compaction: every loop trip, lanes whose walk has finished are
//     ranked with __ballot_sync/__popc and take the next walk indices, so lanes whose walks have
//     terminated are refilled immediately instead of idling until the longest walk ends; the two walks of a
//     pair may run on different lanes (mirrored first-ball samples, same walk stream);
//   * every loop trip each busy lane executes exactly one walk-on-stars step (geometric queries, radial
//     inverse-CDF sample of the ball Green's function, source-grid gather, bookkeeping);
//   * the first-ball source samples of a point (one per antithetic pair, all in the SAME ball, no geometry)
//     are not interleaved with the steps: they are computed 32 pairs at a time by the whole warp in
//     converged code, parked in shared memory, and picked up by whichever lane is handed the pair;
//   * the boundary structure (nodes, primitives, face normals, silhouettes; 16-byte records) is staged
//     in shared memory once per CTA when it fits, and read with vectorised loads;
//   * per-point estimates are reduced in registers/shared memory: control variates come from warp-wide
//     running sums (shuffle-reduced at the refill point), the final means from a shuffle tree; one lane writes p and grad p.
//   * RNG is counter-based: every (seed, global point index, pair, stream) tuple hashes to its own
//     pcg32 state, so results do not depend on which warp/GPU processes a point.
// Statistical (not bitwise) equivalence to the reference: same estimator, same stratification of the
// first-ball directions (Latin hypercube via a keyed permutation instead of a stored shuffle), same
// antithetic pairing and control variates; the radial sample is drawn by inverting the CDF that the
// reference's rejection sampler targets.
#define NMC_FAST_GEOM 1
#ifndef NMC_NO_BESSEL_TAB
#define NMC_BESSEL_TAB 1
#endif
#include "nmc_device.h"
#include "nmc_packet.cuh"
#include "bessel_table.h"
#include "../../include/nmcfs.h"
#include <cstdlib>
#include <mutex>

namespace nmc {

static constexpr int kBlock = 128;
static constexpr int kWarps = kBlock/32;
static constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warpSum(float v) {
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
	return v;
}
__device__ __forceinline__ unsigned warpSumU(unsigned v) {
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
	return v;
}

enum LaneState { kNeedPair = 0, kWalking = 2, kIdle = 3 };
// first-ball record per pair: d0.xyz, e0.xyz, firstSource (walk 0), firstSource (twin), sfr, the pair's control variates (bcv, scv)
static constexpr int kFbFields = 12; // + the first boundary sample's pdf weight (cosine sampling for derivatives)

// resident CTAs per SM the compiler must allow for: 2D 6 (80 registers/thread: the Bessel code spills less), 3D 7
// (72 registers) -- measured on B200 against 8 (64 registers): profiles/README.md, r02 A/B table
#ifdef NMC_MINB
#define NMC_MINB2 NMC_MINB
#define NMC_MINB3 NMC_MINB
#endif
#ifndef NMC_MINB2
#define NMC_MINB2 6
#endif
#ifndef NMC_MINB3
#define NMC_MINB3 7
#endif
__device__ __forceinline__ void stackInit(StridedStack& s, int* base, int slots) {
	s.nodes = base + threadIdx.x; s.dists = reinterpret_cast<float*>(base + slots*kBlock) + threadIdx.x; s.stride = kBlock;
}
__device__ __forceinline__ void stackInit(LocalStack&, int*, int) {}

// STATS: the per-point statistics record of the parity tests (variances, mean walk length) costs five more
// accumulators per lane; the product path (stats12 == nullptr) runs the instantiation without them.
// FLAT: 0 = tree traversals per step, 1 = flat scans over tables staged in shared memory (<= 128 primitives),
// 2 = two-level flat scans over tables in global memory (larger meshes; the tree is still walked once per point),
// 3 = warp-packet tree traversals (nmc_packet.cuh), selectable with NMC_BIG_MESH=packet
template <int DIM, class STACK, int FLAT, bool STATS>
__global__ void __launch_bounds__(kBlock, DIM == 2 ? NMC_MINB2 : NMC_MINB3)
fastKernel(SceneView Sg, SolverParams o, const float* __restrict__ pts, long long n, unsigned long long indexOffset,
		   float* __restrict__ pOut, float* __restrict__ gOut, unsigned int* __restrict__ workCounter,
		   Counters* __restrict__ counters, float* __restrict__ stats12, int stageQuads, int stackSlots) {
	typedef FastMath M;
	extern __shared__ float4 stage[];
	// per-thread traversal stack in shared memory, interleaved across the CTA (after the staged scene)
	STACK stack;
	stackInit(stack, reinterpret_cast<int*>(stage + stageQuads), stackSlots);
	// first-ball chunk of this warp: [field][lane] floats, behind the stacks
	float* fbuf = reinterpret_cast<float*>(stage + stageQuads) + (size_t)2*stackSlots*kBlock + (threadIdx.x >> 5)*(kFbFields*32);

	// ---- stage the boundary structure in shared memory ------------------------------------------------
	// layout: nodes | prims | primN | silhouettes | [FLAT: group boxes (ray, sil) | ray primitives | ray normals]
	SceneView S = Sg;
	FlatTab F = {};
	constexpr int G = FlatGroup<DIM>::n; // the flat-scan lists are padded to whole groups (scene_build.cpp)
	const int nSilP = (Sg.nSilU + G - 1)/G*G, nRayP = (Sg.nRay + G - 1)/G*G;
	if (FLAT == 2) { // large mesh: records and culling boxes stay in global memory (L1 / L2); staging the boxes in shared memory
		// was measured slower (1.9e8 against 2.6e8 walks/s on box_sphere: the 48 KB stage costs resident CTAs)
		F.silsU = Sg.silsU; F.nSilU = Sg.nSilU; F.grpS = Sg.grpS; F.supS = Sg.supS;
		F.rayP = Sg.rayP; F.rayN = Sg.rayN; F.nRay = Sg.nRay; F.grpP = Sg.grpP; F.supP = Sg.supP;
	}
	if (FLAT == 1 || (FLAT == 0 && stageQuads > 0)) { // FLAT == 1 scenes always fit (launchFast)
		// FLAT == 1: the de-duplicated silhouette list replaces the per-leaf references (the tree is only walked once per point)
		const float4* silSrc = FLAT == 1 ? Sg.silsU : Sg.sils;
		const int qN = 4*Sg.nNodes, qP = (DIM == 2 ? 1 : 3)*Sg.nPrims, qF = Sg.nPrims, qS = (DIM == 2 ? 2 : 4)*(FLAT == 1 ? nSilP : Sg.nSilRefs);
#pragma unroll 1
		for (int i = threadIdx.x; i < qN; i += kBlock) stage[i] = Sg.nodes[i];
#pragma unroll 1
		for (int i = threadIdx.x; i < qP; i += kBlock) stage[qN + i] = Sg.prims[i];
#pragma unroll 1
		for (int i = threadIdx.x; i < qF; i += kBlock) stage[qN + qP + i] = Sg.primN[i];
#pragma unroll 1
		for (int i = threadIdx.x; i < qS; i += kBlock) stage[qN + qP + qF + i] = silSrc[i];
		S.nodes = stage; S.prims = stage + qN; S.primN = stage + qN + qP;
		if (FLAT == 1) {
			const int qGP = 2*(nRayP/G), qGS = 2*(nSilP/G), qRP = (DIM == 2 ? 1 : 3)*nRayP;
			const int oG = qN + qP + qF + qS, oR = oG + qGP + qGS;
#pragma unroll 1
			for (int i = threadIdx.x; i < qGP; i += kBlock) stage[oG + i] = Sg.grpP[i];
#pragma unroll 1
			for (int i = threadIdx.x; i < qGS; i += kBlock) stage[oG + qGP + i] = Sg.grpS[i];
#pragma unroll 1
			for (int i = threadIdx.x; i < qRP; i += kBlock) stage[oR + i] = Sg.rayP[i];
#pragma unroll 1
			for (int i = threadIdx.x; i < nRayP; i += kBlock) stage[oR + qRP + i] = Sg.rayN[i];
			// pointers derived from the shared array only: the scans compile to LDS
			F.silsU = stage + (qN + qP + qF); F.nSilU = Sg.nSilU;
			F.grpP = stage + oG; F.grpS = stage + (oG + qGP);
			F.rayP = stage + oR; F.rayN = stage + (oR + qRP); F.nRay = Sg.nRay;
		} else S.sils = stage + qN + qP + qF;
		__syncthreads();
	}

	const int lane = threadIdx.x & 31;
	const unsigned ltMask = (1u << lane) - 1u;
	unsigned cStarted = 0, cCompleted = 0, cSteps = 0, cActive = 0;
	unsigned long long cTrips = 0, cLaneSlices = 0; // warp-level: loop trips and busy lanes per trip (lane 0 only)

	int nPairs = o.nWalks, nAnti = 1;
	if (o.useGradientAntitheticVariates) { nPairs = o.nWalks/2 > 1 ? o.nWalks/2 : 1; nAnti = 2; }
	const unsigned nStrata = 2u*(unsigned)nPairs;
	const float invStrata = 1.0f/(float)nStrata;
	const bool yukawa0 = S.absorption > 0.0f && o.stepsBeforeApplyingTikhonov == 0;

	for (;;) {
		unsigned pi = 0;
		if (lane == 0) pi = atomicAdd(workCounter, 1u);
		pi = __shfl_sync(kFull, pi, 0);
		if ((long long)pi >= n) break;

		// ---- per-point set-up (uniform across the warp): createSolutionGrid + estimationQuantity --------
		const V3 x0 = mk(pts[(size_t)pi*DIM], pts[(size_t)pi*DIM + 1], DIM == 3 ? pts[(size_t)pi*DIM + 2] : 0.0f);
		const float dDist = distDirichlet<DIM>(S, x0);
		float nDist; bool inside = true;
		if (FLAT == 3) { // one packet traversal (all lanes ask for x0) gives the distance and the pseudo-normal of the inside test
			Hit h0; h0.d = kMaxF; h0.p = x0; h0.n = mk(0, 0, 0);
			const bool f0 = S.nPrims > 0 && packetClosestPoint<DIM, WarpOps>(S, x0, kMaxF, S.watertight != 0, h0);
			nDist = f0 ? h0.d : kMaxF;
			if (S.watertight) { // insideDomain (fcpw_scene_loader.h:642-648)
				const float d2s = (f0 && dot(x0 - h0.p, h0.n) > 0.0f ? 1.0f : -1.0f)*nDist;
				inside = fabsf(dDist) < fabsf(d2s) ? dDist < 0.0f : d2s < 0.0f;
			}
		} else {
			nDist = distNeumann<DIM>(S, stack, x0, false);
			inside = S.watertight ? insideDomain<DIM>(S, stack, x0) : true; // pseudo-normals (nrmV) stay in global memory
		}
		// points inside the boundary mask are zeroed on output (demo/grid.h:174,227); do not walk them
		const bool masked = fabsf(nDist) < o.boundaryDistanceMask;
		const bool active = (inside || S.doubleSided) && !masked && nDist > 0.0f;

		float sTot = 0.0f, sTot2 = 0.0f, sG[3] = {0, 0, 0}, sG2[3] = {0, 0, 0}, sFirst = 0.0f;
		unsigned nDone = 0, lenSum = 0;

		if (active) {
			const unsigned long long gidx = indexOffset + pi;
			const unsigned long long key = pointSeed(o.seed, gidx);
			const unsigned permKey0 = (unsigned)(key >> 32), permKey1 = (unsigned)key;
			const float firstR = kShrink*fminf(dDist, nDist);
			BallFast<DIM> fb; fb.init(yukawa0, S.absorption); fb.update(firstR);
			const float normG0 = fb.normG(), exitT = fb.exitThroughput(), bfr = fb.bdyGradFactor()/firstR;
			float cvTot = 0.0f, cvCnt = 0.0f, cvFirst = 0.0f;   // warp-uniform running sums over finished walks
			float pendTot = 0.0f, pendCnt = 0.0f, pendFirst = 0.0f; // this lane's finished walks not yet folded in

			int state = kNeedPair;
			int nextWalk = 0, chunkBase = 0, chunkEnd = 0; // warp-uniform; walks are handed out one by one (walk w = pair w/nAnti, twin w%nAnti);
			const int nWalksPt = nPairs*nAnti;             // pairs [chunkBase, chunkEnd) are parked in fbuf
			int pair = 0, anti = 0, walkLength = 0;
			bool onNeumann = false;
			Pcg32 rng; rng.state = 0; rng.inc = 1;
			unsigned long long walkSeed = 0;
			V3 pt = x0, normal = mk(0, 0, 0), d0 = mk(0, 0, 0), e0 = mk(0, 0, 0);
		bool flipNext = false; // double-sided boundaries: the walk arrived on the back side of the face it sits on
			float throughput = 1.0f, totalSource = 0.0f, firstSource = 0.0f, sfr = 0.0f, bcv = 0.0f, scv = 0.0f;
			BallFast<DIM> bl = fb;

			for (;;) {
				// ---- refill finished lanes with the next pairs (ballot + popc compaction) -------------------
				unsigned need = __ballot_sync(kFull, state == kNeedPair);
				if (need) {
					if (o.useGradientControlVariates) { // fold the finished walks into the running means (converged code)
						cvTot += warpSum(pendTot); cvCnt += warpSum(pendCnt); cvFirst += warpSum(pendFirst);
						pendTot = pendCnt = pendFirst = 0.0f;
					}
					// the unit handed to a lane is ONE walk, so the two antithetic walks of a pair may run on different lanes
					// (same first-ball samples mirrored, same walk stream): half the granularity, a shorter tail per point
					const int mineW = nextWalk + __popc(need & ltMask), total = __popc(need);
					const int lastW = nextWalk + total < nWalksPt ? nextWalk + total : nWalksPt;      // exclusive
					const int mine = nAnti == 2 ? mineW >> 1 : mineW, myAnti = nAnti == 2 ? mineW & 1 : 0;
					const int lastNeeded = nAnti == 2 ? (lastW + 1) >> 1 : lastW;                      // pairs, exclusive
					const bool want = state == kNeedPair && mineW < nWalksPt;
					bool fetched = false;
					float pdfWeight = 1.0f; // first boundary sample: uniform pdf / pdf used (cosine sampling only)
					// Control variates = running means over the walks finished so far (walk_on_stars.h:501-506), fixed ONCE per
					// pair as in the reference: the lane that takes a pair's first walk parks them next to the pair's first-ball
					// samples and the twin reads them, so a twin handed out in a later refill does not see a mean that already
					// contains its partner's (correlated) total.
					const float icnt_ = 1.0f/fmaxf(cvCnt, 1.0f);
					const float bcvNow = o.useGradientControlVariates ? cvTot*icnt_ : 0.0f, scvNow = o.useGradientControlVariates ? cvFirst*icnt_ : 0.0f;
#define NMC_FETCH_FIRST_BALL(slot) do { const int sl_ = (slot); \
						d0 = mk(fbuf[sl_], fbuf[32 + sl_], DIM == 3 ? fbuf[64 + sl_] : 0.0f); \
						e0 = mk(fbuf[96 + sl_], fbuf[128 + sl_], DIM == 3 ? fbuf[160 + sl_] : 0.0f); \
						firstSource = fbuf[(myAnti ? 224 : 192) + sl_]; sfr = fbuf[256 + sl_]; \
						if (o.useCosineSampling) pdfWeight = fbuf[352 + sl_]; \
						if (myAnti == 0) { fbuf[288 + sl_] = bcvNow; fbuf[320 + sl_] = scvNow; } fetched = true; } while (0)
					if (want && mine < chunkEnd) NMC_FETCH_FIRST_BALL(mine - chunkBase);
					bool cvTaken = false;
					if (lastNeeded > chunkEnd) {
						// ---- first-ball source samples of the next 32 pairs, one per lane, converged ---------------
						__syncwarp();
						// a twin served from the chunk that is about to be replaced takes its pair's control variates with it
						if (want && fetched && myAnti == 1) { bcv = fbuf[288 + mine - chunkBase]; scv = fbuf[320 + mine - chunkBase]; cvTaken = true; }
						__syncwarp();
						chunkBase = chunkEnd; chunkEnd = chunkBase + 32 < nPairs ? chunkBase + 32 : nPairs;
						const int cp = chunkBase + lane;
						if (cp < chunkEnd) {
							// stratified first-ball directions (generateStratifiedSamples, sampling.h:434-457):
							// sample 2w drives the source direction, 2w+1 the boundary direction
							const unsigned long long ws = splitmix64(key ^ (0xD1B54A32D192ED03ull*(unsigned long long)(cp + 1)));
							Pcg32 r0; r0.state = splitmix64(ws ^ 0xA0761D6478BD642Full); r0.inc = 3;
							float us0 = ((float)permute(2u*cp, nStrata, permKey0) + r0.nextFloat())*invStrata;
							float ub0 = ((float)permute(2u*cp + 1u, nStrata, permKey0) + r0.nextFloat())*invStrata;
							float us1 = 0.0f, ub1 = 0.0f;
							if (DIM == 3) {
								us1 = ((float)permute(2u*cp, nStrata, permKey1) + r0.nextFloat())*invStrata;
								ub1 = ((float)permute(2u*cp + 1u, nStrata, permKey1) + r0.nextFloat())*invStrata;
							}
							const V3 sdir = sphereDir<DIM, M>(fminf(us0, 1.0f - kEps), fminf(us1, 1.0f - kEps));
							V3 bdir;
							if (o.useCosineSampling) { // walk_on_stars.h:550-554: cosine lobe around +/- directionForDerivative = e_x
								bdir = cosineHemisphere<DIM, M>(fminf(ub0, 1.0f - kEps), fminf(ub1, 1.0f - kEps));
								float last = DIM == 2 ? bdir.y : bdir.z;
								if (r0.nextFloat() < 0.5f) last = -last;
								if (DIM == 2) bdir.y = last; else bdir.z = last;
								// throughput = poissonKernel / pdf: relative to the uniform pdf the default path assumes
								fbuf[352 + lane] = pdfSphere<DIM>(1.0f)/(0.5f*pdfCosineHemisphere<DIM>(fabsf(last)));
								bdir = toFrame<DIM>(mk(1.0f, 0.0f, 0.0f), bdir);
							} else bdir = sphereDir<DIM, M>(fminf(ub0, 1.0f - kEps), fminf(ub1, 1.0f - kEps));
							const V3 be = firstR*bdir;
							const float uA = r0.nextFloat(), uB = r0.nextFloat();
							float rs = 0.0f, fsrc = 0.0f, fsrc1 = 0.0f, sf = 0.0f;
							if (!o.ignoreSource) {
								float gs, qs; bool hframe;
								float xs = fb.sampleX(uA, uB, gs, qs, hframe, true); // g, q at the returned x
								rs = hframe ? xs*fb.R : xs/fb.mu;
								rs = fminf(fmaxf(rs, 1e-4f), fb.R);        // rClamp, distributions.h:378-379
								fsrc = normG0*sourceAt<DIM>(S, x0 + rs*sdir);
								if (nAnti == 2) fsrc1 = normG0*sourceAt<DIM>(S, x0 - rs*sdir); // antithetic twin: mirrored source sample (:532-536)
								// sourceGradientDirection = d * gradientNorm / G(r)  (walk_on_stars.h:542)
								sf = hframe ? fb.srcGradFactorHarmonic(xs) : fb.srcGradFactor(qs, gs);
							}
							fbuf[lane] = rs*sdir.x; fbuf[32 + lane] = rs*sdir.y; if (DIM == 3) fbuf[64 + lane] = rs*sdir.z;
							fbuf[96 + lane] = be.x; fbuf[128 + lane] = be.y; if (DIM == 3) fbuf[160 + lane] = be.z;
							fbuf[192 + lane] = fsrc; fbuf[224 + lane] = fsrc1; fbuf[256 + lane] = sf/fmaxf(rs, 1e-20f);
						}
						cTrips++; cLaneSlices += (unsigned)(chunkEnd - chunkBase);
						__syncwarp();
						if (want && !fetched) NMC_FETCH_FIRST_BALL(mine - chunkBase);
					}
#undef NMC_FETCH_FIRST_BALL
					__syncwarp();
					if (state == kNeedPair) {
						if (want) {
							pair = mine; anti = myAnti;
							if (myAnti == 0) { bcv = bcvNow; scv = scvNow; }
							else if (!cvTaken) { bcv = fbuf[288 + mine - chunkBase]; scv = fbuf[320 + mine - chunkBase]; }
							walkSeed = splitmix64(key ^ (0xD1B54A32D192ED03ull*(unsigned long long)(pair + 1)));
							// start from the boundary sample, mirrored for the antithetic twin (:564-567), same walk stream (:579)
							pt = x0 + (anti ? neg(e0) : e0); normal = mk(0, 0, 0); onNeumann = false; flipNext = false; walkLength = 0;
							throughput = exitT*pdfWeight; totalSource = firstSource;
							rng.state = walkSeed; rng.inc = 1;
							bl = fb;
							state = kWalking; cStarted++;
						} else state = kIdle;
					}
					nextWalk += total;
				}
				const unsigned busy = __ballot_sync(kFull, state == kWalking);
				if (busy == 0u) break;
				cTrips++; cLaneSlices += __popc(busy);

				// ---- phase 1: geometry ---------------------------------------------------------------------------
				V3 dir = mk(0, 0, 0), ipt = pt, inrm = mk(0, 0, 0);
				float idist = 0.0f, uRad = 0.0f, uRad2 = 0.0f, uRR = 1.0f;
				bool hit = false, sliceActive = false, terminated = false, completed = false;
				if (FLAT == 3) {
					// packet traversals: the per-lane parts of the step run under the lane's own predicate, the two tree queries are
					// made by the whole warp (lanes without a query pass a negative radius / ray length)
					bool stepLive = false, flipOrient = false, needSil = false;
					float dirichletDist = 0.0f, starR = 0.0f, dsil = 0.0f;
					if (state == kWalking) {
						cSteps++;
						dirichletDist = distDirichlet<DIM>(S, pt);
						if (!(dirichletDist > o.epsilonShell)) { terminated = true; completed = true; }
						else {
							stepLive = true;
							if (S.doubleSided && onNeumann && flipNext) { normal = normal*-1.0f; flipOrient = true; } // :154-160
							starR = dirichletDist;
							needSil = o.stepsBeforeUsingMaximalSpheres > walkLength && o.minStarRadius <= dirichletDist && S.nPrims > 0;
						}
					}
					const bool fsil = packetClosestSilhouette<DIM, WarpOps>(S, pt, needSil ? (dirichletDist < kMaxF ? dirichletDist*dirichletDist : kMaxF) : -1.0f,
						!flipOrient, o.minStarRadius*o.minStarRadius, o.silhouettePrecision, dsil);
					V3 ro = pt;
					if (stepLive) {
						if (o.stepsBeforeUsingMaximalSpheres > walkLength && o.minStarRadius <= dirichletDist) { // starRadius(), fcpw_scene_loader.h:621-641
							starR = fsil ? fmaxf(dsil, o.minStarRadius) : fmaxf(dirichletDist, o.minStarRadius);
							starR = fmaxf(kShrink*starR, o.minStarRadius);
						}
						bl.update(starR);
						float u0 = rng.nextFloat(), u1 = rng.nextFloat();
						uRad = rng.nextFloat(); uRad2 = rng.nextFloat(); uRR = rng.nextFloat();
						dir = sphereDir<DIM, M>(u0, u1);
						if (onNeumann && dot(normal, dir) > 0.0f) dir = dir*-1.0f;
						if (onNeumann) ro = offsetPoint<DIM>(pt, neg(normal));
						if (DIM == 2) { ro.z = 0.0f; dir.z = 0.0f; }
					}
					Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
					hit = packetRay<DIM, WarpOps>(S, ro, dir, stepLive && S.nPrims > 0 ? starR : -1.0f, h);
					if (stepLive) {
						if (hit) { ipt = h.p; inrm = h.n; idist = h.d; }
						else { ipt = ro + starR*dir; idist = starR; }
						sliceActive = !o.ignoreSource;
					}
				} else
				if (state == kWalking) {
					cSteps++;
					float dirichletDist = distDirichlet<DIM>(S, pt);
					if (!(dirichletDist > o.epsilonShell)) { terminated = true; completed = true; }
					else {
						bool flipOrient = false;
						if (S.doubleSided && onNeumann && flipNext) { normal = normal*-1.0f; flipOrient = true; } // :154-160
						float starR;
						if (o.stepsBeforeUsingMaximalSpheres <= walkLength) starR = dirichletDist;
						else {
							if (FLAT) { // fcpw_scene_loader.h:621-641 with the flat scan
								starR = dirichletDist;
								if (o.minStarRadius <= dirichletDist) {
									float dsil;
									float r2max = dirichletDist < kMaxF ? dirichletDist*dirichletDist : kMaxF;
									bool f = flatClosestSilhouette<DIM, FLAT == 2>(F, pt, r2max, !flipOrient, o.minStarRadius*o.minStarRadius, o.silhouettePrecision, dsil);
									starR = f ? fmaxf(dsil, o.minStarRadius) : fmaxf(dirichletDist, o.minStarRadius);
								}
							} else
							starR = starRadius<DIM, M>(S, stack, pt, o.minStarRadius, dirichletDist, o.silhouettePrecision, flipOrient);
							if (o.minStarRadius <= dirichletDist) starR = fmaxf(kShrink*starR, o.minStarRadius);
						}
						bl.update(starR);
						float u0 = rng.nextFloat(), u1 = rng.nextFloat();
						uRad = rng.nextFloat(); uRad2 = rng.nextFloat(); uRR = rng.nextFloat();
						dir = sphereDir<DIM, M>(u0, u1);
						if (onNeumann && dot(normal, dir) > 0.0f) dir = dir*-1.0f;
						Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
						if (FLAT) {
							V3 ro = onNeumann ? offsetPoint<DIM>(pt, neg(normal)) : pt;
							hit = flatRay<DIM, FLAT == 2>(F, ro, dir, starR, h);
						} else
						hit = intersectNeumann<DIM>(S, stack, pt, normal, dir, starR, onNeumann, h);
						if (hit) { ipt = h.p; inrm = h.n; idist = h.d; }
						else {
							V3 cp = onNeumann ? offsetPoint<DIM>(pt, neg(normal)) : pt;
							ipt = cp + starR*dir; idist = starR;
						}
						sliceActive = !o.ignoreSource;
					}
				}

				// ---- phase 2 (converged): radial sample of the ball Green's function + source gather ------
				float contribution = 0.0f, xs = 0.0f, gs = 0.0f, qs = 0.0f, rs = 0.0f; bool hframe = false;
				if (sliceActive) {
					xs = bl.sampleX(uRad, uRad2, gs, qs, hframe);
					rs = hframe ? xs*bl.R : xs/bl.mu;
					rs = fminf(fmaxf(rs, 1e-4f), bl.R);        // rClamp, distributions.h:378-379
					if (rs <= idist) contribution = bl.normG()*sourceAt<DIM>(S, pt + rs*dir);
				}

				// ---- phase 3: bookkeeping ---------------------------------------------------------------------
				if (state == kWalking) {
					if (!terminated) {
						totalSource += throughput*contribution;
						if (!hit && outsideBox<DIM>(S, ipt)) terminated = true; // EscapedDomain: discarded
						else {
							// directionSampledPoissonKernel at the new position: T(starR) is already known when the
							// walk lands on the sphere, otherwise evaluate T at the hit distance
							throughput *= hit ? bl.stepThroughput(idist) : bl.exitThroughput();
							// Reference quirk kept on purpose: its float Bessel/exp members overflow once r*sqrt(lambda)
							// exceeds 91.906 (2D, bessi1 -> inf) / 103.9 (3D, expf -> 0), the throughput becomes NaN, the
							// walk can no longer be stopped by Russian roulette and is eventually discarded
							// (distributions.h:669-677, 801-813).
							if (bl.yukawa && fmaxf(1e-4f, hit ? idist : bl.R)*bl.mu > (DIM == 2 ? 91.9063f : 103.9f)) terminated = true;
							flipNext = hit && dot(dir, inrm) < 0.0f; // what dot(prevDirection, currentNormal) < 0 will say at the next step (:154-160)
							pt = ipt; normal = inrm; onNeumann = hit;
							if (!(throughput == throughput)) { terminated = true; } // NaN guard: discard
							if (!terminated && throughput < o.russianRouletteThreshold) {
								if (throughput/o.russianRouletteThreshold < uRR) { throughput = 0.0f; terminated = true; completed = true; }
								else throughput = o.russianRouletteThreshold;
							}
							if (!terminated) {
								walkLength++;
								if (walkLength > o.maxWalkLength) terminated = true; // ExceededMaxWalkLength: discarded
								else if (S.absorption > 0.0f && o.stepsBeforeApplyingTikhonov == walkLength) bl.init(true, S.absorption);
							}
						}
					}
					if (terminated) {
						if (completed) { // walk_on_stars.h:583-614
							float total = totalSource;
							float sgn = anti == 0 ? 1.0f : -1.0f;
							float bE = (total - firstSource - bcv)*bfr*sgn, sE = (firstSource - scv)*sfr*sgn;
							float g0 = bE*e0.x + sE*d0.x, g1 = bE*e0.y + sE*d0.y, g2 = bE*e0.z + sE*d0.z;
							sTot += total; sFirst += firstSource;
							sG[0] += g0; sG[1] += g1; sG[2] += g2;
							nDone++;
							if (STATS) { sTot2 += total*total; sG2[0] += g0*g0; sG2[1] += g1*g1; sG2[2] += g2*g2; lenSum += (unsigned)walkLength; }
							pendTot += total; pendCnt += 1.0f; pendFirst += firstSource;
						}
						state = kNeedPair;
					}
				}
			}
			cActive += lane == 0 ? 1u : 0u;
		}

		// ---- reduce the per-lane sums and write the point's estimate -------------------------------------
		float tot = warpSum(sTot), first = warpSum(sFirst);
		float g0 = warpSum(sG[0]), g1 = warpSum(sG[1]), g2 = warpSum(sG[2]);
		unsigned cnt = warpSumU(nDone);
		cCompleted += lane == 0 ? cnt : 0u;
		float inv = 1.0f/(float)(cnt > 0u ? cnt : 1u);
		if (STATS && stats12) {
			float tot2 = warpSum(sTot2);
			float q0 = warpSum(sG2[0]), q1 = warpSum(sG2[1]), q2 = warpSum(sG2[2]);
			unsigned len = warpSumU(lenSum);
			if (lane == 0) {
				float* t = stats12 + (size_t)pi*12;
				float nv = (float)(cnt > 1u ? cnt - 1u : 1u);
				t[0] = tot*inv; t[1] = fmaxf(tot2 - tot*tot*inv, 0.0f)/nv;
				t[2] = g0*inv; t[3] = g1*inv; t[4] = DIM == 3 ? g2*inv : 0.0f;
				t[5] = fmaxf(q0 - g0*g0*inv, 0.0f)/nv; t[6] = fmaxf(q1 - g1*g1*inv, 0.0f)/nv; t[7] = DIM == 3 ? fmaxf(q2 - g2*g2*inv, 0.0f)/nv : 0.0f;
				t[8] = first*inv; t[9] = (float)cnt; t[10] = (float)len*inv; t[11] = active ? 1.0f : 0.0f;
			}
		}
		if (lane == 0) { // getSolution / getGradient masks, demo/grid.h:155-179, 207-237
			bool maskP = fabsf(nDist) < o.boundaryDistanceMask;
			bool maskG = (!inside && !S.doubleSided) || maskP;
			pOut[pi] = maskP ? 0.0f : tot*inv;
			gOut[(size_t)pi*DIM] = maskG ? 0.0f : g0*inv;
			gOut[(size_t)pi*DIM + 1] = maskG ? 0.0f : g1*inv;
			if (DIM == 3) gOut[(size_t)pi*DIM + 2] = maskG ? 0.0f : g2*inv;
		}
		__syncwarp();
	}

	cStarted = warpSumU(cStarted); cSteps = warpSumU(cSteps);
	if (lane == 0 && counters) {
		atomicAdd(&counters->walksStarted, (unsigned long long)cStarted);
		atomicAdd(&counters->walksCompleted, (unsigned long long)cCompleted);
		atomicAdd(&counters->steps, (unsigned long long)cSteps);
		atomicAdd(&counters->activePoints, (unsigned long long)cActive);
		atomicAdd(&counters->trips, cTrips);
		atomicAdd(&counters->laneSlices, cLaneSlices);
	}
}

#if defined(NMC_BESSEL_TAB)
// uploads the Bessel table to the current device once and points cBesselTab at it
static cudaError_t ensureBesselTable() {
	static std::mutex mu;
	static const float4* ptr[64] = {};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	std::lock_guard<std::mutex> lock(mu);
	if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
	if (ptr[dev]) return cudaSuccess;
	const BesselTable& T = besselTable();
	float4* d = nullptr;
	e = cudaMalloc((void**)&d, T.coef.size()*sizeof(float));
	if (e != cudaSuccess) return e;
	e = cudaMemcpy(d, T.coef.data(), T.coef.size()*sizeof(float), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(d); return e; }
	BesselTabView v; v.c = d; v.t0 = T.t0; v.perOctave = (float)T.perOctave; v.n = T.n;
	e = cudaMemcpyToSymbol(cBesselTab, &v, sizeof(v));
	if (e != cudaSuccess) { cudaFree(d); return e; }
	ptr[dev] = d;
	return cudaSuccess;
}
#endif

cudaError_t launchFast(const SceneView& S, const SolverParams& o, const float* d_pts, long long n,
					   unsigned long long indexOffset, float* d_p, float* d_g, unsigned int* d_workCounter,
					   Counters* d_counters, float* d_stats12, int smCount, int maxDepth, cudaStream_t stream, FastLaunchInfo* info) {
	if (n <= 0) return cudaSuccess;
	if (n >= (1ll << 32) - 65536) return cudaErrorInvalidValue;
	const int dim = S.dim;
#if defined(NMC_BESSEL_TAB)
	if (dim == 2) { cudaError_t eb = ensureBesselTable(); if (eb != cudaSuccess) return eb; }
#endif
	// small scenes are scanned flat (no per-step tree traversal); the flat kernels stage every table in shared memory
	// unconditionally, so a scene whose tables do not fit the 48 KB stage falls back to the tree kernels
	bool flat = S.nPrims <= 128 && S.nSilU <= 128; // <= 43 KB of tables (asserted by tests/test_host_logic.py)
	const int G = dim == 2 ? FlatGroup<2>::n : FlatGroup<3>::n;
	const size_t nSilP = (size_t)(S.nSilU + G - 1)/G*G, nRayP = (size_t)(S.nRay + G - 1)/G*G;
	auto stageQuadsFor = [&](bool fl) {
		return (size_t)4*S.nNodes + (size_t)(dim == 2 ? 1 : 3)*S.nPrims + (size_t)S.nPrims + (size_t)(dim == 2 ? 2 : 4)*(fl ? nSilP : (size_t)S.nSilRefs)
			 + (fl ? 2*(nRayP/G) + 2*(nSilP/G) + (size_t)((dim == 2 ? 1 : 3) + 1)*nRayP : 0);
	};
	if (flat && stageQuadsFor(true)*sizeof(float4) > 48*1024) flat = false;
	size_t quads = stageQuadsFor(flat);
	// Beyond the flat-scan limit, by measurement (profiles/r02_mbvh.jsonl, B200, 1e3 .. 1.3e5 primitives): 2D per-lane tree
	// traversals; 3D the two-level flat scan for small meshes (1292 triangles: 1.9e8 walks/s against 1.2e8 tree / 1.5e8 packet), per-lane
	// tree traversals for large ones (20 k: 2.1e7 against 9.6e6 flat2; 82 k: 8.3e6).  The flat scan falls like 1/N, the tree like
	// N^-0.6: the two measured end points cross near 4700 triangles, hence the switch at 4096.  The warp-packet traversals
	// (FLAT == 3, nmc_packet.cuh: one traversal per warp, warp-uniform node loads, register stack) are kept selectable: they never beat
	// the better of the two by much and lose in 2D (the union of 32 lanes' search regions is several times one lane's).
	// NMC_BIG_MESH = packet | tree | flat2 forces one path (A/B measurements, tests).
	static const int bigMode = [] { const char* e = getenv("NMC_BIG_MESH"); return !e ? -1 : e[0] == 't' ? 0 : e[0] == 'f' ? 2 : e[0] == 'p' ? 3 : -1; }();
	const bool packet = !flat && bigMode == 3 && maxDepth + 2 <= NMC_STACK;
	const bool flat2 = !flat && (bigMode == 2 || (bigMode == -1 && S.nPrims <= 4096)) && dim == 3 && maxDepth + 3 <= 24 && S.supP && S.supS;
	if (flat2 || packet) quads = 0; // FLAT == 2 / 3 stage nothing (see the kernel)
	size_t bytes = quads*sizeof(float4);
	int stageQuads = bytes <= 48*1024 ? (int)quads : 0; // larger structures are read through L1/L2
	// traversal stacks: depth of the tree + 2 entries per thread in shared memory when that is small
	int stackSlots = packet ? 0 : maxDepth + 3; // the packet stack lives in registers
	bool smemStack = stackSlots <= 24;
	if (!smemStack) stackSlots = 0; // LocalStack: the first-ball chunks sit right behind the staged scene
	size_t smem = (stageQuads ? bytes : 0) + (size_t)stackSlots*kBlock*8 + (size_t)kWarps*kFbFields*32*sizeof(float);
	void (*kern)(SceneView, SolverParams, const float*, long long, unsigned long long, float*, float*, unsigned int*, Counters*, float*, int, int);
#define NMC_PICK(ST) do { \
		if (flat) kern = dim == 2 ? fastKernel<2, StridedStack, 1, ST> : fastKernel<3, StridedStack, 1, ST>; \
		else if (flat2) kern = dim == 2 ? fastKernel<2, StridedStack, 2, ST> : fastKernel<3, StridedStack, 2, ST>; \
		else if (packet) kern = dim == 2 ? fastKernel<2, StridedStack, 3, ST> : fastKernel<3, StridedStack, 3, ST>; \
		else if (dim == 2) kern = smemStack ? fastKernel<2, StridedStack, 0, ST> : fastKernel<2, LocalStack, 0, ST>; \
		else kern = smemStack ? fastKernel<3, StridedStack, 0, ST> : fastKernel<3, LocalStack, 0, ST>; } while (0)
	if (d_stats12) NMC_PICK(true); else NMC_PICK(false);
#undef NMC_PICK
	if (flat && (!smemStack || stageQuads == 0)) return cudaErrorInvalidConfiguration; // cannot happen: <= 128 primitives give a shallow tree, and see above
	int perSM = 0;
	cudaError_t e = cudaSuccess;
	if (smem > 48*1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); // <= 48 + 24 + 4 KB
	if (e != cudaSuccess) return e;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kBlock, smem);
	if (e != cudaSuccess) return e;
	if (perSM < 1) perSM = 1;
	long long grid = (long long)smCount*perSM;
	long long gridNeeded = (n + kWarps - 1)/kWarps; // one point per warp at a time
	if (grid > gridNeeded) grid = gridNeeded;
	if (info) { info->grid = (int)grid; info->block = kBlock; info->smemBytes = (int)smem; }
	kern<<<(unsigned)grid, kBlock, smem, stream>>>(S, o, d_pts, n, indexOffset, d_p, d_g, d_workCounter, d_counters, d_stats12, stageQuads, stackSlots);
	return cudaGetLastError();
}

// ---- fast-mode probes -----------------------------------------------------------------------------------
template <int DIM>
__global__ void probeFastKernel(int kind, long long n, const float* __restrict__ a0, const float* __restrict__ a1,
								const float* __restrict__ params, float* __restrict__ out) {
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	float lambda = params[0];
	BallFast<DIM> b; b.init(lambda > 0.0f, lambda); b.update(a0[i]);
	if (kind == NMC_PROBE_GREENS_FAST) {
		float r = a1[i];
		float x = b.yukawa ? r*b.mu : r/b.R, T = 1.0f, g = 0.0f, q = 0.0f;
		if (b.yukawa) b.evalTgq(x, T, g, q);
		float* o = out + i*10;
		o[0] = T; o[1] = g; o[2] = b.normG(); o[3] = b.exitThroughput(); o[4] = b.bdyGradFactor();
		o[5] = b.yukawa ? b.srcGradFactor(q, g) : b.srcGradFactorHarmonic(x);
		o[6] = b.stepThroughput(r); o[7] = o[8] = o[9] = 0.0f;
	} else {
		float g, q; bool hf;
		float x = b.sampleX(a1[i], params[1], g, q, hf);
		out[i*2] = hf ? x*b.R : x/b.mu; out[i*2 + 1] = g;
	}
}
cudaError_t launchProbeFast(const SceneView& S, int kind, long long n, const float* a0, const float* a1,
							const float* params, float* d_out, cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	unsigned grid = (unsigned)((n + 127)/128);
	if (S.dim == 2) probeFastKernel<2><<<grid, 128, 0, stream>>>(kind, n, a0, a1, params, d_out);
	else probeFastKernel<3><<<grid, 128, 0, stream>>>(kind, n, a0, a1, params, d_out);
	return cudaGetLastError();
}

// ---- packet probes: the FLAT == 3 kernel's tree queries, 32 consecutive inputs per warp -----------------------------
template <int DIM>
__global__ void probePacketKernel(SceneView S, int kind, long long n, const float* __restrict__ pts, const float* __restrict__ a0,
								  const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ a3,
								  const float* __restrict__ params, float* __restrict__ out) {
	const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	const bool on = i < n; // lanes past the end take part in the traversal without a query
	const V3 x = on ? mk(pts[i*DIM], pts[i*DIM + 1], DIM == 3 ? pts[i*DIM + 2] : 0.0f) : mk(0, 0, 0);
	if (kind == NMC_PROBE_STAR_RADIUS_PACKET) { // starRadius(), nmc_geom.cuh
		const float minR = params[0], prec = params[1]; const bool flipOrient = params[2] != 0.0f;
		const float maxR = on ? a0[i] : 0.0f;
		const bool ask = on && !(minR > maxR) && S.nPrims > 0;
		float d = 0.0f;
		const bool f = packetClosestSilhouette<DIM, WarpOps>(S, x, ask ? (maxR < kMaxF ? maxR*maxR : kMaxF) : -1.0f, !flipOrient, minR*minR, prec, d);
		if (on) out[i] = minR > maxR ? maxR : f ? fmaxf(d, minR) : fmaxf(maxR, minR);
	} else if (kind == NMC_PROBE_RAY_PACKET) { // intersectNeumann(), nmc_geom.cuh
		V3 nn = mk(0, 0, 0), d = mk(0, 0, 0);
		if (on) { nn = mk(a0[i*DIM], a0[i*DIM + 1], DIM == 3 ? a0[i*DIM + 2] : 0.0f); d = mk(a1[i*DIM], a1[i*DIM + 1], DIM == 3 ? a1[i*DIM + 2] : 0.0f); }
		V3 o = on && a3[i] != 0.0f ? offsetPoint<DIM>(x, neg(nn)) : x;
		if (DIM == 2) { o.z = 0.0f; d.z = 0.0f; }
		Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
		const bool hit = packetRay<DIM, WarpOps>(S, o, d, on && S.nPrims > 0 ? a2[i] : -1.0f, h);
		if (on) {
			float* r = out + i*(2 + 2*DIM);
			r[0] = hit ? 1.0f : 0.0f; r[1] = h.d;
			r[2] = h.p.x; r[3] = h.p.y; if (DIM == 3) r[4] = h.p.z;
			r[2 + DIM] = h.n.x; r[3 + DIM] = h.n.y; if (DIM == 3) r[4 + DIM] = h.n.z;
		}
	} else {
		Hit h; h.d = kMaxF; h.p = x; h.n = mk(0, 0, 0);
		const bool f = packetClosestPoint<DIM, WarpOps>(S, x, on && S.nPrims > 0 ? kMaxF : -1.0f, true, h);
		if (on) { out[2*i] = f ? h.d : kMaxF; out[2*i + 1] = f ? (dot(x - h.p, h.n) > 0.0f ? 1.0f : -1.0f)*h.d : kMaxF; }
	}
}
cudaError_t launchProbePacket(const SceneView& S, int kind, long long n, const float* pts, const float* a0, const float* a1,
							  const float* a2, const float* a3, const float* params, float* d_out, cudaStream_t stream) {
	if (n <= 0) return cudaSuccess;
	unsigned grid = (unsigned)((n + 127)/128);
	if (S.dim == 2) probePacketKernel<2><<<grid, 128, 0, stream>>>(S, kind, n, pts, a0, a1, a2, a3, params, d_out);
	else probePacketKernel<3><<<grid, 128, 0, stream>>>(S, kind, n, pts, a0, a1, a2, a3, params, d_out);
	return cudaGetLastError();
}

} // namespace nmc
