// tests/host_emu/emu.cpp -- TEST HARNESS ONLY (never shipped, never loaded by the product).
// Compiles the product's device headers (csrc/*.cuh) for the HOST with g++ so that their logic can be
// checked against the oracle in the CPU-only container before any GPU time is spent.  libnmcfs.so
// itself contains no host execution path for these functions.
#include <vector>
#include <cstring>
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/nmc_estimator.cuh"
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/scene_build.h"
using namespace nmc;

struct EmuScene { FlatScene flat; SceneView v; std::vector<float> src; };

template <int DIM>
static void wostT(EmuScene* s, const SolverParams& o, const float* pts, int n, uint64_t off, float* p, float* g, float* st) {
	std::vector<float> scratch((size_t)4*(o.nWalks + 2));
	for (int i = 0; i < n; i++) {
		V3 x = mk(pts[i*DIM], pts[i*DIM + 1], DIM == 3 ? pts[i*DIM + 2] : 0.0f);
		LhsScratch sc; sc.base = scratch.data(); sc.stride = 1;
		PointResult r;
		detEstimatePoint<DIM>(s->v, o, x, off + i, sc, r);
		p[i] = r.p; for (int k = 0; k < DIM; k++) g[i*DIM + k] = r.g[k];
		if (st) { float* t = st + (size_t)i*12; int nv = r.nSol - 1 > 1 ? r.nSol - 1 : 1;
			t[0] = r.solMean; t[1] = r.solM2/nv; for (int k = 0; k < 3; k++) { t[2 + k] = k < DIM ? r.gradMean[k] : 0; t[5 + k] = k < DIM ? r.gradM2[k]/nv : 0; }
			t[8] = r.meanFirstSource; t[9] = (float)r.nSol; t[10] = (float)r.totalWalkLength/(r.nSol > 1 ? r.nSol : 1); t[11] = (float)r.active;
			if (!r.active) for (int k = 0; k < 11; k++) t[k] = 0; }
	}
}

extern "C" {
void* emu_scene_create(int dim, const float* verts, int nV, const int* prims, int nP, const float* src, int n0, int n1, int n2,
					   float absorption, int watertight, int doubleSided) {
	EmuScene* s = new EmuScene();
	buildFlatScene(dim, verts, nV, prims, nP, doubleSided != 0, s->flat);
	size_t cnt = (size_t)n0*n1*(dim == 3 ? n2 : 1);
	s->src.assign(src, src + cnt);
	SceneView& v = s->v; memset(&v, 0, sizeof(v));
	v.dim = dim; v.nNodes = s->flat.nNodes; v.nPrims = s->flat.nPrims; v.nSilRefs = s->flat.nSilRefs;
	v.nodes = (const float4*)s->flat.nodes.data(); v.prims = (const float4*)s->flat.prims.data();
	v.primN = (const float4*)s->flat.primN.data(); v.nrmV = (const float4*)s->flat.nrmV.data(); v.sils = (const float4*)s->flat.sils.data();
	for (int k = 0; k < 3; k++) { v.bboxLo[k] = s->flat.bboxLo[k]; v.bboxHi[k] = s->flat.bboxHi[k]; }
	v.src = s->src.data(); v.n0 = n0; v.n1 = n1; v.n2 = dim == 3 ? n2 : 1;
	v.absorption = absorption; v.watertight = watertight; v.doubleSided = doubleSided;
	return s;
}
void emu_scene_destroy(void* h) { delete (EmuScene*)h; }
int emu_num_nodes(void* h) { return ((EmuScene*)h)->flat.nNodes; }
void emu_nodes(void* h, float* out) {
	EmuScene* s = (EmuScene*)h;
	for (int i = 0; i < s->flat.nNodes; i++) {
		const Q4* q = &s->flat.nodes[(size_t)4*i]; float* o = out + (size_t)i*16;
		o[0] = q[0].x; o[1] = q[0].y; o[2] = q[0].z; o[3] = q[1].x; o[4] = q[1].y; o[5] = q[1].z;
		o[6] = q[2].x; o[7] = q[2].y; o[8] = q[2].z; o[9] = q[2].w;
		o[10] = (float)asInt(q[3].x); o[11] = (float)asInt(q[3].y); o[12] = (float)asInt(q[0].w); o[13] = (float)asInt(q[3].z);
		o[14] = (float)asInt(q[1].w); o[15] = 0;
	}
}
void emu_wost(void* h, const SolverParams* o, const float* pts, int n, uint64_t off, float* p, float* g, float* st) {
	EmuScene* s = (EmuScene*)h;
	if (s->v.dim == 2) wostT<2>(s, *o, pts, n, off, p, g, st); else wostT<3>(s, *o, pts, n, off, p, g, st);
}
// fast-mode ball functions: out per entry: T(x) g(x) normG exitThroughput bdyGradFactor srcGradFactor(x)
void emu_ball_fast(int dim, float lambda, const float* R, const float* r, int n, float* out) {
	for (int i = 0; i < n; i++) {
		float* o = out + (size_t)i*6;
		if (dim == 2) { BallFast<2> b; b.init(lambda > 0, lambda); b.update(R[i]); float x = b.yukawa ? r[i]*b.mu : r[i]/R[i]; float T = 1, g = 0, q = 0; if (b.yukawa) b.evalTgq(x, T, g, q);
			o[0] = T; o[1] = g; o[2] = b.normG(); o[3] = b.exitThroughput(); o[4] = b.bdyGradFactor(); o[5] = b.yukawa ? b.srcGradFactor(q, g) : b.srcGradFactorHarmonic(x); }
		else { BallFast<3> b; b.init(lambda > 0, lambda); b.update(R[i]); float x = b.yukawa ? r[i]*b.mu : r[i]/R[i]; float T = 1, g = 0, q = 0; if (b.yukawa) b.evalTgq(x, T, g, q);
			o[0] = T; o[1] = g; o[2] = b.normG(); o[3] = b.exitThroughput(); o[4] = b.bdyGradFactor(); o[5] = b.yukawa ? b.srcGradFactor(q, g) : b.srcGradFactorHarmonic(x); }
	}
}
// inverse-CDF sampler: out r per entry
void emu_sample_fast(int dim, float lambda, const float* R, const float* u, const float* u2, int n, float* out) {
	for (int i = 0; i < n; i++) {
		float g, q; bool hf;
		if (dim == 2) { BallFast<2> b; b.init(lambda > 0, lambda); b.update(R[i]); float x = b.sampleX(u[i], u2[i], g, q, hf); out[i] = hf ? x*R[i] : x/b.mu; }
		else { BallFast<3> b; b.init(lambda > 0, lambda); b.update(R[i]); float x = b.sampleX(u[i], u2[i], g, q, hf); out[i] = hf ? x*R[i] : x/b.mu; }
	}
}
}

// star radius through the product's traversal (LocalStack), for both geometry flavours of the headers
extern "C" void emu_star_radius(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flip, float* out) {
	EmuScene* s = (EmuScene*)h;
	for (int i = 0; i < n; i++) {
		if (s->v.dim == 2) out[i] = starRadius<2, FastMath>(s->v, mk(pts[2*i], pts[2*i + 1], 0.0f), minR, maxR[i], prec, flip != 0);
		else out[i] = starRadius<3, FastMath>(s->v, mk(pts[3*i], pts[3*i + 1], pts[3*i + 2]), minR, maxR[i], prec, flip != 0);
	}
}

// sizes of the flattened structure: nodes, prims, silhouette refs, distinct silhouettes, ray primitives, depth
extern "C" void emu_scene_info(void* h, int* out) {
	EmuScene* s = (EmuScene*)h;
	out[0] = s->flat.nNodes; out[1] = s->flat.nPrims; out[2] = s->flat.nSilRefs; out[3] = s->flat.nSilU; out[4] = s->flat.nRay; out[5] = s->flat.maxDepth;
}

// default-mode flat scans (nmc_geom.cuh: flatClosestSilhouette / flatRay) on the host, with the kernel's calling
// convention (wost_fast.cu, phase 1), next to the tree traversals they replace
static FlatTab flatTab(EmuScene* s) {
	FlatTab F;
	F.silsU = (const float4*)s->flat.silsU.data(); F.grpS = (const float4*)s->flat.grpS.data(); F.nSilU = s->flat.nSilU;
	F.rayP = (const float4*)s->flat.rayP.data(); F.rayN = (const float4*)s->flat.rayN.data(); F.grpP = (const float4*)s->flat.grpP.data(); F.nRay = s->flat.nRay;
	return F;
}
extern "C" void emu_flat_star_radius(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flipOrient, float* out) {
	EmuScene* s = (EmuScene*)h;
	FlatTab F = flatTab(s);
	for (int i = 0; i < n; i++) {
		const int D = s->v.dim;
		V3 pt = mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f);
		float dd = maxR[i], starR = dd;
		if (minR <= dd) {
			float dsil = 0.0f, r2max = dd < kMaxF ? dd*dd : kMaxF;
			bool f = D == 2 ? flatClosestSilhouette<2>(F, pt, r2max, flipOrient == 0, minR*minR, prec, dsil)
							: flatClosestSilhouette<3>(F, pt, r2max, flipOrient == 0, minR*minR, prec, dsil);
			starR = f ? fmaxf(dsil, minR) : fmaxf(dd, minR);
		}
		out[i] = starR;
	}
}
// out per ray: hit, distance, normal.xyz  (both variants)
extern "C" void emu_rays(void* h, const float* o, const float* d, const float* tmax, int n, float* outFlat, float* outTree) {
	EmuScene* s = (EmuScene*)h;
	FlatTab F = flatTab(s);
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 ro = mk(o[D*i], o[D*i + 1], D == 3 ? o[D*i + 2] : 0.0f), dir = mk(d[D*i], d[D*i + 1], D == 3 ? d[D*i + 2] : 0.0f);
		Hit a; a.d = kMaxF; a.p = mk(0, 0, 0); a.n = mk(0, 0, 0);
		Hit b = a;
		bool ha = D == 2 ? flatRay<2>(F, ro, dir, tmax[i], a) : flatRay<3>(F, ro, dir, tmax[i], a);
		bool hb = D == 2 ? intersectNeumann<2>(s->v, ro, mk(0, 0, 0), dir, tmax[i], false, b) : intersectNeumann<3>(s->v, ro, mk(0, 0, 0), dir, tmax[i], false, b);
		float* x = outFlat + 5*i; float* y = outTree + 5*i;
		x[0] = ha; x[1] = ha ? a.d : 0.0f; x[2] = ha ? a.n.x : 0.0f; x[3] = ha ? a.n.y : 0.0f; x[4] = ha ? a.n.z : 0.0f;
		y[0] = hb; y[1] = hb ? b.d : 0.0f; y[2] = hb ? b.n.x : 0.0f; y[3] = hb ? b.n.y : 0.0f; y[4] = hb ? b.n.z : 0.0f;
	}
}

// the same with the walk's on-boundary convention: origin offset against the normal (wost_fast.cu phase 1 /
// intersectNeumann); out additionally carries the hit point
extern "C" void emu_rays_onb(void* h, const float* o, const float* nrm, const float* d, const float* tmax, int n, float* outFlat, float* outTree) {
	EmuScene* s = (EmuScene*)h;
	FlatTab F = flatTab(s);
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 pt = mk(o[D*i], o[D*i + 1], D == 3 ? o[D*i + 2] : 0.0f), dir = mk(d[D*i], d[D*i + 1], D == 3 ? d[D*i + 2] : 0.0f);
		V3 nn = mk(nrm[D*i], nrm[D*i + 1], D == 3 ? nrm[D*i + 2] : 0.0f);
		Hit a; a.d = kMaxF; a.p = mk(0, 0, 0); a.n = mk(0, 0, 0);
		Hit b = a;
		V3 ro = D == 2 ? offsetPoint<2>(pt, neg(nn)) : offsetPoint<3>(pt, neg(nn));
		bool ha = D == 2 ? flatRay<2>(F, ro, dir, tmax[i], a) : flatRay<3>(F, ro, dir, tmax[i], a);
		bool hb = D == 2 ? intersectNeumann<2>(s->v, pt, nn, dir, tmax[i], true, b) : intersectNeumann<3>(s->v, pt, nn, dir, tmax[i], true, b);
		float* x = outFlat + 8*i; float* y = outTree + 8*i;
		x[0] = ha; x[1] = ha ? a.d : 0.0f; x[2] = a.n.x; x[3] = a.n.y; x[4] = a.n.z; x[5] = a.p.x; x[6] = a.p.y; x[7] = a.p.z;
		y[0] = hb; y[1] = hb ? b.d : 0.0f; y[2] = b.n.x; y[3] = b.n.y; y[4] = b.n.z; y[5] = b.p.x; y[6] = b.p.y; y[7] = b.p.z;
	}
}

extern "C" void emu_dist_neumann(void* h, const float* pts, int n, int signedDist, float* out) {
	EmuScene* s = (EmuScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 x = mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f);
		out[i] = D == 2 ? distNeumann<2>(s->v, x, signedDist != 0) : distNeumann<3>(s->v, x, signedDist != 0);
	}
}

// keyed permutation of the default mode's Latin-hypercube strata (nmc_math.cuh: permute)
extern "C" void emu_permute(unsigned n, unsigned key, unsigned* out) { for (unsigned i = 0; i < n; i++) out[i] = permute(i, n, key); }
