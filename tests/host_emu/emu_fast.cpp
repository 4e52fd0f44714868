// tests/host_emu/emu_fast.cpp -- TEST HARNESS ONLY.  The product's geometry headers compiled for the host in the
// DEFAULT-MODE flavour (NMC_FAST_GEOM: trig-free normal-cone test, fminf/fmaxf, reciprocal divisions, one-FMA source
// lookup), so that the tree queries the default mode uses on meshes beyond the flat-scan limit can be checked against
// the oracle without a GPU.
#define NMC_FAST_GEOM 1
#define NMC_TRAV_INLINE 1
#include <vector>
#include <thread>
#include <atomic>
#include <functional>
#include <pthread.h>
#include <cstring>
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/nmc_math.cuh"
namespace nmc { static inline int min(int a, int b) { return a < b ? a : b; } static inline int max(int a, int b) { return a > b ? a : b; } } // CUDA's integer min/max
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/nmc_geom.cuh"
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/nmc_packet.cuh"
#include "../../neural-monte-carlo-fluid-simulation_b200/csrc/scene_build.h"
using namespace nmc;

// A warp of 32 host threads in lockstep at every cross-lane operation: ballots meet at a barrier, the stack entry sp lives
// with lane sp & 31 as on the device (WarpOps), reading it back is a synchronising shuffle.
struct HostWarpCtx { pthread_barrier_t bar; std::atomic<unsigned> acc{0}; int s0[32], s1[32]; unsigned q0[32], q1[32]; };
static thread_local HostWarpCtx* tlCtx = nullptr;
static thread_local int tlLane = 0;
struct HostWarp {
	static unsigned ballot(bool p) {
		HostWarpCtx& c = *tlCtx;
		if (p) c.acc.fetch_or(1u << tlLane);
		pthread_barrier_wait(&c.bar);
		unsigned r = c.acc.load();
		pthread_barrier_wait(&c.bar);
		if (tlLane == 0) c.acc.store(0);
		pthread_barrier_wait(&c.bar);
		return r;
	}
	static bool any(bool p) { return ballot(p) != 0; }
	static int popc(unsigned m) { return __builtin_popcount(m); }
	static bool mine(unsigned m) { return (m >> tlLane) & 1u; }
	void put(int sp, int node, unsigned lanes) {
		if (tlLane == (sp & 31)) { (sp < 32 ? tlCtx->s0 : tlCtx->s1)[sp & 31] = node; (sp < 32 ? tlCtx->q0 : tlCtx->q1)[sp & 31] = lanes; }
	}
	int get(int sp, unsigned& lanes) const {
		HostWarpCtx& c = *tlCtx;
		pthread_barrier_wait(&c.bar);
		int v = (sp < 32 ? c.s0 : c.s1)[sp & 31];
		lanes = (sp < 32 ? c.q0 : c.q1)[sp & 31];
		pthread_barrier_wait(&c.bar);
		return v;
	}
};
// runs body(lane, packet) on 32 lockstep threads for every packet of 32 consecutive queries
static void runWarp(int nPackets, const std::function<void(int, int)>& body) {
	HostWarpCtx ctx; pthread_barrier_init(&ctx.bar, nullptr, 32);
	std::vector<std::thread> th;
	for (int l = 0; l < 32; l++) th.emplace_back([&, l] { tlCtx = &ctx; tlLane = l; for (int k = 0; k < nPackets; k++) body(l, k); });
	for (auto& t : th) t.join();
	pthread_barrier_destroy(&ctx.bar);
}

struct FastScene { FlatScene flat; SceneView v; std::vector<float> src; };

extern "C" {
void* emuf_scene_create(int dim, const float* verts, int nV, const int* prims, int nP, const float* src, int n0, int n1, int n2,
						float absorption, int watertight, int doubleSided) {
	FastScene* s = new FastScene();
	buildFlatScene(dim, verts, nV, prims, nP, doubleSided != 0, s->flat);
	size_t cnt = (size_t)n0*n1*(dim == 3 ? n2 : 1);
	s->src.assign(src, src + cnt);
	SceneView& v = s->v; memset(&v, 0, sizeof(v));
	v.dim = dim; v.nNodes = s->flat.nNodes; v.nPrims = s->flat.nPrims; v.nSilRefs = s->flat.nSilRefs;
	v.nodes = (const float4*)s->flat.nodes.data(); v.coneF = (const float4*)s->flat.coneF.data(); v.silsF = (const float4*)s->flat.silsF.data(); v.treeF = (const float4*)s->flat.treeF.data(); v.prims = (const float4*)s->flat.prims.data();
	v.primN = (const float4*)s->flat.primN.data(); v.nrmV = (const float4*)s->flat.nrmV.data(); v.sils = (const float4*)s->flat.sils.data();
	for (int k = 0; k < 3; k++) { v.bboxLo[k] = s->flat.bboxLo[k]; v.bboxHi[k] = s->flat.bboxHi[k]; }
	v.src = s->src.data(); v.n0 = n0; v.n1 = n1; v.n2 = dim == 3 ? n2 : 1;
	v.absorption = absorption; v.watertight = watertight; v.doubleSided = doubleSided;
	// as csrc/capi.cu setSource: texel index along box axis k = (int)(x_k*srcScale[k] + srcOff[k])
	const int nAx[3] = {dim == 2 ? n1 : n0, dim == 2 ? n0 : n1, dim == 3 ? n2 : 1};
	for (int k = 0; k < 3; k++) {
		float ext = v.bboxHi[k] - v.bboxLo[k];
		v.srcScale[k] = ext > 0.0f ? (float)nAx[k]/ext : 0.0f;
		v.srcOff[k] = -v.bboxLo[k]*v.srcScale[k];
	}
	return s;
}
void emuf_scene_destroy(void* h) { delete (FastScene*)h; }
void emuf_star_radius(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flip, float* out) {
	FastScene* s = (FastScene*)h;
	for (int i = 0; i < n; i++) {
		if (s->v.dim == 2) out[i] = starRadius<2, FastMath>(s->v, mk(pts[2*i], pts[2*i + 1], 0.0f), minR, maxR[i], prec, flip != 0);
		else out[i] = starRadius<3, FastMath>(s->v, mk(pts[3*i], pts[3*i + 1], pts[3*i + 2]), minR, maxR[i], prec, flip != 0);
	}
}
void emuf_source(void* h, const float* pts, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 x = mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f);
		out[i] = D == 2 ? sourceAt<2>(s->v, x) : sourceAt<3>(s->v, x);
	}
}
// out per ray: hit, distance
void emuf_rays(void* h, const float* o, const float* d, const float* tmax, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 ro = mk(o[D*i], o[D*i + 1], D == 3 ? o[D*i + 2] : 0.0f), dir = mk(d[D*i], d[D*i + 1], D == 3 ? d[D*i + 2] : 0.0f);
		Hit b; b.d = kMaxF; b.p = mk(0, 0, 0); b.n = mk(0, 0, 0);
		bool hb = D == 2 ? intersectNeumann<2>(s->v, ro, mk(0, 0, 0), dir, tmax[i], false, b) : intersectNeumann<3>(s->v, ro, mk(0, 0, 0), dir, tmax[i], false, b);
		out[2*i] = hb; out[2*i + 1] = hb ? b.d : 0.0f;
	}
}
// ---- the warp-packet traversals (csrc/nmc_packet.cuh) as packets of one query: the wiring of wost_fast.cu's FLAT == 3 branch
void emuf_star_radius_packet(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flipOrient, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 x = mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f);
		if (minR > maxR[i]) { out[i] = maxR[i]; continue; }
		float d = 0.0f, r2 = maxR[i] < kMaxF ? maxR[i]*maxR[i] : kMaxF;
		bool f = D == 2 ? packetClosestSilhouette<2, HostLane>(s->v, x, r2, !flipOrient, minR*minR, prec, d)
						: packetClosestSilhouette<3, HostLane>(s->v, x, r2, !flipOrient, minR*minR, prec, d);
		out[i] = f ? maxS(d, minR) : maxS(maxR[i], minR);
	}
}
void emuf_rays_packet(void* h, const float* o, const float* d, const float* tmax, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 ro = mk(o[D*i], o[D*i + 1], D == 3 ? o[D*i + 2] : 0.0f), dir = mk(d[D*i], d[D*i + 1], D == 3 ? d[D*i + 2] : 0.0f);
		Hit b; b.d = kMaxF; b.p = mk(0, 0, 0); b.n = mk(0, 0, 0);
		bool hb = D == 2 ? packetRay<2, HostLane>(s->v, ro, dir, tmax[i], b) : packetRay<3, HostLane>(s->v, ro, dir, tmax[i], b);
		out[2*i] = hb; out[2*i + 1] = hb ? b.d : 0.0f;
	}
}
// out per point: unsigned distance, signed distance (pseudo-normal side)
void emuf_closest_packet(void* h, const float* pts, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	for (int i = 0; i < n; i++) {
		V3 x = mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f);
		Hit b; b.d = kMaxF; b.p = x; b.n = mk(0, 0, 0);
		bool f = D == 2 ? packetClosestPoint<2, HostLane>(s->v, x, kMaxF, true, b) : packetClosestPoint<3, HostLane>(s->v, x, kMaxF, true, b);
		out[2*i] = f ? b.d : kMaxF;
		out[2*i + 1] = f ? (dot(x - b.p, b.n) > 0.0f ? 1.0f : -1.0f)*b.d : kMaxF;
	}
}
// ---- the same queries as packets of 32 (lanes beyond n idle, as the kernel's finished lanes)
void emuf_star_radius_warp(void* h, const float* pts, int n, float minR, const float* maxR, float prec, int flipOrient, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	runWarp((n + 31)/32, [&](int lane, int k) {
		const int i = 32*k + lane; const bool on = i < n && !(minR > maxR[i]);
		V3 x = i < n ? mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f) : mk(0, 0, 0);
		float d = 0.0f, r2 = on ? (maxR[i] < kMaxF ? maxR[i]*maxR[i] : kMaxF) : -1.0f;
		bool f = D == 2 ? packetClosestSilhouette<2, HostWarp>(s->v, x, r2, !flipOrient, minR*minR, prec, d)
						: packetClosestSilhouette<3, HostWarp>(s->v, x, r2, !flipOrient, minR*minR, prec, d);
		if (i < n) out[i] = !on ? maxR[i] : f ? maxS(d, minR) : maxS(maxR[i], minR);
	});
}
void emuf_rays_warp(void* h, const float* o, const float* d, const float* tmax, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	runWarp((n + 31)/32, [&](int lane, int k) {
		const int i = 32*k + lane;
		V3 ro = mk(0, 0, 0), dir = mk(0, 0, 0);
		if (i < n) { ro = mk(o[D*i], o[D*i + 1], D == 3 ? o[D*i + 2] : 0.0f); dir = mk(d[D*i], d[D*i + 1], D == 3 ? d[D*i + 2] : 0.0f); }
		Hit b; b.d = kMaxF; b.p = mk(0, 0, 0); b.n = mk(0, 0, 0);
		bool hb = D == 2 ? packetRay<2, HostWarp>(s->v, ro, dir, i < n ? tmax[i] : -1.0f, b) : packetRay<3, HostWarp>(s->v, ro, dir, i < n ? tmax[i] : -1.0f, b);
		if (i < n) { out[2*i] = hb; out[2*i + 1] = hb ? b.d : 0.0f; }
	});
}
void emuf_closest_warp(void* h, const float* pts, int n, float* out) {
	FastScene* s = (FastScene*)h;
	const int D = s->v.dim;
	runWarp((n + 31)/32, [&](int lane, int k) {
		const int i = 32*k + lane;
		V3 x = i < n ? mk(pts[D*i], pts[D*i + 1], D == 3 ? pts[D*i + 2] : 0.0f) : mk(0, 0, 0);
		Hit b; b.d = kMaxF; b.p = x; b.n = mk(0, 0, 0);
		bool f = D == 2 ? packetClosestPoint<2, HostWarp>(s->v, x, i < n ? kMaxF : -1.0f, true, b) : packetClosestPoint<3, HostWarp>(s->v, x, i < n ? kMaxF : -1.0f, true, b);
		if (i < n) { out[2*i] = f ? b.d : kMaxF; out[2*i + 1] = f ? (dot(x - b.p, b.n) > 0.0f ? 1.0f : -1.0f)*b.d : kMaxF; }
	});
}
// builder tables for the cone test: nodes (16 floats each), the default mode's cones (4 floats per node), silhouette records
int emuf_num_nodes(void* h) { return ((FastScene*)h)->flat.nNodes; }
int emuf_num_sil_refs(void* h) { return ((FastScene*)h)->flat.nSilRefs; }
void emuf_tables(void* h, float* nodes, float* cones, float* sils) {
	FastScene* s = (FastScene*)h;
	memcpy(nodes, s->flat.nodes.data(), s->flat.nodes.size()*sizeof(Q4));
	memcpy(cones, s->flat.coneF.data(), s->flat.coneF.size()*sizeof(Q4));
	memcpy(sils, s->flat.sils.data(), s->flat.sils.size()*sizeof(Q4));
}
}
