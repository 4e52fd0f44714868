// csrc/nmc_packet.cuh -- warp-packet traversals of the flattened BVH / SNCH (default mode, meshes beyond the flat-scan limit).
//
// The lanes of a warp work on walks of ONE query point (wost_fast.cu), so their queries sit within a ball or two of each
// other.  Instead of 32 private traversals (32 stacks, 32 divergent node fetches per trip, leaf and inner-node code
// serialised) the warp walks the tree once: a node is entered when ANY lane can still improve inside it, every node record is
// fetched with one warp-uniform load (a single L1 transaction, broadcast), the stack (node index + the lanes that asked for
// the node) lives in registers spread over the lanes (entry i in lane i & 31, read back with a shuffle), and the two children
// are ordered by a vote of the lanes that reach both.  Every lane keeps its own search radius / ray length and only works on
// nodes its own tests selected, so each lane gets exactly the result of its private traversal (Sbvh::findClosestSilhouettePointFromNode sbvh.inl:1093-1255, intersectFromNode :538-683,
// findClosestPointFromNode :948-1074) -- only ties between equidistant records may resolve differently.
// The reference's own wide traversal (mbvh.inl:702-818, 1661-1810) vectorises over the four children of one query; here the
// vector lanes are the queries.
//
// All functions must be called by the whole (converged) warp; a lane without a query passes r2 < 0 / tMax < 0.
// Policies: `W` supplies the cross-lane operations and the stack (WarpOps on the device; HostLane = a packet of one query,
// used by tests/host_emu to pin the traversal logic to the oracle without a GPU).
#pragma once
#include "nmc_geom.cuh"

namespace nmc {

#if defined(__CUDACC__)
struct WarpOps {
	// stack entry sp (node index + the lanes that asked for it) lives in lane sp & 31: first / second 32 slots (NMC_STACK = 64 bounds the depth)
	int s0 = 0, s1 = 0; unsigned q0 = 0, q1 = 0;
	__device__ __forceinline__ static bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
	__device__ __forceinline__ static unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
	__device__ __forceinline__ static int popc(unsigned m) { return __popc(m); }
	__device__ __forceinline__ static bool mine(unsigned m) { return (m >> (threadIdx.x & 31)) & 1u; }
	__device__ __forceinline__ void put(int sp, int node, unsigned lanes) {
		const bool here = (int)(threadIdx.x & 31) == (sp & 31);
		if (sp < 32) { s0 = here ? node : s0; q0 = here ? lanes : q0; } else { s1 = here ? node : s1; q1 = here ? lanes : q1; }
	}
	__device__ __forceinline__ int get(int sp, unsigned& lanes) const {
		lanes = __shfl_sync(0xffffffffu, sp < 32 ? q0 : q1, sp & 31);
		return __shfl_sync(0xffffffffu, sp < 32 ? s0 : s1, sp & 31);
	}
};
#define NMC_PK __device__ __forceinline__
#define NMC_PKI __device__ __forceinline__
#else
#define NMC_PK inline
#define NMC_PKI inline
#endif
struct HostLane {
	int st[NMC_STACK + 2]; unsigned q[NMC_STACK + 2];
	static bool any(bool p) { return p; }
	static unsigned ballot(bool p) { return p ? 1u : 0u; }
	static int popc(unsigned m) { return (int)(m & 1u); }
	static bool mine(unsigned m) { return m & 1u; }
	void put(int sp, int node, unsigned lanes) { st[sp] = node; q[sp] = lanes; }
	int get(int sp, unsigned& lanes) const { lanes = q[sp]; return st[sp]; }
};

// pushes the children the packet still needs, the one most lanes prefer on top.  Every entry carries the lanes that asked for
// it: a lane works on a node only if its OWN tests led there, exactly as in its private traversal -- it matters for the
// silhouette query, where the cone test also drops records the leaf test would accept through its precision band.
template <class W>
NMC_PKI void packetPush(W& w, int& sp, int c0, int c1, bool hit0, bool hit1, bool prefer1) {
	const unsigned m0 = W::ballot(hit0), m1 = W::ballot(hit1);
	if (m0 && m1) {
		// lanes that reach one child only vote for it
		const unsigned v1 = W::ballot(hit1 && (!hit0 || prefer1));
		const bool first1 = 2*W::popc(v1) > W::popc(m0 | m1);
		w.put(++sp, first1 ? c0 : c1, first1 ? m0 : m1);
		w.put(++sp, first1 ? c1 : c0, first1 ? m1 : m0);
	} else if (m0) w.put(++sp, c0, m0);
	else if (m1) w.put(++sp, c1, m1);
}

// closest silhouette point within sqrt(r2) of x; false when there is none (dOut untouched)
template <int DIM, class W>
NMC_PK bool packetClosestSilhouette(const SceneView& S, V3 x, float r2, bool flip, float sqMinR, float precision, float& dOut) {
	if (S.nNodes == 0) return false;
	W w;
	bool live = r2 >= 0.0f && sqMinR < r2; // a lane leaves the search for good once its radius is down to minRadius
	if (!live) r2 = -1.0f;
	float b0, b1, tmp;
	bool found = false; int lastId = -1;
	int sp = 0;
	w.put(0, 0, 0xffffffffu);
	while (sp >= 0) {
		unsigned asked;
		const int ni = w.get(sp, asked); sp--;
		const float4 na = S.nodes[4*ni], nb = S.nodes[4*ni + 1];
		boxSqDist(xyz(na), xyz(nb), x, b0, tmp);
		const bool here = W::mine(asked) && live && b0 <= r2; // the radius may have shrunk since the node was pushed
		if (!W::any(here)) continue;
		const int nRefs = asInt(na.w);
		if (nRefs > 0) {
			const float4 nd = S.nodes[4*ni + 3];
			const int silOffset = asInt(nd.y), nSil = asInt(nd.z);
			for (int p = 0; p < nSil; p++) {
				const int ri = silOffset + p;
				V3 viewDir, n0, n1; float d, concavity; int flags, id;
				if (DIM == 2) {
					const float4 s0 = S.sils[2*ri], s1 = S.sils[2*ri + 1];
					flags = asInt(s0.z); id = asInt(s0.w);
					viewDir = x - mk(s0.x, s0.y, 0.0f);
					if (!here || !live || id == lastId || dot(viewDir, viewDir) > r2) continue; // reject on the squared distance before paying for the sqrt
					d = norm(viewDir);
					n0 = mk(s1.x, s1.y, 0.0f); n1 = mk(s1.z, s1.w, 0.0f);
					concavity = n0.x*n1.y - n1.x*n0.y;
				} else {
					const float4 s0 = S.sils[4*ri], s1 = S.sils[4*ri + 1], s2 = S.sils[4*ri + 2];
					flags = asInt(s0.w); id = asInt(s1.w);
					n0 = xyz(s2); concavity = s2.w; n1 = xyz(S.sils[4*ri + 3]);
					if (!here || !live || id == lastId) continue;
					if ((flags & 3) == 3 && !silhouetteCandidate(n0, n1, x - xyz(s0), precision, r2)) continue;
					V3 pt; float t;
					d = closestOnSegment(xyz(s0), xyz(s1), x, pt, t);
					viewDir = x - pt;
				}
				if (d*d > r2) continue;
				bool isSil = (flags & 3) != 3;
				if (!isSil) isSil = isSilhouette(concavity, n0, n1, viewDir, d, flip, precision);
				if (isSil) {
					found = true;
					r2 = minS(r2, d*d);
					dOut = d; lastId = id;
					live = sqMinR < r2;
				}
			}
		} else {
			const int c0 = ni + 1, c1 = ni + asInt(nb.w);
			bool hit0 = false, hit1 = false;
			b0 = b1 = kMaxF;
			const float4 k0 = S.coneF[c0];
			if (k0.w != 2.0f) { // the subtree holds silhouettes
				const V3 lo = xyz(S.nodes[4*c0]), hi = xyz(S.nodes[4*c0 + 1]);
				boxSqDist(lo, hi, x, b0, tmp);
				hit0 = here && live && b0 <= r2 && coneOverlapFast(xyz(k0), k0.w, x, lo, hi, b0, 2.0f*precision);
			}
			const float4 k1 = S.coneF[c1];
			if (k1.w != 2.0f) {
				const V3 lo = xyz(S.nodes[4*c1]), hi = xyz(S.nodes[4*c1 + 1]);
				boxSqDist(lo, hi, x, b1, tmp);
				hit1 = here && live && b1 <= r2 && coneOverlapFast(xyz(k1), k1.w, x, lo, hi, b1, 2.0f*precision);
			}
			packetPush(w, sp, c0, c1, hit0, hit1, b1 < b0);
		}
	}
	return found;
}

// closest hit of the ray segment [0, tMax]; out is written for lanes that return true
template <int DIM, class W>
NMC_PK bool packetRay(const SceneView& S, V3 o, V3 dir, float tMax, Hit& out) {
	if (S.nNodes == 0) return false;
	W w;
	const V3 invD = mk(1.0f/dir.x, 1.0f/dir.y, 1.0f/dir.z);
	float b0, b1, b2, b3;
	bool hitAny = false;
	int sp = 0;
	w.put(0, 0, 0xffffffffu);
	while (sp >= 0) {
		unsigned asked;
		const int ni = w.get(sp, asked); sp--;
		const float4 na = S.nodes[4*ni], nb = S.nodes[4*ni + 1];
		const bool reach = W::mine(asked) && tMax >= 0.0f && boxRay(xyz(na), xyz(nb), o, invD, tMax, b0, b1);
		if (!W::any(reach)) continue;
		const int nRefs = asInt(na.w);
		if (nRefs > 0) {
			const int refOffset = asInt(S.nodes[4*ni + 3].x);
			for (int p = 0; p < nRefs; p++) {
				Hit h;
				if (reach && primRay<DIM>(S, refOffset + p, o, dir, tMax, false, h)) {
					hitAny = true;
					tMax = minS(tMax, h.d);
					out = h;
				}
			}
		} else {
			const int c0 = ni + 1, c1 = ni + asInt(nb.w);
			b0 = b2 = kMaxF;
			const bool hit0 = reach && boxRay(xyz(S.nodes[4*c0]), xyz(S.nodes[4*c0 + 1]), o, invD, tMax, b0, b1);
			const bool hit1 = reach && boxRay(xyz(S.nodes[4*c1]), xyz(S.nodes[4*c1 + 1]), o, invD, tMax, b2, b3);
			packetPush(w, sp, c0, c1, hit0, hit1, b2 < b0);
		}
	}
	return hitAny;
}

// closest point on the mesh within sqrt(r2); wantNormal: the pseudo-normal of closestPoint() (nmc_geom.cuh)
template <int DIM, class W>
NMC_PK bool packetClosestPoint(const SceneView& S, V3 x, float r2, bool wantNormal, Hit& out) {
	if (S.nNodes == 0) return false;
	W w;
	float b0, b1, b2, b3;
	bool found = false;
	out.d = kMaxF; out.ref = -1; out.u = 0.0f; out.v = 0.0f;
	{
		boxSqDist(xyz(S.nodes[0]), xyz(S.nodes[1]), x, b0, b1);
		if (r2 >= 0.0f && b0 <= r2) r2 = minS(r2, b1);
	}
	int sp = 0;
	w.put(0, 0, 0xffffffffu);
	while (sp >= 0) {
		unsigned asked;
		const int ni = w.get(sp, asked); sp--;
		const float4 na = S.nodes[4*ni], nb = S.nodes[4*ni + 1];
		boxSqDist(xyz(na), xyz(nb), x, b0, b1);
		const bool reach = W::mine(asked) && b0 <= r2; // r2 < 0: never
		if (!W::any(reach)) continue;
		const int nRefs = asInt(na.w);
		if (nRefs > 0) {
			const int refOffset = asInt(S.nodes[4*ni + 3].x);
			for (int p = 0; p < nRefs; p++) {
				const int ri = refOffset + p;
				V3 pt; float u = 0.0f, v = 0.0f, d;
				if (DIM == 2) {
					const float4 q = S.prims[ri];
					d = closestOnSegment(mk(q.x, q.y, 0.0f), mk(q.z, q.w, 0.0f), x, pt, u); v = -1.0f;
				} else {
					d = closestOnTriangle(xyz(S.prims[3*ri]), xyz(S.prims[3*ri + 1]), xyz(S.prims[3*ri + 2]), x, pt, u, v);
				}
				if (reach && d*d <= r2) {
					found = true;
					r2 = minS(r2, d*d);
					out.d = d; out.p = pt; out.u = u; out.v = v; out.ref = ri;
				}
			}
		} else {
			const int c0 = ni + 1, c1 = ni + asInt(nb.w);
			// every box holds a primitive: its farthest corner bounds the answer whether or not the lane descends (as closestPoint())
			boxSqDist(xyz(S.nodes[4*c0]), xyz(S.nodes[4*c0 + 1]), x, b0, b1);
			const bool hit0 = reach && b0 <= r2;
			r2 = minS(r2, b1);
			boxSqDist(xyz(S.nodes[4*c1]), xyz(S.nodes[4*c1 + 1]), x, b2, b3);
			const bool hit1 = reach && b2 <= r2;
			r2 = minS(r2, b3);
			bool prefer1 = b2 < b0;
			if (b0 == 0.0f && b2 == 0.0f) prefer1 = b3 < b1;
			packetPush(w, sp, c0, c1, hit0, hit1, prefer1);
		}
	}
	if (found && wantNormal) {
		const int ri = out.ref;
		if (DIM == 2) {
			int vi = -1;
			if (out.u <= kEps) vi = 0; else if (out.u >= 1.0f - kEps) vi = 1;
			out.n = vi >= 0 ? xyz(S.nrmV[2*ri + vi]) : xyz(S.primN[ri]);
		} else {
			const float ome = 1.0f - kEps;
			int vi = -1;
			if (out.u >= ome && out.v <= kEps) vi = 0;
			else if (out.u <= kEps && out.v >= ome) vi = 1;
			else if (out.u <= kEps && out.v <= kEps) vi = 2;
			int ei = -1;
			if (vi == -1) {
				if (out.u <= kEps) ei = 1;
				else if (out.v <= kEps) ei = 2;
				else if (out.u + out.v >= ome) ei = 0;
			}
			out.n = vi >= 0 ? xyz(S.nrmV[6*ri + vi]) : (ei >= 0 ? xyz(S.nrmV[6*ri + 3 + ei]) : xyz(S.primN[ri]));
		}
	}
	return found;
}

} // namespace nmc
