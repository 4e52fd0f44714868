// csrc/scene_build.cpp -- builds the flattened BVH + normal-cone hierarchy ("SNCH") that the walk
// kernels traverse.
//
// The tree TOPOLOGY and the silhouette bookkeeping deliberately follow FCPW's scalar builder,
// because closest-point / ray / silhouette results depend on traversal order whenever two
// candidates tie, and the deterministic mode must return what the reference returns:
//   object-split build, 8 centroid buckets, OverlapSurfaceArea cost, leaf size 4
//     (deps/fcpw/include/fcpw/aggregates/sbvh.inl:4-234, fcpw.inl:523-573 without FCPW_USE_ENOKI)
//   vertex / edge pseudo-normals (fcpw.inl:300-354), silhouette vertices / edges (fcpw.inl:224-291),
//   per-leaf silhouette references filtered by ignoreCandidateSilhouette (sbvh.inl:314-443,
//   demo/scene.h:84-90), bounding cones (sbvh.inl:236-301),
//   2D vertex renumbering and its double wiring of silhouette vertices (fcpw.inl:374-410, 469-490).
// The MEMORY LAYOUT is ours: fixed 16-byte records with pre-gathered vertex data, pre-normalised
// face normals and pre-evaluated dihedral angles, so the device never chases an index.
#include "scene_build.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <map>
#include <array>

namespace nmc {
namespace {

struct P3 { float x, y, z; };
inline P3 operator+(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 operator*(P3 a, float s) { return {a.x*s, a.y*s, a.z*s}; }
inline float dot(P3 a, P3 b) { return a.x*b.x + (a.y*b.y + a.z*b.z); } // Eigen redux order
inline P3 cross(P3 a, P3 b) { return {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
inline P3 unit(P3 a) { float z = dot(a, a); if (z > 0.0f) { float s = std::sqrt(z); return {a.x/s, a.y/s, a.z/s}; } return a; }
inline float at(const P3& a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
inline float mn(float a, float b) { return (b < a) ? b : a; }
inline float mx(float a, float b) { return (a < b) ? b : a; }
inline float bits(int i) { float f; std::memcpy(&f, &i, 4); return f; }

struct Aabb {
	P3 lo{FLT_MAX, FLT_MAX, FLT_MAX}, hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
	void grow(P3 p) { // BoundingBox::expandToInclude(point), bounding_volumes.h:45-49
		const float e = FLT_EPSILON;
		lo = {mn(lo.x, p.x - e), mn(lo.y, p.y - e), mn(lo.z, p.z - e)};
		hi = {mx(hi.x, p.x + e), mx(hi.y, p.y + e), mx(hi.z, p.z + e)};
	}
	void grow(const Aabb& b) {
		lo = {mn(lo.x, b.lo.x), mn(lo.y, b.lo.y), mn(lo.z, b.lo.z)};
		hi = {mx(hi.x, b.hi.x), mx(hi.y, b.hi.y), mx(hi.z, b.hi.z)};
	}
	float area() const { // surfaceArea(), bounding_volumes.h:139-142
		float e0 = mx(hi.x - lo.x, 1e-5f), e1 = mx(hi.y - lo.y, 1e-5f), e2 = mx(hi.z - lo.z, 1e-5f);
		float P = e0*(e1*e2);
		return 2.0f*(P/e0 + (P/e1 + P/e2));
	}
};

struct BuildNode {
	Aabb box;
	P3 axis{0, 0, 0}; float halfAngle = (float)M_PI;
	int refOffset = 0, nRefs = 0, second = 0, silOffset = 0, nSil = 0;
};
struct Sil { int v[4] = {-1, -1, -1, -1}; int id = -1; };

struct Builder {
	int dim, nV, nP;
	std::vector<P3> pos, vnrm, enrm;
	std::vector<int> prim;      // nP x dim, permuted by the build
	std::vector<int> primId;    // original primitive index of each slot
	std::vector<int> eIdx;      // 3D: 3 edge ids per ORIGINAL primitive
	std::vector<Sil> sil;
	std::vector<BuildNode> nodes;
	std::vector<Aabb> rbox; std::vector<P3> rcen;
	int maxDepth = 0;

	const int* pv(int slot) const { return &prim[(size_t)slot*dim]; }
	P3 faceNormal(const int* v, bool normalize) const {
		P3 n;
		if (dim == 2) { P3 d = pos[v[1]] - pos[v[0]]; n = {d.y, -d.x, 0.0f}; }
		else n = cross(pos[v[1]] - pos[v[0]], pos[v[2]] - pos[v[0]]);
		return normalize ? unit(n) : n;
	}
	bool silHasFace(const Sil& s, int f) const { return f == 0 ? s.v[dim == 2 ? 2 : 3] != -1 : s.v[0] != -1; }
	P3 silFaceNormal(const Sil& s, int f) const { // vertex_silhouettes.inl:36-46, edge_silhouettes.inl:45-68
		if (dim == 2) { int i = f == 0 ? 1 : 0; P3 d = pos[s.v[i + 1]] - pos[s.v[i]]; return unit({d.y, -d.x, 0.0f}); }
		int i = f == 0 ? 3 : 0, j = f == 0 ? 1 : 2, k = f == 0 ? 2 : 1;
		return unit(cross(pos[s.v[k]] - pos[s.v[j]], pos[s.v[i]] - pos[s.v[j]]));
	}

	static float splitCost(const Aabb& L, const Aabb& R, int nL, int nR) { // sbvh.inl:18-23
		Aabb I;
		I.lo = {mx(L.lo.x, R.lo.x), mx(L.lo.y, R.lo.y), mx(L.lo.z, R.lo.z)};
		I.hi = {mn(L.hi.x, R.hi.x), mn(L.hi.y, R.hi.y), mn(L.hi.z, R.hi.z)};
		float cost = (nL/R.area() + nR/L.area())*std::fabs(I.area());
		bool valid = I.hi.x >= I.lo.x && I.hi.y >= I.lo.y && I.hi.z >= I.lo.z;
		return valid ? cost : cost*-1;
	}
	void swapSlots(int i, int j) {
		for (int k = 0; k < dim; k++) std::swap(prim[(size_t)i*dim + k], prim[(size_t)j*dim + k]);
		std::swap(primId[i], primId[j]); std::swap(rbox[i], rbox[j]); std::swap(rcen[i], rcen[j]);
	}
	void split(int parent, int start, int end, int depth) { // buildRecursive, sbvh.inl:141-207
		const int kLeaf = 4, kBuckets = 8, kMaxDepth = 64;
		maxDepth = std::max(maxDepth, depth);
		int cur = (int)nodes.size();
		nodes.emplace_back();
		Aabb bb, bc;
		for (int p = start; p < end; p++) { bb.grow(rbox[p]); bc.grow(rcen[p]); }
		nodes[cur].box = bb;
		bool leaf = end - start <= kLeaf || depth == kMaxDepth - 2;
		if (leaf) { nodes[cur].refOffset = start; nodes[cur].nRefs = end - start; }
		if (parent >= 0 && cur != parent + 1) nodes[parent].second = cur - parent;
		if (leaf) return;

		float best = FLT_MAX; int bestDim = -1; float bestCoord = 0.0f;
		P3 ext = bb.hi - bb.lo;
		for (int d = 0; d < 3; d++) { // computeObjectSplit, sbvh.inl:41-112
			if (at(ext, d) < 1e-6f) continue;
			float width = at(ext, d)/kBuckets;
			Aabb bk[kBuckets], right[kBuckets]; int cnt[kBuckets] = {0}, rcnt[kBuckets] = {0};
			for (int p = start; p < end; p++) {
				int b = (int)((at(rcen[p], d) - at(bb.lo, d))/width);
				b = b < 0 ? 0 : (b > kBuckets - 1 ? kBuckets - 1 : b);
				bk[b].grow(rbox[p]); cnt[b]++;
			}
			Aabb acc;
			for (int b = kBuckets - 1; b > 0; b--) {
				acc.grow(bk[b]); right[b] = acc;
				rcnt[b] = cnt[b] + (b != kBuckets - 1 ? rcnt[b + 1] : 0);
			}
			Aabb left; int nL = 0;
			for (int b = 1; b < kBuckets; b++) {
				left.grow(bk[b - 1]); nL += cnt[b - 1];
				if (nL > 0 && rcnt[b] > 0) {
					float c = splitCost(left, right[b], nL, rcnt[b]);
					if (c < best) { best = c; bestDim = d; bestCoord = at(bb.lo, d) + b*width; }
				}
			}
		}
		if (bestDim == -1) { // centroid-box longest axis, sbvh.inl:105-109
			P3 e = bc.hi - bc.lo;
			bestDim = 0; if (e.y > at(e, bestDim)) bestDim = 1; if (e.z > at(e, bestDim)) bestDim = 2;
			bestCoord = (at(bc.lo, bestDim) + at(bc.hi, bestDim))*0.5f;
		}
		int mid = start; // performObjectSplit, sbvh.inl:114-139
		for (int i = start; i < end; i++) if (at(rcen[i], bestDim) < bestCoord) { swapSlots(i, mid); mid++; }
		if (mid == start || mid == end) mid = start + (end - start)/2;
		split(cur, start, mid, depth + 1);
		split(cur, mid, end, depth + 1);
	}

	void cones(const std::vector<int>& refs, const std::vector<P3>& refN, const std::vector<std::array<P3, 2>>& refFN,
			   int start, int end) { // computeBoundingConesRecursive, sbvh.inl:236-301
		BuildNode& node = nodes[start];
		P3 axis{0, 0, 0}; bool any = false, two = true;
		for (int i = start; i < end; i++) for (int j = 0; j < nodes[i].nSil; j++) {
			int r = nodes[i].silOffset + j;
			axis = axis + refN[r];
			two = two && silHasFace(sil[refs[r]], 0) && silHasFace(sil[refs[r]], 1);
			any = true;
		}
		if (!any) node.halfAngle = (float)-M_PI;
		else if (!two) node.halfAngle = (float)M_PI;
		else {
			float an = std::sqrt(dot(axis, axis));
			if (an > FLT_EPSILON) {
				axis = {axis.x/an, axis.y/an, axis.z/an};
				float ha = 0.0f;
				for (int i = start; i < end; i++) for (int j = 0; j < nodes[i].nSil; j++) {
					int r = nodes[i].silOffset + j;
					for (int k = 0; k < 2; k++) ha = mx(ha, std::acos(mx(-1.0f, mn(1.0f, dot(axis, refFN[r][k])))));
				}
				node.axis = axis; node.halfAngle = ha;
			}
		}
		if (node.nRefs == 0) {
			cones(refs, refN, refFN, start + 1, start + node.second);
			cones(refs, refN, refFN, start + node.second, end);
		}
	}

	// The default mode's own cones.  The reference takes the axis from the silhouettes' own normals (edge / vertex normals) and
	// the half-angle from the face normals; on a closed triangle mesh seen from outside the two can point to opposite sides
	// (half-angles near pi: nothing is ever culled and a closest-silhouette query visits a third of the tree -- measured on
	// box_sphere).  Here the axis is the mean of the very face normals the silhouette test multiplies with the view direction
	// (refFN, the records' n0 / n1), so a cone of half-angle < pi/2 around it is a valid bound for
	// dot(view, n0)*dot(view, n1) < 0.  Only culling changes; the deterministic mode keeps the reference's cones.
	void conesFast(const std::vector<int>& refs, const std::vector<std::array<P3, 2>>& refFN, int start, int end, std::vector<Q4>& out) {
		const BuildNode& node = nodes[start];
		P3 axis{0, 0, 0}; bool any = false, two = true;
		for (int i = start; i < end; i++) for (int j = 0; j < nodes[i].nSil; j++) {
			int r = nodes[i].silOffset + j;
			axis = axis + refFN[r][0] + refFN[r][1];
			two = two && silHasFace(sil[refs[r]], 0) && silHasFace(sil[refs[r]], 1);
			any = true;
		}
		float cosH = any ? -1.0f : 2.0f;
		if (any && two) {
			float an = std::sqrt(dot(axis, axis));
			if (an > FLT_EPSILON) {
				axis = {axis.x/an, axis.y/an, axis.z/an};
				cosH = 1.0f;
				for (int i = start; i < end; i++) for (int j = 0; j < nodes[i].nSil; j++) {
					int r = nodes[i].silOffset + j;
					for (int k = 0; k < 2; k++) cosH = mn(cosH, dot(axis, refFN[r][k]));
				}
			}
		}
		// Where the reference's cone does cull (half-angle below pi/2: the 2D meshes) it is kept as it is, so that the default mode
		// drops exactly what the reference drops -- its exact bound also cuts off records the leaf test would accept through its
		// precision band, e.g. the nearly collinear vertices of a finely subdivided circle.  The own cone (stored as 4 + cos) is
		// evaluated with a slack of two precisions and keeps them, as the reference's non-culling cone does.
		if (node.halfAngle >= 0.0f && node.halfAngle < (float)M_PI_2) out[start] = {node.axis.x, node.axis.y, node.axis.z, std::cos(node.halfAngle)};
		else out[start] = {axis.x, axis.y, axis.z, cosH > 0.0f && cosH <= 1.0f ? 4.0f + cosH : cosH};
		if (node.nRefs == 0) {
			conesFast(refs, refFN, start + 1, start + node.second, out);
			conesFast(refs, refFN, start + node.second, end, out);
		}
	}
};

} // namespace

void buildFlatScene(int dim, const float* verts, int nV, const int* prims, int nP, bool doubleSided, FlatScene& out) {
	out = FlatScene();
	out.dim = dim;
	// zombie::computeBoundingBox over DIM components (fcpw_scene_loader.h:75-93)
	for (int k = 0; k < 3; k++) { out.bboxLo[k] = FLT_MAX; out.bboxHi[k] = -FLT_MAX; }
	for (int i = 0; i < nV; i++) for (int k = 0; k < dim; k++) {
		float p = verts[(size_t)i*dim + k]*1.0f;
		out.bboxLo[k] = mn(out.bboxLo[k], p - FLT_EPSILON); out.bboxHi[k] = mx(out.bboxHi[k], p + FLT_EPSILON);
	}
	if (nP <= 0) return;

	Builder B; B.dim = dim; B.nV = nV; B.nP = nP;
	B.pos.resize(nV);
	for (int i = 0; i < nV; i++) B.pos[i] = {verts[(size_t)i*dim], verts[(size_t)i*dim + 1], dim == 3 ? verts[(size_t)i*dim + 2] : 0.0f};
	B.prim.assign(prims, prims + (size_t)nP*dim);
	B.primId.resize(nP);
	for (int i = 0; i < nP; i++) B.primId[i] = i;

	// pseudo-normals, unweighted at vertices, area-weighted at edges (fcpw.inl:300-354)
	B.vnrm.assign(nV, P3{0, 0, 0});
	if (dim == 3) { // assignEdgeIndices, fcpw.inl:200-221
		std::map<std::pair<int, int>, int> ids;
		B.eIdx.resize((size_t)3*nP);
		for (int i = 0; i < nP; i++) for (int j = 0; j < 3; j++) {
			int I = prims[3*i + j], J = prims[3*i + (j + 1)%3];
			if (I > J) std::swap(I, J);
			auto it = ids.find({I, J});
			if (it == ids.end()) it = ids.emplace(std::make_pair(I, J), (int)ids.size()).first;
			B.eIdx[3*i + j] = it->second;
		}
		B.enrm.assign(ids.size(), P3{0, 0, 0});
	}
	for (int i = 0; i < nP; i++) {
		P3 n = B.faceNormal(B.pv(i), true);
		if (dim == 2) { for (int j = 0; j < 2; j++) B.vnrm[B.pv(i)[j]] = B.vnrm[B.pv(i)[j]] + n*1.0f; }
		else {
			P3 un = B.faceNormal(B.pv(i), false);
			float area = 0.5f*std::sqrt(dot(un, un));
			for (int j = 0; j < 3; j++) {
				B.vnrm[B.pv(i)[j]] = B.vnrm[B.pv(i)[j]] + n*1.0f;
				B.enrm[B.eIdx[3*i + j]] = B.enrm[B.eIdx[3*i + j]] + n*area;
			}
		}
	}
	for (auto& n : B.vnrm) n = unit(n);
	for (auto& n : B.enrm) n = unit(n);

	// silhouettes wired in input order (computeSilhouettes, fcpw.inl:224-291)
	if (dim == 2) {
		B.sil.assign(nV, Sil());
		for (int i = 0; i < nP; i++) {
			int a = prims[2*i], b = prims[2*i + 1];
			B.sil[a].v[1] = a; B.sil[a].v[2] = b; B.sil[a].id = a;
			B.sil[b].v[0] = a; B.sil[b].v[1] = b; B.sil[b].id = b;
		}
	} else {
		B.sil.assign(B.enrm.size(), Sil());
		for (int i = 0; i < nP; i++) for (int j = 0; j < 3; j++) {
			int I = j - 1 < 0 ? 2 : j - 1, J = j, K = j + 1 > 2 ? 0 : j + 1;
			bool fwd = true;
			if (prims[3*i + J] > prims[3*i + K]) { std::swap(J, K); fwd = false; }
			Sil& s = B.sil[B.eIdx[3*i + j]];
			s.v[fwd ? 0 : 3] = prims[3*i + I]; s.v[1] = prims[3*i + J]; s.v[2] = prims[3*i + K];
			s.id = B.eIdx[3*i + j];
		}
	}

	// tree
	B.rbox.resize(nP); B.rcen.resize(nP);
	for (int i = 0; i < nP; i++) {
		const int* v = B.pv(i);
		Aabb b; for (int k = 0; k < dim; k++) b.grow(B.pos[v[k]]);
		B.rbox[i] = b;
		B.rcen[i] = dim == 2 ? (B.pos[v[0]] + B.pos[v[1]])*0.5f
							 : P3{((B.pos[v[0]].x + B.pos[v[1]].x) + B.pos[v[2]].x)/3.0f,
								  ((B.pos[v[0]].y + B.pos[v[1]].y) + B.pos[v[2]].y)/3.0f,
								  ((B.pos[v[0]].z + B.pos[v[1]].z) + B.pos[v[2]].z)/3.0f};
	}
	B.nodes.reserve((size_t)2*nP);
	B.split(-1, 0, nP, 0);

	// 2D: renumber vertices in leaf order and wire the silhouette vertices a second time on top of the
	// first wiring (fcpw.inl:374-410, 469-490) -- stale neighbours at open polyline ends are reference behaviour
	if (dim == 2) {
		std::vector<int> remap(nV, -1);
		std::vector<P3> npos(nV, P3{0, 0, 0}), nnrm(nV, P3{0, 0, 0});
		int next = 0;
		for (const BuildNode& n : B.nodes) for (int j = 0; j < n.nRefs; j++) for (int k = 0; k < 2; k++) {
			int v = B.prim[(size_t)(n.refOffset + j)*2 + k];
			if (remap[v] == -1) { npos[next] = B.pos[v]; nnrm[next] = B.vnrm[v]; remap[v] = next++; }
		}
		for (int& v : B.prim) v = remap[v];
		B.pos.swap(npos); B.vnrm.swap(nnrm);
		for (int i = 0; i < nP; i++) {
			int a = B.prim[2*i], b = B.prim[2*i + 1];
			B.sil[a].v[1] = a; B.sil[a].v[2] = b; B.sil[a].id = a;
			B.sil[b].v[0] = a; B.sil[b].v[1] = b; B.sil[b].id = b;
		}
	}

	// silhouette references per leaf (assignSilhouettesToNodes, sbvh.inl:314-443)
	std::vector<int> refs; std::vector<P3> refN; std::vector<std::array<P3, 2>> refFN; std::vector<float> refAngle;
	std::vector<int> stamp(B.sil.size(), -1);
	for (int ni = 0; ni < (int)B.nodes.size(); ni++) {
		BuildNode& node = B.nodes[ni];
		node.silOffset = (int)refs.size();
		for (int j = 0; j < node.nRefs; j++) {
			int slot = node.refOffset + j;
			for (int k = 0; k < dim; k++) {
				int si = dim == 2 ? B.prim[(size_t)slot*2 + k] : B.eIdx[(size_t)3*B.primId[slot] + k];
				if (stamp[si] == ni) continue;
				stamp[si] = ni;
				const Sil& s = B.sil[si];
				P3 n{0, 0, 0}, n0{0, 0, 0}, n1{0, 0, 0}; float angle = 0.0f; bool ignore = false;
				if (B.silHasFace(s, 0) && B.silHasFace(s, 1)) {
					n = dim == 2 ? B.vnrm[s.v[1]] : B.enrm[s.id];
					n0 = B.silFaceNormal(s, 0); n1 = B.silFaceNormal(s, 1);
					if (dim == 2) angle = n0.x*n1.y - n1.x*n0.y;
					else angle = std::atan2(dot(unit(B.pos[s.v[2]] - B.pos[s.v[1]]), cross(n0, n1)), dot(n0, n1));
					ignore = doubleSided ? false : angle < 1e-3f; // demo/scene.h:84-90
				}
				if (!ignore) { refs.push_back(si); refN.push_back(n); refFN.push_back({n0, n1}); refAngle.push_back(angle); }
			}
		}
		node.nSil = (int)refs.size() - node.silOffset;
	}
	B.cones(refs, refN, refFN, 0, (int)B.nodes.size());
	out.coneF.assign(B.nodes.size(), Q4{0, 0, 0, 2.0f});
	if (!B.nodes.empty()) B.conesFast(refs, refFN, 0, (int)B.nodes.size(), out.coneF);

	// silhouette-traversal blocks of the default mode (nmc_geom.cuh closestSilhouetteFast): everything an inner-node visit needs about
	// its two children in 96 contiguous bytes -- one round trip instead of two dependent ones over ten scattered words
	out.treeF.assign((size_t)6*B.nodes.size(), Q4{0, 0, 0, 0});
	for (size_t i = 0; i < B.nodes.size(); i++) {
		const BuildNode& n = B.nodes[i];
		if (n.nRefs > 0) continue;
		const size_t c[2] = {i + 1, i + (size_t)n.second};
		for (int k = 0; k < 2; k++) {
			const BuildNode& ch = B.nodes[c[k]];
			const Q4& cone = out.coneF[c[k]];
			out.treeF[6*i + 3*k + 0] = {ch.box.lo.x, ch.box.lo.y, ch.box.lo.z, cone.w};
			out.treeF[6*i + 3*k + 1] = {ch.box.hi.x, ch.box.hi.y, ch.box.hi.z, bits(k == 1 ? n.second : 1)};
			out.treeF[6*i + 3*k + 2] = {cone.x, cone.y, cone.z, bits(ch.nRefs)};
		}
	}

	// flatten
	out.nNodes = (int)B.nodes.size(); out.nPrims = nP; out.nSilRefs = (int)refs.size(); out.maxDepth = B.maxDepth;
	out.nodes.resize((size_t)4*out.nNodes);
	for (int i = 0; i < out.nNodes; i++) {
		const BuildNode& n = B.nodes[i];
		out.nodes[4*i + 0] = {n.box.lo.x, n.box.lo.y, n.box.lo.z, bits(n.nRefs)};
		out.nodes[4*i + 1] = {n.box.hi.x, n.box.hi.y, n.box.hi.z, bits(n.second)};
		out.nodes[4*i + 2] = {n.axis.x, n.axis.y, n.axis.z, n.halfAngle};
		out.nodes[4*i + 3] = {bits(n.refOffset), bits(n.silOffset), bits(n.nSil), n.halfAngle < 0.0f ? 2.0f : std::cos(n.halfAngle)};
	}
	out.prims.resize((size_t)(dim == 2 ? 1 : 3)*nP); out.primN.resize(nP); out.nrmV.resize((size_t)(dim == 2 ? 2 : 6)*nP);
	for (int i = 0; i < nP; i++) {
		const int* v = B.pv(i);
		P3 fn = B.faceNormal(v, true);
		out.primN[i] = {fn.x, fn.y, fn.z, bits(B.primId[i])};
		if (dim == 2) {
			out.prims[i] = {B.pos[v[0]].x, B.pos[v[0]].y, B.pos[v[1]].x, B.pos[v[1]].y};
			for (int k = 0; k < 2; k++) out.nrmV[2*i + k] = {B.vnrm[v[k]].x, B.vnrm[v[k]].y, B.vnrm[v[k]].z, 0.0f};
		} else {
			for (int k = 0; k < 3; k++) {
				out.prims[3*i + k] = {B.pos[v[k]].x, B.pos[v[k]].y, B.pos[v[k]].z, 0.0f};
				out.nrmV[6*i + k] = {B.vnrm[v[k]].x, B.vnrm[v[k]].y, B.vnrm[v[k]].z, 0.0f};
				const P3& en = B.enrm[B.eIdx[(size_t)3*B.primId[i] + k]];
				out.nrmV[6*i + 3 + k] = {en.x, en.y, en.z, 0.0f};
			}
		}
	}
	out.sils.resize((size_t)(dim == 2 ? 2 : 4)*refs.size());
	out.silsF.assign(dim == 3 ? 2*refs.size() : 0, Q4{0, 0, 0, 0});
	std::vector<char> emitted(B.sil.size(), 0);
	for (size_t r = 0; r < refs.size(); r++) {
		const Sil& s = B.sil[refs[r]];
		int flags = (B.silHasFace(s, 0) ? 1 : 0) | (B.silHasFace(s, 1) ? 2 : 0);
		// query-time face normals (SilhouetteVertex/Edge::normal(fIndex), normalised) -- only defined with two faces
		P3 n0{0, 0, 0}, n1{0, 0, 0};
		if (flags == 3) { n0 = refFN[r][0]; n1 = refFN[r][1]; }
		if (dim == 2) {
			const P3& p = B.pos[s.v[1]];
			out.sils[2*r + 0] = {p.x, p.y, bits(flags), bits(s.id)};
			out.sils[2*r + 1] = {n0.x, n0.y, n1.x, n1.y};
		} else {
			const P3 &pa = B.pos[s.v[1]], &pb = B.pos[s.v[2]];
			out.sils[4*r + 0] = {pa.x, pa.y, pa.z, bits(flags)};
			out.sils[4*r + 1] = {pb.x, pb.y, pb.z, bits(s.id)};
			out.sils[4*r + 2] = {n0.x, n0.y, n0.z, refAngle[r]};
			out.sils[4*r + 3] = {n1.x, n1.y, n1.z, 0.0f};
			// plane form of the prefilter (nmc_geom.cuh silhouetteCandidate): s = n.x + d = n.(x - pa); a record with fewer than two
			// faces is always a silhouette candidate (zeros pass the band test)
			if (flags == 3) {
				out.silsF[2*r + 0] = {n0.x, n0.y, n0.z, -dot(n0, pa)};
				out.silsF[2*r + 1] = {n1.x, n1.y, n1.z, -dot(n1, pa)};
			}
		}
		if (!emitted[refs[r]]) { // de-duplicated copy: a silhouette shared by several leaves appears once
			emitted[refs[r]] = 1;
			const int w = dim == 2 ? 2 : 4;
			for (int q = 0; q < w; q++) out.silsU.push_back(out.sils[w*r + q]);
			out.nSilU++;
		}
	}
	// ray-scan primitives (default mode): merge chains of connected, collinear, equally oriented 2D segments
	if (dim == 2) {
		std::vector<int> outSeg(nV, -1), inSeg(nV, -1), degOut(nV, 0), degIn(nV, 0);
		for (int i = 0; i < nP; i++) { int a = B.prim[2*i], b = B.prim[2*i + 1]; outSeg[a] = i; degOut[a]++; inSeg[b] = i; degIn[b]++; }
		auto collinear = [&](int i, int j) { // segment i ends where j starts
			P3 u = B.pos[B.prim[2*i + 1]] - B.pos[B.prim[2*i]], v = B.pos[B.prim[2*j + 1]] - B.pos[B.prim[2*j]];
			float cr = u.x*v.y - u.y*v.x, dt = u.x*v.x + u.y*v.y;
			return dt > 0.0f && std::fabs(cr) <= 1e-6f*std::sqrt(dot(u, u)*dot(v, v));
		};
		std::vector<char> used(nP, 0);
		for (int i = 0; i < nP; i++) {
			if (used[i]) continue;
			int first = i, lastSeg = i;
			used[i] = 1;
			for (;;) { // grow backwards
				int a = B.prim[2*first];
				if (degIn[a] != 1 || degOut[a] != 1) break;
				int pr = inSeg[a];
				if (pr < 0 || used[pr] || !collinear(pr, first)) break;
				used[pr] = 1; first = pr;
			}
			for (;;) { // grow forwards
				int b = B.prim[2*lastSeg + 1];
				if (degIn[b] != 1 || degOut[b] != 1) break;
				int nx = outSeg[b];
				if (nx < 0 || used[nx] || !collinear(lastSeg, nx)) break;
				used[nx] = 1; lastSeg = nx;
			}
			const P3 &pa = B.pos[B.prim[2*first]], &pb = B.pos[B.prim[2*lastSeg + 1]];
			P3 d = pb - pa, nn = unit({d.y, -d.x, 0.0f});
			out.rayP.push_back({pa.x, pa.y, pb.x, pb.y});
			out.rayN.push_back({nn.x, nn.y, 0.0f, 0.0f});
		}
		out.nRay = (int)out.rayN.size();
	} else { out.rayP = out.prims; out.rayN = out.primN; out.nRay = nP; }

	// group boxes for the flat scans (tree order keeps neighbours together); group = 8 records in 2D, 4 in 3D
	const int G = dim == 2 ? 8 : 4;
	auto groupBoxes = [&](const std::vector<Q4>& rec, int perItem, int nItems, int ptsPerItem, std::vector<Q4>& outBoxes) {
		for (int g0 = 0; g0 < nItems; g0 += G) {
			float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
			for (int i = g0; i < std::min(nItems, g0 + G); i++) for (int q = 0; q < ptsPerItem; q++) {
				const Q4& r = rec[(size_t)perItem*i + q];
				float pts[2][3] = {{r.x, r.y, dim == 3 ? r.z : 0.0f}, {r.z, r.w, 0.0f}};
				int np = (dim == 2 && perItem == 1) ? 2 : 1; // a 2D segment record packs both end points
				for (int k = 0; k < np; k++) for (int c = 0; c < 3; c++) { lo[c] = std::min(lo[c], pts[k][c]); hi[c] = std::max(hi[c], pts[k][c]); }
			}
			outBoxes.push_back({lo[0], lo[1], lo[2], 0.0f}); outBoxes.push_back({hi[0], hi[1], hi[2], 0.0f});
		}
	};
	groupBoxes(out.rayP, dim == 2 ? 1 : 3, out.nRay, dim == 2 ? 1 : 3, out.grpP);
	groupBoxes(out.silsU, dim == 2 ? 2 : 4, out.nSilU, dim == 2 ? 1 : 2, out.grpS); // 2D: the vertex; 3D: both edge end points
	auto superBoxes = [&](const std::vector<Q4>& grp, std::vector<Q4>& outBoxes) { // bounds of 32 consecutive group boxes
		const size_t nG = grp.size()/2;
		for (size_t s0 = 0; s0 < nG; s0 += 32) {
			Q4 lo = grp[2*s0], hi = grp[2*s0 + 1];
			for (size_t g = s0 + 1; g < std::min(nG, s0 + 32); g++) {
				const Q4 &a = grp[2*g], &b = grp[2*g + 1];
				lo.x = std::min(lo.x, a.x); lo.y = std::min(lo.y, a.y); lo.z = std::min(lo.z, a.z);
				hi.x = std::max(hi.x, b.x); hi.y = std::max(hi.y, b.y); hi.z = std::max(hi.z, b.z);
			}
			outBoxes.push_back(lo); outBoxes.push_back(hi);
		}
	};
	superBoxes(out.grpP, out.supP);
	superBoxes(out.grpS, out.supS);

	// scan-friendly records (after the boxes, which need the end points):
	//  * ray primitives: 2D (origin, edge vector) = (pa.xy, pb - pa); 3D plane form (N, d0)(A, d1)(B, d2), see below;
	//  * both lists are padded to a whole number of groups with records no query can accept (a silhouette at
	//    infinity, a degenerate primitive), so the scans run fixed-trip inner loops;
	if (dim == 2) for (int i = 0; i < out.nRay; i++) { Q4& q = out.rayP[i]; q.z -= q.x; q.w -= q.y; }
	else for (int i = 0; i < out.nRay; i++) {
		// 3D: plane form.  With v1 = pb - pa, v2 = pc - pa, N = v1 x v2 (not normalised), A = (v2 x N)/|N|^2, B = (N x v1)/|N|^2:
		//   t = -(N.o + d0)/(N.dir),  P = o + t dir,  P = pa + v v1 + w v2 with v = A.P + d1, w = B.P + d2
		// (d0 = -N.pa, d1 = -A.pa, d2 = -B.pa): 15 FMA + 1 reciprocal per triangle and no branches, against two cross
		// products and three early exits for the Moeller-Trumbore form of Triangle::intersect (triangles.inl:219-256).
		// A degenerate triangle gives N = 0: the denominator test rejects it.  Evaluated in double, stored in float.
		Q4 &a = out.rayP[3*i], &b = out.rayP[3*i + 1], &c = out.rayP[3*i + 2];
		const double pa[3] = {a.x, a.y, a.z}, v1[3] = {(double)b.x - a.x, (double)b.y - a.y, (double)b.z - a.z},
					 v2[3] = {(double)c.x - a.x, (double)c.y - a.y, (double)c.z - a.z};
		const double N[3] = {v1[1]*v2[2] - v1[2]*v2[1], v1[2]*v2[0] - v1[0]*v2[2], v1[0]*v2[1] - v1[1]*v2[0]};
		const double nn = N[0]*N[0] + N[1]*N[1] + N[2]*N[2];
		if (nn > 0.0) {
			const double A[3] = {(v2[1]*N[2] - v2[2]*N[1])/nn, (v2[2]*N[0] - v2[0]*N[2])/nn, (v2[0]*N[1] - v2[1]*N[0])/nn};
			const double Bv[3] = {(N[1]*v1[2] - N[2]*v1[1])/nn, (N[2]*v1[0] - N[0]*v1[2])/nn, (N[0]*v1[1] - N[1]*v1[0])/nn};
			a = {(float)N[0], (float)N[1], (float)N[2], (float)-(N[0]*pa[0] + N[1]*pa[1] + N[2]*pa[2])};
			b = {(float)A[0], (float)A[1], (float)A[2], (float)-(A[0]*pa[0] + A[1]*pa[1] + A[2]*pa[2])};
			c = {(float)Bv[0], (float)Bv[1], (float)Bv[2], (float)-(Bv[0]*pa[0] + Bv[1]*pa[1] + Bv[2]*pa[2])};
		} else { a = {0, 0, 0, 0}; b = {0, 0, 0, 0}; c = {0, 0, 0, 0}; }
	}
	//  * silhouettes are rewritten for a two-stage test.  Both faces of a silhouette vertex / edge contain it, so
	//    dot(x - p, n_k) is the signed distance of x to the plane of face k: s_k(x) = n_k.x + c_k with c_k = -n_k.p.
	//    Stage 1 needs only (n0, c0, n1, c1): a record can be a silhouette from x only if s_0 s_1 < 0 (or x is within
	//    the precision band of a plane).  Stage 2 (distance, exact rules) reads the position.  Records with a single
	//    face (always silhouettes) carry zero normals and c0 = 1, c1 = -1, so stage 1 always passes them on.
	//      2D: (n0.x, n0.y, c0, n1.x) (n1.y, c1, p.x, p.y)
	//      3D: (n0.xyz, c0) (n1.xyz, c1) (pa.xyz, flags) (pb.xyz, dihedral)
	for (int i = 0; i < out.nSilU; i++) {
		if (dim == 2) {
			Q4 a = out.silsU[2*i], b = out.silsU[2*i + 1]; // (p.x, p.y, flags, id) (n0.x, n0.y, n1.x, n1.y)
			int flags; std::memcpy(&flags, &a.z, 4);
			if (flags == 3) {
				out.silsU[2*i] = {b.x, b.y, -(b.x*a.x + b.y*a.y), b.z};
				out.silsU[2*i + 1] = {b.w, -(b.z*a.x + b.w*a.y), a.x, a.y};
			} else { out.silsU[2*i] = {0.0f, 0.0f, 1.0f, 0.0f}; out.silsU[2*i + 1] = {0.0f, -1.0f, a.x, a.y}; }
		} else {
			Q4 a = out.silsU[4*i], b = out.silsU[4*i + 1], c = out.silsU[4*i + 2], d = out.silsU[4*i + 3]; // (pa, flags) (pb, id) (n0, dihedral) (n1, -)
			int flags; std::memcpy(&flags, &a.w, 4);
			if (flags == 3) {
				out.silsU[4*i] = {c.x, c.y, c.z, -(c.x*a.x + c.y*a.y + c.z*a.z)};
				out.silsU[4*i + 1] = {d.x, d.y, d.z, -(d.x*a.x + d.y*a.y + d.z*a.z)};
			} else { out.silsU[4*i] = {0.0f, 0.0f, 0.0f, 1.0f}; out.silsU[4*i + 1] = {0.0f, 0.0f, 0.0f, -1.0f}; }
			out.silsU[4*i + 2] = a;
			out.silsU[4*i + 3] = {b.x, b.y, b.z, c.w};
		}
	}
	const float far = 1e30f; // (x - far)^2 overflows to +inf > any search radius
	for (int i = out.nSilU; i % G != 0; i++) { // never a candidate (s0 s1 = 1, |s| = 1 outside any sane band), and infinitely far if it is
		if (dim == 2) { out.silsU.push_back({0.0f, 0.0f, 1.0f, 0.0f}); out.silsU.push_back({0.0f, 1.0f, far, far}); }
		else { out.silsU.push_back({0, 0, 0, 1.0f}); out.silsU.push_back({0, 0, 0, 1.0f}); out.silsU.push_back({far, far, far, bits(3)}); out.silsU.push_back({far, far, far, 0.0f}); }
	}
	for (int i = out.nRay; i % G != 0; i++) {
		if (dim == 2) out.rayP.push_back({far, far, 0.0f, 0.0f}); // zero edge vector: determinant 0, rejected
		else { out.rayP.push_back({0, 0, 0, 0}); out.rayP.push_back({0, 0, 0, 0}); out.rayP.push_back({0, 0, 0, 0}); } // N = 0: rejected
		out.rayN.push_back({0, 0, 0, 0});
	}
}

} // namespace nmc
