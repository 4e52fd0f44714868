// csrc/nmc_geom.cuh -- boundary acceleration structure (flattened BVH with normal cones, "SNCH")
// and the geometric queries of the walk, as device functions over a SceneView.
//
// Replaces, for this path, FCPW's Sbvh traversal (deps/fcpw/include/fcpw/aggregates/sbvh.inl:538-1255),
// its primitives (geometry/{line_segments,triangles,vertex_silhouettes,edge_silhouettes}.inl),
// core/bounding_volumes.h and zombie's query adapters (include/zombie/utils/fcpw_scene_loader.h:292-652).
//
// Layout (all 16-byte records so that a node / primitive is fetched with vectorised loads and the
// whole structure can be staged in shared memory when it is small):
//   nodes  4 x float4 per node : (lo.xyz, nRefs) (hi.xyz, secondChildOffset) (coneAxis.xyz, halfAngle)
//                                (refOffset, silOffset, nSilRefs, cos(halfAngle))  [ints stored bit-cast]
//   prims  2D 1 x float4 per segment (pa.xy, pb.xy); 3D 3 x float4 per triangle (pa, pb, pc)
//   primN  1 x float4 per primitive: unit face normal, w = primitive index in the input mesh
//   nrmV   pseudo-normals for signed distance: 2D 2 x float4 (vertex a, b); 3D 6 x float4
//          (vertex a, b, c, edge 0, 1, 2)
//   sils   silhouette references in node order: 2D 2 x float4 (p.xy, flags, id)(n0.xy, n1.xy);
//          3D 4 x float4 (pa, flags)(pb, id)(n0, dihedral)(n1, -); flags bit0/bit1 = has face 0/1
// Traversal order, tie-breaking and arithmetic association follow the reference so that the
// deterministic mode returns the same primitive, point and distance.
#pragma once
#include "nmc_math.cuh"

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
#endif

namespace nmc {

struct SceneView {
	int dim, nNodes, nPrims, nSilRefs;
	const float4* nodes;
	const float4* coneF;   // default mode: own normal cone per node (axis, cos halfAngle; scene_build.cpp conesFast)
	const float4* prims;
	const float4* primN;
	const float4* nrmV;
	const float4* sils;
	const float4* treeF;   // default mode: 6 words per inner node, both children's boxes / cone codes / cone axes (scene_build.cpp)
	const float4* silsF;   // 3D default mode: the two face planes of every silhouette reference (prefilter, scene_build.cpp)
	const float4* silsU; int nSilU;   // distinct silhouettes (flat scan)
	const float4* grpP; const float4* grpS; // (lo, hi) per group (8 in 2D, 4 in 3D) of ray primitives / silhouettes
	const float4* supP; const float4* supS; // (lo, hi) per 32 groups: second level of the flat scans on large meshes
	const float4* rayP; const float4* rayN; int nRay; // ray-scan primitives as (origin, edge vectors); 2D: collinear chains merged; padded to whole groups
	float bboxLo[3], bboxHi[3];
	const float* src; int n0, n1, n2;
	float srcScale[3], srcOff[3];     // default mode: texel index along box axis k = (int)(x_k*srcScale[k] + srcOff[k])
	float absorption; int watertight, doubleSided;
};

struct Box { V3 lo, hi; };
struct Hit { float d; V3 p, n; float u, v; int ref; };

#define NMC_STACK 64 // FCPW_SBVH_MAX_DEPTH (aggregates/sbvh.h:5)

NMC_HD V3 xyz(const float4& q) { return mk(q.x, q.y, q.z); }

// BoundingBox::computeSquaredDistance (bounding_volumes.h:62-67)
NMC_HD void boxSqDist(V3 lo, V3 hi, V3 p, float& d2Min, float& d2Max) {
	float ux = lo.x - p.x, vx = p.x - hi.x, uy = lo.y - p.y, vy = p.y - hi.y, uz = lo.z - p.z, vz = p.z - hi.z;
	V3 a = mk(maxS(maxS(ux, vx), 0.0f), maxS(maxS(uy, vy), 0.0f), maxS(maxS(uz, vz), 0.0f));
	V3 c = mk(minS(ux, vx), minS(uy, vy), minS(uz, vz));
	d2Min = dot(a, a); d2Max = dot(c, c);
}
// BoundingBox::intersect(ray) (bounding_volumes.h:99-114)
NMC_HD bool boxRay(V3 lo, V3 hi, V3 o, V3 invD, float rtMax, float& tMin, float& tMax) {
	float t0 = (lo.x - o.x)*invD.x, t1 = (hi.x - o.x)*invD.x;
	float nx = minS(t0, t1), fx = maxS(t0, t1);
	t0 = (lo.y - o.y)*invD.y; t1 = (hi.y - o.y)*invD.y;
	float ny = minS(t0, t1), fy = maxS(t0, t1);
	t0 = (lo.z - o.z)*invD.z; t1 = (hi.z - o.z)*invD.z;
	float nz = minS(t0, t1), fz = maxS(t0, t1);
	float tNearMax = maxS(0.0f, maxS(nx, maxS(ny, nz)));
	float tFarMin = minS(rtMax, minS(fx, minS(fy, fz)));
	if (tNearMax > tFarMin) return false;
	tMin = tNearMax; tMax = tFarMin;
	return true;
}
NMC_HD bool inRange(float val, float low, float high) { return val >= low && val <= high; }
// projectToPlane<3> + computeOrthonormalBasis (bounding_volumes.h:175-209)
NMC_HD float projectToPlane(V3 n, V3 e) {
	float sign = copysignf(1.0f, n.z);
	const float a = -1.0f/(sign + n.z);
	const float b = n.x*n.y*a;
	V3 b1 = mk(fabsf(1.0f + sign*n.x*n.x*a), fabsf(sign*b), fabsf(-sign*n.x));
	V3 b2 = mk(fabsf(b), fabsf(sign + n.y*n.y*a), fabsf(-n.y));
	float r1 = dot(e, b1), r2 = dot(e, b2);
	return sqrtf(r1*r1 + r2*r2);
}
// BoundingCone::overlap (bounding_volumes.h:225-271)
template <class M>
NMC_HD bool coneOverlap(V3 axis, float halfAngle, V3 o, V3 lo, V3 hi, float distToBox) {
	if (halfAngle >= kPi2 || distToBox < kEps) return true;
	V3 c = (lo + hi)*0.5f;
	V3 vca = c - o;
	float l = norm(vca);
	vca = vca/l;
	float dAxisAngle = M::acos_(maxS(-1.0f, minS(1.0f, dot(axis, vca))));
	if (inRange((float)kPi2, dAxisAngle - halfAngle, dAxisAngle + halfAngle)) return true;
	V3 e = hi - c;
	float r2 = dot(e, e);
	if (l*l > r2) {
		float r = sqrtf(r2);
		float vha = M::asin_(r/l);
		float sum = halfAngle + vha;
		return sum >= kPi2 ? true : inRange((float)kPi2, dAxisAngle - sum, dAxisAngle + sum);
	}
	float d = dot(e, mk(fabsf(vca.x), fabsf(vca.y), fabsf(vca.z)));
	float s = l - d;
	if (s <= 0.0f) return true;
	d = projectToPlane(vca, e);
	float vha = M::atan2_(d, s);
	float sum = halfAngle + vha;
	return sum >= kPi2 ? true : inRange((float)kPi2, dAxisAngle - sum, dAxisAngle + sum);
}

// Same predicate without inverse trigonometry (default mode): with c = axis.view,
//   inRange(pi/2, a - h, a + h), a = acos(c)  <=>  |asin c| <= h  <=>  |c| <= sin h      (h < pi/2)
// and for the widened test  |c| <= sin(h + v) = sin h cos v + cos h sin v, "h + v >= pi/2" <=> cos(h + v) <= 0,
// where (sin v, cos v) = (r/l, sqrt(1 - r^2/l^2)) or (d, s)/sqrt(d^2 + s^2).  cosH = cos(halfAngle) is stored per node.
// slack (a small multiple of the silhouette precision) keeps the records the leaf test accepts through its precision band:
// |dot(view, n)| <= precision counts as "on the face plane" there (isSilhouette), which an exact cone bound would cut off.
NMC_HD bool coneOverlapFast(V3 axis, float w, V3 o, V3 lo, V3 hi, float distToBox, float ownSlack) {
	if (w <= 0.0f || distToBox < kEps) return true;
	const bool own = w > 3.0f; // scene_build.cpp conesFast: the reference's cone as cos, the default mode's own cone as 4 + cos
	const float cosH = own ? w - 4.0f : w, slack = own ? ownSlack : 0.0f;
	float sinH = sqrtf(fmaxf(0.0f, 1.0f - cosH*cosH));
	V3 c = (lo + hi)*0.5f;
	V3 vca = c - o;
	float l2 = dot(vca, vca);
	float il = rsqrtf(l2);
	vca = vca*il;
	float ca = fabsf(fminf(1.0f, fmaxf(-1.0f, dot(axis, vca))));
	if (ca <= sinH + slack) return true;
	V3 e = hi - c;
	float r2 = dot(e, e);
	float sv, cv;
	if (l2 > r2) { sv = sqrtf(r2)*il; cv = sqrtf(fmaxf(0.0f, 1.0f - sv*sv)); }
	else {
		float l = l2*il;
		float d = dot(e, mk(fabsf(vca.x), fabsf(vca.y), fabsf(vca.z)));
		float sgap = l - d;
		if (sgap <= 0.0f) return true;
		d = projectToPlane(vca, e);
		float ih = rsqrtf(d*d + sgap*sgap);
		sv = d*ih; cv = sgap*ih;
	}
	if (cosH*cv - sinH*sv <= slack) return true;
	return ca <= sinH*cv + cosH*sv + slack;
}

// findClosestPointLineSegment (line_segments.inl:184-209)
NMC_HD float closestOnSegment(V3 pa, V3 pb, V3 x, V3& pt, float& t) {
	V3 u = pb - pa, v = x - pa;
	float c1 = dot(u, v);
	if (c1 <= 0.0f) { pt = pa; t = 0.0f; return norm(x - pt); }
	float c2 = dot(u, u);
	if (c2 <= c1) { pt = pb; t = 1.0f; return norm(x - pt); }
	t = c1/c2;
	pt = pa + u*t;
	return norm(x - pt);
}
// findClosestPointTriangle (triangles.inl:258-341)
NMC_HD float closestOnTriangle(V3 pa, V3 pb, V3 pc, V3 x, V3& pt, float& t0, float& t1) {
	V3 ab = pb - pa, ac = pc - pa, ax = x - pa;
	float d1 = dot(ab, ax), d2 = dot(ac, ax);
	if (d1 <= 0.0f && d2 <= 0.0f) { t0 = 1.0f; t1 = 0.0f; pt = pa; return norm(x - pt); }
	V3 bx = x - pb;
	float d3 = dot(ab, bx), d4 = dot(ac, bx);
	if (d3 >= 0.0f && d4 <= d3) { t0 = 0.0f; t1 = 1.0f; pt = pb; return norm(x - pt); }
	V3 cx = x - pc;
	float d5 = dot(ab, cx), d6 = dot(ac, cx);
	if (d6 >= 0.0f && d5 <= d6) { t0 = 0.0f; t1 = 0.0f; pt = pc; return norm(x - pt); }
	float vc = d1*d4 - d3*d2;
	if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
		float v = d1/(d1 - d3);
		t0 = 1.0f - v; t1 = v; pt = pa + ab*v; return norm(x - pt);
	}
	float vb = d5*d2 - d1*d6;
	if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
		float w = d2/(d2 - d6);
		t0 = 1.0f - w; t1 = 0.0f; pt = pa + ac*w; return norm(x - pt);
	}
	float va = d3*d6 - d5*d4;
	if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
		float w = (d4 - d3)/((d4 - d3) + (d5 - d6));
		t0 = 0.0f; t1 = 1.0f - w; pt = pb + (pc - pb)*w; return norm(x - pt);
	}
	float denom = 1.0f/(va + vb + vc);
	float v = vb*denom, w = vc*denom;
	t0 = 1.0f - v - w; t1 = v;
	pt = (pa + ab*v) + ac*w;
	return norm(x - pt);
}

// Traversal stack policies.  LocalStack: a plain per-thread array (deterministic mode, host harness).
// StridedStack: entries of one thread interleaved with those of the other threads of the CTA in shared
// memory (slot i of thread t at base[i*stride + t]) -- bank-conflict free and no local-memory traffic.
struct LocalStack {
	int n[NMC_STACK]; float d[NMC_STACK];
	NMC_HD void put(int i, int node, float dist) { n[i] = node; d[i] = dist; }
	NMC_HD int node(int i) const { return n[i]; }
	NMC_HD float dist(int i) const { return d[i]; }
};
struct StridedStack {
	int* nodes; float* dists; int stride;
	NMC_HD void put(int i, int node, float dist) { nodes[i*stride] = node; dists[i*stride] = dist; }
	NMC_HD int node(int i) const { return nodes[i*stride]; }
	NMC_HD float dist(int i) const { return dists[i*stride]; }
};

// closest point on the boundary mesh: Sbvh::findClosestPointFromNode (sbvh.inl:948-1074).
// Returns false when nothing lies within sqrt(r2).  wantNormal: pseudo-normal as in
// LineSegment/Triangle::normal(uv) with soup normals present (line_segments.inl:60-77, triangles.inl:62-90).
template <int DIM, class Stack>
NMC_TRAV bool closestPoint(const SceneView& S, Stack& stack, V3 x, float r2, bool wantNormal, Hit& out) {
	if (S.nNodes == 0) return false;
	float b0, b1, b2, b3;
	bool found = false;
	{
		float4 a = S.nodes[0], b = S.nodes[1];
		boxSqDist(xyz(a), xyz(b), x, b0, b1);
	}
	if (!(b0 <= r2)) return false;
	r2 = minS(r2, b1);
	stack.put(0, 0, b0);
	int sp = 0;
	out.d = kMaxF; out.ref = -1; out.u = 0.0f; out.v = 0.0f;
	while (sp >= 0) {
		int ni = stack.node(sp); float cd = stack.dist(sp); sp--;
		if (cd > r2) continue;
		float4 na = S.nodes[4*ni], nb = S.nodes[4*ni + 1];
		int nRefs = asInt(na.w);
		if (nRefs > 0) {
			int refOffset = asInt(S.nodes[4*ni + 3].x);
			for (int p = 0; p < nRefs; p++) {
				int ri = refOffset + p;
				V3 pt; float u = 0.0f, v = 0.0f, d;
				if (DIM == 2) {
					float4 q = S.prims[ri];
					d = closestOnSegment(mk(q.x, q.y, 0.0f), mk(q.z, q.w, 0.0f), x, pt, u); v = -1.0f;
				} else {
					d = closestOnTriangle(xyz(S.prims[3*ri]), xyz(S.prims[3*ri + 1]), xyz(S.prims[3*ri + 2]), x, pt, u, v);
				}
				if (d*d <= r2) {
					found = true;
					r2 = minS(r2, d*d);
					out.d = d; out.p = pt; out.u = u; out.v = v; out.ref = ri;
				}
			}
		} else {
			int c0 = ni + 1, c1 = ni + asInt(nb.w);
			boxSqDist(xyz(S.nodes[4*c0]), xyz(S.nodes[4*c0 + 1]), x, b0, b1); bool hit0 = b0 <= r2;
			r2 = minS(r2, b1);
			boxSqDist(xyz(S.nodes[4*c1]), xyz(S.nodes[4*c1 + 1]), x, b2, b3); bool hit1 = b2 <= r2;
			r2 = minS(r2, b3);
			if (hit0 && hit1) {
				int closer = c0, other = c1;
				if (b0 == 0.0f && b2 == 0.0f) {
					if (b3 < b1) { closer = c1; other = c0; }
				} else if (b2 < b0) {
					float t = b0; b0 = b2; b2 = t;
					closer = c1; other = c0;
				}
				sp++; stack.put(sp, other, b2);
				sp++; stack.put(sp, closer, b0);
			} else if (hit0) { sp++; stack.put(sp, c0, b0); }
			else if (hit1) { sp++; stack.put(sp, c1, b2); }
		}
	}
	if (found && wantNormal) {
		int ri = out.ref;
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

// LineSegment::intersect(ray) (line_segments.inl:146-182) / Triangle::intersect(ray) (triangles.inl:219-256)
template <int DIM>
NMC_HD bool primRay(const SceneView& S, int ri, V3 o, V3 dir, float tMax, bool occl, Hit& h) {
	if (DIM == 2) {
		float4 q = S.prims[ri];
		V3 pa = mk(q.x, q.y, 0.0f), pb = mk(q.z, q.w, 0.0f);
		V3 u = pa - o, v = pb - pa;
		float dv = dir.x*v.y - dir.y*v.x;
		if (fabsf(dv) <= kEps) return false;
		float ud = u.x*dir.y - u.y*dir.x;
		float s = ud/dv;
		if (s >= 0.0f && s <= 1.0f) {
			float uv = u.x*v.y - u.y*v.x;
			float t = uv/dv;
			if (t >= 0.0f && t <= tMax) {
				if (occl) return true;
				h.d = t; h.p = pa + s*v; h.n = xyz(S.primN[ri]); h.u = s; h.v = -1.0f; h.ref = ri;
				return true;
			}
		}
		return false;
	}
	V3 pa = xyz(S.prims[3*ri]), pb = xyz(S.prims[3*ri + 1]), pc = xyz(S.prims[3*ri + 2]);
	V3 v1 = pb - pa, v2 = pc - pa;
	V3 p = cross(dir, v2);
	float det = dot(v1, p);
	if (fabsf(det) <= kEps) return false;
	float invDet = 1.0f/det;
	V3 s = o - pa;
	float v = dot(s, p)*invDet;
	if (v < 0 || v > 1) return false;
	V3 q = cross(s, v1);
	float w = dot(dir, q)*invDet;
	if (w < 0 || v + w > 1) return false;
	float t = dot(v2, q)*invDet;
	if (t >= 0.0f && t <= tMax) {
		if (occl) return true;
		h.d = t; h.p = (pa + v1*v) + v2*w; h.n = xyz(S.primN[ri]); h.u = 1.0f - v - w; h.v = v; h.ref = ri;
		return true;
	}
	return false;
}
// NMC_WHILE_WHILE: `if` = one node per trip (a leaf is processed right after it is popped); `while` (-DNMC_TRAV_WHILE_WHILE) = a lane
// walks inner nodes until it holds a leaf, then the lanes work on their leaves together.  The second form was tried because ncu showed
// the 3D record loop running with 7 of 32 lanes on a 20 k-triangle obstacle; measured on the B200 it LOSES (16.5 k segments 1.84e8 ->
// 1.28e8 walks/s, 131 k 5.4e7 -> 3.3e7, 20 k triangles 1.29e7 -> 1.16e7: lanes holding a leaf wait for the slowest inner-node walk), so
// it is off; profiles/experiments/r02_while_while_mbvh.jsonl.
#if defined(NMC_FAST_GEOM) && defined(NMC_TRAV_WHILE_WHILE)
#define NMC_WHILE_WHILE while
#else
#define NMC_WHILE_WHILE if
#endif
// closest-hit / any-hit ray: Sbvh::intersectFromNode + processSubtreeForIntersection (sbvh.inl:538-683)
template <int DIM, class Stack>
NMC_TRAV bool rayIntersect(const SceneView& S, Stack& stack, V3 o, V3 dir, float tMax, bool occl, Hit& out) {
	if (S.nNodes == 0) return false;
	V3 invD = mk(1.0f/dir.x, 1.0f/dir.y, 1.0f/dir.z);
	float b0, b1, b2, b3;
	int hits = 0;
	if (!boxRay(xyz(S.nodes[0]), xyz(S.nodes[1]), o, invD, tMax, b0, b1)) return false;
	stack.put(0, 0, b0);
	int sp = 0;
	int refOffset = 0, nLeaf = 0; // see NMC_WHILE_WHILE
	while (sp >= 0 || nLeaf > 0) {
		NMC_WHILE_WHILE (sp >= 0 && nLeaf == 0) {
			int ni = stack.node(sp); float cd = stack.dist(sp); sp--;
			if (cd > tMax) continue;
			float4 na = S.nodes[4*ni];
			int nRefs = asInt(na.w);
			if (nRefs > 0) { refOffset = asInt(S.nodes[4*ni + 3].x); nLeaf = nRefs; }
			else {
				int c0 = ni + 1, c1 = ni + asInt(S.nodes[4*ni + 1].w);
				bool hit0 = boxRay(xyz(S.nodes[4*c0]), xyz(S.nodes[4*c0 + 1]), o, invD, tMax, b0, b1);
				bool hit1 = boxRay(xyz(S.nodes[4*c1]), xyz(S.nodes[4*c1 + 1]), o, invD, tMax, b2, b3);
				if (hit0 && hit1) {
					int closer = c0, other = c1;
					if (b2 < b0) { float t = b0; b0 = b2; b2 = t; closer = c1; other = c0; }
					sp++; stack.put(sp, other, b2);
					sp++; stack.put(sp, closer, b0);
				} else if (hit0) { sp++; stack.put(sp, c0, b0); }
				else if (hit1) { sp++; stack.put(sp, c1, b2); }
			}
		}
		for (int p = 0; p < nLeaf; p++) {
			Hit h;
			if (primRay<DIM>(S, refOffset + p, o, dir, tMax, occl, h)) {
				if (occl) return true;
				hits++;
				tMax = minS(tMax, h.d);
				out = h;
			}
		}
		nLeaf = 0;
	}
	return hits > 0;
}

// isSilhouetteVertex (vertex_silhouettes.inl:62-87) / isSilhouetteEdge (edge_silhouettes.inl:83-110)
NMC_HD bool isSilhouette(float concavity, V3 n0, V3 n1, V3 viewDir, float d, bool flip, float precision) {
	float sign = flip ? 1.0f : -1.0f;
	if (d <= precision) return sign*concavity > precision;
	V3 vu = viewDir/d;
	float dot0 = dot(vu, n0), dot1 = dot(vu, n1);
	if (fabsf(dot0) <= precision) return sign*dot1 > precision;
	if (fabsf(dot1) <= precision) return sign*dot0 > precision;
	return dot0*dot1 < 0.0f;
}
// Cheap necessary condition for isSilhouette() on an edge with two faces (default mode): the face normals are perpendicular
// to the edge, so n.(x - pa) = d*dot(view, n) for the closest point's view direction whatever that point is.  The edge can only
// pass if x sees the two face planes from opposite sides, or sits within the precision band of one of them
// (|dot| <= precision and d <= sqrt(r2)), or within `precision` of the edge itself (|s| <= d <= precision).  On a smooth
// obstacle almost every record of the visited leaves fails this before the closest point (a division, a square root) is computed.
NMC_HD bool silhouetteCandidate(V3 n0, V3 n1, V3 rel, float precision, float r2) {
	const float s0 = dot(n0, rel), s1 = dot(n1, rel);
	return s0*s1 < 0.0f || fminf(fabsf(s0), fabsf(s1)) <= precision*fmaxf(1.0f, sqrtf(r2));
}
#if defined(NMC_FAST_GEOM)
// One record of a 3D leaf after the prefilter (default mode): the tests of the generic loop below, in the same order.
NMC_HD void silhouetteRecord3(const SceneView& S, int ri, V3 x, bool flip, float sqMinR, float precision, float& r2, bool& found, int& lastId, float& dOut) {
	const float4 s0 = S.sils[4*ri], s1 = S.sils[4*ri + 1], s2 = S.sils[4*ri + 2], s3 = S.sils[4*ri + 3];
	const int flags = asInt(s0.w), id = asInt(s1.w);
	if (id == lastId || sqMinR >= r2) return;
	V3 pt; float t;
	const float d = closestOnSegment(xyz(s0), xyz(s1), x, pt, t);
	if (d*d > r2) return;
	bool isSil = (flags & 3) != 3;
	if (!isSil) isSil = isSilhouette(s2.w, xyz(s2), xyz(s3), x - pt, d, flip, precision);
	if (isSil) { found = true; r2 = minS(r2, d*d); dOut = d; lastId = id; }
}
// The records of a 3D leaf, two at a time: the plane pairs of both (32 bytes each, SceneView::silsF) are requested together, so a
// leaf of seven records costs four load round trips instead of seven (the loop is latency-bound: ncu), and
// the full record (64 bytes) is only read for the few that pass.
NMC_HD void silhouetteLeaf3(const SceneView& S, int silOffset, int nSil, V3 x, bool flip, float sqMinR, float precision, float& r2, bool& found, int& lastId, float& dOut) {
	for (int p = 0; p < nSil; p += 2) {
		const int ri = silOffset + p;
		const bool two = p + 1 < nSil;
		const float4 a0 = S.silsF[2*ri], a1 = S.silsF[2*ri + 1];
		const float4 b0 = S.silsF[2*(two ? ri + 1 : ri)], b1 = S.silsF[2*(two ? ri + 1 : ri) + 1];
		const float band = precision*fmaxf(1.0f, sqrtf(r2));
		const float sa0 = fmaf(a0.x, x.x, fmaf(a0.y, x.y, fmaf(a0.z, x.z, a0.w))), sa1 = fmaf(a1.x, x.x, fmaf(a1.y, x.y, fmaf(a1.z, x.z, a1.w)));
		const float sb0 = fmaf(b0.x, x.x, fmaf(b0.y, x.y, fmaf(b0.z, x.z, b0.w))), sb1 = fmaf(b1.x, x.x, fmaf(b1.y, x.y, fmaf(b1.z, x.z, b1.w)));
		const bool ca = sa0*sa1 < 0.0f || fminf(fabsf(sa0), fabsf(sa1)) <= band;
		const bool cb = two && (sb0*sb1 < 0.0f || fminf(fabsf(sb0), fabsf(sb1)) <= band);
		if (ca) silhouetteRecord3(S, ri, x, flip, sqMinR, precision, r2, found, lastId, dOut);
		if (cb) silhouetteRecord3(S, ri + 1, x, flip, sqMinR, precision, r2, found, lastId, dOut);
		if (sqMinR >= r2) break;
	}
}
#endif
// closest silhouette point: Sbvh::findClosestSilhouettePointFromNode (sbvh.inl:1093-1255) with
// SilhouetteVertex/Edge::findClosestSilhouettePoint (vertex_silhouettes.inl:89-118, edge_silhouettes.inl:112-143)
template <int DIM, class M, class Stack>
NMC_TRAV bool closestSilhouette(const SceneView& S, Stack& stack, V3 x, float r2, bool flip, float sqMinR, float precision, float& dOut) {
	if (S.nNodes == 0) return false;
	if (sqMinR >= r2) return false;
	float b0, b1, tmp;
	bool found = false; int lastId = -1;
	boxSqDist(xyz(S.nodes[0]), xyz(S.nodes[1]), x, b0, tmp);
	if (!(b0 <= r2)) return false;
	stack.put(0, 0, b0);
	int sp = 0;
	// one node per trip, or (NMC_WHILE_WHILE = while) inner nodes until the lane holds a leaf; the nodes are visited in the same order
	int silOffset = 0, nSil = 0;
	while (sp >= 0 || nSil > 0) {
		NMC_WHILE_WHILE (sp >= 0 && nSil == 0) {
			int ni = stack.node(sp); float cd = stack.dist(sp); sp--;
			if (cd > r2) continue;
			int nRefs = asInt(S.nodes[4*ni].w);
			if (nRefs > 0) {
				float4 nd = S.nodes[4*ni + 3];
				silOffset = asInt(nd.y); nSil = asInt(nd.z);
			} else {
				int c0 = ni + 1, c1 = ni + asInt(S.nodes[4*ni + 1].w);
				bool hit0 = false, hit1 = false;
				float4 k0 = S.nodes[4*c0 + 2];
				if (k0.w >= 0.0f) {
					V3 lo = xyz(S.nodes[4*c0]), hi = xyz(S.nodes[4*c0 + 1]);
					boxSqDist(lo, hi, x, b0, tmp);
#if defined(NMC_FAST_GEOM)
					const float4 f0 = S.coneF[c0];
					hit0 = b0 <= r2 && coneOverlapFast(xyz(f0), f0.w, x, lo, hi, b0, 2.0f*precision);
#else
					hit0 = b0 <= r2 && coneOverlap<M>(xyz(k0), k0.w, x, lo, hi, b0);
#endif
				}
				float4 k1 = S.nodes[4*c1 + 2];
				if (k1.w >= 0.0f) {
					V3 lo = xyz(S.nodes[4*c1]), hi = xyz(S.nodes[4*c1 + 1]);
					boxSqDist(lo, hi, x, b1, tmp);
#if defined(NMC_FAST_GEOM)
					const float4 f1 = S.coneF[c1];
					hit1 = b1 <= r2 && coneOverlapFast(xyz(f1), f1.w, x, lo, hi, b1, 2.0f*precision);
#else
					hit1 = b1 <= r2 && coneOverlap<M>(xyz(k1), k1.w, x, lo, hi, b1);
#endif
				}
				if (hit0 && hit1) {
					int closer = c0, other = c1;
					if (b1 < b0) { float t = b0; b0 = b1; b1 = t; closer = c1; other = c0; }
					sp++; stack.put(sp, other, b1);
					sp++; stack.put(sp, closer, b0);
				} else if (hit0) { sp++; stack.put(sp, c0, b0); }
				else if (hit1) { sp++; stack.put(sp, c1, b1); }
			}
		}
#if defined(NMC_FAST_GEOM)
		if (DIM == 3) { silhouetteLeaf3(S, silOffset, nSil, x, flip, sqMinR, precision, r2, found, lastId, dOut); nSil = 0; }
#endif
		{
			for (int p = 0; p < nSil; p++) {
				int ri = silOffset + p;
				V3 viewDir, n0, n1; float d, concavity; int flags, id;
				if (DIM == 2) {
					float4 s0 = S.sils[2*ri], s1 = S.sils[2*ri + 1];
					flags = asInt(s0.z); id = asInt(s0.w);
					if (id == lastId) continue;
					if (sqMinR >= r2) continue;
					viewDir = x - mk(s0.x, s0.y, 0.0f);
#if defined(NMC_FAST_GEOM)
					if (dot(viewDir, viewDir) > r2) continue; // reject on the squared distance before paying for the sqrt
#endif
					d = norm(viewDir);
					n0 = mk(s1.x, s1.y, 0.0f); n1 = mk(s1.z, s1.w, 0.0f);
					concavity = n0.x*n1.y - n1.x*n0.y;
				} else {
					float4 s0 = S.sils[4*ri], s1 = S.sils[4*ri + 1], s2 = S.sils[4*ri + 2], s3 = S.sils[4*ri + 3]; // one round trip for the record
					flags = asInt(s0.w); id = asInt(s1.w);
					if (id == lastId) continue;
					if (sqMinR >= r2) continue;
					n0 = xyz(s2); concavity = s2.w; n1 = xyz(s3);
#if defined(NMC_FAST_GEOM)
					if ((flags & 3) == 3 && !silhouetteCandidate(n0, n1, x - xyz(s0), precision, r2)) continue;
#endif
					V3 pt; float t;
					d = closestOnSegment(xyz(s0), xyz(s1), x, pt, t);
					viewDir = x - pt;
				}
				if (d*d > r2) continue;
				bool isSil = (flags & 3) != 3;
				if (!isSil) isSil = isSilhouette(concavity, n0, n1, viewDir, d, flip, precision);
				if (isSil && d*d <= r2) {
					found = true;
					r2 = minS(r2, d*d);
					dOut = d; lastId = id;
					if (sqMinR >= r2) break;
				}
			}
			nSil = 0;
		}
	}
	return found;
}

#if defined(NMC_FAST_GEOM)
// closestSilhouette() of the default mode over the per-node child blocks (SceneView::treeF): the same tests in the same order,
// but an inner-node visit is ONE round trip of six independent 16-byte loads (the generic form reads the node's child offset
// first and then eight words of the two children plus their cones: two dependent trips, ten scattered words), and a leaf is
// recognised from the sign of its stack entry.  ncu on the big-mesh kernels: dependent-load latency is what bounds them.
template <int DIM, class Stack>
NMC_TRAV bool closestSilhouetteFast(const SceneView& S, Stack& stack, V3 x, float r2, bool flip, float sqMinR, float precision, float& dOut) {
	if (S.nNodes == 0) return false;
	if (sqMinR >= r2) return false;
	float b0, b1, tmp;
	bool found = false; int lastId = -1;
	boxSqDist(xyz(S.nodes[0]), xyz(S.nodes[1]), x, b0, tmp);
	if (!(b0 <= r2)) return false;
	stack.put(0, asInt(S.nodes[0].w) > 0 ? ~0 : 0, b0);
	int sp = 0;
	while (sp >= 0) {
		const int e = stack.node(sp); const float cd = stack.dist(sp); sp--;
		if (cd > r2) continue;
		if (e < 0) { // leaf
			const float4 nd = S.nodes[4*(~e) + 3];
			const int silOffset = asInt(nd.y), nSil = asInt(nd.z);
			if (DIM == 3) silhouetteLeaf3(S, silOffset, nSil, x, flip, sqMinR, precision, r2, found, lastId, dOut);
			else for (int p = 0; p < nSil; p++) { // SilhouetteVertex::findClosestSilhouettePoint, as in closestSilhouette()
				const int ri = silOffset + p;
				const float4 s0 = S.sils[2*ri], s1 = S.sils[2*ri + 1];
				const int flags = asInt(s0.z), id = asInt(s0.w);
				if (id == lastId) continue;
				if (sqMinR >= r2) continue;
				const V3 viewDir = x - mk(s0.x, s0.y, 0.0f);
				if (dot(viewDir, viewDir) > r2) continue;
				const float d = norm(viewDir);
				const V3 n0 = mk(s1.x, s1.y, 0.0f), n1 = mk(s1.z, s1.w, 0.0f);
				if (d*d > r2) continue;
				bool isSil = (flags & 3) != 3;
				if (!isSil) isSil = isSilhouette(n0.x*n1.y - n1.x*n0.y, n0, n1, viewDir, d, flip, precision);
				if (isSil && d*d <= r2) {
					found = true;
					r2 = minS(r2, d*d);
					dOut = d; lastId = id;
					if (sqMinR >= r2) break;
				}
			}
			continue;
		}
		const float4* q = S.treeF + 6*(size_t)e;
		const float4 l0 = q[0], h0 = q[1], a0 = q[2], l1 = q[3], h1 = q[4], a1 = q[5];
		const int c0 = e + 1, c1 = e + asInt(h1.w);
		bool hit0 = false, hit1 = false;
		if (l0.w != 2.0f) { // the subtree holds silhouettes
			boxSqDist(xyz(l0), xyz(h0), x, b0, tmp);
			hit0 = b0 <= r2 && coneOverlapFast(xyz(a0), l0.w, x, xyz(l0), xyz(h0), b0, 2.0f*precision);
		}
		if (l1.w != 2.0f) {
			boxSqDist(xyz(l1), xyz(h1), x, b1, tmp);
			hit1 = b1 <= r2 && coneOverlapFast(xyz(a1), l1.w, x, xyz(l1), xyz(h1), b1, 2.0f*precision);
		}
		const int e0 = asInt(a0.w) > 0 ? ~c0 : c0, e1 = asInt(a1.w) > 0 ? ~c1 : c1;
		if (hit0 && hit1) {
			int closer = e0, other = e1;
			if (b1 < b0) { float t = b0; b0 = b1; b1 = t; closer = e1; other = e0; }
			sp++; stack.put(sp, other, b1);
			sp++; stack.put(sp, closer, b0);
		} else if (hit0) { sp++; stack.put(sp, e0, b0); }
		else if (hit1) { sp++; stack.put(sp, e1, b1); }
	}
	return found;
}
#endif

// ---- flat scans for small scenes (default mode only) ------------------------------------------------------
// With a few dozen primitives the tree bookkeeping (stack traffic, box sorting, cone tests) costs more than it
// saves, and it makes the lanes of a warp diverge.  These scans visit the records group by group (8 in 2D, 4 in
// 3D; the lists are padded to whole groups by scene_build.cpp), skip a group whose box is out of reach, and
// return the same minimum as the traversals above (ties may pick another primitive of equal distance).
// The tables live in shared memory: FlatTab's pointers are derived from the kernel's shared array only, so the
// compiler emits LDS (no generic-address loads).
template <int DIM> struct FlatGroup { static constexpr int n = DIM == 2 ? 8 : 4; };
struct FlatTab {
	const float4 *silsU, *grpS; int nSilU;   // distinct silhouettes, (lo, hi) per group
	const float4 *rayP, *rayN, *grpP; int nRay; // ray primitives as (origin, edge vectors), unit normals, (lo, hi) per group
	const float4 *supS, *supP;               // SUPER scans (large meshes, tables in global memory): (lo, hi) per 32 groups
};
template <int DIM>
NMC_HD float boxSqDistMin(float4 lo, float4 hi, V3 p) {
	float ax = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.0f), ay = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.0f);
	float d = ax*ax + ay*ay;
	if (DIM == 3) { float az = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.0f); d += az*az; }
	return d;
}
NMC_HD int lowestBit(unsigned m) {
#if defined(__CUDA_ARCH__)
	return __ffs((int)m) - 1;
#else
	return __builtin_ffs((int)m) - 1;
#endif
}
// Two-stage closest-silhouette scan (record layout: scene_build.cpp).  Stage 1, unrolled and branch-free, marks
// the records whose two face planes see x from opposite sides (or within the precision band |s| <= precision*d,
// bounded with d <= sqrt(r2)); stage 2 runs SilhouetteVertex/Edge::findClosestSilhouettePoint
// (vertex_silhouettes.inl:89-118, edge_silhouettes.inl:112-143) on the marked records only.  Every lane walks
// its own candidate list, so a warp pays for the longest list, not for the union of the lanes' candidates.
// SUPER: two-level scan for meshes of hundreds to thousands of primitives (tables in global memory, L1/L2-resident):
// a block of 32 groups whose common box is out of reach is skipped as a whole.
template <int DIM, bool SUPER = false>
NMC_HD bool flatClosestSilhouette(const FlatTab& F, V3 x, float r2, bool flip, float sqMinR, float precision, float& dOut) {
	if (sqMinR >= r2) return false;
	constexpr int G = FlatGroup<DIM>::n;
	const float bandW = precision*sqrtf(r2);
	const bool cull = F.nSilU > 4*G; // a handful of records (a box): testing them costs less than culling
	bool found = false;
	for (int g0 = 0, gi = 0; g0 < F.nSilU; g0 += G, gi += 2) {
		if (SUPER && (gi & 63) == 0 && boxSqDistMin<DIM>(F.supS[gi >> 5], F.supS[(gi >> 5) + 1], x) > r2) { g0 += 31*G; gi += 62; continue; }
		// every lane of the warp sits near the same query point, so this cull is nearly warp-coherent
		if (cull && boxSqDistMin<DIM>(F.grpS[gi], F.grpS[gi + 1], x) > r2) continue;
		unsigned cand = 0u;
#pragma unroll
		for (int j = 0; j < G; j++) {
			float s0, s1;
			if (DIM == 2) {
				const float4 q0 = F.silsU[2*(g0 + j)], q1 = F.silsU[2*(g0 + j) + 1];
				s0 = fmaf(q0.x, x.x, fmaf(q0.y, x.y, q0.z)); s1 = fmaf(q0.w, x.x, fmaf(q1.x, x.y, q1.y));
			} else {
				const float4 e0 = F.silsU[4*(g0 + j)], e1 = F.silsU[4*(g0 + j) + 1];
				s0 = fmaf(e0.x, x.x, fmaf(e0.y, x.y, fmaf(e0.z, x.z, e0.w))); s1 = fmaf(e1.x, x.x, fmaf(e1.y, x.y, fmaf(e1.z, x.z, e1.w)));
			}
			const bool c = (s0*s1 < 0.0f) | (fminf(fabsf(s0), fabsf(s1)) <= bandW);
			cand = c ? (cand | (1u << j)) : cand;
		}
		while (cand) {
			const int i = g0 + lowestBit(cand);
			cand &= cand - 1u;
			V3 viewDir, n0, n1; float d2, concavity;
			if (DIM == 2) {
				const float4 q0 = F.silsU[2*i], q1 = F.silsU[2*i + 1];
				viewDir = mk(x.x - q1.z, x.y - q1.w, 0.0f);
				d2 = viewDir.x*viewDir.x + viewDir.y*viewDir.y;
				if (d2 > r2) continue;
				n0 = mk(q0.x, q0.y, 0.0f); n1 = mk(q0.w, q1.x, 0.0f);
				concavity = n0.x*n1.y - n1.x*n0.y;
			} else {
				const float4 e2 = F.silsU[4*i + 2], e3 = F.silsU[4*i + 3];
				V3 pt; float t;
				const float d = closestOnSegment(xyz(e2), xyz(e3), x, pt, t);
				d2 = d*d;
				if (d2 > r2) continue;
				viewDir = x - pt;
				n0 = xyz(F.silsU[4*i]); n1 = xyz(F.silsU[4*i + 1]); concavity = e3.w;
			}
			const bool twoFaces = (n0.x != 0.0f) | (n0.y != 0.0f) | (n0.z != 0.0f);
			if (!twoFaces || isSilhouette(concavity, n0, n1, viewDir, sqrtf(d2), flip, precision)) { found = true; r2 = d2; }
		}
		if (sqMinR >= r2) break;
	}
	if (found) dOut = sqrtf(r2);
	return found;
}
#ifndef NMC_RAY_UNROLL
#define NMC_RAY_UNROLL 1
#endif
template <int DIM, bool SUPER = false>
NMC_HD bool flatRay(const FlatTab& F, V3 o, V3 dir, float tMax, Hit& out) {
	constexpr int G = FlatGroup<DIM>::n;
	constexpr int kRayUnroll = DIM == 3 ? NMC_RAY_UNROLL : 1; // 3D: the plane-form test is branch-free, unrolling buys instruction-level parallelism
	int best = -1; float bu = 0.0f;
	// slab test of the ray segment [0, tMax] against each group's box (fminf/fmaxf drop the NaN of 0*inf).
	// Lanes shoot different rays: each lane first collects the groups ITS ray can reach, then walks its own
	// list, so a warp pays for the longest list and not for the union of the lanes' groups.
	const float ix = 1.0f/dir.x, iy = 1.0f/dir.y, iz = DIM == 3 ? 1.0f/dir.z : 0.0f;
	const float ox = -o.x*ix, oy = -o.y*iy, oz = DIM == 3 ? -o.z*iz : 0.0f;
	auto reachBox = [&](const float4 lo, const float4 hi, float tm) {
		const float a0 = fmaf(lo.x, ix, ox), a1 = fmaf(hi.x, ix, ox), b0 = fmaf(lo.y, iy, oy), b1 = fmaf(hi.y, iy, oy);
		float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), 0.0f), tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), tm);
		if (DIM == 3) {
			const float c0 = fmaf(lo.z, iz, oz), c1 = fmaf(hi.z, iz, oz);
			tn = fmaxf(tn, fminf(c0, c1)); tf = fminf(tf, fmaxf(c0, c1));
		}
		return tn <= tf*1.0001f + 1e-6f; // boxes are exact bounds of the records: leave a rounding margin
	};
	auto reach = [&](int gi, float tm) { return reachBox(F.grpP[gi], F.grpP[gi + 1], tm); };
	const int nGAll = (F.nRay + G - 1)/G; // <= 32 unless SUPER
	const bool cull = F.nRay > 4*G;    // a handful of primitives (a box): testing them costs less than culling
#pragma unroll 1
	for (int sg0 = 0; sg0 < (SUPER ? nGAll : 1); sg0 += 32) {
	if (SUPER && !reachBox(F.supP[sg0 >> 4], F.supP[(sg0 >> 4) + 1], tMax)) continue;
	const int nG = SUPER ? (nGAll - sg0 < 32 ? nGAll - sg0 : 32) : nGAll;
	unsigned todo = nG >= 32 ? 0xffffffffu : (1u << nG) - 1u;
	if (cull) {
		todo = 0u;
#pragma unroll 1
		for (int g = 0; g < nG; g++) todo |= (reach(2*(sg0 + g), tMax) ? 1u : 0u) << g;
	}
	while (todo) {
		const int g = sg0 + lowestBit(todo), g0 = g*G;
		todo &= todo - 1u;
		if (cull && best >= 0 && !reach(2*g, tMax)) continue; // the ray got shorter since the list was made
#pragma unroll kRayUnroll
		for (int j = 0; j < G; j++) {
			const int i = g0 + j;
			if (DIM == 2) { // LineSegment::intersect (line_segments.inl:96-140); the division only runs for hits
				const float4 q = F.rayP[i];
				const float ux = q.x - o.x, uy = q.y - o.y;
				const float dv = dir.x*q.w - dir.y*q.z;
				const float a = ux*dir.y - uy*dir.x, b = ux*q.w - uy*q.z;
				const float adv = fabsf(dv);
				if ((adv > kEps) & (a*dv >= 0.0f) & (fabsf(a) <= adv) & (b*dv >= 0.0f) & (fabsf(b) <= tMax*adv)) { // no short-circuit branches
					const float inv = 1.0f/dv;
					const float s = a*inv, t = b*inv;
					if (s >= 0.0f && s <= 1.0f && t >= 0.0f && t <= tMax) { tMax = t; best = i; bu = s; }
				}
			} else { // Triangle::intersect (triangles.inl:219-256) in plane form (record layout: scene_build.cpp), branch-free
				const float4 r0 = F.rayP[3*i], r1 = F.rayP[3*i + 1], r2 = F.rayP[3*i + 2];
				const float den = fmaf(r0.x, dir.x, fmaf(r0.y, dir.y, r0.z*dir.z));
				const float sd = fmaf(r0.x, o.x, fmaf(r0.y, o.y, fmaf(r0.z, o.z, r0.w)));
				const float t = -sd*(1.0f/den);
				const float px = fmaf(t, dir.x, o.x), py = fmaf(t, dir.y, o.y), pz = fmaf(t, dir.z, o.z);
				const float v = fmaf(r1.x, px, fmaf(r1.y, py, fmaf(r1.z, pz, r1.w)));
				const float w = fmaf(r2.x, px, fmaf(r2.y, py, fmaf(r2.z, pz, r2.w)));
				const bool ok = (fabsf(den) > kEps) & (t >= 0.0f) & (t <= tMax) & (v >= 0.0f) & (w >= 0.0f) & (v + w <= 1.0f);
				tMax = ok ? t : tMax; best = ok ? i : best;
			}
		}
	}
	}
	if (best < 0) return false;
	out.d = tMax; out.ref = best; out.n = xyz(F.rayN[best]);
	if (DIM == 2) { const float4 q = F.rayP[best]; out.p = mk(q.x + bu*q.z, q.y + bu*q.w, 0.0f); out.u = bu; out.v = -1.0f; }
	else { out.p = mk(fmaf(tMax, dir.x, o.x), fmaf(tMax, dir.y, o.y), fmaf(tMax, dir.z, o.z)); out.u = 0.0f; out.v = 0.0f; } // barycentrics are not used by the walk
	return true;
}

// ---- zombie's query adapters (fcpw_scene_loader.h:292-652) --------------------------------------
// computeDistToDirichlet with no Dirichlet geometry: sqrt(d2Max) to the scene box (:299-315)
template <int DIM>
NMC_HD float distDirichlet(const SceneView& S, V3 x) {
	float cx = minS(S.bboxLo[0] - x.x, x.x - S.bboxHi[0]);
	float cy = minS(S.bboxLo[1] - x.y, x.y - S.bboxHi[1]);
	if (DIM == 2) return sqrtf(cx*cx + cy*cy);
	float cz = minS(S.bboxLo[2] - x.z, x.z - S.bboxHi[2]);
	return sqrtf(cx*cx + (cy*cy + cz*cz));
}
// computeDistToNeumann (:316-330) + Interaction::signedDistance (core/interaction.h:32-34)
template <int DIM, class Stack>
NMC_HD float distNeumann(const SceneView& S, Stack& stack, V3 x, bool sgn) {
	if (S.nPrims == 0) return kMaxF;
	Hit h; h.d = kMaxF; h.p = mk(0, 0, 0); h.n = mk(0, 0, 0);
	closestPoint<DIM>(S, stack, x, kMaxF, sgn, h);
	if (!sgn) return h.d;
	return (dot(x - h.p, h.n) > 0.0f ? 1.0f : -1.0f)*h.d;
}
template <int DIM>
NMC_HD float distNeumann(const SceneView& S, V3 x, bool sgn) { LocalStack st; return distNeumann<DIM>(S, st, x, sgn); }
template <int DIM, class Stack>
NMC_HD bool insideDomain(const SceneView& S, Stack& stack, V3 x) { // :642-648
	if (!S.watertight) return true;
	float d1 = distDirichlet<DIM>(S, x);
	float d2 = distNeumann<DIM>(S, stack, x, true);
	return fabsf(d1) < fabsf(d2) ? d1 < 0.0f : d2 < 0.0f;
}
template <int DIM>
NMC_HD bool insideDomain(const SceneView& S, V3 x) { LocalStack st; return insideDomain<DIM>(S, st, x); }
template <int DIM>
NMC_HD bool outsideBox(const SceneView& S, V3 x) { // :649-651
	bool in = x.x >= S.bboxLo[0] && x.x <= S.bboxHi[0] && x.y >= S.bboxLo[1] && x.y <= S.bboxHi[1];
	if (DIM == 3) in = in && x.z >= S.bboxLo[2] && x.z <= S.bboxHi[2];
	return !in;
}
// offsetPointAlongDirection (:252-290): integer-ULP ray-origin offset
NMC_HD float offsetComp(float p, float n) {
	const float origin = 1.0f/32.0f, floatScale = 1.0f/65536.0f, intScale = 256.0f;
	int nOff = (int)(n*intScale);
	float pOff = asFloat(asInt(p) + (p < 0 ? -nOff : nOff));
	return fabsf(p) < origin ? p + floatScale*n : pOff;
}
template <int DIM>
NMC_HD V3 offsetPoint(V3 p, V3 n) {
	return mk(offsetComp(p.x, n.x), offsetComp(p.y, n.y), DIM == 3 ? offsetComp(p.z, n.z) : 0.0f);
}
template <int DIM, class M, class Stack>
NMC_HD float starRadius(const SceneView& S, Stack& stack, V3 x, float minR, float maxR, float prec, bool flipOrient) { // :621-641
	if (minR > maxR) return maxR;
	if (S.nPrims > 0) {
		bool flip = !flipOrient; // FCPW's convention needs flipped normals (:629)
		float r2 = maxR < kMaxF ? maxR*maxR : kMaxF;
		float d;
#if defined(NMC_FAST_GEOM) && !defined(NMC_NO_TREE_BLOCKS)
		if (closestSilhouetteFast<DIM>(S, stack, x, r2, flip, minR*minR, prec, d)) return maxS(d, minR);
#else
		if (closestSilhouette<DIM, M>(S, stack, x, r2, flip, minR*minR, prec, d)) return maxS(d, minR);
#endif
	}
	return maxS(maxR, minR);
}
template <int DIM, class M>
NMC_HD float starRadius(const SceneView& S, V3 x, float minR, float maxR, float prec, bool flipOrient) {
	LocalStack st; return starRadius<DIM, M>(S, st, x, minR, maxR, prec, flipOrient);
}
template <int DIM, class Stack>
NMC_HD bool intersectNeumann(const SceneView& S, Stack& stack, V3 org, V3 nrm, V3 dir, float tMax, bool onB, Hit& h) { // :458-484
	if (S.nPrims == 0) return false;
	V3 o = onB ? offsetPoint<DIM>(org, neg(nrm)) : org;
	if (DIM == 2) { o.z = 0.0f; dir.z = 0.0f; }
	return rayIntersect<DIM>(S, stack, o, dir, tMax, false, h);
}
template <int DIM>
NMC_HD bool intersectNeumann(const SceneView& S, V3 org, V3 nrm, V3 dir, float tMax, bool onB, Hit& h) {
	LocalStack st; return intersectNeumann<DIM>(S, st, org, nrm, dir, tMax, onB, h);
}
// pde.source: nearest-texel lookup (demo/scene.h:194-198 + image.h:70-75; zombie3d scene_3d.h:120-126)
template <int DIM>
NMC_HD float sourceAt(const SceneView& S, V3 x) {
#if defined(NMC_FAST_GEOM)
	{ // default mode: one FMA per axis ((x - lo)/(hi - lo)*n folded into scale and offset); texel boundaries move by an ulp
		int a = (int)fmaf(x.x, S.srcScale[0], S.srcOff[0]), b = (int)fmaf(x.y, S.srcScale[1], S.srcOff[1]);
		if (DIM == 2) { // rows <-> y, columns <-> x
			a = min(max(a, 0), S.n1 - 1); b = min(max(b, 0), S.n0 - 1);
			return S.src[b*S.n1 + a];
		}
		int c = (int)fmaf(x.z, S.srcScale[2], S.srcOff[2]);
		a = min(max(a, 0), S.n0 - 1); b = min(max(b, 0), S.n1 - 1); c = min(max(c, 0), S.n2 - 1);
		return S.src[(a*S.n1 + b)*S.n2 + c];
	}
#endif
	float ux = (x.x - S.bboxLo[0])/(S.bboxHi[0] - S.bboxLo[0]);
	float uy = (x.y - S.bboxLo[1])/(S.bboxHi[1] - S.bboxLo[1]);
	if (DIM == 2) {
		int h = S.n0, w = S.n1;
		int i = (int)(uy*h); i = i < 0 ? 0 : (i > h - 1 ? h - 1 : i);
		int j = (int)(ux*w); j = j < 0 ? 0 : (j > w - 1 ? w - 1 : j);
		return S.src[(size_t)i*w + j];
	}
	float uz = (x.z - S.bboxLo[2])/(S.bboxHi[2] - S.bboxLo[2]);
	int i = (int)(ux*S.n0); i = i < 0 ? 0 : (i > S.n0 - 1 ? S.n0 - 1 : i);
	int j = (int)(uy*S.n1); j = j < 0 ? 0 : (j > S.n1 - 1 ? S.n1 - 1 : j);
	int k = (int)(uz*S.n2); k = k < 0 ? 0 : (k > S.n2 - 1 ? S.n2 - 1 : k);
	return S.src[((size_t)i*S.n1 + j)*S.n2 + k];
}

} // namespace nmc
