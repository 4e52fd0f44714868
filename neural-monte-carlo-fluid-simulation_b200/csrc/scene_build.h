// csrc/scene_build.h -- host-side construction of the flattened boundary structure (see nmc_geom.cuh
// for the record layout).  Pure C++, no CUDA.
#pragma once
#include <vector>
#include <cstdint>

namespace nmc {

struct Q4 { float x, y, z, w; };

struct FlatScene {
	int dim = 0;
	int nNodes = 0, nPrims = 0, nSilRefs = 0, maxDepth = 0;
	std::vector<Q4> nodes;  // 4 per node
	std::vector<Q4> coneF;  // 1 per node: the default mode's own normal cone (axis, w); w = 2: no silhouettes below, w <= 0: no culling,
	                        // 0 < w <= 1: the reference's cone, w = cos(halfAngle); w > 4: own cone, w = 4 + cos(halfAngle) (scene_build.cpp conesFast)
	std::vector<Q4> prims;  // 1 (2D) / 3 (3D) per primitive
	std::vector<Q4> primN;  // 1 per primitive
	std::vector<Q4> nrmV;   // 2 (2D) / 6 (3D) per primitive
	std::vector<Q4> sils;   // 2 (2D) / 4 (3D) per silhouette reference
	std::vector<Q4> treeF;  // default mode: 6 per node -- an inner node's two children as (lo, cone code), (hi, -), (cone axis, nRefs of the child) each
	std::vector<Q4> silsF;  // 3D, default mode: 2 per silhouette reference, the two face planes (n, -n.pa) of the plane-side prefilter (zeros: always a candidate)
	std::vector<Q4> silsU;  // the same records, one per distinct silhouette (flat scan of small scenes)
	int nSilU = 0;
	// flat-scan culling boxes: one (lo, hi) pair per group of 8 consecutive primitives / distinct silhouettes
	std::vector<Q4> grpP, grpS;
	// second level for meshes beyond the shared-memory flat scan: one (lo, hi) pair per 32 consecutive groups
	std::vector<Q4> supP, supS;
	// ray-scan primitives of the default mode: in 2D, chains of connected collinear segments are merged into one
	// segment (a subdivided straight wall is one ray target); in 3D a copy of `prims`.  rayN: unit normals.
	std::vector<Q4> rayP, rayN;
	int nRay = 0;
	float bboxLo[3] = {0, 0, 0}, bboxHi[3] = {0, 0, 0};
};

// verts nV x dim, prims nP x dim. ignoreConvex: the bindings' ignoreCandidateSilhouette
// (demo/scene.h:84-90) = !isDoubleSided.
void buildFlatScene(int dim, const float* verts, int nV, const int* prims, int nP, bool doubleSided, FlatScene& out);

} // namespace nmc
