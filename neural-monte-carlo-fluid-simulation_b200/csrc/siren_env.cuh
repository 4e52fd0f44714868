// csrc/siren_env.cuh -- the boundary envelope of query_velocity (src/2d/models/base.py:158-224,
// src/3d/models/base.py:172-260) evaluated inside the SIREN kernels.  Shared by siren.cu and siren_tc.cu.
//
// For output component j at sample x, in the reference's order of operations:
//   1. region override   v_j = (x in region && bit j of regionMask) ? regionVel[j] : net_j      (inlet strip / inlet ball)
//   2. obstacle weight   v_j *= clamp(|x - c| - r, 0, eps)/eps                                     (smoothstep_circular_obs, NOT detached)
//   3. wall weights      v_j *= min(clamp|x_j - lo_j|, clamp|x_j - hi_j|)/eps if bit j of wallMask (detached)
// taylorgreen / vortex_collide: walls on every component.  karman: region = inlet strip on u, obstacle = cylinder,
// wall on v only.  smoke_obs: region = inlet ball on w, obstacle sphere, walls on every component.
#pragma once
#include "../../include/nmcfs_siren.h"

namespace nmc_siren_detail {

struct Env {
	int active;          // 0: identity
	int wallMask;
	float lo[3], hi[3], eps;
	int sphere; float sc[3], sr;
	int region;          // 0 none, 1 box [rlo, rhi], 2 ball (centre rlo, radius rhi[0])
	int regionMask;
	float rlo[3], rhi[3], rvel[3];
	int sphereAxes;      // coordinates entering the obstacle distance (7: sphere / circle; 5: cylinder along y, karman3d)
	float rnoise[3];     // region value = rvel[j] + rnoise[j] * u(x, seed), u uniform in [-1, 1) (3D smoke inlet)
	const unsigned* noiseSeed; // device pointer (the time step), may be null
};

// returns 0, or a static message describing what is wrong with `e`
inline const char* toEnv(const nmc_siren_envelope* e, Env& v) {
	v = Env();
	v.eps = 1.0f;
	if (!e || e->kind == 0) return nullptr;
	if (e->kind != 1 && e->kind != 2) return "unknown envelope kind";
	if (!(e->eps > 0.0f)) return "envelope eps must be positive";
	v.active = 1; v.eps = e->eps;
	for (int i = 0; i < 3; i++) { v.lo[i] = e->lo[i]; v.hi[i] = e->hi[i]; }
	if (e->kind == 1) { v.wallMask = 7; return nullptr; }
	v.wallMask = e->wall_mask & 7;
	v.sphere = e->has_sphere != 0; v.sr = e->sphere_r;
	v.region = e->region_kind; v.regionMask = e->region_mask & 7;
	if (v.region < 0 || v.region > 2) return "unknown envelope region kind";
	for (int i = 0; i < 3; i++) { v.sc[i] = e->sphere_c[i]; v.rlo[i] = e->region_lo[i]; v.rhi[i] = e->region_hi[i]; v.rvel[i] = e->region_vel[i]; v.rnoise[i] = e->region_noise[i]; }
	v.sphereAxes = (e->sphere_axes & 7) ? (e->sphere_axes & 7) : 7;
	v.noiseSeed = e->noise_seed;
	return nullptr;
}

#if defined(__CUDACC__)
__device__ __forceinline__ float envWall(const Env& e, int i, float xi) {
	float a = fminf(fmaxf(fabsf(xi - e.lo[i]), 0.0f), e.eps), b = fminf(fmaxf(fabsf(xi - e.hi[i]), 0.0f), e.eps);
	return fminf(a, b)/e.eps;
}
__device__ __forceinline__ bool envInRegion(const Env& e, int inDim, const float* x) {
	if (e.region == 1) {
		bool in = true;
		for (int i = 0; i < 3; i++) if (i < inDim) in = in && x[i] >= e.rlo[i] && x[i] <= e.rhi[i];
		return in;
	}
	if (e.region == 2) {
		float d2 = 0.0f;
		for (int i = 0; i < 3; i++) if (i < inDim) { float d = x[i] - e.rlo[i]; d2 += d*d; }
		return sqrtf(d2) < e.rhi[0];
	}
	return false;
}
// obstacle weight and (optionally) its gradient with respect to x
__device__ __forceinline__ float envObstacle(const Env& e, int inDim, const float* x, float* grad) {
	if (grad) { grad[0] = grad[1] = grad[2] = 0.0f; }
	if (!e.sphere) return 1.0f;
	float d[3] = {0.0f, 0.0f, 0.0f}, d2 = 0.0f;
	for (int i = 0; i < 3; i++) if (i < inDim && ((e.sphereAxes >> i) & 1)) { d[i] = x[i] - e.sc[i]; d2 += d[i]*d[i]; }
	float r = sqrtf(d2), dist = r - e.sr;
	if (grad && dist > 0.0f && dist < e.eps && r > 0.0f) for (int i = 0; i < 3; i++) grad[i] = d[i]/(r*e.eps);
	return fminf(fmaxf(dist, 0.0f), e.eps)/e.eps;
}
// Inlet noise of the 3D smoke branch (src/3d/models/base.py:203-209: numpy re-seeded with the time step on every call,
// one uniform number per sample inside the inlet ball, shared by the three components).  Here: a hash of the sample's
// coordinate bits and the time step -- a fixed function of position within a step, re-drawn every step.
__device__ __forceinline__ float envNoise(const Env& e, int inDim, const float* x) {
	if (e.rnoise[0] == 0.0f && e.rnoise[1] == 0.0f && e.rnoise[2] == 0.0f) return 0.0f;
	unsigned h = e.noiseSeed ? *e.noiseSeed*0x9E3779B9u + 0x7F4A7C15u : 0x7F4A7C15u;
	for (int i = 0; i < 3; i++) if (i < inDim) {
		h ^= __float_as_uint(x[i]); h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
	}
	return (float)(h >> 8)*(2.0f/16777216.0f) - 1.0f;
}
// forward: y[0..outDim) holds the network output on entry, the enveloped velocity on return
__device__ __forceinline__ void envForward(const Env& e, int inDim, int outDim, const float* x, float* y) {
	if (!e.active) return;
	const bool in = envInRegion(e, inDim, x);
	const float wo = envObstacle(e, inDim, x, nullptr);
	const float un = in ? envNoise(e, inDim, x) : 0.0f;
	for (int j = 0; j < 3; j++) if (j < outDim) {
		float v = (in && ((e.regionMask >> j) & 1)) ? fmaf(e.rnoise[j], un, e.rvel[j]) : y[j];
		v *= wo;
		if (((e.wallMask >> j) & 1) && j < inDim) v *= envWall(e, j, x[j]);
		y[j] = v;
	}
}
// backward: gy[] holds dL/d(enveloped output) on entry and dL/d(network output) on return; gxExtra (may be null)
// receives the gradient that reaches x through the obstacle weight, which needs the network output ynet[]
__device__ __forceinline__ void envBackward(const Env& e, int inDim, int outDim, const float* x, const float* ynet, float* gy, float* gxExtra) {
	if (gxExtra) { gxExtra[0] = gxExtra[1] = gxExtra[2] = 0.0f; }
	if (!e.active) return;
	const bool in = envInRegion(e, inDim, x);
	float gw[3];
	const float wo = envObstacle(e, inDim, x, gxExtra ? gw : nullptr);
	float through = 0.0f; // sum_j gy_j * wall_j * v_j
	const float un = (in && gxExtra && ynet) ? envNoise(e, inDim, x) : 0.0f;
	for (int j = 0; j < 3; j++) if (j < outDim) {
		const bool over = in && ((e.regionMask >> j) & 1);
		const float wall = (((e.wallMask >> j) & 1) && j < inDim) ? envWall(e, j, x[j]) : 1.0f;
		if (gxExtra && ynet) through += gy[j]*wall*(over ? fmaf(e.rnoise[j], un, e.rvel[j]) : ynet[j]);
		gy[j] = over ? 0.0f : gy[j]*wo*wall;
	}
	if (gxExtra) for (int i = 0; i < 3; i++) gxExtra[i] = through*gw[i];
}
#endif

} // namespace nmc_siren_detail
