// csrc/nmc_math.cuh -- scalar building blocks shared by the deterministic and the fast estimator:
// 3-vectors, the pcg32 generator, the seeding rule, and the "exact" transcendental wrappers.
//
// The deterministic mode has to reproduce the reference's float results (glibc logf/expf/sinf/...
// and bessel.hpp's double polynomials).  On the device every such function is evaluated in
// double and rounded once to float, which agrees with glibc's (almost always correctly rounded)
// float functions except in rare double-rounding cases; + - * / sqrt are IEEE on both sides as
// long as this header is compiled with -fmad=false (see csrc/wost_det.cu).
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define NMC_HD __host__ __device__ __forceinline__
#define NMC_D __device__ __forceinline__
#else
#define NMC_HD inline
#define NMC_D inline
#endif
// NMC_FAST_GEOM (defined by wost_fast.cu only): the default mode needs statistical, not bitwise, agreement,
// so min/max map to single FMNMX instructions, vector/scalar divisions become one reciprocal + multiplies,
// the normal-cone test is evaluated without inverse trigonometric functions, and the three BVH traversals
// are kept out of line (one copy each) to keep the kernel inside the instruction cache.
#if defined(NMC_FAST_GEOM) && defined(__CUDACC__) && !defined(NMC_TRAV_INLINE)
#define NMC_TRAV __host__ __device__ __noinline__
#else
#define NMC_TRAV NMC_HD
#endif
// NMC_OUTLINE: one out-of-line copy in the default-mode kernels (instruction-cache footprint, see profiles/README.md)
#if defined(NMC_FAST_GEOM) && defined(__CUDACC__) && !defined(NMC_NO_OUTLINE)
#define NMC_OUTLINE __host__ __device__ __noinline__
#else
#define NMC_OUTLINE NMC_HD
#endif

namespace nmc {

#if !defined(__CUDA_ARCH__)
// host stand-ins (found before CUDA's device intrinsics by name lookup inside this namespace) so that the
// headers also compile for the host pass and for the CPU-side unit harness (tests/host_emu)
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float rsqrtf(float x) { return 1.0f/sqrtf(x); }
#endif

static constexpr float kEps = FLT_EPSILON;
static constexpr float kMaxF = FLT_MAX;
static constexpr float kShrink = 0.99f;          // RADIUS_SHRINK_PERCENTAGE, walk_on_stars.h:9
static constexpr double kPi = 3.14159265358979323846;
static constexpr double kPi2 = 1.57079632679489661923;

struct V3 { float x, y, z; };

NMC_HD V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
NMC_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
NMC_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
NMC_HD V3 operator*(V3 a, float s) { return mk(a.x*s, a.y*s, a.z*s); }
NMC_HD V3 operator*(float s, V3 a) { return mk(s*a.x, s*a.y, s*a.z); }
#if defined(NMC_FAST_GEOM)
NMC_HD V3 operator/(V3 a, float s) { float r = 1.0f/s; return mk(a.x*r, a.y*r, a.z*r); }
#else
NMC_HD V3 operator/(V3 a, float s) { return mk(a.x/s, a.y/s, a.z/s); }
#endif
NMC_HD V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// std::min / std::max argument order and NaN behaviour
#if defined(NMC_FAST_GEOM)
NMC_HD float minS(float a, float b) { return fminf(a, b); }
NMC_HD float maxS(float a, float b) { return fmaxf(a, b); }
#else
NMC_HD float minS(float a, float b) { return (b < a) ? b : a; }
NMC_HD float maxS(float a, float b) { return (a < b) ? b : a; }
#endif
// Eigen fixed-size reductions associate as x0 + (x1 + x2)  (Eigen/src/Core/Redux.h:99-113)
NMC_HD float dot(V3 a, V3 b) { return a.x*b.x + (a.y*b.y + a.z*b.z); }
NMC_HD float norm(V3 a) { return sqrtf(dot(a, a)); }
NMC_HD V3 cross(V3 a, V3 b) { return mk(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x); }
NMC_HD V3 normalized(V3 a) { float z = dot(a, a); if (z > 0.0f) { float s = sqrtf(z); return a/s; } return a; }
NMC_HD float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
NMC_HD void setComp(V3& a, int k, float v) { if (k == 0) a.x = v; else if (k == 1) a.y = v; else a.z = v; }

NMC_HD float asFloat(int a) {
#if defined(__CUDA_ARCH__)
	return __int_as_float(a);
#else
	union { int i; float f; } u; u.i = a; return u.f;
#endif
}
NMC_HD int asInt(float a) {
#if defined(__CUDA_ARCH__)
	return __float_as_int(a);
#else
	union { float f; int i; } u; u.f = a; return u.i;
#endif
}

// ---- exact-mode transcendentals: double evaluation, one rounding -------------------------------
struct ExactMath {
	static NMC_HD float exp_(float x) { return (float)::exp((double)x); }
	static NMC_HD float log_(float x) { return (float)::log((double)x); }
	static NMC_HD float sin_(float x) { return (float)::sin((double)x); }
	static NMC_HD float cos_(float x) { return (float)::cos((double)x); }
	static NMC_HD float cbrt_(float x) { return (float)::cbrt((double)x); }
	static NMC_HD float acos_(float x) { return (float)::acos((double)x); }
	static NMC_HD float asin_(float x) { return (float)::asin((double)x); }
	static NMC_HD float atan2_(float y, float x) { return (float)::atan2((double)y, (double)x); }
};
// ---- fast-mode transcendentals: fp32 -----------------------------------------------------------
struct FastMath {
	static NMC_HD float exp_(float x) { return ::expf(x); }
	static NMC_HD float log_(float x) { return ::logf(x); }
	static NMC_HD float sin_(float x) { return ::sinf(x); }
	static NMC_HD float cos_(float x) { return ::cosf(x); }
	static NMC_HD float cbrt_(float x) { return ::cbrtf(x); }
	static NMC_HD float acos_(float x) { return ::acosf(x); }
	static NMC_HD float asin_(float x) { return ::asinf(x); }
	static NMC_HD float atan2_(float y, float x) { return ::atan2f(y, x); }
};

// ---- pcg32 (reference: deps/pcg32/pcg32.h:40-112), bit-exact -----------------------------------
struct Pcg32 {
	uint64_t state, inc;
	NMC_HD uint32_t nextUInt() {
		uint64_t old = state;
		state = old*0x5851f42d4c957f2dULL + inc;
		uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
		uint32_t rot = (uint32_t)(old >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	NMC_HD void seed(uint64_t initstate, uint64_t initseq = 1u) {
		state = 0u; inc = (initseq << 1u) | 1u;
		nextUInt(); state += initstate; nextUInt();
	}
	NMC_HD uint32_t nextBounded(uint32_t bound) {
		uint32_t threshold = (~bound + 1u) % bound;
		for (;;) { uint32_t r = nextUInt(); if (r >= threshold) return r % bound; }
	}
	NMC_HD float nextFloat() { return asFloat((int)((nextUInt() >> 9) | 0x3f800000u)) - 1.0f; }
};

NMC_HD uint64_t splitmix64(uint64_t z) {
	z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27))*0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
// seeding rule of the deterministic mode (include/nmcfs.h: nmc_point_seed)
NMC_HD uint64_t pointSeed(uint64_t seed, uint64_t index) {
	return splitmix64(seed + 0x9E3779B97F4A7C15ull*(index + 1ull));
}

// Keyed bijection on [0, n): multiply/xorshift rounds on the enclosing power of two with cycle walking.
NMC_HD unsigned permute(unsigned i, unsigned n, unsigned key) {
	unsigned w = n - 1;
	w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
	do {
		i ^= key; i *= 0xe170893du;
		i ^= key >> 16;
		i ^= (i & w) >> 4;
		i ^= key >> 8; i *= 0x0929eb3fu;
		i ^= key >> 23;
		i ^= (i & w) >> 1; i *= 1u | key >> 27;
		i *= 0x6935fa69u;
		i ^= (i & w) >> 11; i *= 0x74dcb303u;
		i ^= (i & w) >> 2; i *= 0x9e501cc3u;
		i ^= (i & w) >> 2; i *= 0xc860a3dfu;
		i &= w;
		i ^= i >> 5;
	} while (i >= n);
	return i;
}

// sampleUnitSphereUniform<DIM>(float* u)  (reference: include/zombie/core/sampling.h:29-45)
template <int DIM, class M>
NMC_HD V3 sphereDir(float u0, float u1) {
#if defined(NMC_FAST_GEOM) && defined(__CUDA_ARCH__)
	{ // default mode: the SFU's sin/cos on [-pi, pi) (abs. error 2^-21, two instructions each):
	  // cos(2 pi u) = -cos(2 pi (u - 1/2)), likewise sin
		const float a0 = 6.2831853f*((DIM == 2 ? u0 : u1) - 0.5f);
		const float sn = -__sinf(a0), cs = -__cosf(a0);
		if (DIM == 2) return mk(cs, sn, 0.0f);
		float z = 1.0f - 2.0f*u0;
		float r = sqrtf(fmaxf(0.0f, 1.0f - z*z));
		return mk(r*cs, r*sn, z);
	}
#endif
	if (DIM == 2) {
		float phi = (float)(2.0f*kPi*u0);
		return mk(M::cos_(phi), M::sin_(phi), 0.0f);
	}
	float z = 1.0f - 2.0f*u0;
	float r = sqrtf(maxS(0.0f, 1.0f - z*z));
	float phi = (float)(2.0f*kPi*u1);
	return mk(r*M::cos_(phi), r*M::sin_(phi), z);
}
// pdfSampleSphereUniform<DIM>(r)  (sampling.h:55-65): double expression narrowed once
template <int DIM>
NMC_HD float pdfSphere(float r) {
	return DIM == 2 ? (float)(1.0f/(2.0f*kPi*r)) : (float)(1.0f/(4.0f*kPi*r*r));
}

// sampleUnitHemisphereCosine<DIM>(float* u) (sampling.h:113-154): 2D (2u - 1, sqrt(1 - .^2)); 3D the concentric disk map
// lifted to the hemisphere around +z.  The reference evaluates the disk angle in double (float * M_PI) and narrows it once.
template <int DIM, class M>
NMC_HD V3 cosineHemisphere(float u0, float u1) {
	if (DIM == 2) {
		float a = 2.0f*u0 - 1.0f;
		return mk(a, sqrtf(maxS(0.0f, 1.0f - a*a)), 0.0f);
	}
	float a = 2.0f*u0 - 1.0f, b = 2.0f*u1 - 1.0f;
	float dx = 0.0f, dy = 0.0f;
	if (!(a == 0.0f && b == 0.0f)) {
		float r, theta;
		if (fabsf(a) > fabsf(b)) { r = a; theta = (float)(0.25f*kPi*(b/a)); }
		else { r = b; theta = (float)(0.5f*kPi*(1.0f - 0.5f*(a/b))); }
		dx = r*M::cos_(theta); dy = r*M::sin_(theta);
	}
	return mk(dx, dy, sqrtf(maxS(0.0f, 1.0f - (dx*dx + dy*dy))));
}
// pdfSampleUnitHemisphereCosine<DIM>(cosTheta) (sampling.h:163-173)
template <int DIM>
NMC_HD float pdfCosineHemisphere(float c) { return DIM == 2 ? c/2.0f : (float)(c/kPi); }
// transformCoordinates<DIM>(n, d) (sampling.h:181-204): the local frame's last axis becomes n
template <int DIM>
NMC_HD V3 toFrame(V3 n, V3 d) {
	if (DIM == 2) return mk(d.x*n.y + d.y*n.x, d.x*(-n.x) + d.y*n.y, 0.0f);
	const float sign = copysignf(1.0f, n.z);
	const float a = -1.0f/(sign + n.z);
	const float b = n.x*n.y*a;
	const V3 b1 = mk(1.0f + sign*n.x*n.x*a, sign*b, -sign*n.x), b2 = mk(b, sign + n.y*n.y*a, -n.y);
	return mk(d.x*b1.x + d.y*b2.x + d.z*n.x, d.x*b1.y + d.y*b2.y + d.z*n.y, d.x*b1.z + d.y*b2.z + d.z*n.z);
}

} // namespace nmc
