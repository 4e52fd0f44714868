/* oracle/nmc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded-per-point restatement of the reference's Monte Carlo
 * pressure-projection path.  Every block cites the reference file:line it follows
 * (paths relative to /root/reference/bindings/zombie, identical under bindings/zombie3d).
 * Pinned bit-for-bit against the reference's own code (oracle/_ref, built by oracle/Makefile
 * from oracle/ref_harness.cpp) by tests/test_oracle_vs_ref.py, and against the committed
 * vectors under tests/golden/ by tests/test_oracle_golden.py.
 *
 * Arithmetic conventions that matter for bit parity (and that the CUDA path mirrors):
 *  - no FMA contraction (-ffp-contract=off); Eigen fixed-size reductions associate as
 *    x0 + (x1 + x2) (Eigen/src/Core/Redux.h:99-113); min/max follow std::min/std::max;
 *  - every expression that touches M_PI is evaluated in double and narrowed once, EXCEPT
 *    Eigen "vector / double" which narrows the scalar to float first;
 *  - Bessel polynomials in double, results narrowed to float members (distributions.h:581-588);
 *  - geometry is carried in 3-vectors with z = 0 in 2D, as the reference does inside FCPW.
 *
 * Deterministic seeding rule (the reference itself seeds from the wall clock,
 * walk_on_stars.h:498,639; see ref_harness.cpp): point i uses pcg32(nmo_point_seed(seed, i), 1),
 * every antithetic pair draws its walk seed as the next nextUInt() of the point's own stream.
 *
 * Deliberate omission: the Neumann boundary sample (walk_on_stars.h:212-260) consumes DIM
 * random numbers but its contribution is multiplied by pde.neumann == 0 in both bindings
 * (demo/scene.h:176-181, scene_3d.h:107-110), so only the RNG consumption is restated.
 */
#include "nmc_oracle.h"

#include <math.h>
#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define EPS FLT_EPSILON
#define MAXF FLT_MAX
#define MINF (-FLT_MAX)
#define SHRINK 0.99f

typedef float v3[3];

static inline float fmin_std(float a, float b) { return (b < a) ? b : a; } /* std::min(a,b) */
static inline float fmax_std(float a, float b) { return (a < b) ? b : a; } /* std::max(a,b) */
static inline float dot3(const float* a, const float* b) { return a[0]*b[0] + (a[1]*b[1] + a[2]*b[2]); }
static inline float norm3(const float* a) { return sqrtf(dot3(a, a)); }
static inline void sub3(const float* a, const float* b, float* o) { o[0] = a[0]-b[0]; o[1] = a[1]-b[1]; o[2] = a[2]-b[2]; }
static inline void cross3(const float* a, const float* b, float* o) {
	float x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
	o[0] = x; o[1] = y; o[2] = z;
}
/* Eigen normalized(): n / sqrt(squaredNorm) when squaredNorm > 0 (Eigen/src/Core/Dot.h:124-134) */
static inline void normalize3(float* a) {
	float z = dot3(a, a);
	if (z > 0.0f) { float s = sqrtf(z); a[0] /= s; a[1] /= s; a[2] /= s; }
}

/* ---- pcg32 (deps/pcg32/pcg32.h:40-112) -------------------------------------------------- */
#define PCG32_MULT 0x5851f42d4c957f2dULL
uint32_t nmo_pcg32_uint(nmo_pcg32* s) {
	uint64_t old = s->state;
	s->state = old*PCG32_MULT + s->inc;
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
void nmo_pcg32_seed(nmo_pcg32* s, uint64_t initstate, uint64_t initseq) {
	s->state = 0u; s->inc = (initseq << 1u) | 1u;
	nmo_pcg32_uint(s); s->state += initstate; nmo_pcg32_uint(s);
}
uint32_t nmo_pcg32_bounded(nmo_pcg32* s, uint32_t bound) {
	uint32_t threshold = (~bound + 1u) % bound;
	for (;;) { uint32_t r = nmo_pcg32_uint(s); if (r >= threshold) return r % bound; }
}
float nmo_pcg32_float(nmo_pcg32* s) {
	union { uint32_t u; float f; } x;
	x.u = (nmo_pcg32_uint(s) >> 9) | 0x3f800000u;
	return x.f - 1.0f;
}
/* splitmix64 finaliser; same rule in ref_harness.cpp:75-82 and the CUDA library */
uint64_t nmo_point_seed(uint64_t seed, uint64_t index) {
	uint64_t z = seed + 0x9E3779B97F4A7C15ull*(index + 1ull);
	z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27))*0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

/* ---- samplers (include/zombie/core/sampling.h) ------------------------------------------- */
/* generateStratifiedSamples<D>(samples, n, sampler), sampling.h:434-457; D = dim-1 */
static void stratified(int D, int n, nmo_pcg32* s, float* out) {
	const float oneMinusEps = 1.0f - EPS;
	float inv = 1.0f/n;
	for (int i = 0; i < n; i++) for (int j = 0; j < D; j++) {
		float sj = (i + nmo_pcg32_float(s))*inv;
		out[D*i + j] = fmin_std(sj, oneMinusEps);
	}
	for (int i = 0; i < D; i++) for (int j = 0; j < n; j++) {
		int other = j + (int)nmo_pcg32_bounded(s, (uint32_t)(n - j));
		float t = out[D*j + i]; out[D*j + i] = out[D*other + i]; out[D*other + i] = t;
	}
}
void nmo_stratified(int dim, uint64_t initstate, int nSamples, float* out, uint64_t* state_out) {
	nmo_pcg32 s; nmo_pcg32_seed(&s, initstate, 1);
	stratified(dim - 1, nSamples, &s, out);
	state_out[0] = s.state; state_out[1] = s.inc;
}
/* sampleUnitSphereUniform<DIM>(float* u), sampling.h:29-45 */
static void sphere_dir(int dim, const float* u, float* d) {
	if (dim == 2) {
		float phi = (float)(2.0f*M_PI*u[0]);
		d[0] = cosf(phi); d[1] = sinf(phi); d[2] = 0.0f;
	} else {
		float z = 1.0f - 2.0f*u[0];
		float r = sqrtf(fmax_std(0.0f, 1.0f - z*z));
		float phi = (float)(2.0f*M_PI*u[1]);
		d[0] = r*cosf(phi); d[1] = r*sinf(phi); d[2] = z;
	}
}
void nmo_sphere_dir(int dim, const float* u, int n, float* out) {
	for (int i = 0; i < n; i++) {
		v3 d; sphere_dir(dim, u + (size_t)i*(dim - 1), d);
		for (int k = 0; k < dim; k++) out[(size_t)i*dim + k] = d[k];
	}
}
/* sampleUnitHemisphereCosine<DIM>(float* u), sampling.h:113-154 (concentric disk map :122-146) */
static void cosine_hemisphere(int dim, const float* u, float* d) {
	if (dim == 2) {
		float u1 = 2.0f*u[0] - 1.0f;
		d[0] = u1; d[1] = sqrtf(fmax_std(0.0f, 1.0f - u1*u1)); d[2] = 0.0f;
		return;
	}
	float u1 = 2.0f*u[0] - 1.0f, u2 = 2.0f*u[1] - 1.0f;
	float dx = 0.0f, dy = 0.0f;
	if (!(u1 == 0 && u2 == 0)) {
		float theta, r;
		if (fabsf(u1) > fabsf(u2)) { r = u1; theta = 0.25f*M_PI*(u2/u1); }
		else { r = u2; theta = 0.5f*M_PI*(1.0f - 0.5f*(u1/u2)); }
		dx = r*cosf(theta); dy = r*sinf(theta);
	}
	d[0] = dx; d[1] = dy; d[2] = sqrtf(fmax_std(0.0f, 1.0f - (dx*dx + dy*dy)));
}
/* pdfSampleUnitHemisphereCosine<DIM>(cosTheta), sampling.h:163-173 */
static inline float pdf_cosine_hemisphere(int dim, float c) { return dim == 2 ? c/2.0f : (float)(c/M_PI); }
/* transformCoordinates<DIM>(n, d), sampling.h:181-204 */
static void transform_coordinates(int dim, const float* n, float* d) {
	if (dim == 2) {
		float sx = n[1], sy = -n[0];
		float x = d[0]*sx + d[1]*n[0], y = d[0]*sy + d[1]*n[1];
		d[0] = x; d[1] = y;
		return;
	}
	float sign = copysignf(1.0f, n[2]);
	const float a = -1.0f/(sign + n[2]);
	const float b = n[0]*n[1]*a;
	float b1[3] = {1.0f + sign*n[0]*n[0]*a, sign*b, -sign*n[0]}, b2[3] = {b, sign + n[1]*n[1]*a, -n[1]};
	float r[3];
	for (int k = 0; k < 3; k++) r[k] = d[0]*b1[k] + d[1]*b2[k] + d[2]*n[k];
	d[0] = r[0]; d[1] = r[1]; d[2] = r[2];
}
/* pdfSampleSphereUniform<DIM>(r), sampling.h:55-65 */
static inline float pdf_sphere(int dim, float r) {
	return dim == 2 ? (float)(1.0f/(2.0f*M_PI*r)) : (float)(1.0f/(4.0f*M_PI*r*r));
}

/* ---- bessel (deps/bessel/bessel.hpp:373-556) --------------------------------------------- */
static double bessi0(double x) {
	double ax, ans, y;
	if ((ax = fabs(x)) < 3.75) {
		y = x/3.75; y = y*y;
		ans = 1.0+y*(3.5156229+y*(3.0899424+y*(1.2067492+y*(0.2659732+y*(0.360768e-1+y*0.45813e-2)))));
	} else {
		y = 3.75/ax;
		ans = (exp(ax)/sqrt(ax))*(0.39894228+y*(0.1328592e-1+y*(0.225319e-2+y*(-0.157565e-2+y*(0.916281e-2
			+y*(-0.2057706e-1+y*(0.2635537e-1+y*(-0.1647633e-1+y*0.392377e-2))))))));
	}
	return ans;
}
static double bessi1(double x) {
	double ax, ans, y;
	if ((ax = fabs(x)) < 3.75) {
		y = x/3.75; y = y*y;
		ans = ax*(0.5+y*(0.87890594+y*(0.51498869+y*(0.15084934+y*(0.2658733e-1+y*(0.301532e-2+y*0.32411e-3))))));
	} else {
		y = 3.75/ax;
		ans = 0.2282967e-1+y*(-0.2895312e-1+y*(0.1787654e-1-y*0.420059e-2));
		ans = 0.39894228+y*(-0.3988024e-1+y*(-0.362018e-2+y*(0.163801e-2+y*(-0.1031555e-1+y*ans))));
		ans *= (exp(ax)/sqrt(ax));
	}
	return x < 0.0 ? -ans : ans;
}
static double bessk0(double x) {
	double y, ans;
	if (x <= 2.0) {
		y = x*x/4.0;
		ans = (-log(x/2.0)*bessi0(x))+(-0.57721566+y*(0.42278420+y*(0.23069756+y*(0.3488590e-1+y*(0.262698e-2
			+y*(0.10750e-3+y*0.74e-5))))));
	} else {
		y = 2.0/x;
		ans = (exp(-x)/sqrt(x))*(1.25331414+y*(-0.7832358e-1+y*(0.2189568e-1+y*(-0.1062446e-1+y*(0.587872e-2
			+y*(-0.251540e-2+y*0.53208e-3))))));
	}
	return ans;
}
static double bessk1(double x) {
	double y, ans;
	if (x <= 2.0) {
		y = x*x/4.0;
		ans = (log(x/2.0)*bessi1(x))+(1.0/x)*(1.0+y*(0.15443144+y*(-0.67278579+y*(-0.18156897+y*(-0.1919402e-1
			+y*(-0.110404e-2+y*(-0.4686e-4)))))));
	} else {
		y = 2.0/x;
		ans = (exp(-x)/sqrt(x))*(1.25331414+y*(0.23498619+y*(-0.3655620e-1+y*(0.1504268e-1+y*(-0.780353e-2
			+y*(0.325614e-2+y*(-0.68245e-3)))))));
	}
	return ans;
}
void nmo_bessel(int kind, const double* x, int n, double* out) {
	for (int i = 0; i < n; i++) {
		switch (kind) {
			case 0: out[i] = bessi0(x[i]); break;
			case 1: out[i] = bessi1(x[i]); break;
			case 2: out[i] = bessk0(x[i]); break;
			case 3: out[i] = bessk1(x[i]); break;
			default: { /* bessk(2, x), bessel.hpp:584-612 */
				double tox = 2.0/x[i], bkm = bessk0(x[i]), bk = bessk1(x[i]);
				out[i] = bkm + 1*tox*bk;
			}
		}
	}
}

/* ---- ball Green's functions (include/zombie/core/distributions.h:273-832) ----------------- */
typedef struct {
	int dim, yukawa;
	float lambda, sqrtLambda;
	v3 c, yVol, ySurf;
	float R, r, rClamp;
	float muR, K0muR, I0muR, K1muR, I1muR;      /* Yukawa 2D :695 */
	float expmuR, sinhmuR, K32muR, I32muR;      /* Yukawa 3D :831 */
} ball_t;

static void ball_init(ball_t* g, int dim, int yukawa, float lambda) {
	memset(g, 0, sizeof(*g));
	g->dim = dim; g->yukawa = yukawa; g->lambda = lambda; g->sqrtLambda = sqrtf(lambda);
	g->rClamp = 1e-4f;
}
/* updateBall: :285-292, :581-588, :706-715 */
static void ball_update(ball_t* g, const float* c, float R) {
	g->c[0] = c[0]; g->c[1] = c[1]; g->c[2] = c[2];
	memset(g->yVol, 0, sizeof(v3)); memset(g->ySurf, 0, sizeof(v3));
	g->R = R; g->r = 0.0f; g->rClamp = 1e-4f;
	if (!g->yukawa) return;
	g->muR = R*g->sqrtLambda;
	if (g->dim == 2) {
		g->K0muR = (float)bessk0(g->muR); g->I0muR = (float)bessi0(g->muR);
		g->K1muR = (float)bessk1(g->muR); g->I1muR = (float)bessi1(g->muR);
	} else {
		g->expmuR = expf(-g->muR);
		float exp2muR = g->expmuR*g->expmuR;
		float coshmuR = (1.0f + exp2muR)/(2.0f*g->expmuR);
		g->sinhmuR = (1.0f - exp2muR)/(2.0f*g->expmuR);
		g->K32muR = g->expmuR*(1.0f + 1.0f/g->muR);
		g->I32muR = coshmuR - g->sinhmuR/g->muR;
	}
}
/* evaluate(): :417-419, :504-506, :607-613, :734-740 */
static float ball_eval(const ball_t* g) {
	float R = g->R, r = g->r;
	if (!g->yukawa) {
		if (g->dim == 2) return (float)(logf(R/r)/(2.0f*M_PI));
		return (float)((1.0f/r - 1.0f/R)/(4.0f*M_PI));
	}
	float mur = r*g->sqrtLambda;
	if (g->dim == 2) {
		float K0mur = (float)bessk0(mur);
		float I0mur = (float)bessi0(mur);
		return (float)((K0mur - I0mur*g->K0muR/g->I0muR)/(2.0*M_PI));
	}
	float expmur = expf(-mur);
	float sinhmur = (1.0f - expmur*expmur)/(2.0f*expmur);
	return (float)((expmur - g->expmuR*sinhmur/g->sinhmuR)/(4.0f*M_PI*r));
}
/* poissonKernel(): :453-455, :540-542, :663-665, :795-797 */
static float ball_poisson(const ball_t* g) {
	if (!g->yukawa) return g->dim == 2 ? (float)(1.0f/(2.0f*M_PI)) : (float)(1.0f/(4.0f*M_PI));
	if (g->dim == 2) return (float)(1.0f/(2.0f*M_PI*g->I0muR));
	return (float)(g->muR/(4.0f*M_PI*g->sinhmuR));
}
/* norm(): :440-442, :527-529, :650-652, :782-784 */
static float ball_norm(const ball_t* g) {
	if (!g->yukawa) return g->dim == 2 ? g->R*g->R/4.0f : g->R*g->R/6.0f;
	if (g->dim == 2) return (float)((1.0f - 2.0*M_PI*ball_poisson(g))/g->lambda);
	return (float)((1.0f - 4.0*M_PI*ball_poisson(g))/g->lambda);
}
/* gradientNorm(): :428-431, :515-518, :634-641, :761-773 */
static float ball_grad_norm(const ball_t* g) {
	float R = g->R, r = g->r;
	if (!g->yukawa) {
		if (g->dim == 2) { float r2 = r*r; return (float)((1.0f/r2 - 1.0f/(R*R))/(2.0f*M_PI)); }
		float r3 = r*r*r; return (float)((1.0f/r3 - 1.0f/(R*R*R))/(4.0f*M_PI));
	}
	float mur = r*g->sqrtLambda;
	if (g->dim == 2) {
		float K1mur = (float)bessk1(mur);
		float I1mur = (float)bessi1(mur);
		float Qr = g->sqrtLambda*(K1mur - I1mur*g->K1muR/g->I1muR);
		return (float)(Qr/(2.0f*M_PI*r));
	}
	float r2 = r*r;
	float expmur = expf(-mur);
	float exp2mur = expmur*expmur;
	float coshmur = (1.0f + exp2mur)/(2.0f*expmur);
	float sinhmur = (1.0f - exp2mur)/(2.0f*expmur);
	float K32mur = expmur*(1.0f + 1.0f/mur);
	float I32mur = coshmur - sinhmur/mur;
	float Qr = g->sqrtLambda*(K32mur - I32mur*g->K32muR/g->I32muR);
	return (float)(Qr/(4.0f*M_PI*r2));
}
/* gradient(): d*gradientNorm(), d = yVol - c */
static void ball_gradient(const ball_t* g, float* o) {
	float gn = ball_grad_norm(g);
	for (int k = 0; k < 3; k++) o[k] = (g->yVol[k] - g->c[k])*gn;
}
/* poissonKernelGradient(): :464-468, :551-555, :680-685, :816-821.  NOTE the divisor is a
 * double expression that Eigen narrows to float before the per-component division. */
static void ball_poisson_grad(const ball_t* g, float* o) {
	v3 d; sub3(g->ySurf, g->c, d);
	if (!g->yukawa) {
		if (g->dim == 2) { float den = (float)(2.0f*M_PI*g->R*g->R); for (int k = 0; k < 3; k++) o[k] = (2.0f*d[k])/den; }
		else { float den = (float)(4.0f*M_PI*g->R*g->R); for (int k = 0; k < 3; k++) o[k] = (3.0f*d[k])/den; }
		return;
	}
	if (g->dim == 2) {
		float QR = g->sqrtLambda/(g->R*g->I1muR);
		float den = (float)(2.0f*M_PI);
		for (int k = 0; k < 3; k++) o[k] = (d[k]*QR)/den;
	} else {
		float QR = g->lambda/g->I32muR;
		float den = (float)(4.0f*M_PI);
		for (int k = 0; k < 3; k++) o[k] = (d[k]*QR)/den;
	}
}
/* directionSampledPoissonKernel(y): :459-461, :546-548, :669-677, :801-813 */
static float ball_dir_poisson(const ball_t* g, const float* y) {
	if (!g->yukawa) return 1.0f;
	v3 d; sub3(y, g->c, d);
	float r = fmax_std(g->rClamp, norm3(d));
	float mur = r*g->sqrtLambda;
	if (g->dim == 2) {
		float K1mur = (float)bessk1(mur);
		float I1mur = (float)bessi1(mur);
		float Q = K1mur + I1mur*g->K0muR/g->I0muR;
		return mur*Q;
	}
	float expmur = expf(-mur);
	float exp2mur = expmur*expmur;
	float coshmur = (1.0f + exp2mur)/(2.0f*expmur);
	float sinhmur = (1.0f - exp2mur)/(2.0f*expmur);
	float K32mur = expmur*(1.0f + 1.0f/mur);
	float I32mur = coshmur - sinhmur/mur;
	float Q = K32mur + I32mur*g->expmuR/g->sinhmuR;
	return mur*Q;
}
/* evaluate(x, y): :422-425, :509-512, :616-631, :743-758 */
static float ball_eval_xy(const ball_t* g, const float* x, const float* y) {
	v3 yx, xc, yc; sub3(y, x, yx); sub3(x, g->c, xc); sub3(y, g->c, yc);
	float R = g->R;
	float r1 = fmax_std(g->rClamp, norm3(yx));
	if (!g->yukawa) {
		if (g->dim == 2) return (float)((logf(R*R - dot3(xc, yc)) - logf(R*r1))/(2.0f*M_PI));
		return (float)((1.0f/r1 - R/(R*R - dot3(xc, yc)))/(4.0f*M_PI));
	}
	float r2 = (R*R - dot3(xc, yc))/R;
	float mur1 = r1*g->sqrtLambda, mur2 = r2*g->sqrtLambda;
	if (g->dim == 2) {
		float K0mur1 = (float)bessk0(mur1), K0mur2 = (float)bessk0(mur2);
		float I0mur1 = (float)bessi0(mur1), I0mur2 = (float)bessi0(mur2);
		float Q1 = K0mur1 - I0mur1*g->K0muR/g->I0muR;
		float Q2 = K0mur2 - I0mur2*g->K0muR/g->I0muR;
		return (float)((Q1 - Q2)/(2.0f*M_PI));
	}
	float expmur1 = expf(-mur1), expmur2 = expf(-mur2);
	float sinhmur1 = (1.0f - expmur1*expmur1)/(2.0f*expmur1);
	float sinhmur2 = (1.0f - expmur2*expmur2)/(2.0f*expmur2);
	float Q1 = (expmur1 - g->expmuR*sinhmur1/g->sinhmuR)/r1;
	float Q2 = (expmur2 - g->expmuR*sinhmur2/g->sinhmuR)/r2;
	return (float)((Q1 - Q2)/(4.0f*M_PI));
}
static float ball_potential(const ball_t* g) {
	return g->dim == 2 ? (float)(2.0f*M_PI*ball_poisson(g)) : (float)(4.0f*M_PI*ball_poisson(g));
}
/* sampleVolume(dir, sampler, pdf): rejection sampler :362-383 with the bounds of :403-409,
 * :591-599, :718-726; 3D harmonic closed form :483-496 */
static void ball_sample_volume(ball_t* g, const float* dir, nmo_pcg32* s, float* pdf) {
	float R = g->R;
	if (!g->yukawa && g->dim == 3) {
		float u1 = nmo_pcg32_float(s);
		float u2 = nmo_pcg32_float(s);
		float phi = (float)(2.0f*M_PI*u2);
		g->r = (1.0f + sqrtf(1.0f - cbrtf(u1*u1))*cosf(phi))*R/2.0f;
		g->r = fmax_std(g->rClamp, g->r);
		if (g->r > R) g->r = R/2.0f;
		for (int k = 0; k < 3; k++) g->yVol[k] = g->c[k] + g->r*dir[k];
		*pdf = ball_eval(g)/ball_norm(g);
		return;
	}
	float bound;
	if (!g->yukawa) bound = 1.5f/R;
	else {
		float a = g->dim == 2 ? 2.2f : 2.0f, b = g->dim == 2 ? 0.6f : 0.5f;
		float lam = g->lambda, sl = g->sqrtLambda;
		bound = R <= lam ?
			fmax_std(fmax_std(a/R, a/lam), fmax_std(b*sqrtf(R), b*sl)) :
			fmax_std(fmin_std(a/R, a/lam), fmin_std(b*sqrtf(R), b*sl));
	}
	int iter = 0;
	do {
		float u = nmo_pcg32_float(s);
		g->r = nmo_pcg32_float(s)*R;
		*pdf = ball_eval(g)/ball_norm(g);
		float pdfRadius = *pdf/pdf_sphere(g->dim, g->r);
		iter++;
		if (u < pdfRadius/bound) break;
	} while (iter < 1000);
	g->r = fmax_std(g->rClamp, g->r);
	if (g->r > R) g->r = R/2.0f;
	for (int k = 0; k < 3; k++) g->yVol[k] = g->c[k] + g->r*dir[k];
}

/* probe with the layout of ref_greens_ball (ref_harness.cpp:107-127) */
void nmo_greens_ball(int dim, float lambda, const float* R, const float* r, int n, float* out) {
	for (int i = 0; i < n; i++) {
		ball_t g; ball_init(&g, dim, lambda > 0.0f, lambda);
		v3 c = {0, 0, 0}; ball_update(&g, c, R[i]);
		g.r = r[i];
		v3 ex = {1, 0, 0}, el = {0, 0, 0}; el[dim - 1] = 1.0f;
		for (int k = 0; k < 3; k++) { g.yVol[k] = c[k] + r[i]*ex[k]; g.ySurf[k] = c[k] + R[i]*el[k]; }
		float* o = out + (size_t)i*10;
		o[0] = ball_eval(&g); o[1] = ball_norm(&g); o[2] = ball_grad_norm(&g); o[3] = ball_poisson(&g);
		o[4] = ball_dir_poisson(&g, g.yVol);
		v3 pg; ball_poisson_grad(&g, pg); o[5] = pg[dim - 1];
		v3 x; for (int k = 0; k < 3; k++) x[k] = c[k] + 0.25f*R[i]*el[k];
		o[6] = ball_eval_xy(&g, x, g.yVol);
		o[7] = ball_potential(&g);
		v3 gr; ball_gradient(&g, gr); o[8] = gr[0];
		o[9] = 0.0f;
	}
}
void nmo_sample_volume(int dim, float lambda, const float* R, const uint64_t* seeds, int n,
					   float* r_out, float* pdf_out, int* draws_out) {
	for (int i = 0; i < n; i++) {
		nmo_pcg32 s; nmo_pcg32_seed(&s, seeds[i], 1);
		ball_t g; ball_init(&g, dim, lambda > 0.0f, lambda);
		v3 c = {0, 0, 0}, ex = {1, 0, 0}; ball_update(&g, c, R[i]);
		float pdf = 0.0f;
		ball_sample_volume(&g, ex, &s, &pdf);
		r_out[i] = g.r; pdf_out[i] = pdf;
		nmo_pcg32 t; nmo_pcg32_seed(&t, seeds[i], 1);
		int k = 0; while (t.state != s.state && k < 4096) { nmo_pcg32_uint(&t); k++; }
		draws_out[i] = k;
	}
}

/* ---- geometry: boxes, cones (deps/fcpw/include/fcpw/core/bounding_volumes.h) --------------- */
typedef struct { v3 lo, hi; } box_t;
static void box_empty(box_t* b) { for (int k = 0; k < 3; k++) { b->lo[k] = MAXF; b->hi[k] = MINF; } }
static void box_add_pt(box_t* b, const float* p) { /* expandToInclude(p) :45-49 */
	for (int k = 0; k < 3; k++) { b->lo[k] = fmin_std(b->lo[k], p[k] - EPS); b->hi[k] = fmax_std(b->hi[k], p[k] + EPS); }
}
static void box_add_box(box_t* b, const box_t* o) {
	for (int k = 0; k < 3; k++) { b->lo[k] = fmin_std(b->lo[k], o->lo[k]); b->hi[k] = fmax_std(b->hi[k], o->hi[k]); }
}
static float box_area(const box_t* b) { /* surfaceArea() :139-142 */
	v3 e; for (int k = 0; k < 3; k++) e[k] = fmax_std(b->hi[k] - b->lo[k], 1e-5f);
	float P = e[0]*(e[1]*e[2]);
	return 2.0f*(P/e[0] + (P/e[1] + P/e[2]));
}
static void box_sqdist(const box_t* b, const float* p, float* d2Min, float* d2Max) { /* :62-67 */
	v3 a, c;
	for (int k = 0; k < 3; k++) {
		float u = b->lo[k] - p[k], v = p[k] - b->hi[k];
		a[k] = fmax_std(fmax_std(u, v), 0.0f);
		c[k] = fmin_std(u, v);
	}
	*d2Min = dot3(a, a); *d2Max = dot3(c, c);
}
static int box_ray(const box_t* b, const float* o, const float* invD, float rtMax, float* tMin, float* tMax) { /* :99-114 */
	v3 tn, tf;
	for (int k = 0; k < 3; k++) {
		float t0 = (b->lo[k] - o[k])*invD[k], t1 = (b->hi[k] - o[k])*invD[k];
		tn[k] = fmin_std(t0, t1); tf[k] = fmax_std(t0, t1);
	}
	float tNearMax = fmax_std(0.0f, fmax_std(tn[0], fmax_std(tn[1], tn[2])));
	float tFarMin = fmin_std(rtMax, fmin_std(tf[0], fmin_std(tf[1], tf[2])));
	if (tNearMax > tFarMin) return 0;
	*tMin = tNearMax; *tMax = tFarMin;
	return 1;
}
static inline int in_range(float val, float low, float high) { return val >= low && val <= high; }
/* computeOrthonormalBasis + projectToPlane<3> :175-209 */
static float project_to_plane(const float* n, const float* e) {
	float sign = copysignf(1.0f, n[2]);
	const float a = -1.0f/(sign + n[2]);
	const float b = n[0]*n[1]*a;
	v3 b1 = {1.0f + sign*n[0]*n[0]*a, sign*b, -sign*n[0]};
	v3 b2 = {b, sign + n[1]*n[1]*a, -n[1]};
	v3 a1 = {fabsf(b1[0]), fabsf(b1[1]), fabsf(b1[2])}, a2 = {fabsf(b2[0]), fabsf(b2[1]), fabsf(b2[2])};
	float r1 = dot3(e, a1), r2 = dot3(e, a2);
	return sqrtf(r1*r1 + r2*r2);
}
/* BoundingCone::overlap :225-271 */
static int cone_overlap(const float* axis, float halfAngle, const float* o, const box_t* b, float distToBox) {
	if (halfAngle >= M_PI_2 || distToBox < EPS) return 1;
	v3 c, vca;
	for (int k = 0; k < 3; k++) c[k] = (b->lo[k] + b->hi[k])*0.5f;
	sub3(c, o, vca);
	float l = norm3(vca);
	for (int k = 0; k < 3; k++) vca[k] /= l;
	float dAxisAngle = acosf(fmax_std(-1.0f, fmin_std(1.0f, dot3(axis, vca))));
	if (in_range((float)M_PI_2, dAxisAngle - halfAngle, dAxisAngle + halfAngle)) return 1;
	v3 e; sub3(b->hi, c, e);
	float r2 = dot3(e, e);
	if (l*l > r2) {
		float r = sqrtf(r2);
		float vha = asinf(r/l);
		float sum = halfAngle + vha;
		return sum >= M_PI_2 ? 1 : in_range((float)M_PI_2, dAxisAngle - sum, dAxisAngle + sum);
	}
	v3 av = {fabsf(vca[0]), fabsf(vca[1]), fabsf(vca[2])};
	float d = dot3(e, av);
	float s = l - d;
	if (s <= 0.0f) return 1;
	d = project_to_plane(vca, e);
	float vha = atan2f(d, s);
	float sum = halfAngle + vha;
	return sum >= M_PI_2 ? 1 : in_range((float)M_PI_2, dAxisAngle - sum, dAxisAngle + sum);
}

/* ---- scene ---------------------------------------------------------------------------------- */
typedef struct {
	box_t box;
	v3 axis; float halfAngle;
	int refOffset, nRefs, secondChild;
	int silOffset, nSilRefs;
} node_t;

typedef struct { int idx[4]; int pIndex; } sil_t; /* SilhouetteVertex (3 used) / SilhouetteEdge (4) */

struct nmo_scene {
	int dim;
	int nV, nP;
	v3* pos;          /* vertex positions (z = 0 in 2D) */
	int* prim;        /* nP x dim vertex indices, BVH order */
	int* primIndex;   /* pIndex of each BVH-ordered primitive */
	v3* vNormal;      /* soup.vNormals */
	int nE; int* eIdx; v3* eNormal; /* 3D: soup.eIndices (by pIndex), soup.eNormals */
	sil_t* sil; int nSil;          /* all silhouette vertices (2D: one per vertex) / edges (3D) */
	int* silRef; int nSilRef;      /* silhouetteRefs */
	node_t* nodes; int nNodes;
	float bboxLo[3], bboxHi[3];    /* zombie::computeBoundingBox over DIM components */
	float* src; int n0, n1, n2;
	float absorption; int watertight, doubleSided;
};

static void prim_box(const nmo_scene* s, const int* pv, box_t* b) { /* line_segments.inl:12-21, triangles.inl:12-23 */
	box_empty(b); /* BoundingBox(pa) == empty expanded by pa */
	for (int k = 0; k < s->dim; k++) box_add_pt(b, s->pos[pv[k]]);
}
static void prim_centroid(const nmo_scene* s, const int* pv, float* c) { /* :23-29 / :25-32 */
	const float *pa = s->pos[pv[0]], *pb = s->pos[pv[1]];
	if (s->dim == 2) for (int k = 0; k < 3; k++) c[k] = (pa[k] + pb[k])*0.5f;
	else { const float* pc = s->pos[pv[2]]; for (int k = 0; k < 3; k++) c[k] = ((pa[k] + pb[k]) + pc[k])/3.0f; }
}
/* unnormalised / normalised face normal: line_segments.inl:49-58, triangles.inl:49-60 */
static void face_normal(const nmo_scene* s, const int* pv, int normalize, float* n) {
	const float *pa = s->pos[pv[0]], *pb = s->pos[pv[1]];
	if (s->dim == 2) { v3 d; sub3(pb, pa, d); n[0] = d[1]; n[1] = -d[0]; n[2] = 0.0f; }
	else { const float* pc = s->pos[pv[2]]; v3 v1, v2; sub3(pb, pa, v1); sub3(pc, pa, v2); cross3(v1, v2, n); }
	if (normalize) normalize3(n);
}

/* -- BVH build: aggregates/sbvh.inl:4-234 (OverlapSurfaceArea, leafSize 4, 8 buckets, no packing) -- */
#define LEAF_SIZE 4
#define NBUCKETS 8
#define SBVH_MAX_DEPTH 64
typedef struct {
	nmo_scene* s; box_t* rbox; v3* rcen; int cap;
} build_t;

static float split_cost(const box_t* L, const box_t* R, int nL, int nR) { /* :4-39 */
	box_t I;
	for (int k = 0; k < 3; k++) { I.lo[k] = fmax_std(L->lo[k], R->lo[k]); I.hi[k] = fmin_std(L->hi[k], R->hi[k]); }
	float cost = (nL/box_area(R) + nR/box_area(L))*fabsf(box_area(&I));
	int valid = I.hi[0] >= I.lo[0] && I.hi[1] >= I.lo[1] && I.hi[2] >= I.lo[2];
	if (!valid) cost *= -1;
	return cost;
}
static void swap_refs(build_t* B, int i, int j) {
	nmo_scene* s = B->s; int d = s->dim;
	for (int k = 0; k < d; k++) { int t = s->prim[i*d + k]; s->prim[i*d + k] = s->prim[j*d + k]; s->prim[j*d + k] = t; }
	{ int t = s->primIndex[i]; s->primIndex[i] = s->primIndex[j]; s->primIndex[j] = t; }
	{ box_t t = B->rbox[i]; B->rbox[i] = B->rbox[j]; B->rbox[j] = t; }
	for (int k = 0; k < 3; k++) { float t = B->rcen[i][k]; B->rcen[i][k] = B->rcen[j][k]; B->rcen[j][k] = t; }
}
static void build_rec(build_t* B, int parent, int start, int end, int depth) { /* :141-207 */
	nmo_scene* s = B->s;
	int cur = s->nNodes++;
	node_t* node = &s->nodes[cur];
	memset(node, 0, sizeof(*node));
	node->halfAngle = (float)M_PI;
	int nRefs = end - start;
	box_t bb, bc; box_empty(&bb); box_empty(&bc);
	for (int p = start; p < end; p++) { box_add_box(&bb, &B->rbox[p]); box_add_pt(&bc, B->rcen[p]); }
	node->box = bb;
	int leaf = nRefs <= LEAF_SIZE || depth == SBVH_MAX_DEPTH - 2;
	if (leaf) { node->refOffset = start; node->nRefs = nRefs; }
	else { node->secondChild = -1; node->nRefs = 0; }
	if (parent >= 0) {
		/* second visit of the parent fixes the right-child offset (:180-189) */
		if (s->nodes[parent].secondChild == -1) s->nodes[parent].secondChild = -2;
		else if (s->nodes[parent].secondChild == -2) s->nodes[parent].secondChild = cur - parent;
	}
	if (leaf) return;

	/* computeObjectSplit :41-112 */
	float splitCost = MAXF; int splitDim = -1; float splitCoord = 0.0f;
	v3 extent; sub3(bb.hi, bb.lo, extent);
	for (int dim = 0; dim < 3; dim++) {
		if (extent[dim] < 1e-6f) continue;
		box_t bk[NBUCKETS], rb[NBUCKETS]; int cnt[NBUCKETS], rcnt[NBUCKETS];
		float bucketWidth = extent[dim]/NBUCKETS;
		for (int b = 0; b < NBUCKETS; b++) { box_empty(&bk[b]); cnt[b] = 0; rcnt[b] = 0; box_empty(&rb[b]); }
		for (int p = start; p < end; p++) {
			int bi = (int)((B->rcen[p][dim] - bb.lo[dim])/bucketWidth);
			bi = bi < 0 ? 0 : (bi > NBUCKETS - 1 ? NBUCKETS - 1 : bi);
			box_add_box(&bk[bi], &B->rbox[p]); cnt[bi] += 1;
		}
		box_t right; box_empty(&right);
		for (int b = NBUCKETS - 1; b > 0; b--) {
			box_add_box(&right, &bk[b]);
			rb[b] = right; rcnt[b] = cnt[b];
			if (b != NBUCKETS - 1) rcnt[b] += rcnt[b + 1];
		}
		box_t left; box_empty(&left); int nL = 0;
		for (int b = 1; b < NBUCKETS; b++) {
			box_add_box(&left, &bk[b - 1]); nL += cnt[b - 1];
			if (nL > 0 && rcnt[b] > 0) {
				float cost = split_cost(&left, &rb[b], nL, rcnt[b]);
				if (cost < splitCost) { splitCost = cost; splitDim = dim; splitCoord = bb.lo[dim] + b*bucketWidth; }
			}
		}
	}
	if (splitDim == -1) { /* fallback :105-109 (maxDimension of the centroid box) */
		v3 e; sub3(bc.hi, bc.lo, e);
		splitDim = 0; if (e[1] > e[splitDim]) splitDim = 1; if (e[2] > e[splitDim]) splitDim = 2;
		splitCoord = (bc.lo[splitDim] + bc.hi[splitDim])*0.5f;
	}
	/* performObjectSplit :114-139 */
	int mid = start;
	for (int i = start; i < end; i++) {
		if (B->rcen[i][splitDim] < splitCoord) { swap_refs(B, i, mid); mid++; }
	}
	if (mid == start || mid == end) mid = start + (end - start)/2;
	build_rec(B, cur, start, mid, depth + 1);
	build_rec(B, cur, mid, end, depth + 1);
}

/* -- silhouettes: fcpw.inl:224-291 (computeSilhouettes), sbvh.inl:236-443 (assign + cones) ----- */
static int sil_has_face(const nmo_scene* s, const sil_t* v, int f) {
	if (s->dim == 2) return f == 0 ? v->idx[2] != -1 : v->idx[0] != -1;  /* vertex_silhouettes.inl:31-34 */
	return f == 0 ? v->idx[3] != -1 : v->idx[0] != -1;                  /* edge_silhouettes.inl:40-43 */
}
static void sil_face_normal(const nmo_scene* s, const sil_t* v, int f, int normalize, float* n) {
	if (s->dim == 2) { /* vertex_silhouettes.inl:36-46 */
		int i = f == 0 ? 1 : 0;
		const float *pa = s->pos[v->idx[i]], *pb = s->pos[v->idx[i + 1]];
		v3 d; sub3(pb, pa, d); n[0] = d[1]; n[1] = -d[0]; n[2] = 0.0f;
	} else { /* edge_silhouettes.inl:45-68 */
		int i, j, k;
		if (f == 0) { i = 3; j = 1; k = 2; } else { i = 0; j = 2; k = 1; }
		const float *pa = s->pos[v->idx[j]], *pb = s->pos[v->idx[k]], *pc = s->pos[v->idx[i]];
		v3 v1, v2; sub3(pb, pa, v1); sub3(pc, pa, v2); cross3(v1, v2, n);
	}
	if (normalize) normalize3(n);
}
/* SilhouetteVertex::normal() / SilhouetteEdge::normal() with soup normals present */
static const float* sil_normal(const nmo_scene* s, const sil_t* v) {
	return s->dim == 2 ? s->vNormal[v->idx[1]] : s->eNormal[v->pIndex];
}

static void cones_rec(nmo_scene* s, const v3* refN, const v3* refFN, int start, int end) { /* sbvh.inl:236-301 */
	node_t* node = &s->nodes[start];
	v3 axis = {0, 0, 0}; float halfAngle = (float)M_PI;
	int any = 0, twoFaces = 1;
	for (int i = start; i < end; i++) {
		node_t* ch = &s->nodes[i];
		for (int j = 0; j < ch->nSilRefs; j++) {
			int ri = ch->silOffset + j;
			const sil_t* sv = &s->sil[s->silRef[ri]];
			for (int k = 0; k < 3; k++) axis[k] += refN[ri][k];
			twoFaces = twoFaces && sil_has_face(s, sv, 0) && sil_has_face(s, sv, 1);
			any = 1;
		}
	}
	if (!any) node->halfAngle = (float)-M_PI;
	else if (!twoFaces) node->halfAngle = (float)M_PI;
	else {
		float an = norm3(axis);
		if (an > EPS) {
			for (int k = 0; k < 3; k++) axis[k] /= an;
			halfAngle = 0.0f;
			for (int i = start; i < end; i++) {
				node_t* ch = &s->nodes[i];
				for (int j = 0; j < ch->nSilRefs; j++) {
					int ri = ch->silOffset + j;
					for (int k = 0; k < 2; k++) {
						float angle = acosf(fmax_std(-1.0f, fmin_std(1.0f, dot3(axis, refFN[2*ri + k]))));
						halfAngle = fmax_std(halfAngle, angle);
					}
				}
			}
			memcpy(node->axis, axis, sizeof(v3)); node->halfAngle = halfAngle;
		}
	}
	if (node->nRefs == 0) {
		cones_rec(s, refN, refFN, start + 1, start + node->secondChild);
		cones_rec(s, refN, refFN, start + node->secondChild, end);
	}
}

static void build_scene_geometry(nmo_scene* s) {
	int d = s->dim, nP = s->nP, nV = s->nV;
	/* vertex / edge normals: fcpw.inl:300-354 (computeWeighted = false) */
	s->vNormal = (v3*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(v3));
	if (d == 3) {
		/* assignEdgeIndices fcpw.inl:200-221: edges numbered by first appearance, key = sorted pair */
		s->eIdx = (int*)malloc(sizeof(int)*3*(size_t)nP);
		int* ekey = (int*)malloc(sizeof(int)*2*3*(size_t)nP);
		int E = 0;
		for (int i = 0; i < nP; i++) for (int j = 0; j < 3; j++) {
			int I = s->prim[3*i + j], J = s->prim[3*i + (j + 1)%3];
			if (I > J) { int t = I; I = J; J = t; }
			int f = -1;
			for (int e = 0; e < E; e++) if (ekey[2*e] == I && ekey[2*e + 1] == J) { f = e; break; }
			if (f < 0) { ekey[2*E] = I; ekey[2*E + 1] = J; f = E++; }
			s->eIdx[3*i + j] = f;
		}
		free(ekey);
		s->nE = E;
		s->eNormal = (v3*)calloc((size_t)(E > 0 ? E : 1), sizeof(v3));
	}
	for (int i = 0; i < nP; i++) {
		v3 n; face_normal(s, &s->prim[i*d], 1, n);
		if (d == 2) {
			for (int j = 0; j < 2; j++) for (int k = 0; k < 3; k++) s->vNormal[s->prim[2*i + j]][k] += 1.0f*n[k];
		} else {
			v3 un; face_normal(s, &s->prim[i*d], 0, un);
			float area = 0.5f*norm3(un);
			for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) {
				s->vNormal[s->prim[3*i + j]][k] += 1.0f*n[k];
				s->eNormal[s->eIdx[3*i + j]][k] += area*n[k];
			}
		}
	}
	for (int i = 0; i < nV; i++) normalize3(s->vNormal[i]);
	for (int i = 0; i < s->nE; i++) normalize3(s->eNormal[i]);

	/* silhouette edges are laid out BEFORE the build, in original triangle order (fcpw.inl:224-291) */
	if (d == 3) {
		s->nSil = s->nE;
		s->sil = (sil_t*)malloc(sizeof(sil_t)*(size_t)(s->nE > 0 ? s->nE : 1));
		for (int e = 0; e < s->nE; e++) { for (int k = 0; k < 4; k++) s->sil[e].idx[k] = -1; s->sil[e].pIndex = -1; }
		for (int i = 0; i < nP; i++) for (int j = 0; j < 3; j++) {
			int I = j - 1 < 0 ? 2 : j - 1, J = j, K = j + 1 > 2 ? 0 : j + 1;
			int e = s->eIdx[3*i + j];
			int orient = 1;
			if (s->prim[3*i + J] > s->prim[3*i + K]) { int t = J; J = K; K = t; orient = -1; }
			sil_t* se = &s->sil[e];
			se->idx[orient == 1 ? 0 : 3] = s->prim[3*i + I];
			se->idx[1] = s->prim[3*i + J];
			se->idx[2] = s->prim[3*i + K];
			se->pIndex = e;
		}
	}

	/* BVH */
	s->nodes = (node_t*)malloc(sizeof(node_t)*(size_t)(2*nP + 1));
	s->nNodes = 0;
	s->primIndex = (int*)malloc(sizeof(int)*(size_t)nP);
	for (int i = 0; i < nP; i++) s->primIndex[i] = i;
	build_t B; B.s = s;
	B.rbox = (box_t*)malloc(sizeof(box_t)*(size_t)nP); B.rcen = (v3*)malloc(sizeof(v3)*(size_t)nP);
	for (int i = 0; i < nP; i++) { prim_box(s, &s->prim[i*d], &B.rbox[i]); prim_centroid(s, &s->prim[i*d], B.rcen[i]); }
	build_rec(&B, -1, 0, nP, 0);
	free(B.rbox); free(B.rcen);
	/* eIdx was indexed by pIndex; keep a BVH-order copy accessor via primIndex */

	/* 2D: sortSoupPositions<3,true,LineSegment,SilhouetteVertex> (fcpw.inl:374-410, 469-490):
	 * vertices are renumbered by first appearance in BVH leaf order, then the silhouette
	 * vertices are wired a SECOND time under the new numbering WITHOUT clearing the first
	 * wiring (done by computeSilhouettes before the build, fcpw.inl:243-262, under the
	 * original numbering).  For open polylines an end vertex therefore keeps a stale
	 * neighbour index from whichever vertex used to own its slot; this is reference
	 * behaviour (it decides which end points count as silhouettes) and is restated as is.
	 * 3D needs no renumbering: SilhouetteEdge indices are remapped consistently (:504-520). */
	if (d == 2) {
		s->nSil = nV;
		s->sil = (sil_t*)malloc(sizeof(sil_t)*(size_t)(nV > 0 ? nV : 1));
		for (int v = 0; v < nV; v++) { for (int k = 0; k < 4; k++) s->sil[v].idx[k] = -1; s->sil[v].pIndex = -1; }
		/* first wiring: original segment order == pIndex order, original vertex numbering */
		int* inv = (int*)malloc(sizeof(int)*(size_t)nP);
		for (int i = 0; i < nP; i++) inv[s->primIndex[i]] = i;
		for (int q = 0; q < nP; q++) {
			int i = inv[q];
			int a = s->prim[2*i], b = s->prim[2*i + 1];
			s->sil[a].idx[1] = a; s->sil[a].idx[2] = b; s->sil[a].pIndex = a;
			s->sil[b].idx[0] = a; s->sil[b].idx[1] = b; s->sil[b].pIndex = b;
		}
		free(inv);
		/* renumber */
		int* map = (int*)malloc(sizeof(int)*(size_t)(nV > 0 ? nV : 1));
		v3* npos = (v3*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(v3));
		v3* nnrm = (v3*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(v3));
		for (int v = 0; v < nV; v++) map[v] = -1;
		int nv = 0;
		for (int i = 0; i < s->nNodes; i++) {
			node_t* node = &s->nodes[i];
			for (int j = 0; j < node->nRefs; j++) {
				int ri = node->refOffset + j;
				for (int k = 0; k < 2; k++) {
					int vi = s->prim[2*ri + k];
					if (map[vi] == -1) {
						memcpy(npos[nv], s->pos[vi], sizeof(v3)); memcpy(nnrm[nv], s->vNormal[vi], sizeof(v3));
						map[vi] = nv++;
					}
				}
			}
		}
		for (int i = 0; i < 2*nP; i++) s->prim[i] = map[s->prim[i]];
		free(s->pos); free(s->vNormal); free(map);
		s->pos = npos; s->vNormal = nnrm;
		/* second wiring: BVH order, new numbering */
		for (int i = 0; i < nP; i++) {
			int a = s->prim[2*i], b = s->prim[2*i + 1];
			s->sil[a].idx[1] = a; s->sil[a].idx[2] = b; s->sil[a].pIndex = a;
			s->sil[b].idx[0] = a; s->sil[b].idx[1] = b; s->sil[b].pIndex = b;
		}
	}

	/* assignSilhouettesToNodes sbvh.inl:314-443 with ignoreCandidateSilhouette (demo/scene.h:84-90) */
	int capRefs = (d == 2 ? 2 : 3)*nP + 1;
	s->silRef = (int*)malloc(sizeof(int)*(size_t)capRefs);
	v3* refN = (v3*)malloc(sizeof(v3)*(size_t)capRefs);
	v3* refFN = (v3*)malloc(sizeof(v3)*2*(size_t)capRefs);
	int* seen = (int*)malloc(sizeof(int)*(size_t)(s->nSil > 0 ? s->nSil : 1));
	s->nSilRef = 0;
	for (int i = 0; i < s->nNodes; i++) {
		node_t* node = &s->nodes[i];
		int start = s->nSilRef;
		for (int q = 0; q < s->nSil; q++) seen[q] = 0;
		for (int j = 0; j < node->nRefs; j++) {
			int ri = node->refOffset + j;
			for (int k = 0; k < d; k++) {
				int si = d == 2 ? s->prim[2*ri + k] : s->eIdx[3*s->primIndex[ri] + k];
				if (seen[si]) continue;
				seen[si] = 1;
				const sil_t* sv = &s->sil[si];
				v3 n = {0, 0, 0}, n0 = {0, 0, 0}, n1 = {0, 0, 0};
				int two = sil_has_face(s, sv, 0) && sil_has_face(s, sv, 1);
				int ignore = 0;
				if (two) {
					memcpy(n, sil_normal(s, sv), sizeof(v3));
					sil_face_normal(s, sv, 0, 1, n0);
					sil_face_normal(s, sv, 1, 1, n1);
					float angle;
					if (d == 2) angle = n0[0]*n1[1] - n1[0]*n0[1];
					else {
						v3 ed, cr; sub3(s->pos[sv->idx[2]], s->pos[sv->idx[1]], ed); normalize3(ed);
						cross3(n0, n1, cr);
						angle = atan2f(dot3(ed, cr), dot3(n0, n1));
					}
					ignore = s->doubleSided ? 0 : angle < 1e-3f;
				}
				if (!ignore) {
					int r = s->nSilRef++;
					s->silRef[r] = si;
					memcpy(refN[r], n, sizeof(v3)); memcpy(refFN[2*r], n0, sizeof(v3)); memcpy(refFN[2*r + 1], n1, sizeof(v3));
				}
			}
		}
		node->silOffset = start; node->nSilRefs = s->nSilRef - start;
	}
	if (s->nNodes > 0) cones_rec(s, refN, refFN, 0, s->nNodes);
	free(refN); free(refFN); free(seen);
}

nmo_scene* nmo_scene_create(int dim, const float* verts, int nV, const int* prims, int nP,
							const float* src, int n0, int n1, int n2,
							float absorption, int watertight, int doubleSided) {
	nmo_scene* s = (nmo_scene*)calloc(1, sizeof(nmo_scene));
	s->dim = dim; s->nV = nV; s->nP = nP;
	s->absorption = absorption; s->watertight = watertight; s->doubleSided = doubleSided;
	s->pos = (v3*)calloc((size_t)(nV > 0 ? nV : 1), sizeof(v3));
	for (int i = 0; i < nV; i++) for (int k = 0; k < dim; k++) s->pos[i][k] = verts[(size_t)i*dim + k];
	s->prim = (int*)malloc(sizeof(int)*(size_t)(nP > 0 ? nP*dim : 1));
	memcpy(s->prim, prims, sizeof(int)*(size_t)nP*dim);
	/* zombie::computeBoundingBox over DIM components (fcpw_scene_loader.h:75-93) */
	for (int k = 0; k < 3; k++) { s->bboxLo[k] = MAXF; s->bboxHi[k] = MINF; }
	for (int i = 0; i < nV; i++) for (int k = 0; k < dim; k++) {
		float p = s->pos[i][k]*1.0f;
		s->bboxLo[k] = fmin_std(s->bboxLo[k], p - EPS); s->bboxHi[k] = fmax_std(s->bboxHi[k], p + EPS);
	}
	s->n0 = n0; s->n1 = n1; s->n2 = dim == 3 ? n2 : 1;
	size_t ns = (size_t)n0*n1*(dim == 3 ? n2 : 1);
	s->src = (float*)malloc(sizeof(float)*ns);
	memcpy(s->src, src, sizeof(float)*ns);
	if (nP > 0) build_scene_geometry(s);
	return s;
}
void nmo_scene_destroy(nmo_scene* s) {
	if (!s) return;
	free(s->pos); free(s->prim); free(s->primIndex); free(s->vNormal); free(s->eIdx); free(s->eNormal);
	free(s->sil); free(s->silRef); free(s->nodes); free(s->src); free(s);
}
void nmo_scene_bbox(const nmo_scene* s, float* out) {
	for (int k = 0; k < s->dim; k++) { out[k] = s->bboxLo[k]; out[s->dim + k] = s->bboxHi[k]; }
}
int nmo_scene_num_nodes(const nmo_scene* s) { return s->nNodes; }
void nmo_scene_nodes(const nmo_scene* s, float* out) {
	for (int i = 0; i < s->nNodes; i++) {
		const node_t* n = &s->nodes[i]; float* o = out + (size_t)i*16;
		for (int k = 0; k < 3; k++) { o[k] = n->box.lo[k]; o[3 + k] = n->box.hi[k]; o[6 + k] = n->axis[k]; }
		o[9] = n->halfAngle; o[10] = (float)n->refOffset; o[11] = (float)n->silOffset;
		o[12] = (float)n->nRefs; o[13] = (float)n->nSilRefs; o[14] = (float)n->secondChild; o[15] = 0;
	}
}

/* ---- primitive queries ------------------------------------------------------------------------ */
/* findClosestPointLineSegment line_segments.inl:184-209 */
static float closest_on_segment(const float* pa, const float* pb, const float* x, float* pt, float* t) {
	v3 u, v, d; sub3(pb, pa, u); sub3(x, pa, v);
	float c1 = dot3(u, v);
	if (c1 <= 0.0f) { memcpy(pt, pa, sizeof(v3)); *t = 0.0f; sub3(x, pt, d); return norm3(d); }
	float c2 = dot3(u, u);
	if (c2 <= c1) { memcpy(pt, pb, sizeof(v3)); *t = 1.0f; sub3(x, pt, d); return norm3(d); }
	*t = c1/c2;
	for (int k = 0; k < 3; k++) pt[k] = pa[k] + u[k]*(*t);
	sub3(x, pt, d); return norm3(d);
}
/* findClosestPointTriangle triangles.inl:258-341 */
static float closest_on_triangle(const float* pa, const float* pb, const float* pc, const float* x, float* pt, float* t) {
	v3 ab, ac, ax, d; sub3(pb, pa, ab); sub3(pc, pa, ac); sub3(x, pa, ax);
	float d1 = dot3(ab, ax), d2 = dot3(ac, ax);
	if (d1 <= 0.0f && d2 <= 0.0f) { t[0] = 1.0f; t[1] = 0.0f; memcpy(pt, pa, sizeof(v3)); sub3(x, pt, d); return norm3(d); }
	v3 bx; sub3(x, pb, bx);
	float d3 = dot3(ab, bx), d4 = dot3(ac, bx);
	if (d3 >= 0.0f && d4 <= d3) { t[0] = 0.0f; t[1] = 1.0f; memcpy(pt, pb, sizeof(v3)); sub3(x, pt, d); return norm3(d); }
	v3 cx; sub3(x, pc, cx);
	float d5 = dot3(ab, cx), d6 = dot3(ac, cx);
	if (d6 >= 0.0f && d5 <= d6) { t[0] = 0.0f; t[1] = 0.0f; memcpy(pt, pc, sizeof(v3)); sub3(x, pt, d); return norm3(d); }
	float vc = d1*d4 - d3*d2;
	if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
		float v = d1/(d1 - d3);
		t[0] = 1.0f - v; t[1] = v;
		for (int k = 0; k < 3; k++) pt[k] = pa[k] + ab[k]*v;
		sub3(x, pt, d); return norm3(d);
	}
	float vb = d5*d2 - d1*d6;
	if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
		float w = d2/(d2 - d6);
		t[0] = 1.0f - w; t[1] = 0.0f;
		for (int k = 0; k < 3; k++) pt[k] = pa[k] + ac[k]*w;
		sub3(x, pt, d); return norm3(d);
	}
	float va = d3*d6 - d5*d4;
	if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
		float w = (d4 - d3)/((d4 - d3) + (d5 - d6));
		t[0] = 0.0f; t[1] = 1.0f - w;
		for (int k = 0; k < 3; k++) pt[k] = pb[k] + (pc[k] - pb[k])*w;
		sub3(x, pt, d); return norm3(d);
	}
	float denom = 1.0f/(va + vb + vc);
	float v = vb*denom, w = vc*denom;
	t[0] = 1.0f - v - w; t[1] = v;
	for (int k = 0; k < 3; k++) pt[k] = (pa[k] + ab[k]*v) + ac[k]*w;
	sub3(x, pt, d); return norm3(d);
}

typedef struct { float d; v3 p, n; float uv[2]; int ref; int prim; } hit_t;

/* normal(uv) with soup normals present: line_segments.inl:60-77, triangles.inl:62-90 */
static void prim_normal_uv(const nmo_scene* s, int ref, const float* uv, float* n) {
	const int* pv = &s->prim[ref*s->dim];
	if (s->dim == 2) {
		int vi = -1;
		if (uv[0] <= EPS) vi = 0; else if (uv[0] >= 1.0f - EPS) vi = 1;
		if (vi >= 0) memcpy(n, s->vNormal[pv[vi]], sizeof(v3)); else face_normal(s, pv, 1, n);
		return;
	}
	const float ome = 1.0f - EPS;
	int vi = -1;
	if (uv[0] >= ome && uv[1] <= EPS) vi = 0;
	else if (uv[0] <= EPS && uv[1] >= ome) vi = 1;
	else if (uv[0] <= EPS && uv[1] <= EPS) vi = 2;
	int ei = -1;
	if (vi == -1) {
		if (uv[0] <= EPS) ei = 1;
		else if (uv[1] <= EPS) ei = 2;
		else if (uv[0] + uv[1] >= ome) ei = 0;
	}
	if (vi >= 0) memcpy(n, s->vNormal[pv[vi]], sizeof(v3));
	else if (ei >= 0) memcpy(n, s->eNormal[s->eIdx[3*s->primIndex[ref] + ei]], sizeof(v3));
	else face_normal(s, pv, 1, n);
}

typedef struct { int node; float dist; } trav_t;

/* findClosestPoint: sbvh.inl:948-1074 + primitive findClosestPoint */
static int closest_point(const nmo_scene* s, const float* x, float r2, int recordNormal, hit_t* out) {
	if (s->nNodes == 0) return 0;
	trav_t stack[SBVH_MAX_DEPTH]; float bh[4];
	int found = 0;
	box_sqdist(&s->nodes[0].box, x, &bh[0], &bh[1]);
	if (!(bh[0] <= r2)) return 0;
	r2 = fmin_std(r2, bh[1]);
	stack[0].node = 0; stack[0].dist = bh[0];
	int sp = 0;
	out->d = MAXF; out->ref = -1; out->prim = -1;
	while (sp >= 0) {
		int ni = stack[sp].node; float cd = stack[sp].dist; sp--;
		if (cd > r2) continue;
		const node_t* node = &s->nodes[ni];
		if (node->nRefs > 0) {
			for (int p = 0; p < node->nRefs; p++) {
				int ri = node->refOffset + p;
				const int* pv = &s->prim[ri*s->dim];
				v3 pt; float uv[2] = {0, 0}; float d;
				if (s->dim == 2) { d = closest_on_segment(s->pos[pv[0]], s->pos[pv[1]], x, pt, &uv[0]); uv[1] = -1; }
				else d = closest_on_triangle(s->pos[pv[0]], s->pos[pv[1]], s->pos[pv[2]], x, pt, uv);
				if (d*d <= r2) {
					found = 1;
					r2 = fmin_std(r2, d*d);
					out->d = d; memcpy(out->p, pt, sizeof(v3)); out->uv[0] = uv[0]; out->uv[1] = uv[1];
					out->ref = ri; out->prim = s->primIndex[ri];
				}
			}
		} else {
			const node_t* n0 = &s->nodes[ni + 1]; const node_t* n1 = &s->nodes[ni + node->secondChild];
			box_sqdist(&n0->box, x, &bh[0], &bh[1]); int hit0 = bh[0] <= r2;
			r2 = fmin_std(r2, bh[1]);
			box_sqdist(&n1->box, x, &bh[2], &bh[3]); int hit1 = bh[2] <= r2;
			r2 = fmin_std(r2, bh[3]);
			if (hit0 && hit1) {
				int closer = ni + 1, other = ni + node->secondChild;
				if (bh[0] == 0.0f && bh[2] == 0.0f) {
					if (bh[3] < bh[1]) { int t = closer; closer = other; other = t; }
				} else if (bh[2] < bh[0]) {
					float t = bh[0]; bh[0] = bh[2]; bh[2] = t;
					int u = closer; closer = other; other = u;
				}
				sp++; stack[sp].node = other; stack[sp].dist = bh[2];
				sp++; stack[sp].node = closer; stack[sp].dist = bh[0];
			} else if (hit0) { sp++; stack[sp].node = ni + 1; stack[sp].dist = bh[0]; }
			else if (hit1) { sp++; stack[sp].node = ni + node->secondChild; stack[sp].dist = bh[2]; }
		}
	}
	if (found && recordNormal) prim_normal_uv(s, out->ref, out->uv, out->n);
	return found;
}

/* ray: sbvh.inl:538-683 + LineSegment::intersect line_segments.inl:146-182 / Triangle::intersect triangles.inl:219-256 */
static int prim_ray(const nmo_scene* s, int ri, const float* o, const float* dir, float tMax, int occl, hit_t* h) {
	const int* pv = &s->prim[ri*s->dim];
	const float *pa = s->pos[pv[0]], *pb = s->pos[pv[1]];
	if (s->dim == 2) {
		v3 u, v; sub3(pa, o, u); sub3(pb, pa, v);
		float dv = dir[0]*v[1] - dir[1]*v[0];
		if (fabsf(dv) <= EPS) return 0;
		float ud = u[0]*dir[1] - u[1]*dir[0];
		float sp = ud/dv;
		if (sp >= 0.0f && sp <= 1.0f) {
			float uv = u[0]*v[1] - u[1]*v[0];
			float t = uv/dv;
			if (t >= 0.0f && t <= tMax) {
				if (occl) return 1;
				h->d = t;
				for (int k = 0; k < 3; k++) h->p[k] = pa[k] + sp*v[k];
				h->n[0] = v[1]; h->n[1] = -v[0]; h->n[2] = 0.0f; normalize3(h->n);
				h->uv[0] = sp; h->uv[1] = -1; h->ref = ri; h->prim = s->primIndex[ri];
				return 1;
			}
		}
		return 0;
	}
	const float* pc = s->pos[pv[2]];
	v3 v1, v2, p, sv, q; sub3(pb, pa, v1); sub3(pc, pa, v2);
	cross3(dir, v2, p);
	float det = dot3(v1, p);
	if (fabsf(det) <= EPS) return 0;
	float invDet = 1.0f/det;
	sub3(o, pa, sv);
	float v = dot3(sv, p)*invDet;
	if (v < 0 || v > 1) return 0;
	cross3(sv, v1, q);
	float w = dot3(dir, q)*invDet;
	if (w < 0 || v + w > 1) return 0;
	float t = dot3(v2, q)*invDet;
	if (t >= 0.0f && t <= tMax) {
		if (occl) return 1;
		h->d = t;
		for (int k = 0; k < 3; k++) h->p[k] = (pa[k] + v1[k]*v) + v2[k]*w;
		cross3(v1, v2, h->n); normalize3(h->n);
		h->uv[0] = 1.0f - v - w; h->uv[1] = v; h->ref = ri; h->prim = s->primIndex[ri];
		return 1;
	}
	return 0;
}
static int ray_intersect(const nmo_scene* s, const float* o, const float* dir, float tMax, int occl, hit_t* out) {
	if (s->nNodes == 0) return 0;
	v3 invD = {1.0f/dir[0], 1.0f/dir[1], 1.0f/dir[2]};
	trav_t stack[SBVH_MAX_DEPTH]; float bh[4];
	int hits = 0;
	if (!box_ray(&s->nodes[0].box, o, invD, tMax, &bh[0], &bh[1])) return 0;
	stack[0].node = 0; stack[0].dist = bh[0];
	int sp = 0;
	while (sp >= 0) {
		int ni = stack[sp].node; float cd = stack[sp].dist; sp--;
		if (cd > tMax) continue;
		const node_t* node = &s->nodes[ni];
		if (node->nRefs > 0) {
			for (int p = 0; p < node->nRefs; p++) {
				hit_t h;
				if (prim_ray(s, node->refOffset + p, o, dir, tMax, occl, &h)) {
					if (occl) return 1;
					hits++;
					tMax = fmin_std(tMax, h.d);
					*out = h;
				}
			}
		} else {
			int c0 = ni + 1, c1 = ni + node->secondChild;
			int hit0 = box_ray(&s->nodes[c0].box, o, invD, tMax, &bh[0], &bh[1]);
			int hit1 = box_ray(&s->nodes[c1].box, o, invD, tMax, &bh[2], &bh[3]);
			if (hit0 && hit1) {
				int closer = c0, other = c1;
				if (bh[2] < bh[0]) {
					float t = bh[0]; bh[0] = bh[2]; bh[2] = t;
					t = bh[1]; bh[1] = bh[3]; bh[3] = t;
					closer = c1; other = c0;
				}
				sp++; stack[sp].node = other; stack[sp].dist = bh[2];
				sp++; stack[sp].node = closer; stack[sp].dist = bh[0];
			} else if (hit0) { sp++; stack[sp].node = c0; stack[sp].dist = bh[0]; }
			else if (hit1) { sp++; stack[sp].node = c1; stack[sp].dist = bh[2]; }
		}
	}
	return hits > 0;
}

/* isSilhouetteVertex vertex_silhouettes.inl:62-87 / isSilhouetteEdge edge_silhouettes.inl:83-110 */
static int is_silhouette(const nmo_scene* s, const sil_t* sv, const float* n0, const float* n1,
						 const float* viewDir, float d, int flip, float precision) {
	float sign = flip ? 1.0f : -1.0f;
	if (d <= precision) {
		float a;
		if (s->dim == 2) a = n0[0]*n1[1] - n1[0]*n0[1];
		else {
			v3 ed, cr; sub3(s->pos[sv->idx[2]], s->pos[sv->idx[1]], ed); normalize3(ed);
			cross3(n0, n1, cr);
			a = atan2f(dot3(ed, cr), dot3(n0, n1));
		}
		return sign*a > precision;
	}
	v3 vu = {viewDir[0]/d, viewDir[1]/d, viewDir[2]/d};
	float dot0 = dot3(vu, n0), dot1 = dot3(vu, n1);
	if (fabsf(dot0) <= precision) return sign*dot1 > precision;
	if (fabsf(dot1) <= precision) return sign*dot0 > precision;
	return dot0*dot1 < 0.0f;
}
/* SilhouetteVertex/Edge::findClosestSilhouettePoint vertex_silhouettes.inl:89-118, edge_silhouettes.inl:112-143 */
static int sil_closest(const nmo_scene* s, const sil_t* sv, const float* x, float r2, int flip,
					   float sqMinR, float precision, float* dOut) {
	if (sqMinR >= r2) return 0;
	v3 viewDir; float d;
	if (s->dim == 2) {
		sub3(x, s->pos[sv->idx[1]], viewDir);
		d = norm3(viewDir);
	} else {
		v3 pt; float t;
		d = closest_on_segment(s->pos[sv->idx[1]], s->pos[sv->idx[2]], x, pt, &t);
		sub3(x, pt, viewDir);
	}
	if (d*d > r2) return 0;
	int isSil = !sil_has_face(s, sv, 0) || !sil_has_face(s, sv, 1);
	if (!isSil) {
		v3 n0, n1; sil_face_normal(s, sv, 0, 1, n0); sil_face_normal(s, sv, 1, 1, n1);
		isSil = is_silhouette(s, sv, n0, n1, viewDir, d, flip, precision);
	}
	if (isSil && d*d <= r2) { *dOut = d; return 1; }
	return 0;
}
/* findClosestSilhouettePoint: sbvh.inl:1093-1255 */
static int closest_silhouette(const nmo_scene* s, const float* x, float r2, int flip, float sqMinR,
							  float precision, float* dOut) {
	if (s->nNodes == 0) return 0;
	if (sqMinR >= r2) return 0;
	trav_t stack[SBVH_MAX_DEPTH]; float bh[2], tmp;
	int found = 0, lastPrim = -1;
	box_sqdist(&s->nodes[0].box, x, &bh[0], &tmp);
	if (!(bh[0] <= r2)) return 0;
	stack[0].node = 0; stack[0].dist = bh[0];
	int sp = 0;
	while (sp >= 0) {
		int ni = stack[sp].node; float cd = stack[sp].dist; sp--;
		if (cd > r2) continue;
		const node_t* node = &s->nodes[ni];
		if (node->nRefs > 0) {
			for (int p = 0; p < node->nSilRefs; p++) {
				const sil_t* sv = &s->sil[s->silRef[node->silOffset + p]];
				if (sv->pIndex == lastPrim) continue;
				float d;
				if (sil_closest(s, sv, x, r2, flip, sqMinR, precision, &d)) {
					found = 1;
					r2 = fmin_std(r2, d*d);
					*dOut = d; lastPrim = sv->pIndex;
					if (sqMinR >= r2) break;
				}
			}
		} else {
			const node_t* n0 = &s->nodes[ni + 1]; const node_t* n1 = &s->nodes[ni + node->secondChild];
			int hit0 = 0, hit1 = 0;
			if (n0->halfAngle >= 0.0f) { box_sqdist(&n0->box, x, &bh[0], &tmp); hit0 = bh[0] <= r2 && cone_overlap(n0->axis, n0->halfAngle, x, &n0->box, bh[0]); }
			if (n1->halfAngle >= 0.0f) { box_sqdist(&n1->box, x, &bh[1], &tmp); hit1 = bh[1] <= r2 && cone_overlap(n1->axis, n1->halfAngle, x, &n1->box, bh[1]); }
			if (hit0 && hit1) {
				int closer = ni + 1, other = ni + node->secondChild;
				if (bh[1] < bh[0]) { float t = bh[0]; bh[0] = bh[1]; bh[1] = t; int u = closer; closer = other; other = u; }
				sp++; stack[sp].node = other; stack[sp].dist = bh[1];
				sp++; stack[sp].node = closer; stack[sp].dist = bh[0];
			} else if (hit0) { sp++; stack[sp].node = ni + 1; stack[sp].dist = bh[0]; }
			else if (hit1) { sp++; stack[sp].node = ni + node->secondChild; stack[sp].dist = bh[1]; }
		}
	}
	return found;
}

/* ---- zombie geometric queries (include/zombie/utils/fcpw_scene_loader.h:292-652) -------------- */
static float dist_dirichlet(const nmo_scene* s, const float* x) { /* :299-315, no Dirichlet geometry */
	float a = 0.0f;
	v3 c = {0, 0, 0};
	for (int k = 0; k < s->dim; k++) { float u = s->bboxLo[k] - x[k], v = x[k] - s->bboxHi[k]; c[k] = fmin_std(u, v); }
	a = s->dim == 2 ? c[0]*c[0] + c[1]*c[1] : dot3(c, c);
	return sqrtf(a);
}
static float dist_neumann(const nmo_scene* s, const float* x, int sgn) { /* :316-330 */
	if (s->nP == 0) return MAXF;
	hit_t h; memset(&h, 0, sizeof(h)); h.d = MAXF;
	closest_point(s, x, MAXF, sgn, &h);
	if (!sgn) return h.d;
	v3 xp; sub3(x, h.p, xp);
	return (dot3(xp, h.n) > 0.0f ? 1.0f : -1.0f)*h.d; /* Interaction::signedDistance interaction.h:32-34 */
}
static int inside_domain(const nmo_scene* s, const float* x) { /* :642-648 */
	if (!s->watertight) return 1;
	float d1 = dist_dirichlet(s, x);
	float d2 = dist_neumann(s, x, 1);
	return fabsf(d1) < fabsf(d2) ? d1 < 0.0f : d2 < 0.0f;
}
static int outside_bbox(const nmo_scene* s, const float* x) { /* :649-651 */
	for (int k = 0; k < s->dim; k++) if (!(x[k] >= s->bboxLo[k] && x[k] <= s->bboxHi[k])) return 1;
	return 0;
}
static inline float i2f(int a) { union { int a; float b; } u; u.a = a; return u.b; }
static inline int f2i(float a) { union { float a; int b; } u; u.a = a; return u.b; }
static void offset_point(int dim, const float* p, const float* n, float* o) { /* :252-290 */
	const float origin = 1.0f/32.0f, floatScale = 1.0f/65536.0f, intScale = 256.0f;
	o[2] = dim == 2 ? p[2] : 0.0f;
	for (int k = 0; k < dim; k++) {
		int nOff = (int)(n[k]*intScale);
		float pOff = i2f(f2i(p[k]) + (p[k] < 0 ? -nOff : nOff));
		o[k] = fabsf(p[k]) < origin ? p[k] + floatScale*n[k] : pOff;
	}
}
static float star_radius(const nmo_scene* s, const float* x, float minR, float maxR, float prec, int flipOrient) { /* :621-641 */
	if (minR > maxR) return maxR;
	if (s->nP > 0) {
		int flip = 1; if (flipOrient) flip = !flip;
		float r2 = maxR < MAXF ? maxR*maxR : MAXF;
		float d;
		if (closest_silhouette(s, x, r2, flip, minR*minR, prec, &d)) return fmax_std(d, minR);
	}
	return fmax_std(maxR, minR);
}
static int intersect_neumann(const nmo_scene* s, const float* org, const float* nrm, const float* dir,
							 float tMax, int onB, hit_t* h) { /* :458-484 */
	if (s->nP == 0) return 0;
	v3 o = {0, 0, 0}, d = {0, 0, 0};
	if (onB) { v3 nn = {-nrm[0], -nrm[1], -nrm[2]}; offset_point(s->dim, org, nn, o); if (s->dim == 2) o[2] = 0.0f; }
	else { for (int k = 0; k < s->dim; k++) o[k] = org[k]; }
	for (int k = 0; k < s->dim; k++) d[k] = dir[k];
	return ray_intersect(s, o, d, tMax, 0, h);
}
static int blocked(const nmo_scene* s, const float* xi, const float* xj, const float* ni, const float* nj, int offi, int offj) { /* :485-499 + primitive.h hasLineOfSight */
	if (s->nP == 0) return 0;
	v3 p1 = {0, 0, 0}, p2 = {0, 0, 0};
	if (offi) { v3 nn = {-ni[0], -ni[1], -ni[2]}; offset_point(s->dim, xi, nn, p1); if (s->dim == 2) p1[2] = 0.0f; } else for (int k = 0; k < s->dim; k++) p1[k] = xi[k];
	if (offj) { v3 nn = {-nj[0], -nj[1], -nj[2]}; offset_point(s->dim, xj, nn, p2); if (s->dim == 2) p2[2] = 0.0f; } else for (int k = 0; k < s->dim; k++) p2[k] = xj[k];
	v3 d; sub3(p2, p1, d);
	float dn = norm3(d);
	for (int k = 0; k < 3; k++) d[k] /= dn;
	hit_t h;
	return ray_intersect(s, p1, d, dn, 1, &h);
}
/* pde.source: demo/scene.h:194-198 + image.h:70-75; zombie3d demo/scene_3d.h:120-126 */
static float source(const nmo_scene* s, const float* x) {
	v3 uv = {0, 0, 0};
	for (int k = 0; k < s->dim; k++) uv[k] = (x[k] - s->bboxLo[k])/(s->bboxHi[k] - s->bboxLo[k]);
	if (s->dim == 2) {
		int h = s->n0, w = s->n1;
		int i = (int)(uv[1]*h); i = i < 0 ? 0 : (i > h - 1 ? h - 1 : i);
		int j = (int)(uv[0]*w); j = j < 0 ? 0 : (j > w - 1 ? w - 1 : j);
		return s->src[(size_t)i*w + j];
	}
	int i = (int)(uv[0]*s->n0); i = i < 0 ? 0 : (i > s->n0 - 1 ? s->n0 - 1 : i);
	int j = (int)(uv[1]*s->n1); j = j < 0 ? 0 : (j > s->n1 - 1 ? s->n1 - 1 : j);
	int k = (int)(uv[2]*s->n2); k = k < 0 ? 0 : (k > s->n2 - 1 ? s->n2 - 1 : k);
	return s->src[((size_t)i*s->n1 + j)*s->n2 + k];
}

/* ---- probes --------------------------------------------------------------------------------- */
#define LOADPT(dst, arr, i) do { (dst)[0] = (dst)[1] = (dst)[2] = 0.0f; for (int _k = 0; _k < s->dim; _k++) (dst)[_k] = (arr)[(size_t)(i)*s->dim + _k]; } while (0)
void nmo_dist_neumann(const nmo_scene* s, const float* pts, int n, int signed_, float* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = dist_neumann(s, x, signed_); }
}
void nmo_dist_dirichlet(const nmo_scene* s, const float* pts, int n, float* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = dist_dirichlet(s, x); }
}
void nmo_inside_domain(const nmo_scene* s, const float* pts, int n, int* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = inside_domain(s, x); }
}
void nmo_outside_bbox(const nmo_scene* s, const float* pts, int n, int* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = outside_bbox(s, x); }
}
void nmo_star_radius(const nmo_scene* s, const float* pts, int n, float minR, const float* maxR,
					 float prec, int flip, float* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = star_radius(s, x, minR, maxR[i], prec, flip); }
}
void nmo_intersect_neumann(const nmo_scene* s, const float* org, const float* nrm, const float* dir,
						   const float* tmax, const int* onb, int n, float* out) {
	const int W = 2 + 2*s->dim;
	for (int i = 0; i < n; i++) {
		v3 o, nn, d; LOADPT(o, org, i); LOADPT(nn, nrm, i); LOADPT(d, dir, i);
		hit_t h; memset(&h, 0, sizeof(h)); h.d = MAXF;
		int hit = intersect_neumann(s, o, nn, d, tmax[i], onb[i], &h);
		float* r = out + (size_t)i*W;
		r[0] = hit ? 1.0f : 0.0f; r[1] = h.d;
		for (int k = 0; k < s->dim; k++) { r[2 + k] = h.p[k]; r[2 + s->dim + k] = h.n[k]; }
	}
}
void nmo_blocked(const nmo_scene* s, const float* xi, const float* xj, const float* ni, const float* nj,
				 const int* offi, const int* offj, int n, int* out) {
	for (int i = 0; i < n; i++) {
		v3 a, b, na, nb; LOADPT(a, xi, i); LOADPT(b, xj, i); LOADPT(na, ni, i); LOADPT(nb, nj, i);
		out[i] = blocked(s, a, b, na, nb, offi[i], offj[i]);
	}
}
void nmo_offset_point(int dim, const float* p, const float* nrm, int n, float* out) {
	for (int i = 0; i < n; i++) {
		v3 a = {0, 0, 0}, b = {0, 0, 0}, o;
		for (int k = 0; k < dim; k++) { a[k] = p[(size_t)i*dim + k]; b[k] = nrm[(size_t)i*dim + k]; }
		offset_point(dim, a, b, o);
		for (int k = 0; k < dim; k++) out[(size_t)i*dim + k] = o[k];
	}
}
void nmo_source(const nmo_scene* s, const float* pts, int n, float* out) {
	for (int i = 0; i < n; i++) { v3 x; LOADPT(x, pts, i); out[i] = source(s, x); }
}

/* ---- the estimator (include/zombie/point_estimation/walk_on_stars.h) --------------------------- */
typedef struct { /* WalkState :880-913 */
	v3 pt, normal, prevDir, srcGradDir, bdyGradDir;
	float prevDist, throughput;
	int onNeumann;
	float terminal, totalNeumann, totalSource, firstSource;
	int walkLength;
} wstate_t;

typedef struct { /* SampleStatistics :744-877 */
	float solMean, solM2, gradMean[3], gradM2[3], totalFirstSource, totalDeriv;
	int nSol, nGrad, totalWalkLength;
} stats_t;

static void welford(float est, float* mean, float* M2, int N) { /* :863-868 */
	float delta = est - *mean;
	*mean += delta/N;
	float delta2 = est - *mean;
	*M2 += delta*delta2;
}

enum { REACHED_DIRICHLET = 0, RUSSIAN_ROULETTE = 1, EXCEEDED_LENGTH = 2, ESCAPED = 3 };

/* walk(): :135-329 (firstSphereRadius is always 0 on this path, :580) */
static int walk(const nmo_scene* s, const nmo_solver_opts* o, float dirichletDist, nmo_pcg32* rng,
				ball_t* g, wstate_t* st) {
	while (dirichletDist > o->epsilonShell) {
		float starRadius;
		/* solveDoubleSided normal flip :154-160 */
		int flipOrient = 0;
		if (s->doubleSided && st->onNeumann) {
			if (st->prevDist > 0.0f && dot3(st->prevDir, st->normal) < 0.0f) {
				for (int k = 0; k < 3; k++) st->normal[k] *= -1.0f;
				flipOrient = 1;
			}
		}
		if (o->stepsBeforeUsingMaximalSpheres <= st->walkLength) starRadius = dirichletDist;
		else {
			starRadius = star_radius(s, st->pt, o->minStarRadius, dirichletDist, o->silhouettePrecision, flipOrient);
			if (o->minStarRadius <= dirichletDist) starRadius = fmax_std(SHRINK*starRadius, o->minStarRadius);
		}
		ball_update(g, st->pt, starRadius);

		float u[2]; u[0] = nmo_pcg32_float(rng); if (s->dim == 3) u[1] = nmo_pcg32_float(rng);
		v3 dir; sphere_dir(s->dim, u, dir);
		if (st->onNeumann && dot3(st->normal, dir) > 0.0f) for (int k = 0; k < 3; k++) dir[k] *= -1.0f;

		hit_t h; memset(&h, 0, sizeof(h)); h.d = MAXF;
		int hit = intersect_neumann(s, st->pt, st->normal, dir, starRadius, st->onNeumann, &h);
		v3 ipt, inrm = {0, 0, 0}; float idist;
		if (hit) { memcpy(ipt, h.p, sizeof(v3)); memcpy(inrm, h.n, sizeof(v3)); idist = h.d; if (s->dim == 2) { ipt[2] = 0; inrm[2] = 0; } }
		else {
			v3 cp;
			if (st->onNeumann) { v3 nn = {-st->normal[0], -st->normal[1], -st->normal[2]}; offset_point(s->dim, st->pt, nn, cp); if (s->dim == 2) cp[2] = 0.0f; }
			else memcpy(cp, st->pt, sizeof(v3));
			for (int k = 0; k < 3; k++) ipt[k] = cp[k] + starRadius*dir[k];
			idist = starRadius;
		}
		if (!o->ignoreNeumann) { /* :212-260: h == 0, only the DIM draws are observable */
			for (int k = 0; k < s->dim; k++) (void)nmo_pcg32_float(rng);
		}
		if (!o->ignoreSource) { /* :262-276 */
			float pdf;
			ball_sample_volume(g, dir, rng, &pdf);
			if (g->r <= idist) {
				float sc = ball_norm(g)*source(s, g->yVol);
				st->totalSource += st->throughput*sc;
			}
		}
		if (!hit && outside_bbox(s, ipt)) return ESCAPED;

		st->prevDist = idist;
		memcpy(st->prevDir, dir, sizeof(v3));
		memcpy(st->pt, ipt, sizeof(v3));
		memcpy(st->normal, inrm, sizeof(v3));
		st->onNeumann = hit;

		st->throughput *= ball_dir_poisson(g, st->pt);
		if (st->throughput < o->russianRouletteThreshold) {
			float survival = st->throughput/o->russianRouletteThreshold;
			if (survival < nmo_pcg32_float(rng)) { st->throughput = 0.0f; return RUSSIAN_ROULETTE; }
			st->throughput = o->russianRouletteThreshold;
		}
		st->walkLength++;
		if (st->walkLength > o->maxWalkLength) return EXCEEDED_LENGTH;
		if (s->absorption > 0.0f && o->stepsBeforeApplyingTikhonov == st->walkLength) {
			ball_init(g, s->dim, 1, s->absorption); /* :319-321 */
		}
		dirichletDist = dist_dirichlet(s, st->pt);
	}
	return REACHED_DIRICHLET;
}

/* estimateSolutionAndGradient(): :466-617 */
static void estimate_point(const nmo_scene* s, const nmo_solver_opts* o, const float* pt, float dDist, float nDist,
						   nmo_pcg32* rng, stats_t* S, float* scratch) {
	memset(S, 0, sizeof(*S));
	int nWalks = o->nWalks, nAnti = 1;
	if (o->useGradientAntitheticVariates) { nWalks = nWalks/2 > 1 ? nWalks/2 : 1; nAnti = 2; }
	float boundaryDist = fmin_std(dDist, nDist);
	float firstR = SHRINK*boundaryDist;
	const int D = s->dim - 1;
	stratified(D, 2*nWalks, rng, scratch);
	const v3 dirForDeriv = {1.0f, 0.0f, 0.0f}; /* SampleEstimationData default :664-667 */

	for (int w = 0; w < nWalks; w++) {
		float boundaryPdf = 0, sourcePdf = 0;
		v3 boundaryPt = {0, 0, 0}, sourcePt = {0, 0, 0};
		uint32_t seed = nmo_pcg32_uint(rng); /* deterministic stand-in for the clock read :498 */
		float bcv = 0.0f, scv = 0.0f;
		if (o->useGradientControlVariates) {
			bcv = S->solMean;
			int N = S->nSol > 1 ? S->nSol : 1;
			scv = S->totalFirstSource/N;
		}
		for (int a = 0; a < nAnti; a++) {
			ball_t g; ball_init(&g, s->dim, s->absorption > 0.0f && o->stepsBeforeApplyingTikhonov == 0, s->absorption);
			wstate_t st; memset(&st, 0, sizeof(st));
			memcpy(st.pt, pt, sizeof(v3)); st.throughput = 1.0f;
			ball_update(&g, st.pt, firstR);
			if (!o->ignoreSource) {
				if (a == 0) {
					v3 sd; sphere_dir(s->dim, &scratch[D*(2*w + 0)], sd);
					ball_sample_volume(&g, sd, rng, &sourcePdf);
					memcpy(sourcePt, g.yVol, sizeof(v3));
				} else {
					v3 sd; sub3(sourcePt, st.pt, sd);
					for (int k = 0; k < 3; k++) g.yVol[k] = st.pt[k] - sd[k];
					g.r = norm3(sd);
				}
				float gn = ball_norm(&g);
				float sc = gn*source(s, g.yVol);
				st.totalSource += st.throughput*sc;
				st.firstSource = sc;
				v3 gr; ball_gradient(&g, gr);
				float den = sourcePdf*gn;
				for (int k = 0; k < 3; k++) st.srcGradDir[k] = gr[k]/den;
			}
			if (a == 0) {
				v3 bd;
				if (o->useCosineSamplingForDerivatives) { /* :550-554 */
					cosine_hemisphere(s->dim, &scratch[D*(2*w + 1)], bd);
					if (nmo_pcg32_float(rng) < 0.5f) bd[s->dim - 1] *= -1.0f;
					boundaryPdf = 0.5f*pdf_cosine_hemisphere(s->dim, fabsf(bd[s->dim - 1]));
					transform_coordinates(s->dim, dirForDeriv, bd);
				} else {
					sphere_dir(s->dim, &scratch[D*(2*w + 1)], bd);
					boundaryPdf = pdf_sphere(s->dim, 1.0f);
				}
				for (int k = 0; k < 3; k++) g.ySurf[k] = g.c[k] + g.R*bd[k];
				memcpy(boundaryPt, g.ySurf, sizeof(v3));
			} else {
				v3 bd; sub3(boundaryPt, st.pt, bd);
				for (int k = 0; k < 3; k++) g.ySurf[k] = st.pt[k] - bd[k];
			}
			st.prevDist = g.R;
			for (int k = 0; k < 3; k++) st.prevDir[k] = (g.ySurf[k] - st.pt[k])/g.R;
			memcpy(st.pt, g.ySurf, sizeof(v3));
			st.throughput *= ball_poisson(&g)/boundaryPdf;
			{
				v3 pg; ball_poisson_grad(&g, pg);
				float den = boundaryPdf*st.throughput;
				for (int k = 0; k < 3; k++) st.bdyGradDir[k] = pg[k]/den;
			}
			float dirichletDist = dist_dirichlet(s, st.pt);
			nmo_pcg32_seed(rng, seed, 1);
			int code = walk(s, o, dirichletDist, rng, &g, &st);
			if (code == REACHED_DIRICHLET || code == RUSSIAN_ROULETTE) {
				st.terminal = 0.0f; /* :331-351 with pde.dirichlet == 0 and initVal == 0 */
				float total = st.throughput*st.terminal + st.totalNeumann + st.totalSource;
				float bEst[3], sEst[3];
				float bContribution = total - st.firstSource;
				float deriv = 0.0f;
				for (int i = 0; i < s->dim; i++) {
					bEst[i] = (bContribution - bcv)*st.bdyGradDir[i];
					sEst[i] = (st.firstSource - scv)*st.srcGradDir[i];
					deriv += bEst[i]*dirForDeriv[i];
					deriv += sEst[i]*dirForDeriv[i];
				}
				S->nSol += 1; welford(total, &S->solMean, &S->solM2, S->nSol);
				S->totalFirstSource += st.firstSource;
				S->nGrad += 1;
				for (int i = 0; i < s->dim; i++) welford(bEst[i] + sEst[i], &S->gradMean[i], &S->gradM2[i], S->nGrad);
				S->totalDeriv += deriv;
				S->totalWalkLength += st.walkLength;
			}
		}
	}
}

typedef struct {
	const nmo_scene* s; const nmo_solver_opts* o; const float* pts; int n;
	uint64_t seed, index_offset; float* p; float* g; float* stats;
	volatile int* next; int maxPairs;
} job_t;

static void solve_point(job_t* J, int i, float* scratch) {
	const nmo_scene* s = J->s; const nmo_solver_opts* o = J->o;
	v3 x; LOADPT(x, J->pts, i);
	/* createSolutionGrid demo/grid.h:69-102 */
	float dDist = dist_dirichlet(s, x);
	float nDist = dist_neumann(s, x, 0);
	int inside = inside_domain(s, x);
	int active = inside || s->doubleSided; /* demo.cpp:152-158 */
	stats_t S; memset(&S, 0, sizeof(S));
	if (active) {
		nmo_pcg32 rng; nmo_pcg32_seed(&rng, nmo_point_seed(J->seed, J->index_offset + (uint64_t)i), 1);
		estimate_point(s, o, x, dDist, nDist, &rng, &S, scratch);
	}
	/* getSolution / getGradient demo/grid.h:155-179, 207-237 */
	int maskP = fabsf(nDist) < o->boundaryDistanceMask;
	int maskG = (!inside && !s->doubleSided) || maskP;
	J->p[i] = maskP ? 0.0f : S.solMean;
	for (int k = 0; k < s->dim; k++) J->g[(size_t)i*s->dim + k] = maskG ? 0.0f : S.gradMean[k];
	if (J->stats) {
		float* t = J->stats + (size_t)i*12;
		for (int k = 0; k < 12; k++) t[k] = 0.0f;
		t[11] = active ? 1.0f : 0.0f;
		if (active) {
			t[0] = S.solMean; t[1] = S.solM2/(S.nSol - 1 > 1 ? S.nSol - 1 : 1);
			for (int k = 0; k < s->dim; k++) { t[2 + k] = S.gradMean[k]; t[5 + k] = S.gradM2[k]/(S.nGrad - 1 > 1 ? S.nGrad - 1 : 1); }
			t[8] = S.totalFirstSource/(S.nSol > 1 ? S.nSol : 1);
			t[9] = (float)S.nSol;
			t[10] = (float)S.totalWalkLength/(S.nSol > 1 ? S.nSol : 1);
		}
	}
}
static void* worker(void* arg) {
	job_t* J = (job_t*)arg;
	float* scratch = (float*)malloc(sizeof(float)*2*(size_t)(2*J->maxPairs + 2));
	for (;;) {
		int b = __sync_fetch_and_add(J->next, 16);
		if (b >= J->n) break;
		int e = b + 16 < J->n ? b + 16 : J->n;
		for (int i = b; i < e; i++) solve_point(J, i, scratch);
	}
	free(scratch);
	return 0;
}
int nmo_wost(const nmo_scene* s, const nmo_solver_opts* o, const float* pts, int n,
			 uint64_t seed, uint64_t index_offset, int nthreads,
			 float* p_out, float* grad_out, float* stats) {
	volatile int next = 0;
	job_t J = {s, o, pts, n, seed, index_offset, p_out, grad_out, stats, &next, o->nWalks > 1 ? o->nWalks : 1};
	if (nthreads <= 1) { worker(&J); return 0; }
	pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t)*(size_t)nthreads);
	for (int t = 0; t < nthreads; t++) pthread_create(&th[t], 0, worker, &J);
	for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
	free(th);
	return 0;
}
