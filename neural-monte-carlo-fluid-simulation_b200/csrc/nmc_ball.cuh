// csrc/nmc_ball.cuh -- Green's functions of a ball (harmonic and screened/"Yukawa", 2D and 3D),
// their radial samplers, Poisson kernels and gradients.
//
// BallExact<DIM> replays the reference's evaluation order and float/double narrowing points
// (include/zombie/core/distributions.h:273-832, deps/bessel/bessel.hpp:373-556) and its rejection
// sampler (:362-383); it is what the deterministic mode uses.
// BallFast<DIM> (below) is the fp32 formulation of the same functions used by the default mode:
// exponentially scaled Bessel functions (no overflow for large R*sqrt(lambda)) and an inverse-CDF
// radial sampler instead of the rejection loop.
#pragma once
#include "nmc_math.cuh"

namespace nmc {

// ---- bessel.hpp polynomials, double ---------------------------------------------------------------
NMC_HD double bessi0(double x) {
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
NMC_HD double bessi1(double x) {
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
NMC_HD double bessk0(double x) {
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
NMC_HD double bessk1(double x) {
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

// ---- BallExact -----------------------------------------------------------------------------------
template <int DIM>
struct BallExact {
	typedef ExactMath M;
	bool yukawa;
	float lambda, sqrtLambda;
	V3 c, yVol, ySurf;
	float R, r, rClamp;
	float muR, a0, a1, a2, a3; // 2D: K0muR I0muR K1muR I1muR ; 3D: expmuR sinhmuR K32muR I32muR

	NMC_HD void init(bool yukawa_, float lambda_) {
		yukawa = yukawa_; lambda = lambda_; sqrtLambda = sqrtf(lambda_);
		c = yVol = ySurf = mk(0, 0, 0); R = 0.0f; r = 0.0f; rClamp = 1e-4f;
		muR = a0 = a1 = a2 = a3 = 0.0f;
	}
	// updateBall (:285-292, :581-588, :706-715)
	NMC_HD void update(V3 c_, float R_) {
		c = c_; yVol = ySurf = mk(0, 0, 0); R = R_; r = 0.0f; rClamp = 1e-4f;
		if (!yukawa) return;
		muR = R*sqrtLambda;
		if (DIM == 2) {
			a0 = (float)bessk0(muR); a1 = (float)bessi0(muR); a2 = (float)bessk1(muR); a3 = (float)bessi1(muR);
		} else {
			a0 = M::exp_(-muR);
			float exp2muR = a0*a0;
			float coshmuR = (1.0f + exp2muR)/(2.0f*a0);
			a1 = (1.0f - exp2muR)/(2.0f*a0);
			a2 = a0*(1.0f + 1.0f/muR);
			a3 = coshmuR - a1/muR;
		}
	}
	// evaluate() (:417-419, :504-506, :607-613, :734-740)
	NMC_HD float evaluate() const {
		if (!yukawa) {
			if (DIM == 2) return (float)(M::log_(R/r)/(2.0f*kPi));
			return (float)((1.0f/r - 1.0f/R)/(4.0f*kPi));
		}
		float mur = r*sqrtLambda;
		if (DIM == 2) {
			float K0mur = (float)bessk0(mur);
			float I0mur = (float)bessi0(mur);
			return (float)((K0mur - I0mur*a0/a1)/(2.0*kPi));
		}
		float expmur = M::exp_(-mur);
		float sinhmur = (1.0f - expmur*expmur)/(2.0f*expmur);
		return (float)((expmur - a0*sinhmur/a1)/(4.0f*kPi*r));
	}
	// poissonKernel() (:453-455, :540-542, :663-665, :795-797)
	NMC_HD float poissonKernel() const {
		if (!yukawa) return DIM == 2 ? (float)(1.0f/(2.0f*kPi)) : (float)(1.0f/(4.0f*kPi));
		if (DIM == 2) return (float)(1.0f/(2.0f*kPi*a1));
		return (float)(muR/(4.0f*kPi*a1));
	}
	// norm() (:440-442, :527-529, :650-652, :782-784)
	NMC_HD float norm_() const {
		if (!yukawa) return DIM == 2 ? R*R/4.0f : R*R/6.0f;
		if (DIM == 2) return (float)((1.0f - 2.0*kPi*poissonKernel())/lambda);
		return (float)((1.0f - 4.0*kPi*poissonKernel())/lambda);
	}
	// gradientNorm() (:428-431, :515-518, :634-641, :761-773)
	NMC_HD float gradientNorm() const {
		if (!yukawa) {
			if (DIM == 2) { float r2 = r*r; return (float)((1.0f/r2 - 1.0f/(R*R))/(2.0f*kPi)); }
			float r3 = r*r*r; return (float)((1.0f/r3 - 1.0f/(R*R*R))/(4.0f*kPi));
		}
		float mur = r*sqrtLambda;
		if (DIM == 2) {
			float K1mur = (float)bessk1(mur);
			float I1mur = (float)bessi1(mur);
			float Qr = sqrtLambda*(K1mur - I1mur*a2/a3);
			return (float)(Qr/(2.0f*kPi*r));
		}
		float r2 = r*r;
		float expmur = M::exp_(-mur);
		float exp2mur = expmur*expmur;
		float coshmur = (1.0f + exp2mur)/(2.0f*expmur);
		float sinhmur = (1.0f - exp2mur)/(2.0f*expmur);
		float K32mur = expmur*(1.0f + 1.0f/mur);
		float I32mur = coshmur - sinhmur/mur;
		float Qr = sqrtLambda*(K32mur - I32mur*a2/a3);
		return (float)(Qr/(4.0f*kPi*r2));
	}
	NMC_HD V3 gradient() const { return (yVol - c)*gradientNorm(); }
	// poissonKernelGradient() (:464-468, :551-555, :680-685, :816-821): Eigen narrows the double
	// divisor to float before the per-component division
	NMC_HD V3 poissonKernelGradient() const {
		V3 d = ySurf - c;
		if (!yukawa) {
			if (DIM == 2) return (2.0f*d)/(float)(2.0f*kPi*R*R);
			return (3.0f*d)/(float)(4.0f*kPi*R*R);
		}
		if (DIM == 2) { float QR = sqrtLambda/(R*a3); return (d*QR)/(float)(2.0f*kPi); }
		float QR = lambda/a3;
		return (d*QR)/(float)(4.0f*kPi);
	}
	// directionSampledPoissonKernel(y) (:459-461, :546-548, :669-677, :801-813)
	NMC_HD float directionSampledPoissonKernel(V3 y) const {
		if (!yukawa) return 1.0f;
		float rr = maxS(rClamp, norm(y - c));
		float mur = rr*sqrtLambda;
		if (DIM == 2) {
			float K1mur = (float)bessk1(mur);
			float I1mur = (float)bessi1(mur);
			float Q = K1mur + I1mur*a0/a1;
			return mur*Q;
		}
		float expmur = M::exp_(-mur);
		float exp2mur = expmur*expmur;
		float coshmur = (1.0f + exp2mur)/(2.0f*expmur);
		float sinhmur = (1.0f - exp2mur)/(2.0f*expmur);
		float K32mur = expmur*(1.0f + 1.0f/mur);
		float I32mur = coshmur - sinhmur/mur;
		float Q = K32mur + I32mur*a0/a1;
		return mur*Q;
	}
	// evaluate(x, y) (:422-425, :509-512, :616-631, :743-758) -- only reached with non-zero Neumann data
	NMC_HD float evaluate(V3 x, V3 y) const {
		float r1 = maxS(rClamp, norm(y - x));
		float dd = dot(x - c, y - c);
		if (!yukawa) {
			if (DIM == 2) return (float)((M::log_(R*R - dd) - M::log_(R*r1))/(2.0f*kPi));
			return (float)((1.0f/r1 - R/(R*R - dd))/(4.0f*kPi));
		}
		float r2 = (R*R - dd)/R;
		float mur1 = r1*sqrtLambda, mur2 = r2*sqrtLambda;
		if (DIM == 2) {
			float K0mur1 = (float)bessk0(mur1), K0mur2 = (float)bessk0(mur2);
			float I0mur1 = (float)bessi0(mur1), I0mur2 = (float)bessi0(mur2);
			float Q1 = K0mur1 - I0mur1*a0/a1;
			float Q2 = K0mur2 - I0mur2*a0/a1;
			return (float)((Q1 - Q2)/(2.0f*kPi));
		}
		float e1 = M::exp_(-mur1), e2 = M::exp_(-mur2);
		float s1 = (1.0f - e1*e1)/(2.0f*e1);
		float s2 = (1.0f - e2*e2)/(2.0f*e2);
		float Q1 = (e1 - a0*s1/a1)/r1;
		float Q2 = (e2 - a0*s2/a1)/r2;
		return (float)((Q1 - Q2)/(4.0f*kPi));
	}
	NMC_HD float potential() const {
		return DIM == 2 ? (float)(2.0f*kPi*poissonKernel()) : (float)(4.0f*kPi*poissonKernel());
	}
	// sampleVolume(dir, sampler, pdf): rejection sampler (:362-383) with the bounds of :403-409,
	// :591-599, :718-726; 3D harmonic closed form (:483-496)
	NMC_HD void sampleVolume(V3 dir, Pcg32& rng, float& pdf) {
		if (!yukawa && DIM == 3) {
			float u1 = rng.nextFloat();
			float u2 = rng.nextFloat();
			float phi = (float)(2.0f*kPi*u2);
			r = (1.0f + sqrtf(1.0f - M::cbrt_(u1*u1))*M::cos_(phi))*R/2.0f;
			r = maxS(rClamp, r);
			if (r > R) r = R/2.0f;
			yVol = c + r*dir;
			pdf = evaluate()/norm_();
			return;
		}
		float bound;
		if (!yukawa) bound = 1.5f/R;
		else {
			const float a = DIM == 2 ? 2.2f : 2.0f, b = DIM == 2 ? 0.6f : 0.5f;
			bound = R <= lambda ?
				maxS(maxS(a/R, a/lambda), maxS(b*sqrtf(R), b*sqrtLambda)) :
				maxS(minS(a/R, a/lambda), minS(b*sqrtf(R), b*sqrtLambda));
		}
		float nrm = norm_();
		int iter = 0;
		do {
			float u = rng.nextFloat();
			r = rng.nextFloat()*R;
			pdf = evaluate()/nrm;
			float pdfRadius = pdf/pdfSphere<DIM>(r);
			iter++;
			if (u < pdfRadius/bound) break;
		} while (iter < 1000);
		r = maxS(rClamp, r);
		if (r > R) r = R/2.0f;
		yVol = c + r*dir;
	}
};

// ---- fast-mode special functions (fp32) -------------------------------------------------------------

// Exponentially scaled modified Bessel functions i0e = e^-x I0, i1e = e^-x I1, k0e = e^x K0,
// k1e = e^x K1 from the Abramowitz-Stegun 9.8.1-9.8.8 fits (the ones bessel.hpp uses; |err| < 2e-7),
// evaluated in float with the exponential factored out so that nothing overflows for large x.
struct Bessel4 { float i0e, i1e, k0e, k1e; };
NMC_OUTLINE Bessel4 besselScaled(float x) {
	Bessel4 b;
	if (x < 3.75f) {
		float y = x*(1.0f/3.75f); y = y*y;
		float i0 = 1.0f+y*(3.5156229f+y*(3.0899424f+y*(1.2067492f+y*(0.2659732f+y*(0.360768e-1f+y*0.45813e-2f)))));
		float i1 = x*(0.5f+y*(0.87890594f+y*(0.51498869f+y*(0.15084934f+y*(0.2658733e-1f+y*(0.301532e-2f+y*0.32411e-3f))))));
		float ex = __expf(-x);
		b.i0e = i0*ex; b.i1e = i1*ex;
		if (x <= 2.0f) {
			float z = x*x*0.25f;
			float lg = __logf(0.5f*x);
			float k0 = (-lg*i0)+(-0.57721566f+z*(0.42278420f+z*(0.23069756f+z*(0.3488590e-1f+z*(0.262698e-2f+z*(0.10750e-3f+z*0.74e-5f))))));
			float k1 = (lg*i1)+(1.0f/x)*(1.0f+z*(0.15443144f+z*(-0.67278579f+z*(-0.18156897f+z*(-0.1919402e-1f+z*(-0.110404e-2f+z*(-0.4686e-4f)))))));
			float ep = 1.0f/ex;
			b.k0e = k0*ep; b.k1e = k1*ep;
			return b;
		}
	} else {
		float y = 3.75f/x;
		float rs = rsqrtf(x);
		b.i0e = rs*(0.39894228f+y*(0.1328592e-1f+y*(0.225319e-2f+y*(-0.157565e-2f+y*(0.916281e-2f+y*(-0.2057706e-1f+y*(0.2635537e-1f+y*(-0.1647633e-1f+y*0.392377e-2f))))))));
		float a = 0.2282967e-1f+y*(-0.2895312e-1f+y*(0.1787654e-1f-y*0.420059e-2f));
		b.i1e = rs*(0.39894228f+y*(-0.3988024e-1f+y*(-0.362018e-2f+y*(0.163801e-2f+y*(-0.1031555e-1f+y*a)))));
	}
	float y = 2.0f/x;
	float rs = rsqrtf(x);
	b.k0e = rs*(1.25331414f+y*(-0.7832358e-1f+y*(0.2189568e-1f+y*(-0.1062446e-1f+y*(0.587872e-2f+y*(-0.251540e-2f+y*0.53208e-3f))))));
	b.k1e = rs*(1.25331414f+y*(0.23498619f+y*(-0.3655620e-1f+y*(0.1504268e-1f+y*(-0.780353e-2f+y*(0.325614e-2f+y*(-0.68245e-3f)))))));
	return b;
}

#if defined(NMC_BESSEL_TAB) && defined(__CUDACC__)
// Default mode on the device: the same four functions from a table of cubic pieces in t = log2(x) (bessel_table.cpp),
// one code path for every x, so the lanes of a warp do not split over the polynomial regimes above.  Only ONE
// translation unit may define NMC_BESSEL_TAB (wost_fast.cu): it owns cBesselTab and fills it (ensureBesselTable).
struct BesselTabView { const float4* c; float t0, perOctave; int n; };
__constant__ BesselTabView cBesselTab;
#endif
#if defined(NMC_BESSEL_TAB) && defined(__CUDA_ARCH__)
__device__ __forceinline__ Bessel4 besselScaledTab(const BesselTabView& tab, float x) {
	const float u = (__log2f(x) - tab.t0)*tab.perOctave;
	if (!(u >= 0.0f && u < (float)tab.n)) return besselScaled(x); // outside the table: the polynomials (cold path)
	const int i = (int)u;
	const float f = u - (float)i;
	const float4* c = tab.c + 4*i;
	const float4 a = __ldg(c), b = __ldg(c + 1), k = __ldg(c + 2), l = __ldg(c + 3);
	Bessel4 r;
	r.i0e = fmaf(fmaf(fmaf(a.w, f, a.z), f, a.y), f, a.x);
	r.i1e = fmaf(fmaf(fmaf(b.w, f, b.z), f, b.y), f, b.x);
	r.k0e = fmaf(fmaf(fmaf(k.w, f, k.z), f, k.y), f, k.x);
	r.k1e = fmaf(fmaf(fmaf(l.w, f, l.z), f, l.y), f, l.x);
	return r;
}
#define NMC_BESSEL4(x) besselScaledTab(cBesselTab, (x))
#else
#define NMC_BESSEL4(x) besselScaled(x)
#endif

// BallFast: centred ball Green's function written in x = r*mu, X = R*mu (mu = sqrt(lambda)):
//   2D  G = g(x)/(2 pi),       g = K0(x) - I0(x) K0(X)/I0(X)        T = x [K1(x) + I1(x) K0(X)/I0(X)]
//   3D  G = mu g(x)/(4 pi x),  g = sinh(X - x)/sinh X               T = [x cosh(X - x) + sinh(X - x)]/sinh X
// T(x) is the un-absorbed exit probability through the sphere of radius r (the reference's
// directionSampledPoissonKernel, distributions.h:669-677, 801-813); T(0) = 1, T' = -x g (both dims),
// |G| = (1 - T(X))/lambda (:650-652, :782-784) and the radial CDF of a source sample is
// F(x) = (1 - T(x))/(1 - T(X)), which sampleX() inverts by a safeguarded Halley iteration instead of the
// reference's rejection loop (:362-383).  Harmonic (lambda = 0) limits are handled in the same frame
// with y = r/R:  2D F = y^2 (1 - 2 ln y),  3D: Ulrich's polar method (:483-496).
template <int DIM>
struct BallFast {
	bool yukawa;
	float lambda, mu;
	float R, X, TX, oneMinusTX;
	float ratio0, ratio1; // 2D: k0e(X)/i0e(X), k1e(X)/i1e(X) (multiply by e^{x-2X}); 3D: 1/(1 - e^{-2X}), K32(X)/I32(X) scaled
	float bdyFac;         // |boundaryGradientDirection| (see bdyGradFactor)

	NMC_HD void init(bool yukawa_, float lambda_) {
		yukawa = yukawa_ && lambda_ > 0.0f; lambda = lambda_; mu = sqrtf(lambda_);
		R = X = 0.0f; TX = 1.0f; oneMinusTX = 0.0f; ratio0 = ratio1 = 0.0f; bdyFac = 0.0f;
	}
	NMC_HD void update(float R_) {
		R = R_;
		if (!yukawa) { TX = 1.0f; oneMinusTX = 0.0f; bdyFac = (DIM == 2 ? 2.0f : 3.0f)/R; return; }
		X = R*mu;
		if (DIM == 2) {
			Bessel4 b = NMC_BESSEL4(X);
			ratio0 = b.k0e/b.i0e; ratio1 = b.k1e/b.i1e;
			TX = __expf(-X)/b.i0e;                       // 1/I0(X)
			oneMinusTX = X < 0.25f ? X*X*0.25f*(1.0f - X*X*(3.0f/16.0f)*(1.0f - X*X*(19.0f/108.0f))) : 1.0f - TX;
			bdyFac = mu*b.i0e/b.i1e;                     // mu I0(X)/I1(X)
		} else {
			float e2 = __expf(-2.0f*X);
			ratio0 = 1.0f/(1.0f - e2);
			TX = 2.0f*X*__expf(-X)*ratio0;               // X/sinh X
			float X2 = X*X;
			oneMinusTX = X < 0.5f ? X2*(1.0f/6.0f)*(1.0f - X2*(7.0f/60.0f)*(1.0f - X2*(31.0f/294.0f))) : 1.0f - TX;
			// K32(X)/I32(X) with I32 = cosh - sinh/X, K32 = e^-X (1 + 1/X): keep e^{-2X} explicit
			float i32e = X < 0.3f ? X2*(1.0f/3.0f)*(1.0f + X2*0.1f)*__expf(-X) : 0.5f*(1.0f + e2) - 0.5f*(1.0f - e2)/X; // e^-X I32(X)
			ratio1 = (1.0f + 1.0f/X)/i32e;               // e^{X} K32(X) / (e^{-X} I32(X))
			// mu sinh X/(cosh X - sinh X/X) = mu (1 - e2)/2 / i32e
			bdyFac = mu*0.5f*(1.0f - e2)/i32e;
		}
	}
	// T(x), g(x) and q(x) = K1(x) - I1(x) K1(X)/I1(X) (2D) / K32(x) - I32(x) K32(X)/I32(X) (3D), 0 < x <= X
	NMC_HD void evalTgq(float x, float& T, float& g, float& q) const {
		if (DIM == 2) {
			Bessel4 b = NMC_BESSEL4(x);
			float em = __expf(-x), e2 = __expf(x - 2.0f*X);
			float ep = ratio0*e2;
			T = x*(b.k1e*em + b.i1e*ep);
			g = b.k0e*em - b.i0e*ep;
			q = b.k1e*em - b.i1e*ratio1*e2;
		} else {
			float em = __expf(-x), ed = __expf(-2.0f*(X - x));
			float sh = em*(1.0f - ed)*ratio0, ch = em*(1.0f + ed)*ratio0;
			T = x*ch + sh;
			g = sh;
			float e2x = em*em, x2 = x*x;
			float k32 = em*(1.0f + 1.0f/x);
			float i32e = x < 0.3f ? x2*(1.0f/3.0f)*(1.0f + x2*0.1f)*em : 0.5f*(1.0f + e2x) - 0.5f*(1.0f - e2x)/x; // e^-x I32(x)
			q = k32 - i32e*ratio1*(em*ed);       // e^{x - 2X} = e^{-x} e^{-2(X - x)}
		}
	}
	NMC_HD void evalTg(float x, float& T, float& g) const { float q; evalTgq(x, T, g, q); }
	// |G| = integral of the Green's function over the ball
	NMC_HD float normG() const {
		if (!yukawa) return DIM == 2 ? R*R*0.25f : R*R*(1.0f/6.0f);
		return oneMinusTX/lambda;
	}
	// initial throughput after the first ball: poissonKernel()/pdf = T(X)
	NMC_HD float exitThroughput() const { return yukawa ? TX : 1.0f; }
	// throughput factor of a step ending at distance r from the centre
	NMC_HD float stepThroughput(float r) const {
		if (!yukawa) return 1.0f;
		float x = fminf(fmaxf(1e-4f, r)*mu, X);
		float T, g; evalTg(x, T, g);
		return T;
	}
	// boundaryGradientDirection = bd * bdyFac  (poissonKernelGradient()/poissonKernel(), :464-468,:680-685,:816-821)
	NMC_HD float bdyGradFactor() const { return bdyFac; }
	// sourceGradientDirection = dir * srcGradFactor  (gradient()/(pdf*norm) = d*gradientNorm()/G(r))
	NMC_HD float srcGradFactorHarmonic(float y) const { // y = r/R
		if (DIM == 2) return (1.0f/(y*R))*(1.0f - y*y)/fmaxf(-__logf(y), 1e-20f);  // r (1/r^2 - 1/R^2)/ln(R/r)
		return (1.0f/(y*R))*(1.0f + y + y*y);                                        // r (1/r^3 - 1/R^3)/(1/r - 1/R)
	}
	NMC_HD float srcGradFactor(float q, float g) const { return mu*q/fmaxf(g, 1e-30f); }
	// inverse CDF of the radial density: returns x = r*mu (Yukawa) or y = r/R (harmonic) for u in [0,1);
	// g receives g(x) (Yukawa only).  F vanishes quadratically at both ends of the interval, so Newton
	// runs on sqrt(F) (u < 1/2) or sqrt(1 - F) (u >= 1/2), which are close to linear there.
	// Tiny balls (X < 0.05) use the harmonic law: the two densities differ by O(X^2).
	NMC_HD float sampleX(float u, float u2, float& g, float& q, bool& harmonicFrame, bool refine = false) const {
		harmonicFrame = !yukawa || X < 0.05f;
		if (harmonicFrame) {
			g = 0.0f; q = 0.0f;
			if (DIM == 3) { // Ulrich's polar method, r/R = (1 + sqrt(1 - cbrt(u^2)) cos(2 pi u2))/2
				float y = 0.5f*(1.0f + sqrtf(fmaxf(0.0f, 1.0f - cbrtf(u*u)))*cosf(6.2831853f*u2));
				return fminf(fmaxf(y, 1e-6f), 1.0f);
			}
			// F = y^2 (1 - 2 ln y), F' = -4 y ln y
			bool lower = u < 0.5f;
			float target = sqrtf(lower ? u : 1.0f - u);
			float y = lower ? sqrtf(u/(1.0f - __logf(fmaxf(u, 1e-12f)))) : 1.0f - sqrtf(0.5f*(1.0f - u));
			y = fminf(fmaxf(y, 1e-6f), 0.9999f);
			for (int it = 0; it < 5; it++) {
				float ly = __logf(y);
				float F = y*y*(1.0f - 2.0f*ly), dF = fmaxf(-4.0f*y*ly, 1e-30f);
				float q = sqrtf(fmaxf(lower ? F : 1.0f - F, 1e-30f));
				float step = 2.0f*q*(q - target)/dF;              // Newton on sqrt(F) or sqrt(1-F)
				float yn = lower ? y - step : y + step;
				yn = fminf(fmaxf(yn, 0.25f*y), 0.5f*(y + 1.0f));
				bool done = fabsf(yn - y) <= 1e-6f;
				y = yn;
				if (done) break;
			}
			return fminf(fmaxf(y, 1e-6f), 1.0f);
		}
		bool lower = u < 0.5f;
		float mass = lower ? u*oneMinusTX : (1.0f - u)*oneMinusTX;  // target value of 1 - T (lower) or T - TX (upper)
		float target = sqrtf(mass);
		float x;
		if (lower) { // 1 - T ~ (x^2/2)(-ln(x/2) - 0.0772 - c) (2D), x^2/2 ... (3D: (x^2/2) coth X ~ x^2/2 (1 + ...))
			x = sqrtf(2.0f*mass);
			if (DIM == 2) { float L = fmaxf(-0.0772f - __logf(0.5f*fminf(x, 1.5f)), 0.35f); x = sqrtf(2.0f*mass/L); L = fmaxf(-0.0772f - __logf(0.5f*fminf(x, 1.5f)), 0.35f); x = sqrtf(2.0f*mass/L); }
			x = fminf(x, 0.7f*X);
		} else {
			x = X - fminf(sqrtf(2.0f*mass/TX), 0.7f*X);   // T - TX ~ TX (X - x)^2/2 near the sphere
			if (X >= 8.0f) { // large balls: the root sits where T(x) ~ sqrt(pi x/2) e^-x (2D) / (1 + x) e^-x (3D) equals TX + mass
				float l = -__logf(fminf(TX + mass, 0.999f)), xa;
				if (DIM == 2) { xa = l + 0.5f*__logf(fmaxf(1.5707963f*fmaxf(l, 0.3f), 1.0f)); xa = l + 0.5f*__logf(fmaxf(1.5707963f*xa, 1.0f)); }
				else { xa = l + __logf(1.0f + l); xa = l + __logf(1.0f + xa); }
				x = fminf(x, xa);
			}
		}
		x = fminf(fmaxf(x, 1e-6f*X), X*(1.0f - 1e-6f));
		// Fixed trip count, no early exit: every lane of the warp runs the same instruction stream (a data-dependent
		// break serialises the warp on its slowest lane anyway).  sqrt(1 - T) / sqrt(T - TX) are close to linear in x,
		// and T' = -x g, g' = -T/x (2D) / -(T - g)/x (3D) come for free, so the iteration is Halley's on
		// h(x) = sqrt(m(x)) - sqrt(mass): two steps from the analytic start reach |F(x) - u| < 1e-4 over
		// 0.05 <= X <= 100 (tests/test_host_logic.py).  With `refine` one more evaluation returns g and q AT the
		// returned x (the first-ball gradient weights need them); otherwise they belong to the last iterate.
		float T, gg = 0.0f, qq = 0.0f;
#pragma unroll 1
		for (int it = 0; it < 2; it++) {
			evalTgq(x, T, gg, qq);
			float m = fmaxf(lower ? 1.0f - T : T - TX, 1e-30f);
			float sq = sqrtf(m), h = sq - target;
			float a = fmaxf(x*gg, 1e-30f);                                  // |m'|
			float m2 = gg - (DIM == 2 ? T : T - gg);                          // g + x g'
			// lower: m' = a, m'' = m2; upper: m' = -a, m'' = -m2.  Halley step = 4 h m' m / (m'^2 (2 sq + h) - 2 h m'' m)
			float sgn = lower ? 1.0f : -1.0f;
			float den = a*a*(2.0f*sq + h) - 2.0f*h*(sgn*m2)*m;
			float step = den > 0.0f ? 4.0f*h*a*m/den : 2.0f*sq*h/a;            // Newton if Halley's denominator degenerates
			float xn = x - sgn*step;
			x = fminf(fmaxf(xn, 0.25f*x), 0.5f*(x + X));
		}
		if (refine) evalTgq(x, T, gg, qq);
		g = gg; q = qq;
		return x;
	}
};

} // namespace nmc
