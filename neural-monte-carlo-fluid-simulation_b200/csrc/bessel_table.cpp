// csrc/bessel_table.cpp -- host side of the default mode's Bessel lookup table (2D screened Poisson).
//
// The 2D Yukawa ball functions need the exponentially scaled modified Bessel functions
//   i0e(x) = e^-x I0(x), i1e(x) = e^-x I1(x), k0e(x) = e^x K0(x), k1e(x) = e^x K1(x)
// several times per walk step.  The reference evaluates the Numerical-Recipes / Abramowitz-Stegun piecewise
// polynomials of deps/bessel/bessel.hpp:373-556 (three regimes: x <= 2, x < 3.75, x >= 3.75).  On a GPU the lanes
// of a warp sit at different x, so a warp pays for all three regimes; the default mode therefore reads one table of
// cubic pieces in t = log2(x) (every function is smooth in t: the logarithm of K0 at 0 becomes linear, the
// algebraic / exponential tails become exponentials of t), built here in double precision from the integral
// representations
//   i_ne(x) = (1/pi) int_0^pi e^{x (cos s - 1)} cos(n s) ds,     k_ne(x) = int_0^inf e^{-x (cosh s - 1)} cosh(n s) ds
// with the trapezoidal rule (exponentially convergent for both).  Layout: per interval 4 x (c0, c1, c2, c3), one
// float4 per function in the order i0e, i1e, k0e, k1e; f(t) = ((c3 u + c2) u + c1) u + c0 with u in [0, 1).
#include "bessel_table.h"

#include <cmath>
#include <mutex>

namespace nmc {

static void scaledBessel(double x, double f[4], double dfdx[4]) {
	const double pi = 3.14159265358979323846;
	// I: periodic trapezoid on [0, pi]; the integrand is peaked at 0 with width 1/sqrt(x)
	const int N = 8192;
	double i0 = 0.0, i1 = 0.0;
	for (int k = 0; k <= N; k++) {
		const double s = pi*k/N, w = (k == 0 || k == N) ? 0.5 : 1.0;
		const double e = std::exp(x*(std::cos(s) - 1.0));
		i0 += w*e; i1 += w*e*std::cos(s);
	}
	i0 /= N; i1 /= N;
	// K: trapezoid on [0, S] until the integrand underflows
	const double h = 0.01;
	double k0 = 0.0, k1 = 0.0;
	for (int k = 0;; k++) {
		const double s = h*k, c = std::cosh(s), a = x*(c - 1.0);
		if (a > 745.0) break;
		const double e = std::exp(-a), w = k == 0 ? 0.5 : 1.0;
		k0 += w*e; k1 += w*e*c;
	}
	k0 *= h; k1 *= h;
	f[0] = i0; f[1] = i1; f[2] = k0; f[3] = k1;
	// I0' = I1, I1' = I0 - I1/x, K0' = -K1, K1' = -K0 - K1/x, and the scaling factors e^-x / e^x
	dfdx[0] = i1 - i0; dfdx[1] = i0 - i1/x - i1; dfdx[2] = -k1 + k0; dfdx[3] = -k0 - k1/x + k1;
}

const BesselTable& besselTable() {
	static BesselTable T;
	static std::once_flag once;
	std::call_once(once, [] {
		T.t0 = -14.0f; T.perOctave = 16; T.n = 22*T.perOctave;   // x in [2^-14, 2^8): rClamp*sqrt(lambda) .. beyond the reference's float overflow (91.9)
		T.coef.resize((size_t)T.n*16);
		const double ln2 = 0.69314718055994530942, hT = 1.0/T.perOctave;
		std::vector<double> val((size_t)(T.n + 1)*4), der((size_t)(T.n + 1)*4);
		for (int i = 0; i <= T.n; i++) {
			const double t = (double)T.t0 + hT*i, x = std::exp2(t);
			double f[4], d[4];
			scaledBessel(x, f, d);
			for (int k = 0; k < 4; k++) { val[(size_t)4*i + k] = f[k]; der[(size_t)4*i + k] = d[k]*x*ln2*hT; } // df/du, u = (t - t_i)/hT
		}
		for (int i = 0; i < T.n; i++) for (int k = 0; k < 4; k++) { // cubic Hermite piece on u in [0, 1]
			const double f0 = val[(size_t)4*i + k], f1 = val[(size_t)4*(i + 1) + k], d0 = der[(size_t)4*i + k], d1 = der[(size_t)4*(i + 1) + k];
			float* c = &T.coef[(size_t)16*i + 4*k];
			c[0] = (float)f0; c[1] = (float)d0; c[2] = (float)(3.0*(f1 - f0) - 2.0*d0 - d1); c[3] = (float)(2.0*(f0 - f1) + d0 + d1);
		}
	});
	return T;
}

} // namespace nmc
