// csrc/bessel_table.h -- table of cubic pieces of the scaled modified Bessel functions (see bessel_table.cpp)
#pragma once
#include <vector>

namespace nmc {

struct BesselTable {
	float t0;        // log2 of the first node
	int perOctave;   // intervals per octave of x
	int n;           // intervals
	std::vector<float> coef; // n x 4 functions x 4 coefficients
};
const BesselTable& besselTable(); // built once per process (thread-safe)

} // namespace nmc
