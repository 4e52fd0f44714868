// oracle/shim/tbb/blocked_range.h -- TEST INFRASTRUCTURE.  Stand-in for the vendored TBB 2020 (which needs its own cmake
// build): the two TBB names the reference's solver headers use, tbb::blocked_range<int> and tbb::parallel_for(range, body),
// implemented on std::thread.  Put on the include path BEFORE the reference's deps/tbb/include by oracle/Makefile; the
// reference sources themselves are compiled unmodified.
#pragma once
namespace tbb {
template <typename T>
class blocked_range {
public:
	blocked_range(T b, T e): b_(b), e_(e) {}
	T begin() const { return b_; }
	T end() const { return e_; }
private:
	T b_, e_;
};
}
