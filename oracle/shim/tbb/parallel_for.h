// oracle/shim/tbb/parallel_for.h -- TEST INFRASTRUCTURE (see blocked_range.h).
#pragma once
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "blocked_range.h"
namespace tbb {
namespace this_task_arena { inline int current_thread_index() { return 0; } } // progress reporting only (walk_on_stars.h:97)
template <typename T, typename Body>
void parallel_for(const blocked_range<T>& range, const Body& body) {
	const T b = range.begin(), e = range.end();
	if (e <= b) return;
	unsigned nt = std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	const T chunk = 16;
	if ((e - b) <= chunk || nt == 1) { body(range); return; }
	std::atomic<T> next(b);
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < nt; t++) pool.emplace_back([&]() {
		for (;;) {
			T s = next.fetch_add(chunk);
			if (s >= e) break;
			body(blocked_range<T>(s, std::min(e, s + chunk)));
		}
	});
	for (auto& th : pool) th.join();
}
}
