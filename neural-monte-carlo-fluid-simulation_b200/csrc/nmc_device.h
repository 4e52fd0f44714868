// csrc/nmc_device.h -- internal interface between the C ABI (capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "nmc_estimator.cuh"

namespace nmc {

// device-side counters accumulated by the estimator kernels
struct Counters { unsigned long long walksStarted, walksCompleted, steps, activePoints, trips, laneSlices; };

// wost_det.cu (compiled with -fmad=false)
cudaError_t launchDeterministic(const SceneView& S, const SolverParams& o, const float* d_pts, long long n,
								unsigned long long indexOffset, float* d_p, float* d_g, float* d_lhs,
								Counters* d_counters, float* d_stats12, cudaStream_t stream);
size_t deterministicScratchFloats(int dim, const SolverParams& o, long long n);
cudaError_t launchProbe(const SceneView& S, int kind, long long n, const float* d_pts, const float* a0, const float* a1,
						const float* a2, const float* a3, const float* params, float* d_out, cudaStream_t stream);
int probeWidth(int dim, int kind);
cudaError_t launchSolutionEstimator(const SceneView& S, const SolverParams& o, const float* d_pts, const float* d_normals,
									const int* d_types, const int* d_aligned, long long n, int nWalks, unsigned long long indexOffset,
									float* d_sol, float* d_stats4, Counters* d_counters, cudaStream_t stream);

// wost_fast.cu
struct FastLaunchInfo { int grid, block, smemBytes; };
cudaError_t launchFast(const SceneView& S, const SolverParams& o, const float* d_pts, long long n,
					   unsigned long long indexOffset, float* d_p, float* d_g, unsigned int* d_workCounter,
					   Counters* d_counters, float* d_stats12, int smCount, int maxDepth, cudaStream_t stream, FastLaunchInfo* info);
cudaError_t launchProbeFast(const SceneView& S, int kind, long long n, const float* a0, const float* a1,
							const float* params, float* d_out, cudaStream_t stream);
cudaError_t launchProbePacket(const SceneView& S, int kind, long long n, const float* pts, const float* a0, const float* a1,
							  const float* a2, const float* a3, const float* params, float* d_out, cudaStream_t stream);

} // namespace nmc
