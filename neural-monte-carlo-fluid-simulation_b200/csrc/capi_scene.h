// csrc/capi_scene.h -- the scene object behind the C ABI handle, shared by capi.cu and bvc.cu.
#pragma once
#include "../../include/nmcfs.h"
#include "nmc_device.h"
#include "scene_build.h"
#include <mutex>
#include <string>
#include <vector>

struct nmc_scene {
	int device = 0, smCount = 148;
	nmc::FlatScene flat;
	std::vector<float> verts;   // the mesh as passed to nmc_scene_create (boundary value caching samples its primitives)
	std::vector<int> prims;
	nmc::SceneView view;
	float4 *d_nodes = nullptr, *d_prims = nullptr, *d_primN = nullptr, *d_nrmV = nullptr, *d_sils = nullptr, *d_silsU = nullptr, *d_grpP = nullptr, *d_grpS = nullptr, *d_rayP = nullptr, *d_rayN = nullptr, *d_supP = nullptr, *d_supS = nullptr, *d_coneF = nullptr, *d_silsF = nullptr, *d_treeF = nullptr;
	float* d_src = nullptr; size_t srcCap = 0;
	// grow-only work buffers
	float* d_work = nullptr; size_t workCap = 0;      // points + outputs for the host-buffer entry point
	float* d_lhs = nullptr; size_t lhsCap = 0;        // deterministic-mode Latin-hypercube scratch
	nmc::Counters* d_counters = nullptr;
	unsigned int* d_workCounter = nullptr;
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	// the work buffers, counters and events above are per-scene mutable state: calls on one scene are serialised
	// (the reference serialises them by holding the GIL, demo.cpp:119; the bindings here release it)
	std::mutex mu;
};
typedef std::lock_guard<std::mutex> Lock;

// capi.cu
int nmcFail(int code, const std::string& msg);
int nmcToParams(const nmc_solver_opts* o, nmc::SolverParams& p);
int nmcProbeUnlocked(nmc_scene* s, int kind, int64_t n, const float* pts, const float* aux0, const float* aux1,
					 const float* aux2, const float* aux3, const float* params, float* out);
