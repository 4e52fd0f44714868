// csrc/zombie_bindings.cpp -- the drop-in Python module `zombie_bindings` (compiled twice: NMC_BIND_DIM=2
// replaces bindings/zombie, NMC_BIND_DIM=3 replaces bindings/zombie3d), on top of the C ABI.
//
// Python-visible surface, identical to the reference:
//   2D (bindings/zombie/demo/demo.cpp:393-401):   wost(scene, solverConfig, outputConfig, sample_points),
//        bvc(scene, solverConfig, outputConfig), Scene(config), Scene(config, sourceValue)
//   3D (bindings/zombie3d/demo/demo.cpp:119-125): wost(...), Scene(config, sourceValue)
// wost returns tuple(sample_points, solution, gradient) as nested Python lists (the reference's STL
// casters).  Additive: wost_array(...) returning numpy arrays, set_mode(), set_seed(), last_stats().
// Where the reference abort()s or exit()s (demo/config.h:6-11, scene.h:106-109) this module raises.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/nmcfs.h"
#include "image_io.h"
#include <filesystem>

#ifndef NMC_BIND_DIM
#define NMC_BIND_DIM 2
#endif

namespace py = pybind11;
static constexpr int DIM = NMC_BIND_DIM;

static int g_mode = NMC_MODE_FAST;
static bool g_haveSeed = false;
static uint64_t g_seed = 0;
static nmc_solve_stats g_lastStats = {};

template <class T>
static T optional(const py::dict& d, const char* key, T def) {
	if (d.contains(key)) return d[key].cast<T>();
	return def;
}
template <class T>
static T required(const py::dict& d, const char* key) {
	if (!d.contains(key)) throw py::key_error(std::string("Missing required setting: ") + key);
	return d[key].cast<T>();
}

static void loadObj(const std::string& path, bool flipOrientation, std::vector<float>& verts, std::vector<int>& prims) {
	std::ifstream obj(path);
	if (!obj) throw std::runtime_error("Error opening file: " + path);
	std::string line;
	int nVerts = 0;
	while (std::getline(obj, line)) {
		std::istringstream ss(line);
		std::string token;
		ss >> token;
		if (token == "v") {
			float c[3] = {0, 0, 0};
			for (int k = 0; k < DIM; k++) ss >> c[k];
			for (int k = 0; k < DIM; k++) verts.push_back(c[k]);
			nVerts++;
		} else if (DIM == 2 && token == "l") { // demo/scene.h:122-129
			long i = 0, j = 0;
			ss >> i >> j;
			if (flipOrientation) { prims.push_back((int)j - 1); prims.push_back((int)i - 1); }
			else { prims.push_back((int)i - 1); prims.push_back((int)j - 1); }
		} else if (DIM == 3 && token == "f") { // fcpw/utilities/scene_loader.inl:124-137: position index before the first '/'
			while (ss >> token) {
				long i = std::strtol(token.c_str(), nullptr, 10);
				prims.push_back(i > 0 ? (int)i - 1 : nVerts + (int)i);
			}
		}
	}
	prims.resize(prims.size()/DIM*DIM);
}

class Scene {
public:
	nmc_scene* handle = nullptr;
	bool isWatertight = false, isDoubleSided = false;

	Scene(const py::dict& config, py::array_t<float, py::array::c_style | py::array::forcecast> source) {
		if (source.ndim() != DIM) throw py::type_error("sourceValue must be a " + std::to_string(DIM) + "-dimensional array");
		const int shape[3] = {(int)source.shape(0), (int)source.shape(1), DIM == 3 ? (int)source.shape(2) : 1};
		init(config, source.data(), shape, false, false);
	}
	// Scene(config) of the 2D module (scene.h:22-52): the source grid comes from the image file config["sourceValue"];
	// isWatertight and flipOrientation default to TRUE here (false in the two-argument form).  PFM files only.
	Scene(const py::dict& config) {
		if (DIM != 2) throw std::runtime_error("Scene(config) exists in the 2D module only");
		const std::string file = required<std::string>(config, "sourceValue");
		if (!nmc_io::hasExtension(file, "pfm")) throw std::runtime_error("Scene(config): only PFM source images are supported (" + file + ")");
		int h = 0, w = 0;
		std::vector<float> grid;
		nmc_io::readPfmGrey(file, h, w, grid);
		const int shape[3] = {h, w, 1};
		init(config, grid.data(), shape, true, true);
	}

private:
	void init(const py::dict& config, const float* source, const int* shape, bool watertightDefault, bool flipDefault) {
		isWatertight = optional<bool>(config, "isWatertight", watertightDefault);
		isDoubleSided = optional<bool>(config, "isDoubleSided", false);
		const std::string boundary = required<std::string>(config, "boundary");
		const bool normalize = DIM == 2 ? optional<bool>(config, "normalizeDomain", false) : false; // zombie3d ignores both
		const bool flip = DIM == 2 ? optional<bool>(config, "flipOrientation", flipDefault) : false;
		std::vector<float> verts; std::vector<int> prims;
		loadObj(boundary, flip, verts, prims);
		const int nV = (int)verts.size()/DIM, nP = (int)prims.size()/DIM;
		if (normalize && nV > 0) { // demo/scene.h:132-142
			float cm[2] = {0, 0};
			for (int i = 0; i < nV; i++) { cm[0] += verts[2*i]; cm[1] += verts[2*i + 1]; }
			cm[0] /= nV; cm[1] /= nV;
			float radius = 0.0f;
			for (int i = 0; i < nV; i++) {
				verts[2*i] -= cm[0]; verts[2*i + 1] -= cm[1];
				radius = std::max(radius, std::sqrt(verts[2*i]*verts[2*i] + verts[2*i + 1]*verts[2*i + 1]));
			}
			for (float& v : verts) v /= radius;
		}
		nmc_scene_opts so;
		so.absorptionCoeff = optional<float>(config, "absorptionCoeff", 0.0f);
		so.isWatertight = isWatertight; so.isDoubleSided = isDoubleSided;
		int device = 0;
		if (const char* lr = std::getenv("LOCAL_RANK")) { if (nmc_device_count() > 1) device = std::atoi(lr); }
		if (config.contains("device")) device = config["device"].cast<int>();
		handle = nmc_scene_create(DIM, verts.data(), nV, prims.data(), nP, source, shape[0], shape[1], shape[2], &so, device);
		if (!handle) throw std::runtime_error(std::string("zombie_bindings.Scene: ") + nmc_last_error());
	}

public:
	~Scene() { nmc_scene_destroy(handle); }
	Scene(const Scene&) = delete;
	Scene& operator=(const Scene&) = delete;
};

static nmc_solver_opts solverOpts(const py::dict& solver, const py::dict& output) { // demo.cpp:121-137
	nmc_solver_opts o;
	o.nWalks = optional<int>(solver, "nWalks", 128);
	o.maxWalkLength = optional<int>(solver, "maxWalkLength", 1024);
	o.stepsBeforeApplyingTikhonov = optional<int>(solver, "setpsBeforeApplyingTikhonov", o.maxWalkLength);
	o.stepsBeforeUsingMaximalSpheres = optional<int>(solver, "setpsBeforeUsingMaximalSpheres", o.maxWalkLength);
	o.epsilonShell = optional<float>(solver, "epsilonShell", 1e-3f);
	o.minStarRadius = optional<float>(solver, "minStarRadius", 1e-3f);
	o.silhouettePrecision = optional<float>(solver, "silhouettePrecision", 1e-3f);
	o.russianRouletteThreshold = optional<float>(solver, "russianRouletteThreshold", 0.0f);
	o.useGradientControlVariates = !optional<bool>(solver, "disableGradientControlVariates", false);
	o.useGradientAntitheticVariates = !optional<bool>(solver, "disableGradientAntitheticVariates", false);
	o.useCosineSamplingForDerivatives = optional<bool>(solver, "useCosineSamplingForDirectionalDerivatives", false);
	o.ignoreDirichlet = optional<bool>(solver, "ignoreDirichlet", false);
	o.ignoreNeumann = optional<bool>(solver, "ignoreNeumann", false);
	o.ignoreSource = optional<bool>(solver, "ignoreSource", false);
	(void)required<int>(output, "gridRes"); // required although unused, demo.cpp:132
	o.boundaryDistanceMask = optional<float>(output, "boundaryDistanceMask", 0.0f);
	o.mode = g_mode;
	if (g_haveSeed) o.seed = g_seed;
	else { std::random_device rd; o.seed = ((uint64_t)rd() << 32) ^ rd(); } // the reference seeds from the clock
	return o;
}

using PtsArray = py::array_t<float, py::array::c_style | py::array::forcecast>;

static void solve(const Scene& scene, const py::dict& solver, const py::dict& output, const PtsArray& pts,
				  std::vector<float>& p, std::vector<float>& g) {
	if (pts.ndim() != 2 || pts.shape(1) != DIM) throw py::type_error("sample_points must have shape (N, " + std::to_string(DIM) + ")");
	const int64_t n = pts.shape(0);
	nmc_solver_opts o = solverOpts(solver, output);
	p.assign((size_t)n, 0.0f); g.assign((size_t)n*DIM, 0.0f);
	int rc;
	nmc_solve_stats st = {};
	{
		// the reference holds the GIL for the whole solve; here other Python threads may run meanwhile -- calls on
		// the same Scene are serialised inside the C ABI (per-scene mutex)
		py::gil_scoped_release nogil;
		rc = nmc_wost_solve(scene.handle, &o, pts.data(), n, 0, p.data(), g.data(), &st);
	}
	g_lastStats = st; // under the GIL again
	if (rc != NMC_OK) throw std::runtime_error(std::string("zombie_bindings.wost: ") + nmc_last_error());
}

// wost(scene, solverConfig, outputConfig, sample_points) -> (sample_points, solution, gradient)
static py::tuple wost(const Scene& scene, const py::dict& solver, const py::dict& output, const PtsArray& pts) {
	std::vector<float> p, g;
	solve(scene, solver, output, pts, p, g);
	const int64_t n = pts.shape(0);
	// The nested lists the pybind STL casters return (demo.cpp:119,204): N x (2 DIM + 1) float objects and 2 N inner lists, built
	// with the C API (stolen references, no accessor temporaries) and with the cyclic collector paused -- millions of fresh
	// container objects would otherwise trigger a full collection every few thousand allocations (1e6 points: 0.87 s -> 0.30 s).
	auto a = pts.unchecked<2>();
#if PY_VERSION_HEX >= 0x030A0000
	const int gcWasOn = PyGC_Disable();
#endif
	PyObject *outPts = PyList_New(n), *outP = PyList_New(n), *outG = PyList_New(n);
	bool ok = outPts && outP && outG;
	for (int64_t i = 0; ok && i < n; i++) {
		PyObject *q = PyList_New(DIM), *gr = PyList_New(DIM), *pv = PyFloat_FromDouble(p[(size_t)i]);
		ok = q && gr && pv;
		for (int k = 0; ok && k < DIM; k++) {
			PyObject *x = PyFloat_FromDouble(a(i, k)), *y = PyFloat_FromDouble(g[(size_t)i*DIM + k]);
			if (!x || !y) { Py_XDECREF(x); Py_XDECREF(y); ok = false; break; }
			PyList_SET_ITEM(q, k, x); PyList_SET_ITEM(gr, k, y);
		}
		if (!ok) { Py_XDECREF(q); Py_XDECREF(gr); Py_XDECREF(pv); break; }
		PyList_SET_ITEM(outPts, i, q); PyList_SET_ITEM(outP, i, pv); PyList_SET_ITEM(outG, i, gr);
	}
#if PY_VERSION_HEX >= 0x030A0000
	if (gcWasOn) PyGC_Enable();
#endif
	if (!ok) { Py_XDECREF(outPts); Py_XDECREF(outP); Py_XDECREF(outG); throw std::bad_alloc(); }
	return py::make_tuple(py::reinterpret_steal<py::object>(outPts), py::reinterpret_steal<py::object>(outP), py::reinterpret_steal<py::object>(outG));
}

static py::tuple wostArray(const Scene& scene, const py::dict& solver, const py::dict& output, const PtsArray& pts) {
	std::vector<float> p, g;
	solve(scene, solver, output, pts, p, g);
	const int64_t n = pts.shape(0);
	py::array_t<float> ap(n), ag({n, (int64_t)DIM});
	std::copy(p.begin(), p.end(), ap.mutable_data());
	std::copy(g.begin(), g.end(), ag.mutable_data());
	return py::make_tuple(ap, ag);
}

#if NMC_BIND_DIM == 2
// bvc(scene, solverConfig, outputConfig) -> None: runBoundaryValueCaching (demo.cpp:265-363); writes outputConfig["solutionFile"]
// (default "solution.pfm") and, unless saveColormapped is false, <stem>_color<ext> (demo/grid.h:9-33, 370-415).
static py::array_t<float> bvcGrid(const Scene& scene, const py::dict& solver, const py::dict& output) {
	nmc_solver_opts o = solverOpts(solver, output);
	o.mode = NMC_MODE_DETERMINISTIC; // the cache-point estimator is the deterministic replay (its seed comes from set_seed() or the OS)
	nmc_bvc_opts b;
	b.boundaryCacheSize = optional<int>(solver, "boundaryCacheSize", 1024);
	b.domainCacheSize = optional<int>(solver, "domainCacheSize", 1024);
	b.nWalksForCachedSolutionEstimates = optional<int>(solver, "nWalksForCachedSolutionEstimates", 128);
	b.nWalksForCachedGradientEstimates = optional<int>(solver, "nWalksForCachedGradientEstimates", 640);
	b.gridRes = required<int>(output, "gridRes");
	b.normalOffsetForCachedDirichletSamples = optional<float>(solver, "normalOffsetForCachedDirichletSamples", 5.0f*o.epsilonShell);
	b.radiusClampForKernels = optional<float>(solver, "radiusClampForKernels", 1e-3f);
	b.regularizationForKernels = optional<float>(solver, "regularizationForKernels", 0.0f);
	if (b.gridRes <= 0) throw py::value_error("gridRes must be positive");
	py::array_t<float> grid({(int64_t)b.gridRes, (int64_t)b.gridRes});
	int rc;
	{
		py::gil_scoped_release nogil;
		rc = nmc_bvc_solve(scene.handle, &o, &b, grid.mutable_data(), nullptr, 0, nullptr, nullptr);
	}
	if (rc != NMC_OK) throw std::runtime_error(std::string("zombie_bindings.bvc: ") + nmc_last_error());
	return grid;
}

static void bvc(const Scene& scene, const py::dict& solver, const py::dict& output) {
	py::array_t<float> grid = bvcGrid(scene, solver, output);
	const int res = (int)grid.shape(0);
	const std::string solutionFile = optional<std::string>(output, "solutionFile", "solution.pfm");
	const bool saveColormapped = optional<bool>(output, "saveColormapped", true);
	const std::string colormap = optional<std::string>(output, "colormap", "");
	const float lo = optional<float>(output, "colormapMinVal", 0.0f), hi = optional<float>(output, "colormapMaxVal", 1.0f);
	// solution->get(j, i) = value of point (i, j): image row <-> y index, column <-> x index (grid.h:388-411)
	std::vector<float> img((size_t)3*res*res), col((size_t)3*res*res);
	const float* g = grid.data();
	for (int i = 0; i < res; i++) for (int j = 0; j < res; j++) {
		const float v = g[(size_t)i*res + j];
		float* px = &img[(size_t)3*((size_t)j*res + i)];
		px[0] = px[1] = px[2] = v;
		nmc_io::applyColormap(std::min(std::max((v - lo)/(hi - lo), 0.0f), 1.0f), colormap, &col[(size_t)3*((size_t)j*res + i)]);
	}
	std::filesystem::path path(solutionFile);
	if (!path.parent_path().empty()) std::filesystem::create_directories(path.parent_path());
	nmc_io::writeImage3(solutionFile, res, res, img);
	if (saveColormapped) nmc_io::writeImage3((path.parent_path()/path.stem()).string() + "_color" + path.extension().string(), res, res, col);
}
#endif

PYBIND11_MODULE(zombie_bindings, m) {
	m.doc() = "pybind11 WoSt"; // as the reference
	m.def("wost", &wost);
#if NMC_BIND_DIM == 2
	m.def("bvc", &bvc);
	m.def("bvc_grid", &bvcGrid, "additive: the masked evaluation grid of bvc() as a numpy array [i][j] instead of image files");
#endif
	auto cls = py::class_<Scene>(m, "Scene");
#if NMC_BIND_DIM == 2
	cls.def(py::init<const py::dict&>());
#endif
	cls.def(py::init<const py::dict&, PtsArray>());
	// additive entry points (the reference has none of these)
	m.def("wost_array", &wostArray, "numpy in, numpy out: returns (p[N], grad[N, dim])");
	m.def("set_mode", [](const std::string& mode) {
		if (mode == "fast") g_mode = NMC_MODE_FAST;
		else if (mode == "deterministic") g_mode = NMC_MODE_DETERMINISTIC;
		else throw py::value_error("mode must be 'fast' or 'deterministic'");
	});
	m.def("set_seed", [](py::object seed) { if (seed.is_none()) g_haveSeed = false; else { g_haveSeed = true; g_seed = seed.cast<uint64_t>(); } });
	m.def("last_stats", []() {
		py::dict d;
		d["walks_started"] = g_lastStats.walks_started; d["walks_completed"] = g_lastStats.walks_completed;
		d["walk_steps"] = g_lastStats.walk_steps; d["active_points"] = g_lastStats.active_points;
		d["kernel_ms"] = g_lastStats.kernel_ms; d["total_ms"] = g_lastStats.total_ms; d["kernel_launches"] = g_lastStats.kernel_launches;
		return d;
	});
}
