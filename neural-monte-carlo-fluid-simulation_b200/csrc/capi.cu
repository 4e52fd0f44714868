// csrc/capi.cu -- the C ABI of libnmcfs.so (include/nmcfs.h): scene upload, solve entry points, probes.
#include "../../include/nmcfs.h"
#include "nmc_device.h"
#include "scene_build.h"
#include "bessel_table.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace nmc;

static thread_local std::string g_err;
int nmcFail(int code, const std::string& msg) { g_err = msg; return code; }
static int fail(int code, const std::string& msg) { return nmcFail(code, msg); }
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(NMC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

#include "capi_scene.h"

extern "C" const char* nmc_last_error(void) { return g_err.c_str(); }

extern "C" int nmc_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

extern "C" int nmc_bessel_table(float* out, int capacity_floats, float* t0, int* per_octave) {
	const BesselTable& T = besselTable();
	if (t0) *t0 = T.t0;
	if (per_octave) *per_octave = T.perOctave;
	if (out) for (size_t i = 0; i < T.coef.size() && i < (size_t)(capacity_floats > 0 ? capacity_floats : 0); i++) out[i] = T.coef[i];
	return T.n;
}

extern "C" uint64_t nmc_point_seed(uint64_t seed, uint64_t index) { return pointSeed(seed, index); }

template <class T>
static cudaError_t upload(T*& dst, const std::vector<Q4>& src) {
	dst = nullptr;
	if (src.empty()) return cudaSuccess;
	cudaError_t e = cudaMalloc((void**)&dst, src.size()*sizeof(Q4));
	if (e != cudaSuccess) return e;
	return cudaMemcpy(dst, src.data(), src.size()*sizeof(Q4), cudaMemcpyHostToDevice);
}

static int setSource(nmc_scene* s, const float* src, int n0, int n1, int n2, int isDevice, cudaStream_t stream = 0, bool async = false) {
	if (!src || n0 <= 0 || n1 <= 0 || (s->flat.dim == 3 && n2 <= 0)) return fail(NMC_ERR_INVALID, "source grid: null pointer or empty shape");
	size_t count = (size_t)n0*n1*(s->flat.dim == 3 ? n2 : 1);
	if (count > s->srcCap) {
		if (s->d_src) cudaFree(s->d_src);
		s->d_src = nullptr; s->srcCap = 0;
		CK(cudaMalloc((void**)&s->d_src, count*sizeof(float)));
		s->srcCap = count;
	}
	if (async) CK(cudaMemcpyAsync(s->d_src, src, count*sizeof(float), isDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
	else CK(cudaMemcpy(s->d_src, src, count*sizeof(float), isDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
	s->view.src = s->d_src; s->view.n0 = n0; s->view.n1 = n1; s->view.n2 = s->flat.dim == 3 ? n2 : 1;
	{ // texel index = (int)(x*scale + off) for the default mode: axis k of the box maps to n_k texels (2D: rows <-> y)
		const SceneView& v = s->view;
		const int nAx[3] = {s->flat.dim == 2 ? n1 : n0, s->flat.dim == 2 ? n0 : n1, s->flat.dim == 3 ? n2 : 1};
		for (int k = 0; k < 3; k++) {
			float ext = v.bboxHi[k] - v.bboxLo[k];
			s->view.srcScale[k] = ext > 0.0f ? (float)nAx[k]/ext : 0.0f;
			s->view.srcOff[k] = -v.bboxLo[k]*s->view.srcScale[k];
		}
	}
	return NMC_OK;
}

extern "C" nmc_scene* nmc_scene_create(int dim, const float* verts, int nV, const int* prims, int nP,
									   const float* src, int n0, int n1, int n2,
									   const nmc_scene_opts* opts, int device) {
	if ((dim != 2 && dim != 3) || !opts || nV < 0 || nP < 0 || (nV > 0 && !verts) || (nP > 0 && !prims)) {
		fail(NMC_ERR_INVALID, "nmc_scene_create: bad arguments"); return nullptr;
	}
	for (long long i = 0; i < (long long)nP*dim; i++) if (prims[i] < 0 || prims[i] >= nV) {
		fail(NMC_ERR_INVALID, "nmc_scene_create: primitive index out of range"); return nullptr;
	}
	if (nmc_device_count() <= 0) { fail(NMC_ERR_NO_DEVICE, "no CUDA device: libnmcfs has no CPU path"); return nullptr; }
	if (cudaSetDevice(device) != cudaSuccess) { fail(NMC_ERR_CUDA, "cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
	nmc_scene* s = new nmc_scene();
	s->device = device;
	cudaDeviceGetAttribute(&s->smCount, cudaDevAttrMultiProcessorCount, device);
	buildFlatScene(dim, verts, nV, prims, nP, opts->isDoubleSided != 0, s->flat);
	s->verts.assign(verts, verts + (size_t)nV*dim);
	s->prims.assign(prims, prims + (size_t)nP*dim);
	SceneView& v = s->view;
	memset(&v, 0, sizeof(v));
	v.dim = dim; v.nNodes = s->flat.nNodes; v.nPrims = s->flat.nPrims; v.nSilRefs = s->flat.nSilRefs;
	for (int k = 0; k < 3; k++) { v.bboxLo[k] = s->flat.bboxLo[k]; v.bboxHi[k] = s->flat.bboxHi[k]; }
	v.absorption = opts->absorptionCoeff; v.watertight = opts->isWatertight != 0; v.doubleSided = opts->isDoubleSided != 0;
	bool ok = upload(s->d_nodes, s->flat.nodes) == cudaSuccess && upload(s->d_prims, s->flat.prims) == cudaSuccess &&
			  upload(s->d_primN, s->flat.primN) == cudaSuccess && upload(s->d_nrmV, s->flat.nrmV) == cudaSuccess &&
			  upload(s->d_sils, s->flat.sils) == cudaSuccess && upload(s->d_silsU, s->flat.silsU) == cudaSuccess && upload(s->d_grpP, s->flat.grpP) == cudaSuccess && upload(s->d_grpS, s->flat.grpS) == cudaSuccess && upload(s->d_rayP, s->flat.rayP) == cudaSuccess && upload(s->d_rayN, s->flat.rayN) == cudaSuccess && upload(s->d_supP, s->flat.supP) == cudaSuccess && upload(s->d_supS, s->flat.supS) == cudaSuccess && upload(s->d_coneF, s->flat.coneF) == cudaSuccess && upload(s->d_silsF, s->flat.silsF) == cudaSuccess && upload(s->d_treeF, s->flat.treeF) == cudaSuccess &&
			  cudaMalloc((void**)&s->d_counters, sizeof(Counters)) == cudaSuccess &&
			  cudaMalloc((void**)&s->d_workCounter, sizeof(unsigned int)) == cudaSuccess;
	for (int i = 0; ok && i < 4; i++) ok = cudaEventCreate(&s->ev[i]) == cudaSuccess;
	if (!ok) { fail(NMC_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(cudaGetLastError())); nmc_scene_destroy(s); return nullptr; }
	v.nodes = s->d_nodes; v.prims = s->d_prims; v.primN = s->d_primN; v.nrmV = s->d_nrmV; v.sils = s->d_sils;
	v.silsU = s->d_silsU; v.nSilU = s->flat.nSilU; v.grpP = s->d_grpP; v.grpS = s->d_grpS; v.rayP = s->d_rayP; v.rayN = s->d_rayN; v.nRay = s->flat.nRay; v.supP = s->d_supP; v.supS = s->d_supS; v.coneF = s->d_coneF; v.silsF = s->d_silsF; v.treeF = s->d_treeF;
	if (setSource(s, src, n0, n1, n2, 0) != NMC_OK) { nmc_scene_destroy(s); return nullptr; }
	return s;
}

extern "C" void nmc_scene_destroy(nmc_scene* s) {
	if (!s) return;
	cudaSetDevice(s->device);
	cudaFree(s->d_nodes); cudaFree(s->d_prims); cudaFree(s->d_primN); cudaFree(s->d_nrmV); cudaFree(s->d_sils); cudaFree(s->d_silsU); cudaFree(s->d_grpP); cudaFree(s->d_grpS); cudaFree(s->d_rayP); cudaFree(s->d_rayN); cudaFree(s->d_supP); cudaFree(s->d_supS); cudaFree(s->d_coneF); cudaFree(s->d_silsF); cudaFree(s->d_treeF);
	cudaFree(s->d_src); cudaFree(s->d_work); cudaFree(s->d_lhs); cudaFree(s->d_counters); cudaFree(s->d_workCounter);
	for (int i = 0; i < 4; i++) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
	cudaGetLastError();
	delete s;
}

extern "C" int nmc_scene_set_source(nmc_scene* s, const float* src, int n0, int n1, int n2, int src_is_device) {
	if (!s) return fail(NMC_ERR_INVALID, "null scene");
	Lock lock(s->mu);
	CK(cudaSetDevice(s->device));
	return setSource(s, src, n0, n1, n2, src_is_device);
}
extern "C" int nmc_scene_set_source_async(nmc_scene* s, const float* src, int n0, int n1, int n2, int src_is_device, void* stream) {
	if (!s) return fail(NMC_ERR_INVALID, "null scene");
	Lock lock(s->mu);
	CK(cudaSetDevice(s->device));
	return setSource(s, src, n0, n1, n2, src_is_device, (cudaStream_t)stream, true);
}
extern "C" int nmc_scene_dim(const nmc_scene* s) { return s ? s->flat.dim : 0; }
extern "C" int nmc_scene_bbox(const nmc_scene* s, float* out) {
	if (!s || !out) return fail(NMC_ERR_INVALID, "null argument");
	for (int k = 0; k < s->flat.dim; k++) { out[k] = s->flat.bboxLo[k]; out[s->flat.dim + k] = s->flat.bboxHi[k]; }
	return NMC_OK;
}
extern "C" int nmc_scene_num_nodes(const nmc_scene* s) { return s ? s->flat.nNodes : 0; }
extern "C" int nmc_scene_nodes(const nmc_scene* s, float* out) {
	if (!s || !out) return fail(NMC_ERR_INVALID, "null argument");
	for (int i = 0; i < s->flat.nNodes; i++) {
		const Q4* q = &s->flat.nodes[(size_t)4*i]; float* o = out + (size_t)i*16;
		int nRefs, second, refOffset, silOffset, nSil;
		memcpy(&nRefs, &q[0].w, 4); memcpy(&second, &q[1].w, 4);
		memcpy(&refOffset, &q[3].x, 4); memcpy(&silOffset, &q[3].y, 4); memcpy(&nSil, &q[3].z, 4);
		o[0] = q[0].x; o[1] = q[0].y; o[2] = q[0].z; o[3] = q[1].x; o[4] = q[1].y; o[5] = q[1].z;
		o[6] = q[2].x; o[7] = q[2].y; o[8] = q[2].z; o[9] = q[2].w;
		o[10] = (float)refOffset; o[11] = (float)silOffset; o[12] = (float)nRefs; o[13] = (float)nSil; o[14] = (float)second; o[15] = 0.0f;
	}
	return NMC_OK;
}

int nmcToParams(const nmc_solver_opts* o, SolverParams& p);
static int toParams(const nmc_solver_opts* o, SolverParams& p) { return nmcToParams(o, p); }
int nmcToParams(const nmc_solver_opts* o, SolverParams& p) {
	if (!o) return fail(NMC_ERR_INVALID, "null solver options");
	if (o->nWalks < 0 || o->maxWalkLength < 0) return fail(NMC_ERR_INVALID, "negative nWalks / maxWalkLength");
	if (o->mode != NMC_MODE_FAST && o->mode != NMC_MODE_DETERMINISTIC) return fail(NMC_ERR_INVALID, "unknown mode");
	p.nWalks = o->nWalks; p.maxWalkLength = o->maxWalkLength;
	p.stepsBeforeApplyingTikhonov = o->stepsBeforeApplyingTikhonov;
	p.stepsBeforeUsingMaximalSpheres = o->stepsBeforeUsingMaximalSpheres;
	p.epsilonShell = o->epsilonShell; p.minStarRadius = o->minStarRadius;
	p.silhouettePrecision = o->silhouettePrecision; p.russianRouletteThreshold = o->russianRouletteThreshold;
	p.useGradientControlVariates = o->useGradientControlVariates; p.useGradientAntitheticVariates = o->useGradientAntitheticVariates;
	p.useCosineSampling = o->useCosineSamplingForDerivatives != 0;
	p.ignoreDirichlet = o->ignoreDirichlet; p.ignoreNeumann = o->ignoreNeumann; p.ignoreSource = o->ignoreSource;
	p.boundaryDistanceMask = o->boundaryDistanceMask; p.seed = o->seed;
	return NMC_OK;
}

// deterministic-mode scratch is bounded by processing the points in chunks
static const long long kDetChunk = 1ll << 19;

static int solveDevice(nmc_scene* s, const nmc_solver_opts* opts, const float* d_pts, int64_t n, uint64_t indexOffset,
					   float* d_p, float* d_g, float* d_stats12, cudaStream_t stream, nmc_solve_stats* stats) {
	SolverParams p;
	int rc = toParams(opts, p);
	if (rc != NMC_OK) return rc;
	if (n < 0 || (n > 0 && (!d_pts || !d_p || !d_g))) return fail(NMC_ERR_INVALID, "null buffer");
	const int dim = s->flat.dim;
	int launches = 0;
	CK(cudaMemsetAsync(s->d_counters, 0, sizeof(Counters), stream));
	CK(cudaEventRecord(s->ev[0], stream));
	if (opts->mode == NMC_MODE_DETERMINISTIC) {
		for (long long b = 0; b < n; b += kDetChunk) {
			long long m = n - b < kDetChunk ? n - b : kDetChunk;
			size_t need = deterministicScratchFloats(dim, p, m);
			if (need > s->lhsCap) {
				CK(cudaStreamSynchronize(stream));
				if (s->d_lhs) cudaFree(s->d_lhs);
				s->d_lhs = nullptr; s->lhsCap = 0;
				CK(cudaMalloc((void**)&s->d_lhs, (need > 0 ? need : 1)*sizeof(float)));
				s->lhsCap = need;
			}
			CK(launchDeterministic(s->view, p, d_pts + b*dim, m, indexOffset + (uint64_t)b, d_p + b, d_g + b*dim, s->d_lhs,
								   s->d_counters, d_stats12 ? d_stats12 + b*12 : nullptr, stream));
			launches++;
		}
	} else {
		FastLaunchInfo info;
		CK(cudaMemsetAsync(s->d_workCounter, 0, sizeof(unsigned int), stream));
		CK(launchFast(s->view, p, d_pts, n, indexOffset, d_p, d_g, s->d_workCounter, s->d_counters, d_stats12, s->smCount, s->flat.maxDepth, stream, &info));
		launches++;
	}
	CK(cudaEventRecord(s->ev[1], stream));
	if (stats) {
		Counters c;
		CK(cudaMemcpyAsync(&c, s->d_counters, sizeof(c), cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		stats->walks_started = c.walksStarted; stats->walks_completed = c.walksCompleted;
		stats->walk_steps = c.steps; stats->active_points = c.activePoints;
		stats->warp_trips = c.trips; stats->lane_slices = c.laneSlices;
		CK(cudaEventElapsedTime(&stats->kernel_ms, s->ev[0], s->ev[1]));
		stats->kernel_launches = launches;
	}
	return NMC_OK;
}

extern "C" int nmc_wost_solve_device(nmc_scene* s, const nmc_solver_opts* opts, const float* d_pts, int64_t n,
									 uint64_t index_offset, float* d_p_out, float* d_grad_out, void* stream,
									 nmc_solve_stats* stats) {
	if (!s) return fail(NMC_ERR_INVALID, "null scene");
	Lock lock(s->mu);
	CK(cudaSetDevice(s->device));
	if (stats) memset(stats, 0, sizeof(*stats));
	cudaStream_t st = (cudaStream_t)stream;
	CK(cudaEventRecord(s->ev[2], st));
	int rc = solveDevice(s, opts, d_pts, n, index_offset, d_p_out, d_grad_out, nullptr, st, stats);
	if (rc != NMC_OK) return rc;
	if (stats) {
		CK(cudaEventRecord(s->ev[3], st));
		CK(cudaEventSynchronize(s->ev[3]));
		CK(cudaEventElapsedTime(&stats->total_ms, s->ev[2], s->ev[3]));
	}
	return NMC_OK;
}

static int ensureWork(nmc_scene* s, size_t floats) {
	if (floats <= s->workCap) return NMC_OK;
	if (s->d_work) cudaFree(s->d_work);
	s->d_work = nullptr; s->workCap = 0;
	CK(cudaMalloc((void**)&s->d_work, floats*sizeof(float)));
	s->workCap = floats;
	return NMC_OK;
}

// Undocumented extension used by the parity tests: when NMC_STATS12 points are requested the per-point
// statistics (layout of oracle/ref_harness.cpp) are written after grad_out by nmc_wost_solve_stats.
extern "C" int nmc_wost_solve_stats(nmc_scene* s, const nmc_solver_opts* opts, const float* pts, int64_t n,
									uint64_t index_offset, float* p_out, float* grad_out, float* stats12,
									nmc_solve_stats* stats) {
	if (!s) return fail(NMC_ERR_INVALID, "null scene");
	if (n < 0 || (n > 0 && (!pts || !p_out || !grad_out))) return fail(NMC_ERR_INVALID, "null buffer");
	Lock lock(s->mu);
	CK(cudaSetDevice(s->device));
	if (stats) memset(stats, 0, sizeof(*stats));
	if (n == 0) return NMC_OK;
	const int dim = s->flat.dim;
	size_t nn = (size_t)n;
	int rc = ensureWork(s, nn*(size_t)(2*dim + 1 + (stats12 ? 12 : 0)));
	if (rc != NMC_OK) return rc;
	float* d_pts = s->d_work; float* d_p = d_pts + nn*dim; float* d_g = d_p + nn; float* d_st = stats12 ? d_g + nn*dim : nullptr;
	cudaStream_t st = 0;
	CK(cudaEventRecord(s->ev[2], st));
	CK(cudaMemcpyAsync(d_pts, pts, nn*dim*sizeof(float), cudaMemcpyHostToDevice, st));
	rc = solveDevice(s, opts, d_pts, n, index_offset, d_p, d_g, d_st, st, stats);
	if (rc != NMC_OK) return rc;
	CK(cudaMemcpyAsync(p_out, d_p, nn*sizeof(float), cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(grad_out, d_g, nn*dim*sizeof(float), cudaMemcpyDeviceToHost, st));
	if (stats12) CK(cudaMemcpyAsync(stats12, d_st, nn*12*sizeof(float), cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(s->ev[3], st));
	CK(cudaEventSynchronize(s->ev[3]));
	if (stats) CK(cudaEventElapsedTime(&stats->total_ms, s->ev[2], s->ev[3]));
	return NMC_OK;
}

extern "C" int nmc_wost_solve(nmc_scene* s, const nmc_solver_opts* opts, const float* pts, int64_t n,
							  uint64_t index_offset, float* p_out, float* grad_out, nmc_solve_stats* stats) {
	return nmc_wost_solve_stats(s, opts, pts, n, index_offset, p_out, grad_out, nullptr, stats);
}

int nmcProbeUnlocked(nmc_scene* s, int kind, int64_t n, const float* pts, const float* aux0, const float* aux1,
						 const float* aux2, const float* aux3, const float* params, float* out);
extern "C" int nmc_probe(nmc_scene* s, int kind, int64_t n, const float* pts, const float* aux0, const float* aux1,
						 const float* aux2, const float* aux3, const float* params, float* out) {
	if (!s || !out || n < 0) return fail(NMC_ERR_INVALID, "bad arguments");
	if (n == 0) return NMC_OK;
	Lock lock(s->mu);
	return nmcProbeUnlocked(s, kind, n, pts, aux0, aux1, aux2, aux3, params, out);
}
int nmcProbeUnlocked(nmc_scene* s, int kind, int64_t n, const float* pts, const float* aux0, const float* aux1,
						 const float* aux2, const float* aux3, const float* params, float* out) {
	CK(cudaSetDevice(s->device));
	const int dim = s->flat.dim;
	int W = probeWidth(dim, kind);
	size_t nn = (size_t)n;
	// widths of the optional inputs per probe kind
	size_t wp = pts ? dim : 0, w0 = 0, w1 = 0, w2 = 0, w3 = 0;
	switch (kind) {
		case NMC_PROBE_STAR_RADIUS: case NMC_PROBE_STAR_RADIUS_PACKET: w0 = 1; break;
		case NMC_PROBE_RAY: case NMC_PROBE_RAY_PACKET: w0 = dim; w1 = dim; w2 = 1; w3 = 1; break;
		case NMC_PROBE_GREENS: case NMC_PROBE_GREENS_FAST: case NMC_PROBE_SAMPLE_RADIUS_FAST: w0 = 1; w1 = 1; break;
		case NMC_PROBE_SAMPLE_VOLUME: w0 = 1; w1 = 2; break;
		default: break;
	}
	if ((w0 && !aux0) || (w1 && !aux1) || (w2 && !aux2) || (w3 && !aux3)) return fail(NMC_ERR_INVALID, "probe: missing input");
	float* d = nullptr;
	size_t total = nn*(wp + w0 + w1 + w2 + w3 + W) + 8;
	CK(cudaMalloc((void**)&d, total*sizeof(float)));
	float* d_pts = d; float* d0 = d_pts + nn*wp; float* d1 = d0 + nn*w0; float* d2 = d1 + nn*w1; float* d3 = d2 + nn*w2;
	float* d_par = d3 + nn*w3; float* d_out = d_par + 8;
	cudaError_t e = cudaSuccess;
	if (wp) e = cudaMemcpy(d_pts, pts, nn*wp*4, cudaMemcpyHostToDevice);
	if (!e && w0) e = cudaMemcpy(d0, aux0, nn*w0*4, cudaMemcpyHostToDevice);
	if (!e && w1) e = cudaMemcpy(d1, aux1, nn*w1*4, cudaMemcpyHostToDevice);
	if (!e && w2) e = cudaMemcpy(d2, aux2, nn*w2*4, cudaMemcpyHostToDevice);
	if (!e && w3) e = cudaMemcpy(d3, aux3, nn*w3*4, cudaMemcpyHostToDevice);
	float par[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	if (params) memcpy(par, params, 4*sizeof(float));
	if (!e) e = cudaMemcpy(d_par, par, sizeof(par), cudaMemcpyHostToDevice);
	if (!e) {
		if (kind == NMC_PROBE_GREENS_FAST || kind == NMC_PROBE_SAMPLE_RADIUS_FAST) e = launchProbeFast(s->view, kind, n, d0, d1, d_par, d_out, 0);
		else if (kind == NMC_PROBE_STAR_RADIUS_PACKET || kind == NMC_PROBE_RAY_PACKET || kind == NMC_PROBE_CLOSEST_PACKET) {
			if (!wp) e = cudaErrorInvalidValue;
			else e = launchProbePacket(s->view, kind, n, d_pts, d0, d1, d2, d3, d_par, d_out, 0);
		}
		else e = launchProbe(s->view, kind, n, wp ? d_pts : nullptr, d0, d1, d2, d3, d_par, d_out, 0);
	}
	if (!e) e = cudaMemcpy(out, d_out, nn*W*4, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e) return fail(NMC_ERR_CUDA, std::string("probe: ") + cudaGetErrorString(e));
	return NMC_OK;
}
