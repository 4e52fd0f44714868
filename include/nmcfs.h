/* include/nmcfs.h -- C ABI of libnmcfs.so, the B200 (sm_100a) Monte Carlo pressure-projection library.
 *
 * This is the drop-in boundary for the hot path of Pranav-Jain/Neural-Monte-Carlo-Fluid-Simulation:
 * everything the reference's pybind module `zombie_bindings` does between Python and the CPU
 * solver (bindings/zombie/demo/demo.cpp:119-205,393-401; bindings/zombie3d/demo/demo.cpp:15-125)
 * maps onto the entry points below.  Plain pointers and sizes only; no C++ or torch types.
 * OBJ / dict parsing stays in the binding layer (csrc/zombie_bindings.cpp), as in the reference
 * (demo/scene.h:104-145, fcpw/utilities/scene_loader.inl:99-150).
 *
 * All functions return 0 on success (or a handle) and set nmc_last_error() otherwise; nothing
 * aborts the process (the reference abort()s, demo/config.h:9-10).  There is NO CPU fallback:
 * without a CUDA device every compute entry point fails with NMC_ERR_NO_DEVICE.
 */
#ifndef NMCFS_H
#define NMCFS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMC_OK 0
#define NMC_ERR_INVALID 1
#define NMC_ERR_NO_DEVICE 2
#define NMC_ERR_CUDA 3
#define NMC_ERR_UNSUPPORTED 4

typedef struct nmc_scene nmc_scene;

/* Scene options = the "scene" section of wost.json as read by the 2-argument Scene constructors
 * (demo/scene.h:54-63, scene_3d.h:22-31). normalizeDomain / flipOrientation are applied by the
 * binding layer to the vertex/segment arrays before they reach this ABI. */
typedef struct {
	float absorptionCoeff; /* screening coefficient lambda (0 = Poisson) */
	int isWatertight;      /* default 0 for the 2-argument constructors */
	int isDoubleSided;
} nmc_scene_opts;

/* estimator modes */
#define NMC_MODE_FAST 0          /* default: counter-based RNG, fp32 math, inverse-CDF radial sampling;
                                    matches the reference statistically */
#define NMC_MODE_DETERMINISTIC 1 /* replays the reference's arithmetic and RNG consumption; per point
                                    pcg32(nmc_point_seed(seed, global index), 1), per pair seed drawn
                                    from the point's stream (oracle/ref_harness.cpp) */

/* Solver options = the "solver" + "output" sections as read by runWalkOnStars_sampled
 * (demo.cpp:121-137), with the reference's defaults applied by the caller. */
typedef struct {
	int nWalks;                          /* 128 */
	int maxWalkLength;                   /* 1024 */
	int stepsBeforeApplyingTikhonov;     /* key "setpsBeforeApplyingTikhonov" [sic]; default maxWalkLength */
	int stepsBeforeUsingMaximalSpheres;  /* key "setpsBeforeUsingMaximalSpheres" [sic]; default maxWalkLength */
	float epsilonShell;                  /* 1e-3 */
	float minStarRadius;                 /* 1e-3 */
	float silhouettePrecision;           /* 1e-3 */
	float russianRouletteThreshold;      /* 0 */
	int useGradientControlVariates;      /* !disableGradientControlVariates */
	int useGradientAntitheticVariates;   /* !disableGradientAntitheticVariates */
	int useCosineSamplingForDerivatives; /* unsupported (NMC_ERR_UNSUPPORTED when set) */
	int ignoreDirichlet;
	int ignoreNeumann;
	int ignoreSource;
	float boundaryDistanceMask;          /* output.boundaryDistanceMask, default 0 */
	int mode;                            /* NMC_MODE_* */
	uint64_t seed;                       /* global seed; results depend only on (seed, global point index) */
} nmc_solver_opts;

/* Per-call counters and timings (all optional outputs). */
typedef struct {
	uint64_t walks_started;   /* one walk = one sampler.seed()+walk() call (walk_on_stars.h:579-581) */
	uint64_t walks_completed; /* walks that were averaged in (Russian roulette / Dirichlet shell) */
	uint64_t walk_steps;      /* iterations of the walk loop (walk_on_stars.h:145) */
	uint64_t active_points;   /* points with estimationQuantity != None */
	float kernel_ms;          /* device time of the estimator kernel(s), CUDA events on the launch stream */
	float total_ms;           /* device time of the whole call incl. copies */
	int kernel_launches;
	uint64_t warp_trips;      /* default mode: trips of the per-warp slice loop ... */
	uint64_t lane_slices;     /* ... and busy lanes summed over those trips (lane_slices / (32 warp_trips) = lane occupancy) */
} nmc_solve_stats;

const char* nmc_last_error(void);
int nmc_device_count(void);

/* Scene(config, sourceValue): boundary mesh + source grid, uploaded to `device`.
 *  verts  nV x dim floats; prims nP x dim vertex indices (segments in 2D, triangles in 3D).
 *  src    2D: [n0 = rows (y)][n1 = cols (x)] (demo/image.h:59-75); 3D: [n0][n1][n2] <-> (x, y, z)
 *         (scene_3d.h:120-126); n2 ignored in 2D. */
nmc_scene* nmc_scene_create(int dim, const float* verts, int nV, const int* prims, int nP,
							const float* src, int n0, int n1, int n2,
							const nmc_scene_opts* opts, int device);
void nmc_scene_destroy(nmc_scene* scene);
/* Replace the source grid without rebuilding the boundary structure (the reference rebuilds the
 * whole Scene every step, src/2d/models/model_split.py:191). src_is_device != 0: device pointer. */
int nmc_scene_set_source(nmc_scene* scene, const float* src, int n0, int n1, int n2, int src_is_device);
/* Same, enqueued on `stream` (a cudaStream_t): ordered after the kernel that produced a device-resident grid and
 * before a following nmc_wost_solve_device on the same stream; no host synchronisation. */
int nmc_scene_set_source_async(nmc_scene* scene, const float* src, int n0, int n1, int n2, int src_is_device, void* stream);
int nmc_scene_dim(const nmc_scene* scene);
int nmc_scene_bbox(const nmc_scene* scene, float* lo_hi /* 2*dim */);
int nmc_scene_num_nodes(const nmc_scene* scene);
/* per node 16 floats: lo[3] hi[3] axis[3] halfAngle refOffset silOffset nRefs nSilRefs secondChild 0 */
int nmc_scene_nodes(const nmc_scene* scene, float* out);

/* wost(scene, solverConfig, outputConfig, sample_points): HOST buffers in, HOST buffers out.
 *  pts n x dim; p_out n; grad_out n x dim.  index_offset = global index of pts[0] (multi-GPU shards). */
int nmc_wost_solve(nmc_scene* scene, const nmc_solver_opts* opts, const float* pts, int64_t n,
				   uint64_t index_offset, float* p_out, float* grad_out, nmc_solve_stats* stats);
/* Same with DEVICE buffers on the scene's device, enqueued on `stream` (a cudaStream_t, may be 0).
 * stats (if non-null) forces a synchronise at the end of the call. */
int nmc_wost_solve_device(nmc_scene* scene, const nmc_solver_opts* opts, const float* d_pts, int64_t n,
						  uint64_t index_offset, float* d_p_out, float* d_grad_out, void* stream,
						  nmc_solve_stats* stats);

/* nmc_wost_solve plus per-point estimator statistics for the parity tests: stats12 (may be NULL) receives
 * 12 floats per point in the layout of oracle/ref_harness.cpp ref_wost: unmasked solution mean, solution
 * variance, gradient mean[3], gradient variance[3], mean first-source term, number of averaged walks,
 * mean walk length, estimationQuantity != None. */
int nmc_wost_solve_stats(nmc_scene* scene, const nmc_solver_opts* opts, const float* pts, int64_t n,
						 uint64_t index_offset, float* p_out, float* grad_out, float* stats12,
						 nmc_solve_stats* stats);

/* EstimationQuantity::Solution (walk_on_stars.h:354-461) at caller-given sample points, deterministic replay: the
 * estimator boundary value caching runs at its cache points.  HOST buffers.  normals n x dim (may be NULL), types per
 * point 0 = in the domain, 2 = ON the reflecting boundary (zombie::SampleType; NULL = all 0), aligned = the point's
 * estimateBoundaryNormalAligned flag (double-sided scenes; NULL = all 0).  Stream of point i: nmc_point_seed(opts->seed,
 * index_offset + i), not re-seeded between walks.  stats4_out (may be NULL): variance, number of averaged walks, mean walk
 * length, first sphere radius per point. */
int nmc_estimate_solution(nmc_scene* scene, const nmc_solver_opts* opts, const float* pts, const float* normals,
						  const int* types, const int* aligned, int64_t n, int n_walks, uint64_t index_offset,
						  float* solution_out, float* stats4_out);

/* Boundary value caching, bvc(scene, solverConfig, outputConfig) of the 2D bindings (demo.cpp:265-363): the solver options
 * it reads beyond nmc_solver_opts. */
typedef struct {
	int boundaryCacheSize;                        /* 1024 */
	int domainCacheSize;                          /* 1024 */
	int nWalksForCachedSolutionEstimates;         /* 128 */
	int nWalksForCachedGradientEstimates;         /* 640; unused: there are no Dirichlet cache points in the bindings */
	int gridRes;                                  /* outputConfig["gridRes"] */
	float normalOffsetForCachedDirichletSamples;  /* 5 * epsilonShell */
	float radiusClampForKernels;                  /* 1e-3 */
	float regularizationForKernels;               /* 0 */
} nmc_bvc_opts;
/* Runs the whole pipeline on the scene's device and returns the masked evaluation grid: grid_out[i * gridRes + j] is the
 * value at point (i, j) of createEvaluationGrid (demo/grid.h:352-368) after saveEvaluationGrid's masking (:388-411);
 * writing the image files is left to the binding layer.  cache_out (may be NULL): up to cache_cap boundary cache points x 6
 * floats (x, y, nx, ny, estimated solution, pdf).  2D scenes only. */
int nmc_bvc_solve(nmc_scene* scene, const nmc_solver_opts* opts, const nmc_bvc_opts* bvc, float* grid_out,
				  float* cache_out, int cache_cap, int* n_cache_out, int* n_domain_out);
/* The splat stage alone (Splatter::splat, splatter.h:53-116, 203-290), HOST buffers: cache8 = n_cache records of 8 floats
 * (x, y, nx, ny, value, normal derivative, pdf, kind: 0 boundary, 2 boundary normal-aligned, 1 source); out[i] receives the
 * sum of the three groups' means for every evaluation point whose Dirichlet distance is >= the cut-off. */
int nmc_bvc_splat(int dim, float absorption, const float* eval_pts, const float* eval_dirichlet_dist, int64_t n_eval,
				  const float* cache8, int n_cache, float radius_clamp, float regularization, float dirichlet_dist_cutoff,
				  float* out);

/* Seeding rule of the deterministic mode (splitmix64 finaliser of seed + golden*(index+1)). */
uint64_t nmc_point_seed(uint64_t seed, uint64_t index);

/* Device-side probes used by the parity tests: each evaluates one building block on the GPU for n
 * inputs (host buffers).  Layouts follow oracle/ref_harness.cpp. */
#define NMC_PROBE_DIST_NEUMANN 0        /* in: pts            out: 1 float  (unsigned distance) */
#define NMC_PROBE_SIGNED_DIST_NEUMANN 1 /* in: pts            out: 1 float */
#define NMC_PROBE_DIST_DIRICHLET 2      /* in: pts            out: 1 float */
#define NMC_PROBE_INSIDE_DOMAIN 3       /* in: pts            out: 1 float (0/1) */
#define NMC_PROBE_STAR_RADIUS 4         /* in: pts, aux0 = maxRadius[n]; params: minR, prec, flip  out: 1 float */
#define NMC_PROBE_RAY 5                 /* in: pts, aux0 = normal[n*dim], aux1 = dir[n*dim], aux2 = tmax[n], aux3 = onBoundary[n] (as float)
                                           out: 2+2*dim floats: hit, dist, pt, normal */
#define NMC_PROBE_SOURCE 6              /* in: pts            out: 1 float */
#define NMC_PROBE_GREENS 7              /* in: pts unused; aux0 = R[n], aux1 = r[n]; params: lambda  out: 10 floats (ref_greens_ball) */
#define NMC_PROBE_SAMPLE_VOLUME 8       /* aux0 = R[n], aux1 = seeds as 2 x uint32 per entry; params: lambda  out: r, pdf, draws */
#define NMC_PROBE_GREENS_FAST 9         /* fast-mode fp32 ball functions: aux0 = R[n], aux1 = r[n]  out: 10 floats */
#define NMC_PROBE_SAMPLE_RADIUS_FAST 10 /* fast-mode inverse-CDF radial sampler: aux0 = R[n], aux1 = u[n]  out: r, pdf */
/* the default mode's warp-packet tree queries (csrc/nmc_packet.cuh): 32 consecutive inputs form one packet */
#define NMC_PROBE_STAR_RADIUS_PACKET 11 /* inputs and output of NMC_PROBE_STAR_RADIUS */
#define NMC_PROBE_RAY_PACKET 12         /* inputs and output of NMC_PROBE_RAY */
#define NMC_PROBE_CLOSEST_PACKET 13     /* in: pts            out: 2 floats: unsigned distance, signed distance */
int nmc_probe(nmc_scene* scene, int kind, int64_t n, const float* pts, const float* aux0, const float* aux1,
			  const float* aux2, const float* aux3, const float* params, float* out);

/* The default mode's table of the exponentially scaled modified Bessel functions i0e, i1e, k0e, k1e (2D screened
 * Poisson; csrc/bessel_table.cpp): cubic pieces in t = log2(x), 16 floats per interval (4 functions x c0..c3,
 * f = ((c3 u + c2) u + c1) u + c0, u = (t - t0) * per_octave - interval).  Host-only (no device needed): returns the
 * number of intervals and copies up to capacity_floats coefficients.  For the tests. */
int nmc_bessel_table(float* out, int capacity_floats, float* t0, int* per_octave);

/* Pipe-throughput peaks of `device`, measured by micro-benchmarks (csrc/peaks.cu), in warp-instructions per
 * second: out3[0] fp32 FMA (FMA pipe; ~2/3 of the dispatch limit on B200), out3[1] MUFU (ex2), out3[2] fp64 FMA. */
int nmc_measure_peaks(int device, float* out3);

/* Issue (dispatch) limit of `device`: one warp instruction per SM sub-partition per clock.  out2[0] = 4 x SMs x the SM
 * clock measured under load (clock64 against globaltimer) in warp-instructions per second, out2[1] = that clock in Hz.
 * Denominator of the issue-bound roofline bench.py reports for the walk kernels. */
int nmc_measure_issue_peak(int device, float* out2);

#ifdef __cplusplus
}
#endif
#endif
