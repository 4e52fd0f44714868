/* oracle/nmc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Monte Carlo pressure-projection path
 * (bindings/zombie{,3d}: walk-on-stars estimator + FCPW scalar BVH queries + pcg32 + bessel).
 * Every function cites the reference file:line it follows.  Pinned against the reference's own
 * code compiled as oracle/_ref (see oracle/ref_harness.cpp) by tests/test_oracle_vs_ref.py and
 * against the committed vectors in tests/golden/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product (libnmcfs.so) never does.
 */
#ifndef NMC_ORACLE_H
#define NMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t state, inc; } nmo_pcg32;

typedef struct nmo_scene nmo_scene;

typedef struct {
	int nWalks;
	int maxWalkLength;
	int stepsBeforeApplyingTikhonov;
	int stepsBeforeUsingMaximalSpheres;
	float epsilonShell;
	float minStarRadius;
	float silhouettePrecision;
	float russianRouletteThreshold;
	int useGradientControlVariates;
	int useGradientAntitheticVariates;
	int useCosineSamplingForDerivatives;
	int ignoreDirichlet;
	int ignoreNeumann;
	int ignoreSource;
	float boundaryDistanceMask;
} nmo_solver_opts;

/* rng */
void nmo_pcg32_seed(nmo_pcg32* s, uint64_t initstate, uint64_t initseq);
uint32_t nmo_pcg32_uint(nmo_pcg32* s);
uint32_t nmo_pcg32_bounded(nmo_pcg32* s, uint32_t bound);
float nmo_pcg32_float(nmo_pcg32* s);
uint64_t nmo_point_seed(uint64_t seed, uint64_t index);
void nmo_stratified(int dim, uint64_t initstate, int nSamples, float* out, uint64_t* state_out);
void nmo_sphere_dir(int dim, const float* u, int n, float* out);

/* bessel: kind 0 i0, 1 i1, 2 k0, 3 k1, 4 k2 */
void nmo_bessel(int kind, const double* x, int n, double* out);

/* ball green's function probes: same layout as ref_greens_ball / ref_sample_volume */
void nmo_greens_ball(int dim, float lambda, const float* R, const float* r, int n, float* out);
void nmo_sample_volume(int dim, float lambda, const float* R, const uint64_t* seeds, int n,
					   float* r_out, float* pdf_out, int* draws_out);

/* scene: verts nV x dim, prims nP x dim (segments in 2D, triangles in 3D), source grid
 * 2D [n0=h][n1=w], 3D [n0][n1][n2] */
nmo_scene* nmo_scene_create(int dim, const float* verts, int nV, const int* prims, int nP,
							const float* src, int n0, int n1, int n2,
							float absorption, int watertight, int doubleSided);
void nmo_scene_destroy(nmo_scene* s);
void nmo_scene_bbox(const nmo_scene* s, float* out);
int nmo_scene_num_nodes(const nmo_scene* s);
/* per node 16 floats: pMin[3] pMax[3] axis[3] halfAngle offset silOffset nRefs nSilRefs 0 0 */
void nmo_scene_nodes(const nmo_scene* s, float* out);

/* geometric query probes (layouts as the ref_* probes) */
void nmo_dist_neumann(const nmo_scene* s, const float* pts, int n, int signed_, float* out);
void nmo_dist_dirichlet(const nmo_scene* s, const float* pts, int n, float* out);
void nmo_inside_domain(const nmo_scene* s, const float* pts, int n, int* out);
void nmo_outside_bbox(const nmo_scene* s, const float* pts, int n, int* out);
void nmo_star_radius(const nmo_scene* s, const float* pts, int n, float minR, const float* maxR,
					 float prec, int flip, float* out);
void nmo_intersect_neumann(const nmo_scene* s, const float* org, const float* nrm, const float* dir,
						   const float* tmax, const int* onb, int n, float* out);
void nmo_blocked(const nmo_scene* s, const float* xi, const float* xj, const float* ni, const float* nj,
				 const int* offi, const int* offj, int n, int* out);
void nmo_offset_point(int dim, const float* p, const float* nrm, int n, float* out);
void nmo_source(const nmo_scene* s, const float* pts, int n, float* out);

/* the estimator; stats layout as ref_wost (12 floats per point, may be NULL) */
int nmo_wost(const nmo_scene* s, const nmo_solver_opts* o, const float* pts, int n,
			 uint64_t seed, uint64_t index_offset, int nthreads,
			 float* p_out, float* grad_out, float* stats);

#ifdef __cplusplus
}
#endif
#endif
