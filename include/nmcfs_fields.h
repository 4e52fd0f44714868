/* include/nmcfs_fields.h -- C ABI of the grid post-processing kernels in libnmcfs.so:
 * the semi-Lagrangian density advection and the Taylor-Green velocity error the reference evaluates on the CPU
 * with numpy + scipy.ndimage.map_coordinates (src/2d/move_density.py:92-153, src/3d/move_density.py:184-215).
 * All pointers are DEVICE pointers on the current CUDA device; `stream` is a cudaStream_t (may be 0).
 */
#ifndef NMCFS_FIELDS_H
#define NMCFS_FIELDS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* nmc_fields_last_error(void);

/* One advection step of a node-centred density grid d[n0][n1]([n2]) (row-major, fp32):
 *   node (i,j,k) sits at  x = lo + (i,j,k)/n * extent              (np.indices / N * extent + lo, :98-101)
 *   back = x - dt * vel[i][j][k][:]                                (:130-131)
 *   pos  = (back - lo) * n / extent      in index units, per axis  (:133)
 *   out[i][j][k] = linear interpolation of d_in at pos             (map_coordinates(order=1, prefilter=False))
 * mode 0 = 'constant', cval 0 (2D script): a position outside [0, n-1] on any axis gives 0, no blending;
 * mode 1 = 'nearest' (3D script): positions are clamped to [0, n-1].
 * dim = 2 or 3; shape[dim]; vel has dim floats per node; d_out must not alias d_in. */
int nmc_advect_density(int dim, const int* shape, const float* d_in, const float* vel, float dt,
					   const float* lo, const float* extent, int mode, float* d_out, void* stream);

/* Semi-Lagrangian back-trace of the advection fit (model_split.py:97-103): out[s][a] = clamp(x[s][a] - dt u[s][a],
 * lo[a], hi[a]); lo / hi are HOST arrays of dim floats. */
int nmc_backtrace(int dim, const float* x, const float* u, int64_t n, float dt, const float* lo, const float* hi, float* out, void* stream);

/* sum over n samples of || u[s][:] - u_ref[s][:] ||^2 (dim floats per sample) accumulated in double into *out_sum,
 * which the caller zero-fills; the mean is the Taylor-Green error metric (move_density.py:143-146). */
int nmc_sum_squared_error(int dim, const float* u, const float* u_ref, int64_t n, double* out_sum, void* stream);

#ifdef __cplusplus
}
#endif
#endif
