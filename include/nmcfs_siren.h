/* include/nmcfs_siren.h -- C ABI of the fused SIREN neural-field kernels in libnmcfs.so.
 *
 * Replaces the stock-PyTorch evaluation / fit of the velocity network on the hot path
 * (src/2d/models/networks.py:15-68 MLP + Sine; src/2d/models/base.py:83-96 update_network;
 * src/2d/models/model_split.py:88-120, 246-284 fit loops; utils/diff_ops.py:45-51 divergence).
 * All pointers are DEVICE pointers on the current CUDA device; `stream` is a cudaStream_t (may be 0).
 * Layer l has weight W[l] (row-major [out_l][in_l], the nn.Linear layout, so state_dicts are
 * interchangeable) and bias b[l]; l = 0 .. n_hidden_layers + 1.
 */
#ifndef NMCFS_SIREN_H
#define NMCFS_SIREN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int in_dim;          /* 2 or 3 (1..3 accepted) */
	int out_dim;         /* 2 or 3 (1..3 accepted) */
	int hidden;          /* 64 or 128 */
	int n_hidden_layers; /* number of H x H layers */
	float w0;            /* 30 (Sine.forward, networks.py:19-21) */
} nmc_siren_shape;

/* Boundary envelope applied to the network output inside the kernels (query_velocity, src/2d/models/base.py:158-224,
 * src/3d/models/base.py:172-260).  May be passed as NULL (= kind 0, none).
 * kind 1: wall weights on every output component i (< in_dim), the taylorgreen / vortex_collide branches:
 *   w_i(x) = min(|x_i - lo_i|, |x_i - hi_i|) clamped to [0, eps] / eps.  The reference detaches these weights.
 * kind 2: general, in the reference's order of operations:
 *   1. region override: inside the region (region_kind 1: box [region_lo, region_hi]; 2: ball, centre region_lo,
 *      radius region_hi[0], strict <) the components in region_mask take region_vel (karman inlet strip on u,
 *      base.py:170-171; smoke_obs inlet ball on w, 3d base.py:225-229).  No gradient reaches the network there.
 *   2. obstacle weight clamp(|x - sphere_c| - sphere_r, 0, eps)/eps on every component when has_sphere
 *      (smoothstep_circular_obs, base.py:352-358).  NOT detached in the reference: nmc_siren_backward adds the
 *      gradient that reaches x through it.
 *   3. wall weights (as kind 1) on the components in wall_mask (karman: v only, base.py:175-180; karman3d: u and v,
 *      3d base.py:262-270).
 * The fields after eps are ignored for kind 0 and 1. */
typedef struct {
	int kind;
	float lo[3], hi[3];
	float eps;
	int wall_mask;
	int has_sphere; float sphere_c[3]; float sphere_r;
	int region_kind; int region_mask; float region_lo[3], region_hi[3], region_vel[3];
	int sphere_axes;            /* bit i: coordinate i enters the obstacle distance; 0 = all (sphere / circle).  0b101: the
	                               cylinder along y of karman3d (cylinder_obstacle_function, src/3d/sources.py:141-145) */
	float region_noise[3];      /* region value = region_vel[j] + region_noise[j] * u, u uniform in [-1, 1), one u per sample
	                               (the 3D smoke inlet, src/3d/models/base.py:197-209); u is a hash of the sample's
	                               coordinates and *noise_seed, i.e. fixed within a time step */
	const unsigned* noise_seed; /* DEVICE pointer to the time step (read at kernel time: CUDA-graph replays see updates); may be NULL */
} nmc_siren_envelope;

const char* nmc_siren_last_error(void);

/* y[n][out] = net(x[n][in]).  z_saved (may be NULL) receives the pre-activations of every sine layer,
 * (n_hidden_layers + 1) * hidden * n floats laid out [layer][neuron][sample], for nmc_siren_backward. */
int nmc_siren_forward(const nmc_siren_shape* shape, const float* const* W, const float* const* b, const float* x,
					  int64_t n, float* y, float* z_saved, const nmc_siren_envelope* env, void* stream);

/* Backward "delta chain": given grad_y = dL/dy [n][out] and the z_saved of the forward pass, writes
 *   dZ[l][j][s] = dL/dz_l and A[l][j][s] = sin(w0 z_l) for l = 0 .. n_hidden_layers (layout of z_saved;
 *   A has (L+1)*hidden*n floats, dZ has ((L+1)*hidden + out_dim)*n: its last out_dim rows receive grad_y'),
 *   and, if grad_x != NULL, dL/dx [n][in].  dZ and A may both be NULL when only dL/dx is wanted (the divergence of a
 *   frozen network): nothing but grad_x is written then.
 * The parameter gradients are then GEMMs over the batch dimension (done by the caller, one batched call):
 *   dW_0 = dZ_0 x,  dW_l = dZ_l A_{l-1}^T (l = 1..L),  dW_last = grad_y'^T A_L^T,  db_l = rowsum(dZ_l),
 *   with grad_y' = grad_y times the (detached) envelope weights when an envelope is given. */
int nmc_siren_backward(const nmc_siren_shape* shape, const float* const* W, const float* const* b, const float* x,
					   int64_t n, const float* z_saved, const float* grad_y, float* dZ, float* A,
					   float* grad_x, const nmc_siren_envelope* env, void* stream);

/* Backward, stage 2: every weight / bias gradient from the dZ and A of nmc_siren_backward in one launch,
 * accumulated (atomics over batch splits) into gW[l], gb[l], which the caller zero-fills. */
int nmc_siren_weight_grads(const nmc_siren_shape* shape, const float* x, int64_t n, const float* dZ, const float* A,
						   float* const* gW, float* const* gb, void* stream);

/* Tensor-core forward (tcgen05, 3xTF32 split => fp32-level accuracy); same contract as nmc_siren_forward, including the
 * optional z_saved for a following nmc_siren_backward.  Returns an error (never takes another path) on unsupported
 * shapes (hidden 64 | 128, at least one hidden layer). */
int nmc_siren_forward_tc(const nmc_siren_shape* shape, const float* const* W, const float* const* b, const float* x,
						 int64_t n, float* y, float* z_saved, const nmc_siren_envelope* env, void* stream);

/* Tensor-core (tcgen05, 3xTF32) backward of a fit iteration, two launches with the contracts of the fp32 pair above:
 * nmc_siren_backward_tc writes dZ (layout of nmc_siren_backward, including the out_dim rows of grad_y'); it does not
 * produce A or grad_x (the weight-gradient kernel recomputes A = sin(w0 z) from z_saved while staging its operands).
 * nmc_siren_weight_grads_tc accumulates every weight / bias gradient into the zero-filled gW / gb: hidden layers as
 * H x H x batch tcgen05 GEMMs, first / last layer and bias sums on the FMA pipe.  n must be a multiple of 4; dZ, z_saved
 * and the hidden layers' gW must be 16-byte aligned.  Errors (never another path) on unsupported shapes. */
int nmc_siren_backward_tc(const nmc_siren_shape* shape, const float* const* W, const float* const* b, const float* x,
						  int64_t n, const float* z_saved, const float* grad_y, float* dZ,
						  const nmc_siren_envelope* env, void* stream);
int nmc_siren_weight_grads_tc(const nmc_siren_shape* shape, const float* x, int64_t n, const float* dZ, const float* z_saved,
							  float* const* gW, float* const* gb, void* stream);

/* The same backward pass in ONE launch for hidden = 64 networks with at most 6 hidden layers (csrc/siren_tc_fused_bwd.cu):
 * delta chain and every weight / bias gradient, deltas and activations kept in shared memory, gradient tiles accumulated in
 * TMEM and added to the zero-filled gW / gb at the end.  Nothing but gW / gb is written (no dZ).  Errors on other shapes. */
int nmc_siren_backward_fused_tc(const nmc_siren_shape* shape, const float* const* W, const float* x, int64_t n,
								const float* z_saved, const float* grad_y, float* const* gW, float* const* gb,
								const nmc_siren_envelope* env, void* stream);

/* MSE loss of a fit iteration in one launch: diff = y - target, grad_y = dL/dy = diff * 2/count, *loss = mean(diff^2)
 * (count = n * out_dim floats; base.py:83-96 with the loss of model_split.py:113). */
int nmc_mse_grad(const float* y, const float* target, int64_t count, float* diff, float* grad_y, float* loss, void* stream);

/* torch.optim.Adam step (no weight decay, no amsgrad) over one flat parameter buffer; step is 1-based. */
int nmc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
				  float eps, int64_t step, void* stream);

/* The same update with the step counter in DEVICE memory (int64, zero before the first call): the call first adds
 * 1 to *step on the stream, then applies the update with the bias corrections of the new value.  This is the variant
 * to capture in a CUDA graph: a replayed nmc_adam_step would keep the corrections of its capture-time step. */
int nmc_adam_step_device(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
						 float eps, long long* step, void* stream);

/* ---- per-iteration glue of the fit loops, for CUDA-graph replay (csrc/fit_glue.cu) ------------------------------------------
 * The random draws of an iteration are keyed by (seed, *epoch, *step, element): `step` is Adam's device-side counter (the
 * iteration index within a fit), `epoch` a device int64 the host bumps once per fit, so a captured iteration draws fresh
 * numbers at every replay without a host-side generator.  24-bit uniforms in [0, 1) like torch.rand. */

/* sample_in_training with the 'random' pattern (src/2d/models/base.py:225-241, utils/model_utils.py:22-31): n points uniform in
 * the box [lo, hi) (host arrays of `dim` floats).  obstacle = {centre[dim], radius} (host) or NULL: a point inside the ball is
 * redrawn once so that the batch keeps its shape (the reference drops it). */
int nmc_fit_sample_uniform(int dim, const float* lo, const float* hi, int64_t n, float* out, const long long* step,
						   const long long* epoch, uint64_t seed, const float* obstacle, void* stream);

/* The projection fit's batch (src/2d/models/model_split.py:272-277): idx = min(floor(u * *count), cap - 1) per element,
 * out_x[i] = src_x[idx], out_g[i] = src_g[idx] (rows of `dim` floats; *count is a device float). */
int nmc_fit_gather(int dim, int64_t n, const float* src_x, const float* src_g, const float* count, int64_t cap, float* out_x,
				   float* out_g, const long long* step, const long long* epoch, uint64_t seed, void* stream);

/* Targets computed ahead of time, a chunk of iterations per launch (the fit target depends only on the frozen previous network
 * and on the batch): ring_x / ring_t / ring_s (optional) hold `slots` batches of `count` floats each; the captured iteration
 * copies slot (*step % slots) -- Adam's device-side counter = the iteration index -- into the fixed buffers it reads. */
int nmc_fit_fetch(int64_t count, int slots, const float* ring_x, const float* ring_t, const float* ring_s, const long long* step,
				  float* out_x, float* out_t, float* out_s, void* stream);

/* nmc_mse_grad with the rest of an iteration's bookkeeping in the same launch: the target is target - sub when sub != NULL
 * (u_prev - grad p of the projection fit), `zero` (zero_count floats: the flat gradient buffer) is cleared, and *step_advance
 * (Adam's device-side counter) is incremented -- every kernel that reads the counter as the iteration index was launched
 * earlier on the stream, nmc_adam_update_device later.  Early stop without a host round trip (base.py:148 leaves the loop at the
 * first iteration whose loss is <= 1.1e-10): when stop_flag != NULL and the loss is <= stop_threshold, *stop_flag (device int,
 * cleared by the caller at the start of a fit) is set and stays set. */
int nmc_mse_grad_fit(const float* y, const float* target, const float* sub, int64_t count, float* diff, float* grad_y, float* loss,
					 float* zero, int64_t zero_count, long long* step_advance, float stop_threshold, int* stop_flag, void* stream);

/* nmc_adam_step_device without the increment (the counter was advanced by nmc_mse_grad_fit).  With stop_flag != NULL and
 * *stop_flag != 0 the parameters are left alone: iterations replayed between the loss reaching the threshold and the host's
 * next test of the flag do not move a fit the reference has already ended. */
int nmc_adam_update_device(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
						   float eps, const long long* step, const int* stop_flag, void* stream);
/* nmc_adam_update_device of one iteration and nmc_fit_fetch of the next in ONE launch (both read the device-side step, neither
 * writes it): used between the iterations of an unrolled fit graph. */
int nmc_adam_update_fetch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
						  float eps, const long long* step, const int* stop_flag, int64_t count, int slots, const float* ring_x,
						  const float* ring_t, const float* ring_s, float* out_x, float* out_t, float* out_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif
