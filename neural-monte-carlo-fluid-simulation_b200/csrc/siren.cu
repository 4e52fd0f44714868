// csrc/siren.cu -- fused SIREN neural-field kernels (fp32): forward, backward (parameter and input
// gradients) and Adam, one launch each.
//
// Replaces, for the velocity network of the time-stepper, the chain of stock PyTorch kernels behind
//   MLP / Sine            src/2d/models/networks.py:15-68 (nn.Sequential of Linear -> sin(30 .) ...)
//   update_network        src/2d/models/base.py:83-96  (backward + Adam.step)
//   divergence (autograd) src/2d/utils/diff_ops.py:45-51 (needs d(out)/d(coords): grad_x below)
// The network is  x[in] -> (Linear H, sin(w0 .)) -> L x (Linear HxH, sin(w0 .)) -> Linear out ; w0 = 30.
// One thread owns one sample; its activation vector lives in registers, the layer's weights are staged
// in shared memory and read as broadcast float4s, so a layer costs H*H FMAs + H*H/4 LDS.128 per sample.
// The backward pass recomputes activations from the saved pre-activations z_l (written by the training
// forward, layout [layer][neuron][sample], coalesced) and emits the per-layer deltas dZ_l and activations A_l;
// the weight gradients dW_l = dZ_l A_{l-1}^T are batch-dimension GEMMs issued by the host layer (siren.py).
// This is the exact-fp32 path (parity with torch within summation-order error); the tensor-core path for
// large inference batches is csrc/siren_tc.cu.
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>
#include "../../include/nmcfs_siren.h"
#include "siren_env.cuh"
#include "pdl.cuh"

namespace {

// sin/cos of the SIREN pre-activation w0*z.  |w0 z| stays below a few hundred, so one explicit reduction to
// [-pi, pi] (t = x/2pi - rint(x/2pi), exact subtraction) followed by the SFU sine/cosine is accurate to
// ~5e-7 absolute -- the same size as the fp32 rounding of the argument itself (ulp(100) = 7.6e-6) -- and costs
// 4 instructions instead of the ~40 of sinf's generic range reduction.  -DNMC_SIREN_LIBM_SIN restores sinf/cosf.
__device__ __forceinline__ float sinReduced(float x) {
#ifdef NMC_SIREN_LIBM_SIN
	return sinf(x);
#else
	float t = x*0.15915494309189535f;
	t -= rintf(t);
	return __sinf(6.283185307179586f*t);
#endif
}

constexpr int kTile = 128;     // samples per CTA tile == threads per CTA
constexpr int kMaxLayers = 18; // first + hidden + last

struct Params {
	const float* W[kMaxLayers];
	const float* b[kMaxLayers];
	float* gW[kMaxLayers];
	float* gb[kMaxLayers];
};

using nmc_siren_detail::Env;


template <int H>
__global__ void __launch_bounds__(kTile)
sirenForward(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
			 float* __restrict__ y, float* __restrict__ zSaved) {
	extern __shared__ float smem[];
	float* Wt = smem;                 // [H][H + 4]  transposed weights of the current hidden layer: Wt[k][n]
	float* act = smem + H*(H + 4);    // [H][kTile]  activation exchange (column per thread)
	const int tid = threadIdx.x;
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x) {
		const long long s = tile*kTile + tid;
		const bool live = s < n;
		float a[H];
		float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
		if (live) { x0 = x[s*inDim]; if (inDim > 1) x1 = x[s*inDim + 1]; if (inDim > 2) x2 = x[s*inDim + 2]; }
		{ // first layer: in -> H
#pragma unroll
			for (int j = 0; j < H; j++) {
				const float* w = &P.W[0][j*inDim];
				float z = __ldg(&P.b[0][j]) + __ldg(w)*x0;
				if (inDim > 1) z += __ldg(w + 1)*x1;
				if (inDim > 2) z += __ldg(w + 2)*x2;
				if (zSaved && live) zSaved[(size_t)j*n + s] = z;
				a[j] = sinReduced(w0*z);
			}
		}
		for (int l = 1; l <= nHidden; l++) {
			__syncthreads();
			for (int i = tid; i < H*H; i += kTile) { int nn = i/H, k = i - nn*H; Wt[k*(H + 4) + nn] = __ldg(&P.W[l][i]); }
			__syncthreads();
#pragma unroll 1
			for (int n0 = 0; n0 < H; n0 += 16) {
				float acc[16];
#pragma unroll
				for (int j = 0; j < 16; j++) acc[j] = __ldg(&P.b[l][n0 + j]);
#pragma unroll
				for (int k = 0; k < H; k++) {
					const float4* w = reinterpret_cast<const float4*>(&Wt[k*(H + 4) + n0]);
#pragma unroll
					for (int q = 0; q < 4; q++) {
						float4 v = w[q];
						acc[4*q + 0] += a[k]*v.x; acc[4*q + 1] += a[k]*v.y; acc[4*q + 2] += a[k]*v.z; acc[4*q + 3] += a[k]*v.w;
					}
				}
#pragma unroll
				for (int j = 0; j < 16; j++) {
					if (zSaved && live) zSaved[((size_t)l*H + n0 + j)*n + s] = acc[j];
					act[(n0 + j)*kTile + tid] = sinReduced(w0*acc[j]);
				}
			}
#pragma unroll
			for (int k = 0; k < H; k++) a[k] = act[k*kTile + tid];
		}
		// last layer: H -> out (no activation, outermost_linear=True)
		const int last = nHidden + 1;
		float yo[3] = {0.0f, 0.0f, 0.0f};
		for (int j = 0; j < outDim; j++) {
			float z = __ldg(&P.b[last][j]);
#pragma unroll
			for (int k = 0; k < H; k++) z += __ldg(&P.W[last][j*H + k])*a[k];
			if (j == 0) yo[0] = z; else if (j == 1) yo[1] = z; else yo[2] = z;
		}
		if (env.active) { const float xs[3] = {x0, x1, x2}; nmc_siren_detail::envForward(env, inDim, outDim, xs, yo); }
		if (live) {
			y[s*outDim] = yo[0];
			if (outDim > 1) y[s*outDim + 1] = yo[1];
			if (outDim > 2) y[s*outDim + 2] = yo[2];
		}
	}
}


// Backward, stage 1 ("delta chain"): per sample, back-propagates dL/dy through the layers and writes
//   dZ[l][j][s] = dL/dz_l  (l = 0..L)   and   A[l][j][s] = sin(w0 z_l)  (l = 0..L)
// coalesced, plus dL/dx.  The weight gradients are then plain GEMMs over the batch dimension,
//   dW_l = dZ_l A_{l-1}^T  (K = n),  db_l = rowsum(dZ_l),
// which the host layer issues as ONE batched cuBLAS call for the hidden layers (a library GEMM is the right tool
// for a plain GEMM; a 128-sample tile per CTA cannot fill 148 SMs at the fit loops' batch sizes of 4096-16384).
template <int H>
__global__ void __launch_bounds__(kTile)
sirenBackwardChain(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
				   const float* __restrict__ zSaved, const float* __restrict__ gy, float* __restrict__ gx,
				   float* __restrict__ dZ, float* __restrict__ A) {
	extern __shared__ float smem[];
	constexpr int LD = H + 4;
	float* Ws = smem;                  // [H][LD]  W_l row-major (n, k)
	float* ex = Ws + H*LD;             // [H][kTile] exchange buffer (column per thread)
	const int tid = threadIdx.x;
	const int last = nHidden + 1;
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x) {
		const long long s = tile*kTile + tid;
		const bool live = s < n;
		float g[H];
		float gxe[3] = {0.0f, 0.0f, 0.0f}; // gradient reaching x through the obstacle weight of the envelope
		{ // last layer: g_L = W_last^T gy  (and A_L for dW_last)
			float gy0 = 0.0f, gy1 = 0.0f, gy2 = 0.0f;
			if (live) {
				gy0 = gy[s*outDim]; if (outDim > 1) gy1 = gy[s*outDim + 1]; if (outDim > 2) gy2 = gy[s*outDim + 2];
				if (env.active) {
					const float xs[3] = {x[s*inDim], inDim > 1 ? x[s*inDim + 1] : 0.0f, inDim > 2 ? x[s*inDim + 2] : 0.0f};
					float gys[3] = {gy0, gy1, gy2}, yn[3] = {0.0f, 0.0f, 0.0f};
					const bool viaObstacle = env.sphere && gx != nullptr; // the obstacle weight is not detached (base.py:352-358)
					if (viaObstacle) { // network output y = W_last sin(w0 z_L) + b, needed for d(weight)/dx * y
						for (int j = 0; j < outDim; j++) yn[j] = __ldg(&P.b[last][j]);
#pragma unroll 4
						for (int k = 0; k < H; k++) {
							float t = w0*zSaved[((size_t)nHidden*H + k)*n + s]*0.15915494309189535f;
							t -= rintf(t);
							float ak = __sinf(6.283185307179586f*t);
							for (int j = 0; j < outDim; j++) yn[j] += __ldg(&P.W[last][j*H + k])*ak;
						}
					}
					nmc_siren_detail::envBackward(env, inDim, outDim, xs, yn, gys, viaObstacle ? gxe : nullptr);
					gy0 = gys[0]; gy1 = gys[1]; gy2 = gys[2];
				}
				// rows (L+1)*H .. of dZ: the (envelope-scaled) output gradient, for dW_last = gy'^T A_L^T
				const size_t r0 = (size_t)(nHidden + 1)*H;
				if (dZ) {
					dZ[(r0 + 0)*n + s] = gy0;
					if (outDim > 1) dZ[(r0 + 1)*n + s] = gy1;
					if (outDim > 2) dZ[(r0 + 2)*n + s] = gy2;
				}
			}
#pragma unroll
			for (int k = 0; k < H; k++) {
				float acc = __ldg(&P.W[last][k])*gy0;
				if (outDim > 1) acc += __ldg(&P.W[last][H + k])*gy1;
				if (outDim > 2) acc += __ldg(&P.W[last][2*H + k])*gy2;
				g[k] = acc;
			}
		}
		for (int l = nHidden; l >= 0; l--) {
			// dz_l = g * w0 cos(w0 z_l);  A_l = sin(w0 z_l)
#pragma unroll
			for (int j = 0; j < H; j++) {
				float zl = live ? zSaved[((size_t)l*H + j)*n + s] : 0.0f;
				float t = w0*zl*0.15915494309189535f;
				t -= rintf(t);
				float sn, cs;
				__sincosf(6.283185307179586f*t, &sn, &cs);
				g[j] = g[j]*w0*cs;
				if (live && dZ) { dZ[((size_t)l*H + j)*n + s] = g[j]; A[((size_t)l*H + j)*n + s] = sn; }
			}
			if (l == 0) break;
			__syncthreads();
			for (int i = tid; i < H*H; i += kTile) { int nn = i/H, k = i - nn*H; Ws[nn*LD + k] = __ldg(&P.W[l][i]); }
			__syncthreads();
			// g_{l-1}[k] = sum_nn W_l[nn][k] dz[nn]
#pragma unroll 1
			for (int k0 = 0; k0 < H; k0 += 16) {
				float acc[16];
#pragma unroll
				for (int j = 0; j < 16; j++) acc[j] = 0.0f;
#pragma unroll
				for (int nn = 0; nn < H; nn++) {
					const float4* w = reinterpret_cast<const float4*>(&Ws[nn*LD + k0]);
#pragma unroll
					for (int q = 0; q < 4; q++) {
						float4 v = w[q];
						acc[4*q + 0] += g[nn]*v.x; acc[4*q + 1] += g[nn]*v.y; acc[4*q + 2] += g[nn]*v.z; acc[4*q + 3] += g[nn]*v.w;
					}
				}
#pragma unroll
				for (int j = 0; j < 16; j++) ex[(k0 + j)*kTile + tid] = acc[j];
			}
#pragma unroll
			for (int k = 0; k < H; k++) g[k] = ex[k*kTile + tid];
		}
		if (gx && live) { // dL/dx = W_0^T dz_0
			float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
			for (int j = 0; j < H; j++) {
				const float* w = &P.W[0][j*inDim];
				a0 += __ldg(w)*g[j];
				if (inDim > 1) a1 += __ldg(w + 1)*g[j];
				if (inDim > 2) a2 += __ldg(w + 2)*g[j];
			}
			gx[s*inDim] = a0 + gxe[0]; if (inDim > 1) gx[s*inDim + 1] = a1 + gxe[1]; if (inDim > 2) gx[s*inDim + 2] = a2 + gxe[2];
		}
	}
}

// ---- small-batch variants: one sample's layer split over H/16 threads -------------------------------------------
// The fit loops run on 4096-16384 samples.  With a thread per sample that is 32-128 CTAs of 4 warps: most of the
// 148 SMs idle and nothing hides latency.  Here a CTA takes 32 samples (the lanes of a warp) and warp g computes
// neurons [16 g, 16 g + 16) of every layer for them, so a batch of 4096 is 128 CTAs of H/16 warps and the weights
// are read from shared memory as warp-wide broadcasts.  Activations (forward) / deltas (backward) of a layer are
// exchanged through a double-buffered [H][32] array; the last layer and dL/dx are reduced across the warps.
constexpr int kTS = 32;

template <int H, int NPT>
__global__ void __launch_bounds__(32*(H/NPT), H == 128 ? 2 : 4)
sirenForwardSplit(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
				  float* __restrict__ y, float* __restrict__ zSaved) {
	nmc_pdl::gridEnter();
	extern __shared__ float smem[];
	constexpr int LD = H + 4, NG = H/NPT, NT = 32*NG;
	// H = 64: two weight buffers, the next layer's weights stream in with cp.async while this layer is computed
	// (a layer is ~2.5 us, of which the L2 -> shared copy was more than a third); H = 128 keeps one buffer (occupancy).
	constexpr int NB = H == 64 ? 2 : 1;
	float* Wt = smem;                   // [NB][H][LD]  Wt[k][nn] = W_l[nn][k]
	float* act = Wt + NB*H*LD;          // [2][H][kTS]
	float* part = act + 2*H*kTS;        // [H/NPT][3][kTS]
	const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5, n0 = NPT*g;
	const int last = nHidden + 1;
	for (long long tile = blockIdx.x; tile*kTS < n; tile += gridDim.x) {
		const long long s = tile*kTS + lane;
		const bool live = s < n;
		float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
		if (live) { x0 = x[s*inDim]; if (inDim > 1) x1 = x[s*inDim + 1]; if (inDim > 2) x2 = x[s*inDim + 2]; }
		auto stageAsync = [&](int l, float* dst) { // transposing copy (8 (k) x 4 (nn) patches per warp), asynchronous
			for (int idx = tid; idx < H*H; idx += NT) {
				const int b = idx >> 5, kk = idx & 7, nq = (idx >> 3) & 3;
				const int k = (b % (H/8))*8 + kk, nn = (b/(H/8))*4 + nq;
				const unsigned d = (unsigned)__cvta_generic_to_shared(&dst[k*LD + nn]);
				asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(&P.W[l][nn*H + k]) : "memory");
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		};
		if (NB == 2 && nHidden >= 1) stageAsync(1, Wt); // nobody reads buffer 0 any more: every warp passed the barrier before the output
		float a16[NPT];
#pragma unroll
		for (int j = 0; j < NPT; j++) { // first layer: in -> H
			const float* w = &P.W[0][(n0 + j)*inDim];
			float z = __ldg(&P.b[0][n0 + j]) + __ldg(w)*x0;
			if (inDim > 1) z += __ldg(w + 1)*x1;
			if (inDim > 2) z += __ldg(w + 2)*x2;
			if (zSaved && live) zSaved[(size_t)(n0 + j)*n + s] = z;
			a16[j] = sinReduced(w0*z);
			act[(n0 + j)*kTS + lane] = a16[j];
		}
		int cur = 0;
		for (int l = 1; l <= nHidden; l++) {
			const float* Wl = Wt;
			if (NB == 2) {
				asm volatile("cp.async.wait_group 0;" ::: "memory");
				__syncthreads(); // this layer's weights have landed (all threads' copies), act[cur] is complete
				Wl = Wt + ((l - 1) & 1)*H*LD;
				if (l < nHidden) stageAsync(l + 1, Wt + (l & 1)*H*LD); // the other buffer was last read by layer l - 1
			} else {
				__syncthreads(); // act[cur] is complete, nobody reads Wt any more
				// transposing copy, bank-conflict free: a warp writes an 8 (k) x 4 (nn) patch = banks 4 kk + nq
				for (int idx = tid; idx < H*H; idx += NT) {
					const int b = idx >> 5, kk = idx & 7, nq = (idx >> 3) & 3;
					const int k = (b % (H/8))*8 + kk, nn = (b/(H/8))*4 + nq;
					Wt[k*LD + nn] = __ldg(&P.W[l][nn*H + k]);
				}
				__syncthreads();
			}
			float acc[NPT];
#pragma unroll
			for (int j = 0; j < NPT; j++) acc[j] = __ldg(&P.b[l][n0 + j]);
			const float* ac = act + cur*H*kTS + lane;
#pragma unroll 8
			for (int k = 0; k < H; k++) {
				const float av = ac[k*kTS];
				const float4* w = reinterpret_cast<const float4*>(&Wl[k*LD + n0]);
#pragma unroll
				for (int q = 0; q < NPT/4; q++) {
					const float4 v = w[q];
					acc[4*q + 0] += av*v.x; acc[4*q + 1] += av*v.y; acc[4*q + 2] += av*v.z; acc[4*q + 3] += av*v.w;
				}
			}
			cur ^= 1;
			float* an = act + cur*H*kTS + lane;
#pragma unroll
			for (int j = 0; j < NPT; j++) {
				if (zSaved && live) zSaved[((size_t)l*H + n0 + j)*n + s] = acc[j];
				a16[j] = sinReduced(w0*acc[j]);
				an[(n0 + j)*kTS] = a16[j];
			}
		}
		// last layer: H -> out, partial dot products over this warp's 16 neurons
		for (int j = 0; j < outDim; j++) {
			float pj = 0.0f;
#pragma unroll
			for (int i = 0; i < NPT; i++) pj += __ldg(&P.W[last][j*H + n0 + i])*a16[i];
			part[(g*3 + j)*kTS + lane] = pj;
		}
		__syncthreads();
		if (g == 0) {
			float yo[3] = {0.0f, 0.0f, 0.0f};
			for (int j = 0; j < outDim; j++) {
				float z = __ldg(&P.b[last][j]);
				for (int q = 0; q < NG; q++) z += part[(q*3 + j)*kTS + lane];
				if (j == 0) yo[0] = z; else if (j == 1) yo[1] = z; else yo[2] = z;
			}
			if (env.active) { const float xs[3] = {x0, x1, x2}; nmc_siren_detail::envForward(env, inDim, outDim, xs, yo); }
			if (live) {
				y[s*outDim] = yo[0];
				if (outDim > 1) y[s*outDim + 1] = yo[1];
				if (outDim > 2) y[s*outDim + 2] = yo[2];
			}
		}
		__syncthreads(); // `part` (and, without hidden layers, act[0]) is rewritten by the next tile
	}
}

template <int H, int NPT>
__global__ void __launch_bounds__(32*(H/NPT), H == 128 ? 2 : 4)
sirenBackwardSplit(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
				   const float* __restrict__ zSaved, const float* __restrict__ gy, float* __restrict__ gx,
				   float* __restrict__ dZ, float* __restrict__ A) {
	extern __shared__ float smem[];
	constexpr int LD = H + 4, NG = H/NPT, NT = 32*NG;
	constexpr int NB = H == 64 ? 2 : 1; // H = 64: the next layer's weights stream in with cp.async (see sirenForwardSplit)
	float* Ws = smem;                   // [NB][H][LD]  W_l row-major (nn, k)
	float* ex = Ws + NB*H*LD;           // [2][H][kTS] delta exchange
	float* part = ex + 2*H*kTS;         // [H/NPT][3][kTS]
	const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5, n0 = NPT*g;
	const int last = nHidden + 1;
	for (long long tile = blockIdx.x; tile*kTS < n; tile += gridDim.x) {
		const long long s = tile*kTS + lane;
		const bool live = s < n;
		float gy0 = 0.0f, gy1 = 0.0f, gy2 = 0.0f;
		float gxe[3] = {0.0f, 0.0f, 0.0f}; // gradient reaching x through the obstacle weight of the envelope
		if (live) {
			gy0 = gy[s*outDim]; if (outDim > 1) gy1 = gy[s*outDim + 1]; if (outDim > 2) gy2 = gy[s*outDim + 2];
			if (env.active) {
				const float xs[3] = {x[s*inDim], inDim > 1 ? x[s*inDim + 1] : 0.0f, inDim > 2 ? x[s*inDim + 2] : 0.0f};
				float gys[3] = {gy0, gy1, gy2}, yn[3] = {0.0f, 0.0f, 0.0f};
				const bool viaObstacle = env.sphere && gx != nullptr;
				if (viaObstacle) { // network output, recomputed by every warp (only the divergence grid takes this path)
					for (int j = 0; j < outDim; j++) yn[j] = __ldg(&P.b[last][j]);
#pragma unroll 4
					for (int k = 0; k < H; k++) {
						const float ak = sinReduced(w0*zSaved[((size_t)nHidden*H + k)*n + s]);
						for (int j = 0; j < outDim; j++) yn[j] += __ldg(&P.W[last][j*H + k])*ak;
					}
				}
				nmc_siren_detail::envBackward(env, inDim, outDim, xs, yn, gys, viaObstacle ? gxe : nullptr);
				gy0 = gys[0]; gy1 = gys[1]; gy2 = gys[2];
			}
			if (g == 0) { // rows (L+1)*H .. of dZ: the (envelope-scaled) output gradient, for dW_last = gy'^T A_L^T
				const size_t r0 = (size_t)(nHidden + 1)*H;
				if (dZ) {
					dZ[(r0 + 0)*n + s] = gy0;
					if (outDim > 1) dZ[(r0 + 1)*n + s] = gy1;
					if (outDim > 2) dZ[(r0 + 2)*n + s] = gy2;
				}
			}
		}
		auto stageAsync = [&](int l, float* dst) {
			for (int idx = tid; idx < H*H/4; idx += NT) {
				const int nn = idx/(H/4), k4 = idx - nn*(H/4);
				const unsigned d = (unsigned)__cvta_generic_to_shared(&dst[nn*LD + 4*k4]);
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(&P.W[l][nn*H + 4*k4]) : "memory");
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		};
		if (NB == 2 && nHidden >= 1) stageAsync(nHidden, Ws); // layer index l uses buffer (nHidden - l) & 1; buffer 0 is free (barrier at the tile's end)
		float g16[NPT];
#pragma unroll
		for (int i = 0; i < NPT; i++) { // g_L = W_last^T gy'
			float acc = __ldg(&P.W[last][n0 + i])*gy0;
			if (outDim > 1) acc += __ldg(&P.W[last][H + n0 + i])*gy1;
			if (outDim > 2) acc += __ldg(&P.W[last][2*H + n0 + i])*gy2;
			g16[i] = acc;
		}
		int cur = 0;
		for (int l = nHidden; l >= 0; l--) {
			float* exw = ex + cur*H*kTS + lane;
#pragma unroll
			for (int i = 0; i < NPT; i++) { // dz_l = g * w0 cos(w0 z_l);  A_l = sin(w0 z_l)
				const float zl = live ? zSaved[((size_t)l*H + n0 + i)*n + s] : 0.0f;
				float t = w0*zl*0.15915494309189535f;
				t -= rintf(t);
				float sn, cs;
				__sincosf(6.283185307179586f*t, &sn, &cs);
				g16[i] = g16[i]*w0*cs;
				if (live && dZ) { dZ[((size_t)l*H + n0 + i)*n + s] = g16[i]; A[((size_t)l*H + n0 + i)*n + s] = sn; }
				exw[(n0 + i)*kTS] = g16[i];
			}
			if (l == 0) break;
			const float* Wl = Ws;
			if (NB == 2) {
				asm volatile("cp.async.wait_group 0;" ::: "memory");
				__syncthreads(); // W_l has landed, ex[cur] complete
				Wl = Ws + ((nHidden - l) & 1)*H*LD;
				if (l > 1) stageAsync(l - 1, Ws + ((nHidden - l + 1) & 1)*H*LD); // the other buffer was last read by layer l + 1
			} else {
				__syncthreads(); // ex[cur] complete, nobody reads Ws any more
				for (int idx = tid; idx < H*H/4; idx += NT) {
					const int nn = idx/(H/4), k4 = idx - nn*(H/4);
					*reinterpret_cast<float4*>(&Ws[nn*LD + 4*k4]) = __ldg(reinterpret_cast<const float4*>(&P.W[l][nn*H + 4*k4]));
				}
				__syncthreads();
			}
			float acc[NPT];
#pragma unroll
			for (int i = 0; i < NPT; i++) acc[i] = 0.0f;
			const float* er = ex + cur*H*kTS + lane;
#pragma unroll 8
			for (int nn = 0; nn < H; nn++) { // g_{l-1}[k] = sum_nn W_l[nn][k] dz[nn]
				const float dv = er[nn*kTS];
				const float4* w = reinterpret_cast<const float4*>(&Wl[nn*LD + n0]);
#pragma unroll
				for (int q = 0; q < NPT/4; q++) {
					const float4 v = w[q];
					acc[4*q + 0] += dv*v.x; acc[4*q + 1] += dv*v.y; acc[4*q + 2] += dv*v.z; acc[4*q + 3] += dv*v.w;
				}
			}
#pragma unroll
			for (int i = 0; i < NPT; i++) g16[i] = acc[i];
			cur ^= 1;
		}
		if (gx) { // dL/dx = W_0^T dz_0: partial sums over this warp's 16 neurons, reduced by warp 0
			float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
			for (int i = 0; i < NPT; i++) {
				const float* w = &P.W[0][(n0 + i)*inDim];
				a0 += __ldg(w)*g16[i];
				if (inDim > 1) a1 += __ldg(w + 1)*g16[i];
				if (inDim > 2) a2 += __ldg(w + 2)*g16[i];
			}
			__syncthreads(); // `part` of the previous tile has been consumed
			part[(g*3 + 0)*kTS + lane] = a0; part[(g*3 + 1)*kTS + lane] = a1; part[(g*3 + 2)*kTS + lane] = a2;
			__syncthreads();
			if (g == 0 && live) {
				float r[3] = {gxe[0], gxe[1], gxe[2]};
				for (int q = 0; q < NG; q++) { r[0] += part[(q*3 + 0)*kTS + lane]; r[1] += part[(q*3 + 1)*kTS + lane]; r[2] += part[(q*3 + 2)*kTS + lane]; }
				gx[s*inDim] = r[0]; if (inDim > 1) gx[s*inDim + 1] = r[1]; if (inDim > 2) gx[s*inDim + 2] = r[2];
			}
		}
		__syncthreads(); // the next tile reuses ex[0]
	}
}

// Backward, stage 2: all weight and bias gradients in ONE launch.  Every gradient is C[i][j] = sum_s P[i][s] Q[j][s]
// over the batch (P = dZ_l or the scaled output gradient, Q = A_{l-1} or x^T), so the grid is (K-splits, layers):
// a CTA stages 32-sample tiles of its two operands in shared memory, accumulates a 4x4 register tile per thread
// (hidden layers) or a strided set of scalars (first / last layer), and adds its partial result with atomics.
constexpr int kGT = 256;   // threads
constexpr int kGS = 32;    // samples per staged tile
constexpr int kGChunkMax = 512; // samples per CTA at most; small batches use smaller chunks to fill the SMs
template <int H>
__global__ void __launch_bounds__(kGT)
sirenWeightGrad(Params P, int inDim, int outDim, int nHidden, const float* __restrict__ x, long long n,
				const float* __restrict__ dZ, const float* __restrict__ A, int chunk) {
	// staged tiles are sample-major, [sample][row], so that a thread's four rows are one 16-byte load and the lanes of a
	// warp read consecutive 16-byte words (the row-major layout cost four 4-way-conflicted scalar loads instead)
	constexpr int LDS = H + 4;
	__shared__ __align__(16) float Ps[kGS][LDS];
	__shared__ __align__(16) float Qs[kGS][LDS];
	const int tid = threadIdx.x;
	const int l = blockIdx.y;                 // 0 .. nHidden + 1
	const int last = nHidden + 1;
	const long long s0 = (long long)blockIdx.x*chunk;
	const long long s1 = s0 + chunk < n ? s0 + chunk : n;
	const int RP = l == last ? outDim : H;    // rows of P
	const int RQ = l == 0 ? inDim : H;        // rows of Q
	const float* Pg = l == last ? dZ + (size_t)(nHidden + 1)*H*n : dZ + (size_t)l*H*n;
	const float* Qg = l == 0 ? nullptr : A + (size_t)(l - 1)*H*n;
	constexpr int NP = (H*H)/(16*kGT);        // 4x4 tiles per thread: 1 (H = 64) or 4 (H = 128)
	constexpr int RB = (kGT/(H/4))*4;          // rows covered by one pass
	float acc[NP][4][4];
#pragma unroll
	for (int q = 0; q < NP; q++)
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 4; j++) acc[q][i][j] = 0.0f;
	float small[(3*H + kGT - 1)/kGT];         // first/last layer: <= 3*H outputs
#pragma unroll
	for (int q = 0; q < (3*H + kGT - 1)/kGT; q++) small[q] = 0.0f;
	float bsum = 0.0f;
	const bool hidden = l >= 1 && l <= nHidden;
	const int ti = (tid/(H/4))*4, tj = (tid%(H/4))*4; // 4x4 tile origin (H = 64: 16x16 threads; H = 128: uses two passes)
	for (long long sb = s0; sb < s1; sb += kGS) {
		const int ns = (int)(s1 - sb < kGS ? s1 - sb : kGS);
		__syncthreads();
		for (int idx = tid; idx < RP*kGS; idx += kGT) { int r = idx/kGS, c = idx - r*kGS; Ps[c][r] = c < ns ? Pg[(size_t)r*n + sb + c] : 0.0f; }
		if (l == 0) { for (int idx = tid; idx < kGS*inDim; idx += kGT) { int c = idx/inDim, r = idx - c*inDim; Qs[c][r] = c < ns ? x[(sb + c)*inDim + r] : 0.0f; } }
		else { for (int idx = tid; idx < RQ*kGS; idx += kGT) { int r = idx/kGS, c = idx - r*kGS; Qs[c][r] = c < ns ? Qg[(size_t)r*n + sb + c] : 0.0f; } }
		__syncthreads();
		if (hidden) {
#pragma unroll 4
			for (int c = 0; c < kGS; c++) {
				const float4 qv = *reinterpret_cast<const float4*>(&Qs[c][tj]);
				const float q0 = qv.x, q1 = qv.y, q2 = qv.z, q3 = qv.w;
#pragma unroll
				for (int ps = 0; ps < NP; ps++) {
					const int i0 = ti + ps*RB;
					const float4 pv = *reinterpret_cast<const float4*>(&Ps[c][i0]);
					const float p0 = pv.x, p1 = pv.y, p2 = pv.z, p3 = pv.w;
					acc[ps][0][0] += p0*q0; acc[ps][0][1] += p0*q1; acc[ps][0][2] += p0*q2; acc[ps][0][3] += p0*q3;
					acc[ps][1][0] += p1*q0; acc[ps][1][1] += p1*q1; acc[ps][1][2] += p1*q2; acc[ps][1][3] += p1*q3;
					acc[ps][2][0] += p2*q0; acc[ps][2][1] += p2*q1; acc[ps][2][2] += p2*q2; acc[ps][2][3] += p2*q3;
					acc[ps][3][0] += p3*q0; acc[ps][3][1] += p3*q1; acc[ps][3][2] += p3*q2; acc[ps][3][3] += p3*q3;
				}
			}
		} else {
			int q = 0;
			for (int o = tid; o < RP*RQ; o += kGT, q++) {
				int i = o/RQ, j = o - i*RQ;
				float a = 0.0f;
				for (int c = 0; c < kGS; c++) a += Ps[c][i]*Qs[c][j];
				small[q] += a;
			}
		}
		if (tid < RP) { float a = 0.0f; for (int c = 0; c < kGS; c++) a += Ps[c][tid]; bsum += a; }
	}
	if (hidden) {
#pragma unroll
		for (int ps = 0; ps < NP; ps++)
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) atomicAdd(&P.gW[l][(ti + ps*RB + i)*H + tj + j], acc[ps][i][j]);
	} else {
		int q = 0;
		for (int o = tid; o < RP*RQ; o += kGT, q++) atomicAdd(&P.gW[l][o], small[q]);
	}
	if (tid < RP) atomicAdd(&P.gb[l][tid], bsum);
}

// MSE loss of a fit iteration in one launch: diff = y - target, grad_y = diff * 2/count, loss = mean(diff^2).
// A few CTAs (one SM moves ~100 bytes per clock: a single CTA took 10 us for 49152 floats), 16-byte accesses; every CTA
// leaves its partial sum in a scratch slot, the last one to finish adds the slots in index order (deterministic) and
// writes the loss, so nothing needs a zero-fill.  The scratch is per device context: calls are stream-ordered by the
// callers (one fit at a time), not re-entrant across streams.
constexpr int kMseBlocks = 32;
__device__ float g_msePart[kMseBlocks];
__device__ unsigned g_mseDone;
// Fit loops (nmc_mse_grad_fit): the target is target - sub, the flat gradient buffer `zero` is cleared and Adam's device-side
// step counter is advanced in the same launch (three 2 us launches less per iteration).
__global__ void __launch_bounds__(512) mseGrad(const float* __restrict__ y, const float* __restrict__ target, const float* __restrict__ sub,
												long long count, float* __restrict__ diff, float* __restrict__ gy, float* __restrict__ loss, int vec,
												float* __restrict__ zero, long long zeroCount, long long* __restrict__ stepAdvance,
												float stopThreshold, int* __restrict__ stopFlag) {
	nmc_pdl::gridEnter();
	const float scale = 2.0f/(float)count;
	const long long tid = (long long)blockIdx.x*blockDim.x + threadIdx.x, nth = (long long)gridDim.x*blockDim.x;
	if (stepAdvance && tid == 0) *stepAdvance += 1;
	if (zero) {
		float4* z4 = reinterpret_cast<float4*>(zero);   // 16-byte aligned (checked by the host)
		for (long long i = tid; i < (zeroCount >> 2); i += nth) z4[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		for (long long i = (zeroCount & ~3ll) + tid; i < zeroCount; i += nth) zero[i] = 0.0f;
	}
	float acc = 0.0f;
	long long done = 0;
	if (vec) {
		const long long c4 = count >> 2;
		const float4* y4 = reinterpret_cast<const float4*>(y); const float4* t4 = reinterpret_cast<const float4*>(target);
		const float4* s4 = reinterpret_cast<const float4*>(sub);
		float4* d4 = reinterpret_cast<float4*>(diff); float4* g4 = reinterpret_cast<float4*>(gy);
#pragma unroll 2
		for (long long i = tid; i < c4; i += nth) {
			const float4 a = y4[i];
			float4 b = t4[i];
			if (sub) { const float4 c = s4[i]; b = make_float4(b.x - c.x, b.y - c.y, b.z - c.z, b.w - c.w); }
			const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
			d4[i] = d; g4[i] = make_float4(d.x*scale, d.y*scale, d.z*scale, d.w*scale);
			acc += (d.x*d.x + d.y*d.y) + (d.z*d.z + d.w*d.w);
		}
		done = c4 << 2;
	}
	for (long long i = done + tid; i < count; i += nth) {
		const float d = y[i] - (sub ? target[i] - sub[i] : target[i]);
		diff[i] = d; gy[i] = d*scale;
		acc += d*d;
	}
	for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
	__shared__ float part[32];
	__shared__ bool isLast;
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x < 32) {
		float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0f;
		for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
		if (threadIdx.x == 0) {
			g_msePart[blockIdx.x] = v;
			__threadfence();
			isLast = atomicAdd(&g_mseDone, 1u) == gridDim.x - 1;
		}
	}
	__syncthreads();
	if (isLast && threadIdx.x == 0) {
		__threadfence();
		float v = 0.0f;
		for (unsigned b = 0; b < gridDim.x; b++) v += *(volatile float*)&g_msePart[b];
		*loss = v/(float)count;
		if (stopFlag && v/(float)count <= stopThreshold) *stopFlag = 1;   // sticky until the next fit: the Adam update checks it
		g_mseDone = 0;
	}
}

// Adam with the step counter in device memory: a CUDA graph that replays one fit iteration must not freeze the bias
// corrections at their capture-time values, so the counter is advanced by a one-thread kernel and read by the update.
__global__ void adamAdvance(long long* step) { *step += 1; }
__global__ void adamKernelDev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
							  long long n, float lr, float b1, float b2, float eps, const long long* __restrict__ step, const int* __restrict__ stopFlag) {
	nmc_pdl::gridEnter();
	if (stopFlag && *stopFlag) return;   // the fit has reached the early-stop threshold (base.py:148): the reference has left the loop
	__shared__ float bc[2];
	if (threadIdx.x == 0) { // 1 - beta^t = -expm1(t log1p(-(1 - beta))): fp32 throughout (1 - beta is exact), ~2e-7 relative; the double-precision
		const float t = (float)*step;   // pow it replaces took ~1.5 of the kernel's 3.6 us on the fp64 pipe
		bc[0] = -expm1f(t*log1pf(-(1.0f - b1)));
		bc[1] = sqrtf(-expm1f(t*log1pf(-(1.0f - b2))));
	}
	__syncthreads();
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	float gi = g[i];
	float mi = b1*m[i] + (1.0f - b1)*gi;
	float vi = b2*v[i] + (1.0f - b2)*gi*gi;
	m[i] = mi; v[i] = vi;
	float denom = sqrtf(vi)/bc[1] + eps;
	p[i] -= (lr/bc[0])*(mi/denom);
}

// adamKernelDev of iteration i and the ring fetch of iteration i + 1 (fit_glue.cu: fitFetch) in one launch: the first `adamBlocks`
// CTAs update the parameters, the others copy slot (step % slots) of the target ring into the fixed buffers the next captured
// iteration reads.  Both read the step counter, nobody writes it (nmc_mse_grad_fit advances it earlier in the iteration), so the
// launch does what the two did back to back; one ~2 us launch less per iteration of an unrolled fit graph.
__global__ void adamFetchKernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
								long long n, float lr, float b1, float b2, float eps, const long long* __restrict__ step, const int* __restrict__ stopFlag,
								int adamBlocks, long long count, int slots, const float* __restrict__ ringX, const float* __restrict__ ringT,
								const float* __restrict__ ringS, float* __restrict__ outX, float* __restrict__ outT, float* __restrict__ outS, int vec) {
	nmc_pdl::gridEnter();
	if ((int)blockIdx.x >= adamBlocks) {
		const long long base = (*step % slots)*count;
		const long long tid = (long long)(blockIdx.x - adamBlocks)*blockDim.x + threadIdx.x, nth = (long long)(gridDim.x - adamBlocks)*blockDim.x;
		if (vec) {
			const float4* x4 = reinterpret_cast<const float4*>(ringX + base); const float4* t4 = reinterpret_cast<const float4*>(ringT + base);
			const float4* s4 = ringS ? reinterpret_cast<const float4*>(ringS + base) : nullptr;
			for (long long i = tid; i < (count >> 2); i += nth) {
				reinterpret_cast<float4*>(outX)[i] = x4[i]; reinterpret_cast<float4*>(outT)[i] = t4[i];
				if (s4) reinterpret_cast<float4*>(outS)[i] = s4[i];
			}
		} else {
			for (long long i = tid; i < count; i += nth) {
				outX[i] = ringX[base + i]; outT[i] = ringT[base + i];
				if (ringS) outS[i] = ringS[base + i];
			}
		}
		return;
	}
	if (stopFlag && *stopFlag) return;
	__shared__ float bc[2];
	if (threadIdx.x == 0) {
		const float t = (float)*step;
		bc[0] = -expm1f(t*log1pf(-(1.0f - b1)));
		bc[1] = sqrtf(-expm1f(t*log1pf(-(1.0f - b2))));
	}
	__syncthreads();
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	float gi = g[i];
	float mi = b1*m[i] + (1.0f - b1)*gi;
	float vi = b2*v[i] + (1.0f - b2)*gi*gi;
	m[i] = mi; v[i] = vi;
	float denom = sqrtf(vi)/bc[1] + eps;
	p[i] -= (lr/bc[0])*(mi/denom);
}

// torch.optim.Adam (no amsgrad, no weight decay) over one flat buffer; step is 1-based
__global__ void adamKernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
						   long long n, float lr, float b1, float b2, float eps, float bc1, float bc2sqrt) {
	long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	float gi = g[i];
	float mi = b1*m[i] + (1.0f - b1)*gi;
	float vi = b2*v[i] + (1.0f - b2)*gi*gi;
	m[i] = mi; v[i] = vi;
	float denom = sqrtf(vi)/bc2sqrt + eps;
	p[i] -= (lr/bc1)*(mi/denom);
}

thread_local const char* g_err = "";
int fail(const char* m) { g_err = m; return 1; }

int fill(Params& P, const nmc_siren_shape* sh, const float* const* W, const float* const* b, float* const* gW, float* const* gb) {
	if (!sh || !W || !b) return fail("null argument");
	if (sh->hidden != 64 && sh->hidden != 128) return fail("hidden_features must be 64 or 128");
	if (sh->in_dim < 1 || sh->in_dim > 3 || sh->out_dim < 1 || sh->out_dim > 3) return fail("in/out features must be 1..3");
	if (sh->n_hidden_layers < 0 || sh->n_hidden_layers + 2 > kMaxLayers) return fail("too many layers");
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) {
		P.W[l] = W[l]; P.b[l] = b[l];
		P.gW[l] = gW ? gW[l] : nullptr; P.gb[l] = gb ? gb[l] : nullptr;
		if (!W[l] || !b[l] || (gW && (!gW[l] || !gb[l]))) return fail("null layer pointer");
	}
	return 0;
}

// batches below this size run the split kernels (NMC_SIREN_SPLIT_MAX overrides; 0 disables them)
bool useSplit(long long n) {
	static long long limit = -1;
	if (limit < 0) { const char* e = getenv("NMC_SIREN_SPLIT_MAX"); limit = e ? atoll(e) : 65536; }
	return n <= limit;
}
constexpr int kNpt64 = 8, kNpt128 = 8; // neurons per thread: 8 warps per CTA for both widths
size_t splitSmem(int H) { return ((size_t)(H == 64 ? 2 : 1)*H*(H + 4) + (size_t)2*H*kTS + (size_t)(H/(H == 64 ? kNpt64 : kNpt128))*3*kTS)*sizeof(float); }

int smCount() {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms;
}

} // namespace

namespace nmc_siren_detail { void setError(const char* m) { g_err = m; } }

extern "C" const char* nmc_siren_last_error(void) { return g_err; }


extern "C" int nmc_siren_forward(const nmc_siren_shape* sh, const float* const* W, const float* const* b, const float* x,
								 int64_t n, float* y, float* z_saved, const nmc_siren_envelope* envp, void* stream) {
	Params P;
	if (fill(P, sh, W, b, nullptr, nullptr)) return 1;
	Env env;
	if (const char* bad = nmc_siren_detail::toEnv(envp, env)) return fail(bad);
	if (n <= 0) return 0;
	if (!x || !y) return fail("null buffer");
	const int H = sh->hidden;
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e;
	if (useSplit(n)) { // small batches: a sample's layer split over several threads (see sirenForwardSplit)
		size_t smemS = splitSmem(H);
		long long tilesS = (n + kTS - 1)/kTS;
		int gridS = (int)(tilesS < 8ll*smCount() ? tilesS : 8ll*smCount());
		if (H == 64) {
			e = cudaFuncSetAttribute(sirenForwardSplit<64, kNpt64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS);
			if (!e) e = nmc_pdl::launch(sirenForwardSplit<64, kNpt64>, dim3(gridS), dim3(32*(64/kNpt64)), smemS, st, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, y, z_saved);
		} else {
			e = cudaFuncSetAttribute(sirenForwardSplit<128, kNpt128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS);
			if (!e) e = nmc_pdl::launch(sirenForwardSplit<128, kNpt128>, dim3(gridS), dim3(32*(128/kNpt128)), smemS, st, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, y, z_saved);
		}
		if (!e) e = cudaGetLastError();
		return e ? fail(cudaGetErrorString(e)) : 0;
	}
	size_t smem = ((size_t)H*(H + 4) + (size_t)H*kTile)*sizeof(float);
	long long tiles = (n + kTile - 1)/kTile;
	int grid = (int)(tiles < 4ll*smCount() ? tiles : 4ll*smCount());
	if (H == 64) {
		e = cudaFuncSetAttribute(sirenForward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) sirenForward<64><<<grid, kTile, smem, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, y, z_saved);
	} else {
		e = cudaFuncSetAttribute(sirenForward<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) sirenForward<128><<<grid, kTile, smem, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, y, z_saved);
	}
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_siren_backward(const nmc_siren_shape* sh, const float* const* W, const float* const* b, const float* x,
								  int64_t n, const float* z_saved, const float* grad_y, float* dZ, float* A,
								  float* grad_x, const nmc_siren_envelope* envp, void* stream) {
	Params P;
	if (fill(P, sh, W, b, nullptr, nullptr)) return 1;
	Env env;
	if (const char* bad = nmc_siren_detail::toEnv(envp, env)) return fail(bad);
	if (n <= 0) return 0;
	if (!x || !z_saved || !grad_y || (!dZ) != (!A) || (!dZ && !grad_x)) return fail("null buffer");
	const int H = sh->hidden;
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e;
	if (useSplit(n)) {
		size_t smemS = splitSmem(H);
		long long tilesS = (n + kTS - 1)/kTS;
		int gridS = (int)(tilesS < 8ll*smCount() ? tilesS : 8ll*smCount());
		if (H == 64) {
			e = cudaFuncSetAttribute(sirenBackwardSplit<64, kNpt64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS);
			if (!e) sirenBackwardSplit<64, kNpt64><<<gridS, 32*(64/kNpt64), smemS, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, z_saved, grad_y, grad_x, dZ, A);
		} else {
			e = cudaFuncSetAttribute(sirenBackwardSplit<128, kNpt128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemS);
			if (!e) sirenBackwardSplit<128, kNpt128><<<gridS, 32*(128/kNpt128), smemS, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, z_saved, grad_y, grad_x, dZ, A);
		}
		if (!e) e = cudaGetLastError();
		return e ? fail(cudaGetErrorString(e)) : 0;
	}
	size_t smem = ((size_t)H*(H + 4) + (size_t)H*kTile)*sizeof(float);
	long long tiles = (n + kTile - 1)/kTile;
	int grid = (int)(tiles < 4ll*smCount() ? tiles : 4ll*smCount());
	if (H == 64) {
		e = cudaFuncSetAttribute(sirenBackwardChain<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) sirenBackwardChain<64><<<grid, kTile, smem, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, z_saved, grad_y, grad_x, dZ, A);
	} else {
		e = cudaFuncSetAttribute(sirenBackwardChain<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) sirenBackwardChain<128><<<grid, kTile, smem, st>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, z_saved, grad_y, grad_x, dZ, A);
	}
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_siren_weight_grads(const nmc_siren_shape* sh, const float* x, int64_t n, const float* dZ, const float* A,
									  float* const* gW, float* const* gb, void* stream) {
	if (!sh || !gW || !gb) return fail("null argument");
	if (sh->hidden != 64 && sh->hidden != 128) return fail("hidden_features must be 64 or 128");
	if (sh->n_hidden_layers < 0 || sh->n_hidden_layers + 2 > kMaxLayers) return fail("too many layers");
	if (n <= 0) return 0;
	if (!x || !dZ || !A) return fail("null buffer");
	Params P;
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) { P.W[l] = nullptr; P.b[l] = nullptr; P.gW[l] = gW[l]; P.gb[l] = gb[l]; if (!gW[l] || !gb[l]) return fail("null layer pointer"); }
	// samples per CTA: as large as possible (fewer atomics) while the hidden layers' CTAs still cover the SMs
	int chunk = kGChunkMax;
	const int heavy = sh->n_hidden_layers > 0 ? sh->n_hidden_layers : 1;
	const long long want = (sh->hidden == 64 ? 2ll : 1ll)*smCount(); // H = 128 threads carry 4x the atomics: fewer, larger chunks (measured)
	while (chunk > 64 && ((n + chunk - 1)/chunk)*heavy < want) chunk >>= 1;
	dim3 grid((unsigned)((n + chunk - 1)/chunk), (unsigned)(sh->n_hidden_layers + 2));
	cudaStream_t st = (cudaStream_t)stream;
	if (sh->hidden == 64) sirenWeightGrad<64><<<grid, kGT, 0, st>>>(P, sh->in_dim, sh->out_dim, sh->n_hidden_layers, x, n, dZ, A, chunk);
	else sirenWeightGrad<128><<<grid, kGT, 0, st>>>(P, sh->in_dim, sh->out_dim, sh->n_hidden_layers, x, n, dZ, A, chunk);
	cudaError_t e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_mse_grad_fit(const float* y, const float* target, const float* sub, int64_t count, float* diff, float* grad_y, float* loss,
								float* zero, int64_t zero_count, long long* step_advance, float stop_threshold, int* stop_flag, void* stream) {
	if (count <= 0) return 0;
	if (!y || !target || !diff || !grad_y || !loss) return fail("null buffer");
	if (zero && (((uintptr_t)zero & 15) || zero_count < 0)) return fail("the buffer to clear must be 16-byte aligned");
	const int vec = (((uintptr_t)y | (uintptr_t)target | (uintptr_t)diff | (uintptr_t)grad_y | (uintptr_t)sub) & 15) == 0;
	int blocks = (int)((count/4 + 511)/512);
	blocks = blocks < 1 ? 1 : (blocks > kMseBlocks ? kMseBlocks : blocks);
	cudaError_t e = nmc_pdl::launch(mseGrad, dim3(blocks), dim3(512), 0, (cudaStream_t)stream, y, target, sub, (long long)count, diff, grad_y, loss, vec,
									zero_count > 0 ? zero : (float*)nullptr, (long long)zero_count, step_advance, stop_threshold, stop_flag);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_mse_grad(const float* y, const float* target, int64_t count, float* diff, float* grad_y, float* loss, void* stream) {
	return nmc_mse_grad_fit(y, target, nullptr, count, diff, grad_y, loss, nullptr, 0, nullptr, 0.0f, nullptr, stream);
}

extern "C" int nmc_adam_update_device(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
									  float eps, const long long* step, const int* stop_flag, void* stream) {
	if (n <= 0) return 0;
	if (!p || !g || !m || !v || !step) return fail("bad arguments");
	cudaError_t e = nmc_pdl::launch(adamKernelDev, dim3((unsigned)((n + 255)/256)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (long long)n, lr, beta1, beta2, eps, step, stop_flag);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_adam_update_fetch(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
									 float eps, const long long* step, const int* stop_flag, int64_t count, int slots, const float* ring_x,
									 const float* ring_t, const float* ring_s, float* out_x, float* out_t, float* out_s, void* stream) {
	if (n <= 0 || count <= 0) return fail("nmc_adam_update_fetch: empty update or fetch");
	if (!p || !g || !m || !v || !step || slots < 1 || !ring_x || !ring_t || !out_x || !out_t || (ring_s && !out_s)) return fail("bad arguments");
	const int vec = (count % 4 == 0) && ((((uintptr_t)ring_x | (uintptr_t)ring_t | (uintptr_t)ring_s | (uintptr_t)out_x | (uintptr_t)out_t | (uintptr_t)out_s) & 15) == 0);
	const int adamBlocks = (int)((n + 255)/256);
	long long fb = ((vec ? count/4 : count) + 255)/256;
	fb = fb < 1 ? 1 : (fb > 1184 ? 1184 : fb);
	cudaError_t e = nmc_pdl::launch(adamFetchKernel, dim3((unsigned)(adamBlocks + fb)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (long long)n, lr, beta1, beta2,
									eps, step, stop_flag, adamBlocks, (long long)count, slots, ring_x, ring_t, ring_s, out_x, out_t, out_s, vec);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_adam_step_device(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
									float eps, long long* step, void* stream) {
	if (n <= 0) return 0;
	if (!p || !g || !m || !v || !step) return fail("bad arguments");
	adamAdvance<<<1, 1, 0, (cudaStream_t)stream>>>(step);
	adamKernelDev<<<(unsigned)((n + 255)/256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, step, nullptr);
	cudaError_t e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
							 float eps, int64_t step, void* stream) {
	if (n <= 0) return 0;
	if (!p || !g || !m || !v || step < 1) return fail("bad arguments");
	float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
	adamKernel<<<(unsigned)((n + 255)/256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2));
	cudaError_t e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}
