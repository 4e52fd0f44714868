// csrc/siren_tc_fused_bwd.cu -- the whole backward pass of a fit iteration of the H = 64 networks in ONE tcgen05 kernel:
// delta chain AND every weight / bias gradient (update_network, src/2d/models/base.py:83-96; loss.backward() through
// src/2d/models/networks.py:47-57).  The two-kernel form (siren_tc_bwd.cu) writes every delta to global memory and reads it
// back, with the activations, in a second launch whose MMA batches wait on those loads; here a CTA keeps the deltas dZ_l and
// the activations A_{l-1} of its 128 samples in shared memory and uses the SAME buffers twice:
//   dA_{l-1} = dZ_l W_l            A operand: dZ_l [sample x neuron], K-major (K = neurons)                      M 128, N 64
//   dW_l    += dZ_l^T [A_{l-1} 1]  A operand: the dZ_l buffer read MN-major (M = neurons, K = samples);          M 64, N 72
//                                  B operand: the A_{l-1} buffer read MN-major, with a column of ones appended,
//                                  so that column 64 of the product is the bias gradient sum_s dZ_l[s][.]
// In the un-swizzled canonical layout the element (r, c) of a [rows x K] K-major operand sits at
// (r/8)(32 K) + (c/4) 128 + (r%8) 16 + (c%4) 4; read MN-major with MN = c and K = r the same bytes are a canonical operand
// with SBO = 128 (16-byte chunks of four neurons) and LBO = 32 K (groups of eight samples) -- cute::UMMA make_umma_desc<Major::MN>,
// mma_traits_sm100.hpp:238-270 -- and one K step (8 samples) advances the start address by LBO.
// First layer: dW_0 | db_0 = dZ_0^T [x 1] (N = 8); last layer: dW_last^T = A_L^T gy' (N = 8); both as M = 64 MMAs over small
// [128 x 8] operands.  All products are 3xTF32 (hi.hi + hi.lo + lo.hi).  The gradient tiles accumulate in TMEM across the
// CTA's tiles (512 columns: 64 for the chain, 72 per hidden layer, 16 for the small layers: up to 6 hidden layers) and are added
// to the gradient buffer once, with 16-byte vector reductions.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nmcfs_siren.h"
#include "siren_env.cuh"
#include "siren_tc.cuh"

namespace nmc_siren_detail { void setError(const char* m); }

namespace {

using namespace nmc_siren_tc;
using nmc_siren_detail::Env;

constexpr int H = 64;
constexpr int kTile = 128, kThreads = 256, kMaxLayers = 18, kMaxHidden = 6;
constexpr int KA = H + 8;        // columns of the activation buffer: 64 activations, a one, seven zeros
constexpr int HC = H/2;          // columns per thread (two threads per sample row)

struct Params {
	const float* W[kMaxLayers];
	float* gW[kMaxLayers];
	float* gb[kMaxLayers];
};

// instruction descriptor with both operands MN-major (bits 15 / 16 of cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t instrDescTf32MN(int M, int N) { return instrDescTf32(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ void tmemLoad8(uint32_t taddr, uint32_t (&v)[8]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
				 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
sirenBackwardFusedTc(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
					 const float* __restrict__ zSaved, const float* __restrict__ gy) {
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned char* Dhi = smem;                        // dZ_l [128 x 64]
	unsigned char* Dlo = Dhi + kTile*H*4;
	unsigned char* Ahi = Dlo + kTile*H*4;             // [A_{l-1} | 1 | 0..] [128 x 72]
	unsigned char* Alo = Ahi + kTile*KA*4;
	unsigned char* Bhi = Alo + kTile*KA*4;            // W_l^T [64 x 64]
	unsigned char* Blo = Bhi + H*H*4;
	unsigned char* Shi = Blo + H*H*4;                 // small operand [128 x 8]: gy' (last layer), then [x 1] (first layer)
	unsigned char* Slo = Shi + kTile*8*4;
	float* sWL = reinterpret_cast<float*>(Slo + kTile*8*4);   // last layer's weights [outDim][H]
	__shared__ __align__(8) unsigned long long mbar;
	__shared__ uint32_t tmemBaseSh;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & (kTile - 1), half = tid >> 7;
	const int cBeg = half*HC;
	const int last = nHidden + 1;

	if (warp == 0) tmemAlloc(&tmemBaseSh, 512u);
	if (tid == 0) mbarInit(smemAddr(&mbar), 1);
	for (int i = tid; i < outDim*H; i += kThreads) sWL[i] = __ldg(&P.W[last][i]);
	fenceBeforeSync();
	__syncthreads();
	fenceAfterSync();
	const uint32_t tmemBase = tmemBaseSh;
	const uint32_t barAddr = smemAddr(&mbar);
	// TMEM columns: [0, 64) chain accumulator; hidden layer l: [64 + 72 (l - 1), + 72); small layers: 16 columns at the end
	const uint32_t colSmall = 64u + 72u*(uint32_t)nHidden;   // +0: dW_0 | db_0 (8 columns), +8: dW_last^T (8 columns)
#if defined(NMC_DBG_VARIANT) && NMC_DBG_VARIANT == 1
	const uint32_t idChain = instrDescTf32(kTile, H), idGrad = instrDescTf32(H, KA), idSmall = instrDescTf32MN(H, 8);
#elif defined(NMC_DBG_VARIANT) && NMC_DBG_VARIANT == 2
	const uint32_t idChain = instrDescTf32(kTile, H), idGrad = instrDescTf32MN(H, 64), idSmall = instrDescTf32MN(H, 8);
#elif defined(NMC_DBG_VARIANT) && NMC_DBG_VARIANT == 3
	const uint32_t idChain = instrDescTf32(kTile, H), idGrad = instrDescTf32(H, 64) | (1u << 16), idSmall = instrDescTf32MN(H, 8);
#else
	const uint32_t idChain = instrDescTf32(kTile, H), idGrad = instrDescTf32MN(H, KA), idSmall = instrDescTf32MN(H, 8);
#endif
	uint32_t phase = 0;

	constexpr int RW = H*H/4/kThreads; // 4
	float4 wreg[RW];
	auto loadW = [&](int l) { // B(n = input neuron, k = output neuron) = W_l[k][n]: four consecutive k of one n per 16-byte word
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/H, r = idx - k4*H;
			const float* w = &P.W[l][(size_t)(4*k4)*H + r];
			wreg[i] = make_float4(__ldg(w), __ldg(w + H), __ldg(w + 2*H), __ldg(w + 3*H));
		}
	};
	auto storeW = [&]() {
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/H, r = idx - k4*H;
			float4 h, o;
			splitTf32(wreg[i], h, o);
			const int off = coreOffsetBytes<H>(r, 4*k4);
			*reinterpret_cast<float4*>(Bhi + off) = h;
			*reinterpret_cast<float4*>(Blo + off) = o;
		}
	};
	// three-product MMA batch over `ksteps` K steps; the descriptors advance by aStep / bStep (16-byte units) per step
	auto issue = [&](uint32_t d, uint64_t aH, uint64_t aL, uint64_t bH, uint64_t bL, uint32_t aStep, uint32_t bStep, int ksteps, uint32_t idesc, bool accumulate) {
#pragma unroll 1
		for (int ks = 0; ks < ksteps; ks++) {
			mmaTf32(d, aH + (uint64_t)(aStep*ks), bH + (uint64_t)(bStep*ks), idesc, (accumulate || ks > 0) ? 1u : 0u);
			mmaTf32(d, aH + (uint64_t)(aStep*ks), bL + (uint64_t)(bStep*ks), idesc, 1u);
			mmaTf32(d, aL + (uint64_t)(aStep*ks), bH + (uint64_t)(bStep*ks), idesc, 1u);
		}
	};
	// descriptors (start address in 16-byte units in the low bits: adding to the 64-bit value moves the operand)
	const uint64_t dD_K_h = smemDesc(smemAddr(Dhi), 128, H*32), dD_K_l = smemDesc(smemAddr(Dlo), 128, H*32);          // chain A: K-major, K = neurons
	const uint64_t dB_K_h = smemDesc(smemAddr(Bhi), 128, H*32), dB_K_l = smemDesc(smemAddr(Blo), 128, H*32);          // chain B: K-major
	const uint64_t dD_MN_h = smemDesc(smemAddr(Dhi), H*32, 128), dD_MN_l = smemDesc(smemAddr(Dlo), H*32, 128);        // gradient A: MN-major (LBO = sample groups, SBO = neuron chunks)
	const uint64_t dA_MN_h = smemDesc(smemAddr(Ahi), KA*32, 128), dA_MN_l = smemDesc(smemAddr(Alo), KA*32, 128);      // gradient B / last-layer A
	const uint64_t dS_MN_h = smemDesc(smemAddr(Shi), 8*32, 128), dS_MN_l = smemDesc(smemAddr(Slo), 8*32, 128);        // small B

	// the ones column of the activation buffer (and its zero padding) never changes
	if (half == 1) {
		const int off = coreOffsetBytes<KA>(row, H);
		*reinterpret_cast<float4*>(Ahi + off) = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
		*reinterpret_cast<float4*>(Ahi + off + 128) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		*reinterpret_cast<float4*>(Alo + off) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		*reinterpret_cast<float4*>(Alo + off + 128) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	if ((long long)blockIdx.x*kTile < n) loadW(nHidden);

	int tilesDone = 0;
	float bl0 = 0.0f, bl1 = 0.0f, bl2 = 0.0f; // last layer's bias gradient: sum of gy' over this thread's samples (half 0 only)
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x, tilesDone++) {
		const bool acc = tilesDone > 0;
		const long long s = tile*kTile + row;
		const bool live = s < n;
		const long long sc = live ? s : n - 1;   // rows past the end read the last sample; their deltas are zeroed
		const float* zrow = zSaved + sc + (size_t)cBeg*n;
		float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
		if (live) {
			g0 = gy[s*outDim]; if (outDim > 1) g1 = gy[s*outDim + 1]; if (outDim > 2) g2 = gy[s*outDim + 2];
			x0 = x[s*inDim]; if (inDim > 1) x1 = x[s*inDim + 1]; if (inDim > 2) x2 = x[s*inDim + 2];
			if (env.active) { // dL/d(network output) = dL/d(enveloped output) x (detached) envelope weights
				const float xs[3] = {x0, x1, x2};
				float gys[3] = {g0, g1, g2};
				nmc_siren_detail::envBackward(env, inDim, outDim, xs, nullptr, gys, nullptr);
				g0 = gys[0]; g1 = gys[1]; g2 = gys[2];
			}
		}
		if (half == 0) { // small operand: gy' padded to eight columns
			bl0 += g0; bl1 += g1; bl2 += g2;
			float4 h, o;
			splitTf32(make_float4(g0, g1, g2, 0.0f), h, o);
			const int off = coreOffsetBytes<8>(row, 0);
			*reinterpret_cast<float4*>(Shi + off) = h; *reinterpret_cast<float4*>(Slo + off) = o;
			*reinterpret_cast<float4*>(Shi + off + 128) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
			*reinterpret_cast<float4*>(Slo + off + 128) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		}
		float zreg[HC];
		{
			const float* zp = zrow + (size_t)nHidden*H*n;
#pragma unroll
			for (int q = 0; q < HC; q++) { zreg[q] = __ldg(zp); zp += n; }
		}
		// layer L: A_L = sin(w0 z_L) (for dW_last) and dZ_L = (W_last^T gy') w0 cos(w0 z_L)
#pragma unroll
		for (int q4 = 0; q4 < HC; q4 += 4) {
			float d[4], a[4];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const int c = cBeg + q4 + q;
				float v = sWL[c]*g0;
				if (outDim > 1) v += sWL[H + c]*g1;
				if (outDim > 2) v += sWL[2*H + c]*g2;
				const float t = 6.283185307179586f*turnsReduced(w0*zreg[q4 + q]);
				a[q] = live ? __sinf(t) : 0.0f;
				d[q] = v*w0*__cosf(t);
			}
			float4 h, o;
			splitTf32(make_float4(d[0], d[1], d[2], d[3]), h, o);
			int off = coreOffsetBytes<H>(row, cBeg + q4);
			*reinterpret_cast<float4*>(Dhi + off) = h; *reinterpret_cast<float4*>(Dlo + off) = o;
			splitTf32(make_float4(a[0], a[1], a[2], a[3]), h, o);
			off = coreOffsetBytes<KA>(row, cBeg + q4);
			*reinterpret_cast<float4*>(Ahi + off) = h; *reinterpret_cast<float4*>(Alo + off) = o;
		}
		fenceProxyAsync();
		fenceBeforeSync();
		__syncthreads();
		if (tid == 0) { // dW_last^T [64 x 8] += A_L^T gy'
			fenceAfterSync();
			issue(tmemBase + colSmall + 8u, dA_MN_h, dA_MN_l, dS_MN_h, dS_MN_l, KA*2, 16, kTile/8, idSmall, acc);
			mmaCommit(barAddr);
		}
		for (int l = nHidden; l >= 1; l--) {
			// pre-activations of layer l - 1 (in flight during the wait below)
			{
				const float* zp = zrow + (size_t)(l - 1)*H*n;
#pragma unroll
				for (int q = 0; q < HC; q++) { zreg[q] = __ldg(zp); zp += n; }
			}
			if (l == nHidden) { // the last-layer batch has finished reading the activation and small operands
				mbarWait(barAddr, phase);
				phase ^= 1u;
				fenceAfterSync();
			}
			storeW();
			float cs[HC];
#pragma unroll
			for (int q4 = 0; q4 < HC; q4 += 4) { // A_{l-1} = sin(w0 z_{l-1}) -> operand; cos kept for the epilogue
				float a[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					const float t = 6.283185307179586f*turnsReduced(w0*zreg[q4 + q]);
					a[q] = live ? __sinf(t) : 0.0f;
					cs[q4 + q] = live ? w0*__cosf(t) : 0.0f;
				}
				float4 h, o;
				splitTf32(make_float4(a[0], a[1], a[2], a[3]), h, o);
				const int off = coreOffsetBytes<KA>(row, cBeg + q4);
				*reinterpret_cast<float4*>(Ahi + off) = h; *reinterpret_cast<float4*>(Alo + off) = o;
			}
			fenceProxyAsync();
			fenceBeforeSync();
			__syncthreads();
			if (l > 1) loadW(l - 1);
			else if ((tile + gridDim.x)*kTile < n) loadW(nHidden);
			if (tid == 0) {
				fenceAfterSync();
				issue(tmemBase, dD_K_h, dD_K_l, dB_K_h, dB_K_l, 16, 16, H/8, idChain, false);                                   // dA_{l-1}
				issue(tmemBase + 64u + 72u*(uint32_t)(l - 1), dD_MN_h, dD_MN_l, dA_MN_h, dA_MN_l, H*2, KA*2, kTile/8, idGrad, acc); // dW_l | db_l
				mmaCommit(barAddr);
			}
			mbarWait(barAddr, phase);
			phase ^= 1u;
			fenceAfterSync();
			// epilogue: dZ_{l-1} = dA_{l-1} w0 cos(w0 z_{l-1}) -> the D operand of the next layer (and of dW_{l-1})
			uint32_t v[HC];
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 16) tmemLoad16Async(tmemBase + ((uint32_t)((warp & 3)*32) << 16) + (uint32_t)(cBeg + c0), &v[c0]);
			tmemLoadWait();
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 4) {
				float4 h, o;
				splitTf32(make_float4(__uint_as_float(v[c0])*cs[c0], __uint_as_float(v[c0 + 1])*cs[c0 + 1], __uint_as_float(v[c0 + 2])*cs[c0 + 2], __uint_as_float(v[c0 + 3])*cs[c0 + 3]), h, o);
				const int off = coreOffsetBytes<H>(row, cBeg + c0);
				*reinterpret_cast<float4*>(Dhi + off) = h; *reinterpret_cast<float4*>(Dlo + off) = o;
			}
		}
		// first layer: dW_0 | db_0 = dZ_0^T [x 1]
		if (half == 0) {
			float4 h, o;
			splitTf32(make_float4(x0, x1, x2, 1.0f), h, o);
			const int off = coreOffsetBytes<8>(row, 0);
			*reinterpret_cast<float4*>(Shi + off) = h; *reinterpret_cast<float4*>(Slo + off) = o; // columns 4..7 stay zero
		}
		fenceProxyAsync();
		fenceBeforeSync();
		__syncthreads();
		if (tid == 0) {
			fenceAfterSync();
			issue(tmemBase + colSmall, dD_MN_h, dD_MN_l, dS_MN_h, dS_MN_l, H*2, 16, kTile/8, idSmall, acc);
			mmaCommit(barAddr);
		}
		mbarWait(barAddr, phase); // D and the small operand are rewritten by the next tile
		phase ^= 1u;
		fenceAfterSync();
	}

	// ---- gradient tiles -> gradient buffer.  M = 64 accumulators: row i in TMEM lane (i % 16) + 32 (i / 16) -------------------
#ifdef NMC_FUSED_DEBUG
	if (tilesDone > 0 && blockIdx.x == 0 && warp < 4) { // raw dump: TMEM lane (warp*32 + lane), 32 columns of hidden layer 1's region and 16 of the chain's
		uint32_t v[16];
		for (int c0 = 0; c0 < 32; c0 += 16) {
			tmemLoad16(tmemBase + ((uint32_t)(warp*32) << 16) + 64u + (uint32_t)c0, v);
			for (int q = 0; q < 16; q++) P.gW[1][(warp*32 + lane)*32 + c0 + q] = __uint_as_float(v[q]);
		}
	}
	fenceBeforeSync();
	__syncthreads();
	if (warp == 0) tmemFree(tmemBase, 512u);
	return;
#endif
	if (tilesDone > 0) {
		const int sp = warp & 3, ch = warp >> 2;
		const int i = sp*16 + (lane & 15);
		const bool valid = lane < 16;
		const uint32_t laneBase = tmemBase + ((uint32_t)(sp*32) << 16);
		for (int l = 1; l <= nHidden; l++) {
			const uint32_t col = 64u + 72u*(uint32_t)(l - 1);
			float* gw = P.gW[l] + (size_t)i*H;
			for (int c0 = ch*32; c0 < ch*32 + 32; c0 += 16) {
				uint32_t v[16];
				tmemLoad16(laneBase + col + (uint32_t)c0, v);
				if (valid) {
#pragma unroll
					for (int q = 0; q < 16; q += 4)
						redAdd4(gw + c0 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
				}
			}
			if (ch == 1) { // column 64: the bias gradient
				uint32_t v[8];
				tmemLoad8(laneBase + col + 64u, v);
				if (valid) atomicAdd(&P.gb[l][i], __uint_as_float(v[0]));
			}
		}
		if (ch == 0) {
			uint32_t v[16];
			tmemLoad16(laneBase + colSmall, v);
			if (valid) {
				for (int j = 0; j < inDim; j++) atomicAdd(&P.gW[0][i*inDim + j], __uint_as_float(v[j]));
				atomicAdd(&P.gb[0][i], __uint_as_float(v[3]));
				for (int j = 0; j < outDim; j++) atomicAdd(&P.gW[last][j*H + i], __uint_as_float(v[8 + j]));
			}
		}
		if (half == 0) { // last layer's bias gradient
			for (int off = 16; off > 0; off >>= 1) {
				bl0 += __shfl_xor_sync(0xffffffffu, bl0, off); bl1 += __shfl_xor_sync(0xffffffffu, bl1, off); bl2 += __shfl_xor_sync(0xffffffffu, bl2, off);
			}
			if (lane == 0) {
				atomicAdd(&P.gb[last][0], bl0);
				if (outDim > 1) atomicAdd(&P.gb[last][1], bl1);
				if (outDim > 2) atomicAdd(&P.gb[last][2], bl2);
			}
		}
	}
	fenceBeforeSync();
	__syncthreads();
	if (warp == 0) tmemFree(tmemBase, 512u);
}

int smCount() {
	static int sms = 0;
	if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
	return sms;
}
int fail(const char* m) { nmc_siren_detail::setError(m); return 1; }

} // namespace

extern "C" int nmc_siren_backward_fused_tc(const nmc_siren_shape* sh, const float* const* W, const float* x, int64_t n,
										   const float* z_saved, const float* grad_y, float* const* gW, float* const* gb,
										   const nmc_siren_envelope* envp, void* stream) {
	if (!sh || !W || !gW || !gb) return fail("null argument");
	if (sh->hidden != 64 || sh->n_hidden_layers < 1 || sh->n_hidden_layers > kMaxHidden || sh->in_dim < 1 || sh->in_dim > 3 || sh->out_dim < 1 || sh->out_dim > 3)
		return fail("fused tensor-core backward: unsupported shape (hidden 64, 1..6 hidden layers, in/out 1..3)");
	if (n <= 0) return 0;
	if (!x || !z_saved || !grad_y) return fail("null buffer");
	Params P;
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) {
		P.W[l] = W[l]; P.gW[l] = gW[l]; P.gb[l] = gb[l];
		if (!W[l] || !gW[l] || !gb[l]) return fail("null layer pointer");
		if (l >= 1 && l <= sh->n_hidden_layers && ((uintptr_t)gW[l] & 15)) return fail("fused tensor-core backward: hidden-layer gradient buffers must be 16-byte aligned");
	}
	Env env;
	if (const char* bad = nmc_siren_detail::toEnv(envp, env)) return fail(bad);
	const size_t smem = (size_t)(2*kTile*H + 2*kTile*KA + 2*H*H + 2*kTile*8)*4 + (size_t)sh->out_dim*H*4;
	const long long tiles = (n + kTile - 1)/kTile;
	const int grid = (int)(tiles < smCount() ? tiles : smCount());
	cudaError_t e = cudaFuncSetAttribute(sirenBackwardFusedTc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (!e) sirenBackwardFusedTc<<<grid, kThreads, smem, (cudaStream_t)stream>>>(P, env, sh->in_dim, sh->out_dim, sh->n_hidden_layers, sh->w0, x, n, z_saved, grad_y);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}
