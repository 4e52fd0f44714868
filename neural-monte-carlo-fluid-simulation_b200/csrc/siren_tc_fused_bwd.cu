// csrc/siren_tc_fused_bwd.cu -- the whole backward pass of a fit iteration of the H = 64 networks in ONE tcgen05 kernel:
// delta chain AND every weight / bias gradient (update_network, src/2d/models/base.py:83-96; loss.backward() through
// src/2d/models/networks.py:47-57).  The two-kernel form (siren_tc_bwd.cu) writes every delta to global memory and reads it
// back, with the activations, in a second launch whose MMA batches wait on those loads; here a CTA keeps the deltas dZ_l and
// the activations A_{l-1} of its 128 samples in shared memory and uses the SAME buffers twice:
//   dA_{l-1} = dZ_l W_l            A operand: dZ_l [sample x neuron], K-major (K = neurons), no swizzle            M 128, N 64
//   dW_l    += dZ_l^T A_{l-1}      A operand: dZ_l read MN-major (M = neurons, K = samples);                       M 64, N 64
//                                  B operand: A_{l-1} read MN-major
// tcgen05.mma.kind::tf32 reads an MN-major operand ONLY in the 128-byte-swizzle / 32-byte-base layout (descriptor layout type
// 1, cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B): with no swizzle or the 32 / 64 / 128-byte swizzles the instruction runs
// and multiplies zeros (profiles/tools/umma_mn_probe.cu reads the addresses back through one-hot operands).  That layout is
// rows of 32 MN elements (128 bytes) per K index, four K rows per 512-byte atom, the 32-byte chunk index XORed with k % 4;
// SBO = stride between atoms along K, LBO = stride between groups of 32 along MN.  A K-major read of the same bytes does not
// exist (layout type 1 faults for K-major operands, the 128-byte K-major swizzle permutes 16-byte chunks by row % 8), so the
// deltas are written twice: K-major for the chain, MN-major for the gradients; the activations only MN-major.
// First layer: dW_0 | db_0 = dZ_0^T [x 1] (N = 8); last layer: dW_last^T = A_L^T gy' (N = 8): the small operand [8 x 128] is
// K-major (K = samples).  All products are 3xTF32 (hi.hi + hi.lo + lo.hi).  The gradient tiles accumulate in TMEM across the
// CTA's tiles (512 columns: 64 for the chain, 72 per hidden layer, 16 for the small layers: up to 6 hidden layers) and are
// added to the gradient buffer once, with 16-byte vector reductions.  The bias gradients (column sums of the deltas) are
// reduced with warp shuffles in the epilogues and shared-memory atomics.
// Per layer the chain batch and the gradient batch are committed to two mbarriers: the epilogue (TMEM load, cosine factor,
// K-major store of the next delta, sine / cosine of the next layer's pre-activations) only waits for the chain and overlaps
// the gradient MMAs; the MN-major stores wait for the gradient batch that still reads those buffers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/nmcfs_siren.h"
#include "siren_env.cuh"
#include "siren_tc.cuh"
#include "pdl.cuh"

namespace nmc_siren_detail { void setError(const char* m); }

namespace {

using namespace nmc_siren_tc;
using nmc_siren_detail::Env;

constexpr int H = 64;
constexpr int kTile = 128, kWorkers = 256, kThreads = kWorkers + 128, kMaxLayers = 18, kMaxHidden = 6;   // warps 0..7: operands and epilogues; warps 8..11: MMA issue (one thread)
constexpr int HC = H/2;          // columns per thread (two threads per sample row) = one MN group of 32 neurons
constexpr int kMnGroup = kTile*128;   // bytes of one [128 samples x 32 neurons] block of an MN-major buffer
constexpr uint32_t kTransA = 1u << 15, kTransB = 1u << 16;   // instruction descriptor: operand is MN-major
constexpr uint64_t kLayoutMn = 1ull << 61;                   // shared-memory descriptor: 128-byte swizzle, 32-byte base

// MN-major operand (MN = neurons, K = samples): byte offset of the 32-byte chunk holding neurons 8j .. 8j+7 (j = 0..3 within
// the thread's group of 32) of sample r
__device__ __forceinline__ int mnChunkOffset(int group, int r, int j) {
	return group*kMnGroup + (r >> 2)*512 + (r & 3)*128 + ((j ^ (r & 3)) << 5);
}
// small K-major operand [8 x 128 samples]: element (n, sample r)
__device__ __forceinline__ int smallOffset(int n, int r) { return (r >> 2)*128 + n*16 + (r & 3)*4; }

// -DNMC_TC_TRACE: CTA 0, threads 256 (the MMA issuer) and 64 stamp clock64() at the phase boundaries of the first tile
// (profiles/tools/fused_bwd_trace.py)
#ifdef NMC_TC_TRACE
__device__ long long g_ftrace[2][256];
__device__ int g_ftraceN[2];
#define FTRACE(tag) do { if (blockIdx.x == 0 && (tid == 256 || tid == 64) && (tilesDone == 0 || (tag) >= 22) && tn < 127) { const int sl = tid == 256 ? 0 : 1; g_ftrace[sl][2*tn] = (tag); g_ftrace[sl][2*tn + 1] = clock64(); tn++; g_ftraceN[sl] = tn; } } while (0)
#else
#define FTRACE(tag) do {} while (0)
#endif

// warp specialisation: the issuing warpgroup hands its registers to the two working warpgroups (168 per thread at launch for 384
// threads; 256 x 240 + 128 x 24 = 64512 afterwards).  The roles meet at a named barrier (every thread of the CTA arrives).
__device__ __forceinline__ void regsInc240() { asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory"); }
__device__ __forceinline__ void regsDec24() { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory"); }
// thread-block cluster: barrier over every thread of every CTA, and a 16-byte load from a peer CTA's shared memory
__device__ __forceinline__ void clusterSync() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t clusterRank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ float4 loadPeer16(uint32_t localAddr, uint32_t rank) {
	uint32_t remote;
	float4 v;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(localAddr), "r"(rank));
	asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
	return v;
}
__device__ __forceinline__ void ctaBarrier() { asm volatile("bar.sync 1, %0;" :: "n"(kThreads) : "memory"); }

struct Params {
	const float* W[kMaxLayers];
	float* gW[kMaxLayers];
	float* gb[kMaxLayers];
};

__global__ void __launch_bounds__(kThreads, 1)
sirenBackwardFusedTc(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
					 const float* __restrict__ zSaved, const float* __restrict__ gy, int cluster) {
	nmc_pdl::gridEnter();
	extern __shared__ __align__(1024) unsigned char smem[];
	unsigned char* Dhi = smem;                        // dZ_l [128 x 64], K-major (chain operand)
	unsigned char* Dlo = Dhi + kTile*H*4;
	unsigned char* Mhi = Dlo + kTile*H*4;             // dZ_l, MN-major (gradient operand)
	unsigned char* Mlo = Mhi + kTile*H*4;
	unsigned char* Ahi = Mlo + kTile*H*4;             // A_{l-1}, MN-major
	unsigned char* Alo = Ahi + kTile*H*4;
	unsigned char* Bhi = Alo + kTile*H*4;             // W_l^T [64 x 64]
	unsigned char* Blo = Bhi + H*H*4;
	// the small operand [8 x 128] (gy' for the last layer, then [x 1] for the first) lives in the weight buffer: it is written
	// and read while no chain batch is in flight (before the first storeW of a tile / after the last chain batch)
	unsigned char* Shi = Bhi;
	unsigned char* Slo = Blo;
	float* sWL = reinterpret_cast<float*>(Blo + H*H*4);             // last layer's weights [3][H]
	float* sBias = sWL + 3*H;                                       // bias-gradient sums [nHidden + 1][H]
	float* sBl = sBias + (kMaxHidden + 1)*H;                        // last layer's bias gradient (cluster reduction), 16 bytes
	unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sBl + 4);   // [0] gradient batches, [1] chain batches
	uint32_t& tmemBaseSh = *reinterpret_cast<uint32_t*>(mbar + 2);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & (kTile - 1), half = tid >> 7;
	const int cBeg = half*HC;
	const int last = nHidden + 1;
	const bool worker = tid < kWorkers;        // thread 256 only issues the MMAs: issuing the 72 MMAs of a layer takes ~5000 cycles, which a
	const bool issuer = tid == kWorkers;       // working warp would add to its own epilogue (and to everybody's barrier wait)

	if (warp == 0) tmemAlloc(&tmemBaseSh, 512u);
	if (tid == 0) { mbarInit(smemAddr(&mbar[0]), 1); mbarInit(smemAddr(&mbar[1]), 1); }
	for (int i = tid; i < (kMaxHidden + 1)*H + 4; i += kThreads) sBias[i] = 0.0f;   // and sBl
	for (int i = tid; i < outDim*H; i += kThreads) sWL[i] = __ldg(&P.W[last][i]);
	fenceBeforeSync();
	__syncthreads();
	fenceAfterSync();
	const uint32_t tmemBase = tmemBaseSh;
	const uint32_t barG = smemAddr(&mbar[0]), barC = smemAddr(&mbar[1]);
	// TMEM columns: [0, 64) chain accumulator; hidden layer l: [64 + 72 (l - 1), + 72); small layers: 16 columns at the end
	const uint32_t colSmall = 64u + 72u*(uint32_t)nHidden;   // +0: dW_0 | db_0 (8 columns), +8: dW_last^T (8 columns)
	const uint32_t idChain = instrDescTf32(kTile, H), idGrad = instrDescTf32(H, H) | kTransA | kTransB, idSmall = instrDescTf32(H, 8) | kTransA;
	uint32_t phaseG = 0, phaseC = 0;
	bool liveRow = false;

	constexpr int RW = H*H/4/kWorkers; // 4
	float4 wreg[RW];
	auto loadW = [&](int l) { // B(n = input neuron, k = output neuron) = W_l[k][n]: four consecutive k of one n per 16-byte word
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kWorkers, k4 = idx/H, r = idx - k4*H;
			const float* w = &P.W[l][(size_t)(4*k4)*H + r];
			wreg[i] = make_float4(__ldg(w), __ldg(w + H), __ldg(w + 2*H), __ldg(w + 3*H));
		}
	};
	auto storeW = [&]() {
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kWorkers, k4 = idx/H, r = idx - k4*H;
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
	const uint64_t dD_MN_h = smemDesc(smemAddr(Mhi), kMnGroup, 512) | kLayoutMn, dD_MN_l = smemDesc(smemAddr(Mlo), kMnGroup, 512) | kLayoutMn;   // gradient A
	const uint64_t dA_MN_h = smemDesc(smemAddr(Ahi), kMnGroup, 512) | kLayoutMn, dA_MN_l = smemDesc(smemAddr(Alo), kMnGroup, 512) | kLayoutMn;   // gradient B / last-layer A
	const uint64_t dS_K_h = smemDesc(smemAddr(Shi), 128, 256), dS_K_l = smemDesc(smemAddr(Slo), 128, 256);            // small B [8 x 128], K-major
	constexpr uint32_t kMnStep = 1024/16, kSmallStep = 256/16;   // one K step = 8 samples

	if (worker && (long long)blockIdx.x*kTile < n) loadW(nHidden);

	// this thread's 32 values of a row (sample `row`, neurons cBeg .. cBeg + 31) -> the operand buffers, split into TF32 hi / lo
	auto storeKMajor = [&](const float (&val)[HC]) {
#pragma unroll
		for (int q4 = 0; q4 < HC; q4 += 4) {
			float4 h, o;
			splitTf32(make_float4(val[q4], val[q4 + 1], val[q4 + 2], val[q4 + 3]), h, o);
			const int off = coreOffsetBytes<H>(row, cBeg + q4);
			*reinterpret_cast<float4*>(Dhi + off) = h; *reinterpret_cast<float4*>(Dlo + off) = o;
		}
	};
	// MN-major rows are 128 bytes per sample with the 32-byte chunks rotated by r % 4: lanes r and r + 4 of a quarter warp would
	// hit the same 16 bytes' banks, so lanes with bit 2 set write the two halves of a chunk in the opposite order
	const bool swapHalves = (row & 4) != 0;
	auto storeMnMajor = [&](unsigned char* hi, unsigned char* lo, const float (&val)[HC]) {
#pragma unroll
		for (int q8 = 0; q8 < HC; q8 += 8) {
			float4 h0, o0, h1, o1;
			splitTf32(make_float4(val[q8], val[q8 + 1], val[q8 + 2], val[q8 + 3]), h0, o0);
			splitTf32(make_float4(val[q8 + 4], val[q8 + 5], val[q8 + 6], val[q8 + 7]), h1, o1);
			const int off = mnChunkOffset(half, row, q8 >> 3);
			const int offA = off + (swapHalves ? 16 : 0), offB = off + (swapHalves ? 0 : 16);
			*reinterpret_cast<float4*>(hi + offA) = swapHalves ? h1 : h0; *reinterpret_cast<float4*>(lo + offA) = swapHalves ? o1 : o0;
			*reinterpret_cast<float4*>(hi + offB) = swapHalves ? h0 : h1; *reinterpret_cast<float4*>(lo + offB) = swapHalves ? o0 : o1;
		}
	};
	auto loadZ = [&](float (&z)[HC], const float* zrow, int layer) {
		const float* zp = zrow + (size_t)layer*H*n;
#pragma unroll
		for (int q = 0; q < HC; q++) { z[q] = __ldg(zp); zp += n; }
	};
	// bias gradients db_l = sum over samples of dZ_l: transposed butterfly over the warp's 32 rows (lane j ends with the sum of
	// column cBeg + j), then one shared-memory atomic per lane; the sums leave with the gradient tiles
	auto biasSum = [&](const float (&val)[HC], int layer) {
		float t[HC];
#pragma unroll
		for (int q = 0; q < HC; q++) t[q] = val[q];
#pragma unroll
		for (int off = 16; off >= 1; off >>= 1) {
			const bool upper = (lane & off) != 0;
#pragma unroll
			for (int q = 0; q < off; q++) {
				const float send = upper ? t[q] : t[q + off];
				const float keep = upper ? t[q + off] : t[q];
				t[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
			}
		}
		atomicAdd(&sBias[layer*H + cBeg + lane], t[0]);
	};
	auto activations = [&](const float (&z)[HC], float (&av)[HC], float (&cv)[HC]) {
#pragma unroll
		for (int q = 0; q < HC; q++) {
			const float t = 6.283185307179586f*turnsReduced(w0*z[q]);
			av[q] = liveRow ? __sinf(t) : 0.0f;
			cv[q] = liveRow ? w0*__cosf(t) : 0.0f;
		}
	};

	int tilesDone = 0;
#ifdef NMC_TC_TRACE
	int tn = 0;
#endif
	float bl0 = 0.0f, bl1 = 0.0f, bl2 = 0.0f; // last layer's bias gradient: sum of gy' over this thread's samples (half 0 only)
	if (!worker) {
		// ---- issuing warpgroup: the same barrier sequence as the workers, one thread issues after each -------------------------
		regsDec24();
		for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x, tilesDone++) {
			const bool acc = tilesDone > 0;
			FTRACE(1);
			ctaBarrier();
			if (issuer) { // dW_last^T [64 x 8] += A_L^T gy'
				fenceAfterSync();
				issue(tmemBase + colSmall + 8u, dA_MN_h, dA_MN_l, dS_K_h, dS_K_l, kMnStep, kSmallStep, kTile/8, idSmall, acc);
				mmaCommit(barG);
			}
			FTRACE(4);
			for (int l = nHidden; l >= 1; l--) {
				ctaBarrier();
				FTRACE(13);
				if (issuer) {
					fenceAfterSync();
					issue(tmemBase, dD_K_h, dD_K_l, dB_K_h, dB_K_l, 16, 16, H/8, idChain, false);                                   // dA_{l-1}
					mmaCommit(barC);
					issue(tmemBase + 64u + 72u*(uint32_t)(l - 1), dD_MN_h, dD_MN_l, dA_MN_h, dA_MN_l, kMnStep, kMnStep, kTile/8, idGrad, acc);   // dW_l
					mmaCommit(barG);
				}
				FTRACE(14);
			}
			ctaBarrier();
			if (issuer) { // dW_0 += dZ_0^T x
				fenceAfterSync();
				issue(tmemBase + colSmall, dD_MN_h, dD_MN_l, dS_K_h, dS_K_l, kMnStep, kSmallStep, kTile/8, idSmall, acc);
				mmaCommit(barG);
			}
			FTRACE(21);
		}
		if (cluster > 1) { clusterSync(); clusterSync(); }   // the working warpgroups' gradient reduction
	} else {
	regsInc240();
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x, tilesDone++) {
		const long long s = tile*kTile + row;
		const bool live = s < n;
		liveRow = live;
		FTRACE(1);
		const long long sc = live ? s : n - 1;   // rows past the end read the last sample; their deltas are zeroed
		const float* zrow = zSaved + sc + (size_t)cBeg*n;
		float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
		float zreg[HC];
		float d[HC];    // dZ of the layer the loop is about to process (this thread's half row)
		float a[HC];    // A_{l-1} = sin(w0 z_{l-1})
		float cs[HC];   // w0 cos(w0 z_{l-1})
		{
			float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
			loadZ(zreg, zrow, nHidden);
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
			if (half == 0) { // small operand: gy' padded to eight rows
				bl0 += g0; bl1 += g1; bl2 += g2;
				const float gs[8] = {g0, g1, g2, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
				for (int q = 0; q < 8; q++) {
					float h, o;
					splitTf32(gs[q], h, o);
					*reinterpret_cast<float*>(Shi + smallOffset(q, row)) = h; *reinterpret_cast<float*>(Slo + smallOffset(q, row)) = o;
				}
			}
			// layer L: A_L = sin(w0 z_L) (for dW_last) and dZ_L = (W_last^T gy') w0 cos(w0 z_L)
#pragma unroll
			for (int q = 0; q < HC; q++) {
				const int c = cBeg + q;
				float v = sWL[c]*g0;
				if (outDim > 1) v += sWL[H + c]*g1;
				if (outDim > 2) v += sWL[2*H + c]*g2;
				const float t = 6.283185307179586f*turnsReduced(w0*zreg[q]);
				a[q] = live ? __sinf(t) : 0.0f;
				d[q] = live ? v*w0*__cosf(t) : 0.0f;
			}
		}
		FTRACE(2);
		storeKMajor(d);
		storeMnMajor(Mhi, Mlo, d);
		storeMnMajor(Ahi, Alo, a);
		biasSum(d, nHidden);
		FTRACE(3);
		fenceProxyAsync();
		fenceBeforeSync();
		ctaBarrier();               // -> the last layer's batch
		FTRACE(4);
		loadZ(zreg, zrow, nHidden - 1);
		activations(zreg, a, cs);
		for (int l = nHidden; l >= 1; l--) {
			// here: Dk holds dZ_l; d[] = dZ_l (not yet in Dm when l < nHidden); a[], cs[] belong to layer l - 1; a gradient batch
			// reading Am, Dm (and, for l = nHidden, the small operand in the weight buffer) may still be in flight
			FTRACE(10);
			mbarWait(barG, phaseG);
			phaseG ^= 1u;
			fenceAfterSync();
			FTRACE(11);
			if (l < nHidden) storeMnMajor(Mhi, Mlo, d);
			storeW();
			storeMnMajor(Ahi, Alo, a);
			fenceProxyAsync();
			fenceBeforeSync();
			FTRACE(12);
			ctaBarrier();           // -> chain batch (barC), gradient batch (barG)
			FTRACE(13);
			float csn[HC];
			if (l > 1) { // the next layer's activations while the chain batch runs (a[] is in the operand buffer by now)
				loadW(l - 1);
				loadZ(zreg, zrow, l - 2);
				activations(zreg, a, csn);
			} else if ((tile + gridDim.x)*kTile < n) loadW(nHidden);
			FTRACE(14);
			mbarWait(barC, phaseC);
			phaseC ^= 1u;
			fenceAfterSync();
			FTRACE(15);
			// epilogue: dZ_{l-1} = dA_{l-1} w0 cos(w0 z_{l-1}); the chain batch has finished reading Dk, the gradient batch still reads Dm
			{
				uint32_t v[HC];
#pragma unroll
				for (int c0 = 0; c0 < HC; c0 += 16) tmemLoad16Async(tmemBase + ((uint32_t)((warp & 3)*32) << 16) + (uint32_t)(cBeg + c0), &v[c0]);
				tmemLoadWait();
#pragma unroll
				for (int q = 0; q < HC; q++) { d[q] = __uint_as_float(v[q])*cs[q]; cs[q] = csn[q]; }
			}
			FTRACE(16);
			if (l > 1) storeKMajor(d);
			FTRACE(17);
			biasSum(d, l - 1);
			FTRACE(18);
		}
		FTRACE(20);
		// first layer: dW_0 = dZ_0^T x
		mbarWait(barG, phaseG);
		phaseG ^= 1u;
		fenceAfterSync();
		storeMnMajor(Mhi, Mlo, d);
		if (half == 0) {
			const float xs4[4] = {x0, x1, x2, 0.0f};   // rows 3..7 of the operand stay zero
#pragma unroll
			for (int q = 0; q < 8; q++) {
				float h = 0.0f, o = 0.0f;
				if (q < 4) splitTf32(xs4[q], h, o);
				*reinterpret_cast<float*>(Shi + smallOffset(q, row)) = h; *reinterpret_cast<float*>(Slo + smallOffset(q, row)) = o;
			}
		}
		fenceProxyAsync();
		fenceBeforeSync();
		ctaBarrier();               // -> the first layer's batch
		mbarWait(barG, phaseG);     // Dm, Am and the small operand are rewritten by the next tile
		phaseG ^= 1u;
		fenceAfterSync();
		FTRACE(21);
	}
	// the gradients leave from the working warpgroups' branch: after the roles merge the code is compiled for the issuing
	// warpgroup's 24 registers (the read-out spilled everything and took 15-19 k cycles there)
	// ---- gradient tiles -> gradient buffer.  M = 64 accumulators: row i in TMEM lane (i % 16) + 32 (i / 16) -------------------
	if (tilesDone > 0 && cluster <= 1) {
		const int sp = warp & 3, ch = warp >> 2;
		const int i = sp*16 + (lane & 15);
		const bool valid = lane < 16;
		const uint32_t laneBase = tmemBase + ((uint32_t)(sp*32) << 16);
		// every CTA adds to the same 20 k addresses: the order is rotated by the CTA index, or the reductions of all CTAs queue
		// up at the same L2 slices at the same time (25 k cycles for 5 layers against 19 k rotated)
		for (int li = 0; li < nHidden; li++) {
			const int l = 1 + (li + (int)blockIdx.x) % nHidden;
			const uint32_t col = 64u + 72u*(uint32_t)(l - 1);
			float* gw = P.gW[l] + (size_t)i*H;
			for (int cc = 0; cc < 2; cc++) {
				const int c0 = ch*32 + 16*((cc + (int)blockIdx.x/nHidden) & 1);
				uint32_t v[16];
				tmemLoad16(laneBase + col + (uint32_t)c0, v);
				if (valid) {
#pragma unroll
					for (int q = 0; q < 16; q += 4)
						redAdd4(gw + c0 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
				}
			}
		}
		if (ch == 0) {
			uint32_t v[16];
			tmemLoad16(laneBase + colSmall, v);
			if (valid) {
				for (int j = 0; j < inDim; j++) atomicAdd(&P.gW[0][i*inDim + j], __uint_as_float(v[j]));
				for (int j = 0; j < outDim; j++) atomicAdd(&P.gW[last][j*H + i], __uint_as_float(v[8 + j]));
			}
		}
		for (int o = tid; o < (nHidden + 1)*H; o += kWorkers) atomicAdd(&P.gb[o/H][o % H], sBias[o]);   // complete: every biasSum precedes the tile's last __syncthreads
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
	FTRACE(23);
	if (cluster > 1) {
		// ---- every gradient reduced over the cluster first: a CTA leaves its tiles in its own shared memory (the operand buffers
		// are idle), CTA r of the cluster sums slice r of all peers through distributed shared memory and issues the global
		// reductions for that slice only: `cluster` times fewer L2 atomics on the same 21 k addresses (they were a quarter of
		// the kernel, the 128-deep same-address queues of the small layers and biases half of that)
		float4* stage4 = reinterpret_cast<float4*>(smem);   // [nHidden][64][16] hidden tiles, [64] dW_0 rows, [64] dW_last columns
		const int T4 = nHidden*H*H/4, B4 = (nHidden + 1)*H/4;
		{
			const int sp = warp & 3, ch = warp >> 2;
			const int i = sp*16 + (lane & 15);
			const uint32_t laneBase = tmemBase + ((uint32_t)(sp*32) << 16);
			for (int l = 1; l <= nHidden; l++) {
				uint32_t v[32];
#pragma unroll
				for (int q = 0; q < 32; q++) v[q] = 0u;
				if (tilesDone > 0) {
					tmemLoad16Async(laneBase + 64u + 72u*(uint32_t)(l - 1) + (uint32_t)(ch*32), &v[0]);
					tmemLoad16Async(laneBase + 64u + 72u*(uint32_t)(l - 1) + (uint32_t)(ch*32 + 16), &v[16]);
					tmemLoadWait();
				}
				if (lane < 16) { // a row is 16 float4: slot (column chunk ^ row % 16), or the 16 lanes of a warp would write one bank group
					float4* dst = stage4 + (size_t)(l - 1)*H*H/4 + i*(H/4);
#pragma unroll
					for (int q = 0; q < 32; q += 4) dst[(ch*8 + (q >> 2)) ^ (i & 15)] = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
				}
			}
			if (ch == 0) {
				uint32_t v[16];
#pragma unroll
				for (int q = 0; q < 16; q++) v[q] = 0u;
				if (tilesDone > 0) tmemLoad16(laneBase + colSmall, v);
				if (lane < 16) {
					stage4[T4 + i] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), 0.0f);
					stage4[T4 + H + i] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), 0.0f);
				}
			}
			if (half == 0) { // last layer's bias gradient
				for (int off = 16; off > 0; off >>= 1) {
					bl0 += __shfl_xor_sync(0xffffffffu, bl0, off); bl1 += __shfl_xor_sync(0xffffffffu, bl1, off); bl2 += __shfl_xor_sync(0xffffffffu, bl2, off);
				}
				if (lane == 0) { atomicAdd(&sBl[0], bl0); atomicAdd(&sBl[1], bl1); atomicAdd(&sBl[2], bl2); }
			}
		}
		FTRACE(24);
		clusterSync();
		FTRACE(25);
		{
			const int total = T4 + 2*H + B4 + 1, per = (total + cluster - 1)/cluster;
			const int beg = (int)clusterRank()*per, end = beg + per < total ? beg + per : total;
			const uint32_t stageAddr = smemAddr(stage4), biasAddr = smemAddr(sBias), blAddr = smemAddr(sBl);
#pragma unroll 2
			for (int idx = beg + tid; idx < end; idx += kWorkers) {
				const uint32_t addr = idx < T4 + 2*H ? stageAddr + (uint32_t)idx*16u : (idx < T4 + 2*H + B4 ? biasAddr + (uint32_t)(idx - T4 - 2*H)*16u : blAddr);
				float4 part[8];
#pragma unroll
				for (int p = 0; p < 8; p++) part[p] = p < cluster ? loadPeer16(addr, (uint32_t)p) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				float4 sum = part[0];
#pragma unroll
				for (int p = 1; p < 8; p++) { sum.x += part[p].x; sum.y += part[p].y; sum.z += part[p].z; sum.w += part[p].w; }
				if (idx < T4) {
					const int l = idx/(H*H/4), w = idx - l*(H*H/4), i = w >> 4, c4 = (w & 15) ^ (i & 15);
					redAdd4(P.gW[l + 1] + i*H + 4*c4, sum.x, sum.y, sum.z, sum.w);
				} else if (idx < T4 + H) {
					const int i = idx - T4;
					const float sv[3] = {sum.x, sum.y, sum.z};
					for (int j = 0; j < inDim; j++) atomicAdd(&P.gW[0][i*inDim + j], sv[j]);
				} else if (idx < T4 + 2*H) {
					const int i = idx - T4 - H;
					const float sv[3] = {sum.x, sum.y, sum.z};
					for (int j = 0; j < outDim; j++) atomicAdd(&P.gW[last][j*H + i], sv[j]);
				} else if (idx < T4 + 2*H + B4) {
					const int o = 4*(idx - T4 - 2*H), l = o/H, c = o - l*H;
					atomicAdd(&P.gb[l][c], sum.x); atomicAdd(&P.gb[l][c + 1], sum.y); atomicAdd(&P.gb[l][c + 2], sum.z); atomicAdd(&P.gb[l][c + 3], sum.w);
				} else {
					const float sv[3] = {sum.x, sum.y, sum.z};
					for (int j = 0; j < outDim; j++) atomicAdd(&P.gb[last][j], sv[j]);
				}
			}
		}
		FTRACE(26);
		clusterSync();   // no CTA leaves while a peer still reads its shared memory
	}
	}

	fenceBeforeSync();
	__syncthreads();
	FTRACE(22);
	if (warp == 0) tmemFree(tmemBase, 512u);
}

int smCount() {
	static int sms = 0;
	if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
	return sms;
}
int fail(const char* m) { nmc_siren_detail::setError(m); return 1; }

} // namespace

#ifdef NMC_TC_TRACE
extern "C" int nmc_siren_trace_read_fused(int slot, long long* out, int cap) { // (tag, clock) pairs of the last traced launch
	int n = 0;
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(&n, g_ftraceN, sizeof(int), sizeof(int)*slot);
	if (n > cap) n = cap;
	cudaMemcpyFromSymbol(out, g_ftrace, sizeof(long long)*2*n, sizeof(long long)*256*slot);
	return n;
}
#endif

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
	const size_t smem = (size_t)(6*kTile*H + 2*H*H)*4 + (size_t)(3 + kMaxHidden + 1)*H*4 + 48;
	const long long tiles = (n + kTile - 1)/kTile;
	cudaError_t e = cudaFuncSetAttribute(sirenBackwardFusedTc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e) return fail(cudaGetErrorString(e));
	// thread-block clusters for the gradient reduction (batches that fill the GPU): the largest size the device can co-schedule
	static int clusterMax = -1;
	if (clusterMax < 0) {
		clusterMax = 1;
		const char* envc = getenv("NMC_FUSED_BWD_CLUSTER");
		const int want = envc ? atoi(envc) : 4;   // 8: the 128 CTAs of a 16384 batch no longer fit in one wave (82 us against 44)
		for (int c = want; c > 1; c >>= 1) {
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3((unsigned)(smCount()/c*c)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
			cfg.attrs = at; cfg.numAttrs = 1;
			int nc = 0;
			if (cudaOccupancyMaxActiveClusters(&nc, sirenBackwardFusedTc, &cfg) == cudaSuccess && nc*c >= 96) { clusterMax = c; break; }
			cudaGetLastError();
		}
	}
	int cluster = tiles >= 64 ? clusterMax : 1;
	long long cap = cluster > 1 ? (long long)(smCount()/cluster)*cluster : smCount();
	long long want = cluster > 1 ? (tiles + cluster - 1)/cluster*cluster : tiles;
	const int grid = (int)(want < cap ? want : cap);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
	cudaLaunchAttribute at[2];
	at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	at[1] = nmc_pdl::attribute();
	cfg.attrs = at; cfg.numAttrs = nmc_pdl::enabled() ? 2 : 1;
	e = cudaLaunchKernelEx(&cfg, sirenBackwardFusedTc, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, z_saved, grad_y, cluster);
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}
