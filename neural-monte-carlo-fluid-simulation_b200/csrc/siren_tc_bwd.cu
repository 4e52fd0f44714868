// csrc/siren_tc_bwd.cu -- tensor-core (tcgen05 / TMEM) backward pass of the SIREN velocity-network fit for sm_100a:
// the delta chain and the weight gradients of update_network (src/2d/models/base.py:83-96, loss.backward() through
// src/2d/models/networks.py:47-57), i.e. the two dense contractions of the backward pass
//     dA_{l-1} = dZ_l W_l                    (samples x H_out) . (H_out x H_in)         sirenBackwardTc
//     dW_l     = dZ_l^T A_{l-1}              (H_out x samples) . (samples x H_in)       sirenWeightGradTc
// Both run as 3xTF32 (hi/lo split of both operands, three MMAs per K step: hi.hi + hi.lo + lo.hi) so the gradients keep
// fp32-level accuracy (the fit runs at lr = 1e-5 with Adam; parity is against torch's fp32 autograd).
//
// sirenBackwardTc mirrors the forward kernel (siren_tc.cu): a CTA owns 128 samples (thread <-> TMEM lane <-> sample), the
// deltas dZ_l live in shared memory as the K-major A operand, W_l^T is staged 64 input neurons at a time as the B operand
// (next chunk prefetched into registers during the MMAs), the epilogue multiplies the accumulator row by
// w0 cos(w0 z_{l-1}) (pre-activations saved by the training forward, prefetched during the MMAs), writes dZ_{l-1} to
// global memory for the weight-gradient kernel and back into shared memory as the next operand.
//
// sirenWeightGradTc: grid (sample chunks, layers).  Hidden layers: both operands are staged K-major with K = samples
// straight from the [neuron][sample] global layout (A_{l-1} = sin(w0 z_{l-1}) is recomputed while staging, so the
// activations never travel through global memory), 64 samples per MMA batch accumulated in one TMEM tile of H x H;
// the tile is added to the gradient buffer with 16-byte vector reductions.  First / last layer (K = 2|3 or N = 2|3) and
// the bias sums are not GEMM-shaped and run on the FMA pipe in the same launch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include "../../include/nmcfs_siren.h"
#include "siren_env.cuh"
#include "siren_tc.cuh"
#include "pdl.cuh"

namespace nmc_siren_detail { void setError(const char* m); }

namespace {

using namespace nmc_siren_tc;
using nmc_siren_detail::Env;

constexpr int kTile = 128;        // samples per CTA tile of the delta chain
constexpr int kThreads = 256;     // two threads per accumulator row
constexpr int kMaxLayers = 18;
constexpr int kNChunk = 64;       // UMMA N of the delta chain: input neurons per staged chunk of W^T

struct Params {
	const float* W[kMaxLayers];
	const float* b[kMaxLayers];
	float* gW[kMaxLayers];
	float* gb[kMaxLayers];
};

// -DNMC_TC_TRACE: thread 0 of the first CTA of layer 0 (FMA path, slot 0) and of layer 1 (tensor-core path, slot 1), and of the delta chain (slot 2), stamps
// clock64() at its phase boundaries (profiles/tools/tc_trace.py)
#ifdef NMC_TC_TRACE
__device__ long long g_traceW[3][256];
__device__ int g_traceWN[3];
#define TRACEW(slot, tag) do { if (blockIdx.x == 0 && tid == 0 && tn < 127) { g_traceW[slot][2*tn] = (tag); g_traceW[slot][2*tn + 1] = clock64(); tn++; g_traceWN[slot] = tn; } } while (0)
#else
#define TRACEW(slot, tag) do {} while (0)
#endif

// ---- delta chain ---------------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(kThreads, H == 64 ? 2 : 1)
sirenBackwardTc(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
				const float* __restrict__ zSaved, const float* __restrict__ gy, float* __restrict__ dZ) {
	nmc_pdl::gridEnter();
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned char* Dhi = smem;                       // dZ_l [128 samples x H neurons], K-major (K = neurons of layer l)
	unsigned char* Dlo = Dhi + kTile*H*4;
	unsigned char* Bhi = Dlo + kTile*H*4;            // W_l^T chunk [64 input neurons x H output neurons], K-major
	unsigned char* Blo = Bhi + kNChunk*H*4;
	float* sWL = reinterpret_cast<float*>(Blo + kNChunk*H*4);   // last layer's weights [outDim][H]: warp-wide broadcasts
	__shared__ __align__(8) unsigned long long mbar;
	__shared__ uint32_t tmemBaseSh;
	const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), half = tid >> 7;
	const int cBeg = half*(H/2);
	constexpr int HC = H/2;                          // columns per thread

	if (warp == 0) tmemAlloc(&tmemBaseSh, (uint32_t)H);
	if (tid == 0) mbarInit(smemAddr(&mbar), 1);
	for (int i = tid; i < outDim*H; i += kThreads) sWL[i] = __ldg(&P.W[nHidden + 1][i]);
	fenceBeforeSync();
	__syncthreads();
	fenceAfterSync();
	const uint32_t tmemBase = tmemBaseSh;
	const uint32_t barAddr = smemAddr(&mbar);
	const uint32_t idesc = instrDescTf32(kTile, kNChunk);
	uint32_t phase = 0;

	// B operand of dA = dZ_l W_l:  B(nrow = input neuron, k = output neuron) = W_l[k][nrow]; a thread gathers four
	// consecutive k of one input neuron (lanes walk the input neurons: coalesced rows of W_l) into one 16-byte word
	constexpr int RW = kNChunk*H/4/kThreads;
	float4 wreg[RW];
	auto loadW = [&](int l, int nc) {
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/kNChunk, r = idx - k4*kNChunk;
			const float* w = &P.W[l][(size_t)(4*k4)*H + nc*kNChunk + r];
			wreg[i] = make_float4(__ldg(w), __ldg(w + H), __ldg(w + 2*H), __ldg(w + 3*H));
		}
	};
	auto storeW = [&]() {
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/kNChunk, r = idx - k4*kNChunk;
			float4 h, o;
			splitTf32(wreg[i], h, o);
			const int off = coreOffsetBytes<H>(r, 4*k4);
			*reinterpret_cast<float4*>(Bhi + off) = h;
			*reinterpret_cast<float4*>(Blo + off) = o;
		}
	};
	if (nHidden >= 1 && (long long)blockIdx.x*kTile < n) loadW(nHidden, 0);

#ifdef NMC_TC_TRACE
	int tn = 0;
#endif
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x) {
		const long long s = tile*kTile + row;
		const bool live = s < n;
		// rows past the end read the last sample (always in bounds, no per-element branch); their results are zeroed / not stored
		const long long sc = live ? s : n - 1;
		const float* zrow = zSaved + sc + (size_t)cBeg*n;   // + layer*H*n: this thread's first column, stride n per column
		float* drow = dZ + sc + (size_t)cBeg*n;
		TRACEW(2, 1);
		float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
		if (live) {
			g0 = gy[s*outDim]; if (outDim > 1) g1 = gy[s*outDim + 1]; if (outDim > 2) g2 = gy[s*outDim + 2];
			if (env.active) { // dL/d(network output) = dL/d(enveloped output) x (detached) envelope weights
				const float xs[3] = {x[s*inDim], inDim > 1 ? x[s*inDim + 1] : 0.0f, inDim > 2 ? x[s*inDim + 2] : 0.0f};
				float gys[3] = {g0, g1, g2};
				nmc_siren_detail::envBackward(env, inDim, outDim, xs, nullptr, gys, nullptr);
				g0 = gys[0]; g1 = gys[1]; g2 = gys[2];
			}
			if (half == 0) { // rows (L+1)*H .. of dZ: the scaled output gradient, for dW_last = gy'^T A_L^T
				const size_t r0 = (size_t)(nHidden + 1)*H;
				dZ[(r0 + 0)*n + s] = g0;
				if (outDim > 1) dZ[(r0 + 1)*n + s] = g1;
				if (outDim > 2) dZ[(r0 + 2)*n + s] = g2;
			}
		}
		float zreg[HC];
		{ // dZ_L = (W_last^T gy') * w0 cos(w0 z_L) on the FMA pipe, written straight into the A operand
			{
				const float* zp = zrow + (size_t)nHidden*H*n;
#pragma unroll
				for (int q = 0; q < HC; q++) { zreg[q] = __ldg(zp); zp += n; }
			}
			float* dp = drow + (size_t)nHidden*H*n;
			if (!live) { g0 = 0.0f; g1 = 0.0f; g2 = 0.0f; }
#pragma unroll
			for (int q4 = 0; q4 < HC; q4 += 4) {
				float d[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					const int c = cBeg + q4 + q;
					float a = sWL[c]*g0;
					if (outDim > 1) a += sWL[H + c]*g1;
					if (outDim > 2) a += sWL[2*H + c]*g2;
					a = a*w0*cosReduced(w0*zreg[q4 + q]);
					if (live) *dp = a;
					dp += n;
					d[q] = a;
				}
				float4 h, o;
				splitTf32(make_float4(d[0], d[1], d[2], d[3]), h, o);
				const int off = coreOffsetBytes<H>(row, cBeg + q4);
				*reinterpret_cast<float4*>(Dhi + off) = h;
				*reinterpret_cast<float4*>(Dlo + off) = o;
			}
		}
		for (int l = nHidden; l >= 1; l--) {
			// pre-activations of layer l - 1 for this thread's columns: in flight during the MMAs
			{
				const float* zp = zrow + (size_t)(l - 1)*H*n;
#pragma unroll
				for (int q = 0; q < HC; q++) { zreg[q] = __ldg(zp); zp += n; }
			}
			for (int nc = 0; nc < H/kNChunk; nc++) {
				TRACEW(2, 2);
				storeW();
				TRACEW(2, 3);
				fenceProxyAsync();
				fenceBeforeSync();
				__syncthreads();
				TRACEW(2, 4);
				{ // the chunk that follows: this layer, the layer below, or the first one of the next tile (requested before the
				  // MMA issue so that thread 0's share is not late)
					int ln = l, ncn = nc + 1;
					if (ncn == H/kNChunk) { ncn = 0; ln = l - 1; }
					if (ln >= 1) loadW(ln, ncn);
					else if ((tile + gridDim.x)*kTile < n) loadW(nHidden, 0);
				}
				if (tid == 0) {
					fenceAfterSync();
					const uint32_t d = tmemBase + (uint32_t)(nc*kNChunk);
					const uint32_t aH = smemAddr(Dhi), aL = smemAddr(Dlo), bH = smemAddr(Bhi), bL = smemAddr(Blo);
					const uint32_t sbo = H*32;
					const uint64_t dAh = smemDesc(aH, 128, sbo), dAl = smemDesc(aL, 128, sbo);
					const uint64_t dBh = smemDesc(bH, 128, sbo), dBl = smemDesc(bL, 128, sbo);
#pragma unroll
					for (int ks = 0; ks < H/8; ks++) { // one K step = 256 bytes = 16 descriptor address units
						mmaTf32(d, dAh + 16*ks, dBh + 16*ks, idesc, ks > 0 ? 1u : 0u);
						mmaTf32(d, dAh + 16*ks, dBl + 16*ks, idesc, 1u);
						mmaTf32(d, dAl + 16*ks, dBh + 16*ks, idesc, 1u);
					}
					mmaCommit(barAddr);
				}
				TRACEW(2, 5);
				mbarWait(barAddr, phase);
				phase ^= 1u;
				fenceAfterSync();
				TRACEW(2, 6);
			}
			// epilogue: dZ_{l-1} = dA_{l-1} * w0 cos(w0 z_{l-1}) -> global (weight gradients) and the next A operand
			float* dp = drow + (size_t)(l - 1)*H*n;
			const float w0l = live ? w0 : 0.0f;   // dead rows: zero deltas
			uint32_t v[HC]; // the whole half row in flight, one wait
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 16) tmemLoad16Async(tmemBase + ((uint32_t)((warp & 3)*32) << 16) + (uint32_t)(cBeg + c0), &v[c0]);
			tmemLoadWait();
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 4) {
				float d[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					const float a = __uint_as_float(v[c0 + q])*w0l*cosReduced(w0*zreg[c0 + q]);
					if (live) *dp = a;
					dp += n;
					d[q] = a;
				}
				if (l > 1) {
					float4 h, o;
					splitTf32(make_float4(d[0], d[1], d[2], d[3]), h, o);
					const int off = coreOffsetBytes<H>(row, cBeg + c0);
					*reinterpret_cast<float4*>(Dhi + off) = h;
					*reinterpret_cast<float4*>(Dlo + off) = o;
				}
			}
		}
		TRACEW(2, 7);
		// the next tile's first deltas overwrite D: every thread is past its last use (MMAs completed via the mbarrier)
		fenceBeforeSync();
		__syncthreads();
	}
	fenceBeforeSync();
	__syncthreads();
	if (warp == 0) tmemFree(tmemBase, (uint32_t)H);
}

// ---- weight gradients ----------------------------------------------------------------------------------------------------

constexpr int kKS = 64;           // samples per MMA batch (K of one staged operand tile)
constexpr int kGS = 32;           // samples per staged tile of the FMA path (first / last layer)

template <int H>
__global__ void __launch_bounds__(kThreads)
sirenWeightGradTc(Params P, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n,
				  const float* __restrict__ dZ, const float* __restrict__ zSaved, int chunkTc, int chunkFma) {
	nmc_pdl::gridEnter();
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ __align__(8) unsigned long long mbar;
	__shared__ uint32_t tmemBaseSh;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int l = blockIdx.y;                 // 0 .. nHidden + 1
	const int last = nHidden + 1;
	// the grid's x extent covers the finer of the two partitions; the CTAs beyond a layer's own chunk count leave at once
	const int chunk = (l == 0 || l == last) ? chunkFma : chunkTc;
	const long long s0 = (long long)blockIdx.x*chunk;
	if (s0 >= n) return;
	const long long s1 = s0 + chunk < n ? s0 + chunk : n;
#ifdef NMC_TC_TRACE
	int tn = 0;
#endif

	if (l == 0 || l == last) { // ---- FMA path: C[i][j] = sum_s P[i][s] Q[j][s] with 2-3 rows on one side; bias sums
		constexpr int LDS = H + 4;
		float (*Ps)[LDS] = reinterpret_cast<float (*)[LDS]>(smem);
		float (*Qs)[LDS] = reinterpret_cast<float (*)[LDS]>(smem + kGS*LDS*4);
		const int RP = l == last ? outDim : H;    // rows of P
		const int RQ = l == 0 ? inDim : H;        // rows of Q
		const float* Pg = l == last ? dZ + (size_t)(nHidden + 1)*H*n : dZ;
		const float* Zg = l == 0 ? nullptr : zSaved + (size_t)(l - 1)*H*n;
		float small[(3*H + kThreads - 1)/kThreads];
#pragma unroll
		for (int q = 0; q < (3*H + kThreads - 1)/kThreads; q++) small[q] = 0.0f;
		float bsum = 0.0f;
		for (long long sb = s0; sb < s1; sb += kGS) {
			const int ns = (int)(s1 - sb < kGS ? s1 - sb : kGS);
			if (l == 0) TRACEW(0, 1);
			__syncthreads();
			for (int idx = tid; idx < RP*kGS; idx += kThreads) { int r = idx/kGS, c = idx - r*kGS; Ps[c][r] = c < ns ? Pg[(size_t)r*n + sb + c] : 0.0f; }
			if (l == 0) { for (int idx = tid; idx < kGS*inDim; idx += kThreads) { int c = idx/inDim, r = idx - c*inDim; Qs[c][r] = c < ns ? x[(sb + c)*inDim + r] : 0.0f; } }
			else { for (int idx = tid; idx < RQ*kGS; idx += kThreads) { int r = idx/kGS, c = idx - r*kGS; Qs[c][r] = c < ns ? sinReduced(w0*Zg[(size_t)r*n + sb + c]) : 0.0f; } }
			__syncthreads();
			int q = 0;
			for (int o = tid; o < RP*RQ; o += kThreads, q++) {
				int i = o/RQ, j = o - i*RQ;
				float a = 0.0f;
				for (int c = 0; c < kGS; c++) a += Ps[c][i]*Qs[c][j];
				small[q] += a;
			}
			if (tid < RP) { float a = 0.0f; for (int c = 0; c < kGS; c++) a += Ps[c][tid]; bsum += a; }
		}
		int q = 0;
		if (l == 0) TRACEW(0, 2);
		for (int o = tid; o < RP*RQ; o += kThreads, q++) atomicAdd(&P.gW[l][o], small[q]);
		if (tid < RP) atomicAdd(&P.gb[l][tid], bsum);
		if (l == 0) TRACEW(0, 3);
		return;
	}

	// ---- tensor-core path: dW_l [H x H] += dZ_l[:, chunk] . A_{l-1}[:, chunk]^T --------------------------------------------
	constexpr int OPB = H*kKS*4;              // bytes of one operand tile [H rows x 64 samples]
	unsigned char* Phi = smem;
	unsigned char* Plo = Phi + OPB;
	unsigned char* Qhi = Plo + OPB;
	unsigned char* Qlo = Qhi + OPB;
	if (warp == 0) tmemAlloc(&tmemBaseSh, (uint32_t)H);
	if (tid == 0) mbarInit(smemAddr(&mbar), 1);
	fenceBeforeSync();
	__syncthreads();
	fenceAfterSync();
	const uint32_t tmemBase = tmemBaseSh;
	const uint32_t barAddr = smemAddr(&mbar);
	const uint32_t idesc = instrDescTf32(H, H);
	const float* Pg = dZ + (size_t)l*H*n;
	const float* Zg = zSaved + (size_t)(l - 1)*H*n;

	// staging map: 32 consecutive work items = 8 rows x 4 float4 (64 bytes of one row): conflict-free 16-byte stores into
	// the core-matrix layout, fully used 32-byte sectors on the global side
	constexpr int RV = H*(kKS/4)/kThreads;    // float4 per thread and operand: 4 (H = 64), 8 (H = 128)
	float4 preg[RV], qreg[RV];
	float bacc[RV];
#pragma unroll
	for (int i = 0; i < RV; i++) bacc[i] = 0.0f;
	auto itemRow = [&](int i) { const int idx = tid + i*kThreads; return ((idx >> 7) << 3) | (idx & 7); };
	auto itemK4 = [&](int i) { const int idx = tid + i*kThreads; return (((idx >> 5) & 3) << 2) | ((idx >> 3) & 3); };
	auto loadStage = [&](long long sb) {
#pragma unroll
		for (int i = 0; i < RV; i++) {
			const int r = itemRow(i), k4 = itemK4(i);
			const long long s = sb + 4*k4;
			if (s < s1) { // n and the chunk bounds are multiples of 4: a 16-byte word is inside or outside as a whole
				preg[i] = __ldg(reinterpret_cast<const float4*>(&Pg[(size_t)r*n + s]));
				qreg[i] = __ldg(reinterpret_cast<const float4*>(&Zg[(size_t)r*n + s]));
			} else {
				preg[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
				qreg[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); // sin(0) = 0: no contribution
			}
		}
	};
	auto storeStage = [&]() {
#pragma unroll
		for (int i = 0; i < RV; i++) {
			const int r = itemRow(i), k4 = itemK4(i);
			const int off = coreOffsetBytes<kKS>(r, 4*k4);
			float4 h, o;
			splitTf32(preg[i], h, o);
			*reinterpret_cast<float4*>(Phi + off) = h;
			*reinterpret_cast<float4*>(Plo + off) = o;
			bacc[i] += (preg[i].x + preg[i].y) + (preg[i].z + preg[i].w);
			const float4 a = make_float4(sinReduced(w0*qreg[i].x), sinReduced(w0*qreg[i].y), sinReduced(w0*qreg[i].z), sinReduced(w0*qreg[i].w));
			splitTf32(a, h, o);
			*reinterpret_cast<float4*>(Qhi + off) = h;
			*reinterpret_cast<float4*>(Qlo + off) = o;
		}
	};
	uint32_t phase = 0;
	int stage = 0;
	if (l == 1) TRACEW(1, 1);
	loadStage(s0);
	for (long long sb = s0; sb < s1; sb += kKS, stage++) {
		if (l == 1) TRACEW(1, 2);
		if (stage > 0) { // the MMAs of the previous batch have finished reading the operand tiles
			mbarWait(barAddr, phase);
			phase ^= 1u;
			fenceAfterSync();
		}
		if (l == 1) TRACEW(1, 3);
		storeStage();
		if (l == 1) TRACEW(1, 4);
		fenceProxyAsync();
		fenceBeforeSync();
		__syncthreads();
		if (l == 1) TRACEW(1, 5);
		if (sb + kKS < s1) loadStage(sb + kKS); // in flight during the MMAs (requested before the issue: thread 0's share is not late)
		if (tid == 0) {
			fenceAfterSync();
			const uint32_t pH = smemAddr(Phi), pL = smemAddr(Plo), qH = smemAddr(Qhi), qL = smemAddr(Qlo);
			const uint32_t sbo = kKS*32;
			const uint64_t dPh = smemDesc(pH, 128, sbo), dPl = smemDesc(pL, 128, sbo);
			const uint64_t dQh = smemDesc(qH, 128, sbo), dQl = smemDesc(qL, 128, sbo);
#pragma unroll
			for (int ks = 0; ks < kKS/8; ks++) { // one K step = 256 bytes = 16 descriptor address units
				mmaTf32(tmemBase, dPh + 16*ks, dQh + 16*ks, idesc, (stage > 0 || ks > 0) ? 1u : 0u);
				mmaTf32(tmemBase, dPh + 16*ks, dQl + 16*ks, idesc, 1u);
				mmaTf32(tmemBase, dPl + 16*ks, dQh + 16*ks, idesc, 1u);
			}
			mmaCommit(barAddr);
		}
		if (l == 1) TRACEW(1, 6);
	}
	if (l == 1) TRACEW(1, 7);
	mbarWait(barAddr, phase);
	fenceAfterSync();
	if (l == 1) TRACEW(1, 8);
	{ // accumulator tile -> gradient buffer.  M = 128: row i in TMEM lane i; M = 64: row i in lane (i % 16) + 32 (i / 16)
		const int sp = warp & 3, ch = warp >> 2;  // TMEM sub-partition of this warp, column half
		const int i = H == 128 ? sp*32 + lane : sp*16 + (lane & 15);
		const bool valid = H == 128 || lane < 16;
		float* gw = P.gW[l] + (size_t)i*H;
		for (int c0 = ch*(H/2); c0 < (ch + 1)*(H/2); c0 += 16) {
			uint32_t v[16];
			tmemLoad16(tmemBase + ((uint32_t)(sp*32) << 16) + (uint32_t)c0, v);
			if (valid) {
#pragma unroll
				for (int q = 0; q < 16; q += 4)
					redAdd4(gw + c0 + q, __uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]), __uint_as_float(v[q + 3]));
			}
		}
	}
	{ // bias gradient: row sums of dZ_l over this CTA's samples; the four lanes (k4 low bits) of a row first
#pragma unroll
		for (int i = 0; i < RV; i++) {
			float b = bacc[i];
			b += __shfl_xor_sync(0xffffffffu, b, 8);
			b += __shfl_xor_sync(0xffffffffu, b, 16);
			if (lane < 8) atomicAdd(&P.gb[l][itemRow(i)], b);
		}
	}
	if (l == 1) TRACEW(1, 9);
	fenceBeforeSync();
	__syncthreads();
	if (warp == 0) tmemFree(tmemBase, (uint32_t)H);
}

bool shapeOk(const nmc_siren_shape* sh) {
	return sh && (sh->hidden == 64 || sh->hidden == 128) && sh->n_hidden_layers >= 1 && sh->n_hidden_layers + 2 <= kMaxLayers &&
		   sh->in_dim >= 1 && sh->in_dim <= 3 && sh->out_dim >= 1 && sh->out_dim <= 3;
}
int smCount() {
	static int sms = 0;
	if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
	return sms;
}
int fail(const char* m) { nmc_siren_detail::setError(m); return 1; }

} // namespace

extern "C" int nmc_siren_backward_tc(const nmc_siren_shape* sh, const float* const* W, const float* const* b, const float* x,
									 int64_t n, const float* z_saved, const float* grad_y, float* dZ,
									 const nmc_siren_envelope* envp, void* stream) {
	if (!sh || !W || !b) return fail("null argument");
	if (!shapeOk(sh)) return fail("tensor-core backward: unsupported shape (hidden 64|128, >= 1 hidden layer, in/out 1..3)");
	if (n <= 0) return 0;
	if (!x || !z_saved || !grad_y || !dZ) return fail("null buffer");
	Params P;
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) { P.W[l] = W[l]; P.b[l] = b[l]; P.gW[l] = nullptr; P.gb[l] = nullptr; }
	Env env;
	if (const char* bad = nmc_siren_detail::toEnv(envp, env)) return fail(bad);
	const int H = sh->hidden;
	const size_t smem = (size_t)(2*kTile*H + 2*kNChunk*H)*4 + (size_t)sh->out_dim*H*4;
	const long long tiles = (n + kTile - 1)/kTile;
	const int perSM = H == 64 ? 2 : 1;
	const int grid = (int)(tiles < (long long)perSM*smCount() ? tiles : (long long)perSM*smCount());
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e;
	if (H == 64) {
		e = cudaFuncSetAttribute(sirenBackwardTc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) e = nmc_pdl::launch(sirenBackwardTc<64>, dim3(grid), dim3(kThreads), smem, st, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, z_saved, grad_y, dZ);
	} else {
		e = cudaFuncSetAttribute(sirenBackwardTc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) e = nmc_pdl::launch(sirenBackwardTc<128>, dim3(grid), dim3(kThreads), smem, st, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, z_saved, grad_y, dZ);
	}
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

extern "C" int nmc_siren_weight_grads_tc(const nmc_siren_shape* sh, const float* x, int64_t n, const float* dZ, const float* z_saved,
										 float* const* gW, float* const* gb, void* stream) {
	if (!sh || !gW || !gb) return fail("null argument");
	if (!shapeOk(sh)) return fail("tensor-core weight gradients: unsupported shape (hidden 64|128, >= 1 hidden layer, in/out 1..3)");
	if (n <= 0) return 0;
	if (n % 4) return fail("tensor-core weight gradients: the batch size must be a multiple of 4 (16-byte operand loads)");
	if (!x || !dZ || !z_saved) return fail("null buffer");
	Params P;
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) {
		P.W[l] = nullptr; P.b[l] = nullptr; P.gW[l] = gW[l]; P.gb[l] = gb[l];
		if (!gW[l] || !gb[l]) return fail("null layer pointer");
		if (l >= 1 && l <= sh->n_hidden_layers && ((uintptr_t)gW[l] & 15)) return fail("tensor-core weight gradients: hidden-layer gradient buffers must be 16-byte aligned");
	}
	if (((uintptr_t)dZ & 15) || ((uintptr_t)z_saved & 15)) return fail("tensor-core weight gradients: dZ and z_saved must be 16-byte aligned");
	const int H = sh->hidden;
	// samples per CTA: as large as possible (fewer reductions into the gradient buffer) while the hidden layers' CTAs cover the SMs
	// (measured, batch 16384: H = 64 (64 KB per CTA, three resident per SM, which hide each other's load -> stage -> MMA
	// round trips) is fastest with ~2 CTAs per SM; H = 128 (128 KB per CTA, one per SM) with a single wave)
	int chunk = 1024;
	const long long want = H == 64 ? 2ll*smCount() : (3ll*smCount())/4;
	while (chunk > 2*kKS && ((n + chunk - 1)/chunk)*sh->n_hidden_layers < want) chunk >>= 1;
	if (const char* ov = getenv("NMC_WGRAD_CHUNK")) { const int v = atoi(ov); if (v >= kKS && v % kKS == 0) chunk = v; } // A/B measurements
	// first / last layer (FMA path, 32-sample tiles with a full load -> barrier -> compute round trip each): short chunks,
	// so that these CTAs finish with the tensor-core ones (phase trace: 3000 cycles per tile)
	const int chunkFma = chunk < 128 ? chunk : 128;
	dim3 grid((unsigned)((n + chunkFma - 1)/chunkFma), (unsigned)(sh->n_hidden_layers + 2));
	const size_t smemTc = (size_t)4*H*kKS*4, smemFma = (size_t)2*kGS*(H + 4)*4;
	const size_t smem = smemTc > smemFma ? smemTc : smemFma;
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e;
	if (H == 64) {
		e = cudaFuncSetAttribute(sirenWeightGradTc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) e = nmc_pdl::launch(sirenWeightGradTc<64>, grid, dim3(kThreads), smem, st, P, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, dZ, z_saved, chunk, chunkFma);
	} else {
		e = cudaFuncSetAttribute(sirenWeightGradTc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (!e) e = nmc_pdl::launch(sirenWeightGradTc<128>, grid, dim3(kThreads), smem, st, P, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, dZ, z_saved, chunk, chunkFma);
	}
	if (!e) e = cudaGetLastError();
	return e ? fail(cudaGetErrorString(e)) : 0;
}

#ifdef NMC_TC_TRACE
extern "C" int nmc_siren_trace_read_wgrad(int slot, long long* out, int cap) { // (tag, clock) pairs of the last traced launch
	int n[3] = {0, 0, 0};
	if (slot < 0 || slot > 2) return 0;
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(n, g_traceWN, sizeof(n));
	int k = n[slot] > cap ? cap : n[slot];
	cudaMemcpyFromSymbol(out, g_traceW, sizeof(long long)*2*k, sizeof(long long)*256*slot);
	return k;
}
#endif
