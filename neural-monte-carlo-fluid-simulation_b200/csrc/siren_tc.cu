// csrc/siren_tc.cu -- tensor-core (tcgen05 / TMEM) forward pass of the SIREN velocity network for sm_100a.
//
// The hidden layers  A_l = sin(w0 (A_{l-1} W_l^T + b_l))  are the only dense contractions on the hot path
// (src/2d/models/networks.py:47-57).  One CTA of 128 threads owns a tile of 128
// samples: thread t <-> sample row t <-> TMEM lane t.
//   * operands live in shared memory in the K-major, un-swizzled UMMA canonical layout (8 x 16-byte core
//     matrices); activations A [128 x H] and one 64-row chunk of W_l [64 x H] at a time;
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 64, K = 8) into a TMEM
//     accumulator of H fp32 columns and commits to an mbarrier; everybody else waits on the barrier;
//   * fp32 accuracy is recovered with the 3xTF32 split  A = A_hi + A_lo, W = W_hi + W_lo,
//     D = A_hi W_hi + A_hi W_lo + A_lo W_hi  (SIREN's sin(30 .) amplifies TF32 rounding by 30, and the fit
//     runs at lr = 1e-5: plain TF32 is not accurate enough);
//   * the epilogue reads each thread's row with tcgen05.ld (32x32b), adds the bias, applies sin(w0 .),
//     splits the result and writes it back as the next layer's A operand -- activations never leave the SM.
// The first (in -> H, K = 2|3) and last (H -> out, N = 2|3) layers are not GEMM-shaped and run on the FMA pipe
// inside the same kernel.  No TMA: a layer's weights are 16-64 KB, L2-resident and re-split on load.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nmcfs_siren.h"
#include "siren_env.cuh"
#include "siren_tc.cuh"
#include "pdl.cuh"

namespace {

using nmc_siren_tc::sinReduced;

constexpr int kTile = 128;
constexpr int kMaxLayers = 18;

// -DNMC_TC_TRACE: CTA 0 / thread 0 stamps clock64() at the phase boundaries of its first tile (profiles/tools/tc_trace.py)
#ifdef NMC_TC_TRACE
__device__ long long g_trace[256];
__device__ int g_traceN;
#define TRACE(tag) do { if (blockIdx.x == 0 && tid == 0 && tile == blockIdx.x && tn < 127) { g_trace[2*tn] = (tag); g_trace[2*tn + 1] = clock64(); tn++; g_traceN = tn; } } while (0)
#else
#define TRACE(tag) do {} while (0)
#endif
constexpr int kNChunk = 64;   // output neurons per MMA group (UMMA N)

struct Params {
	const float* W[kMaxLayers];
	const float* b[kMaxLayers];
};

using nmc_siren_detail::Env;


__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (r, k) of an operand with K columns (fp32/tf32, 4 per 16 bytes)
template <int K>
__device__ __forceinline__ int coreOffsetBytes(int r, int k) {
	return (r >> 3)*(K*32) + (k >> 2)*128 + (r & 7)*16 + (k & 3)*4;
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, mma_sm100_desc.hpp:98-123)
__device__ __forceinline__ uint64_t smemDesc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
	return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lboBytes >> 4) << 16) | ((uint64_t)(sboBytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mmaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
				 :: "r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
					 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void splitTf32(float v, float& hi, float& lo) {
	hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); // the 10 mantissa bits the tensor core keeps
	lo = v - hi;
}

// Two threads per accumulator row: thread `row` and thread `row + 128` split the columns of the row between them in
// the first layer and in the epilogues (warps w and w + 4 read the same 32 TMEM lanes), and all 256 threads stage
// the weight operand, so the serial part between two MMA batches is half as long.
constexpr int kThreads = 2*kTile;

// Everything that is not a GEMM operand -- every bias, the first layer's weights (H x in), the last layer's (out x H) and
// its bias -- is staged once per CTA in shared memory and read as warp-wide broadcasts: the epilogues used to fetch
// them with one dependent global load per element (phase trace: 5500 of the 8400 cycles of a hidden layer).
__host__ __device__ inline int smallParamFloats(int H, int nHidden, int inDim, int outDim) {
	return (nHidden + 1)*H + 4*H + outDim*H + outDim + 0*inDim;
}

template <int H, bool SAVEZ>
__global__ void __launch_bounds__(kThreads, H == 64 ? 2 : 1)
sirenForwardTc(Params P, Env env, int inDim, int outDim, int nHidden, float w0, const float* __restrict__ x, long long n, float* __restrict__ y,
			   float* __restrict__ zSaved) {
	nmc_pdl::gridEnter();
	extern __shared__ __align__(128) unsigned char smem[];
	unsigned char* Ahi = smem;
	unsigned char* Alo = Ahi + kTile*H*4;
	unsigned char* Bhi = Alo + kTile*H*4;
	unsigned char* Blo = Bhi + kNChunk*H*4;
	float* sBias = reinterpret_cast<float*>(Blo + kNChunk*H*4);   // [nHidden + 1][H]
	float4* sW0 = reinterpret_cast<float4*>(sBias + (nHidden + 1)*H); // [H] (w_0, w_1, w_2, bias) of the first layer: one 16-byte broadcast per neuron
	float* sWL = reinterpret_cast<float*>(sW0 + H);                // [outDim][H]
	float* sbL = sWL + outDim*H;                                   // [outDim]
	__shared__ __align__(8) unsigned long long mbar;
	__shared__ uint32_t tmemBaseSh;
	__shared__ float ypart[3][kTile];
	const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kTile - 1), half = tid >> 7;
	const int cBeg = half*(H/2);
	constexpr int HC = H/2;
	const int last = nHidden + 1;

	if (warp == 0) { // TMEM: H fp32 columns x 128 lanes
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(&tmemBaseSh)), "r"((uint32_t)H) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smemAddr(&mbar)) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int l = 0; l <= nHidden; l++) for (int i = tid; i < H; i += kThreads) sBias[l*H + i] = __ldg(&P.b[l][i]);
	for (int i = tid; i < H; i += kThreads)
		sW0[i] = make_float4(__ldg(&P.W[0][i*inDim]), inDim > 1 ? __ldg(&P.W[0][i*inDim + 1]) : 0.0f, inDim > 2 ? __ldg(&P.W[0][i*inDim + 2]) : 0.0f, __ldg(&P.b[0][i]));
	for (int i = tid; i < outDim*H; i += kThreads) sWL[i] = __ldg(&P.W[last][i]);
	if (tid < outDim) sbL[tid] = __ldg(&P.b[last][tid]);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmemBase = tmemBaseSh;
	const uint32_t barAddr = smemAddr(&mbar);
	// instruction descriptor (cute::UMMA::InstrDescriptor, mma_sm100_desc.hpp:412-439):
	// D = F32, A = B = TF32, both K-major, N = 64, M = 128
	const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNChunk >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
	uint32_t phase = 0;
	// Weight operand pipeline: the next 64-row chunk of W is fetched from L2 into registers while the tensor core works
	// on the current one (the single B buffer can only be rewritten once those MMAs have completed), so the
	// global-load latency no longer sits between two MMA batches.
	constexpr int RW = kNChunk*H/4/kThreads;
	float4 wreg[RW];
	auto loadW = [&](int l, int nc) {
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/kNChunk, r = idx - k4*kNChunk; // lanes walk down the rows of one core-matrix column
			wreg[i] = __ldg(reinterpret_cast<const float4*>(&P.W[l][(size_t)(nc*kNChunk + r)*H + 4*k4]));
		}
	};
	auto storeW = [&]() { // hi/lo TF32 split into the K-major core-matrix layout
#pragma unroll
		for (int i = 0; i < RW; i++) {
			const int idx = tid + i*kThreads, k4 = idx/kNChunk, r = idx - k4*kNChunk;
			float4 h, o;
			splitTf32(wreg[i].x, h.x, o.x); splitTf32(wreg[i].y, h.y, o.y); splitTf32(wreg[i].z, h.z, o.z); splitTf32(wreg[i].w, h.w, o.w);
			const int off = coreOffsetBytes<H>(r, 4*k4);
			*reinterpret_cast<float4*>(Bhi + off) = h;
			*reinterpret_cast<float4*>(Blo + off) = o;
		}
	};
	if (nHidden >= 1 && (long long)blockIdx.x*kTile < n) loadW(1, 0);

#ifdef NMC_TC_TRACE
	int tn = 0;
#endif
	for (long long tile = blockIdx.x; tile*kTile < n; tile += gridDim.x) {
		const long long s = tile*kTile + row;
		const bool live = s < n;
		TRACE(1);
		float* zrow = SAVEZ ? zSaved + s : nullptr;
		float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f;
		// activation of this thread's columns c .. c + 3 of layer l from the pre-activations z: optional save, sine, the last
		// layer's dot products (when l is the last hidden layer) or the hi/lo operand words of the next GEMM
		auto emit4 = [&](int l, int c, const float (&z)[4], bool lastHidden) {
			float hi[4], lo[4];
#pragma unroll
			for (int q = 0; q < 4; q++) {
				if (SAVEZ) { if (live) zrow[((size_t)l*H + c + q)*n] = z[q]; } // a warp writes 32 consecutive samples of one neuron
				float a = sinReduced(w0*z[q]);
				if (!live) a = 0.0f;
				if (lastHidden) {
					y0 += sWL[c + q]*a;
					if (outDim > 1) y1 += sWL[H + c + q]*a;
					if (outDim > 2) y2 += sWL[2*H + c + q]*a;
				}
				splitTf32(a, hi[q], lo[q]);
			}
			if (!lastHidden) {
				const int off = coreOffsetBytes<H>(row, c);
				*reinterpret_cast<float4*>(Ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
				*reinterpret_cast<float4*>(Alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
			}
		};
		{ // first layer on the FMA pipe, written straight into the A operand
			float x0 = 0.0f, x1 = 0.0f, x2 = 0.0f;
			if (live) { x0 = x[s*inDim]; if (inDim > 1) x1 = x[s*inDim + 1]; if (inDim > 2) x2 = x[s*inDim + 2]; }
#pragma unroll 4
			for (int c4 = 0; c4 < HC; c4 += 4) {
				float z[4];
#pragma unroll
				for (int q = 0; q < 4; q++) {
					const float4 w = sW0[cBeg + c4 + q];
					z[q] = fmaf(w.z, x2, fmaf(w.y, x1, fmaf(w.x, x0, w.w))); // absent inputs carry zero weights
				}
				emit4(0, cBeg + c4, z, nHidden == 0);
			}
		}
		for (int l = 1; l <= nHidden; l++) {
			for (int nc = 0; nc < H/kNChunk; nc++) {
				// stage one 64-row chunk of W_l (rows = output neurons, K-major) as hi/lo TF32 operands
				TRACE(2);
				storeW();
				TRACE(3);
				// generic-proxy writes (A from the previous epilogue, B from above) -> visible to the tensor core
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncthreads();
				TRACE(4);
				{ // fetch the chunk that follows (this layer, the next layer, or the first one of the next tile) during the MMAs;
				  // requested before the MMA issue so that thread 0's share is not late
					int ln = l, ncn = nc + 1;
					if (ncn == H/kNChunk) { ncn = 0; ln = l + 1; }
					if (ln <= nHidden) loadW(ln, ncn);
					else if ((tile + gridDim.x)*kTile < n) loadW(1, 0);
				}
				if (tid == 0) {
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t d = tmemBase + (uint32_t)(nc*kNChunk);
					const uint32_t aH = smemAddr(Ahi), aL = smemAddr(Alo), bH = smemAddr(Bhi), bL = smemAddr(Blo);
					const uint32_t sbo = H*32;
					const uint64_t dAh = smemDesc(aH, 128, sbo), dAl = smemDesc(aL, 128, sbo);
					const uint64_t dBh = smemDesc(bH, 128, sbo), dBl = smemDesc(bL, 128, sbo);
#pragma unroll
					for (int ks = 0; ks < H/8; ks++) { // UMMA K = 8 tf32 = two 16-byte core-matrix columns = 256 bytes (16 in descriptor units)
						mmaTf32(d, dAh + 16*ks, dBh + 16*ks, idesc, ks > 0 ? 1u : 0u);
						mmaTf32(d, dAh + 16*ks, dBl + 16*ks, idesc, 1u);
						mmaTf32(d, dAl + 16*ks, dBh + 16*ks, idesc, 1u);
					}
					// arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(barAddr) : "memory");
				}
				TRACE(5);
				TRACE(6);
				mbarWait(barAddr, phase);
				phase ^= 1u;
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				TRACE(7);
			}
			// epilogue: this thread's row of the accumulator -> bias, sin, next layer's operand
			const float* bl = sBias + l*H;
			const bool lastHidden = l == nHidden;
			uint32_t v[HC]; // the whole half row in flight, one wait (two round trips per 32 columns before: the epilogue is latency-bound at 8 warps per SM)
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 16)
				nmc_siren_tc::tmemLoad16Async(tmemBase + ((uint32_t)((warp & 3)*32) << 16) + (uint32_t)(cBeg + c0), &v[c0]);
			nmc_siren_tc::tmemLoadWait();
#pragma unroll
			for (int c0 = 0; c0 < HC; c0 += 4) {
				const int c = cBeg + c0;
				const float4 bv = *reinterpret_cast<const float4*>(bl + c);
				const float z[4] = {__uint_as_float(v[c0]) + bv.x, __uint_as_float(v[c0 + 1]) + bv.y, __uint_as_float(v[c0 + 2]) + bv.z, __uint_as_float(v[c0 + 3]) + bv.w};
				if (lastHidden) emit4(l, c, z, true);
				else emit4(l, c, z, false);
			}
		}
		TRACE(8);
		if (half == 1) { ypart[0][row] = y0; ypart[1][row] = y1; ypart[2][row] = y2; } // the other half of this row's last-layer dot product
		__syncthreads();
		if (live && half == 0) {
			y0 += ypart[0][row]; y1 += ypart[1][row]; y2 += ypart[2][row];
			float yo[3] = {y0 + sbL[0], outDim > 1 ? y1 + sbL[1] : 0.0f, outDim > 2 ? y2 + sbL[2] : 0.0f};
			if (env.active) {
				const float xs[3] = {x[s*inDim], inDim > 1 ? x[s*inDim + 1] : 0.0f, inDim > 2 ? x[s*inDim + 2] : 0.0f};
				nmc_siren_detail::envForward(env, inDim, outDim, xs, yo);
			}
			y[s*outDim] = yo[0];
			if (outDim > 1) y[s*outDim + 1] = yo[1];
			if (outDim > 2) y[s*outDim + 2] = yo[2];
		}
		// the next tile's first layer overwrites A: every thread is past its last use (MMAs completed via the mbarrier)
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		__syncthreads();
		TRACE(9);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmemBase), "r"((uint32_t)H) : "memory");
}

} // namespace

extern "C" int nmc_siren_forward_tc(const nmc_siren_shape* sh, const float* const* W, const float* const* b, const float* x,
									int64_t n, float* y, float* z_saved, const nmc_siren_envelope* envp, void* stream);
#ifdef NMC_TC_TRACE
extern "C" int nmc_siren_trace_read(long long* out, int cap) { // (tag, clock) pairs of the last traced launch
	int n = 0;
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(&n, g_traceN, sizeof(int));
	if (n > cap) n = cap;
	cudaMemcpyFromSymbol(out, g_trace, sizeof(long long)*2*n);
	return n;
}
#endif

namespace nmc_siren_detail { void setError(const char* m); }

extern "C" int nmc_siren_forward_tc(const nmc_siren_shape* sh, const float* const* W, const float* const* b, const float* x,
									int64_t n, float* y, float* z_saved, const nmc_siren_envelope* envp, void* stream) {
	if (!sh || !W || !b) { nmc_siren_detail::setError("null argument"); return 1; }
	if ((sh->hidden != 64 && sh->hidden != 128) || sh->n_hidden_layers < 1 || sh->n_hidden_layers + 2 > kMaxLayers ||
		sh->in_dim < 1 || sh->in_dim > 3 || sh->out_dim < 1 || sh->out_dim > 3) {
		nmc_siren_detail::setError("tensor-core forward: unsupported shape (hidden 64|128, >= 1 hidden layer, in/out 1..3)");
		return 1;
	}
	if (n <= 0) return 0;
	if (!x || !y) { nmc_siren_detail::setError("null buffer"); return 1; }
	Params P;
	for (int l = 0; l < sh->n_hidden_layers + 2; l++) { P.W[l] = W[l]; P.b[l] = b[l]; }
	Env env;
	if (const char* bad = nmc_siren_detail::toEnv(envp, env)) { nmc_siren_detail::setError(bad); return 1; }
	const int H = sh->hidden;
	size_t smem = (size_t)(2*kTile*H + 2*kNChunk*H)*4 + (size_t)smallParamFloats(H, sh->n_hidden_layers, sh->in_dim, sh->out_dim)*4;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	long long tiles = (n + kTile - 1)/kTile;
	int perSM = H == 64 ? 2 : 1;
	int grid = (int)(tiles < (long long)perSM*sms ? tiles : (long long)perSM*sms);
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e;
#define NMC_LAUNCH_FWD(HH, SZ) do { \
		e = cudaFuncSetAttribute(sirenForwardTc<HH, SZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
		if (!e) e = nmc_pdl::launch(sirenForwardTc<HH, SZ>, dim3(grid), dim3(kThreads), smem, st, P, env, (int)sh->in_dim, (int)sh->out_dim, (int)sh->n_hidden_layers, (float)sh->w0, x, (long long)n, y, z_saved); \
	} while (0)
	if (H == 64) { if (z_saved) NMC_LAUNCH_FWD(64, true); else NMC_LAUNCH_FWD(64, false); }
	else { if (z_saved) NMC_LAUNCH_FWD(128, true); else NMC_LAUNCH_FWD(128, false); }
#undef NMC_LAUNCH_FWD
	if (!e) e = cudaGetLastError();
	if (e) { nmc_siren_detail::setError(cudaGetErrorString(e)); return 1; }
	return 0;
}
