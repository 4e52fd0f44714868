// profiles/tools/umma_mn_probe.cu -- which shared-memory words does tcgen05.mma.kind::tf32 read for an MN-major operand
// (no swizzle)?  One operand is a K-major one-hot selector (layout known from the forward kernels), the other is a region of
// shared memory whose every 4-byte word holds its own index; D[m][n] then IS the word index the tensor core read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../neural-monte-carlo-fluid-simulation_b200/csrc -o umma_mn_probe umma_mn_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "siren_tc.cuh"
using namespace nmc_siren_tc;

struct Cfg { int which; int M, N; uint32_t lbo, sbo; uint32_t extraBits; uint32_t layoutType; uint32_t startOff; };

__global__ void probe(Cfg c, float* out) {
	extern __shared__ __align__(1024) unsigned char smem[];
	float* region = reinterpret_cast<float*>(smem);              // 16 KB: word i holds (i % 2048)
	unsigned char* sel = smem + 16384;                           // K-major one-hot selector [128 x 8]
	__shared__ __align__(8) unsigned long long mbar;
	__shared__ uint32_t tmemBaseSh;
	const int tid = threadIdx.x, warp = tid >> 5;
	if (warp == 0) tmemAlloc(&tmemBaseSh, 128u);
	if (tid == 0) mbarInit(smemAddr(&mbar), 1);
	for (int i = tid; i < 4096; i += blockDim.x) region[i] = (float)(i % 2048);
	if (tid == 0 && (smemAddr(region) & 1023)) printf("region not 1024-aligned: %x\n", smemAddr(region));
	for (int r = tid; r < 128; r += blockDim.x)
		for (int k = 0; k < 8; k++) *reinterpret_cast<float*>(sel + coreOffsetBytes<8>(r, k)) = (k == (r & 7)) ? 1.0f : 0.0f;
	fenceProxyAsync();
	fenceBeforeSync();
	__syncthreads();
	fenceAfterSync();
	const uint32_t tmemBase = tmemBaseSh;
	if (tid == 0) {
		const uint64_t dSel = smemDesc(smemAddr(sel), 128, 8*32);
		const uint64_t dReg = smemDesc(smemAddr(region) + c.startOff, c.lbo, c.sbo) | ((uint64_t)c.layoutType << 61);
		// sentinel: D = sel * sel^T (K-major both) = identity-ish pattern, to see whether the probed MMA overwrites it
		mmaTf32(tmemBase, dSel, dSel, instrDescTf32(c.M, c.N), 0u);
		uint32_t id = instrDescTf32(c.M, c.N) | c.extraBits;
		if (c.which == 0) mmaTf32(tmemBase, dReg, dSel, id, 0u);   // A probed: D[m][n] = A[m][n % 8]
		else mmaTf32(tmemBase, dSel, dReg, id, 0u);                // B probed: D[m][n] = B[n][m % 8]
		mmaCommit(smemAddr(&mbar));
	}
	mbarWait(smemAddr(&mbar), 0);
	fenceAfterSync();
	uint32_t v[16];
	for (int c0 = 0; c0 < c.N; c0 += 16) {
		tmemLoad16(tmemBase + ((uint32_t)(warp*32) << 16) + (uint32_t)c0, v);
		for (int q = 0; q < 16; q++) out[(size_t)tid*256 + c0 + q] = __uint_as_float(v[q]);
	}
	fenceBeforeSync();
	__syncthreads();
	if (warp == 0) tmemFree(tmemBase, 128u);
}

int main(int argc, char** argv) {
	const int only = argc > 1 ? atoi(argv[1]) : -1;
	float* d; cudaMalloc(&d, 128*256*4);
	std::vector<float> h(128*256);
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 4096);
	const uint32_t TA = 1u << 15, TB = 1u << 16;
	Cfg cfgs[] = {
		{0, 128, 16, 16, 1024, 0, 2, 0},        // 0: A K-major, standard 128B swizzle, for reference
		{0, 128, 16, 16, 1024, 0, 2, 32},       // 1
		{0, 64, 16, 16, 1024, 0, 1, 0},         // 2: A K-major read with layout type 1?
		{0, 64, 16, 16, 1024, 0, 1, 32},        // 3
		{0, 128, 16, 16, 1024, 0, 1, 0},        // 4
		{1, 64, 32, 16, 1024, 0, 1, 0},         // 5: B K-major, layout type 1
		{1, 64, 32, 16, 1024, 0, 1, 64},        // 6
		{1, 64, 64, 4096, 512, TB, 1, 0},       // 7: B MN-major type 1 (known good)
	};



	int idx = -1;
	for (const Cfg& c : cfgs) {
		if (++idx != only && only >= 0) continue;
		cudaMemset(d, 0xFF, 128*256*4);
		probe<<<1, 128, 16384 + 4096>>>(c, d);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h.data(), d, h.size()*4, cudaMemcpyDeviceToHost);
		printf("== %s probed, M %d N %d LBO %u SBO %u transpose-bits %x layout %u (%s)\n", c.which ? "B" : "A", c.M, c.N, c.lbo, c.sbo, c.extraBits >> 15, c.layoutType*1000 + c.startOff, cudaGetErrorString(e));
		if (c.which == 1) { // rows m = 0..7 (k), columns n
			for (int m = 0; m < 8; m++) { printf("  k=%d:", m); for (int n = 0; n < c.N; n += (c.N > 32 && n >= 12 ? 4 : 1)) printf(" %5.0f", h[(size_t)m*256 + n]); printf("\n"); }
		} else {            // columns n = 0..7 (k), rows m: TMEM lane of row m = m (M = 128) or (m % 16) + 32 (m / 16) (M = 64)
			for (int k = 0; k < 8; k++) {
				printf("  k=%d:", k);
				for (int mi = 0; mi < 24; mi++) { int m = mi < 12 ? mi : (mi - 12)*(c.M/12) + 12; if (m >= c.M) break; int lane = c.M == 128 ? m : (m % 16) + 32*(m/16); printf(" %5.0f", h[(size_t)lane*256 + k]); }
				printf("\n");
			}
		}
	}
	return 0;
}
