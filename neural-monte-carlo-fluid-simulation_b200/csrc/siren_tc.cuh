// csrc/siren_tc.cuh -- tcgen05 / TMEM building blocks shared by the tensor-core SIREN kernels (siren_tc.cu forward,
// siren_tc_bwd.cu delta chain and weight gradients): shared-memory operand layout, descriptors, MMA issue, mbarrier wait,
// the 3xTF32 split and the reduced-argument sine / cosine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmc_siren_tc {

// sin/cos of the SIREN pre-activation w0*z.  |w0 z| stays below a few hundred, so one explicit reduction to
// [-pi, pi] (t = x/2pi - rint(x/2pi), exact subtraction) followed by the SFU sine/cosine is accurate to
// ~5e-7 absolute -- the same size as the fp32 rounding of the argument itself (ulp(100) = 7.6e-6) -- and costs
// 4 instructions instead of the ~40 of sinf's generic range reduction.  -DNMC_SIREN_LIBM_SIN restores sinf/cosf.
// The rounding t - rint(t) is done with the 1.5 * 2^23 magic constant (two FADDs, exact for |t| < 2^22, the same
// round-to-nearest-even result as rintf) instead of FRND, which shares the 16-lane XU pipe with the sine itself.
__device__ __forceinline__ float turnsReduced(float x) {
	const float t = x*0.15915494309189535f;
	const float r = __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f);
	return t - r;
}
__device__ __forceinline__ float sinReduced(float x) {
#ifdef NMC_SIREN_LIBM_SIN
	return sinf(x);
#else
	return __sinf(6.283185307179586f*turnsReduced(x));
#endif
}
__device__ __forceinline__ float cosReduced(float x) {
#ifdef NMC_SIREN_LIBM_SIN
	return cosf(x);
#else
	return __cosf(6.283185307179586f*turnsReduced(x));
#endif
}

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (r, k) of an operand with K columns (fp32/tf32, 4 per 16 bytes); 8 x 16-byte core
// matrices, the two core-matrix columns of one UMMA K step (8 tf32) 128 bytes apart, 8-row groups K*32 bytes apart
template <int K>
__device__ __forceinline__ int coreOffsetBytes(int r, int k) {
	return (r >> 3)*(K*32) + (k >> 2)*128 + (r & 7)*16 + (k & 3)*4;
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, mma_sm100_desc.hpp:98-123)
__device__ __forceinline__ uint64_t smemDesc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
	return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lboBytes >> 4) << 16) | ((uint64_t)(sboBytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor, mma_sm100_desc.hpp:412-439): D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t instrDescTf32(int M, int N) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mmaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
				 :: "r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mmaCommit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
					 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}
__device__ __forceinline__ void fenceBeforeSync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fenceAfterSync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core (async proxy)
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmemAlloc(uint32_t* slot, uint32_t cols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(slot)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmemFree(uint32_t base, uint32_t cols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32 (w & 3) .. + 31)
__device__ __forceinline__ void tmemLoad16(uint32_t taddr, uint32_t (&v)[16]) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
				 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
				   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
				 : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same without the wait: several loads can be in flight before one tmemLoadWait()
__device__ __forceinline__ void tmemLoad16Async(uint32_t taddr, uint32_t* v) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
				 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
				   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
				 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmemLoadWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void splitTf32(float v, float& hi, float& lo) {
	hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); // the 10 mantissa bits the tensor core keeps
	lo = v - hi;
}
__device__ __forceinline__ void splitTf32(const float4& v, float4& hi, float4& lo) {
	splitTf32(v.x, hi.x, lo.x); splitTf32(v.y, hi.y, lo.y); splitTf32(v.z, hi.z, lo.z); splitTf32(v.w, hi.w, lo.w);
}
// 16-byte vector reduction into global memory (sm_90+): four fp32 atomic adds in one L2 transaction
__device__ __forceinline__ void redAdd4(float* addr, float a, float b, float c, float d) {
	asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

} // namespace nmc_siren_tc
