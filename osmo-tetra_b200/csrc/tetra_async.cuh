/*
 * tetra_async.cuh - the pieces the staged kernels share: cp.async.bulk (TMA bulk copy) + mbarrier plumbing in PTX
 * (plain copies under the CPU emulator), and the IDP.4A byte -> bit packing.
 */
#pragma once
#include "tetra_kernels.cuh"

namespace tb {

/* ---- async-copy plumbing (PTX on the GPU, plain copies under the CPU emulator) ---- */
#ifdef TB_SIMT_EMULATION
__device__ __forceinline__ void mbar_init(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_arrive(uint64_t *) {}
__device__ __forceinline__ void mbar_wait(uint64_t *, unsigned) { __syncwarp(); }
__device__ __forceinline__ void fence_async_shared() {}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, unsigned bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
	for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
	return c;
}
#else
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"W_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra W_%=;\n\t}"
		:: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
/* generic-proxy accesses to shared memory before this point are ordered before later async-proxy (bulk copy) accesses */
__device__ __forceinline__ void fence_async_shared()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
/* the same with an L2 evict-first hint: stream data that is read exactly once should not push the decode pass's
 * survivor histories (written once, read once ~100 us later) out of L2 */
__device__ __forceinline__ void bulk_g2s_stream(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
	             :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
#endif

/* 16 bytes holding 0/1 -> 16 bits (first byte -> bit 0) with four integer dot products */
__device__ __forceinline__ uint32_t pack16_dp4a(uint4 v)
{
	const uint32_t lo = dp4a_u(v.y, 0x80402010u, dp4a_u(v.x, 0x08040201u, 0));
	const uint32_t hi = dp4a_u(v.w, 0x80402010u, dp4a_u(v.z, 0x08040201u, 0));
	return lo | (hi << 8);
}

}  // namespace tb
