/* TEST INFRASTRUCTURE ONLY.
 *
 * A tiny SIMT emulator that lets the CUDA sources under osmo-tetra_b200/csrc be
 * compiled with g++ and executed on the CPU, so that the `-m "not gpu"` tests can
 * check kernel logic and the host-side lock state machine end to end in a
 * container without a GPU.  It is never part of the product: the product library
 * is built by nvcc for sm_100a only and fails loudly without a CUDA device.
 *
 * Model: one fiber per CUDA thread, blocks executed one after the other, lanes of
 * a warp meet at warp collectives (__shfl*_sync, __ballot_sync, __syncwarp) and
 * all threads of a block meet at __syncthreads().  Only the subset of CUDA used
 * by the kernels is provided.
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
/* standard headers the host code uses: before the CUDA keywords below become macros (libstdc++ spells attributes like
 * __noinline__ itself) */
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstddef>
#include <mutex>
#include <thread>

#define TB_SIMT_EMULATION 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }

namespace simt {

struct Fiber {
	void *sp = nullptr;
	char *stack = nullptr;
	bool done = false;
};

struct Block {
	std::vector<Fiber> fibers;
	void *main_sp = nullptr;
	int cur = -1;
	unsigned nthreads = 0;
	/* warp barriers */
	unsigned warp_arrived[64];
	unsigned warp_gen[64];
	unsigned warp_live[64];
	uint64_t xchg[64][32];
	/* block barrier */
	unsigned blk_arrived = 0, blk_gen = 0, blk_live = 0;
	std::function<void()> body;
};

extern thread_local Block *g_blk;
extern thread_local dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

extern "C" void simt_ctx_switch(void **save_sp, void *load_sp);
void yield_to_main();
void run_block(Block &b);
void launch(dim3 grid, dim3 block, const std::function<void()> &body);

static inline unsigned lane_id() { return g_threadIdx.x & 31; }
static inline unsigned warp_id() { return g_threadIdx.x >> 5; }

void warp_barrier();
void block_barrier();
void set_dyn_smem(size_t bytes);
uint8_t *dyn_smem();

template <typename T>
static inline T warp_exchange(T v, unsigned src)
{
	static_assert(sizeof(T) <= 8, "exchange type too large");
	Block *b = g_blk;
	unsigned w = warp_id();
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	b->xchg[w][lane_id()] = raw;
	warp_barrier();
	uint64_t got = b->xchg[w][src & 31];
	warp_barrier();
	T out;
	memcpy(&out, &got, sizeof(T));
	return out;
}

}  // namespace simt

#define threadIdx (simt::g_threadIdx)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)

static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_barrier(); }
static inline void __syncthreads() { simt::block_barrier(); }

template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return simt::warp_exchange(v, (unsigned)src); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
	unsigned l = simt::lane_id();
	return simt::warp_exchange(v, l + d < 32 ? l + d : l);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
	unsigned l = simt::lane_id();
	return simt::warp_exchange(v, l >= d ? l - d : l);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m)
{
	return simt::warp_exchange(v, simt::lane_id() ^ (unsigned)m);
}
static inline unsigned __ballot_sync(unsigned, int pred)
{
	simt::Block *b = simt::g_blk;
	unsigned w = simt::warp_id();
	b->xchg[w][simt::lane_id()] = pred ? 1 : 0;
	simt::warp_barrier();
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if (b->xchg[w][i]) r |= 1u << i;
	simt::warp_barrier();
	return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s)
{
	s &= 31;
	return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s)
{
	s &= 31;
	return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline uint32_t __brev(uint32_t v)
{
	uint32_t r = 0;
	for (int i = 0; i < 32; i++) if (v & (1u << i)) r |= 1u << (31 - i);
	return r;
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
	uint64_t t = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		unsigned sel = (s >> (4 * i)) & 7;
		r |= (uint32_t)((t >> (8 * sel)) & 0xff) << (8 * i);
	}
	return r;
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }

/* fibers of a block run one at a time, so plain read-modify-write is atomic */
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }

/* ---- the slice of the CUDA runtime API the host code uses ---- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
static inline const char *cudaGetErrorString(cudaError_t) { return "simt-emulation"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
struct cudaDeviceProp { int multiProcessorCount; int major, minor; char name[256]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); p->multiProcessorCount = 2; p->major = 10; strcpy(p->name, "simt-emulation"); return 0; }

#define TB_LAUNCH(kernel, grid, block, stream, ...) \
	simt::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })
#define TB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
	(simt::set_dyn_smem(smem), simt::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); }))
#define TB_DYN_SMEM() (simt::dyn_smem())
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

/* packed 16x2 SIMD intrinsics used by the lane Viterbi */
static inline uint32_t __vadd2(uint32_t a, uint32_t b)
{
	return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
}
static inline uint32_t __vminu2(uint32_t a, uint32_t b)
{
	uint32_t lo = (a & 0xffff) < (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
	uint32_t hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
	return lo | (hi << 16);
}
static inline uint32_t __viaddmin_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vminu2(__vadd2(a, b), c); }
