/*
 * tetra_stage_tma.cuh - the fused descramble + de-interleave stage on its own, SCH/F geometry
 * (K = 432, a = 103): type-5 bytes in, type-3 bytes out (tetra_scramb.c:77-85 then
 * tetra_interleave.c:51-60), one THREAD per block.
 *
 * Memory-bound by construction (432 B read + 432 B written per block), so everything is arranged
 * around the copy engines: each lane bulk-copies (cp.async.bulk) its 432-byte block into its own
 * shared-memory row (27 x 16 B: an odd number of 16-byte units, so 128-bit row accesses of the 32
 * lanes are conflict free), packs it to bits with IDP.4A, XORs the scrambling sequence built from
 * nibble tables, and writes the permuted bits back as bytes into a dense output tile that leaves
 * the SM as ONE 13 824-byte bulk store per warp.  Two row sets per warp are in flight.
 */
#pragma once
#include "tetra_async.cuh"

namespace tb {

constexpr int ST_K = 432, ST_A = 103;
constexpr int ST_WARPS = 8;
constexpr int ST_STAGES = 2;      /* tiles per warp: the one being transformed and the one loading */
constexpr int ST_LF_WORDS = 14;
/* per warp: ST_STAGES tiles of 32 x 432 B, transformed IN PLACE (each thread owns its row);
 * per CTA: nibble tables 8 x 14 x 16 words */
constexpr size_t ST_TILE = 32 * ST_K;
constexpr size_t ST_SMEM = (size_t)ST_WARPS * ST_STAGES * ST_TILE + 8 * ST_LF_WORDS * 16 * 4 + ST_WARPS * ST_STAGES * 8 + 16;

#ifdef TB_SIMT_EMULATION
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) { memcpy(dst, src, bytes); }
template <int N> __device__ __forceinline__ void bulk_store_wait_read() {}
__device__ __forceinline__ void fence_async_smem() {}
#else
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
	             :: "l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_store_wait_read()      /* at most N stores still reading smem */
{
	asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

__global__ void __launch_bounds__(ST_WARPS * 32)
k_stage_tma(const uint8_t *__restrict__ type5, uint8_t *__restrict__ type3, const uint32_t *__restrict__ codes,
            uint64_t n, const Tables *__restrict__ tab)
{
	uint8_t *smem = TB_DYN_SMEM();
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint8_t *tiles = smem + (size_t)wib * ST_STAGES * ST_TILE;
	uint32_t *nib = reinterpret_cast<uint32_t *>(smem + (size_t)ST_WARPS * ST_STAGES * ST_TILE);   /* [8][14][16] */
	uint64_t *bars = reinterpret_cast<uint64_t *>(nib + 8 * ST_LF_WORDS * 16) + wib * ST_STAGES;

	/* nibble tables of the (linear) scrambler: word w of the sequence = XOR_n nib[n][w][(code >> 4n) & 15] */
	for (int i = threadIdx.x; i < 8 * ST_LF_WORDS * 16; i += blockDim.x) {
		const int v = i & 15, w = (i >> 4) % ST_LF_WORDS, nn = i / (16 * ST_LF_WORDS);
		uint32_t x = 0;
#pragma unroll
		for (int b = 0; b < 4; ++b)
			if ((v >> b) & 1) x ^= tab->lfsr_col[4 * nn + b][w];
		nib[i] = x;
	}
	if (lane == 0)
		for (int s = 0; s < ST_STAGES; ++s) mbar_init(&bars[s], 1);
	__syncthreads();

	const uint64_t ngroups = (n + 31) / 32;
	const uint64_t nwarps = (uint64_t)gridDim.x * ST_WARPS;
	const uint64_t w0 = (uint64_t)blockIdx.x * ST_WARPS + wib;

	/* rows are dense (stride K), so a warp's 32 blocks are one contiguous 13 824-byte copy */
	auto issue = [&](uint64_t grp, int st) {
		if (lane == 0) {
			if (grp < ngroups) {
				const uint64_t first = grp * 32;
				const unsigned bytes = (unsigned)((n - first < 32 ? n - first : 32) * ST_K);
				mbar_expect_tx(&bars[st], bytes);
				bulk_g2s(tiles + (size_t)st * ST_TILE, type5 + first * ST_K, bytes, &bars[st]);
			} else {
				mbar_arrive(&bars[st]);
			}
		}
	};

	unsigned phase_bits = 0;
	int st = 0;
	issue(w0, 0);
	/* per-block scrambling codes are fetched one round ahead so their latency is off the critical path */
	uint32_t code_next = (w0 < ngroups && w0 * 32 + lane < n) ? codes[w0 * 32 + lane] : 0u;
	for (uint64_t grp = w0; grp < ngroups; grp += nwarps) {
		const uint32_t code = code_next;
		{
			const uint64_t inext = (grp + nwarps) * 32 + lane;
			code_next = (grp + nwarps < ngroups && inext < n) ? codes[inext] : 0u;
		}
		/* the next tile goes where the previous tile was stored from: wait until that store has
		 * finished reading shared memory, then start the copy so it overlaps this tile's work */
		if (lane == 0) bulk_store_wait_read<0>();
		__syncwarp();
		issue(grp + nwarps, st ^ 1);
		mbar_wait(&bars[st], (phase_bits >> st) & 1u);
		phase_bits ^= 1u << st;
		const uint64_t i = grp * 32 + lane;
		const bool have = i < n;
		uint8_t *tile = tiles + (size_t)st * ST_TILE;
		if (have) {
			uint4 *row = reinterpret_cast<uint4 *>(tile + (size_t)lane * ST_K);
			uint32_t t4[ST_LF_WORDS];
#pragma unroll
			for (int j = 0; j < 13; ++j)
				t4[j] = pack16_dp4a(row[2 * j]) | (pack16_dp4a(row[2 * j + 1]) << 16);
			t4[13] = pack16_dp4a(row[26]);
#pragma unroll
			for (int w = 0; w < ST_LF_WORDS; ++w) {
				uint32_t x = 0;
#pragma unroll
				for (int nn = 0; nn < 8; ++nn) x ^= nib[(nn * ST_LF_WORDS + w) * 16 + ((code >> (4 * nn)) & 15)];
				t4[w] ^= x;
			}
			/* permuted bits back to bytes, in place: this thread is the only user of its row */
#pragma unroll
			for (int u = 0; u < 27; ++u) {
				uint32_t wd[4];
#pragma unroll
				for (int q = 0; q < 4; ++q) {
					uint32_t v = 0;
#pragma unroll
					for (int b = 0; b < 4; ++b) {
						const int j = 16 * u + 4 * q + b;
						const unsigned m = (unsigned)(ST_A * (j + 1)) % ST_K;
						const int sh = (int)(m & 31) - 8 * b;
						const uint32_t src = t4[m >> 5];
						const uint32_t x = sh >= 0 ? (src >> sh) : (src << (-sh));
						v |= x & (1u << (8 * b));
					}
					wd[q] = v;
				}
				row[u] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
			}
		}
		fence_async_smem();               /* generic-proxy writes above -> visible to the bulk store */
		__syncwarp();
		const uint64_t first = grp * 32;
		const uint64_t cnt = n - first < 32 ? n - first : 32;
		if (lane == 0) bulk_s2g(type3 + first * ST_K, tile, (unsigned)(cnt * ST_K));
		st ^= 1;
	}
	if (lane == 0) bulk_store_wait_read<0>();
}

}  // namespace tb
