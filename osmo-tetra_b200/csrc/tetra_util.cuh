/*
 * tetra_util.cuh - two small HBM-bound helpers around the receive chain:
 *
 *   k_slots_digest   order-independent 64-bit digest of slot records + packed type-1 bits (the device twin of the
 *                    "running digest of the parity records" BASELINE.md asks for): the digests of the shards of a
 *                    sharded run add up, mod 2^64, to the digest of the single-GPU run of the same stream.
 *   k_pack_bits      the reference's one-bit-per-byte stream (tetra-rx.c:82-95) -> TB200_IN_PACKED, eight bits per
 *                    byte: what travels over NVLink when one stream is decoded by several GPUs (8x fewer bytes).
 */
#pragma once
#include "tetra_kernels.cuh"
#include "tetra_async.cuh"

namespace tb {

/* one slot's contribution: a multiply-xorshift chain over (global slot index, 4 record words, 9 type-1 words) */
TB_HD inline uint64_t digest_mix(uint64_t h, uint32_t w)
{
	h = (h ^ w) * 0x9E3779B97F4A7C15ull;
	return h ^ (h >> 32);
}

TB_HD inline uint64_t slot_digest(uint64_t k_global, const SlotOut &s, const uint32_t *p9)
{
	uint64_t h = (k_global + 1) * 0xD6E8FEB86659FD93ull;
	h ^= h >> 29;
	h = digest_mix(h, s.slot_bit);
	h = digest_mix(h, s.scrambling_code);
	h = digest_mix(h, (uint32_t)s.find_off | ((uint32_t)s.window << 16));
	h = digest_mix(h, (uint32_t)s.time | ((uint32_t)(uint8_t)s.find_rc << 16) | ((uint32_t)s.flags << 24));
	for (int i = 0; i < TYPE1_WORDS; ++i) h = digest_mix(h, p9 ? p9[i] : 0u);
	return h;
}

__global__ void __launch_bounds__(256)
k_slots_digest(const SlotOut *__restrict__ slots, const uint32_t *__restrict__ packed, uint64_t n, uint64_t k_base,
               unsigned long long *out)
{
	uint64_t acc = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint32_t p[TYPE1_WORDS];
#pragma unroll
		for (int j = 0; j < TYPE1_WORDS; ++j) p[j] = packed ? packed[i * TYPE1_WORDS + j] : 0u;
		acc += slot_digest(k_base + i, slots[i], p);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, (unsigned long long)acc);
}

/* n_bits bytes holding 0/1 -> bits, stream bit i = byte i>>3 bit i&7; a thread turns 32 bytes into one word
 * (two 16-byte loads, IDP.4A packing, coalesced 4-byte stores).  `in` may have any alignment. */
__global__ void __launch_bounds__(256)
k_pack_bits(const uint8_t *__restrict__ in, uint64_t n_bits, uint32_t *__restrict__ out)
{
	const uint64_t n_words = (n_bits + 31) >> 5;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
	for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
		const uint64_t b0 = w << 5;
		uint32_t v;
		if (aligned && b0 + 32 <= n_bits) {
			const uint4 *q = reinterpret_cast<const uint4 *>(in + b0);
			v = pack16_dp4a(__ldcs(q)) | (pack16_dp4a(__ldcs(q + 1)) << 16);
		} else {
			v = 0;
			for (int i = 0; i < 32; ++i)
				if (b0 + i < n_bits) v |= (uint32_t)(in[b0 + i] & 1u) << i;
		}
		out[w] = v;
	}
}

}  // namespace tb
