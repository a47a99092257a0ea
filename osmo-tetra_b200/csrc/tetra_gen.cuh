/*
 * tetra_gen.cuh - synthetic TETRA downlink stream generator (TX side), one thread per burst.
 *
 * Bench / test input only: builds SYNC and normal continuous downlink bursts the way the
 * reference's own test generator does (conv_enc_test.c:88-156,198-305: type-1 -> CRC16 ->
 * 4 tail bits -> rate-1/4 mother code -> 2/3 puncturing -> block interleaving ->
 * scrambling -> burst layout of phy/tetra_burst.c:169-260), from a counter-based RNG keyed
 * by (seed, burst index) so any shard of a large stream can be regenerated anywhere.
 * The CPU twin in oracle/tetra_oracle.c produces the same bits; tests compare the two.
 */
#pragma once
#include <stdint.h>

namespace tb {

struct GenCfg {
	uint64_t seed;
	uint32_t sb_period;
	uint32_t lead_sb;
	uint32_t ndb2_per_256;
	uint32_t ber_per_65536;
	uint32_t random_cell;
	uint32_t lead_in_bits;
};

enum { G_KIND = 1, G_CELL = 2, G_BLKB = 3, G_BLKA = 4, G_BBK = 5, G_NOISE = 6, G_LEADIN = 7 };

TB_HD inline uint64_t gen_mix(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

TB_HD inline uint64_t gen_rng(uint64_t seed, uint64_t k, uint32_t lane, uint32_t w)
{
	return gen_mix(gen_mix(seed + 0x9e3779b97f4a7c15ull * (k + 1)) + ((uint64_t)lane << 32) + w);
}

TB_HD inline int gen_kind(const GenCfg &c, uint64_t k)
{
	if (k < c.lead_sb || (c.sb_period && k % c.sb_period == 0)) return 3;       /* SYNC */
	if ((gen_rng(c.seed, k, G_KIND, 0) & 255) < c.ndb2_per_256) return 1;         /* NORM_2 */
	return 0;                                                                     /* NORM_1 */
}

TB_HD inline uint64_t gen_last_sb(const GenCfg &c, uint64_t k)
{
	uint64_t best = 0;
	if (c.lead_sb) best = (k < c.lead_sb) ? k : c.lead_sb - 1;
	if (c.sb_period) {
		uint64_t p = k - k % c.sb_period;
		if (p > best || !c.lead_sb) best = p;
	}
	return best;
}

/* little helpers over one-bit-per-byte arrays in thread-local memory */
struct GenBlock {
	uint8_t t2[288];
	uint8_t t3[432];
	uint8_t t5[432];
};

TB_HD inline uint16_t gen_crc16(const uint8_t *bits, int len)
{
	uint32_t crc = 0xffff;
	for (int i = 0; i < len; i++) {
		uint32_t top = ((crc >> 15) ^ bits[i]) & 1;
		crc = (crc << 1) & 0xffff;
		if (top) crc ^= 0x1021;
	}
	return (uint16_t)crc;
}

/* type-1 bits already in g.t2[0..T1); produces g.t5[0..K) */
TB_HD inline void gen_encode(GenBlock &g, int K, int N, int T1, int A, uint32_t code)
{
	uint16_t crc = (uint16_t)~gen_crc16(g.t2, T1);
	for (int i = 0; i < 16; i++) g.t2[T1 + i] = (crc >> (15 - i)) & 1;
	for (int i = T1 + 16; i < N; i++) g.t2[i] = 0;
	/* mother code + 2/3 puncturing: of each pair of steps keep G1,G2 of the even one, G1 of the odd one */
	unsigned st = 0;
	for (int t = 0; t < N; t++) {
		unsigned b = g.t2[t], o = mother_out(st, b);
		if ((t & 1) == 0) {
			g.t3[3 * (t >> 1) + 0] = (o >> 3) & 1;
			g.t3[3 * (t >> 1) + 1] = (o >> 2) & 1;
		} else {
			g.t3[3 * (t >> 1) + 2] = (o >> 3) & 1;
		}
		st = ((st << 1) | b) & 15;
	}
	/* interleave (tetra_interleave.c:41-49) and scramble (tetra_scramb.c:77-85) */
	for (int j = 0; j < K; j++) g.t5[(A * (j + 1)) % K] = g.t3[j];
	uint32_t l = code;
	for (int i = 0; i < K; i++) {
		uint32_t fb = 0, x = l & 0xDB710641u;
		x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; fb = x & 1;
		l = (l >> 1) | (fb << 31);
		g.t5[i] ^= (uint8_t)fb;
	}
}

TB_HD inline void gen_noise(uint8_t *bits, int n, const GenCfg &c, uint64_t k, uint32_t sub)
{
	if (!c.ber_per_65536) return;
	for (int i = 0; i < n; i += 4) {
		uint64_t r = gen_rng(c.seed, k, G_NOISE, (sub << 8) + (i >> 2));
		for (int q = 0; q < 4 && i + q < n; q++)
			if (((r >> (16 * q)) & 0xffff) < c.ber_per_65536) bits[i + q] ^= 1;
	}
}

TB_HD inline void gen_random_bits(uint8_t *dst, int n, uint64_t seed, uint64_t k, uint32_t lane)
{
	for (int i = 0; i < n; i += 64) {
		uint64_t r = gen_rng(seed, k, lane, i >> 6);
		for (int q = 0; q < 64 && i + q < n; q++) dst[i + q] = (r >> q) & 1;
	}
}

TB_HD inline void gen_put(uint8_t *dst, uint64_t v, int n)
{
	for (int i = 0; i < n; i++) dst[i] = (v >> (n - 1 - i)) & 1;
}

TB_HD inline uint32_t gen_rm3014(uint32_t info)   /* tetra_rm3014.c:28-43,74-86 */
{
	const uint16_t par[14] = { 0x9b60, 0x2de0, 0xfc20, 0xe03c, 0x983a, 0x5436, 0x2c2e,
	                           0xffdf, 0x8339, 0x42b5, 0x21ad, 0x1273, 0x096b, 0x04e7 };
	uint32_t v = 0;
	for (int i = 0; i < 14; i++)
		if ((info >> (13 - i)) & 1) v ^= (1u << (29 - i)) | par[i];
	return v;
}

TB_HD inline void gen_burst(const GenCfg &c, uint64_t k, uint8_t *burst)
{
	const int kind = gen_kind(c, k);
	uint32_t mcc = 262, mnc = 42, cc = 1;
	if (c.random_cell) {
		uint64_t r = gen_rng(c.seed, gen_last_sb(c, k), G_CELL, 0);
		mcc = r & 0x3ff; mnc = (r >> 10) & 0x3fff; cc = (r >> 24) & 0x3f;
	}
	const uint32_t code = scramb_init_from(mcc, mnc, cc);
	GenBlock g;

	for (int i = 0; i < 510; i++) burst[i] = 0;
	for (int i = 0; i < 12; i++) burst[i] = (SEQ_Q >> (10 + i)) & 1;        /* q11..q22 */
	for (int i = 0; i < 10; i++) burst[500 + i] = (SEQ_Q >> i) & 1;         /* q1..q10 */

	/* AACH: header bits 00 so the upper MAC never switches the slot to traffic */
	uint8_t bb[30];
	{
		uint64_t rb = gen_rng(c.seed, k, G_BBK, 0);
		gen_put(bb, gen_rm3014((uint32_t)(rb & 0x0fff)), 30);
		uint32_t l = code;
		for (int i = 0; i < 30; i++) {
			uint32_t x = l & 0xDB710641u;
			x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
			uint32_t fb = x & 1;
			l = (l >> 1) | (fb << 31);
			bb[i] ^= (uint8_t)fb;
		}
	}

	if (kind == 3) {
		/* SYNC PDU (Table 21.73 order as in testpdu.c:41-57) */
		uint8_t *p = g.t2;
		gen_put(p, 0, 4); p += 4;
		gen_put(p, cc, 6); p += 6;
		gen_put(p, k % 4, 2); p += 2;
		gen_put(p, (k / 4) % 18 + 1, 5); p += 5;
		gen_put(p, (k / 72) % 60 + 1, 6); p += 6;
		gen_put(p, 0, 8); p += 8;
		gen_put(p, mcc, 10); p += 10;
		gen_put(p, mnc, 14); p += 14;
		gen_put(p, 0, 5);
		gen_encode(g, 120, 80, 60, 11, 3);
		gen_noise(g.t5, 120, c, k, 0);
		for (int i = 0; i < 120; i++) burst[94 + i] = g.t5[i];
		gen_random_bits(g.t2, 124, c.seed, k, G_BLKA);
		gen_encode(g, 216, 144, 124, 101, code);
		gen_noise(g.t5, 216, c, k, 1);
		for (int i = 0; i < 216; i++) burst[282 + i] = g.t5[i];
		for (int i = 0; i < 8; i++) { burst[14 + i] = 1; burst[86 + i] = 1; }   /* f1..f8, f73..f80 */
		for (int i = 0; i < 38; i++) burst[214 + i] = (SEQ_Y >> i) & 1;
		for (int i = 0; i < 30; i++) burst[252 + i] = bb[i];
	} else {
		if (kind == 0) {
			gen_random_bits(g.t2, 268, c.seed, k, G_BLKA);
			gen_encode(g, 432, 288, 268, 103, code);
			gen_noise(g.t5, 432, c, k, 0);
			for (int i = 0; i < 216; i++) { burst[14 + i] = g.t5[i]; burst[282 + i] = g.t5[216 + i]; }
		} else {
			gen_random_bits(g.t2, 124, c.seed, k, G_BLKA);
			gen_encode(g, 216, 144, 124, 101, code);
			gen_noise(g.t5, 216, c, k, 0);
			for (int i = 0; i < 216; i++) burst[14 + i] = g.t5[i];
			gen_random_bits(g.t2, 124, c.seed, k, G_BLKB);
			gen_encode(g, 216, 144, 124, 101, code);
			gen_noise(g.t5, 216, c, k, 1);
			for (int i = 0; i < 216; i++) burst[282 + i] = g.t5[i];
		}
		for (int i = 0; i < 14; i++) burst[230 + i] = bb[i];
		const uint32_t ts = (kind == 1) ? SEQ_P : SEQ_N;
		for (int i = 0; i < 22; i++) burst[244 + i] = (ts >> i) & 1;
		for (int i = 0; i < 16; i++) burst[266 + i] = bb[14 + i];
	}
}

__global__ void __launch_bounds__(64)
k_gen_bursts(GenCfg c, uint64_t k0, uint64_t n, uint8_t *out)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint8_t burst[510];
		gen_burst(c, k0 + i, burst);
		uint8_t *dst = out + 510 * i;
		for (int j = 0; j < 510; j++) dst[j] = burst[j];
	}
}

__global__ void __launch_bounds__(256)
k_gen_lead_in(GenCfg c, uint8_t *out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < c.lead_in_bits) {
		uint64_t r = gen_rng(c.seed, 0, G_LEADIN, i >> 6);
		out[i] = (r >> (i & 63)) & 1;
	}
}

}  // namespace tb
