/*
 * tetra_kernels.cuh - device side of the B200 TETRA lower-MAC receive chain.
 *
 * Written for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a).  The same source
 * also compiles under tests/simt/cpu_simt.h, a test-only SIMT emulator used by the CPU
 * test-suite to check kernel logic without a GPU; that build is never shipped.
 *
 * Conventions
 *   - bit streams arrive as one bit per byte (reference ABI); on chip they are packed
 *     32 bits per word, LSB first: stream bit i <-> word[i >> 5] bit (i & 31).
 *   - one warp owns one 510-bit slot for the memory-shaped stages (load, pack, search,
 *     descramble + de-interleave gather, output); the Viterbi stage exists in two forms
 *     (warp-shuffle butterflies over 16 lanes per trellis, and one lane per coded block).
 *
 * Reference behaviour each routine reproduces is cited as file:line under osmo-tetra/src.
 */
#pragma once
#include <cstddef>
#include <stdint.h>

#ifndef TB_SIMT_EMULATION
#include <cuda_runtime.h>
#define TB_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define TB_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
extern __shared__ __align__(16) uint8_t tb_dyn_smem_[];
#define TB_DYN_SMEM() (tb_dyn_smem_)
#endif

#define TB_HD __host__ __device__

namespace tb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SLOT_BITS = 510;

/* ------------------------------------------------------------------ tables -- */

/* Built once on the host (tetra_b200.cu: build_tables) and kept in device memory. */
struct Tables {
	/* LFSR output word w (bits 32w..32w+31 of the scrambling sequence) is the XOR over the
	 * set bits b of the 32-bit init value of lfsr_col[b][w] (tetra_scramb.c:34-50 is linear) */
	uint32_t lfsr_col[32][16];
	uint32_t lfsr_sb1[16];          /* sequence for SCRAMB_INIT = 3 (tetra_scramb.h:14) */
	/* CRC-16-CCITT: contribution of a set message bit at distance d from the end,
	 * x^(16+d) mod 0x11021 (crc_simple.c:65-82 is linear in the message) */
	uint16_t crc_pow[288];
	uint16_t crc_init[4];           /* 0xffff pushed through L zero bits, L = 76, 140, 284, 108 */
	/* tetra_find_train_seq's pre-filter never sees in[20] (tetra_burst.c:288-294): a true
	 * match at offset k <= 20 is reported only if the distorted filter value still equals
	 * one of the five 22-bit prefixes.  blind_ok[seq][prev] bit k: seq 0=y 1=n 2=p,
	 * prev = stream bit k-1 (ignored for k = 0). */
	uint32_t blind_ok[3][2];
	uint32_t pad[2];
	/* the same CRC with the register bit-reversed: byte table [0,256) and nibble table [256,272),
	 * so that LSB-first packed bytes feed it directly (lane kernels) */
	uint32_t crc_tab_r[272];
	/* RM(30,14), tetra_rm3014.c:28-43: parity (low 16 bits of the code word) of the low / high 7 information bits */
	uint16_t rm_par[2][128];
	/* the scrambler 32 steps at a time: the register after 32 steps of tetra_scramb.c:34-50 IS the word of the 32 bits it
	 * put out, so sequence word i+1 = L(word i), word 0 = L(init), L linear: XOR of one entry per byte of the argument */
	uint32_t lfsr_leap[4][256];
};

__device__ __forceinline__ uint32_t lfsr_leap32(const uint32_t *__restrict__ leap, uint32_t x)
{
	return leap[x & 255u] ^ leap[256 + ((x >> 8) & 255u)] ^ leap[512 + ((x >> 16) & 255u)] ^ leap[768 + (x >> 24)];
}

/* ---- RM(30,14) decoding (the AACH, tetra_lower_mac.c:268-274 has a FIXME where this belongs) ----
 * Words follow tetra_rm3014_compute: 30 bits, bit 29 first on air, information = word >> 16.  The code is linear, so
 * the nearest code word is word ^ leader[syndrome], leader = the lightest (then numerically smallest) error pattern
 * with that syndrome: a 2^16-entry table built on the host.  Result: information bits of the nearest code word |
 * distance << 16 | (syndrome != 0) << 24. */
TB_HD inline uint32_t rm3014_syndrome(const Tables *tab, uint32_t word)
{
	const uint32_t info = (word >> 16) & 0x3fffu;
	return (uint32_t)(tab->rm_par[0][info & 127] ^ tab->rm_par[1][info >> 7] ^ (word & 0xffffu));
}

__device__ __forceinline__ uint32_t rm3014_decode(const Tables *__restrict__ tab, const uint32_t *__restrict__ leader, uint32_t word)
{
	word &= 0x3fffffffu;
	const uint32_t syn = rm3014_syndrome(tab, word);
	const uint32_t e = leader[syn];
	return ((word ^ e) >> 16) | ((uint32_t)__popc(e) << 16) | (syn ? 1u << 24 : 0u);
}

/* the slot's 30 descrambled BBK bits, first bit on air in bit 0 -> the word convention above */
__device__ __forceinline__ uint32_t rm3014_word_from_air(uint32_t bits_lsb_first) { return __brev(bits_lsb_first) >> 2; }

/* training sequences, LSB = first bit on air (values checked against the reference's
 * arrays tetra_burst.c:59-70 by tests/test_oracle.py) */
constexpr uint64_t SEQ_Y = 0x3983973983ull;   /* 38 bits  1100 0001 1001 1100 1110 1001 1100 0001 1001 11 */
constexpr uint32_t SEQ_N = 0x0b970bu;         /* 22 bits  1101 0000 1110 1001 1101 00 */
constexpr uint32_t SEQ_P = 0x1ec25eu;         /* 22 bits  0111 1010 0100 0011 0111 10 */
constexpr uint32_t SEQ_Q = 0x2d60edu;         /* 22 bits  1011 0111 0000 0110 1011 01 */
constexpr uint32_t SEQ_X = 0x30b970b9u;       /* 30 bits  1001 1101 0000 1110 1001 1101 0000 11 */

/* block parameters, tetra_lower_mac.c:55-102 */
template <int BT> struct Blk;
template <> struct Blk<0> { static constexpr int K = 120, N = 80,  T1 = 60,  A = 11,  CRCI = 0; };  /* SB1 */
template <> struct Blk<1> { static constexpr int K = 216, N = 144, T1 = 124, A = 101, CRCI = 1; };  /* SB2 / NDB half */
template <> struct Blk<5> { static constexpr int K = 432, N = 288, T1 = 268, A = 103, CRCI = 2; };  /* SCH/F */
template <> struct Blk<4> { static constexpr int K = 168, N = 112, T1 = 92,  A = 13,  CRCI = 3; };  /* SCH/HU (uplink; leaf operator only) */

/* where type-5 bit m of a block sits inside the 510-bit burst, tetra_burst.c:31-47,347-372 */
enum Placement { PL_RAW = 0, PL_SB1 = 1, PL_BLK2 = 2, PL_BLK1 = 3, PL_SCHF = 4 };
template <int PL> __device__ __forceinline__ unsigned place(unsigned m)
{
	if (PL == PL_SB1)  return 94 + m;
	if (PL == PL_BLK2) return 282 + m;
	if (PL == PL_BLK1) return 14 + m;
	if (PL == PL_SCHF) return m < 216 ? 14 + m : 66 + m;
	return m;
}

/* ------------------------------------------------------------- TDMA time -- */

struct Tm { uint32_t tn, fn, mn; };

/* tetra_tdma_time_add_tn(tm, 1), tetra_tdma.c:27-53,75-79 (including its wrap rules) */
TB_HD inline Tm tm_step(Tm t)
{
	t.tn += 1;
	if (t.tn > 4) { t.fn += t.tn / 4; t.tn %= 4; }
	if (t.fn > 18) { t.mn += t.fn / 18; t.fn %= 18; }
	if (t.mn > 60) t.mn %= 60;
	return t;
}

/* n applications of tm_step in O(1): one explicit step brings every field into its
 * steady range (tn 1..4, fn 0..18, mn 0..60), after which tn cycles 1..4, fn cycles
 * 1..18 and mn cycles 1..60 (0 is left once and never re-entered). */
TB_HD inline Tm tm_advance(Tm t, uint64_t n)
{
	if (n == 0) return t;
	t = tm_step(t);
	if (--n == 0) return t;
	uint64_t a = (uint64_t)(t.tn - 1) + n;
	t.tn = (uint32_t)(a % 4) + 1;
	uint64_t cf = a / 4;
	if (cf) {
		if (t.fn == 0) { t.fn = 1; cf--; }
		uint64_t f = (uint64_t)(t.fn - 1) + cf;
		t.fn = (uint32_t)(f % 18) + 1;
		uint64_t cm = f / 18;
		if (cm) {
			if (t.mn == 0) { t.mn = 1; cm--; }
			uint64_t m = (uint64_t)(t.mn - 1) + cm;
			t.mn = (uint32_t)(m % 60) + 1;
		}
	}
	return t;
}

/* the same in 32-bit arithmetic (n < 2^31): what the kernels use - the distance to the last SYNC burst is small, and
 * 64-bit remainders by 18 and 60 cost a subroutine call each */
TB_HD inline Tm tm_advance_small(Tm t, uint32_t n)
{
	if (n == 0) return t;
	t = tm_step(t);
	if (--n == 0) return t;
	const uint32_t a = (t.tn - 1) + n;
	t.tn = (a & 3u) + 1;
	uint32_t cf = a >> 2;
	if (cf) {
		if (t.fn == 0) { t.fn = 1; cf--; }
		const uint32_t f = (t.fn - 1) + cf;
		t.fn = f % 18u + 1;
		uint32_t cm = f / 18u;
		if (cm) {
			if (t.mn == 0) { t.mn = 1; cm--; }
			const uint32_t m = (t.mn - 1) + cm;
			t.mn = m % 60u + 1;
		}
	}
	return t;
}

TB_HD inline bool tm_is_bnch(Tm t)   /* tetra_lower_mac.c:122-127 */
{
	return t.fn == 18 && t.tn == 4 - ((t.mn + 3) % 4);
}

TB_HD inline uint32_t scramb_init_from(uint32_t mcc, uint32_t mnc, uint32_t cc)  /* tetra_scramb.c:87-99 */
{
	return ((((cc & 0x3f) | ((mnc & 0x3fff) << 6) | ((mcc & 0x3ff) << 20))) << 2) | 3;
}

/* mother code output nibble (MSB = G1) for register state `st` (bit0 = newest) and input b,
 * tetra_conv_enc.c:43-74 == tables viterbi_cch.c:35-40 */
TB_HD inline unsigned mother_out(unsigned st, unsigned b)
{
	unsigned d1 = st & 1, d2 = (st >> 1) & 1, d3 = (st >> 2) & 1, d4 = (st >> 3) & 1;
	unsigned g1 = b ^ d1 ^ d4, g2 = b ^ d2 ^ d3 ^ d4, g3 = b ^ d1 ^ d2 ^ d4, g4 = b ^ d1 ^ d3 ^ d4;
	return (g1 << 3) | (g2 << 2) | (g3 << 1) | g4;
}

/* ------------------------------------------------------ per-slot work area -- */

/* what the classify pass leaves for the decode pass, one per slot */
struct SlotWs {
	uint32_t sb1_t1[2];      /* SB1 type-1 bits (60), only for delivered SYNC bursts */
	uint32_t sb_code;        /* scrambling code announced by a CRC-good SB1 */
	uint16_t find_off;
	uint16_t window;
	int8_t   find_rc;
	uint8_t  good_sb;        /* SB1 CRC good */
	uint8_t  kind;           /* TB200_KIND_* */
	uint8_t  unlock;         /* receiver leaves LOCKED after this slot */
	uint8_t  tn, fn, mn, cc; /* SYNC PDU fields (raw) */
	uint16_t mcc, mnc;
	uint32_t sb1_crc;        /* CRC-16 register after SB1's 76 bits (0x1d0f = good), for the optional CRC output */
};
static_assert(sizeof(SlotWs) == 32, "SlotWs layout");
static_assert(offsetof(SlotWs, find_rc) == 16 && offsetof(SlotWs, good_sb) == 17 && offsetof(SlotWs, kind) == 18 && offsetof(SlotWs, unlock) == 19,
              "k_scan_blocks reads these four bytes as one word");

struct SlotOut {                 /* == struct tb200_slot */
	uint32_t slot_bit;
	uint32_t scrambling_code;
	uint16_t find_off;
	uint16_t window;
	uint16_t time;
	int8_t   find_rc;
	uint8_t  flags;
};
static_assert(sizeof(SlotOut) == 16, "SlotOut layout");

/* receiver state carried on the device between launches (t_phy_state + _tcd) */
struct DevCarry {
	uint32_t scramb_init;
	uint32_t tn, fn, mn;
	uint32_t mcc, mnc, cc;
	uint32_t seen_good;        /* a CRC-good SB1 has been seen since the chain started (sharded decode: the state no longer depends on the carry-in) */
	uint64_t first_good;       /* slot index (relative to the chain start) of that first SB1, ~0 if none yet */
};

constexpr int KIND_NONE = 0, KIND_SB = 1, KIND_NDB_F = 2, KIND_NDB_2 = 3;
constexpr int F_CRC_A = 0x04, F_CRC_B = 0x08, F_BNCH = 0x10, F_UNLOCK = 0x20;
constexpr int TS_NORM_1 = 0, TS_NORM_2 = 1, TS_SYNC = 3;
constexpr int TYPE1_STRIDE = 288, TYPE1_WORDS = 9;

/* per-warp shared memory */
struct WarpSmem {
	uint32_t bw[20];         /* the slot's 510 bits, packed (16 words used) */
	uint32_t lf[16];         /* scrambling sequence words */
	uint32_t t3[2][16];      /* type-3 bits of up to two blocks (+1 word slack for funnel reads) */
	uint32_t t2[2][10];      /* decoded type-2 bits */
	uint32_t outw[12];       /* the slot's type-1 string */
	uint32_t bbk[2];         /* AACH type-1 bits (14) */
	uint32_t sb1[4];         /* SB1 type-1 bits (60) handed over by the classify pass */
	uint32_t dec[296];       /* survivor decisions, one word per trellis step */
};

/* ----------------------------------------------------------- bit packing -- */

/* 16 bytes holding 0/1 -> 16 bits, first byte -> bit 0.  (w & 0x01010101) * 0x01020408 moves
 * the four byte LSBs into bits 24..27 without carries. */
__device__ __forceinline__ uint32_t pack16(uint4 v)
{
	uint32_t n0 = ((v.x & 0x01010101u) * 0x01020408u) >> 24;
	uint32_t n1 = ((v.y & 0x01010101u) * 0x01020408u) >> 24;
	uint32_t n2 = ((v.z & 0x01010101u) * 0x01020408u) >> 24;
	uint32_t n3 = ((v.w & 0x01010101u) * 0x01020408u) >> 24;
	return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}

/* 4 bits -> 4 bytes holding 0/1, bit 0 -> first byte */
__device__ __forceinline__ uint32_t unpack4(uint32_t nib)
{
	return ((nib & 0xf) * 0x00204081u) & 0x01010101u;
}

/* 16-byte unit at aligned address `ua`; bytes outside [lo, hi) read as zero */
__device__ __forceinline__ uint4 load_unit(uintptr_t ua, uintptr_t lo, uintptr_t hi)
{
	if (ua + 16 <= lo || ua >= hi)
		return make_uint4(0, 0, 0, 0);
	if (ua >= lo && ua + 16 <= hi)
		return *reinterpret_cast<const uint4 *>(ua);
	uint32_t w[4] = {0, 0, 0, 0};
	for (int i = 0; i < 16; i++) {
		uintptr_t a = ua + i;
		if (a >= lo && a < hi)
			w[i >> 2] |= (uint32_t)(*reinterpret_cast<const uint8_t *>(a)) << (8 * (i & 3));
	}
	return make_uint4(w[0], w[1], w[2], w[3]);
}

/* Warp-cooperative load of up to 1024+64 stream bits starting at byte address `p`
 * (any alignment); bytes at or beyond `end` read as zero.  On return lane l holds
 *   x0 = bits [32l, 32l+32),  x1 = the following word, x2 = the one after that,
 * i.e. everything a 38-bit pattern anchored anywhere in the lane's 32 positions needs. */
__device__ __forceinline__ void load_window(const uint8_t *p, const uint8_t *end, unsigned lane,
                                            uint32_t &x0, uint32_t &x1, uint32_t &x2)
{
	const uintptr_t pa = reinterpret_cast<uintptr_t>(p);
	const uintptr_t ea = reinterpret_cast<uintptr_t>(end);
	const uintptr_t a0 = pa & ~(uintptr_t)15;
	const unsigned d = (unsigned)(pa - a0);            /* 0..15 bits of misalignment after packing */
	/* aligned words: lane l packs bytes [32l, 32l+32); lanes 0..2 also pack words 32..34 */
	uint32_t A = pack16(load_unit(a0 + 32 * lane, pa, ea)) |
	             (pack16(load_unit(a0 + 32 * lane + 16, pa, ea)) << 16);
	uint32_t B = 0;
	if (lane < 3)
		B = pack16(load_unit(a0 + 1024 + 32 * lane, pa, ea)) |
		    (pack16(load_unit(a0 + 1024 + 32 * lane + 16, pa, ea)) << 16);
	/* realign by d bits: X_j = (A_j >> d) | (A_{j+1} << (32-d)) */
	uint32_t An = __shfl_sync(FULL, A, (lane + 1) & 31);
	uint32_t B0 = __shfl_sync(FULL, B, 0);
	uint32_t Bn = __shfl_sync(FULL, B, (lane + 1) & 31);
	if (lane == 31) An = B0;
	uint32_t XA = __funnelshift_r(A, An, d);
	uint32_t XB = __funnelshift_r(B, Bn, d);                 /* valid for lanes 0,1: words 32,33 */
	uint32_t n1 = __shfl_sync(FULL, XA, (lane + 1) & 31);
	uint32_t n2 = __shfl_sync(FULL, XA, (lane + 2) & 31);
	uint32_t b0 = __shfl_sync(FULL, XB, 0);
	uint32_t b1 = __shfl_sync(FULL, XB, 1);
	x0 = XA;
	x1 = (lane == 31) ? b0 : n1;
	x2 = (lane == 30) ? b0 : (lane == 31) ? b1 : n2;
}

/* ---- input formats of the bit stream (tb200_options.input) ------------------------------
 * BYTES   one bit per byte, the reference's file format (tetra-rx.c:82-95)
 * PACKED  eight bits per byte, stream bit i = byte i>>3 bit i&7
 * F32SYM  one float per pi/4-DQPSK symbol as the demodulators write it; the two hard bits of a symbol
 *         follow float_to_bits.c:33-72 (no AFC): >2 -> 01, (0,2] -> 00, < -2 -> 11, else 10 */
enum { IN_BYTES = 0, IN_PACKED = 1, IN_F32SYM = 2 };

/* float_to_bits.c:33-72: process_sym_fl + sym_int2bits; first bit of the symbol in bit 0 */
__device__ __forceinline__ uint32_t slice_symbol(float fl)
{
	if (fl > 2.0f) return 2u;          /* sym  3 -> bits 0,1 */
	if (fl > 0.0f) return 0u;          /* sym  1 -> bits 0,0 */
	if (fl < -2.0f) return 3u;         /* sym -3 -> bits 1,1 */
	return 1u;                         /* sym -1 -> bits 1,0 (also NaN, like the reference's comparisons) */
}

/* 32-bit word i of a bit-packed buffer of nbytes bytes (4-byte aligned base); bytes past the end read as zero */
__device__ __forceinline__ uint32_t packed_word(const uint8_t *base, uint64_t i, uint64_t nbytes)
{
	if (4 * (i + 1) <= nbytes) return reinterpret_cast<const uint32_t *>(base)[i];
	uint32_t v = 0;
	for (uint64_t b = 4 * i; b < nbytes; ++b) v |= (uint32_t)base[b] << (8 * (b - 4 * i));
	return v;
}

/* 32 stream bits [pos, pos+32) of a PACKED / F32SYM buffer that starts at stream bit 0 of `base`
 * and holds `avail` bits; bits at or beyond avail read as zero.  Any pos; base 4-byte aligned. */
__device__ __forceinline__ uint32_t fetch32(const uint8_t *base, int fmt, uint64_t pos, uint64_t avail)
{
	if (pos >= avail) return 0;
	uint32_t v;
	if (fmt == IN_PACKED) {
		/* the caller's buffer ends with the byte that holds bit avail-1: a word that reaches past it is put
		 * together from its bytes (sub-allocated buffers, compute-sanitizer) */
		const uint64_t i = pos >> 5, nbytes = (avail + 7) >> 3;
		const uint32_t lo = packed_word(base, i, nbytes), hi = packed_word(base, i + 1, nbytes);
		v = __funnelshift_r(lo, hi, (uint32_t)(pos & 31));
	} else {
		const float *f = reinterpret_cast<const float *>(base);
		const uint64_t s0 = pos >> 1, ns = (avail + 1) >> 1;
		uint64_t acc = 0;
#pragma unroll
		for (int t = 0; t < 17; ++t)
			if (s0 + t < ns) acc |= (uint64_t)slice_symbol(f[s0 + t]) << (2 * t);
		v = (uint32_t)(acc >> (pos & 1));
	}
	const uint64_t left = avail - pos;
	if (left < 32) v &= (1u << left) - 1;
	return v;
}

/* load_window for the PACKED / F32SYM formats: same result layout */
__device__ __forceinline__ void load_window_fmt(const uint8_t *base, int fmt, uint64_t pos, uint64_t avail, unsigned lane,
                                                uint32_t &x0, uint32_t &x1, uint32_t &x2)
{
	const uint32_t XA = fetch32(base, fmt, pos + 32 * lane, avail);
	const uint32_t XB = lane < 2 ? fetch32(base, fmt, pos + 1024 + 32 * lane, avail) : 0u;
	uint32_t n1 = __shfl_sync(FULL, XA, (lane + 1) & 31);
	uint32_t n2 = __shfl_sync(FULL, XA, (lane + 2) & 31);
	uint32_t b0 = __shfl_sync(FULL, XB, 0);
	uint32_t b1 = __shfl_sync(FULL, XB, 1);
	x0 = XA;
	x1 = (lane == 31) ? b0 : n1;
	x2 = (lane == 30) ? b0 : (lane == 31) ? b1 : n2;
}

/* --------------------------------------------- training sequence search -- */

__device__ __forceinline__ uint32_t low_mask(long long lim)   /* bits 0..lim set, none if lim < 0 */
{
	if (lim < 0) return 0;
	if (lim >= 31) return 0xffffffffu;
	return (2u << lim) - 1;
}

/* tetra_find_train_seq (tetra_burst.c:269-339) over the window [p, p+W) for the sequences
 * enabled in `mask` (SYNC, NORM_1, NORM_2; the uplink ones are never enabled by the
 * receiver, tetra_burst_sync.c:76,118-120).  Returns the type or -1; *off = first offset.
 * On return from the first 1024-bit pass lane l < 16 also holds word l of the slot in *slotw. */
template <typename LOAD>
__device__ __forceinline__ int find_train_seq_core(LOAD load, unsigned W,
                                                   uint32_t mask, const Tables *__restrict__ tab,
                                                   unsigned *off, uint32_t *slotw)
{
	const unsigned lane = threadIdx.x & 31;
	const bool en_y = mask & (1u << TS_SYNC), en_n = mask & (1u << TS_NORM_1), en_p = mask & (1u << TS_NORM_2);
	int rc = -1;
	unsigned found_off = 0;
	for (unsigned it = 0; it * 1024 < W; ++it) {
		uint32_t x0, x1, x2;
		load(it, lane, x0, x1, x2);
		if (it == 0 && slotw) *slotw = x0;
		uint32_t My = 0xffffffffu, Mn = 0xffffffffu, Mp = 0xffffffffu;
#pragma unroll
		for (int b = 0; b < 22; ++b) {
			uint32_t s = __funnelshift_r(x0, x1, b);
			My &= ((SEQ_Y >> b) & 1) ? s : ~s;
			Mn &= ((SEQ_N >> b) & 1) ? s : ~s;
			Mp &= ((SEQ_P >> b) & 1) ? s : ~s;
		}
#pragma unroll
		for (int b = 22; b < 32; ++b) {
			uint32_t s = __funnelshift_r(x0, x1, b);
			My &= ((SEQ_Y >> b) & 1) ? s : ~s;
		}
#pragma unroll
		for (int b = 32; b < 38; ++b) {
			uint32_t s = __funnelshift_r(x1, x2, b - 32);
			My &= ((SEQ_Y >> b) & 1) ? s : ~s;
		}
		/* remain_len gate: a sequence of length L at offset o needs o + L <= W */
		const long long rel = (long long)W - 1024ll * it - 32ll * lane;   /* bits left from this lane's bit 0 */
		My &= en_y ? low_mask(rel - 38) : 0;
		Mn &= en_n ? low_mask(rel - 22) : 0;
		Mp &= en_p ? low_mask(rel - 22) : 0;
		if (it == 0 && lane == 0) {
			/* pre-filter blind spot for offsets 0..20 */
			const uint32_t prev = x0 << 1;
			const uint32_t keep_hi = 0xffe00000u;
			My &= keep_hi | (prev & tab->blind_ok[0][1]) | (~prev & tab->blind_ok[0][0]);
			Mn &= keep_hi | (prev & tab->blind_ok[1][1]) | (~prev & tab->blind_ok[1][0]);
			Mp &= keep_hi | (prev & tab->blind_ok[2][1]) | (~prev & tab->blind_ok[2][0]);
		}
		const uint32_t Mall = My | Mn | Mp;
		const unsigned vote = __ballot_sync(FULL, Mall != 0);
		if (vote) {
			const int src = __ffs((int)vote) - 1;
			const unsigned i0 = (unsigned)__ffs((int)Mall) - 1;      /* garbage on lanes without a match */
			const int ty = ((My >> (i0 & 31)) & 1) ? TS_SYNC : ((Mn >> (i0 & 31)) & 1) ? TS_NORM_1 : TS_NORM_2;
			const unsigned packed = (i0 & 31) | ((unsigned)ty << 8);
			const unsigned got = __shfl_sync(FULL, packed, src);
			found_off = 1024 * it + 32 * (unsigned)src + (got & 31);
			rc = (int)(got >> 8);
			break;
		}
	}
	*off = found_off;
	return rc;
}

__device__ inline int find_train_seq_warp(const uint8_t *p, const uint8_t *end, unsigned W,
                                          uint32_t mask, const Tables *__restrict__ tab,
                                          unsigned *off, uint32_t *slotw)
{
	const uint8_t *wend = p + W < end ? p + W : end;
	return find_train_seq_core([&](unsigned it, unsigned lane, uint32_t &x0, uint32_t &x1, uint32_t &x2) {
		load_window(p + 1024 * it, wend, lane, x0, x1, x2);
	}, W, mask, tab, off, slotw);
}

/* the same over a buffer in any input format: the window starts at stream bit `pos` of a buffer that
 * holds `avail` bits from its stream bit 0 */
__device__ inline int find_train_seq_warp_fmt(const uint8_t *base, int fmt, uint64_t pos, uint64_t avail, unsigned W,
                                              uint32_t mask, const Tables *__restrict__ tab,
                                              unsigned *off, uint32_t *slotw)
{
	if (fmt == IN_BYTES)
		return find_train_seq_warp(base + pos, base + avail, W, mask, tab, off, slotw);
	const uint64_t wend = pos + W < avail ? pos + W : avail;
	return find_train_seq_core([&](unsigned it, unsigned lane, uint32_t &x0, uint32_t &x1, uint32_t &x2) {
		load_window_fmt(base, fmt, pos + 1024ull * it, wend, lane, x0, x1, x2);
	}, W, mask, tab, off, slotw);
}

/* ------------------------------------------------------------- scrambler -- */

/* lane l < 16 returns word l of the scrambling sequence for `init` (tetra_scramb.c:66-85) */
__device__ __forceinline__ uint32_t lfsr_word(uint32_t init, unsigned lane, const Tables *__restrict__ tab)
{
	uint32_t w = 0;
	const unsigned l = lane & 15;
#pragma unroll 8
	for (int b = 0; b < 32; ++b)
		w ^= ((init >> b) & 1) ? tab->lfsr_col[b][l] : 0u;
	return w;
}

/* --------------------------------- descramble + de-interleave (+ de-puncture) -- */

/* The fused stage: type-3 bit j of the block = burst[place(m)] ^ lfsr[m], m = A(j+1) mod K
 * (tetra_scramb.c:77-85 then tetra_interleave.c:51-60); bits are produced 32 at a time,
 * one per lane, and assembled with a ballot.  De-puncturing (tetra_conv_enc.c:226-248) is
 * pure addressing: Viterbi step t reads type-3 bits 3(t/2) + {0,1} (t even) or + 2 (t odd). */
template <int BT, int PL>
__device__ __forceinline__ void gather_type3(const uint32_t *bw, const uint32_t *lf, uint32_t *t3, unsigned lane)
{
	constexpr int K = Blk<BT>::K, A = Blk<BT>::A;
	constexpr int NW = (K + 31) / 32;
#pragma unroll
	for (int i = 0; i < NW; ++i) {
		const unsigned j = lane + 32 * i;
		const unsigned m = (A * (j + 1)) % K;
		const unsigned pos = place<PL>(m);
		const uint32_t bit = ((bw[pos >> 5] >> (pos & 31)) ^ (lf[m >> 5] >> (m & 31))) & 1;
		const uint32_t w = __ballot_sync(FULL, (j < (unsigned)K) && bit);
		if (lane == 0) t3[i] = w;
	}
	if (lane == 0) { t3[NW] = 0; }
	__syncwarp();
}

/* ------------------------------------------------ Viterbi, warp-shuffle form -- */

/* K=5 16-state decoder with the reference's semantics (viterbi.c:6-25, viterbi_cch.c:58-66,
 * libosmocore osmo_conv_decode): hard inputs with erasures, cost = disagreeing non-erased
 * symbols, start state 0, N data steps + 4 erased flush steps, trace back from state 0,
 * equal cost keeps the predecessor with the older bit 0 (state s>>1).
 * Lane = state.  Lanes 0..15 run the trellis of t3a; lanes 16..31 run t3b when `two`,
 * otherwise they mirror lanes 0..15.  One ballot per step packs the 16 (or 32) survivor
 * decisions into a word of `dec` for the bit-packed trace back. */
template <int N>
__device__ inline void viterbi_warp(const uint32_t *t3a, const uint32_t *t3b, bool two,
                                    uint32_t *dec, uint32_t *t2a, uint32_t *t2b, bool tie_hi)
{
	const unsigned lane = threadIdx.x & 31, s = lane & 15, half = lane >> 4, base = two ? (lane & 16) : 0;
	const uint32_t *t3 = (two && half) ? t3b : t3a;
	const unsigned b = s & 1, p0 = s >> 1, p1 = p0 | 8;
	const unsigned o = mother_out(p0, b);
	const unsigned g1 = (o >> 3) & 1, g2 = (o >> 2) & 1;
	const unsigned E = g1 | (g2 << 1) | (g1 << 2);
	uint32_t pm = s ? (1u << 16) : 0;
	const unsigned src0 = base | p0, src1 = base | p1;

	for (int q = 0; q < (N + 4) / 2; ++q) {
		const bool data = q < N / 2;
		uint32_t v = 0;
		if (data) {
			const unsigned bp = 3 * q;
			v = (__funnelshift_r(t3[bp >> 5], t3[(bp >> 5) + 1], bp & 31) & 7) ^ E;
		}
		{   /* even step: G1, G2 received */
			const uint32_t m0 = data ? (((v & 3) + 1) >> 1) : 0, m1 = data ? 2 - m0 : 0;
			const uint32_t c0 = __shfl_sync(FULL, pm, src0) + m0;
			const uint32_t c1 = __shfl_sync(FULL, pm, src1) + m1;
			const bool d = tie_hi ? c1 <= c0 : c1 < c0;      /* include/tetra_tie_rule.h */
			pm = d ? c1 : c0;
			const uint32_t bal = __ballot_sync(FULL, d);
			if (lane == 0) dec[2 * q] = bal;
		}
		{   /* odd step: G1 received */
			const uint32_t m0 = data ? (v >> 2) : 0, m1 = data ? 1 - m0 : 0;
			const uint32_t c0 = __shfl_sync(FULL, pm, src0) + m0;
			const uint32_t c1 = __shfl_sync(FULL, pm, src1) + m1;
			const bool d = tie_hi ? c1 <= c0 : c1 < c0;      /* include/tetra_tie_rule.h */
			pm = d ? c1 : c0;
			const uint32_t bal = __ballot_sync(FULL, d);
			if (lane == 0) dec[2 * q + 1] = bal;
		}
	}
	__syncwarp();

	/* trace back: every lane of a half walks the same path, the half's lane 0 stores */
	uint32_t *t2 = (two && half) ? t2b : t2a;
	const bool writer = (s == 0) && (two || half == 0);
	unsigned st = 0;
	uint32_t acc = 0;
	for (int t = N + 3; t >= 0; --t) {
		const uint32_t w = dec[t];
		if (t < N) {
			acc |= (st & 1) << (t & 31);
			if ((t & 31) == 0) {
				if (writer) t2[t >> 5] = acc;
				acc = 0;
			}
		}
		st = (st >> 1) | (((w >> (base + st)) & 1) << 3);
	}
	__syncwarp();
}

/* CRC-16-CCITT over type-2 bits [0, L) (crc_simple.c:103-106), 16 lanes per block:
 * linear form, each lane folds the table entries of its bits, XOR-reduce. */
__device__ __forceinline__ uint32_t crc16_half(const uint32_t *t2, int L, int crci, const Tables *__restrict__ tab)
{
	const unsigned s = threadIdx.x & 15;
	uint32_t acc = 0;
	for (int i = s; i < L; i += 16) {
		const uint32_t bit = (t2[i >> 5] >> (i & 31)) & 1;
		acc ^= bit ? (uint32_t)tab->crc_pow[L - 1 - i] : 0u;
	}
	acc ^= __shfl_xor_sync(FULL, acc, 8);
	acc ^= __shfl_xor_sync(FULL, acc, 4);
	acc ^= __shfl_xor_sync(FULL, acc, 2);
	acc ^= __shfl_xor_sync(FULL, acc, 1);
	return acc ^ tab->crc_init[crci];
}

/* -------------------------------------------- Viterbi, one lane per block -- */

/* Same decoder, but a single thread owns the whole trellis: 16 path metrics in registers,
 * 8 butterflies per step, no shuffles.  Differential metric form: branch cost for the
 * (j, input 0) / (j+8, input 1) branches is +d, for the other two -d, with
 * d = mismatches(out_j) - mismatches(~out_j); identical arg-min and identical ties as the
 * mismatch count.  Decisions go to `dec` (16 bits per step, 2 steps per word). */
struct LaneVit {
	int pm[16];
	__device__ __forceinline__ void init()
	{
#pragma unroll
		for (int i = 0; i < 16; ++i) pm[i] = i ? (1 << 20) : 0;
	}
	/* d[j] for butterflies j = 0..7, from the (G1,G2) classes of viterbi_cch.c:35-40:
	 * j: 0 1 2 3 4 5 6 7 -> (G1,G2) of out(j,0): 00 10 01 11 01 11 00 10 */
	__device__ __forceinline__ uint32_t step(const int dA, const int dB, const int dC, const int dD, const bool tie_hi)
	{
		/* dA: class 00, dB: class 10, dC: class 01, dD: class 11 */
		const int dj[8] = { dA, dB, dC, dD, dC, dD, dA, dB };
		int nm[16];
		uint32_t dec = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int a0 = pm[j] + dj[j], a1 = pm[j + 8] - dj[j];
			const int b0 = pm[j] - dj[j], b1 = pm[j + 8] + dj[j];
			const bool e = tie_hi ? a1 <= a0 : a1 < a0, f = tie_hi ? b1 <= b0 : b1 < b0;
			nm[2 * j] = e ? a1 : a0;
			nm[2 * j + 1] = f ? b1 : b0;
			dec |= (e ? 1u : 0u) << (2 * j);
			dec |= (f ? 1u : 0u) << (2 * j + 1);
		}
#pragma unroll
		for (int i = 0; i < 16; ++i) pm[i] = nm[i];
		return dec;
	}
};

/* t3: packed type-3 bits (own copy, any memory), dec: N/2+2 words, out: packed type-2 bits */
template <int N>
__device__ inline void viterbi_lane(const uint32_t *t3, uint32_t *dec, uint32_t *out, bool tie_hi)
{
	LaneVit v;
	v.init();
	for (int q = 0; q < N / 2; ++q) {
		const unsigned bp = 3 * q;
		const uint32_t r = __funnelshift_r(t3[bp >> 5], t3[(bp >> 5) + 1], bp & 31) & 7;
		const int r1 = r & 1, r2 = (r >> 1) & 1, r3 = (r >> 2) & 1;
		/* even step: d = m(out) - m(~out) = 2*m(out) - 2 over (G1,G2) */
		const int m00 = r1 + r2, m10 = (1 - r1) + r2, m01 = r1 + (1 - r2), m11 = 2 - m00;
		uint32_t d0 = v.step(2 * m00 - 2, 2 * m10 - 2, 2 * m01 - 2, 2 * m11 - 2, tie_hi);
		/* odd step: only G1: d = 2*m - 1 */
		const int n0 = r3, n1 = 1 - r3;
		uint32_t d1 = v.step(2 * n0 - 1, 2 * n1 - 1, 2 * n0 - 1, 2 * n1 - 1, tie_hi);
		dec[q] = d0 | (d1 << 16);
	}
	for (int q = N / 2; q < N / 2 + 2; ++q) {
		uint32_t d0 = v.step(0, 0, 0, 0, tie_hi);
		uint32_t d1 = v.step(0, 0, 0, 0, tie_hi);
		dec[q] = d0 | (d1 << 16);
	}
	unsigned st = 0;
	uint32_t acc = 0;
	for (int t = N + 3; t >= 0; --t) {
		const uint32_t w = dec[t >> 1] >> ((t & 1) * 16);
		if (t < N) {
			acc |= (st & 1) << (t & 31);
			if ((t & 31) == 0) { out[t >> 5] = acc; acc = 0; }
		}
		st = (st >> 1) | (((w >> st) & 1) << 3);
	}
}

/* bit-serial CRC for the lane form (crc_simple.c:65-82) */
__device__ inline uint32_t crc16_serial(const uint32_t *t2, int L)
{
	uint32_t crc = 0xffff;
	for (int i = 0; i < L; ++i) {
		const uint32_t bit = (t2[i >> 5] >> (i & 31)) & 1;
		const uint32_t top = ((crc >> 15) ^ bit) & 1;
		crc = (crc << 1) & 0xffff;
		if (top) crc ^= 0x1021;
	}
	return crc;
}

/* ------------------------------------------------------ output assembly -- */

/* OR `len` bits of src (bit 0 first) into the slot's type-1 string at bit `dst`.
 * Lanes 0..8 each own one output word; src needs one readable word past the data. */
__device__ __forceinline__ void put_bits(uint32_t *outw, int dst, const uint32_t *src, int len, unsigned lane)
{
	if (lane < 9) {
		const int sidx = 32 * (int)lane - dst;
		uint32_t val = 0;
		if (sidx >= 0) {
			if (sidx < len)
				val = __funnelshift_r(src[sidx >> 5], src[(sidx >> 5) + 1], sidx & 31);
		} else if (sidx > -32) {
			val = src[0] << (-sidx);
		}
		const int hi = len - sidx;          /* output bits i < hi map inside the source */
		if (hi <= 0) val = 0;
		else if (hi < 32) val &= (1u << hi) - 1;
		outw[lane] |= val;
	}
}

/* store the slot's type-1 string: unpacked (one bit per byte, 288-byte stride) and/or packed */
__device__ __forceinline__ void store_type1(const uint32_t *outw, uint8_t *type1, uint32_t *type1_packed,
                                            uint64_t k, unsigned lane)
{
	if (type1 && lane < 18) {
		const uint32_t h = (outw[lane >> 1] >> (16 * (lane & 1))) & 0xffff;
		uint4 v = make_uint4(unpack4(h), unpack4(h >> 4), unpack4(h >> 8), unpack4(h >> 12));
		*reinterpret_cast<uint4 *>(type1 + k * TYPE1_STRIDE + 16 * lane) = v;
	}
	if (type1_packed && lane < 9)
		type1_packed[k * TYPE1_WORDS + lane] = outw[lane];
}

__device__ __forceinline__ uint32_t extract_bits(const uint32_t *w, unsigned pos, unsigned n)  /* n <= 32 */
{
	uint32_t v = __funnelshift_r(w[pos >> 5], w[(pos >> 5) + 1], pos & 31);
	return n >= 32 ? v : (v & ((1u << n) - 1));
}

/* reverse the low n bits (type-1 bit 0 is the MSB of a protocol field, tetra_common.c:31-39) */
__device__ __forceinline__ uint32_t field_msb_first(const uint32_t *w, unsigned pos, unsigned n)
{
	return __brev(extract_bits(w, pos, n)) >> (32 - n);
}

/* =================================================================== kernels == */

/* The modelled tetra_burst_sync_in() calls of the current RUN of equal-length reads (tetra-rx.c:82-95 reads 64 bytes
 * at a time, a pipe may hand out less): calls c_base+1, c_base+2, ... deliver `chunk` bits each, so after call c the
 * receiver has been given T(c) = min(t_base + (c - c_base) * chunk, n_end) bits.  A stream fed with changing read sizes
 * is a sequence of such runs (one library call per run). */
struct CallGeom {
	uint64_t c_base;           /* calls made before the run */
	uint64_t t_base;           /* bits they delivered */
	uint64_t n_end;            /* bits delivered when the run's last call is done */
	uint32_t chunk;            /* bits per call of the run */
	uint32_t pad;
};

/* the call that processes a slot ending at stream bit end_bit: the first one that has delivered it, but one slot per
 * call (tetra_burst_sync.c:107-150), i.e. not before call cfloor */
TB_HD inline uint64_t call_for(const CallGeom &cg, uint64_t end_bit, uint64_t cfloor)
{
	uint64_t need = cg.c_base;
	if (end_bit > cg.t_base) {
		const uint64_t x = end_bit - cg.t_base + cg.chunk - 1;
		need += cg.chunk == 64 ? x >> 6 : (x <= 0xffffffffull ? (uint64_t)((uint32_t)x / cg.chunk) : x / cg.chunk);
	}
	return need > cfloor ? need : cfloor;
}

TB_HD inline uint64_t bits_at_call(const CallGeom &cg, uint64_t c)
{
	const uint64_t t = cg.t_base + (c - cg.c_base) * cg.chunk;
	return t < cg.n_end ? t : cg.n_end;
}

struct RxGeom {
	const uint8_t *bits;       /* device buffer holding stream bits [base_bit, base_bit + n_bytes) */
	uint64_t n_bytes;          /* stream BITS available from base_bit (= bytes in the 1-bit-per-byte format) */
	uint64_t base_bit;
	uint64_t a0;               /* absolute bit of slot 0 of this launch */
	uint64_t cmin;             /* first tetra_burst_sync_in() call that may process slot 0 */
	CallGeom cg;               /* the modelled calls */
	uint32_t n_slots;
	int fmt;                   /* IN_BYTES / IN_PACKED / IN_F32SYM: how `bits` encodes the stream (base_bit % 128 == 0 unless bytes) */
	int tie_hi;                /* Viterbi tie rule (include/tetra_tie_rule.h), used by the warp-form SB1 decode */
};

/* bits the search sees for slot k: bits_in_buf when the slot is processed
 * (tetra_burst_sync.c:107-120 with one slot per call and `chunk` new bits per call) */
__device__ __forceinline__ unsigned slot_window(const RxGeom &g, uint64_t k, uint64_t ak)
{
	return (unsigned)(bits_at_call(g.cg, call_for(g.cg, ak + SLOT_BITS, g.cmin + k)) - ak);
}

/* Pass 1, one warp per slot: load + pack the slot, search the training sequence with the
 * reference's first-match semantics, classify, and decode SB1 of SYNC bursts (its scrambling
 * code is fixed, so it needs no cell state).  Leaves SlotWs + the packed slot for pass 2. */
template <bool DO_SB1>
__global__ void __launch_bounds__(256)
k_classify(RxGeom g, const Tables *__restrict__ tab, SlotWs *__restrict__ ws, uint32_t *__restrict__ slot_bits)
{
	__shared__ WarpSmem sm[8];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	WarpSmem &S = sm[wib];
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	const uint8_t *end = g.bits + g.n_bytes;

	for (uint64_t k = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib; k < g.n_slots; k += nwarps) {
		const uint64_t ak = g.a0 + (uint64_t)SLOT_BITS * k;
		const unsigned W = slot_window(g, k, ak);
		const uint8_t *p = g.bits + (ak - g.base_bit);
		unsigned off = 0;
		uint32_t xw = 0;
		const uint32_t mask = (1u << TS_SYNC) | (1u << TS_NORM_1) | (1u << TS_NORM_2);
		const int rc = find_train_seq_warp(p, end, W, mask, tab, &off, &xw);

		/* tetra_burst_sync.c:121-143 */
		int kind = KIND_NONE;
		bool unlock = false;
		if (rc == TS_SYNC) { if (off == 214) kind = KIND_SB; else unlock = true; }
		else if (rc == TS_NORM_1) { if (off == 244) kind = KIND_NDB_F; }
		else if (rc == TS_NORM_2) { if (off == 244) kind = KIND_NDB_2; }
		else unlock = true;

		if (lane < 16) {
			/* bits past the slot's 510 belong to the next slot */
			if (lane == 15) xw &= 0x3fffffffu;
			slot_bits[k * 16 + lane] = xw;
			S.bw[lane] = xw;
		}
		uint32_t good = 0, t1lo = 0, t1hi = 0, code = 0;
		uint32_t tn = 0, fn = 0, mn = 0, cc = 0, mcc = 0, mnc = 0, sb1_crc = 0;
		if (DO_SB1 && kind == KIND_SB) {
			if (lane < 16) S.lf[lane] = tab->lfsr_sb1[lane];
			if (lane < 4) S.bw[16 + lane] = 0;
			__syncwarp();
			gather_type3<0, PL_SB1>(S.bw, S.lf, S.t3[0], lane);
			viterbi_warp<80>(S.t3[0], S.t3[0], false, S.dec, S.t2[0], S.t2[0], g.tie_hi != 0);
			const uint32_t crc = crc16_half(S.t2[0], 76, 0, tab);
			good = (crc == 0x1d0f);
			sb1_crc = crc;
			t1lo = S.t2[0][0];
			t1hi = S.t2[0][1] & 0x0fffffffu;
			/* SYNC PDU fields, tetra_lower_mac.c:291-297 */
			cc  = field_msb_first(S.t2[0], 4, 6);
			tn  = field_msb_first(S.t2[0], 10, 2) + 1;
			fn  = field_msb_first(S.t2[0], 12, 5);
			mn  = field_msb_first(S.t2[0], 17, 6);
			mcc = field_msb_first(S.t2[0], 31, 10);
			mnc = field_msb_first(S.t2[0], 41, 14);
			code = scramb_init_from(mcc, mnc, cc);
		}
		if (lane == 0) {
			SlotWs w;
			w.sb1_t1[0] = t1lo; w.sb1_t1[1] = t1hi;
			w.sb_code = code;
			w.find_off = (uint16_t)off; w.window = (uint16_t)W;
			w.find_rc = (int8_t)rc; w.good_sb = (uint8_t)good; w.kind = (uint8_t)kind; w.unlock = unlock;
			w.tn = (uint8_t)tn; w.fn = (uint8_t)fn; w.mn = (uint8_t)mn; w.cc = (uint8_t)cc;
			w.mcc = (uint16_t)mcc; w.mnc = (uint16_t)mnc; w.sb1_crc = sb1_crc;
			ws[k] = w;
		}
		__syncwarp();
	}
}

/* Scan, step a: per block of 1024 slots, running "index of the latest CRC-good SB1 at or
 * before me" (or -1), the block's last one, and the first slot that loses lock.
 * 256 threads, four consecutive slots each: the block is one round trip to memory and two barriers, so what counts is how
 * many blocks an SM has in flight (8 of these; the earlier 1024-thread form had 2 and took 7 waves for a 2^21-slot piece). */
constexpr int SCAN_BLOCK = 1024, SCAN_THREADS = 256, SCAN_PER = SCAN_BLOCK / SCAN_THREADS;
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_blocks(const SlotWs *__restrict__ ws, uint32_t n_slots, int32_t *__restrict__ last_good,
              int32_t *__restrict__ blk_last, uint32_t *__restrict__ first_unlock, uint32_t *__restrict__ first_good,
              uint32_t *__restrict__ kind_count, uint32_t *__restrict__ kind_list, uint32_t list_stride)
{
	constexpr int NW = SCAN_THREADS / 32;
	__shared__ int32_t warp_last[NW];
	__shared__ uint32_t kcnt[4][NW];     /* slots of kind c in warp w, then exclusive offsets */
	__shared__ uint32_t kbase[4];
	const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint64_t k0 = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)SCAN_PER * tid;
	/* the byte quartet find_rc | good_sb | kind | unlock of every slot (SlotWs bytes 16..19) */
	uint32_t q[SCAN_PER];
#pragma unroll
	for (int i = 0; i < SCAN_PER; ++i)
		q[i] = k0 + i < n_slots ? reinterpret_cast<const uint32_t *>(ws + k0 + i)[4] : 0xffffffffu;
	int32_t incl[SCAN_PER];              /* running maximum inside the thread */
	int kind[SCAN_PER];
	uint32_t cnt[4] = { 0, 0, 0, 0 }, rank[SCAN_PER];
	int32_t run = -1;
#pragma unroll
	for (int i = 0; i < SCAN_PER; ++i) {
		const bool have = k0 + i < n_slots;
		const bool good = have && ((q[i] >> 8) & 0xffu) != 0;
		kind[i] = have ? (int)((q[i] >> 16) & 0xffu) : -1;
		if (have && (q[i] >> 24) != 0) atomicMin(first_unlock, (uint32_t)(k0 + i));
		if (good) run = (int32_t)(k0 + i);
		incl[i] = run;
		rank[i] = 0;
#pragma unroll
		for (int c = 0; c < 4; ++c)
			if (kind[i] == c) { rank[i] = cnt[c]; cnt[c]++; }
	}
	/* inclusive max-scan of the threads' maxima inside the warp */
	int32_t v = run;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int32_t o = __shfl_up_sync(FULL, v, d);
		if (lane >= (unsigned)d && o > v) v = o;
	}
	int32_t before = __shfl_up_sync(FULL, v, 1);        /* what precedes this thread inside the warp */
	if (lane == 0) before = -1;
	if (lane == 31) warp_last[w] = v;
	/* slots grouped by kind (one list per TB200_KIND_*), so that the lane decode pass can give every warp
	 * blocks of one length: exclusive sums of the per-thread counts inside the warp, warp offsets by a scan, one atomic
	 * per kind and thread block.  Lists keep the slots of a block in stream order. */
	uint32_t pre[4] = { 0, 0, 0, 0 };
	if (kind_count) {
#pragma unroll
		for (int c = 0; c < 4; ++c) {
			uint32_t x = cnt[c];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(FULL, x, d);
				if (lane >= (unsigned)d) x += o;
			}
			pre[c] = x - cnt[c];
			if (lane == 31) kcnt[c][w] = x;
		}
	}
	__syncthreads();
	if (w == 0) {
		int32_t x = lane < NW ? warp_last[lane] : -1;
#pragma unroll
		for (int d = 1; d < NW; d <<= 1) {
			const int32_t o = __shfl_up_sync(FULL, x, d);
			if (lane >= (unsigned)d && o > x) x = o;
		}
		if (lane < NW) warp_last[lane] = x;
	} else if (kind_count && w <= 4) {
		/* warps 1..4: exclusive scan of the per-warp counts of kind w-1, block total -> one atomic */
		const int c = (int)w - 1;
		const uint32_t mine = lane < NW ? kcnt[c][lane] : 0u;
		uint32_t x = mine;
#pragma unroll
		for (int d = 1; d < NW; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL, x, d);
			if (lane >= (unsigned)d) x += o;
		}
		if (lane < NW) kcnt[c][lane] = x - mine;
		if (lane == NW - 1) kbase[c] = x ? atomicAdd(&kind_count[c], x) : 0u;
	}
	__syncthreads();
	if (w > 0) { const int32_t o = warp_last[w - 1]; if (o > before) before = o; }
#pragma unroll
	for (int i = 0; i < SCAN_PER; ++i) {
		if (k0 + i >= n_slots) break;
		const int32_t lg = incl[i] > before ? incl[i] : before;
		last_good[k0 + i] = lg;
		/* first CRC-good SB1 of the piece: only a block's first one competes */
		if (incl[i] == (int32_t)(k0 + i) && before < 0 && (i == 0 || incl[i - 1] < 0)) atomicMin(first_good, (uint32_t)(k0 + i));
		if (kind_count && kind[i] >= 0) {
			uint32_t at = rank[i];
#pragma unroll
			for (int c = 0; c < 4; ++c)
				if (kind[i] == c) at += kbase[c] + kcnt[c][w] + pre[c];
			kind_list[(size_t)kind[i] * list_stride + at] = (uint32_t)(k0 + i);
		}
	}
	if (tid == SCAN_THREADS - 1) blk_last[blockIdx.x] = run > before ? run : before;
}

/* Scan, step b: exclusive running max over the block results (single thread block). */
__global__ void __launch_bounds__(1024)
k_scan_prefix(const int32_t *__restrict__ blk_last, uint32_t n_blocks, int32_t *__restrict__ blk_prev)
{
	__shared__ int32_t warp_last[32];
	__shared__ int32_t running;
	const unsigned tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	if (tid == 0) running = -1;
	__syncthreads();
	for (uint32_t base = 0; base < n_blocks; base += 1024) {
		const uint32_t i = base + tid;
		int32_t v = i < n_blocks ? blk_last[i] : -1;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int32_t o = __shfl_up_sync(FULL, v, d);
			if (lane >= (unsigned)d && o > v) v = o;
		}
		if (lane == 31) warp_last[w] = v;
		__syncthreads();
		if (w == 0) {
			int32_t x = warp_last[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int32_t o = __shfl_up_sync(FULL, x, d);
				if (lane >= (unsigned)d && o > x) x = o;
			}
			warp_last[lane] = x;
		}
		__syncthreads();
		int32_t incl = v;
		if (w > 0 && warp_last[w - 1] > incl) incl = warp_last[w - 1];
		const int32_t run = running;
		if (run > incl) incl = run;
		/* exclusive: what precedes block i */
		int32_t excl = __shfl_up_sync(FULL, incl, 1);
		if (lane == 0) excl = (w > 0) ? (warp_last[w - 1] > run ? warp_last[w - 1] : run) : run;
		if (i < n_blocks) blk_prev[i] = excl;
		__syncthreads();
		if (tid == 1023) running = incl;
		__syncthreads();
	}
}

/* cell state in force for slot k (tetra_lower_mac.c:167,283-302 + tetra_burst_sync.c:113):
 * X_k = PDU time of a CRC-good SB in this slot, else one slot after X_{k-1}; the scrambling
 * code is the one announced by the latest CRC-good SB1 at or before k. */
/* returns true when the state came from the carry (no CRC-good SB1 in this launch at or before k) */
__device__ __forceinline__ bool cell_state(uint64_t k, const SlotWs *__restrict__ ws,
                                           const int32_t *__restrict__ last_good,
                                           const int32_t *__restrict__ blk_prev,
                                           const DevCarry *__restrict__ carry, Tm *tm, uint32_t *code)
{
	int32_t j = last_good[k];
	if (j < 0) j = blk_prev[k >> 10];
	if (j >= 0) {
		const SlotWs s = ws[j];
		Tm t = { s.tn, s.fn, s.mn };
		*tm = tm_advance_small(t, (uint32_t)(k - (uint64_t)j));      /* k indexes the slots of one launch: < 2^32 */
		*code = s.sb_code;
		return false;
	}
	Tm t = { carry->tn, carry->fn, carry->mn };
	*tm = k + 1 < 0x7fffffffull ? tm_advance_small(t, (uint32_t)(k + 1)) : tm_advance(t, k + 1);
	*code = carry->scramb_init;
	return true;
}

/* after the launch's last valid slot: the state the next launch starts from */
__global__ void k_finalize_carry(const SlotWs *__restrict__ ws, const int32_t *__restrict__ last_good,
                                 const int32_t *__restrict__ blk_prev, uint32_t n_valid, DevCarry *carry,
                                 const uint32_t *__restrict__ first_good, uint64_t k_base)
{
	if (threadIdx.x != 0 || blockIdx.x != 0 || n_valid == 0) return;
	const uint64_t k = n_valid - 1;
	Tm tm; uint32_t code;
	cell_state(k, ws, last_good, blk_prev, carry, &tm, &code);
	int32_t j = last_good[k];
	if (j < 0) j = blk_prev[k >> 10];
	DevCarry c = *carry;
	c.scramb_init = code; c.tn = tm.tn; c.fn = tm.fn; c.mn = tm.mn;
	if (j >= 0) { c.mcc = ws[j].mcc; c.mnc = ws[j].mnc; c.cc = ws[j].cc; }
	if (j >= 0 && !c.seen_good) {
		/* first_good may name a slot behind n_valid when the piece was cut at a lock loss and issued whole before;
		 * it is recomputed by the scan of the cut piece, so it is below n_valid here */
		c.seen_good = 1;
		c.first_good = k_base + (first_good ? (uint64_t)*first_good : (uint64_t)j);
	}
	*carry = c;
}

struct DecodeArgs {
	const SlotWs *ws;
	const uint32_t *slot_bits;
	const int32_t *last_good;
	const int32_t *blk_prev;
	const DevCarry *carry;
	const Tables *tab;
	SlotOut *slots;
	uint8_t *type1;
	uint32_t *type1_packed;
	uint64_t a0;
	uint64_t out_base;       /* index of this launch's slot 0 in the output arrays */
	uint32_t n_slots;
	const uint32_t *kind_count;   /* [4] slots per TB200_KIND_* (lane form only) */
	const uint32_t *kind_list;    /* [4][list_stride] their indices */
	uint32_t list_stride;
	uint32_t *crc;                /* optional: CRC-16 registers per slot, block A (SB1 / SCH-F / BLK1) | block B (SB2 / BLK2) << 16 */
	int tie_hi;                   /* Viterbi tie rule (warp form; the lane kernels are templates) */
	unsigned long long *stats;    /* [3] counters of this piece: slots handed to the lower MAC, primitives, CRC-good blocks */
	uint32_t *aach;               /* optional: RM(30,14) decoding of the slot's AACH (rm3014_decode), ~0 for slots without one */
	const uint32_t *rm_leader;    /* its coset-leader table */
	int skip_dependent;           /* sharded decode, first pass: leave out the slots whose cell state would come from a carry-in that
	                               * is not known yet (no CRC-good SB1 since the shard began); they are decoded once it is */
};

/* what a slot adds to the counters (tb200_stats): packed as bursts | primitives << 8 | CRC-good blocks << 16 */
__device__ __forceinline__ uint32_t slot_counts(int kind, uint32_t flags)
{
	if (kind == KIND_NONE) return 0;
	return 1u | ((kind == KIND_NDB_F ? 2u : 3u) << 8) | ((uint32_t)__popc(flags & (F_CRC_A | F_CRC_B)) << 16);
}

/* warp-wide sum of the packed per-lane counts (each field stays below 256: at most two slots per lane), one
 * atomic per field and warp */
__device__ __forceinline__ void add_counts(unsigned long long *stats, uint32_t v)
{
#ifdef TB_SIMT_EMULATION
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
#else
	v = __reduce_add_sync(FULL, v);
#endif
	if ((threadIdx.x & 31) == 0 && v) {
		atomicAdd(&stats[0], (unsigned long long)(v & 0xffu));
		atomicAdd(&stats[1], (unsigned long long)((v >> 8) & 0xffu));
		atomicAdd(&stats[2], (unsigned long long)(v >> 16));
	}
}

/* Pass 2 (warp-shuffle Viterbi form), one warp per slot: everything of tp_sap_udata_ind
 * (tetra_lower_mac.c:143-357) that depends on the cell state: BBK, SB2, SCH/F, BLK1+BLK2. */
__global__ void __launch_bounds__(256)
k_decode_warp(DecodeArgs a)
{
	__shared__ WarpSmem sm[8];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	WarpSmem &S = sm[wib];
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	const Tables *__restrict__ tab = a.tab;

	for (uint64_t k = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib; k < a.n_slots; k += nwarps) {
		const SlotWs w = a.ws[k];
		Tm tm; uint32_t code;
		const bool dep = cell_state(k, a.ws, a.last_good, a.blk_prev, a.carry, &tm, &code);
		if (a.skip_dependent && dep && !a.carry->seen_good) continue;          /* warp-uniform */
		const int kind = w.kind;
		uint32_t flags = (uint32_t)kind | (w.unlock ? F_UNLOCK : 0);
		uint32_t crcs = 0, bb30 = 0;

		if (lane < 16) S.bw[lane] = a.slot_bits[k * 16 + lane];
		if (lane < 4) S.bw[16 + lane] = 0;
		if (lane < 12) S.outw[lane] = 0;
		if (kind != KIND_NONE) {
			const uint32_t lw = lfsr_word(code, lane, tab);
			if (lane < 16) S.lf[lane] = lw;
		}
		__syncwarp();

		if (kind == KIND_SB) {
			/* SB1 | BBK | SB2 (tetra_burst.c:347-353) */
			if (w.good_sb) flags |= F_CRC_A;
			if (lane == 0) {
				S.sb1[0] = w.sb1_t1[0]; S.sb1[1] = w.sb1_t1[1]; S.sb1[2] = 0;
				S.bbk[0] = (extract_bits(S.bw, 252, 14) ^ S.lf[0]) & 0x3fff; S.bbk[1] = 0;
				bb30 = (extract_bits(S.bw, 252, 30) ^ S.lf[0]) & 0x3fffffffu;
			}
			gather_type3<1, PL_BLK2>(S.bw, S.lf, S.t3[0], lane);
			viterbi_warp<144>(S.t3[0], S.t3[0], false, S.dec, S.t2[0], S.t2[0], a.tie_hi != 0);
			const uint32_t crc = crc16_half(S.t2[0], 140, 1, tab);
			if (crc == 0x1d0f) flags |= F_CRC_B;
			crcs = (w.sb1_crc & 0xffffu) | (crc << 16);
			if (tm_is_bnch(tm)) flags |= F_BNCH;
			put_bits(S.outw, 0, S.sb1, 60, lane);
			put_bits(S.outw, 60, S.bbk, 14, lane);
			put_bits(S.outw, 74, S.t2[0], 124, lane);
		} else if (kind == KIND_NDB_F) {
			/* BBK | SCH/F (tetra_burst.c:363-373) */
			if (lane == 0) {
				S.bbk[0] = (extract_bits(S.bw, 230, 14) ^ S.lf[0]) & 0x3fff; S.bbk[1] = 0;
				bb30 = ((extract_bits(S.bw, 230, 14) | (extract_bits(S.bw, 266, 16) << 14)) ^ S.lf[0]) & 0x3fffffffu;
			}
			gather_type3<5, PL_SCHF>(S.bw, S.lf, S.t3[0], lane);
			viterbi_warp<288>(S.t3[0], S.t3[0], false, S.dec, S.t2[0], S.t2[0], a.tie_hi != 0);
			const uint32_t crc = crc16_half(S.t2[0], 284, 2, tab);
			if (crc == 0x1d0f) flags |= F_CRC_A;
			crcs = crc & 0xffffu;
			put_bits(S.outw, 0, S.bbk, 14, lane);
			put_bits(S.outw, 14, S.t2[0], 268, lane);
		} else if (kind == KIND_NDB_2) {
			/* BBK | BLK1 | BLK2 (tetra_burst.c:354-362), the two halves decode side by side */
			if (lane == 0) {
				S.bbk[0] = (extract_bits(S.bw, 230, 14) ^ S.lf[0]) & 0x3fff; S.bbk[1] = 0;
				bb30 = ((extract_bits(S.bw, 230, 14) | (extract_bits(S.bw, 266, 16) << 14)) ^ S.lf[0]) & 0x3fffffffu;
			}
			gather_type3<1, PL_BLK1>(S.bw, S.lf, S.t3[0], lane);
			gather_type3<1, PL_BLK2>(S.bw, S.lf, S.t3[1], lane);
			viterbi_warp<144>(S.t3[0], S.t3[1], true, S.dec, S.t2[0], S.t2[1], a.tie_hi != 0);
			const uint32_t crc = crc16_half(S.t2[lane >> 4], 140, 1, tab);
			const uint32_t okA = __shfl_sync(FULL, crc == 0x1d0f, 0), okB = __shfl_sync(FULL, crc == 0x1d0f, 16);
			if (okA) flags |= F_CRC_A;
			if (okB) flags |= F_CRC_B;
			crcs = (__shfl_sync(FULL, crc, 0) & 0xffffu) | (__shfl_sync(FULL, crc, 16) << 16);
			put_bits(S.outw, 0, S.bbk, 14, lane);
			put_bits(S.outw, 14, S.t2[0], 124, lane);
			put_bits(S.outw, 138, S.t2[1], 124, lane);
		}
		__syncwarp();
		const uint64_t ko = a.out_base + k;
		store_type1(S.outw, a.type1, a.type1_packed, ko, lane);
		if (lane == 0) {
			SlotOut o;
			o.slot_bit = (uint32_t)(a.a0 + (uint64_t)SLOT_BITS * k);
			o.scrambling_code = code;
			o.find_off = w.find_off; o.window = w.window;
			o.time = (uint16_t)(tm.tn | (tm.fn << 3) | (tm.mn << 8));
			o.find_rc = w.find_rc; o.flags = (uint8_t)flags;
			a.slots[ko] = o;
			if (a.crc) a.crc[ko] = crcs;
			if (a.aach) a.aach[ko] = kind == KIND_NONE ? 0xffffffffu : rm3014_decode(tab, a.rm_leader, rm3014_word_from_air(bb30));
		}
		if (a.stats) add_counts(a.stats, lane == 0 ? slot_counts(kind, flags) : 0u);
		__syncwarp();
	}
}

}  // namespace tb
