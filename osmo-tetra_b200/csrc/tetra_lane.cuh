/*
 * tetra_lane.cuh - the throughput form of the decode pass: one THREAD owns two coded blocks.
 *
 * The warp-per-slot kernels of tetra_kernels.cuh spend ~20k warp instructions per SCH/F burst,
 * almost all of it shuffles and lane-redundant work around a 16-state trellis that cannot use
 * 32 lanes.  Here the trellis lives entirely in one thread's registers and two independent
 * trellises are packed into the 16-bit halves of each register, so every add-compare-select
 * is one add + one VIADDMNMX.U16x2 for two blocks at once.  The survivor decisions are not
 * extracted per step: a tag below the metric's unit makes the packed minimum carry them along
 * (acs2_step), so that after 4 (or 8) steps the low nibble (byte) of every metric is the history
 * of its survivor; histories go to a per-CTA scratch area in global memory and the trace back
 * walks them one group at a time.  Descrambling is a handful of word XORs in the burst domain
 * and de-interleaving + de-puncturing is a fully unrolled constant-index bit gather.
 *
 * Semantics are those of the reference chain (tetra_lower_mac.c:143-357, viterbi.c:6-25,
 * libosmocore osmo_conv_decode): see viterbi_warp() in tetra_kernels.cuh for the rules.
 */
#pragma once
#include "tetra_kernels.cuh"
#include "tetra_async.cuh"

namespace tb {

constexpr int LANE_DEC_ROWS = 296;     /* words per thread: 73 groups of 4 trellis steps x uint4, padded */
constexpr int LANE_T3_ROWS = 14;
constexpr int LANE_NT = 32;           /* threads per CTA of the lane kernels: one warp, 5 CTAs fit an SM's shared memory */

/* dynamic shared memory of a lane kernel with NT threads, in 32-bit words.  The survivor
 * decisions (296 x 4 B per thread) do NOT live here: at 37.9 KB per warp they would cap an SM at
 * five warps; they go to a per-CTA scratch area in global memory that stays L2 resident
 * (written once, read once a few microseconds later, 128-byte coalesced rows). */
__host__ __device__ constexpr size_t lane_smem_words(int nt)
{
	return (size_t)2 * LANE_T3_ROWS * nt + 256 + 16 + (nt / 32) * 16 + 1024;
}
__host__ __device__ constexpr size_t lane_scratch_words_per_cta(int nt) { return (size_t)LANE_DEC_ROWS * nt; }

struct LaneSmem {
	uint4 *dec;          /* [LANE_DEC_ROWS / 4][NT] in GLOBAL scratch: survivor histories, one uint4 per thread and group of 4 steps */
	uint32_t *t3;        /* [2][LANE_T3_ROWS][NT] type-3 bits, later the decoded type-2 bits */
	uint32_t *crc_tab;   /* [256] reflected CRC-CCITT byte table, then [16] nibble table */
	uint32_t *lfb;       /* [NT/32][16] per-warp scrambling sequence broadcast */
	uint32_t *leap;      /* [4][256] the scrambler 32 steps at a time (Tables::lfsr_leap) */
	static constexpr int nt = LANE_NT;
	__device__ __forceinline__ LaneSmem(uint8_t *base, uint32_t *scratch)
	{
		uint32_t *p = reinterpret_cast<uint32_t *>(base);
		dec = reinterpret_cast<uint4 *>(scratch + (size_t)blockIdx.x * lane_scratch_words_per_cta(nt));
		t3 = p; p += 2 * LANE_T3_ROWS * nt;
		crc_tab = p; p += 256 + 16;
		lfb = p; p += (nt / 32) * 16;
		leap = p;
	}
	__device__ __forceinline__ uint32_t *t3col(int tr, int tid) const { return t3 + (size_t)tr * LANE_T3_ROWS * nt + tid; }
};

/* class of the (G1,G2) outputs of branch (state j, input 0): idx = 2*G1 + G2 */
__host__ __device__ constexpr unsigned branch_class(unsigned j)
{
	return ((((j >> 0) & 1) ^ ((j >> 3) & 1)) << 1) | (((j >> 1) & 1) ^ ((j >> 2) & 1) ^ ((j >> 3) & 1));
}

__host__ __device__ constexpr unsigned rev4(unsigned v)
{
	return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3);
}

/* a * 16 + b as one multiply-add on the FMA pipe (opaque, so that the optimiser does not turn the
 * nibble packing into shifts + masks on the ALU pipe, which is the one the ACS loop saturates) */
__device__ __forceinline__ uint32_t mad16_opaque(uint32_t a, uint32_t b)
{
#ifdef TB_SIMT_EMULATION
	return a * 16 + b;
#else
	uint32_t r;
	asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(r) : "r"(a), "r"(b));
	return r;
#endif
}

/* a * K + c as an integer multiply-add with a non-unit K.  The ACS loop saturates the ALU pipe (VIADDMNMX,
 * LOP3, SHF live there) while the multiply-add pipe has room, so the branch metrics are built from the raw
 * 0/1 received bits with the metric scale (16 per mismatch) folded into the multiplier: ptxas keeps such
 * multiply-adds on the FMA pipe (it turns x*1+c and x*-1+c back into ALU adds). */
template <int K>
__device__ __forceinline__ uint32_t mad_k(uint32_t a, uint32_t c)
{
#ifdef TB_SIMT_EMULATION
	return a * (uint32_t)K + c;
#else
	uint32_t r;
	asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "n"(K), "r"(c));
	return r;
#endif
}

__device__ __forceinline__ uint32_t sub_opaque(uint32_t a, uint32_t b)
{
#ifdef TB_SIMT_EMULATION
	return a - b;
#else
	uint32_t r;
	asm("sub.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	return r;
#endif
}

/* One trellis step for two packed trellises (16-bit halves of every register).
 *
 * Path metrics are kept in units of 16 (one mismatching symbol costs 16), which leaves the low
 * four bits of every metric free.  Step q = t mod 4 of a group of four adds 2^q to the candidate
 * that comes from the older-bit-1 predecessor (s>>1)|8.  The bits below 2^q of any metric only
 * hold tags of earlier steps of the group, i.e. they are < 2^q, so a single packed add-min
 *   - compares the true metrics exactly,
 *   - gives ties to the predecessor s>>1 (the reference rule: libosmocore keeps the first/lower
 *     predecessor on equal cost), and
 *   - leaves in bit q of the winner "which predecessor won", while bits < q are inherited from
 *     the winner: after four steps the low nibble of a state's metric is the decision history of
 *     its survivor path through the group (register exchange for free, no per-step extraction).
 * The trellis has 4 bits of memory, so that nibble IS the state the survivor had four steps
 * earlier (bit-reversed), and also the four decoded bits of the previous group.
 * M0[c] = 16 * mismatches of output class c (per half); the complementary branch has class c ^ 3.
 * Cost: one add + one VIADDMNMX.U16x2 per state and step for two trellises. */
/* TIE_HI (include/tetra_tie_rule.h, the other tie rule): the tag goes on the candidate from s>>1 instead, ties
 * then go to (s>>1)|8 and the history bits come out complemented (take_history* flips them back). */
template <bool TIE_HI>
__device__ __forceinline__ void acs2_step(uint32_t (&pm)[16], const uint32_t (&M0)[4], const uint32_t (&M1)[4])
{
	uint32_t nm[16];
#pragma unroll
	for (int s = 0; s < 16; ++s) {
		const unsigned c = branch_class(s >> 1) ^ ((s & 1) ? 3u : 0u);
		if (TIE_HI) {
			const uint32_t c1 = pm[(s >> 1) | 8] + M0[c ^ 3];
			nm[s] = __viaddmin_u16x2(pm[s >> 1], M1[c], c1);
		} else {
			const uint32_t c1 = pm[(s >> 1) | 8] + M1[c ^ 3];
			nm[s] = __viaddmin_u16x2(pm[s >> 1], M0[c], c1);
		}
	}
#pragma unroll
	for (int s = 0; s < 16; ++s) pm[s] = nm[s];
}

/* The odd trellis steps (only G1 was sent, rate 2/3) with the metrics taken relative to the branches whose G1 is 0:
 * those cost nothing, the others d = U * (1 - 2 r) (U = metric unit, r = received bit), i.e. +U or -U.  Every path
 * collects the same offset per step, so nothing changes for the comparisons, but half of the candidates need no add at all:
 *   G1(lo branch) = 0:  new = min(pm[hi] + (d + tag), pm[lo])            one VIADDMNMX
 *   G1(lo branch) = 1:  new = min(pm[lo] + d, pm[hi] + tag)              one add with a constant + one VIADDMNMX
 * 24 instead of 32 instructions per step for two trellises.  d can be negative; VIADDMNMX.U16x2 adds half by half, so d
 * and d + tag come in two's complement PER HALF (mad with U * 254 instead of -2 U: the 2^16 of a negative low half lands
 * where it cancels the borrow), and the metrics carry a floor that the -U steps cannot eat up (callers).
 * TIE_HI mirrors it: the tag sits on the lo candidate, the free candidates are the ones whose hi branch has G1 = 0. */
template <bool TIE_HI>
__device__ __forceinline__ void acs2_odd_step(uint32_t (&pm)[16], const uint32_t d, const uint32_t dT, uint32_t tagw)
{
#if !defined(TB_SIMT_EMULATION) && !defined(TB_ODD_TAG_IMMEDIATE)
	/* the tag as a register the optimiser cannot see through (threadIdx.y is 0, but only at run time): register + register
	 * adds go to the multiply-add pipe (IMAD.IADD), register + immediate ones to the ALU pipe (VIADD), and that is the
	 * pipe the VIADDMNMX already saturate */
	tagw += threadIdx.y;
#endif
	uint32_t nm[16];
#pragma unroll
	for (int s = 0; s < 16; ++s) {
		const unsigned c = branch_class(s >> 1) ^ ((s & 1) ? 3u : 0u);
		const bool g1_lo = (c >> 1) & 1;                      /* G1 of the branch from s>>1; the one from (s>>1)|8 has the complement */
		const uint32_t lo = pm[s >> 1], hi = pm[(s >> 1) | 8];
		if (!TIE_HI) {
			if (!g1_lo) nm[s] = __viaddmin_u16x2(hi, dT, lo);
			else        nm[s] = __viaddmin_u16x2(lo, d, hi + tagw);
		} else {
			if (g1_lo)  nm[s] = __viaddmin_u16x2(lo, dT, hi);
			else        nm[s] = __viaddmin_u16x2(hi, d, lo + tagw);
		}
	}
#pragma unroll
	for (int s = 0; s < 16; ++s) pm[s] = nm[s];
}

/* multiplier that turns the received bit r (per half, carried with weight W) into -2 U r in two's complement per half:
 * (2^16 - 2 U) / W; added to U that is U (1 - 2 r), and nothing carries from the low half into the high one */
__host__ __device__ constexpr int odd_mult(int unit, int weight) { return (65536 - 2 * unit) / weight; }

/* After the fourth step of a group: strip the history nibbles off the 16 metrics and pack them.
 * The nibble of state s goes to position rev4(s), so that the trace back can index it with the
 * previous group's decoded nibble directly.  Result: x,y = trellis X positions 0-7, 8-15; z,w = Y. */
template <bool TIE_HI>
__device__ __forceinline__ uint4 take_history(uint32_t (&pm)[16])
{
	uint32_t W[4];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		uint32_t acc = 0;
#pragma unroll
		for (int i = 3; i >= 0; --i) {
			const unsigned s = rev4(4 * j + i);
			const uint32_t h = pm[s] & 0x000f000fu;
			pm[s] = sub_opaque(pm[s], h);
			acc = i == 3 ? h : mad16_opaque(acc, h);
		}
		W[j] = TIE_HI ? ~acc : acc;
	}
	return make_uint4(__byte_perm(W[0], W[1], 0x5410), __byte_perm(W[2], W[3], 0x5410),
	                  __byte_perm(W[0], W[1], 0x7632), __byte_perm(W[2], W[3], 0x7632));
}

/* nibble f (0..15) of the 64-bit value hi:lo */
__device__ __forceinline__ uint32_t nibble64(uint32_t lo, uint32_t hi, uint32_t f)
{
	const uint32_t src = (f & 8) ? hi : lo;
	return (src >> ((f & 7) * 4)) & 15u;
}

/* Forward pass + trace back for the two blocks whose type-3 bits sit in sm.t3[0] / sm.t3[1]
 * (columns of this thread).  nx, ny: type-2 lengths (0 = no block); nmax: warp-wide maximum so
 * the loop is uniform.  Leaves the decoded type-2 bits in the same columns.
 * dec: this thread's column of the history scratch, one uint4 per group of four steps. */
template <bool MASKED, bool TIE_HI>
__device__ __noinline__ void viterbi_pair_t(uint4 *dec, uint32_t *cx, uint32_t *cy, int nx, int ny, int nmax)
{
	uint32_t pm[16];
	/* start in state 0 (osmo_conv_decode); the floor 0x1000 is what the odd steps may take off (at most 16 each, 146 of them) */
#pragma unroll
	for (int i = 0; i < 16; ++i) pm[i] = i ? 0x50005000u : 0x10001000u;
	constexpr int nt = LANE_NT;
	const int groups = nmax / 8;                       /* 4 step pairs = 12 type-3 bits per group */
	for (int g = 0; g < groups; ++g) {
		const unsigned bp = 12 * g, w = bp >> 5, sh = bp & 31;
		const uint32_t vx = __funnelshift_r(cx[w * nt], cx[(w + 1) * nt], sh) & 0xfffu;
		const uint32_t vy = __funnelshift_r(cy[w * nt], cy[(w + 1) * nt], sh) & 0xfffu;
		const uint32_t z = vx | (vy << 16);                 /* received bits of both trellises, 12 per half */
		const uint32_t nz = ~z;
		uint32_t live = 0x00010001u;
		if (MASKED) live = ((8 * g < nx) ? 0x0001u : 0u) | ((8 * g < ny) ? 0x00010000u : 0u);
#pragma unroll
		for (int p = 0; p < 4; ++p) {
			/* raw 0/1 per half: r1, r2 = the two symbols of the even step, r3 = the one of the odd step;
			 * q1 = 1 - r1.  Blocks that have ended in this lane (MASKED) see no symbols: all costs zero. */
			const uint32_t r1 = (z >> (3 * p)) & live, q1 = (nz >> (3 * p)) & live;
			const uint32_t r2 = (z >> (3 * p + 1)) & live;
			const uint32_t r3 = (z >> (3 * p + 2)) & live;
			const uint32_t t = r1 + r2;                       /* mismatches if 00 was sent: 0..2 */
			const uint32_t u = q1 + r2;                       /* mismatches if G1=1, G2=0 was sent */
			const uint32_t two = MASKED ? live * 32u : 0x00200020u, one = MASKED ? live * 16u : 0x00100010u;
			uint32_t M0[4], M1[4];
			const uint32_t te = (p & 1) ? 0x00040004u : 0x00010001u;     /* tag of step q = 0 / 2 */
			const uint32_t to = (p & 1) ? 0x00080008u : 0x00020002u;     /* tag of step q = 1 / 3 */
			M0[0] = mad_k<16>(t, 0u);           M1[0] = mad_k<16>(t, te);
			M0[3] = mad_k<-16>(t, two);         M1[3] = mad_k<-16>(t, two + te);
			M0[2] = mad_k<16>(u, 0u);           M1[2] = mad_k<16>(u, te);
			M0[1] = mad_k<-16>(u, two);         M1[1] = mad_k<-16>(u, two + te);
			acs2_step<TIE_HI>(pm, M0, M1);
			/* odd step: only G1 was sent; 16 * (1 - 2 r3) per half, two's complement per half (16 * 254 = 65536 / 16 - 32) */
			acs2_odd_step<TIE_HI>(pm, mad_k<odd_mult(16, 1)>(r3, one), mad_k<odd_mult(16, 1)>(r3, one + to), to);
			if (p & 1) dec[(2 * g + (p >> 1)) * nt] = take_history<TIE_HI>(pm);
		}
	}
	{
		/* four flush steps: no received symbols, so no cost; both inputs stay allowed, the trace back
		 * starts in state 0, which only the all-zero tail can reach */
		const uint32_t Z0[4] = { 0, 0, 0, 0 };
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const uint32_t tq = 0x00010001u << q;
			const uint32_t Z1[4] = { tq, tq, tq, tq };
			acs2_step<TIE_HI>(pm, Z0, Z1);
		}
		dec[(nmax >> 2) * nt] = take_history<TIE_HI>(pm);
	}
	/* Trace back, one group (= one decoded nibble) per step.  f = decoded nibble of group g = bit-reversed
	 * state after the group; the history nibble at position f of group g is the decoded nibble of group
	 * g-1.  The block with n type-2 bits ends (after its flush group n/4) in state 0. */
	uint32_t fx = 0, fy = 0, hx = 0, hy = 0;
	/* output word wi holds the nibbles of groups 8wi .. 8wi+7, read from the histories of groups 8wi+1 .. 8wi+8 */
	const int gtop = nmax >> 2;                         /* index of the flush group of a full-length block */
	for (int wi = (nmax - 1) >> 5; wi >= 0; --wi) {
		const int ghi = 8 * wi + 8 < gtop ? 8 * wi + 8 : gtop;
		const uint4 *dp = dec + ghi * nt;
		const int cnt = ghi - 8 * wi;                      /* 4 or 8 */
		for (int i = 0; i < cnt; i += 4) {
			uint4 h4[4];
#pragma unroll
			for (int u = 0; u < 4; ++u) h4[u] = dp[-u * nt];     /* loads first: L2 latency overlaps */
			dp -= 4 * nt;
#pragma unroll
			for (int u = 0; u < 4; ++u) {
				const uint4 h = h4[u];
				if (MASKED) {
					const int g = ghi - i - u;
					if (4 * g <= nx) { fx = nibble64(h.x, h.y, fx); hx = hx * 16 + fx; }
					if (4 * g <= ny) { fy = nibble64(h.z, h.w, fy); hy = hy * 16 + fy; }
				} else {
					fx = nibble64(h.x, h.y, fx); hx = hx * 16 + fx;
					fy = nibble64(h.z, h.w, fy); hy = hy * 16 + fy;
				}
			}
		}
		cx[wi * nt] = hx;
		cy[wi * nt] = hy;
	}
}

/* ---- the same decoder for warps whose lanes all carry two blocks of the same length (the normal case
 * once the slots are grouped by kind): groups of EIGHT steps.  Metrics in units of 256, tags 2^q for
 * q = 0..7, so the low BYTE of a state's metric is the decision history of its survivor through eight
 * steps = the eight decoded bits that end four steps before the group does, and its low nibble is (bit
 * reversed) the state the survivor had when the group began.  History extraction, packing and the trace
 * back cost half as much per step as in the four-step form; the mismatch counts of two step pairs are
 * formed field-wise in one word (step_pair_u8<256> / <32>).  65 535 / 256 = 255 mismatches fit a metric,
 * a 288-step block can collect 432, so every eighth group the common minimum is subtracted (metrics only
 * ever get compared, so this changes nothing).  The four flush steps stay a four-step group. ---- */

/* strip the history bytes; word i of the result holds positions 4i..4i+3 (position = rev4(state)).
 * One PRMT pulls the two history bytes (trellis X, trellis Y) of two states into one word, one AND per
 * state clears them. */
template <bool TIE_HI>
__device__ __forceinline__ void take_history8(uint32_t (&pm)[16], uint4 &hx, uint4 &hy)
{
	uint32_t W[8];          /* W[j]: positions 2j, 2j+1; low half = trellis X, high half = trellis Y */
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const unsigned s1 = rev4(2 * j + 1), s0 = rev4(2 * j);
		W[j] = __byte_perm(pm[s0], pm[s1], 0x6240);     /* s0.b0, s1.b0, s0.b2, s1.b2 */
		if (TIE_HI) W[j] = ~W[j];
		pm[s1] &= 0xff00ff00u;
		pm[s0] &= 0xff00ff00u;
	}
	hx = make_uint4(__byte_perm(W[0], W[1], 0x5410), __byte_perm(W[2], W[3], 0x5410),
	                __byte_perm(W[4], W[5], 0x5410), __byte_perm(W[6], W[7], 0x5410));
	hy = make_uint4(__byte_perm(W[0], W[1], 0x7632), __byte_perm(W[2], W[3], 0x7632),
	                __byte_perm(W[4], W[5], 0x7632), __byte_perm(W[6], W[7], 0x7632));
}

/* byte f (0..15) of the 128-bit value h */
__device__ __forceinline__ uint32_t byte128(const uint4 &h, uint32_t f)
{
	const uint32_t lo = (f & 4) ? h.y : h.x, hi = (f & 4) ? h.w : h.z;
	const uint32_t w = (f & 8) ? hi : lo;
	return (w >> ((f & 3) * 8)) & 0xffu;
}

/* two trellis steps (an even one with two received symbols, an odd one with one) of the eight-step form.
 * t, u, r3: mismatch counts per half, scaled so that count * K = 256 per mismatch; p = pair index in the group. */
template <int K, bool TIE_HI>
__device__ __forceinline__ void step_pair_u8(uint32_t (&pm)[16], uint32_t t, uint32_t u, uint32_t r3, int p)
{
	const uint32_t to = 0x00020002u << (2 * p);
	uint32_t te = 0x00010001u << (2 * p);
#if !defined(TB_SIMT_EMULATION) && defined(TB_EVEN_TAG_REG)
	/* as in acs2_odd_step: a tag the optimiser cannot fold keeps "metric + tag" a register + register add on the
	 * multiply-add pipe instead of a VIADD with an immediate on the ALU pipe */
	te += threadIdx.y;
#endif
	uint32_t M0[4], M1[4];
	M0[0] = mad_k<K>(t, 0u);                M1[0] = M0[0] + te;
	M0[3] = mad_k<-K>(t, 0x02000200u);      M1[3] = M0[3] + te;
	M0[2] = mad_k<K>(u, 0u);                M1[2] = M0[2] + te;
	M0[1] = mad_k<-K>(u, 0x02000200u);      M1[1] = M0[1] + te;
	acs2_step<TIE_HI>(pm, M0, M1);
	/* odd step: only G1 was sent; 256 * (1 - 2 r3) per half (r3 carries weight 256 / K) */
	acs2_odd_step<TIE_HI>(pm, mad_k<odd_mult(256, 256 / K)>(r3, 0x01000100u), mad_k<odd_mult(256, 256 / K)>(r3, 0x01000100u + to), to);
}

template <bool TIE_HI>
__device__ __noinline__ void viterbi_pair_u8(uint4 *dec, uint32_t *cx, uint32_t *cy, int n)
{
	uint32_t pm[16];
	/* start in state 0 (osmo_conv_decode); the floor 0x2400 covers what the odd steps take off between two re-centrings
	 * (at most 256 each, 32 of them) */
#pragma unroll
	for (int i = 0; i < 16; ++i) pm[i] = i ? 0x64006400u : 0x24002400u;
	constexpr int nt = LANE_NT;
	const int groups = n / 8;                          /* 4 step pairs = 12 type-3 bits per group */
	for (int g = 0; g < groups; ++g) {
		const unsigned bp = 12 * g, w = bp >> 5, sh = bp & 31;
		const uint32_t vx = __funnelshift_r(cx[w * nt], cx[(w + 1) * nt], sh) & 0xfffu;
		const uint32_t vy = __funnelshift_r(cy[w * nt], cy[(w + 1) * nt], sh) & 0xfffu;
		const uint32_t z = vx | (vy << 16);                 /* received bits of both trellises, 12 per half */
		/* the mismatch counts of two step pairs at a time: the symbols of pairs 2h, 2h+1 sit at bits 0-2 and 3-5;
		 * field-wise sums leave t, u of pair 2h in bits 0-1 and of pair 2h+1 in bits 3-4 (weight 8, made up for by
		 * the multiplier 32 instead of 256) */
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			const uint32_t zz = h ? (z >> 6) : z;
			const uint32_t A = zz & 0x00090009u, B = (zz >> 1) & 0x00090009u, C3 = (zz >> 2) & 0x00090009u;
			const uint32_t T = A + B;                         /* mismatches if 00 was sent: 0..2 per field */
			const uint32_t U = (A ^ 0x00090009u) + B;         /* mismatches if G1=1, G2=0 was sent */
			step_pair_u8<256, TIE_HI>(pm, T & 0x00030003u, U & 0x00030003u, C3 & 0x00010001u, 2 * h);
			step_pair_u8<32, TIE_HI>(pm, T & 0x00180018u, U & 0x00180018u, C3 & 0x00080008u, 2 * h + 1);
		}
		uint4 hx, hy;
		take_history8<TIE_HI>(pm, hx, hy);
		dec[(2 * g) * nt] = hx;
		dec[(2 * g + 1) * nt] = hy;
		if ((g & 7) == 7) {            /* keep the metrics small: bring the per-trellis minimum back to the floor */
			uint32_t m = pm[0];
#pragma unroll
			for (int i = 1; i < 16; ++i) m = __viaddmin_u16x2(pm[i], 0u, m);
			m = sub_opaque(m, 0x24002400u);
#pragma unroll
			for (int i = 0; i < 16; ++i) pm[i] = sub_opaque(pm[i], m);
		}
	}
	uint4 hf;
	{
		/* four flush steps: no received symbols, so no cost; both inputs stay allowed, the trace back
		 * starts in state 0, which only the all-zero tail can reach.  The metrics are multiples of 256
		 * here, so the four-step form (tags 1, 2, 4, 8; low nibble) applies unchanged. */
		const uint32_t Z0[4] = { 0, 0, 0, 0 };
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const uint32_t tq = 0x00010001u << q;
			const uint32_t Z1[4] = { tq, tq, tq, tq };
			acs2_step<TIE_HI>(pm, Z0, Z1);
		}
		hf = take_history<TIE_HI>(pm);
	}
	/* Trace back.  The flush group gives the last four decoded bits (= index into the last eight-step
	 * group); the byte at position f of group g holds decoded bits [8g-4, 8g+4), its low nibble is the index
	 * into group g-1.  A 64-bit shift register collects them; after group g its bit 0 is decoded bit 8g-4,
	 * so output word wi leaves as bits [4, 36) right after group 4*wi. */
	uint32_t fx = nibble64(hf.x, hf.y, 0), fy = nibble64(hf.z, hf.w, 0);
	uint32_t xlo = fx, xhi = 0, ylo = fy, yhi = 0;
	for (int wi = (n - 1) >> 5; wi >= 0; --wi) {
		const int ghi = 4 * wi + 3 < groups - 1 ? 4 * wi + 3 : groups - 1;
		const int cnt = ghi - 4 * wi + 1;                  /* 2 or 4 groups */
		const uint4 *dp = dec + (size_t)(2 * ghi) * nt;
		for (int i = 0; i < cnt; i += 2) {
			uint4 bx[2], by[2];
#pragma unroll
			for (int u = 0; u < 2; ++u) { bx[u] = dp[-2 * u * nt]; by[u] = dp[(1 - 2 * u) * nt]; }     /* loads first: latency overlaps */
			dp -= 4 * nt;
#pragma unroll
			for (int u = 0; u < 2; ++u) {
				const uint32_t b0 = byte128(bx[u], fx), b1 = byte128(by[u], fy);
				fx = b0 & 15u; fy = b1 & 15u;
				xhi = __funnelshift_l(xlo, xhi, 8); xlo = (xlo << 8) | b0;
				yhi = __funnelshift_l(ylo, yhi, 8); ylo = (ylo << 8) | b1;
			}
		}
		cx[wi * nt] = __funnelshift_r(xlo, xhi, 4);
		cy[wi * nt] = __funnelshift_r(ylo, yhi, 4);
	}
}

template <bool TIE_HI>
__device__ __forceinline__ void viterbi_pair(uint4 *dec, uint32_t *cx, uint32_t *cy, int nx, int ny, int nmax)
{
	/* warp-uniform choice: no masking work when every lane carries two full-length blocks */
	const bool uniform = __all_sync(FULL, nx == nmax && ny == nmax);
	if (uniform) viterbi_pair_u8<TIE_HI>(dec, cx, cy, nmax);
	else         viterbi_pair_t<true, TIE_HI>(dec, cx, cy, nx, ny, nmax);
}

/* CRC-16-CCITT of the first L type-2 bits of a column, reflected byte-table form of
 * crc_simple.c:65-82 (register bit-reversed, so packed LSB-first bytes feed it directly);
 * returns true when the residue is TETRA_CRC_OK (0x1d0f, tetra_common.h:69). L % 8 == 4. */
__device__ __forceinline__ uint32_t crc_col(const uint32_t *crc_tab, const uint32_t *col, int L);
/* the CRC register itself, in the reference's bit order (what tetra-rx prints as "CRC COMP: 0x%04x") */
__device__ __forceinline__ uint32_t crc_col(const uint32_t *crc_tab, const uint32_t *col, int L)
{
	uint32_t r = 0xffff;
	constexpr int nt = LANE_NT;
	const int nbytes = L >> 3;
	for (int i = 0; i < nbytes; ++i) {
		const uint32_t b = (col[(i >> 2) * nt] >> (8 * (i & 3))) & 0xff;
		r = (r >> 8) ^ crc_tab[(r ^ b) & 0xff];
	}
	const int bp = L & ~7;
	const uint32_t nib = (col[(bp >> 5) * nt] >> (bp & 31)) & 0xf;
	r = (r >> 4) ^ crc_tab[256 + ((r ^ nib) & 0xf)];
	return __brev(r) >> 16;           /* the table form keeps the register bit-reversed (0xf0b8 = good) */
}

/* ---- descrambling in the burst domain ------------------------------------------- */

template <int S>
__device__ __forceinline__ uint32_t lf_window(const uint32_t (&lf)[LANE_T3_ROWS])   /* bits [S, S+32) of the sequence */
{
	if constexpr (S <= -32 || S >= 32 * LANE_T3_ROWS) {
		return 0;
	} else if constexpr (S < 0) {
		return lf[0] << (-S);
	} else {
		constexpr int w = S >> 5, sh = S & 31;
		const uint32_t lo = lf[w];
		if constexpr (sh == 0) {
			return lo;
		} else if constexpr (w + 1 < LANE_T3_ROWS) {
			return __funnelshift_r(lo, lf[w + 1], sh);
		} else {
			return lo >> sh;
		}
	}
}

/* burst bits [DST, DST+LEN) ^= sequence bits [SRC, SRC+LEN)  (tetra_scramb.c:77-85 per block) */
template <int DST, int SRC, int LEN>
__device__ __forceinline__ void xor_region(uint32_t (&bw)[16], const uint32_t (&lf)[LANE_T3_ROWS])
{
#pragma unroll
	for (int w = 0; w < 16; ++w) {
		const int lo = 32 * w, hi = 32 * w + 32;
		if (hi <= DST || lo >= DST + LEN) continue;
		uint32_t mask = 0xffffffffu;
		if (DST > lo) mask &= 0xffffffffu << (DST - lo);
		if (DST + LEN < hi) mask &= 0xffffffffu >> (hi - (DST + LEN));
		uint32_t v = 0;
		switch (w) {   /* compile-time window selection */
#define TB_CASE(W) case W: v = lf_window<SRC + 32 * W - DST>(lf); break;
		TB_CASE(0) TB_CASE(1) TB_CASE(2) TB_CASE(3) TB_CASE(4) TB_CASE(5) TB_CASE(6) TB_CASE(7)
		TB_CASE(8) TB_CASE(9) TB_CASE(10) TB_CASE(11) TB_CASE(12) TB_CASE(13) TB_CASE(14) TB_CASE(15)
#undef TB_CASE
		}
		bw[w] ^= v & mask;
	}
}

/* ---- de-interleave + de-puncture addressing, constant indices ---------------------- */

/* type-3 bit j = (descrambled) burst bit place(A(j+1) mod K)  (tetra_interleave.c:51-60);
 * K bits written as words to the thread's column */
template <int BT, int PL>
__device__ __forceinline__ void gather_lane(const uint32_t (&bw)[16], uint32_t *col, int nt)
{
	constexpr int K = Blk<BT>::K, A = Blk<BT>::A, NW = (K + 31) / 32;
#pragma unroll
	for (int w = 0; w < NW; ++w) {
		uint32_t acc = 0;
#pragma unroll
		for (int i = 0; i < 32; ++i) {
			const int j = 32 * w + i;
			if (j < K) {
				const unsigned m = (unsigned)(A * (j + 1)) % K;
				const unsigned pos = place<PL>(m);
				const int sh = (int)(pos & 31) - i;
				const uint32_t src = bw[pos >> 5];
				const uint32_t v = sh >= 0 ? (src >> sh) : (src << (-sh));
				acc |= v & (1u << i);
			}
		}
		col[w * nt] = acc;
	}
}

/* OR `len` bits into the slot's type-1 string at bit `dst`; get(i) returns source word i */
template <typename F>
__device__ __forceinline__ void put_lane(uint32_t (&outw)[9], int dst, int len, F get)
{
#pragma unroll
	for (int w = 0; w < 9; ++w) {
		const int sidx = 32 * w - dst;
		uint32_t val = 0;
		if (sidx >= 0) {
			if (sidx < len) {
				const uint32_t lo = get(sidx >> 5), hi = get((sidx >> 5) + 1);
				val = __funnelshift_r(lo, hi, sidx & 31);
			}
		} else if (sidx > -32) {
			val = get(0) << (-sidx);
		}
		const int hi_ = len - sidx;
		if (hi_ <= 0) val = 0;
		else if (hi_ < 32) val &= (1u << hi_) - 1;
		outw[w] |= val;
	}
}

/* scrambling sequence words for `code`, for every lane of the warp.  One cell per warp (the normal case): cooperative,
 * 16 lanes build one word each from the column table and broadcast it.  Lanes with different codes (a stream whose
 * SYNC bursts announce ever new cells, BASELINE config 4): every lane runs its own scrambler 32 steps at a time with
 * the byte tables in shared memory - 14 dependent look-ups instead of one cooperative round per distinct code. */
__device__ inline void lane_lfsr(uint32_t code, bool need, uint32_t (&lf)[LANE_T3_ROWS], uint32_t *bcast,
                                 const Tables *__restrict__ tab, const uint32_t *__restrict__ leap)
{
	const unsigned lane = threadIdx.x & 31;
	const unsigned pending = __ballot_sync(FULL, need);
	if (!pending) return;
	const int leader = __ffs((int)pending) - 1;
	const uint32_t c = __shfl_sync(FULL, code, leader);
	if (__all_sync(FULL, !need || code == c)) {
		const uint32_t wv = lfsr_word(c, lane, tab);
		if (lane < 16) bcast[lane] = wv;
		__syncwarp();
		if (need) {
#pragma unroll
			for (int i = 0; i < LANE_T3_ROWS; ++i) lf[i] = bcast[i];
		}
		__syncwarp();
		return;
	}
	if (need) {
		uint32_t w = code;
#pragma unroll
		for (int i = 0; i < LANE_T3_ROWS; ++i) {
			w = lfsr_leap32(leap, w);
			lf[i] = w;
		}
	}
}

__device__ __forceinline__ void lane_load_tables(const LaneSmem &sm, const Tables *__restrict__ tab, bool with_leap = false)
{
	for (int i = threadIdx.x; i < 256 + 16; i += blockDim.x)
		sm.crc_tab[i] = tab->crc_tab_r[i];
	if (with_leap)
		for (int i = threadIdx.x; i < 1024; i += blockDim.x)
			sm.leap[i] = (&tab->lfsr_leap[0][0])[i];
	__syncthreads();
}

__device__ __forceinline__ void load_slot_bits(const uint32_t *__restrict__ slot_bits, uint64_t k, uint32_t (&bw)[16])
{
	const uint4 *p = reinterpret_cast<const uint4 *>(slot_bits + k * 16);
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const uint4 v = __ldcs(p + i);          /* read once: streaming, keeps L2 for the history scratch */
		bw[4 * i] = v.x; bw[4 * i + 1] = v.y; bw[4 * i + 2] = v.z; bw[4 * i + 3] = v.w;
	}
}

/* =============================================================== SB1 pass ==
 * SYNC bursts only: SB1 always uses scrambling code 3 (tetra_lower_mac.c:181-183), so it can be
 * decoded before the cell state is known.  Fills the SYNC-PDU part of SlotWs. */
template <bool TIE_HI>
__global__ void __launch_bounds__(32)
k_sb1_lane(SlotWs *__restrict__ ws, const uint32_t *__restrict__ slot_bits,
           const uint32_t *__restrict__ sb_list, const uint32_t *__restrict__ sb_count,
           const Tables *__restrict__ tab, uint32_t *__restrict__ scratch)
{
	LaneSmem sm(TB_DYN_SMEM(), scratch);
	lane_load_tables(sm, tab);
	const int tid = threadIdx.x;
	const uint32_t n_sb = *sb_count;                 /* SYNC bursts the classify pass listed */
	const uint32_t nthreads = gridDim.x * blockDim.x;
	const uint32_t npairs = (n_sb + 1) / 2;
	const uint32_t rounds = (npairs + nthreads - 1) / nthreads;
	uint32_t lf[LANE_T3_ROWS];
#pragma unroll
	for (int i = 0; i < LANE_T3_ROWS; ++i) lf[i] = tab->lfsr_sb1[i];

	for (uint32_t r = 0; r < rounds; ++r) {
		const uint32_t pair = r * nthreads + blockIdx.x * blockDim.x + tid;
		uint32_t k[2] = { 0, 0 };
		int n[2] = { 0, 0 };
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			if (2 * pair + h < n_sb) {
				k[h] = sb_list[2 * pair + h];
				uint32_t bw[16];
				load_slot_bits(slot_bits, k[h], bw);
				xor_region<94, 0, 120>(bw, lf);
				gather_lane<0, PL_SB1>(bw, sm.t3col(h, tid), sm.nt);
				n[h] = 80;
			}
		}
		if (!__ballot_sync(FULL, n[0] != 0)) continue;
		viterbi_pair<TIE_HI>(sm.dec + tid, sm.t3col(0, tid), sm.t3col(1, tid), n[0], n[1], 80);
#pragma unroll
		for (int h = 0; h < 2; ++h) {
			if (n[h]) {
				const uint32_t *col = sm.t3col(h, tid);
				constexpr int nt = LANE_NT;
				const uint32_t crc1 = crc_col(sm.crc_tab, col, 76);
				const bool good = crc1 == 0x1d0f;
				const uint32_t w0 = col[0], w1 = col[nt], w2 = col[2 * nt];
				const uint32_t t2[4] = { w0, w1, w2, 0 };
				SlotWs w = ws[k[h]];
				w.sb1_t1[0] = w0; w.sb1_t1[1] = w1 & 0x0fffffffu;
				w.good_sb = good;
				w.cc = (uint8_t)field_msb_first(t2, 4, 6);
				w.tn = (uint8_t)(field_msb_first(t2, 10, 2) + 1);
				w.fn = (uint8_t)field_msb_first(t2, 12, 5);
				w.mn = (uint8_t)field_msb_first(t2, 17, 6);
				w.mcc = (uint16_t)field_msb_first(t2, 31, 10);
				w.mnc = (uint16_t)field_msb_first(t2, 41, 14);
				w.sb_code = scramb_init_from(w.mcc, w.mnc, w.cc);
				w.sb1_crc = crc1;
				ws[k[h]] = w;
			}
		}
		__syncwarp();
	}
}

/* ============================================================= decode pass ==
 * Everything of tp_sap_udata_ind that needs the cell state.  A thread decodes two coded blocks at once
 * (the two packed trellises).  The slots were grouped by kind (k_scan_blocks), so a warp gets blocks of
 * one length: 288 steps for two SCH/F slots, 144 for the two halves BLK1 / BLK2 of ONE two-block slot or
 * for the SB2 blocks of two SYNC bursts; dropped slots only get their record.  Units are handed out
 * longest first.  Warps that straddle two lists run the masked form of the trellis loop.
 *
 * The pass has three parts with very different needs: PREPARE (cell state, scrambling sequence, descramble,
 * de-interleave: dependent global loads and ~2000 instructions of straight-line code per unit, few registers),
 * TRELLIS (the ACS loop + trace back: ~100 live registers, pure integer issue) and FINISH (CRC, type-1 assembly,
 * stores).  Fused into one kernel (k_decode_lane, kept as the second form) a warp holds the trellis loop's
 * registers while it waits for the loads of the other two parts: at 16 warps per SM more than half of the warps'
 * time went there (ncu r02: 53 % of the stall samples in 22 % of the instructions) and the issue slots were half
 * empty.  Split (k_lane_prepare -> k_lane_trellis [-> k_lane_finish]) every kernel runs at the occupancy its own
 * registers allow, and the type-3 bits travel between them as one 4.6 KB block per warp of units in global
 * memory ([row][lane], so that every access is a coalesced 128-byte row and the trellis kernel can pull the next
 * block into shared memory with one bulk copy while it works on the current one). */

constexpr uint32_t LANE_NO_SLOT = 0xffffffffu;
constexpr int LANE_META_ROW = 2 * LANE_T3_ROWS;                /* rows 28..35 of a unit block: see lane_unit_store */
constexpr int LANE_UNIT_ROWS = 2 * LANE_T3_ROWS + 8;
constexpr int LANE_UNIT_WORDS = LANE_UNIT_ROWS * LANE_NT;      /* 1152 words = 4608 bytes per warp of units */
__host__ __device__ constexpr size_t lane_unit_blocks(size_t slots) { return (slots + 31) / 32 + 4; }   /* worst case: every slot its own unit, + one partial warp per list */

/* what a thread knows about its unit: two slots, or the two halves of one two-block slot (then k[1] is unused) */
struct LaneUnit {
	uint32_t k[2];           /* slot index in the piece, LANE_NO_SLOT = none */
	uint32_t code[2], flags[2];
	uint32_t bbk[2];         /* the 30 descrambled AACH bits, first on air in bit 0 */
	uint32_t tm16[2];        /* tn | fn << 3 | mn << 8, the slot record's encoding */
	int kind[2], n[2];
};

__device__ __forceinline__ int lane_ncode(int n) { return n == 288 ? 2 : (n == 144 ? 1 : 0); }
__device__ __forceinline__ int lane_nsteps(uint32_t mw) { const uint32_t c = (mw >> 28) & 3u; return c == 2 ? 288 : (c == 1 ? 144 : 0); }

/* the unit's part of the block: rows 28/29 tm16 | flags << 16 | kind << 24 | step code << 28 per half, then k, code, bbk */
__device__ __forceinline__ void lane_unit_store(uint32_t *blk_lane, const LaneUnit &m)
{
	constexpr int nt = LANE_NT;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		blk_lane[(LANE_META_ROW + h) * nt] = m.tm16[h] | (m.flags[h] << 16) | ((uint32_t)m.kind[h] << 24) | ((uint32_t)lane_ncode(m.n[h]) << 28);
		blk_lane[(LANE_META_ROW + 2 + h) * nt] = m.k[h];
		blk_lane[(LANE_META_ROW + 4 + h) * nt] = m.code[h];
		blk_lane[(LANE_META_ROW + 6 + h) * nt] = m.bbk[h];
	}
}
__device__ __forceinline__ void lane_unit_load(const uint32_t *blk_lane, LaneUnit &m)
{
	constexpr int nt = LANE_NT;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t mw = blk_lane[(LANE_META_ROW + h) * nt];
		m.tm16[h] = mw & 0xffffu; m.flags[h] = (mw >> 16) & 0xffu; m.kind[h] = (int)((mw >> 24) & 0xfu); m.n[h] = lane_nsteps(mw);
		m.k[h] = blk_lane[(LANE_META_ROW + 2 + h) * nt];
		m.code[h] = blk_lane[(LANE_META_ROW + 4 + h) * nt];
		m.bbk[h] = blk_lane[(LANE_META_ROW + 6 + h) * nt];
	}
}

/* units per list: pairs of SCH/F slots, single two-block slots, pairs of SYNC bursts, pairs of dropped slots */
struct LaneLists {
	uint32_t cF, c2, cS, c0;
	uint64_t uF, u2, uS, units;
	const uint32_t *lF, *l2, *lS, *l0;
	__device__ __forceinline__ explicit LaneLists(const DecodeArgs &a)
	{
		cF = a.kind_count[KIND_NDB_F]; c2 = a.kind_count[KIND_NDB_2]; cS = a.kind_count[KIND_SB]; c0 = a.kind_count[KIND_NONE];
		uF = (cF + 1) / 2; u2 = c2; uS = (cS + 1) / 2;
		units = uF + u2 + uS + (c0 + 1) / 2;
		lF = a.kind_list + (size_t)KIND_NDB_F * a.list_stride; l2 = a.kind_list + (size_t)KIND_NDB_2 * a.list_stride;
		lS = a.kind_list + (size_t)KIND_SB * a.list_stride; l0 = a.kind_list + (size_t)KIND_NONE * a.list_stride;
	}
	__device__ __forceinline__ uint64_t warps() const { return (units + 31) / 32; }
	__device__ __forceinline__ void slots_of(uint64_t u, uint32_t (&k)[2]) const
	{
		k[0] = k[1] = LANE_NO_SLOT;
		if (u < uF)                { k[0] = lF[2 * u]; if (2 * u + 1 < cF) k[1] = lF[2 * u + 1]; }
		else if (u < uF + u2)      { k[0] = l2[u - uF]; }
		else if (u < uF + u2 + uS) { const uint64_t i = u - uF - u2; k[0] = lS[2 * i]; if (2 * i + 1 < cS) k[1] = lS[2 * i + 1]; }
		else if (u < units)        { const uint64_t i = u - uF - u2 - uS; k[0] = l0[2 * i]; if (2 * i + 1 < c0) k[1] = l0[2 * i + 1]; }
	}
};

/* PREPARE: load, descramble and de-interleave every block of the unit's slots; the type-3 bits go to the two columns
 * (stride LANE_NT words: shared memory in the fused kernel, the unit block in global memory otherwise).
 * The cell state of a slot hangs on a chain of dependent loads (list -> slot state + index of the last CRC-good SYNC burst ->
 * that burst's state); the chains of the unit's two slots are walked level by level together and the slots' bits are
 * requested as soon as their index is known, so a unit waits for three load latencies instead of eight. */
__device__ __forceinline__ void lane_prepare(const DecodeArgs &a, const Tables *__restrict__ tab, uint32_t *bcast, const uint32_t *leap,
                                             uint64_t u, const LaneLists &L, uint32_t *col0, uint32_t *col1, LaneUnit &m)
{
	constexpr int nt = LANE_NT;
	L.slots_of(u, m.k);
	/* level 2: everything addressed by the slot index */
	SlotWs w[2];
	int32_t j[2];
	uint32_t bw[2][16];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		m.code[h] = 0; m.bbk[h] = 0; m.kind[h] = KIND_NONE; m.n[h] = 0;
		j[h] = -1;
		if (m.k[h] != LANE_NO_SLOT) {
			w[h] = a.ws[m.k[h]];
			j[h] = a.last_good[m.k[h]];
			load_slot_bits(a.slot_bits, m.k[h], bw[h]);
		}
	}
	/* level 3: the CRC-good SYNC burst the slot's cell state comes from (cell_state(), spelled out for two slots) */
	Tm tm[2];
	bool dep[2];
#pragma unroll
	for (int h = 0; h < 2; ++h)
		if (m.k[h] != LANE_NO_SLOT && j[h] < 0) j[h] = a.blk_prev[m.k[h] >> 10];
	SlotWs sj[2];
#pragma unroll
	for (int h = 0; h < 2; ++h)
		if (m.k[h] != LANE_NO_SLOT && j[h] >= 0) sj[h] = a.ws[j[h]];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		tm[h].tn = tm[h].fn = tm[h].mn = 0;
		dep[h] = false;
		if (m.k[h] == LANE_NO_SLOT) continue;
		if (j[h] >= 0) {
			const Tm t = { sj[h].tn, sj[h].fn, sj[h].mn };
			tm[h] = tm_advance_small(t, m.k[h] - (uint32_t)j[h]);
			m.code[h] = sj[h].sb_code;
		} else {
			const Tm t = { a.carry->tn, a.carry->fn, a.carry->mn };
			tm[h] = tm_advance_small(t, m.k[h] + 1);          /* k indexes the slots of one launch: < 2^31 */
			m.code[h] = a.carry->scramb_init;
			dep[h] = true;
		}
	}
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		bool good_sb = false, unlock = false;
		if (m.k[h] != LANE_NO_SLOT) {
			m.kind[h] = w[h].kind; good_sb = w[h].good_sb; unlock = w[h].unlock;
			if (a.skip_dependent && dep[h] && !a.carry->seen_good) {     /* sharded decode: decoded later, with the real carry-in */
				m.k[h] = LANE_NO_SLOT; m.kind[h] = KIND_NONE; good_sb = unlock = false;
			}
		}
		m.flags[h] = (uint32_t)m.kind[h] | (unlock ? F_UNLOCK : 0) | ((m.kind[h] == KIND_SB && good_sb) ? F_CRC_A : 0);
		if (m.kind[h] == KIND_SB && tm_is_bnch(tm[h])) m.flags[h] |= F_BNCH;
		m.tm16[h] = (tm[h].tn | (tm[h].fn << 3) | (tm[h].mn << 8)) & 0xffffu;
		uint32_t lf[LANE_T3_ROWS];
		lane_lfsr(m.code[h], m.kind[h] != KIND_NONE, lf, bcast, tab, leap);
		if (m.kind[h] != KIND_NONE) {
			uint32_t *col = h ? col1 : col0;
			if (m.kind[h] == KIND_SB) {
				xor_region<252, 0, 30>(bw[h], lf);
				xor_region<282, 0, 216>(bw[h], lf);
				m.bbk[h] = extract_bits(bw[h], 252, 30);
				gather_lane<1, PL_BLK2>(bw[h], col, nt); m.n[h] = 144;
			} else if (m.kind[h] == KIND_NDB_F) {
				xor_region<14, 0, 216>(bw[h], lf);
				xor_region<282, 216, 216>(bw[h], lf);
				xor_region<230, 0, 14>(bw[h], lf);
				m.bbk[h] = extract_bits(bw[h], 230, 14);
				if (a.aach) m.bbk[h] |= ((extract_bits(bw[h], 266, 16) ^ (lf[0] >> 14)) & 0xffffu) << 14;       /* the broadcast block's second part */
				gather_lane<5, PL_SCHF>(bw[h], col, nt); m.n[h] = 288;
			} else if (h == 0) {
				/* two-block slot: BLK1 on trellis X, BLK2 on trellis Y of this thread (k[1] is unused) */
				xor_region<14, 0, 216>(bw[h], lf);
				xor_region<282, 0, 216>(bw[h], lf);
				xor_region<230, 0, 14>(bw[h], lf);
				m.bbk[h] = extract_bits(bw[h], 230, 14);
				if (a.aach) m.bbk[h] |= ((extract_bits(bw[h], 266, 16) ^ (lf[0] >> 14)) & 0xffffu) << 14;
				gather_lane<1, PL_BLK1>(bw[h], col0, nt);
				gather_lane<1, PL_BLK2>(bw[h], col1, nt);
				m.n[0] = m.n[1] = 144;
			}
		}
	}
}

/* FINISH: CRCs of the decoded blocks (columns col0 / col1, stride LANE_NT), the slot's type-1 string in the reference's
 * delivery order, the slot record and the side outputs.
 * The type-1 strings leave through `stage` (LANE_STAGE_WORDS words of the warp's shared memory, may overlay the columns):
 * a thread that stored its own slot's 288 bytes would issue 18 stores whose 32 lanes each hit a different 32-byte sector
 * half way (32 partial-sector requests per instruction, 18 LSU cycles per slot - this, not HBM, bounded the finish part);
 * staged, lane l stores chunk (32 i + l) of the warp's 64 x 18 chunks, so an instruction covers 512 contiguous bytes. */
constexpr int LANE_STAGE_STRIDE = 10;                          /* 9 type-1 words + the output index of the slot */
constexpr int LANE_STAGE_WORDS = 64 * LANE_STAGE_STRIDE;
__device__ __forceinline__ void lane_finish(const DecodeArgs &a, const Tables *__restrict__ tab, const uint32_t *crc_tab,
                                            const uint32_t *col0, const uint32_t *col1, LaneUnit &m, uint32_t *stage)
{
	constexpr int nt = LANE_NT;
	const int tid = threadIdx.x & 31;
	uint32_t crcs[2] = { 0, 0 };            /* block A | block B << 16 of each slot */
	if (m.kind[0] == KIND_NDB_2) {
		const uint32_t ca = crc_col(crc_tab, col0, 140), cb = crc_col(crc_tab, col1, 140);
		if (ca == 0x1d0f) m.flags[0] |= F_CRC_A;
		if (cb == 0x1d0f) m.flags[0] |= F_CRC_B;
		crcs[0] = ca | (cb << 16);
	}
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t *col = h ? col1 : col0;
		if (m.kind[h] == KIND_SB) {
			const uint32_t cb = crc_col(crc_tab, col, 140);
			if (cb == 0x1d0f) m.flags[h] |= F_CRC_B;
			crcs[h] = cb << 16;             /* SB1's half is added from the slot state below */
		} else if (m.kind[h] == KIND_NDB_F) {
			const uint32_t ca = crc_col(crc_tab, col, 284);
			if (ca == 0x1d0f) m.flags[h] |= F_CRC_A;
			crcs[h] = ca;
		}
	}
	/* assemble the slots' type-1 strings in the reference's delivery order; records and side outputs straight from here */
	uint32_t outw[2][9];
#pragma unroll
	for (int h = 0; h < 2; ++h) {
#pragma unroll
		for (int i = 0; i < 9; ++i) outw[h][i] = 0;
		if (m.k[h] == LANE_NO_SLOT) continue;
		const uint32_t *col = h ? col1 : col0;
		const uint32_t bb = m.bbk[h] & 0x3fffu;
		const SlotWs w = a.ws[m.k[h]];
		if (m.kind[h] == KIND_SB) {
			const uint32_t s0 = w.sb1_t1[0], s1 = w.sb1_t1[1];
			put_lane(outw[h], 0, 60, [&](int i) { return i == 0 ? s0 : (i == 1 ? s1 : 0u); });
			put_lane(outw[h], 60, 14, [&](int i) { return i == 0 ? bb : 0u; });
			put_lane(outw[h], 74, 124, [&](int i) { return col[i * nt]; });
		} else if (m.kind[h] == KIND_NDB_F) {
			put_lane(outw[h], 0, 14, [&](int i) { return i == 0 ? bb : 0u; });
			put_lane(outw[h], 14, 268, [&](int i) { return col[i * nt]; });
		} else if (m.kind[h] == KIND_NDB_2) {
			put_lane(outw[h], 0, 14, [&](int i) { return i == 0 ? bb : 0u; });
			put_lane(outw[h], 14, 124, [&](int i) { return col0[i * nt]; });
			put_lane(outw[h], 138, 124, [&](int i) { return col1[i * nt]; });
		}
		const uint64_t ko = a.out_base + m.k[h];
		SlotOut o;
		o.slot_bit = (uint32_t)(a.a0 + (uint64_t)SLOT_BITS * m.k[h]);
		o.scrambling_code = m.code[h];
		o.find_off = w.find_off; o.window = w.window;
		o.time = (uint16_t)m.tm16[h];
		o.find_rc = w.find_rc; o.flags = (uint8_t)m.flags[h];
		a.slots[ko] = o;
		if (a.crc) a.crc[ko] = crcs[h] | (m.kind[h] == KIND_SB ? (w.sb1_crc & 0xffffu) : 0u);
		if (a.aach) a.aach[ko] = m.kind[h] == KIND_NONE ? 0xffffffffu : rm3014_decode(tab, a.rm_leader, rm3014_word_from_air(m.bbk[h]));
	}
	if (a.stats)
		add_counts(a.stats, (m.k[0] != LANE_NO_SLOT ? slot_counts(m.kind[0], m.flags[0]) : 0u) +
		                    (m.k[1] != LANE_NO_SLOT && m.kind[0] != KIND_NDB_2 ? slot_counts(m.kind[1], m.flags[1]) : 0u));
	if (!a.type1 && !a.type1_packed) return;
	__syncwarp();                             /* every lane is done with the columns: the stage may overlay them */
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		uint32_t *rec = stage + (2 * tid + h) * LANE_STAGE_STRIDE;
#pragma unroll
		for (int i = 0; i < 9; ++i) rec[i] = outw[h][i];
		rec[9] = m.k[h];
	}
	__syncwarp();
	if (a.type1) {
		for (int q = tid; q < 64 * 18; q += 32) {
			const int sl = q / 18, c = q - 18 * sl;
			const uint32_t *rec = stage + sl * LANE_STAGE_STRIDE;
			const uint32_t k = rec[9];
			if (k == LANE_NO_SLOT) continue;
			const uint32_t hbits = (rec[c >> 1] >> (16 * (c & 1))) & 0xffff;
			__stcs(reinterpret_cast<uint4 *>(a.type1 + (a.out_base + k) * TYPE1_STRIDE) + c,
			       make_uint4(unpack4(hbits), unpack4(hbits >> 4), unpack4(hbits >> 8), unpack4(hbits >> 12)));
		}
	}
	if (a.type1_packed) {
		for (int q = tid; q < 64 * 9; q += 32) {
			const int sl = q / 9, i = q - 9 * sl;
			const uint32_t *rec = stage + sl * LANE_STAGE_STRIDE;
			const uint32_t k = rec[9];
			if (k == LANE_NO_SLOT) continue;
			__stcs(a.type1_packed + (a.out_base + k) * TYPE1_WORDS + i, rec[i]);
		}
	}
}

__device__ __forceinline__ int lane_warp_max(int v)
{
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) {
		const int o = __shfl_xor_sync(FULL, v, d);
		v = o > v ? o : v;
	}
	return v;
}

/* the fused form: one kernel, the type-3 bits never leave shared memory */
template <bool TIE_HI>
__global__ void __launch_bounds__(32, 16)
k_decode_lane(DecodeArgs a, uint32_t *__restrict__ scratch)
{
	LaneSmem sm(TB_DYN_SMEM(), scratch);
	const Tables *__restrict__ tab = a.tab;
	lane_load_tables(sm, tab, true);
	const int tid = threadIdx.x;
	uint32_t *bcast = sm.lfb + (tid >> 5) * 16;
	const LaneLists L(a);
	const uint64_t nw = L.warps();
	for (uint64_t wu = blockIdx.x; wu < nw; wu += gridDim.x) {
		LaneUnit m;
		lane_prepare(a, tab, bcast, sm.leap, wu * 32 + tid, L, sm.t3col(0, tid), sm.t3col(1, tid), m);
		const int nmax = lane_warp_max(m.n[0] > m.n[1] ? m.n[0] : m.n[1]);
		if (nmax) viterbi_pair<TIE_HI>(sm.dec + tid, sm.t3col(0, tid), sm.t3col(1, tid), m.n[0], m.n[1], nmax);
		lane_finish(a, tab, sm.crc_tab, sm.t3col(0, tid), sm.t3col(1, tid), m, sm.t3);
		__syncwarp();
	}
}

/* ---- the split form ---- */

/* PREPARE as its own kernel: one warp per CTA, no trellis registers, so an SM holds 24+ of them and the dependent
 * loads of the cell state hide behind each other.  Shared memory: the scrambler's leap tables + one broadcast row. */
__host__ __device__ constexpr size_t lane_prepare_smem_words() { return 1024 + 16; }

/* one slot's share of PREPARE (k = LANE_NO_SLOT: nothing to do, but the warp collectives of lane_lfsr are still joined).
 * first: the slot is the unit's first one (only that one can be a two-block slot, which fills both columns). */
struct LaneSlotPrep { uint32_t k, code, flags, bbk, tm16; int kind, n0, n1; };
__device__ __forceinline__ LaneSlotPrep lane_prepare_slot(const DecodeArgs &a, const Tables *__restrict__ tab, uint32_t *bcast, const uint32_t *leap,
                                                          uint32_t k, bool first, uint32_t *col0, uint32_t *col1)
{
	constexpr int nt = LANE_NT;
	LaneSlotPrep r;
	r.k = k; r.code = 0; r.bbk = 0; r.kind = KIND_NONE; r.n0 = r.n1 = 0;
	Tm tm; tm.tn = tm.fn = tm.mn = 0;
	bool good_sb = false, unlock = false;
	uint32_t bw[16];
	if (k != LANE_NO_SLOT) {
		const SlotWs w = a.ws[k];
		load_slot_bits(a.slot_bits, k, bw);           /* asked for now, needed behind the cell state's chain of loads */
		r.kind = w.kind; good_sb = w.good_sb; unlock = w.unlock;
		const bool dep = cell_state(k, a.ws, a.last_good, a.blk_prev, a.carry, &tm, &r.code);
		if (a.skip_dependent && dep && !a.carry->seen_good) {     /* sharded decode: decoded later, with the real carry-in */
			r.k = LANE_NO_SLOT; r.kind = KIND_NONE; good_sb = unlock = false;
		}
	}
	r.flags = (uint32_t)r.kind | (unlock ? F_UNLOCK : 0) | ((r.kind == KIND_SB && good_sb) ? F_CRC_A : 0);
	if (r.kind == KIND_SB && tm_is_bnch(tm)) r.flags |= F_BNCH;
	r.tm16 = (tm.tn | (tm.fn << 3) | (tm.mn << 8)) & 0xffffu;
	uint32_t lf[LANE_T3_ROWS];
	lane_lfsr(r.code, r.kind != KIND_NONE, lf, bcast, tab, leap);
	if (r.kind != KIND_NONE) {
		uint32_t *col = first ? col0 : col1;
		if (r.kind == KIND_SB) {
			xor_region<252, 0, 30>(bw, lf);
			xor_region<282, 0, 216>(bw, lf);
			r.bbk = extract_bits(bw, 252, 30);
			gather_lane<1, PL_BLK2>(bw, col, nt); r.n0 = 144;
		} else if (r.kind == KIND_NDB_F) {
			xor_region<14, 0, 216>(bw, lf);
			xor_region<282, 216, 216>(bw, lf);
			xor_region<230, 0, 14>(bw, lf);
			r.bbk = extract_bits(bw, 230, 14);
			if (a.aach) r.bbk |= ((extract_bits(bw, 266, 16) ^ (lf[0] >> 14)) & 0xffffu) << 14;       /* the broadcast block's second part */
			gather_lane<5, PL_SCHF>(bw, col, nt); r.n0 = 288;
		} else if (first) {
			/* two-block slot: BLK1 on trellis X, BLK2 on trellis Y of the unit's thread (the unit has no second slot) */
			xor_region<14, 0, 216>(bw, lf);
			xor_region<282, 0, 216>(bw, lf);
			xor_region<230, 0, 14>(bw, lf);
			r.bbk = extract_bits(bw, 230, 14);
			if (a.aach) r.bbk |= ((extract_bits(bw, 266, 16) ^ (lf[0] >> 14)) & 0xffffu) << 14;
			gather_lane<1, PL_BLK1>(bw, col0, nt);
			gather_lane<1, PL_BLK2>(bw, col1, nt);
			r.n0 = r.n1 = 144;
		}
	}
	return r;
}

/* A warp takes ONE slot of each of 32 units (task = 2 * block + half): half the registers of the two-slot form, so an SM
 * holds twice the warps for the same work, and both the load chains and the ALU-pipe bound gather have more to overlap with.
 * The halves of a unit block are disjoint (column h, rows 28+h, 30+h, 32+h, 34+h); a two-block slot is the first half's
 * business alone, it fills both columns and both halves of the unit's state. */
#ifndef TB_PREP_MIN_CTAS
#define TB_PREP_MIN_CTAS 24
#endif
__global__ void __launch_bounds__(32, TB_PREP_MIN_CTAS)
k_lane_prepare(DecodeArgs a, uint32_t *__restrict__ units)
{
	uint32_t *leap = reinterpret_cast<uint32_t *>(TB_DYN_SMEM());
	uint32_t *bcast = leap + 1024;
	const Tables *__restrict__ tab = a.tab;
	const int tid = threadIdx.x;
	constexpr int nt = LANE_NT;
	for (int i = tid; i < 1024; i += 32) leap[i] = (&tab->lfsr_leap[0][0])[i];
	__syncwarp();
	const LaneLists L(a);
	const uint64_t ntask = 2 * L.warps();
	for (uint64_t task = blockIdx.x; task < ntask; task += gridDim.x) {
		const uint64_t wu = task >> 1, u = wu * 32 + tid;
		const int h = (int)(task & 1);
		uint32_t *blk = units + wu * LANE_UNIT_WORDS + tid;
		uint32_t k[2];
		L.slots_of(u, k);
		const bool two_block_unit = u >= L.uF && u < L.uF + L.u2;
		const bool mine = !(h == 1 && two_block_unit);
		const LaneSlotPrep r = lane_prepare_slot(a, tab, bcast, leap, mine ? k[h] : LANE_NO_SLOT, h == 0, blk, blk + LANE_T3_ROWS * nt);
		if (mine) {
			blk[(LANE_META_ROW + h) * nt] = r.tm16 | (r.flags << 16) | ((uint32_t)r.kind << 24) | ((uint32_t)lane_ncode(r.n0) << 28);
			blk[(LANE_META_ROW + 2 + h) * nt] = r.k;
			blk[(LANE_META_ROW + 4 + h) * nt] = r.code;
			blk[(LANE_META_ROW + 6 + h) * nt] = r.bbk;
			if (h == 0 && two_block_unit) {       /* the second trellis carries BLK2 of the same slot; no second slot */
				blk[(LANE_META_ROW + 1) * nt] = ((uint32_t)KIND_NONE << 24) | ((uint32_t)lane_ncode(r.n1) << 28);
				blk[(LANE_META_ROW + 3) * nt] = LANE_NO_SLOT;
				blk[(LANE_META_ROW + 5) * nt] = 0;
				blk[(LANE_META_ROW + 7) * nt] = 0;
			}
		}
		__syncwarp();
	}
}

/* TRELLIS: the ACS loop and the trace back over the prepared blocks.  One warp per CTA, 16 CTAs per SM (registers); the
 * block of the CTA's next unit arrives by bulk copy while the current one is decoded.  FINISH = true runs the finish part
 * right here out of shared memory, otherwise the decoded bits go back into the block (rows 0..8 and 14..22) for
 * k_lane_finish. */
template <bool FINISH>
__host__ __device__ constexpr size_t lane_trellis_smem_words() { return 2 * LANE_UNIT_WORDS + 4 + (FINISH ? 256 + 16 : 0); }

template <bool TIE_HI, bool FINISH>
__global__ void __launch_bounds__(32, 16)
k_lane_trellis(DecodeArgs a, uint32_t *__restrict__ scratch, uint32_t *__restrict__ units)
{
	uint32_t *buf = reinterpret_cast<uint32_t *>(TB_DYN_SMEM());
	uint64_t *bars = reinterpret_cast<uint64_t *>(buf + 2 * LANE_UNIT_WORDS);
	uint32_t *crc_tab = buf + 2 * LANE_UNIT_WORDS + 4;
	const Tables *__restrict__ tab = a.tab;
	const int tid = threadIdx.x;
	constexpr int nt = LANE_NT;
	constexpr unsigned BYTES = LANE_UNIT_WORDS * sizeof(uint32_t);
	uint4 *dec = reinterpret_cast<uint4 *>(scratch + (size_t)blockIdx.x * lane_scratch_words_per_cta(nt)) + tid;
	if (FINISH)
		for (int i = tid; i < 256 + 16; i += 32) crc_tab[i] = tab->crc_tab_r[i];
	if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
	__syncwarp();
	const LaneLists L(a);
	const uint64_t nw = L.warps();
	if (blockIdx.x < nw && tid == 0) {
		mbar_expect_tx(&bars[0], BYTES);
		bulk_g2s(buf, units + (size_t)blockIdx.x * LANE_UNIT_WORDS, BYTES, &bars[0]);
	}
	unsigned phase = 0, b = 0;
	for (uint64_t wu = blockIdx.x; wu < nw; wu += gridDim.x, b ^= 1u) {
		const uint64_t nxt = wu + gridDim.x;
		if (nxt < nw && tid == 0) {           /* the other buffer is free: its unit was finished before the __syncwarp below */
			fence_async_shared();
			mbar_expect_tx(&bars[b ^ 1u], BYTES);
			bulk_g2s(buf + (b ^ 1u) * LANE_UNIT_WORDS, units + nxt * LANE_UNIT_WORDS, BYTES, &bars[b ^ 1u]);
		}
		mbar_wait(&bars[b], (phase >> b) & 1u);
		phase ^= 1u << b;
		uint32_t *cx = buf + b * LANE_UNIT_WORDS + tid, *cy = cx + LANE_T3_ROWS * nt;
		if (FINISH) {
			LaneUnit m;
			lane_unit_load(cx, m);
			const int nmax = lane_warp_max(m.n[0] > m.n[1] ? m.n[0] : m.n[1]);
			if (nmax) viterbi_pair<TIE_HI>(dec, cx, cy, m.n[0], m.n[1], nmax);
			lane_finish(a, tab, crc_tab, cx, cy, m, buf + b * LANE_UNIT_WORDS);
		} else {
			const int n0 = lane_nsteps(cx[LANE_META_ROW * nt]), n1 = lane_nsteps(cx[(LANE_META_ROW + 1) * nt]);
			const int nmax = lane_warp_max(n0 > n1 ? n0 : n1);
			if (nmax) {
				viterbi_pair<TIE_HI>(dec, cx, cy, n0, n1, nmax);
				uint32_t *blk = units + wu * LANE_UNIT_WORDS + tid;
				const int rows = (nmax + 31) >> 5;
				for (int i = 0; i < rows; ++i) {
					if (n0) blk[i * nt] = cx[i * nt];
					if (n1) blk[(LANE_T3_ROWS + i) * nt] = cy[i * nt];
				}
			}
		}
		__syncwarp();
	}
}

/* FINISH as its own kernel: reads the decoded rows and the unit's part of the block */
__host__ __device__ constexpr size_t lane_finish_smem_words() { return 256 + 16 + LANE_STAGE_WORDS; }
__global__ void __launch_bounds__(32, 24)
k_lane_finish(DecodeArgs a, const uint32_t *__restrict__ units)
{
	uint32_t *crc_tab = reinterpret_cast<uint32_t *>(TB_DYN_SMEM());
	uint32_t *stage = crc_tab + 256 + 16;
	const Tables *__restrict__ tab = a.tab;
	const int tid = threadIdx.x;
	for (int i = tid; i < 256 + 16; i += 32) crc_tab[i] = tab->crc_tab_r[i];
	__syncwarp();
	const LaneLists L(a);
	const uint64_t nw = L.warps();
	for (uint64_t wu = blockIdx.x; wu < nw; wu += gridDim.x) {
		const uint32_t *blk = units + wu * LANE_UNIT_WORDS + tid;
		LaneUnit m;
		lane_unit_load(blk, m);
		lane_finish(a, tab, crc_tab, blk, blk + LANE_T3_ROWS * LANE_NT, m, stage);
		__syncwarp();
	}
}

}  // namespace tb
