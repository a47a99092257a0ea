/*
 * tetra_classify_tile.cuh - pass 1 (training-sequence search + slot packing), one CTA per TILE of slots.
 *
 * (An earlier thread-per-slot form staged 528 bytes per slot, could keep only six warps on an SM and was
 * latency bound at 65 % of the HBM roofline; profiles/r01c.)  Slots of a LOCKED stream are back to back, so
 * a CTA takes 64 consecutive slots = one contiguous 32 640-byte piece of the stream:
 *
 *   copy   one cp.async.bulk (TMA bulk copy) brings the 16-byte aligned superset of the tile into shared
 *          memory; two tile buffers per CTA, the copy of the CTA's next tile is in flight while the
 *          current one is searched (3 CTAs per SM -> ~96 KB in flight per SM, every byte read once);
 *   pack   all threads turn the 0/1 bytes into a dense bit string (16 bytes -> 16 bits with IDP.4A);
 *   emit   the slot's 510 bits, re-aligned to a 64-byte record (funnel shifts of the bit string), written
 *          coalesced for the decode pass and kept transposed in shared memory for the search;
 *   search position-parallel: warp j tests the 32 positions of word j of a slot against the three downlink
 *          sequences, two pattern bits per LOP3; offsets 0..255 are covered (the expected offsets 214 /
 *          244 are inside; a hit there is the first hit whatever the search window).  Words 0..5 almost
 *          never hold a match, so six of the eight warps leave after 12 pattern bits.  Hits (about one
 *          per slot) go through the reference's pre-filter blind spot rule and an atomicMin per slot
 *          keeps the first;
 *   decide one thread per slot: search window, kind, lock loss; slots without a hit in the tested range go
 *          through the exact warp-cooperative search over their whole window (rare).
 *
 * Semantics: tetra_find_train_seq (tetra_burst.c:269-339) + LOCKED arm of tetra_burst_sync_in
 * (tetra_burst_sync.c:107-143).
 */
#pragma once
#include "tetra_kernels.cuh"
#include "tetra_async.cuh"

namespace tb {

/* bits_in_buf when slot k of the launch is processed (same value as slot_window()) */
struct WinGeom {
	CallGeom cg;
	uint64_t cmin, a0;
};

__device__ __forceinline__ unsigned slot_window32(const WinGeom &g, uint32_t k)
{
	/* 64-bit: a shard of the sharded path can hold more than 2^32 / 510 slots in one launch */
	const uint64_t ak = g.a0 + 510ull * k;
	return (unsigned)(bits_at_call(g.cg, call_for(g.cg, ak + 510u, g.cmin + k)) - ak);
}

constexpr int CT_THREADS = 256;
constexpr int CT_RAW = ((15 + 510 * 64 + 15) / 16) * 16;       /* bytes of one raw tile buffer (32 672) */
constexpr int CT_BITW = ((CT_RAW / 32 + 8 + 3) / 4) * 4;       /* words of the packed tile (+ lead of a packed copy + read-ahead), 16-byte multiple */

/* per input format: slots per tile and the size of one copy buffer.
 *   BYTES   64 slots = 32 640 stream bytes, packed to bits by the CTA
 *   PACKED  64 slots = 4 080 bytes that ARE the bit string: copied straight into it, no pack phase
 *   F32SYM  32 slots = 8 160 symbols = 32 640 bytes, sliced to 2 bits per symbol by the CTA */
template <int FMT> struct TileFmt;
template <> struct TileFmt<IN_BYTES>  { static constexpr int SLOTS = 64, BUF = CT_RAW; };
template <> struct TileFmt<IN_PACKED> { static constexpr int SLOTS = 64, BUF = CT_BITW * 4; };
template <> struct TileFmt<IN_F32SYM> { static constexpr int SLOTS = 32, BUF = CT_RAW; };
constexpr int CT_AROW = 64 + 1;                                /* row pitch of the aligned slot words (odd: no bank conflicts on the transposed write) */
template <int FMT> constexpr size_t ct_smem()
{
	return (size_t)2 * TileFmt<FMT>::BUF + (FMT == IN_PACKED ? 0 : (size_t)CT_BITW * 4) + (size_t)16 * CT_AROW * 4 + 64 * 4 + 2 * 8;
}
constexpr int CT_SLOTS = 64;                                   /* slots per tile of the byte format (launch geometry helper) */
constexpr size_t CT_SMEM = ct_smem<IN_BYTES>();

/* two pattern bits (B, B+1) of the three downlink sequences against the 32 positions of word x0 */
template <int B>
__device__ __forceinline__ void match_pair(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t &My, uint32_t &Mn, uint32_t &Mp)
{
	uint32_t s0, s1;
	if constexpr (B < 32) s0 = __funnelshift_r(x0, x1, B); else s0 = __funnelshift_r(x1, x2, B - 32);
	if constexpr (B + 1 < 32) s1 = __funnelshift_r(x0, x1, B + 1); else s1 = __funnelshift_r(x1, x2, B + 1 - 32);
	My &= (((SEQ_Y >> B) & 1) ? s0 : ~s0) & (((SEQ_Y >> (B + 1)) & 1) ? s1 : ~s1);
	if constexpr (B < 22) {
		Mn &= (((SEQ_N >> B) & 1) ? s0 : ~s0) & (((SEQ_N >> (B + 1)) & 1) ? s1 : ~s1);
		Mp &= (((SEQ_P >> B) & 1) ? s0 : ~s0) & (((SEQ_P >> (B + 1)) & 1) ? s1 : ~s1);
	}
}
#define TB_MATCH_PAIR(b) match_pair<(b)>(x0, x1, x2, My, Mn, Mp)

struct TileGeom {
	const uint8_t *src;      /* 16-byte aligned start of the copy */
	uint32_t lead;           /* stream BITS between the start of the copied bit string and the first slot of the tile */
	uint32_t bytes;          /* copy size, multiple of 16 */
	uint32_t k0, ns;         /* first slot, slots in the tile */
	bool staged;             /* the aligned superset lies inside the caller's buffer: TMA copy */
};

template <int FMT>
__device__ __forceinline__ TileGeom tile_geom(const RxGeom &g, uint32_t tile, uintptr_t buf_lo, uintptr_t buf_hi)
{
	constexpr uint32_t S = TileFmt<FMT>::SLOTS;
	TileGeom t;
	t.k0 = tile * S;
	t.ns = g.n_slots - t.k0 < S ? g.n_slots - t.k0 : S;
	const uint64_t bit0 = (g.a0 - g.base_bit) + 510ull * t.k0;          /* first bit of the tile, relative to g.bits */
	uintptr_t p;                                                         /* address of the unit that holds it */
	uint32_t sub;                                                        /* bit inside that unit */
	uint64_t span;                                                       /* bytes from p to the end of the tile's last bit */
	if (FMT == IN_BYTES)       { p = reinterpret_cast<uintptr_t>(g.bits) + bit0;            sub = 0;                   span = 510ull * t.ns; }
	else if (FMT == IN_PACKED) { p = reinterpret_cast<uintptr_t>(g.bits) + (bit0 >> 3);     sub = (uint32_t)(bit0 & 7); span = (sub + 510ull * t.ns + 7) >> 3; }
	else                       { p = reinterpret_cast<uintptr_t>(g.bits) + 4 * (bit0 >> 1); sub = (uint32_t)(bit0 & 1); span = 4 * ((sub + 510ull * t.ns + 1) >> 1); }
	const uintptr_t a = p & ~(uintptr_t)15;
	t.src = reinterpret_cast<const uint8_t *>(a);
	const uint32_t lead_bytes = (uint32_t)(p - a);
	t.lead = FMT == IN_BYTES ? lead_bytes : FMT == IN_PACKED ? 8 * lead_bytes + sub : lead_bytes / 2 + sub;
	t.bytes = (uint32_t)((lead_bytes + span + 15u) & ~15ull);
	t.staged = a >= buf_lo && a + t.bytes <= buf_hi;
	return t;
}

#ifndef TB_CT_MIN_CTAS
#define TB_CT_MIN_CTAS 3      /* resident CTAs per SM the register budget is cut for (A/B builds: -DTB_CT_MIN_CTAS=4 -> 64 registers) */
#endif
template <int FMT>
__global__ void __launch_bounds__(CT_THREADS, TB_CT_MIN_CTAS)
k_classify_tile(RxGeom g, WinGeom wg, const Tables *__restrict__ tab, SlotWs *__restrict__ ws,
                uint32_t *__restrict__ slot_bits, uint32_t *__restrict__ sb_list, uint32_t *__restrict__ sb_count)
{
	constexpr int S = TileFmt<FMT>::SLOTS, BUF = TileFmt<FMT>::BUF;
	uint8_t *smem = TB_DYN_SMEM();
	uint8_t *raw = smem;                                                 /* [2][BUF] copy buffers */
	uint32_t *bitw_own = reinterpret_cast<uint32_t *>(smem + 2 * BUF);    /* [CT_BITW] the tile as a dense bit string (the copy buffer itself when PACKED) */
	uint32_t *al = bitw_own + (FMT == IN_PACKED ? 0 : CT_BITW);          /* [16][CT_AROW] word j of slot s, slot-aligned */
	uint32_t *first = al + 16 * CT_AROW;                                 /* [64] (offset << 3 | type), ~0 = none */
	uint64_t *bars = reinterpret_cast<uint64_t *>(first + 64);           /* [2] */
	const unsigned tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;

	/* bytes of the caller's buffer: g.n_bytes counts stream bits */
	const uint64_t buf_bytes = FMT == IN_BYTES ? g.n_bytes : FMT == IN_PACKED ? (g.n_bytes + 7) >> 3 : 4 * ((g.n_bytes + 1) >> 1);
	const uintptr_t buf_lo = (reinterpret_cast<uintptr_t>(g.bits) + 15) & ~(uintptr_t)15;
	const uintptr_t buf_hi = (reinterpret_cast<uintptr_t>(g.bits) + buf_bytes) & ~(uintptr_t)15;
	const uint8_t *end = g.bits + buf_bytes;
	const uint32_t ntiles = (g.n_slots + S - 1) / S;

	if (tid == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
	}
	__syncthreads();
	auto issue = [&](uint32_t tile, int b) {            /* thread 0 only */
		if (tile >= ntiles) return;
		const TileGeom t = tile_geom<FMT>(g, tile, buf_lo, buf_hi);
		if (!t.staged) return;
		mbar_expect_tx(&bars[b], t.bytes);
#ifdef TB_CT_EVICT_FIRST
		bulk_g2s_stream(raw + (size_t)b * BUF, t.src, t.bytes, &bars[b]);
#else
		bulk_g2s(raw + (size_t)b * BUF, t.src, t.bytes, &bars[b]);
#endif
	};
	if (tid == 0) issue(blockIdx.x, 0);
	__syncthreads();

	unsigned phase_bits = 0;                     /* bit b = parity the next wait on buffer b expects */
	int b = 0;
	for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, b ^= 1) {
		const TileGeom t = tile_geom<FMT>(g, tile, buf_lo, buf_hi);
		uint8_t *rb = raw + (size_t)b * BUF;
		uint32_t *bitw = FMT == IN_PACKED ? reinterpret_cast<uint32_t *>(rb) : bitw_own;
		if (tid == 0) issue(tile + gridDim.x, b ^ 1);         /* prefetch this CTA's next tile */
		if (t.staged) {
			mbar_wait(&bars[b], (phase_bits >> b) & 1u);
			phase_bits ^= 1u << b;
		} else {
			/* tile touches the ends of the caller's buffer (or the buffer is not 16-byte aligned there) */
			for (uint32_t i = tid; i < t.bytes; i += CT_THREADS) {
				const uint8_t *a = t.src + i;
				rb[i] = (a >= g.bits && a < end) ? *a : 0;
			}
			__syncthreads();
		}
		/* ---- pack: the copied tile -> dense bit string, two 16-byte quads per thread and round */
		if (FMT == IN_BYTES) {                       /* 16 bytes -> 16 bits */
			const uint32_t nq = t.bytes >> 4;
			const uint4 *q4 = reinterpret_cast<const uint4 *>(rb);
			uint16_t *bh = reinterpret_cast<uint16_t *>(bitw);
			for (uint32_t q = tid; q < nq + 8; q += 2 * CT_THREADS) {
				const uint32_t q1 = q + CT_THREADS;
				const uint4 v0 = q < nq ? q4[q] : make_uint4(0, 0, 0, 0);
				const uint4 v1 = q1 < nq ? q4[q1] : make_uint4(0, 0, 0, 0);
				bh[q] = (uint16_t)pack16_dp4a(v0);
				if (q1 < nq + 8) bh[q1] = (uint16_t)pack16_dp4a(v1);
			}
		} else if (FMT == IN_F32SYM) {               /* 4 symbols -> 8 bits (float_to_bits.c:33-72) */
			const uint32_t nq = t.bytes >> 4;
			const float4 *q4 = reinterpret_cast<const float4 *>(rb);
			uint8_t *bb = reinterpret_cast<uint8_t *>(bitw);
			for (uint32_t q = tid; q < nq + 16; q += CT_THREADS) {
				uint32_t v = 0;
				if (q < nq) {
					const float4 f = q4[q];
					v = slice_symbol(f.x) | (slice_symbol(f.y) << 2) | (slice_symbol(f.z) << 4) | (slice_symbol(f.w) << 6);
				}
				bb[q] = (uint8_t)v;
			}
		}
		if (tid < 64) first[tid] = 0xffffffffu;
		__syncthreads();
		/* ---- align + emit: word j of slot s (bit i of the slot -> word i>>5 bit i&31); thread = (j, S/16 slots).
		 * 16 slots further on the stream is 255 words further on and equally aligned. */
		{
			const uint32_t j = tid & 15, s0 = tid >> 4;
			const uint32_t dt = t.lead + 510u * s0, sh = dt & 31;
			const uint32_t *src = bitw + (dt >> 5) + j;
			uint32_t *dst = slot_bits + (size_t)t.k0 * 16 + tid;
#pragma unroll
			for (int r = 0; r < S / 16; ++r) {
				const uint32_t s = s0 + 16 * r;
				uint32_t v = __funnelshift_r(src[255 * r], src[255 * r + 1], sh);
				if (j == 15) v &= 0x3fffffffu;
				al[j * CT_AROW + s] = v;
				if (s < t.ns) dst[256 * r] = v;
			}
		}
		__syncthreads();
		/* ---- search offsets 0..255 of every slot: warp j tests the 32 positions of word j of the slots
		 * lane and lane + 32.  The expected offsets 214 (SYNC) and 244 (normal) sit in words 6 and 7, so
		 * the other warps leave after 12 pattern bits almost always. */
#pragma unroll 1
		for (int r = 0; r < S / 32; ++r) {
			const uint32_t s = lane + 32 * r, j = wib;
			const uint32_t x0 = al[j * CT_AROW + s], x1 = al[(j + 1) * CT_AROW + s];
			uint32_t x2 = 0;
			uint32_t My = 0xffffffffu, Mn = 0xffffffffu, Mp = 0xffffffffu;
			TB_MATCH_PAIR(0); TB_MATCH_PAIR(2); TB_MATCH_PAIR(4); TB_MATCH_PAIR(6); TB_MATCH_PAIR(8); TB_MATCH_PAIR(10);
			if (s >= t.ns) My = Mn = Mp = 0;
			if (__any_sync(FULL, (My | Mn | Mp) != 0)) {
				TB_MATCH_PAIR(12); TB_MATCH_PAIR(14); TB_MATCH_PAIR(16); TB_MATCH_PAIR(18); TB_MATCH_PAIR(20);
				if (My) {            /* the 16 remaining bits of the SYNC sequence, only where its prefix matched */
					x2 = al[(j + 2) * CT_AROW + s];
					TB_MATCH_PAIR(22); TB_MATCH_PAIR(24); TB_MATCH_PAIR(26); TB_MATCH_PAIR(28);
					TB_MATCH_PAIR(30); TB_MATCH_PAIR(32); TB_MATCH_PAIR(34); TB_MATCH_PAIR(36);
				}
				uint32_t Mall = My | Mn | Mp;
				while (Mall) {
					const int i0 = __ffs((int)Mall) - 1;
					Mall &= Mall - 1;
					const uint32_t rel = 32u * j + (uint32_t)i0;
					const int seq = ((My >> i0) & 1) ? 0 : ((Mn >> i0) & 1) ? 1 : 2;
					bool ok = true;
					if (rel <= 20) {      /* pre-filter blind spot, offsets 0..20 (tetra_burst.c:288-294); only word 0 */
						const uint32_t prev = rel > 0 ? (x0 >> (rel - 1)) & 1u : 0u;
						ok = (tab->blind_ok[seq][prev] >> rel) & 1u;
					}
					if (ok) {
						const uint32_t ty = seq == 0 ? (uint32_t)TS_SYNC : seq == 1 ? (uint32_t)TS_NORM_1 : (uint32_t)TS_NORM_2;
						atomicMin(&first[s], (rel << 3) | ty);
					}
				}
			}
		}
		__syncthreads();
		/* ---- decide: one thread per slot */
		if (tid < S) {
			const bool have = tid < t.ns;
			const uint32_t k = t.k0 + tid;
			const uint64_t ak = g.a0 + 510ull * k;
			const unsigned W = have ? slot_window32(wg, k) : 0;
			const uint32_t f = first[tid];
			int rc = -1;
			unsigned off = 0;
			if (have && f != 0xffffffffu) { rc = (int)(f & 7u); off = f >> 3; }
			/* exact warp-cooperative search for the slots the fast path could not settle */
			unsigned todo = __ballot_sync(FULL, have && rc < 0);
			while (todo) {
				const int src = __ffs((int)todo) - 1;
				todo &= todo - 1;
				const uint64_t off_b = __shfl_sync(FULL, (uint64_t)(ak - g.base_bit), src);
				const unsigned Ws = __shfl_sync(FULL, W, src);
				unsigned o2 = 0;
				const uint32_t mask = (1u << TS_SYNC) | (1u << TS_NORM_1) | (1u << TS_NORM_2);
				const int r2 = find_train_seq_warp_fmt(g.bits, FMT, off_b, g.n_bytes, Ws, mask, tab, &o2, nullptr);
				if ((int)lane == src) { rc = r2; off = o2; }
			}
			int kind = KIND_NONE;
			bool unlock = false;
			if (have) {
				if (rc == TS_SYNC) { if (off == 214) kind = KIND_SB; else unlock = true; }
				else if (rc == TS_NORM_1) { if (off == 244) kind = KIND_NDB_F; }
				else if (rc == TS_NORM_2) { if (off == 244) kind = KIND_NDB_2; }
				else unlock = true;
			}
			/* SYNC bursts are listed so that the SB1 pass only touches them (one atomic per warp) */
			{
				const unsigned m = __ballot_sync(FULL, kind == KIND_SB);
				if (m) {
					uint32_t base = 0;
					if (lane == (unsigned)(__ffs((int)m) - 1)) base = atomicAdd(sb_count, (uint32_t)__popc(m));
					base = __shfl_sync(FULL, base, __ffs((int)m) - 1);
					if (kind == KIND_SB) sb_list[base + __popc(m & ((1u << lane) - 1))] = k;
				}
			}
			if (have) {
				SlotWs w;
				w.sb1_t1[0] = 0; w.sb1_t1[1] = 0; w.sb_code = 0;
				w.find_off = (uint16_t)off; w.window = (uint16_t)W;
				w.find_rc = (int8_t)rc; w.good_sb = 0; w.kind = (uint8_t)kind; w.unlock = unlock;
				w.tn = w.fn = w.mn = w.cc = 0; w.mcc = w.mnc = 0; w.sb1_crc = 0;
				ws[k] = w;
			}
		}
		__syncthreads();
	}
}
#undef TB_MATCH_PAIR

}  // namespace tb
