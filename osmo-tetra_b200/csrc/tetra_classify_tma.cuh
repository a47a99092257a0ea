/*
 * tetra_classify_tma.cuh - pass 1 (training-sequence search + slot packing), one THREAD per slot.
 *
 * The warp-per-slot search of tetra_kernels.cuh is ALU-pipe bound (600 warp instructions per slot,
 * 10 % of the HBM roofline): half of its lanes have no positions to test and every lane repeats the
 * realign / select work.  Here each thread owns a slot:
 *   - every lane issues one 528-byte cp.async.bulk (TMA bulk copy, 16-byte aligned superset of its
 *     510-byte slot) into its own PADDED shared-memory row (33 x 16 B, so the 32 rows start in
 *     different bank groups) completing on one mbarrier per warp; two row sets per warp are in flight
 *     so the copy of the next 32 slots overlaps the search of the current ones;
 *   - bytes are packed to bits with IDP.4A (dp4a with weights 1,2,4,8 / 16,32,64,128);
 *   - the search tests the 256 first positions bit-parallel in registers (the expected offsets are 214
 *     and 244; a hit below 256 is the first hit whatever the window), with the reference's pre-filter
 *     blind spot applied to positions 0..20; slots with no hit there, with windows the fast path cannot
 *     judge, or too close to the buffer ends go through the exact warp-cooperative search.
 * Semantics: tetra_find_train_seq (tetra_burst.c:269-339) + LOCKED arm of tetra_burst_sync_in
 * (tetra_burst_sync.c:107-143).
 */
#pragma once
#include "tetra_kernels.cuh"

namespace tb {

constexpr int CLS_ROW = 528;                 /* bytes staged per slot */
constexpr int CLS_WARPS = 6;                 /* warps per CTA (6 x 2 x 16.5 KB of rows = 203 KB: one CTA per SM) */
constexpr int CLS_STAGES = 2;
constexpr size_t CLS_SMEM = (size_t)CLS_WARPS * CLS_STAGES * 32 * CLS_ROW + CLS_WARPS * CLS_STAGES * 8 + 16;

/* ---- async-copy plumbing (PTX on the GPU, plain copies under the CPU emulator) ---- */
#ifdef TB_SIMT_EMULATION
__device__ __forceinline__ void mbar_init(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_arrive(uint64_t *) {}
__device__ __forceinline__ void mbar_wait(uint64_t *, unsigned) { __syncwarp(); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *) { memcpy(dst, src, bytes); }
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
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

/* bits_in_buf when slot k is processed, in 32-bit arithmetic relative to the launch
 * (same value as slot_window(); rel = first slot's offset inside its read() chunk) */
struct WinGeom {
	uint32_t chunk, rel0;        /* rel0 = a0 mod chunk */
	uint64_t c00;                /* a0 div chunk */
	uint64_t cmin, n_end, a0;
};

__device__ __forceinline__ unsigned slot_window32(const WinGeom &g, uint32_t k)
{
	/* 64-bit: a shard of the sharded path can hold more than 2^32 / 510 slots in one launch */
	const uint64_t num = (uint64_t)g.rel0 + 510ull * k + 510u + g.chunk - 1;
	const uint64_t q = g.chunk == 64 ? num >> 6 : (num <= 0xffffffffull ? (uint64_t)((uint32_t)num / g.chunk) : num / g.chunk);
	uint64_t c = g.c00 + q;
	const uint64_t lo = g.cmin + k;
	if (c < lo) c = lo;
	uint64_t t = c * g.chunk;
	if (t > g.n_end) t = g.n_end;
	return (unsigned)(t - (g.a0 + 510ull * k));
}

/* match masks of the three downlink sequences at the 32 positions of word x0 (x1, x2 follow it) */
__device__ __forceinline__ void match32(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t &My, uint32_t &Mn, uint32_t &Mp)
{
	My = Mn = Mp = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < 22; b += 2) {
		const uint32_t s0 = __funnelshift_r(x0, x1, b), s1 = __funnelshift_r(x0, x1, b + 1);
		My &= (((SEQ_Y >> b) & 1) ? s0 : ~s0) & (((SEQ_Y >> (b + 1)) & 1) ? s1 : ~s1);
		Mn &= (((SEQ_N >> b) & 1) ? s0 : ~s0) & (((SEQ_N >> (b + 1)) & 1) ? s1 : ~s1);
		Mp &= (((SEQ_P >> b) & 1) ? s0 : ~s0) & (((SEQ_P >> (b + 1)) & 1) ? s1 : ~s1);
	}
	if (My) {            /* the 16 remaining bits of the SYNC sequence, only where its prefix matched */
#pragma unroll
		for (int b = 22; b < 38; b += 2) {
			const uint32_t s0 = b < 32 ? __funnelshift_r(x0, x1, b) : __funnelshift_r(x1, x2, b - 32);
			const uint32_t s1 = b + 1 < 32 ? __funnelshift_r(x0, x1, b + 1) : __funnelshift_r(x1, x2, b + 1 - 32);
			My &= (((SEQ_Y >> b) & 1) ? s0 : ~s0) & (((SEQ_Y >> (b + 1)) & 1) ? s1 : ~s1);
		}
	}
}

__global__ void __launch_bounds__(CLS_WARPS * 32)
k_classify_tma(RxGeom g, WinGeom wg, const Tables *__restrict__ tab, SlotWs *__restrict__ ws,
               uint32_t *__restrict__ slot_bits, uint32_t *__restrict__ sb_list, uint32_t *__restrict__ sb_count)
{
	uint8_t *smem = TB_DYN_SMEM();
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint8_t *rows = smem + (size_t)wib * CLS_STAGES * 32 * CLS_ROW;
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)CLS_WARPS * CLS_STAGES * 32 * CLS_ROW) + wib * CLS_STAGES;
	if (lane == 0) {
		for (int s = 0; s < CLS_STAGES; ++s) mbar_init(&bars[s], 32);
	}
	__syncwarp();

	const uintptr_t buf_lo = (reinterpret_cast<uintptr_t>(g.bits) + 15) & ~(uintptr_t)15;
	const uintptr_t buf_hi = (reinterpret_cast<uintptr_t>(g.bits) + g.n_bytes) & ~(uintptr_t)15;
	const uint8_t *end = g.bits + g.n_bytes;
	const uint32_t ngroups = (g.n_slots + 31) / 32;
	const uint32_t nwarps = gridDim.x * CLS_WARPS;
	const uint32_t w0 = blockIdx.x * CLS_WARPS + wib;
	const uint32_t blind0_y = tab->blind_ok[0][0], blind1_y = tab->blind_ok[0][1];
	const uint32_t blind0_n = tab->blind_ok[1][0], blind1_n = tab->blind_ok[1][1];
	const uint32_t blind0_p = tab->blind_ok[2][0], blind1_p = tab->blind_ok[2][1];

	/* issue the copy of group `grp` into stage `st`; returns whether this lane's slot is staged */
	auto issue = [&](uint32_t grp, int st) -> bool {
		const uint32_t k = grp * 32 + lane;
		bool staged = false;
		if (grp < ngroups && k < g.n_slots) {
			const uintptr_t p = reinterpret_cast<uintptr_t>(g.bits) + (size_t)((g.a0 - g.base_bit) + 510ull * k);
			const uintptr_t a = p & ~(uintptr_t)15;
			staged = a >= buf_lo && a + CLS_ROW <= buf_hi;
			if (staged) {
				mbar_expect_tx(&bars[st], CLS_ROW);
				bulk_g2s(rows + ((size_t)st * 32 + lane) * CLS_ROW, reinterpret_cast<const void *>(a), CLS_ROW, &bars[st]);
			}
		}
		if (!staged) mbar_arrive(&bars[st]);
		return staged;
	};

	unsigned phase_bits = 0;                     /* bit s = parity the next wait on stage s expects */
	bool staged_next = issue(w0, 0);
	int st = 0;
	for (uint32_t grp = w0; grp < ngroups; grp += nwarps) {
		const bool staged = staged_next;
		staged_next = issue(grp + nwarps, st ^ 1);          /* prefetch the warp's next group */
		mbar_wait(&bars[st], (phase_bits >> st) & 1u);
		phase_bits ^= 1u << st;

		const uint32_t k = grp * 32 + lane;
		const bool have = k < g.n_slots;
		const uint64_t ak = g.a0 + 510ull * k;
		const uint8_t *p = g.bits + (size_t)(ak - g.base_bit);
		const unsigned W = have ? slot_window32(wg, k) : 0;
		uint32_t X[17];
		int rc = -1;
		unsigned off = 0;
		bool slow = have;
		if (have && staged) {
			const unsigned d = (unsigned)(reinterpret_cast<uintptr_t>(p) & 15);
			const uint4 *row = reinterpret_cast<const uint4 *>(rows + ((size_t)st * 32 + lane) * CLS_ROW);
			uint32_t A[17];
#pragma unroll
			for (int j = 0; j < 16; ++j)
				A[j] = pack16_dp4a(row[2 * j]) | (pack16_dp4a(row[2 * j + 1]) << 16);
			A[16] = pack16_dp4a(row[32]);
#pragma unroll
			for (int j = 0; j < 16; ++j) X[j] = __funnelshift_r(A[j], A[j + 1], d);
			X[16] = A[16] >> d;
			/* first hit among positions 0..255 */
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				if (rc < 0) {
					uint32_t My, Mn, Mp;
					match32(X[j], X[j + 1], X[j + 2], My, Mn, Mp);
					if (j == 0) {          /* pre-filter blind spot, offsets 0..20 (tetra_burst.c:288-294) */
						const uint32_t prev = X[0] << 1, keep = 0xffe00000u;
						My &= keep | (prev & blind1_y) | (~prev & blind0_y);
						Mn &= keep | (prev & blind1_n) | (~prev & blind0_n);
						Mp &= keep | (prev & blind1_p) | (~prev & blind0_p);
					}
					const uint32_t Mall = My | Mn | Mp;
					if (Mall) {
						const int i0 = __ffs((int)Mall) - 1;
						off = 32 * j + i0;
						rc = ((My >> i0) & 1) ? TS_SYNC : ((Mn >> i0) & 1) ? TS_NORM_1 : TS_NORM_2;
					}
				}
			}
			slow = rc < 0;
			X[15] &= 0x3fffffffu;
		}
		/* exact warp-cooperative search for the slots the fast path could not settle */
		unsigned todo = __ballot_sync(FULL, slow);
		while (todo) {
			const int src = __ffs((int)todo) - 1;
			todo &= todo - 1;
			const uint64_t off_b = __shfl_sync(FULL, (uint64_t)(ak - g.base_bit), src);
			const unsigned Ws = __shfl_sync(FULL, W, src);
			unsigned o2 = 0;
			uint32_t xw = 0;
			const uint32_t mask = (1u << TS_SYNC) | (1u << TS_NORM_1) | (1u << TS_NORM_2);
			const int r2 = find_train_seq_warp(g.bits + off_b, end, Ws, mask, tab, &o2, &xw);
			/* hand the packed slot words (lane j holds word j) to the owning lane */
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const uint32_t v = __shfl_sync(FULL, xw, j);
				if ((int)lane == src) X[j] = v;
			}
			if ((int)lane == src) { rc = r2; off = o2; X[15] &= 0x3fffffffu; }
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
			uint4 *sb = reinterpret_cast<uint4 *>(slot_bits + (size_t)k * 16);
#pragma unroll
			for (int j = 0; j < 4; ++j) sb[j] = make_uint4(X[4 * j], X[4 * j + 1], X[4 * j + 2], X[4 * j + 3]);
			SlotWs w;
			w.sb1_t1[0] = 0; w.sb1_t1[1] = 0; w.sb_code = 0;
			w.find_off = (uint16_t)off; w.window = (uint16_t)W;
			w.find_rc = (int8_t)rc; w.good_sb = 0; w.kind = (uint8_t)kind; w.unlock = unlock;
			w.tn = w.fn = w.mn = w.cc = 0; w.mcc = w.mnc = 0; w.sb1_crc = 0;
			ws[k] = w;
		}
		__syncwarp();
		st ^= 1;
	}
}

}  // namespace tb
