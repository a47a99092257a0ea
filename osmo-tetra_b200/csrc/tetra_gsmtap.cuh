/*
 * tetra_gsmtap.cuh - GSMTAP framing of the decoded blocks on the device.
 *
 * The reference wraps every CRC-good block that reaches its upper MAC into a GSMTAP frame for wireshark
 * (tetra_upper_mac.c:483-488 -> tetra_gsmtap_makemsg, tetra_gsmtap.c:31-63): a 16-byte struct gsmtap_hdr
 * (version 2, hdr_len 4 words, type TETRA_I1, timeslot = tn-1, frame_number = htonl((hn*60+mn)*18+fn) with
 * hn = 0, sub_type from lchan2gsmtap[], everything else 0) followed by the type-1 bits packed eight per byte,
 * first bit in the MSB (osmo_ubit2pbit), the last byte zero-padded.  Here: slot records + the LSB-first
 * packed type-1 words of the decode pass in, one contiguous byte stream of frames in delivery order out
 * (SB1, AACH, SB2 / AACH, SCH/F / AACH, BLK1, BLK2; blocks with a wrong CRC are skipped like
 * tetra_upper_mac.c:480-481 skips them).
 *
 * Three launches: per-tile byte/frame totals, one exclusive scan over the tiles, and the emit pass, which
 * assembles a tile's frames in shared memory and writes them out as one contiguous, coalesced run.
 * HBM-bound: 52 B read and at most 82 B written per slot.
 */
#pragma once
#include "tetra_kernels.cuh"

namespace tb {

constexpr int GT_THREADS = 256;               /* slots per tile */
constexpr int GT_SLOT_MAX = 82;               /* AACH 18 + 2 x (16 + 16) bytes */
constexpr int GT_HDR = 16;
/* enum values of libosmocore's gsmtap.h: GSMTAP_VERSION, GSMTAP_TYPE_TETRA_I1, GSMTAP_TETRA_* */
constexpr uint32_t GT_VERSION = 2, GT_TYPE_TETRA_I1 = 5;
constexpr uint32_t GT_BSCH = 1, GT_AACH = 2, GT_SCH_F = 5, GT_BNCH = 6;

__host__ __device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

/* bytes (low 40 bits) and frames (high 24 bits) of one slot */
__device__ __forceinline__ uint64_t gsmtap_slot_total(uint32_t flags)
{
	const int kind = flags & 3;
	const uint32_t a = (flags & F_CRC_A) ? 1 : 0, b = (flags & F_CRC_B) ? 1 : 0;
	uint32_t bytes = 0, frames = 0;
	if (kind == KIND_SB) { bytes = 18 + a * 24 + b * 32; frames = 1 + a + b; }
	else if (kind == KIND_NDB_F) { bytes = 18 + a * 50; frames = 1 + a; }
	else if (kind == KIND_NDB_2) { bytes = 18 + (a + b) * 32; frames = 1 + a + b; }
	return (uint64_t)bytes | ((uint64_t)frames << 40);
}

/* block-wide exclusive scan of one 64-bit value per thread (GT_THREADS threads); returns the exclusive prefix,
 * *total = sum over the block */
__device__ __forceinline__ uint64_t gt_block_scan(uint64_t v, uint64_t *warp_tot, uint64_t *total)
{
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint64_t inc = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= (unsigned)d) inc += o;
	}
	if (lane == 31) warp_tot[wib] = inc;
	__syncthreads();
	uint64_t base = 0, all = 0;
	for (unsigned w = 0; w < nw; w++) {
		const uint64_t t = warp_tot[w];
		if (w < wib) base += t;
		all += t;
	}
	__syncthreads();
	*total = all;
	return base + inc - v;
}

__global__ void __launch_bounds__(GT_THREADS)
k_gsmtap_sizes(const SlotOut *__restrict__ slots, uint64_t n, uint64_t *__restrict__ tile_tot)
{
	__shared__ uint64_t warp_tot[GT_THREADS / 32];
	const uint64_t i = (uint64_t)blockIdx.x * GT_THREADS + threadIdx.x;
	const uint64_t v = i < n ? gsmtap_slot_total(slots[i].flags) : 0;
	uint64_t total;
	gt_block_scan(v, warp_tot, &total);
	if (threadIdx.x == 0) tile_tot[blockIdx.x] = total;
}

/* one CTA: tile totals -> exclusive prefixes in place, grand total to tile_tot[n_tiles] */
__global__ void __launch_bounds__(1024)
k_gsmtap_scan(uint64_t *__restrict__ tile_tot, uint64_t n_tiles)
{
	__shared__ uint64_t warp_tot[32];
	const uint64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
	const uint64_t lo = umin64(per * threadIdx.x, n_tiles), hi = umin64(lo + per, n_tiles);
	uint64_t s = 0;
	for (uint64_t k = lo; k < hi; k++) s += tile_tot[k];
	uint64_t total;
	uint64_t run = gt_block_scan(s, warp_tot, &total);
	for (uint64_t k = lo; k < hi; k++) {
		const uint64_t t = tile_tot[k];
		tile_tot[k] = run;
		run += t;
	}
	if (threadIdx.x == 0) tile_tot[n_tiles] = total;
}

/* one frame into the tile's staging area (16-bit units; every frame length and offset is even).  OFF / LEN: the
 * block's place in the slot's type-1 string, compile-time so that every word index and shift is an immediate.
 * h1 = type | timeslot << 8, f_hi / f_lo = htonl(frame_number) as two 16-bit units. */
template <int OFF, int LEN>
__device__ __forceinline__ uint16_t *gt_put_frame(uint16_t *dst, const uint32_t *w, uint32_t h1, uint32_t f_hi, uint32_t f_lo,
                                                  uint32_t sub)
{
	dst[0] = (uint16_t)(GT_VERSION | (GT_HDR / 4) << 8);
	dst[1] = (uint16_t)h1;
	dst[2] = 0;                                                  /* arfcn */
	dst[3] = 0;                                                  /* signal_dbm, snr_db */
	dst[4] = (uint16_t)f_hi;
	dst[5] = (uint16_t)f_lo;
	dst[6] = (uint16_t)sub;                                      /* sub_type, antenna_nr */
	dst[7] = 0;                                                  /* sub_slot, res */
	constexpr int WORDS = LEN / 32, REM = LEN % 32;
	/* 32 type-1 bits (LSB first) -> 4 frame bytes (first bit in the MSB of the first byte): reverse the bits,
	 * then the bytes */
#pragma unroll
	for (int k = 0; k <= WORDS; k++) {
		const int p = OFF + 32 * k;
		if (k == WORDS && REM == 0) break;
		uint32_t x = (p & 31) ? __funnelshift_r(w[p >> 5], w[(p >> 5) + 1], p & 31) : w[p >> 5];
		if (k == WORDS) x &= (1u << REM) - 1;
		const uint32_t sw = __byte_perm(__brev(x), 0, 0x0123);
		dst[8 + 2 * k] = (uint16_t)sw;
		if (k < WORDS || REM > 16) dst[9 + 2 * k] = (uint16_t)(sw >> 16);
	}
	return dst + 8 + (LEN + 15) / 16;
}

__global__ void __launch_bounds__(GT_THREADS)
k_gsmtap_emit(const SlotOut *__restrict__ slots, const uint32_t *__restrict__ packed, uint64_t n,
              const uint64_t *__restrict__ tile_base, uint16_t *__restrict__ frames, uint64_t cap_bytes,
              uint64_t *__restrict__ slot_off)
{
	__shared__ uint64_t warp_tot[GT_THREADS / 32];
	__shared__ __align__(16) uint32_t pw[GT_THREADS * TYPE1_WORDS + 4];
	__shared__ __align__(16) uint16_t stage[GT_THREADS * GT_SLOT_MAX / 2 + 8];
	const uint64_t tile0 = (uint64_t)blockIdx.x * GT_THREADS;
	const unsigned cnt = (unsigned)umin64((uint64_t)GT_THREADS, n - tile0);
	const uint64_t n_tiles = gridDim.x;
	/* the host launches this pass before it knows the total: nothing is written when the buffer is too small
	 * (the host then reports the size needed) */
	if ((tile_base[n_tiles] & ((1ull << 40) - 1)) > cap_bytes)
		return;
	/* the tile's packed type-1 words, coalesced */
	if (cnt == GT_THREADS && ((uintptr_t)packed & 15) == 0) {        /* a full tile is 9216 bytes: whole uint4s */
		const uint4 *src = reinterpret_cast<const uint4 *>(packed + tile0 * TYPE1_WORDS);
		for (unsigned k = threadIdx.x; k < GT_THREADS * TYPE1_WORDS / 4; k += GT_THREADS)
			reinterpret_cast<uint4 *>(pw)[k] = src[k];
	} else {
		for (unsigned k = threadIdx.x; k < cnt * TYPE1_WORDS; k += GT_THREADS)
			pw[k] = packed[tile0 * TYPE1_WORDS + k];
	}
	if (threadIdx.x == 0) pw[cnt * TYPE1_WORDS] = 0;
	uint32_t flags = 0, time = 0;
	if (threadIdx.x < cnt) {
		flags = slots[tile0 + threadIdx.x].flags;
		time = slots[tile0 + threadIdx.x].time;
	}
	const uint64_t v = gsmtap_slot_total(flags);
	uint64_t total;
	const uint64_t excl = gt_block_scan(v, warp_tot, &total);    /* also orders the pw[] writes before the reads */
	const uint64_t base = tile_base[blockIdx.x] & ((1ull << 40) - 1);
	/* the staging area starts at the same 16-byte phase as the tile's run in the output, so that the copy-out
	 * can move whole uint4s */
	uint16_t *out = frames + base / 2;
	const unsigned phase = (unsigned)(((uintptr_t)out >> 1) & 7);
	const unsigned my_off = (unsigned)(excl & ((1ull << 40) - 1)), tile_bytes = (unsigned)(total & ((1ull << 40) - 1));
	if (threadIdx.x < cnt) {
		if (slot_off) slot_off[tile0 + threadIdx.x] = base + my_off;
		const int kind = flags & 3;
		const bool a = flags & F_CRC_A, b = flags & F_CRC_B;
		const uint32_t tn = time & 7u, fn = (time >> 3) & 31u, mn = (time >> 8) & 63u;
		const uint32_t fnum = mn * 18 + fn;                      /* tetra_tdma.c:96-99 with hn = 0 */
		const uint32_t h1 = GT_TYPE_TETRA_I1 | ((tn - 1) & 0xff) << 8;          /* timeslot tn - 1, tetra_upper_mac.c:484 */
		const uint32_t f_hi = ((fnum >> 24) & 0xff) | ((fnum >> 16) & 0xff) << 8, f_lo = ((fnum >> 8) & 0xff) | (fnum & 0xff) << 8;
		const uint32_t *w = pw + threadIdx.x * TYPE1_WORDS;
		uint16_t *d = stage + phase + my_off / 2;
		if (kind == KIND_SB) {
			if (a) d = gt_put_frame<0, 60>(d, w, h1, f_hi, f_lo, GT_BSCH);
			d = gt_put_frame<60, 14>(d, w, h1, f_hi, f_lo, GT_AACH);
			if (b) d = gt_put_frame<74, 124>(d, w, h1, f_hi, f_lo, (flags & F_BNCH) ? GT_BNCH : 0);
		} else if (kind == KIND_NDB_F) {
			d = gt_put_frame<0, 14>(d, w, h1, f_hi, f_lo, GT_AACH);
			if (a) d = gt_put_frame<14, 268>(d, w, h1, f_hi, f_lo, GT_SCH_F);
		} else if (kind == KIND_NDB_2) {
			d = gt_put_frame<0, 14>(d, w, h1, f_hi, f_lo, GT_AACH);
			if (a) d = gt_put_frame<14, 124>(d, w, h1, f_hi, f_lo, 0);
			if (b) d = gt_put_frame<138, 124>(d, w, h1, f_hi, f_lo, 0);
		}
	}
	__syncthreads();
	const unsigned nh = tile_bytes / 2;                          /* halves to write: stage[phase, phase + nh) -> out[0, nh) */
	const unsigned head = umin((8 - phase) & 7, nh), body = (nh - head) / 8, tail0 = head + body * 8;
	if (threadIdx.x < head) out[threadIdx.x] = stage[phase + threadIdx.x];
	const uint4 *sv = reinterpret_cast<const uint4 *>(stage + phase + head);
	uint4 *ov = reinterpret_cast<uint4 *>(out + head);
	for (unsigned k = threadIdx.x; k < body; k += GT_THREADS)
		ov[k] = sv[k];
	if (threadIdx.x < nh - tail0) out[tail0 + threadIdx.x] = stage[phase + tail0 + threadIdx.x];
}

}  // namespace tb
