/*
 * tetra_afc.cuh - float_to_bits with its pseudo-AFC (float_to_bits.c:128-164, option -a) on the device.
 *
 * The reference's live pipeline slices the demodulator's float symbols with `float_to_bits -a` (src/receiver1udp:62): a
 * one-pole tracker of the symbols' mean,  filter = filter * (1.0 - filter_val) + (fl - filter_goal) * filter_val  for
 * symbols inside (-5, 5), is subtracted before the slicer.  The state is a float that is rounded once per symbol, so the
 * result is only reproduced by doing exactly these operations in exactly this order - a serial recurrence over the
 * whole stream.  It is a contraction (factor 1 - filter_val), which is what makes it parallel all the same:
 *
 *   speculate  a thread takes a chunk of L symbols, starts W symbols earlier from state 0 and runs the recurrence;
 *              after W = 24 / filter_val steps the start value has decayed below any float's last bit and the rounded
 *              sequences have, as a rule, merged with the true one.  It slices its chunk and notes the state it had at
 *              the chunk's start and at its end;
 *   verify     chunk 0 starts from the true state.  Chunk c is exact if chunk c-1 is and the state chunk c assumed at its
 *              start IS (bit for bit) the state chunk c-1 ended with - the host walks the two small arrays;
 *   repair     the first chunk that fails is redone from the true state by one thread, and the walk goes on.
 *
 * The outcome is the serial program's, not an approximation of it; a wrong guess costs time, never a bit.
 * Output: the hard bits packed eight per byte (TB200_IN_PACKED), which is what the search kernel reads next.
 */
#pragma once
#include "tetra_kernels.cuh"

namespace tb {

struct AfcParams {
	float filter_val, filter_goal;
	double keep;               /* 1.0 - (double)filter_val */
};

/* one step of float_to_bits.c:140-147; every operation rounded where the C program rounds it, no fused multiply-add */
__device__ __forceinline__ float afc_track(float f, float fl, const AfcParams &p)
{
	if (fl > -5.0f && fl < 5.0f) {
#ifdef TB_SIMT_EMULATION
		volatile double a = (double)f * p.keep;
		volatile float b = (fl - p.filter_goal) * p.filter_val;
		volatile float r = (float)(a + (double)b);
		return r;
#else
		const double a = __dmul_rn((double)f, p.keep);
		const float b = __fmul_rn(__fsub_rn(fl, p.filter_goal), p.filter_val);
		return __double2float_rn(__dadd_rn(a, (double)b));
#endif
	}
	return f;
}

__device__ __forceinline__ float afc_sub(float fl, float f)
{
#ifdef TB_SIMT_EMULATION
	volatile float r = fl - f;
	return r;
#else
	return __fsub_rn(fl, f);
#endif
}

/* thread c: symbols [c * L, (c + 1) * L), L a multiple of 16 (one output word = 16 symbols); warm-up from max(0, c * L - W).
 * first_only >= 0: only that chunk, from the given state (the repair step). */
__global__ void __launch_bounds__(128)
k_afc_chunks(const float *__restrict__ sym, uint64_t n_sym, AfcParams p, float f0, uint32_t L, uint64_t W,
             uint32_t *__restrict__ out, float *__restrict__ f_start, float *__restrict__ f_end, long long only, float only_state)
{
	const uint64_t n_chunks = (n_sym + L - 1) / L;
	uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (only >= 0) { if (c != 0) return; c = (uint64_t)only; }
	if (c >= n_chunks) return;
	const uint64_t start = c * L, end = start + L < n_sym ? start + L : n_sym;
	float f;
	if (only >= 0) {
		f = only_state;
	} else {
		const uint64_t w0 = start > W ? start - W : 0;
		f = w0 == 0 ? f0 : 0.0f;
		for (uint64_t i = w0; i < start; ++i) f = afc_track(f, sym[i], p);
	}
	f_start[c] = f;
	for (uint64_t i0 = start; i0 < end; i0 += 16) {
		uint32_t word = 0;
#pragma unroll 4
		for (int j = 0; j < 16; ++j) {
			if (i0 + j < end) {
				const float fl = sym[i0 + j];
				f = afc_track(f, fl, p);
				word |= slice_symbol(afc_sub(fl, f)) << (2 * j);
			}
		}
		out[i0 >> 4] = word;
	}
	f_end[c] = f;
}

/* the same program without -a: plain slicing, a word per thread */
__global__ void __launch_bounds__(256)
k_slice_symbols(const float *__restrict__ sym, uint64_t n_sym, uint32_t *__restrict__ out)
{
	const uint64_t n_words = (n_sym + 15) >> 4, stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
		uint32_t word = 0;
#pragma unroll
		for (int j = 0; j < 16; ++j)
			if (16 * w + j < n_sym) word |= slice_symbol(sym[16 * w + j]) << (2 * j);
		out[w] = word;
	}
}

}  // namespace tb
