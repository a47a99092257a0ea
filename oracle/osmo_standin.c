/* TEST INFRASTRUCTURE ONLY - not part of the product path.
 *
 * Stand-in for the handful of libosmocore symbols the reference hot-path files
 * link against.  libosmocore itself is a third-party
 * dependency of osmo-tetra that is NOT present in /root/reference nor installed
 * in this image (src/Makefile:1-2 takes it from pkg-config, version un-pinned;
 * contrib/jenkins.sh:20 builds its master branch).
 *
 * The one piece of real arithmetic here is osmo_conv_decode(), which the
 * reference calls from lower_mac/viterbi_cch.c:65.  Its published algorithm
 * (libosmocore src/conv.c "osmo_conv_decode*" and src/conv_acc.c +
 * conv_acc_generic.c "osmo_conv_decode_acc", the latter taken for N<=4, K in
 * {5,7}) is restated below TWICE, independently:
 *
 *   variant ACC  - the accelerated path: int16 correlation metrics, start bias
 *                  127*N*K on state 0, butterflies over (2i, 2i+1) -> (i, i+8)
 *                  with the "sum0 >= sum1" survivor rule, min-subtraction every
 *                  INT16_MAX/(N*127) - K steps, trace back from state 0.
 *   variant GEN  - the generic path: accumulated |soft - expected| distance,
 *                  states scanned ascending, survivor replaced only on strict
 *                  improvement, K-1 zero-input flush steps, trace back from 0.
 *
 * PARITY STATUS: the tie-break behaviour of both variants is written from the
 * published algorithm, not diffed against a libosmocore build, and no test in
 * the reference pins Viterbi output on noisy input.  On
 * noisy input this is therefore "parity unpinned" for the libosmocore part;
 * clean-channel behaviour is pinned by the reference's own loop-back
 * (conv_enc_test.c:52-85).  tests/test_oracle.py checks ACC == GEN bit for bit
 * on random noisy blocks.  Select with TETRA_ORACLE_VITERBI=gen (default acc).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <osmocom/core/bits.h>
#include <osmocom/core/conv.h>
#include <osmocom/core/utils.h>
#include <osmocom/core/bitvec.h>

#include "../include/tetra_tie_rule.h"

/* the tie rule both restatements apply (include/tetra_tie_rule.h); run-time override for the tests */
static int g_tie = TETRA_VITERBI_TIE_DEFAULT;
void oracle_conv_set_tie(int tie) { g_tie = tie ? TETRA_TIE_KEEPS_HIGH_PRED : TETRA_TIE_KEEPS_LOW_PRED; }

/* ---------------------------------------------------------------- Viterbi -- */

#define MAX_STEPS 1024

static unsigned bitrev(unsigned v, int n)
{
	unsigned r = 0;
	for (int i = 0; i < n; i++)
		if (v & (1u << i))
			r |= 1u << (n - 1 - i);
	return r;
}

static int16_t sat16(int v)
{
	if (v > INT16_MAX) return INT16_MAX;
	if (v < INT16_MIN) return INT16_MIN;
	return (int16_t)v;
}

/* conv_acc-style decoder (flush termination, non-recursive codes, K=5 only). */
int oracle_conv_decode_acc(const struct osmo_conv_code *code, const sbit_t *in, ubit_t *out)
{
	const int N = code->N, K = code->K, ns = 1 << (K - 1), len = code->len;
	const int steps = len + K - 1;
	const int intrvl = INT16_MAX / (N * INT8_MAX) - K;
	int16_t sums[64], nsums[64];
	static uint8_t paths[MAX_STEPS][64];

	if (K != 5 || N > 4 || steps > MAX_STEPS || code->term != CONV_TERM_FLUSH)
		return -1;

	memset(sums, 0, sizeof(sums));
	sums[0] = INT8_MAX * N * K;

	for (int t = 0; t < steps; t++) {
		const sbit_t *seq = &in[N * t];
		for (int i = 0; i < ns / 2; i++) {
			/* branch (acc state 2i --input 0--> acc state i); acc numbering keeps
			 * the newest bit in the MSB, the code tables keep it in the LSB */
			unsigned o = code->next_output[bitrev(2 * i, K - 1)][0];
			int m = 0;
			for (int j = 0; j < N; j++)
				m += seq[j] * (((o >> (N - 1 - j)) & 1) ? -1 : 1);
			int s0 = sat16(sums[2 * i] + m), s1 = sat16(sums[2 * i + 1] - m);
			int s2 = sat16(sums[2 * i] - m), s3 = sat16(sums[2 * i + 1] + m);
			/* libosmocore: "sum0 >= sum1" keeps acc state 2i (oldest bit 0); the other rule makes it strict */
			const int hi = g_tie == TETRA_TIE_KEEPS_HIGH_PRED;
			if (hi ? s0 > s1 : s0 >= s1) { nsums[i] = s0; paths[t][i] = 0; }
			else                         { nsums[i] = s1; paths[t][i] = 1; }
			if (hi ? s2 > s3 : s2 >= s3) { nsums[i + ns / 2] = s2; paths[t][i + ns / 2] = 0; }
			else                         { nsums[i + ns / 2] = s3; paths[t][i + ns / 2] = 1; }
		}
		if (t % intrvl == 0) {
			int16_t mn = nsums[0];
			for (int i = 1; i < ns; i++)
				if (nsums[i] < mn) mn = nsums[i];
			for (int i = 0; i < ns; i++)
				nsums[i] -= mn;
		}
		memcpy(sums, nsums, sizeof(int16_t) * ns);
	}

	unsigned state = 0;
	for (int t = steps - 1; t >= 0; t--) {
		unsigned d = paths[t][state];
		if (t < len)
			out[t] = (state >> (K - 2)) & 1;
		state = ((state << 1) & (ns - 2)) | d;
	}
	return 0;
}

/* conv.c-style generic decoder (flush termination). */
int oracle_conv_decode_gen(const struct osmo_conv_code *code, const sbit_t *in, ubit_t *out)
{
	const int N = code->N, K = code->K, ns = 1 << (K - 1), len = code->len;
	const int steps = len + K - 1;
	const unsigned MAX_AE = 0x00ffffff;
	unsigned ae[64], ae_next[64];
	static uint8_t hist[MAX_STEPS][64];

	if (steps > MAX_STEPS || ns > 64 || code->term != CONV_TERM_FLUSH)
		return -1;

	for (int s = 0; s < ns; s++)
		ae[s] = MAX_AE;
	ae[0] = 0;

	for (int t = 0; t < steps; t++) {
		const sbit_t *sym = &in[N * t];
		int nb = (t < len) ? 2 : 1;     /* flush steps: input 0 only */
		for (int s = 0; s < ns; s++)
			ae_next[s] = MAX_AE;
		for (int s = 0; s < ns; s++) {
			for (int b = 0; b < nb; b++) {
				unsigned o = code->next_output[s][b];
				unsigned st = code->next_state[s][b];
				unsigned nae = ae[s];
				for (int m = 0; m < N; m++) {
					int ov = (o >> (N - 1 - m)) & 1;
					if (sym[m])
						nae += (ov ? sym[m] : -sym[m]) + 127;
				}
				/* libosmocore: strict, so the predecessor scanned first (the lower state) keeps a tie */
				if (g_tie == TETRA_TIE_KEEPS_HIGH_PRED ? ae_next[st] >= nae : ae_next[st] > nae) {
					ae_next[st] = nae;
					hist[t][st] = s;
				}
			}
		}
		memcpy(ae, ae_next, sizeof(unsigned) * ns);
	}

	unsigned state = 0;
	for (int t = steps - 1; t >= 0; t--) {
		unsigned prev = hist[t][state];
		if (t < len)
			out[t] = (code->next_state[prev][1] == state && code->next_state[prev][0] != state) ? 1 : 0;
		state = prev;
	}
	return 0;
}

#ifndef ORACLE_STANDIN_CONV_ONLY      /* pin_conv.c links the real libosmocore next to the two restatements */
int osmo_conv_decode(const struct osmo_conv_code *code, const sbit_t *input, ubit_t *output)
{
	static int variant = -1;
	if (variant < 0) {
		const char *e = getenv("TETRA_ORACLE_VITERBI");
		variant = (e && !strcmp(e, "gen")) ? 1 : 0;
	}
	return variant ? oracle_conv_decode_gen(code, input, output)
		       : oracle_conv_decode_acc(code, input, output);
}

/* ------------------------------------------------------------ small utils -- */

const char *get_value_string(const struct value_string *vs, uint32_t val)
{
	static char unk[32];
	for (; vs->str; vs++)
		if (vs->value == val)
			return vs->str;
	snprintf(unk, sizeof(unk), "unknown 0x%x", val);
	return unk;
}

char *osmo_ubit_dump(const uint8_t *bits, unsigned int len)
{
	static char buf[4100];
	unsigned int i;
	if (len > sizeof(buf) - 1)
		len = sizeof(buf) - 1;
	for (i = 0; i < len; i++) {
		switch (bits[i]) {
		case 0: buf[i] = '0'; break;
		case 0xfe: buf[i] = '?'; break;
		case 0xff: buf[i] = '-'; break;
		case 1: buf[i] = '1'; break;
		default: buf[i] = 'E'; break;
		}
	}
	buf[i] = 0;
	return buf;
}

char *osmo_hexdump(const unsigned char *b, int len)
{
	static char buf[4100];
	int o = 0;
	buf[0] = 0;
	for (int i = 0; i < len && o < (int)sizeof(buf) - 4; i++)
		o += snprintf(buf + o, sizeof(buf) - o, "%02x ", b[i]);
	return buf;
}

int osmo_pbit2ubit(ubit_t *out, const pbit_t *in, unsigned int num_bits)
{
	for (unsigned int i = 0; i < num_bits; i++)
		out[i] = (in[i / 8] >> (7 - (i % 8))) & 1;
	return num_bits;
}

int osmo_ubit2pbit(pbit_t *out, const ubit_t *in, unsigned int num_bits)
{
	unsigned int nbytes = (num_bits + 7) / 8;
	memset(out, 0, nbytes);
	for (unsigned int i = 0; i < num_bits; i++)
		if (in[i] & 1)
			out[i / 8] |= 1 << (7 - (i % 8));
	return nbytes;
}

int bitvec_set_bit(struct bitvec *bv, int bit)
{
	unsigned int n = bv->cur_bit;
	if (n / 8 >= bv->data_len)
		return -1;
	bv->data[n / 8] &= ~(1 << (7 - (n % 8)));
	if (bit)
		bv->data[n / 8] |= 1 << (7 - (n % 8));
	bv->cur_bit++;
	return 0;
}

int bitvec_set_uint(struct bitvec *bv, unsigned int in, unsigned int count)
{
	for (unsigned int i = 0; i < count; i++)
		if (bitvec_set_bit(bv, (in >> (count - 1 - i)) & 1) < 0)
			return -1;
	return 0;
}

#endif /* ORACLE_STANDIN_CONV_ONLY */

/* test hook: decode with an explicit variant (0 = acc, 1 = generic) and the TETRA mother
 * code tables rebuilt from the generator polynomials (viterbi_cch.c:28-48) */
int oracle_tetra_cch_decode(int variant, const sbit_t *in, int n, ubit_t *out)
{
	static uint8_t nout[16][2], nst[16][2];
	for (unsigned s = 0; s < 16; s++)
		for (unsigned b = 0; b < 2; b++) {
			unsigned d1 = s & 1, d2 = (s >> 1) & 1, d3 = (s >> 2) & 1, d4 = (s >> 3) & 1;
			unsigned g1 = b ^ d1 ^ d4, g2 = b ^ d2 ^ d3 ^ d4, g3 = b ^ d1 ^ d2 ^ d4, g4 = b ^ d1 ^ d3 ^ d4;
			nout[s][b] = (g1 << 3) | (g2 << 2) | (g3 << 1) | g4;
			nst[s][b] = ((s << 1) | b) & 15;
		}
	struct osmo_conv_code code;
	memset(&code, 0, sizeof(code));
	code.N = 4; code.K = 5; code.len = n;
	code.next_output = nout; code.next_state = nst;
	return variant ? oracle_conv_decode_gen(&code, in, out) : oracle_conv_decode_acc(&code, in, out);
}
