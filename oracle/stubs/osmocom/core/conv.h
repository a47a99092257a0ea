/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/conv.h>
 * (see bits.h in this directory for why). */
#pragma once
#include <stdint.h>
#include <osmocom/core/bits.h>

enum osmo_conv_term {
	CONV_TERM_FLUSH = 0,
	CONV_TERM_TRUNCATION,
	CONV_TERM_TAIL_BITING,
};

struct osmo_conv_code {
	int N;                               /* outputs per input bit */
	int K;                               /* constraint length */
	int len;                             /* number of data bits */
	enum osmo_conv_term term;
	const uint8_t (*next_output)[2];
	const uint8_t (*next_state)[2];
	const uint8_t *next_term_output;
	const uint8_t *next_term_state;
	const int *puncture;
};

int osmo_conv_decode(const struct osmo_conv_code *code, const sbit_t *input, ubit_t *output);
