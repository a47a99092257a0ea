/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/bitvec.h>,
 * needed only by the reference's test-PDU generator (testpdu.c). */
#pragma once
#include <stdint.h>

struct bitvec {
	unsigned int cur_bit;
	unsigned int data_len;
	uint8_t *data;
};

int bitvec_set_bit(struct bitvec *bv, int bit);
int bitvec_set_uint(struct bitvec *bv, unsigned int in, unsigned int count);
