/*
 * tetra_text.h - the text the reference's PHY and lower MAC write to stdout / stderr, rebuilt from the
 * results of libtetra_b200 (host side, C).  Shared by tetra_rx_b200.c (file in, text out) and tetra_shim.c
 * (so that `tetra-rx` linked against the shim prints what the all-reference `tetra-rx` prints, line for
 * line, interleaved with the upper MAC's own output).
 *
 *   found SYNC training sequence in bit #N            phy/tetra_burst_sync.c:79
 *   (empty line) BURST                                phy/tetra_burst_sync.c:114-116
 *   #### SYNC burst at offset N?!?          (stderr)  phy/tetra_burst_sync.c:126,136
 *   #### could not find successive burst training sequence   (stderr)  :139
 *   BNCH FOLLOWS                                      lower_mac/tetra_lower_mac.c:170-173
 *   CRC COMP: 0x1d0f OK / CRC COMP: 0x.... WRONG      lower_mac/tetra_lower_mac.c:258-267
 *   <SB1|SB2|NDB|SCH/F> mn/fn/tn/sn type1: <bits>     lower_mac/tetra_lower_mac.c:264-265
 *   TMB-SAP SYNC CC ... TN ... FN ... MN ... MCC ... MNC ...   lower_mac/tetra_lower_mac.c:283-289
 */
#ifndef TETRA_TEXT_H
#define TETRA_TEXT_H

#include <stdint.h>
#include <stdio.h>

#include "tetra_b200.h"

/* t_phy_state.time as the text needs it: zero at start (tetra_burst_sync.c:34), +1 slot per LOCKED slot,
 * replaced by the SYNC PDU's time after a CRC-good SB1 */
struct tb200_text {
	uint32_t tn, fn, mn;
};

static inline const char *tb200_text_bits(const uint8_t *bits, unsigned int len)   /* osmo_ubit_dump */
{
	static char buf[512];
	unsigned int i;
	for (i = 0; i < len && i < sizeof(buf) - 1; i++)
		buf[i] = bits[i] ? '1' : '0';
	buf[i] = 0;
	return buf;
}

static inline unsigned int tb200_text_uint(const uint8_t *bits, unsigned int len)   /* bits_to_uint, tetra_common.c:31-39 */
{
	unsigned int v = 0;
	while (len--)
		v = (v << 1) | (*bits++ & 1);
	return v;
}

static inline const char *tb200_text_time(const struct tb200_text *t)              /* tetra_tdma_time_dump; sn is never set */
{
	static char buf[64];
	snprintf(buf, sizeof(buf), "%02u/%02u/%u/%03u", t->mn, t->fn, t->tn, 0u);
	return buf;
}

static inline void tb200_text_lock(uint32_t offset)
{
	printf("found SYNC training sequence in bit #%u\n", offset);
}

/* start of a LOCKED slot (tetra_burst_sync.c:113-143); returns the number of blocks the slot delivers */
static inline int tb200_text_slot(struct tb200_text *t, const struct tb200_slot *s)
{
	const int kind = s->flags & TB200_F_KIND_MASK;
	tb200_debug_time_advance(&t->tn, &t->fn, &t->mn, 1);
	printf("\nBURST\n");
	if (kind == TB200_KIND_NONE) {
		if (s->find_rc < 0)
			fprintf(stderr, "#### could not find successive burst training sequence\n");
		else
			fprintf(stderr, "#### SYNC burst at offset %u?!?\n", (unsigned int)s->find_off);
		return 0;
	}
	return kind == TB200_KIND_NDB_F ? 2 : 3;
}

static inline void tb200_text_crc(const char *name, const struct tb200_text *t, uint32_t crc, int ok,
                                  const uint8_t *type1, unsigned int len)
{
	printf("CRC COMP: 0x%04x ", crc & 0xffff);
	if (ok) {
		printf("OK\n");
		printf("%s %s type1: %s\n", name, tb200_text_time(t), tb200_text_bits(type1, len));
	} else
		printf("WRONG\n");
}

/* block `b` of the slot in the reference's delivery order (tetra_burst.c:347-373): SYNC burst SB1, BBK, SB2;
 * normal burst BBK, SCH/F or BBK, BLK1, BLK2.  type1 = the block's type-1 bits (one per byte), crc = the
 * slot's CRC word (tb200_set_crc_buffer).  Everything tp_sap_udata_ind prints before it calls the upper MAC. */
static inline void tb200_text_block(struct tb200_text *t, const struct tb200_slot *s, int b, uint32_t crc, const uint8_t *type1)
{
	const int kind = s->flags & TB200_F_KIND_MASK;
	const int ok_a = (s->flags & TB200_F_CRC_A) != 0, ok_b = (s->flags & TB200_F_CRC_B) != 0;
	if (kind == TB200_KIND_SB) {
		if (b == 0) {
			/* SB1 is printed with the time before the SYNC PDU is applied (time_str is taken on entry) */
			tb200_text_crc("SB1", t, crc, ok_a, type1, 60);
			printf("TMB-SAP SYNC CC %s(0x%02x) ", tb200_text_bits(type1 + 4, 6), tb200_text_uint(type1 + 4, 6));
			printf("TN %s(%u) ", tb200_text_bits(type1 + 10, 2), tb200_text_uint(type1 + 10, 2) + 1);
			printf("FN %s(%2u) ", tb200_text_bits(type1 + 12, 5), tb200_text_uint(type1 + 12, 5));
			printf("MN %s(%2u) ", tb200_text_bits(type1 + 17, 6), tb200_text_uint(type1 + 17, 6));
			printf("MCC %s(%u) ", tb200_text_bits(type1 + 31, 10), tb200_text_uint(type1 + 31, 10));
			printf("MNC %s(%u)\n", tb200_text_bits(type1 + 41, 14), tb200_text_uint(type1 + 41, 14));
			t->tn = s->time & 7u; t->fn = (s->time >> 3) & 31u; t->mn = (s->time >> 8) & 63u;   /* tetra_lower_mac.c:302 */
		} else if (b == 2) {
			if (s->flags & TB200_F_BNCH)
				printf("BNCH FOLLOWS\n");
			tb200_text_crc("SB2", t, crc >> 16, ok_b, type1, 124);
		}
	} else if (kind == TB200_KIND_NDB_F) {
		if (b == 1)
			tb200_text_crc("SCH/F", t, crc, ok_a, type1, 268);
	} else if (kind == TB200_KIND_NDB_2) {
		if (b == 1)
			tb200_text_crc("NDB", t, crc, ok_a, type1, 124);
		else if (b == 2)
			tb200_text_crc("NDB", t, crc >> 16, ok_b, type1, 124);
	}
}

/* offset of block b's type-1 bits inside the slot's type-1 string (include/tetra_b200.h, TB200_KIND_*) */
static inline unsigned int tb200_text_block_offset(int kind, int b)
{
	if (kind == TB200_KIND_SB) return b == 0 ? 0 : b == 1 ? 60 : 74;
	if (kind == TB200_KIND_NDB_F) return b == 0 ? 0 : 14;
	return b == 0 ? 0 : b == 1 ? 14 : 138;
}

#endif /* TETRA_TEXT_H */
