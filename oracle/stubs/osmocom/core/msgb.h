/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/msgb.h>
 * (see bits.h in this directory for why).  Only the members and helpers that
 * tetra_lower_mac.c touches; the message body lives in the same allocation so a
 * plain free() (talloc_free stand-in) releases it. */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <osmocom/core/linuxlist.h>
#include <osmocom/core/utils.h>       /* the real msgb.h pulls these in; upper-MAC files rely on it */
#include <osmocom/core/bits.h>
#include <osmocom/core/talloc.h>

struct msgb {
	struct llist_head list;
	unsigned char *l1h, *l2h, *l3h, *l4h;
	unsigned long cb[5];
	uint16_t data_len;
	uint16_t len;
	unsigned char *head;
	unsigned char *tail;
	unsigned char *data;
	unsigned char _data[0];
};

#ifndef MSGB_STANDIN_SLACK
#define MSGB_STANDIN_SLACK 16384
#endif
static inline struct msgb *msgb_alloc(uint16_t size, const char *name)
{
	/* MSGB_STANDIN_SLACK zeroed bytes behind the buffer: the reference's upper MAC reads (and prints) well past the
	 * end of short or damaged PDUs (negative length fields); with the slack both builds print the same zeros
	 * there instead of whatever the heap holds, which makes whole-program output comparable */
	struct msgb *m = (struct msgb *)calloc(1, sizeof(*m) + size + MSGB_STANDIN_SLACK);
	(void)name;
	if (!m)
		return NULL;
	m->data_len = size;
	m->head = m->data = m->tail = m->_data;
	return m;
}

static inline unsigned char *msgb_put(struct msgb *m, unsigned int len)
{
	unsigned char *tmp = m->tail;
	m->tail += len;
	m->len += len;
	return tmp;
}

static inline unsigned int msgb_l1len(const struct msgb *m) { return m->tail - m->l1h; }
static inline unsigned int msgb_l2len(const struct msgb *m) { return m->tail - m->l2h; }
static inline unsigned int msgb_l3len(const struct msgb *m) { return m->tail - m->l3h; }
