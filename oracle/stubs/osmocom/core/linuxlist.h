/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/linuxlist.h>
 * (see bits.h in this directory for why). */
#pragma once

struct llist_head {
	struct llist_head *next, *prev;
};

#define INIT_LLIST_HEAD(ptr) do { (ptr)->next = (ptr); (ptr)->prev = (ptr); } while (0)

/* the usual circular doubly linked list helpers the upper MAC uses (tetra_llc.c:36-105), written
 * out here because libosmocore is absent; only needed when the real upper MAC is linked (tests/test_program.py) */
#include <stddef.h>
#define LLIST_HEAD_INIT(name) { &(name), &(name) }
#define llist_entry(ptr, type, member) ((type *)((char *)(ptr) - offsetof(type, member)))
#define llist_for_each_entry(pos, head, member) \
	for (pos = llist_entry((head)->next, __typeof__(*pos), member); &pos->member != (head); \
	     pos = llist_entry(pos->member.next, __typeof__(*pos), member))
static inline void llist_add(struct llist_head *n, struct llist_head *head)
{
	n->next = head->next; n->prev = head; head->next->prev = n; head->next = n;
}
static inline void llist_del(struct llist_head *e)
{
	e->next->prev = e->prev; e->prev->next = e->next; e->next = e->prev = NULL;
}
