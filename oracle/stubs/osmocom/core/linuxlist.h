/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/linuxlist.h>
 * (see bits.h in this directory for why). */
#pragma once

struct llist_head {
	struct llist_head *next, *prev;
};

#define INIT_LLIST_HEAD(ptr) do { (ptr)->next = (ptr); (ptr)->prev = (ptr); } while (0)
