/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/prim.h>
 * (see bits.h in this directory for why). */
#pragma once
#include <osmocom/core/msgb.h>      /* the real header pulls these in, crypto/tetra_crypto.c relies on it */
#include <osmocom/core/talloc.h>
#include <stdint.h>

struct msgb;

enum osmo_prim_operation {
	PRIM_OP_REQUEST,
	PRIM_OP_RESPONSE,
	PRIM_OP_INDICATION,
	PRIM_OP_CONFIRM,
};

struct osmo_prim_hdr {
	unsigned int sap;
	unsigned int primitive;
	enum osmo_prim_operation operation;
	struct msgb *msg;
};
