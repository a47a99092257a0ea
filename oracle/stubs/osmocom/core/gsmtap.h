/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/gsmtap.h> (see bits.h in this directory
 * for why).  The GSMTAP v2 header and the constants tetra_gsmtap.c uses, restated from the published GSMTAP
 * header format (the same numbers wireshark's packet-gsmtap dissector uses); libosmocore is absent, so these
 * values are not checked against its header here: "parity unpinned" for the constants, while the way the frame
 * is put together comes from the reference's own tetra_gsmtap.c compiled in place. */
#pragma once
#include <stdint.h>

#define GSMTAP_VERSION           0x02
#define GSMTAP_TYPE_TETRA_I1     0x05     /* TETRA air interface */

#define GSMTAP_TETRA_BSCH        0x01
#define GSMTAP_TETRA_AACH        0x02
#define GSMTAP_TETRA_SCH_HU      0x03
#define GSMTAP_TETRA_SCH_HD      0x04
#define GSMTAP_TETRA_SCH_F       0x05
#define GSMTAP_TETRA_BNCH        0x06
#define GSMTAP_TETRA_STCH        0x07
#define GSMTAP_TETRA_TCH_F       0x08

struct gsmtap_hdr {
	uint8_t  version;        /* GSMTAP_VERSION */
	uint8_t  hdr_len;        /* length in 32-bit words */
	uint8_t  type;           /* GSMTAP_TYPE_* */
	uint8_t  timeslot;
	uint16_t arfcn;
	int8_t   signal_dbm;
	int8_t   snr_db;
	uint32_t frame_number;   /* network byte order */
	uint8_t  sub_type;       /* channel type */
	uint8_t  antenna_nr;
	uint8_t  sub_slot;
	uint8_t  res;
} __attribute__((packed));
