/* TEST INFRASTRUCTURE ONLY - record layouts shared by the reference recorder
 * (ref_glue.c), the CPU restatement (tetra_oracle.c) and the Python tests.
 *
 * One tb_record per TMV-SAP primitive that reaches upper_mac_prim_recv()
 * (tetra_upper_mac.h:22), i.e. per tp_sap_udata_ind() call, in call order
 * (SURVEY.md section 8a "parity record").  Fields mirror struct
 * tmv_unitdata_param (tetra_prim.h:25-33) plus the type-1 bits at msg->l1h. */
#pragma once
#include <stdint.h>

struct tb_record {
	uint32_t slot_bit;          /* absolute bit number of the burst start (uint32 like tetra_burst_sync.h:16) */
	uint8_t  lchan;             /* enum tetra_log_chan, tetra_common.h:22-39 */
	uint8_t  crc_ok;
	uint8_t  blk_num;
	uint8_t  tn;
	uint8_t  fn;
	uint8_t  mn;
	uint16_t type1_len;
	uint32_t scrambling_code;
	uint8_t  type1[272];        /* one bit per byte, 268 max (SCH/F) */
};                                  /* 288 bytes */

/* One per tetra_find_train_seq() call made by the lock FSM. */
struct tb_fsm_event {
	uint32_t call_index;        /* 1-based tetra_burst_sync_in() call */
	uint32_t buf_start_bit;     /* trs->bitbuf_start_bitnum at the search */
	uint32_t window;            /* end_of_in */
	uint32_t mask;
	int32_t  rc;                /* enum tetra_train_seq or -1 */
	uint32_t offset;
};                                  /* 24 bytes */

/* What the reference's upper MAC feeds back into the lower MAC when it parses an AACH block
 * (rx_aach, tetra_upper_mac.c:423-455 with macpdu_decode_access_assign, tetra_mac_pdu.c:257-330):
 * on frames other than 18, headers 1..3 carry the downlink usage marker in field 1 (bits [2,8));
 * a marker above 3 means the slot carries traffic.  The stealing flags are reset.  Used by the
 * recorders of the tests to exercise tetra_lower_mac.c:190-241 and the shim's mirror of it. */
#define TB_EMULATE_RX_AACH(cur_burst, bits, fn) do { \
		unsigned hdr_ = ((bits)[0] << 1) | (bits)[1], f1_ = 0, dl_ = 0; \
		for (int i_ = 0; i_ < 6; i_++) f1_ = (f1_ << 1) | ((bits)[2 + i_] & 1); \
		if ((fn) != 18 && hdr_ != 0) dl_ = f1_; \
		(cur_burst).is_traffic = dl_ > 3 ? (int)dl_ : 0; \
		(cur_burst).blk1_stolen = 0; (cur_burst).blk2_stolen = 0; \
	} while (0)
