/* TEST INFRASTRUCTURE ONLY - drives the reference's own GSMTAP framing (src/tetra_gsmtap.c, compiled in place
 * into oracle/_ref/libtetra_ref.so) the way its upper MAC does (tetra_upper_mac.c:480-488): one
 * tetra_gsmtap_makemsg + tetra_gsmtap_sendmsg per CRC-good primitive.  The UDP socket of libosmocore's
 * gsmtap_util is replaced by a capture buffer. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <osmocom/core/msgb.h>
#include <osmocom/core/gsmtap.h>
#include <osmocom/core/gsmtap_util.h>

#include "tetra_common.h"
#include "tetra_tdma.h"
#include "tetra_gsmtap.h"
#include "oracle_records.h"

struct gsmtap_inst { int dummy; };
static struct gsmtap_inst g_inst;
static uint8_t *g_out;
static size_t g_cap, g_len, g_frames;

struct gsmtap_inst *gsmtap_source_init(const char *host, uint16_t port, int ofd_wq_mode)
{
	(void)host; (void)port; (void)ofd_wq_mode;
	return &g_inst;
}

int gsmtap_source_add_sink(struct gsmtap_inst *gti) { (void)gti; return 0; }

/* libosmocore's gsmtap_sendmsg writes the message to the socket and frees it */
int gsmtap_sendmsg(struct gsmtap_inst *gti, struct msgb *msg)
{
	(void)gti;
	if (g_out && g_len + msg->len <= g_cap)
		memcpy(g_out + g_len, msg->data, msg->len);
	g_len += msg->len;
	g_frames++;
	free(msg);
	return 0;
}

/* frames of the records with crc_ok, back to back; returns the bytes needed (written if they fit) */
size_t ref_gsmtap_frames(const struct tb_record *rec, size_t n, uint8_t *out, size_t cap, size_t *n_frames)
{
	static int inited;
	static struct tetra_mac_state tms;
	if (!inited) {
		tetra_gsmtap_init("localhost", 0);
		inited = 1;
	}
	g_out = out; g_cap = cap; g_len = 0; g_frames = 0;
	for (size_t i = 0; i < n; i++) {
		const struct tb_record *r = &rec[i];
		if (!r->crc_ok)                                   /* tetra_upper_mac.c:480-481 */
			continue;
		struct tetra_tdma_time tm;
		memset(&tm, 0, sizeof(tm));
		tm.tn = r->tn; tm.fn = r->fn; tm.mn = r->mn;
		struct msgb *m = tetra_gsmtap_makemsg(&tm, (enum tetra_log_chan)r->lchan, (uint8_t)(r->tn - 1), 0, 0, 0,
		                                      r->type1, r->type1_len, &tms);
		if (m)
			tetra_gsmtap_sendmsg(m);
	}
	if (n_frames)
		*n_frames = g_frames;
	g_out = NULL;
	return g_len;
}
