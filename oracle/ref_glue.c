/* TEST INFRASTRUCTURE ONLY - not part of the product path.
 *
 * Glue that turns the UNMODIFIED reference lower-MAC receive chain (compiled in
 * place from /root/reference/src by oracle/Makefile into oracle/_ref/) into a
 * recorder.  It supplies exactly the symbols the 13 hot-path files leave
 * undefined:
 *
 *   upper_mac_prim_recv()      tetra_upper_mac.h:22   - records one parity record
 *                                                       per TMV-SAP primitive and
 *                                                       returns -1 (one call per
 *                                                       block, tetra_lower_mac.c:326-352)
 *   update_current_network()   crypto/tetra_crypto.c:416 - no-op (no key store)
 *   __wrap_tetra_find_train_seq  link-time wrapper (ld --wrap) around
 *                              phy/tetra_burst.c:269 that logs every search the
 *                              lock FSM performs (window, mask, result).
 *
 * tetra_lower_mac.c is #included (not copied) so that its file-static cell
 * state `_tcd` (tetra_lower_mac.c:113) can be reset between scenarios.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>

#include <osmocom/core/bits.h>
#include <lower_mac/tetra_lower_mac.c>   /* compiled in place, see header comment */

#include "oracle_records.h"

/* -------------------------------------------------------------- recorder -- */

static struct tb_record *g_rec;
static size_t g_nrec, g_caprec;
static struct tb_fsm_event *g_ev;
static size_t g_nev, g_capev;
static struct tetra_rx_state *g_trs;
static struct tetra_mac_state *g_tms;
static struct tetra_crypto_state *g_tcs;
static uint32_t g_call_index;
static int g_record_enabled = 1;
static int g_feedback;              /* emulate the upper MAC's AACH feedback (is_traffic), see oracle_records.h */
static char g_dumpdir[512];

int upper_mac_prim_recv(struct osmo_prim_hdr *op, void *priv)
{
	struct tetra_tmvsap_prim *tmvp = (struct tetra_tmvsap_prim *)op;
	struct tmv_unitdata_param *tup = &tmvp->u.unitdata;
	struct msgb *msg = op->msg;
	if (g_feedback && priv && tup->lchan == TETRA_LC_AACH) {
		struct tetra_mac_state *tms = priv;
		TB_EMULATE_RX_AACH(tms->cur_burst, msg->l1h, tup->tdma_time.fn);
	}
	if (!g_record_enabled)
		return -1;
	if (g_nrec == g_caprec) {
		g_caprec = g_caprec ? g_caprec * 2 : 4096;
		g_rec = realloc(g_rec, g_caprec * sizeof(*g_rec));
	}
	struct tb_record *r = &g_rec[g_nrec++];
	memset(r, 0, sizeof(*r));
	r->slot_bit = g_trs ? g_trs->bitbuf_start_bitnum : 0;
	r->lchan = tup->lchan;
	r->crc_ok = tup->crc_ok;
	r->blk_num = tup->blk_num;
	r->scrambling_code = tup->scrambling_code;
	r->tn = tup->tdma_time.tn;
	r->fn = tup->tdma_time.fn;
	r->mn = tup->tdma_time.mn;
	r->type1_len = msgb_l1len(msg);
	if (r->type1_len > sizeof(r->type1))
		r->type1_len = sizeof(r->type1);
	memcpy(r->type1, msg->l1h, r->type1_len);
	return -1;
}

void update_current_network(struct tetra_crypto_state *tcs, int mcc, int mnc)
{
	/* what crypto/tetra_crypto.c:416 does to the fields the lower MAC compares */
	tcs->mcc = mcc;
	tcs->mnc = mnc;
}

int __real_tetra_find_train_seq(const uint8_t *in, unsigned int end_of_in,
				uint32_t mask, unsigned int *offset);

int __wrap_tetra_find_train_seq(const uint8_t *in, unsigned int end_of_in,
				uint32_t mask, unsigned int *offset)
{
	unsigned int off = 0;
	int rc = __real_tetra_find_train_seq(in, end_of_in, mask, &off);
	if (rc >= 0)
		*offset = off;
	if (g_record_enabled) {
		if (g_nev == g_capev) {
			g_capev = g_capev ? g_capev * 2 : 4096;
			g_ev = realloc(g_ev, g_capev * sizeof(*g_ev));
		}
		struct tb_fsm_event *e = &g_ev[g_nev++];
		e->call_index = g_call_index;
		e->buf_start_bit = g_trs ? g_trs->bitbuf_start_bitnum : 0;
		e->window = end_of_in;
		e->mask = mask;
		e->rc = rc;
		e->offset = rc >= 0 ? off : 0;
	}
	return rc;
}

/* --------------------------------------------------------------- control -- */

static int g_saved_stdout = -1, g_saved_stderr = -1;

static void silence(void)
{
	fflush(stdout); fflush(stderr);
	int nul = open("/dev/null", O_WRONLY);
	g_saved_stdout = dup(1); g_saved_stderr = dup(2);
	dup2(nul, 1); dup2(nul, 2);
	close(nul);
}

static void unsilence(void)
{
	fflush(stdout); fflush(stderr);
	if (g_saved_stdout >= 0) { dup2(g_saved_stdout, 1); close(g_saved_stdout); g_saved_stdout = -1; }
	if (g_saved_stderr >= 0) { dup2(g_saved_stderr, 2); close(g_saved_stderr); g_saved_stderr = -1; }
}

/* fresh receiver: what tetra-rx.c:48-54 sets up, plus zeroed static state */
void ref_reset(void)
{
	free(g_trs); free(g_tms); free(g_tcs);
	g_trs = calloc(1, sizeof(*g_trs));
	g_tms = calloc(1, sizeof(*g_tms));
	g_tcs = calloc(1, sizeof(*g_tcs));
	tetra_mac_state_init(g_tms);
	g_tms->tcs = g_tcs;
	g_trs->burst_cb_priv = g_tms;
	memset(&_tcd, 0, sizeof(_tcd));
	memset(&t_phy_state, 0, sizeof(t_phy_state));
	g_nrec = 0;
	g_nev = 0;
	g_call_index = 0;
}

void ref_set_recording(int on) { g_record_enabled = on; }

/* switch the emulated AACH feedback on; traffic blocks are then dumped into `dumpdir` by the reference
 * (tetra_lower_mac.c:198-241) instead of being decoded.  Call after ref_reset(). */
void ref_set_feedback(int on, const char *dumpdir)
{
	g_feedback = on;
	snprintf(g_dumpdir, sizeof(g_dumpdir), "%s", dumpdir ? dumpdir : "");
	if (g_tms)
		g_tms->dumpdir = g_dumpdir[0] ? g_dumpdir : NULL;
}

/* feed like tetra-rx.c:82-95: read() chunks of `chunk` bytes (64 there) */
long ref_feed(const uint8_t *bits, size_t n, unsigned int chunk, int quiet)
{
	uint8_t buf[4096];
	if (!g_trs)
		ref_reset();
	if (chunk == 0 || chunk > sizeof(buf))
		return -1;
	if (quiet)
		silence();
	for (size_t pos = 0; pos < n; pos += chunk) {
		unsigned int len = (n - pos < chunk) ? (unsigned int)(n - pos) : chunk;
		memcpy(buf, bits + pos, len);
		g_call_index++;
		tetra_burst_sync_in(g_trs, buf, len);
	}
	if (quiet)
		unsilence();
	return (long)g_nrec;
}

/* The same with a read() size per call and the reference's return value of every call
 * (tetra_burst_sync.c:66-106: len, 0 while KNOW_FSTART waits, -1 when the UNLOCKED search fails). */
long ref_feed_calls(const uint8_t *bits, const uint32_t *lens, size_t n_calls, int32_t *rc_out, int quiet)
{
	uint8_t buf[4096];
	size_t pos = 0;
	if (!g_trs)
		ref_reset();
	if (quiet)
		silence();
	for (size_t i = 0; i < n_calls; i++) {
		if (lens[i] > sizeof(buf))
			break;
		memcpy(buf, bits + pos, lens[i]);
		pos += lens[i];
		g_call_index++;
		const int rc = tetra_burst_sync_in(g_trs, buf, lens[i]);
		if (rc_out)
			rc_out[i] = rc;
	}
	if (quiet)
		unsilence();
	return (long)g_nrec;
}

/* Drive the lower MAC directly at the TP-SAP seam (drop-in depth C of SURVEY 8b). */
void ref_tp_sap(int type, int blk_num, const uint8_t *bits, unsigned int len, int quiet)
{
	if (!g_trs)
		ref_reset();
	if (quiet)
		silence();
	tp_sap_udata_ind((enum tp_sap_data_type)type, blk_num, bits, len, g_tms);
	if (quiet)
		unsilence();
}

size_t ref_num_records(void) { return g_nrec; }
const struct tb_record *ref_records(void) { return g_rec; }
size_t ref_num_events(void) { return g_nev; }
const struct tb_fsm_event *ref_events(void) { return g_ev; }
uint32_t ref_cell_scramb_init(void) { return _tcd.scramb_init; }
int ref_rx_state(void) { return g_trs ? (int)g_trs->state : -1; }
void ref_set_time(uint32_t tn, uint32_t fn, uint32_t mn)
{
	t_phy_state.time.tn = tn; t_phy_state.time.fn = fn; t_phy_state.time.mn = mn;
}
void ref_get_time(uint32_t *tn, uint32_t *fn, uint32_t *mn)
{
	*tn = t_phy_state.time.tn; *fn = t_phy_state.time.fn; *mn = t_phy_state.time.mn;
}

/* silence wrappers for the chatty TX-side builders (they printf phase sums) */
int ref_build_sync_burst(uint8_t *buf, const uint8_t *sb, const uint8_t *bb, const uint8_t *bkn)
{
	silence();
	int rc = build_sync_c_d_burst(buf, sb, bb, bkn);
	unsilence();
	return rc;
}

int ref_build_norm_burst(uint8_t *buf, const uint8_t *bkn1, const uint8_t *bb, const uint8_t *bkn2, int two)
{
	silence();
	int rc = build_norm_c_d_burst(buf, bkn1, bb, bkn2, two);
	unsilence();
	return rc;
}

void tetra_rm3014_init(void);
void ref_rm3014_init(void)
{
	silence();
	tetra_rm3014_init();
	unsilence();
}

/* ---- config 1: the reference's own conv_enc_test generator -------------------
 * build_sb() (conv_enc_test.c:198-305) prints the 510-bit SYNC burst as ASCII at :302;
 * capture that line.  `r` plays the role of rand() at conv_enc_test.c:337-339. */
extern uint8_t pdu_sync[8];
void testpdu_init(void);
int build_sb(void);
int build_ndb_schf(void);

static int capture_burst(int (*fn)(void), const char *tag, uint8_t *out510)
{
	char path[] = "/tmp/tetra_ref_capXXXXXX";
	int fd = mkstemp(path);
	if (fd < 0)
		return -1;
	fflush(stdout);
	int saved = dup(1);
	dup2(fd, 1);
	fn();
	fflush(stdout);
	dup2(saved, 1);
	close(saved);
	off_t sz = lseek(fd, 0, SEEK_END);
	if (sz <= 0 || sz > (1 << 24)) { close(fd); unlink(path); return -1; }
	char *buf = malloc((size_t)sz + 1);
	lseek(fd, 0, SEEK_SET);
	ssize_t got = read(fd, buf, (size_t)sz);
	close(fd);
	unlink(path);
	int rc = -1;
	if (got > 0) {
		buf[got] = 0;
		char *p = strstr(buf, tag);
		if (p) {
			p += strlen(tag);
			rc = 0;
			for (int i = 0; i < 510; i++) {
				if (p[i] != '0' && p[i] != '1') { rc = -1; break; }
				out510[i] = p[i] - '0';
			}
		}
	}
	free(buf);
	return rc;
}

int ref_conv_enc_test_sb(uint32_t r, uint8_t *out510)
{
	static int inited;
	if (!inited) { silence(); testpdu_init(); unsilence(); inited = 1; }
	osmo_store32le(r, pdu_sync);
	osmo_store32le(r, pdu_sync + 4);
	return capture_burst(build_sb, "cont sync DL burst: ", out510);
}

extern uint8_t pdu_schf[268];
int ref_conv_enc_test_ndb(uint32_t r, uint8_t *out510)
{
	static int inited;
	if (!inited) { silence(); testpdu_init(); unsilence(); inited = 1; }
	osmo_store32le(r, pdu_schf);
	osmo_store32le(r, pdu_schf + 4);
	return capture_burst(build_ndb_schf, "cont norm DL burst: ", out510);
}
