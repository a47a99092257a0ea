/*
 * tetra_shim.c - drop-in for the reference's PHY + lower MAC objects.
 *
 * Link this file and libtetra_b200.so INSTEAD of libosmo-tetra-phy.a (phy/tetra_burst_sync.o,
 * phy/tetra_burst.o) and lower_mac/tetra_lower_mac.o (src/Makefile:13-26); tetra-rx.c and the
 * whole upper MAC stay untouched.  It exports the symbols tetra-rx links against
 *     int tetra_burst_sync_in(struct tetra_rx_state *, uint8_t *, unsigned int)   tetra_burst_sync.h:24
 *     struct tetra_phy_state t_phy_state                                          tetra_common.h:44-47
 * and delivers, per decoded block and in stream order on the calling thread, the same
 * TMV-SAP primitive the reference builds (tetra_lower_mac.c:129-140,162-167,276-352) to
 *     int upper_mac_prim_recv(struct osmo_prim_hdr *op, void *priv)              tetra_upper_mac.h:22
 *
 * Bits are queued and decoded on the GPU in batches: a batch ends when TETRA_B200_BATCH_BITS (default 8 Mi bits) are
 * queued, when TETRA_B200_FLUSH_MS (default 250) milliseconds have passed since the last one (a demodulator piped into
 * tetra-rx delivers 36 kbit/s: primitives then reach the upper MAC four times a second instead of every four minutes),
 * and whenever the length of the reads changes: read() on a pipe returns what is there (tetra-rx.c:82-95), the library
 * models one run of equal-length reads per call, so any sequence of read sizes gives the reference's search windows.
 * TETRA_B200_BATCH_BITS=0 decodes on every call and returns the reference's own return value (len, 0 while
 * KNOW_FSTART waits, -1 when the UNLOCKED search fails, tetra_burst_sync.c:77-78,93-94); batched calls return len
 * (what the result will be is not known yet; tetra-rx.c:94 ignores it).
 * tetra-rx has no end-of-stream call: it reads until read() returns 0, prints "EOF", frees its state and
 * exits (tetra-rx.c:82-102).  Built with -DTETRA_B200_SHIM_WRAP_READ and linked with -Wl,--wrap=read the
 * shim sees that read() itself: when the descriptor that feeds tetra_burst_sync_in() reports end of file
 * the queued tail is decoded and delivered before read() returns, i.e. before "EOF" is printed and before
 * the receiver state is freed - the program's output is then identical to the all-reference build
 * (tests/test_program.py).  Other callers end the stream with tetra_b200_shim_flush() while the receiver
 * state is still alive; an atexit() handler is only the last resort.  Compiled against the reference's
 * own headers.
 *
 * Drop-in depths:
 *   default                    PHY + lower MAC on the GPU, primitives to upper_mac_prim_recv()  (depth B)
 *   -DTETRA_B200_SHIM_L0_ONLY  only the PHY (lock, training-sequence search, slot classification) on the
 *                              GPU; every delivered slot goes to the REFERENCE's own tetra_burst_rx_cb()
 *                              (phy/tetra_burst.c:341-379, keep phy/tetra_burst.o) and from there into the
 *                              reference's tp_sap_udata_ind() / lower MAC: isolates sync + slicing parity (depth A)
 *   depth C (GPU lower MAC behind the reference's PHY) is the batched leaf tb200_decode_blocks().
 *
 * Upper-MAC feedback: when the upper MAC marks the current slot as traffic (tms->cur_burst.is_traffic, set in
 * rx_aach, tetra_upper_mac.c:444-452) the shim withholds the same primitives the reference lower MAC withholds
 * (tetra_lower_mac.c:190-241) and writes the same <dumpdir>/traffic_*.out / .txt files for SCH/F-shaped
 * traffic slots.  Not reproduced: the dump of a 216-bit second block (the reference fills half of it from
 * uninitialised memory).  The stdout / stderr text of the reference's PHY and lower MAC is printed from
 * tetra_text.h in the reference's order (TETRA_B200_TEXT=0 switches it off).  Reads of more than 296 bits are rejected
 * (one slot per call could not keep up with them: the reference then drops bits in make_bitbuf_space).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <osmocom/core/msgb.h>
#include <osmocom/core/talloc.h>

#include <tetra_common.h>
#include <tetra_tdma.h>
#include <tetra_prim.h>
#include <tetra_upper_mac.h>
#include <phy/tetra_burst.h>
#include <phy/tetra_burst_sync.h>
#include <crypto/tetra_crypto.h>

#include "tetra_b200.h"
#include "tetra_text.h"

struct tetra_phy_state t_phy_state;

/* crypto/tetra_crypto.c:416; weak so that callers without the crypto archive (test recorders) still link */
void update_current_network(struct tetra_crypto_state *tcs, int mcc, int mnc) __attribute__((weak));

#ifdef TETRA_B200_SHIM_L0_ONLY
/* the reference's slicer, phy/tetra_burst.c:341-379 (tetra_burst_sync.c:36 forward-declares it the same way) */
void tetra_burst_rx_cb(const uint8_t *burst, unsigned int len, enum tetra_train_seq type, void *priv);
#endif

#define SHIM_HIST 8192u      /* bits of earlier batches kept in front of the batch: a slot may start in them */

static struct {
	tb200_ctx *ctx;
	uint8_t *bits;
	size_t n_bits, cap_bits, batch_bits;
	uint8_t *hist;         /* [SHIM_HIST bits of history][the batch = bits] */
	size_t n_hist;
	uint64_t fed_before;   /* stream bits handed to the library before this batch */
	unsigned int chunk;    /* length of the reads of the queued run */
	int started;
	int flush_ms;          /* wall-clock flush interval (0: off) */
	struct timespec last_flush;
	unsigned int calls_since_clock;
	int finished;          /* the stream was flushed: a second flush (atexit after an explicit one) does nothing */
	struct tetra_rx_state *trs;
	struct tb200_slot *slots;
	uint8_t *type1;
	struct tb200_record *rec;
	uint32_t *crc;         /* CRC registers per slot, for the text */
	int text;              /* print what the reference's PHY + lower MAC print (TETRA_B200_TEXT=0 switches it off) */
	struct tb200_text txt;
	size_t max_slots;
} S;

static void shim_die(const char *msg)
{
	fprintf(stderr, "tetra_b200 shim: %s\n", msg);
	exit(1);
}

#ifndef TETRA_B200_SHIM_L0_ONLY
/* what tp_sap_udata_ind does after the arithmetic: allocate the primitive, fill it, hand it up,
 * re-invoke while the upper MAC consumed only part of the block (tetra_lower_mac.c:326-352).
 * This function and dump_traffic_schf() below follow the reference statement by statement ON PURPOSE: they are the
 * contract with the untouched upper MAC (which fields of struct tetra_tmvsap_prim it reads, how it walks msg->head /
 * l1h between re-invocations, the names and the 690-word frame layout of the traffic dump files it appends to) -
 * host-side glue at the seam, not part of the accelerated path. */
static void deliver(const struct tb200_record *r, void *priv)
{
	struct tetra_tmvsap_prim *ttp = talloc_zero(NULL, struct tetra_tmvsap_prim);
	struct tmv_unitdata_param *tup = &ttp->u.unitdata;
	struct msgb *msg = ttp->oph.msg = msgb_alloc(412, "tmvsap_prim");
	ttp->oph.sap = TETRA_SAP_TMV;
	ttp->oph.primitive = PRIM_TMV_UNITDATA;
	ttp->oph.operation = PRIM_OP_INDICATION;
	tup->lchan = r->lchan;
	tup->crc_ok = r->crc_ok;
	tup->scrambling_code = r->scrambling_code;
	tup->blk_num = r->blk_num;
	tup->tdma_time.tn = r->tn;
	tup->tdma_time.fn = r->fn;
	tup->tdma_time.mn = r->mn;
	msg->l1h = msgb_put(msg, r->type1_len);
	memcpy(msg->l1h, r->type1, r->type1_len);

	uint32_t offset = 0;
	uint8_t *orig_head = msg->head, *orig_tail = msg->tail;
	while (offset < (uint32_t)(r->type1_len - 16)) {
		int pdu_bits = upper_mac_prim_recv(&ttp->oph, priv);
		if (pdu_bits < 0)
			break;
		offset += pdu_bits;
		msg->head = orig_head + offset;
		msg->tail = orig_tail;
		msg->len = msg->tail - msg->head;
		msg->l1h = msg->head;
		msg->l2h = msg->l3h = msg->l4h = 0;
	}
	talloc_free(msg);
	talloc_free(ttp);
}
#endif

/* raw bits of the slot that starts at (wrapping) stream bit slot_bit, from the retained history + batch */
static const uint8_t *slot_raw_bits(uint32_t slot_bit)
{
	const uint64_t buf0 = S.fed_before - S.n_hist;             /* stream bit of hist[SHIM_HIST - n_hist] */
	const uint64_t abs = buf0 + (uint32_t)(slot_bit - (uint32_t)buf0);
	if (abs < buf0 || abs + TB200_BITS_PER_SLOT > S.fed_before + S.n_bits)
		shim_die("slot outside the retained bits");
	return S.hist + (SHIM_HIST - S.n_hist) + (abs - buf0);
}

#ifndef TETRA_B200_SHIM_L0_ONLY
/* The traffic dump of an SCH/F-shaped slot (tetra_lower_mac.c:198-241): the 432 descrambled type-4 bits as
 * +-127 soft values in the 690-word frame the ETSI codec tools read, appended to
 * <dumpdir>/traffic_<usage>_<tsn>.out, and the SSI appended to the matching .txt.  Host-side mirror: the
 * scrambling sequence is the LFSR of tetra_scramb.c:34-50 (fb = parity(state & 0xDB710641), shifted in at bit 31). */
static void dump_traffic_schf(const struct tetra_mac_state *tms, const uint8_t *burst, uint32_t code)
{
	char fname[4096];
	int16_t block[690];
	uint8_t type4[432];
	uint32_t st = code;
	for (int m = 0; m < 432; m++) {
		uint32_t fb = 0;
		if (code) {
			fb = (uint32_t)__builtin_parity(st & 0xDB710641u);
			st = (st >> 1) | (fb << 31);
		}
		type4[m] = burst[m < 216 ? 14 + m : 66 + m] ^ (uint8_t)fb;       /* tetra_burst.c:363-373 */
	}
	snprintf(fname, sizeof(fname), "%s/traffic_%d_%d.out", tms->dumpdir, tms->cur_burst.is_traffic, tms->tsn);
	FILE *f = fopen(fname, "ab");
	if (!f) {
		fprintf(stderr, "Could not open dump file %s for writing\n", fname);
		exit(1);
	}
	memset(block, 0x00, sizeof(block));
	for (int i = 0; i < 6; i++)
		block[115 * i] = 0x6b21 + i;
	for (int i = 0; i < 114; i++) block[1 + i] = type4[i] ? -127 : 127;
	for (int i = 0; i < 114; i++) block[116 + i] = type4[114 + i] ? -127 : 127;
	for (int i = 0; i < 114; i++) block[231 + i] = type4[228 + i] ? -127 : 127;
	for (int i = 0; i < 90; i++) block[346 + i] = type4[342 + i] ? -127 : 127;
	fwrite(block, sizeof(int16_t), 690, f);
	fclose(f);
	snprintf(fname, sizeof(fname), "%s/traffic_%d_%d.txt", tms->dumpdir, tms->cur_burst.is_traffic, tms->tsn);
	f = fopen(fname, "a");
	if (f) {
		fprintf(f, "%d\n", tms->ssi);
		fclose(f);
	}
}
#endif

static void shim_run(int final)
{
	if (!S.ctx || S.finished || (!S.n_bits && !final))
		return;
	if (final)
		S.finished = 1;
	uint32_t flags = (S.started ? 0 : TB200_FRESH) | (final ? TB200_FINAL : 0);
	if (S.chunk) {                    /* the reads of this run */
		struct tb200_options opt;
		tb200_default_options(&opt);
		opt.chunk_bits = S.chunk;
		opt.output = TB200_OUT_UNPACKED;
		if (tb200_set_options(S.ctx, &opt) != 0)
			shim_die(tb200_last_error(S.ctx));
	}
	clock_gettime(CLOCK_MONOTONIC, &S.last_flush);
	long n = tb200_rx_stream_host(S.ctx, S.bits, S.n_bits, flags, S.slots, S.type1, NULL, S.max_slots);
	if (n < 0) {
		fprintf(stderr, "tetra_b200 shim: %s\n", tb200_last_error(S.ctx));
		exit(1);
	}
	S.started = 1;
	void *priv = S.trs ? S.trs->burst_cb_priv : NULL;
#ifdef TETRA_B200_SHIM_L0_ONLY
	/* what the LOCKED arm of tetra_burst_sync_in does per slot (phy/tetra_burst_sync.c:113-143): advance the
	 * slot counter, then hand a slot whose training sequence sits where it should to the reference's slicer */
	for (long i = 0; i < n; i++) {
		const struct tb200_slot *sl = &S.slots[i];
		tetra_tdma_time_add_tn(&t_phy_state.time, 1);
		if ((sl->flags & TB200_F_KIND_MASK) == TB200_KIND_NONE)
			continue;
		tetra_burst_rx_cb(slot_raw_bits(sl->slot_bit), TB200_BITS_PER_SLOT, (enum tetra_train_seq)sl->find_rc, priv);
	}
#else
	size_t nrec = tb200_expand_records(S.slots, S.type1, (size_t)n, S.rec, 3 * S.max_slots);
	struct tetra_mac_state *tms = priv;
	/* lock acquisitions of this batch, for the "found SYNC training sequence" lines */
	size_t n_ev = S.text ? tb200_get_lock_events(S.ctx, NULL, 0) : 0, e = 0, ri = 0;
	struct tb200_lock_event *ev = n_ev ? malloc(n_ev * sizeof(*ev)) : NULL;      /* a noisy batch can hold any number */
	if (n_ev && !ev)
		shim_die("out of memory");
	if (n_ev)
		tb200_get_lock_events(S.ctx, ev, n_ev);
	for (long i = 0; i <= n; i++) {
		while (e < n_ev && ev[e].next_slot == (uint64_t)i)
			tb200_text_lock(ev[e++].offset);
		if (i == n)
			break;
		const struct tb200_slot *sl = &S.slots[i];
		const int kind = sl->flags & TB200_F_KIND_MASK;
		const int nblk = S.text ? tb200_text_slot(&S.txt, sl) : (kind == TB200_KIND_NONE ? 0 : kind == TB200_KIND_NDB_F ? 2 : 3);
		for (int b = 0; b < nblk && ri < nrec; b++) {
			const struct tb200_record *r = &S.rec[ri++];
			/* the reference takes the traffic branch before it prints anything about the block (:190-241 precede :258) */
			if (tms && tms->cur_burst.is_traffic) {
				/* the upper MAC has just seen an AACH that marks this slot as traffic (tetra_upper_mac.c:444-452):
				 * what tetra_lower_mac.c:190-241 does with the blocks that follow.  BLK1 of a normal burst counts as
				 * stolen and is still decoded; SCH/F and an un-stolen block 2 are NOT handed to the upper MAC: the
				 * reference writes their descrambled bits to <dumpdir>/traffic_*.out for an external codec.  The
				 * SCH/F dump is reproduced; the dump of a 216-bit block 2 is not (the reference reads 216
				 * uninitialised stack bytes into it, tetra_lower_mac.c:221-228: there is nothing to be exact against). */
				if (r->type1_len == 124 && r->blk_num == 1)
					tms->cur_burst.blk1_stolen = true;
				if (r->type1_len == 268) {
					dump_traffic_schf(tms, slot_raw_bits(r->slot_bit), r->scrambling_code);
					continue;
				}
				if (r->blk_num == 2 && !tms->cur_burst.blk2_stolen) {
					static int warned;
					if (!warned++)
						fprintf(stderr, "tetra_b200 shim: traffic in a second half slot: block withheld, its dump is not written\n");
					continue;
				}
			}
			if (S.text)
				tb200_text_block(&S.txt, sl, b, S.crc[i], r->type1);
			if (kind == TB200_KIND_SB && b == 0 && tms && tms->tcs) {
				/* after every SB1 the reference hands the cell in force to the crypto state
				 * (tetra_lower_mac.c:304-308); the cell is the one behind the slot's scrambling code
				 * (tetra_scramb.c:87-99: ((cc | mnc << 6 | mcc << 20) << 2) | 3, all zero before the first good SB1) */
				struct tetra_crypto_state *tcs = tms->tcs;
				const uint32_t code = sl->scrambling_code;
				const int cc = (code >> 2) & 0x3f, mnc = (code >> 8) & 0x3fff, mcc = (code >> 22) & 0x3ff;
				tcs->cc = cc;
				if ((tcs->mcc != mcc || tcs->mnc != mnc) && update_current_network)
					update_current_network(tcs, mcc, mnc);
			}
			deliver(r, priv);
		}
	}
	free(ev);
	if (S.text)
		fflush(stdout);
#endif
	{       /* keep the last SHIM_HIST bits for slots that start in this batch and complete in the next */
		const size_t have = S.n_hist + S.n_bits, keep = have < SHIM_HIST ? have : SHIM_HIST;
		memmove(S.hist + (SHIM_HIST - keep), S.hist + (SHIM_HIST - S.n_hist) + (have - keep), keep);
		S.n_hist = keep;
	}
	S.fed_before += S.n_bits;
	S.n_bits = 0;
	struct tb200_rx_carry c;
	tb200_get_carry(S.ctx, &c);
	if (S.trs) {                      /* mirror what callers could look at */
		S.trs->state = (enum rx_state)c.state;
		S.trs->bits_in_buf = c.bits_in_buf;
		S.trs->bitbuf_start_bitnum = (unsigned int)c.buf_start_bit;
		S.trs->next_frame_start_bitnum = (unsigned int)c.next_frame_start;
	}
#ifndef TETRA_B200_SHIM_L0_ONLY      /* at the L0-only depth the reference's lower MAC owns the time (tetra_lower_mac.c:302) */
	t_phy_state.time.tn = c.tn; t_phy_state.time.fn = c.fn; t_phy_state.time.mn = c.mn;
#endif
}

void tetra_b200_shim_flush(void)
{
	shim_run(1);
}

#ifdef TETRA_B200_SHIM_WRAP_READ
/* link with -Wl,--wrap=read: end of file on the descriptor that feeds the receiver ends the stream */
#include <unistd.h>
ssize_t __real_read(int fd, void *buf, size_t count);
static const void *g_last_read_buf;
static int g_last_read_fd = -1, g_stream_fd = -1;

ssize_t __wrap_read(int fd, void *buf, size_t count)
{
	const ssize_t n = __real_read(fd, buf, count);
	if (n > 0) {
		g_last_read_buf = buf;
		g_last_read_fd = fd;
	} else if (n == 0 && fd == g_stream_fd) {
		shim_run(1);
	}
	return n;
}
#endif

static void shim_init(unsigned int first_len)
{
	const char *e = getenv("TETRA_B200_BATCH_BITS");
	const char *d = getenv("TETRA_B200_DEVICE");
	S.batch_bits = e ? strtoull(e, NULL, 0) : (8u << 20);        /* 0: decode on every call */
	const char *fm = getenv("TETRA_B200_FLUSH_MS");
	S.flush_ms = fm ? atoi(fm) : 250;
	clock_gettime(CLOCK_MONOTONIC, &S.last_flush);
	if (tb200_create(&S.ctx, d ? atoi(d) : 0) != 0)
		shim_die("no usable CUDA device - this build has no CPU lower MAC");
	struct tb200_options opt;
	tb200_default_options(&opt);
	opt.chunk_bits = first_len;
	opt.output = TB200_OUT_UNPACKED;
	if (tb200_set_options(S.ctx, &opt) != 0)
		shim_die(tb200_last_error(S.ctx));
	S.chunk = first_len;
	S.cap_bits = (S.batch_bits < 4096 ? 4096 : S.batch_bits) + 4096;
	S.hist = tb200_host_alloc(S.cap_bits + SHIM_HIST);
	S.bits = S.hist ? S.hist + SHIM_HIST : NULL;
	S.max_slots = tb200_max_slots(S.cap_bits) + 16;
	S.slots = tb200_host_alloc(S.max_slots * sizeof(*S.slots));
	S.type1 = tb200_host_alloc(S.max_slots * TB200_TYPE1_STRIDE);
	S.rec = malloc(3 * S.max_slots * sizeof(*S.rec));
	S.crc = tb200_host_alloc(S.max_slots * sizeof(*S.crc));
	if (!S.bits || !S.slots || !S.type1 || !S.rec || !S.crc)
		shim_die("out of memory");
	const char *t = getenv("TETRA_B200_TEXT");
	S.text = !(t && t[0] == '0');
	tb200_set_crc_buffer(S.ctx, S.crc);
	atexit(tetra_b200_shim_flush);
}

int tetra_burst_sync_in(struct tetra_rx_state *trs, uint8_t *bits, unsigned int len)
{
	if (!S.ctx)
		shim_init(len);
	S.trs = trs;
#ifdef TETRA_B200_SHIM_WRAP_READ
	if (bits == g_last_read_buf)
		g_stream_fd = g_last_read_fd;         /* the descriptor whose data reaches the receiver */
#endif
	if (len > 296)
		shim_die("reads of more than 296 bits are not modelled");
	if (S.n_bits && len != S.chunk)
		shim_run(0);                          /* the length of the reads changes: a new run */
	if (len > S.cap_bits - S.n_bits)
		shim_run(0);
	S.chunk = len;
	memcpy(S.bits + S.n_bits, bits, len);
	S.n_bits += len;
	if (S.batch_bits == 0) {
		/* synchronous: decode now and return what the reference returns (tetra_burst_sync.c:66-106) */
		struct tb200_rx_carry before, after;
		tb200_get_carry(S.ctx, &before);
		shim_run(0);
		tb200_get_carry(S.ctx, &after);
		if (before.state == TB200_RX_UNLOCKED && after.state == TB200_RX_UNLOCKED && after.bits_in_buf >= 2 * TB200_BITS_PER_SLOT)
			return -1;                        /* the SYNC search over the buffer failed */
		if (before.state == TB200_RX_KNOW_FSTART && after.state == TB200_RX_KNOW_FSTART)
			return 0;                         /* the frame start is not in the buffer yet */
		return len;
	}
	if (S.n_bits >= S.batch_bits)
		shim_run(0);
	else if (S.flush_ms && ++S.calls_since_clock >= 16) {
		struct timespec now;
		S.calls_since_clock = 0;
		clock_gettime(CLOCK_MONOTONIC, &now);
		if ((now.tv_sec - S.last_flush.tv_sec) * 1000 + (now.tv_nsec - S.last_flush.tv_nsec) / 1000000 >= S.flush_ms)
			shim_run(0);
	}
	return len;
}
