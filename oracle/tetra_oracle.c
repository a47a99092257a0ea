/* TEST INFRASTRUCTURE ONLY - not part of the product path.
 *
 * CPU restatement ("port" oracle) of the osmo-tetra lower-MAC receive chain from
 * type-5 bits to type-1 bits, written from the behaviour of the reference
 * (file:line cited per function, relative to /root/reference/src), not copied
 * from it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * Pinning: every function here is checked in tests/test_oracle.py against the
 * reference's own code compiled in place (oracle/_ref/libtetra_ref.so) and
 * against the known-answer vectors the reference carries (crc_test.c:43-57,
 * tetra_punct_test tuples tetra_conv_enc.c:257-267, the conv_enc_test.c loop-back).
 * The Viterbi tie-break follows the published libosmocore algorithm (library not
 * available here, version un-pinned by the reference): on noisy input that part
 * is "parity unpinned" - see oracle/osmo_standin.c and DESIGN.md.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_records.h"
#include "../include/tetra_tie_rule.h"

/* Viterbi tie rule (include/tetra_tie_rule.h): compile-time default, run-time override for the tests */
static int g_tie = TETRA_VITERBI_TIE_DEFAULT;
void orc_set_tie(int tie) { g_tie = tie ? TETRA_TIE_KEEPS_HIGH_PRED : TETRA_TIE_KEEPS_LOW_PRED; }
int orc_get_tie(void) { return g_tie; }

/* ------------------------------------------------------------ constants -- */

/* training sequences, EN 300 392-2 9.4.4.3.2-4 as listed in phy/tetra_burst.c:59-70,
 * written here as bit strings (first bit on air first) */
static const char SEQ_Y[] = "11000001100111001110100111000001100111";  /* sync, 38 */
static const char SEQ_N[] = "1101000011101001110100";                  /* normal 1, 22 */
static const char SEQ_P[] = "0111101001000011011110";                  /* normal 2, 22 */
static const char SEQ_Q[] = "1011011100000110101101";                  /* normal 3, 22 */
static const char SEQ_X[] = "100111010000111010011101000011";          /* extended, 30 */

enum { TS_NORM_1 = 0, TS_NORM_2 = 1, TS_NORM_3 = 2, TS_SYNC = 3, TS_EXT = 4 };   /* tetra_burst.h:27-33 */
enum { T_SB1 = 0, T_SB2 = 1, T_NDB = 2, T_BBK = 3, T_SCH_HU = 4, T_SCH_F = 5 }; /* tetra_burst.h:9-16 */
enum { LC_UNKNOWN = 0, LC_SCH_F = 1, LC_AACH = 8, LC_BSCH = 10, LC_BNCH = 11 }; /* tetra_common.h:22-39 */

struct blk_param { uint16_t k, n, type1, a; };            /* tetra_lower_mac.c:55-102 */
static const struct blk_param BLK[6] = {
	[T_SB1]    = { 120,  80,  60,  11 },
	[T_SB2]    = { 216, 144, 124, 101 },
	[T_NDB]    = { 216, 144, 124, 101 },
	[T_BBK]    = {  30,  30,  14,   0 },
	[T_SCH_HU] = { 168, 112,  92,  13 },
	[T_SCH_F]  = { 432, 288, 268, 103 },
};

/* ---------------------------------------------------- leaf: scrambler -- */

/* tetra_scramb.c:34-50 - Fibonacci LFSR, taps 32 26 23 22 16 12 11 10 8 7 5 4 2 1
 * counted from the MSB side; written here as a parity over a tap mask. */
#define LFSR_TAPS 0xDB710641u

static inline unsigned lfsr_step(uint32_t *st)
{
	unsigned fb = __builtin_parity(*st & LFSR_TAPS);
	*st = (*st >> 1) | ((uint32_t)fb << 31);
	return fb;
}

void orc_scramb_get_bits(uint32_t init, uint8_t *out, int len)     /* tetra_scramb.c:66 */
{
	for (int i = 0; i < len; i++)
		out[i] = lfsr_step(&init);
}

void orc_scramb_bits(uint32_t init, uint8_t *io, int len)          /* tetra_scramb.c:77-85 */
{
	for (int i = 0; i < len; i++)
		io[i] ^= lfsr_step(&init);
}

uint32_t orc_scramb_get_init(unsigned mcc, unsigned mnc, unsigned cc) /* tetra_scramb.c:87-99 */
{
	uint32_t v = (cc & 0x3f) | ((mnc & 0x3fff) << 6) | ((uint32_t)(mcc & 0x3ff) << 20);
	return (v << 2) | 3;
}

/* -------------------------------------------------- leaf: interleaver -- */

void orc_deinterleave(unsigned K, unsigned a, const uint8_t *in, uint8_t *out)  /* tetra_interleave.c:51-60 */
{
	for (unsigned j = 0; j < K; j++)
		out[j] = in[(a * (j + 1)) % K];
}

void orc_interleave(unsigned K, unsigned a, const uint8_t *in, uint8_t *out)    /* tetra_interleave.c:41-49 */
{
	for (unsigned j = 0; j < K; j++)
		out[(a * (j + 1)) % K] = in[j];
}

/* ---------------------------------------------- leaf: RCPC 2/3 (de)puncture -- */

/* tetra_conv_enc.c:96,128-134,226-248: type-3 bit j lands on mother bit
 * 8*(j/3) + {0,1,4}[j%3]; everything else stays 0xff (erased). */
static inline unsigned mother_index_2_3(unsigned j)
{
	static const uint8_t P[3] = { 0, 1, 4 };
	return 8 * (j / 3) + P[j % 3];
}

void orc_depunct_2_3(const uint8_t *type3, int len, uint8_t *mother)
{
	for (int j = 0; j < len; j++)
		mother[mother_index_2_3(j)] = type3[j];
}

void orc_punct_2_3(const uint8_t *mother, int len, uint8_t *type3)              /* tetra_conv_enc.c:201-223 */
{
	for (int j = 0; j < len; j++)
		type3[j] = mother[mother_index_2_3(j)];
}

/* ------------------------------------------------ leaf: mother code TX -- */

/* tetra_conv_enc.c:43-74: rate 1/4, K=5; state bit0 = newest input (D^1). */
static inline unsigned mother_out(unsigned state, unsigned b)
{
	unsigned d1 = state & 1, d2 = (state >> 1) & 1, d3 = (state >> 2) & 1, d4 = (state >> 3) & 1;
	unsigned g1 = b ^ d1 ^ d4;
	unsigned g2 = b ^ d2 ^ d3 ^ d4;
	unsigned g3 = b ^ d1 ^ d2 ^ d4;
	unsigned g4 = b ^ d1 ^ d3 ^ d4;
	return (g1 << 3) | (g2 << 2) | (g3 << 1) | g4;   /* nibble, MSB = G1 as viterbi_cch.c:35-40 */
}

void orc_conv_encode(const uint8_t *in, int len, uint8_t *mother)
{
	unsigned st = 0;
	for (int i = 0; i < len; i++) {
		unsigned b = in[i] & 1, o = mother_out(st, b);
		mother[4 * i + 0] = (o >> 3) & 1;
		mother[4 * i + 1] = (o >> 2) & 1;
		mother[4 * i + 2] = (o >> 1) & 1;
		mother[4 * i + 3] = o & 1;
		st = ((st << 1) | b) & 15;
	}
}

/* ---------------------------------------------------- leaf: Viterbi -- */

/* viterbi.c:6-25 + viterbi_cch.c:58-66 + libosmocore osmo_conv_decode (absent,
 * see file header).  `mother` is the 4n-byte de-punctured stream of the wrapper:
 * 0 -> strong 0, 0xff -> erased, anything else -> strong 1.  Integer formulation
 * of the same maximum-likelihood search: cost = number of non-erased symbols that
 * disagree with the branch output; start in state 0; n data steps then 4 flush
 * steps whose symbols are all erased; trace back from state 0; on equal cost the
 * survivor is the predecessor whose oldest register bit is 0 (state t>>1 rather
 * than (t>>1)|8 in the table numbering of viterbi_cch.c:42-48) - or the other one
 * when the tie switch of include/tetra_tie_rule.h says so. */
int orc_viterbi(const uint8_t *mother, uint8_t *out, int n)
{
	enum { NS = 16 };
	const int steps = n + 4;
	uint32_t pm[NS], nm[NS];
	uint16_t *dec = malloc(sizeof(uint16_t) * steps);
	const uint32_t INF = 1u << 30;

	if (!dec)
		return -1;
	for (int s = 0; s < NS; s++)
		pm[s] = INF;
	pm[0] = 0;

	for (int t = 0; t < steps; t++) {
		unsigned known = 0, val = 0;    /* per-symbol masks, bit3 = G1 */
		if (t < n) {
			for (int g = 0; g < 4; g++) {
				uint8_t v = mother[4 * t + g];
				if (v != 0xff) {
					known |= 8u >> g;
					if (v != 0)
						val |= 8u >> g;
				}
			}
		}
		uint16_t d = 0;
		for (unsigned s = 0; s < NS; s++) {
			unsigned b = s & 1, p0 = s >> 1, p1 = p0 | 8;
			uint32_t c0 = pm[p0] + __builtin_popcount((mother_out(p0, b) ^ val) & known);
			uint32_t c1 = pm[p1] + __builtin_popcount((mother_out(p1, b) ^ val) & known);
			if (g_tie == TETRA_TIE_KEEPS_HIGH_PRED ? c1 <= c0 : c1 < c0) { nm[s] = c1; d |= 1u << s; }
			else         { nm[s] = c0; }
		}
		dec[t] = d;
		memcpy(pm, nm, sizeof(pm));
	}

	unsigned s = 0;
	for (int t = steps - 1; t >= 0; t--) {
		if (t < n)
			out[t] = s & 1;
		s = (s >> 1) | (((dec[t] >> s) & 1) << 3);
	}
	free(dec);
	return 0;
}

/* --------------------------------------------------------- leaf: CRC16 -- */

uint16_t orc_crc16(const uint8_t *bits, int len)      /* crc_simple.c:65-82,103-106 */
{
	uint16_t crc = 0xffff;
	for (int i = 0; i < len; i++) {
		unsigned top = ((crc >> 15) ^ bits[i]) & 1;
		crc = (uint16_t)(crc << 1);
		if (top)
			crc ^= 0x1021;
	}
	return crc;
}

/* ------------------------------------------------------ leaf: RM(30,14) -- */

/* tetra_rm3014.c:28-43,59-86: systematic code word = info<<16 | parity; the parity
 * columns of the generator matrix, one 16-bit row per info bit (MSB first). */
static const uint16_t RM_PARITY[14] = {
	0x9b60, 0x2de0, 0xfc20, 0xe03c, 0x983a, 0x5436, 0x2c2e,
	0xffdf, 0x8339, 0x42b5, 0x21ad, 0x1273, 0x096b, 0x04e7,
};

uint32_t orc_rm3014_compute(uint16_t info)
{
	uint32_t v = 0;
	for (int i = 0; i < 14; i++)
		if ((info >> (13 - i)) & 1)
			v ^= ((uint32_t)1 << (29 - i)) | RM_PARITY[i];
	return v;
}

/* Maximum-likelihood decoding of a received 30-bit word (bit 29 = first bit on air, as tetra_rm3014_compute
 * returns it) by brute force over all 2^14 code words: the nearest code word wins, among equally near ones the
 * one whose error pattern word ^ cw is numerically smallest.  The reference has no RM decoder to compare with
 * (tetra_rm3014.c:92-96 returns inp >> 16, tetra_lower_mac.c:268-274 "FIXME: RM3014-decode"): the code itself is
 * pinned through orc_rm3014_compute == tetra_rm3014_compute, the decoder by this exhaustive search.
 * Returns the distance; *info = the 14 information bits of the winner (as tetra_rm3014_decode's *out). */
int orc_rm3014_decode_ml(uint32_t word, uint16_t *info)
{
	static uint32_t cw[1 << 14];
	static int have;
	if (!have) {
		for (uint32_t i = 0; i < (1u << 14); i++)
			cw[i] = orc_rm3014_compute((uint16_t)i);
		have = 1;
	}
	word &= 0x3fffffffu;
	int best_d = 31;
	uint32_t best_e = 0xffffffffu, best_i = 0;
	for (uint32_t i = 0; i < (1u << 14); i++) {
		const uint32_t e = word ^ cw[i];
		const int d = __builtin_popcount(e);
		if (d < best_d || (d == best_d && e < best_e)) {
			best_d = d; best_e = e; best_i = i;
		}
	}
	*info = (uint16_t)best_i;
	return best_d;
}

/* ------------------------------------------------------ leaf: TDMA time -- */

struct orc_time { uint32_t tn, fn, mn; };

void orc_time_add_slot(struct orc_time *t)              /* tetra_tdma.c:27-53,75-79 with tn_count = 1 */
{
	t->tn += 1;
	if (t->tn > 4) { t->fn += t->tn / 4; t->tn %= 4; }
	if (t->fn > 18) { t->mn += t->fn / 18; t->fn %= 18; }
	if (t->mn > 60) t->mn %= 60;
}

static unsigned bits_to_uint(const uint8_t *b, int len)  /* tetra_common.c:31-39 */
{
	unsigned v = 0;
	while (len--)
		v = (v << 1) | (*b++ & 1);
	return v;
}

/* -------------------------------------------- training sequence search -- */

static uint32_t prefix22(const char *s)
{
	uint32_t v = 0;
	for (int i = 0; i < 22; i++)
		v = (v << 1) | (uint32_t)(s[i] - '0');
	return v;
}

static int seq_matches(const uint8_t *p, const char *s, int len)
{
	for (int i = 0; i < len; i++)
		if (p[i] != (uint8_t)(s[i] - '0'))
			return 0;
	return 1;
}

/* tetra_burst.c:269-339.  First position (ascending) at which an enabled sequence
 * matches exactly, gated by a 22-bit rolling pre-filter that is primed with
 * in[0..19] and then shifts in in[k+21] at position k - so in[20] never enters it
 * and positions 0..20 see a distorted window.  Reads in[] up to
 * index end+20 like the reference does; callers keep that readable. */
int orc_find_train_seq(const uint8_t *in, unsigned end, uint32_t mask, unsigned *offset)
{
	const uint32_t pre[5] = { prefix22(SEQ_Y), prefix22(SEQ_N), prefix22(SEQ_P),
				  prefix22(SEQ_Q), prefix22(SEQ_X) };
	uint32_t filt = 0;
	for (int i = 0; i < 20; i++)
		filt = (filt << 1) | in[i];
	for (unsigned k = 0; k < end; k++) {
		filt = ((filt << 1) | in[k + 21]) & 0x3fffff;
		if (filt != pre[0] && filt != pre[1] && filt != pre[2] && filt != pre[3] && filt != pre[4])
			continue;
		unsigned remain = end - k;
		if ((mask & (1u << TS_SYNC)) && remain >= 38 && seq_matches(in + k, SEQ_Y, 38)) { *offset = k; return TS_SYNC; }
		if ((mask & (1u << TS_NORM_1)) && remain >= 22 && seq_matches(in + k, SEQ_N, 22)) { *offset = k; return TS_NORM_1; }
		if ((mask & (1u << TS_NORM_2)) && remain >= 22 && seq_matches(in + k, SEQ_P, 22)) { *offset = k; return TS_NORM_2; }
		if ((mask & (1u << TS_NORM_3)) && remain >= 22 && seq_matches(in + k, SEQ_Q, 22)) { *offset = k; return TS_NORM_3; }
		if ((mask & (1u << TS_EXT)) && remain >= 30 && seq_matches(in + k, SEQ_X, 30)) { *offset = k; return TS_EXT; }
	}
	return -1;
}

/* ------------------------------------------------------- receiver state -- */

enum { RX_UNLOCKED = 0, RX_KNOW_FSTART = 1, RX_LOCKED = 2 };   /* tetra_burst_sync.h:6-10 */

struct orc_rx {
	/* tetra_burst_sync.h:12-20; bitbuf has 64 bytes of slack because the search
	 * pre-filter reads up to 21 entries past the window */
	int state;
	unsigned bits_in_buf;
	uint8_t bitbuf[4096 + 64];
	uint32_t buf_start_bit;
	uint32_t next_frame_start;
	/* t_phy_state (tetra_burst_sync.c:34) and _tcd (tetra_lower_mac.c:104-113) */
	struct orc_time phy_time;
	struct orc_time cell_time;
	unsigned mcc, mnc, cc;
	uint32_t scramb_init;
	uint32_t call_index;
	/* recorder */
	struct tb_record *rec; size_t nrec, caprec;
	struct tb_fsm_event *ev; size_t nev, capev;
	int recording;
};

static struct orc_rx *G;

void orc_reset(void)
{
	if (G) { free(G->rec); free(G->ev); free(G); }
	G = calloc(1, sizeof(*G));
	G->recording = 1;
}

void orc_set_recording(int on) { if (!G) orc_reset(); G->recording = on; }

static struct tb_record *new_record(struct orc_rx *rx)
{
	if (rx->nrec == rx->caprec) {
		rx->caprec = rx->caprec ? rx->caprec * 2 : 4096;
		rx->rec = realloc(rx->rec, rx->caprec * sizeof(*rx->rec));
	}
	struct tb_record *r = &rx->rec[rx->nrec++];
	memset(r, 0, sizeof(*r));
	return r;
}

/* --------------------------------------------------------- lower MAC -- */

static int is_bnch(const struct orc_time *t)            /* tetra_lower_mac.c:122-127 */
{
	return t->fn == 18 && t->tn == 4 - ((t->mn + 3) % 4);
}

/* tetra_lower_mac.c:143-357 (the arithmetic; traffic-dump side path excluded, the
 * recorder stands where upper_mac_prim_recv is called, one record per block) */
static void orc_tp_sap_rx(struct orc_rx *rx, int type, int blk_num, const uint8_t *bits)
{
	const struct blk_param *bp = &BLK[type];
	uint8_t type4[512], type3[512], type2[512], mother[4 * 512];
	unsigned lchan = LC_UNKNOWN, crc_ok = 0;
	uint32_t code;

	rx->cell_time = rx->phy_time;                                   /* :167 */
	if (type == T_SB2 && is_bnch(&rx->cell_time))                   /* :170-173 */
		lchan = LC_BNCH;

	memcpy(type4, bits, bp->k);                                     /* :179-186 */
	code = (type == T_SB1) ? 3 : rx->scramb_init;
	orc_scramb_bits(code, type4, bp->k);

	if (bp->a) {                                                    /* :243-256 */
		orc_deinterleave(bp->k, bp->a, type4, type3);
		memset(mother, 0xff, 4 * bp->n);
		orc_depunct_2_3(type3, bp->k, mother);
		orc_viterbi(mother, type2, bp->n);
		crc_ok = orc_crc16(type2, bp->type1 + 16) == 0x1d0f;        /* :258-267 */
	} else {                                                        /* BBK :268-274 */
		crc_ok = 1;
		memcpy(type2, type4, bp->n);
	}

	switch (type) {                                                 /* :282-324 */
	case T_SB1:
		if (crc_ok) {
			rx->cc = bits_to_uint(type2 + 4, 6);
			rx->cell_time.tn = bits_to_uint(type2 + 10, 2) + 1;
			rx->cell_time.fn = bits_to_uint(type2 + 12, 5);
			rx->cell_time.mn = bits_to_uint(type2 + 17, 6);
			rx->mcc = bits_to_uint(type2 + 31, 10);
			rx->mnc = bits_to_uint(type2 + 41, 14);
			rx->scramb_init = orc_scramb_get_init(rx->mcc, rx->mnc, rx->cc);
		}
		rx->phy_time = rx->cell_time;
		lchan = LC_BSCH;
		break;
	case T_BBK:   lchan = LC_AACH;  break;
	case T_SCH_F: lchan = LC_SCH_F; break;
	default: break;
	}

	if (rx->recording) {
		struct tb_record *r = new_record(rx);
		r->slot_bit = rx->buf_start_bit;
		r->lchan = lchan;
		r->crc_ok = crc_ok;
		r->blk_num = blk_num;
		r->tn = rx->cell_time.tn; r->fn = rx->cell_time.fn; r->mn = rx->cell_time.mn;
		r->type1_len = bp->type1;
		r->scrambling_code = code;
		memcpy(r->type1, type2, bp->type1);
	}
}

void orc_tp_sap(int type, int blk_num, const uint8_t *bits)
{
	if (!G) orc_reset();
	orc_tp_sap_rx(G, type, blk_num, bits);
}

/* ------------------------------------------------------- burst slicing -- */

/* tetra_burst.c:31-47,341-379: sub-block offsets inside the 510-bit burst and the
 * fixed delivery order SB1,BBK,SB2 / BBK,BLK1,BLK2 / BBK,SCH-F */
static void orc_burst_rx(struct orc_rx *rx, const uint8_t *burst, int ts)
{
	uint8_t bbk[30], schf[432];
	switch (ts) {
	case TS_SYNC:
		orc_tp_sap_rx(rx, T_SB1, 1, burst + 94);
		orc_tp_sap_rx(rx, T_BBK, 0, burst + 252);
		orc_tp_sap_rx(rx, T_SB2, 2, burst + 282);
		break;
	case TS_NORM_2:
		memcpy(bbk, burst + 230, 14); memcpy(bbk + 14, burst + 266, 16);
		orc_tp_sap_rx(rx, T_BBK, 0, bbk);
		orc_tp_sap_rx(rx, T_NDB, 1, burst + 14);
		orc_tp_sap_rx(rx, T_NDB, 2, burst + 282);
		break;
	case TS_NORM_1:
		memcpy(bbk, burst + 230, 14); memcpy(bbk + 14, burst + 266, 16);
		memcpy(schf, burst + 14, 216); memcpy(schf + 216, burst + 282, 216);
		orc_tp_sap_rx(rx, T_BBK, 0, bbk);
		orc_tp_sap_rx(rx, T_SCH_F, 0, schf);
		break;
	default:
		break;
	}
}

/* ------------------------------------------------------------ lock FSM -- */

static void log_event(struct orc_rx *rx, unsigned window, uint32_t mask, int rc, unsigned off)
{
	if (!rx->recording)
		return;
	if (rx->nev == rx->capev) {
		rx->capev = rx->capev ? rx->capev * 2 : 4096;
		rx->ev = realloc(rx->ev, rx->capev * sizeof(*rx->ev));
	}
	struct tb_fsm_event *e = &rx->ev[rx->nev++];
	e->call_index = rx->call_index;
	e->buf_start_bit = rx->buf_start_bit;
	e->window = window;
	e->mask = mask;
	e->rc = rc;
	e->offset = rc >= 0 ? off : 0;
}

/* tetra_burst_sync.c:38-154 */
static int orc_sync_in(struct orc_rx *rx, const uint8_t *bits, unsigned len)
{
	unsigned off = 0;
	int rc;

	unsigned space = 4096 - rx->bits_in_buf;                         /* :38-51 */
	if (space < len) {
		unsigned delta = len - space;
		memmove(rx->bitbuf, rx->bitbuf + delta, rx->bits_in_buf - delta);
		rx->bits_in_buf -= delta;
		rx->buf_start_bit += delta;
	}
	memcpy(rx->bitbuf + rx->bits_in_buf, bits, len);                 /* :62-64 */
	rx->bits_in_buf += len;

	if (rx->state == RX_UNLOCKED) {                                  /* :67-90 */
		if (rx->bits_in_buf < 1020)
			return len;
		rc = orc_find_train_seq(rx->bitbuf, rx->bits_in_buf, 1u << TS_SYNC, &off);
		log_event(rx, rx->bits_in_buf, 1u << TS_SYNC, rc, off);
		if (rc < 0)
			return rc;
		rx->state = RX_KNOW_FSTART;
		rx->next_frame_start = rx->buf_start_bit + off + 296;
		return len;
	}
	if (rx->state == RX_KNOW_FSTART) {                               /* :91-106 */
		if (rx->buf_start_bit + rx->bits_in_buf < rx->next_frame_start)
			return 0;
		int shift = (int)(rx->next_frame_start - rx->buf_start_bit);
		int remaining = (int)rx->bits_in_buf - shift;
		memmove(rx->bitbuf, rx->bitbuf + shift, remaining);
		rx->bits_in_buf = remaining;
		rx->buf_start_bit += shift;
		rx->next_frame_start += 510;
		rx->state = RX_LOCKED;                                   /* falls through */
	}
	if (rx->bits_in_buf < 510)                                       /* :107-150 */
		return len;
	orc_time_add_slot(&rx->phy_time);
	const uint32_t mask = (1u << TS_NORM_1) | (1u << TS_NORM_2) | (1u << TS_SYNC);
	rc = orc_find_train_seq(rx->bitbuf, rx->bits_in_buf, mask, &off);
	log_event(rx, rx->bits_in_buf, mask, rc, off);
	if (rc == TS_SYNC) {
		if (off == 214) orc_burst_rx(rx, rx->bitbuf, rc);
		else rx->state = RX_UNLOCKED;
	} else if (rc == TS_NORM_1 || rc == TS_NORM_2) {
		if (off == 244) orc_burst_rx(rx, rx->bitbuf, rc);
	} else {
		rx->state = RX_UNLOCKED;
	}
	rx->bits_in_buf -= 510;
	memmove(rx->bitbuf, rx->bitbuf + 510, rx->bits_in_buf);
	rx->buf_start_bit += 510;
	rx->next_frame_start += 510;
	return len;
}

/* feed like tetra-rx.c:82-95 (chunk = 64 there) */
long orc_feed(const uint8_t *bits, size_t n, unsigned chunk)
{
	if (!G) orc_reset();
	if (chunk == 0 || chunk > 4096)
		return -1;
	for (size_t pos = 0; pos < n; pos += chunk) {
		unsigned len = (n - pos < chunk) ? (unsigned)(n - pos) : chunk;
		G->call_index++;
		orc_sync_in(G, bits + pos, len);
	}
	return (long)G->nrec;
}

size_t orc_num_records(void) { return G ? G->nrec : 0; }
const struct tb_record *orc_records(void) { return G ? G->rec : NULL; }
size_t orc_num_events(void) { return G ? G->nev : 0; }
const struct tb_fsm_event *orc_events(void) { return G ? G->ev : NULL; }
int orc_rx_state(void) { return G ? G->state : -1; }
uint32_t orc_cell_scramb_init(void) { return G ? G->scramb_init : 0; }
void orc_set_cell(uint32_t scramb_init) { if (!G) orc_reset(); G->scramb_init = scramb_init; }
void orc_set_time(uint32_t tn, uint32_t fn, uint32_t mn)
{
	if (!G) orc_reset();
	G->phy_time.tn = tn; G->phy_time.fn = fn; G->phy_time.mn = mn;
}
void orc_get_time(uint32_t *tn, uint32_t *fn, uint32_t *mn)
{
	if (!G) orc_reset();
	*tn = G->phy_time.tn; *fn = G->phy_time.fn; *mn = G->phy_time.mn;
}

/* FNV-1a over all records, a cheap "checksum of checksums" for large runs */
uint64_t orc_records_digest(const struct tb_record *r, size_t n)
{
	uint64_t h = 0xcbf29ce484222325ull;
	const uint8_t *p = (const uint8_t *)r;
	for (size_t i = 0; i < n * sizeof(*r); i++) {
		h ^= p[i];
		h *= 0x100000001b3ull;
	}
	return h;
}

/* ============================================================ generator ==
 * Synthetic downlink stream generator (TX side), the CPU twin of the CUDA
 * generator in osmo-tetra_b200/csrc/tetra_gen.cu: same counter-based RNG, same
 * burst schedule, bit-identical output, so any shard of a large GPU-generated
 * stream can be regenerated here for parity.  TX chain follows
 * conv_enc_test.c:88-156,198-305 (type-1 -> CRC -> tail -> mother code -> 2/3
 * puncture -> interleave -> scramble -> burst, burst layouts tetra_burst.c:169-260),
 * except that SB2 and BBK ARE scrambled with the cell code (the reference test
 * generator forgets to, conv_enc_test.c:284,296) and phase-adjustment bits are 0.
 */

struct orc_gen_cfg {
	uint64_t seed;
	uint32_t sb_period;      /* burst k is a SYNC burst iff k % sb_period == 0 (0: only k < lead_sb) */
	uint32_t lead_sb;        /* bursts [0, lead_sb) are always SYNC bursts */
	uint32_t ndb2_per_256;   /* of the remaining bursts, this many in 256 carry two half-slot blocks (p training seq) */
	uint32_t ber_per_65536;  /* i.i.d. flip probability on payload bits, in 1/65536 */
	uint32_t random_cell;    /* 0: every SB announces MCC 262 / MNC 42 / CC 1; 1: random per SB */
	uint32_t lead_in_bits;   /* random bits before burst 0 (not generated by orc_gen_burst) */
};

static inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

/* word `w` of stream `lane` of burst `k` */
static inline uint64_t gen_rng(uint64_t seed, uint64_t k, uint32_t lane, uint32_t w)
{
	return mix64(mix64(seed + 0x9e3779b97f4a7c15ull * (k + 1)) + ((uint64_t)lane << 32) + w);
}

enum { LANE_KIND = 1, LANE_CELL = 2, LANE_SB1 = 3, LANE_BLK = 4, LANE_BBK = 5, LANE_NOISE = 6, LANE_LEADIN = 7 };

int orc_gen_kind(const struct orc_gen_cfg *c, uint64_t k)   /* returns TS_SYNC / TS_NORM_1 / TS_NORM_2 */
{
	if (k < c->lead_sb || (c->sb_period && k % c->sb_period == 0))
		return TS_SYNC;
	if ((gen_rng(c->seed, k, LANE_KIND, 0) & 255) < c->ndb2_per_256)
		return TS_NORM_2;
	return TS_NORM_1;
}

static uint64_t last_sb_before(const struct orc_gen_cfg *c, uint64_t k)   /* index of the latest SB at or before k */
{
	uint64_t best = 0;
	if (c->lead_sb)
		best = (k < c->lead_sb) ? k : c->lead_sb - 1;
	if (c->sb_period) {
		uint64_t p = k - k % c->sb_period;
		if (p > best || !c->lead_sb) best = p;
	}
	return best;
}

void orc_gen_cell(const struct orc_gen_cfg *c, uint64_t sb_index, unsigned *mcc, unsigned *mnc, unsigned *cc)
{
	if (!c->random_cell) { *mcc = 262; *mnc = 42; *cc = 1; return; }
	uint64_t r = gen_rng(c->seed, sb_index, LANE_CELL, 0);
	*mcc = r & 0x3ff; *mnc = (r >> 10) & 0x3fff; *cc = (r >> 24) & 0x3f;
}

static void put_bits(uint8_t *dst, uint64_t v, int n)    /* MSB first */
{
	for (int i = 0; i < n; i++)
		dst[i] = (v >> (n - 1 - i)) & 1;
}

static void fill_random(uint8_t *dst, int n, uint64_t seed, uint64_t k, uint32_t lane)
{
	for (int i = 0; i < n; i++) {
		uint64_t r = gen_rng(seed, k, lane, i >> 6);
		dst[i] = (r >> (i & 63)) & 1;
	}
}

/* type-1 -> type-5 for one coded block */
static void encode_block(const uint8_t *type1, int type, uint32_t code, uint8_t *type5)
{
	const struct blk_param *bp = &BLK[type];
	uint8_t type2[512], mother[4 * 512], type3[512];
	memset(type2, 0, sizeof(type2));
	memcpy(type2, type1, bp->type1);
	uint16_t crc = ~orc_crc16(type2, bp->type1);
	put_bits(type2 + bp->type1, crc, 16);               /* + 4 zero tail bits */
	orc_conv_encode(type2, bp->n, mother);
	orc_punct_2_3(mother, bp->k, type3);
	orc_interleave(bp->k, bp->a, type3, type5);
	orc_scramb_bits(code, type5, bp->k);
}

static void add_noise(uint8_t *bits, int n, const struct orc_gen_cfg *c, uint64_t k, uint32_t sub)
{
	if (!c->ber_per_65536)
		return;
	for (int i = 0; i < n; i++) {
		uint64_t r = gen_rng(c->seed, k, LANE_NOISE, (sub << 8) + (i >> 2));
		if (((r >> (16 * (i & 3))) & 0xffff) < c->ber_per_65536)
			bits[i] ^= 1;
	}
}

/* returns the type-1 payloads too (may be NULL) so tests can check recovery */
void orc_gen_burst(const struct orc_gen_cfg *c, uint64_t k, uint8_t *burst)
{
	int kind = orc_gen_kind(c, k);
	unsigned mcc, mnc, cc;
	uint64_t sb = last_sb_before(c, k);
	orc_gen_cell(c, sb, &mcc, &mnc, &cc);
	/* before any SB has been seen the receiver has code 0; the generator only
	 * produces streams that start with an SB, so `code` is always the SB's */
	uint32_t code = orc_scramb_get_init(mcc, mnc, cc);
	uint8_t t1[272], blk1[432], blk2[216], bb[30];

	/* AACH: 14 bits, first two (header) zero so the upper MAC never flags traffic */
	uint64_t rb = gen_rng(c->seed, k, LANE_BBK, 0);
	uint32_t cw = orc_rm3014_compute((uint16_t)(rb & 0x0fff));
	put_bits(bb, cw, 30);
	orc_scramb_bits(code, bb, 30);

	memset(burst, 0, 510);
	/* q11..q22 | 2 phase adj (0) */
	for (int i = 0; i < 12; i++) burst[i] = SEQ_Q[10 + i] - '0';

	if (kind == TS_SYNC) {
		/* SYNC PDU, Table 21.73 field order as in testpdu.c:41-57 */
		uint8_t *p = t1;
		uint64_t slot = k;
		put_bits(p, 0, 4); p += 4;                         /* system code */
		put_bits(p, cc, 6); p += 6;
		put_bits(p, slot % 4, 2); p += 2;                  /* tn - 1 */
		put_bits(p, (slot / 4) % 18 + 1, 5); p += 5;       /* fn */
		put_bits(p, (slot / 72) % 60 + 1, 6); p += 6;      /* mn */
		put_bits(p, 0, 8); p += 8;                         /* sharing, reserved frames, dtx, f18 ext, reserved */
		put_bits(p, mcc, 10); p += 10;
		put_bits(p, mnc, 14); p += 14;
		put_bits(p, 0, 5); p += 5;
		encode_block(t1, T_SB1, 3, blk1);
		add_noise(blk1, 120, c, k, 0);
		fill_random(t1, 124, c->seed, k, LANE_BLK);
		encode_block(t1, T_SB2, code, blk2);
		add_noise(blk2, 216, c, k, 1);
		/* 9.4.4.2.6: q11-22, hc, f1-80, sb 120, y 38, bb 30, bkn2 216, hd, q1-10 */
		for (int i = 0; i < 8; i++) { burst[14 + i] = 1; burst[14 + 72 + i] = 1; }
		memcpy(burst + 94, blk1, 120);
		for (int i = 0; i < 38; i++) burst[214 + i] = SEQ_Y[i] - '0';
		memcpy(burst + 252, bb, 30);
		memcpy(burst + 282, blk2, 216);
	} else {
		if (kind == TS_NORM_1) {
			fill_random(t1, 268, c->seed, k, LANE_BLK);
			encode_block(t1, T_SCH_F, code, blk1);
			add_noise(blk1, 432, c, k, 0);
			memcpy(burst + 14, blk1, 216);
			memcpy(burst + 282, blk1 + 216, 216);
		} else {
			fill_random(t1, 124, c->seed, k, LANE_BLK);
			encode_block(t1, T_NDB, code, blk1);
			add_noise(blk1, 216, c, k, 0);
			fill_random(t1, 124, c->seed, k, LANE_SB1);
			encode_block(t1, T_NDB, code, blk2);
			add_noise(blk2, 216, c, k, 1);
			memcpy(burst + 14, blk1, 216);
			memcpy(burst + 282, blk2, 216);
		}
		/* 9.4.4.2.5: q11-22, ha, bkn1 216, bb 14, n/p 22, bb 16, bkn2 216, hb, q1-10 */
		memcpy(burst + 230, bb, 14);
		const char *ts = (kind == TS_NORM_2) ? SEQ_P : SEQ_N;
		for (int i = 0; i < 22; i++) burst[244 + i] = ts[i] - '0';
		memcpy(burst + 266, bb + 14, 16);
	}
	for (int i = 0; i < 10; i++) burst[500 + i] = SEQ_Q[i] - '0';
}

/* whole stream: lead_in random bits, then bursts [k0, k0+n) back to back */
void orc_gen_stream(const struct orc_gen_cfg *c, uint64_t k0, uint64_t n, uint8_t *out, int with_lead_in)
{
	if (with_lead_in) {
		for (uint32_t i = 0; i < c->lead_in_bits; i++) {
			uint64_t r = gen_rng(c->seed, 0, LANE_LEADIN, i >> 6);
			out[i] = (r >> (i & 63)) & 1;
		}
		out += c->lead_in_bits;
	}
	for (uint64_t i = 0; i < n; i++)
		orc_gen_burst(c, k0 + i, out + 510 * i);
}

/* ------------------------------------------------------------ symbol slicer --
 * float_to_bits.c, the program between the demodulator and tetra-rx (receiver1:8):
 * one float per pi/4-DQPSK symbol (phase step in units of pi/4) -> two unpacked bits.
 * process_sym_fl (float_to_bits.c:33-50): strict comparisons, > 2 -> 3, > 0 -> 1, < -2 -> -3, else -1
 * (so 2.0 -> 1, 0.0 -> -1, -2.0 -> -1, NaN -> -1); sym_int2bits (:52-76): 3 -> 0,1   1 -> 0,0
 * -3 -> 1,1   -1 -> 1,0.  The pseudo-AFC (-a, :138-147) follows below. */
void orc_float_to_bits(const float *sym, size_t n, uint8_t *bits)
{
	for (size_t i = 0; i < n; i++) {
		const float fl = sym[i];
		int s;
		if (fl > 2) s = 3;
		else if (fl > 0) s = 1;
		else if (fl < -2) s = -3;
		else s = -1;
		switch (s) {
		case -3: bits[2 * i] = 1; bits[2 * i + 1] = 1; break;
		case 1:  bits[2 * i] = 0; bits[2 * i + 1] = 0; break;
		case 3:  bits[2 * i] = 0; bits[2 * i + 1] = 1; break;
		default: bits[2 * i] = 1; bits[2 * i + 1] = 0; break;
		}
	}
}

/* float_to_bits -a (float_to_bits.c:140-147): a one-pole tracker of the symbols' mean, subtracted before slicing.
 *     if (-5 < fl < 5) filter = filter * (1.0 - filter_val) + (fl - filter_goal) * filter_val;     slice(fl - filter)
 * filter, filter_val, filter_goal are floats; `1.0` is a double, so the first product and the sum are formed in double
 * and rounded to float once per symbol, the second product in float (C's usual arithmetic conversions; no fused
 * multiply-add: plain x86-64 code).  *filter is the tracker's state: zero at the start of the program. */
void orc_float_to_bits_afc(const float *sym, size_t n, float filter_val, float filter_goal, float *filter, uint8_t *bits)
{
	volatile float f = *filter;            /* volatile: every step is rounded to float, whatever the compiler would like */
	for (size_t i = 0; i < n; i++) {
		const float fl = sym[i];
		if ((fl > -5.0) && (fl < 5.0)) {
			volatile double a = (double)f * (1.0 - (double)filter_val);
			volatile float b = (fl - filter_goal) * filter_val;
			f = (float)(a + (double)b);
		}
		orc_float_to_bits(&(float){ fl - f }, 1, bits + 2 * i);
	}
	*filter = f;
}

/* ----------------------------------------------------------- GSMTAP framing --
 * What tetra_gsmtap_makemsg (tetra_gsmtap.c:31-63) builds for every CRC-good primitive the upper MAC sees
 * (rx_tmv_unitdata_ind, tetra_upper_mac.c:480-488: timeslot = tn - 1, sub-slot / level / snr 0):
 * struct gsmtap_hdr of libosmocore's gsmtap.h (absent here; restated from the published GSMTAP v2 format:
 * version 2, hdr_len 4 words, type 5 = TETRA_I1, timeslot, arfcn 0, signal 0, snr 0, frame number in network
 * order, sub_type, antenna 0, sub_slot, reserved 0; msgb_alloc zeroes what is not set) + osmo_ubit2pbit of the
 * type-1 bits (eight per byte, first bit in the MSB, zero padding).  frame number = (hn*60 + mn)*18 + fn
 * (tetra_tdma.c:96-99) with hn = 0: the lower MAC never sets it.  sub_type = lchan2gsmtap[lchan]
 * (tetra_gsmtap.c:18-27): SCH_F 5, AACH 2, BSCH 1, BNCH 6, 0 for TETRA_LC_UNKNOWN. */
size_t orc_gsmtap_frames(const struct tb_record *rec, size_t n, uint8_t *out, size_t cap, size_t *n_frames)
{
	size_t len = 0, frames = 0;
	for (size_t i = 0; i < n; i++) {
		const struct tb_record *r = &rec[i];
		if (!r->crc_ok)
			continue;
		const unsigned nb = (r->type1_len + 7u) / 8u;
		if (out && len + 16 + nb <= cap) {
			uint8_t *h = out + len;
			const uint32_t fnum = (uint32_t)r->mn * 18u + r->fn;
			unsigned sub = 0;
			switch (r->lchan) {          /* enum tetra_log_chan, tetra_common.h:22-39 */
			case 1: sub = 5; break;      /* SCH/F */
			case 8: sub = 2; break;      /* AACH */
			case 10: sub = 1; break;     /* BSCH */
			case 11: sub = 6; break;     /* BNCH */
			default: break;
			}
			memset(h, 0, 16 + nb);
			h[0] = 2; h[1] = 4; h[2] = 5; h[3] = (uint8_t)(r->tn - 1);
			h[8] = fnum >> 24; h[9] = fnum >> 16; h[10] = fnum >> 8; h[11] = fnum;
			h[12] = (uint8_t)sub;
			for (unsigned b = 0; b < r->type1_len; b++)
				h[16 + b / 8] |= (uint8_t)((r->type1[b] & 1) << (7 - b % 8));
		}
		len += 16 + nb;
		frames++;
	}
	if (n_frames)
		*n_frames = frames;
	return len;
}
