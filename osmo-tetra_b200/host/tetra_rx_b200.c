/*
 * tetra_rx_b200.c - file-in / text-out receiver on top of libtetra_b200.so: the counterpart of
 * `tetra-rx <bits-file>` (src/tetra-rx.c:40-103) for the part of its output that the PHY and the lower MAC
 * produce.  It reads a stream file, decodes it on the GPU in one batch call and prints, in stream order,
 * exactly the lines the reference prints from phy/tetra_burst_sync.c and lower_mac/tetra_lower_mac.c:
 *
 *   found SYNC training sequence in bit #N            tetra_burst_sync.c:79
 *   (empty line) BURST                                tetra_burst_sync.c:114-116
 *   BNCH FOLLOWS                                      tetra_lower_mac.c:170-173
 *   CRC COMP: 0x1d0f OK / CRC COMP: 0x.... WRONG      tetra_lower_mac.c:258-267
 *   <SB1|SB2|NDB|SCH/F> mn/fn/tn/sn type1: <bits>     tetra_lower_mac.c:264-265
 *   TMB-SAP SYNC CC ... TN ... FN ... MN ... MCC ... MNC ...   tetra_lower_mac.c:283-289
 *
 * and on stderr the "#### ..." complaints of tetra_burst_sync.c:126,136,139.  That is the text the
 * reference's regression harness counts (tetra-rx-tests.sh:56 greps "^CRC COMP: 0x.+ OK").  The upper MAC
 * (everything tetra-rx prints from upper_mac_prim_recv() on) is not part of this tool: link the real one
 * through tetra_shim.c for that.
 *
 * usage: tetra-rx-b200 [-f bytes|packed|f32] [-c read_size_bits] [-g cuda_device] [-G gpus] [-p out.pcap] <stream-file>
 *   bytes   one bit per byte, what tetra-rx reads (default)      packed  eight bits per byte
 *   f32     float32 symbols, what float_to_bits reads
 *   -p      also write the GSMTAP frames of the CRC-good blocks (tb200_gsmtap_pack: what the reference sends
 *           to UDP port 4729 through tetra_gsmtap_sendmsg, tetra_upper_mac.c:480-488) as a pcap file of
 *           IPv4/UDP packets, one per frame, 127.0.0.1 -> 127.0.0.1:4729, for wireshark
 *   -G N    decode the ONE stream on N GPUs (devices g .. g+N-1): one host thread per GPU, the C sharding driver
 *           tb200_dist_rx_stream (NCCL inside the library: bit-packed shards scattered from the first GPU, one
 *           all-gather for the cell state, lock losses handled); the text is the same, line for line
 */
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tetra_b200.h"
#include "tetra_text.h"

/* ---- one stream on several GPUs: a host thread per GPU around tb200_dist_rx_stream ---- */
struct mg_shared {
	int world, device0;
	unsigned int chunk;
	uint32_t input;
	const uint8_t *data; size_t data_bytes; uint64_t n_bits;
	uint8_t id[TB200_DIST_ID_BYTES];
	struct tb200_slot *slots; uint8_t *type1; uint32_t *crc; uint64_t cap;      /* host, global slot order */
	struct tb200_lock_event *ev; size_t n_ev;
	long n_total;
	pthread_mutex_t mu;
	int failed;
};
struct mg_rank { struct mg_shared *sh; int rank; };

static void *mg_thread(void *arg)
{
	struct mg_rank *me = arg;
	struct mg_shared *sh = me->sh;
	tb200_ctx *rx = NULL;
	tb200_dist *dd = NULL;
	const char *why = NULL;
#define MG_FAIL(msg) do { why = (msg); goto out; } while (0)
	if (tb200_create(&rx, sh->device0 + me->rank) != 0) MG_FAIL("no usable CUDA device");
	struct tb200_options opt;
	tb200_default_options(&opt);
	opt.chunk_bits = sh->chunk; opt.output = TB200_OUT_UNPACKED; opt.input = sh->input;
	if (tb200_set_options(rx, &opt) != 0) MG_FAIL(tb200_last_error(rx));
	if (tb200_dist_create(&dd, rx, me->rank, sh->world, sh->id) != 0) MG_FAIL(tb200_last_error(rx));
	uint8_t *d_bits = NULL;
	if (me->rank == 0) {
		d_bits = tb200_dev_alloc(rx, sh->data_bytes + 64);
		if (!d_bits || tb200_dev_copy(rx, d_bits, sh->data, sh->data_bytes, 1) != 0) MG_FAIL("stream does not fit the first GPU");
	}
	const uint64_t cap = tb200_dist_max_local_slots(sh->n_bits, sh->world);
	struct tb200_slot *d_slots = tb200_dev_alloc(rx, cap * sizeof(*d_slots));
	uint8_t *d_type1 = tb200_dev_alloc(rx, cap * TB200_TYPE1_STRIDE);
	uint32_t *d_crc = tb200_dev_alloc(rx, cap * sizeof(uint32_t));
	if (!d_slots || !d_type1 || !d_crc) MG_FAIL("out of device memory");
	tb200_set_crc_buffer(rx, d_crc);
	struct tb200_dist_run runs[1024];
	uint32_t n_runs = 0;
	const uint32_t mode = TB200_DIST_SCATTER | (sh->input == TB200_IN_BYTES ? TB200_DIST_PACK : 0);
	const long n = tb200_dist_rx_stream(dd, d_bits, sh->n_bits, mode, d_slots, d_type1, NULL, cap, runs, 1024, &n_runs);
	if (n < 0) MG_FAIL(tb200_dist_last_error(dd));
	for (uint32_t i = 0; i < n_runs; i++) {          /* results into the arrays of the whole stream, in delivery order */
		const uint64_t g = runs[i].global_slot, l = runs[i].local_slot, c = runs[i].n_slots;
		if (g + c > sh->cap) MG_FAIL("more slots than the stream can hold");
		if (tb200_dev_copy(rx, sh->slots + g, d_slots + l, c * sizeof(*d_slots), 0) ||
		    tb200_dev_copy(rx, sh->type1 + g * TB200_TYPE1_STRIDE, d_type1 + l * TB200_TYPE1_STRIDE, c * TB200_TYPE1_STRIDE, 0) ||
		    tb200_dev_copy(rx, sh->crc + g, d_crc + l, c * sizeof(uint32_t), 0))
			MG_FAIL(tb200_last_error(rx));
	}
	pthread_mutex_lock(&sh->mu);
	sh->n_total += n;
	if (me->rank == 0) {
		sh->n_ev = tb200_get_lock_events(rx, NULL, 0);
		sh->ev = calloc(sh->n_ev ? sh->n_ev : 1, sizeof(*sh->ev));
		tb200_get_lock_events(rx, sh->ev, sh->n_ev);
	}
	pthread_mutex_unlock(&sh->mu);
out:
	if (why) {
		fprintf(stderr, "tetra-rx-b200: GPU %d: %s\n", sh->device0 + me->rank, why);
		pthread_mutex_lock(&sh->mu); sh->failed = 1; pthread_mutex_unlock(&sh->mu);
		if (!dd) exit(1);          /* the other ranks would wait for this one inside the communicator set-up */
	}
	if (dd) tb200_dist_destroy(dd);
	if (rx) tb200_destroy(rx);
	return NULL;
#undef MG_FAIL
}

static void put16be(uint8_t *p, unsigned v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }

/* one GSMTAP frame as a raw-IPv4 pcap record (LINKTYPE_RAW): pcap record header, IPv4, UDP, payload */
static void pcap_packet(FILE *f, const uint8_t *frame, unsigned len, uint64_t usec)
{
	const uint32_t rec[4] = { (uint32_t)(usec / 1000000u), (uint32_t)(usec % 1000000u), 28 + len, 28 + len };
	uint8_t h[28] = { 0x45, 0 };
	put16be(h + 2, 28 + len);                    /* total length */
	h[8] = 64; h[9] = 17;                        /* TTL, UDP */
	h[12] = h[16] = 127; h[15] = h[19] = 1;      /* 127.0.0.1 -> 127.0.0.1 */
	uint32_t sum = 0;
	for (int i = 0; i < 20; i += 2)
		sum += (uint32_t)h[i] << 8 | h[i + 1];
	while (sum >> 16)
		sum = (sum & 0xffff) + (sum >> 16);
	put16be(h + 10, ~sum & 0xffff);
	put16be(h + 20, 4729); put16be(h + 22, 4729);    /* GSMTAP_UDP_PORT */
	put16be(h + 24, 8 + len);                    /* UDP length; checksum 0 = none */
	fwrite(rec, sizeof(rec), 1, f);
	fwrite(h, sizeof(h), 1, f);
	fwrite(frame, len, 1, f);
}

int main(int argc, char **argv)
{
	const char *fmt = "bytes", *path = NULL, *pcap = NULL;
	unsigned int chunk = 64;
	int device = 0, gpus = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-f") && i + 1 < argc) fmt = argv[++i];
		else if (!strcmp(argv[i], "-c") && i + 1 < argc) chunk = (unsigned int)atoi(argv[++i]);
		else if (!strcmp(argv[i], "-g") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-G") && i + 1 < argc) gpus = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-p") && i + 1 < argc) pcap = argv[++i];
		else path = argv[i];
	}
	if (!path) {
		fprintf(stderr, "usage: %s [-f bytes|packed|f32] [-c read_size_bits] [-g cuda_device] [-G gpus] [-p out.pcap] <stream-file>\n", argv[0]);
		return 2;
	}
	uint32_t input = !strcmp(fmt, "packed") ? TB200_IN_PACKED : !strcmp(fmt, "f32") ? TB200_IN_F32SYM : TB200_IN_BYTES;

	FILE *f = fopen(path, "rb");
	if (!f) { perror("open"); return 1; }
	fseek(f, 0, SEEK_END);
	const long fsize = ftell(f);
	fseek(f, 0, SEEK_SET);
	uint8_t *data = tb200_host_alloc((size_t)fsize + 64);
	if (!data || fread(data, 1, (size_t)fsize, f) != (size_t)fsize) { fprintf(stderr, "read failed\n"); return 1; }
	fclose(f);
	const uint64_t n_bits = input == TB200_IN_BYTES ? (uint64_t)fsize : input == TB200_IN_PACKED ? 8ull * fsize : (uint64_t)fsize / 4 * 2;

	tb200_ctx *rx = NULL;
	struct tb200_slot *slots; uint8_t *type1; uint32_t *crc;
	struct tb200_lock_event *ev; size_t n_ev;
	long n;
	const uint64_t cap = tb200_max_slots(n_bits) + 16;
	if (gpus >= 1) {
		if (pcap || input == TB200_IN_F32SYM) { fprintf(stderr, "-G goes with -f bytes|packed and without -p\n"); return 2; }
		struct mg_shared sh;
		memset(&sh, 0, sizeof(sh));
		sh.world = gpus; sh.device0 = device; sh.chunk = chunk; sh.input = input;
		sh.data = data; sh.data_bytes = (size_t)fsize; sh.n_bits = n_bits; sh.cap = cap;
		sh.slots = calloc(cap, sizeof(*sh.slots)); sh.type1 = calloc(cap, TB200_TYPE1_STRIDE); sh.crc = calloc(cap, sizeof(uint32_t));
		pthread_mutex_init(&sh.mu, NULL);
		if (!sh.slots || !sh.type1 || !sh.crc) { fprintf(stderr, "out of memory\n"); return 1; }
		/* stdout carries the receiver's text and nothing else: while the GPUs work (NCCL prints its version line to
		 * stdout when NCCL_DEBUG asks for it) descriptor 1 points at stderr */
		fflush(stdout);
		const int saved_stdout = dup(1);
		dup2(2, 1);
		if (tb200_dist_get_id(sh.id) != 0) { fprintf(stderr, "NCCL (libnccl.so.2) is not available\n"); return 1; }
		pthread_t th[64];
		struct mg_rank rk[64];
		if (gpus > 64) { fprintf(stderr, "at most 64 GPUs\n"); return 2; }
		for (int r = 0; r < gpus; r++) { rk[r].sh = &sh; rk[r].rank = r; pthread_create(&th[r], NULL, mg_thread, &rk[r]); }
		for (int r = 0; r < gpus; r++) pthread_join(th[r], NULL);
		fflush(stdout);
		dup2(saved_stdout, 1);
		close(saved_stdout);
		if (sh.failed) return 1;
		slots = sh.slots; type1 = sh.type1; crc = sh.crc; ev = sh.ev; n_ev = sh.n_ev; n = sh.n_total;
	} else {
	if (tb200_create(&rx, device) != 0) { fprintf(stderr, "no usable CUDA device (there is no CPU lower MAC in this build)\n"); return 1; }
	struct tb200_options opt;
	tb200_default_options(&opt);
	opt.chunk_bits = chunk; opt.output = TB200_OUT_UNPACKED | (pcap ? TB200_OUT_PACKED : 0); opt.input = input;
	if (tb200_set_options(rx, &opt) != 0) { fprintf(stderr, "%s\n", tb200_last_error(rx)); return 1; }
	slots = tb200_host_alloc(cap * sizeof(*slots));
	type1 = tb200_host_alloc(cap * TB200_TYPE1_STRIDE);
	crc = tb200_host_alloc(cap * sizeof(*crc));
	uint32_t *packed = pcap ? tb200_host_alloc(cap * TB200_TYPE1_WORDS * sizeof(uint32_t)) : NULL;
	if (!slots || !type1 || !crc || (pcap && !packed)) { fprintf(stderr, "out of memory\n"); return 1; }
	tb200_set_crc_buffer(rx, crc);
	n = tb200_rx_stream_host(rx, data, n_bits, TB200_FRESH | TB200_FINAL, slots, type1, packed, cap);
	if (n < 0) { fprintf(stderr, "%s\n", tb200_last_error(rx)); return 1; }
	if (pcap) {
		/* frames of the whole file in one device pass, then one pcap record per frame; the capture time of a
		 * frame is its slot's place in the stream (85/6 ms per slot) */
		uint64_t n_frames = 0;
		const long long need = tb200_gsmtap_pack(rx, slots, packed, (uint64_t)n, NULL, 0, NULL, &n_frames, 0);
		uint8_t *frames = need >= 0 ? malloc((size_t)need + 2) : NULL;
		uint64_t *off = malloc(((size_t)n + 1) * sizeof(*off));
		if (need < 0 || !frames || !off ||
		    tb200_gsmtap_pack(rx, slots, packed, (uint64_t)n, frames, (uint64_t)need, off, &n_frames, 0) != need) {
			fprintf(stderr, "GSMTAP framing failed: %s\n", tb200_last_error(rx));
			return 1;
		}
		FILE *pf = fopen(pcap, "wb");
		if (!pf) { perror("open pcap"); return 1; }
		const uint32_t gh[6] = { 0xa1b2c3d4u, 2u | 4u << 16, 0, 0, 65535, 101 /* LINKTYPE_RAW */ };
		fwrite(gh, sizeof(gh), 1, pf);
		for (long i = 0; i < n; i++) {
			uint64_t p = off[i];
			while (p < off[i + 1]) {
				const unsigned len = TB200_GSMTAP_FRAME_LEN(frames[p + 12]);
				pcap_packet(pf, frames + p, len, (uint64_t)i * 85000u / 6u);
				p += len;
			}
		}
		fclose(pf);
		free(frames); free(off);
	}
	n_ev = tb200_get_lock_events(rx, NULL, 0);
	ev = calloc(n_ev ? n_ev : 1, sizeof(*ev));
	tb200_get_lock_events(rx, ev, n_ev);
	}

	struct tb200_text txt = { 0, 0, 0 };             /* t_phy_state.time: zero at start (tetra_burst_sync.c:34) */
	size_t e = 0;
	for (long i = 0; i <= n; i++) {
		while (e < n_ev && ev[e].next_slot == (uint64_t)i)
			tb200_text_lock(ev[e++].offset);
		if (i == n)
			break;
		const struct tb200_slot *s = &slots[i];
		const uint8_t *t1 = type1 + (size_t)i * TB200_TYPE1_STRIDE;
		const int kind = s->flags & TB200_F_KIND_MASK;
		const int nblk = tb200_text_slot(&txt, s);
		for (int b = 0; b < nblk; b++)
			tb200_text_block(&txt, s, b, crc[i], t1 + tb200_text_block_offset(kind, b));
		if (nblk && (txt.tn != (s->time & 7u) || txt.fn != ((s->time >> 3) & 31u) || txt.mn != ((s->time >> 8) & 63u))) {
			fprintf(stderr, "internal error: slot %ld time %s differs from the device's\n", i, tb200_text_time(&txt));
			return 3;
		}
	}
	fflush(stdout);
	if (rx) tb200_destroy(rx);
	return 0;
}
