/*
 * tetra_b200.h - C ABI of the B200-native TETRA lower-MAC receive chain
 *                (type-5 bits -> type-1 bits), libtetra_b200.so.
 *
 * Every entry point states the reference interface it replaces (file:line relative
 * to osmo-tetra/src).  Plain pointers and sizes only; no C++/torch types.
 *
 * Bit streams use the reference's ABI: one bit per byte, values 0/1
 * (tetra-rx.c:82-95 reads such a file 64 bytes at a time).
 */
#ifndef TETRA_B200_H
#define TETRA_B200_H

#include <stdint.h>
#include <stddef.h>
#include "tetra_tie_rule.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_BITS_PER_SLOT   510     /* TETRA_BITS_PER_TS, tetra_common.h:19 */
#define TB200_TYPE1_STRIDE    288     /* bytes of unpacked type-1 bits kept per slot (282 used max) */
#define TB200_TYPE1_WORDS     9       /* 32-bit words of packed type-1 bits per slot */

/* enum tetra_train_seq (phy/tetra_burst.h:27-33) */
enum { TB200_TRAIN_NORM_1 = 0, TB200_TRAIN_NORM_2 = 1, TB200_TRAIN_NORM_3 = 2,
       TB200_TRAIN_SYNC = 3, TB200_TRAIN_EXT = 4 };

/* enum tp_sap_data_type (phy/tetra_burst.h:9-16) */
enum { TB200_T_SB1 = 0, TB200_T_SB2 = 1, TB200_T_NDB = 2, TB200_T_BBK = 3,
       TB200_T_SCH_HU = 4, TB200_T_SCH_F = 5 };

/* enum rx_state (phy/tetra_burst_sync.h:6-10) */
enum { TB200_RX_UNLOCKED = 0, TB200_RX_KNOW_FSTART = 1, TB200_RX_LOCKED = 2 };

/* what a slot turned out to be (decides the layout of its type-1 bits) */
enum { TB200_KIND_NONE = 0,    /* nothing delivered to the lower MAC (dropped / lock lost) */
       TB200_KIND_SB = 1,      /* SYNC burst:   SB1[0,60)  BBK[60,74)  SB2[74,198)          tetra_burst.c:347-353 */
       TB200_KIND_NDB_F = 2,   /* normal burst: BBK[0,14)  SCH/F[14,282)                    tetra_burst.c:363-373 */
       TB200_KIND_NDB_2 = 3 }; /* normal burst: BBK[0,14)  BLK1[14,138)  BLK2[138,262)      tetra_burst.c:354-362 */

/* tb200_slot.flags */
#define TB200_F_KIND_MASK   0x03
#define TB200_F_CRC_A       0x04   /* CRC of SB1 / SCH-F / BLK1 good (tetra_lower_mac.c:258-267) */
#define TB200_F_CRC_B       0x08   /* CRC of SB2 / BLK2 good */
#define TB200_F_BNCH        0x10   /* SB2 carries BNCH (is_bnch, tetra_lower_mac.c:122-127,170-173) */
#define TB200_F_UNLOCK      0x20   /* this slot made the receiver lose lock (tetra_burst_sync.c:127,140) */

/* One per slot the LOCKED receiver consumed (tetra_burst_sync.c:107-150), in stream order. */
struct tb200_slot {
	uint32_t slot_bit;         /* absolute bit number of the slot start (uint32, wraps like tetra_burst_sync.h:16) */
	uint32_t scrambling_code;  /* cell scrambling code used for every block but SB1 (tetra_lower_mac.c:185) */
	uint16_t find_off;         /* offset returned by the training-sequence search (valid if find_rc >= 0) */
	uint16_t window;           /* bits the search was allowed to look at (bits_in_buf, tetra_burst_sync.c:117) */
	uint16_t time;             /* tdma time on the slot's primitives: tn | fn << 3 | mn << 8 */
	int8_t   find_rc;          /* enum tetra_train_seq or -1 (tetra_find_train_seq, tetra_burst.c:269-339) */
	uint8_t  flags;            /* TB200_F_* */
};

/* Receiver state that survives from one call to the next: the reference keeps it in
 * struct tetra_rx_state (tetra_burst_sync.h:12-20), t_phy_state (tetra_burst_sync.c:34)
 * and the static _tcd (tetra_lower_mac.c:104-113). */
struct tb200_rx_carry {
	uint64_t stream_bits;      /* bits fed so far */
	uint64_t buf_start_bit;    /* bitbuf_start_bitnum */
	uint64_t next_frame_start; /* next_frame_start_bitnum */
	uint64_t calls;            /* tetra_burst_sync_in() calls modelled so far */
	uint32_t state;            /* TB200_RX_* */
	uint32_t bits_in_buf;
	uint32_t scramb_init;      /* tcd->scramb_init */
	uint16_t mcc, mnc;
	uint8_t  colour_code;
	uint8_t  tn, fn, mn;       /* t_phy_state.time */
};

struct tb200_stats {             /* since the last TB200_FRESH call, except kernel_launches */
	uint64_t slots;            /* slots consumed while LOCKED */
	uint64_t bursts_decoded;   /* slots handed to the lower MAC (kind != NONE) */
	uint64_t blocks;           /* TMV-SAP primitives produced */
	uint64_t crc_ok_blocks;    /* of those with a CRC, how many were good */
	uint64_t lock_losses;
	uint64_t lock_acquisitions;
	uint64_t kernel_launches;  /* CUDA kernels launched by the last call */
};

typedef struct tb200_ctx tb200_ctx;

/* ---- life cycle ---------------------------------------------------------- */

/* One context per GPU/process; `device` is the CUDA ordinal.  Returns 0 or a negative
 * TB200_E_* code.  Fails (never falls back to the CPU) when no CUDA device is usable.
 * A context is a single receiver (like the reference's globals t_phy_state / _tcd): calls on one context
 * must not overlap; several contexts may be used from several threads, also on the same GPU
 * (tools/overlap_probe.py: two receivers on one GPU give +10 % aggregate throughput). */
int  tb200_create(tb200_ctx **out, int device);
void tb200_destroy(tb200_ctx *ctx);
const char *tb200_last_error(const tb200_ctx *ctx);
const char *tb200_version(void);

#define TB200_E_CUDA     (-1)
#define TB200_E_ARG      (-2)
#define TB200_E_NOMEM    (-3)
#define TB200_E_STATE    (-4)

/* ---- options -------------------------------------------------------------- */

#define TB200_OUT_UNPACKED  1   /* type-1 bits one per byte (reference ABI, msg->l1h) */
#define TB200_OUT_PACKED    2   /* type-1 bits 32 per word, LSB first */

#define TB200_VITERBI_WARP  0   /* one warp per burst, warp-shuffle ACS butterflies */
#define TB200_VITERBI_LANE  1   /* one lane per coded block, registers-only ACS */

/* how the bit stream handed to tb200_rx_stream_* is encoded */
#define TB200_IN_BYTES   0   /* one bit per byte: the file format tetra-rx reads (tetra-rx.c:82-95), float_to_bits' output */
#define TB200_IN_PACKED  1   /* eight bits per byte, stream bit i = byte i>>3, bit i&7 (4-byte aligned buffer); a stream fed over several
                              * tb200_rx_stream_host calls continues on 128-bit boundaries (n_bits % 128 == 0 except for the last call) */
#define TB200_IN_F32SYM  2   /* one float32 per symbol, the demodulator output float_to_bits reads (float_to_bits.c:128-164):
                              * two hard bits per symbol, sliced on the device exactly like process_sym_fl + sym_int2bits
                              * (float_to_bits.c:33-72), with its pseudo-AFC (-a) when options.afc is set; n_bits = 2 * symbols */

struct tb200_options {
	uint32_t chunk_bits;        /* read() size the caller models; tetra-rx.c:83 uses 64. 1..296 */
	uint32_t output;            /* TB200_OUT_* bit mask */
	uint32_t viterbi;           /* TB200_VITERBI_* */
	uint32_t pipeline_slots;    /* slots per pipelined piece in the host-buffer path (0 = default) */
	uint32_t profile;           /* 1: bracket every kernel with CUDA events on its stream (tb200_get_timing) */
	uint32_t input;             /* TB200_IN_*; the packed / symbol formats need the lane kernels */
	uint32_t serial_passes;     /* 1: pass 1 (search, SB1, scans) of a piece on the same CUDA stream as the decode passes, so that no two
	                             * kernels of the receiver ever run side by side (clean per-kernel timings); 0 (default): pass 1 of
	                             * piece i+1 on its own stream under the decode pass of piece i */
	uint32_t afc;               /* TB200_IN_F32SYM only: 1 = slice like `float_to_bits -a`, i.e. with its pseudo-AFC (float_to_bits.c:140-147;
	                             * what the reference's live pipeline runs, src/receiver1udp:62), exactly (tb200_float_to_bits) */
	float    afc_filter_val;    /* float_to_bits -f (default 0.0001) */
	float    afc_filter_goal;   /* float_to_bits -F (default 0) */
	uint32_t viterbi_tie;       /* survivor on equal path metrics (tetra_tie_rule.h): 0 = predecessor s>>1 (libosmocore's
	                             * osmo_conv_decode as restated, viterbi_cch.c:58-66), 1 = predecessor (s>>1)|8.  The default is
	                             * TETRA_VITERBI_TIE_DEFAULT, the compile-time switch the CPU oracle shares */
	uint32_t host_pack_threads; /* tb200_rx_stream_host with TB200_IN_BYTES input: 0 (default) = the caller's bytes cross PCIe as they are
	                             * (510 bytes per burst); n > 0 = n host threads of the library pack them to eight bits per byte first
	                             * (into pinned staging, piece by piece, under the GPU's work on the previous piece), so that 64 bytes per
	                             * burst cross the bus.  Pays when the host has cores and memory bandwidth to spare: ~8 threads match
	                             * PCIe 5 x16, 16 threads on one socket pack ~120 GB/s (tools/microbench/host_pack.c).  Results are
	                             * identical either way */
};
void tb200_default_options(struct tb200_options *opt);
int  tb200_set_options(tb200_ctx *ctx, const struct tb200_options *opt);

/* ---- batch receive: replaces the tetra-rx read loop + tetra_burst_sync_in -----
 *
 * Both calls behave like
 *     for each chunk_bits-sized piece of bits[0..n_bits): tetra_burst_sync_in(trs, piece, len)
 * (tetra-rx.c:82-95, phy/tetra_burst_sync.c:54-154) followed by the whole lower MAC
 * (lower_mac/tetra_lower_mac.c:143-357) for every delivered block, except that results
 * are returned as arrays instead of upper_mac_prim_recv() callbacks.
 *
 * flags: TB200_FRESH starts from a zeroed receiver (what tetra-rx.c:48-54 allocates),
 * otherwise the call continues the stream where the previous call on this ctx stopped.
 * TB200_FINAL says the stream ends with this call, so a trailing piece shorter than
 * chunk_bits is fed as one last short read (read() at EOF, tetra-rx.c:85-94); without it
 * the bits after the last full chunk boundary are kept for the next call.
 */
#define TB200_FRESH  1u
#define TB200_FINAL  2u

/* Device-resident: `d_bits` is a device pointer; results stay on the device.
 * d_slots[max_slots], d_type1 (max_slots * TB200_TYPE1_STRIDE bytes, 16-byte aligned, may
 * be NULL), d_type1_packed (max_slots * TB200_TYPE1_WORDS words, may be NULL) are device
 * buffers owned by the caller.  Returns the number of slots written (>= 0) or TB200_E_*.
 * The call waits for the device on entry (the caller's buffers may still be in flight on other
 * streams) and all its device work is complete when it returns.  A stream continues across calls like with
 * tb200_rx_stream_host (flags = 0): the bits the receiver may still look at are kept on the device, a continuing call
 * copies them and the first 32 Kbit of the new buffer into one piece, works through that and goes on in place in the
 * caller's buffer.  Host and device calls may alternate inside a stream. */
long tb200_rx_stream_dev(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint32_t flags,
                         struct tb200_slot *d_slots, uint8_t *d_type1, uint32_t *d_type1_packed,
                         uint64_t max_slots);

/* Host buffers in and out (pinned memory makes it faster, pageable works):
 * H2D copies, kernels and D2H copies are pipelined over CUDA streams inside the call. */
long tb200_rx_stream_host(tb200_ctx *ctx, const uint8_t *bits, uint64_t n_bits, uint32_t flags,
                          struct tb200_slot *slots, uint8_t *type1, uint32_t *type1_packed,
                          uint64_t max_slots);

/* Upper bound on the slots a stream of n_bits can yield. */
uint64_t tb200_max_slots(uint64_t n_bits);

/* Optional side outputs, for callers that reproduce the reference's text output (tetra-rx prints them).
 *
 * CRC registers: after tb200_set_crc_buffer(ctx, buf) every rx call also writes, per slot and at the same index as
 * the slot record, the CRC-16 register values the reference prints as "CRC COMP: 0x%04x" (tetra_lower_mac.c:258-267):
 * bits 0-15 = SB1 / SCH-F / BLK1, bits 16-31 = SB2 / BLK2 (0x1d0f = good, 0 = no such block).  `buf` lives where the
 * slot records live (device memory for tb200_rx_stream_dev, host memory for tb200_rx_stream_host) and holds
 * max_slots words; NULL switches the output off.
 *
 * Lock acquisitions of the last rx call, i.e. every "found SYNC training sequence in bit #%u"
 * (tetra_burst_sync.c:79): `offset` is that number (relative to bitbuf[0]), `call` the tetra_burst_sync_in()
 * call that found it, `next_slot` the index of the first slot record delivered after it. */
struct tb200_lock_event {
	uint64_t next_slot;
	uint64_t call;
	uint32_t offset;
	uint32_t pad;
};
int    tb200_set_crc_buffer(tb200_ctx *ctx, uint32_t *crc);
/* RM(30,14) decoding of the AACH, which the reference leaves out: its lower MAC hands the first 14 descrambled bits of
 * the broadcast block up unchecked (tetra_lower_mac.c:268-274 "FIXME: RM3014-decode", tetra_rm3014.c:92-96 returns
 * inp >> 16).  The type-1 bits of the slot stay what the reference delivers; this is an opt-in side output at the same
 * index as the slot record (residency as for tb200_set_crc_buffer; NULL switches it off): the result of
 * tb200_rm3014_decode() for the slot's 30 descrambled broadcast bits, 0xffffffff for slots handed to nobody. */
int    tb200_set_aach_buffer(tb200_ctx *ctx, uint32_t *aach);
size_t tb200_get_lock_events(const tb200_ctx *ctx, struct tb200_lock_event *ev, size_t max_events);

int tb200_get_carry(const tb200_ctx *ctx, struct tb200_rx_carry *out);
int tb200_get_stats(const tb200_ctx *ctx, struct tb200_stats *out);

/* Device time of the kernels of the last rx call, measured with CUDA events recorded on the
 * stream the kernels were launched on (needs options.profile = 1, which serialises nothing
 * but adds event records).  *_ms are sums over all launches of that kernel in the call. */
struct tb200_timing {
	float total_ms;             /* first launch to last launch of the call, compute stream */
	float classify_ms;          /* search kernel (load + pack + training-sequence search + slot records) + SB1 pass */
	float scan_ms;              /* k_scan_blocks + k_scan_prefix + k_finalize_carry */
	float decode_ms;            /* k_decode_*: descramble + de-interleave + Viterbi + CRC + output */
	uint32_t launches_classify, launches_scan, launches_decode;
	uint32_t pieces;
	uint64_t slots;             /* slots those launches covered */
	float leaf_ms;              /* last tb200_descramble_deinterleave kernel (device pointers) */
	float search_ms;            /* the training-sequence search kernel alone (classify_ms minus the SB1 pass) */
	float prepare_ms;           /* split decode pass: k_lane_prepare (cell state, descramble, de-interleave); 0 for the fused kernel */
	float trellis_ms;           /* split decode pass: k_lane_trellis (ACS loop + trace back [+ finish]); decode_ms - prepare_ms - trellis_ms = k_lane_finish */
};
int tb200_get_timing(const tb200_ctx *ctx, struct tb200_timing *out);

/* Integer-ALU peak of this GPU (32-bit add / min / compare results per second), measured with a
 * register-only kernel; the denominator for the Viterbi add-compare-select roofline. */
double tb200_measure_int_peak(tb200_ctx *ctx);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) */
void *tb200_host_alloc(size_t bytes);
void  tb200_host_free(void *p);

/* ---- records: what crosses TMV-SAP ------------------------------------------
 * Expands slots + type-1 bits into one record per tp_sap_udata_ind() call, in the
 * reference's call order (tetra_burst.c:347-373), with the fields of
 * struct tmv_unitdata_param (tetra_prim.h:25-33).  Host-side helper. */
struct tb200_record {
	uint32_t slot_bit;
	uint8_t  lchan;            /* enum tetra_log_chan, tetra_common.h:22-39 */
	uint8_t  crc_ok;
	uint8_t  blk_num;
	uint8_t  tn, fn, mn;
	uint16_t type1_len;
	uint32_t scrambling_code;
	uint8_t  type1[272];       /* one bit per byte */
};
/* returns records written; `records` may be NULL to count only */
size_t tb200_expand_records(const struct tb200_slot *slots, const uint8_t *type1, size_t n_slots,
                            struct tb200_record *records, size_t max_records);

/* ---- leaf operators (batched, device side; used by the stage-parity tests) ----
 * Each mirrors one reference function over `n` independent blocks laid out back to back. */

/* tetra_scramb_bits + block_deinterleave (tetra_scramb.c:77, tetra_interleave.c:51):
 * type-5 bytes -> type-3 bytes, n blocks of K bits each with interleaver constant a and
 * per-block scrambling code d_codes[i].  Host or device pointers (is_device). */
int tb200_descramble_deinterleave(tb200_ctx *ctx, const uint8_t *type5, uint8_t *type3,
                                  const uint32_t *codes, uint64_t n, uint32_t K, uint32_t a,
                                  int is_device);

/* whole tp_sap_udata_ind arithmetic for n blocks of one type (tetra_lower_mac.c:143-280), every row of
 * tetra_blk_param[] (tetra_lower_mac.c:55-102): SB1, SB2, NDB, SCH/F, the uplink SCH/HU (168 -> 92 bits, a = 13)
 * and BBK (first 14 descrambled bits, crc_ok = 1).  type-5 bytes in; type-1 bytes (type1_bits each) and crc_ok
 * flags out.  Host pointers. */
int tb200_decode_blocks(tb200_ctx *ctx, int blk_type, const uint8_t *type5, const uint32_t *codes,
                        uint64_t n, uint8_t *type1, uint8_t *crc_ok);

/* tetra_find_train_seq (tetra_burst.c:269-339) for n independent windows: window i is
 * bits[starts[i] .. starts[i]+lens[i]); mask as in the reference. Host pointers. */
int tb200_find_train_seq(tb200_ctx *ctx, const uint8_t *bits, uint64_t n_bits,
                         const uint64_t *starts, const uint32_t *lens, uint64_t n, uint32_t mask,
                         int32_t *rc, uint32_t *offset);

/* tetra_rcpc_depunct (tetra_conv_enc.c:226-248) for n blocks laid out back to back, any of the reference's seven
 * puncturers (enum tetra_rcpc_puncturer, tetra_conv_enc.h:16-24: 0 = 2/3, 1 = 1/3, 2 = 292/432, 3 = 148/432,
 * 4 = 112/168, 5 = 72/162, 6 = 38/80): `len` type-3 bytes per block in, `mother_len` mother-code bytes per block out,
 * filled with 0xff (erased) first like the caller of the reference does (tetra_lower_mac.c:249).  The receive chain
 * itself only ever uses 2/3 and folds it into the trellis loop; this is the function on its own, as a parity
 * checkpoint.  TB200_E_ARG for an unknown puncturer (-EINVAL in the reference). */
int tb200_rcpc_depunct(tb200_ctx *ctx, int puncturer, const uint8_t *type3, uint32_t len, uint64_t n,
                       uint8_t *mother, uint32_t mother_len, int is_device);

/* viterbi_dec_sb1_wrapper (viterbi.c:6-25: byte 0 -> strong 0, 0xff -> erased, anything else -> strong 1, four erased
 * flush steps) + conv_cch_decode (viterbi_cch.c:58-66, the K = 5 rate-1/4 mother code with all four generators) for n
 * independent blocks of sym_count type-2 bits: 4 * sym_count mother-code bytes per block in, sym_count bytes (0/1) out.
 * With tb200_rcpc_depunct in front of it every RCPC rate of tetra_conv_enc.c:128-198 (1/3, 292/432, 148/432 and the
 * speech rates next to 2/3) decodes end to end; the receive chain itself only ever needs 2/3 and folds de-puncturing
 * and decoding into one packed trellis loop.  Tie rule: options.viterbi_tie.  Host or device pointers. */
int tb200_viterbi_decode(tb200_ctx *ctx, const uint8_t *mother, uint64_t n, uint32_t sym_count, uint8_t *out, int is_device);

/* Maximum-likelihood decoding of the (30,14) Reed-Muller code of the AACH (generator: tetra_rm3014.c:28-43) for n
 * received 30-bit words laid out like tetra_rm3014_compute's result (bit 29 = first bit on air, information = word >> 16,
 * the argument of the reference's tetra_rm3014_decode(inp, out), tetra_rm3014.c:92-96).  The code has minimum distance 8:
 * up to three bit errors are always corrected.  out[i] = the 14 information bits of the nearest code word (what
 * tetra_rm3014_decode's *out would be for it) | distance to it << 16 | (received word is not a code word) << 24; among
 * equally near code words the one with the numerically smallest error pattern wins.  Host or device pointers. */
int tb200_rm3014_decode(tb200_ctx *ctx, const uint32_t *words, uint64_t n, uint32_t *out, int is_device);

/* GSMTAP framing of the decoded blocks: what tetra_gsmtap_makemsg (tetra_gsmtap.c:31-63) builds
 * when rx_tmv_unitdata_ind hands it a CRC-good block (tetra_upper_mac.c:480-488), for every block of n_slots slots:
 * a 16-byte GSMTAP v2 header (type TETRA_I1, timeslot tn-1, frame_number (mn*18+fn) in network order, sub_type
 * BSCH / AACH / BNCH / SCH_F or 0 for an unknown channel) + the type-1 bits eight per byte, first bit in the MSB
 * (osmo_ubit2pbit).  Frames are written back to back in the reference's delivery order (SB1, AACH, SB2 / AACH, SCH/F /
 * AACH, BLK1, BLK2); blocks with a wrong CRC produce none.  Frame lengths: AACH 18, SB1 24, SB2 / NDB half 32, SCH/F 50
 * bytes.  `slots` / `type1_packed` are the outputs of tb200_rx_stream_* (TB200_OUT_PACKED).  slot_off (may be NULL)
 * receives n_slots + 1 byte offsets: the frames of slot i are frames[slot_off[i] .. slot_off[i+1]).  frames == NULL
 * only counts.  Returns the number of bytes (written or needed), *n_frames (host pointer, may be NULL) the number of
 * frames; TB200_E_ARG if cap_bytes is too small.  is_device: slots, type1_packed, frames and slot_off are device
 * pointers (else host pointers, copied inside the call).  Not reproduced: the extra frames the reference sends when its
 * upper MAC re-enters a block that holds several PDUs (tetra_lower_mac.c:330-352) - those depend on the parse. */
#define TB200_GSMTAP_SLOT_MAX 82
/* length of a frame from byte 12 of its header (sub_type): AACH 18, BSCH 24, SCH/F 50, BNCH / unknown channel 32 */
#define TB200_GSMTAP_FRAME_LEN(sub_type) ((sub_type) == 2 ? 18u : (sub_type) == 1 ? 24u : (sub_type) == 5 ? 50u : 32u)
long long tb200_gsmtap_pack(tb200_ctx *ctx, const struct tb200_slot *slots, const uint32_t *type1_packed, uint64_t n_slots,
                            uint8_t *frames, uint64_t cap_bytes, uint64_t *slot_off, uint64_t *n_frames, int is_device);

/* ---- synthetic downlink generator (bench / test input, TX side) ----------------- */
struct tb200_gen_cfg {
	uint64_t seed;
	uint32_t sb_period;      /* burst k is a SYNC burst iff k % sb_period == 0 (0: none) */
	uint32_t lead_sb;        /* bursts [0, lead_sb) are SYNC bursts */
	uint32_t ndb2_per_256;   /* share of the other bursts that carry two half-slot blocks */
	uint32_t ber_per_65536;  /* i.i.d. payload bit-flip probability */
	uint32_t random_cell;    /* per-SB random MCC/MNC/colour code */
	uint32_t lead_in_bits;   /* random bits in front of burst 0 */
};
/* writes lead_in (if k0 == 0 and with_lead_in) + bursts [k0, k0+n) to a DEVICE buffer */
int tb200_gen_stream_dev(tb200_ctx *ctx, const struct tb200_gen_cfg *cfg, uint64_t k0, uint64_t n,
                         uint8_t *d_out, int with_lead_in);

/* ---- one stream sharded over several GPUs --------------------------------------
 * Slots of a LOCKED stream sit at fixed positions, so contiguous slot ranges decode independently
 * except for the cell state (scrambling code + TDMA time of the latest CRC-good SB1).  Pass 1
 * (search, classification, SB1) needs no cell state; the ranks then exchange one 32-byte summary
 * each (an all-gather), derive their carry-in with tb200_shard_carry_in() and run pass 2.
 * All pointers are device pointers except `summary` / `carry`. */
struct tb200_shard_summary {
	uint32_t n_slots;          /* slots this rank classified */
	uint32_t first_unlock;     /* first rank-local slot that lost lock, 0xffffffff if none */
	uint32_t has_good_sb;      /* the shard contains a CRC-good SB1 ... */
	uint32_t scramb_init;      /* ... announcing this cell code */
	uint32_t slots_after;      /* slots of the shard after that SB1 (n_slots if none) */
	uint16_t mcc, mnc;
	uint8_t  tn, fn, mn, cc;   /* its raw SYNC-PDU time and colour code */
	uint32_t pad;
};
/* lock acquisition on the head of the stream (UNLOCKED / KNOW_FSTART of tetra_burst_sync.c:67-106):
 * returns 1 and the absolute bit of the first LOCKED slot + the call that may process it, 0 if no lock */
int  tb200_find_lock(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint64_t *a0, uint64_t *cmin);
/* d_bits holds stream bits [base_bit, base_bit + n_bytes) encoded as options.input says (n_bytes counts stream
 * BITS; bit-packed shards start on a 128-bit boundary of the stream: base_bit % 128 == 0); the shard's slots
 * start at absolute bit a0 (= lock a0 + 510 * first slot index), cmin = lock cmin + first slot index,
 * n_end = stream length */
int  tb200_shard_pass1(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t base_bit, uint64_t n_bytes,
                       uint64_t a0, uint64_t cmin, uint64_t n_end, uint32_t n_slots,
                       struct tb200_shard_summary *summary);
void tb200_shard_carry_in(const struct tb200_shard_summary *all, int rank, const struct tb200_rx_carry *initial,
                          struct tb200_rx_carry *out);
long tb200_shard_pass2(tb200_ctx *ctx, const struct tb200_rx_carry *carry_in, struct tb200_slot *d_slots,
                       uint8_t *d_type1, uint32_t *d_type1_packed);

/* One stream read by several GPUs without a copy: the rank that holds the stream allocates it with
 * tb200_dev_alloc (plain device memory, exportable), exports a 64-byte handle, the other ranks (one process
 * per GPU) import it and hand the mapped pointer (+ byte offset of their shard) to tb200_shard_pass1: the
 * search kernel then pulls its shard straight out of the owner's HBM over NVLink (TMA bulk copies from peer
 * memory), instead of waiting for a scatter to finish first.  The pointer passed to export must be the
 * start of a tb200_dev_alloc allocation. */
void *tb200_dev_alloc(tb200_ctx *ctx, size_t bytes);
void  tb200_dev_free(tb200_ctx *ctx, void *p);
/* host <-> device copy for callers that do not link the CUDA runtime themselves (the C programs) */
int   tb200_dev_copy(tb200_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device);
int   tb200_ipc_export(tb200_ctx *ctx, const void *d_ptr, uint8_t handle[64]);
int   tb200_ipc_import(tb200_ctx *ctx, const uint8_t handle[64], void **d_ptr);
int   tb200_ipc_close(tb200_ctx *ctx, void *d_ptr);

/* ---- one stream decoded by several GPUs: the driver (C, NCCL) ---------------------------------
 * The calls above are the building blocks; this is the whole job, all ranks calling collectively (one process
 * or thread per GPU): rank 0 holds the stream (device memory, encoded as options.input says), acquires lock
 * (tetra_burst_sync.c:67-106), the slot range is cut into contiguous shards, every rank gets at its shard
 * (TB200_DIST_SCATTER: one grouped ncclSend / ncclRecv from rank 0; TB200_DIST_PEER: nothing is copied, the search
 * kernels read the shard out of rank 0's memory over NVLink - the stream must then live in tb200_dev_alloc memory),
 * runs its piece pipeline speculatively, the 56-byte summaries are all-gathered (the one exchange step of the path: the cell state of
 * tetra_lower_mac.c:291-302), every rank derives its carry-in and decodes the few slots that depended on it.  A lock loss inside a shard
 * (tetra_burst_sync.c:123-142) ends the segment right after the losing slot: the ranks behind it discard their
 * speculative work, rank 0 runs the UNLOCKED search from there exactly like the single-GPU receiver, and the rest of
 * the stream is sharded again as a new segment.  Results stay rank-local: `runs` says which global slot range
 * (in the order a single receiver would deliver them) each local run of slots is.
 *
 * TB200_DIST_PACK: the stream on rank 0 is one bit per byte (TB200_IN_BYTES); rank 0 first packs it to eight
 * bits per byte on the device (inside the call, it is part of the timed work), so that 8x fewer bytes travel.
 *
 * Plumbing: NCCL (libnccl.so.2, loaded at run time) - tb200_dist_create; or callbacks supplied by the caller
 * (the CPU tests drive the same code with gloo) - tb200_dist_create_with_ops. */
typedef struct tb200_dist tb200_dist;
#define TB200_DIST_ID_BYTES 128                 /* ncclUniqueId */
#define TB200_DIST_SCATTER  0u
#define TB200_DIST_PEER     1u
#define TB200_DIST_PACK     0x100u

struct tb200_dist_ops {
	void *user;
	/* `bytes` of host memory from rank `root` to everyone */
	int (*bcast)(void *user, void *host_buf, size_t bytes, int root);
	/* bytes_each of host memory from every rank, concatenated in rank order, to everyone */
	int (*allgather)(void *user, const void *host_send, void *host_recv, size_t bytes_each);
	/* device memory: bytes [offsets[r], offsets[r] + sizes[r]) of dev_src on rank `root` to dev_dst on rank r */
	int (*scatter)(void *user, const void *dev_src, const uint64_t *offsets, const uint64_t *sizes, void *dev_dst, int root);
};

struct tb200_dist_run {
	uint64_t global_slot;      /* index of the run's first slot in single-receiver delivery order */
	uint64_t local_slot;       /* where the run starts in this rank's output arrays */
	uint64_t n_slots;
};

struct tb200_dist_timing {       /* wall-clock of this rank's last tb200_dist_rx_stream, milliseconds */
	float total_ms, pack_ms, lock_ms, transfer_ms, pass1_ms, exchange_ms, pass2_ms;
	uint64_t bytes_sent;       /* bytes this rank shipped to other ranks (scatter), or served to them (peer: what they pulled) */
	uint32_t segments;         /* 1 + lock losses handled */
	uint32_t pad;
};

int  tb200_dist_get_id(uint8_t id[TB200_DIST_ID_BYTES]);     /* rank 0; hand the id to the other ranks out of band */
int  tb200_dist_create(tb200_dist **out, tb200_ctx *ctx, int rank, int world, const uint8_t id[TB200_DIST_ID_BYTES]);
int  tb200_dist_create_with_ops(tb200_dist **out, tb200_ctx *ctx, int rank, int world, const struct tb200_dist_ops *ops);
void tb200_dist_destroy(tb200_dist *d);
const char *tb200_dist_last_error(const tb200_dist *d);
/* d_bits / n_bits: the stream, rank 0 only (ignored elsewhere).  Outputs as for tb200_rx_stream_dev, rank-local.
 * A stream that keeps lock gives every rank ceil(slots / world) slots; lock losses skew that towards the first ranks
 * (see tb200_dist_max_local_slots, the bound that holds for any stream).  Arrays that turn out too small make the
 * call fail with TB200_E_ARG on that rank.  Returns the slots written on this rank or TB200_E_*. */
long tb200_dist_rx_stream(tb200_dist *d, const uint8_t *d_bits, uint64_t n_bits, uint32_t mode,
                          struct tb200_slot *d_slots, uint8_t *d_type1, uint32_t *d_type1_packed, uint64_t max_slots,
                          struct tb200_dist_run *runs, uint32_t max_runs, uint32_t *n_runs);
uint64_t tb200_dist_max_local_slots(uint64_t n_bits, int world);
int  tb200_dist_get_timing(const tb200_dist *d, struct tb200_dist_timing *out);

/* ---- helpers around the chain ------------------------------------------------------------------ */

/* Order-independent 64-bit digest of n slot records + their packed type-1 words (may be NULL).  Slot i counts as
 * global slot k_base + i, so the digests of the shards of a sharded run add up (mod 2^64) to the digest of a
 * single-GPU run of the same stream: a whole-run parity check that moves 8 bytes.  Host or device pointers. */
int tb200_slots_digest(tb200_ctx *ctx, const struct tb200_slot *slots, const uint32_t *type1_packed, uint64_t n,
                       uint64_t k_base, int is_device, uint64_t *digest);

/* TB200_IN_BYTES -> TB200_IN_PACKED on the device: n_bits bytes holding 0/1 (tetra-rx.c:82-95) to (n_bits+7)/8 bytes,
 * stream bit i = byte i>>3 bit i&7.  d_packed: 4-byte aligned, room for 4 * ((n_bits + 31) / 32) bytes. */
int tb200_pack_bits_dev(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint8_t *d_packed);

/* float_to_bits (float_to_bits.c:128-164) on the device: n_sym float32 symbols -> 2 * n_sym hard bits, packed eight per
 * byte (TB200_IN_PACKED: bit i of the stream = byte i>>3 bit i&7; room for 4 * ceil(n_sym / 16) bytes).  afc = 1 is the
 * program's -a, its pseudo-AFC (:140-147) with -f filter_val (default 0.0001) and -F filter_goal (default 0): a serial
 * float recurrence, reproduced bit for bit (speculative chunks + verification, csrc/tetra_afc.cuh).  *state (may be NULL):
 * the tracker's value in front of the first symbol (0 at the start of a stream), replaced by its value behind the last, so
 * a stream can be sliced piece by piece.  Returns the number of chunks that had to be redone (>= 0) or TB200_E_*.
 * The receive calls do this themselves for TB200_IN_F32SYM input when options.afc is set. */
int tb200_float_to_bits(tb200_ctx *ctx, const float *sym, uint64_t n_sym, int afc, float filter_val, float filter_goal,
                        float *state, uint8_t *packed_bits, int is_device);

/* ---- introspection used by the tests ------------------------------------------------ */

/* n x tetra_tdma_time_add_tn(tm, 1) (tetra_tdma.c:75-79) in closed form, as the kernels do it */
void tb200_debug_time_advance(uint32_t *tn, uint32_t *fn, uint32_t *mn, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* TETRA_B200_H */
