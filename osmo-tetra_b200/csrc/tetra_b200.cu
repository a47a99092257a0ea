/*
 * tetra_b200.cu - host side of libtetra_b200.so: tables, the lock state machine of
 * tetra_burst_sync_in(), the batch driver and the C ABI declared in include/tetra_b200.h.
 *
 * The arithmetic (search, descramble, de-interleave, Viterbi, CRC) runs in the CUDA kernels
 * of tetra_kernels.cuh; what stays on the host is control: which modelled read() call
 * processes which slot, when the receiver is UNLOCKED / KNOW_FSTART / LOCKED
 * (phy/tetra_burst_sync.c:54-154), and the stream plumbing.  There is no CPU decode path:
 * without a CUDA device tb200_create() fails.
 */
#include "tetra_kernels.cuh"
#include "tetra_lane.cuh"
#include "tetra_classify_tile.cuh"
#include "tetra_stage_tma.cuh"
#include "tetra_gen.cuh"
#include "tetra_gsmtap.cuh"
#include "tetra_util.cuh"
#include "tetra_afc.cuh"
#include "../../include/tetra_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>

using namespace tb;

static_assert(sizeof(tb200_slot) == sizeof(SlotOut), "slot ABI");
static_assert(sizeof(tb200_gen_cfg) == sizeof(GenCfg), "gen cfg ABI");
static_assert(sizeof(tb200_record) == 288, "record ABI");

/* ------------------------------------------------------------------ tables -- */

static inline unsigned lfsr_step_host(uint32_t *st)   /* tetra_scramb.c:34-50 */
{
	unsigned fb = __builtin_parity(*st & 0xDB710641u);
	*st = (*st >> 1) | ((uint32_t)fb << 31);
	return fb;
}

static void lfsr_words_host(uint32_t init, uint32_t *w, int nwords)
{
	for (int i = 0; i < nwords; i++) {
		uint32_t v = 0;
		for (int b = 0; b < 32; b++)
			v |= (uint32_t)lfsr_step_host(&init) << b;
		w[i] = v;
	}
}

static uint16_t crc16_host(const uint8_t *bits, int len, uint16_t crc)   /* crc_simple.c:65-82 */
{
	for (int i = 0; i < len; i++) {
		unsigned top = ((crc >> 15) ^ bits[i]) & 1;
		crc = (uint16_t)(crc << 1);
		if (top) crc ^= 0x1021;
	}
	return crc;
}

static void build_tables(Tables *t)
{
	memset(t, 0, sizeof(*t));
	for (int b = 0; b < 32; b++)
		lfsr_words_host(1u << b, t->lfsr_col[b], 16);
	lfsr_words_host(3, t->lfsr_sb1, 16);

	std::vector<uint8_t> msg(300, 0);
	for (int d = 0; d < 288; d++) {
		std::fill(msg.begin(), msg.end(), 0);
		msg[0] = 1;
		t->crc_pow[d] = crc16_host(msg.data(), d + 1, 0);
	}
	std::fill(msg.begin(), msg.end(), 0);
	const int Ls[4] = {76, 140, 284, 108};
	for (int i = 0; i < 4; i++)
		t->crc_init[i] = crc16_host(msg.data(), Ls[i], 0xffff);

	/* reflected CRC tables: register bit-reversed, polynomial 0x1021 -> 0x8408 */
	for (int x = 0; x < 256; x++) {
		uint32_t r = x;
		for (int i = 0; i < 8; i++) r = (r & 1) ? (r >> 1) ^ 0x8408 : (r >> 1);
		t->crc_tab_r[x] = r;
	}
	for (int x = 0; x < 16; x++) {
		uint32_t r = x;
		for (int i = 0; i < 4; i++) r = (r & 1) ? (r >> 1) ^ 0x8408 : (r >> 1);
		t->crc_tab_r[256 + x] = r;
	}

	/* the scrambler 32 steps at a time, one table per byte of the register */
	for (int byte = 0; byte < 4; byte++)
		for (uint32_t v = 0; v < 256; v++) {
			uint32_t st = v << (8 * byte);
			for (int i = 0; i < 32; i++) lfsr_step_host(&st);
			t->lfsr_leap[byte][v] = st;
		}

	/* RM(30,14): parity halves of the systematic generator matrix (tetra_rm3014.c:28-43; rows as tetra_rm3014_init
	 * builds them, :59-72: information bit i of the 14 sits at word bit 29 - i) */
	static const uint16_t rm_parity[14] = { 0x9b60, 0x2de0, 0xfc20, 0xe03c, 0x983a, 0x5436, 0x2c2e,
	                                        0xffdf, 0x8339, 0x42b5, 0x21ad, 0x1273, 0x096b, 0x04e7 };
	for (unsigned half = 0; half < 2; half++)
		for (unsigned v = 0; v < 128; v++) {
			uint16_t p = 0;
			for (unsigned b = 0; b < 7; b++)
				if ((v >> b) & 1) p ^= rm_parity[13 - (7 * half + b)];     /* info bit j (LSB = last on air) <-> row 13 - j */
			t->rm_par[half][v] = p;
		}

	/* pre-filter blind spot, by running the filter of tetra_burst.c:286-303 on a buffer
	 * that carries the sequence at offset k with the bit before it set to `prev` */
	const uint64_t seqs[3] = {SEQ_Y, SEQ_N, SEQ_P};
	const uint32_t pre[5] = {(uint32_t)(SEQ_Y & 0x3fffff), SEQ_N, SEQ_P, SEQ_Q, SEQ_X & 0x3fffff};
	uint32_t pre_msb[5];
	for (int i = 0; i < 5; i++) {          /* the filter shifts left: first bit ends up in bit 21 */
		uint32_t v = 0;
		for (int b = 0; b < 22; b++)
			v = (v << 1) | ((pre[i] >> b) & 1);
		pre_msb[i] = v;
	}
	for (int s = 0; s < 3; s++)
		for (int prev = 0; prev < 2; prev++) {
			uint32_t ok = 0;
			for (int k = 0; k <= 20; k++) {
				uint8_t in[96];
				memset(in, 0, sizeof(in));
				for (int b = 0; b < 38; b++)
					in[k + b] = (seqs[s] >> b) & 1;     /* bits past the sequence do not reach the filter at k */
				if (k > 0) in[k - 1] = prev;
				uint32_t filt = 0;
				for (int i = 0; i < 20; i++)
					filt = (filt << 1) | in[i];
				for (int c = 0; c <= k; c++)
					filt = ((filt << 1) | in[c + 21]) & 0x3fffff;
				for (int i = 0; i < 5; i++)
					if (filt == pre_msb[i]) ok |= 1u << k;
			}
			t->blind_ok[s][prev] = ok;
		}
}

/* --------------------------------------------------------------- context -- */

#define NBUF 3

struct RxHost {
	uint32_t state = TB200_RX_UNLOCKED;
	uint64_t calls = 0;          /* modelled tetra_burst_sync_in() calls done */
	uint64_t buf_start = 0;      /* bitbuf_start_bitnum (64-bit here) */
	uint32_t bits_in_buf = 0;
	uint64_t next_frame_start = 0;
	/* the current run of equal-length reads (CallGeom): calls and bits before it */
	uint64_t c_base = 0, t_base = 0;
	uint64_t delivered = 0;      /* bits the modelled calls have delivered so far (the end of the last run) */
};

struct SyncHit { uint64_t pos; uint32_t prev; };

/* ---- host threads that pack the caller's one-bit-per-byte stream before it crosses PCIe (options.host_pack_threads) ----
 * Eight 0/1 bytes -> one byte with a 64-bit multiply (byte i lands on bit 56 + i, no two partial products share a bit):
 * stream bit i = byte i >> 3, bit i & 7, the layout of TB200_IN_PACKED.  A persistent pool; the calling thread hands out one
 * job (a range of bytes) at a time and waits for it, the GPU works on the previous piece meanwhile. */
static inline uint8_t pack8_bytes(const uint8_t *p)
{
	uint64_t x;
	memcpy(&x, p, 8);
	return (uint8_t)(((x & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
}
static void pack_range(const uint8_t *src, size_t n_bits, uint8_t *dst)        /* n_bits bytes -> (n_bits + 7) / 8 bytes */
{
	const size_t full = n_bits >> 3;
	for (size_t i = 0; i < full; i++) dst[i] = pack8_bytes(src + 8 * i);
	if (n_bits & 7) {
		uint8_t tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		memcpy(tmp, src + 8 * full, n_bits & 7);
		dst[full] = pack8_bytes(tmp);
	}
}
struct PackPool {
	std::vector<std::thread> th;
	std::mutex m;
	std::condition_variable cv_go, cv_done;
	uint64_t gen = 0;
	unsigned pending = 0;
	bool quit = false;
	const uint8_t *src = nullptr; uint8_t *dst = nullptr; size_t n_bits = 0;
	unsigned n_workers = 0;          /* workers of the current job (set with it) */
	void worker(unsigned id, uint64_t seen)       /* seen: the job counter when the thread was made - only later jobs are its own */
	{
		for (;;) {
			const uint8_t *s; uint8_t *d; size_t nb; unsigned nth;
			{
				std::unique_lock<std::mutex> lk(m);
				cv_go.wait(lk, [&] { return quit || gen != seen; });
				if (quit) return;
				seen = gen; s = src; d = dst; nb = n_bits; nth = n_workers;
			}
			/* slices of whole 64-byte output lines */
			const size_t out_bytes = (nb + 7) >> 3, per = ((out_bytes + nth - 1) / nth + 63) & ~(size_t)63;
			const size_t o0 = std::min(out_bytes, (size_t)id * per), o1 = std::min(out_bytes, o0 + per);
			if (o1 > o0) pack_range(s + 8 * o0, std::min(nb, 8 * o1) - 8 * o0, d + o0);
			{
				std::lock_guard<std::mutex> lk(m);
				if (--pending == 0) cv_done.notify_one();
			}
		}
	}
	void resize(unsigned n)
	{
		if (n == th.size()) return;
		stop();
		quit = false; pending = 0;
		const uint64_t g0 = gen;
		th.reserve(n);
		for (unsigned i = 0; i < n; i++) th.emplace_back([this, i, g0] { worker(i, g0); });
	}
	/* start a job and come back; wait() returns when it is done.  One job at a time. */
	void start(const uint8_t *s, size_t nb, uint8_t *d)
	{
		if (th.empty()) { pack_range(s, nb, d); return; }
		std::lock_guard<std::mutex> lk(m);
		src = s; dst = d; n_bits = nb; pending = n_workers = (unsigned)th.size(); ++gen;
		cv_go.notify_all();
	}
	void wait()
	{
		if (th.empty()) return;
		std::unique_lock<std::mutex> lk(m);
		cv_done.wait(lk, [&] { return pending == 0; });
	}
	void run(const uint8_t *s, size_t nb, uint8_t *d) { start(s, nb, d); wait(); }
	void stop()
	{
		{ std::lock_guard<std::mutex> lk(m); quit = true; }
		cv_go.notify_all();
		for (auto &t : th) t.join();
		th.clear();
	}
	~PackPool() { stop(); }
};


struct tb200_ctx {
	int device = 0;
	int sm_count = 148;
	char err[256] = {0};
	tb200_options opt;
	tb200_stats stats;
	Tables h_tab;
	Tables *d_tab = nullptr;
	DevCarry *d_carry = nullptr;     /* chain of carries, one per piece (+1) */
	size_t carry_cap = 0;
	/* per-piece workspace: what pass 1 (search, SB1, scans; stream s_front) leaves for pass 2 (decode; stream
	 * s_compute).  Two sets, so that pass 1 of piece i+1 runs while pass 2 of piece i does: the search is HBM
	 * bound, the decode pass integer-issue bound. */
	struct WorkSet {
		SlotWs *ws = nullptr;
		uint32_t *slot_bits = nullptr;
		int32_t *last_good = nullptr, *blk_last = nullptr, *blk_prev = nullptr;
		uint32_t *sb_list = nullptr;     /* slots classified as SYNC bursts; [ws_slots] + counter at the end */
		uint32_t *kind_list = nullptr;   /* [4][ws_slots] slots grouped by kind + [4] counters at the end (k_scan_blocks) */
	} wset[2];
	size_t ws_slots = 0;
	cudaStream_t s_front = nullptr;
	cudaEvent_t ev_front[2] = {nullptr, nullptr}, ev_back[2] = {nullptr, nullptr};   /* pass 1 / pass 2 of the set's last piece done */
	bool back_pending[2] = {false, false};
	uint32_t *d_lane_scratch = nullptr;   /* survivor decisions of the decode pass, one area per resident CTA */
	uint32_t *d_sb1_scratch = nullptr;    /* the same for the SB1 pass (it runs next to the decode pass of the previous piece) */
	unsigned lane_ctas = 0;               /* resident CTAs of the lane kernels (grid size) */
	uint32_t *d_units = nullptr;          /* split decode pass: the unit blocks between prepare / trellis / finish (lane_unit_blocks(ws_slots)) */
	int lane_form = 1;                    /* 0 fused kernel, 1 prepare | trellis+finish, 2 prepare | trellis | finish (TB200_LANE_FORM) */
	unsigned prep_ctas = 0, fin_ctas = 0; /* grids of the prepare / finish kernels */
	unsigned form_ctas[3] = {0, 0, 0};    /* grid of the trellis-carrying kernel of each form (<= lane_ctas, which sizes the history scratch) */
	uint32_t *d_flags = nullptr;     /* first unlocking slot per piece */
	uint32_t *h_flags = nullptr;     /* pinned mirror */
	size_t flags_cap = 0;
	unsigned long long *d_pstats = nullptr;   /* [flags_cap][3] per-piece counters of the decode pass (tb200_stats) */
	unsigned long long *h_pstats = nullptr;   /* pinned mirror */
	/* host path staging */
	uint8_t *d_in[NBUF] = {nullptr, nullptr, nullptr};
	size_t in_cap = 0;
	uint8_t *h_pack[NBUF] = {nullptr, nullptr, nullptr};   /* pinned: a piece of the caller's bytes, packed by the host threads */
	size_t h_pack_cap = 0;
	PackPool pack_pool;
	long pack_inflight = -1;         /* piece whose bytes the pool is packing (or has packed) ahead of its issue */
	uint64_t pack_inflight_base = 0, pack_inflight_hi = 0;
	SlotOut *d_oslots[NBUF] = {nullptr, nullptr, nullptr};
	uint8_t *d_otype1[NBUF] = {nullptr, nullptr, nullptr};
	uint32_t *d_opacked[NBUF] = {nullptr, nullptr, nullptr};
	uint32_t *d_ocrc[NBUF] = {nullptr, nullptr, nullptr};
	uint32_t *user_crc = nullptr;    /* optional CRC-register output (tb200_set_crc_buffer), same residency as the slots */
	uint32_t *user_aach = nullptr;   /* optional RM(30,14)-decoded AACH output (tb200_set_aach_buffer) */
	uint32_t *d_oaach[NBUF] = {nullptr, nullptr, nullptr};
	uint32_t *d_rm_leader = nullptr; /* [2^16] coset leaders, built on first use */
	std::vector<tb200_lock_event> lock_events;   /* lock acquisitions of the last rx call */
	size_t out_cap = 0;
	/* UNLOCKED search */
	uint32_t *d_hits = nullptr;      /* [0] = count, then (pos_lo, pos_hi|prev<<31) pairs */
	uint32_t *h_hits = nullptr;
	uint8_t *d_region = nullptr;
	size_t region_cap = 0;
	std::vector<SyncHit> hits;
	uint64_t hits_lo = 0, hits_hi = 0;
	cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
	cudaEvent_t ev_h2d[NBUF], ev_comp[NBUF], ev_d2h[NBUF];
	/* stream state across calls */
	RxHost rx;
	std::vector<uint8_t> tail;       /* bits [tail_base, fed_end) kept from earlier calls (host path) */
	uint8_t *d_tail = nullptr;       /* the same for a stream fed with device buffers; tail_dev says which of the two is current */
	size_t d_tail_cap = 0, d_tail_bytes = 0;
	bool tail_dev = false;
	uint8_t *d_cont = nullptr;       /* device path: kept tail + head of the new buffer, contiguous */
	size_t d_cont_cap = 0;
	uint64_t tail_base = 0;
	int tail_fmt = 0;                /* encoding of the tail = input format of the stream */
	uint64_t fed_end = 0;            /* absolute bits handed to the ctx so far */
	DevCarry h_carry;
	DevCarry *h_carry_pin = nullptr;   /* pinned: [0] upload staging, [1] download staging (ordered on s_compute) */
	bool stop_at_lock = false;       /* tb200_find_lock: return from rx_run as soon as LOCKED is reached */
	/* sharded decode: what pass 1 left for pass 2 */
	uint64_t shard_a0 = 0;
	uint32_t shard_slots = 0;
	/* float_to_bits -a (options.afc): the tracker's state between calls, scratch of the pre-pass */
	float afc_state = 0.f;
	float *d_afc_sym = nullptr; size_t afc_sym_cap = 0;
	uint8_t *d_afc_bits = nullptr; size_t afc_bits_cap = 0;
	float *d_afc_f = nullptr; size_t afc_f_cap = 0;
	std::vector<float> h_afc_f;
	std::vector<uint8_t> h_afc_bits;
	uint64_t afc_repairs = 0;        /* chunks of the last pre-pass whose speculative start state was wrong */
	/* grow-only device scratch of the leaf operators: no cudaMalloc / cudaFree per call, nothing to leak on an error path */
	void *leaf_mem[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	size_t leaf_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	cudaEvent_t leaf_ev[2] = {nullptr, nullptr};
	/* profiling (options.profile) */
	std::vector<cudaEvent_t> prof_ev;    /* PE_COUNT per piece (enum PE_*) */
	size_t prof_used = 0;
	tb200_timing timing;
};

/* coset leaders of the RM(30,14) code: for every 16-bit syndrome the lightest error pattern that has it, the
 * numerically smallest among equally light ones (patterns visited by weight, then value: Gosper's hack) */
static void build_rm_leaders(const Tables *t, std::vector<uint32_t> &leader)
{
	leader.assign(1u << 16, 0xffffffffu);
	size_t left = leader.size();
	leader[0] = 0; left--;
	for (int w = 1; w <= 30 && left; w++) {
		uint32_t e = (1u << w) - 1;
		while (e < (1u << 30)) {
			const uint32_t syn = rm3014_syndrome(t, e);
			if (leader[syn] == 0xffffffffu) { leader[syn] = e; left--; }
			const uint32_t c = e & (0u - e), r = e + c;
			e = (((r ^ e) >> 2) / c) | r;
		}
	}
}

static int fail(tb200_ctx *c, int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(c->err, sizeof(c->err), fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(ctx, TB200_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *tb200_version(void)
{
#ifdef TB_SIMT_EMULATION
	return "tetra_b200 0.1 (SIMT emulation build - tests only)";
#else
	return "tetra_b200 0.1 (sm_100a)";
#endif
}

extern "C" void tb200_default_options(tb200_options *o)
{
	o->chunk_bits = 64;
	o->output = TB200_OUT_UNPACKED | TB200_OUT_PACKED;
	o->viterbi = TB200_VITERBI_LANE;
	o->pipeline_slots = 0;
	o->profile = 0;
	o->input = TB200_IN_BYTES;
	o->serial_passes = getenv("TB200_SERIAL") ? 1 : 0;
	o->afc = 0; o->afc_filter_val = 0.0001f; o->afc_filter_goal = 0.f;      /* float_to_bits.c:83-86 */
	o->viterbi_tie = TETRA_VITERBI_TIE_DEFAULT;
	o->host_pack_threads = 0;
}

extern "C" int tb200_set_options(tb200_ctx *ctx, const tb200_options *o)
{
	if (!ctx || !o) return TB200_E_ARG;
	if (o->chunk_bits < 1 || o->chunk_bits > 296)
		return fail(ctx, TB200_E_ARG, "chunk_bits must be 1..296");
	if (o->viterbi > TB200_VITERBI_LANE)
		return fail(ctx, TB200_E_ARG, "unknown viterbi variant");
	if (o->input > TB200_IN_F32SYM)
		return fail(ctx, TB200_E_ARG, "unknown input format");
	if (o->viterbi_tie > 1)
		return fail(ctx, TB200_E_ARG, "viterbi_tie must be 0 or 1");
	if (o->host_pack_threads > 256)
		return fail(ctx, TB200_E_ARG, "host_pack_threads must be 0..256");
	if (o->afc > 1 || (o->afc && (!(o->afc_filter_val > 0.f) || !(o->afc_filter_val <= 1.f))))
		return fail(ctx, TB200_E_ARG, "afc must be 0 or 1 with afc_filter_val in (0, 1]");
	if (o->input != TB200_IN_BYTES && o->viterbi != TB200_VITERBI_LANE)
		return fail(ctx, TB200_E_ARG, "packed / symbol input needs the lane kernels (TB200_VITERBI_LANE)");
	ctx->opt = *o;
	return 0;
}

extern "C" const char *tb200_last_error(const tb200_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

/* ---- device memory another process can map (one stream read by several GPUs over NVLink) ---- */

extern "C" void *tb200_dev_alloc(tb200_ctx *ctx, size_t bytes)
{
	if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
	void *p = nullptr;
	if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

extern "C" void tb200_dev_free(tb200_ctx *ctx, void *p)
{
	if (ctx) cudaSetDevice(ctx->device);
	cudaFree(p);
}

extern "C" int tb200_dev_copy(tb200_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device)
{
	if (!ctx || (bytes && (!dst || !src))) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	CU(cudaMemcpy(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int tb200_ipc_export(tb200_ctx *ctx, const void *d_ptr, uint8_t handle[64])
{
	if (!ctx || !d_ptr || !handle) return TB200_E_ARG;
#ifdef TB_SIMT_EMULATION
	return fail(ctx, TB200_E_STATE, "no IPC in the emulation build");
#else
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
	memcpy(handle, &h, 64);
	return 0;
#endif
}

extern "C" int tb200_ipc_import(tb200_ctx *ctx, const uint8_t handle[64], void **d_ptr)
{
	if (!ctx || !d_ptr || !handle) return TB200_E_ARG;
#ifdef TB_SIMT_EMULATION
	return fail(ctx, TB200_E_STATE, "no IPC in the emulation build");
#else
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return 0;
#endif
}

extern "C" int tb200_ipc_close(tb200_ctx *ctx, void *d_ptr)
{
	if (!ctx) return TB200_E_ARG;
#ifdef TB_SIMT_EMULATION
	(void)d_ptr;
	return 0;
#else
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	CU(cudaIpcCloseMemHandle(d_ptr));
	return 0;
#endif
}

extern "C" int tb200_set_crc_buffer(tb200_ctx *ctx, uint32_t *crc)
{
	if (!ctx) return TB200_E_ARG;
	ctx->user_crc = crc;
	return 0;
}

static int ensure_rm_leaders(tb200_ctx *ctx)
{
	if (ctx->d_rm_leader) return 0;
	std::vector<uint32_t> leader;
	build_rm_leaders(&ctx->h_tab, leader);
	CU(cudaMalloc((void **)&ctx->d_rm_leader, leader.size() * sizeof(uint32_t)));
	CU(cudaMemcpy(ctx->d_rm_leader, leader.data(), leader.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
	return 0;
}

extern "C" int tb200_set_aach_buffer(tb200_ctx *ctx, uint32_t *aach)
{
	if (!ctx) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	if (aach) { int rc = ensure_rm_leaders(ctx); if (rc) return rc; }
	ctx->user_aach = aach;
	return 0;
}

extern "C" size_t tb200_get_lock_events(const tb200_ctx *ctx, tb200_lock_event *ev, size_t max_events)
{
	if (!ctx) return 0;
	const size_t n = ctx->lock_events.size();
	for (size_t i = 0; ev && i < n && i < max_events; i++) ev[i] = ctx->lock_events[i];
	return n;
}

extern "C" int tb200_create(tb200_ctx **out, int device)
{
	if (!out) return TB200_E_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
		fprintf(stderr, "tetra_b200: no usable CUDA device (asked for %d of %d); there is no CPU path\n", device, ndev);
		return TB200_E_CUDA;
	}
	tb200_ctx *ctx = new tb200_ctx();
	ctx->device = device;
	tb200_default_options(&ctx->opt);
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	memset(&ctx->h_carry, 0, sizeof(ctx->h_carry));
	auto bail = [&](const char *what) { fprintf(stderr, "tetra_b200: %s failed\n", what); delete ctx; return TB200_E_CUDA; };
	if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice");
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail("cudaGetDeviceProperties");
	ctx->sm_count = prop.multiProcessorCount;
	build_tables(&ctx->h_tab);
	if (cudaMalloc((void **)&ctx->d_tab, sizeof(Tables)) != cudaSuccess) return bail("cudaMalloc");
	if (cudaMemcpy(ctx->d_tab, &ctx->h_tab, sizeof(Tables), cudaMemcpyHostToDevice) != cudaSuccess) return bail("cudaMemcpy");
	if (cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
	if (cudaStreamCreateWithFlags(&ctx->s_front, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
	for (int i = 0; i < 2; i++) {
		cudaEventCreateWithFlags(&ctx->ev_front[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&ctx->ev_back[i], cudaEventDisableTiming);
	}
	if (cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
	if (cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) != cudaSuccess) return bail("stream");
	for (int i = 0; i < NBUF; i++) {
		cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming);
		cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming);
	}
	{
		int per_sm = 8;
		if (const char *e = getenv("TB200_LANE_FORM")) { const int v = atoi(e); if (v >= 0 && v <= 2) ctx->lane_form = v; }
		int prep_per_sm = 8, fin_per_sm = 8, fused = 8, split = 8, split3 = 8;
#ifndef TB_SIMT_EMULATION
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fused, k_decode_lane<false>, LANE_NT,
		                                                   lane_smem_words(LANE_NT) * sizeof(uint32_t)) != cudaSuccess || fused < 1 ||
		    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&split, k_lane_trellis<false, true>, LANE_NT,
		                                                   lane_trellis_smem_words<true>() * sizeof(uint32_t)) != cudaSuccess || split < 1 ||
		    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&split3, k_lane_trellis<false, false>, LANE_NT,
		                                                   lane_trellis_smem_words<false>() * sizeof(uint32_t)) != cudaSuccess || split3 < 1 ||
		    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&prep_per_sm, k_lane_prepare, LANE_NT,
		                                                   lane_prepare_smem_words() * sizeof(uint32_t)) != cudaSuccess || prep_per_sm < 1 ||
		    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fin_per_sm, k_lane_finish, LANE_NT,
		                                                   lane_finish_smem_words() * sizeof(uint32_t)) != cudaSuccess || fin_per_sm < 1)
			return bail("cudaOccupancyMaxActiveBlocksPerMultiprocessor");
#endif
		if (const char *e = getenv("TB200_LANE_CTAS_PER_SM")) {
			const int v = atoi(e);
			if (v >= 1) { fused = std::min(fused, v); split = std::min(split, v); split3 = std::min(split3, v); }
		}
		if (const char *e = getenv("TB200_PREP_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < prep_per_sm) prep_per_sm = v; }
		if (const char *e = getenv("TB200_FIN_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < fin_per_sm) fin_per_sm = v; }
		per_sm = std::max(fused, std::max(split, split3));      /* the history scratch is sized for the largest grid */
		ctx->form_ctas[0] = (unsigned)(ctx->sm_count * fused);
		ctx->form_ctas[1] = (unsigned)(ctx->sm_count * split);
		ctx->form_ctas[2] = (unsigned)(ctx->sm_count * split3);
		ctx->lane_ctas = (unsigned)(ctx->sm_count * per_sm);
		ctx->prep_ctas = (unsigned)(ctx->sm_count * prep_per_sm);
		ctx->fin_ctas = (unsigned)(ctx->sm_count * fin_per_sm);
		const size_t scratch_bytes = (size_t)ctx->lane_ctas * lane_scratch_words_per_cta(LANE_NT) * sizeof(uint32_t);
		if (cudaMalloc((void **)&ctx->d_lane_scratch, scratch_bytes) != cudaSuccess ||
		    cudaMalloc((void **)&ctx->d_sb1_scratch, scratch_bytes) != cudaSuccess)
			return bail("cudaMalloc");
#ifndef TB_SIMT_EMULATION
		/* Optional (TB200_L2_PERSIST=1): pin the survivor histories (written once, read once ~100 us later by the
		 * same CTA) in the L2 set-aside.  Measured: decode DRAM traffic 805 -> 551 MB per 10^6 bursts, but the
		 * decode kernel is not DRAM bound (no gain) and the set-aside costs the HBM-bound kernels of the same
		 * context 10-50 % (search 0.110 -> 0.120 ms, stand-alone stage 6.1 -> 2.7 TB/s), so it is off by default. */
		if (getenv("TB200_L2_PERSIST") && prop.persistingL2CacheMaxSize > 0) {
			const size_t set_aside = std::min<size_t>(scratch_bytes, (size_t)prop.persistingL2CacheMaxSize);
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside) == cudaSuccess) {
				cudaStreamAttrValue av;
				memset(&av, 0, sizeof(av));
				av.accessPolicyWindow.base_ptr = ctx->d_lane_scratch;
				av.accessPolicyWindow.num_bytes = std::min<size_t>(scratch_bytes, (size_t)prop.accessPolicyMaxWindowSize);
				av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)av.accessPolicyWindow.num_bytes);
				av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
				av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
				cudaStreamSetAttribute(ctx->s_compute, cudaStreamAttributeAccessPolicyWindow, &av);
			}
			cudaGetLastError();
		}
#endif
	}
#ifndef TB_SIMT_EMULATION
	if (cudaFuncSetAttribute(k_stage_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM) != cudaSuccess)
		return bail("cudaFuncSetAttribute");
	if (cudaFuncSetAttribute(k_classify_tile<IN_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct_smem<IN_BYTES>()) != cudaSuccess ||
	    cudaFuncSetAttribute(k_classify_tile<IN_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct_smem<IN_PACKED>()) != cudaSuccess ||
	    cudaFuncSetAttribute(k_classify_tile<IN_F32SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ct_smem<IN_F32SYM>()) != cudaSuccess)
		return bail("cudaFuncSetAttribute");
#endif
	if (cudaMalloc((void **)&ctx->d_hits, sizeof(uint32_t) * (2 * 8192 + 2)) != cudaSuccess) return bail("cudaMalloc");
	if (cudaHostAlloc((void **)&ctx->h_hits, sizeof(uint32_t) * (2 * 8192 + 2), cudaHostAllocDefault) != cudaSuccess) return bail("cudaHostAlloc");
	if (cudaHostAlloc((void **)&ctx->h_carry_pin, 2 * sizeof(DevCarry), cudaHostAllocDefault) != cudaSuccess) return bail("cudaHostAlloc");
	*out = ctx;
	return 0;
}

extern "C" void tb200_destroy(tb200_ctx *ctx)
{
	if (!ctx) return;
	ctx->pack_pool.stop();            /* before its staging buffers go */
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	cudaFree(ctx->d_tab); cudaFree(ctx->d_carry);
	for (int i = 0; i < 2; i++) {
		tb200_ctx::WorkSet &w = ctx->wset[i];
		cudaFree(w.ws); cudaFree(w.slot_bits); cudaFree(w.last_good); cudaFree(w.blk_last); cudaFree(w.blk_prev);
		cudaFree(w.sb_list); cudaFree(w.kind_list);
		cudaEventDestroy(ctx->ev_front[i]); cudaEventDestroy(ctx->ev_back[i]);
	}
	cudaFree(ctx->d_flags); cudaFreeHost(ctx->h_flags); cudaFree(ctx->d_lane_scratch); cudaFree(ctx->d_sb1_scratch); cudaFree(ctx->d_units);
	cudaFree(ctx->d_pstats); cudaFreeHost(ctx->h_pstats); cudaFree(ctx->d_rm_leader);
	cudaFree(ctx->d_afc_sym); cudaFree(ctx->d_afc_bits); cudaFree(ctx->d_afc_f); cudaFree(ctx->d_tail); cudaFree(ctx->d_cont);
	for (int i = 0; i < NBUF; i++) cudaFree(ctx->d_oaach[i]);
	for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
	for (int i = 0; i < 8; i++) cudaFree(ctx->leaf_mem[i]);
	for (int i = 0; i < 2; i++) if (ctx->leaf_ev[i]) cudaEventDestroy(ctx->leaf_ev[i]);
	for (int i = 0; i < NBUF; i++) {
		cudaFree(ctx->d_in[i]); cudaFreeHost(ctx->h_pack[i]); cudaFree(ctx->d_oslots[i]); cudaFree(ctx->d_otype1[i]); cudaFree(ctx->d_opacked[i]); cudaFree(ctx->d_ocrc[i]);
		cudaEventDestroy(ctx->ev_h2d[i]); cudaEventDestroy(ctx->ev_comp[i]); cudaEventDestroy(ctx->ev_d2h[i]);
	}
	cudaFree(ctx->d_hits); cudaFreeHost(ctx->h_hits); cudaFreeHost(ctx->h_carry_pin); cudaFree(ctx->d_region);
	cudaStreamDestroy(ctx->s_compute); cudaStreamDestroy(ctx->s_front); cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_d2h);
	delete ctx;
}

extern "C" void *tb200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
	return p;
}
extern "C" void tb200_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" uint64_t tb200_max_slots(uint64_t n_bits) { return n_bits / SLOT_BITS + 1; }

/* ---------------------------------------------------------- small helpers -- */

template <typename T>
static int grow(tb200_ctx *ctx, T **p, size_t n)
{
	if (*p) cudaFree(*p);
	*p = nullptr;
	CU(cudaMalloc((void **)p, n * sizeof(T)));
	return 0;
}

static int ensure_workspace(tb200_ctx *ctx, size_t slots)
{
	if (slots <= ctx->ws_slots) return 0;
	CU(cudaDeviceSynchronize());
	int rc;
	for (int i = 0; i < 2; i++) {
		tb200_ctx::WorkSet &w = ctx->wset[i];
		if ((rc = grow(ctx, &w.ws, slots))) return rc;
		if ((rc = grow(ctx, &w.slot_bits, slots * 16))) return rc;
		if ((rc = grow(ctx, &w.last_good, slots))) return rc;
		if ((rc = grow(ctx, &w.blk_last, slots / 1024 + 2))) return rc;
		if ((rc = grow(ctx, &w.blk_prev, slots / 1024 + 2))) return rc;
		if ((rc = grow(ctx, &w.sb_list, slots + 4))) return rc;
		if ((rc = grow(ctx, &w.kind_list, 4 * slots + 4))) return rc;
	}
	if ((rc = grow(ctx, &ctx->d_units, lane_unit_blocks(slots) * LANE_UNIT_WORDS))) return rc;
	ctx->back_pending[0] = ctx->back_pending[1] = false;
	ctx->ws_slots = slots;
	return 0;
}

static int ensure_pieces(tb200_ctx *ctx, size_t npieces)
{
	if (npieces + 2 > ctx->carry_cap) {
		CU(cudaDeviceSynchronize());
		DevCarry keep = ctx->h_carry;
		size_t cap = npieces + 64;
		int rc;
		if ((rc = grow(ctx, &ctx->d_carry, cap))) return rc;
		ctx->carry_cap = cap;
		CU(cudaMemcpy(ctx->d_carry, &keep, sizeof(keep), cudaMemcpyHostToDevice));
	}
	if (npieces + 2 > ctx->flags_cap) {
		CU(cudaDeviceSynchronize());
		size_t cap = npieces + 64;
		int rc;
		if ((rc = grow(ctx, &ctx->d_flags, 2 * cap))) return rc;
		if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
		CU(cudaHostAlloc((void **)&ctx->h_flags, 2 * cap * sizeof(uint32_t), cudaHostAllocDefault));
		if ((rc = grow(ctx, &ctx->d_pstats, 3 * cap))) return rc;
		if (ctx->h_pstats) cudaFreeHost(ctx->h_pstats);
		CU(cudaHostAlloc((void **)&ctx->h_pstats, 3 * cap * sizeof(unsigned long long), cudaHostAllocDefault));
		ctx->flags_cap = cap;
	}
	return 0;
}

static int ensure_staging(tb200_ctx *ctx, size_t in_bytes, size_t slots)
{
	if (in_bytes > ctx->in_cap) {
		CU(cudaDeviceSynchronize());
		for (int i = 0; i < NBUF; i++) { int rc = grow(ctx, &ctx->d_in[i], in_bytes); if (rc) return rc; }
		ctx->in_cap = in_bytes;
	}
	if (slots > ctx->out_cap) {
		CU(cudaDeviceSynchronize());
		for (int i = 0; i < NBUF; i++) {
			int rc;
			if ((rc = grow(ctx, &ctx->d_oslots[i], slots))) return rc;
			if ((rc = grow(ctx, &ctx->d_otype1[i], slots * TYPE1_STRIDE))) return rc;
			if ((rc = grow(ctx, &ctx->d_opacked[i], slots * TYPE1_WORDS))) return rc;
			if ((rc = grow(ctx, &ctx->d_ocrc[i], slots))) return rc;
			if ((rc = grow(ctx, &ctx->d_oaach[i], slots))) return rc;
		}
		ctx->out_cap = slots;
	}
	return 0;
}

/* TB200_TRACE=1: host-side time stamps of one call (stderr), for hunting host round trips */
#include <chrono>
struct HostTrace {
	bool on = getenv("TB200_TRACE") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void mark(const char *what)
	{
		if (!on) return;
		const auto t = std::chrono::steady_clock::now();
		fprintf(stderr, "[tb200 trace] %-28s +%7.1f us\n", what, std::chrono::duration<double, std::micro>(t - t0).count());
	}
};
static HostTrace *g_trace = nullptr;
#define TB_TRACE(what) do { if (g_trace) g_trace->mark(what); } while (0)

static int float_to_bits_dev(tb200_ctx *ctx, const float *d_sym, uint64_t n_sym, int afc, float filter_val, float filter_goal,
                             float *state, uint32_t *d_out);

/* ------------------------------------------------------------ the source -- */

/* Where the stream bits of this call live.  Absolute bit i of the stream is
 *   tail[i - tail_base]            for tail_base <= i < new_base   (host path only)
 *   data[i - new_base]             for new_base  <= i < end                          */
struct Source {
	bool on_device;
	const uint8_t *data;
	uint64_t new_base;
	uint64_t end;
	int fmt;               /* IN_BYTES / IN_PACKED / IN_F32SYM; the packed formats continue on 128-bit boundaries of the stream */
	/* device-resident data that is still arriving (sharded decode: chunks received or packed on other streams):
	 * stream bits below ready[i].first are in place once event ready[i].second has fired; ascending */
	const std::vector<std::pair<uint64_t, cudaEvent_t>> *ready = nullptr;
	bool skip_dependent = false;     /* sharded decode, first pass (DecodeArgs::skip_dependent) */
};

/* bytes that hold `nbits` stream bits in format fmt */
static inline size_t fmt_bytes(int fmt, uint64_t nbits)
{
	return fmt == IN_BYTES ? (size_t)nbits : fmt == IN_PACKED ? (size_t)((nbits + 7) / 8) : (size_t)(4 * ((nbits + 1) / 2));
}

/* copy stream bits [lo, hi) to device memory `dst` (host path); *dbase = stream bit that dst[0] bit 0 holds
 * (lo for the byte format, lo rounded down to a 16-byte unit of the packed formats) */
static int stage_bits(tb200_ctx *ctx, const Source &src, uint64_t lo, uint64_t hi, uint8_t *dst, cudaStream_t st, uint64_t *dbase)
{
	*dbase = lo;
	if (hi <= lo) return 0;
	if (src.fmt != IN_BYTES) {
		/* the packed formats are copied in 128-bit units of the stream; the caller's buffer of this call starts at
		 * stream bit new_base, what earlier calls left is in the tail (both on 128-bit boundaries) */
		const uint64_t lo_al = lo & ~(uint64_t)127;
		*dbase = lo_al;
		uint64_t p = lo_al;
		if (p < src.new_base) {
			const uint64_t h = std::min(hi, src.new_base);
			const size_t b0 = fmt_bytes(src.fmt, p - ctx->tail_base), b1 = fmt_bytes(src.fmt, h - ctx->tail_base);
			CU(cudaMemcpyAsync(dst, ctx->tail.data() + b0, b1 - b0, cudaMemcpyHostToDevice, st));
			dst += fmt_bytes(src.fmt, src.new_base - p);      /* (h == new_base unless the piece ends inside the tail) */
			p = src.new_base;
		}
		if (hi > p) {
			const size_t b0 = fmt_bytes(src.fmt, p - src.new_base), b1 = fmt_bytes(src.fmt, hi - src.new_base);
			CU(cudaMemcpyAsync(dst, src.data + b0, b1 - b0, cudaMemcpyHostToDevice, st));
		}
		return 0;
	}
	if (lo < src.new_base) {
		uint64_t h = std::min(hi, src.new_base);
		CU(cudaMemcpyAsync(dst, ctx->tail.data() + (lo - ctx->tail_base), h - lo, cudaMemcpyHostToDevice, st));
		dst += h - lo;
		lo = h;
	}
	if (hi > lo)
		CU(cudaMemcpyAsync(dst, src.data + (lo - src.new_base), hi - lo, cudaMemcpyHostToDevice, st));
	return 0;
}

/* -------------------------------------------------- UNLOCKED: SYNC search -- */

/* every position p in [lo, hi) where the 38-bit SYNC training sequence starts; bits are
 * readable up to `avail` bytes from `bits`; position p <-> bits[p - base] */
__global__ void __launch_bounds__(256)
k_scan_sync(const uint8_t *bits, int fmt, uint64_t base, uint64_t avail, uint64_t lo, uint64_t hi,
            const Tables *__restrict__ tab, uint32_t *hits, uint32_t cap)
{
	const unsigned lane = threadIdx.x & 31;
	const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	const uint8_t *end = bits + avail;
	for (uint64_t p0 = lo + warp * 1024; p0 < hi; p0 += nwarps * 1024) {
		uint32_t x0, x1, x2;
		if (fmt == IN_BYTES) load_window(bits + (p0 - base), end, lane, x0, x1, x2);
		else                 load_window_fmt(bits, fmt, p0 - base, avail, lane, x0, x1, x2);
		uint32_t My = 0xffffffffu;
#pragma unroll
		for (int b = 0; b < 32; ++b) {
			uint32_t s = __funnelshift_r(x0, x1, b);
			My &= ((SEQ_Y >> b) & 1) ? s : ~s;
		}
#pragma unroll
		for (int b = 32; b < 38; ++b) {
			uint32_t s = __funnelshift_r(x1, x2, b - 32);
			My &= ((SEQ_Y >> b) & 1) ? s : ~s;
		}
		const uint64_t pl = p0 + 32 * lane;
		/* keep positions < hi whose 38 bits are inside the available data */
		const uint64_t data_end = base + avail;
		const uint64_t conf = data_end >= 37 ? data_end - 37 : 0;
		const long long lim = (long long)(hi < conf ? hi : conf) - (long long)pl - 1;
		My &= low_mask(lim);
		/* the bit before each position: bit i-1 of this lane's word, bit 31 of the previous lane's for i = 0 */
		uint32_t prev_word = __shfl_up_sync(FULL, x0, 1);
		uint32_t first_prev = 0;
		if (lane == 0) {
			if (p0 > base) first_prev = fmt == IN_BYTES ? (bits[p0 - base - 1] & 1u) : (fetch32(bits, fmt, p0 - base - 1, avail) & 1u);
		} else {
			first_prev = prev_word >> 31;
		}
		const uint32_t prevbits = (x0 << 1) | first_prev;
		while (My) {
			const int i = __ffs((int)My) - 1;
			My &= My - 1;
			const uint32_t slot = atomicAdd(&hits[0], 1u);
			if (slot < cap) {
				const uint64_t pos = pl + i;
				hits[2 + 2 * slot] = (uint32_t)pos;
				hits[3 + 2 * slot] = (uint32_t)(pos >> 32) | (((prevbits >> i) & 1) << 31);
			}
		}
	}
}

#define REGION_BITS (1u << 18)
#define HIT_CAP 8192u

/* make sure the hit list covers [from, upto) as far as the data of this call allows */
static int scan_more_hits(tb200_ctx *ctx, const Source &src, uint64_t from, uint64_t upto)
{
	if (from < ctx->hits_lo || from > ctx->hits_hi) {
		ctx->hits.clear();
		ctx->hits_lo = ctx->hits_hi = from;
	}
	/* a sequence starting in the last 37 bits of the data cannot be confirmed yet */
	const uint64_t confirmable = src.end >= 37 ? src.end - 37 : 0;
	const uint64_t first_bit = src.on_device ? src.new_base : ctx->tail_base;
	while (ctx->hits_hi < upto && ctx->hits_hi < confirmable) {
		const uint64_t lo = ctx->hits_hi;
		uint64_t hi = std::min<uint64_t>(lo + REGION_BITS, confirmable);
		uint32_t n = 0;
		/* A region that holds more matches than the list (the SYNC sequence overlaps itself at shift 24: a
		 * period-24 stream gives 10 923 hits per 2^18 bits) is scanned again at half the length; 8192 bits
		 * can never overflow a list of 8192. */
		for (;;) {
			const uint8_t *dbits;
			uint64_t dbase, davail;
			if (src.on_device) {
				dbits = src.data; dbase = src.new_base; davail = src.end - src.new_base;
			} else {
				const uint64_t rd_lo = (lo > first_bit) ? lo - 1 : lo;      /* one bit back for the blind-spot rule */
				const uint64_t rd_hi = std::min<uint64_t>(hi + 64, src.end);
				const size_t need = fmt_bytes(IN_F32SYM, (uint64_t)REGION_BITS + 512);
				if (need > ctx->region_cap) {
					CU(cudaDeviceSynchronize());
					int rc = grow(ctx, &ctx->d_region, need);
					if (rc) return rc;
					ctx->region_cap = need;
				}
				int rc = stage_bits(ctx, src, rd_lo, rd_hi, ctx->d_region, ctx->s_compute, &dbase);
				if (rc) return rc;
				dbits = ctx->d_region; davail = rd_hi - dbase;
			}
			CU(cudaMemsetAsync(ctx->d_hits, 0, sizeof(uint32_t) * (2 + 2 * 64), ctx->s_compute));
			const unsigned blocks = (unsigned)std::min<uint64_t>((hi - lo + 8191) / 8192, (uint64_t)ctx->sm_count * 4);
			TB_LAUNCH(k_scan_sync, blocks, 256, ctx->s_compute, dbits, src.fmt, dbase, davail, lo, hi, ctx->d_tab, ctx->d_hits, HIT_CAP);
			ctx->stats.kernel_launches++;
			CU(cudaGetLastError());
			/* the list is short (a SYNC sequence per 18-odd bursts): fetch the count and the first entries, the rest only if needed */
			const uint32_t first = 64;
			CU(cudaMemcpyAsync(ctx->h_hits, ctx->d_hits, sizeof(uint32_t) * (2 + 2 * first), cudaMemcpyDeviceToHost, ctx->s_compute));
			CU(cudaStreamSynchronize(ctx->s_compute));
			TB_TRACE("sync-hit list on host");
			n = ctx->h_hits[0];
			if (n > HIT_CAP) {
				if (hi - lo <= HIT_CAP)
					return fail(ctx, TB200_E_STATE, "SYNC hit list overflow (%u hits in %llu bits)", n, (unsigned long long)(hi - lo));
				hi = lo + std::max<uint64_t>((hi - lo) / 2, HIT_CAP);
				continue;
			}
			if (n > first) {
				CU(cudaMemcpyAsync(ctx->h_hits + 2 + 2 * first, ctx->d_hits + 2 + 2 * first, sizeof(uint32_t) * 2 * (n - first),
				                   cudaMemcpyDeviceToHost, ctx->s_compute));
				CU(cudaStreamSynchronize(ctx->s_compute));
			}
			break;
		}
		const size_t old = ctx->hits.size();
		for (uint32_t i = 0; i < n; i++) {
			SyncHit h;
			h.pos = (uint64_t)ctx->h_hits[2 + 2 * i] | ((uint64_t)(ctx->h_hits[3 + 2 * i] & 0x7fffffffu) << 32);
			h.prev = ctx->h_hits[3 + 2 * i] >> 31;
			ctx->hits.push_back(h);
		}
		std::sort(ctx->hits.begin() + old, ctx->hits.end(), [](const SyncHit &x, const SyncHit &y) { return x.pos < y.pos; });
		ctx->hits_hi = hi;
	}
	/* forget hits that no later search can reach */
	if (ctx->hits.size() > 65536) {
		auto it = std::lower_bound(ctx->hits.begin(), ctx->hits.end(), from,
		                           [](const SyncHit &h, uint64_t v) { return h.pos < v; });
		ctx->hits.erase(ctx->hits.begin(), it);
		ctx->hits_lo = from;
	}
	return 0;
}

/* tetra_find_train_seq(bitbuf, bits_in_buf, SYNC only) as the UNLOCKED state calls it
 * (tetra_burst_sync.c:75-76): first SYNC sequence fully inside the buffer, subject to the
 * pre-filter blind spot relative to bitbuf[0] */
static bool first_sync_in_buffer(tb200_ctx *ctx, uint64_t buf_start, uint32_t bits_in_buf, uint64_t *pos)
{
	auto it = std::lower_bound(ctx->hits.begin(), ctx->hits.end(), buf_start,
	                           [](const SyncHit &h, uint64_t v) { return h.pos < v; });
	for (; it != ctx->hits.end(); ++it) {
		if (it->pos + 38 > buf_start + bits_in_buf) return false;
		uint64_t k = it->pos - buf_start;
		if (k <= 20 && !((ctx->h_tab.blind_ok[0][it->prev] >> k) & 1)) continue;
		*pos = it->pos;
		return true;
	}
	return false;
}

/* ----------------------------------------------------- LOCKED: the pieces -- */

struct Outputs {
	bool on_device;
	tb200_slot *slots;
	uint8_t *type1;
	uint32_t *packed;
	uint32_t *crc;               /* optional (tb200_set_crc_buffer) */
	uint32_t *aach = nullptr;    /* optional (tb200_set_aach_buffer) */
	uint64_t max_slots;
	uint64_t n;                  /* slots written so far */
};

struct Segment {
	uint64_t a0;                 /* absolute bit of slot 0 */
	uint64_t cmin;               /* first call that may process slot 0 */
	CallGeom cg;                 /* the modelled calls: run of equal-length reads */
};

static inline uint64_t slot_call(const Segment &s, uint64_t k)   /* call index that processes slot k */
{
	return call_for(s.cg, s.a0 + (uint64_t)SLOT_BITS * k + SLOT_BITS, s.cmin + k);
}

/* events per piece when options.profile is on */
enum { PE_START = 0, PE_SB1 = 1, PE_SCAN = 2, PE_DECODE_END = 3, PE_DECODE_START = 4, PE_SEARCH = 5, PE_PREPARE_END = 6, PE_TRELLIS_END = 7, PE_COUNT = 8 };

/* pass 1 of a piece on stream s_front: search + classification, SB1, the scan over "last CRC-good SB1" and - unless
 * the cell state in front of the piece is not known yet (sharded decode) - the carry for the next piece: the state
 * after a piece follows from pass 1 alone (tetra_lower_mac.c:291-302: only SB1 results change it), so the carry
 * chain runs ahead of the decode pass. */
static int enqueue_pass1(tb200_ctx *ctx, const RxGeom &g, size_t piece_idx, int set, cudaEvent_t *pe, bool with_carry, uint64_t k_base = 0)
{
	cudaStream_t st = ctx->opt.serial_passes ? ctx->s_compute : ctx->s_front;
	tb200_ctx::WorkSet &w = ctx->wset[set];
	const uint32_t nb = g.n_slots;
	const unsigned wpb = 8;
	const unsigned blocks = (unsigned)std::min<uint64_t>((nb + wpb - 1) / wpb, (uint64_t)ctx->sm_count * 16);
	if (ctx->back_pending[set]) CU(cudaStreamWaitEvent(st, ctx->ev_back[set], 0));     /* the set's previous piece has been decoded */
	CU(cudaMemsetAsync(ctx->d_flags + 2 * piece_idx, 0xff, 2 * sizeof(uint32_t), st));      /* first lock loss, first CRC-good SB1 */
	if (pe) CU(cudaEventRecord(pe[PE_START], st));
	const bool lane = ctx->opt.viterbi == TB200_VITERBI_LANE;
	const unsigned lane_nt = LANE_NT;
	const size_t lane_smem = lane_smem_words(lane_nt) * sizeof(uint32_t);
	if (lane) {
		WinGeom wg;
		wg.cg = g.cg; wg.cmin = g.cmin; wg.a0 = g.a0;
		uint32_t *sb_count = w.sb_list + ctx->ws_slots;
		CU(cudaMemsetAsync(sb_count, 0, sizeof(uint32_t), st));
		const unsigned per_tile = g.fmt == IN_F32SYM ? TileFmt<IN_F32SYM>::SLOTS : CT_SLOTS;
		const unsigned tiles = (nb + per_tile - 1) / per_tile;
		unsigned cls_per_sm = g.fmt == IN_PACKED ? 8 : 3;
		if (const char *e = getenv("TB200_CLS_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 8) cls_per_sm = (unsigned)v; }
		const unsigned cls_blocks = std::min<unsigned>(tiles, (unsigned)ctx->sm_count * cls_per_sm);
		if (g.fmt == IN_BYTES)
			TB_LAUNCH_SMEM(k_classify_tile<IN_BYTES>, cls_blocks, CT_THREADS, ct_smem<IN_BYTES>(), st, g, wg, ctx->d_tab, w.ws,
			               w.slot_bits, w.sb_list, sb_count);
		else if (g.fmt == IN_PACKED)
			TB_LAUNCH_SMEM(k_classify_tile<IN_PACKED>, cls_blocks, CT_THREADS, ct_smem<IN_PACKED>(), st, g, wg, ctx->d_tab, w.ws,
			               w.slot_bits, w.sb_list, sb_count);
		else
			TB_LAUNCH_SMEM(k_classify_tile<IN_F32SYM>, cls_blocks, CT_THREADS, ct_smem<IN_F32SYM>(), st, g, wg, ctx->d_tab, w.ws,
			               w.slot_bits, w.sb_list, sb_count);
		if (pe) CU(cudaEventRecord(pe[PE_SEARCH], st));
		/* at most one SYNC burst per slot, two per thread; the kernel reads the real count from sb_count */
		const uint64_t npairs = ((uint64_t)nb + 1) / 2;
		const unsigned sb1_blocks = (unsigned)std::min<uint64_t>((npairs + lane_nt - 1) / lane_nt, (uint64_t)ctx->lane_ctas);
		if (ctx->opt.viterbi_tie)
			TB_LAUNCH_SMEM(k_sb1_lane<true>, sb1_blocks, lane_nt, lane_smem, st, w.ws, w.slot_bits, w.sb_list, sb_count,
			               ctx->d_tab, ctx->d_sb1_scratch);
		else
			TB_LAUNCH_SMEM(k_sb1_lane<false>, sb1_blocks, lane_nt, lane_smem, st, w.ws, w.slot_bits, w.sb_list, sb_count,
			               ctx->d_tab, ctx->d_sb1_scratch);
		ctx->stats.kernel_launches++;
	} else {
		TB_LAUNCH(k_classify<true>, blocks, 256, st, g, ctx->d_tab, w.ws, w.slot_bits);
		if (pe) CU(cudaEventRecord(pe[PE_SEARCH], st));
	}
	if (pe) CU(cudaEventRecord(pe[PE_SB1], st));
	const unsigned nblk = (nb + 1023) / 1024;
	uint32_t *kind_count = w.kind_list + 4 * ctx->ws_slots;
	CU(cudaMemsetAsync(kind_count, 0, 4 * sizeof(uint32_t), st));
	TB_LAUNCH(k_scan_blocks, nblk, SCAN_THREADS, st, w.ws, nb, w.last_good, w.blk_last, ctx->d_flags + 2 * piece_idx, ctx->d_flags + 2 * piece_idx + 1,
	          kind_count, w.kind_list, (uint32_t)ctx->ws_slots);
	TB_LAUNCH(k_scan_prefix, 1, 1024, st, w.blk_last, nblk, w.blk_prev);
	ctx->stats.kernel_launches += 3;
	if (with_carry) {
		CU(cudaMemcpyAsync(ctx->d_carry + piece_idx + 1, ctx->d_carry + piece_idx, sizeof(DevCarry), cudaMemcpyDeviceToDevice, st));
		TB_LAUNCH(k_finalize_carry, 1, 32, st, w.ws, w.last_good, w.blk_prev, nb, ctx->d_carry + piece_idx + 1,
		          ctx->d_flags + 2 * piece_idx + 1, k_base);
		ctx->stats.kernel_launches++;
	}
	if (pe) CU(cudaEventRecord(pe[PE_SCAN], st));
	CU(cudaEventRecord(ctx->ev_front[set], st));
	CU(cudaGetLastError());
	return 0;
}

/* pass 2 of a piece on stream s_compute: everything that needs the cell state carried in d_carry[piece_idx] */
static int enqueue_pass2(tb200_ctx *ctx, uint64_t a0, uint32_t nb, size_t piece_idx, int set, cudaEvent_t *pe, bool with_carry,
                         SlotOut *o_slots, uint8_t *o_type1, uint32_t *o_packed, uint64_t out_base, uint32_t *o_crc = nullptr,
                         bool skip_dependent = false, uint32_t *o_aach = nullptr)
{
	cudaStream_t st = ctx->s_compute;
	tb200_ctx::WorkSet &w = ctx->wset[set];
	const bool lane = ctx->opt.viterbi == TB200_VITERBI_LANE;
	const unsigned lane_nt = LANE_NT;
	const size_t lane_smem = lane_smem_words(lane_nt) * sizeof(uint32_t);
	const uint64_t npairs = ((uint64_t)nb + 1) / 2;
	const unsigned lane_blocks = (unsigned)std::min<uint64_t>((npairs + lane_nt - 1) / lane_nt, (uint64_t)ctx->form_ctas[0]);
	const unsigned blocks = (unsigned)std::min<uint64_t>((nb + 7) / 8, (uint64_t)ctx->sm_count * 16);
	DecodeArgs a;
	a.ws = w.ws; a.slot_bits = w.slot_bits; a.last_good = w.last_good; a.blk_prev = w.blk_prev;
	a.carry = ctx->d_carry + piece_idx; a.tab = ctx->d_tab;
	a.slots = o_slots;
	a.type1 = (ctx->opt.output & TB200_OUT_UNPACKED) ? o_type1 : nullptr;
	a.type1_packed = (ctx->opt.output & TB200_OUT_PACKED) ? o_packed : nullptr;
	a.a0 = a0; a.out_base = out_base; a.n_slots = nb;
	a.kind_count = w.kind_list + 4 * ctx->ws_slots; a.kind_list = w.kind_list; a.list_stride = (uint32_t)ctx->ws_slots;
	a.crc = o_crc;
	a.tie_hi = (int)ctx->opt.viterbi_tie;
	a.stats = ctx->d_pstats + 3 * piece_idx;
	a.skip_dependent = skip_dependent ? 1 : 0;
	a.aach = o_aach; a.rm_leader = ctx->d_rm_leader;
	CU(cudaStreamWaitEvent(st, ctx->ev_front[set], 0));
	CU(cudaMemsetAsync(a.stats, 0, 3 * sizeof(unsigned long long), st));
	if (pe) CU(cudaEventRecord(pe[PE_DECODE_START], st));
	if (pe && !(lane && ctx->lane_form != 0)) { CU(cudaEventRecord(pe[PE_PREPARE_END], st)); CU(cudaEventRecord(pe[PE_TRELLIS_END], st)); }
	if (lane && ctx->lane_form == 0) {
		if (ctx->opt.viterbi_tie) TB_LAUNCH_SMEM(k_decode_lane<true>, lane_blocks, lane_nt, lane_smem, st, a, ctx->d_lane_scratch);
		else                      TB_LAUNCH_SMEM(k_decode_lane<false>, lane_blocks, lane_nt, lane_smem, st, a, ctx->d_lane_scratch);
	} else if (lane) {
		/* split form: the unit count is only known on the device (kind_count), so the grids are sized for the worst case
		 * (every slot a unit of its own) and capped at what is resident */
		const uint64_t wmax = lane_unit_blocks(nb);
		const unsigned pblocks = (unsigned)std::min<uint64_t>(2 * wmax, (uint64_t)ctx->prep_ctas);      /* a task = one slot of each unit of a block */
		const unsigned tblocks = (unsigned)std::min<uint64_t>(wmax, (uint64_t)ctx->form_ctas[ctx->lane_form]);
		const unsigned fblocks = (unsigned)std::min<uint64_t>(wmax, (uint64_t)ctx->fin_ctas);
		TB_LAUNCH_SMEM(k_lane_prepare, pblocks, lane_nt, lane_prepare_smem_words() * sizeof(uint32_t), st, a, ctx->d_units);
		if (pe) CU(cudaEventRecord(pe[PE_PREPARE_END], st));
		if (ctx->lane_form == 1) {
			const size_t sm_b = lane_trellis_smem_words<true>() * sizeof(uint32_t);
			if (ctx->opt.viterbi_tie) TB_LAUNCH_SMEM((k_lane_trellis<true, true>), tblocks, lane_nt, sm_b, st, a, ctx->d_lane_scratch, ctx->d_units);
			else                      TB_LAUNCH_SMEM((k_lane_trellis<false, true>), tblocks, lane_nt, sm_b, st, a, ctx->d_lane_scratch, ctx->d_units);
			if (pe) CU(cudaEventRecord(pe[PE_TRELLIS_END], st));
			ctx->stats.kernel_launches += 1;
		} else {
			const size_t sm_b = lane_trellis_smem_words<false>() * sizeof(uint32_t);
			if (ctx->opt.viterbi_tie) TB_LAUNCH_SMEM((k_lane_trellis<true, false>), tblocks, lane_nt, sm_b, st, a, ctx->d_lane_scratch, ctx->d_units);
			else                      TB_LAUNCH_SMEM((k_lane_trellis<false, false>), tblocks, lane_nt, sm_b, st, a, ctx->d_lane_scratch, ctx->d_units);
			if (pe) CU(cudaEventRecord(pe[PE_TRELLIS_END], st));
			TB_LAUNCH_SMEM(k_lane_finish, fblocks, lane_nt, lane_finish_smem_words() * sizeof(uint32_t), st, a, ctx->d_units);
			ctx->stats.kernel_launches += 2;
		}
	} else {
		TB_LAUNCH(k_decode_warp, blocks, 256, st, a);
	}
	ctx->stats.kernel_launches++;
	if (pe) CU(cudaEventRecord(pe[PE_DECODE_END], st));
	if (!with_carry) {
		CU(cudaMemcpyAsync(ctx->d_carry + piece_idx + 1, ctx->d_carry + piece_idx, sizeof(DevCarry), cudaMemcpyDeviceToDevice, st));
		TB_LAUNCH(k_finalize_carry, 1, 32, st, w.ws, w.last_good, w.blk_prev, nb, ctx->d_carry + piece_idx + 1,
		          ctx->d_flags + 2 * piece_idx + 1, (uint64_t)0);
		ctx->stats.kernel_launches++;
	}
	CU(cudaEventRecord(ctx->ev_back[set], st));
	ctx->back_pending[set] = true;
	CU(cudaGetLastError());
	return 0;
}

/* enqueue classify + scan + carry (s_front) and decode (s_compute) for slots [k0, k0+nb) of the segment */
static int enqueue_piece(tb200_ctx *ctx, const Segment &seg, uint64_t k0, uint32_t nb, const uint8_t *d_bits,
                         uint64_t d_base, uint64_t d_avail, int fmt, size_t piece_idx,
                         SlotOut *o_slots, uint8_t *o_type1, uint32_t *o_packed, uint64_t out_base, uint32_t *o_crc,
                         bool skip_dependent, uint32_t *o_aach)
{
	RxGeom g;
	g.bits = d_bits; g.n_bytes = d_avail; g.base_bit = d_base; g.fmt = fmt;
	g.a0 = seg.a0 + (uint64_t)SLOT_BITS * k0; g.cmin = seg.cmin + k0; g.cg = seg.cg;
	g.n_slots = nb; g.tie_hi = (int)ctx->opt.viterbi_tie;
	cudaEvent_t *pe = nullptr;
	if (ctx->opt.profile) {
		while (ctx->prof_ev.size() < ctx->prof_used + PE_COUNT) {
			cudaEvent_t e;
			CU(cudaEventCreateWithFlags(&e, 0));
			ctx->prof_ev.push_back(e);
		}
		pe = &ctx->prof_ev[ctx->prof_used];
		ctx->prof_used += PE_COUNT;
		ctx->timing.pieces++;
		ctx->timing.slots += nb;
	}
	const int set = (int)(piece_idx & 1);
	int rc = enqueue_pass1(ctx, g, piece_idx, set, pe, true, k0);
	if (rc) return rc;
	return enqueue_pass2(ctx, g.a0, nb, piece_idx, set, pe, true, o_slots, o_type1, o_packed, out_base, o_crc, skip_dependent, o_aach);
}

/* Process slots [0, n_slots) of a LOCKED segment, optimistically assuming lock is kept;
 * the first piece that reports a lock loss is redone up to and including the losing slot.
 * *valid = slots consumed, *lost = whether the last one lost lock. */
static int run_locked(tb200_ctx *ctx, const Source &src, const Segment &seg, uint64_t n_slots,
                      Outputs &out, uint64_t *valid, bool *lost)
{
	*valid = 0; *lost = false;
	if (n_slots == 0) return 0;
	if (out.n + n_slots > out.max_slots)
		return fail(ctx, TB200_E_ARG, "output arrays too small: need %llu slots, have %llu",
		            (unsigned long long)(out.n + n_slots), (unsigned long long)out.max_slots);
	/* slots per piece.  Device-resident input 2^21: pass 1 of piece i+1 (stream s_front) fills the SMs that the last
	 * round of piece i's decode pass (stream s_compute) leaves idle; the two passes do not share an SM well (the
	 * decode grid holds every register, and the search pushes the survivor histories out of L2), so smaller pieces
	 * only add round-quantisation loss (measured at 8*10^6 SCH/F bursts: 303 104 slots 1.50, 2^20 1.69, 2^21 1.74 *10^9
	 * bursts/s; everything on one stream 1.64).  Host input 2^17 with a ramp (below), bit-packed host input 2^18 without
	 * one: its copies are 8x smaller, so the per-piece launch and synchronisation cost weighs more (measured: 10^6 bursts
	 * packed in and out 2.36 -> 2.00 ms) */
	const bool packed_host = !src.on_device && src.fmt == IN_PACKED;
	uint32_t dev_piece = 1u << 21;
	if (const char *e = getenv("TB200_PIECE_SLOTS")) { const long v = atol(e); if (v >= 1024 && v <= (1 << 24)) dev_piece = (uint32_t)v; }
	uint32_t P = ctx->opt.pipeline_slots ? ctx->opt.pipeline_slots : (src.on_device ? dev_piece : packed_host ? (1u << 18) : (1u << 17));
	if (P > n_slots) P = (uint32_t)n_slots;
	/* piece boundaries; the host path ramps the first pieces up (P/16, P/8, ...) so that the first
	 * copy is short and compute / copy-back start early */
	std::vector<uint64_t> pstart;
	{
		uint64_t k = 0;
		uint32_t cur = (!src.on_device && !packed_host && !ctx->opt.pipeline_slots && P >= 16384) ? P / 16 : P;
		while (k < n_slots) {
			pstart.push_back(k);
			k += cur;
			if (cur < P) cur = std::min<uint32_t>(P, cur * 2);
		}
		pstart.push_back(n_slots);
	}
	const size_t npieces = pstart.size() - 1;
	int rc;
	if ((rc = ensure_workspace(ctx, P))) return rc;
	if ((rc = ensure_pieces(ctx, npieces))) return rc;
	/* bits a piece may touch: its slots plus the largest search window (<= 4096) */
	const size_t piece_in = fmt_bytes(src.fmt, (uint64_t)P * SLOT_BITS + 4096 + 64 + 128) + 64;
	if (!src.on_device && (rc = ensure_staging(ctx, piece_in, P))) return rc;
	const bool host_pack = !src.on_device && src.fmt == IN_BYTES && ctx->opt.host_pack_threads > 0;
	if (host_pack) {
		const size_t need = ((size_t)P * SLOT_BITS + 4096 + 64 + 256) / 8 + 64;
		if (need > ctx->h_pack_cap) {
			CU(cudaDeviceSynchronize());
			for (int i = 0; i < NBUF; i++) {
				if (ctx->h_pack[i]) cudaFreeHost(ctx->h_pack[i]);
				ctx->h_pack[i] = nullptr;
				CU(cudaHostAlloc((void **)&ctx->h_pack[i], need, cudaHostAllocDefault));
			}
			ctx->h_pack_cap = need;
		}
		if (ctx->pack_inflight >= 0) { ctx->pack_pool.wait(); ctx->pack_inflight = -1; }
		ctx->pack_pool.resize(ctx->opt.host_pack_threads > 1 ? ctx->opt.host_pack_threads : 0);      /* 1: the calling thread packs */
	}
	/* every TB200_HOST_PACK_EVERY-th piece goes through the host threads, the others cross the bus as bytes (both roads busy) */
	unsigned pack_every = 1;
	if (const char *e = getenv("TB200_HOST_PACK_EVERY")) { const int v = atoi(e); if (v >= 1 && v <= 64) pack_every = (unsigned)v; }
	auto pack_this = [&](size_t i) { return (i % pack_every) == 0; };
	const bool host_out = !out.on_device;
	if (src.on_device && host_out)
		return fail(ctx, TB200_E_ARG, "device input with host output is not supported");

	auto piece_range = [&](size_t i, uint64_t *k0, uint32_t *nb) {
		*k0 = pstart[i];
		*nb = (uint32_t)(pstart[i + 1] - pstart[i]);
	};
	auto issue = [&](size_t i, uint32_t nb_override) -> int {
		uint64_t k0; uint32_t nb;
		piece_range(i, &k0, &nb);
		if (nb_override) nb = nb_override;
		const int b = (int)(i % NBUF);
		const uint64_t lo = seg.a0 + (uint64_t)SLOT_BITS * k0;
		const uint64_t hi = std::min<uint64_t>(seg.cg.n_end, lo + (uint64_t)SLOT_BITS * nb + 4096);
		const uint8_t *dbits; uint64_t dbase, davail;
		int piece_fmt = src.fmt;
		if (src.on_device) {
			dbits = src.data; dbase = src.new_base; davail = std::min<uint64_t>(seg.cg.n_end, src.end) - src.new_base;
			if (src.ready) {
				for (const auto &rv : *src.ready)
					if (rv.first >= hi || &rv == &src.ready->back()) { CU(cudaStreamWaitEvent(ctx->s_front, rv.second, 0)); break; }
			}
		} else if (host_pack && pack_this(i) && (lo & ~(uint64_t)127) >= src.new_base) {
			/* the piece lies in this call's buffer: the host threads pack it (from a 128-bit boundary of the stream on, as
			 * the packed format is staged), 8x fewer bytes cross the bus, the search kernel takes it as TB200_IN_PACKED.
			 * h_pack[b] is free: the piece that used it last has been copied back (ev_d2h, waited for by the caller's loop) */
			dbase = lo & ~(uint64_t)127;
			const size_t nbits_p = (size_t)(hi - dbase);
			if (ctx->pack_inflight == (long)i && ctx->pack_inflight_base == dbase && ctx->pack_inflight_hi == hi) {
				ctx->pack_pool.wait();                   /* started when the previous piece was issued */
			} else {
				if (ctx->pack_inflight >= 0) ctx->pack_pool.wait();
				ctx->pack_pool.run(src.data + (dbase - src.new_base), nbits_p, ctx->h_pack[b]);
			}
			ctx->pack_inflight = -1;
			CU(cudaMemcpyAsync(ctx->d_in[b], ctx->h_pack[b], ((nbits_p + 7) >> 3), cudaMemcpyHostToDevice, ctx->s_h2d));
			CU(cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
			CU(cudaStreamWaitEvent(ctx->s_front, ctx->ev_h2d[b], 0));
			dbits = ctx->d_in[b]; davail = hi - dbase; piece_fmt = IN_PACKED;
		} else {
			int r = stage_bits(ctx, src, lo, hi, ctx->d_in[b], ctx->s_h2d, &dbase);
			if (r) return r;
			CU(cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
			CU(cudaStreamWaitEvent(ctx->s_front, ctx->ev_h2d[b], 0));
			dbits = ctx->d_in[b]; davail = hi - dbase;
		}
		SlotOut *os; uint8_t *ot; uint32_t *op; uint64_t ob; uint32_t *oc, *oa;
		if (out.on_device) {
			os = (SlotOut *)out.slots; ot = out.type1; op = out.packed; ob = out.n + k0; oc = out.crc; oa = out.aach;
		} else {
			os = ctx->d_oslots[b]; ot = ctx->d_otype1[b]; op = ctx->d_opacked[b]; ob = 0; oc = out.crc ? ctx->d_ocrc[b] : nullptr;
			oa = out.aach ? ctx->d_oaach[b] : nullptr;
		}
		int r = enqueue_piece(ctx, seg, k0, nb, dbits, dbase, davail, piece_fmt, i, os, ot, op, ob, oc, src.skip_dependent, oa);
		if (r) return r;
		CU(cudaEventRecord(ctx->ev_comp[b], ctx->s_compute));
		cudaStream_t so = host_out ? ctx->s_d2h : ctx->s_compute;
		if (host_out) {
			CU(cudaStreamWaitEvent(so, ctx->ev_comp[b], 0));
			const uint64_t o0 = out.n + k0;
			CU(cudaMemcpyAsync(out.slots + o0, os, (size_t)nb * sizeof(SlotOut), cudaMemcpyDeviceToHost, so));
			if (out.type1 && (ctx->opt.output & TB200_OUT_UNPACKED))
				CU(cudaMemcpyAsync(out.type1 + o0 * TYPE1_STRIDE, ot, (size_t)nb * TYPE1_STRIDE, cudaMemcpyDeviceToHost, so));
			if (out.packed && (ctx->opt.output & TB200_OUT_PACKED))
				CU(cudaMemcpyAsync(out.packed + o0 * TYPE1_WORDS, op, (size_t)nb * TYPE1_WORDS * 4, cudaMemcpyDeviceToHost, so));
			if (out.crc)
				CU(cudaMemcpyAsync(out.crc + o0, oc, (size_t)nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, so));
			if (out.aach)
				CU(cudaMemcpyAsync(out.aach + o0, oa, (size_t)nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, so));
		}
		CU(cudaMemcpyAsync(ctx->h_flags + 2 * i, ctx->d_flags + 2 * i, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, so));
		CU(cudaMemcpyAsync(ctx->h_pstats + 3 * i, ctx->d_pstats + 3 * i, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, so));
		CU(cudaEventRecord(ctx->ev_d2h[b], so));
		/* the next piece's bytes are packed while this one is copied and decoded (its staging buffer is free: the piece that
		 * used it last, i + 1 - NBUF, was waited for before this one was issued) */
		if (host_pack && !nb_override && i + 1 < npieces && pack_this(i + 1)) {
			uint64_t k1; uint32_t nb1;
			piece_range(i + 1, &k1, &nb1);
			const uint64_t lo1 = seg.a0 + (uint64_t)SLOT_BITS * k1;
			const uint64_t hi1 = std::min<uint64_t>(seg.cg.n_end, lo1 + (uint64_t)SLOT_BITS * nb1 + 4096);
			const uint64_t base1 = lo1 & ~(uint64_t)127;
			if (base1 >= src.new_base) {
				ctx->pack_pool.start(src.data + (base1 - src.new_base), (size_t)(hi1 - base1), ctx->h_pack[(i + 1) % NBUF]);
				ctx->pack_inflight = (long)(i + 1); ctx->pack_inflight_base = base1; ctx->pack_inflight_hi = hi1;
			}
		}
		return 0;
	};

	size_t bad_piece = npieces;
	for (size_t i = 0; i < npieces; i++) {
		if (i >= NBUF) CU(cudaEventSynchronize(ctx->ev_d2h[i % NBUF]));
		if ((rc = issue(i, 0))) return rc;
		if (i >= 1) {
			CU(cudaEventSynchronize(ctx->ev_d2h[(i - 1) % NBUF]));
			if (ctx->h_flags[2 * (i - 1)] != 0xffffffffu) { bad_piece = i - 1; break; }
		}
	}
	/* the state after the last valid slot becomes the start of the chain again: optimistically from the last
	 * piece, enqueued behind its kernels so that it costs no extra round trip */
	auto carry_over = [&](size_t last) -> int {
		/* on s_compute: behind the decode pass of the last piece, which still reads the head of the chain when the
		 * call has a single piece (and which waited for the front stream, where the chain is written) */
		CU(cudaMemcpyAsync(&ctx->h_carry_pin[1], ctx->d_carry + last + 1, sizeof(DevCarry), cudaMemcpyDeviceToHost, ctx->s_compute));
		CU(cudaMemcpyAsync(ctx->d_carry, ctx->d_carry + last + 1, sizeof(DevCarry), cudaMemcpyDeviceToDevice, ctx->s_compute));
		return 0;
	};
	if (bad_piece == npieces && (rc = carry_over(npieces - 1))) return rc;
	TB_TRACE("pieces enqueued");
	CU(cudaStreamSynchronize(ctx->s_h2d));
	CU(cudaStreamSynchronize(ctx->s_front));
	CU(cudaStreamSynchronize(ctx->s_compute));
	CU(cudaStreamSynchronize(ctx->s_d2h));
	TB_TRACE("pieces done");
	if (bad_piece == npieces && ctx->h_flags[2 * (npieces - 1)] != 0xffffffffu) bad_piece = npieces - 1;

	if (bad_piece < npieces) {
		/* redo the piece that lost lock, cut right after the losing slot, so that outputs and
		 * the carried cell state stop exactly where the reference's LOCKED state stops */
		const uint32_t u = ctx->h_flags[2 * bad_piece];
		uint64_t k0; uint32_t nb;
		piece_range(bad_piece, &k0, &nb);
		if (bad_piece == 0) {
			/* the chain start was overwritten by the optimistic carry-over: put it back */
			ctx->h_carry_pin[0] = ctx->h_carry;
			CU(cudaMemcpyAsync(ctx->d_carry, &ctx->h_carry_pin[0], sizeof(DevCarry), cudaMemcpyHostToDevice, ctx->s_front));
		}
		if ((rc = issue(bad_piece, u + 1))) return rc;
		if ((rc = carry_over(bad_piece))) return rc;
		CU(cudaStreamSynchronize(ctx->s_h2d));
		CU(cudaStreamSynchronize(ctx->s_front));
		CU(cudaStreamSynchronize(ctx->s_compute));
		CU(cudaStreamSynchronize(ctx->s_d2h));
		*valid = k0 + u + 1;
		*lost = true;
	} else {
		*valid = n_slots;
	}
	ctx->h_carry = ctx->h_carry_pin[1];
	out.n += *valid;
	/* counters of the pieces that count (a piece enqueued behind the one that lost lock does not) */
	for (size_t i = 0; i < npieces && i <= bad_piece; i++) {
		ctx->stats.bursts_decoded += ctx->h_pstats[3 * i];
		ctx->stats.blocks += ctx->h_pstats[3 * i + 1];
		ctx->stats.crc_ok_blocks += ctx->h_pstats[3 * i + 2];
	}
	return 0;
}

/* ------------------------------------------------------- the state machine -- */

/* the LOCKED run that starts at the receiver's present position: where its slots sit and how many the modelled
 * calls up to c_max can process (tetra_burst_sync.c:107-150, one slot per call) */
static uint64_t locked_extent(const RxHost &rx, const CallGeom &cg, uint64_t c_max, Segment *seg)
{
	seg->a0 = rx.buf_start; seg->cmin = rx.calls + 1; seg->cg = cg;
	if (cg.n_end < seg->a0 + SLOT_BITS || c_max < seg->cmin) return 0;
	const uint64_t by_bits = (cg.n_end - SLOT_BITS - seg->a0) / SLOT_BITS + 1;
	const uint64_t by_calls = c_max - seg->cmin + 1;
	return std::min(by_bits, by_calls);
}

/* receiver state after `valid` slots of that run were consumed, the last of which lost lock if `lost` */
static void locked_advance(tb200_ctx *ctx, const Segment &seg, uint64_t valid, bool lost)
{
	RxHost &rx = ctx->rx;
	ctx->stats.slots += valid;
	const uint64_t c_last = slot_call(seg, valid - 1);
	rx.calls = c_last;
	rx.buf_start = seg.a0 + (uint64_t)SLOT_BITS * valid;
	rx.bits_in_buf = (uint32_t)(bits_at_call(seg.cg, c_last) - rx.buf_start);
	rx.next_frame_start += (uint64_t)SLOT_BITS * valid;
	if (lost) {
		rx.state = TB200_RX_UNLOCKED;
		ctx->stats.lock_losses++;
	}
}

/* the calls of this run: everything the stream holds beyond what earlier runs delivered, `chunk` bits per call; the
 * last call of a FINAL run may be short (read() at EOF), a non-final run ends on a call boundary */
static CallGeom run_geometry(const tb200_ctx *ctx, uint64_t total, bool final, uint64_t *c_max)
{
	const RxHost &rx = ctx->rx;
	CallGeom cg;
	cg.chunk = ctx->opt.chunk_bits; cg.pad = 0;
	cg.c_base = rx.c_base; cg.t_base = rx.t_base;
	const uint64_t fresh = total - rx.t_base;
	const uint64_t ncalls = final ? (fresh + cg.chunk - 1) / cg.chunk : fresh / cg.chunk;
	*c_max = rx.c_base + ncalls;
	cg.n_end = final ? total : rx.t_base + ncalls * cg.chunk;
	return cg;
}

static int rx_run(tb200_ctx *ctx, const Source &src, bool final, Outputs &out, bool same_run = false)
{
	const uint64_t total = src.end;
	RxHost &rx = ctx->rx;
	/* a new run starts where the calls modelled so far ended (same_run: the caller comes back into the run it left
	 * with stop_at_lock) */
	if (!same_run) { rx.c_base = rx.calls; rx.t_base = rx.delivered; }
	uint64_t c_max = 0;
	const CallGeom cg = run_geometry(ctx, total, final, &c_max);
	const uint64_t n_end = cg.n_end;
	auto T = [&](uint64_t c) { return bits_at_call(cg, c); };
	int rc;
	struct Delivered { RxHost &rx; uint64_t n_end; ~Delivered() { rx.delivered = n_end; } } mark_delivered{rx, n_end};

	while (rx.calls < c_max) {
		if (rx.state == TB200_RX_LOCKED) {
			if (ctx->stop_at_lock) return 0;
			Segment seg;
			const uint64_t n_slots = locked_extent(rx, cg, c_max, &seg);
			if (n_slots == 0) {
				rx.calls = c_max;
				rx.bits_in_buf = (uint32_t)(T(c_max) - rx.buf_start);
				break;
			}
			uint64_t valid = 0; bool lost = false;
			if ((rc = run_locked(ctx, src, seg, n_slots, out, &valid, &lost))) return rc;
			locked_advance(ctx, seg, valid, lost);
			continue;
		}
		/* UNLOCKED / KNOW_FSTART: one modelled call at a time (tetra_burst_sync.c:60-106) */
		const uint64_t c = rx.calls + 1;
		const uint64_t t_prev = T(rx.calls), t_now = T(c);
		uint64_t bib = (uint64_t)rx.bits_in_buf + (t_now - t_prev);
		if (bib > 4096) { rx.buf_start += bib - 4096; bib = 4096; }   /* make_bitbuf_space, :38-51 */
		rx.bits_in_buf = (uint32_t)bib;
		rx.calls = c;
		if (rx.state == TB200_RX_UNLOCKED) {
			if (bib < 2 * SLOT_BITS) continue;
			if ((rc = scan_more_hits(ctx, src, rx.buf_start, rx.buf_start + bib))) return rc;
			uint64_t pos;
			if (!first_sync_in_buffer(ctx, rx.buf_start, rx.bits_in_buf, &pos)) continue;
			rx.state = TB200_RX_KNOW_FSTART;
			rx.next_frame_start = pos + 296;
			ctx->stats.lock_acquisitions++;
			{
				tb200_lock_event le;
				le.next_slot = ctx->stop_at_lock ? ctx->stats.slots : out.n;      /* (sharded decode: rank 0 counts the slots of all ranks) */
				le.call = c; le.offset = (uint32_t)(pos - rx.buf_start); le.pad = 0;
				ctx->lock_events.push_back(le);
			}
			continue;
		}
		/* KNOW_FSTART */
		if (rx.buf_start + rx.bits_in_buf < rx.next_frame_start) continue;
		const uint64_t shift = rx.next_frame_start - rx.buf_start;
		rx.bits_in_buf -= (uint32_t)shift;
		rx.buf_start = rx.next_frame_start;
		rx.next_frame_start += SLOT_BITS;
		rx.state = TB200_RX_LOCKED;
		rx.calls = c - 1;            /* the LOCKED arm runs in this same call (fall-through, :105-107) */
	}
	return 0;
}

#ifdef TB_SIMT_EMULATION
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
#endif

static void profile_begin(tb200_ctx *ctx)
{
	ctx->prof_used = 0;
	memset(&ctx->timing, 0, sizeof(ctx->timing));
}

static int profile_end(tb200_ctx *ctx)
{
	if (!ctx->opt.profile || ctx->prof_used == 0) return 0;
	CU(cudaStreamSynchronize(ctx->s_compute));
	float ms = 0.f;
	for (size_t i = 0; i + PE_COUNT <= ctx->prof_used; i += PE_COUNT) {
		cudaEvent_t *e = &ctx->prof_ev[i];
		CU(cudaEventElapsedTime(&ms, e[PE_START], e[PE_SB1])); ctx->timing.classify_ms += ms; ctx->timing.launches_classify++;
		CU(cudaEventElapsedTime(&ms, e[PE_START], e[PE_SEARCH])); ctx->timing.search_ms += ms;
		CU(cudaEventElapsedTime(&ms, e[PE_SB1], e[PE_SCAN])); ctx->timing.scan_ms += ms; ctx->timing.launches_scan += 3;
		CU(cudaEventElapsedTime(&ms, e[PE_DECODE_START], e[PE_DECODE_END])); ctx->timing.decode_ms += ms; ctx->timing.launches_decode++;
		CU(cudaEventElapsedTime(&ms, e[PE_DECODE_START], e[PE_PREPARE_END])); ctx->timing.prepare_ms += ms;
		CU(cudaEventElapsedTime(&ms, e[PE_PREPARE_END], e[PE_TRELLIS_END])); ctx->timing.trellis_ms += ms;
	}
	/* first launch of the call (pass 1 of the first piece) to the end of the last decode pass */
	CU(cudaEventElapsedTime(&ms, ctx->prof_ev[PE_START], ctx->prof_ev[ctx->prof_used - PE_COUNT + PE_DECODE_END]));
	ctx->timing.total_ms = ms;
	return 0;
}

static void reset_stream(tb200_ctx *ctx)
{
	ctx->rx = RxHost();
	ctx->tail.clear();
	ctx->d_tail_bytes = 0; ctx->tail_dev = false;
	ctx->tail_base = 0;
	ctx->fed_end = 0;
	ctx->hits.clear();
	ctx->hits_lo = ctx->hits_hi = 0;
	memset(&ctx->h_carry, 0, sizeof(ctx->h_carry));
	memset(&ctx->stats, 0, sizeof(ctx->stats));
}

static int push_carry(tb200_ctx *ctx)
{
	int rc = ensure_pieces(ctx, 1);
	if (rc) return rc;
	/* ordered with the kernels: async copy on the compute stream from pinned staging (a plain cudaMemcpy from
	 * pageable memory runs on the legacy stream, which our non-blocking streams do not wait for) */
	CU(cudaStreamSynchronize(ctx->s_compute));
	CU(cudaStreamSynchronize(ctx->s_front));
	ctx->h_carry_pin[0] = ctx->h_carry;
	/* on s_front: pass 1 of the first piece reads it there, the decode pass waits for pass 1 */
	CU(cudaMemcpyAsync(ctx->d_carry, &ctx->h_carry_pin[0], sizeof(DevCarry), cudaMemcpyHostToDevice, ctx->s_front));
	return 0;
}

/* bits of a device-resident stream a continuing call looks at first: what a receiver in any state can still have in its
 * buffer (4096) plus a slot and the search's look-ahead, generously */
#define DEV_CONT_HEAD_BITS (1u << 15)

static int dev_reserve(tb200_ctx *ctx, uint8_t **p, size_t *cap, size_t need)
{
	if (need <= *cap) return 0;
	CU(cudaDeviceSynchronize());
	cudaFree(*p); *p = nullptr; *cap = 0;
	CU(cudaMalloc((void **)p, need + 256));
	*cap = need + 256;
	return 0;
}

extern "C" long tb200_rx_stream_dev(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint32_t flags,
                                    tb200_slot *d_slots, uint8_t *d_type1, uint32_t *d_type1_packed, uint64_t max_slots)
{
	if (!ctx) return TB200_E_ARG;
	if ((!d_bits && n_bits) || !d_slots) return fail(ctx, TB200_E_ARG, "null buffer");
	if (d_type1 && ((uintptr_t)d_type1 & 15)) return fail(ctx, TB200_E_ARG, "d_type1 must be 16-byte aligned");
	if (ctx->opt.input != TB200_IN_BYTES && ((uintptr_t)d_bits & 3))
		return fail(ctx, TB200_E_ARG, "packed / symbol input must be 4-byte aligned");
	const bool fresh = (flags & TB200_FRESH) != 0, fin = (flags & TB200_FINAL) != 0;
	if (ctx->opt.input != TB200_IN_BYTES && !fin && (n_bits & 127))
		return fail(ctx, TB200_E_ARG, "a bit-packed / symbol stream continues on 128-bit boundaries: n_bits of every call but the last must be a multiple of 128");
	const bool afc = ctx->opt.input == TB200_IN_F32SYM && ctx->opt.afc;
	const int eff_fmt = afc ? IN_PACKED : (int)ctx->opt.input;          /* how the chain sees this call's bits */
	if (!fresh && ctx->tail_fmt != eff_fmt && ctx->fed_end)
		return fail(ctx, TB200_E_ARG, "the input format cannot change inside a stream");
	HostTrace trace;
	g_trace = trace.on ? &trace : nullptr;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	CU(cudaDeviceSynchronize());        /* the caller's buffers may still be in flight on its own streams */
	TB_TRACE("entry sync");
	if (fresh) { reset_stream(ctx); ctx->afc_state = 0.f; }
	int rc = push_carry(ctx);
	if (rc) return rc;
	TB_TRACE("carry pushed");
	ctx->stats.kernel_launches = 0;          /* "by the last call" */
	const uint8_t *data = d_bits;
	if (afc && n_bits) {
		/* float_to_bits -a first (the tracker is a recurrence over the whole stream, its state carries over between calls),
		 * then the chain on packed bits */
		const size_t need = 4 * (size_t)((n_bits / 2 + 15) / 16) + 256;
		if (need > ctx->afc_bits_cap) {
			cudaFree(ctx->d_afc_bits); ctx->d_afc_bits = nullptr; ctx->afc_bits_cap = 0;
			CU(cudaMalloc((void **)&ctx->d_afc_bits, need));
			ctx->afc_bits_cap = need;
		}
		if ((rc = float_to_bits_dev(ctx, reinterpret_cast<const float *>(d_bits), n_bits / 2, 1, ctx->opt.afc_filter_val,
		                            ctx->opt.afc_filter_goal, &ctx->afc_state, reinterpret_cast<uint32_t *>(ctx->d_afc_bits)))) return rc;
		data = ctx->d_afc_bits;
	}
	Outputs out; out.on_device = true; out.slots = d_slots; out.type1 = d_type1; out.packed = d_type1_packed;
	out.crc = ctx->user_crc; out.aach = ctx->user_aach; out.max_slots = max_slots; out.n = 0;
	ctx->lock_events.clear();
	const uint64_t base = ctx->fed_end;             /* stream bit of the caller's bit 0 */
	ctx->fed_end += n_bits;
	if (base == 0 || ctx->tail_base > base) ctx->tail_base = base;
	const uint64_t tail_bits = base - ctx->tail_base;
	profile_begin(ctx);
	bool whole_staged = false;
	if (tail_bits == 0) {
		Source src; src.on_device = true; src.data = data; src.new_base = base; src.end = ctx->fed_end; src.fmt = eff_fmt;
		rc = rx_run(ctx, src, fin, out);
	} else {
		/* a continuing call.  The receiver may still look at bits of earlier calls (its buffer, a slot that straddles
		 * the calls): those were kept, they go in front of the head of the new buffer into one contiguous piece, and a first
		 * run works on that.  After DEV_CONT_HEAD_BITS the receiver's buffer lies inside the caller's memory and a second
		 * run of the same modelled calls goes on in place - nothing else of the new buffer is copied. */
		const size_t tb = fmt_bytes(eff_fmt, tail_bits);
		if (!ctx->tail_dev) {                       /* the stream came through host buffers so far */
			if ((rc = dev_reserve(ctx, &ctx->d_tail, &ctx->d_tail_cap, tb))) return rc;
			CU(cudaMemcpy(ctx->d_tail, ctx->tail.data(), tb, cudaMemcpyHostToDevice));
			ctx->d_tail_bytes = tb; ctx->tail_dev = true;
		}
		/* (whole modelled reads only count for a run that is not the stream's last: a few of them must fit the head) */
		const uint64_t head = std::min<uint64_t>(n_bits, (std::max<uint64_t>(DEV_CONT_HEAD_BITS, 4ull * ctx->opt.chunk_bits + 8192) + 127) & ~127ull);
		whole_staged = head == n_bits;
		const size_t hb = fmt_bytes(eff_fmt, head);
		if ((rc = dev_reserve(ctx, &ctx->d_cont, &ctx->d_cont_cap, tb + hb + 64))) return rc;
		CU(cudaMemcpyAsync(ctx->d_cont, ctx->d_tail, tb, cudaMemcpyDeviceToDevice, ctx->s_compute));
		if (hb) CU(cudaMemcpyAsync(ctx->d_cont + tb, data, hb, cudaMemcpyDeviceToDevice, ctx->s_compute));
		CU(cudaMemsetAsync(ctx->d_cont + tb + hb, 0, 64, ctx->s_compute));
		CU(cudaStreamSynchronize(ctx->s_compute));
		Source s1; s1.on_device = true; s1.data = ctx->d_cont; s1.new_base = ctx->tail_base; s1.end = base + head; s1.fmt = eff_fmt;
		rc = rx_run(ctx, s1, fin && whole_staged, out);
		if (!rc && !whole_staged) {
			if (ctx->rx.buf_start < base)
				return fail(ctx, TB200_E_STATE, "device continuation: the receiver still needs bits of the previous call");
			Source s2; s2.on_device = true; s2.data = data; s2.new_base = base; s2.end = ctx->fed_end; s2.fmt = eff_fmt;
			rc = rx_run(ctx, s2, fin, out, true);
		}
	}
	TB_TRACE("rx_run done");
	g_trace = nullptr;
	if (rc) return rc;
	if ((rc = profile_end(ctx))) return rc;
	/* keep what a later call may still look at: everything from bitbuf[0] on (whole 128-bit units of the packed formats) */
	if (!fin) {
		uint64_t keep_from = std::min<uint64_t>(ctx->rx.buf_start, ctx->fed_end);
		if (eff_fmt != IN_BYTES) keep_from &= ~(uint64_t)127;
		if (keep_from < ctx->tail_base) keep_from = ctx->tail_base;
		const size_t kb = fmt_bytes(eff_fmt, ctx->fed_end - keep_from);
		const uint8_t *from;
		if (keep_from >= base) from = data + fmt_bytes(eff_fmt, keep_from - base);
		else if (whole_staged) from = ctx->d_cont + fmt_bytes(eff_fmt, keep_from - ctx->tail_base);
		else return fail(ctx, TB200_E_STATE, "device continuation: tail outside the staged head");
		/* d_cont / the caller's buffer -> d_tail (never d_tail -> d_tail: the old tail was copied into d_cont) */
		if ((rc = dev_reserve(ctx, &ctx->d_tail, &ctx->d_tail_cap, kb))) return rc;
		CU(cudaMemcpyAsync(ctx->d_tail, from, kb, cudaMemcpyDeviceToDevice, ctx->s_compute));
		CU(cudaStreamSynchronize(ctx->s_compute));
		ctx->d_tail_bytes = kb; ctx->tail_dev = true;
		ctx->tail.clear();
		ctx->tail_base = keep_from;
	} else {
		ctx->d_tail_bytes = 0; ctx->tail.clear(); ctx->tail_base = ctx->fed_end;
	}
	ctx->tail_fmt = eff_fmt;
	return (long)out.n;
}

extern "C" long tb200_rx_stream_host(tb200_ctx *ctx, const uint8_t *bits, uint64_t n_bits, uint32_t flags,
                                     tb200_slot *slots, uint8_t *type1, uint32_t *type1_packed, uint64_t max_slots)
{
	if (!ctx) return TB200_E_ARG;
	if ((!bits && n_bits) || !slots) return fail(ctx, TB200_E_ARG, "null buffer");
	if (ctx->opt.input != TB200_IN_BYTES && !(flags & TB200_FINAL) && (n_bits & 127))
		return fail(ctx, TB200_E_ARG, "a bit-packed / symbol stream continues on 128-bit boundaries: n_bits of every call but the last must be a multiple of 128");
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	if (flags & TB200_FRESH) { reset_stream(ctx); ctx->afc_state = 0.f; }
	const bool afc = ctx->opt.input == TB200_IN_F32SYM && ctx->opt.afc;
	const int eff_fmt = afc ? IN_PACKED : (int)ctx->opt.input;          /* how the chain sees this call's bits */
	if (!(flags & TB200_FRESH) && ctx->tail_fmt != eff_fmt && ctx->fed_end)
		return fail(ctx, TB200_E_ARG, "the input format cannot change inside a stream");
	if (ctx->tail_dev) {                     /* the stream came through device buffers so far: its kept tail moves to the host */
		ctx->tail.resize(ctx->d_tail_bytes);
		if (ctx->d_tail_bytes) CU(cudaMemcpy(ctx->tail.data(), ctx->d_tail, ctx->d_tail_bytes, cudaMemcpyDeviceToHost));
		ctx->tail_dev = false; ctx->d_tail_bytes = 0;
	}
	if (afc && n_bits) {
		/* float_to_bits -a over this call's symbols, from the tracker state the last call left: symbols up, packed bits back */
		const uint64_t n_sym = n_bits / 2;
		const size_t out_bytes = 4 * (size_t)((n_sym + 15) / 16);
		if (n_sym > ctx->afc_sym_cap) {
			cudaFree(ctx->d_afc_sym); ctx->d_afc_sym = nullptr; ctx->afc_sym_cap = 0;
			CU(cudaMalloc((void **)&ctx->d_afc_sym, n_sym * sizeof(float) + 64));
			ctx->afc_sym_cap = n_sym;
		}
		if (out_bytes + 256 > ctx->afc_bits_cap) {
			cudaFree(ctx->d_afc_bits); ctx->d_afc_bits = nullptr; ctx->afc_bits_cap = 0;
			CU(cudaMalloc((void **)&ctx->d_afc_bits, out_bytes + 256));
			ctx->afc_bits_cap = out_bytes + 256;
		}
		CU(cudaMemcpyAsync(ctx->d_afc_sym, bits, n_sym * sizeof(float), cudaMemcpyHostToDevice, ctx->s_compute));
		int rc0 = float_to_bits_dev(ctx, ctx->d_afc_sym, n_sym, 1, ctx->opt.afc_filter_val, ctx->opt.afc_filter_goal, &ctx->afc_state,
		                            reinterpret_cast<uint32_t *>(ctx->d_afc_bits));
		if (rc0) return rc0;
		ctx->h_afc_bits.resize(out_bytes + 64);
		CU(cudaMemcpyAsync(ctx->h_afc_bits.data(), ctx->d_afc_bits, out_bytes, cudaMemcpyDeviceToHost, ctx->s_compute));
		CU(cudaStreamSynchronize(ctx->s_compute));
		bits = ctx->h_afc_bits.data();
	}
	int rc = push_carry(ctx);
	if (rc) return rc;
	ctx->stats.kernel_launches = 0;          /* "by the last call" */
	Source src; src.on_device = false; src.data = bits; src.new_base = ctx->fed_end; src.end = ctx->fed_end + n_bits;
	src.fmt = eff_fmt;
	Outputs out; out.on_device = false; out.slots = slots; out.type1 = type1; out.packed = type1_packed;
	out.crc = ctx->user_crc; out.aach = ctx->user_aach; out.max_slots = max_slots; out.n = 0;
	ctx->lock_events.clear();
	ctx->fed_end += n_bits;
	profile_begin(ctx);
	rc = rx_run(ctx, src, (flags & TB200_FINAL) != 0, out);
	if (rc) return rc;
	if ((rc = profile_end(ctx))) return rc;
	/* keep what a later call may still look at: everything from bitbuf[0] on */
	uint64_t keep_from = std::min<uint64_t>(ctx->rx.buf_start, ctx->fed_end);
	std::vector<uint8_t> nt;
	if (src.fmt == IN_BYTES) {
		nt.reserve(ctx->fed_end - keep_from);
		for (uint64_t i = keep_from; i < ctx->fed_end; i++)
			nt.push_back(i < src.new_base ? ctx->tail[i - ctx->tail_base] : bits[i - src.new_base]);
	} else if (!(flags & TB200_FINAL)) {
		keep_from &= ~(uint64_t)127;                     /* whole 128-bit units, as the staging copies them */
		if (keep_from < src.new_base) {
			const size_t b0 = fmt_bytes(src.fmt, keep_from - ctx->tail_base), b1 = fmt_bytes(src.fmt, src.new_base - ctx->tail_base);
			nt.insert(nt.end(), ctx->tail.begin() + b0, ctx->tail.begin() + b1);
		}
		const uint64_t from_new = std::max(keep_from, src.new_base);
		const size_t c0 = fmt_bytes(src.fmt, from_new - src.new_base), c1 = fmt_bytes(src.fmt, ctx->fed_end - src.new_base);
		nt.insert(nt.end(), bits + c0, bits + c1);
	}
	ctx->tail.swap(nt);
	ctx->tail_base = keep_from;
	ctx->tail_fmt = src.fmt;
	/* hits before the kept range are of no further use */
	return (long)out.n;
}

extern "C" int tb200_get_carry(const tb200_ctx *ctx, tb200_rx_carry *o)
{
	if (!ctx || !o) return TB200_E_ARG;
	memset(o, 0, sizeof(*o));
	o->stream_bits = ctx->fed_end;
	o->buf_start_bit = ctx->rx.buf_start;
	o->next_frame_start = ctx->rx.next_frame_start;
	o->calls = ctx->rx.calls;
	o->state = ctx->rx.state;
	o->bits_in_buf = ctx->rx.bits_in_buf;
	o->scramb_init = ctx->h_carry.scramb_init;
	o->mcc = (uint16_t)ctx->h_carry.mcc; o->mnc = (uint16_t)ctx->h_carry.mnc; o->colour_code = (uint8_t)ctx->h_carry.cc;
	o->tn = (uint8_t)ctx->h_carry.tn; o->fn = (uint8_t)ctx->h_carry.fn; o->mn = (uint8_t)ctx->h_carry.mn;
	return 0;
}

extern "C" int tb200_get_timing(const tb200_ctx *ctx, tb200_timing *o)
{
	if (!ctx || !o) return TB200_E_ARG;
	*o = ctx->timing;
	return 0;
}

/* register-only integer kernel: 8 independent add/min chains per thread */
__global__ void __launch_bounds__(256)
k_int_peak(uint32_t *out, int iters, uint32_t seed)
{
	uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
	const uint32_t k = seed | 1;
	for (int i = 0; i < iters; ++i) {
#pragma unroll 8
		for (int u = 0; u < 8; ++u) {
			a0 = umin(a0 + k, a1); a1 = umin(a1 + k, a2); a2 = umin(a2 + k, a3); a3 = umin(a3 + k, a4);
			a4 = umin(a4 + k, a5); a5 = umin(a5 + k, a6); a6 = umin(a6 + k, a7); a7 = umin(a7 + k, a0);
		}
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

extern "C" double tb200_measure_int_peak(tb200_ctx *ctx)
{
	if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return 0.0;
	const int blocks = ctx->sm_count * 8, iters = 4096;
	uint32_t *d = nullptr;
	if (cudaMalloc((void **)&d, sizeof(uint32_t) * blocks * 256) != cudaSuccess) return 0.0;
	double best = 0.0;
#ifndef TB_SIMT_EMULATION
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(e0, ctx->s_compute);
		k_int_peak<<<blocks, 256, 0, ctx->s_compute>>>(d, iters, 12345u + rep);
		cudaEventRecord(e1, ctx->s_compute);
		cudaEventSynchronize(e1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		/* 8 x 8 (add + min) pairs per iteration per thread = 128 integer results */
		const double ops = (double)blocks * 256 * iters * 128.0;
		if (rep > 0 && ms > 0.f) best = std::max(best, ops / (ms * 1e-3));
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
#endif
	cudaFree(d);
	return best;
}

extern "C" int tb200_get_stats(const tb200_ctx *ctx, tb200_stats *o)
{
	if (!ctx || !o) return TB200_E_ARG;
	*o = ctx->stats;
	return 0;
}

/* ------------------------------------------------------------- records -- */

extern "C" size_t tb200_expand_records(const tb200_slot *slots, const uint8_t *type1, size_t n_slots,
                                       tb200_record *rec, size_t max_records)
{
	/* enum tetra_log_chan values, tetra_common.h:22-39 */
	enum { LC_UNKNOWN = 0, LC_SCH_F = 1, LC_AACH = 8, LC_BSCH = 10, LC_BNCH = 11 };
	size_t n = 0;
	auto emit = [&](const tb200_slot &s, const uint8_t *t1, int off, int len, int lchan, int crc_ok, int blk, uint32_t code) {
		if (rec && n < max_records) {
			tb200_record &r = rec[n];
			memset(&r, 0, sizeof(r));
			r.slot_bit = s.slot_bit;
			r.lchan = (uint8_t)lchan; r.crc_ok = (uint8_t)crc_ok; r.blk_num = (uint8_t)blk;
			r.tn = s.time & 7; r.fn = (s.time >> 3) & 31; r.mn = (s.time >> 8) & 63;
			r.type1_len = (uint16_t)len;
			r.scrambling_code = code;
			if (t1) memcpy(r.type1, t1 + off, len);
		}
		n++;
	};
	for (size_t i = 0; i < n_slots; i++) {
		const tb200_slot &s = slots[i];
		const uint8_t *t1 = type1 ? type1 + i * TB200_TYPE1_STRIDE : nullptr;
		const int a = (s.flags & TB200_F_CRC_A) != 0, b = (s.flags & TB200_F_CRC_B) != 0;
		switch (s.flags & TB200_F_KIND_MASK) {
		case TB200_KIND_SB:       /* tetra_burst.c:350-352 */
			emit(s, t1, 0, 60, LC_BSCH, a, 1, 3);
			emit(s, t1, 60, 14, LC_AACH, 1, 0, s.scrambling_code);
			emit(s, t1, 74, 124, (s.flags & TB200_F_BNCH) ? LC_BNCH : LC_UNKNOWN, b, 2, s.scrambling_code);
			break;
		case TB200_KIND_NDB_F:    /* tetra_burst.c:371-372 */
			emit(s, t1, 0, 14, LC_AACH, 1, 0, s.scrambling_code);
			emit(s, t1, 14, 268, LC_SCH_F, a, 0, s.scrambling_code);
			break;
		case TB200_KIND_NDB_2:    /* tetra_burst.c:359-361 */
			emit(s, t1, 0, 14, LC_AACH, 1, 0, s.scrambling_code);
			emit(s, t1, 14, 124, LC_UNKNOWN, a, 1, s.scrambling_code);
			emit(s, t1, 138, 124, LC_UNKNOWN, b, 2, s.scrambling_code);
			break;
		default:
			break;
		}
	}
	return n;
}

/* -------------------------------------------------------- leaf operators -- */

/* one warp per window */
__global__ void __launch_bounds__(256)
k_leaf_find(const uint8_t *bits, uint64_t n_bytes, const uint64_t *starts, const uint32_t *lens, uint64_t n,
            uint32_t mask, const Tables *__restrict__ tab, int32_t *rc, uint32_t *offs)
{
	const unsigned lane = threadIdx.x & 31;
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
		unsigned off = 0;
		int r = find_train_seq_warp(bits + starts[i], bits + n_bytes, lens[i], mask, tab, &off, nullptr);
		if (lane == 0) { rc[i] = r; offs[i] = off; }
	}
}

template <int BT>
__device__ __forceinline__ void leaf_decode_one(WarpSmem &S, const uint8_t *type5, uint32_t code, const Tables *tab,
                                                uint8_t *type1, uint8_t *crc_ok, unsigned lane, int variant, bool tie_hi)
{
	constexpr int K = Blk<BT>::K, N = Blk<BT>::N, T1 = Blk<BT>::T1;
	uint32_t x0, x1, x2;
	load_window(type5, type5 + K, lane, x0, x1, x2);
	if (lane < 16) S.bw[lane] = x0;
	if (lane < 4) S.bw[16 + lane] = 0;
	const uint32_t lw = lfsr_word(code, lane, tab);
	if (lane < 16) S.lf[lane] = lw;
	if (lane < 12) S.outw[lane] = 0;
	__syncwarp();
	gather_type3<BT, PL_RAW>(S.bw, S.lf, S.t3[0], lane);
	uint32_t crc;
	if (variant == TB200_VITERBI_LANE) {
		if (lane == 0) viterbi_lane<N>(S.t3[0], S.dec, S.t2[0], tie_hi);
		__syncwarp();
		crc = crc16_serial(S.t2[0], T1 + 16);
	} else {
		viterbi_warp<N>(S.t3[0], S.t3[0], false, S.dec, S.t2[0], S.t2[0], tie_hi);
		crc = crc16_half(S.t2[0], T1 + 16, Blk<BT>::CRCI, tab);
	}
	put_bits(S.outw, 0, S.t2[0], T1, lane);
	__syncwarp();
	for (int i = lane; i < T1; i += 32)
		type1[i] = (S.outw[i >> 5] >> (i & 31)) & 1;
	if (lane == 0) *crc_ok = (crc == 0x1d0f);
	__syncwarp();
}

__global__ void __launch_bounds__(256)
k_leaf_decode(int blk_type, const uint8_t *type5, const uint32_t *codes, uint64_t n, const Tables *__restrict__ tab,
              uint8_t *type1, uint8_t *crc_ok, int variant, int tie)
{
	const bool tie_hi = tie != 0;
	__shared__ WarpSmem sm[8];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib; i < n; i += nwarps) {
		if (blk_type == TB200_T_SB1)
			leaf_decode_one<0>(sm[wib], type5 + i * 120, codes[i], tab, type1 + i * 60, crc_ok + i, lane, variant, tie_hi);
		else if (blk_type == TB200_T_SCH_F)
			leaf_decode_one<5>(sm[wib], type5 + i * 432, codes[i], tab, type1 + i * 268, crc_ok + i, lane, variant, tie_hi);
		else if (blk_type == TB200_T_SCH_HU)
			leaf_decode_one<4>(sm[wib], type5 + i * 168, codes[i], tab, type1 + i * 92, crc_ok + i, lane, variant, tie_hi);
		else if (blk_type == TB200_T_BBK) {
			/* no channel decoding in the reference (tetra_lower_mac.c:268-274): the first 14 descrambled bits, CRC flag 1 */
			const uint32_t lw = lfsr_word(codes[i], 0, tab);
			if (lane < 14) type1[i * 14 + lane] = (type5[i * 30 + lane] ^ (lw >> lane)) & 1;
			if (lane == 0) crc_ok[i] = 1;
		} else
			leaf_decode_one<1>(sm[wib], type5 + i * 216, codes[i], tab, type1 + i * 124, crc_ok + i, lane, variant, tie_hi);
	}
}

/* The fused descramble + de-interleave stage on its own (type-5 bytes -> type-3 bytes):
 * a CTA stages 8 blocks, each warp permutes one.  Memory-bound by construction:
 * K bytes read + K bytes written per block. */
__global__ void __launch_bounds__(256)
k_descramble_deinterleave(const uint8_t *__restrict__ type5, uint8_t *__restrict__ type3,
                          const uint32_t *__restrict__ codes, uint64_t n, uint32_t K, uint32_t a,
                          const Tables *__restrict__ tab)
{
	__shared__ uint32_t s_bits[8][16];
	__shared__ uint32_t s_lf[8][16];
	const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
	for (uint64_t i = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib; i < n; i += nwarps) {
		const uint8_t *src = type5 + i * K;
		uint32_t x0, x1, x2;
		load_window(src, src + K, lane, x0, x1, x2);
		const uint32_t lw = lfsr_word(codes[i], lane, tab);
		if (lane < 16) { s_bits[wib][lane] = x0 ^ lw; }
		(void)s_lf;
		__syncwarp();
		uint8_t *dst = type3 + i * K;
		/* each lane produces 4 consecutive output bytes per round -> coalesced 128-byte stores */
		for (uint32_t j0 = 4 * lane; j0 < K; j0 += 128) {
			uint32_t word = 0;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t j = j0 + q;
				if (j < K) {
					const uint32_t m = (a * (j + 1)) % K;
					word |= ((s_bits[wib][m >> 5] >> (m & 31)) & 1) << (8 * q);
				}
			}
			if (j0 + 4 <= K && (((uintptr_t)(dst + j0)) & 3) == 0) {
				*reinterpret_cast<uint32_t *>(dst + j0) = word;
			} else {
				for (int q = 0; q < 4 && j0 + q < K; ++q)
					dst[j0 + q] = (word >> (8 * q)) & 0xff;
			}
		}
		__syncwarp();
	}
}

template <typename T>
static int leaf_buf(tb200_ctx *ctx, int i, size_t n, T **p)
{
	const size_t bytes = n * sizeof(T) + 64;
	if (bytes > ctx->leaf_cap[i]) {
		if (ctx->leaf_mem[i]) { CU(cudaDeviceSynchronize()); cudaFree(ctx->leaf_mem[i]); }
		ctx->leaf_mem[i] = nullptr; ctx->leaf_cap[i] = 0;
		CU(cudaMalloc(&ctx->leaf_mem[i], bytes + bytes / 4));
		ctx->leaf_cap[i] = bytes + bytes / 4;
	}
	*p = reinterpret_cast<T *>(ctx->leaf_mem[i]);
	return 0;
}

static int leaf_events(tb200_ctx *ctx)
{
	for (int i = 0; i < 2; i++)
		if (!ctx->leaf_ev[i]) CU(cudaEventCreateWithFlags(&ctx->leaf_ev[i], 0));
	return 0;
}

static int leaf_common(tb200_ctx *ctx)
{
	if (!ctx) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	/* callers hand in device buffers they may have just produced on THEIR streams (e.g. a torch.zeros
	 * still running); our streams are non-blocking, so wait for the device before touching them */
	CU(cudaDeviceSynchronize());
	return 0;
}

extern "C" int tb200_find_train_seq(tb200_ctx *ctx, const uint8_t *bits, uint64_t n_bits, const uint64_t *starts,
                                    const uint32_t *lens, uint64_t n, uint32_t mask, int32_t *rc, uint32_t *offset)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (n == 0) return 0;
	uint8_t *d_bits = nullptr; uint64_t *d_st = nullptr; uint32_t *d_len = nullptr, *d_off = nullptr; int32_t *d_rc = nullptr;
	if ((r = leaf_buf(ctx, 0, n_bits + 64, &d_bits)) || (r = leaf_buf(ctx, 1, n, &d_st)) || (r = leaf_buf(ctx, 2, n, &d_len)) ||
	    (r = leaf_buf(ctx, 3, n, &d_off)) || (r = leaf_buf(ctx, 4, n, &d_rc))) return r;
	cudaStream_t st = ctx->s_compute;
	CU(cudaMemcpyAsync(d_bits, bits, n_bits, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_st, starts, n * 8, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(d_len, lens, n * 4, cudaMemcpyHostToDevice, st));
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 7) / 8, 4096);
	TB_LAUNCH(k_leaf_find, blocks, 256, st, d_bits, n_bits, d_st, d_len, n, mask, ctx->d_tab, d_rc, d_off);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(rc, d_rc, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(offset, d_off, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int tb200_decode_blocks(tb200_ctx *ctx, int blk_type, const uint8_t *type5, const uint32_t *codes,
                                   uint64_t n, uint8_t *type1, uint8_t *crc_ok)
{
	int r = leaf_common(ctx);
	if (r) return r;
	int K, T1;
	switch (blk_type) {
	case TB200_T_SB1: K = 120; T1 = 60; break;
	case TB200_T_SB2: case TB200_T_NDB: K = 216; T1 = 124; break;
	case TB200_T_SCH_F: K = 432; T1 = 268; break;
	case TB200_T_SCH_HU: K = 168; T1 = 92; break;
	case TB200_T_BBK: K = 30; T1 = 14; break;
	default: return fail(ctx, TB200_E_ARG, "block type %d is not decoded by the receive path", blk_type);
	}
	if (n == 0) return 0;
	uint8_t *d5 = nullptr, *d1 = nullptr, *dc = nullptr; uint32_t *dcode = nullptr;
	if ((r = leaf_buf(ctx, 0, n * K + 64, &d5)) || (r = leaf_buf(ctx, 1, n * T1, &d1)) || (r = leaf_buf(ctx, 2, n, &dc)) ||
	    (r = leaf_buf(ctx, 3, n, &dcode))) return r;
	cudaStream_t st = ctx->s_compute;
	CU(cudaMemcpyAsync(d5, type5, n * K, cudaMemcpyHostToDevice, st));
	CU(cudaMemcpyAsync(dcode, codes, n * 4, cudaMemcpyHostToDevice, st));
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 7) / 8, 4096);
	TB_LAUNCH(k_leaf_decode, blocks, 256, st, blk_type, d5, dcode, n, ctx->d_tab, d1, dc, (int)ctx->opt.viterbi, (int)ctx->opt.viterbi_tie);
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(type1, d1, n * T1, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(crc_ok, dc, n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

extern "C" int tb200_descramble_deinterleave(tb200_ctx *ctx, const uint8_t *type5, uint8_t *type3, const uint32_t *codes,
                                             uint64_t n, uint32_t K, uint32_t a, int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (K == 0 || K > 432 || a == 0) return fail(ctx, TB200_E_ARG, "K must be 1..432");
	if (n == 0) return 0;
	const uint8_t *d5 = type5; uint8_t *d3 = type3; const uint32_t *dcode = codes;
	uint8_t *a5 = nullptr, *a3 = nullptr; uint32_t *ac = nullptr;
	cudaStream_t st = ctx->s_compute;
	if (!is_device) {
		if ((r = leaf_buf(ctx, 0, n * K + 64, &a5)) || (r = leaf_buf(ctx, 1, n * K, &a3)) || (r = leaf_buf(ctx, 2, n, &ac))) return r;
		CU(cudaMemcpyAsync(a5, type5, n * K, cudaMemcpyHostToDevice, st));
		CU(cudaMemcpyAsync(ac, codes, n * 4, cudaMemcpyHostToDevice, st));
		d5 = a5; d3 = a3; dcode = ac;
	}
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 8);
	if (ctx->opt.profile) {
		if ((r = leaf_events(ctx))) return r;
		CU(cudaEventRecord(ctx->leaf_ev[0], st));
	}
	if (K == ST_K && a == ST_A && (((uintptr_t)d5 | (uintptr_t)d3) & 15) == 0) {
		const uint64_t groups = (n + 31) / 32;
		const unsigned sb = (unsigned)std::min<uint64_t>((groups + ST_WARPS - 1) / ST_WARPS, (uint64_t)ctx->sm_count);
		TB_LAUNCH_SMEM(k_stage_tma, sb, ST_WARPS * 32, ST_SMEM, st, d5, d3, dcode, n, ctx->d_tab);
	} else {
		TB_LAUNCH(k_descramble_deinterleave, blocks, 256, st, d5, d3, dcode, n, K, a, ctx->d_tab);
	}
	if (ctx->opt.profile) CU(cudaEventRecord(ctx->leaf_ev[1], st));
	CU(cudaGetLastError());
	if (!is_device) CU(cudaMemcpyAsync(type3, a3, n * K, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	if (ctx->opt.profile) CU(cudaEventElapsedTime(&ctx->timing.leaf_ms, ctx->leaf_ev[0], ctx->leaf_ev[1]));
	return 0;
}

/* ------------------------------------------------------------ de-puncturing -- */

/* one of the reference's seven puncturers (struct puncturer, tetra_conv_enc.c:103-110,128-198): type-3 bit j
 * (1-based) is mother-code bit k = period * ((i-1)/t) + P[i - t*((i-1)/t)] with i = i_func(j) */
struct PunctDef {
	uint8_t P[24];
	uint8_t t, period, ifunc;      /* ifunc: 0 i = j, 1 i = j + (j-1)/65 (292/432), 2 i = j + (j-1)/35 (148/432) */
};

static const PunctDef k_punct_defs[7] = {
	{ { 0, 1, 2, 5 }, 3, 8, 0 },                                                          /* TETRA_RCPC_PUNCT_2_3 */
	{ { 0, 1, 2, 3, 5, 6, 7 }, 6, 8, 0 },                                                 /* 1_3 */
	{ { 0, 1, 2, 5 }, 3, 8, 1 },                                                          /* 292_432 */
	{ { 0, 1, 2, 3, 5, 6, 7 }, 6, 8, 2 },                                                 /* 148_432 */
	{ { 0, 1, 2, 4 }, 3, 6, 0 },                                                          /* 112_168 (speech) */
	{ { 0, 1, 2, 3, 4, 5, 7, 8, 10, 11 }, 9, 12, 0 },                                     /* 72_162 */
	{ { 0, 1, 2, 3, 4, 5, 7, 8, 10, 11, 13, 14, 16, 17, 19, 20, 22, 23 }, 17, 24, 0 },    /* 38_80 */
};

__global__ void __launch_bounds__(256)
k_rcpc_depunct(const uint8_t *__restrict__ type3, uint32_t len, uint64_t n, uint8_t *__restrict__ mother, uint32_t mother_len,
               PunctDef pd)
{
	const uint64_t total = n * len, stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
		const uint64_t b = idx / len;
		const uint32_t j = (uint32_t)(idx - b * len) + 1;
		const uint32_t i = pd.ifunc == 1 ? j + (j - 1) / 65 : pd.ifunc == 2 ? j + (j - 1) / 35 : j;
		const uint32_t q = (i - 1) / pd.t;
		const uint32_t k = pd.period * q + pd.P[i - pd.t * q];
		if (k - 1 < mother_len)
			mother[b * mother_len + (k - 1)] = type3[idx];
	}
}

extern "C" int tb200_rcpc_depunct(tb200_ctx *ctx, int puncturer, const uint8_t *type3, uint32_t len, uint64_t n,
                                  uint8_t *mother, uint32_t mother_len, int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (puncturer < 0 || puncturer >= 7) return fail(ctx, TB200_E_ARG, "unknown puncturer %d", puncturer);   /* -EINVAL in the reference */
	if (len == 0 || mother_len == 0 || !type3 || !mother) return fail(ctx, TB200_E_ARG, "bad argument");
	if (n == 0) return 0;
	const uint8_t *d3 = type3; uint8_t *dm = mother;
	uint8_t *a3 = nullptr, *am = nullptr;
	cudaStream_t st = ctx->s_compute;
	if (!is_device) {
		if ((r = leaf_buf(ctx, 0, n * len, &a3)) || (r = leaf_buf(ctx, 1, n * mother_len, &am))) return r;
		CU(cudaMemcpyAsync(a3, type3, n * len, cudaMemcpyHostToDevice, st));
		d3 = a3; dm = am;
	}
	/* the caller of the reference function fills the mother buffer with 0xff first (tetra_lower_mac.c:249) */
	CU(cudaMemsetAsync(dm, 0xff, n * mother_len, st));
	const unsigned blocks = (unsigned)std::min<uint64_t>((n * len + 255) / 256, (uint64_t)ctx->sm_count * 16);
	TB_LAUNCH(k_rcpc_depunct, blocks, 256, st, d3, len, n, dm, mother_len, k_punct_defs[puncturer]);
	CU(cudaGetLastError());
	if (!is_device) CU(cudaMemcpyAsync(mother, am, n * mother_len, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

/* ----------------------------------------------------------- GSMTAP framing -- */

extern "C" long long tb200_gsmtap_pack(tb200_ctx *ctx, const tb200_slot *slots, const uint32_t *type1_packed, uint64_t n_slots,
                                       uint8_t *frames, uint64_t cap_bytes, uint64_t *slot_off, uint64_t *n_frames,
                                       int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (n_slots && (!slots || (frames && !type1_packed))) return fail(ctx, TB200_E_ARG, "null argument");
	if ((uintptr_t)frames & 1) return fail(ctx, TB200_E_ARG, "frames must be 2-byte aligned");
	if (n_frames) *n_frames = 0;
	if (n_slots == 0) return 0;
	const uint64_t tiles = (n_slots + GT_THREADS - 1) / GT_THREADS;
	if (tiles > 0x7fffffffull) return fail(ctx, TB200_E_ARG, "too many slots for one call");
	const SlotOut *d_slots = reinterpret_cast<const SlotOut *>(slots);
	const uint32_t *d_packed = type1_packed;
	SlotOut *a_slots = nullptr; uint32_t *a_packed = nullptr; uint8_t *a_frames = nullptr; uint64_t *a_off = nullptr, *d_tot = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	auto release = [&]() {
		cudaFree(a_slots); cudaFree(a_packed); cudaFree(a_frames); cudaFree(a_off); cudaFree(d_tot);
		if (e0) cudaEventDestroy(e0);
		if (e1) cudaEventDestroy(e1);
	};
#define CUR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); \
	return fail(ctx, TB200_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
	if (!is_device) {
		CUR(cudaMalloc((void **)&a_slots, n_slots * sizeof(SlotOut)));
		CUR(cudaMemcpy(a_slots, slots, n_slots * sizeof(SlotOut), cudaMemcpyHostToDevice));
		d_slots = a_slots;
		if (frames) {
			CUR(cudaMalloc((void **)&a_packed, n_slots * TYPE1_WORDS * 4));
			CUR(cudaMemcpy(a_packed, type1_packed, n_slots * TYPE1_WORDS * 4, cudaMemcpyHostToDevice));
			d_packed = a_packed;
		}
	}
	CUR(cudaMalloc((void **)&d_tot, (tiles + 1) * 8));
	if (ctx->opt.profile) {
		CUR(cudaEventCreateWithFlags(&e0, 0)); CUR(cudaEventCreateWithFlags(&e1, 0));
		CUR(cudaEventRecord(e0, ctx->s_compute));
	}
	TB_LAUNCH(k_gsmtap_sizes, (unsigned)tiles, GT_THREADS, ctx->s_compute, d_slots, n_slots, d_tot);
	TB_LAUNCH(k_gsmtap_scan, 1, 1024, ctx->s_compute, d_tot, tiles);
	uint8_t *d_frames = frames; uint64_t *d_off = slot_off;
	if (frames) {
		/* the emit pass is queued before the total is known (no host round trip between the passes); it checks
		 * the capacity itself */
		if (!is_device) {
			CUR(cudaMalloc((void **)&a_frames, std::min<uint64_t>(cap_bytes, n_slots * GT_SLOT_MAX) + 16));
			d_frames = a_frames;
			if (slot_off) { CUR(cudaMalloc((void **)&a_off, (n_slots + 1) * 8)); d_off = a_off; }
		}
		TB_LAUNCH(k_gsmtap_emit, (unsigned)tiles, GT_THREADS, ctx->s_compute, d_slots, d_packed, n_slots, d_tot,
		          reinterpret_cast<uint16_t *>(d_frames), cap_bytes, d_off);
		if (ctx->opt.profile) CUR(cudaEventRecord(e1, ctx->s_compute));
	}
	uint64_t total = 0;
	CUR(cudaMemcpyAsync(&total, d_tot + tiles, 8, cudaMemcpyDeviceToHost, ctx->s_compute));
	CUR(cudaGetLastError());
	CUR(cudaStreamSynchronize(ctx->s_compute));
	const uint64_t bytes = total & ((1ull << 40) - 1);
	if (n_frames) *n_frames = total >> 40;
	if (frames) {
		if (ctx->opt.profile) CUR(cudaEventElapsedTime(&ctx->timing.leaf_ms, e0, e1));
		if (bytes > cap_bytes) {
			release();
			return fail(ctx, TB200_E_ARG, "GSMTAP frames need %llu bytes, the buffer holds %llu", (unsigned long long)bytes, (unsigned long long)cap_bytes);
		}
		if (d_off) CUR(cudaMemcpy(d_off + n_slots, &bytes, 8, cudaMemcpyHostToDevice));
		if (!is_device) {
			CUR(cudaMemcpy(frames, a_frames, bytes, cudaMemcpyDeviceToHost));
			if (slot_off) CUR(cudaMemcpy(slot_off, a_off, (n_slots + 1) * 8, cudaMemcpyDeviceToHost));
		}
	}
	release();
#undef CUR
	return (long long)bytes;
}

/* -------------------------------------------------------------- generator -- */

extern "C" int tb200_gen_stream_dev(tb200_ctx *ctx, const tb200_gen_cfg *cfg, uint64_t k0, uint64_t n,
                                    uint8_t *d_out, int with_lead_in)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (!cfg || !d_out) return fail(ctx, TB200_E_ARG, "null argument");
	GenCfg g;
	memcpy(&g, cfg, sizeof(g));
	uint8_t *bursts = d_out;
	if (with_lead_in && g.lead_in_bits) {
		const unsigned blocks = (g.lead_in_bits + 255) / 256;
		TB_LAUNCH(k_gen_lead_in, blocks, 256, ctx->s_compute, g, d_out);
		bursts += g.lead_in_bits;
	}
	if (n) {
		const unsigned blocks = (unsigned)std::min<uint64_t>((n + 63) / 64, 1u << 20);
		TB_LAUNCH(k_gen_bursts, blocks, 64, ctx->s_compute, g, k0, n, bursts);
	}
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(ctx->s_compute));
	return 0;
}

extern "C" void tb200_debug_time_advance(uint32_t *tn, uint32_t *fn, uint32_t *mn, uint64_t n)
{
	Tm t = { *tn, *fn, *mn };
	t = tm_advance(t, n);
	*tn = t.tn; *fn = t.fn; *mn = t.mn;
}

/* ------------------------------------------- Viterbi wrapper as a leaf (any rate) -- */

/* viterbi_dec_sb1_wrapper (viterbi.c:6-25) + conv_cch_decode (viterbi_cch.c:58-66) for independent blocks: the full
 * rate-1/4 mother trellis with all four generators, any of the 4 * sym_count symbols may be erased, so together with
 * tb200_rcpc_depunct every RCPC rate of tetra_conv_enc.c:128-198 decodes end to end (the receive chain itself only
 * uses 2/3 and runs the packed two-symbol form).  One thread per block, decisions in a scratch row. */
__global__ void __launch_bounds__(128)
k_viterbi_mother(const uint8_t *__restrict__ mother, uint64_t n, uint32_t sym_count, uint8_t *__restrict__ out,
                 uint16_t *__restrict__ dec, int tie_hi)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint32_t steps = sym_count + 4;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint8_t *in = mother + i * 4ull * sym_count;
		uint16_t *d = dec + i * steps;
		uint32_t pm[16];
#pragma unroll
		for (int s = 0; s < 16; ++s) pm[s] = s ? (1u << 28) : 0u;          /* start in state 0 (osmo_conv_decode) */
		for (uint32_t t = 0; t < steps; ++t) {
			uint32_t known = 0, val = 0;                                    /* bit 3 = G1 (viterbi_cch.c:35-40 output nibbles) */
			if (t < sym_count) {
#pragma unroll
				for (int g = 0; g < 4; ++g) {
					const uint8_t v = in[4 * t + g];                          /* viterbi.c:13-22: 0 -> +127, 0xff -> 0, else -127 */
					if (v != 0xff) { known |= 8u >> g; if (v != 0) val |= 8u >> g; }
				}
			}
			uint32_t nm[16], bits = 0;
#pragma unroll
			for (unsigned s = 0; s < 16; ++s) {
				const unsigned b = s & 1, p0 = s >> 1, p1 = p0 | 8;
				const uint32_t c0 = pm[p0] + __popc((mother_out(p0, b) ^ val) & known);
				const uint32_t c1 = pm[p1] + __popc((mother_out(p1, b) ^ val) & known);
				const bool hi = tie_hi ? c1 <= c0 : c1 < c0;                /* include/tetra_tie_rule.h */
				nm[s] = hi ? c1 : c0;
				bits |= (hi ? 1u : 0u) << s;
			}
#pragma unroll
			for (int s = 0; s < 16; ++s) pm[s] = nm[s];
			d[t] = (uint16_t)bits;
		}
		unsigned st = 0;
		for (int t = (int)steps - 1; t >= 0; --t) {
			if ((uint32_t)t < sym_count) out[i * sym_count + t] = st & 1;
			st = (st >> 1) | (((d[t] >> st) & 1u) << 3);
		}
	}
}

extern "C" int tb200_viterbi_decode(tb200_ctx *ctx, const uint8_t *mother, uint64_t n, uint32_t sym_count, uint8_t *out, int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (sym_count == 0 || sym_count > 1024) return fail(ctx, TB200_E_ARG, "sym_count must be 1..1024");
	if (n && (!mother || !out)) return fail(ctx, TB200_E_ARG, "null argument");
	if (n == 0) return 0;
	const uint8_t *dm = mother; uint8_t *dout = out;
	cudaStream_t st = ctx->s_compute;
	if (!is_device) {
		uint8_t *a = nullptr, *b = nullptr;
		if ((r = leaf_buf(ctx, 0, n * 4 * sym_count, &a)) || (r = leaf_buf(ctx, 1, n * sym_count, &b))) return r;
		CU(cudaMemcpyAsync(a, mother, n * 4 * sym_count, cudaMemcpyHostToDevice, st));
		dm = a; dout = b;
	}
	uint16_t *dec = nullptr;
	if ((r = leaf_buf(ctx, 2, n * (sym_count + 4), &dec))) return r;
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 127) / 128, (uint64_t)ctx->sm_count * 8);
	TB_LAUNCH(k_viterbi_mother, blocks, 128, st, dm, n, sym_count, dout, dec, (int)ctx->opt.viterbi_tie);
	CU(cudaGetLastError());
	if (!is_device) CU(cudaMemcpyAsync(out, dout, n * sym_count, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

/* ------------------------------------------------------------ RM(30,14) leaf -- */

__global__ void __launch_bounds__(256)
k_rm3014_decode(const uint32_t *__restrict__ words, uint64_t n, const Tables *__restrict__ tab,
                const uint32_t *__restrict__ leader, uint32_t *__restrict__ out)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		out[i] = rm3014_decode(tab, leader, words[i]);
}

extern "C" int tb200_rm3014_decode(tb200_ctx *ctx, const uint32_t *words, uint64_t n, uint32_t *out, int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (n && (!words || !out)) return fail(ctx, TB200_E_ARG, "null argument");
	if (n == 0) return 0;
	if ((r = ensure_rm_leaders(ctx))) return r;
	const uint32_t *dw = words; uint32_t *dout = out;
	cudaStream_t st = ctx->s_compute;
	if (!is_device) {
		uint32_t *a = nullptr, *b = nullptr;
		if ((r = leaf_buf(ctx, 0, n, &a)) || (r = leaf_buf(ctx, 1, n, &b))) return r;
		CU(cudaMemcpyAsync(a, words, n * 4, cudaMemcpyHostToDevice, st));
		dw = a; dout = b;
	}
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
	TB_LAUNCH(k_rm3014_decode, blocks, 256, st, dw, n, ctx->d_tab, ctx->d_rm_leader, dout);
	CU(cudaGetLastError());
	if (!is_device) CU(cudaMemcpyAsync(out, dout, n * 4, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return 0;
}

/* ------------------------------------------------- float_to_bits (with -a) -- */

/* d_sym[0..n_sym) -> packed hard bits d_out (4 * ceil(n_sym / 16) bytes), on stream s_compute; *state: the tracker's
 * state in front of symbol 0, updated to the state behind the last one.  Synchronous (the verification walks on the host). */
static int float_to_bits_dev(tb200_ctx *ctx, const float *d_sym, uint64_t n_sym, int afc, float filter_val, float filter_goal,
                             float *state, uint32_t *d_out)
{
	cudaStream_t st = ctx->s_compute;
	ctx->afc_repairs = 0;
	if (n_sym == 0) return 0;
	if (!afc) {
		const uint64_t words = (n_sym + 15) / 16;
		const unsigned blocks = (unsigned)std::min<uint64_t>((words + 255) / 256, (uint64_t)ctx->sm_count * 16);
		TB_LAUNCH(k_slice_symbols, blocks, 256, st, d_sym, n_sym, d_out);
		CU(cudaGetLastError());
		CU(cudaStreamSynchronize(st));
		return 0;
	}
	if (!(filter_val > 0.f) || !(filter_val <= 1.f)) return fail(ctx, TB200_E_ARG, "afc_filter_val must be in (0, 1]");
	AfcParams p;
	p.filter_val = filter_val; p.filter_goal = filter_goal; p.keep = 1.0 - (double)filter_val;
	/* warm-up: the start value has decayed by e^-24 < 2^-34; chunks as long as the warm-up (twice the serial work),
	 * but short enough to give every SM a few hundred threads when the stream is long */
	uint64_t W = (uint64_t)std::min<double>(24.0 / (double)filter_val + 64.0, 4e9);
	uint64_t L = std::max<uint64_t>(4096, std::min<uint64_t>(W, n_sym / ((uint64_t)ctx->sm_count * 256) + 1));
	L = (L + 15) & ~(uint64_t)15;
	if (L > 0x7ffffff0ull) L = 0x7ffffff0ull;
	const uint64_t n_chunks = (n_sym + L - 1) / L;
	if (2 * n_chunks > ctx->afc_f_cap) {
		if (ctx->d_afc_f) { CU(cudaDeviceSynchronize()); cudaFree(ctx->d_afc_f); ctx->d_afc_f = nullptr; ctx->afc_f_cap = 0; }
		CU(cudaMalloc((void **)&ctx->d_afc_f, 2 * n_chunks * sizeof(float) + 64));
		ctx->afc_f_cap = 2 * n_chunks;
	}
	float *d_fs = ctx->d_afc_f, *d_fe = ctx->d_afc_f + n_chunks;
	const unsigned blocks = (unsigned)((n_chunks + 127) / 128);
	TB_LAUNCH(k_afc_chunks, blocks, 128, st, d_sym, n_sym, p, *state, (uint32_t)L, W, d_out, d_fs, d_fe, (long long)-1, 0.f);
	CU(cudaGetLastError());
	ctx->h_afc_f.resize(2 * n_chunks);
	CU(cudaMemcpyAsync(ctx->h_afc_f.data(), ctx->d_afc_f, 2 * n_chunks * sizeof(float), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	/* verify: chunk c assumed h_fs[c] at its start; the truth is what chunk c-1 ended with */
	float *h_fs = ctx->h_afc_f.data(), *h_fe = h_fs + n_chunks;
	for (uint64_t c = 1; c < n_chunks; c++) {
		uint32_t a, b;
		memcpy(&a, &h_fs[c], 4); memcpy(&b, &h_fe[c - 1], 4);
		if (a == b) continue;
		TB_LAUNCH(k_afc_chunks, 1, 128, st, d_sym, n_sym, p, 0.f, (uint32_t)L, W, d_out, d_fs, d_fe, (long long)c, h_fe[c - 1]);
		CU(cudaGetLastError());
		CU(cudaMemcpyAsync(&h_fe[c], d_fe + c, sizeof(float), cudaMemcpyDeviceToHost, st));
		CU(cudaStreamSynchronize(st));
		ctx->afc_repairs++;
	}
	*state = h_fe[n_chunks - 1];
	return 0;
}

extern "C" int tb200_float_to_bits(tb200_ctx *ctx, const float *sym, uint64_t n_sym, int afc, float filter_val, float filter_goal,
                                   float *state, uint8_t *packed_bits, int is_device)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (n_sym && (!sym || !packed_bits)) return fail(ctx, TB200_E_ARG, "null argument");
	if (n_sym == 0) return 0;
	float st0 = state ? *state : 0.f;
	const float *d_sym = sym; uint32_t *d_out = reinterpret_cast<uint32_t *>(packed_bits);
	const size_t out_bytes = 4 * (size_t)((n_sym + 15) / 16);
	if (!is_device) {
		float *a = nullptr; uint32_t *b = nullptr;
		if ((r = leaf_buf(ctx, 0, n_sym, &a)) || (r = leaf_buf(ctx, 1, out_bytes / 4, &b))) return r;
		CU(cudaMemcpyAsync(a, sym, n_sym * sizeof(float), cudaMemcpyHostToDevice, ctx->s_compute));
		d_sym = a; d_out = b;
	} else if ((uintptr_t)packed_bits & 3) {
		return fail(ctx, TB200_E_ARG, "packed_bits must be 4-byte aligned");
	}
	if ((r = float_to_bits_dev(ctx, d_sym, n_sym, afc, filter_val, filter_goal, &st0, d_out))) return r;
	if (!is_device) {
		CU(cudaMemcpyAsync(packed_bits, d_out, (size_t)((2 * n_sym + 7) / 8), cudaMemcpyDeviceToHost, ctx->s_compute));
		CU(cudaStreamSynchronize(ctx->s_compute));
	}
	if (state) *state = st0;
	return (int)std::min<uint64_t>(ctx->afc_repairs, 0x7fffffff);
}

/* ---------------------------------------------------------- digest, packing -- */

extern "C" int tb200_slots_digest(tb200_ctx *ctx, const tb200_slot *slots, const uint32_t *type1_packed, uint64_t n,
                                  uint64_t k_base, int is_device, uint64_t *digest)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (!digest || (n && !slots)) return fail(ctx, TB200_E_ARG, "null argument");
	*digest = 0;
	if (n == 0) return 0;
	const SlotOut *d_slots = reinterpret_cast<const SlotOut *>(slots);
	const uint32_t *d_packed = type1_packed;
	cudaStream_t st = ctx->s_compute;
	if (!is_device) {
		SlotOut *a = nullptr; uint32_t *b = nullptr;
		if ((r = leaf_buf(ctx, 0, n, &a))) return r;
		CU(cudaMemcpyAsync(a, slots, n * sizeof(SlotOut), cudaMemcpyHostToDevice, st));
		d_slots = a;
		if (type1_packed) {
			if ((r = leaf_buf(ctx, 1, n * TYPE1_WORDS, &b))) return r;
			CU(cudaMemcpyAsync(b, type1_packed, n * TYPE1_WORDS * 4, cudaMemcpyHostToDevice, st));
			d_packed = b;
		}
	}
	unsigned long long *d_out = nullptr;
	if ((r = leaf_buf(ctx, 2, 1, &d_out))) return r;
	CU(cudaMemsetAsync(d_out, 0, sizeof(*d_out), st));
	const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
	TB_LAUNCH(k_slots_digest, blocks, 256, st, d_slots, d_packed, n, k_base, d_out);
	CU(cudaGetLastError());
	unsigned long long h = 0;
	CU(cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	*digest = h;
	return 0;
}

static int pack_bits_async(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint8_t *d_packed, cudaStream_t st)
{
	if (n_bits == 0) return 0;
	const uint64_t words = (n_bits + 31) / 32;
	const unsigned blocks = (unsigned)std::min<uint64_t>((words + 255) / 256, (uint64_t)ctx->sm_count * 16);
	TB_LAUNCH(k_pack_bits, blocks, 256, st, d_bits, n_bits, reinterpret_cast<uint32_t *>(d_packed));
	CU(cudaGetLastError());
	return 0;
}

extern "C" int tb200_pack_bits_dev(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint8_t *d_packed)
{
	int r = leaf_common(ctx);
	if (r) return r;
	if (n_bits && (!d_bits || !d_packed)) return fail(ctx, TB200_E_ARG, "null argument");
	if ((uintptr_t)d_packed & 3) return fail(ctx, TB200_E_ARG, "d_packed must be 4-byte aligned");
	if ((r = pack_bits_async(ctx, d_bits, n_bits, d_packed, ctx->s_compute))) return r;
	CU(cudaStreamSynchronize(ctx->s_compute));
	return 0;
}

/* ------------------------------------------------- sharded decode (multi-GPU) -- */

static_assert(sizeof(tb200_shard_summary) == 32, "shard summary ABI");

/* what a rank tells the others about its shard: last CRC-good SB1 and first lock loss */
__global__ void k_shard_summary(const SlotWs *__restrict__ ws, const int32_t *__restrict__ last_good,
                                const int32_t *__restrict__ blk_prev, const uint32_t *__restrict__ first_unlock,
                                uint32_t n, tb200_shard_summary *out)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	tb200_shard_summary s;
	memset(&s, 0, sizeof(s));
	s.n_slots = n;
	s.first_unlock = *first_unlock;
	s.slots_after = n;
	if (n) {
		int32_t j = last_good[n - 1];
		if (j < 0) j = blk_prev[(n - 1) >> 10];
		if (j >= 0) {
			const SlotWs w = ws[j];
			s.has_good_sb = 1; s.scramb_init = w.sb_code; s.slots_after = n - 1 - (uint32_t)j;
			s.mcc = w.mcc; s.mnc = w.mnc; s.tn = w.tn; s.fn = w.fn; s.mn = w.mn; s.cc = w.cc;
		}
	}
	*out = s;
}

extern "C" int tb200_find_lock(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t n_bits, uint64_t *a0, uint64_t *cmin)
{
	if (!ctx || !d_bits || !a0 || !cmin) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	CU(cudaDeviceSynchronize());
	reset_stream(ctx);
	Source src; src.on_device = true; src.data = d_bits; src.new_base = 0; src.end = n_bits; src.fmt = (int)ctx->opt.input;
	Outputs out; out.on_device = true; out.slots = nullptr; out.type1 = nullptr; out.packed = nullptr; out.crc = nullptr; out.max_slots = 0; out.n = 0;
	ctx->fed_end = n_bits;
	ctx->stop_at_lock = true;
	int rc = rx_run(ctx, src, true, out);
	ctx->stop_at_lock = false;
	if (rc) return rc;
	if (ctx->rx.state != TB200_RX_LOCKED) return 0;
	*a0 = ctx->rx.buf_start;
	*cmin = ctx->rx.calls + 1;
	return 1;
}

extern "C" int tb200_shard_pass1(tb200_ctx *ctx, const uint8_t *d_bits, uint64_t base_bit, uint64_t n_bytes,
                                 uint64_t a0, uint64_t cmin, uint64_t n_end, uint32_t n_slots,
                                 tb200_shard_summary *summary)
{
	if (!ctx || !d_bits || !summary) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	if (a0 < base_bit) return fail(ctx, TB200_E_ARG, "shard bits start after the first slot");
	CU(cudaDeviceSynchronize());
	int rc;
	if ((rc = ensure_workspace(ctx, n_slots ? n_slots : 1))) return rc;
	if ((rc = ensure_pieces(ctx, 1))) return rc;
	RxGeom g;
	g.bits = d_bits; g.n_bytes = n_bytes; g.base_bit = base_bit; g.a0 = a0; g.cmin = cmin; g.fmt = (int)ctx->opt.input;
	if (g.fmt != IN_BYTES && ((base_bit & 127) || ((uintptr_t)d_bits & 3)))
		return fail(ctx, TB200_E_ARG, "bit-packed / symbol shards start on a 128-bit boundary of the stream, in a 4-byte aligned buffer");
	g.cg.c_base = 0; g.cg.t_base = 0; g.cg.n_end = n_end; g.cg.chunk = ctx->opt.chunk_bits; g.cg.pad = 0;
	g.n_slots = n_slots; g.tie_hi = (int)ctx->opt.viterbi_tie;
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	if (n_slots && (rc = enqueue_pass1(ctx, g, 0, 0, nullptr, false))) return rc;
	tb200_shard_summary *d_sum = reinterpret_cast<tb200_shard_summary *>(ctx->d_hits);    /* small scratch */
	cudaStream_t st = ctx->s_front;
	if (!n_slots) CU(cudaMemsetAsync(ctx->d_flags, 0xff, 2 * sizeof(uint32_t), st));
	TB_LAUNCH(k_shard_summary, 1, 32, st, ctx->wset[0].ws, ctx->wset[0].last_good, ctx->wset[0].blk_prev, ctx->d_flags, n_slots, d_sum);
	ctx->stats.kernel_launches++;
	CU(cudaMemcpyAsync(summary, d_sum, sizeof(*summary), cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	ctx->shard_a0 = a0;
	ctx->shard_slots = n_slots;
	return 0;
}

/* receiver state in front of rank `rank`'s shard, from the initial state and the summaries of ranks 0..rank-1 */
extern "C" void tb200_shard_carry_in(const tb200_shard_summary *all, int rank, const tb200_rx_carry *initial, tb200_rx_carry *out)
{
	tb200_rx_carry c = *initial;
	for (int r = 0; r < rank; r++) {
		const tb200_shard_summary &s = all[r];
		Tm t;
		if (s.has_good_sb) {
			t.tn = s.tn; t.fn = s.fn; t.mn = s.mn;
			t = tm_advance(t, s.slots_after);
			c.scramb_init = s.scramb_init; c.mcc = s.mcc; c.mnc = s.mnc; c.colour_code = s.cc;
		} else {
			t.tn = c.tn; t.fn = c.fn; t.mn = c.mn;
			t = tm_advance(t, s.n_slots);
		}
		c.tn = (uint8_t)t.tn; c.fn = (uint8_t)t.fn; c.mn = (uint8_t)t.mn;
	}
	*out = c;
}

extern "C" long tb200_shard_pass2(tb200_ctx *ctx, const tb200_rx_carry *carry_in, tb200_slot *d_slots, uint8_t *d_type1,
                                  uint32_t *d_type1_packed)
{
	if (!ctx || !carry_in || !d_slots) return TB200_E_ARG;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, TB200_E_CUDA, "cudaSetDevice");
	if (d_type1 && ((uintptr_t)d_type1 & 15)) return fail(ctx, TB200_E_ARG, "d_type1 must be 16-byte aligned");
	CU(cudaDeviceSynchronize());
	DevCarry dc;
	memset(&dc, 0, sizeof(dc));
	dc.scramb_init = carry_in->scramb_init; dc.tn = carry_in->tn; dc.fn = carry_in->fn; dc.mn = carry_in->mn;
	dc.mcc = carry_in->mcc; dc.mnc = carry_in->mnc; dc.cc = carry_in->colour_code;
	CU(cudaMemcpyAsync(ctx->d_carry, &dc, sizeof(dc), cudaMemcpyHostToDevice, ctx->s_compute));
	if (ctx->shard_slots) {
		int rc = enqueue_pass2(ctx, ctx->shard_a0, ctx->shard_slots, 0, 0, nullptr, false, (SlotOut *)d_slots, d_type1, d_type1_packed, 0);
		if (rc) return rc;
	}
	CU(cudaMemcpyAsync(&ctx->h_carry, ctx->d_carry + (ctx->shard_slots ? 1 : 0), sizeof(DevCarry), cudaMemcpyDeviceToHost, ctx->s_compute));
	if (ctx->shard_slots)
		CU(cudaMemcpyAsync(ctx->h_pstats, ctx->d_pstats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->s_compute));
	CU(cudaStreamSynchronize(ctx->s_compute));
	if (ctx->shard_slots) {
		ctx->stats.slots += ctx->shard_slots;
		ctx->stats.bursts_decoded += ctx->h_pstats[0]; ctx->stats.blocks += ctx->h_pstats[1]; ctx->stats.crc_ok_blocks += ctx->h_pstats[2];
	}
	return (long)ctx->shard_slots;
}

#include "tetra_dist.cuh"
