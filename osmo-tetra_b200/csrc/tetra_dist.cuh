/*
 * tetra_dist.cuh - one stream decoded by several GPUs: the driver behind tb200_dist_* (include/tetra_b200.h).
 * Included at the end of tetra_b200.cu (it drives the receiver's own machinery: rx_run, run_locked).
 *
 * Per LOCKED run ("segment") of the stream:
 *   rank 0   acquires lock like the single receiver (UNLOCKED / KNOW_FSTART of tetra_burst_sync.c:67-106 on the
 *            host, SYNC hits from k_scan_sync) and broadcasts where the run's slots sit + the receiver state;
 *   plan     contiguous slot ranges, one per rank, each with the look-ahead halo of the search window;
 *   data     SCATTER: rank 0 sends every shard in chunks (one grouped ncclSend per chunk index, all peers at once),
 *            the receivers post the matching ncclRecv's up front; every chunk has its own event, so a rank starts on
 *            its first piece while the later ones are still on the wire.  PEER: nothing is sent, the search kernels
 *            pull the shard out of rank 0's memory over NVLink.  PACK: rank 0 first turns the one-bit-per-byte stream
 *            into eight bits per byte, chunk by chunk in the order the chunks leave (k_pack_bits, HBM bound), so
 *            packing, sending and decoding overlap;
 *   decode   every rank runs the receiver's normal piece pipeline (run_locked) over its shard SPECULATIVELY: the
 *            cell state in front of the shard is not known yet, so the slots that would take it from the carry-in
 *            (those before the shard's first CRC-good SB1, tetra_lower_mac.c:291-302) are left out;
 *   exchange one all-gather of a 48-byte summary per rank (the only exchange step of the path): last cell state,
 *            first CRC-good SB1, first lock loss;
 *   fix-up   every rank derives its true carry-in and decodes the slots it left out (a handful);
 *   lock loss inside a shard (tetra_burst_sync.c:123-142): the segment ends right behind the losing slot, the ranks
 *            behind it drop their speculative slots, rank 0 takes the receiver state after that slot and goes back
 *            to the UNLOCKED search; what follows is a new segment.
 */
#pragma once

#include <chrono>
#include <unistd.h>
#ifndef TB_SIMT_EMULATION
#include <dlfcn.h>
#endif

/* ---- NCCL, loaded at run time (libnccl.so.2; torch brings its own copy into the process) ---- */
#ifndef TB_SIMT_EMULATION
struct NcclApi {
	typedef struct ncclComm *comm_t;
	struct uid { char internal[TB200_DIST_ID_BYTES]; };
	void *lib = nullptr;
	int (*GetUniqueId)(uid *) = nullptr;
	int (*CommInitRank)(comm_t *, int, uid, int) = nullptr;
	int (*CommDestroy)(comm_t) = nullptr;
	int (*Broadcast)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, comm_t, cudaStream_t) = nullptr;
	int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	static constexpr int kUint8 = 1;           /* ncclUint8 */
	bool load(char *err, size_t n)
	{
		if (lib) return true;
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for (const char *nm : names)
			if ((lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
		if (!lib) { snprintf(err, n, "libnccl.so.2 not found: %s", dlerror()); return false; }
#define TB_SYM(field, name) do { *(void **)(&field) = dlsym(lib, name); if (!field) { snprintf(err, n, "NCCL symbol %s missing", name); return false; } } while (0)
		TB_SYM(GetUniqueId, "ncclGetUniqueId"); TB_SYM(CommInitRank, "ncclCommInitRank"); TB_SYM(CommDestroy, "ncclCommDestroy");
		TB_SYM(Broadcast, "ncclBroadcast"); TB_SYM(AllGather, "ncclAllGather"); TB_SYM(Send, "ncclSend"); TB_SYM(Recv, "ncclRecv");
		TB_SYM(GroupStart, "ncclGroupStart"); TB_SYM(GroupEnd, "ncclGroupEnd"); TB_SYM(GetErrorString, "ncclGetErrorString");
#undef TB_SYM
		return true;
	}
};
static NcclApi g_nccl;
#endif

/* what every rank tells the others after its speculative pass */
struct DistSummary {
	uint64_t n_valid;          /* slots of the shard that count: all of it, or up to and including the one that lost lock */
	uint64_t first_good;       /* shard-relative index of the first CRC-good SB1, ~0 if none */
	uint32_t lost;             /* the shard's last valid slot lost lock */
	uint32_t seen_good;
	uint32_t scramb_init, tn, fn, mn, mcc, mnc, cc;     /* receiver state behind the last valid slot, if seen_good */
	uint32_t pad;
};
static_assert(sizeof(DistSummary) == 56, "summary layout");

struct DistMeta {              /* rank 0 -> all, per segment */
	uint32_t ok, chunk, fmt, tie;
	uint64_t a0, cmin, n_end, n_slots, k_done;
	DevCarry carry;
	uint64_t src_ptr;          /* PEER: the (packed or original) stream on rank 0 */
	int64_t pid;
	int32_t src_device;
	uint32_t feeder_ppm;       /* share of the whole segment that rank 0 gives up because it also packs the stream (parts per million) */
	uint8_t ipc[64];
};

struct tb200_dist {
	tb200_ctx *ctx = nullptr;
	int rank = 0, world = 1;
	char err[256] = {0};
	tb200_dist_ops ops;
	bool own_nccl = false;
#ifndef TB_SIMT_EMULATION
	NcclApi::comm_t comm = nullptr;
#endif
	uint8_t *d_stage = nullptr, *h_stage = nullptr;      /* small collectives: 16 KB each */
	uint8_t *d_packed = nullptr; size_t packed_cap = 0;  /* rank 0: the stream, eight bits per byte (TB200_DIST_PACK) */
	uint8_t *d_shard = nullptr; size_t shard_cap = 0;    /* the shard a rank received */
	uint64_t peer_key = 0; void *peer_map = nullptr; bool peer_ipc = false;
	cudaStream_t s_pack = nullptr, s_xfer = nullptr;
	cudaEvent_t ev_pack[2] = {nullptr, nullptr};         /* timing: first pack launch, behind the last one */
	std::vector<cudaEvent_t> ev_pool; size_t ev_used = 0;
	tb200_dist_timing timing;
};

static int dfail(tb200_dist *d, int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(d->err, sizeof(d->err), fmt, ap);
	va_end(ap);
	return code;
}
/* TB200_DIST_TRACE=1: progress of every rank on stderr (hunting a rank that waits for another) */
#define DTRACE(...) do { static const bool on_ = getenv("TB200_DIST_TRACE") != nullptr; \
	if (on_) { fprintf(stderr, "[tb200 dist r%d] ", d->rank); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } } while (0)
#define DCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return dfail(d, TB200_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

/* ---- the three collectives on NCCL ---- */
#ifndef TB_SIMT_EMULATION
#define DNC(call) do { int r_ = (call); if (r_ != 0) \
	return dfail(d, TB200_E_CUDA, "%s: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

static int nccl_bcast(void *user, void *host_buf, size_t bytes, int root)
{
	tb200_dist *d = (tb200_dist *)user;
	if (bytes > 8192) return dfail(d, TB200_E_ARG, "broadcast too large");
	cudaStream_t st = d->s_xfer;
	if (d->rank == root) { memcpy(d->h_stage, host_buf, bytes); DCU(cudaMemcpyAsync(d->d_stage, d->h_stage, bytes, cudaMemcpyHostToDevice, st)); }
	DNC(g_nccl.Broadcast(d->d_stage, d->d_stage, bytes, NcclApi::kUint8, root, d->comm, st));
	DCU(cudaMemcpyAsync(d->h_stage, d->d_stage, bytes, cudaMemcpyDeviceToHost, st));
	DCU(cudaStreamSynchronize(st));
	memcpy(host_buf, d->h_stage, bytes);
	return 0;
}

static int nccl_allgather(void *user, const void *host_send, void *host_recv, size_t bytes_each)
{
	tb200_dist *d = (tb200_dist *)user;
	if (bytes_each > 1024 || bytes_each * d->world > 8192) return dfail(d, TB200_E_ARG, "all-gather too large");
	cudaStream_t st = d->s_xfer;
	memcpy(d->h_stage, host_send, bytes_each);
	DCU(cudaMemcpyAsync(d->d_stage, d->h_stage, bytes_each, cudaMemcpyHostToDevice, st));
	DNC(g_nccl.AllGather(d->d_stage, d->d_stage + 8192, bytes_each, NcclApi::kUint8, d->comm, st));
	DCU(cudaMemcpyAsync(d->h_stage + 8192, d->d_stage + 8192, bytes_each * d->world, cudaMemcpyDeviceToHost, st));
	DCU(cudaStreamSynchronize(st));
	memcpy(host_recv, d->h_stage + 8192, bytes_each * d->world);
	return 0;
}
#endif

/* ---- life cycle ---- */

extern "C" int tb200_dist_get_id(uint8_t id[TB200_DIST_ID_BYTES])
{
#ifdef TB_SIMT_EMULATION
	(void)id;
	return TB200_E_STATE;
#else
	char err[256];
	if (!id || !g_nccl.load(err, sizeof(err))) return TB200_E_STATE;
	NcclApi::uid u;
	if (g_nccl.GetUniqueId(&u) != 0) return TB200_E_CUDA;
	memcpy(id, &u, TB200_DIST_ID_BYTES);
	return 0;
#endif
}

static int dist_common_init(tb200_dist *d)
{
	tb200_ctx *ctx = d->ctx;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return dfail(d, TB200_E_CUDA, "cudaSetDevice");
	DCU(cudaMalloc((void **)&d->d_stage, 16384));
	DCU(cudaHostAlloc((void **)&d->h_stage, 16384, cudaHostAllocDefault));
	/* packing and shipping feed every other rank: their blocks go first whenever an SM has room */
	int prio_lo = 0, prio_hi = 0;
	DCU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	DCU(cudaStreamCreateWithPriority(&d->s_pack, cudaStreamNonBlocking, prio_hi));
	DCU(cudaStreamCreateWithPriority(&d->s_xfer, cudaStreamNonBlocking, prio_hi));
	DCU(cudaEventCreateWithFlags(&d->ev_pack[0], 0));
	DCU(cudaEventCreateWithFlags(&d->ev_pack[1], 0));
	memset(&d->timing, 0, sizeof(d->timing));
	return 0;
}

extern "C" int tb200_dist_create_with_ops(tb200_dist **out, tb200_ctx *ctx, int rank, int world, const tb200_dist_ops *ops)
{
	if (!out || !ctx || !ops || world < 1 || rank < 0 || rank >= world || !ops->bcast || !ops->allgather) return TB200_E_ARG;
	tb200_dist *d = new tb200_dist();
	d->ctx = ctx; d->rank = rank; d->world = world; d->ops = *ops;
	int rc = dist_common_init(d);
	if (rc) { fprintf(stderr, "tetra_b200: %s\n", d->err); delete d; return rc; }
	*out = d;
	return 0;
}

extern "C" int tb200_dist_create(tb200_dist **out, tb200_ctx *ctx, int rank, int world, const uint8_t id[TB200_DIST_ID_BYTES])
{
	if (!out || !ctx || world < 1 || rank < 0 || rank >= world || !id) return TB200_E_ARG;
#ifdef TB_SIMT_EMULATION
	return fail(ctx, TB200_E_STATE, "no NCCL in the emulation build: use tb200_dist_create_with_ops");
#else
	tb200_dist *d = new tb200_dist();
	d->ctx = ctx; d->rank = rank; d->world = world; d->own_nccl = true;
	auto bail = [&](int rc) { fail(ctx, rc, "tb200_dist_create: %s", d->err); delete d; return rc; };
	if (!g_nccl.load(d->err, sizeof(d->err))) return bail(TB200_E_STATE);
	int rc = dist_common_init(d);
	if (rc) return bail(rc);
	NcclApi::uid u;
	memcpy(&u, id, TB200_DIST_ID_BYTES);
	const int r = g_nccl.CommInitRank(&d->comm, world, u, rank);
	if (r != 0) { snprintf(d->err, sizeof(d->err), "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); return bail(TB200_E_CUDA); }
	d->ops.user = d; d->ops.bcast = nccl_bcast; d->ops.allgather = nccl_allgather; d->ops.scatter = nullptr;
	*out = d;
	return 0;
#endif
}

extern "C" void tb200_dist_destroy(tb200_dist *d)
{
	if (!d) return;
	cudaSetDevice(d->ctx->device);
	cudaDeviceSynchronize();
#ifndef TB_SIMT_EMULATION
	if (d->peer_map && d->peer_ipc) cudaIpcCloseMemHandle(d->peer_map);
	if (d->comm) g_nccl.CommDestroy(d->comm);
#endif
	cudaFree(d->d_stage); cudaFreeHost(d->h_stage); cudaFree(d->d_packed); cudaFree(d->d_shard);
	for (cudaEvent_t e : d->ev_pool) cudaEventDestroy(e);
	for (int i = 0; i < 2; i++) if (d->ev_pack[i]) cudaEventDestroy(d->ev_pack[i]);
	if (d->s_pack) cudaStreamDestroy(d->s_pack);
	if (d->s_xfer) cudaStreamDestroy(d->s_xfer);
	delete d;
}

extern "C" const char *tb200_dist_last_error(const tb200_dist *d) { return d ? d->err : "null"; }

extern "C" uint64_t tb200_dist_max_local_slots(uint64_t n_bits, int world)
{
	/* Without lock losses a rank gets ceil(slots / world).  Every lock loss starts a new segment that is cut evenly
	 * again, while what a rank decoded in front of the loss stays with it - a stream that keeps losing lock early in
	 * its segments piles its slots up on the first ranks.  The bound that holds for every stream is all of them. */
	(void)world;
	return n_bits / SLOT_BITS + 1;
}

extern "C" int tb200_dist_get_timing(const tb200_dist *d, tb200_dist_timing *out)
{
	if (!d || !out) return TB200_E_ARG;
	*out = d->timing;
	return 0;
}

/* ---- helpers ---- */

static int dist_event(tb200_dist *d, cudaEvent_t *e)
{
	if (d->ev_used == d->ev_pool.size()) {
		cudaEvent_t n;
		DCU(cudaEventCreateWithFlags(&n, cudaEventDisableTiming));
		d->ev_pool.push_back(n);
	}
	*e = d->ev_pool[d->ev_used++];
	return 0;
}

static constexpr uint64_t DIST_HALO = 4096 + 64 + 128;      /* look-ahead of the search window (tetra_burst_sync.c:117) + read-ahead + alignment */

struct DistPlan {
	uint64_t k0, k1;           /* slot range of the rank inside the segment */
	uint64_t lo, hi, base;     /* stream bits: first slot, end of the data the shard needs, start of the data it holds */
};

/* Contiguous slot ranges.  Even, unless rank 0 also has to turn the whole one-bit-per-byte stream into packed bits
 * (TB200_DIST_PACK): that costs it rho = feeder_ppm / 10^6 of the time one GPU needs to decode the whole segment (it
 * reads 510 bytes per slot at HBM speed, decoding runs at ~1.6 * 10^9 slots/s), so the other ranks get
 * n (1 + rho) / world slots each and rank 0 what is left - with enough ranks nothing: it then only feeds the others. */
static void dist_bounds(const DistMeta &m, int world, int r, uint64_t *k0, uint64_t *k1)
{
	const uint64_t n = m.n_slots;
	if (world == 1) { *k0 = 0; *k1 = n; return; }
	const double rho = m.feeder_ppm * 1e-6;
	double others = (double)n * (1.0 + rho) / world;
	double first = others - rho * (double)n;
	if (first < 0) { first = 0; others = (double)n / (world - 1); }
	const uint64_t n0 = std::min<uint64_t>(n, (uint64_t)first);
	const uint64_t per = (n - n0 + world - 2) / (world - 1);
	if (r == 0) { *k0 = 0; *k1 = n0; return; }
	*k0 = std::min<uint64_t>(n, n0 + (uint64_t)(r - 1) * per);
	*k1 = std::min<uint64_t>(n, n0 + (uint64_t)r * per);
}

static DistPlan dist_plan(const DistMeta &m, int world, int r)
{
	DistPlan p;
	dist_bounds(m, world, r, &p.k0, &p.k1);
	p.lo = m.a0 + SLOT_BITS * p.k0;
	p.hi = std::max(p.lo, std::min<uint64_t>(m.n_end, m.a0 + SLOT_BITS * p.k1 + DIST_HALO));
	p.base = m.fmt == IN_BYTES ? p.lo : (p.lo & ~(uint64_t)127);
	return p;
}

/* the data of a shard travels in chunks so that the first piece can start early: chunk boundaries in stream bits
 * (ascending, multiples of 128, the last one = hi) */
static std::vector<uint64_t> dist_chunks(const DistPlan &p, uint64_t chunk_slots)
{
	std::vector<uint64_t> ends;
	for (uint64_t k = p.k0; k < p.k1; k += chunk_slots) {
		const uint64_t e = p.lo + SLOT_BITS * (std::min(k + chunk_slots, p.k1) - p.k0) + DIST_HALO;
		ends.push_back(std::min(p.hi, (e + 127) & ~(uint64_t)127));
	}
	if (!ends.empty()) ends.back() = p.hi;
	return ends;
}

static inline size_t dist_byte_of(int fmt, uint64_t bit) { return fmt == IN_BYTES ? (size_t)bit : (size_t)(bit >> 3); }
static inline size_t dist_bytes_upto(int fmt, uint64_t bit) { return fmt == IN_BYTES ? (size_t)bit : (size_t)((bit + 7) >> 3); }

/* ---- the call ---- */

extern "C" long tb200_dist_rx_stream(tb200_dist *d, const uint8_t *d_bits, uint64_t n_bits, uint32_t mode,
                                     tb200_slot *d_slots, uint8_t *d_type1, uint32_t *d_type1_packed, uint64_t max_slots,
                                     tb200_dist_run *runs, uint32_t max_runs, uint32_t *n_runs)
{
	if (!d || !d_slots || !n_runs || (max_runs && !runs)) return TB200_E_ARG;
	tb200_ctx *ctx = d->ctx;
	const int rank = d->rank, world = d->world;
	const bool peer = (mode & 0xff) == TB200_DIST_PEER;
	const bool pack = (mode & TB200_DIST_PACK) != 0;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return dfail(d, TB200_E_CUDA, "cudaSetDevice");
	if (ctx->opt.viterbi != TB200_VITERBI_LANE) return dfail(d, TB200_E_ARG, "the sharded decode needs the lane kernels");
	if (rank == 0 && (!d_bits || (pack && ctx->opt.input != TB200_IN_BYTES) || ctx->opt.input == TB200_IN_F32SYM))
		return dfail(d, TB200_E_ARG, "rank 0 needs the stream (one bit per byte, or bit-packed without TB200_DIST_PACK)");
	if (d_type1 && ((uintptr_t)d_type1 & 15)) return dfail(d, TB200_E_ARG, "d_type1 must be 16-byte aligned");
#ifdef TB_SIMT_EMULATION
	if (peer && world > 1) return dfail(d, TB200_E_STATE, "no peer mapping in the emulation build");
#endif
	DCU(cudaDeviceSynchronize());
	typedef std::chrono::steady_clock clk;
	auto ms_since = [](clk::time_point t) { return std::chrono::duration<float, std::milli>(clk::now() - t).count(); };
	const clk::time_point t_call = clk::now();
	memset(&d->timing, 0, sizeof(d->timing));
	*n_runs = 0;
	uint64_t n_local = 0;
	bool first_segment = true;
	const tb200_options saved_opt = ctx->opt;
	auto restore = [&]() { ctx->opt = saved_opt; };

	/* rank 0: the receiver whose state machine this is */
	reset_stream(ctx);
	ctx->lock_events.clear();
	const int src_fmt = rank == 0 ? (int)ctx->opt.input : IN_BYTES;
	Source fsm_src; fsm_src.on_device = true; fsm_src.data = d_bits; fsm_src.new_base = 0; fsm_src.end = n_bits; fsm_src.fmt = src_fmt;
	const uint8_t *xfer_src = d_bits;        /* what the shards are cut from */
	int xfer_fmt = src_fmt;
	if (rank == 0) {
		ctx->fed_end = n_bits;
		if (pack) {
			const size_t need = 4 * (size_t)((n_bits + 31) / 32) + 256;
			if (need > d->packed_cap) {
				cudaFree(d->d_packed); d->d_packed = nullptr; d->packed_cap = 0;
				DCU(cudaMalloc((void **)&d->d_packed, need));
				d->packed_cap = need;
			}
			xfer_src = d->d_packed; xfer_fmt = IN_PACKED;
		}
	}
	bool packed_any = false;

	for (;;) {
		/* ---- lock acquisition on rank 0, then the segment's geometry to everyone */
		DistMeta m;
		memset(&m, 0, sizeof(m));
		clk::time_point t0 = clk::now();
		Segment seg;
		uint64_t c_max = 0;
		if (rank == 0) {
			const uint32_t C = ctx->opt.chunk_bits;
			Outputs none; none.on_device = true; none.slots = nullptr; none.type1 = nullptr; none.packed = nullptr; none.crc = nullptr; none.max_slots = 0; none.n = 0;
			ctx->stop_at_lock = true;
			ctx->opt.input = (uint32_t)src_fmt;
			int rc = rx_run(ctx, fsm_src, true, none, !first_segment);
			first_segment = false;
			ctx->stop_at_lock = false;
			if (rc) { restore(); return dfail(d, rc, "%s", ctx->err); }
			CallGeom cg;
			cg.c_base = 0; cg.t_base = 0; cg.n_end = n_bits; cg.chunk = C; cg.pad = 0;      /* one run: the whole stream, C bits per call */
			c_max = (n_bits + C - 1) / C;
			if (ctx->rx.state == TB200_RX_LOCKED) {
				m.n_slots = locked_extent(ctx->rx, cg, c_max, &seg);
				m.ok = m.n_slots > 0;
			}
			m.chunk = C; m.fmt = (uint32_t)xfer_fmt; m.tie = ctx->opt.viterbi_tie;
			m.a0 = seg.a0; m.cmin = seg.cmin; m.n_end = n_bits; m.k_done = ctx->stats.slots;
			m.carry = ctx->h_carry; m.carry.seen_good = 0; m.carry.first_good = ~0ull;
			m.src_ptr = (uint64_t)(uintptr_t)xfer_src; m.pid = (int64_t)getpid(); m.src_device = ctx->device;
			if (pack && world > 1) {
				/* packing reads 510 B per slot at ~6 TB/s = 1.2e10 slots/s, decoding runs at ~1.6e9 slots/s */
				m.feeder_ppm = 130000;
				if (const char *e = getenv("TB200_DIST_FEEDER_PPM")) m.feeder_ppm = (uint32_t)atoi(e);
			}
#ifndef TB_SIMT_EMULATION
			if (peer && world > 1 && m.ok) {
				cudaIpcMemHandle_t h;
				if (cudaIpcGetMemHandle(&h, const_cast<uint8_t *>(xfer_src)) == cudaSuccess) memcpy(m.ipc, &h, 64);
				else cudaGetLastError();
			}
#endif
		}
		if (world > 1) {
			int rc = d->ops.bcast(d->ops.user, &m, sizeof(m), 0);
			if (rc) { restore(); return rc; }
		}
		d->timing.lock_ms += ms_since(t0);
		DTRACE("segment meta: ok %u a0 %llu slots %llu fmt %u", m.ok, (unsigned long long)m.a0, (unsigned long long)m.n_slots, m.fmt);
		if (!m.ok) break;
		d->timing.segments++;
		xfer_fmt = (int)m.fmt;            /* every rank: how the shards are encoded on the wire */
		seg.a0 = m.a0; seg.cmin = m.cmin;
		seg.cg.c_base = 0; seg.cg.t_base = 0; seg.cg.n_end = m.n_end; seg.cg.chunk = m.chunk; seg.cg.pad = 0;
		ctx->opt.chunk_bits = m.chunk; ctx->opt.viterbi_tie = m.tie; ctx->opt.input = m.fmt;

		const DistPlan me = dist_plan(m, world, rank);
		const uint64_t n_mine = me.k1 - me.k0;
		const uint64_t chunk_slots = 1u << 20;
		d->ev_used = 0;
		std::vector<std::pair<uint64_t, cudaEvent_t>> ready;

		/* ---- data: pack (rank 0), scatter or peer mapping */
		t0 = clk::now();
		const uint8_t *shard_ptr = nullptr;      /* holds stream bits from me.base on */
		if (rank == 0) {
			/* chunk index major, rank minor: every rank's first chunk leaves first */
			std::vector<DistPlan> plans(world);
			std::vector<std::vector<uint64_t>> ends(world);
			size_t max_chunks = 0;
			for (int r = 0; r < world; r++) {
				plans[r] = dist_plan(m, world, r);
				ends[r] = dist_chunks(plans[r], chunk_slots);
				max_chunks = std::max(max_chunks, ends[r].size());
			}
			/* packing runs front to back over the stream, so the order of the chunks is also the order of the bits as
			 * long as chunk c of rank r+1 lies behind chunk c of rank r - it does not (rank-minor order jumps back and
			 * forth), so every chunk packs exactly its own bit range [start, end) */
			/* two sweeps over the chunk indices: first the chunks of the other ranks (every rank's first chunk leaves
			 * first, so all of them start early), then rank 0's own - it decodes once it has fed everybody */
			for (int sweep = 0; sweep < 2; sweep++)
			for (size_t c = 0; c < max_chunks; c++) {
				bool group_open = false;
				for (int r = sweep == 0 ? 1 : 0; r < (sweep == 0 ? world : 1); r++) {
					if (c >= ends[r].size()) continue;
					const uint64_t b0 = c == 0 ? plans[r].base : ends[r][c - 1];
					const uint64_t b1 = ends[r][c];
					cudaEvent_t ev_packed = nullptr;
					if (pack) {
						/* [b0, b1) of the byte stream -> packed words; b0 is a multiple of 128 */
						if (!packed_any) { DCU(cudaEventRecord(d->ev_pack[0], d->s_pack)); packed_any = true; }
						/* whole words, also the last one: the ranges of neighbouring ranks overlap by the halo, and whoever
						 * packs a word must write all of it */
						const uint64_t b1w = std::min<uint64_t>((b1 + 31) & ~(uint64_t)31, n_bits);
						int rc = pack_bits_async(ctx, d_bits + b0, b1w - b0, d->d_packed + (b0 >> 3), d->s_pack);
						if (rc) { restore(); return dfail(d, rc, "%s", ctx->err); }
						int er = dist_event(d, &ev_packed);
						if (er) { restore(); return er; }
						DCU(cudaEventRecord(ev_packed, d->s_pack));
					}
					if (r == 0) {
						if (ev_packed) ready.push_back({ b1, ev_packed });
						continue;
					}
					if (peer) continue;      /* the readers pull; they wait for the whole pack below */
#ifndef TB_SIMT_EMULATION
					if (d->own_nccl) {
						if (ev_packed) DCU(cudaStreamWaitEvent(d->s_xfer, ev_packed, 0));
						if (!group_open) { DNC(g_nccl.GroupStart()); group_open = true; }
						const size_t o0 = dist_byte_of(xfer_fmt, b0), o1 = dist_bytes_upto(xfer_fmt, b1);
						DNC(g_nccl.Send(xfer_src + o0, o1 - o0, NcclApi::kUint8, r, d->comm, d->s_xfer));
						d->timing.bytes_sent += o1 - o0;
					}
#endif
				}
#ifndef TB_SIMT_EMULATION
				if (group_open) DNC(g_nccl.GroupEnd());
#endif
				(void)group_open;
			}
			shard_ptr = xfer_src + dist_byte_of(xfer_fmt, me.base);
			if (!d->own_nccl && !peer && world > 1) {
				/* caller-supplied plumbing moves whole shards */
				if (pack) DCU(cudaStreamSynchronize(d->s_pack));
				std::vector<uint64_t> offs(world), sizes(world);
				for (int r = 0; r < world; r++) {
					offs[r] = dist_byte_of(xfer_fmt, plans[r].base);
					sizes[r] = r == 0 || plans[r].k1 == plans[r].k0 ? 0 : dist_bytes_upto(xfer_fmt, plans[r].hi) - offs[r];
					d->timing.bytes_sent += sizes[r];
				}
				if (!d->ops.scatter) { restore(); return dfail(d, TB200_E_ARG, "scatter callback missing"); }
				int rc = d->ops.scatter(d->ops.user, xfer_src, offs.data(), sizes.data(), nullptr, 0);
				if (rc) { restore(); return rc; }
			}
			if (peer && world > 1) {
				/* the readers start once everything they may touch is packed */
				if (pack) DCU(cudaStreamSynchronize(d->s_pack));
				uint32_t go = 1;
				int rc = d->ops.bcast(d->ops.user, &go, sizeof(go), 0);
				if (rc) { restore(); return rc; }
				for (int r = 1; r < world; r++)
					if (plans[r].k1 > plans[r].k0) d->timing.bytes_sent += dist_bytes_upto(xfer_fmt, plans[r].hi) - dist_byte_of(xfer_fmt, plans[r].base);
			}
		} else if (peer) {
#ifndef TB_SIMT_EMULATION
			if (d->peer_key != m.src_ptr || !d->peer_map) {
				if (d->peer_map && d->peer_ipc) cudaIpcCloseMemHandle(d->peer_map);
				d->peer_map = nullptr;
				if (m.pid == (int64_t)getpid()) {
					/* the owner is a thread of this process: plain peer access */
					cudaError_t e = cudaDeviceEnablePeerAccess(m.src_device, 0);
					if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { restore(); return dfail(d, TB200_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); }
					cudaGetLastError();
					d->peer_map = (void *)(uintptr_t)m.src_ptr; d->peer_ipc = false;
				} else {
					cudaIpcMemHandle_t h;
					memcpy(&h, m.ipc, 64);
					DCU(cudaIpcOpenMemHandle(&d->peer_map, h, cudaIpcMemLazyEnablePeerAccess));
					d->peer_ipc = true;
				}
				d->peer_key = m.src_ptr;
			}
			uint32_t go = 0;
			int rc = d->ops.bcast(d->ops.user, &go, sizeof(go), 0);
			if (rc) { restore(); return rc; }
			shard_ptr = (const uint8_t *)d->peer_map + dist_byte_of(xfer_fmt, me.base);
#endif
		} else {
			const size_t o0 = dist_byte_of(xfer_fmt, me.base), o1 = dist_bytes_upto(xfer_fmt, me.hi);
			const size_t need = (o1 - o0) + 256;
			if (need > d->shard_cap) {
				cudaFree(d->d_shard); d->d_shard = nullptr; d->shard_cap = 0;
				DCU(cudaMalloc((void **)&d->d_shard, need + need / 8));
				d->shard_cap = need + need / 8;
			}
			shard_ptr = d->d_shard;
			if (n_mine) {
#ifndef TB_SIMT_EMULATION
				if (d->own_nccl) {
					const std::vector<uint64_t> ends = dist_chunks(me, chunk_slots);
					for (size_t c = 0; c < ends.size(); c++) {
						const uint64_t b0 = c == 0 ? me.base : ends[c - 1], b1 = ends[c];
						const size_t c0 = dist_byte_of(xfer_fmt, b0), c1 = dist_bytes_upto(xfer_fmt, b1);
						DNC(g_nccl.Recv(d->d_shard + (c0 - o0), c1 - c0, NcclApi::kUint8, 0, d->comm, d->s_xfer));
						cudaEvent_t ev;
						int er = dist_event(d, &ev);
						if (er) { restore(); return er; }
						DCU(cudaEventRecord(ev, d->s_xfer));
						ready.push_back({ b1, ev });
					}
				} else
#endif
				{
					if (!d->ops.scatter) { restore(); return dfail(d, TB200_E_ARG, "scatter callback missing"); }
				}
			}
			if (!d->own_nccl) {
				std::vector<uint64_t> offs(world, 0), sizes(world, 0);
				sizes[rank] = n_mine ? o1 - o0 : 0;
				int rc = d->ops.scatter(d->ops.user, nullptr, offs.data(), sizes.data(), d->d_shard, 0);
				if (rc) { restore(); return rc; }
			}
		}
		d->timing.transfer_ms += ms_since(t0);
		DTRACE("data queued: %llu slots, %zu ready events", (unsigned long long)n_mine, ready.size());

		/* ---- speculative pass over the shard */
		t0 = clk::now();
		DistSummary mine;
		memset(&mine, 0, sizeof(mine));
		mine.first_good = ~0ull;
		const uint64_t local_base = n_local;
		Segment sseg;                 /* the shard as a LOCKED run of its own */
		sseg.a0 = me.lo; sseg.cmin = m.cmin + me.k0; sseg.cg = seg.cg;
		Source ssrc; ssrc.on_device = true; ssrc.data = shard_ptr; ssrc.new_base = me.base; ssrc.end = me.hi; ssrc.fmt = (int)m.fmt;
		ssrc.ready = ready.empty() ? nullptr : &ready;
		ssrc.skip_dependent = rank > 0;
		Outputs out; out.on_device = true; out.slots = d_slots; out.type1 = d_type1; out.packed = d_type1_packed;
		out.crc = ctx->user_crc; out.aach = ctx->user_aach;     /* side outputs: rank-local, indexed like the slots */
		out.max_slots = max_slots; out.n = n_local;
		DevCarry chain_end = m.carry;
		if (n_mine) {
			ctx->h_carry = m.carry;
			int rc = push_carry(ctx);
			uint64_t valid = 0; bool lost = false;
			if (!rc) rc = run_locked(ctx, ssrc, sseg, n_mine, out, &valid, &lost);
			if (rc) { restore(); return dfail(d, rc, "%s", ctx->err); }
			chain_end = ctx->h_carry;
			mine.n_valid = valid; mine.lost = lost; mine.seen_good = chain_end.seen_good; mine.first_good = chain_end.first_good;
			mine.scramb_init = chain_end.scramb_init; mine.tn = chain_end.tn; mine.fn = chain_end.fn; mine.mn = chain_end.mn;
			mine.mcc = chain_end.mcc; mine.mnc = chain_end.mnc; mine.cc = chain_end.cc;
		}
		d->timing.pass1_ms += ms_since(t0);
		DTRACE("first pass done: valid %llu lost %u seen_good %u", (unsigned long long)mine.n_valid, mine.lost, mine.seen_good);

		/* ---- the exchange step: everyone's summary to everyone */
		t0 = clk::now();
		std::vector<DistSummary> all(world);
		if (world > 1) {
			int rc = d->ops.allgather(d->ops.user, &mine, all.data(), sizeof(DistSummary));
			if (rc) { restore(); return rc; }
		} else {
			all[0] = mine;
		}
		d->timing.exchange_ms += ms_since(t0);
		DTRACE("summaries exchanged");

		/* the segment ends behind the first lock loss; ranks behind it drop what they decoded */
		int r_lost = -1;
		for (int r = 0; r < world && r_lost < 0; r++)
			if (all[r].lost) r_lost = r;
		uint64_t seg_valid = 0;
		for (int r = 0; r < world; r++) {
			const DistPlan p = dist_plan(m, world, r);
			if (r_lost >= 0 && r > r_lost) break;
			seg_valid = (p.k0) + all[r].n_valid;
		}
		const bool i_count = r_lost < 0 || rank <= r_lost;
		/* carry-in of a rank = state behind the previous ranks' slots: the last summary with a CRC-good SB1, advanced
		 * over the slots of the ranks after it (tetra_tdma_time_add_tn per slot, tm_advance in closed form) */
		auto state_before = [&](int upto) {
			DevCarry c = m.carry;
			for (int r = 0; r < upto; r++) {
				const DistSummary &s = all[r];
				if (s.seen_good) {
					c.scramb_init = s.scramb_init; c.tn = s.tn; c.fn = s.fn; c.mn = s.mn; c.mcc = s.mcc; c.mnc = s.mnc; c.cc = s.cc;
				} else {
					Tm t = { c.tn, c.fn, c.mn };
					t = tm_advance(t, s.n_valid);
					c.tn = t.tn; c.fn = t.fn; c.mn = t.mn;
				}
			}
			c.seen_good = 0; c.first_good = ~0ull;
			return c;
		};

		/* ---- fix-up: the slots that waited for the carry-in */
		t0 = clk::now();
		if (i_count && rank > 0 && mine.n_valid) {
			const uint64_t head = mine.seen_good ? std::min<uint64_t>(mine.first_good, mine.n_valid) : mine.n_valid;
			if (head) {
				ctx->h_carry = state_before(rank);
				int rc = push_carry(ctx);
				Source hsrc = ssrc; hsrc.skip_dependent = false; hsrc.ready = nullptr;
				Outputs hout = out; hout.n = local_base;
				uint64_t valid = 0; bool lost = false;
				/* counters of the head were not taken in the first pass (its slots were left out) */
				if (!rc) rc = run_locked(ctx, hsrc, sseg, head, hout, &valid, &lost);
				if (rc) { restore(); return dfail(d, rc, "%s", ctx->err); }
				if (valid != head || (lost && head != mine.n_valid)) { restore(); return dfail(d, TB200_E_STATE, "fix-up pass disagrees with the first pass"); }
			}
		}
		d->timing.pass2_ms += ms_since(t0);

		if (i_count && mine.n_valid) {
			if (*n_runs < max_runs) {
				runs[*n_runs].global_slot = m.k_done + me.k0;
				runs[*n_runs].local_slot = local_base;
				runs[*n_runs].n_slots = mine.n_valid;
			}
			(*n_runs)++;
			n_local = local_base + mine.n_valid;
		} else {
			n_local = local_base;
		}
		if (*n_runs > max_runs) { restore(); return dfail(d, TB200_E_ARG, "more runs than the caller's array holds"); }

		/* ---- receiver state behind the segment's last valid slot, on every rank (rank 0 needs it to go on) */
		const int r_last = r_lost >= 0 ? r_lost : world - 1;
		DevCarry after = state_before(r_last + 1);
		ctx->h_carry = after;
		if (rank == 0) {
			/* the run_locked calls above counted rank 0's slots only */
			ctx->stats.slots = m.k_done;
			locked_advance(ctx, seg, seg_valid, r_lost >= 0);
			if (r_lost < 0) {         /* what rx_run does when no further slot fits: the remaining calls only fill the buffer */
				ctx->rx.calls = c_max;
				ctx->rx.bits_in_buf = (uint32_t)(n_bits - ctx->rx.buf_start);
			}
		}
		if (r_lost < 0) break;            /* the run reached the end of the stream */
	}
	if (packed_any) DCU(cudaEventRecord(d->ev_pack[1], d->s_pack));
	DCU(cudaStreamSynchronize(d->s_pack));
	DCU(cudaStreamSynchronize(d->s_xfer));
#ifndef TB_SIMT_EMULATION
	if (packed_any) DCU(cudaEventElapsedTime(&d->timing.pack_ms, d->ev_pack[0], d->ev_pack[1]));
#endif
	restore();
	d->timing.total_ms = ms_since(t_call);
	return (long)n_local;
}
